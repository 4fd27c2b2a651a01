// rgb_tiles.h -- the 8 x 8 pixel tiles of RGBImgPartialObsWrapper (gym_minigrid 1.0.x `wrappers.py` / `minigrid.py`
// Grid.render_tile / `rendering.py`; used by the reference when 'pixel' is in the architecture name:
// scripts/train_rl.py:54-58, babyai/evaluate.py:91-92), rasterised ONCE on the host when a pool first renders.
//
// The wrapper's image is a pure function of the 7 x 7 x 3 observation: every view cell becomes one tile that depends only
// on (type, color, state), on whether the cell is visible (highlight) and on whether it is the agent's cell (3, 6) --
// 513 distinct tiles of 192 bytes:
//   id 0..255     visible cell with cell byte  type | color << 3 | state << 6   (object drawn, highlighted)
//   id 256        unseen cell (type 0): grid lines only, no highlight
//   id 257..512   the agent's own cell: 257 + cell byte (what it carries, or empty), the agent triangle on top, highlighted
// Each tile is drawn exactly as the reference package draws it: shapes are predicates on the unit square sampled at the
// pixel centres of a 24 x 24 supersampled tile (float64, same operation order as the Python code so that the comparisons
// and the box filter round identically), highlight = img + 0.3 (255 - img), 3 x 3 box filter as two successive means,
// truncation to uint8.  tests/test_rgb.py compares every tile with the oracle shim's literal restatement of that code.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

namespace bb_rgb {

constexpr int TILE = 8, SUB = 3, RES = TILE * SUB;              // 24 x 24 samples per tile
constexpr int N_TILES = 513, TILE_BYTES = TILE * TILE * 3;      // 192
constexpr int ID_UNSEEN = 256, ID_AGENT0 = 257;

struct Canvas { uint8_t px[RES][RES][3]; };

struct Rect { double x0, x1, y0, y1; bool in(double x, double y) const { return x >= x0 && x <= x1 && y >= y0 && y <= y1; } };
struct Circle { double cx, cy, r; bool in(double x, double y) const { return (x - cx) * (x - cx) + (y - cy) * (y - cy) <= r * r; } };

template <class F>
static void fill(Canvas &c, const F &f, const double col[3])
{
    for (int y = 0; y < RES; y++)
        for (int x = 0; x < RES; x++) {
            const double yf = (y + 0.5) / RES, xf = (x + 0.5) / RES;
            if (f.in(xf, yf)) for (int k = 0; k < 3; k++) c.px[y][x][k] = (uint8_t)col[k];     // float -> uint8: truncation
        }
}

// point_in_triangle(a, b, c) behind rotate_fn(.., cx = cy = 0.5, theta)
struct RotTriangle {
    double ax, ay, bx, by, cx_, cy_, cs, sn;
    bool in(double x, double y) const
    {
        x = x - 0.5; y = y - 0.5;
        const double x2 = 0.5 + x * cs - y * sn;
        const double y2 = 0.5 + y * cs + x * sn;
        const double v0x = cx_ - ax, v0y = cy_ - ay, v1x = bx - ax, v1y = by - ay, v2x = x2 - ax, v2y = y2 - ay;
        const double dot00 = v0x * v0x + v0y * v0y, dot01 = v0x * v1x + v0y * v1y, dot02 = v0x * v2x + v0y * v2y;
        const double dot11 = v1x * v1x + v1y * v1y, dot12 = v1x * v2x + v1y * v2y;
        const double inv = 1 / (dot00 * dot11 - dot01 * dot01);
        const double u = (dot11 * dot02 - dot01 * dot12) * inv, v = (dot00 * dot12 - dot01 * dot02) * inv;
        return u >= 0 && v >= 0 && (u + v) < 1;
    }
};

static const double COLORS[6][3] = { { 255, 0, 0 }, { 0, 255, 0 }, { 0, 0, 255 }, { 112, 39, 195 }, { 255, 255, 0 }, { 100, 100, 100 } };

// WorldObj.render of the object a cell byte decodes to (WorldObj.decode: empty / unseen -> nothing)
static void draw_object(Canvas &cv, int type, int color, int state)
{
    if (color > 5) return;
    const double *c = COLORS[color];
    const double black[3] = { 0, 0, 0 };
    if (type == 2) fill(cv, Rect{ 0, 1, 0, 1 }, c);                                   // wall
    else if (type == 4) {                                                             // door
        if (state == 0) {
            fill(cv, Rect{ 0.88, 1.00, 0.00, 1.00 }, c);
            fill(cv, Rect{ 0.92, 0.96, 0.04, 0.96 }, black);
        } else if (state == 2) {
            const double dim[3] = { 0.45 * c[0], 0.45 * c[1], 0.45 * c[2] };
            fill(cv, Rect{ 0.00, 1.00, 0.00, 1.00 }, c);
            fill(cv, Rect{ 0.06, 0.94, 0.06, 0.94 }, dim);
            fill(cv, Rect{ 0.52, 0.75, 0.50, 0.56 }, c);
        } else {
            fill(cv, Rect{ 0.00, 1.00, 0.00, 1.00 }, c);
            fill(cv, Rect{ 0.04, 0.96, 0.04, 0.96 }, black);
            fill(cv, Rect{ 0.08, 0.92, 0.08, 0.92 }, c);
            fill(cv, Rect{ 0.12, 0.88, 0.12, 0.88 }, black);
            fill(cv, Circle{ 0.75, 0.50, 0.08 }, c);
        }
    } else if (type == 5) {                                                           // key
        fill(cv, Rect{ 0.50, 0.63, 0.31, 0.88 }, c);
        fill(cv, Rect{ 0.38, 0.50, 0.59, 0.66 }, c);
        fill(cv, Rect{ 0.38, 0.50, 0.81, 0.88 }, c);
        fill(cv, Circle{ 0.56, 0.28, 0.190 }, c);
        fill(cv, Circle{ 0.56, 0.28, 0.064 }, black);
    } else if (type == 6) fill(cv, Circle{ 0.5, 0.5, 0.31 }, c);                      // ball
    else if (type == 7) {                                                             // box
        fill(cv, Rect{ 0.12, 0.88, 0.12, 0.88 }, c);
        fill(cv, Rect{ 0.18, 0.82, 0.18, 0.82 }, black);
        fill(cv, Rect{ 0.16, 0.84, 0.47, 0.53 }, c);
    }
    // (floor / goal / lava never occur in BabyAI levels: drawn as empty)
}

// Grid.render_tile(obj, agent_dir = 3 if agent else None, highlight, tile_size = 8)
static void render_tile(int cell_byte, bool has_obj, bool agent, bool highlight, uint8_t out[TILE_BYTES])
{
    Canvas cv;
    memset(&cv, 0, sizeof cv);
    const double grey[3] = { 100, 100, 100 }, red[3] = { 255, 0, 0 };
    fill(cv, Rect{ 0, 0.031, 0, 1 }, grey);
    fill(cv, Rect{ 0, 1, 0, 0.031 }, grey);
    if (has_obj) draw_object(cv, cell_byte & 7, (cell_byte >> 3) & 7, cell_byte >> 6);
    if (agent) {
        const double theta = 0.5 * M_PI * 3;
        fill(cv, RotTriangle{ 0.12, 0.19, 0.87, 0.50, 0.12, 0.81, cos(-theta), sin(-theta) }, red);
    }
    if (highlight)
        for (int y = 0; y < RES; y++)
            for (int x = 0; x < RES; x++)
                for (int k = 0; k < 3; k++) {
                    const uint8_t p = cv.px[y][x][k];
                    double b = (double)p + 0.30 * (double)(uint8_t)(255 - p);
                    b = b < 0 ? 0 : b > 255 ? 255 : b;
                    cv.px[y][x][k] = (uint8_t)b;
                }
    // downsample: mean over the 3 sub-columns, then mean over the 3 sub-rows (numpy float64, in that order), truncated
    for (int ty = 0; ty < TILE; ty++)
        for (int tx = 0; tx < TILE; tx++)
            for (int k = 0; k < 3; k++) {
                double m[SUB];
                for (int sy = 0; sy < SUB; sy++) {
                    const uint8_t *row = &cv.px[ty * SUB + sy][tx * SUB][0];
                    m[sy] = (((double)row[k] + (double)row[3 + k]) + (double)row[6 + k]) / 3.0;
                }
                const double v = ((m[0] + m[1]) + m[2]) / 3.0;
                out[(ty * TILE + tx) * 3 + k] = (uint8_t)v;
            }
}

// the whole table: N_TILES x 8 x 8 x 3 bytes
static void render_all_tiles(uint8_t *lut)
{
    for (int b = 0; b < 256; b++) {
        const bool obj = (b & 7) >= 2;                   // types 0 (unseen) and 1 (empty) decode to no object
        render_tile(b, obj, false, true, lut + (size_t)b * TILE_BYTES);
        render_tile(b, obj, true, true, lut + (size_t)(ID_AGENT0 + b) * TILE_BYTES);
    }
    render_tile(0, false, false, false, lut + (size_t)ID_UNSEEN * TILE_BYTES);
}

}  // namespace bb_rgb
