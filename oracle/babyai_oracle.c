/*
 * babyai_oracle.c -- CPU restatement of the BabyAI environment hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the *checker* (and the CPU baseline
 * "port"): only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  Nothing under babyai_b200/ links, imports
 * or calls it.
 *
 * What it restates (each function cites what it follows):
 *   - gym_minigrid 1.0.x MiniGridEnv / Grid / RoomGrid  -- third-party, absent
 *     from /root/reference; restated from SURVEY.md Appendix A and kept
 *     literal: object-per-cell grid, real slice + rotate_left x (dir+1),
 *     nested-loop process_vis, per-cell encode.
 *   - /root/reference/babyai/levels/levelgen.py  (RoomGridLevel, LevelGen)
 *   - /root/reference/babyai/levels/verifier.py  (ObjDesc, *Instr)
 *   - /root/reference/babyai/levels/iclr19_levels.py (level parameterisation)
 *   - /root/reference/babyai/rl/utils/penv.py:4-16 (auto-reset on done)
 *
 * Pinning: tests/test_oracle_vs_reference.py runs the reference's own,
 * unmodified babyai.levels on the clean-room gym_minigrid shim (oracle/shim)
 * with the Philox back-end and checks this file against it step by step
 * (obs bytes, reward, done, mission, full grid); tests/golden/ holds traces
 * generated that way.  Below the gym_minigrid boundary parity is unpinned
 * (no copy of the third-party package exists here) -- see DESIGN.md.
 *
 * RNG: Philox4x32-10, stream definition in oracle/philox.py.
 *
 * Build: oracle/build.py  (gcc -O2 -ffp-contract=off -shared -fPIC -pthread)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <pthread.h>

/* ---- constants: SURVEY App. A.1 ------------------------------------------ */
enum { T_UNSEEN = 0, T_EMPTY = 1, T_WALL = 2, T_DOOR = 4, T_KEY = 5, T_BALL = 6, T_BOX = 7 };
enum { C_RED = 0, C_GREEN = 1, C_BLUE = 2, C_PURPLE = 3, C_YELLOW = 4, C_GREY = 5 };
/* COLOR_NAMES = sorted(COLORS) = blue green grey purple red yellow */
static const int COLOR_NAMES[6] = { C_BLUE, C_GREEN, C_GREY, C_PURPLE, C_RED, C_YELLOW };
static const char *COLOR_STR[6] = { "red", "green", "blue", "purple", "yellow", "grey" };
static const char *TYPE_STR[8] = { "", "", "wall", "", "door", "key", "ball", "box" };
static const int DIR_X[4] = { 1, 0, -1, 0 };   /* DIR_TO_VEC: right, down, left, up */
static const int DIR_Y[4] = { 0, 1, 0, -1 };
enum { A_LEFT = 0, A_RIGHT = 1, A_FORWARD = 2, A_PICKUP = 3, A_DROP = 4, A_TOGGLE = 5, A_DONE = 6 };
/* add_object / add_distractors draw from ['key','ball','box'] (App. A.6) */
static const int KBB[3] = { T_KEY, T_BALL, T_BOX };
/* verifier.py:7,10,13 */
static const int OBJ_TYPES[4] = { T_BOX, T_BALL, T_KEY, T_DOOR };
enum { LOC_LEFT = 0, LOC_RIGHT = 1, LOC_FRONT = 2, LOC_BEHIND = 3 };

#define VIEW 7
#define OBS_BYTES (VIEW * VIEW * 3)
#define MAXW 25
#define MAXCELLS (MAXW * MAXW)
#define MAXOBJ 64
#define MAXROOM 16
#define MAXDESC 8
#define MAXNODE 8
#define NONE (-1)
#define WALL (-2)

/* "exceptions" */
#define OK 0
#define REJECT 1      /* levelgen.RejectSampling */
#define RECURSION 2   /* RecursionError from bounded rejection loops */
#define TRY(x) do { int _r = (x); if (_r) return _r; } while (0)

/* level kinds */
enum { KIND_REDBALL = 0, KIND_OBJ = 1, KIND_LEVELGEN = 2, KIND_IMPUNLOCK = 3, KIND_UNLOCK = 4 };
enum { I_GOTO = 0, I_PICKUP = 1, I_OPEN = 2, I_PUTNEXT = 3, I_BEFORE = 4, I_AFTER = 5, I_AND = 6 };
enum { K_ACTION = 0, K_AND = 1, K_SEQ = 2 };

typedef struct {
    int32_t kind;
    int32_t room_size, num_rows, num_cols, num_dists;   /* KIND_IMPUNLOCK / KIND_UNLOCK: num_dists per unlocked room */
    int32_t instr;            /* KIND_OBJ: I_GOTO / I_PICKUP / I_OPEN / I_PUTNEXT */
    int32_t doors_open;       /* Level_GoTo(doors_open=...) */
    int32_t grey_dists;       /* Level_GoToRedBallGrey */
    double  locked_room_prob; /* LevelGen ... */
    int32_t locations, unblocking, implicit_unlock;
    int32_t n_action_kinds; int32_t action_kinds[4];
    int32_t n_instr_kinds;  int32_t instr_kinds[3];
    int32_t all_unique;           /* add_distractors(all_unique=...) of the KIND_OBJ levels */
    int32_t require_unreachable;  /* Level_UnblockPickup: reject when every object IS reachable */
} LevelSpec;

typedef struct { int type, color, is_open, is_locked, cur_x, cur_y; } Obj;

typedef struct {
    int top_x, top_y, size;
    int doors[4];       /* NONE or object id */
    int has_pos[4], door_x[4], door_y[4];
    int neighbors[4];   /* room index or NONE */
    int locked;
    int nobjs; int objs[MAXOBJ];
} Room;

typedef struct {
    int type, color, loc;            /* NONE = unspecified */
    int nset; int set[MAXOBJ];       /* obj_set: object identities */
    int nposs; int px[MAXOBJ], py[MAXOBJ]; /* obj_poss: snapshot positions */
} Desc;

typedef struct {
    int kind; int a, b; int desc, desc2;
    int pre_carrying;
    int a_done, b_done;              /* 0 = False, 1 = 'continue', 2 = 'success' */
} Node;

typedef struct {
    LevelSpec sp;
    int W, H;
    /* RNG */
    uint64_t seed, draws;
    /* grid + objects */
    int cell[MAXCELLS];
    int nobj; Obj obj[MAXOBJ];
    int nroom; Room room[MAXROOM];
    int agent_x, agent_y, agent_dir, agent_placed;
    int carrying;
    int step_count, max_steps;
    /* instruction */
    int ndesc; Desc desc[MAXDESC];
    int nnode; Node node[MAXNODE];
    int root;
    /* LevelGen.locked_room: persists across episodes (levelgen.py:284) */
    int locked_room_idx;       /* NONE = None */
    int locked_room_serial;    /* attempt serial in which it was set */
    int attempt_serial;
    char mission[512];
    uint64_t n_attempts;
} Env;

typedef struct { int n; Env *env; } Pool;

/* ---- Philox4x32-10 (oracle/philox.py) ------------------------------------ */
static void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
{
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static uint32_t rng_u32(Env *e)
{
    uint64_t i = e->draws++;
    uint64_t blk = i >> 2;
    uint32_t out[4];
    philox((uint32_t)blk, (uint32_t)(blk >> 32), 0, 0, (uint32_t)e->seed, (uint32_t)(e->seed >> 32), out);
    return out[i & 3];
}

/* MiniGridEnv._rand_int / _rand_float / _rand_bool (App. A.3) */
static int rand_int(Env *e, int lo, int hi)
{
    uint32_t n = (uint32_t)(hi - lo);
    if (n == 1) return lo;                   /* consumes nothing */
    return lo + (int)(((uint64_t)rng_u32(e) * n) >> 32);
}
static double rand_float01(Env *e) { return (double)rng_u32(e) * (1.0 / 4294967296.0); }
static int rand_bool(Env *e) { return rand_int(e, 0, 2) == 0; }

/* ---- Grid ------------------------------------------------------------------ */
static int cell_get(const Env *e, int x, int y) { return e->cell[y * e->W + x]; }
static void cell_set(Env *e, int x, int y, int v) { e->cell[y * e->W + x] = v; }

static void wall_rect(Env *e, int x, int y, int w, int h)
{
    for (int i = 0; i < w; i++) { cell_set(e, x + i, y, WALL); cell_set(e, x + i, y + h - 1, WALL); }
    for (int j = 0; j < h; j++) { cell_set(e, x, y + j, WALL); cell_set(e, x + w - 1, y + j, WALL); }
}

static int cell_type(const Env *e, int c) { return c == NONE ? T_EMPTY : c == WALL ? T_WALL : e->obj[c].type; }

/* WorldObj predicates (App. A.2) */
static int can_overlap(const Env *e, int c) { return c >= 0 && e->obj[c].type == T_DOOR && e->obj[c].is_open; }
static int can_pickup(const Env *e, int c) { return c >= 0 && e->obj[c].type != T_DOOR; }
static int see_behind(const Env *e, int c)
{
    if (c == WALL) return 0;
    if (c >= 0 && e->obj[c].type == T_DOOR) return e->obj[c].is_open;
    return 1;
}
static void encode_cell(const Env *e, int c, uint8_t out[3])
{
    if (c == NONE) { out[0] = T_EMPTY; out[1] = 0; out[2] = 0; return; }
    if (c == WALL) { out[0] = T_WALL; out[1] = C_GREY; out[2] = 0; return; }
    const Obj *o = &e->obj[c];
    out[0] = (uint8_t)o->type; out[1] = (uint8_t)o->color; out[2] = 0;
    if (o->type == T_DOOR) out[2] = o->is_open ? 0 : o->is_locked ? 2 : 1;
}

/* ---- RoomGrid (App. A.6) ------------------------------------------------- */
static Room *get_room(Env *e, int i, int j) { return &e->room[j * e->sp.num_cols + i]; }
static int room_index_from_pos(const Env *e, int x, int y)
{
    int i = x / (e->sp.room_size - 1), j = y / (e->sp.room_size - 1);
    return j * e->sp.num_cols + i;
}
static int room_pos_inside(const Room *r, int x, int y)
{
    if (x < r->top_x || y < r->top_y) return 0;
    if (x >= r->top_x + r->size || y >= r->top_y + r->size) return 0;
    return 1;
}

/* RoomGrid._gen_grid */
static void roomgrid_gen_grid(Env *e)
{
    const int S = e->sp.room_size, R = e->sp.num_rows, C = e->sp.num_cols;
    for (int k = 0; k < e->W * e->H; k++) e->cell[k] = NONE;
    e->nobj = 0;
    e->nroom = R * C;
    for (int j = 0; j < R; j++)
        for (int i = 0; i < C; i++) {
            Room *r = get_room(e, i, j);
            memset(r, 0, sizeof *r);
            r->top_x = i * (S - 1); r->top_y = j * (S - 1); r->size = S;
            for (int k = 0; k < 4; k++) { r->doors[k] = NONE; r->neighbors[k] = NONE; }
            wall_rect(e, r->top_x, r->top_y, S, S);
        }
    for (int j = 0; j < R; j++)
        for (int i = 0; i < C; i++) {
            Room *r = get_room(e, i, j);
            int x_l = r->top_x + 1, y_l = r->top_y + 1;
            int x_m = r->top_x + S - 1, y_m = r->top_y + S - 1;
            if (i < C - 1) {
                r->neighbors[0] = j * C + i + 1;
                r->has_pos[0] = 1; r->door_x[0] = x_m; r->door_y[0] = rand_int(e, y_l, y_m);
            }
            if (j < R - 1) {
                r->neighbors[1] = (j + 1) * C + i;
                r->has_pos[1] = 1; r->door_x[1] = rand_int(e, x_l, x_m); r->door_y[1] = y_m;
            }
            if (i > 0) {
                Room *n = get_room(e, i - 1, j);
                r->neighbors[2] = j * C + i - 1;
                r->has_pos[2] = 1; r->door_x[2] = n->door_x[0]; r->door_y[2] = n->door_y[0];
            }
            if (j > 0) {
                Room *n = get_room(e, i, j - 1);
                r->neighbors[3] = (j - 1) * C + i;
                r->has_pos[3] = 1; r->door_x[3] = n->door_x[1]; r->door_y[3] = n->door_y[1];
            }
        }
    /* The agent starts in the middle, facing right */
    e->agent_x = (C / 2) * (S - 1) + S / 2;
    e->agent_y = (R / 2) * (S - 1) + S / 2;
    e->agent_dir = 0;
    e->agent_placed = 1;
}

/* MiniGridEnv.place_obj (App. A.3); obj = NONE places nothing (place_agent).
 * reject_next_to: Manhattan distance to the agent's CURRENT position < 2. */
static int place_obj(Env *e, int obj, int top_x, int top_y, int size, int use_reject, int max_tries, int *ox, int *oy)
{
    if (top_x < 0) top_x = 0;
    if (top_y < 0) top_y = 0;
    int num_tries = 0, x, y;
    for (;;) {
        if (num_tries > max_tries) return RECURSION;
        num_tries++;
        int hx = top_x + size < e->W ? top_x + size : e->W;
        int hy = top_y + size < e->H ? top_y + size : e->H;
        x = rand_int(e, top_x, hx);
        y = rand_int(e, top_y, hy);
        if (cell_get(e, x, y) != NONE) continue;
        if (e->agent_placed && x == e->agent_x && y == e->agent_y) continue;
        if (use_reject) {
            int d = abs(e->agent_x - x) + abs(e->agent_y - y);
            if (d < 2) continue;
        }
        break;
    }
    if (obj != NONE) {
        cell_set(e, x, y, obj);
        e->obj[obj].cur_x = x; e->obj[obj].cur_y = y;
    }
    *ox = x; *oy = y;
    return OK;
}

static int new_obj(Env *e, int type, int color)
{
    int id = e->nobj++;
    Obj *o = &e->obj[id];
    o->type = type; o->color = color; o->is_open = 0; o->is_locked = 0; o->cur_x = o->cur_y = -1;
    return id;
}

/* RoomGrid.place_in_room / add_object */
static int add_object(Env *e, int i, int j, int type, int color, int *out_id)
{
    Room *r = get_room(e, i, j);
    int id = new_obj(e, type, color), x, y;
    TRY(place_obj(e, id, r->top_x, r->top_y, r->size, 1, 1000, &x, &y));
    r->objs[r->nobjs++] = id;
    if (out_id) *out_id = id;
    return OK;
}

/* RoomGrid.add_door (door_idx, color, locked all given or drawn) */
static int add_door(Env *e, int i, int j, int door_idx, int color, int locked)
{
    Room *r = get_room(e, i, j);
    if (door_idx == NONE) {
        for (;;) {
            door_idx = rand_int(e, 0, 4);
            if (r->neighbors[door_idx] != NONE && r->doors[door_idx] == NONE) break;
        }
    }
    if (color == NONE) color = COLOR_NAMES[rand_int(e, 0, 6)];
    if (locked == NONE) locked = rand_bool(e);
    r->locked = locked;
    int id = new_obj(e, T_DOOR, color);
    e->obj[id].is_locked = locked;
    int x = r->door_x[door_idx], y = r->door_y[door_idx];
    cell_set(e, x, y, id);
    e->obj[id].cur_x = x; e->obj[id].cur_y = y;
    r->doors[door_idx] = id;
    e->room[r->neighbors[door_idx]].doors[(door_idx + 2) % 4] = id;
    return id;
}

/* RoomGrid.place_agent(i=None, j=None, rand_dir=True) */
static int place_agent(Env *e)
{
    int i = rand_int(e, 0, e->sp.num_cols);
    int j = rand_int(e, 0, e->sp.num_rows);
    Room *r = get_room(e, i, j);
    {   /* KNOWN DIVERGENCE: the reference's loop below (roomgrid.py place_agent, `while True`) never returns when no
         * empty cell of the room has an empty or wall cell in front of it for any heading -- it happens with 3x3
         * rooms packed with distractors and doors (MiniBossLevel, seed 698, 57th level; reproduced with the
         * reference itself: tests/test_oracle_vs_reference.py::test_reference_place_agent_hang_is_rejected).
         * The reference hangs there; the oracle (and the kernels) reject the level like any other failed rejection
         * sampling.  Levels the reference can generate are unaffected. */
        int any = 0;
        for (int y = r->top_y + 1; y < r->top_y + r->size - 1 && !any; y++)
            for (int x = r->top_x + 1; x < r->top_x + r->size - 1 && !any; x++) {
                if (cell_get(e, x, y) != NONE) continue;
                for (int d = 0; d < 4; d++) {
                    int fc = cell_get(e, x + DIR_X[d], y + DIR_Y[d]);
                    if (fc == NONE || fc == WALL) any = 1;
                }
            }
        if (!any) return RECURSION;
    }
    for (;;) {
        /* MiniGridEnv.place_agent: agent_pos = None while sampling */
        int x, y;
        e->agent_placed = 0;
        TRY(place_obj(e, NONE, r->top_x, r->top_y, r->size, 0, 1000, &x, &y));
        e->agent_x = x; e->agent_y = y; e->agent_placed = 1;
        e->agent_dir = rand_int(e, 0, 4);
        int fc = cell_get(e, x + DIR_X[e->agent_dir], y + DIR_Y[e->agent_dir]);
        if (fc == NONE || fc == WALL) break;
    }
    return OK;
}

/* RoomGrid.connect_all(door_colors=COLOR_NAMES, max_itrs=5000) */
static int connect_all_colors(Env *e, int exclude_color);
static int connect_all(Env *e) { return connect_all_colors(e, NONE); }
/* connect_all(door_colors=[c for c in COLOR_NAMES if c is not exclude_color]) (iclr19_levels.py:441-446) */
static int connect_all_colors(Env *e, int exclude_color)
{
    int start = room_index_from_pos(e, e->agent_x, e->agent_y);
    int num_itrs = 0;
    for (;;) {
        if (num_itrs > 5000) return RECURSION;
        num_itrs++;
        /* find_reach: DFS over rooms through doors[k] */
        int reach[MAXROOM] = { 0 }, stack[MAXROOM * 5], sp = 0, nreach = 0;
        stack[sp++] = start;
        while (sp > 0) {
            int r = stack[--sp];
            if (reach[r]) continue;
            reach[r] = 1; nreach++;
            for (int k = 0; k < 4; k++)
                if (e->room[r].doors[k] != NONE) stack[sp++] = e->room[r].neighbors[k];
        }
        if (nreach == e->sp.num_rows * e->sp.num_cols) break;
        int i = rand_int(e, 0, e->sp.num_cols);
        int j = rand_int(e, 0, e->sp.num_rows);
        int k = rand_int(e, 0, 4);
        Room *r = get_room(e, i, j);
        if (!r->has_pos[k] || r->doors[k] != NONE) continue;
        if (r->locked || e->room[r->neighbors[k]].locked) continue;
        int color;
        if (exclude_color == NONE) color = COLOR_NAMES[rand_int(e, 0, 6)];
        else {                                  /* _rand_elem over the five remaining names, same order */
            int pick = rand_int(e, 0, 5);
            for (int c = 0, n = 0; ; c++) if (COLOR_NAMES[c] != exclude_color && n++ == pick) { color = COLOR_NAMES[c]; break; }
        }
        add_door(e, i, j, k, color, 0);
    }
    return OK;
}

/* RoomGrid.add_distractors(i=None, j=None, num, all_unique); ids appended to out */
static int add_distractors_at(Env *e, int i, int j, int num, int all_unique, int *out, int *nout);
static int add_distractors(Env *e, int num, int all_unique, int *out, int *nout)
{
    return add_distractors_at(e, NONE, NONE, num, all_unique, out, nout);
}
/* i / j == NONE: the room column / row is drawn per object */
static int add_distractors_at(Env *e, int i, int j, int num, int all_unique, int *out, int *nout)
{
    int seen_t[MAXOBJ * 2], seen_c[MAXOBJ * 2], nseen = 0;
    for (int r = 0; r < e->nroom; r++)
        for (int k = 0; k < e->room[r].nobjs; k++) {
            seen_t[nseen] = e->obj[e->room[r].objs[k]].type;
            seen_c[nseen++] = e->obj[e->room[r].objs[k]].color;
        }
    int n = 0;
    while (n < num) {
        int color = COLOR_NAMES[rand_int(e, 0, 6)];
        int type = KBB[rand_int(e, 0, 3)];
        if (all_unique) {
            int dup = 0;
            for (int k = 0; k < nseen; k++) if (seen_t[k] == type && seen_c[k] == color) dup = 1;
            if (dup) continue;
        }
        int ri = i != NONE ? i : rand_int(e, 0, e->sp.num_cols);
        int rj = j != NONE ? j : rand_int(e, 0, e->sp.num_rows);
        int id;
        TRY(add_object(e, ri, rj, type, color, &id));
        seen_t[nseen] = type; seen_c[nseen++] = color;
        if (out) out[n] = id;
        n++;
    }
    if (nout) *nout = n;
    return OK;
}

/* ---- levelgen.py:201-253 check_objs_reachable ------------------------------ */
static int check_objs_reachable(Env *e)
{
    static __thread uint8_t reachable[MAXCELLS];
    static __thread int stack[MAXCELLS * 4 + 4];
    memset(reachable, 0, (size_t)(e->W * e->H));
    int sp = 0;
    stack[sp++] = e->agent_y * e->W + e->agent_x;
    while (sp > 0) {
        int p = stack[--sp];
        int i = p % e->W, j = p / e->W;
        if (reachable[p]) continue;
        reachable[p] = 1;
        int c = e->cell[p];
        /* anything but a door blocks */
        if (c != NONE && !(c >= 0 && e->obj[c].type == T_DOOR)) continue;
        if (i + 1 < e->W) stack[sp++] = p + 1;
        if (i - 1 >= 0) stack[sp++] = p - 1;
        if (j + 1 < e->H) stack[sp++] = p + e->W;
        if (j - 1 >= 0) stack[sp++] = p - e->W;
    }
    for (int p = 0; p < e->W * e->H; p++) {
        int c = e->cell[p];
        if (c == NONE || c == WALL) continue;
        if (!reachable[p]) return REJECT;
    }
    return OK;
}

/* ---- verifier.py:96-161 ObjDesc.find_matching_objs -------------------------- */
static void find_matching_objs(Env *e, Desc *d, int use_location)
{
    if (use_location) d->nset = 0;
    d->nposs = 0;
    int agent_room = room_index_from_pos(e, e->agent_x, e->agent_y);
    for (int i = 0; i < e->W; i++)
        for (int j = 0; j < e->H; j++) {
            int c = cell_get(e, i, j);
            if (c == NONE) continue;
            if (!use_location) {
                int tracked = 0;
                for (int k = 0; k < d->nset; k++) if (d->set[k] == c) tracked = 1;
                if (!tracked) continue;
            }
            int ctype = cell_type(e, c);
            int ccolor = c == WALL ? C_GREY : e->obj[c].color;
            if (d->type != NONE && ctype != d->type) continue;
            if (d->color != NONE && ccolor != d->color) continue;
            if (use_location && d->loc != NONE) {
                if (!room_pos_inside(&e->room[agent_room], i, j)) continue;
                int vx = i - e->agent_x, vy = j - e->agent_y;
                int d1x = DIR_X[e->agent_dir], d1y = DIR_Y[e->agent_dir];
                int d2x = -d1y, d2y = d1x;
                int dot1 = vx * d1x + vy * d1y, dot2 = vx * d2x + vy * d2y;
                int m = d->loc == LOC_LEFT ? dot2 < 0 : d->loc == LOC_RIGHT ? dot2 > 0
                      : d->loc == LOC_FRONT ? dot1 > 0 : dot1 < 0;
                if (!m) continue;
            }
            if (use_location) d->set[d->nset++] = c;
            d->px[d->nposs] = i; d->py[d->nposs] = j; d->nposs++;
        }
}

static int new_desc(Env *e, int type, int color, int loc)
{
    int id = e->ndesc++;
    Desc *d = &e->desc[id];
    d->type = type; d->color = color; d->loc = loc; d->nset = 0; d->nposs = 0;
    return id;
}
static int new_node(Env *e, int kind, int a, int b, int desc, int desc2)
{
    int id = e->nnode++;
    Node *n = &e->node[id];
    n->kind = kind; n->a = a; n->b = b; n->desc = desc; n->desc2 = desc2;
    n->pre_carrying = NONE; n->a_done = n->b_done = 0;
    return id;
}

/* ---- verifier.py reset_verifier (:251-255,290-294,321-328,369-377,442-447..) */
static void reset_verifier(Env *e, int n)
{
    Node *nd = &e->node[n];
    switch (nd->kind) {
    case I_OPEN: case I_GOTO:
        find_matching_objs(e, &e->desc[nd->desc], 1); break;
    case I_PICKUP:
        nd->pre_carrying = NONE; find_matching_objs(e, &e->desc[nd->desc], 1); break;
    case I_PUTNEXT:
        nd->pre_carrying = NONE;
        find_matching_objs(e, &e->desc[nd->desc], 1);
        find_matching_objs(e, &e->desc[nd->desc2], 1); break;
    default:
        reset_verifier(e, nd->a); reset_verifier(e, nd->b);
        nd->a_done = 0; nd->b_done = 0; break;
    }
}

/* levelgen.py:68-75 + verifier.py:195-202 update_objs_poss */
static void update_objs_poss(Env *e, int n)
{
    Node *nd = &e->node[n];
    if (nd->kind >= I_BEFORE) { update_objs_poss(e, nd->a); update_objs_poss(e, nd->b); return; }
    find_matching_objs(e, &e->desc[nd->desc], 0);
    if (nd->kind == I_PUTNEXT) find_matching_objs(e, &e->desc[nd->desc2], 0);
}

enum { V_CONTINUE = 1, V_SUCCESS = 2 };

/* verifier.py verify: Open :257-274, GoTo :296-303, Pickup :330-350,
 * PutNext :393-417, Before :449-471, After :490-512, And :536-550
 * (strict=False and use_done_actions=False: the defaults of every ICLR level) */
static int verify(Env *e, int n, int action)
{
    Node *nd = &e->node[n];
    int fx = e->agent_x + DIR_X[e->agent_dir], fy = e->agent_y + DIR_Y[e->agent_dir];
    switch (nd->kind) {
    case I_OPEN: {
        if (action != A_TOGGLE) return V_CONTINUE;
        int fc = cell_get(e, fx, fy);
        Desc *d = &e->desc[nd->desc];
        for (int k = 0; k < d->nset; k++)
            if (fc != NONE && fc == d->set[k] && e->obj[fc].is_open) return V_SUCCESS;
        return V_CONTINUE;
    }
    case I_GOTO: {
        Desc *d = &e->desc[nd->desc];
        for (int k = 0; k < d->nposs; k++)
            if (d->px[k] == fx && d->py[k] == fy) return V_SUCCESS;
        return V_CONTINUE;
    }
    case I_PICKUP: {
        int pre = nd->pre_carrying;
        nd->pre_carrying = e->carrying;
        if (action != A_PICKUP) return V_CONTINUE;
        Desc *d = &e->desc[nd->desc];
        for (int k = 0; k < d->nset; k++)
            if (pre == NONE && e->carrying == d->set[k]) return V_SUCCESS;
        nd->pre_carrying = e->carrying;
        return V_CONTINUE;
    }
    case I_PUTNEXT: {
        int pre = nd->pre_carrying;
        nd->pre_carrying = e->carrying;
        if (action != A_DROP) return V_CONTINUE;
        Desc *dm = &e->desc[nd->desc], *df = &e->desc[nd->desc2];
        for (int k = 0; k < dm->nset; k++) {
            int oa = dm->set[k];
            if (pre != oa) continue;
            int ax = e->obj[oa].cur_x, ay = e->obj[oa].cur_y;
            for (int q = 0; q < df->nposs; q++)
                if (abs(ax - df->px[q]) + abs(ay - df->py[q]) == 1) return V_SUCCESS;
        }
        return V_CONTINUE;
    }
    case I_BEFORE:
        if (nd->a_done == V_SUCCESS) {
            nd->b_done = verify(e, nd->b, action);
            if (nd->b_done == V_SUCCESS) return V_SUCCESS;
        } else {
            nd->a_done = verify(e, nd->a, action);
            if (nd->a_done == V_SUCCESS) return verify(e, n, action);
        }
        return V_CONTINUE;
    case I_AFTER:
        if (nd->b_done == V_SUCCESS) {
            nd->a_done = verify(e, nd->a, action);
            if (nd->a_done == V_SUCCESS) return V_SUCCESS;
        } else {
            nd->b_done = verify(e, nd->b, action);
            if (nd->b_done == V_SUCCESS) return verify(e, n, action);
        }
        return V_CONTINUE;
    case I_AND:
        if (nd->a_done != V_SUCCESS) nd->a_done = verify(e, nd->a, action);
        if (nd->b_done != V_SUCCESS) nd->b_done = verify(e, nd->b, action);
        if (nd->a_done == V_SUCCESS && nd->b_done == V_SUCCESS) return V_SUCCESS;
        return V_CONTINUE;
    }
    return V_CONTINUE;
}

/* ---- verifier.py surface(): ObjDesc :64-94, instrs :248,288,319,367,440,481,527 */
static void desc_surface(Env *e, Desc *d, char *out)
{
    find_matching_objs(e, d, 1);
    char s[128] = "";
    if (d->color != NONE) { strcat(s, COLOR_STR[d->color]); strcat(s, " "); }
    strcat(s, d->type != NONE ? TYPE_STR[d->type] : "object");
    if (d->loc == LOC_FRONT) strcat(s, " in front of you");
    else if (d->loc == LOC_BEHIND) strcat(s, " behind you");
    else if (d->loc == LOC_LEFT) strcat(s, " on your left");
    else if (d->loc == LOC_RIGHT) strcat(s, " on your right");
    strcat(out, d->nset > 1 ? "a " : "the ");
    strcat(out, s);
}
static void surface(Env *e, int n, char *out)
{
    Node *nd = &e->node[n];
    switch (nd->kind) {
    case I_OPEN: strcat(out, "open "); desc_surface(e, &e->desc[nd->desc], out); break;
    case I_GOTO: strcat(out, "go to "); desc_surface(e, &e->desc[nd->desc], out); break;
    case I_PICKUP: strcat(out, "pick up "); desc_surface(e, &e->desc[nd->desc], out); break;
    case I_PUTNEXT:
        strcat(out, "put "); desc_surface(e, &e->desc[nd->desc], out);
        strcat(out, " next to "); desc_surface(e, &e->desc[nd->desc2], out); break;
    case I_BEFORE: surface(e, nd->a, out); strcat(out, ", then "); surface(e, nd->b, out); break;
    case I_AFTER: surface(e, nd->a, out); strcat(out, " after you "); surface(e, nd->b, out); break;
    case I_AND: surface(e, nd->a, out); strcat(out, " and "); surface(e, nd->b, out); break;
    }
}

/* ---- levelgen.py:354-395 LevelGen.rand_obj ----------------------------------- */
static int rand_obj(Env *e, const int *types, int ntypes, int *out_desc)
{
    int num_tries = 0;
    int id = new_desc(e, NONE, NONE, NONE);   /* slot reused across tries */
    for (;;) {
        if (num_tries > 100) return RECURSION;
        num_tries++;
        int ci = rand_int(e, 0, 7);                       /* [None, *COLOR_NAMES] */
        int color = ci == 0 ? NONE : COLOR_NAMES[ci - 1];
        int type = types[rand_int(e, 0, ntypes)];
        int loc = NONE;
        if (e->sp.locations && rand_bool(e)) loc = rand_int(e, 0, 4);   /* LOC_NAMES */
        Desc *d = &e->desc[id];
        d->type = type; d->color = color; d->loc = loc;
        find_matching_objs(e, d, 1);
        if (d->nset == 0) continue;
        if (!e->sp.implicit_unlock && e->locked_room_idx != NONE) {
            const Room *lr = &e->room[e->locked_room_idx];   /* geometry of the (possibly stale) room */
            int n_not_locked = 0;
            for (int k = 0; k < d->nposs; k++) if (!room_pos_inside(lr, d->px[k], d->py[k])) n_not_locked++;
            if (n_not_locked == 0) continue;
        }
        *out_desc = id;
        return OK;
    }
}

/* levelgen.py:397-460 LevelGen.rand_instr */
static int rand_instr(Env *e, const int *action_kinds, int n_ak, const int *instr_kinds, int n_ik, int *out_node)
{
    int kind = instr_kinds[rand_int(e, 0, n_ik)];
    if (kind == K_ACTION) {
        int action = action_kinds[rand_int(e, 0, n_ak)];
        int d1, d2;
        if (action == I_GOTO) { TRY(rand_obj(e, OBJ_TYPES, 4, &d1)); *out_node = new_node(e, I_GOTO, NONE, NONE, d1, NONE); }
        else if (action == I_PICKUP) { TRY(rand_obj(e, OBJ_TYPES, 3, &d1)); *out_node = new_node(e, I_PICKUP, NONE, NONE, d1, NONE); }
        else if (action == I_OPEN) { static const int door_only[1] = { T_DOOR }; TRY(rand_obj(e, door_only, 1, &d1)); *out_node = new_node(e, I_OPEN, NONE, NONE, d1, NONE); }
        else { TRY(rand_obj(e, OBJ_TYPES, 3, &d1)); TRY(rand_obj(e, OBJ_TYPES, 4, &d2)); *out_node = new_node(e, I_PUTNEXT, NONE, NONE, d1, d2); }
        return OK;
    }
    if (kind == K_AND) {
        static const int only_action[1] = { K_ACTION };
        int a, b;
        TRY(rand_instr(e, action_kinds, n_ak, only_action, 1, &a));
        TRY(rand_instr(e, action_kinds, n_ak, only_action, 1, &b));
        *out_node = new_node(e, I_AND, a, b, NONE, NONE);
        return OK;
    }
    {
        static const int action_and[2] = { K_ACTION, K_AND };
        int a, b;
        TRY(rand_instr(e, action_kinds, n_ak, action_and, 2, &a));
        TRY(rand_instr(e, action_kinds, n_ak, action_and, 2, &b));
        int which = rand_int(e, 0, 2);   /* ['before', 'after'] */
        *out_node = new_node(e, which == 0 ? I_BEFORE : I_AFTER, a, b, NONE, NONE);
        return OK;
    }
}

/* levelgen.py:104-155 validate_instrs */
static int validate_instrs(Env *e, int n)
{
    Node *nd = &e->node[n];
    int locked_colors[MAXOBJ], nlc = 0;
    if (e->sp.kind == KIND_LEVELGEN && e->sp.unblocking) {
        for (int i = 0; i < e->sp.num_cols; i++)
            for (int j = 0; j < e->sp.num_rows; j++) {
                Room *r = get_room(e, i, j);
                for (int k = 0; k < 4; k++)
                    if (r->doors[k] != NONE && e->obj[r->doors[k]].is_locked) locked_colors[nlc++] = e->obj[r->doors[k]].color;
            }
    }
    if (nd->kind == I_PUTNEXT) {
        reset_verifier(e, n);
        Desc *dm = &e->desc[nd->desc], *df = &e->desc[nd->desc2];
        for (int a = 0; a < dm->nset; a++)
            for (int b = 0; b < df->nset; b++)
                if (dm->set[a] == df->set[b]) return REJECT;
        /* objs_next() verifier.py:379-391 */
        for (int a = 0; a < dm->nset; a++)
            for (int q = 0; q < df->nposs; q++)
                if (abs(e->obj[dm->set[a]].cur_x - df->px[q]) + abs(e->obj[dm->set[a]].cur_y - df->py[q]) == 1) return REJECT;
        if (dm->nset == 1 && df->nset == 1 && dm->set[0] == df->set[0]) return REJECT;
    }
    if (nd->kind <= I_PUTNEXT) {
        if (!(e->sp.kind == KIND_LEVELGEN && e->sp.unblocking)) return OK;
        int ds[2] = { nd->desc, nd->kind == I_PUTNEXT ? nd->desc2 : NONE };
        for (int q = 0; q < 2; q++) {
            if (ds[q] == NONE) continue;
            Desc *d = &e->desc[ds[q]];
            if (d->type == T_KEY)
                for (int k = 0; k < nlc; k++) if (d->color == locked_colors[k]) return REJECT;
        }
        return OK;
    }
    TRY(validate_instrs(e, nd->a));
    TRY(validate_instrs(e, nd->b));
    return OK;
}

/* levelgen.py:172-187 num_navs_needed */
static int num_navs_needed(Env *e, int n)
{
    Node *nd = &e->node[n];
    if (nd->kind == I_PUTNEXT) return 2;
    if (nd->kind <= I_OPEN) return 1;
    return num_navs_needed(e, nd->a) + num_navs_needed(e, nd->b);
}

/* levelgen.py:321-352 LevelGen.add_locked_room */
static int add_locked_room(Env *e)
{
    int door;
    for (;;) {
        int i = rand_int(e, 0, e->sp.num_cols);
        int j = rand_int(e, 0, e->sp.num_rows);
        int door_idx = rand_int(e, 0, 4);
        e->locked_room_idx = j * e->sp.num_cols + i;
        e->locked_room_serial = e->attempt_serial;
        if (get_room(e, i, j)->neighbors[door_idx] == NONE) continue;
        door = add_door(e, i, j, door_idx, NONE, 1);
        break;
    }
    for (;;) {
        int i = rand_int(e, 0, e->sp.num_cols);
        int j = rand_int(e, 0, e->sp.num_rows);
        if (j * e->sp.num_cols + i == e->locked_room_idx) continue;
        TRY(add_object(e, i, j, T_KEY, e->obj[door].color, NULL));
        break;
    }
    return OK;
}

/* gen_mission of the supported level families */
static int gen_mission(Env *e)
{
    const LevelSpec *sp = &e->sp;
    e->ndesc = 0; e->nnode = 0;
    if (sp->kind == KIND_REDBALL) {
        /* iclr19_levels.py:26-37 (Grey) / :55-63 */
        int obj, dists[MAXOBJ], nd;
        TRY(place_agent(e));
        TRY(add_object(e, 0, 0, T_BALL, C_RED, &obj));
        TRY(add_distractors(e, sp->num_dists, 0, dists, &nd));
        if (sp->grey_dists) for (int k = 0; k < nd; k++) e->obj[dists[k]].color = C_GREY;
        TRY(check_objs_reachable(e));
        e->root = new_node(e, I_GOTO, NONE, NONE, new_desc(e, e->obj[obj].type, e->obj[obj].color, NONE), NONE);
        return OK;
    }
    if (sp->kind == KIND_OBJ) {
        /* iclr19_levels.py:88-92 GoToObj, :119-124 GoToLocal, :202-211 PutNextLocal, :247-257 GoTo, :365-371 Pickup,
         * :380-391 UnblockPickup, :399-415 Open, :482-491 PutNext */
        int objs[MAXOBJ], n;
        TRY(place_agent(e));
        TRY(connect_all(e));
        TRY(add_distractors(e, sp->num_dists, sp->all_unique, objs, &n));
        if (sp->require_unreachable) { if (check_objs_reachable(e) == OK) return REJECT; }
        else TRY(check_objs_reachable(e));
        if (sp->instr == I_OPEN) {
            /* every door once per adjacent room: columns outer, rows inner, sides right/down/left/up */
            int doors[MAXROOM * 4], nd = 0;
            for (int i = 0; i < sp->num_cols; i++)
                for (int j = 0; j < sp->num_rows; j++)
                    for (int k = 0; k < 4; k++)
                        if (get_room(e, i, j)->doors[k] != NONE) doors[nd++] = get_room(e, i, j)->doors[k];
            int door = doors[rand_int(e, 0, nd)];
            e->root = new_node(e, I_OPEN, NONE, NONE, new_desc(e, T_DOOR, e->obj[door].color, NONE), NONE);
        } else if (sp->instr == I_PUTNEXT) {
            /* o1, o2 = self._rand_subset(objs, 2) */
            int i1 = rand_int(e, 0, n);
            int o1 = objs[i1];
            for (int k = i1; k + 1 < n; k++) objs[k] = objs[k + 1];
            int o2 = objs[rand_int(e, 0, n - 1)];
            e->root = new_node(e, I_PUTNEXT, NONE, NONE, new_desc(e, e->obj[o1].type, e->obj[o1].color, NONE),
                               new_desc(e, e->obj[o2].type, e->obj[o2].color, NONE));
        } else {
            int obj = objs[rand_int(e, 0, n)];
            e->root = new_node(e, sp->instr, NONE, NONE, new_desc(e, e->obj[obj].type, e->obj[obj].color, NONE), NONE);
        }
        if (sp->doors_open)   /* levelgen.py:189-199 open_all_doors */
            for (int k = 0; k < e->nobj; k++) if (e->obj[k].type == T_DOOR) e->obj[k].is_open = 1;
        return OK;
    }
    if (sp->kind == KIND_UNLOCK) {
        /* iclr19_levels.py:418-474 Level_Unlock.gen_mission; num_dists = distractors per unlocked room (3) */
        const int id = rand_int(e, 0, sp->num_cols);
        const int jd = rand_int(e, 0, sp->num_rows);
        const int door = add_door(e, id, jd, NONE, NONE, 1);
        for (;;) {
            int ik = rand_int(e, 0, sp->num_cols);
            int jk = rand_int(e, 0, sp->num_rows);
            if (ik == id && jk == jd) continue;
            TRY(add_object(e, ik, jk, T_KEY, e->obj[door].color, NULL));
            break;
        }
        /* with probability 1/2 the locked door is the only door of its colour */
        if (rand_bool(e)) TRY(connect_all_colors(e, e->obj[door].color));
        else TRY(connect_all(e));
        for (int i = 0; i < sp->num_cols; i++)
            for (int j = 0; j < sp->num_rows; j++)
                if (i != id || j != jd) TRY(add_distractors_at(e, i, j, sp->num_dists, 0, NULL, NULL));
        for (;;) {
            TRY(place_agent(e));
            if (room_index_from_pos(e, e->agent_x, e->agent_y) == jd * sp->num_cols + id) continue;
            break;
        }
        TRY(check_objs_reachable(e));
        e->root = new_node(e, I_OPEN, NONE, NONE, new_desc(e, T_DOOR, e->obj[door].color, NONE), NONE);
        return OK;
    }
    if (sp->kind == KIND_IMPUNLOCK) {
        /* iclr19_levels.py:311-355 Level_GoToImpUnlock.gen_mission; num_dists = distractors per unlocked room (2) */
        const int id = rand_int(e, 0, sp->num_cols);
        const int jd = rand_int(e, 0, sp->num_rows);
        const int door = add_door(e, id, jd, NONE, NONE, 1);
        for (;;) {                                    /* the key goes to a different room */
            int ik = rand_int(e, 0, sp->num_cols);
            int jk = rand_int(e, 0, sp->num_rows);
            if (ik == id && jk == jd) continue;
            TRY(add_object(e, ik, jk, T_KEY, e->obj[door].color, NULL));
            break;
        }
        TRY(connect_all(e));
        for (int i = 0; i < sp->num_cols; i++)        /* columns outer, rows inner (:334-342) */
            for (int j = 0; j < sp->num_rows; j++)
                if (i != id || j != jd) TRY(add_distractors_at(e, i, j, sp->num_dists, 0, NULL, NULL));
        for (;;) {
            TRY(place_agent(e));
            if (room_index_from_pos(e, e->agent_x, e->agent_y) == jd * sp->num_cols + id) continue;
            break;
        }
        TRY(check_objs_reachable(e));
        int obj, n;
        TRY(add_distractors_at(e, id, jd, 1, 0, &obj, &n));
        e->root = new_node(e, I_GOTO, NONE, NONE, new_desc(e, e->obj[obj].type, e->obj[obj].color, NONE), NONE);
        return OK;
    }
    /* levelgen.py:293-319 LevelGen.gen_mission */
    if (rand_float01(e) < sp->locked_room_prob) TRY(add_locked_room(e));
    TRY(connect_all(e));
    TRY(add_distractors(e, sp->num_dists, 0, NULL, NULL));
    for (;;) {
        TRY(place_agent(e));
        int start = room_index_from_pos(e, e->agent_x, e->agent_y);
        /* `start_room is self.locked_room`: identity, so only a room of THIS attempt */
        if (e->locked_room_idx != NONE && e->locked_room_serial == e->attempt_serial && start == e->locked_room_idx) continue;
        break;
    }
    if (!sp->unblocking) TRY(check_objs_reachable(e));
    TRY(rand_instr(e, sp->action_kinds, sp->n_action_kinds, sp->instr_kinds, sp->n_instr_kinds, &e->root));
    return OK;
}

/* levelgen.py:77-102 RoomGridLevel._gen_grid */
static void level_gen_grid(Env *e)
{
    for (;;) {
        e->attempt_serial++;
        e->n_attempts++;
        roomgrid_gen_grid(e);
        if (gen_mission(e)) continue;
        if (validate_instrs(e, e->root)) continue;
        break;
    }
    e->mission[0] = 0;
    surface(e, e->root, e->mission);
}

/* ---- observation (App. A.5), literal ------------------------------------------ */
typedef struct { int w, h; int c[VIEW * VIEW]; } View;   /* c[j*w+i] */

static void view_rotate_left(const View *in, View *out)
{
    out->w = in->h; out->h = in->w;
    for (int i = 0; i < in->w; i++)
        for (int j = 0; j < in->h; j++)
            out->c[(out->h - 1 - i) * out->w + j] = in->c[j * in->w + i];
}

static void gen_obs(const Env *e, uint8_t *image)
{
    int topX, topY;
    switch (e->agent_dir) {          /* get_view_exts */
    case 0: topX = e->agent_x; topY = e->agent_y - VIEW / 2; break;
    case 1: topX = e->agent_x - VIEW / 2; topY = e->agent_y; break;
    case 2: topX = e->agent_x - VIEW + 1; topY = e->agent_y - VIEW / 2; break;
    default: topX = e->agent_x - VIEW / 2; topY = e->agent_y - VIEW + 1; break;
    }
    View a, b; a.w = a.h = VIEW;
    for (int j = 0; j < VIEW; j++)          /* Grid.slice: OOB -> Wall() */
        for (int i = 0; i < VIEW; i++) {
            int x = topX + i, y = topY + j;
            a.c[j * VIEW + i] = (x >= 0 && x < e->W && y >= 0 && y < e->H) ? cell_get(e, x, y) : WALL;
        }
    View *cur = &a, *oth = &b;
    for (int r = 0; r < e->agent_dir + 1; r++) { view_rotate_left(cur, oth); View *t = cur; cur = oth; oth = t; }

    /* Grid.process_vis(agent_pos=(3,6)) */
    uint8_t mask[VIEW][VIEW];   /* [i][j] */
    memset(mask, 0, sizeof mask);
    mask[VIEW / 2][VIEW - 1] = 1;
    for (int j = VIEW - 1; j >= 0; j--) {
        for (int i = 0; i < VIEW - 1; i++) {
            if (!mask[i][j]) continue;
            if (!see_behind(e, cur->c[j * VIEW + i])) continue;
            mask[i + 1][j] = 1;
            if (j > 0) { mask[i + 1][j - 1] = 1; mask[i][j - 1] = 1; }
        }
        for (int i = VIEW - 1; i >= 1; i--) {
            if (!mask[i][j]) continue;
            if (!see_behind(e, cur->c[j * VIEW + i])) continue;
            mask[i - 1][j] = 1;
            if (j > 0) { mask[i - 1][j - 1] = 1; mask[i][j - 1] = 1; }
        }
    }
    /* the agent sees what it carries at its own cell */
    cur->c[(VIEW - 1) * VIEW + VIEW / 2] = e->carrying;
    /* Grid.encode(vis_mask): array[i, j, :] */
    for (int i = 0; i < VIEW; i++)
        for (int j = 0; j < VIEW; j++) {
            uint8_t *px = image + (i * VIEW + j) * 3;
            if (mask[i][j]) encode_cell(e, cur->c[j * VIEW + i], px);
            else { px[0] = px[1] = px[2] = 0; }
        }
}

/* ---- RoomGridLevel.reset (levelgen.py:35-47) ----------------------------------- */
static void env_reset(Env *e, uint8_t *image)
{
    level_gen_grid(e);
    e->carrying = NONE;
    e->step_count = 0;
    if (image) gen_obs(e, image);
    reset_verifier(e, e->root);
    int nav_time_room = e->sp.room_size * e->sp.room_size;
    int nav_time_maze = nav_time_room * e->sp.num_rows * e->sp.num_cols;
    e->max_steps = num_navs_needed(e, e->root) * nav_time_maze;
}

/* ---- MiniGridEnv.step (App. A.4) + RoomGridLevel.step (levelgen.py:49-66) ------- */
static void env_step(Env *e, int action, uint8_t *image, float *reward, uint8_t *done)
{
    e->step_count++;
    double rew = 0; int dn = 0;
    int fx = e->agent_x + DIR_X[e->agent_dir], fy = e->agent_y + DIR_Y[e->agent_dir];
    int fc = cell_get(e, fx, fy);
    switch (action) {
    case A_LEFT: e->agent_dir -= 1; if (e->agent_dir < 0) e->agent_dir += 4; break;
    case A_RIGHT: e->agent_dir = (e->agent_dir + 1) % 4; break;
    case A_FORWARD: if (fc == NONE || can_overlap(e, fc)) { e->agent_x = fx; e->agent_y = fy; } break;
    case A_PICKUP:
        if (fc != NONE && can_pickup(e, fc) && e->carrying == NONE) {
            e->carrying = fc; e->obj[fc].cur_x = -1; e->obj[fc].cur_y = -1; cell_set(e, fx, fy, NONE);
        }
        break;
    case A_DROP:
        if (fc == NONE && e->carrying != NONE) {
            cell_set(e, fx, fy, e->carrying);
            e->obj[e->carrying].cur_x = fx; e->obj[e->carrying].cur_y = fy; e->carrying = NONE;
        }
        break;
    case A_TOGGLE:
        if (fc >= 0) {
            Obj *o = &e->obj[fc];
            if (o->type == T_DOOR) {
                if (o->is_locked) {
                    if (e->carrying != NONE && e->obj[e->carrying].type == T_KEY && e->obj[e->carrying].color == o->color) {
                        o->is_locked = 0; o->is_open = 1;
                    }
                } else o->is_open = !o->is_open;
            } else if (o->type == T_BOX) cell_set(e, fx, fy, NONE);   /* contains = None */
        }
        break;
    default: break;   /* done */
    }
    if (e->step_count >= e->max_steps) dn = 1;
    if (image) gen_obs(e, image);
    /* RoomGridLevel.step */
    if (action == A_DROP) update_objs_poss(e, e->root);
    if (verify(e, e->root, action) == V_SUCCESS) {
        dn = 1;
        rew = 1 - 0.9 * ((double)e->step_count / (double)e->max_steps);   /* _reward() */
    }
    *reward = (float)rew;
    *done = (uint8_t)dn;
}

/* ================= exported C API (ctypes) ===================================== */
Pool *oracle_create(const LevelSpec *sp, int n)
{
    Pool *p = calloc(1, sizeof *p);
    p->n = n;
    p->env = calloc((size_t)n, sizeof(Env));
    for (int i = 0; i < n; i++) {
        Env *e = &p->env[i];
        e->sp = *sp;
        e->W = (sp->room_size - 1) * sp->num_cols + 1;
        e->H = (sp->room_size - 1) * sp->num_rows + 1;
        e->locked_room_idx = NONE;
        e->carrying = NONE;
    }
    return p;
}
void oracle_destroy(Pool *p) { free(p->env); free(p); }

void oracle_seed(Pool *p, const uint64_t *seeds)
{
    for (int i = 0; i < p->n; i++) { p->env[i].seed = seeds[i]; p->env[i].draws = 0; }
}

void oracle_reset(Pool *p, uint8_t *obs, int8_t *dir)
{
    for (int i = 0; i < p->n; i++) {
        env_reset(&p->env[i], obs ? obs + (size_t)i * OBS_BYTES : NULL);
        if (dir) dir[i] = (int8_t)p->env[i].agent_dir;
    }
}

/* penv.py:7-11 worker: step, and on done replace obs with reset()'s */
static void step_range(Pool *p, int lo, int hi, const int8_t *actions, uint8_t *obs, float *reward, uint8_t *done,
                       int8_t *dir, int autoreset)
{
    for (int i = lo; i < hi; i++) {
        Env *e = &p->env[i];
        uint8_t *im = obs ? obs + (size_t)i * OBS_BYTES : NULL;
        env_step(e, actions[i], im, &reward[i], &done[i]);
        if (done[i] && autoreset) env_reset(e, im);
        if (dir) dir[i] = (int8_t)e->agent_dir;
    }
}

void oracle_step(Pool *p, const int8_t *actions, uint8_t *obs, float *reward, uint8_t *done, int8_t *dir, int autoreset)
{
    step_range(p, 0, p->n, actions, obs, reward, done, dir, autoreset);
}

typedef struct { Pool *p; int lo, hi; const int8_t *a; uint8_t *obs; float *rew; uint8_t *done; int8_t *dir; int ar; } Job;
static void *job_main(void *arg)
{
    Job *j = arg;
    step_range(j->p, j->lo, j->hi, j->a, j->obs, j->rew, j->done, j->dir, j->ar);
    return NULL;
}
/* same, env range split over nthreads host threads (CPU baseline) */
void oracle_step_mt(Pool *p, const int8_t *actions, uint8_t *obs, float *reward, uint8_t *done, int8_t *dir,
                    int autoreset, int nthreads)
{
    if (nthreads <= 1) { oracle_step(p, actions, obs, reward, done, dir, autoreset); return; }
    pthread_t th[256]; Job jb[256];
    if (nthreads > 256) nthreads = 256;
    for (int t = 0; t < nthreads; t++) {
        jb[t] = (Job){ p, (int)((int64_t)p->n * t / nthreads), (int)((int64_t)p->n * (t + 1) / nthreads),
                       actions, obs, reward, done, dir, autoreset };
        pthread_create(&th[t], NULL, job_main, &jb[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
}

const char *oracle_mission(Pool *p, int i) { return p->env[i].mission; }

/* packed cell byte = type | color << 3 | state << 6  (empty = 0x01) */
void oracle_get_state(Pool *p, int i, uint8_t *grid /* H*W */, int32_t *info /* 8 */)
{
    Env *e = &p->env[i];
    for (int k = 0; k < e->W * e->H; k++) {
        uint8_t t[3]; encode_cell(e, e->cell[k], t);
        grid[k] = (uint8_t)(t[0] | (t[1] << 3) | (t[2] << 6));
    }
    uint8_t t[3] = { 0, 0, 0 };
    if (e->carrying != NONE) encode_cell(e, e->carrying, t);
    info[0] = e->agent_x; info[1] = e->agent_y; info[2] = e->agent_dir;
    info[3] = e->carrying == NONE ? 0 : (t[0] | (t[1] << 3) | (t[2] << 6));
    info[4] = e->step_count; info[5] = e->max_steps;
    info[6] = (int32_t)e->draws; info[7] = (int32_t)e->n_attempts;
}

int oracle_width(Pool *p) { return p->env[0].W; }
int oracle_height(Pool *p) { return p->env[0].H; }

/* Philox known-answer hook for tests */
void oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    philox(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}
