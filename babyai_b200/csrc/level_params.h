// level_params.h -- host-side derivation of the kernels' LevelParams from the
// reference-style constructor arguments (bb_level_spec).  Shared by pool.cu and
// the test-only host build (tests/hostemu) so both see the same layout.
#pragma once
#include <math.h>
#include <string.h>
#include "../../include/babyai_b200.h"
#include "env_logic.cuh"

namespace bb {

// returns nullptr on success, else a static error message
static inline const char *make_level_params(const bb_level_spec *s, LevelParams *lp)
{
    memset(lp, 0, sizeof *lp);
    if (s->kind < 0 || s->kind > 5) return "bad level kind";
    if (s->kind == BB_KIND_BONUS) { if (s->room_size < 3 || s->room_size > 20 || s->bonus < 1 || s->bonus > 20) return "bad bonus level"; }
    else if (s->room_size < 4 || s->room_size > 8) return "room_size must be in 4..8";
    if (s->num_rows < 1 || s->num_cols < 1 || s->num_rows * s->num_cols > MAXROOMS) return "too many rooms";
    lp->kind = s->kind; lp->room_size = s->room_size; lp->num_rows = s->num_rows; lp->num_cols = s->num_cols;
    lp->num_dists = s->num_dists; lp->instr = s->instr; lp->doors_open = s->doors_open; lp->grey_dists = s->grey_dists;
    lp->locations = s->locations; lp->unblocking = s->unblocking; lp->implicit_unlock = s->implicit_unlock;
    lp->all_unique = s->all_unique; lp->require_unreachable = s->require_unreachable;
    lp->strict_mask = s->strict_mask & 0x1F; lp->done_actions = s->done_actions ? 1 : 0;
    // the instruction leaves a family can produce, and whether it only ever builds ONE ActionInstr (the verifier's fast paths)
    lp->kinds_mask = 0xF; lp->single_instr = 1;
    if (s->kind == BB_KIND_REDBALL || s->kind == BB_KIND_IMPUNLOCK) lp->kinds_mask = 1 << BB_I_GOTO;
    else if (s->kind == BB_KIND_OBJ) lp->kinds_mask = 1 << s->instr;
    else if (s->kind == BB_KIND_UNLOCK) lp->kinds_mask = 1 << BB_I_OPEN;
    else if (s->kind == BB_KIND_LEVELGEN) {
        lp->kinds_mask = 0;
        for (int i = 0; i < s->n_action_kinds && i < 4; i++) lp->kinds_mask |= 1 << s->action_kinds[i];
        lp->single_instr = s->n_instr_kinds == 1 && s->instr_kinds[0] == BB_K_ACTION;
    } else if (s->kind == BB_KIND_BONUS) lp->single_instr = !(s->bonus == 14 || s->bonus == 19 || s->bonus == 20);   // OpenTwoDoors, MoveTwoAcross, OpenDoorsOrder
    lp->bonus = s->kind == BB_KIND_BONUS ? s->bonus : 0; lp->bonus_a = s->bonus_a; lp->bonus_b = s->bonus_b;
    lp->box_contains = lp->bonus == BN_KEY_IN_BOX ? 2 : 0;           // door = object 0, box = object 1, its key = object 2
    if (s->kind == BB_KIND_OBJ && (s->instr < BB_I_GOTO || s->instr > BB_I_PUTNEXT)) return "bad instruction kind";
    if (s->kind == BB_KIND_OBJ && s->instr == BB_I_PUTNEXT && s->num_dists < 2) return "PutNext needs two objects";
    lp->n_action_kinds = s->n_action_kinds; lp->n_instr_kinds = s->n_instr_kinds;
    for (int i = 0; i < 4; i++) lp->action_kinds[i] = s->action_kinds[i];
    for (int i = 0; i < 3; i++) lp->instr_kinds[i] = s->instr_kinds[i];
    lp->col_mul = s->num_cols > 0 ? (64 + s->num_cols - 1) / s->num_cols : 64;
    for (int r = 0; r < s->num_rows * s->num_cols; r++)
        if (s->num_cols <= 0 || ((r * lp->col_mul) >> 6) != r / s->num_cols) return "unsupported room grid (room / num_cols by multiplication)";
    lp->W = (s->room_size - 1) * s->num_cols + 1;
    lp->H = (s->room_size - 1) * s->num_rows + 1;
    if (lp->W > MAXH || lp->H > MAXH) return "grid too large";
    lp->cells = lp->W * lp->H;
    // the grid is stored twice: row-major G (H rows, stride rs_g) and column-major GT (W rows, stride rs_t);
    // strides are multiples of 4 so that a 7-cell window is three aligned 32-bit loads; padding bytes are walls
    lp->rs_g = (lp->W + 3) / 4 * 4;
    lp->rs_t = (lp->H + 3) / 4 * 4;
    lp->gt_off = (lp->H * lp->rs_g + 15) / 16 * 16;
    lp->cells_pad = lp->gt_off + (lp->W * lp->rs_t + 15) / 16 * 16;
    lp->nav_time_maze = s->room_size * s->room_size * s->num_rows * s->num_cols;   // levelgen.py:42-43
    int max_objs = s->num_dists + (s->kind == BB_KIND_OBJ ? 0 : 1);   // + the red ball / the key of the locked room
    if (s->kind == BB_KIND_IMPUNLOCK) {      // num_dists per unlocked room + the target in the locked room + the key
        if (s->num_rows * s->num_cols < 2) return "GoToImpUnlock needs at least two rooms";
        max_objs = s->num_dists * (s->num_rows * s->num_cols - 1) + 2;
    }
    if (s->kind == BB_KIND_UNLOCK) {         // only the doors are table objects; the key and the distractors are untracked
        if (s->num_rows * s->num_cols < 2) return "Unlock needs at least two rooms";
        if (s->num_dists * (s->num_rows * s->num_cols - 1) + 1 > MAXUNTRACKED) return "too many untracked objects";
        max_objs = 0;
    }
    if (s->kind == BB_KIND_LEVELGEN) {
        if (s->n_action_kinds < 1 || s->n_action_kinds > 4 || s->n_instr_kinds < 1 || s->n_instr_kinds > 3)
            return "bad LevelGen kinds";
        double t = ceil(s->locked_room_prob * 4294967296.0);
        lp->locked_thr = t <= 0 ? 0ull : (uint64_t)t;
    }
    // doors: one per internal wall at most
    max_objs += s->num_rows * (s->num_cols - 1) + s->num_cols * (s->num_rows - 1);
    if (s->kind == BB_KIND_BONUS) max_objs = MAXOBJ;               // (at most 18 objects + 4 doors, MoveTwoAcrossS8N9)
    if (max_objs > MAXOBJ) return "too many objects for the 32-entry object table";
    lp->obj_words = max_objs < 1 ? 1 : (max_objs + 3) / 4;
    // longest mission in tokens
    int per_desc = 3 + (s->kind == BB_KIND_LEVELGEN && s->locations ? 4 : 0);
    int leaf = 2 + per_desc;
    if (s->kind == BB_KIND_LEVELGEN) {
        bool putnext = false, has_and = false, has_seq = false;
        for (int i = 0; i < s->n_action_kinds; i++) if (s->action_kinds[i] == BB_I_PUTNEXT) putnext = true;
        for (int i = 0; i < s->n_instr_kinds; i++) { if (s->instr_kinds[i] == BB_K_AND) has_and = true; if (s->instr_kinds[i] == BB_K_SEQ) has_seq = true; }
        if (putnext) leaf = 1 + per_desc + 2 + per_desc;
        int side = (has_and || has_seq) ? 2 * leaf + 1 : leaf;
        lp->max_tokens = has_seq ? 2 * side + 2 : side;
    } else lp->max_tokens = (s->kind == BB_KIND_OBJ && s->instr == BB_I_PUTNEXT) ? 1 + per_desc + 2 + per_desc : leaf;
    if (s->kind == BB_KIND_BONUS) lp->max_tokens = 24;             // longest: MoveTwoAcross, two PutNext of 8 words + 'then'
    lp->max_tokens = (lp->max_tokens + 7) / 8 * 8;          // 16-byte rows
    if (lp->max_tokens > MAXTOK) lp->max_tokens = MAXTOK;
    // wall template of the empty RoomGrid (Grid.wall_rect per room)
    for (int y = 0; y < lp->H; y++) {
        uint32_t row = 0;
        for (int x = 0; x < lp->W; x++)
            if (x % (s->room_size - 1) == 0 || y % (s->room_size - 1) == 0) row |= 1u << x;
        lp->wall_rows[y] = row;
    }
    // small single-room levels: bitboard walls + byte rows of the empty room (generate_small / emit_small_level)
    lp->small = 0;
    const bool lg_ok = s->kind != BB_KIND_LEVELGEN ||
        (s->locked_room_prob <= 0 && s->n_instr_kinds == 1 && s->instr_kinds[0] == BB_K_ACTION && s->n_action_kinds == 1 &&
         (s->action_kinds[0] == BB_I_GOTO || s->action_kinds[0] == BB_I_PICKUP));
    const bool obj_ok = s->kind != BB_KIND_OBJ || ((s->instr == BB_I_GOTO || s->instr == BB_I_PICKUP) && !s->all_unique && !s->require_unreachable);
    if (s->kind <= BB_KIND_LEVELGEN && s->num_rows == 1 && s->num_cols == 1 && lp->W <= 8 && lp->H <= 8 && s->num_dists + 1 <= 10 && lg_ok && obj_ok) {
        lp->small = 1;
        lp->wall64 = 0;
        for (int y = 0; y < 8; y++)
            for (int x = 0; x < 8; x++)
                if (x >= lp->W || y >= lp->H || ((lp->wall_rows[y] >> x) & 1u)) lp->wall64 |= 1ull << (8 * y + x);
        for (int r = 0; r < 16; r++) {
            uint64_t row = 0;
            for (int c = 0; c < 8; c++) {
                const int x = r < 8 ? c : r - 8, y = r < 8 ? r : c;      // G row r = y; GT row r - 8 = x
                const bool wall = x >= lp->W || y >= lp->H || ((lp->wall_rows[y < lp->H ? y : 0] >> x) & 1u);
                row |= (uint64_t)(wall ? CELL_WALL : CELL_EMPTY) << (8 * c);
            }
            lp->row_tmpl[r] = row;
        }
    }
    return nullptr;
}

}  // namespace bb
