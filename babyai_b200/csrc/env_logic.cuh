// env_logic.cuh -- per-environment logic of the batched BabyAI pool: level
// generation, the 7 MiniGrid actions, the instruction verifier and the
// egocentric 7x7 observation.  One CUDA thread runs these for one environment;
// the kernels in pool.cu add the warp-level parts (coalesced staging of the
// observation bytes, ballot compaction of finished episodes).
//
// Everything here is BB_HD (__host__ __device__) so that tests/hostemu can
// compile the *same* source for the host and single-step it in a debugger
// against the oracle; the product only ever runs the device build.
//
// Representation (B200-first, not the reference's object graph):
//   * a grid cell is ONE byte  type | color<<3 | state<<6  (exactly the three
//     observation channels, so encoding a visible cell is bit-slicing);
//   * object identity -- which the reference verifier relies on through `is`
//     (verifier.py:120,266,340,408) -- lives in a 32-entry object table
//     (position + type/colour) and 32-bit sets: obj_set(desc) is a bitmask,
//     the `obj_poss` snapshot is the bitmask of objects that were on the grid
//     at the last refresh (positions of on-grid objects only change at a
//     successful drop, which is also a refresh, so no position copy is needed);
//   * level generation never reads the grid: occupancy is kept as one 32-bit
//     row mask per grid row (W <= 25), reachability is a bit-parallel flood
//     fill on those rows, descriptors are matched against the object table,
//     and the byte grid is rendered once at the end.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BB_HD __host__ __device__ __forceinline__
#define BB_HD_NOINLINE __host__ __device__ __noinline__
#define BB_ALIGN16 __align__(16)
#else
#define BB_ALIGN16 alignas(16)
#define BB_HD inline
#define BB_HD_NOINLINE
#endif

// Level generation for everything but the small single-room levels (generate_level) runs ONE WARP PER LEVEL on the device: the
// 32 lanes execute the same level with identical control flow, Philox blocks and grid rows split across them.  Compile
// with -DBB_GEN_COOP=0 for one LANE per level (the scalar code of the host build, working arrays in local memory): measured
// 3x SLOWER (r02f: GoTo 32 768 envs 8.5e8 vs 2.8e9 env-steps/s; 4 working lanes per warp beat 8, 16 and 32 -- the lanes of
// a warp sit in different rejection loops, and the scalar working set lives in local memory).
#ifndef BB_GEN_COOP
#define BB_GEN_COOP 1
#endif
#if defined(__CUDA_ARCH__) && BB_GEN_COOP
#define BB_GEN_WARP 1
#else
#define BB_GEN_WARP 0
#endif
// The cooperative generator is bound by INSTRUCTION FETCH (k_gen: 41 700 instructions = 667 KB against a 32 KB L1.5
// instruction cache; `no_instruction` the top stall, ncu r02c): the helpers that the grammar inlines at dozens of call sites
// are real functions there (BB_GEN_OUTLINE=0: everything inlined, as in round 1).
#ifndef BB_GEN_OUTLINE
#define BB_GEN_OUTLINE 1
#endif
// the generator's working memory is the warp's slice of the kernel's __shared__ array, but it reaches the out-of-line functions
// as a generic pointer: without the hint every access is a generic LD.E / ST.E with 64-bit address arithmetic (three more
// instructions than an LDS / STS, and a longer latency)
#if defined(__CUDA_ARCH__) && BB_GEN_COOP
#define BB_ASSUME_SHARED(p) __builtin_assume(__isShared(p))
#else
#define BB_ASSUME_SHARED(p) ((void)0)
#endif
#if defined(__CUDACC__) && BB_GEN_COOP && BB_GEN_OUTLINE
#define BB_GEN_FN __host__ __device__ __noinline__
#else
#define BB_GEN_FN BB_HD
#endif

namespace bb {

// ---- constants (gym_minigrid OBJECT_TO_IDX / COLOR_TO_IDX / STATE_TO_IDX) ----
enum : int { T_UNSEEN = 0, T_EMPTY = 1, T_WALL = 2, T_DOOR = 4, T_KEY = 5, T_BALL = 6, T_BOX = 7 };
enum : int { C_RED = 0, C_GREEN = 1, C_BLUE = 2, C_PURPLE = 3, C_YELLOW = 4, C_GREY = 5 };
enum : int { A_LEFT = 0, A_RIGHT = 1, A_FORWARD = 2, A_PICKUP = 3, A_DROP = 4, A_TOGGLE = 5, A_DONE = 6 };
enum : int { KIND_REDBALL = 0, KIND_OBJ = 1, KIND_LEVELGEN = 2, KIND_IMPUNLOCK = 3, KIND_UNLOCK = 4, KIND_BONUS = 5 };
// KIND_BONUS: the gen_mission of babyai/levels/bonus_levels.py, selected by LevelParams::bonus (bonus_a / bonus_b: its arguments)
enum : int { BN_GOTO_REDBLUE = 1, BN_OPEN_RED_DOOR, BN_OPEN_DOOR, BN_GOTO_DOOR, BN_GOTO_OBJ_DOOR, BN_ACTION_OBJ_DOOR, BN_UNLOCK_LOCAL,
             BN_KEY_IN_BOX, BN_UNLOCK_PICKUP, BN_BLOCKED_UNLOCK_PICKUP, BN_UNLOCK_TO_UNLOCK, BN_PICKUP_DIST, BN_PICKUP_ABOVE,
             BN_OPEN_TWO_DOORS, BN_FIND_OBJ, BN_KEY_CORRIDOR, BN_ONE_ROOM, BN_PUTNEXT, BN_MOVE_TWO_ACROSS, BN_OPEN_DOORS_ORDER };
enum : int { I_GOTO = 0, I_PICKUP = 1, I_OPEN = 2, I_PUTNEXT = 3, I_NONE = 0xFF };
enum : int { K_ACTION = 0, K_AND = 1, K_SEQ = 2 };
enum : int { R_SINGLE = 0, R_BEFORE = 1, R_AFTER = 2 };
enum : int { LOC_LEFT = 0, LOC_RIGHT = 1, LOC_FRONT = 2, LOC_BEHIND = 3, LOC_NONE = 7 };
enum : int { ANY = 7, ANY_TYPE = 0 };   // "None" for descriptor colour / type (7 is T_BOX)

constexpr int CELL_EMPTY = T_EMPTY;                       // 0x01
constexpr int CELL_WALL = T_WALL | (C_GREY << 3);         // 0x2A
constexpr int NO_OBJ = 0xFF;
// Level_Unlock places 24 distractors + a key + up to 12 doors: more than the 32-entry object table.  Its instruction
// (open a door) only ever names doors, so only the doors are TRACKED (table entry, set-mask bit); every other object of
// that level is UNTRACKED: it exists as its cell byte alone, and while carried as CARRY_UNTRACKED | type | color << 3 in
// EnvHot.carry.  Only the kernels instantiated with UNTR = true (KIND_UNLOCK pools) know this encoding.
constexpr int CARRY_UNTRACKED = 0x40;
constexpr int MAXUNTRACKED = 32;
constexpr int MAXOBJ = 32;
constexpr int MAXH = 25;
constexpr int MAXROOMS = 16;
constexpr int MAXTOK = 72;
constexpr int OBS_BYTES = 147;
constexpr int OBS_WORDS = 37;

// vocabulary of the baby language (verifier.py surface() strings); id 0 = pad.
enum : int {
    W_PAD = 0, W_GO, W_TO, W_PICK, W_UP, W_OPEN, W_PUT, W_NEXT, W_THE, W_A, W_OBJECT,
    W_RED, W_GREEN, W_BLUE, W_PURPLE, W_YELLOW, W_GREY,       // W_RED + COLOR_TO_IDX
    W_BOX, W_BALL, W_KEY, W_DOOR,
    W_IN, W_FRONT, W_OF, W_YOU, W_BEHIND, W_ON, W_YOUR, W_LEFT, W_RIGHT,
    W_THEN, W_AFTER, W_AND, W_COUNT
};

// ---- level parameters (kernel argument, lives in the constant bank) ---------
struct LevelParams {
    int32_t kind, room_size, num_rows, num_cols, num_dists, instr, doors_open, grey_dists;
    int32_t locations, unblocking, implicit_unlock;
    int32_t all_unique, require_unreachable;
    int32_t n_action_kinds, action_kinds[4];
    int32_t n_instr_kinds, instr_kinds[3];
    int32_t W, H, cells, cells_pad, max_tokens, nav_time_maze;
    int32_t strict_mask, done_actions;   // verifier modes (see verify_action / verify_leaf)
    int32_t col_mul;                     // room / num_cols == (room * col_mul) >> 6 for every room of the level (make_level_params checks it)
    int32_t kinds_mask, single_instr;    // bit k: the family can produce an instruction leaf of kind k; 1: it only produces a single ActionInstr
    int32_t bonus, bonus_a, bonus_b;     // KIND_BONUS: which bonus_levels.py family and its constructor arguments
    int32_t box_contains;                // box object id + 1 whose contents is the NEXT table entry (Level_KeyInBox), else 0
    int32_t obj_words;            // ceil(most object-table entries a level of this family uses / 4): words of the packed x / y arrays in use
    int32_t rs_g, rs_t, gt_off;   // grid bytes of one env: G = H rows x rs_g at 0, GT = W rows x rs_t at gt_off
    uint64_t locked_thr;          // rand_float(0,1) < p  <=>  u32 < ceil(p * 2^32)
    uint32_t wall_rows[MAXH];     // bit x of row y: (x, y) is a wall of the empty RoomGrid
    // single-room levels up to 8x8 ("small"): bitboard (bit 8 y + x) of walls and of everything outside the grid,
    // and the byte rows of the empty room for both stored orientations (G rows 0..7, GT rows 8..15)
    int32_t small;
    uint64_t wall64;
    uint64_t row_tmpl[16];
};

// ---- per-environment records (struct-of-arrays over envs, one array each) ---
struct BB_ALIGN16 EnvHot {
    uint8_t x, y;
    uint8_t dirflags;             // bits 0-1 agent_dir, bit 2 frozen (ManyEnvs flavour)
    uint8_t carry;                // object id, NO_OBJ, or (KIND_UNLOCK) CARRY_UNTRACKED | type | color << 3
    uint16_t step_count, max_steps;
    uint32_t cur_mask;            // objects currently on the grid
    uint32_t snap_mask;           // objects on the grid at the last obj_poss refresh
};
struct BB_ALIGN16 ObjTab {
    uint8_t x[MAXOBJ], y[MAXOBJ]; // last on-grid position (== WorldObj.cur_pos while not carried)
    uint8_t tc[MAXOBJ];           // type | color << 3
};
struct BB_ALIGN16 InstrRec {
    uint32_t desc_mask[8];        // obj_set of descriptor d; leaf i owns descs 2i (and 2i+1: PutNext fixed)
    uint8_t leaf_kind[4];         // leaves 0,1 = side A; 2,3 = side B
    uint8_t leaf_pre[4];          // preCarrying (verifier.py:325,373)
    uint8_t root_kind;            // R_SINGLE / R_BEFORE / R_AFTER
    uint8_t side_and;             // bit 0: side A is an AndInstr, bit 1: side B; bits 4-7: lastStepMatch of leaf 0-3 (done-action mode)
    uint8_t flags;                // 'success' latches: 0 root.a 1 root.b 2 A.a 3 A.b 4 B.a 5 B.b
    uint8_t pad0;
    uint32_t pad1;
};
struct BB_ALIGN16 RngRec { uint64_t seed, draws; };

// ---- Philox4x32-10; stream layout documented in DESIGN.md --------------------
BB_HD uint32_t mulhi32(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

// Per-env random stream, scalar form: one lane owns the stream (generate_small, host build).
struct RngScalar {
    uint32_t k0, k1;
    uint64_t draws;
    uint64_t blk;                 // block held in b0..b3; ~0 = none
    uint32_t b0, b1, b2, b3;

    BB_HD void init(uint64_t seed, uint64_t d)
    {
        k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32); draws = d; blk = ~0ull;
        b0 = b1 = b2 = b3 = 0;
    }
    BB_HD void refill(uint64_t n)
    {
        uint32_t c0 = (uint32_t)n, c1 = (uint32_t)(n >> 32), c2 = 0, c3 = 0, x0 = k0, x1 = k1;
#pragma unroll
        for (int r = 0; r < 10; r++) {
            uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            uint32_t n0 = hi1 ^ c1 ^ x0, n2 = hi0 ^ c3 ^ x1;
            c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
            x0 += 0x9E3779B9u; x1 += 0xBB67AE85u;
        }
        b0 = c0; b1 = c1; b2 = c2; b3 = c3;
    }
    BB_HD uint32_t u32()
    {
        const uint64_t i = draws++;
        const uint64_t nb = i >> 2;
        const uint32_t w = (uint32_t)i & 3u;
        if (nb != blk) { blk = nb; refill(nb); }
        return w == 0 ? b0 : w == 1 ? b1 : w == 2 ? b2 : b3;
    }
    // MiniGridEnv._rand_int(lo, hi); a range of one value consumes no draw
    BB_HD int randint(int lo, int hi)
    {
        uint32_t n = (uint32_t)(hi - lo);
        if (n == 1) return lo;
        return lo + (int)mulhi32(u32(), n);
    }
    BB_HD bool randbool() { return randint(0, 2) == 0; }
};

// Warp-cooperative form (generate_level on the device): the WHOLE WARP runs one environment with identical
// control flow, the 32 lanes compute 32 consecutive Philox blocks (128 draws) at once and a draw is a shuffle
// from the lane that holds its block.  In the host build it is the scalar generator.
#if BB_GEN_WARP
// One Philox block, OUT OF LINE: the cooperative generator draws at ~600 call sites and needs a new block at one draw in
// 128; inlined, the ten rounds were 45 % of k_gen's 41 700 instructions (667 KB of code against a 32 KB L1.5 instruction
// cache: `no_instruction` was the top stall of the kernel, ncu r02c).
static __device__ __noinline__ uint4 philox_block_ool(uint32_t k0, uint32_t k1, uint64_t n)
{
    uint32_t c0 = (uint32_t)n, c1 = (uint32_t)(n >> 32), c2 = 0, c3 = 0, x0 = k0, x1 = k1;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ x0, n2 = hi0 ^ c3 ^ x1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        x0 += 0x9E3779B9u; x1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
#endif
#if BB_GEN_WARP
// The warp's random stream: 128 consecutive draws (32 Philox blocks, one per lane) sit in the warp's shared GenMem; a draw is
// one broadcast shared-memory load.  (Round 1 kept the four words of a lane's block in registers and shuffled: a 64-bit
// counter, three selects and a shuffle per draw, ~25 instructions at each of ~600 call sites.)
// the cold path of a draw, one copy, arguments by value (the Rng itself stays in registers)
static __device__ __noinline__ void rng_fill_ool(uint32_t k0, uint32_t k1, uint64_t base, uint32_t sbuf)
{
    __syncwarp();
    const uint4 b = philox_block_ool(k0, k1, (base >> 2) + (threadIdx.x & 31));
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sbuf + 16u * (threadIdx.x & 31)), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
    __syncwarp();
}
struct Rng {
    uint32_t k0, k1;
    uint64_t base;                // draw index of buf[0], a multiple of 4
    uint32_t pos;                 // draws consumed from buf; 128 = used up
    uint32_t sbuf;                // GenMem::draw_buf as a SHARED-space address: the draw is an explicit ld.shared (a generic
                                  // pointer that has passed through an out-of-line call loses its address space: LD.E + 64-bit
                                  // address arithmetic at the hottest 13 M loads of a pass, ncu r02w)
    __device__ __forceinline__ void fill() { rng_fill_ool(k0, k1, base, sbuf); }
    __device__ __forceinline__ void init(uint64_t seed, uint64_t d, uint32_t *b)
    {
        k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32); sbuf = (uint32_t)__cvta_generic_to_shared(b);
        base = d & ~3ull; pos = (uint32_t)d & 3u;
        fill();
    }
    __device__ __forceinline__ uint64_t count() const { return base + pos; }
    __device__ __forceinline__ uint32_t u32()
    {
        if (__builtin_expect(pos >= 128u, 0)) { base += 128ull; pos = 0; fill(); }      // warp-uniform, one draw in 128
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(sbuf + 4u * pos));
        pos++;
        return v;
    }
    __device__ __forceinline__ int randint(int lo, int hi)
    {
        uint32_t n = (uint32_t)(hi - lo);
        if (n == 1) return lo;
        return lo + (int)__umulhi(u32(), n);
    }
    __device__ __forceinline__ bool randbool() { return randint(0, 2) == 0; }
};
#else
struct Rng : RngScalar {
    BB_HD void init(uint64_t seed, uint64_t d, uint32_t *) { RngScalar::init(seed, d); }
    BB_HD uint64_t count() const { return draws; }
};
#endif

BB_HD int popc32(uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
BB_HD int ffs32(uint32_t v)   // index of lowest set bit, v != 0
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}
BB_HD int iabs(int v) { return v < 0 ? -v : v; }

// DIR_TO_VEC: right, down, left, up
BB_HD int dir_dx(int d) { return d == 0 ? 1 : d == 2 ? -1 : 0; }
BB_HD int dir_dy(int d) { return d == 1 ? 1 : d == 3 ? -1 : 0; }

// COLOR_NAMES = sorted(['red','green','blue','purple','yellow','grey'])
BB_HD int color_by_name_rank(int k)   // blue green grey purple red yellow
{
    // one nibble per rank: C_BLUE, C_GREEN, C_GREY, C_PURPLE, C_RED, C_YELLOW
    return (int)((((uint32_t)C_BLUE) | ((uint32_t)C_GREEN << 4) | ((uint32_t)C_GREY << 8) | ((uint32_t)C_PURPLE << 12) |
                  ((uint32_t)C_RED << 16) | ((uint32_t)C_YELLOW << 20)) >> (4 * k)) & 7;
}

// =============================================================================
// Level generation (RoomGridLevel._gen_grid, levelgen.py:77-102, and below)
// =============================================================================
// the two orientations of one env's grid (see LevelParams::rs_g / rs_t / gt_off)
BB_HD int get_cell(const LevelParams &lp, const uint8_t *grid, int x, int y) { return grid[y * lp.rs_g + x]; }
BB_HD void set_cell(const LevelParams &lp, uint8_t *grid, int x, int y, int v)
{
    grid[y * lp.rs_g + x] = (uint8_t)v;
    grid[lp.gt_off + x * lp.rs_t + y] = (uint8_t)v;
}

struct LevelOut {           // where one generated level is written (live or spare slot)
    uint8_t *grid; EnvHot *hot; ObjTab *obj; InstrRec *ins; int16_t *tok;
};

// Working arrays of one generation.  Device: ONE copy per warp in shared memory (all 32 lanes run the
// same environment in lockstep and would otherwise each keep a private copy in local memory: 40 KB of
// L1 per warp, which round 1 showed evicts the step kernel's working set).  Host build: on the stack.
struct GenMem {
    uint32_t occ[MAXH];            // walls + doors + objects, one bit per cell
    uint32_t doorcell[MAXH];       // door cells (subset of occ)
    uint32_t wallmask[MAXH];       // the walls of THIS level: lp.wall_rows minus what remove_wall() took out (bonus levels)
    uint32_t pass[MAXH], fill[MAXH];  // reachability flood fill
    uint8_t door_y_right[MAXROOMS];   // Room.door_pos[0].y of room r
    uint8_t door_x_down[MAXROOMS];    // Room.door_pos[1].x of room r
    uint8_t door_id_right[MAXROOMS];  // object id of the door in room r's right / down slot
    uint8_t door_id_down[MAXROOMS];
    // instruction being built
    int leaf_kind[4]; int desc_type[8], desc_color[8], desc_loc[8];
    uint32_t desc_mask[8];
    int16_t tok[MAXTOK];
    ObjTab obj;                    // object table under construction (copied to the slot at the end)
#if defined(__CUDACC__) && BB_GEN_COOP
    BB_ALIGN16 uint32_t draw_buf[128];   // the warp's next 128 draws (struct Rng)
#endif
};

struct GenMemX : GenMem {          // KIND_UNLOCK only: the untracked objects of the level under construction
    uint8_t ux[MAXUNTRACKED], uy[MAXUNTRACKED], utc[MAXUNTRACKED];
};

template <bool X> struct GenMemFor { typedef GenMem type; };
template <> struct GenMemFor<true> { typedef GenMemX type; };

struct GenCtx {
    Rng rng;
    GenMem *m;
    int nun;                          // untracked objects placed so far (KIND_UNLOCK)
    uint32_t door_right, door_down;   // bit r: a door exists in room r's right / down slot
    uint32_t room_locked;             // bit r: Room.locked
    int nobj;
    int ax, ay, adir; bool agent_placed;
    int locked_door;                  // object id of the locked door or -1
    uint32_t locked_mask;             // every locked door (bonus levels have up to two)
    uint32_t hidden_mask;             // objects that are not on the grid at reset (the key inside KeyInBox's box)
    int start_carry;                  // object the agent picks up before its first step (PutNext*Carrying), or NO_OBJ
    int root_kind, side_and;
    // LevelGen.locked_room (persists across episodes, levelgen.py:284)
    int locked_room; bool locked_room_fresh;
};

enum : int { GEN_OK = 0, GEN_REJECT = 1, GEN_RECURSION = 2 };
#define BB_TRY(x) do { int _r = (x); if (_r) return _r; } while (0)

// room index -> (column, row) without an integer division (a dozen instructions each on the device; g_place alone was 6.6 % of
// k_gen's instructions, ncu r02u)
BB_HD int room_row(const LevelParams &lp, int room) { return (room * lp.col_mul) >> 6; }
BB_HD int room_col(const LevelParams &lp, int room) { return room - room_row(lp, room) * lp.num_cols; }

// RoomGrid._gen_grid (gym_minigrid.roomgrid; SURVEY App. A.6 / App. B "G0")
BB_HD void g_roomgrid(const LevelParams &lp, GenCtx &g)
{
    const int S = lp.room_size, R = lp.num_rows, C = lp.num_cols;
#if BB_GEN_WARP
    // the lanes split the rows and the object-table words (everything else in the generator is the same store from every lane)
    const int l0 = threadIdx.x & 31, lstep = 32;
    __syncwarp();
#else
    const int l0 = 0, lstep = 1;
#endif
    for (int y = l0; y < lp.H; y += lstep) { g.m->occ[y] = lp.wall_rows[y]; g.m->doorcell[y] = 0; g.m->wallmask[y] = lp.wall_rows[y]; }
    for (int j = 0; j < R; j++)
        for (int i = 0; i < C; i++) {
            int r = j * C + i, tx = i * (S - 1), ty = j * (S - 1);
            if (i < C - 1) g.m->door_y_right[r] = (uint8_t)g.rng.randint(ty + 1, ty + S - 1);
            if (j < R - 1) g.m->door_x_down[r] = (uint8_t)g.rng.randint(tx + 1, tx + S - 1);
        }
    g.door_right = g.door_down = g.room_locked = 0;
    {   // unused object-table entries are zero (deterministic state bytes)
        uint32_t *ow = reinterpret_cast<uint32_t *>(&g.m->obj);
        for (int k = l0; k < (int)(sizeof(ObjTab) / 4); k += lstep) ow[k] = 0;
    }
#if BB_GEN_WARP
    __syncwarp();
#endif
    g.nobj = 0; g.nun = 0;
    g.ax = (C / 2) * (S - 1) + S / 2;
    g.ay = (R / 2) * (S - 1) + S / 2;
    g.adir = 0; g.agent_placed = true;
    g.locked_door = -1; g.locked_mask = 0; g.hidden_mask = 0; g.start_carry = NO_OBJ;
    g.locked_room_fresh = false;
}

// MiniGridEnv.place_obj over one room rectangle (walls included), App. A.3
BB_HD int g_place(const LevelParams &lp, GenCtx &g, int room, bool reject_next_to, int &ox, int &oy)
{
    const int S = lp.room_size;
    const int tx = room_col(lp, room) * (S - 1), ty = room_row(lp, room) * (S - 1);
    const int hx = tx + S < lp.W ? tx + S : lp.W, hy = ty + S < lp.H ? ty + S : lp.H;
    int tries = 0;
    for (;;) {
        if (tries > 1000) return GEN_RECURSION;
        tries++;
        int x = g.rng.randint(tx, hx);
        int y = g.rng.randint(ty, hy);
        if ((g.m->occ[y] >> x) & 1u) continue;
        if (g.agent_placed && x == g.ax && y == g.ay) continue;
        if (reject_next_to && iabs(g.ax - x) + iabs(g.ay - y) < 2) continue;
        ox = x; oy = y;
        return GEN_OK;
    }
}

// RoomGrid.add_object -> place_in_room
BB_HD int g_add_object(const LevelParams &lp, GenCtx &g, const LevelOut &o, int room, int type, int color, int &id, bool untracked = false)
{
    int x, y;
    if (untracked) {                    // KIND_UNLOCK: no table entry, the object is its cell byte (rendered by generate_level_t<true>)
        GenMemX *mx = static_cast<GenMemX *>(g.m);
        id = NO_OBJ;
        BB_TRY(g_place(lp, g, room, true, x, y));
        if (g.nun >= MAXUNTRACKED) return GEN_RECURSION;     // cannot happen: make_level_params bounds the count
        mx->ux[g.nun] = (uint8_t)x; mx->uy[g.nun] = (uint8_t)y; mx->utc[g.nun] = (uint8_t)(type | (color << 3));
        g.nun++;
        g.m->occ[y] |= 1u << x;
        return GEN_OK;
    }
    id = g.nobj++;                      // the Python object exists even if placement then fails
    BB_TRY(g_place(lp, g, room, true, x, y));
    g.m->obj.x[id] = (uint8_t)x; g.m->obj.y[id] = (uint8_t)y; g.m->obj.tc[id] = (uint8_t)(type | (color << 3));
    g.m->occ[y] |= 1u << x;
    return GEN_OK;
}

BB_HD bool g_has_slot(const LevelParams &lp, int room, int k)
{
    int i = room_col(lp, room), j = room_row(lp, room);
    return k == 0 ? i < lp.num_cols - 1 : k == 1 ? j < lp.num_rows - 1 : k == 2 ? i > 0 : j > 0;
}
BB_HD int g_neighbor(const LevelParams &lp, int room, int k)
{
    return k == 0 ? room + 1 : k == 1 ? room + lp.num_cols : k == 2 ? room - 1 : room - lp.num_cols;
}
BB_HD bool g_has_door(const LevelParams &lp, const GenCtx &g, int room, int k)
{
    if (k == 0) return (g.door_right >> room) & 1u;
    if (k == 1) return (g.door_down >> room) & 1u;
    if (k == 2) return (g.door_right >> (room - 1)) & 1u;
    return (g.door_down >> (room - lp.num_cols)) & 1u;
}

// RoomGrid.add_door(i, j, door_idx, color, locked) with everything decided
BB_HD int g_add_door(const LevelParams &lp, GenCtx &g, const LevelOut &o, int room, int k, int color, bool locked)
{
    const int S = lp.room_size;
    int owner = k == 2 ? room - 1 : k == 3 ? room - lp.num_cols : room;
    int x, y;
    if (k == 0 || k == 2) { x = room_col(lp, owner) * (S - 1) + S - 1; y = g.m->door_y_right[owner]; g.door_right |= 1u << owner; }
    else { x = g.m->door_x_down[owner]; y = room_row(lp, owner) * (S - 1) + S - 1; g.door_down |= 1u << owner; }
    if (locked) g.room_locked |= 1u << room; else g.room_locked &= ~(1u << room);   // room.locked = locked
    int id = g.nobj++;
    if (k == 0 || k == 2) g.m->door_id_right[owner] = (uint8_t)id; else g.m->door_id_down[owner] = (uint8_t)id;
    g.m->obj.x[id] = (uint8_t)x; g.m->obj.y[id] = (uint8_t)y; g.m->obj.tc[id] = (uint8_t)(T_DOOR | (color << 3));
    g.m->doorcell[y] |= 1u << x;
    if (locked) { g.locked_door = id; g.locked_mask |= 1u << id; }
    return id;
}

// RoomGrid.place_agent(i=None, j=None, rand_dir=True)
BB_HD int g_place_agent(const LevelParams &lp, GenCtx &g, int room_given = -1)
{
    int room = room_given;
    if (room < 0) {
        int i = g.rng.randint(0, lp.num_cols);
        int j = g.rng.randint(0, lp.num_rows);
        room = j * lp.num_cols + i;
    }
    {   // KNOWN DIVERGENCE: the reference's loop below never returns when no empty cell of the
        // room has an empty or wall cell in front of it for any heading (3x3 rooms packed with distractors and
        // doors: MiniBossLevel seed 698, 57th level).  The reference hangs; here the level is rejected like any
        // other failed rejection sampling and generation starts over.  Levels the reference CAN generate are unaffected.
        const int S = lp.room_size;
        const int tx = room_col(lp, room) * (S - 1), ty = room_row(lp, room) * (S - 1);
        const uint32_t rmask = ((1u << (S - 2)) - 1u) << (tx + 1);
        bool any = false;
        for (int y = ty + 1; y < ty + S - 1; y++) {
            const uint32_t empty = ~g.m->occ[y] & rmask;
            // cells that may be in front of the agent: empty, or wall (a door is neither)
            const uint32_t f0 = ~g.m->occ[y] | (g.m->wallmask[y] & ~g.m->doorcell[y]);
            const uint32_t fu = ~g.m->occ[y - 1] | (g.m->wallmask[y - 1] & ~g.m->doorcell[y - 1]);
            const uint32_t fd = ~g.m->occ[y + 1] | (g.m->wallmask[y + 1] & ~g.m->doorcell[y + 1]);
            if (empty & ((f0 << 1) | (f0 >> 1) | fu | fd)) any = true;
        }
        if (!any) return GEN_RECURSION;
    }
    for (;;) {
        int x, y;
        g.agent_placed = false;          // MiniGridEnv.place_agent: agent_pos = None while sampling
        BB_TRY(g_place(lp, g, room, false, x, y));
        g.ax = x; g.ay = y; g.agent_placed = true;
        g.adir = g.rng.randint(0, 4);
        int fx = x + dir_dx(g.adir), fy = y + dir_dy(g.adir);
        // front cell must be empty or a wall (a door is neither)
        bool occupied = (g.m->occ[fy] >> fx) & 1u;
        bool wall = ((g.m->wallmask[fy] >> fx) & 1u) && !((g.m->doorcell[fy] >> fx) & 1u);
        if (!occupied || wall) break;
    }
    return GEN_OK;
}

// RoomGrid.connect_all(door_colors=COLOR_NAMES, max_itrs=5000)
// exclude >= 0: door_colors = COLOR_NAMES without that colour (Level_Unlock, iclr19_levels.py:441-446)
BB_HD int g_connect_all(const LevelParams &lp, GenCtx &g, const LevelOut &o, int exclude = -1)
{
    const int S = lp.room_size, C = lp.num_cols, NR = lp.num_rows * lp.num_cols;
    const int start = (g.ay / (S - 1)) * C + g.ax / (S - 1);
    const uint32_t all = NR >= 32 ? 0xFFFFFFFFu : ((1u << NR) - 1u);
    int itrs = 0;
    // find_reach only changes when a door is added: the reference recomputes it on every trip of the loop (every random
    // (i, j, k) draw, most of which hit a wall slot or an existing door); here it is recomputed after a door was added only
    // (ncu r02a: the fixpoint below was 52 % of k_gen's instructions on BossLevel).  The draws are the same.
    uint32_t reach = 0;
    bool stale = true;
    for (;;) {
        if (itrs > 5000) return GEN_RECURSION;
        itrs++;
        if (stale) {
            // find_reach as a flood over room bitmasks: door_right bit r joins rooms r and r + 1 (set only where the slot
            // exists), door_down bit r joins r and r + C
            reach = 1u << start;
            for (;;) {
                const uint32_t nr = reach | ((reach & g.door_right) << 1) | ((reach >> 1) & g.door_right) |
                                    ((reach & g.door_down) << C) | ((reach >> C) & g.door_down);
                if (nr == reach) break;
                reach = nr;
            }
            stale = false;
        }
        if (reach == all) break;
        int i = g.rng.randint(0, C);
        int j = g.rng.randint(0, lp.num_rows);
        int k = g.rng.randint(0, 4);
        int room = j * C + i;
        // (the slot test on the drawn (i, j): g_has_slot(room) would divide to get them back)
        const bool slot = k == 0 ? i < C - 1 : k == 1 ? j < lp.num_rows - 1 : k == 2 ? i > 0 : j > 0;
        if (!slot || g_has_door(lp, g, room, k)) continue;
        if (((g.room_locked >> room) & 1u) || ((g.room_locked >> g_neighbor(lp, room, k)) & 1u)) continue;
        int color;
        if (exclude < 0) color = color_by_name_rank(g.rng.randint(0, 6));
        else {                               // _rand_elem over the five remaining names, in name order
            const int pick = g.rng.randint(0, 5);
            int rank = 0;
            for (int c = 0, n = 0; c < 6; c++) if (color_by_name_rank(c) != exclude && n++ == pick) rank = c;
            color = color_by_name_rank(rank);
        }
        g_add_door(lp, g, o, room, k, color, false);
        stale = true;
    }
    return GEN_OK;
}

// RoomGrid.add_distractors(i=None, j=None, num, all_unique): with all_unique a (type, color) pair that was
// already drawn is drawn again before any room / position draw
BB_HD int g_add_distractors(const LevelParams &lp, GenCtx &g, const LevelOut &o, int num, int &first_id, bool all_unique = false, int room = -1, bool untracked = false)
{
    first_id = g.nobj;
    uint32_t seen = 0;                       // bit 6 * type_rank + color
    if (all_unique)                          // everything already placed through place_in_room counts (roomgrid.add_distractors)
        for (int k = 0; k < g.nobj; k++) {
            const int tc = g.m->obj.tc[k], ty = tc & 7;
            if (ty >= T_KEY && !((g.hidden_mask >> k) & 1u)) seen |= 1u << (6 * (ty == T_KEY ? 0 : ty == T_BALL ? 1 : 2) + (tc >> 3));
        }
    for (int n = 0; n < num;) {
        int color = color_by_name_rank(g.rng.randint(0, 6));
        int t = g.rng.randint(0, 3);
        int type = t == 0 ? T_KEY : t == 1 ? T_BALL : T_BOX;
        if (all_unique) {
            const uint32_t bit = 1u << (6 * t + color);
            if (seen & bit) continue;
            seen |= bit;
        }
        int r = room;                        // add_distractors(i, j, ...): no room draws when the room is given
        if (r < 0) {
            int ri = g.rng.randint(0, lp.num_cols);
            int rj = g.rng.randint(0, lp.num_rows);
            r = rj * lp.num_cols + ri;
        }
        int id;
        BB_TRY(g_add_object(lp, g, o, r, type, color, id, untracked));
        n++;
    }
    return GEN_OK;
}

// RoomGridLevel.check_objs_reachable (levelgen.py:201-253) as a bit-parallel
// flood fill: F grows through cells that are empty or doors; a cell counts as
// reachable if it is in F or 4-adjacent to F; every object and door must be.
BB_HD int g_check_reachable(const LevelParams &lp, GenCtx &g)
{
    const uint32_t full = (lp.W >= 32) ? 0xFFFFFFFFu : ((1u << lp.W) - 1u);
#if BB_GEN_WARP
    // warp-per-level form (generate_level runs with the whole warp on one level): lane y owns grid row y, the rows above and
    // below come by shuffle, one iteration spreads the fill by one row and up to four columns in every row at once
    const int lane = threadIdx.x & 31;
    uint32_t pass = 0, things = 0, f = 0;
    if (lane < lp.H) {
        const uint32_t occ = g.m->occ[lane], door = g.m->doorcell[lane];
        pass = (~occ | door) & full;
        things = (occ & ~g.m->wallmask[lane]) | door;              // non-wall, non-empty cells
        if (lane == g.ay) f = 1u << g.ax;
    }
    uint32_t up, dn;
    for (;;) {
        up = __shfl_up_sync(0xFFFFFFFFu, f, 1); if (lane == 0) up = 0;
        dn = __shfl_down_sync(0xFFFFFFFFu, f, 1); if (lane == 31) dn = 0;
        uint32_t n = (f | (f << 1) | (f >> 1) | up | dn) & pass;
        n |= ((n << 1) | (n >> 1)) & pass;
        n |= ((n << 1) | (n >> 1)) & pass;
        n |= ((n << 1) | (n >> 1)) & pass;
        const bool changed = n != f;
        f = n;
        if (!__any_sync(0xFFFFFFFFu, changed)) break;
    }
    const uint32_t near = f | (f << 1) | (f >> 1) | up | dn;
    return __any_sync(0xFFFFFFFFu, (things & ~near) != 0u) ? GEN_REJECT : GEN_OK;
#else
    uint32_t *pass = g.m->pass, *f = g.m->fill;
    for (int y = 0; y < lp.H; y++) { pass[y] = (~g.m->occ[y] | g.m->doorcell[y]) & full; f[y] = 0; }
    f[g.ay] = 1u << g.ax;
    for (;;) {
        bool changed = false;
        for (int y = 0; y < lp.H; y++) {
            uint32_t v = f[y];
            uint32_t n = v | (v << 1) | (v >> 1);
            if (y > 0) n |= f[y - 1];
            if (y + 1 < lp.H) n |= f[y + 1];
            n &= pass[y];
            // run the horizontal spread to its fixpoint inside the row
            for (;;) { uint32_t m = (n | (n << 1) | (n >> 1)) & pass[y]; if (m == n) break; n = m; }
            if (n != v) { f[y] = n; changed = true; }
        }
        if (!changed) break;
    }
    for (int y = 0; y < lp.H; y++) {
        uint32_t v = f[y];
        uint32_t near = v | (v << 1) | (v >> 1);
        if (y > 0) near |= f[y - 1];
        if (y + 1 < lp.H) near |= f[y + 1];
        uint32_t things = (g.m->occ[y] & ~g.m->wallmask[y]) | g.m->doorcell[y];   // non-wall, non-empty cells
        if (things & ~near) return GEN_REJECT;
    }
    return GEN_OK;
#endif
}

// ObjDesc.find_matching_objs(env, use_location=True) over the object table
// (verifier.py:96-161): every object is on the grid at generation time.
// (out of line in the cooperative device build: everything it needs comes by value, so the caller's GenCtx stays in registers)
struct MatchPose { int nobj, ax, ay, adir; };
BB_GEN_FN uint32_t g_match_pose(const LevelParams &lp, const GenMem *gm, const MatchPose mp, int type, int color, int loc)
{
    BB_ASSUME_SHARED(gm);
    const int S = lp.room_size;
    const int rtx = (mp.ax / (S - 1)) * (S - 1), rty = (mp.ay / (S - 1)) * (S - 1);   // agent room top
    const int d1x = dir_dx(mp.adir), d1y = dir_dy(mp.adir), d2x = -d1y, d2y = d1x;
    uint32_t m = 0;
#if BB_GEN_WARP
    const int k0 = threadIdx.x & 31, kstep = 32;       // lane k tests object k, the mask is a ballot
#else
    const int k0 = 0, kstep = 1;
#endif
    for (int k = k0; k < mp.nobj; k += kstep) {
        int tc = gm->obj.tc[k];
        if (type != ANY_TYPE && (tc & 7) != type) continue;
        if (color != ANY && (tc >> 3) != color) continue;
        if (loc != LOC_NONE) {
            int x = gm->obj.x[k], y = gm->obj.y[k];
            if (x < rtx || y < rty || x >= rtx + S || y >= rty + S) continue;   // Room.pos_inside
            int vx = x - mp.ax, vy = y - mp.ay;
            int dot1 = vx * d1x + vy * d1y, dot2 = vx * d2x + vy * d2y;
            bool ok = loc == LOC_LEFT ? dot2 < 0 : loc == LOC_RIGHT ? dot2 > 0 : loc == LOC_FRONT ? dot1 > 0 : dot1 < 0;
            if (!ok) continue;
        }
        m |= 1u << k;
    }
#if BB_GEN_WARP
    m = __reduce_or_sync(0xFFFFFFFFu, m);
#endif
    return m;
}
BB_HD uint32_t g_match(const LevelParams &lp, const GenCtx &g, const LevelOut &, int type, int color, int loc)
{
    const MatchPose mp = { g.nobj, g.ax, g.ay, g.adir };
    return g_match_pose(lp, g.m, mp, type, color, loc);
}

// LevelGen.rand_obj (levelgen.py:354-395); types: 4 = OBJ_TYPES, 3 = NOT_DOOR, 1 = ['door']
// Out of line in the cooperative device build, with its own copy of the random stream: takes and returns it BY VALUE, so that
// the caller's GenCtx never has its address taken and stays in registers.
struct RandObjRes { Rng rng; int status; };
BB_GEN_FN RandObjRes g_rand_obj_v(const LevelParams &lp, GenMem *gm, Rng rng, const MatchPose mp, int locked_room, int ntypes, int d)
{
    BB_ASSUME_SHARED(gm);
    const int S = lp.room_size;
    RandObjRes res;
    int tries = 0;
    for (;;) {
        if (tries > 100) { res.rng = rng; res.status = GEN_RECURSION; return res; }
        tries++;
        int ci = rng.randint(0, 7);                                  // [None, *COLOR_NAMES]
        int color = ci == 0 ? ANY : color_by_name_rank(ci - 1);
        int ti = rng.randint(0, ntypes);
        int type = ntypes == 1 ? T_DOOR : (ti == 0 ? T_BOX : ti == 1 ? T_BALL : ti == 2 ? T_KEY : T_DOOR);
        int loc = LOC_NONE;
        if (lp.locations && rng.randbool()) loc = rng.randint(0, 4);   // LOC_NAMES order
        uint32_t m = g_match_pose(lp, gm, mp, type, color, loc);
        if (m == 0) continue;
        if (!lp.implicit_unlock && locked_room >= 0) {
            // at least one match outside the (possibly stale) locked room's rectangle
            int ltx = (locked_room % lp.num_cols) * (S - 1), lty = (locked_room / lp.num_cols) * (S - 1);
            bool outside = false;
            for (uint32_t mm = m; mm; mm &= mm - 1) {
                int k = ffs32(mm);
                int x = gm->obj.x[k], y = gm->obj.y[k];
                if (x < ltx || y < lty || x >= ltx + S || y >= lty + S) outside = true;
            }
            if (!outside) continue;
        }
        gm->desc_type[d] = type; gm->desc_color[d] = color; gm->desc_loc[d] = loc; gm->desc_mask[d] = m;
        res.rng = rng; res.status = GEN_OK;
        return res;
    }
}
BB_HD int g_rand_obj(const LevelParams &lp, GenCtx &g, const LevelOut &, int ntypes, int d)
{
    const MatchPose mp = { g.nobj, g.ax, g.ay, g.adir };
    const RandObjRes r = g_rand_obj_v(lp, g.m, g.rng, mp, g.locked_room, ntypes, d);
    g.rng = r.rng;
    return r.status;
}

// one ActionInstr of rand_instr (levelgen.py:409-424) into leaf slot `leaf`
BB_HD int g_rand_action(const LevelParams &lp, GenCtx &g, const LevelOut &o, int leaf)
{
    int action = lp.action_kinds[g.rng.randint(0, lp.n_action_kinds)];
    g.m->leaf_kind[leaf] = action;
    if (action == I_GOTO) return g_rand_obj(lp, g, o, 4, 2 * leaf);
    if (action == I_PICKUP) return g_rand_obj(lp, g, o, 3, 2 * leaf);
    if (action == I_OPEN) return g_rand_obj(lp, g, o, 1, 2 * leaf);
    BB_TRY(g_rand_obj(lp, g, o, 3, 2 * leaf));
    return g_rand_obj(lp, g, o, 4, 2 * leaf + 1);
}

// rand_instr for one side of a sequence / the whole instruction: 'action' or 'and'
BB_HD int g_rand_side(const LevelParams &lp, GenCtx &g, const LevelOut &o, int side, int kind)
{
    if (kind == K_ACTION) return g_rand_action(lp, g, o, 2 * side);
    g.side_and |= 1 << side;
    BB_TRY(g_rand_action(lp, g, o, 2 * side));
    return g_rand_action(lp, g, o, 2 * side + 1);
}

// LevelGen.rand_instr (levelgen.py:397-460), canonicalised: side A [, side B]
BB_HD int g_rand_instr(const LevelParams &lp, GenCtx &g, const LevelOut &o)
{
    int kind = lp.instr_kinds[g.rng.randint(0, lp.n_instr_kinds)];
    if (kind != K_SEQ) { g.root_kind = R_SINGLE; return g_rand_side(lp, g, o, 0, kind); }
    int ka = g.rng.randint(0, 2);            // instr_kinds=['action','and']
    BB_TRY(g_rand_side(lp, g, o, 0, ka));
    int kb = g.rng.randint(0, 2);
    BB_TRY(g_rand_side(lp, g, o, 1, kb));
    g.root_kind = g.rng.randint(0, 2) == 0 ? R_BEFORE : R_AFTER;
    return GEN_OK;
}

// RoomGridLevel.validate_instrs (levelgen.py:104-155)
BB_HD int g_validate(const LevelParams &lp, const GenCtx &g, const LevelOut &o)
{
    const bool unblocking = lp.kind == KIND_LEVELGEN && lp.unblocking;
    const int locked_color = g.locked_door >= 0 ? (g.m->obj.tc[g.locked_door] >> 3) : -1;
    for (int leaf = 0; leaf < 4; leaf++) {
        int kind = g.m->leaf_kind[leaf];
        if (kind == I_NONE) continue;
        if (kind == I_PUTNEXT) {
            uint32_t mv = g.m->desc_mask[2 * leaf], fx = g.m->desc_mask[2 * leaf + 1];
            if (mv & fx) return GEN_REJECT;
            for (uint32_t a = mv; a; a &= a - 1) {           // objs_next(), verifier.py:379-391
                int ia = ffs32(a);
                for (uint32_t b = fx; b; b &= b - 1) {
                    int ib = ffs32(b);
                    if (iabs((int)g.m->obj.x[ia] - (int)g.m->obj.x[ib]) + iabs((int)g.m->obj.y[ia] - (int)g.m->obj.y[ib]) == 1)
                        return GEN_REJECT;
                }
            }
        }
        if (unblocking && locked_color >= 0) {
            int nd = kind == I_PUTNEXT ? 2 : 1;
            for (int q = 0; q < nd; q++)
                if (g.m->desc_type[2 * leaf + q] == T_KEY && g.m->desc_color[2 * leaf + q] == locked_color) return GEN_REJECT;
        }
    }
    return GEN_OK;
}

// LevelGen.add_locked_room (levelgen.py:321-352)
BB_HD int g_add_locked_room(const LevelParams &lp, GenCtx &g, const LevelOut &o)
{
    int door;
    for (;;) {
        int i = g.rng.randint(0, lp.num_cols);
        int j = g.rng.randint(0, lp.num_rows);
        int k = g.rng.randint(0, 4);
        g.locked_room = j * lp.num_cols + i; g.locked_room_fresh = true;
        if (!g_has_slot(lp, g.locked_room, k)) continue;
        int color = color_by_name_rank(g.rng.randint(0, 6));     // add_door(color=None) -> _rand_color()
        door = g_add_door(lp, g, o, g.locked_room, k, color, true);
        break;
    }
    for (;;) {
        int i = g.rng.randint(0, lp.num_cols);
        int j = g.rng.randint(0, lp.num_rows);
        int room = j * lp.num_cols + i;
        if (room == g.locked_room) continue;
        int id;
        BB_TRY(g_add_object(lp, g, o, room, T_KEY, g.m->obj.tc[door] >> 3, id));
        break;
    }
    return GEN_OK;
}

BB_HD void g_single_desc(GenCtx &g, const LevelParams &lp, const LevelOut &o, int kind, int obj_id)
{
    int tc = g.m->obj.tc[obj_id];
    g.root_kind = R_SINGLE; g.side_and = 0;
    g.m->leaf_kind[0] = kind;
    g.m->desc_type[0] = tc & 7; g.m->desc_color[0] = tc >> 3; g.m->desc_loc[0] = LOC_NONE;
    g.m->desc_mask[0] = g_match(lp, g, o, tc & 7, tc >> 3, LOC_NONE);
}

// Level_GoToImpUnlock.gen_mission (iclr19_levels.py:311-355)
BB_HD int g_mission_impunlock(const LevelParams &lp, GenCtx &g, const LevelOut &o)
{
    const int id = g.rng.randint(0, lp.num_cols);
    const int jd = g.rng.randint(0, lp.num_rows);
    const int locked = jd * lp.num_cols + id;
    int k;
    do k = g.rng.randint(0, 4); while (!g_has_slot(lp, locked, k));      // add_door(door_idx=None): no door exists yet
    const int door = g_add_door(lp, g, o, locked, k, color_by_name_rank(g.rng.randint(0, 6)), true);
    for (;;) {                                 // the key goes to a different room
        const int ik = g.rng.randint(0, lp.num_cols);
        const int jk = g.rng.randint(0, lp.num_rows);
        if (ik == id && jk == jd) continue;
        int key;
        BB_TRY(g_add_object(lp, g, o, jk * lp.num_cols + ik, T_KEY, g.m->obj.tc[door] >> 3, key));
        break;
    }
    BB_TRY(g_connect_all(lp, g, o));
    int first;
    for (int i = 0; i < lp.num_cols; i++)      // columns outer, rows inner (:334-342); num_dists per unlocked room
        for (int j = 0; j < lp.num_rows; j++)
            if (j * lp.num_cols + i != locked) BB_TRY(g_add_distractors(lp, g, o, lp.num_dists, first, false, j * lp.num_cols + i));
    for (;;) {
        BB_TRY(g_place_agent(lp, g));
        if ((g.ay / (lp.room_size - 1)) * lp.num_cols + g.ax / (lp.room_size - 1) == locked) continue;
        break;
    }
    BB_TRY(g_check_reachable(lp, g));
    BB_TRY(g_add_distractors(lp, g, o, 1, first, false, locked));        // the target, behind the locked door
    g_single_desc(g, lp, o, I_GOTO, first);
    return GEN_OK;
}

// Level_Unlock.gen_mission (iclr19_levels.py:418-474); only the doors are tracked objects (see CARRY_UNTRACKED)
BB_HD int g_mission_unlock(const LevelParams &lp, GenCtx &g, const LevelOut &o)
{
    const int id = g.rng.randint(0, lp.num_cols);
    const int jd = g.rng.randint(0, lp.num_rows);
    const int locked = jd * lp.num_cols + id;
    int k;
    do k = g.rng.randint(0, 4); while (!g_has_slot(lp, locked, k));
    const int door = g_add_door(lp, g, o, locked, k, color_by_name_rank(g.rng.randint(0, 6)), true);
    const int door_color = g.m->obj.tc[door] >> 3;
    for (;;) {
        const int ik = g.rng.randint(0, lp.num_cols);
        const int jk = g.rng.randint(0, lp.num_rows);
        if (ik == id && jk == jd) continue;
        int key;
        BB_TRY(g_add_object(lp, g, o, jk * lp.num_cols + ik, T_KEY, door_color, key, true));
        break;
    }
    // with probability 1/2 the locked door is the only door of its colour (_rand_bool: randint(0, 2) == 0)
    if (g.rng.randint(0, 2) == 0) BB_TRY(g_connect_all(lp, g, o, door_color));
    else BB_TRY(g_connect_all(lp, g, o));
    int first;
    for (int i = 0; i < lp.num_cols; i++)
        for (int j = 0; j < lp.num_rows; j++)
            if (j * lp.num_cols + i != locked) BB_TRY(g_add_distractors(lp, g, o, lp.num_dists, first, false, j * lp.num_cols + i, true));
    for (;;) {
        BB_TRY(g_place_agent(lp, g));
        if ((g.ay / (lp.room_size - 1)) * lp.num_cols + g.ax / (lp.room_size - 1) == locked) continue;
        break;
    }
    BB_TRY(g_check_reachable(lp, g));
    g_single_desc(g, lp, o, I_OPEN, door);
    return GEN_OK;
}

// gen_mission of the level families.  IMPUNLOCK selects the instantiation that only serves KIND_IMPUNLOCK, so that the code
// generated for the other families (register allocation of generate_level inside k_gen) stays the one profiled in round 1.
// ---- bonus_levels.py ------------------------------------------------------------------------------------------------
// RoomGrid.remove_wall(i, j, wall_idx): the wall cells between two rooms (corners stay) become empty; both rooms count as
// connected through that side (room.doors[idx] = True: connect_all sees a door, add_door sees an occupied slot)
BB_HD void g_remove_wall(const LevelParams &lp, GenCtx &g, int room, int k)
{
    const int S = lp.room_size;
    const int owner = k == 2 ? room - 1 : k == 3 ? room - lp.num_cols : room;
    const int tx = (owner % lp.num_cols) * (S - 1), ty = (owner / lp.num_cols) * (S - 1);
    if (k == 0 || k == 2) {                    // the right wall of `owner`
        for (int i = 1; i < S - 1; i++) { g.m->wallmask[ty + i] &= ~(1u << (tx + S - 1)); g.m->occ[ty + i] &= ~(1u << (tx + S - 1)); }
        g.door_right |= 1u << owner;
    } else {                                   // its bottom wall
        const uint32_t bits = ((1u << (S - 2)) - 1u) << (tx + 1);
        g.m->wallmask[ty + S - 1] &= ~bits; g.m->occ[ty + S - 1] &= ~bits;
        g.door_down |= 1u << owner;
    }
}
// RoomGrid.add_door(i, j, door_idx=None, color=None, locked=None): the draws a missing argument costs, in the reference's order
BB_HD int g_add_door_rand(const LevelParams &lp, GenCtx &g, const LevelOut &o, int room, int k, int color, int locked)
{
    if (k < 0) for (;;) { k = g.rng.randint(0, 4); if (g_has_slot(lp, room, k) && !g_has_door(lp, g, room, k)) break; }
    if (color < 0) color = color_by_name_rank(g.rng.randint(0, 6));
    if (locked < 0) locked = g.rng.randbool() ? 1 : 0;
    return g_add_door(lp, g, o, room, k, color, locked != 0);
}
// RoomGrid.add_object(i, j, kind=None, color=None)
BB_HD int g_add_object_rand(const LevelParams &lp, GenCtx &g, const LevelOut &o, int room, int type, int color, int &id)
{
    if (type < 0) { const int t = g.rng.randint(0, 3); type = t == 0 ? T_KEY : t == 1 ? T_BALL : T_BOX; }
    if (color < 0) color = color_by_name_rank(g.rng.randint(0, 6));
    return g_add_object(lp, g, o, room, type, color, id);
}
// a descriptor by (type, color, loc) into slot d; ObjDesc matching happens at reset_verifier, i.e. with the final agent pose
BB_HD void g_desc(const LevelParams &lp, GenCtx &g, const LevelOut &o, int d, int type, int color, int loc)
{
    g.m->desc_type[d] = type; g.m->desc_color[d] = color; g.m->desc_loc[d] = loc;
    g.m->desc_mask[d] = g_match(lp, g, o, type, color, loc) & ~g.hidden_mask;
}
// MiniGridEnv._rand_subset(COLOR_NAMES, n): colours in draw order
BB_HD void g_rand_colors(GenCtx &g, int n, int out[6])
{
    int left[6], nl = 6;
    for (int c = 0; c < 6; c++) left[c] = color_by_name_rank(c);
    for (int k = 0; k < n; k++) {
        const int i = g.rng.randint(0, nl);
        out[k] = left[i];
        for (int c = i; c + 1 < nl; c++) left[c] = left[c + 1];
        nl--;
    }
}

BB_HD int g_mission_bonus(const LevelParams &lp, GenCtx &g, const LevelOut &o)
{
    const int C = lp.num_cols, mid = (lp.num_rows > 1 && C > 1) ? 1 * C + 1 : 0;      // room (1, 1) of a 3 x 3 grid
    int id = 0, first = 0;
    switch (lp.bonus) {
    case BN_GOTO_REDBLUE: {                 // Level_GoToRedBlueBall :24-40
        BB_TRY(g_place_agent(lp, g));
        BB_TRY(g_add_distractors(lp, g, o, lp.num_dists, first));
        for (int k = first; k < g.nobj; k++) {
            const int tc = g.m->obj.tc[k];
            if ((tc & 7) == T_BALL && ((tc >> 3) == C_BLUE || (tc >> 3) == C_RED)) return GEN_REJECT;
        }
        const int color = g.rng.randint(0, 2) == 0 ? C_RED : C_BLUE;
        BB_TRY(g_add_object(lp, g, o, 0, T_BALL, color, id));
        BB_TRY(g_check_reachable(lp, g));
        g.m->leaf_kind[0] = I_GOTO; g_desc(lp, g, o, 0, T_BALL, color, LOC_NONE);
        return GEN_OK;
    }
    case BN_OPEN_RED_DOOR: {                // Level_OpenRedDoor :59-62
        g_add_door(lp, g, o, 0, 0, C_RED, false);
        BB_TRY(g_place_agent(lp, g, 0));
        g.m->leaf_kind[0] = I_OPEN; g_desc(lp, g, o, 0, T_DOOR, C_RED, LOC_NONE);
        return GEN_OK;
    }
    case BN_OPEN_DOOR: {                    // Level_OpenDoor :82-99; bonus_a: 0 = select_by None, 1 = "color", 2 = "loc"
        int colors[6];
        g_rand_colors(g, 4, colors);
        int door0 = -1;
        for (int i = 0; i < 4; i++) { const int d = g_add_door(lp, g, o, mid, i, colors[i], false); if (i == 0) door0 = d; }
        int sel = lp.bonus_a;
        if (sel == 0) sel = g.rng.randint(0, 2) == 0 ? 1 : 2;
        int loc = LOC_NONE;
        if (sel == 2) loc = g.rng.randint(0, 4);
        BB_TRY(g_place_agent(lp, g, mid));
        g.m->leaf_kind[0] = I_OPEN;
        if (sel == 1) g_desc(lp, g, o, 0, T_DOOR, g.m->obj.tc[door0] >> 3, LOC_NONE); else g_desc(lp, g, o, 0, T_DOOR, ANY, loc);
        return GEN_OK;
    }
    case BN_GOTO_DOOR: {                    // Level_GoToDoor :163-171
        int doors[4];
        for (int i = 0; i < 4; i++) doors[i] = g_add_door_rand(lp, g, o, mid, -1, -1, -1);
        BB_TRY(g_place_agent(lp, g, mid));
        const int pick = doors[g.rng.randint(0, 4)];
        g.m->leaf_kind[0] = I_GOTO; g_desc(lp, g, o, 0, T_DOOR, g.m->obj.tc[pick] >> 3, LOC_NONE);
        return GEN_OK;
    }
    case BN_GOTO_OBJ_DOOR: {                // Level_GoToObjDoor :186-197
        BB_TRY(g_place_agent(lp, g, mid));
        BB_TRY(g_add_distractors(lp, g, o, 8, first, false, mid));
        for (int i = 0; i < 4; i++) g_add_door_rand(lp, g, o, mid, -1, -1, -1);
        BB_TRY(g_check_reachable(lp, g));
        const int pick = first + g.rng.randint(0, 12);            // 8 distractors, then the 4 doors, in that order
        const int tc = g.m->obj.tc[pick];
        g.m->leaf_kind[0] = I_GOTO; g_desc(lp, g, o, 0, tc & 7, tc >> 3, LOC_NONE);
        return GEN_OK;
    }
    case BN_ACTION_OBJ_DOOR: {              // Level_ActionObjDoor :214-234 (add_distractors default: all_unique=True)
        BB_TRY(g_add_distractors(lp, g, o, 5, first, true, mid));
        for (int i = 0; i < 4; i++) g_add_door_rand(lp, g, o, mid, -1, -1, 0);
        BB_TRY(g_place_agent(lp, g, mid));
        const int pick = first + g.rng.randint(0, 9);
        const int tc = g.m->obj.tc[pick];
        const bool goto_it = g.rng.randbool();
        g.m->leaf_kind[0] = goto_it ? I_GOTO : ((tc & 7) == T_DOOR ? I_OPEN : I_PICKUP);
        g_desc(lp, g, o, 0, tc & 7, tc >> 3, LOC_NONE);
        return GEN_OK;
    }
    case BN_UNLOCK_LOCAL: {                 // Level_UnlockLocal :247-254; bonus_a: distractors
        const int door = g_add_door_rand(lp, g, o, mid, -1, -1, 1);
        BB_TRY(g_add_object(lp, g, o, mid, T_KEY, g.m->obj.tc[door] >> 3, id));
        if (lp.bonus_a) BB_TRY(g_add_distractors(lp, g, o, 3, first, true, mid));
        BB_TRY(g_place_agent(lp, g, mid));
        g.m->leaf_kind[0] = I_OPEN; g_desc(lp, g, o, 0, T_DOOR, ANY, LOC_NONE);
        return GEN_OK;
    }
    case BN_KEY_IN_BOX: {                   // Level_KeyInBox :277-287: the key is inside the box, not on the grid
        const int door = g_add_door_rand(lp, g, o, mid, -1, -1, 1);
        const int bcolor = color_by_name_rank(g.rng.randint(0, 6));
        BB_TRY(g_add_object(lp, g, o, mid, T_BOX, bcolor, id));   // object `id`; its contents is the next table entry
        const int key = g.nobj++;
        g.m->obj.x[key] = 0; g.m->obj.y[key] = 0; g.m->obj.tc[key] = (uint8_t)(T_KEY | ((g.m->obj.tc[door] >> 3) << 3));
        g.hidden_mask |= 1u << key;
        BB_TRY(g_place_agent(lp, g, mid));
        g.m->leaf_kind[0] = I_OPEN; g_desc(lp, g, o, 0, T_DOOR, ANY, LOC_NONE);
        return GEN_OK;
    }
    case BN_UNLOCK_PICKUP: {                // Level_UnlockPickup :307-319 (1 x 2 rooms); bonus_a: distractors
        int obj;
        BB_TRY(g_add_object_rand(lp, g, o, 1, T_BOX, -1, obj));
        const int door = g_add_door_rand(lp, g, o, 0, 0, -1, 1);
        BB_TRY(g_add_object(lp, g, o, 0, T_KEY, g.m->obj.tc[door] >> 3, id));
        if (lp.bonus_a) BB_TRY(g_add_distractors(lp, g, o, 4, first, true, -1));
        BB_TRY(g_place_agent(lp, g, 0));
        g.m->leaf_kind[0] = I_PICKUP; g_desc(lp, g, o, 0, T_BOX, g.m->obj.tc[obj] >> 3, LOC_NONE);
        return GEN_OK;
    }
    case BN_BLOCKED_UNLOCK_PICKUP: {        // Level_BlockedUnlockPickup :348-361: a ball set directly in front of the locked door
        int obj;
        BB_TRY(g_add_object_rand(lp, g, o, 1, T_BOX, -1, obj));
        const int door = g_add_door_rand(lp, g, o, 0, 0, -1, 1);
        const int bcolor = color_by_name_rank(g.rng.randint(0, 6));
        {   // self.grid.set(pos[0] - 1, pos[1], Ball(color)): no rejection sampling, whatever was there is replaced
            const int bx = g.m->obj.x[door] - 1, by = g.m->obj.y[door];
            for (int k = 0; k < g.nobj; k++)
                if (g.m->obj.x[k] == bx && g.m->obj.y[k] == by && !((g.hidden_mask >> k) & 1u)) g.hidden_mask |= 1u << k;   // (an overwritten object leaves the grid)
            const int ball = g.nobj++;
            g.m->obj.x[ball] = (uint8_t)bx; g.m->obj.y[ball] = (uint8_t)by; g.m->obj.tc[ball] = (uint8_t)(T_BALL | (bcolor << 3));
            g.m->occ[by] |= 1u << bx;
        }
        BB_TRY(g_add_object(lp, g, o, 0, T_KEY, g.m->obj.tc[door] >> 3, id));
        BB_TRY(g_place_agent(lp, g, 0));
        g.m->leaf_kind[0] = I_PICKUP; g_desc(lp, g, o, 0, T_BOX, ANY, LOC_NONE);
        return GEN_OK;
    }
    case BN_UNLOCK_TO_UNLOCK: {             // Level_UnlockToUnlock :379-398 (1 x 3 rooms)
        int colors[6];
        g_rand_colors(g, 2, colors);
        g_add_door(lp, g, o, 0, 0, colors[0], true);
        BB_TRY(g_add_object(lp, g, o, 2, T_KEY, colors[0], id));
        g_add_door(lp, g, o, 1, 0, colors[1], true);
        BB_TRY(g_add_object(lp, g, o, 1, T_KEY, colors[1], id));
        int obj;
        BB_TRY(g_add_object_rand(lp, g, o, 0, T_BALL, -1, obj));
        BB_TRY(g_place_agent(lp, g, 1));
        g.m->leaf_kind[0] = I_PICKUP; g_desc(lp, g, o, 0, T_BALL, ANY, LOC_NONE);
        return GEN_OK;
    }
    case BN_PICKUP_DIST: {                  // Level_PickupDist :418-432
        BB_TRY(g_add_distractors(lp, g, o, 5, first, true, -1));
        BB_TRY(g_place_agent(lp, g, 0));
        const int tc = g.m->obj.tc[first + g.rng.randint(0, 5)];
        const int sel = g.rng.randint(0, 3);                      // ["type", "color", "both"]
        g.m->leaf_kind[0] = I_PICKUP;
        g_desc(lp, g, o, 0, sel == 1 ? ANY_TYPE : (tc & 7), sel == 0 ? ANY : (tc >> 3), LOC_NONE);
        return GEN_OK;
    }
    case BN_PICKUP_ABOVE: {                 // Level_PickupAbove :461-469: the object is in room (1, 0)
        int obj;
        BB_TRY(g_add_object_rand(lp, g, o, 1, -1, -1, obj));
        g_add_door_rand(lp, g, o, mid, 3, -1, 0);
        BB_TRY(g_place_agent(lp, g, mid));
        BB_TRY(g_connect_all(lp, g, o));
        g.m->leaf_kind[0] = I_PICKUP; g_desc(lp, g, o, 0, g.m->obj.tc[obj] & 7, g.m->obj.tc[obj] >> 3, LOC_NONE);
        return GEN_OK;
    }
    case BN_OPEN_TWO_DOORS: {               // Level_OpenTwoDoors :497-515; bonus_a / bonus_b: first / second colour + 1, 0 = drawn
        int colors[6];
        g_rand_colors(g, 2, colors);
        const int c1 = lp.bonus_a ? lp.bonus_a - 1 : colors[0], c2 = lp.bonus_b ? lp.bonus_b - 1 : colors[1];
        g_add_door(lp, g, o, mid, 2, c1, false);
        g_add_door(lp, g, o, mid, 0, c2, false);
        BB_TRY(g_place_agent(lp, g, mid));
        g.root_kind = R_BEFORE;
        g.m->leaf_kind[0] = I_OPEN; g_desc(lp, g, o, 0, T_DOOR, c1, LOC_NONE);
        g.m->leaf_kind[2] = I_OPEN; g_desc(lp, g, o, 4, T_DOOR, c2, LOC_NONE);
        return GEN_OK;
    }
    case BN_FIND_OBJ: {                     // Level_FindObjS5 :579-587 (i from num_rows, j from num_cols, used as (i, j))
        const int i = g.rng.randint(0, lp.num_rows);
        const int j = g.rng.randint(0, lp.num_cols);
        int obj;
        BB_TRY(g_add_object_rand(lp, g, o, j * C + i, -1, -1, obj));
        BB_TRY(g_place_agent(lp, g, mid));
        BB_TRY(g_connect_all(lp, g, o));
        g.m->leaf_kind[0] = I_PICKUP; g_desc(lp, g, o, 0, g.m->obj.tc[obj] & 7, ANY, LOC_NONE);
        return GEN_OK;
    }
    case BN_KEY_CORRIDOR: {                 // KeyCorridor :636-656 (3 columns, num_rows rows); bonus_a: obj_type
        for (int j = 1; j < lp.num_rows; j++) g_remove_wall(lp, g, j * C + 1, 3);
        const int row = g.rng.randint(0, lp.num_rows);
        const int door = g_add_door_rand(lp, g, o, row * C + 2, 2, -1, 1);
        int obj;
        BB_TRY(g_add_object_rand(lp, g, o, row * C + 2, lp.bonus_a, -1, obj));
        BB_TRY(g_add_object(lp, g, o, g.rng.randint(0, lp.num_rows) * C + 0, T_KEY, g.m->obj.tc[door] >> 3, id));
        BB_TRY(g_place_agent(lp, g, (lp.num_rows / 2) * C + 1));
        BB_TRY(g_connect_all(lp, g, o));
        g.m->leaf_kind[0] = I_PICKUP; g_desc(lp, g, o, 0, g.m->obj.tc[obj] & 7, ANY, LOC_NONE);
        return GEN_OK;
    }
    case BN_ONE_ROOM: {                     // Level_1RoomS8 :721-724
        int obj;
        BB_TRY(g_add_object_rand(lp, g, o, 0, T_BALL, -1, obj));
        BB_TRY(g_place_agent(lp, g));
        g.m->leaf_kind[0] = I_PICKUP; g_desc(lp, g, o, 0, T_BALL, ANY, LOC_NONE);
        return GEN_OK;
    }
    case BN_PUTNEXT: case BN_MOVE_TWO_ACROSS: {   // PutNext :793-819 / MoveTwoAcross :931-953 (1 x 2 rooms); bonus_a: objs_per_room
        const int n = lp.bonus_a;
        int fl, fr;
        BB_TRY(g_place_agent(lp, g, 0));
        BB_TRY(g_add_distractors(lp, g, o, n, fl, true, 0));
        BB_TRY(g_add_distractors(lp, g, o, n, fr, true, 1));
        g_remove_wall(lp, g, 0, 0);
        if (lp.bonus == BN_PUTNEXT) {
            int a = fl + g.rng.randint(0, n), b = fr + g.rng.randint(0, n);
            if (g.rng.randbool()) { const int t = a; a = b; b = t; }
            g.m->leaf_kind[0] = I_PUTNEXT;
            g_desc(lp, g, o, 0, g.m->obj.tc[a] & 7, g.m->obj.tc[a] >> 3, LOC_NONE);
            g_desc(lp, g, o, 1, g.m->obj.tc[b] & 7, g.m->obj.tc[b] >> 3, LOC_NONE);
            if (lp.bonus_b) g.start_carry = a;                    // start_carrying: obj_a, from the first step on
        } else {
            int l0 = g.rng.randint(0, n), l1 = g.rng.randint(0, n - 1); if (l1 >= l0) l1++;      // _rand_subset(objs_l, 2)
            int r0 = g.rng.randint(0, n), r1 = g.rng.randint(0, n - 1); if (r1 >= r0) r1++;
            const int a = fl + l0, d = fl + l1, b = fr + r0, c = fr + r1;
            g.root_kind = R_BEFORE;
            g.m->leaf_kind[0] = I_PUTNEXT;
            g_desc(lp, g, o, 0, g.m->obj.tc[a] & 7, g.m->obj.tc[a] >> 3, LOC_NONE);
            g_desc(lp, g, o, 1, g.m->obj.tc[b] & 7, g.m->obj.tc[b] >> 3, LOC_NONE);
            g.m->leaf_kind[2] = I_PUTNEXT;
            g_desc(lp, g, o, 4, g.m->obj.tc[c] & 7, g.m->obj.tc[c] >> 3, LOC_NONE);
            g_desc(lp, g, o, 5, g.m->obj.tc[d] & 7, g.m->obj.tc[d] >> 3, LOC_NONE);
        }
        return GEN_OK;
    }
    case BN_OPEN_DOORS_ORDER: {             // OpenDoorsOrder :996-1016; bonus_a: num_doors
        const int nd = lp.bonus_a;
        int colors[6], doors[6];
        g_rand_colors(g, nd, colors);
        for (int i = 0; i < nd; i++) doors[i] = g_add_door_rand(lp, g, o, mid, -1, colors[i], 0);
        BB_TRY(g_place_agent(lp, g, mid));
        int i1 = g.rng.randint(0, nd), i2 = g.rng.randint(0, nd - 1); if (i2 >= i1) i2++;         // _rand_subset(doors, 2)
        const int mode = g.rng.randint(0, 3);
        g.m->leaf_kind[0] = I_OPEN; g_desc(lp, g, o, 0, T_DOOR, g.m->obj.tc[doors[i1]] >> 3, LOC_NONE);
        if (mode != 0) {
            g.root_kind = mode == 1 ? R_BEFORE : R_AFTER;
            g.m->leaf_kind[2] = I_OPEN; g_desc(lp, g, o, 4, T_DOOR, g.m->obj.tc[doors[i2]] >> 3, LOC_NONE);
        }
        return GEN_OK;
    }
    }
    return GEN_RECURSION;
}

template <bool IMPUNLOCK>
BB_HD int g_mission(const LevelParams &lp, GenCtx &g, const LevelOut &o)
{
    for (int k = 0; k < 4; k++) g.m->leaf_kind[k] = I_NONE;
    for (int k = 0; k < 8; k++) { g.m->desc_mask[k] = 0; g.m->desc_type[k] = ANY_TYPE; g.m->desc_color[k] = ANY; g.m->desc_loc[k] = LOC_NONE; }
    g.side_and = 0; g.root_kind = R_SINGLE;
    if constexpr (IMPUNLOCK) return lp.kind == KIND_UNLOCK ? g_mission_unlock(lp, g, o) : lp.kind == KIND_BONUS ? g_mission_bonus(lp, g, o) : g_mission_impunlock(lp, g, o);
    if (lp.kind == KIND_REDBALL) {                 // iclr19_levels.py:26-37, 55-63
        int ball, first;
        BB_TRY(g_place_agent(lp, g));
        BB_TRY(g_add_object(lp, g, o, 0, T_BALL, C_RED, ball));
        BB_TRY(g_add_distractors(lp, g, o, lp.num_dists, first));
        if (lp.grey_dists)
            for (int k = first; k < g.nobj; k++) g.m->obj.tc[k] = (uint8_t)((g.m->obj.tc[k] & 7) | (C_GREY << 3));
        BB_TRY(g_check_reachable(lp, g));
        g_single_desc(g, lp, o, I_GOTO, ball);
        return GEN_OK;
    }
    if (lp.kind == KIND_OBJ) {                     // iclr19_levels.py:88-92, 119-124, 202-211, 247-257, 365-415, 482-491
        int first;
        BB_TRY(g_place_agent(lp, g));
        BB_TRY(g_connect_all(lp, g, o));
        BB_TRY(g_add_distractors(lp, g, o, lp.num_dists, first, lp.all_unique != 0));
        if (lp.require_unreachable) { if (g_check_reachable(lp, g) == GEN_OK) return GEN_REJECT; }   // Level_UnblockPickup
        else BB_TRY(g_check_reachable(lp, g));
        if (lp.instr == I_OPEN) {
            // doors as Level_Open lists them: every door once per adjacent room; columns outer, rows inner, sides 0..3
            int nd = 0;
            for (int i = 0; i < lp.num_cols; i++)
                for (int j = 0; j < lp.num_rows; j++)
                    for (int k = 0; k < 4; k++) { const int r = j * lp.num_cols + i; if (g_has_slot(lp, r, k) && g_has_door(lp, g, r, k)) nd++; }
            int pick = g.rng.randint(0, nd), door = 0;
            for (int i = 0; i < lp.num_cols; i++)
                for (int j = 0; j < lp.num_rows; j++)
                    for (int k = 0; k < 4; k++) {
                        const int r = j * lp.num_cols + i;
                        if (!(g_has_slot(lp, r, k) && g_has_door(lp, g, r, k))) continue;
                        if (pick-- == 0)
                            door = k == 0 ? g.m->door_id_right[r] : k == 1 ? g.m->door_id_down[r]
                                 : k == 2 ? g.m->door_id_right[r - 1] : g.m->door_id_down[r - lp.num_cols];
                    }
            g_single_desc(g, lp, o, I_OPEN, door);
        } else if (lp.instr == I_PUTNEXT) {          // o1, o2 = self._rand_subset(objs, 2)
            const int i1 = g.rng.randint(0, lp.num_dists);
            int i2 = g.rng.randint(0, lp.num_dists - 1);
            if (i2 >= i1) i2++;                      // index into the list with o1 removed
            g_single_desc(g, lp, o, I_PUTNEXT, first + i1);
            const int tc2 = g.m->obj.tc[first + i2];
            g.m->desc_type[1] = tc2 & 7; g.m->desc_color[1] = tc2 >> 3; g.m->desc_loc[1] = LOC_NONE;
            g.m->desc_mask[1] = g_match(lp, g, o, tc2 & 7, tc2 >> 3, LOC_NONE);
        } else {
            int pick = first + g.rng.randint(0, lp.num_dists);
            g_single_desc(g, lp, o, lp.instr, pick);
        }
        return GEN_OK;
    }
    // LevelGen.gen_mission, levelgen.py:293-319
    if ((uint64_t)g.rng.u32() < lp.locked_thr) BB_TRY(g_add_locked_room(lp, g, o));
    BB_TRY(g_connect_all(lp, g, o));
    int first;
    BB_TRY(g_add_distractors(lp, g, o, lp.num_dists, first));
    for (;;) {
        BB_TRY(g_place_agent(lp, g));
        int start = (g.ay / (lp.room_size - 1)) * lp.num_cols + g.ax / (lp.room_size - 1);
        if (g.locked_room_fresh && start == g.locked_room) continue;   // `start_room is self.locked_room`
        break;
    }
    if (!lp.unblocking) BB_TRY(g_check_reachable(lp, g));
    return g_rand_instr(lp, g, o);
}

// ---- mission tokens (Instr.surface / ObjDesc.surface, verifier.py:64-94 ...) --
BB_GEN_FN int tok_desc(const GenMem *gm, int d, int16_t *tok, int n)
{
    BB_ASSUME_SHARED(gm); BB_ASSUME_SHARED(tok);
    // ('a' when several objects match.  find_matching_objs scans every grid cell, walls included: a description by colour
    // alone -- ObjDesc(None, 'grey'), Level_PickupDist -- also matches the grey walls, so it is always 'a grey object')
    const bool many = popc32(gm->desc_mask[d]) > 1 || (gm->desc_type[d] == ANY_TYPE && gm->desc_color[d] == C_GREY);
    tok[n++] = many ? W_A : W_THE;
    if (gm->desc_color[d] != ANY) tok[n++] = (int16_t)(W_RED + gm->desc_color[d]);
    int t = gm->desc_type[d];
    tok[n++] = (int16_t)(t == ANY_TYPE ? W_OBJECT : t == T_BOX ? W_BOX : t == T_BALL ? W_BALL : t == T_KEY ? W_KEY : W_DOOR);
    int loc = gm->desc_loc[d];
    if (loc == LOC_FRONT) { tok[n++] = W_IN; tok[n++] = W_FRONT; tok[n++] = W_OF; tok[n++] = W_YOU; }
    else if (loc == LOC_BEHIND) { tok[n++] = W_BEHIND; tok[n++] = W_YOU; }
    else if (loc == LOC_LEFT) { tok[n++] = W_ON; tok[n++] = W_YOUR; tok[n++] = W_LEFT; }
    else if (loc == LOC_RIGHT) { tok[n++] = W_ON; tok[n++] = W_YOUR; tok[n++] = W_RIGHT; }
    return n;
}
BB_HD int tok_leaf(const GenCtx &g, int leaf, int16_t *tok, int n)
{
    int k = g.m->leaf_kind[leaf];
    if (k == I_GOTO) { tok[n++] = W_GO; tok[n++] = W_TO; }
    else if (k == I_PICKUP) { tok[n++] = W_PICK; tok[n++] = W_UP; }
    else if (k == I_OPEN) { tok[n++] = W_OPEN; }
    else { tok[n++] = W_PUT; }
    n = tok_desc(g.m, 2 * leaf, tok, n);
    if (k == I_PUTNEXT) { tok[n++] = W_NEXT; tok[n++] = W_TO; n = tok_desc(g.m, 2 * leaf + 1, tok, n); }
    return n;
}
BB_HD int tok_side(const GenCtx &g, int side, int16_t *tok, int n)
{
    n = tok_leaf(g, 2 * side, tok, n);
    if ((g.side_and >> side) & 1) { tok[n++] = W_AND; n = tok_leaf(g, 2 * side + 1, tok, n); }
    return n;
}

// four cells from four wall bits: wall (0x2A) where the bit is set, empty (0x01) elsewhere
BB_HD uint32_t g_cells4(uint32_t bits)
{
    const uint32_t e = (((bits & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu;
    return ((uint32_t)CELL_WALL * 0x01010101u & e) | ((uint32_t)CELL_EMPTY * 0x01010101u & ~e);
}

// One whole RoomGridLevel.reset() worth of generation (levelgen.py:35-47,77-102):
// retries until an attempt is accepted, then renders grid + records into `o`.
// Returns the number of attempts.
template <bool IMPUNLOCK>
BB_HD_NOINLINE int generate_level_t(const LevelParams &lp, const LevelOut &o, RngRec *rngrec, uint8_t *locked_room_persist, GenMem *mem)
{
    BB_ASSUME_SHARED(mem);
    GenCtx g;
    g.m = mem;
#if defined(__CUDACC__) && BB_GEN_COOP
    g.rng.init(rngrec->seed, rngrec->draws, mem->draw_buf);
#else
    g.rng.init(rngrec->seed, rngrec->draws, nullptr);
#endif
    g.locked_room = *locked_room_persist == 0xFF ? -1 : (int)*locked_room_persist;
    int attempts = 0;
    for (;;) {
        attempts++;
        g_roomgrid(lp, g);
        if (g_mission<IMPUNLOCK>(lp, g, o)) continue;
        if (g_validate(lp, g, o)) continue;
        break;
    }
    rngrec->draws = g.rng.count();
    *locked_room_persist = g.locked_room < 0 ? 0xFF : (uint8_t)g.locked_room;

    // ---- render the byte grid: walls, then doors/objects -----------------
    // (warp-per-level form: the warp's lanes split the cells; every lane holds the same level)
#if BB_GEN_WARP
    const int lane = threadIdx.x & 31, nlanes = 32;
#else
    const int lane = 0, nlanes = 1;
#endif
    // G (row-major) and GT (column-major), four cells per 32-bit store; row padding up to the stride is wall
    {
        uint32_t *gw = reinterpret_cast<uint32_t *>(o.grid);
        const int gwpr = lp.rs_g >> 2, twpr = lp.rs_t >> 2;
        for (int c = lane; c < lp.H * gwpr; c += nlanes) {
            const int y = c / gwpr, x0 = (c - y * gwpr) * 4;
            uint32_t bits = (g.m->wallmask[y] >> x0) & 0xFu;
            for (int b = 0; b < 4; b++) if (x0 + b >= lp.W) bits |= 1u << b;
            gw[c] = g_cells4(bits);
        }
        uint32_t *tw = reinterpret_cast<uint32_t *>(o.grid + lp.gt_off);
        for (int c = lane; c < lp.W * twpr; c += nlanes) {
            const int x = c / twpr, y0 = (c - x * twpr) * 4;
            uint32_t bits = 0;
            for (int b = 0; b < 4; b++) bits |= (y0 + b >= lp.H ? 1u : ((g.m->wallmask[y0 + b < lp.H ? y0 + b : 0] >> x) & 1u)) << b;
            tw[c] = g_cells4(bits);
        }
    }
#if BB_GEN_WARP
    __syncwarp();
#endif
    for (int k = lane; k < g.nobj; k += nlanes) {          // objects never share a cell (a hidden one is not on the grid)
        if ((g.hidden_mask >> k) & 1u) continue;
        int tc = g.m->obj.tc[k], st = 0;
        if ((tc & 7) == T_DOOR) st = lp.doors_open ? 0 : (((g.locked_mask >> k) & 1u) ? 2 : 1);   // open_all_doors levelgen.py:189-199
        set_cell(lp, o.grid, g.m->obj.x[k], g.m->obj.y[k], tc | (st << 6));
    }
    if constexpr (IMPUNLOCK) {            // KIND_UNLOCK: the untracked objects are their cell bytes
        const GenMemX *mx = static_cast<const GenMemX *>(g.m);
        for (int k = 0; k < g.nun; k++) set_cell(lp, o.grid, mx->ux[k], mx->uy[k], mx->utc[k]);
    }
    // object table: working copy -> slot
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(&g.m->obj);
        uint32_t *dst = reinterpret_cast<uint32_t *>(o.obj);
        for (int k = lane; k < (int)(sizeof(ObjTab) / 4); k += nlanes) dst[k] = src[k];
    }
    // ---- verifier record (reset_verifier) + max_steps (levelgen.py:42-45) --
    int navs = 0;
    for (int k = 0; k < 4; k++) {
        o.ins->leaf_kind[k] = (uint8_t)g.m->leaf_kind[k];
        o.ins->leaf_pre[k] = NO_OBJ;
        if (g.m->leaf_kind[k] != I_NONE) navs += g.m->leaf_kind[k] == I_PUTNEXT ? 2 : 1;
    }
    for (int k = 0; k < 8; k++) o.ins->desc_mask[k] = g.m->desc_mask[k];
    o.ins->root_kind = (uint8_t)g.root_kind; o.ins->side_and = (uint8_t)g.side_and; o.ins->flags = 0;
    o.ins->pad0 = (uint8_t)(g.start_carry == NO_OBJ ? 0 : g.start_carry + 1); o.ins->pad1 = 0;
    EnvHot h;
    h.x = (uint8_t)g.ax; h.y = (uint8_t)g.ay; h.dirflags = (uint8_t)g.adir; h.carry = NO_OBJ;
    h.step_count = 0; h.max_steps = (uint16_t)(navs * lp.nav_time_maze);
    h.cur_mask = (g.nobj >= 32 ? 0xFFFFFFFFu : ((1u << g.nobj) - 1u)) & ~g.hidden_mask;
    h.snap_mask = h.cur_mask;
    *o.hot = h;
    // ---- mission tokens ---------------------------------------------------
    int16_t *tok = g.m->tok;
    int n = tok_side(g, 0, tok, 0);
    if (g.root_kind == R_BEFORE) { tok[n++] = W_THEN; n = tok_side(g, 1, tok, n); }
    else if (g.root_kind == R_AFTER) { tok[n++] = W_AFTER; tok[n++] = W_YOU; n = tok_side(g, 1, tok, n); }
    for (int k = lane; k < lp.max_tokens; k += nlanes) o.tok[k] = k < n ? tok[k] : (int16_t)0;
    return attempts;
}

BB_HD int generate_level(const LevelParams &lp, const LevelOut &o, RngRec *rngrec, uint8_t *locked_room_persist, GenMem *mem)
{
    return (lp.kind == KIND_IMPUNLOCK || lp.kind == KIND_UNLOCK || lp.kind == KIND_BONUS) ? generate_level_t<true>(lp, o, rngrec, locked_room_persist, mem)
                                                                 : generate_level_t<false>(lp, o, rngrec, locked_room_persist, mem);
}

// =============================================================================
// Small-level generator: ONE LANE PER LEVEL, written as a single flat loop
// =============================================================================
// For single-room levels up to 8x8 (GoToRedBall*, GoToObj*, GoToLocal*, PickupLoc) the whole generator state
// fits in registers: occupancy is a 64-bit board, the objects are two packed 64-bit words.  The reference's
// nested rejection loops (attempts > placements > tries) are flattened into ONE loop whose body performs one
// RNG-consuming event chosen by a phase variable.  Every lane of a warp runs that same loop on a different
// environment: lanes differ only in predicates and trip count and reconverge at every iteration, which the
// nested-loop form (generate_level) cannot do -- that is why it had to be run one warp per level.
// Draw order and accept/reject rules are exactly those of generate_level; the two are cross-checked level by
// level in the host build (tests/hostemu) and both against the oracle.
struct SmallLevel {             // everything emit_small_level() needs
    uint64_t poss;              // 6 bits per object: x | y << 3
    uint64_t tcs;               // 6 bits per object: type | color << 3
    int nobj, ax, ay, adir;
    int leaf_kind, d_type, d_color, d_loc;
    uint32_t d_mask;
};

BB_HD int sm_obj_x(uint64_t poss, int k) { return (int)((poss >> (6 * k)) & 7u); }
BB_HD int sm_obj_y(uint64_t poss, int k) { return (int)((poss >> (6 * k + 3)) & 7u); }
BB_HD int sm_obj_tc(uint64_t tcs, int k) { return (int)((tcs >> (6 * k)) & 63u); }

// ObjDesc.find_matching_objs over the packed objects (single room: every object is in the agent's room)
BB_HD uint32_t sm_match(const SmallLevel &L, int type, int color, int loc)
{
    const int d1x = dir_dx(L.adir), d1y = dir_dy(L.adir), d2x = -d1y, d2y = d1x;
    uint32_t m = 0;
    for (int k = 0; k < L.nobj; k++) {
        const int tc = sm_obj_tc(L.tcs, k);
        if (type != ANY_TYPE && (tc & 7) != type) continue;
        if (color != ANY && (tc >> 3) != color) continue;
        if (loc != LOC_NONE) {
            const int vx = sm_obj_x(L.poss, k) - L.ax, vy = sm_obj_y(L.poss, k) - L.ay;
            const int dot1 = vx * d1x + vy * d1y, dot2 = vx * d2x + vy * d2y;
            const bool ok = loc == LOC_LEFT ? dot2 < 0 : loc == LOC_RIGHT ? dot2 > 0 : loc == LOC_FRONT ? dot1 > 0 : dot1 < 0;
            if (!ok) continue;
        }
        m |= 1u << k;
    }
    return m;
}

// ---- draw source of one lane: a ring of the stream's next 64 draws ---------------------------------------
// The generator below never asks for randomness through a branchy "refill if needed" path: the draws of the
// lane's stream sit in a 64-word ring (16 Philox blocks; word d & 63 holds draw d) that is topped up for ALL lanes
// of a warp together (k_gen_small: converged Philox, the expensive part of generation), and a draw is one indexed
// read.  Memory with dynamic indexing: shared memory on the device (stride `ws` words between a lane's
// consecutive entries), a plain array in the host build.
template <int RING_BLOCKS>
struct DrawRingT {
    static constexpr int RING_WORDS = 4 * RING_BLOCKS;
    uint32_t *w; int ws;
    uint32_t k0, k1;
    uint64_t draws;               // index of the next draw of the stream
    uint64_t gen;                 // blocks gen - RING_BLOCKS .. gen - 1 are in the ring

    BB_HD void init(uint32_t *w_, int ws_, uint64_t seed, uint64_t d)
    {
        w = w_; ws = ws_; k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32); draws = d; gen = d >> 2;
    }
    BB_HD int avail() const { return (int)((int64_t)(gen * 4ull) - (int64_t)draws); }     // draws ready to be read
    // the slot of block `gen` holds block gen - RING_BLOCKS: free once every draw of that block is consumed
    BB_HD bool room() const { return gen < (draws >> 2) + (uint64_t)RING_BLOCKS; }
    BB_HD void gen_block()
    {
        RngScalar r; r.k0 = k0; r.k1 = k1;
        r.refill(gen);
        const int s = (int)(gen & (uint64_t)(RING_BLOCKS - 1)) * 4;
        w[s * ws] = r.b0; w[(s + 1) * ws] = r.b1; w[(s + 2) * ws] = r.b2; w[(s + 3) * ws] = r.b3;
        gen++;
    }
    BB_HD void fill() { while (room()) gen_block(); }                                      // scalar top-up (host, rare device path)
    BB_HD uint32_t peek(int j) const { return w[(int)((draws + (uint64_t)j) & (uint64_t)(RING_WORDS - 1)) * ws]; }   // needs avail() > j
    BB_HD void advance(int n) { draws += (uint64_t)n; }
};
typedef DrawRingT<16> DrawRing;   // k_gen_small, host build; the generator warp inside k_rollout uses 8 blocks
constexpr int RING_LOW = 8;       // every generator step below reads at most 5 draws: top up when fewer than 8 are ready

// ---- one ATTEMPT at a small level, in phases that a warp runs in lock-step ---------------------------------
//   small_attempt_begin     RoomGrid._gen_grid of the one room
//   small_place_try         one placement try of whatever the lane is placing (agent or next object): the
//                           bulk of the work, one shared code path
//   small_flood_sweep/_ok   check_objs_reachable as a bitboard flood fill
//   small_pick / small_desc_try   the instruction's object descriptor
// generate_small() below drives them for one lane (host build); k_gen_small drives them for 32 lanes with
// warp-level loops: every lane of the warp is in the same phase, lanes that are done with a phase wait.
enum : int { ST_OBJ = 0, ST_AGENT, ST_PLACED, ST_FAIL, ST_IDLE };
struct SmallAttempt {
    SmallLevel L;
    uint64_t occ, fill;
    int stage, k, tries, cur_tc;
    bool agent_placed;
};

template <class DS>
BB_HD void small_attempt_begin(const LevelParams &lp, SmallAttempt &a, DS &ds)
{
    const bool levelgen = lp.kind == KIND_LEVELGEN;
    const int S = lp.room_size;
    a.occ = lp.wall64; a.fill = 0; a.k = 0; a.tries = 0; a.cur_tc = 0;
    a.L.poss = 0; a.L.tcs = 0; a.L.nobj = 0;
    a.L.ax = S / 2; a.L.ay = S / 2; a.L.adir = 0; a.agent_placed = true;        // RoomGrid._gen_grid: agent in the middle, facing right
    a.L.leaf_kind = lp.kind == KIND_OBJ ? lp.instr : (levelgen ? lp.action_kinds[0] : I_GOTO);
    a.L.d_type = ANY_TYPE; a.L.d_color = ANY; a.L.d_loc = LOC_NONE; a.L.d_mask = 0;
    if (levelgen) { ds.advance(1); a.stage = lp.num_dists > 0 ? ST_OBJ : ST_AGENT; }   // `_rand_float(0,1) < 0`: one draw; distractors first
    else a.stage = ST_AGENT;                                                             // iclr19 levels: agent first
}

// RoomGrid.place_agent -> MiniGridEnv.place_agent tries / add_object, add_distractors: one placement try
template <class DS>
BB_HD void small_place_try(const LevelParams &lp, SmallAttempt &a, DS &ds)
{
    const int S = lp.room_size;
    const bool levelgen = lp.kind == KIND_LEVELGEN;
    const int nplace = lp.num_dists + (lp.kind == KIND_REDBALL ? 1 : 0);     // objects placed per attempt
    SmallLevel &L = a.L;
    if (a.tries > 1000) { a.stage = ST_FAIL; return; }                       // place_obj: RecursionError
    const bool isobj = a.stage == ST_OBJ;
    const bool fixed = lp.kind == KIND_REDBALL && a.k == 0;                  // the red ball: no colour / type draws
    int used = 0;
    if (isobj && a.tries == 0) {
        if (fixed) a.cur_tc = T_BALL | (C_RED << 3);
        else {
            const int color = color_by_name_rank((int)mulhi32(ds.peek(0), 6u));
            const int t = (int)mulhi32(ds.peek(1), 3u);
            a.cur_tc = (t == 0 ? T_KEY : t == 1 ? T_BALL : T_BOX) | (color << 3);
            used = 2;
        }
    }
    a.tries++;
    const int x = (int)mulhi32(ds.peek(used), (uint32_t)S), y = (int)mulhi32(ds.peek(used + 1), (uint32_t)S);
    used += 2;
    bool ok = !((a.occ >> (8 * y + x)) & 1u);
    if (isobj) ok = ok && !(a.agent_placed && x == L.ax && y == L.ay) && (iabs(L.ax - x) + iabs(L.ay - y) >= 2);   // reject_next_to
    if (!ok) { ds.advance(used); return; }
    a.tries = 0;
    if (isobj) {
        a.occ |= 1ull << (8 * y + x);
        L.poss |= (uint64_t)(x | (y << 3)) << (6 * a.k);
        L.tcs |= (uint64_t)a.cur_tc << (6 * a.k);
        a.k++; L.nobj = a.k;
        if (a.k == nplace) {
            if (lp.kind == KIND_REDBALL && lp.grey_dists)                  // GoToRedBallGrey: distractors turn grey
                for (int q = 1; q < a.k; q++) L.tcs = (L.tcs & ~(56ull << (6 * q))) | ((uint64_t)(C_GREY << 3) << (6 * q));
            if (levelgen) { a.stage = ST_AGENT; a.agent_placed = false; }   // MiniGridEnv.place_agent: agent_pos = None
            else a.stage = ST_PLACED;
        }
    } else {
        L.ax = x; L.ay = y; a.agent_placed = true;
        L.adir = (int)mulhi32(ds.peek(used), 4u);
        used++;
        const int fb = 8 * (y + dir_dy(L.adir)) + x + dir_dx(L.adir);
        const bool front_ok = !((a.occ >> fb) & 1u) || ((lp.wall64 >> fb) & 1u);
        if (front_ok) a.stage = (!levelgen && nplace > 0) ? ST_OBJ : ST_PLACED;
    }
    ds.advance(used);
}

// check_objs_reachable: the fill starts on the agent; two dilations per call; returns whether anything changed
BB_HD void small_flood_begin(SmallAttempt &a) { a.fill = 1ull << (8 * a.L.ay + a.L.ax); }
BB_HD bool small_flood_sweep(SmallAttempt &a)
{
    const uint64_t pass = ~a.occ;
    uint64_t f1 = a.fill | ((a.fill << 1 | a.fill >> 1 | a.fill << 8 | a.fill >> 8) & pass);
    f1 |= (f1 << 1 | f1 >> 1 | f1 << 8 | f1 >> 8) & pass;
    const bool changed = f1 != a.fill;
    a.fill = f1;
    return changed;
}
BB_HD bool small_flood_ok(const LevelParams &lp, const SmallAttempt &a)     // every object touches the filled region
{
    const uint64_t near = a.fill | a.fill << 1 | a.fill >> 1 | a.fill << 8 | a.fill >> 8;
    const uint64_t things = a.occ & ~lp.wall64;
    return (things & ~near) == 0;
}
BB_HD bool small_needs_check(const LevelParams &lp) { return !(lp.kind == KIND_LEVELGEN && lp.unblocking); }

// iclr19 levels: GoToRedBall* describe object 0; the others `obj = self._rand_elem(objs)` (one object: no draw)
template <class DS>
BB_HD void small_pick(const LevelParams &lp, SmallAttempt &a, DS &ds)
{
    SmallLevel &L = a.L;
    int idx = 0;
    if (lp.kind == KIND_OBJ && lp.num_dists > 1) { idx = (int)mulhi32(ds.peek(0), (uint32_t)lp.num_dists); ds.advance(1); }
    const int tc = sm_obj_tc(L.tcs, idx);
    L.d_type = tc & 7; L.d_color = tc >> 3; L.d_loc = LOC_NONE;
    L.d_mask = sm_match(L, L.d_type, L.d_color, LOC_NONE);
}

// LevelGen.rand_obj, one try; returns true when the lane is finished with this phase (matched, or attempt failed)
template <class DS>
BB_HD bool small_desc_try(const LevelParams &lp, SmallAttempt &a, DS &ds)
{
    SmallLevel &L = a.L;
    if (a.tries > 100) { a.stage = ST_FAIL; return true; }                  // rand_obj: RecursionError
    a.tries++;
    const int ci = (int)mulhi32(ds.peek(0), 7u);
    const int color = ci == 0 ? ANY : color_by_name_rank(ci - 1);
    const int ntypes = L.leaf_kind == I_GOTO ? 4 : 3;
    const int ti = (int)mulhi32(ds.peek(1), (uint32_t)ntypes);
    const int type = ti == 0 ? T_BOX : ti == 1 ? T_BALL : ti == 2 ? T_KEY : T_DOOR;
    int used = 2, loc = LOC_NONE;
    if (lp.locations) {
        const bool with_loc = mulhi32(ds.peek(2), 2u) == 0;                // _rand_bool()
        used = 3;
        if (with_loc) { loc = (int)mulhi32(ds.peek(3), 4u); used = 4; }
    }
    ds.advance(used);
    const uint32_t m = sm_match(L, type, color, loc);
    if (m == 0) return false;
    L.d_type = type; L.d_color = color; L.d_loc = loc; L.d_mask = m;
    return true;
}

// Generates one level of a small single-room environment on one lane.  Returns the number of attempts.
BB_HD int generate_small(const LevelParams &lp, RngScalar &rng, SmallLevel &L)
{
    uint32_t words[DrawRing::RING_WORDS];
    DrawRing ds;
    ds.init(words, 1, ((uint64_t)rng.k1 << 32) | rng.k0, rng.draws);
    int attempts = 0;
    for (;;) {
        SmallAttempt a;
        attempts++;
        small_attempt_begin(lp, a, ds);
        while (a.stage == ST_OBJ || a.stage == ST_AGENT) {
            if (ds.avail() < RING_LOW) ds.fill();
            small_place_try(lp, a, ds);
        }
        if (a.stage == ST_FAIL) continue;
        if (small_needs_check(lp)) {
            small_flood_begin(a);
            while (small_flood_sweep(a)) {}
            if (!small_flood_ok(lp, a)) continue;
        }
        a.tries = 0;
        if (lp.kind == KIND_LEVELGEN) {
            for (;;) {
                if (ds.avail() < RING_LOW) ds.fill();
                if (small_desc_try(lp, a, ds)) break;
            }
            if (a.stage == ST_FAIL) continue;
        } else {
            if (ds.avail() < RING_LOW) ds.fill();
            small_pick(lp, a, ds);
        }
        L = a.L;
        break;
    }
    rng.draws = ds.draws; rng.blk = ~0ull;
    return attempts;
}

// SmallLevel -> the records of a level slot (grid in both orientations, object table, verifier, tokens, hot).
// Written without dynamically indexed local arrays: on the device everything stays in registers.
BB_HD void store16(void *dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d)      // dst is 16-byte aligned
{
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<uint4 *>(dst) = make_uint4(a, b, c, d);
#else
    const uint32_t v[4] = {a, b, c, d};
    __builtin_memcpy(dst, v, 16);
#endif
}
BB_HD void emit_small_level(const LevelParams &lp, const SmallLevel &L, const LevelOut &o)
{
    // the empty room in both orientations (row templates), then one byte per object and orientation on top
    // (same thread, same addresses: the later stores win)
    if (lp.rs_g == 8) {
#pragma unroll
        for (int r = 0; r < 8; r++) {                              // 8-byte rows: G rows 0..H-1, then GT rows 0..W-1 (gt_off = 8 H)
            const uint64_t t0 = lp.row_tmpl[2 * r], t1 = 2 * r + 1 < lp.H ? lp.row_tmpl[2 * r + 1] : 0ull;   // padding stays zero
            if (2 * r < lp.H) store16(o.grid + 16 * r, (uint32_t)t0, (uint32_t)(t0 >> 32), (uint32_t)t1, (uint32_t)(t1 >> 32));
        }
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const uint64_t t0 = lp.row_tmpl[8 + 2 * r], t1 = 2 * r + 1 < lp.W ? lp.row_tmpl[8 + 2 * r + 1] : 0ull;
            if (2 * r < lp.W) store16(o.grid + lp.gt_off + 16 * r, (uint32_t)t0, (uint32_t)(t0 >> 32), (uint32_t)t1, (uint32_t)(t1 >> 32));
        }
    } else {                                                       // 4-byte rows (rooms up to 4 x 4)
        store16(o.grid, (uint32_t)lp.row_tmpl[0], (uint32_t)lp.row_tmpl[1], (uint32_t)lp.row_tmpl[2], (uint32_t)lp.row_tmpl[3]);
        store16(o.grid + lp.gt_off, (uint32_t)lp.row_tmpl[8], (uint32_t)lp.row_tmpl[9], (uint32_t)lp.row_tmpl[10], (uint32_t)lp.row_tmpl[11]);
    }
    for (int k = 0; k < L.nobj; k++) {
        const int x = sm_obj_x(L.poss, k), y = sm_obj_y(L.poss, k), tc = sm_obj_tc(L.tcs, k);
        o.grid[y * lp.rs_g + x] = (uint8_t)tc;
        o.grid[lp.gt_off + x * lp.rs_t + y] = (uint8_t)tc;
    }
    // object table: 6-bit fields -> bytes, objects 0..11 (the rest of the 32 slots is zero)
    uint32_t X[3], Y[3], TC[3];
#pragma unroll
    for (int wd = 0; wd < 3; wd++) {
        X[wd] = 0; Y[wd] = 0; TC[wd] = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int k = 4 * wd + i;
            if (k >= 10) continue;                                   // poss / tcs hold at most 10 objects
            X[wd] |= (uint32_t)sm_obj_x(L.poss, k) << (8 * i);
            Y[wd] |= (uint32_t)sm_obj_y(L.poss, k) << (8 * i);
            TC[wd] |= (uint32_t)sm_obj_tc(L.tcs, k) << (8 * i);
        }
    }
    uint8_t *ob = reinterpret_cast<uint8_t *>(o.obj);
    store16(ob, X[0], X[1], X[2], 0); store16(ob + 16, 0, 0, 0, 0);
    store16(ob + 32, Y[0], Y[1], Y[2], 0); store16(ob + 48, 0, 0, 0, 0);
    store16(ob + 64, TC[0], TC[1], TC[2], 0); store16(ob + 80, 0, 0, 0, 0);
    // verifier record: one leaf, one descriptor
    uint8_t *ib = reinterpret_cast<uint8_t *>(o.ins);
    store16(ib, L.d_mask, 0, 0, 0); store16(ib + 16, 0, 0, 0, 0);
    store16(ib + 32, (uint32_t)L.leaf_kind | ((uint32_t)I_NONE << 8) | ((uint32_t)I_NONE << 16) | ((uint32_t)I_NONE << 24),
            (uint32_t)NO_OBJ * 0x01010101u, (uint32_t)R_SINGLE, 0);
    // hot record
    const uint32_t all = (1u << L.nobj) - 1u;
    store16(o.hot, (uint32_t)L.ax | ((uint32_t)L.ay << 8) | ((uint32_t)L.adir << 16) | ((uint32_t)NO_OBJ << 24),
            (uint32_t)(uint16_t)lp.nav_time_maze << 16, all, all);                       // step_count 0, max_steps: one navigation
    // "go to" / "pick up" + ObjDesc.surface: [v0 v1 article (colour) type loc...], the colour word is optional
    const uint64_t v0 = L.leaf_kind == I_GOTO ? W_GO : W_PICK, v1 = L.leaf_kind == I_GOTO ? W_TO : W_UP;
    const uint64_t art = popc32(L.d_mask) > 1 ? W_A : W_THE;
    const uint64_t ty = L.d_type == ANY_TYPE ? W_OBJECT : L.d_type == T_BOX ? W_BOX : L.d_type == T_BALL ? W_BALL : L.d_type == T_KEY ? W_KEY : W_DOOR;
    uint64_t loc = 0;                                              // up to four words, 16 bits each
    if (L.d_loc == LOC_FRONT) loc = (uint64_t)W_IN | ((uint64_t)W_FRONT << 16) | ((uint64_t)W_OF << 32) | ((uint64_t)W_YOU << 48);
    else if (L.d_loc == LOC_BEHIND) loc = (uint64_t)W_BEHIND | ((uint64_t)W_YOU << 16);
    else if (L.d_loc == LOC_LEFT) loc = (uint64_t)W_ON | ((uint64_t)W_YOUR << 16) | ((uint64_t)W_LEFT << 32);
    else if (L.d_loc == LOC_RIGHT) loc = (uint64_t)W_ON | ((uint64_t)W_YOUR << 16) | ((uint64_t)W_RIGHT << 32);
    uint64_t t0 = v0 | (v1 << 16) | (art << 32), t1, t2;
    if (L.d_color != ANY) { t0 |= (uint64_t)(W_RED + L.d_color) << 48; t1 = ty | (loc << 16); t2 = loc >> 48; }
    else { t0 |= ty << 48; t1 = loc; t2 = 0; }
    store16(o.tok, (uint32_t)t0, (uint32_t)(t0 >> 32), (uint32_t)t1, (uint32_t)(t1 >> 32));
    if (lp.max_tokens > 8) store16(o.tok + 8, (uint32_t)t2, (uint32_t)(t2 >> 32), 0, 0);
    for (int k = 16; k < lp.max_tokens; k++) o.tok[k] = 0;
}

// =============================================================================
// step: MiniGridEnv.step (App. A.4) + RoomGridLevel.step (levelgen.py:49-66)
// =============================================================================
// The step / verifier / observation code below is written against a small "environment memory"
// accessor M so that the same source runs on (a) plain pointers into the struct-of-arrays state
// (GlobalMem: host build, large grids) and (b) the warp's shared-memory staging area (pool.cu).
//   cell(x,y) set_cell(x,y,v) word_at(byte_offset)       grid (both orientations)
//   ox(k) oy(k) otc(k) set_oxy(k,x,y)                   object table
//   desc_mask(d) leaf_kind(l) leaf_pre(l) set_leaf_pre(l,v) root_kind() side_and() flags() set_flags(v)
BB_HD uint32_t load_u32_any(const uint8_t *p)                   // p is 4-byte aligned
{
#if defined(__CUDA_ARCH__)
    return *reinterpret_cast<const uint32_t *>(p);
#else
    uint32_t v; __builtin_memcpy(&v, p, 4); return v;
#endif
}
struct GlobalMem {
    const LevelParams &lp; uint8_t *grid; ObjTab *ot; InstrRec *ins;
    BB_HD GlobalMem(const LevelParams &lp_, uint8_t *g, ObjTab *o, InstrRec *i) : lp(lp_), grid(g), ot(o), ins(i) {}
    BB_HD int cell(int x, int y) const { return grid[y * lp.rs_g + x]; }
    BB_HD void set_cell(int x, int y, int v) { bb::set_cell(lp, grid, x, y, v); }
    // aligned 32-bit word at byte offset `off` of the env's grid bytes (G at 0, GT at gt_off)
    BB_HD uint32_t word_at(int off) const
    {
        const uint8_t *p = grid + off;
#if defined(__CUDA_ARCH__)
        return *reinterpret_cast<const uint32_t *>(p);
#else
        uint32_t v; __builtin_memcpy(&v, p, 4); return v;
#endif
    }
    BB_HD int ox(int k) const { return ot->x[k]; }
    BB_HD int oy(int k) const { return ot->y[k]; }
    BB_HD int otc(int k) const { return ot->tc[k]; }
    BB_HD uint32_t oxw(int i) const { return load_u32_any(ot->x + 4 * i); }      // x / y of objects 4i .. 4i+3, one byte each
    BB_HD uint32_t oyw(int i) const { return load_u32_any(ot->y + 4 * i); }
    BB_HD void set_oxy(int k, int x, int y) { ot->x[k] = (uint8_t)x; ot->y[k] = (uint8_t)y; }
    BB_HD uint32_t desc_mask(int d) const { return ins->desc_mask[d]; }
    BB_HD int leaf_kind(int l) const { return ins->leaf_kind[l]; }
    BB_HD int leaf_pre(int l) const { return ins->leaf_pre[l]; }
    BB_HD void set_leaf_pre(int l, int v) { ins->leaf_pre[l] = (uint8_t)v; }
    BB_HD int root_kind() const { return ins->root_kind; }
    BB_HD int side_and() const { return ins->side_and; }
    BB_HD void set_side_and(int v) { ins->side_and = (uint8_t)v; }
    BB_HD int flags() const { return ins->flags; }
    BB_HD void set_flags(int v) { ins->flags = (uint8_t)v; }
    BB_HD int start_carry() const { return ins->pad0; }       // object id + 1 the agent holds from its first step on (PutNext*Carrying), 0 = none
};

// The objects whose TABLE position is (x, y), as a set mask: SWAR over the packed x / y byte arrays, four objects per
// word -- no loop over set bits, no divergence (round 1 walked the masks object by object: with the seven actions and the
// verifier tree in front of them those loops ran at 1.5-3 active lanes, ncu r02b).  Positions are < 32, so a byte of
// (x ^ X) | (y ^ Y) is < 0x40 and "+ 0x7F" raises bit 7 exactly when it is non-zero.  Unused table entries sit at (0, 0),
// an outer-wall corner that is never a front cell, and are excluded by every mask this is intersected with anyway.
template <class M>
BB_HD uint32_t objs_at(const M &mem, int x, int y)
{
    const uint32_t bx = (uint32_t)x * 0x01010101u, by = (uint32_t)y * 0x01010101u;
    uint32_t at = 0;
    for (int i = 0; i < mem.lp.obj_words; i++) {
        const uint32_t d = (mem.oxw(i) ^ bx) | (mem.oyw(i) ^ by);
        const uint32_t z = ((d + 0x7F7F7F7Fu) & 0x80808080u) ^ 0x80808080u;       // 0x80 in the bytes that match
        at |= ((((z >> 7) * 0x01020408u) >> 24) & 0xFu) << (4 * i);
    }
    return at;
}

struct StepCtx {            // what a leaf verifier looks at after the action was applied
    int action, fx, fy;     // front_pos AFTER the move
    int carry;              // env.carrying after the action
    uint32_t cur_mask, snap_mask;
    uint32_t at;            // objects whose table position is front_pos (objs_at)
    int fcell;              // the cell at front_pos after the action
};

// ---- the verifier.  Results are tri-state like the reference's 'continue' / 'success' / 'failure' strings ----
enum : int { V_CONT = 0, V_SUCC = 1, V_FAIL = 2 };
// LevelParams::strict_mask: bit l = leaf l was built with strict=True (OpenInstr / PickupInstr / PutNextInstr of the "Debug"
// bonus levels), bit 4 = the root Before / After is strict.  LevelParams::done_actions = the reference's BABYAI_DONE_ACTIONS
// mode (verifier.py:15-17): an ActionInstr only reports through the `done` action (lastStepMatch).

// ActionInstr.verify_action: Open :257-274, GoTo :296-303, Pickup :330-350, PutNext :393-417.
// Written branch-light: the four kinds are evaluated as predicates and selected by kind, so that the lanes of a warp
// (one env each, all with different instructions) run ONE instruction stream (round 1's if-chain per kind, nested in the
// side / root recursion, ran at 3.6 active lanes: ncu r02c).  Only PutNext's neighbourhood test is a loop, and it runs only
// on the step where a matching object was just dropped.
template <int KM = 0, class M>                    // KM != 0: the family's kinds_mask as a compile-time constant (mem_spec)
BB_HD int verify_action(M &mem, int leaf, const StepCtx &s)
{
    const int kind = mem.leaf_kind(leaf);
    const uint32_t set = mem.desc_mask(2 * leaf);
    const int km = KM ? KM : mem.lp.kinds_mask;    // (kernel argument: the branches on it are uniform)
    const bool goto_ok = (set & s.snap_mask & s.at) != 0;                             // some pos in obj_poss is front_pos
    if (km == (1 << I_GOTO)) return goto_ok ? V_SUCC : V_CONT;                        // GoTo-only families (GoToLocal, GoToRedBall, GoTo, ...)
    const bool strict = ((mem.lp.strict_mask >> leaf) & 1) != 0;
    int pre = NO_OBJ;
    if (km & ((1 << I_PICKUP) | (1 << I_PUTNEXT))) {
        pre = mem.leaf_pre(leaf);
        if (kind == I_PICKUP || kind == I_PUTNEXT) mem.set_leaf_pre(leaf, s.carry);   // preCarrying, refreshed on every evaluation
    }
    // the toggled cell must be a door of the set, and open: at most one object is ON a cell
    bool toggled_door = false, open_ok = false;
    if (km & (1 << I_OPEN)) {
        toggled_door = s.action == A_TOGGLE && (s.fcell & 7) == T_DOOR;
        open_ok = toggled_door && (s.fcell >> 6) == 0 && (set & s.cur_mask & s.at) != 0;
    }
    const bool picked = s.action == A_PICKUP && s.carry != NO_OBJ;                    // (carrying something after a pickup action)
    const bool pick_ok = picked && pre == NO_OBJ && s.carry < MAXOBJ && ((set >> (s.carry & 31)) & 1u);   // (untracked objects are in no set)
    bool put_ok = false;
    if ((km & (1 << I_PUTNEXT)) && kind == I_PUTNEXT && s.action == A_DROP && pre < MAXOBJ && ((set >> (pre & 31)) & 1u) && s.carry != pre) {
        const int ax = mem.ox(pre), ay = mem.oy(pre);      // (a carried object that was not dropped: cur_pos == (-1, -1), excluded above)
        for (uint32_t m = mem.desc_mask(2 * leaf + 1) & s.snap_mask; m; m &= m - 1) {
            const int k = ffs32(m);
            if (iabs(ax - mem.ox(k)) + iabs(ay - mem.oy(k)) == 1) put_ok = true;
        }
    }
    const bool ok = kind == I_GOTO ? goto_ok : kind == I_OPEN ? open_ok : kind == I_PICKUP ? pick_ok : put_ok;
    // strict mode (verifier.py:269-272, 343-346, 398-401): the wrong door toggled / any object picked up is a failure;
    // PutNext tests it BEFORE it looks at the drop, Open and Pickup after their success test
    const bool bad = kind == I_OPEN ? toggled_door : kind == I_GOTO ? false : picked;
    if (strict && kind == I_PUTNEXT && bad) return V_FAIL;
    if (ok) return V_SUCC;
    return strict && bad ? V_FAIL : V_CONT;
}

// ActionInstr.verify :211-231: with done actions the instruction only reports when the agent says `done`
template <class M>
BB_HD int verify_leaf(M &mem, int leaf, const StepCtx &s)
{
    if (!mem.lp.done_actions) return verify_action(mem, leaf, s);
    const int bit = 0x10 << leaf;                              // lastStepMatch of leaf l: bit 4 + l of the side_and byte
    if (s.action == A_DONE) return (mem.side_and() & bit) ? V_SUCC : V_FAIL;
    const int res = verify_action(mem, leaf, s);
    const int sa = mem.side_and();
    const int ns = res == V_SUCC ? (sa | bit) : (sa & ~bit);
    if (ns != sa) mem.set_side_and(ns);
    return V_CONT;                                             // (the reference falls off the end: None)
}

// one side: an ActionInstr, or AndInstr.verify (verifier.py:536-550): each unfinished half is evaluated every step; a
// half's 'failure' is not passed on (the `action is done` test of :544 never holds for the integer actions a vectorised
// env is stepped with)
template <class M>
BB_HD int verify_side(M &mem, int side, const StepCtx &s)
{
    const bool is_and = ((mem.side_and() >> side) & 1) != 0;
    const int ba = 2 + 2 * side, bb_ = 3 + 2 * side;
    int fl = mem.flags();
    const bool a_latched = is_and && ((fl >> ba) & 1);
    const bool b_latched = ((fl >> bb_) & 1) != 0;
    int r0 = V_SUCC;
    if (!a_latched) r0 = verify_leaf(mem, 2 * side, s);
    if (!is_and) return r0;
    if (!a_latched && r0 == V_SUCC) fl |= 1 << ba;
    if (!b_latched && verify_leaf(mem, 2 * side + 1, s) == V_SUCC) fl |= 1 << bb_;
    if (fl != mem.flags()) mem.set_flags(fl);
    return (((fl >> ba) & 1) && ((fl >> bb_) & 1)) ? V_SUCC : V_CONT;
}

// BeforeInstr.verify :449-471, AfterInstr.verify :490-512 (hand-over re-verifies the same action)
template <class M>
BB_HD int verify_root(M &mem, const StepCtx &s)
{
    if (mem.lp.single_instr) return verify_leaf(mem, 0, s);      // families that only build one ActionInstr (uniform branch)
    const int rk = mem.root_kind();
    const int first = rk == R_AFTER ? 1 : 0, second = 1 - first;
    const bool first_done = rk != R_SINGLE && ((mem.flags() >> first) & 1);
    int r = V_SUCC;
    if (!first_done) r = verify_side(mem, first, s);
    if (rk == R_SINGLE) return r;
    if (r == V_FAIL) return V_FAIL;
    if (r == V_CONT) {
        // strict Before / After (:466-469, 507-509): finishing the other instruction first is a failure
        if ((mem.lp.strict_mask & 0x10) && verify_side(mem, second, s) == V_SUCC) return V_FAIL;
        return V_CONT;
    }
    if (!first_done) mem.set_flags(mem.flags() | (1 << first));
    return verify_side(mem, second, s);
}

// Compile-time specialisation of the stepping code by level family, carried by the accessor type: an accessor that declares
// `static constexpr int spec_room_kinds = K` (K != 0) promises  num_rows == num_cols == 1, kinds_mask == K (one instruction
// kind: GoTo or Pickup), single_instr, no strict / done-action mode, no bonus family, no untracked objects  -- GoToRedBall*,
// GoToObj*, GoToLocal* (K = GoTo: BASELINE configs 1 and 2) and PickupLoc (K = Pickup: config 3).
// The generic code is the same with those tests as uniform run-time branches; the specialised instantiations are there for the
// INSTRUCTION FOOTPRINT of the persistent kernel's loop (k_rollout: 7 616 instructions against a 32 KB = 2 048-instruction
// L1.5 instruction cache; `no_instruction` was 17 % of its issue-stall cycles, ncu r02c; 0.30 of 7.2 cycles after, r02k).
template <class M, class = void> struct mem_spec { static constexpr int room_kinds = 0; };
template <class M> struct mem_spec<M, decltype((void)M::spec_room_kinds)> { static constexpr int room_kinds = M::spec_room_kinds; };
BB_HD int level_spec_room_kinds(const LevelParams &lp)          // the K of the instantiation a level qualifies for, or 0
{
    const bool plain = lp.num_rows == 1 && lp.num_cols == 1 && lp.single_instr && !lp.done_actions && lp.strict_mask == 0 &&
                       lp.bonus == 0 && lp.kind != KIND_UNLOCK && lp.kind != KIND_BONUS;
    if (!plain) return 0;
    return lp.kinds_mask == (1 << I_GOTO) || lp.kinds_mask == (1 << I_PICKUP) ? lp.kinds_mask : 0;
}

struct StepResult { bool done; bool success; float reward; };      // done: success, failure (strict / done-action modes) or time-out

// Applies one action to the live state of one env.  `h` is the env's hot record
// held in registers by the caller (written back by the caller).
// UNTR: the pool serves KIND_UNLOCK, whose non-door objects are untracked (CARRY_UNTRACKED)
template <bool UNTR = false, class M>
BB_HD StepResult step_env(EnvHot &h, M &mem, int action)
{
    if (mem_spec<M>::room_kinds == 0 && h.step_count == 0 && mem.lp.bonus == BN_PUTNEXT) {
        // Level_PutNext*Carrying (bonus_levels.py:821-829): reset() returns the observation of the generated level, THEN takes
        // obj_a off the grid into the agent's hands -- so the first step acts on the modified state
        const int sc = mem.start_carry();
        if (sc) {
            mem.set_cell(mem.ox(sc - 1), mem.oy(sc - 1), CELL_EMPTY);
            h.cur_mask &= ~(1u << (sc - 1));
            h.carry = (uint8_t)(sc - 1);
        }
    }
    int x = h.x, y = h.y, dir = h.dirflags & 3, carry = h.carry;
    const int fx = x + dir_dx(dir), fy = y + dir_dy(dir);
    const int fc = mem.cell(fx, fy);
    const int ftype = fc & 7;
    // the three actions that change the pose first: the other four act on the cell in front, which for them is also
    // front_pos AFTER the action -- the position the verifier looks at.  One objs_at() of that cell serves both.
    if (action == A_LEFT) dir = (dir + 3) & 3;
    else if (action == A_RIGHT) dir = (dir + 1) & 3;
    else if (action == A_FORWARD && (fc == CELL_EMPTY || (ftype == T_DOOR && (fc >> 6) == 0))) { x = fx; y = fy; }
    const int nfx = x + dir_dx(dir), nfy = y + dir_dy(dir);
    uint32_t at = objs_at(mem, nfx, nfy);
    if (action == A_PICKUP) {
        if (ftype >= T_KEY && carry == NO_OBJ) {
            const uint32_t m = at & h.cur_mask;
            if (m) { carry = ffs32(m); h.cur_mask &= ~(1u << carry); mem.set_cell(fx, fy, CELL_EMPTY); }
            else if constexpr (UNTR) { carry = CARRY_UNTRACKED | (fc & 0x3F); mem.set_cell(fx, fy, CELL_EMPTY); }
        }
    } else if (action == A_DROP) {
        if constexpr (UNTR) {
            if (fc == CELL_EMPTY && carry != NO_OBJ) {
                if (carry & CARRY_UNTRACKED) mem.set_cell(fx, fy, carry & 0x3F);
                else {
                    mem.set_cell(fx, fy, mem.otc(carry));
                    mem.set_oxy(carry, fx, fy);
                    h.cur_mask |= 1u << carry;
                    at |= 1u << carry;                  // its table position is the front cell now
                }
                carry = NO_OBJ;
            }
        } else if (fc == CELL_EMPTY && carry != NO_OBJ) {
            mem.set_cell(fx, fy, mem.otc(carry));
            mem.set_oxy(carry, fx, fy);
            h.cur_mask |= 1u << carry;
            at |= 1u << carry;
            carry = NO_OBJ;
        }
    } else if (action == A_TOGGLE) {
        if (ftype == T_DOOR) {
            int st = fc >> 6, ns = st;
            if (st == 2) {           // locked: needs a carried key of the door's colour; key stays in hand
                if constexpr (UNTR) {
                    if (carry != NO_OBJ) {
                        const int ctc = (carry & CARRY_UNTRACKED) ? (carry & 0x3F) : mem.otc(carry);
                        if ((ctc & 7) == T_KEY && (ctc >> 3) == ((fc >> 3) & 7)) ns = 0;
                    }
                } else if (carry != NO_OBJ && (mem.otc(carry) & 7) == T_KEY && (mem.otc(carry) >> 3) == ((fc >> 3) & 7)) ns = 0;
            } else ns = st ^ 1;
            if (ns != st) mem.set_cell(fx, fy, (fc & 0x3F) | (ns << 6));
        } else if (ftype == T_BOX) {  // Box.toggle: replaced by its contents (None, or Level_KeyInBox's key)
            const uint32_t m = at & h.cur_mask;
            int inside = -1;
            if (m) {
                const int id = ffs32(m);
                h.cur_mask &= ~(1u << id);
                if (mem.lp.box_contains == id + 1) inside = id + 1;
            }
            if (inside >= 0) {
                mem.set_cell(fx, fy, mem.otc(inside));
                mem.set_oxy(inside, fx, fy);
                h.cur_mask |= 1u << inside;
                at |= 1u << inside;
            } else mem.set_cell(fx, fy, CELL_EMPTY);
        }
    }
    h.step_count = (uint16_t)(h.step_count + 1);
    StepResult r;
    r.done = h.step_count >= h.max_steps;
    // RoomGridLevel.step: any drop action refreshes obj_poss (levelgen.py:53-54)
    if (action == A_DROP) h.snap_mask = h.cur_mask;
    h.x = (uint8_t)x; h.y = (uint8_t)y; h.dirflags = (uint8_t)((h.dirflags & ~3) | dir); h.carry = (uint8_t)carry;
    StepCtx s;
    s.action = action; s.fx = nfx; s.fy = nfy; s.carry = carry;
    s.cur_mask = h.cur_mask; s.snap_mask = h.snap_mask; s.at = at;
    int status;
    if constexpr (mem_spec<M>::room_kinds != 0) {
        s.fcell = 0;
        status = verify_action<mem_spec<M>::room_kinds>(mem, 0, s);        // the one leaf, its kind known at compile time
    } else {
        s.fcell = (mem.lp.kinds_mask & (1 << I_OPEN)) ? mem.cell(nfx, nfy) : 0;
        status = verify_root(mem, s);
    }
    r.success = status == V_SUCC;
    r.reward = 0.0f;
    if (status == V_FAIL) r.done = true;                  // RoomGridLevel.step: 'failure' ends the episode with reward 0
    if (r.success) {
        r.done = true;
        // _reward(): 1 - 0.9 * (step_count / max_steps) in float64, no fused multiply-add
        double q = (double)h.step_count / (double)h.max_steps;
#if defined(__CUDA_ARCH__)
        r.reward = (float)__dsub_rn(1.0, __dmul_rn(0.9, q));
#else
        volatile double p = 0.9 * q;
        r.reward = (float)(1.0 - p);
#endif
    }
    return r;
}

// =============================================================================
// observation: gen_obs_grid + process_vis + encode (App. A.5)
// =============================================================================
// The 7x7 view is handled as seven "view columns" (fixed lateral index vi, the
// seven depths vj as bytes), because that is the order of the output bytes
// (image[vi][vj][c]).  A view column is a 7-byte window of ONE stored grid row:
// of G (row-major) when the agent faces left/right, of GT (column-major) when it
// faces up/down -- three aligned 32-bit loads + a funnel shift, byte-reversed for
// two of the four headings.  Everything after that is SWAR on packed bytes:
// see-through flags, 8x8 bit transposes between column and row domain, the
// row-wise visibility propagation with a carry trick, byte masks, and the
// (type, color, state) expansion with byte permutes.
BB_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, int s)     // low 32 bits of (hi:lo) >> s, s in 0..31
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, s);
#else
    return s == 0 ? lo : (lo >> s) | (hi << (32 - s));
#endif
}
BB_HD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t sel)   // PRMT, selector nibbles 0..7 only
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
#endif
}
BB_HD uint32_t rev7(uint32_t x)                                  // reverse the low 7 bits
{
#if defined(__CUDA_ARCH__)
    return __brev(x & 0x7Fu) >> 25;
#else
    uint32_t r = 0;
    for (int i = 0; i < 7; i++) r |= ((x >> i) & 1u) << (6 - i);
    return r;
#endif
}
BB_HD uint32_t load_u32(const uint8_t *p)                        // p is 4-byte aligned
{
#if defined(__CUDA_ARCH__)
    return *reinterpret_cast<const uint32_t *>(p);
#else
    uint32_t v;
    __builtin_memcpy(&v, p, 4);
    return v;
#endif
}

// 0x80 in every byte of x (all bytes <= 0x7F) that is non-zero
BB_HD uint32_t nz80(uint32_t x) { return (x + 0x7F7F7F7Fu) & 0x80808080u; }
// see-through flags (0x80 per byte) of four packed cells: opaque = wall, or door that is not open
BB_HD uint32_t see80(uint32_t c)
{
    const uint32_t t = c & 0x07070707u;
    const uint32_t notwall = nz80(t ^ 0x02020202u);
    const uint32_t notdoor = nz80(t ^ 0x04040404u);
    const uint32_t closed = nz80((c >> 6) & 0x03030303u);
    return notwall & (notdoor | ~closed);
}
// the four 0x80 flags of a word -> bits 0..3 (partial products land on distinct bits: no carries)
BB_HD uint32_t gather4(uint32_t f80) { return (((f80 >> 7) * 0x01020408u) >> 24) & 0xFu; }
// bits 0..3 -> 0xFF in bytes 0..3
BB_HD uint32_t expand4(uint32_t bits) { return (((bits & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu; }

// 8x8 bit-matrix transpose, bit (i, j) at position 8 i + j of (hi:lo)
BB_HD void transpose8(uint32_t &lo, uint32_t &hi)
{
    uint64_t x = ((uint64_t)hi << 32) | lo, t;
    t = (x ^ (x >> 7)) & 0x00AA00AA00AA00AAull; x = x ^ t ^ (t << 7);
    t = (x ^ (x >> 14)) & 0x0000CCCC0000CCCCull; x = x ^ t ^ (t << 14);
    t = (x ^ (x >> 28)) & 0x00000000F0F0F0F0ull; x = x ^ t ^ (t << 28);
    lo = (uint32_t)x; hi = (uint32_t)(x >> 32);
}

// Visibility of the 7x7 view as 7-bit row masks (bit i = lateral column vi);
// see[j] = see-through cells of view row j, row 6 is the agent's.  Equivalent to the
// reference's nested process_vis loops (exhaustive test: test_vis_rows_bit_trick_equals_literal_loops).
// Row rule: the visible see-through cells flood sideways through see-through cells,
// and the cell just beyond each end of a flooded run is visible too; that same set
// (run + its two neighbours) is what the next row starts from.  The flood towards
// higher bits is one addition: carries ripple through the run of ones in `s` and stop
// on (and set) the first blocked cell; the other direction is the same on reversed bits.
BB_HD uint32_t both7(uint32_t x)            // the low 7 bits of x, and the same 7 bits reversed at bits 8..14
{
#if defined(__CUDA_ARCH__)
    return (x & 0x7Fu) | (__brev(x & 0x7Fu) >> 17);
#else
    return (x & 0x7Fu) | (rev7(x) << 8);
#endif
}
BB_HD uint32_t swap7(uint32_t u)            // exchanges the two 7-bit fields of a both7() word, reversing each
{
#if defined(__CUDA_ARCH__)
    return __brev(u) >> 17;                 // bit i -> 14 - i
#else
    return rev7(u >> 8) | (rev7(u) << 8);
#endif
}
// Both flood directions in ONE addition: a row mask is carried together with its bit-reversed copy (both7), so the carry
// chain that floods towards higher bits in the low field floods towards lower bits of the row in the high field; bit 7
// between the fields absorbs the low field's carry-out (the exhaustive test covers all 2^14 (visible, see-through) pairs
// of every row).
BB_HD void vis_rows(const uint32_t see[7], uint32_t vis[7])
{
    uint32_t s2[7];
#pragma unroll
    for (int j = 0; j < 7; j++) s2[j] = both7(see[j]);
    uint32_t v2 = both7(1u << 3);               // agent cell (3, 6)
#pragma unroll
    for (int j = 6; j >= 0; j--) {
        const uint32_t a2 = v2 & s2[j];         // visible cells that let light through (both orientations)
        const uint32_t u = (((s2[j] + a2) ^ s2[j]) | a2) & 0x7F7Fu;     // low field: flood up; high field: flood down, reversed
        const uint32_t d2 = (u | swap7(u)) & 0x7F7Fu;                   // the union, again in both orientations
        vis[j] = (v2 | d2) & 0x7Fu;
        v2 = d2;                                // what the row above starts from
    }
}

// ---- view columns -------------------------------------------------------------------------------
// Geometry of the view for one agent pose: which stored row holds view column vi and where its
// 7-cell window starts.
//   dir 3 (up):    GT row ax-3+vi, window y = ay-6 .. ay     (byte vj <-> y = ay-6+vj)
//   dir 1 (down):  GT row ax+3-vi, window y = ay .. ay+6     reversed (vj <-> y = ay+6-vj)
//   dir 0 (right): G  row ay-3+vi, window x = ax .. ax+6     reversed (vj <-> x = ax+6-vj)
//   dir 2 (left):  G  row ay+3-vi, window x = ax-6 .. ax     (vj <-> x = ax-6+vj)
struct ViewGeom {
    bool vert; int nrows, c_row, rstep, k0, sh; bool ok0, ok1, ok2; uint32_t sel_lo, sel_hi;
    int off0, dstep;            // byte offset (within the env's grid bytes) of window word 0 of column vi: off0 + vi * dstep
};
BB_HD ViewGeom view_geom(const LevelParams &lp, int ax, int ay, int dir)
{
    ViewGeom v;
    v.vert = (dir & 1) != 0;
    const int rs = v.vert ? lp.rs_t : lp.rs_g;
    v.nrows = v.vert ? lp.W : lp.H;
    v.c_row = v.vert ? ax : ay;
    const int c_win = v.vert ? ay : ax;
    v.rstep = (dir == 3 || dir == 0) ? 1 : -1;
    const bool rev = (dir == 1 || dir == 0);
    const int s0 = rev ? c_win : c_win - 6;
    v.k0 = s0 >> 2;                              // arithmetic shift: floor for negative starts
    v.sh = (s0 & 3) * 8;
    const int nwords = rs >> 2;
    v.ok0 = v.k0 >= 0 && v.k0 < nwords; v.ok1 = v.k0 + 1 >= 0 && v.k0 + 1 < nwords; v.ok2 = v.k0 + 2 >= 0 && v.k0 + 2 < nwords;
    v.sel_lo = rev ? 0x3456u : 0x3210u; v.sel_hi = rev ? 0x7012u : 0x7654u;
    v.dstep = v.rstep * rs;
    v.off0 = (v.vert ? lp.gt_off : 0) + (v.c_row - 3 * v.rstep) * rs + 4 * v.k0;
    return v;
}
// cells of view column vi: lo = depths vj 0..3, hi = vj 4..6 (+ one unused byte)
template <class M>
BB_HD void col_load(const M &mem, const ViewGeom &v, int vi, uint32_t &lo, uint32_t &hi)
{
    const uint32_t WALLW = 0x2A2A2A2Au;
    const int row = v.c_row + v.rstep * (vi - 3);
    const bool rok = row >= 0 && row < v.nrows;
    const int off = v.off0 + vi * v.dstep;
    const uint32_t w0 = (rok && v.ok0) ? mem.word_at(off) : WALLW;          // slice(): out of bounds -> Wall()
    const uint32_t w1 = (rok && v.ok1) ? mem.word_at(off + 4) : WALLW;
    const uint32_t w2 = (rok && v.ok2) ? mem.word_at(off + 8) : WALLW;
    const uint32_t a = funnel_r(w0, w1, v.sh), b = funnel_r(w1, w2, v.sh);
    lo = byte_perm(a, b, v.sel_lo);
    hi = byte_perm(a, b, v.sel_hi);
}
// see-through cells of a column: bit vj
BB_HD uint32_t col_see(uint32_t lo, uint32_t hi) { return (gather4(see80(lo)) | (gather4(see80(hi)) << 4)) & 0x7Fu; }
// four cells -> 12 observation bytes (type, color, state each)
BB_HD void encode4(uint32_t c, uint32_t &o0, uint32_t &o1, uint32_t &o2)
{
    const uint32_t T = c & 0x07070707u, K = (c >> 3) & 0x07070707u, S = (c >> 6) & 0x03030303u;
    o0 = byte_perm(byte_perm(T, K, 0x1040u), S, 0x3410u);      // t0 k0 s0 t1
    o1 = byte_perm(byte_perm(T, K, 0x6205u), S, 0x3250u);      // k1 s1 t2 k2
    o2 = byte_perm(byte_perm(T, K, 0x0730u), S, 0x7216u);      // s2 t3 k3 s3
}
// one view column -> its 21 output bytes (6 words, bytes 21..23 zero); cv = visibility of the column, bit vj
BB_HD void col_encode(uint32_t lo, uint32_t hi, uint32_t cv, uint32_t out[6])
{
    lo &= expand4(cv);
    hi &= expand4(cv >> 4) & 0x00FFFFFFu;
    encode4(lo, out[0], out[1], out[2]);
    encode4(hi, out[3], out[4], out[5]);
}

// 49 masked cells (R[2 vi] = vj 0..3, R[2 vi + 1] = vj 4..6 and a zero byte) -> 37 observation words:
// the cell stream (index 7 vi + vj) four cells at a time, 3 output words each
BB_HD void encode_view(const uint32_t R[14], uint32_t w[OBS_WORDS])
{
#pragma unroll
    for (int k = 0; k < 13; k++) {
        // stream bytes 4k .. 4k+3 live in at most two consecutive registers of R (7-byte records: 4 + 3)
        uint32_t sel = 0; int ra = -1;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int q = 4 * k + i;
            const int vi = q / 7, vj = q % 7;
            const int r = 2 * vi + (vj >= 4 ? 1 : 0), byte = vj >= 4 ? vj - 4 : vj;
            if (q >= 49) { sel |= 3u << (4 * i); continue; }          // byte 3 of R[13] is zero
            if (ra < 0) ra = r;
            sel |= (uint32_t)(r == ra ? byte : 4 + byte) << (4 * i);
        }
        const uint32_t c = byte_perm(R[ra], ra + 1 < 14 ? R[ra + 1] : 0u, sel);
        uint32_t o0, o1, o2;
        encode4(c, o0, o1, o2);
        w[3 * k] = o0;
        if (k < 12) { w[3 * k + 1] = o1; w[3 * k + 2] = o2; }
    }
}

// Single-room levels (num_rows = num_cols = 1): the only opaque cells are the outer walls (no doors; keys, balls
// and boxes let light through, minigrid.py:146-148), so process_vis lights exactly the view cells that lie inside
// the grid: the agent's row floods sideways up to and including the side walls, every interior row above it does
// the same, the far wall row is lit from below (corners through the diagonal rule) and nothing passes it.
// Visible <=> in bounds: no see-through flags, no transposes, no propagation -- the column windows with
// out-of-grid cells zeroed.  (test_room_observation_equals_generic compares it with observe_generic for every pose.)
template <class M>
BB_HD void observe_room_cells(const LevelParams &lp, const M &mem, int ax, int ay, int dir, int carry_cell, uint32_t R[14])
{
    const ViewGeom v = view_geom(lp, ax, ay, dir);
    // cells from the agent to the wall it faces (inclusive): view depths vj >= 6 - dist are inside the grid
    const int dist = dir == 3 ? ay : dir == 1 ? lp.H - 1 - ay : dir == 0 ? lp.W - 1 - ax : ax;
    const uint32_t dm = dist >= 6 ? 0x7Fu : (0x7Fu << (6 - dist)) & 0x7Fu;
    const uint32_t mlo = expand4(dm), mhi = expand4(dm >> 4) & 0x00FFFFFFu;
#pragma unroll
    for (int vi = 0; vi < 7; vi++) {
        const int row = v.c_row + v.rstep * (vi - 3);
        const bool rok = row >= 0 && row < v.nrows;
        const int off = v.off0 + vi * v.dstep;
        const uint32_t w0 = (rok && v.ok0) ? mem.word_at(off) : 0u;
        const uint32_t w1 = (rok && v.ok1) ? mem.word_at(off + 4) : 0u;
        const uint32_t w2 = (rok && v.ok2) ? mem.word_at(off + 8) : 0u;
        const uint32_t a = funnel_r(w0, w1, v.sh), b = funnel_r(w1, w2, v.sh);
        R[2 * vi] = byte_perm(a, b, v.sel_lo) & mlo;
        R[2 * vi + 1] = byte_perm(a, b, v.sel_hi) & mhi;
    }
    R[7] = (R[7] & 0xFF00FFFFu) | ((uint32_t)carry_cell << 16);        // the agent's own cell shows what it carries
}
template <class M>
BB_HD void observe_room(const LevelParams &lp, const M &mem, int ax, int ay, int dir, int carry_cell, uint32_t w[OBS_WORDS])
{
    uint32_t R[14];
    observe_room_cells(lp, mem, ax, ay, dir, carry_cell, R);
    encode_view(R, w);
}

// Writes the 147 observation bytes as 37 little-endian words (last byte 0): one lane does all columns.
template <class M>
BB_HD void observe_generic_cells(const LevelParams &lp, const M &mem, int ax, int ay, int dir, int carry_cell, uint32_t R[14])
{
    const ViewGeom v = view_geom(lp, ax, ay, dir);          // R[2 vi] = cells vj 0..3, R[2 vi + 1] = cells vj 4..6 (+1 unused byte)
#pragma unroll
    for (int vi = 0; vi < 7; vi++) col_load(mem, v, vi, R[2 * vi], R[2 * vi + 1]);
    // see-through bits: per column (bit vj), then transposed to per row (bit vi)
    uint32_t blo = 0, bhi = 0;
#pragma unroll
    for (int vi = 0; vi < 7; vi++) {
        const uint32_t cm = col_see(R[2 * vi], R[2 * vi + 1]);
        if (vi < 4) blo |= cm << (8 * vi); else bhi |= cm << (8 * (vi - 4));
    }
    transpose8(blo, bhi);
    uint32_t see[7], vis[7];
#pragma unroll
    for (int vj = 0; vj < 7; vj++) see[vj] = ((vj < 4 ? blo >> (8 * vj) : bhi >> (8 * (vj - 4)))) & 0x7Fu;
    vis_rows(see, vis);
    uint32_t vlo = 0, vhi = 0;
#pragma unroll
    for (int vj = 0; vj < 7; vj++) { if (vj < 4) vlo |= vis[vj] << (8 * vj); else vhi |= vis[vj] << (8 * (vj - 4)); }
    transpose8(vlo, vhi);                        // byte vi = visibility of column vi, bit vj
    // the agent's own cell (3, 6) shows what it carries (or empty); it is always visible
    R[7] = (R[7] & 0xFF00FFFFu) | ((uint32_t)carry_cell << 16);
#pragma unroll
    for (int vi = 0; vi < 7; vi++) {
        const uint32_t cv = (vi < 4 ? vlo >> (8 * vi) : vhi >> (8 * (vi - 4))) & 0x7Fu;
        R[2 * vi] &= expand4(cv);
        R[2 * vi + 1] &= expand4(cv >> 4);       // also clears the unused fourth byte
    }
}
template <class M>
BB_HD void observe_generic(const LevelParams &lp, const M &mem, int ax, int ay, int dir, int carry_cell, uint32_t w[OBS_WORDS])
{
    uint32_t R[14];
    observe_generic_cells(lp, mem, ax, ay, dir, carry_cell, R);
    encode_view(R, w);
}
// the 49 masked view cells of an observation (what encode_view / encode_stage_view expand to bytes)
template <class M>
BB_HD void observe_cells(const LevelParams &lp, const M &mem, int ax, int ay, int dir, int carry_cell, uint32_t R[14])
{
    if (lp.num_rows == 1 && lp.num_cols == 1) observe_room_cells(lp, mem, ax, ay, dir, carry_cell, R);
    else observe_generic_cells(lp, mem, ax, ay, dir, carry_cell, R);
}


template <class M>
BB_HD void observe(const LevelParams &lp, const M &mem, int ax, int ay, int dir, int carry_cell, uint32_t w[OBS_WORDS])
{
    if constexpr (mem_spec<M>::room_kinds != 0) { observe_room(lp, mem, ax, ay, dir, carry_cell, w); return; }
    if (lp.num_rows == 1 && lp.num_cols == 1) observe_room(lp, mem, ax, ay, dir, carry_cell, w);
    else observe_generic(lp, mem, ax, ay, dir, carry_cell, w);
}

// The same observation assembled from per-column pieces exactly as the 8-lanes-per-env kernel does
// (lane vi = column vi; the ballots become loops here).  Test cross-check only.
template <class M>
BB_HD void observe_columns(const LevelParams &lp, const M &mem, int ax, int ay, int dir, int carry_cell, uint8_t out[OBS_BYTES])
{
    const ViewGeom v = view_geom(lp, ax, ay, dir);
    uint32_t lo[7], hi[7], cm[7], see[7], vis[7];
    for (int vi = 0; vi < 7; vi++) { col_load(mem, v, vi, lo[vi], hi[vi]); cm[vi] = col_see(lo[vi], hi[vi]); }
    for (int vj = 0; vj < 7; vj++) { see[vj] = 0; for (int vi = 0; vi < 7; vi++) see[vj] |= ((cm[vi] >> vj) & 1u) << vi; }   // = 7 ballots
    vis_rows(see, vis);
    for (int vi = 0; vi < 7; vi++) {
        uint32_t cv = 0;
        for (int vj = 0; vj < 7; vj++) cv |= ((vis[vj] >> vi) & 1u) << vj;
        uint32_t h = hi[vi];
        if (vi == 3) h = (h & 0xFF00FFFFu) | ((uint32_t)carry_cell << 16);
        uint32_t o[6];
        col_encode(lo[vi], h, cv, o);
        for (int b = 0; b < 21; b++) out[21 * vi + b] = (uint8_t)(o[b >> 2] >> (8 * (b & 3)));
    }
}

// The same observation computed cell by cell (straightforward form; test cross-check only)
BB_HD void observe_simple(const LevelParams &lp, const uint8_t *grid, int ax, int ay, int dir, int carry_cell, uint8_t out[OBS_BYTES])
{
    const int fx = dir_dx(dir), fy = dir_dy(dir), rx = -fy, ry = fx;
    uint8_t cell[49];
    uint32_t see[7], vis[7];
    for (int vj = 0; vj < 7; vj++) {
        uint32_t s = 0;
        for (int vi = 0; vi < 7; vi++) {
            const int wx = ax + fx * (6 - vj) + rx * (vi - 3), wy = ay + fy * (6 - vj) + ry * (vi - 3);
            const bool inb = wx >= 0 && wx < lp.W && wy >= 0 && wy < lp.H;
            const uint32_t c = inb ? (uint32_t)get_cell(lp, grid, wx, wy) : (uint32_t)CELL_WALL;
            cell[vi * 7 + vj] = (uint8_t)c;
            const uint32_t t = c & 7u;
            s |= ((t == T_WALL || (t == T_DOOR && (c >> 6) != 0)) ? 0u : 1u) << vi;
        }
        see[vj] = s;
    }
    vis_rows(see, vis);
    cell[3 * 7 + 6] = (uint8_t)carry_cell;
    for (int vi = 0; vi < 7; vi++)
        for (int vj = 0; vj < 7; vj++) {
            const uint32_t c = ((vis[vj] >> vi) & 1u) ? cell[vi * 7 + vj] : 0u;
            uint8_t *px = out + (vi * 7 + vj) * 3;
            px[0] = (uint8_t)(c & 7u); px[1] = (uint8_t)((c >> 3) & 7u); px[2] = (uint8_t)(c >> 6);
        }
}

template <bool UNTR = false, class M>
BB_HD int carry_cell_of(const EnvHot &h, const M &mem)
{
    if constexpr (UNTR) { if (h.carry != NO_OBJ && (h.carry & CARRY_UNTRACKED)) return h.carry & 0x3F; }
    return h.carry == NO_OBJ ? CELL_EMPTY : mem.otc(h.carry);
}

// ---- staging of 32 observations of a warp as aligned words ----------------------
// Lane l owns bytes [147 l, 147 l + 147) of the 4704-byte warp tile; 147 is not a
// multiple of 4, so each lane funnel-shifts its 37 words by its own misalignment
// and the word that straddles two lanes is completed with the next lane's first
// word (obtained with a shuffle in the kernel).  Every tile word is written by
// exactly one lane, as one aligned 32-bit store.
BB_HD uint32_t funnel_l(uint32_t lo, uint32_t hi, int s)     // (hi:lo << s) >> 32, s in 0..31
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, s);
#else
    return s == 0 ? hi : (hi << s) | (lo >> (32 - s));
#endif
}
BB_HD void stage_obs_words(uint32_t *tile, const uint32_t w[OBS_WORDS], int lane, uint32_t next_w0)
{
    const int D = OBS_BYTES * lane;
    const int sh = D & 3, wb = D >> 2;
    const int kl = (sh + OBS_BYTES - 1) >> 2;                 // word holding this lane's last byte
    const int nvalid = ((sh + OBS_BYTES - 1) & 3) + 1;        // this lane's bytes in that word
    const int s8 = 8 * sh;
#pragma unroll
    for (int k = 0; k <= OBS_WORDS; k++) {
        const uint32_t lo = k > 0 ? w[k - 1] : 0u;
        const uint32_t hi = k < OBS_WORDS ? w[k] : 0u;
        uint32_t v = funnel_l(lo, hi, s8);
        if (k == 0 && sh != 0) continue;                      // first word belongs to the previous lane
        if (k > kl) continue;
        if (k == kl && nvalid < 4) v |= next_w0 << (8 * nvalid);
        tile[wb + k] = v;
    }
}


// Same idea for records of LBYTES bytes held in NW words (bytes beyond LBYTES must be zero):
// record number q of the tile starts at byte LBYTES * q; next_w0 = first word of record q + 1.
template <int LBYTES, int NW>
BB_HD void stage_record_words(uint32_t *tile, const uint32_t w[NW], int q, uint32_t next_w0)
{
    const int D = LBYTES * q;
    const int sh = D & 3, wb = D >> 2;
    const int kl = (sh + LBYTES - 1) >> 2;
    const int nvalid = ((sh + LBYTES - 1) & 3) + 1;
    const int s8 = 8 * sh;
#pragma unroll
    for (int k = 0; k <= NW; k++) {
        const uint32_t lo = k > 0 ? w[k - 1] : 0u;
        const uint32_t hi = k < NW ? w[k] : 0u;
        uint32_t v = funnel_l(lo, hi, s8);
        if (k == 0 && sh != 0) continue;
        if (k > kl) continue;
        if (k == kl && nvalid < 4) v |= next_w0 << (8 * nvalid);
        tile[wb + k] = v;
    }
}

}  // namespace bb
