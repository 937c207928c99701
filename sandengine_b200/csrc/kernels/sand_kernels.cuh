// sandengine_b200 device code (sm_100a).  Compiled at rule-compile time (NVRTC, -arch=sm_100a) together
// with the generated "rules_gen.cuh"; also compiled ahead of time by build.py for inspection
// (-Xptxas -v, cuobjdump).  No libc / CUDA headers are needed: plain CUDA C + inline PTX only.
//
// What is implemented here (reference: /root/reference/shaders/compute/gen/falling_sand.glsl):
//   se_hash43 / mirror decision          :58-79, :115-120, :698-701
//   se_block (Margolus block transition) :692-718  (rule bodies come from rules_gen.cuh)
//   se_step_global                       :676-733 + :737-799 restructured: ONE thread per 2x2 block
//                                        (the shader evaluates each block 4x, once per cell)
//   modification override                :749-794
//   se_light                             :469-524 (flood-fill lighting relaxation)
//   se_clear_frame1                      :743-746
//
// Cell storage in HBM: one packed uint32 per cell, value = material id (ids >= SE_N_MATERIALS read as
// NULL, gen/materials.glsl:79-86).  Inside kernels a cell is a "fat" 32-bit word (see codegen.h).
//
// Strip decomposition: a kernel sees a local buffer of `Hl` rows that starts at global row `gy0` of a
// grid of `Hg` rows (single GPU: gy0 = 0, Hl = Hg).  Rows outside the GLOBAL grid read as WALL
// (SCREEN_IS_BORDER, operations.glsl:45-51).  A block that needs a row inside the global grid but
// outside the local buffer is skipped: that can only happen in the outermost ghost rows, whose
// staleness is accounted for by the host's halo-exchange schedule (see sim.cpp).

struct SeRand { unsigned u[4]; };

#define SE_ID(c) ((c) & 0xFFu)
#define SE_TYPE(c) (((c) >> 8) & 0xFFu)
#define SE_RANK(c) ((c) >> 24)
#define SE_F_NOSWAP 0x00010000u
#define SE_F_EMISSIVE 0x00020000u
#define SE_F_OBSTACLE 0x00040000u
// density(a) < density(b) on ranks stored in the top byte: (a | 0x00FFFFFF) < b
#define SE_DENS_LT(a, b) ((((a) | 0x00FFFFFFu)) < (b))
#define SE_DENS_GT(a, b) SE_DENS_LT(b, a)
#define SE_DENS_LE(a, b) (!SE_DENS_LT(b, a))
#define SE_DENS_GE(a, b) (!SE_DENS_LT(a, b))
#define SE_DENS_EQ(a, b) (SE_RANK(a) == SE_RANK(b))
#define SE_DENS_NE(a, b) (SE_RANK(a) != SE_RANK(b))
#define SE_ISTYPE32(c, m) ((((m) >> SE_TYPE(c)) & 1u) != 0u)
#define SE_ISTYPE64(c, m) ((((m) >> SE_TYPE(c)) & 1ull) != 0ull)
// guarded swap (operations.glsl:16-23): refuses when either side is WALL / NULL typed
#define SE_SWAP(a, b) do { if ((((a) | (b)) & SE_F_NOSWAP) == 0u) { const unsigned t_ = (a); (a) = (b); (b) = t_; } } while (0)
// hash lane -> float exactly as vec4(hash4i(x)) / float(0xffffffffU): the divisor rounds to 2^32
#define SE_RANDF(k) (__fmul_rn(__uint2float_rn(rnd.u[k]), 2.3283064365386963e-10f))

#include "rules_gen.cuh"

#ifndef SE_EXPERIMENTAL_KERNELS
#define SE_EXPERIMENTAL_KERNELS 0   // se_step_lut_global_census / se_build_popbits / se_light_fused (see the SE_FLAG_*_EXPERIMENTAL flags)
#endif

typedef unsigned long long se_u64;

struct SeMod { int px, py, shape, size, mat, pad0, pad1, pad2; };   // 32 B == simulation.rs:45-56

struct SeStepParams {
    const unsigned* in;    // local buffer, row 0 == global row gy0
    unsigned* out;         // == in for the in-place update
    int W, Hl, gy0, Hg;
    int frame;
    int n_mods;            // 0..256, already cut at the first mod_size == 0 (falling_sand.glsl:754-756)
    const SeMod* mods;     // device pointer
};

// ---- Chris Wellons' lowbias32 as used by the shader (math.glsl:17-25) ----
static __device__ __forceinline__ unsigned se_hashi(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
// (Writing the shifts as high multiplies -- x >> k == umulhi(x, 2^(32-k)) -- to move them from the saturated ALU
// pipe to the FMA pipe was measured SLOWER on B200: 931 -> 851 Gcell/s; IMAD.HI issues at a lower rate.)
// hash43(uvec3(pos_rounded, frame)) -- only the lanes the rule set consumes are evaluated
static __device__ __forceinline__ void se_hash43(int px, int py, int frame, SeRand& rnd) {
    const unsigned x = (unsigned)px * 461u + (unsigned)py * 2131u + (unsigned)frame * (2131u * 2131u);
    rnd.u[0] = se_hashi(x * 213u);
    rnd.u[1] = (SE_RAND_LANES & 2u) ? se_hashi(x * 2131u) : 0u;
    rnd.u[2] = (SE_RAND_LANES & 4u) ? se_hashi(x * 21313u) : 0u;
    rnd.u[3] = (SE_RAND_LANES & 8u) ? se_hashi(x * 213132u) : 0u;
}

// falling_sand.glsl:701-718 for a given RAND.  The caller has already taken the all-EMPTY early-out (:692-694).
static __device__ __forceinline__ void se_block_with_rand(unsigned& s, unsigned& r, unsigned& d, unsigned& dr, const SeRand& rnd, int px, int py, int frame) {
    const bool mirror = rnd.u[0] <= SE_MIRROR_UMAX;   // rand.x < 0.5
    if (mirror) { SE_SWAP(s, r); SE_SWAP(d, dr); }
    se_apply_mirrored(s, r, d, dr, rnd, px, py, frame);
#if SE_HAVE_LEFT_RULES
    if (mirror) se_apply_left(s, r, d, dr, rnd, px, py, frame);
#endif
    if (mirror) { SE_SWAP(s, r); SE_SWAP(d, dr); }
#if SE_HAVE_RIGHT_RULES
    if (!mirror) se_apply_right(s, r, d, dr, rnd, px, py, frame);
#endif
}

// falling_sand.glsl:698-718
static __device__ __forceinline__ void se_block(unsigned& s, unsigned& r, unsigned& d, unsigned& dr, int px, int py, int frame) {
    SeRand rnd;
    se_hash43(px, py, frame, rnd);
    se_block_with_rand(s, r, d, dr, rnd, px, py, frame);
}

static __device__ __forceinline__ void se_margolus_offset(int frame, int& ox, int& oy) {   // operations.glsl:25-34
    const int f = frame & 3;   // frame >= 0
    ox = (f == 1 || f == 3) ? 1 : 0;
    oy = (f == 1 || f == 2) ? 1 : 0;
}

// Per-cell modification scan (falling_sand.glsl:749-794). Returns true and the material id when the
// cell is overridden.  The circle test is evaluated in f32 exactly as written in the shader.
static __device__ __forceinline__ bool se_mod_lookup(const SeMod* __restrict__ mods, int n_mods, int x, int y, unsigned& mat_out) {
    bool got = false;
    int fin = 1;   // MAT_NULL
    for (int i = 0; i < n_mods; ++i) {
        const SeMod m = mods[i];
        const int dx = abs(m.px - x), dy = abs(m.py - y);
        bool hit = false;
        if (m.shape == 0) {
            const float fx = (float)dx, fy = (float)dy;
            const float dist = __fsqrt_rn(__fadd_rn(__fmul_rn(fx, fx), __fmul_rn(fy, fy)));
            hit = dist <= (float)m.size;
        } else if (m.shape == 1) {
            hit = dx <= m.size && dy <= m.size;
        }
        if (hit) { got = true; fin = m.mat; }
    }
    mat_out = (unsigned)fin;
    return got && fin != 1;
}

#ifndef SE_HOST_EMU   // kernels (the pure device functions above are also compiled on the host by tests/emu)
// ---------------------------------------------------------------------------------------------
// K1a: one Margolus step, one thread per 2x2 block, straight from/to global memory.
//   IN_PLACE : out == in, only cells whose id changed are stored (blocks partition the grid, so the
//              update is race-free in place -- SURVEY.md section 7)
//   HAS_MODS : apply the modification override per cell
// Thread (tx, ty) of the grid handles block column bx of block row jb0 + by.
// ---------------------------------------------------------------------------------------------
template <bool IN_PLACE, bool HAS_MODS>
static __device__ __forceinline__ void se_step_global_impl(const SeStepParams& p, const unsigned* __restrict__ fat_sm) {
    int ox, oy;
    se_margolus_offset(p.frame, ox, oy);
    const int bx = blockIdx.x * blockDim.x + threadIdx.x;
    const int x0 = 2 * bx - ox;
    if (x0 >= p.W) return;
    const int jb0 = (p.gy0 + oy) >> 1;                      // first block row that touches the local buffer
    const int jl = blockIdx.y * blockDim.y + threadIdx.y;
    const int y0 = 2 * (jb0 + jl) - oy;                     // global row of the block's top cells
    if (y0 >= p.Hg || y0 >= p.gy0 + p.Hl) return;
    const int y1 = y0 + 1;
    // row status: 0 = in buffer, 1 = outside the global grid (WALL), 2 = missing (inside grid, not local)
    const int st0 = (y0 < 0) ? 1 : (y0 < p.gy0 ? 2 : 0);
    const int st1 = (y1 >= p.Hg) ? 1 : (y1 >= p.gy0 + p.Hl ? 2 : 0);
    if (st0 == 2 || st1 == 2) return;
    const bool cx0 = x0 >= 0, cx1 = (x0 + 1) < p.W;
    const size_t row0 = (size_t)(y0 - p.gy0) * (size_t)p.W, row1 = (size_t)(y1 - p.gy0) * (size_t)p.W;

    unsigned raw[4];
    raw[0] = (st0 == 0 && cx0) ? p.in[row0 + x0] : 2u;
    raw[1] = (st0 == 0 && cx1) ? p.in[row0 + x0 + 1] : 2u;
    raw[2] = (st1 == 0 && cx0) ? p.in[row1 + x0] : 2u;
    raw[3] = (st1 == 0 && cx1) ? p.in[row1 + x0 + 1] : 2u;

    unsigned s = fat_sm[min(raw[0], 255u)], r = fat_sm[min(raw[1], 255u)];
    unsigned d = fat_sm[min(raw[2], 255u)], dr = fat_sm[min(raw[3], 255u)];
    if ((raw[0] | raw[1] | raw[2] | raw[3]) != 0u) {       // all-EMPTY early-out, falling_sand.glsl:692-694
        se_block(s, r, d, dr, x0, y0, p.frame);
    }
    unsigned res[4] = {SE_ID(s), SE_ID(r), SE_ID(d), SE_ID(dr)};
    if (HAS_MODS) {
        unsigned m;
        if (se_mod_lookup(p.mods, p.n_mods, x0, y0, m)) res[0] = m;
        if (se_mod_lookup(p.mods, p.n_mods, x0 + 1, y0, m)) res[1] = m;
        if (se_mod_lookup(p.mods, p.n_mods, x0, y1, m)) res[2] = m;
        if (se_mod_lookup(p.mods, p.n_mods, x0 + 1, y1, m)) res[3] = m;
    }
    if (st0 == 0) {
        if (cx0 && (!IN_PLACE || res[0] != raw[0])) p.out[row0 + x0] = res[0];
        if (cx1 && (!IN_PLACE || res[1] != raw[1])) p.out[row0 + x0 + 1] = res[1];
    }
    if (st1 == 0) {
        if (cx0 && (!IN_PLACE || res[2] != raw[2])) p.out[row1 + x0] = res[2];
        if (cx1 && (!IN_PLACE || res[3] != raw[3])) p.out[row1 + x0 + 1] = res[3];
    }
}

// Modifications are culled per CTA: thread t tests record t against the CTA's cell rectangle (+ the
// record's size) and the survivors are compacted IN ORDER (last match wins, falling_sand.glsl:764-773)
// into shared memory, so the per-cell scan only walks the records that can touch this CTA.
static __device__ __forceinline__ int se_cull_mods(const SeStepParams& p, SeMod* mods_sm, int* warp_counts) {
    int ox, oy;
    se_margolus_offset(p.frame, ox, oy);
    const int t = threadIdx.y * blockDim.x + threadIdx.x;
    const int x_lo = 2 * (int)(blockIdx.x * blockDim.x) - ox, x_hi = x_lo + 2 * (int)blockDim.x - 1;
    const int jb0 = (p.gy0 + oy) >> 1;
    const int y_lo = 2 * (jb0 + (int)(blockIdx.y * blockDim.y)) - oy, y_hi = y_lo + 2 * (int)blockDim.y - 1;
    bool keep = false;
    SeMod m;
    if (t < p.n_mods) {
        m = p.mods[t];
        // both shapes are contained in the square |dx| <= size, |dy| <= size (negative sizes never match)
        keep = m.size >= 0 && m.px + m.size >= x_lo && m.px - m.size <= x_hi && m.py + m.size >= y_lo && m.py - m.size <= y_hi;
    }
    const unsigned ballot = __ballot_sync(0xFFFFFFFFu, keep);
    const int warp = t >> 5, lane = t & 31;
    if (lane == 0) warp_counts[warp] = __popc(ballot);
    __syncthreads();
    int base = 0, total = 0;
    for (int w = 0; w < 8; ++w) { if (w < warp) base += warp_counts[w]; total += warp_counts[w]; }
    if (keep) mods_sm[base + __popc(ballot & ((1u << lane) - 1u))] = m;
    __syncthreads();
    return total;
}

#define SE_DEFINE_STEP_GLOBAL(NAME, IN_PLACE, HAS_MODS)                                  \
    extern "C" __global__ void __launch_bounds__(256) NAME(const SeStepParams p) {        \
        __shared__ unsigned fat_sm[256];                                                  \
        __shared__ SeMod mods_sm[HAS_MODS ? 256 : 1];                                     \
        __shared__ int warp_counts[8];                                                    \
        const int t = threadIdx.y * blockDim.x + threadIdx.x;                             \
        if (t < 256) fat_sm[t] = se_fat_table[t];                                         \
        SeStepParams q = p;                                                               \
        if (HAS_MODS) { q.n_mods = se_cull_mods(p, mods_sm, warp_counts); q.mods = mods_sm; } \
        __syncthreads();                                                                  \
        se_step_global_impl<IN_PLACE, HAS_MODS>(q, fat_sm);                               \
    }
SE_DEFINE_STEP_GLOBAL(se_step_inplace, true, false)
SE_DEFINE_STEP_GLOBAL(se_step_inplace_mods, true, true)
SE_DEFINE_STEP_GLOBAL(se_step_pingpong, false, false)
SE_DEFINE_STEP_GLOBAL(se_step_pingpong_mods, false, true)

#endif  // SE_HOST_EMU (K1a)

// ---------------------------------------------------------------------------------------------
// K3: lighting relaxation (operations.glsl:114-169).
//   old_cells : material ids BEFORE this step (input_data)    new_cells : ids AFTER this step
//   light_in / light_out : float4 per cell, ping-pong (simulation.rs:239)
// The module is compiled with -fmad=false: every sum below is evaluated as written (no FMA), in the
// shader's neighbour order DOWN, UP, DOWNLEFT, UPLEFT, DOWNRIGHT, UPRIGHT, RIGHT, LEFT (math.glsl:154-166).
//
// The per-neighbour term (rgb * keep * a, a) does not depend on which cell reads it, so a CTA computes it ONCE
// per cell of its 32 x 32 tile (+ a one-cell ring) into shared memory -- one id load, one float4 load and one
// table lookup per cell instead of eight.  Then each of the 256 threads walks 4 consecutive rows of one column
// with a sliding 3 x 3 window of terms in registers (3 shared-memory loads per cell instead of 8) and combines
// the 8 neighbours in the shader's order with the shader's arithmetic (the reader-dependent part is only
// `falloff = (a == 0) ? max_falloff : a`).  CTAs whose tile + ring lies strictly inside the grid (and the local
// buffer) take the INTERIOR path: all 8 neighbours exist, so there are no validity tests and the division by
// `num` is the exact multiplication by 0.125; the CTAs on the rim keep the general form.
// The two phases are plain per-thread functions so that tests/emu can run them on the host, CTA by CTA.
// ---------------------------------------------------------------------------------------------
struct SeLightParams {
    const unsigned* old_cells;
    const unsigned* new_cells;
    const float4* light_in;
    float4* light_out;
    int W, Hl, gy0, Hg;
};

#define SE_LT_W 32                       // tile width  (= lanes of a warp: one row segment of 512 B per warp access)
#ifndef SE_LT_ROWS
#define SE_LT_ROWS 4                     // consecutive rows per thread (tunable at rule-compile time: env SE_LT_ROWS = 2 | 4 | 8)
#endif
#ifndef SE_LT_MINCTAS
#define SE_LT_MINCTAS 4                  // resident CTAs per SM the register allocation aims at (env SE_LT_MINCTAS)
#endif
#define SE_LT_H (8 * SE_LT_ROWS)         // tile height (8 warps x SE_LT_ROWS rows)
#define SE_LT_STRIDE (SE_LT_W + 2)       // term row stride (float4)
#define SE_LT_TERMS ((SE_LT_H + 2) * SE_LT_STRIDE)

// true when the tile of CTA (bx, by) and its one-cell ring lie inside the grid and inside the local buffer
static __device__ __forceinline__ bool se_light_tile_is_interior(const SeLightParams& p, int bx, int by) {
    const int x_lo = bx * SE_LT_W - 1, x_hi = bx * SE_LT_W + SE_LT_W;         // ring columns
    const int yl_lo = by * SE_LT_H - 1, yl_hi = by * SE_LT_H + SE_LT_H;       // ring rows (local)
    return x_lo >= 0 && x_hi < p.W && yl_lo >= 0 && yl_hi < p.Hl && p.gy0 + yl_lo >= 0 && p.gy0 + yl_hi < p.Hg;
}

// term of ring/tile cell (i, j): i = local row - (tile row 0 - 1), j = column - (tile column 0 - 1)
template <bool INTERIOR>
static __device__ __forceinline__ void se_light_stage_one(const SeLightParams& p, const unsigned* fat, float4* term, int bx, int by, int i, int j) {
    const int nx = bx * SE_LT_W - 1 + j, nyl = by * SE_LT_H - 1 + i, ny = p.gy0 + nyl;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (INTERIOR || (nx >= 0 && nx < p.W && ny >= 0 && ny < p.Hg && nyl >= 0 && nyl < p.Hl)) {   // outOfBounds, :139-141
        const size_t nidx = (size_t)nyl * p.W + nx;
        const unsigned id = p.old_cells[nidx];
        const float4 li = p.light_in[nidx];
        const unsigned nf = fat[id < 255u ? id : 255u];
        const float keep = (nf & SE_F_OBSTACLE) ? 0.0f : 1.0f;     // vec4(vec3(float(!obstacle)), 1.0), :498
        const float la = li.w;                                     // a * 1.0 == a
        v = make_float4((li.x * keep) * la, (li.y * keep) * la, (li.z * keep) * la, la);
    }
    term[i * SE_LT_STRIDE + j] = v;
}

// phase 1, thread `tid` of 256: tile columns row by row (one warp per row: aligned 512-byte segments), then the
// two ring columns (68 cells) in one pass
template <bool INTERIOR>
static __device__ __forceinline__ void se_light_stage(const SeLightParams& p, const unsigned* fat, float4* term, int bx, int by, int tid) {
    const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int it = 0; it < (SE_LT_H + 2 + 7) / 8; ++it) {
        const int i = warp + 8 * it;
        if (i < SE_LT_H + 2) se_light_stage_one<INTERIOR>(p, fat, term, bx, by, i, lane + 1);
    }
    if (tid < 2 * (SE_LT_H + 2)) se_light_stage_one<INTERIOR>(p, fat, term, bx, by, tid >> 1, (tid & 1) * (SE_LT_W + 1));
}

// one neighbour, in the shader's order (:500-513); `mx.w` is never read by the shader's result
#define SE_LIGHT_ACC(v)                                                               \
    {                                                                                 \
        const float falloff_ = ((v).w == 0.0f) ? max_falloff : (v).w;                 \
        avg.x += (v).x; avg.y += (v).y; avg.z += (v).z; avg.w += falloff_;            \
        max_falloff = fmaxf(falloff_, max_falloff);                                   \
        mx.x = fmaxf(mx.x, (v).x); mx.y = fmaxf(mx.y, (v).y); mx.z = fmaxf(mx.z, (v).z); \
    }

// phase 2, thread `tid` of 256: column tid & 31, rows 4 * (tid >> 5) .. + 3 of the tile
// `new_id(idx, x, y, tile_row)` supplies the cell's material id AFTER this step: se_light reads it from new_cells,
// the fused kernel takes it from its shared-memory tile (and stores it).
struct SeNewIdFromGlobal {
    const unsigned* new_cells;
    __device__ __forceinline__ unsigned operator()(size_t idx, int, int, int, int) const { return new_cells[idx]; }
};

template <bool INTERIOR, class NewId>
static __device__ __forceinline__ void se_light_compute(const SeLightParams& p, const unsigned* fat, const float4* term, int bx, int by, int tid, const NewId& new_id) {
    const int tx = tid & 31, row0 = (tid >> 5) * SE_LT_ROWS;
    const int x = bx * SE_LT_W + tx;
    if (!INTERIOR && x >= p.W) return;
    const float4* tp = term + row0 * SE_LT_STRIDE + tx;            // term (row0 - 1, tx - 1) of the tile
    float4 a0 = tp[0], a1 = tp[1], a2 = tp[2];                     // row above
    float4 b0 = tp[SE_LT_STRIDE], b1 = tp[SE_LT_STRIDE + 1], b2 = tp[SE_LT_STRIDE + 2];   // own row
#pragma unroll
    for (int i = 0; i < SE_LT_ROWS; ++i) {
        const int yl = by * SE_LT_H + row0 + i;
        if (!INTERIOR && yl >= p.Hl) break;
        const float4 c0 = tp[(i + 2) * SE_LT_STRIDE], c1 = tp[(i + 2) * SE_LT_STRIDE + 1], c2 = tp[(i + 2) * SE_LT_STRIDE + 2];   // row below
        const int y = p.gy0 + yl;                                  // global row
        const size_t idx = (size_t)yl * p.W + x;
        const unsigned id = new_id(idx, x, y, row0 + i, tx);
        const unsigned me = id < 255u ? id : 255u;
        float4 light;
        if (fat[me] & SE_F_EMISSIVE) {                             // :126-127
            light = make_float4(se_emission_table[me * 4 + 0], se_emission_table[me * 4 + 1], se_emission_table[me * 4 + 2], se_emission_table[me * 4 + 3]);
        } else if (!INTERIOR && y == 0) {                          // :128-129 (row 0 is never part of an interior tile)
            light = make_float4(1.0f, 1.0f, 1.0f, 0.999999f);
        } else {
            float4 avg = make_float4(0.f, 0.f, 0.f, 0.f), mx = make_float4(0.f, 0.f, 0.f, 0.f);
            float max_falloff = 0.0f;
            if (INTERIOR) {
                // DOWN, UP, DOWNLEFT, UPLEFT, DOWNRIGHT, UPRIGHT, RIGHT, LEFT (math.glsl:154-166)
                SE_LIGHT_ACC(c1) SE_LIGHT_ACC(a1) SE_LIGHT_ACC(c0) SE_LIGHT_ACC(a0)
                SE_LIGHT_ACC(c2) SE_LIGHT_ACC(a2) SE_LIGHT_ACC(b2) SE_LIGHT_ACC(b0)
                // num == 8: x / 8.0f == x * 0.125f exactly (both are the correctly rounded x / 8)
                avg.x *= 0.125f; avg.y *= 0.125f; avg.z *= 0.125f; avg.w *= 0.125f;
            } else {
                const bool up = yl - 1 >= 0 && y - 1 >= 0, down = yl + 1 < p.Hl && y + 1 < p.Hg;
                const bool left = x - 1 >= 0, right = x + 1 < p.W;
                int num = 0;
                if (down) { SE_LIGHT_ACC(c1) ++num; }
                if (up) { SE_LIGHT_ACC(a1) ++num; }
                if (down && left) { SE_LIGHT_ACC(c0) ++num; }
                if (up && left) { SE_LIGHT_ACC(a0) ++num; }
                if (down && right) { SE_LIGHT_ACC(c2) ++num; }
                if (up && right) { SE_LIGHT_ACC(a2) ++num; }
                if (right) { SE_LIGHT_ACC(b2) ++num; }
                if (left) { SE_LIGHT_ACC(b0) ++num; }
                if (num > 0) {                                     // :516-518
                    const float dn = (float)num;
                    avg.x = __fdiv_rn(avg.x, dn); avg.y = __fdiv_rn(avg.y, dn); avg.z = __fdiv_rn(avg.z, dn); avg.w = __fdiv_rn(avg.w, dn);
                }
            }
            // mix(avg.rgb, max.rgb, 0.5) = avg*(1-0.5) + max*0.5, :521
            light = make_float4(avg.x * 0.5f + mx.x * 0.5f, avg.y * 0.5f + mx.y * 0.5f, avg.z * 0.5f + mx.z * 0.5f, avg.w);
        }
        p.light_out[idx] = light;
        a0 = b0; a1 = b1; a2 = b2;
        b0 = c0; b1 = c1; b2 = c2;
    }
}

#ifndef SE_HOST_EMU
extern "C" __global__ void __launch_bounds__(256, SE_LT_MINCTAS) se_light(const SeLightParams p) {
    __shared__ unsigned fat_sm[256];
    __shared__ float4 term[SE_LT_TERMS];
    const int tid = threadIdx.x;
    fat_sm[tid] = se_fat_table[tid];
    __syncthreads();
    const int bx = blockIdx.x, by = blockIdx.y;
    if (se_light_tile_is_interior(p, bx, by)) {
        se_light_stage<true>(p, fat_sm, term, bx, by, tid);
        __syncthreads();
        se_light_compute<true>(p, fat_sm, term, bx, by, tid, SeNewIdFromGlobal{p.new_cells});
    } else {
        se_light_stage<false>(p, fat_sm, term, bx, by, tid);
        __syncthreads();
        se_light_compute<false>(p, fat_sm, term, bx, by, tid, SeNewIdFromGlobal{p.new_cells});
    }
}
#endif  // SE_HOST_EMU (se_light kernel)

// ---------------------------------------------------------------------------------------------
// K5: colour shading (operations.glsl:100-108, math.glsl:82-111) -- the render product of `setCell`:
// colour = material colour, and for non-EMPTY cells rgb -= 0.25 * noise(pos, 3 octaves, lacunarity 2, freq 0.25),
// clamped to [0,1].  A pure function of (material id, position), so it is evaluated on demand from the
// id buffer instead of being stored every step like the reference's output_color texture.
// Uses sinf (not __sinf); the hash `fract(sin(p) * 43758.5453)` amplifies last-bit differences of sin, so
// parity with the CPU restatement is a tolerance statement (tests/test_gpu_parity.py::test_colour_shading).
// ---------------------------------------------------------------------------------------------
static __device__ __forceinline__ float se_fract(float x) { return x - floorf(x); }
static __device__ __forceinline__ float2 se_old_hash2(float px, float py) {   // math.glsl:40-44
    const float a = px * 127.1f + py * 311.7f;
    const float b = px * 269.5f + py * 183.3f;
    return make_float2(-1.0f + 2.0f * se_fract(sinf(a) * 43758.5453123f), -1.0f + 2.0f * se_fract(sinf(b) * 43758.5453123f));
}
static __device__ float se_simplex(float px, float py) {                      // math.glsl:82-96
    const float K1 = 0.366025404f, K2 = 0.211324865f;
    const float sk = (px + py) * K1;
    const float ix = floorf(px + sk), iy = floorf(py + sk);
    const float t = (ix + iy) * K2;
    const float ax = px - ix + t, ay = py - iy + t;
    const float m = (ax < ay) ? 0.0f : 1.0f;                                   // step(a.y, a.x)
    const float ox = m, oy = 1.0f - m;
    const float bx = ax - ox + K2, by = ay - oy + K2;
    const float cx = ax - 1.0f + 2.0f * K2, cy = ay - 1.0f + 2.0f * K2;
    const float h0 = fmaxf(0.5f - (ax * ax + ay * ay), 0.0f);
    const float h1 = fmaxf(0.5f - (bx * bx + by * by), 0.0f);
    const float h2 = fmaxf(0.5f - (cx * cx + cy * cy), 0.0f);
    const float2 g0 = se_old_hash2(ix + 0.0f, iy + 0.0f), g1 = se_old_hash2(ix + ox, iy + oy), g2 = se_old_hash2(ix + 1.0f, iy + 1.0f);
    const float n0 = h0 * h0 * h0 * h0 * (ax * g0.x + ay * g0.y);
    const float n1 = h1 * h1 * h1 * h1 * (bx * g1.x + by * g1.y);
    const float n2 = h2 * h2 * h2 * h2 * (cx * g2.x + cy * g2.y);
    return 0.25f + 0.5f * (n0 * 70.0f + n1 * 70.0f + n2 * 70.0f);
}

// colour of one cell (the kernel below and tests/emu run exactly this)
static __device__ __forceinline__ float4 se_shade_cell(unsigned id, int x, int y) {
    float4 col = make_float4(se_color_table[id * 4 + 0], se_color_table[id * 4 + 1], se_color_table[id * 4 + 2], se_color_table[id * 4 + 3]);
    if (id != 0u) {                                                            // cell.mat != MAT_EMPTY
        float fx = (float)x, fy = (float)y, f = 0.0f;
        for (int o = 1; o < 4; ++o) {                                          // noise(pos, 3, 2.0, 0.25), math.glsl:98-107
            f += 1.0f / (float)o * se_simplex(fx * 0.25f, fy * 0.25f);
            fx *= 2.0f; fy *= 2.0f;
        }
        const float rnd = f * 0.25f;
        col.x = fminf(fmaxf(col.x - rnd, 0.0f), 1.0f);
        col.y = fminf(fmaxf(col.y - rnd, 0.0f), 1.0f);
        col.z = fminf(fmaxf(col.z - rnd, 0.0f), 1.0f);
    }
    return col;
}

#ifndef SE_HOST_EMU
struct SeShadeParams {
    const unsigned* cells;   // first OWNED row
    float4* rgba_f32;        // or nullptr
    unsigned* rgba8;         // or nullptr: packed R | G<<8 | B<<16 | A<<24, round-to-nearest of clamp(c)*255
    int W, rows, y0;         // y0 = global row of the first owned row (positions feed the noise)
};

extern "C" __global__ void __launch_bounds__(256) se_shade(const SeShadeParams p) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int yl = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= p.W || yl >= p.rows) return;
    const size_t idx = (size_t)yl * p.W + x;
    const unsigned id = min(p.cells[idx], 255u);
    const float4 col = se_shade_cell(id, x, p.y0 + yl);
    if (p.rgba_f32) p.rgba_f32[idx] = col;
    if (p.rgba8) {
        const unsigned r = (unsigned)__float2int_rn(fminf(fmaxf(col.x, 0.f), 1.f) * 255.0f), g = (unsigned)__float2int_rn(fminf(fmaxf(col.y, 0.f), 1.f) * 255.0f);
        const unsigned b = (unsigned)__float2int_rn(fminf(fmaxf(col.z, 0.f), 1.f) * 255.0f), a = (unsigned)__float2int_rn(fminf(fmaxf(col.w, 0.f), 1.f) * 255.0f);
        p.rgba8[idx] = r | (g << 8) | (b << 16) | (a << 24);
    }
}
#endif  // SE_HOST_EMU (se_shade kernel)

#ifndef SE_HOST_EMU
// frame == 1: every cell becomes EMPTY (falling_sand.glsl:743-746); lighting (if on) then runs with
// new_cells == all-EMPTY through se_light.
extern "C" __global__ void __launch_bounds__(256) se_fill_cells(unsigned* cells, size_t n, unsigned value) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) cells[i] = value;
}


#endif  // SE_HOST_EMU

// =============================================================================================
// K1b: transition-table kernel with shared-memory tiles and temporal blocking.
//
// Eligible rule sets (SE_LUT_ELIGIBLE, decided by the code generator): the block transition is a pure
// function of (4 material ids, mirror bit, rand.y) with rand.y used only against literal thresholds --
// no pos/frame/other rand use -- there are no non-mirrored rules and N = SE_N_MATERIALS <= 12.  Then
//     T0[idx],  idx = ((a*N + b)*N + c)*N + d            (N^4 entries, 16 bit)
// holds the UNMIRRORED transition of every block state, nibble-packed (a | b<<4 | c<<8 | d<<12), built on
// the device by se_build_lut from the very rule code generated for the generic path.  Entry kinds:
//     0x0000..0xEFFF  the result (ids < 15)
//     0xF000 | k      result depends on rand.y: pool[k] = {thr, A, B}: result = (u1 <= thr) ? A : B, where A/B
//                     are again entries (B may chain to another pool entry for states with several thresholds)
//     0xFFFF          the state or its result holds WALL / NULL (guarded swaps, operations.glsl:16-23):
//                     take the generic path (same generated code as K1a)
// The mirrored transition is  swap_pairs(T0[swap_pairs(state)])  -- exact when no cell of the block
// refuses to swap, which is precisely when the table is used.
//
// One persistent CTA loops over tiles: load (uint4 global -> u8 shared), nsub Margolus sub-steps in
// shared memory with a halo of T columns / T/2+1 rows, store the interior (u8 shared -> uint4 global, ping-pong buffer).
// HBM traffic per cell-update ~ (4 * tile/interior + 4) / T bytes instead of 8.
// =============================================================================================
#if SE_LUT_ELIGIBLE
#define SE_N4 (SE_N_MATERIALS * SE_N_MATERIALS * SE_N_MATERIALS * SE_N_MATERIALS)
// SE_LUT_TWO_TABLES (rule sets with Left/Right rules, EXPERIMENTAL): entries [0, N^4) hold the unmirrored view, entries
// [N^4, 2 N^4) the final result of the MIRRORED evaluation (swap, mirrored + left rules, swap back) of the same state,
// so the lookup needs no byte swaps; without Left/Right rules one table serves both views through the mirror symmetry.
#define SE_LUT_ENTRIES (SE_N4 * (SE_LUT_TWO_TABLES ? 2 : 1))
#define SE_TILE_PW 256          // tile width in cells (= bytes); 64 words per row
#define SE_LUT_POOL_MAX 4095
#define SE_LUT_SLOW 0xFFFFu
#ifndef SE_TILE_THREADS
#define SE_TILE_THREADS 512
#endif

struct SePoolEntry { unsigned thr; unsigned short a, b; };   // 8 bytes

// one block state: evaluate every rand.y class with the generated rule code, then encode
static __device__ __forceinline__ void se_build_lut_entry(int idx, unsigned short* __restrict__ base, SePoolEntry* __restrict__ pool,
                                                          unsigned* __restrict__ counter) {
    const int N = SE_N_MATERIALS;
#if SE_LUT_TWO_TABLES
    const int entry = idx;
    const bool mirror_table = entry >= SE_N4;
    idx = entry - (mirror_table ? SE_N4 : 0);
#endif
    const unsigned ia = idx / (N * N * N), ib = (idx / (N * N)) % N, ic = (idx / N) % N, id = idx % N;
    unsigned short res[SE_LUT_NCLS];
    // class c <=> rand.y hash lane u1 in (U_{c-1}, U_c]  (U_{-1} = -1, U_{NCLS-1} = 2^32-1)
    bool noswap = ((se_fat_table[ia] | se_fat_table[ib] | se_fat_table[ic] | se_fat_table[id]) & SE_F_NOSWAP) != 0u;
#pragma unroll
    for (int cls = 0; cls < SE_LUT_NCLS; ++cls) {
        unsigned s = se_fat_table[ia], r = se_fat_table[ib], d = se_fat_table[ic], dr = se_fat_table[id];
        if (idx != 0) {
            SeRand rnd;
#if SE_LUT_TWO_TABLES
            rnd.u[0] = mirror_table ? 0u : 0xFFFFFFFFu;                   // rand.x < 0.5: the mirrored evaluation, start to finish
#else
            rnd.u[0] = 0xFFFFFFFFu;                                       // rand.x >= 0.5: unmirrored view
#endif
            rnd.u[1] = cls == 0 ? 0u : se_lut_thresholds[cls - 1] + 1u;  // representative of the class
            rnd.u[2] = 0u; rnd.u[3] = 0u;
            se_block_with_rand(s, r, d, dr, rnd, 0, 0, 0);
        }
        if ((s | r | d | dr) & SE_F_NOSWAP) noswap = true;
        res[cls] = (unsigned short)(SE_ID(s) | (SE_ID(r) << 4) | (SE_ID(d) << 8) | (SE_ID(dr) << 12));
    }
#if SE_LUT_TWO_TABLES
    idx = entry;                                                          // the slot written below
#endif
    if (noswap) { base[idx] = SE_LUT_SLOW; return; }
    int n_breaks = 0;
#pragma unroll
    for (int cls = 1; cls < SE_LUT_NCLS; ++cls) n_breaks += (res[cls] != res[cls - 1]) ? 1 : 0;
    if (n_breaks == 0) { base[idx] = res[0]; return; }
    const unsigned k0 = atomicAdd(counter, (unsigned)n_breaks);
    if (k0 + n_breaks > SE_LUT_POOL_MAX) { base[idx] = SE_LUT_SLOW; return; }   // overflow: the host refuses the table
    base[idx] = (unsigned short)(0xF000u | k0);
    unsigned k = k0;
    int left = n_breaks;
    for (int cls = 1; cls < SE_LUT_NCLS; ++cls) {
        if (res[cls] != res[cls - 1]) {
            --left;
            SePoolEntry e;
            e.thr = se_lut_thresholds[cls - 1];            // u1 <= U_{cls-1}  <=>  class < cls
            e.a = res[cls - 1];
            e.b = left ? (unsigned short)(0xF000u | (k + 1)) : res[cls];
            pool[k] = e;
            ++k;
        }
    }
}

#ifndef SE_HOST_EMU
extern "C" __global__ void __launch_bounds__(256) se_build_lut(unsigned short* __restrict__ base, SePoolEntry* __restrict__ pool,
                                                               unsigned* __restrict__ counter) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < SE_LUT_ENTRIES) se_build_lut_entry(idx, base, pool, counter);
}
#endif

struct SeTileParams {
    unsigned* buf0;         // local buffer (row 0 == global row gy0) read by the first T-block of the launch
    unsigned* buf1;         // the other ping-pong buffer; T-block k reads buf[k & 1] and writes buf[(k + 1) & 1]
    int W, Hl, gy0, Hg;
    int frame0;             // frame number of the first sub-step of this launch
    int nblk;               // T-blocks in this launch (>= 1)
    int tsteps;             // sub-steps of every T-block but the last (<= the halo allows)
    int nsub_last;          // sub-steps of the last T-block, 1..tsteps
    unsigned seq_base;      // T-blocks completed by earlier launches (the per-tile flags count in this sequence)
    unsigned* done;         // per tile: sequence number of its last completed T-block (dataflow between T-blocks)
    int HY;                 // halo depth in rows (even, >= nsub/2 + 1: the row offset changes every OTHER frame,
                            // operations.glsl:25-34, so validity shrinks by at most floor(n/2)+1 rows in n steps)
    int HX;                 // halo depth in columns (multiple of 4, >= nsub: the column offset alternates every frame)
    int PH;                 // tile height in cells (even)
    int tiles_x, tiles_y;
    int lut_words;          // 32-bit words of (base + pad + pool) to stage in shared memory
    int pool_offset;        // byte offset of the pool inside the staged table (8-aligned)
    int tile_offset;        // byte offset of the tile inside dynamic shared memory (16-aligned)
    const unsigned* lut;
};

static __device__ __forceinline__ unsigned se_pack_ids(uint4 v) {
    const unsigned a = v.x < SE_N_MATERIALS ? v.x : 1u, b = v.y < SE_N_MATERIALS ? v.y : 1u;
    const unsigned c = v.z < SE_N_MATERIALS ? v.z : 1u, d = v.w < SE_N_MATERIALS ? v.w : 1u;
    return a | (b << 8) | (c << 16) | (d << 24);
}

static __device__ __forceinline__ unsigned se_nibbles_to_bytes(unsigned e) {
    unsigned x = (e | (e << 8)) & 0x00FF00FFu;
    return (x | (x << 4)) & 0x0F0F0F0Fu;
}

// Table access: on the device the table lives in shared memory and is addressed with 32-bit shared
// addresses (explicit ld.shared keeps the compiler from re-deriving the shared window base per access);
// the host emulation reads the same layout through plain pointers.
#ifdef SE_HOST_EMU
typedef const unsigned char* se_tab_t;
static inline unsigned se_tab_u16(se_tab_t t, unsigned byte_off) { unsigned short v; std::memcpy(&v, t + byte_off, 2); return v; }
static inline void se_tab_pool(se_tab_t t, unsigned byte_off, unsigned& thr, unsigned& ab) { std::memcpy(&thr, t + byte_off, 4); std::memcpy(&ab, t + byte_off + 4, 4); }
static inline unsigned se_idx4(unsigned vv) {
    return (((vv & 0xFFu) * SE_N_MATERIALS + ((vv >> 8) & 0xFFu)) * SE_N_MATERIALS + ((vv >> 16) & 0xFFu)) * SE_N_MATERIALS + (vv >> 24);
}
#else
typedef unsigned se_tab_t;   // shared-space address
static __device__ __forceinline__ unsigned se_lds_u8(unsigned a) { unsigned v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
static __device__ __forceinline__ unsigned se_lds_u16(unsigned a) { unsigned v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
static __device__ __forceinline__ unsigned se_lds_u32(unsigned a) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
static __device__ __forceinline__ void se_sts_u8(unsigned a, unsigned v) { asm volatile("st.shared.u8 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
static __device__ __forceinline__ void se_sts_u16(unsigned a, unsigned v) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
static __device__ __forceinline__ void se_sts_u32(unsigned a, unsigned v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
static __device__ __forceinline__ unsigned se_tab_u16(se_tab_t t, unsigned byte_off) { return se_lds_u16(t + byte_off); }
static __device__ __forceinline__ void se_tab_pool(se_tab_t t, unsigned byte_off, unsigned& thr, unsigned& ab) {
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(thr), "=r"(ab) : "r"(t + byte_off));
}
// idx = ((a*N + b)*N + c)*N + d from the four id bytes: one dp4a for a*N^2 + b*N + c
static __device__ __forceinline__ unsigned se_idx4(unsigned vv) {
    return __dp4a(vv, (unsigned)(SE_N_MATERIALS * SE_N_MATERIALS) | ((unsigned)SE_N_MATERIALS << 8) | (1u << 16), 0u) * SE_N_MATERIALS + (vv >> 24);
}
#endif

// one block: v = a | b<<8 | c<<16 | d<<24 (material ids), returns the new ids in the same packing
static __device__ __forceinline__ unsigned se_block_lut(unsigned v, unsigned seed, int px, int py, int frame,
                                                        se_tab_t tab, unsigned pool_off, const unsigned* __restrict__ fat_sm) {
    const unsigned u0 = se_hashi(seed * 213u);
    const bool mirror = u0 <= SE_MIRROR_UMAX;
#if SE_LUT_TWO_TABLES
    unsigned e = se_tab_u16(tab, (se_idx4(v) + (mirror ? (unsigned)SE_N4 : 0u)) * 2u);
#else
    const unsigned vv = mirror ? __byte_perm(v, 0u, 0x2301) : v;
    unsigned e = se_tab_u16(tab, se_idx4(vv) * 2u);
#endif
    if (e >= 0xF000u) {
        if (e != SE_LUT_SLOW) {
            const unsigned u1 = se_hashi(seed * 2131u);
            do {
                unsigned thr, ab;
                se_tab_pool(tab, pool_off + (e & 0xFFFu) * 8u, thr, ab);
                e = (u1 <= thr) ? (ab & 0xFFFFu) : (ab >> 16);
            } while (e >= 0xF000u);
        } else {
            // generic path (the block touches WALL / NULL): same generated code as K1a
            unsigned s = fat_sm[v & 0xFFu], r = fat_sm[(v >> 8) & 0xFFu], d = fat_sm[(v >> 16) & 0xFFu], dr = fat_sm[v >> 24];
            SeRand rnd;
            rnd.u[0] = u0;
            rnd.u[1] = (SE_RAND_LANES & 2u) ? se_hashi(seed * 2131u) : 0u;
            rnd.u[2] = 0u; rnd.u[3] = 0u;
            se_block_with_rand(s, r, d, dr, rnd, px, py, frame);
            return SE_ID(s) | (SE_ID(r) << 8) | (SE_ID(d) << 16) | (SE_ID(dr) << 24);
        }
    }
#if SE_LUT_TWO_TABLES
    return se_nibbles_to_bytes(e);
#else
    const unsigned rr = se_nibbles_to_bytes(e);
    return mirror ? __byte_perm(rr, 0u, 0x2301) : rr;
#endif
}

// ---------------------------------------------------------------------------------------------
// Running census (EXPERIMENTAL, SE_FLAG_RUNNING_CENSUS; off by default): the per-material population of the
// owned rows is kept up to date by the per-frame kernel K1c instead of being recounted by a pass over the grid.
// Guarded swaps only permute the cells of a block, so the population changes only where a SET fired (or where a
// block straddles the first/last owned row of a strip, or holds an id the table does not know).  `popbits` is
// a bit per block state: 1 = some outcome of the state (any rand.y class, either mirror view) is not a
// permutation of its four ids.  It is only a filter: blocks that pass it are compared cell by cell.
// ---------------------------------------------------------------------------------------------
static __device__ __forceinline__ unsigned se_sorted4(unsigned a, unsigned b, unsigned c, unsigned d) {
    unsigned t;
    if (a > b) { t = a; a = b; b = t; }
    if (c > d) { t = c; c = d; d = t; }
    if (a > c) { t = a; a = c; c = t; }
    if (b > d) { t = b; b = d; d = t; }
    if (b > c) { t = b; b = c; c = t; }
    return a | (b << 8) | (c << 16) | (d << 24);
}

// same evaluation as se_build_lut_entry: every rand.y class of the unmirrored view
static __device__ __forceinline__ bool se_popchange_entry(int idx) {
    const int N = SE_N_MATERIALS;
    const unsigned ia = idx / (N * N * N), ib = (idx / (N * N)) % N, ic = (idx / N) % N, id = idx % N;
    if ((se_fat_table[ia] | se_fat_table[ib] | se_fat_table[ic] | se_fat_table[id]) & SE_F_NOSWAP) return true;   // generic path
    if (idx == 0) return false;
    const unsigned before = se_sorted4(ia, ib, ic, id);
#pragma unroll
    for (int cls = 0; cls < SE_LUT_NCLS; ++cls) {
        unsigned s = se_fat_table[ia], r = se_fat_table[ib], d = se_fat_table[ic], dr = se_fat_table[id];
        SeRand rnd;
        rnd.u[0] = 0xFFFFFFFFu;
        rnd.u[1] = cls == 0 ? 0u : se_lut_thresholds[cls - 1] + 1u;
        rnd.u[2] = 0u; rnd.u[3] = 0u;
        se_block_with_rand(s, r, d, dr, rnd, 0, 0, 0);
        if ((s | r | d | dr) & SE_F_NOSWAP) return true;
        if (se_sorted4(SE_ID(s), SE_ID(r), SE_ID(d), SE_ID(dr)) != before) return true;
    }
    return false;
}

// sets the bit of state idx and of its mirror image (the lookup uses the unmirrored block whatever rand.x says)
static __device__ __forceinline__ void se_build_popbits_entry(int idx, unsigned* __restrict__ popbits) {
    if (!se_popchange_entry(idx)) return;
    const int N = SE_N_MATERIALS;
    const int ia = idx / (N * N * N), ib = (idx / (N * N)) % N, ic = (idx / N) % N, id = idx % N;
    const int mirrored = ((ib * N + ia) * N + id) * N + ic;
    atomicOr(popbits + (idx >> 5), 1u << (idx & 31));
    atomicOr(popbits + (mirrored >> 5), 1u << (mirrored & 31));
}

#ifdef SE_HOST_EMU
typedef const unsigned* se_pop_t;
static inline unsigned se_popbit(se_pop_t pop, unsigned idx) { return (pop[idx >> 5] >> (idx & 31)) & 1u; }
#else
typedef unsigned se_pop_t;   // shared-space address
static __device__ __forceinline__ unsigned se_popbit(se_pop_t pop, unsigned idx) {
    unsigned w;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(pop + 4u * (idx >> 5)));
    return (w >> (idx & 31)) & 1u;
}
#endif

// One block of K1c: v = clamped ids (what the table saw), raw* = the words read from memory, nv = new ids,
// cm = bit k set when cell k (a, b, c, d) is inside the grid AND in an owned row.  Adds the population deltas of
// the counted cells to hist[256] (CTA-local, flushed once per launch).
static __device__ __forceinline__ void se_census_block(int* hist, se_pop_t pop, unsigned v, unsigned ra, unsigned rb, unsigned rc, unsigned rd,
                                                       unsigned nv, unsigned cm) {
    const unsigned na = nv & 0xFFu, nb = (nv >> 8) & 0xFFu, nc = (nv >> 16) & 0xFFu, nd = nv >> 24;
    if (cm == 0u || (na == ra && nb == rb && nc == rc && nd == rd)) return;
    // all four cells counted and every outcome of this state is a permutation: nothing to do.  (A raw id the
    // table does not know reads as NULL, NULL states have their bit set, so v is what decides.)
    if (cm == 0xFu && !se_popbit(pop, se_idx4(v))) return;
    if ((cm & 1u) && na != ra) { atomicAdd(hist + (ra < 255u ? ra : 255u), -1); atomicAdd(hist + na, 1); }
    if ((cm & 2u) && nb != rb) { atomicAdd(hist + (rb < 255u ? rb : 255u), -1); atomicAdd(hist + nb, 1); }
    if ((cm & 4u) && nc != rc) { atomicAdd(hist + (rc < 255u ? rc : 255u), -1); atomicAdd(hist + nc, 1); }
    if ((cm & 8u) && nd != rd) { atomicAdd(hist + (rd < 255u ? rd : 255u), -1); atomicAdd(hist + nd, 1); }
}

// ---------------------------------------------------------------------------------------------
// K3f (EXPERIMENTAL, SE_FLAG_FUSED_LIGHT_EXPERIMENTAL; off by default): Margolus step + modification override +
// lighting relaxation in ONE pass over the grid for table-eligible rule sets.  se_light already stages the old ids
// of its 32 x SE_LT_H tile and a one-cell ring; every 2x2 block that covers a tile cell lies inside tile + ring
// (the block offset is 0 or 1), so the CTA can run the transition table on those ids in shared memory, apply the
// modification override per cell, and feed the new ids straight into the light combine: the separate step kernel
// (K1a ping-pong) and one read of the id buffer disappear (48 -> 40 physical bytes per cell).
// Blocks cut by a tile edge are evaluated by both CTAs (deterministic: RAND depends on position and frame only).
// Three per-thread phases, run CTA by CTA on the host by tests/emu:
//   stage  : old id (clamped; WALL outside the grid; MISSING inside the grid but outside the local buffer) and
//            the light term of every tile + ring cell
//   blocks : transition table on the blocks covering the tile, in place in the shared id array
//   compute: override, store the new id, combine the eight neighbour terms (se_light_compute)
// ---------------------------------------------------------------------------------------------
#define SE_LF_IDS_STRIDE (SE_LT_W + 4)
#define SE_LF_IDS_BYTES ((SE_LT_H + 2) * SE_LF_IDS_STRIDE)
#define SE_LF_MISSING 0xFFu

struct SeFusedParams {
    SeLightParams lp;        // old_cells: ids before the step; new_cells: ids after the step (WRITTEN by this kernel)
    int frame;
    int n_mods;              // already cut at the first mod_size == 0
    const SeMod* mods;
    int lut_words, pool_offset;
    int tile_offset;         // byte offset of term[] in dynamic shared memory (behind the staged table, 16-aligned)
    const unsigned* lut;
    int tiles_x, tiles_y;
};

// can record m touch a cell of the rectangle?  (both shapes lie inside the square |dx|,|dy| <= size; negative sizes never match)
static __device__ __forceinline__ bool se_mod_touches(const SeMod& m, int x_lo, int x_hi, int y_lo, int y_hi) {
    return m.size >= 0 && m.px + m.size >= x_lo && m.px - m.size <= x_hi && m.py + m.size >= y_lo && m.py - m.size <= y_hi;
}

template <bool INTERIOR>
static __device__ __forceinline__ void se_fused_stage_one(const SeLightParams& p, const unsigned* fat, float4* term, unsigned char* ids,
                                                          int bx, int by, int i, int j) {
    const int nx = bx * SE_LT_W - 1 + j, nyl = by * SE_LT_H - 1 + i, ny = p.gy0 + nyl;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    unsigned idb;
    const bool in_grid = INTERIOR || (nx >= 0 && nx < p.W && ny >= 0 && ny < p.Hg);
    if (INTERIOR || (in_grid && nyl >= 0 && nyl < p.Hl)) {
        const size_t nidx = (size_t)nyl * p.W + nx;
        const unsigned id = p.old_cells[nidx];
        const float4 li = p.light_in[nidx];
        const unsigned nf = fat[id < 255u ? id : 255u];
        const float keep = (nf & SE_F_OBSTACLE) ? 0.0f : 1.0f;
        const float la = li.w;
        v = make_float4((li.x * keep) * la, (li.y * keep) * la, (li.z * keep) * la, la);
        idb = id < SE_N_MATERIALS ? id : 1u;                       // unknown ids read as NULL (gen/materials.glsl:79-86)
    } else {
        idb = in_grid ? SE_LF_MISSING : 2u;                        // WALL outside the grid (operations.glsl:45-51)
    }
    term[i * SE_LT_STRIDE + j] = v;
    ids[i * SE_LF_IDS_STRIDE + j] = (unsigned char)idb;
}

template <bool INTERIOR>
static __device__ __forceinline__ void se_fused_stage(const SeLightParams& p, const unsigned* fat, float4* term, unsigned char* ids, int bx, int by, int tid) {
    const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int it = 0; it < (SE_LT_H + 2 + 7) / 8; ++it) {
        const int i = warp + 8 * it;
        if (i < SE_LT_H + 2) se_fused_stage_one<INTERIOR>(p, fat, term, ids, bx, by, i, lane + 1);
    }
    if (tid < 2 * (SE_LT_H + 2)) se_fused_stage_one<INTERIOR>(p, fat, term, ids, bx, by, tid >> 1, (tid & 1) * (SE_LT_W + 1));
}

// the blocks that cover the tile: (SE_LT_W/2 + ox) x (SE_LT_H/2 + oy) of them, in place in ids[]
static __device__ __forceinline__ void se_fused_blocks(const SeLightParams& p, int frame, se_tab_t tab, unsigned pool_off, const unsigned* fat,
                                                       unsigned char* ids, int bx, int by, int tid) {
    int ox, oy;
    se_margolus_offset(frame, ox, oy);
    const int nbx = SE_LT_W / 2 + ox, nby = SE_LT_H / 2 + oy;
    const unsigned fterm = (unsigned)frame * (2131u * 2131u);
    for (int b = tid; b < nbx * nby; b += 256) {
        const int bj = b / nbx, bi = b - bj * nbx;
        const int cx = 2 * bi + 1 - ox, cy = 2 * bj + 1 - oy;      // ring coordinates of the block's top-left cell
        const int x0 = bx * SE_LT_W - 1 + cx, y0 = p.gy0 + by * SE_LT_H - 1 + cy;
        unsigned char* q = ids + cy * SE_LF_IDS_STRIDE + cx;
        const unsigned a = q[0], bb = q[1], c = q[SE_LF_IDS_STRIDE], d = q[SE_LF_IDS_STRIDE + 1];
        if ((a | bb | c | d) & 0x80u) continue;                    // a row of the block is not in the local buffer: skipped (see K1a)
        const unsigned v = a | (bb << 8) | (c << 16) | (d << 24);
        const unsigned seed = (unsigned)x0 * 461u + (unsigned)y0 * 2131u + fterm;
        const unsigned nv = se_block_lut(v, seed, x0, y0, frame, tab, pool_off, fat);
        q[0] = (unsigned char)(nv & 0xFFu); q[1] = (unsigned char)((nv >> 8) & 0xFFu);
        q[SE_LF_IDS_STRIDE] = (unsigned char)((nv >> 16) & 0xFFu); q[SE_LF_IDS_STRIDE + 1] = (unsigned char)(nv >> 24);
    }
}

// new id of a tile cell: the table's result, overridden by the (culled) modification list, stored to the id buffer
template <bool HAS_MODS>
struct SeNewIdFused {
    const unsigned char* ids;
    unsigned* new_cells;
    const SeMod* mods;
    int n_mods;
    __device__ __forceinline__ unsigned operator()(size_t idx, int x, int y, int tile_row, int tile_col) const {
        unsigned id = ids[(tile_row + 1) * SE_LF_IDS_STRIDE + tile_col + 1];
        if (HAS_MODS) {
            unsigned m;
            if (se_mod_lookup(mods, n_mods, x, y, m)) id = m;
        }
        new_cells[idx] = id;
        return id;
    }
};

#ifndef SE_HOST_EMU
// one Margolus sub-step over the whole tile; OX (column phase) is a template parameter so that the
// aligned 16-bit and the byte access variants are separate straight-line loops
template <int OX>
static __device__ __forceinline__ void se_tile_substep(unsigned tile_sa, se_tab_t tab, unsigned pool_off, const unsigned* __restrict__ fat_sm,
                                                       int PH, int oy, int frame, int gx_org, int gy_org, int warp, int nwarps, int lane) {
    const int PW = SE_TILE_PW;
    const int nbx = (PW - OX) >> 1, nby = (PH - oy) >> 1;
    const unsigned fterm = (unsigned)frame * (2131u * 2131u) + (unsigned)(gx_org + OX) * 461u;
    for (int j = warp; j < nby; j += nwarps) {
        const int ly = 2 * j + oy;
        const unsigned rowseed = (unsigned)(gy_org + ly) * 2131u + fterm;
        const unsigned row_sa = tile_sa + (unsigned)(ly * PW + OX);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = lane + 32 * k;
            if (OX == 0 || i < nbx) {
                const unsigned c0 = row_sa + 2u * (unsigned)i;
                unsigned v;
                if (OX == 0) v = se_lds_u16(c0) | (se_lds_u16(c0 + PW) << 16);
                else v = se_lds_u8(c0) | (se_lds_u8(c0 + 1) << 8) | (se_lds_u8(c0 + PW) << 16) | (se_lds_u8(c0 + PW + 1) << 24);
                // No branch for the all-EMPTY early-out (falling_sand.glsl:692-694): T0[0] == 0 by construction
                // (se_build_lut_entry skips the rules for state 0), so an empty block maps to itself.
                {
                    const unsigned seed = rowseed + (unsigned)(2 * i) * 461u;
                    const unsigned nv = se_block_lut(v, seed, gx_org + OX + 2 * i, gy_org + ly, frame, tab, pool_off, fat_sm);
                    if (nv != v) {
                        if (OX == 0) {
                            se_sts_u16(c0, nv & 0xFFFFu);
                            se_sts_u16(c0 + PW, nv >> 16);
                        } else {
                            se_sts_u8(c0, nv & 0xFFu); se_sts_u8(c0 + 1, (nv >> 8) & 0xFFu);
                            se_sts_u8(c0 + PW, (nv >> 16) & 0xFFu); se_sts_u8(c0 + PW + 1, nv >> 24);
                        }
                    }
                }
            }
        }
    }
}

extern "C" __global__ void __launch_bounds__(SE_TILE_THREADS, 2) se_step_tiles(const SeTileParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ unsigned fat_sm[256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    unsigned smem_sa;   // laundered through asm so the compiler keeps it in a register instead of re-deriving it (S2R) per access
    asm volatile("mov.u32 %0, %1;" : "=r"(smem_sa) : "r"((unsigned)__cvta_generic_to_shared(smem)));
    const se_tab_t tab = smem_sa;
    const unsigned pool_off = (unsigned)p.pool_offset;
    const unsigned tile_sa = smem_sa + (unsigned)p.tile_offset;

    // stage the table with 16-byte loads (it is re-read by every CTA of every launch: keep this prologue short)
    {
        const uint4* lut4 = reinterpret_cast<const uint4*>(p.lut);
        const int n4 = (p.lut_words + 3) >> 2;           // the device buffer is padded to a multiple of 16 bytes
        for (int i = tid; i < n4; i += blockDim.x) {
            const uint4 v = __ldg(lut4 + i);
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(smem_sa + 16u * i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        }
    }
    if (tid < 256) fat_sm[tid] = se_fat_table[tid];
    __syncthreads();

    const int PW = SE_TILE_PW, PH = p.PH;
    const int TWo = PW - 2 * p.HX, THo = PH - 2 * p.HY;
    const int n_tiles = p.tiles_x * p.tiles_y;
    // Work items (k, t) = (T-block, tile), ordered by w = k * n_tiles + t and dealt round-robin to the persistent
    // CTAs, each taking its items in increasing order.  A tile of T-block k needs the results of T-block k-1 in
    // its 3x3 tile neighbourhood only, so instead of a grid-wide barrier (or a launch) between T-blocks every
    // tile publishes a sequence number when its interior is stored and consumers wait on the nine flags they
    // need.  CTAs that run out of work in T-block k move on to k+1: no launch gap, no table re-staging, and
    // the partially filled last round of one T-block is filled with tiles of the next.  Deadlock-free because
    // every dependency has a smaller w and all CTAs are co-resident (grid <= occupancy x SMs).
    const long long total_items = (long long)p.nblk * n_tiles;
    for (long long w = blockIdx.x; w < total_items; w += gridDim.x) {
        const int k = (int)(w / n_tiles), t = (int)(w - (long long)k * n_tiles);
        const unsigned* in = (k & 1) ? p.buf1 : p.buf0;
        unsigned* out = (k & 1) ? p.buf0 : p.buf1;
        const int nsub = (k == p.nblk - 1) ? p.nsub_last : p.tsteps;
        const int tframe0 = p.frame0 + k * p.tsteps;
        const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
        if (k > 0) {
            if (tid < 9) {
                const int ntx = tx + (tid % 3) - 1, nty = ty + (tid / 3) - 1;
                if (ntx >= 0 && ntx < p.tiles_x && nty >= 0 && nty < p.tiles_y) {
                    const unsigned* flag = p.done + (nty * p.tiles_x + ntx);
                    const unsigned want = p.seq_base + (unsigned)k;
                    unsigned v;
                    while (true) {
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
                        if ((int)(v - want) >= 0) break;
                        __nanosleep(100);
                    }
                }
            }
            __syncthreads();
        }
        const int gx_org = tx * TWo - p.HX;              // global x of tile column 0 (multiple of 4)
        const int gy_org = p.gy0 + ty * THo - p.HY;      // global y of tile row 0 (even)
        const bool border = gx_org < 0 || gx_org + PW > p.W || gy_org < 0 || gy_org + PH > p.Hg;

        // ---- load: uint4 of packed-u32 cells -> 4 id bytes ----
        for (int r = warp; r < PH; r += nwarps) {
            const int gy = gy_org + r, lr = gy - p.gy0;
            const bool row_ok = gy >= 0 && gy < p.Hg && lr >= 0 && lr < p.Hl;
            const uint4* src = reinterpret_cast<const uint4*>(in + (size_t)(row_ok ? lr : 0) * p.W);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int q = lane + 32 * h;
                const int gx = gx_org + 4 * q;
                // __ldcg (L2 only): the data may have been written by another SM earlier in this very launch
                unsigned wv = 0x02020202u;               // WALL outside the grid (operations.glsl:45-51)
                if (row_ok && gx >= 0 && gx < p.W) wv = se_pack_ids(__ldcg(src + (gx >> 2)));
                se_sts_u32(tile_sa + 4u * (unsigned)(r * (PW / 4) + q), wv);
            }
        }
        __syncthreads();

        // ---- nsub Margolus sub-steps in shared memory ----
        for (int sub = 0; sub < nsub; ++sub) {
            const int frame = tframe0 + sub;
            int ox, oy;
            se_margolus_offset(frame, ox, oy);
            if (ox == 0) se_tile_substep<0>(tile_sa, tab, pool_off, fat_sm, PH, oy, frame, gx_org, gy_org, warp, nwarps, lane);
            else se_tile_substep<1>(tile_sa, tab, pool_off, fat_sm, PH, oy, frame, gx_org, gy_org, warp, nwarps, lane);
            __syncthreads();
            if (border) {
                // cells outside the grid are WALL at every step, whatever a SET wrote into them
                for (int r = warp; r < PH; r += nwarps) {
                    const int gy = gy_org + r;
                    const bool row_out = gy < 0 || gy >= p.Hg;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int q = lane + 32 * h;
                        const int gx = gx_org + 4 * q;
                        if (row_out || gx < 0 || gx >= p.W) se_sts_u32(tile_sa + 4u * (unsigned)(r * (PW / 4) + q), 0x02020202u);
                    }
                }
                __syncthreads();
            }
        }

        // ---- store the interior: 4 id bytes -> uint4 of packed-u32 cells ----
        for (int r = p.HY + warp; r < PH - p.HY; r += nwarps) {
            const int gy = gy_org + r, lr = gy - p.gy0;
            if (gy >= p.Hg || lr >= p.Hl) break;
            uint4* dst = reinterpret_cast<uint4*>(out + (size_t)lr * p.W);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int q = lane + 32 * h;
                const int gx = gx_org + 4 * q;
                if (4 * q >= p.HX && 4 * q < PW - p.HX && gx < p.W) {
                    const unsigned wv = se_lds_u32(tile_sa + 4u * (unsigned)(r * (PW / 4) + q));
                    dst[gx >> 2] = make_uint4(wv & 0xFFu, (wv >> 8) & 0xFFu, (wv >> 16) & 0xFFu, wv >> 24);
                }
            }
        }
        __syncthreads();                                  // every thread's stores are issued ...
        if (tid == 0) {                                   // ... and made visible before the tile is published
            __threadfence();
            asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p.done + t), "r"(p.seq_base + (unsigned)k + 1u) : "memory");
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K1c: ONE Margolus step with the transition table, straight from/to global memory, in place.
// This is the per-frame path (`Simulation::run()` called once per frame): a single step cannot amortise the
// tile load/store of K1b, so blocks are read directly (two 8-byte loads per block when the column phase is
// even, four 4-byte loads otherwise), pushed through se_block_lut and only the cells that changed are
// written back.  Persistent CTAs keep the table in shared memory; a warp walks 32 consecutive blocks of
// one block row, a CTA walks block rows grid-stride.  HBM-bound: ~4 B read + (changed fraction) x 4 B
// written per cell.
// ---------------------------------------------------------------------------------------------
struct SeLutStepParams {
    unsigned* cells;        // local buffer, updated in place
    int W, Hl, gy0, Hg;
    int frame;
    int lut_words, pool_offset;
    const unsigned* lut;
};

#define SE_K1C_THREADS 512
#define SE_K1C_SPAN 8           // a work item = one block row x SPAN chunks of 32 blocks (amortises the row set-up)
#ifndef SE_K1C_BATCH
#define SE_K1C_BATCH 4          // chunks whose loads are issued together (the update is in place, so the compiler
#endif                          // cannot hoist loads over the previous chunk's stores by itself); 4: 425 us, 2: 461 us @16384^2
#ifndef SE_K1C_MINCTAS
#define SE_K1C_MINCTAS 2
#endif

// running census (EXPERIMENTAL): extra arguments of se_step_lut_global_census
struct SeLutCensusParams {
    unsigned long long* census;   // 256 bins: population of the owned rows, updated in place
    const unsigned* popbits;      // ceil(N^4 / 32) words
    int pop_words;
    int pop_offset;               // byte offset of the staged popbits in dynamic shared memory (16-aligned, behind the table)
    int own_y0, own_y1;           // owned global rows [own_y0, own_y1)
};

template <bool CENSUS>
static __device__ __forceinline__ void se_k1c_body(const SeLutStepParams& p, const SeLutCensusParams& cx, int* hist_sm) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ unsigned fat_sm[256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    unsigned smem_sa;
    asm volatile("mov.u32 %0, %1;" : "=r"(smem_sa) : "r"((unsigned)__cvta_generic_to_shared(smem)));
    if (CENSUS) {
        for (int i = tid; i < cx.pop_words; i += blockDim.x)
            asm volatile("st.shared.u32 [%0], %1;" :: "r"(smem_sa + (unsigned)cx.pop_offset + 4u * i), "r"(__ldg(cx.popbits + i)) : "memory");
        for (int i = tid; i < 256; i += blockDim.x) hist_sm[i] = 0;
    }
    {
        const uint4* lut4 = reinterpret_cast<const uint4*>(p.lut);
        const int n4 = (p.lut_words + 3) >> 2;
        for (int i = tid; i < n4; i += blockDim.x) {
            const uint4 v = __ldg(lut4 + i);
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(smem_sa + 16u * i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        }
    }
    if (tid < 256) fat_sm[tid] = se_fat_table[tid];
    __syncthreads();
    const se_tab_t tab = smem_sa;
    const unsigned pool_off = (unsigned)p.pool_offset;

    int ox, oy;
    se_margolus_offset(p.frame, ox, oy);
    const int jb0 = (p.gy0 + oy) >> 1;
    const int y_end = min(p.Hg, p.gy0 + p.Hl);
    const unsigned nby = (unsigned)(((y_end + oy + 1) >> 1) - jb0);
    const int nbx = (p.W + ox + 1) >> 1;
    const unsigned chunks_x = (unsigned)((nbx + 31) >> 5);                 // 32 blocks per warp-iteration
    const unsigned spans_x = (chunks_x + SE_K1C_SPAN - 1) / SE_K1C_SPAN;
    const unsigned n_items = nby * spans_x;
    const unsigned fterm = (unsigned)p.frame * (2131u * 2131u);
    const bool vec_ok = (ox == 0) && ((p.W & 1) == 0);                     // 8-byte aligned pairs
    for (unsigned item = blockIdx.x * nwarps + warp; item < n_items; item += gridDim.x * nwarps) {
        const unsigned jl = item / spans_x, sp = item - jl * spans_x;
        const int y0 = 2 * (jb0 + (int)jl) - oy, y1 = y0 + 1;
        const int st0 = (y0 < 0) ? 1 : (y0 < p.gy0 ? 2 : 0);
        const int st1 = (y1 >= p.Hg) ? 1 : (y1 >= p.gy0 + p.Hl ? 2 : 0);
        if (st0 == 2 || st1 == 2) continue;                                // missing ghost row: block row skipped (see K1a)
        unsigned* base0 = p.cells + (size_t)(st0 == 0 ? y0 - p.gy0 : 0) * (size_t)p.W;
        unsigned* base1 = p.cells + (size_t)(st1 == 0 ? y1 - p.gy0 : 0) * (size_t)p.W;
        const unsigned rowseed = (unsigned)y0 * 2131u + fterm;
        // census: which of the two rows are counted (inside the grid and owned by this strip)
        const unsigned cm_rows = !CENSUS ? 0u : ((st0 == 0 && y0 >= cx.own_y0 && y0 < cx.own_y1) ? 3u : 0u) | ((st1 == 0 && y1 >= cx.own_y0 && y1 < cx.own_y1) ? 12u : 0u);
        const int c_begin = (int)sp * SE_K1C_SPAN, c_end = min((int)chunks_x, c_begin + SE_K1C_SPAN);
        for (int cb = c_begin; cb < c_end; cb += SE_K1C_BATCH) {
            unsigned a[SE_K1C_BATCH], b[SE_K1C_BATCH], c[SE_K1C_BATCH], d[SE_K1C_BATCH];
#pragma unroll
            for (int u = 0; u < SE_K1C_BATCH; ++u) {                       // ---- loads of the whole batch first
                const int bx = (cb + u) * 32 + lane;
                const int x0 = 2 * bx - ox;
                a[u] = b[u] = c[u] = d[u] = 2u;                            // WALL outside the grid
                if (cb + u < c_end && bx < nbx) {
                    if (vec_ok) {
                        if (st0 == 0) { const uint2 t = *reinterpret_cast<const uint2*>(base0 + x0); a[u] = t.x; b[u] = t.y; }
                        if (st1 == 0) { const uint2 t = *reinterpret_cast<const uint2*>(base1 + x0); c[u] = t.x; d[u] = t.y; }
                    } else {
                        const bool cx0 = x0 >= 0, cx1 = (x0 + 1) < p.W;
                        if (st0 == 0 && cx0) a[u] = base0[x0];
                        if (st0 == 0 && cx1) b[u] = base0[x0 + 1];
                        if (st1 == 0 && cx0) c[u] = base1[x0];
                        if (st1 == 0 && cx1) d[u] = base1[x0 + 1];
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < SE_K1C_BATCH; ++u) {                       // ---- transitions + write-back
                const int bx = (cb + u) * 32 + lane;
                const int x0 = 2 * bx - ox;
                if (cb + u < c_end && bx < nbx) {
                    const unsigned ia = a[u] < SE_N_MATERIALS ? a[u] : 1u, ib = b[u] < SE_N_MATERIALS ? b[u] : 1u;
                    const unsigned ic = c[u] < SE_N_MATERIALS ? c[u] : 1u, id = d[u] < SE_N_MATERIALS ? d[u] : 1u;
                    const unsigned v = ia | (ib << 8) | (ic << 16) | (id << 24);
                    const unsigned seed = (unsigned)x0 * 461u + rowseed;
                    const unsigned nv = se_block_lut(v, seed, x0, y0, p.frame, tab, pool_off, fat_sm);
                    const unsigned na = nv & 0xFFu, nb = (nv >> 8) & 0xFFu, nc = (nv >> 16) & 0xFFu, nd = nv >> 24;
                    if (CENSUS) {
                        const unsigned cm_cols = (x0 >= 0 ? 5u : 0u) | ((x0 + 1) < p.W ? 10u : 0u);
                        se_census_block(hist_sm, smem_sa + (unsigned)cx.pop_offset, v, a[u], b[u], c[u], d[u], nv, cm_rows & cm_cols);
                    }
                    if (vec_ok) {
                        if (st0 == 0 && (na != a[u] || nb != b[u])) *reinterpret_cast<uint2*>(base0 + x0) = make_uint2(na, nb);
                        if (st1 == 0 && (nc != c[u] || nd != d[u])) *reinterpret_cast<uint2*>(base1 + x0) = make_uint2(nc, nd);
                    } else {
                        const bool cx0 = x0 >= 0, cx1 = (x0 + 1) < p.W;
                        if (st0 == 0 && cx0 && na != a[u]) base0[x0] = na;
                        if (st0 == 0 && cx1 && nb != b[u]) base0[x0 + 1] = nb;
                        if (st1 == 0 && cx0 && nc != c[u]) base1[x0] = nc;
                        if (st1 == 0 && cx1 && nd != d[u]) base1[x0 + 1] = nd;
                    }
                }
            }
        }
    }
    if (CENSUS) {
        __syncthreads();
        for (int i = tid; i < 256; i += blockDim.x) {
            const int dlt = hist_sm[i];
            if (dlt != 0) atomicAdd(cx.census + i, (unsigned long long)(long long)dlt);   // two's complement: negative deltas wrap correctly
        }
    }
}

extern "C" __global__ void __launch_bounds__(SE_K1C_THREADS, SE_K1C_MINCTAS) se_step_lut_global(const SeLutStepParams p) {
    se_k1c_body<false>(p, SeLutCensusParams{}, nullptr);
}

#if SE_EXPERIMENTAL_KERNELS      // compiled only on request (env SE_EXPERIMENTAL_KERNELS=1 when the rules are compiled)
extern "C" __global__ void __launch_bounds__(SE_K1C_THREADS, SE_K1C_MINCTAS) se_step_lut_global_census(const SeLutStepParams p, const SeLutCensusParams cx) {
    __shared__ int hist_sm[256];
    se_k1c_body<true>(p, cx, hist_sm);
}

#ifndef SE_LF_MINCTAS
#define SE_LF_MINCTAS 3
#endif
extern "C" __global__ void __launch_bounds__(256, SE_LF_MINCTAS) se_light_fused(const SeFusedParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ unsigned fat_sm[256];
    __shared__ SeMod mods_sm[256];
    __shared__ int warp_counts[8];
    const int tid = threadIdx.x;
    unsigned smem_sa;
    asm volatile("mov.u32 %0, %1;" : "=r"(smem_sa) : "r"((unsigned)__cvta_generic_to_shared(smem)));
    {
        const uint4* lut4 = reinterpret_cast<const uint4*>(p.lut);
        const int n4 = (p.lut_words + 3) >> 2;
        for (int i = tid; i < n4; i += blockDim.x) {
            const uint4 v = __ldg(lut4 + i);
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(smem_sa + 16u * i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        }
    }
    fat_sm[tid] = se_fat_table[tid];
    __syncthreads();
    const se_tab_t tab = smem_sa;
    float4* term = reinterpret_cast<float4*>(smem + p.tile_offset);
    unsigned char* ids = reinterpret_cast<unsigned char*>(term + SE_LT_TERMS);
    const int n_tiles = p.tiles_x * p.tiles_y;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {        // persistent: the table is staged once per CTA
        const int by = t / p.tiles_x, bx = t - by * p.tiles_x;
        int n_cull = 0;
        if (p.n_mods > 0) {
            // cull the modification list against this tile, keeping the order (last match wins, falling_sand.glsl:764-773)
            const int x_lo = bx * SE_LT_W, y_lo = p.lp.gy0 + by * SE_LT_H;
            bool keep = false;
            SeMod m;
            if (tid < p.n_mods) { m = p.mods[tid]; keep = se_mod_touches(m, x_lo, x_lo + SE_LT_W - 1, y_lo, y_lo + SE_LT_H - 1); }
            const unsigned ballot = __ballot_sync(0xFFFFFFFFu, keep);
            const int warp = tid >> 5, lane = tid & 31;
            if (lane == 0) warp_counts[warp] = __popc(ballot);
            __syncthreads();
            int base = 0;
            for (int w = 0; w < 8; ++w) { if (w < warp) base += warp_counts[w]; n_cull += warp_counts[w]; }
            if (keep) mods_sm[base + __popc(ballot & ((1u << lane) - 1u))] = m;
        }
        if (se_light_tile_is_interior(p.lp, bx, by)) se_fused_stage<true>(p.lp, fat_sm, term, ids, bx, by, tid);
        else se_fused_stage<false>(p.lp, fat_sm, term, ids, bx, by, tid);
        __syncthreads();
        se_fused_blocks(p.lp, p.frame, tab, (unsigned)p.pool_offset, fat_sm, ids, bx, by, tid);
        __syncthreads();
        if (se_light_tile_is_interior(p.lp, bx, by)) {
            if (n_cull) se_light_compute<true>(p.lp, fat_sm, term, bx, by, tid, SeNewIdFused<true>{ids, const_cast<unsigned*>(p.lp.new_cells), mods_sm, n_cull});
            else se_light_compute<true>(p.lp, fat_sm, term, bx, by, tid, SeNewIdFused<false>{ids, const_cast<unsigned*>(p.lp.new_cells), mods_sm, 0});
        } else {
            if (n_cull) se_light_compute<false>(p.lp, fat_sm, term, bx, by, tid, SeNewIdFused<true>{ids, const_cast<unsigned*>(p.lp.new_cells), mods_sm, n_cull});
            else se_light_compute<false>(p.lp, fat_sm, term, bx, by, tid, SeNewIdFused<false>{ids, const_cast<unsigned*>(p.lp.new_cells), mods_sm, 0});
        }
        __syncthreads();                                           // the next tile overwrites term[] / ids[] / mods_sm[]
    }
}

extern "C" __global__ void __launch_bounds__(256) se_build_popbits(unsigned* __restrict__ popbits) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < SE_N4) se_build_popbits_entry(idx, popbits);
}
#endif  // SE_EXPERIMENTAL_KERNELS
#endif  // SE_HOST_EMU
#endif  // SE_LUT_ELIGIBLE
