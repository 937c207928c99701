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

// can record m touch a cell of the rectangle?  (both shapes lie inside the square |dx|,|dy| <= size; negative sizes never
// match; 64-bit sums: a position near INT_MAX must not wrap around)
static __device__ __forceinline__ bool se_mod_touches(const SeMod& m, int x_lo, int x_hi, int y_lo, int y_hi) {
    const long long px = m.px, py = m.py, sz = m.size;
    return m.size >= 0 && px + sz >= x_lo && px - sz <= x_hi && py + sz >= y_lo && py - sz <= y_hi;
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
        // both shapes are contained in the square |dx| <= size, |dy| <= size (negative sizes never match); 64-bit sums: a
        // position near INT_MAX must not wrap around and drop a record the shader's per-cell test would still match
        const long long px = m.px, py = m.py, sz = m.size;
        keep = m.size >= 0 && px + sz >= x_lo && px - sz <= x_hi && py + sz >= y_lo && py - sz <= y_hi;
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
// Transition tables.
//
// Eligible rule sets (SE_LUT_MODE != 0, decided by the code generator): the block transition is a pure function of
// (4 material ids, mirror bit, rand.y) with rand.y used only against literal thresholds -- no pos / frame / other
// rand use -- and at most 127 materials.  Then
//     T[idx],  idx = ((a*N + b)*N + c)*N + d            (N^4 32-bit entries per view)
// holds the transition of every block state, built on the device by se_build_lut from the very rule code generated
// for the generic path (se_block_with_rand).  Entry kinds:
//     bit 31 clear          the result: four id bytes a | b<<8 | c<<16 | d<<24 (ids < 128)
//     SE_E_SPECIAL | k      the result depends on rand.y: pool[k] = {thr, A, B, 0}: result = (u1 <= thr) ? A : B, where A / B
//                           are again entries (B may chain to another pool entry for states with several thresholds)
//     SE_E_SPECIAL|SE_E_SLOW  one-table mode only: the state or its result holds WALL / NULL (guarded swaps,
//                           operations.glsl:16-23): take the generic path (same generated code as K1a)
// One-table mode (no Left/Right rules): the table holds the UNMIRRORED evaluation; the mirrored transition is
// swap_pairs(T[swap_pairs(state)]) -- exact when no cell of the block refuses to swap, which is precisely when the
// table is used.  Two-table mode: entries [N^4, 2 N^4) hold the finished MIRRORED evaluation (guarded swap, mirrored +
// left rules, guarded swap back) of the same state, so there are neither byte swaps nor a slow path.
//   SE_LUT_MODE 1: the table is staged in shared memory (N^4 * tables * 4 B <= 60 KB: the default rule set)
//   SE_LUT_MODE 2: the table stays in global memory (L2 / HBM resident), always two views: up to 127 materials
// =============================================================================================
#if SE_LUT_ELIGIBLE
#define SE_N4 (SE_N_MATERIALS * SE_N_MATERIALS * SE_N_MATERIALS * SE_N_MATERIALS)
#define SE_LUT_ENTRIES (SE_N4 * (SE_LUT_TWO_TABLES ? 2 : 1))
#define SE_E_SPECIAL 0x80000000u
#define SE_E_SLOW 0x40000000u
#define SE_E_INDEX 0x3FFFFFFFu
#define SE_TILE_PW 256          // tile width in cells (= bytes): one warp walks a tile row, 8 cells per lane

static __device__ __forceinline__ unsigned se_ids4(unsigned s, unsigned r, unsigned d, unsigned dr) {
    return SE_ID(s) | (SE_ID(r) << 8) | (SE_ID(d) << 16) | (SE_ID(dr) << 24);
}
// the four ids of a block as a sorted tuple: two blocks hold the same population iff these are equal
static __device__ __forceinline__ unsigned se_sorted4(unsigned a, unsigned b, unsigned c, unsigned d) {
    unsigned t;
    if (a > b) { t = a; a = b; b = t; }
    if (c > d) { t = c; c = d; d = t; }
    if (a > c) { t = a; a = c; c = t; }
    if (b > d) { t = b; b = d; d = t; }
    if (b > c) { t = b; b = c; c = t; }
    return a | (b << 8) | (c << 16) | (d << 24);
}
#define SE_E_POPFLAG 0x80u      // census tables only: bit 7 of the first id byte = "this outcome is not a permutation of the state"

// one block state: evaluate every rand.y class with the generated rule code, then encode.  pool == nullptr: count only.
// flag_pop: build the table of the running census (se_step_lut_global_census): outcomes that change the population of
// the block (a SET fired) carry SE_E_POPFLAG, so the kernel knows from the entry it has read anyway when to look closer.
static __device__ __forceinline__ void se_build_lut_entry(unsigned entry, unsigned* __restrict__ base, unsigned* __restrict__ pool,
                                                          unsigned* __restrict__ counter, unsigned pool_cap, unsigned flag_pop) {
    const unsigned N = SE_N_MATERIALS;
#if SE_LUT_TWO_TABLES
    const bool mirror_table = entry >= (unsigned)SE_N4;
    const unsigned idx = entry - (mirror_table ? (unsigned)SE_N4 : 0u);
#else
    const unsigned idx = entry;
#endif
    const unsigned ia = idx / (N * N * N), ib = (idx / (N * N)) % N, ic = (idx / N) % N, id = idx % N;
    unsigned res[SE_LUT_NCLS];
    // class c <=> rand.y hash lane u1 in (U_{c-1}, U_c]  (U_{-1} = -1, U_{NCLS-1} = 2^32-1)
    bool noswap = ((se_fat_table[ia] | se_fat_table[ib] | se_fat_table[ic] | se_fat_table[id]) & SE_F_NOSWAP) != 0u;
#pragma unroll
    for (int cls = 0; cls < SE_LUT_NCLS; ++cls) {
        unsigned s = se_fat_table[ia], r = se_fat_table[ib], d = se_fat_table[ic], dr = se_fat_table[id];
        if (idx != 0) {                                                   // all-EMPTY early-out (falling_sand.glsl:692-694): T[0] == 0
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
        res[cls] = se_ids4(s, r, d, dr);
        if (flag_pop && se_sorted4(SE_ID(s), SE_ID(r), SE_ID(d), SE_ID(dr)) != se_sorted4(ia, ib, ic, id)) res[cls] |= SE_E_POPFLAG;
    }
#if !SE_LUT_TWO_TABLES
    if (noswap) { base[entry] = SE_E_SPECIAL | SE_E_SLOW; return; }
#else
    (void)noswap;
#endif
    int n_breaks = 0;
#pragma unroll
    for (int cls = 1; cls < SE_LUT_NCLS; ++cls) n_breaks += (res[cls] != res[cls - 1]) ? 1 : 0;
    if (n_breaks == 0) { base[entry] = res[0]; return; }
    const unsigned k0 = atomicAdd(counter, (unsigned)n_breaks);
    base[entry] = SE_E_SPECIAL | (k0 & SE_E_INDEX);
    if (!pool || k0 + (unsigned)n_breaks > pool_cap) return;              // counting pass / overflow: the host looks at the counter
    unsigned k = k0;
    int left = n_breaks;
    for (int cls = 1; cls < SE_LUT_NCLS; ++cls) {
        if (res[cls] != res[cls - 1]) {
            --left;
            pool[4u * k + 0u] = se_lut_thresholds[cls - 1];              // u1 <= U_{cls-1}  <=>  class < cls
            pool[4u * k + 1u] = res[cls - 1];
            pool[4u * k + 2u] = left ? (SE_E_SPECIAL | (k + 1u)) : res[cls];
            pool[4u * k + 3u] = 0u;                                       // 16-byte entries: one 128-bit read
            ++k;
        }
    }
}

#ifndef SE_HOST_EMU
extern "C" __global__ void __launch_bounds__(256) se_build_lut(unsigned* __restrict__ base, unsigned* __restrict__ pool,
                                                               unsigned* __restrict__ counter, unsigned pool_cap, unsigned flag_pop) {
    const unsigned entry = blockIdx.x * blockDim.x + threadIdx.x;
    if (entry < (unsigned)SE_LUT_ENTRIES) se_build_lut_entry(entry, base, pool, counter, pool_cap, flag_pop);
}
#endif

// ---- table access: shared-space addresses (mode 1), global pointers (mode 2), plain pointers on the host ----
#ifdef SE_HOST_EMU
struct SeTab { const unsigned* base; const unsigned* pool; };
static inline unsigned se_tab_entry(const SeTab& t, unsigned idx) { return t.base[idx]; }
static inline void se_tab_pool(const SeTab& t, unsigned k, unsigned& thr, unsigned& a, unsigned& b) { thr = t.pool[4u * k]; a = t.pool[4u * k + 1u]; b = t.pool[4u * k + 2u]; }
static inline unsigned se_dp4a(unsigned x, unsigned w) {
    return (x & 0xFFu) * (w & 0xFFu) + ((x >> 8) & 0xFFu) * ((w >> 8) & 0xFFu) + ((x >> 16) & 0xFFu) * ((w >> 16) & 0xFFu) + (x >> 24) * (w >> 24);
}
#else
static __device__ __forceinline__ unsigned se_lds_u8(unsigned a) { unsigned v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
static __device__ __forceinline__ unsigned se_lds_u16(unsigned a) { unsigned v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
static __device__ __forceinline__ unsigned se_lds_u32(unsigned a) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
static __device__ __forceinline__ void se_lds_u64(unsigned a, unsigned& x, unsigned& y) { asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(a)); }
static __device__ __forceinline__ void se_sts_u8(unsigned a, unsigned v) { asm volatile("st.shared.u8 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
static __device__ __forceinline__ void se_sts_u16(unsigned a, unsigned v) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
static __device__ __forceinline__ void se_sts_u32(unsigned a, unsigned v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
static __device__ __forceinline__ void se_sts_u64(unsigned a, unsigned x, unsigned y) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(a), "r"(x), "r"(y) : "memory"); }
static __device__ __forceinline__ unsigned se_dp4a(unsigned x, unsigned w) { return __dp4a(x, w, 0u); }
struct SeTabG { const unsigned* base; const unsigned* pool; };   // table in global memory, read-only for the whole launch
static __device__ __forceinline__ unsigned se_tab_entry(const SeTabG& t, unsigned idx) {
    unsigned v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(t.base + idx));
    return v;
}
static __device__ __forceinline__ void se_tab_pool(const SeTabG& t, unsigned k, unsigned& thr, unsigned& a, unsigned& b) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(t.pool) + k);
    thr = q.x; a = q.y; b = q.z;
}
#if SE_LUT_MODE == 1
struct SeTabS { unsigned base; unsigned pool; };            // table staged in shared memory: shared-space byte addresses
static __device__ __forceinline__ unsigned se_tab_entry(const SeTabS& t, unsigned idx) { return se_lds_u32(t.base + 4u * idx); }
static __device__ __forceinline__ void se_tab_pool(const SeTabS& t, unsigned k, unsigned& thr, unsigned& a, unsigned& b) {
    unsigned pad;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(thr), "=r"(a), "=r"(b), "=r"(pad) : "r"(t.pool + 16u * k));
}
typedef SeTabS SeTab;
#else
typedef SeTabG SeTab;
#endif
#endif

// idx = ((a*N + b)*N + c)*N + d from the four id bytes: two byte dot products (weights <= 255 for every N)
static __device__ __forceinline__ unsigned se_idx4(unsigned v) {
    const unsigned w = (unsigned)SE_N_MATERIALS | (1u << 8);
    return se_dp4a(v, w) * (unsigned)(SE_N_MATERIALS * SE_N_MATERIALS) + se_dp4a(v, w << 16);
}

// The rare ways out of the table: rand.y-dependent states walk their pool chain, slow states run the generated code.
// `e` is the special entry, `v` the block as stored (unpermuted); the result is returned in the TABLE's view (the
// caller undoes the mirror permutation in one-table mode, so the slow path applies it once more: an involution).
template <class Tab>
#ifndef SE_HOST_EMU
static __device__ __noinline__ unsigned se_block_special(unsigned e, unsigned v, unsigned seed, unsigned mirror_sel, int px, int py, int frame,
                                                         const Tab tab, const unsigned* __restrict__ fat_sm)
#else
static inline unsigned se_block_special(unsigned e, unsigned v, unsigned seed, unsigned mirror_sel, int px, int py, int frame,
                                        const Tab tab, const unsigned* fat_sm)
#endif
{
    const unsigned u1 = se_hashi(seed * 2131u);
#if !SE_LUT_TWO_TABLES
    if (e & SE_E_SLOW) {
        unsigned s = fat_sm[v & 0xFFu], r = fat_sm[(v >> 8) & 0xFFu], d = fat_sm[(v >> 16) & 0xFFu], dr = fat_sm[v >> 24];
        SeRand rnd;
        rnd.u[0] = (mirror_sel == 0x2301u) ? 0u : 0xFFFFFFFFu;
        rnd.u[1] = u1; rnd.u[2] = 0u; rnd.u[3] = 0u;
        se_block_with_rand(s, r, d, dr, rnd, px, py, frame);
        return __byte_perm(se_ids4(s, r, d, dr), 0u, mirror_sel);
    }
#else
    (void)v; (void)mirror_sel; (void)px; (void)py; (void)frame; (void)fat_sm;
#endif
    do {
        unsigned thr, a, b;
        se_tab_pool(tab, e & SE_E_INDEX, thr, a, b);
        e = (u1 <= thr) ? a : b;
    } while (e & SE_E_SPECIAL);
    return e;
}

// one block: v = a | b<<8 | c<<16 | d<<24 (material ids < N), returns the new ids in the same packing.
// No branch for the all-EMPTY early-out (falling_sand.glsl:692-694): T[0] == 0 in both views by construction.
template <class Tab>
static __device__ __forceinline__ unsigned se_block_lut(unsigned v, unsigned seed, int px, int py, int frame, const Tab tab,
                                                        const unsigned* __restrict__ fat_sm) {
    const unsigned u0 = se_hashi(seed * 213u);
    const bool mirror = u0 <= SE_MIRROR_UMAX;
#if SE_LUT_TWO_TABLES
    unsigned e = se_tab_entry(tab, se_idx4(v) + (mirror ? (unsigned)SE_N4 : 0u));
    if (e & SE_E_SPECIAL) e = se_block_special(e, v, seed, 0x3210u, px, py, frame, tab, fat_sm);
    return e;
#else
    const unsigned sel = mirror ? 0x2301u : 0x3210u;
    unsigned e = se_tab_entry(tab, se_idx4(__byte_perm(v, 0u, sel)));
    if (e & SE_E_SPECIAL) e = se_block_special(e, v, seed, sel, px, py, frame, tab, fat_sm);
    return __byte_perm(e, 0u, sel);
#endif
}

static __device__ __forceinline__ unsigned se_clamp_id(unsigned id) { return id < SE_N_MATERIALS ? id : 1u; }   // unknown ids read as NULL (gen/materials.glsl:79-86)
static __device__ __forceinline__ unsigned se_pack_ids(uint4 v) {
    return se_clamp_id(v.x) | (se_clamp_id(v.y) << 8) | (se_clamp_id(v.z) << 16) | (se_clamp_id(v.w) << 24);
}

// ---------------------------------------------------------------------------------------------
// Running census (SE_FLAG_RUNNING_CENSUS, one-table mode): the per-material population of the owned rows is kept up
// to date by the per-frame kernel K1c instead of being recounted by a pass over the grid.
// Guarded swaps only permute the cells of a block, so the population changes only where a SET fired (or where a
// block straddles the first/last owned row of a strip, or holds an id the table does not know).  The census variant of
// K1c reads a copy of the table whose outcomes carry SE_E_POPFLAG when they are not a permutation of the state (the
// generated-code path for WALL / NULL blocks always counts as flagged): one bit test per block, and only flagged or
// partly counted blocks are compared cell by cell.
// ---------------------------------------------------------------------------------------------
#if !SE_LUT_TWO_TABLES
// One block of K1c: r* = the words read from memory, nv = new ids (flag removed), cm = bit k set when cell k (a, b, c, d)
// is inside the grid AND in an owned row, `look` = the outcome was flagged / generated code / a raw id was unknown.
// Adds the population deltas of the counted cells to hist[256] (CTA-local, flushed once per launch).
static __device__ __forceinline__ void se_census_block(int* hist, bool look, unsigned ra, unsigned rb, unsigned rc, unsigned rd, unsigned nv, unsigned cm) {
    if (cm == 0u || (cm == 0xFu && !look)) return;      // all four cells counted and the outcome is a permutation: nothing to do
    const unsigned na = nv & 0xFFu, nb = (nv >> 8) & 0xFFu, nc = (nv >> 16) & 0xFFu, nd = nv >> 24;
    if ((cm & 1u) && na != ra) { atomicAdd(hist + (ra < 255u ? ra : 255u), -1); atomicAdd(hist + na, 1); }
    if ((cm & 2u) && nb != rb) { atomicAdd(hist + (rb < 255u ? rb : 255u), -1); atomicAdd(hist + nb, 1); }
    if ((cm & 4u) && nc != rc) { atomicAdd(hist + (rc < 255u ? rc : 255u), -1); atomicAdd(hist + nc, 1); }
    if ((cm & 8u) && nd != rd) { atomicAdd(hist + (rd < 255u ? rd : 255u), -1); atomicAdd(hist + nd, 1); }
}
#endif  // !SE_LUT_TWO_TABLES

#ifndef SE_HOST_EMU
// =============================================================================================
// K1b: transition table + shared-memory tiles + temporal blocking + (strips) the ghost-row push.
//
// Launch shape: one CTA of 1024 threads per SM.  The CTA stages the table once (mode 1) and then runs as TWO
// independent halves of 16 warps, each with its own tile buffer and its own named barrier: while one half waits
// for HBM (tile load / store) the other one computes, and both share the one copy of the table.
// A tile is 256 cells wide (one byte per cell in shared memory), PH rows high.  Per Margolus sub-step a warp walks
// block rows; a lane owns 8 consecutive cells of both rows of the block row (two 8-byte shared loads), i.e. four
// blocks in the even column phase; in the odd phase its four blocks are shifted by one cell and the straddling byte
// pair travels through one shuffle each way.  The block itself is: lowbias32 hash -> mirror bit -> byte permute ->
// two byte dot products -> one table read -> permute back (se_block_lut).
//
// Work items (k, t) = (T-block, tile), numbered k * n_tiles + position(t), are drawn by the 2 x gridDim.x halves from
// one device-wide queue (atomicAdd), in order.  A tile of T-block k needs the
// results of T-block k-1 in its 3x3 tile neighbourhood only, so instead of a grid-wide barrier (or a launch)
// between T-blocks every tile publishes a sequence number when its interior is stored and consumers wait on the
// flags they need.  Deadlock-free because every dependency has a smaller index and all CTAs are co-resident
// (cooperative launch, grid <= SMs).
//
// Strips: the tiles cover the rows this device OWNS.  A tile in the first / last tile row also
// stores the `push_rows` rows next to the strip boundary into the neighbour device's ghost rows (peer st.global
// over NVLink, same ping-pong parity) and publishes its sequence number in the neighbour's memory (st.release.sys);
// the neighbour's boundary tiles acquire those flags exactly like the local ones, which also covers the
// write-after-read hazard on the ghost rows.  The boundary tile rows are dealt first, so the push overlaps the
// interior of the same T-block.  No host synchronisation and no separate copy: the exchange is part of the store.
// A wait that outlives `spin_limit` polls (a neighbour that was never launched) raises status[0] and lets the
// kernel run to its end instead of hanging the device; the host reports it.
// =============================================================================================
struct SeTileParams {
    unsigned* buf0;         // local buffer (row 0 == global row gy0) read by the first T-block of the launch
    unsigned* buf1;         // the other ping-pong buffer; T-block k reads buf[k & 1] and writes buf[(k + 1) & 1]
    unsigned* nbr_buf0[2];  // [0] strip above, [1] strip below: the neighbour's buffers of the same parity (null: none)
    unsigned* nbr_buf1[2];
    unsigned* nbr_flags[2]; // per tile column, in the NEIGHBOUR's memory: sequence numbers of my boundary tiles
    const unsigned* in_flags[2];   // per tile column, in MY memory: sequence numbers of the neighbours' boundary tiles
    int nbr_gy0[2];         // global row of row 0 of the neighbour's buffers
    int W, Hl, gy0, Hg;
    int own_y0, own_y1;     // owned global rows: the tiles' interiors partition [own_y0, own_y1)
    int frame0;             // frame number of the first sub-step of this launch
    int nblk;               // T-blocks in this launch (>= 1)
    int tsteps;             // sub-steps of every T-block but the last
    int nsub_last;          // sub-steps of the last T-block, 1..tsteps
    unsigned seq_base;      // T-blocks completed by earlier launches (the flags count in this sequence)
    unsigned* done;         // per tile: sequence number of its last completed T-block
    unsigned* status;       // [0] != 0: a flag wait timed out (results invalid)
    unsigned* queue;        // work queue: the halves draw item numbers with atomicAdd (in order: dependencies always have a smaller number)
    unsigned queue_base;    // value of *queue when this launch starts (it is never reset: every half overshoots exactly once)
    int HY;                 // halo depth in rows (even, >= nsub/2 + 1: the row offset changes every OTHER frame,
                            // operations.glsl:25-34, so validity shrinks by at most floor(n/2)+1 rows in n steps)
    int HX;                 // halo depth in columns (multiple of 4, >= nsub: the column offset alternates every frame)
    int PH;                 // tile height in cells (even) = THo + 2 HY
    int THo;                // interior rows per tile (even)
    int tiles_x, tiles_y;
    int push_rows;          // rows pushed into a neighbour's ghost rows (>= HY of the neighbour)
    int table_bytes;        // mode 1: bytes of (table + pool) to stage in shared memory
    int pool_offset;        // mode 1: byte offset of the pool inside the staged image
    int tile_offset;        // byte offset of the first tile buffer in dynamic shared memory (16-aligned)
    int tile_stride;        // bytes between the two halves' tile buffers
    const unsigned* lut;    // table image in global memory (mode 1: staged; mode 2: read in place)
    const unsigned* pool;   // mode 2: pool in global memory
    unsigned spin_limit;
};

#define SE_HALF_WARPS 16
// The sub-step loops are bound by the ALU pipe (LOP3 / SHF / ISETP / PRMT / SEL issue every other cycle per scheduler),
// the FMA pipe (IMAD) idles: wherever it costs nothing, integer work is phrased as a multiply-add.
// (seed + offset) * factor as ONE multiply-add per block instead of one multiply per row and an ALU add per block:
static __device__ __forceinline__ unsigned se_seed_mul(unsigned seed, unsigned offset, unsigned factor) {
    unsigned r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(seed), "r"(factor), "r"(offset * factor));
    return r;
}
// rand.x < 0.5 from the hash lane WITHOUT its last xor-shift.  With y the value before `x ^= x >> 16`:
//   u0 = y ^ (y >> 16) <= SE_MIRROR_UMAX = 2^31 - 65   <=>   y < 2^31  and not  0x7FFF8000 <= y <= 0x7FFF803F
// (the high half-word of u0 is that of y; for y >> 16 == 0x7FFF the low half-word of u0 is that of y xor 0x7FFF, and
// u0 >= 0x7FFFFFC0 means that low half-word lies in [0xFFC0, 0xFFFF], i.e. y's in [0x8000, 0x803F]).  Two compares
// instead of shift + xor + compare; checked exhaustively around every boundary by tests/test_codegen_host_emulation.py.
static __device__ __forceinline__ bool se_mirror_bit(unsigned x) {
#if SE_MIRROR_UMAX == 0x7fffffbfu
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU;
    // y < 0x7FFF8000 (unsigned)  or  y > 0x7FFF803F as a SIGNED number (i.e. 0x7FFF8040 .. 0x7FFFFFFF): two chained SETPs
    unsigned r;
    asm("{\n\t.reg .pred p, q;\n\tsetp.lt.u32 p, %1, 0x7FFF8000;\n\tsetp.gt.or.s32 q, %1, 0x7FFF803F, p;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(r) : "r"(x));
    return r != 0u;
#else
    return se_hashi(x) <= SE_MIRROR_UMAX;
#endif
}
#ifndef SE_LOAD_ROWS
#define SE_LOAD_ROWS 2          // tile rows whose 128-bit loads a lane keeps in flight together (registers: 8 per row)
#endif
#define SE_HALF_THREADS (32 * SE_HALF_WARPS)
static __device__ __forceinline__ void se_half_sync(int half) { asm volatile("bar.sync %0, %1;" :: "r"(half + 1), "n"(SE_HALF_THREADS) : "memory"); }

// one Margolus sub-step over the whole tile; OX (column phase) is a template parameter: separate straight-line loops
// one Margolus sub-step over the whole tile; OX (column phase) is a template parameter: separate straight-line loops
template <int OX>
static __device__ __forceinline__ void se_tile_substep(unsigned tile_sa, const SeTab tab, const unsigned* __restrict__ fat_sm,
                                                       int PH, int oy, int frame, int gx_org, int gy_org, int W, int Hg, int hw, int lane) {
    // block rows whose two cell rows both lie outside the grid hold WALL only (re-imposed after every sub-step): skipped
    const int j_lo = max(0, (-gy_org - oy) >> 1);                      // first j with gy_org + 2j + oy >= -1
    const int j_hi = min((PH - oy) >> 1, ((Hg - 1 - gy_org - oy) >> 1) + 1);   // one past the last j with gy_org + 2j + oy <= Hg - 1
    const int lane_x = gx_org + 8 * lane + OX;                         // global x of the left cell of the lane's first block
    // lanes whose four blocks lie outside the grid (the unused part of the last tile column) only take part in the shuffles
    const bool lane_in = lane_x + 7 >= 0 && lane_x < W;
    const unsigned fterm = (unsigned)frame * (2131u * 2131u) + (unsigned)lane_x * 461u;
    for (int j = j_lo + hw; j < j_hi; j += SE_HALF_WARPS) {
        const int ly = 2 * j + oy;
        const unsigned rowseed = (unsigned)(gy_org + ly) * 2131u + fterm;
        const unsigned a0 = tile_sa + (unsigned)(ly * SE_TILE_PW + 8 * lane);
        unsigned X0, Y0, X1, Y1;
        se_lds_u64(a0, X0, Y0);
        se_lds_u64(a0 + SE_TILE_PW, X1, Y1);
        unsigned v[4], n[4];
        if (OX == 0) {
            v[0] = __byte_perm(X0, X1, 0x5410); v[1] = __byte_perm(X0, X1, 0x7632);
            v[2] = __byte_perm(Y0, Y1, 0x5410); v[3] = __byte_perm(Y0, Y1, 0x7632);
        } else {
            // blocks at cells (1,2) (3,4) (5,6) (7, 8 = the next lane's cell 0)
            const unsigned q = __shfl_down_sync(0xFFFFFFFFu, __byte_perm(X0, X1, 0x0040), 1);
            const unsigned S0 = __funnelshift_r(X0, Y0, 24), S1 = __funnelshift_r(X1, Y1, 24);
            v[0] = __byte_perm(X0, X1, 0x6521);
            v[1] = __byte_perm(S0, S1, 0x5410); v[2] = __byte_perm(S0, S1, 0x7632);
            v[3] = __byte_perm(__byte_perm(Y0, Y1, 0x0703), q, 0x5240);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) n[k] = v[k];
        if (lane_in) {
            // the four table reads are issued back to back; the entries that need more than the table (rand.y-dependent
            // states, WALL / NULL blocks in one-table mode) are picked up afterwards behind ONE test.
            // (Measured and rejected: deferring those blocks to a per-warp list in shared memory and working them off 32 at a
            // time -- the ballots and list bookkeeping cost more than the divergence: 15.5 ms against 13.0 ms per 64 steps.)
            unsigned e[4], mir = 0u;                                   // mir bit k: block k is evaluated in the mirrored view
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const bool mirror = se_mirror_bit(se_seed_mul(rowseed, (unsigned)(2 * k) * 461u, 213u));
                mir |= mirror ? (1u << k) : 0u;
#if SE_LUT_TWO_TABLES
                e[k] = se_tab_entry(tab, se_idx4(v[k]) + (mirror ? (unsigned)SE_N4 : 0u));
#else
                e[k] = se_tab_entry(tab, se_idx4(__byte_perm(v[k], 0u, mirror ? 0x2301u : 0x3210u)));
#endif
            }
            if ((e[0] | e[1] | e[2] | e[3]) & SE_E_SPECIAL) {
                // rand.y-dependent states: the second hash lane, then the pool chain (one 128-bit read per threshold), inline
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if ((e[k] & (SE_E_SPECIAL | SE_E_SLOW)) == SE_E_SPECIAL) {
                        const unsigned u1 = se_hashi(se_seed_mul(rowseed, (unsigned)(2 * k) * 461u, 2131u));
                        do {
                            unsigned thr, a, b;
                            se_tab_pool(tab, e[k] & SE_E_INDEX, thr, a, b);
                            e[k] = (u1 <= thr) ? a : b;
                        } while (e[k] & SE_E_SPECIAL);
                    }
#if !SE_LUT_TWO_TABLES
                // what is left in one-table mode: blocks that hold WALL / NULL (generated code, out of line; table-eligible
                // rule sets never read pos / frame, so the generated code gets zeros for them)
                if ((e[0] | e[1] | e[2] | e[3]) & SE_E_SPECIAL) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (e[k] & SE_E_SPECIAL)
                            e[k] = se_block_special(e[k], v[k], rowseed + (unsigned)(2 * k) * 461u, ((mir >> k) & 1u) ? 0x2301u : 0x3210u, 0, 0, 0, tab, fat_sm);
                }
#endif
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) n[k] = SE_LUT_TWO_TABLES ? e[k] : __byte_perm(e[k], 0u, ((mir >> k) & 1u) ? 0x2301u : 0x3210u);
        }
        if (OX == 0) {
            se_sts_u64(a0, __byte_perm(n[0], n[1], 0x5410), __byte_perm(n[2], n[3], 0x5410));
            se_sts_u64(a0 + SE_TILE_PW, __byte_perm(n[0], n[1], 0x7632), __byte_perm(n[2], n[3], 0x7632));
        } else {
            // cell 0 of the lane is cell 8 of the lane to the left.  Lane 0 / lane 31 have no partner: their edge
            // cells take a (valid-id) byte that is never used -- the outermost tile columns are inside the halo that
            // an odd-phase sub-step invalidates anyway.
            const unsigned ld = __shfl_up_sync(0xFFFFFFFFu, n[3], 1);
            se_sts_u64(a0, __byte_perm(__byte_perm(n[0], n[1], 0x4100), ld, 0x3215), __byte_perm(__byte_perm(n[1], n[2], 0x0541), n[3], 0x4210));
            se_sts_u64(a0 + SE_TILE_PW, __byte_perm(__byte_perm(n[0], n[1], 0x6320), ld, 0x3217), __byte_perm(__byte_perm(n[1], n[2], 0x0763), n[3], 0x6210));
        }
    }
}

// wait until *flag has reached `want`; false (and status raised) when the wait is abandoned
static __device__ __forceinline__ bool se_wait_flag(const unsigned* flag, unsigned want, bool sys_scope, unsigned* status, unsigned spin_limit) {
    unsigned spins = 0;
    while (true) {
        unsigned v;
        if (sys_scope) asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        else asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int)(v - want) >= 0) return true;
        if ((++spins & 63u) == 0u) {
            unsigned st;
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(st) : "l"(status) : "memory");
            if (st != 0u || spins > spin_limit) { atomicExch(status, 1u); return false; }
        }
        __nanosleep(spins < 64u ? 40 : 400);
    }
}

extern "C" __global__ void __launch_bounds__(2 * SE_HALF_THREADS, 1) se_step_tiles(const SeTileParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ unsigned fat_sm[256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int half = warp / SE_HALF_WARPS, hw = warp - half * SE_HALF_WARPS, htid = tid - half * SE_HALF_THREADS;
    unsigned smem_sa;   // laundered through asm so the compiler keeps it in a register instead of re-deriving it per access
    asm volatile("mov.u32 %0, %1;" : "=r"(smem_sa) : "r"((unsigned)__cvta_generic_to_shared(smem)));
    const unsigned tile_sa = smem_sa + (unsigned)p.tile_offset + (unsigned)half * (unsigned)p.tile_stride;
#if SE_LUT_MODE == 1
    {   // stage table + pool with 16-byte loads (the image is padded to a multiple of 16 bytes)
        const uint4* lut4 = reinterpret_cast<const uint4*>(p.lut);
        const int n4 = (p.table_bytes + 15) >> 4;
        for (int i = tid; i < n4; i += blockDim.x) {
            const uint4 v = __ldg(lut4 + i);
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(smem_sa + 16u * i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        }
    }
    const SeTab tab{smem_sa, smem_sa + (unsigned)p.pool_offset};
#else
    const SeTab tab{p.lut, p.pool};
#endif
    if (tid < 256) fat_sm[tid] = se_fat_table[tid];
    __syncthreads();

    const int PW = SE_TILE_PW, PH = p.PH;
    const int TWo = PW - 2 * p.HX;
    const int n_tiles = p.tiles_x * p.tiles_y;
    const bool has_above = p.nbr_flags[0] != nullptr, has_below = p.nbr_flags[1] != nullptr;
    const unsigned total_items = (unsigned)p.nblk * (unsigned)n_tiles;
    __shared__ unsigned next_item[2];
    while (true) {
        // Work items are drawn from one device-wide queue, in order: a half that was held up (a tile at the grid's edge, a
        // slow flag) does not hold up the tiles dealt to it in advance, and the two halves of a CTA drift apart so that one
        // computes while the other waits for HBM.
        if (htid == 0) next_item[half] = atomicAdd(p.queue, 1u) - p.queue_base;
        se_half_sync(half);
        const unsigned w = next_item[half];
        if (w >= total_items) {
            // Strips: the half that draws the first number past the end keeps the kernel alive until the neighbours' pushes of
            // the LAST T-block have landed in this strip's ghost rows (their boundary tiles come first in every T-block, so
            // this is rarely a wait).  A kernel that runs after this one -- a per-step kernel, which looks at no flags -- then
            // finds the ghost rows of the final buffer complete.
            if (w == total_items && (has_above || has_below)) {
                const unsigned want = p.seq_base + (unsigned)p.nblk;
                for (int i = htid; i < 2 * p.tiles_x; i += SE_HALF_THREADS) {
                    const int side = i / p.tiles_x, ntx = i - side * p.tiles_x;
                    if (side == 0 ? has_above : has_below) se_wait_flag(p.in_flags[side] + ntx, want, true, p.status, p.spin_limit);
                }
            }
            break;
        }
        const int k = (int)(w / (unsigned)n_tiles), pos = (int)(w - (unsigned)k * (unsigned)n_tiles);
        // Strips: tile rows from the outside in (0, last, 1, last - 1, ...).  The boundary rows come first, so their push
        // overlaps the interior of the same T-block and the neighbours never wait for the end of a T-block; and a tile's
        // dependencies (its own row and the two next to it, one T-block earlier) are at most two rows behind it in this
        // order, i.e. nearly a whole T-block of items ahead in the queue.  (Rows top to bottom would make the strip below
        // wait for the LAST row of every T-block; "boundary rows first, then top to bottom" would make the own last row wait
        // for its upper neighbour, the last one drawn: 0.5 rounds of stall per T-block at 8 GPUs.)
        const int ry = pos / p.tiles_x, tx = pos - ry * p.tiles_x;
        const int ty = (has_above || has_below) ? ((ry & 1) ? p.tiles_y - 1 - (ry >> 1) : (ry >> 1)) : ry;
        const int t = ty * p.tiles_x + tx;
        unsigned* in = (k & 1) ? p.buf1 : p.buf0;
        unsigned* out = (k & 1) ? p.buf0 : p.buf1;
        const int nsub = (k == p.nblk - 1) ? p.nsub_last : p.tsteps;
        const int tframe0 = p.frame0 + k * p.tsteps;
        const bool top_row = ty == 0, bottom_row = ty == p.tiles_y - 1;
        // ---- dependencies: 3x3 tile neighbourhood of T-block k-1, across the strip boundary too ----
        if (htid < 15) {
            const unsigned want = p.seq_base + (unsigned)k;
            if (htid < 9) {
                const int ntx = tx + (htid % 3) - 1, nty = ty + (htid / 3) - 1;
                if (k > 0 && ntx >= 0 && ntx < p.tiles_x && nty >= 0 && nty < p.tiles_y)
                    se_wait_flag(p.done + (nty * p.tiles_x + ntx), want, false, p.status, p.spin_limit);
            } else {
                // (k == 0 waits for nothing, here as above: every launch ends only when the neighbours' boundary tiles of its
                // last T-block are complete -- the final wait below -- so what the first T-block reads has been delivered and
                // what it overwrites has been read.  The flags of an earlier launch may belong to another tile geometry.)
                const int side = (htid - 9) / 3, ntx = tx + ((htid - 9) % 3) - 1;
                const bool need = k > 0 && (side == 0 ? (has_above && top_row) : (has_below && bottom_row));
                if (need && ntx >= 0 && ntx < p.tiles_x)
                    se_wait_flag(p.in_flags[side] + ntx, want, true, p.status, p.spin_limit);
            }
        }
        se_half_sync(half);
        const int gx_org = tx * TWo - p.HX;                 // global x of tile column 0 (multiple of 4)
        const int y_int0 = p.own_y0 + ty * p.THo;           // first interior row (global, even)
        const int gy_org = y_int0 - p.HY;                   // global y of tile row 0 (even)
        const int y_int1 = min(y_int0 + p.THo, p.own_y1);   // end of the interior rows
        const bool border = gx_org < 0 || gx_org + PW > p.W || gy_org < 0 || gy_org + PH > p.Hg;

        // ---- load: 2 x uint4 of packed-u32 cells -> 8 id bytes per lane; SE_LOAD_ROWS rows (two 128-bit loads each) in flight per lane ----
        for (int r0 = hw; r0 < PH; r0 += SE_LOAD_ROWS * SE_HALF_WARPS) {
            uint4 q0[SE_LOAD_ROWS], q1[SE_LOAD_ROWS];
            bool ok0[SE_LOAD_ROWS], ok1[SE_LOAD_ROWS];
            const int gx = gx_org + 8 * lane;
#pragma unroll
            for (int u = 0; u < SE_LOAD_ROWS; ++u) {
                const int r = r0 + u * SE_HALF_WARPS;
                const int gy = gy_org + r, lr = gy - p.gy0;
                const bool row_ok = r < PH && gy >= 0 && gy < p.Hg && lr >= 0 && lr < p.Hl;
                const uint4* src = reinterpret_cast<const uint4*>(in + (size_t)(row_ok ? lr : 0) * p.W);
                ok0[u] = row_ok && gx >= 0 && gx < p.W;
                ok1[u] = row_ok && gx + 4 >= 0 && gx + 4 < p.W;
                // __ldcg (L2 only): the data may have been written by another SM (or a peer device) earlier in this very launch
                if (ok0[u]) q0[u] = __ldcg(src + (gx >> 2));
                if (ok1[u]) q1[u] = __ldcg(src + ((gx + 4) >> 2));
            }
#pragma unroll
            for (int u = 0; u < SE_LOAD_ROWS; ++u) {
                const int r = r0 + u * SE_HALF_WARPS;
                if (r < PH)                                  // WALL outside the grid (operations.glsl:45-51)
                    se_sts_u64(tile_sa + (unsigned)(r * PW + 8 * lane), ok0[u] ? se_pack_ids(q0[u]) : 0x02020202u, ok1[u] ? se_pack_ids(q1[u]) : 0x02020202u);
            }
        }
        se_half_sync(half);

        // ---- nsub Margolus sub-steps in shared memory ----
        for (int sub = 0; sub < nsub; ++sub) {
            const int frame = tframe0 + sub;
            int ox, oy;
            se_margolus_offset(frame, ox, oy);
            if (ox == 0) se_tile_substep<0>(tile_sa, tab, fat_sm, PH, oy, frame, gx_org, gy_org, p.W, p.Hg, hw, lane);
            else se_tile_substep<1>(tile_sa, tab, fat_sm, PH, oy, frame, gx_org, gy_org, p.W, p.Hg, hw, lane);
            se_half_sync(half);
            if (border) {
                // cells outside the grid are WALL at every step, whatever a SET wrote into them
                for (int r = hw; r < PH; r += SE_HALF_WARPS) {
                    const int gy = gy_org + r;
                    const bool row_out = gy < 0 || gy >= p.Hg;
                    const int gx = gx_org + 8 * lane;
                    if (row_out || gx < 0 || gx >= p.W) se_sts_u32(tile_sa + (unsigned)(r * PW + 8 * lane), 0x02020202u);
                    if (row_out || gx + 4 < 0 || gx + 4 >= p.W) se_sts_u32(tile_sa + (unsigned)(r * PW + 8 * lane + 4), 0x02020202u);
                }
                se_half_sync(half);
            }
        }

        // ---- store the interior: 8 id bytes -> 2 x uint4 of packed-u32 cells; boundary rows also into the neighbours' ghost rows ----
        unsigned* nb_out[2];
        nb_out[0] = (has_above && top_row) ? ((k & 1) ? p.nbr_buf0[0] : p.nbr_buf1[0]) : nullptr;
        nb_out[1] = (has_below && bottom_row) ? ((k & 1) ? p.nbr_buf0[1] : p.nbr_buf1[1]) : nullptr;
        for (int r = p.HY + hw; r < PH - p.HY; r += SE_HALF_WARPS) {
            const int gy = gy_org + r;
            if (gy >= y_int1) break;
            const int gx = gx_org + 8 * lane;
            unsigned w0, w1;
            se_lds_u64(tile_sa + (unsigned)(r * PW + 8 * lane), w0, w1);
            const bool ok0 = 8 * lane >= p.HX && 8 * lane < PW - p.HX && gx < p.W;
            const bool ok1 = 8 * lane + 4 >= p.HX && 8 * lane + 4 < PW - p.HX && gx + 4 < p.W;
            const uint4 c0 = make_uint4(w0 & 0xFFu, (w0 >> 8) & 0xFFu, (w0 >> 16) & 0xFFu, w0 >> 24);
            const uint4 c1 = make_uint4(w1 & 0xFFu, (w1 >> 8) & 0xFFu, (w1 >> 16) & 0xFFu, w1 >> 24);
            uint4* dst = reinterpret_cast<uint4*>(out + (size_t)(gy - p.gy0) * p.W);
            if (ok0) dst[gx >> 2] = c0;
            if (ok1) dst[(gx + 4) >> 2] = c1;
            if (nb_out[0] && gy < p.own_y0 + p.push_rows) {
                uint4* nd = reinterpret_cast<uint4*>(nb_out[0] + (size_t)(gy - p.nbr_gy0[0]) * p.W);
                if (ok0) nd[gx >> 2] = c0;
                if (ok1) nd[(gx + 4) >> 2] = c1;
            }
            if (nb_out[1] && gy >= p.own_y1 - p.push_rows) {
                uint4* nd = reinterpret_cast<uint4*>(nb_out[1] + (size_t)(gy - p.nbr_gy0[1]) * p.W);
                if (ok0) nd[gx >> 2] = c0;
                if (ok1) nd[(gx + 4) >> 2] = c1;
            }
        }
        se_half_sync(half);                               // every thread's stores are issued ...
        if (htid == 0) {                                  // ... and made visible before the tile is published
            const unsigned seq = p.seq_base + (unsigned)k + 1u;
            if ((has_above && top_row) || (has_below && bottom_row)) {
                __threadfence_system();
                if (has_above && top_row) asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p.nbr_flags[0] + tx), "r"(seq) : "memory");
                if (has_below && bottom_row) asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p.nbr_flags[1] + tx), "r"(seq) : "memory");
            } else {
                __threadfence();
            }
            asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p.done + t), "r"(seq) : "memory");
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K1c: ONE Margolus step with the transition table, straight from/to global memory, in place.
// This is the per-frame path (`Simulation::run()` called once per frame): a single step cannot amortise the
// tile load/store of K1b, so blocks are read directly (two 8-byte loads per block when the column phase is
// even, four 4-byte loads otherwise), pushed through se_block_lut and only the cells that changed are
// written back.  Persistent CTAs keep the table in shared memory (mode 1); a warp walks 32 consecutive blocks of
// one block row, a CTA walks block rows grid-stride.  HBM-bound: ~4 B read + (changed fraction) x 4 B
// written per cell.
// ---------------------------------------------------------------------------------------------
struct SeLutStepParams {
    unsigned* cells;        // local buffer, updated in place
    int W, Hl, gy0, Hg;
    int frame;
    int table_bytes, pool_offset;
    const unsigned* lut;
    const unsigned* pool;
    int n_mods;             // the _mods kernels: already cut at the first mod_size == 0 (falling_sand.glsl:754-756)
    const SeMod* mods;
};

#define SE_K1C_THREADS 512
#define SE_K1C_SPAN 8           // a work item = one block row x SPAN chunks of 32 blocks (amortises the row set-up)
#ifndef SE_K1C_BATCH
#define SE_K1C_BATCH 4          // chunks whose loads are issued together (the update is in place, so the compiler
#endif                          // cannot hoist loads over the previous chunk's stores by itself)
#ifndef SE_K1C_MINCTAS
#define SE_K1C_MINCTAS 2
#endif

// running census: extra arguments of se_step_lut_global_census
struct SeLutCensusParams {
    unsigned long long* census;   // 256 bins: population of the owned rows, updated in place
    int own_y0, own_y1;           // owned global rows [own_y0, own_y1)
};

// MODS: the frame's modification records override the cells they cover (falling_sand.glsl:749-794) -- a brush held down keeps the
// per-frame path on the table kernel.  A warp tests the records against the bounding box of its work item once; only items a
// record can touch scan the list per cell.
template <bool CENSUS, bool MODS>
static __device__ __forceinline__ void se_k1c_body(const SeLutStepParams& p, const SeLutCensusParams& cx, int* hist_sm) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ unsigned fat_sm[256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    unsigned smem_sa;
    asm volatile("mov.u32 %0, %1;" : "=r"(smem_sa) : "r"((unsigned)__cvta_generic_to_shared(smem)));
#if !SE_LUT_TWO_TABLES
    if (CENSUS) {
        for (int i = tid; i < 256; i += blockDim.x) hist_sm[i] = 0;
    }
#endif
#if SE_LUT_MODE == 1
    {
        const uint4* lut4 = reinterpret_cast<const uint4*>(p.lut);
        const int n4 = (p.table_bytes + 15) >> 4;
        for (int i = tid; i < n4; i += blockDim.x) {
            const uint4 v = __ldg(lut4 + i);
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(smem_sa + 16u * i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        }
    }
    const SeTab tab{smem_sa, smem_sa + (unsigned)p.pool_offset};
#else
    const SeTab tab{p.lut, p.pool};
#endif
    if (tid < 256) fat_sm[tid] = se_fat_table[tid];
    __syncthreads();

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
        bool item_hit = false;                                             // can a modification record touch this item?
        if (MODS) {
            const int x_lo = 2 * (c_begin * 32) - ox, x_hi = 2 * (c_end * 32) - ox - 1;
            for (int m0 = 0; m0 < p.n_mods; m0 += 32) {
                bool touch = false;
                if (m0 + lane < p.n_mods) touch = se_mod_touches(p.mods[m0 + lane], x_lo, x_hi, y0, y1);
                item_hit = item_hit || __any_sync(0xFFFFFFFFu, touch);
            }
        }
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
            // ---- transitions: the table reads of the whole batch back to back, the rare entries behind one test ----
            unsigned v[SE_K1C_BATCH], e[SE_K1C_BATCH], mir = 0u, unknown = 0u;   // unknown bit u: a raw id the table does not know (reads as NULL)
#pragma unroll
            for (int u = 0; u < SE_K1C_BATCH; ++u) {
                const unsigned vmax = max(max(a[u], b[u]), max(c[u], d[u]));
                if (!(cb + u < c_end && (cb + u) * 32 + lane < nbx)) {
                    v[u] = 0u;                                             // no block here: entry 0 (all EMPTY), nothing written
                } else if (vmax >= (unsigned)SE_N_MATERIALS) {
                    unknown |= 1u << u;
                    v[u] = se_clamp_id(a[u]) | (se_clamp_id(b[u]) << 8) | (se_clamp_id(c[u]) << 16) | (se_clamp_id(d[u]) << 24);
                } else {
                    v[u] = __byte_perm(__byte_perm(a[u], b[u], 0x0040), __byte_perm(c[u], d[u], 0x0040), 0x5410);
                }
                const unsigned seed = (unsigned)(2 * ((cb + u) * 32 + lane) - ox) * 461u + rowseed;
                const bool mirror = se_mirror_bit(seed * 213u);
                mir |= mirror ? (1u << u) : 0u;
#if SE_LUT_TWO_TABLES
                e[u] = se_tab_entry(tab, se_idx4(v[u]) + (mirror ? (unsigned)SE_N4 : 0u));
#else
                e[u] = se_tab_entry(tab, se_idx4(__byte_perm(v[u], 0u, mirror ? 0x2301u : 0x3210u)));
#endif
            }
            unsigned any = 0u;
#pragma unroll
            for (int u = 0; u < SE_K1C_BATCH; ++u) any |= e[u];
            if (any & SE_E_SPECIAL) {
#pragma unroll
                for (int u = 0; u < SE_K1C_BATCH; ++u)
                    if ((e[u] & (SE_E_SPECIAL | SE_E_SLOW)) == SE_E_SPECIAL) {
                        const unsigned seed = (unsigned)(2 * ((cb + u) * 32 + lane) - ox) * 461u + rowseed;
                        const unsigned u1 = se_hashi(seed * 2131u);
                        do {
                            unsigned thr, ea, eb;
                            se_tab_pool(tab, e[u] & SE_E_INDEX, thr, ea, eb);
                            e[u] = (u1 <= thr) ? ea : eb;
                        } while (e[u] & SE_E_SPECIAL);
                    }
#if !SE_LUT_TWO_TABLES
#pragma unroll
                for (int u = 0; u < SE_K1C_BATCH; ++u)
                    if (e[u] & SE_E_SPECIAL) {
                        const unsigned seed = (unsigned)(2 * ((cb + u) * 32 + lane) - ox) * 461u + rowseed;
                        e[u] = se_block_special(e[u], v[u], seed, ((mir >> u) & 1u) ? 0x2301u : 0x3210u, 0, 0, 0, tab, fat_sm);
                        if (CENSUS) unknown |= 1u << u;                    // generated code: no flag to go by, compare the cells
                    }
#endif
            }
#pragma unroll
            for (int u = 0; u < SE_K1C_BATCH; ++u) {                       // ---- write-back of the cells that changed
                const int bx = (cb + u) * 32 + lane;
                const int x0 = 2 * bx - ox;
                if (cb + u < c_end && bx < nbx) {
#if !SE_LUT_TWO_TABLES
                    bool look = CENSUS && ((e[u] & SE_E_POPFLAG) || ((unknown >> u) & 1u));
                    if (CENSUS) e[u] &= ~SE_E_POPFLAG;
#endif
                    unsigned nv = SE_LUT_TWO_TABLES ? e[u] : __byte_perm(e[u], 0u, ((mir >> u) & 1u) ? 0x2301u : 0x3210u);
                    bool overridden = false;
                    if (MODS && item_hit) {
                        unsigned m;
                        if (se_mod_lookup(p.mods, p.n_mods, x0, y0, m)) { nv = (nv & 0xFFFFFF00u) | m; overridden = true; }
                        if (se_mod_lookup(p.mods, p.n_mods, x0 + 1, y0, m)) { nv = (nv & 0xFFFF00FFu) | (m << 8); overridden = true; }
                        if (se_mod_lookup(p.mods, p.n_mods, x0, y1, m)) { nv = (nv & 0xFF00FFFFu) | (m << 16); overridden = true; }
                        if (se_mod_lookup(p.mods, p.n_mods, x0 + 1, y1, m)) { nv = (nv & 0x00FFFFFFu) | (m << 24); overridden = true; }
                    }
#if !SE_LUT_TWO_TABLES
                    // only flagged outcomes and blocks cut by the grid's / the strip's edge are looked at
                    look = look || (CENSUS && overridden);
                    if (CENSUS && (look || cm_rows != 0xFu || x0 < 0 || x0 + 1 >= p.W)) {
                        const unsigned cm_cols = (x0 >= 0 ? 5u : 0u) | ((x0 + 1) < p.W ? 10u : 0u);
                        se_census_block(hist_sm, look, a[u], b[u], c[u], d[u], nv, cm_rows & cm_cols);
                    }
#endif
                    // a cell is written when its id changed -- or when the raw word was an id the table does not know (it is
                    // NULL from now on, like in the reference: operations.glsl:111 stores the id of what getCell returned)
                    const unsigned diff = ((unknown >> u) & 1u) ? 0xFFFFFFFFu : (nv ^ v[u]);
                    if (diff != 0u) {
                        const unsigned na = nv & 0xFFu, nb = (nv >> 8) & 0xFFu, nc = (nv >> 16) & 0xFFu, nd = nv >> 24;
                        if (vec_ok) {
                            if (st0 == 0 && (diff & 0xFFFFu)) *reinterpret_cast<uint2*>(base0 + x0) = make_uint2(na, nb);
                            if (st1 == 0 && (diff >> 16)) *reinterpret_cast<uint2*>(base1 + x0) = make_uint2(nc, nd);
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
    }
#if !SE_LUT_TWO_TABLES
    if (CENSUS) {
        __syncthreads();
        for (int i = tid; i < 256; i += blockDim.x) {
            const int dlt = hist_sm[i];
            if (dlt != 0) atomicAdd(cx.census + i, (unsigned long long)(long long)dlt);   // two's complement: negative deltas wrap correctly
        }
    }
#endif
}

extern "C" __global__ void __launch_bounds__(SE_K1C_THREADS, SE_K1C_MINCTAS) se_step_lut_global(const SeLutStepParams p) {
    se_k1c_body<false, false>(p, SeLutCensusParams{}, nullptr);
}
extern "C" __global__ void __launch_bounds__(SE_K1C_THREADS, SE_K1C_MINCTAS) se_step_lut_global_mods(const SeLutStepParams p) {
    se_k1c_body<false, true>(p, SeLutCensusParams{}, nullptr);
}

#if !SE_LUT_TWO_TABLES
extern "C" __global__ void __launch_bounds__(SE_K1C_THREADS, SE_K1C_MINCTAS) se_step_lut_global_census(const SeLutStepParams p, const SeLutCensusParams cx) {
    __shared__ int hist_sm[256];
    se_k1c_body<true, false>(p, cx, hist_sm);
}
extern "C" __global__ void __launch_bounds__(SE_K1C_THREADS, SE_K1C_MINCTAS) se_step_lut_global_census_mods(const SeLutStepParams p, const SeLutCensusParams cx) {
    __shared__ int hist_sm[256];
    se_k1c_body<true, true>(p, cx, hist_sm);
}

#endif

// =============================================================================================
// K3f: Margolus step + modification override + lighting relaxation in ONE pass over the grid (the reference's single
// dispatch, falling_sand.glsl:737-799 + operations.glsl:99-171), for table-eligible rule sets.  HBM-bound by design:
// 4 B old id + 16 B old light in, 4 B new id + 16 B new light out per cell = the 40 algorithmic bytes.
//
// One persistent CTA of 768 threads per SM, run as THREE INDEPENDENT GROUPS of 8 warps (named barriers, like K1b's halves; `half`
// below is the group index) that read the transition table where it lies (L1 / L2).  A group walks 64 x 16 tiles.  The inputs of a tile -- light and ids of the tile and its one-cell ring -- arrive by TMA
// (cp.async.bulk.tensor: one 3-D box of light in groups of four float4, one 3-D box of ids in groups of 16, out-of-grid
// elements zero-filled) into one of the half's two buffers, signalled by an mbarrier.  Three jobs per tile:
//   B  the 2x2 blocks that cover the tile (they lie inside tile + ring: the block offset is 0 or 1): old ids straight
//      from the TMA buffer (unknown ids NULL, WALL outside the grid, MISSING inside the grid but outside a strip's
//      buffer) through the transition table, one block per thread, new ids as bytes; blocks cut by a tile edge are
//      evaluated by both neighbours (deterministic: RAND depends on position and frame only)
//   A  every ring + tile cell: the reader-independent neighbour term (rgb * keep * a, a) written over the light
//   C  every tile cell: new id (B's byte, overridden by the culled modification list), stored; the eight neighbour
//      terms combined in the shader's order, stored.  Tiles whose ring lies inside the grid and the buffer take a
//      branch-free version: packed adds (add.f32x2) for the sums, three-input maxima, the alpha rule (a zero alpha
//      counts as the running maximum) as one predicated add.
// A and B depend on the TMA data only, C on A and B of the same tile: the loop is software-pipelined -- an iteration is
// C(k), then B(k+1) and A(k+1), then ONE barrier of the half; the loads of tile k+1 are issued at the start of the
// iteration (their buffer was released by the barrier before) and land while C(k) runs.  Id bytes and the culled
// modification list (indices) are double-buffered like the TMA buffers.
// =============================================================================================
#define SE_LF_TW 64
#ifndef SE_LF_NG
#define SE_LF_NG 3                                   // independent groups of warps per CTA (each with its own tiles, buffers and named barrier)
#endif
#ifndef SE_LF_HALF
#define SE_LF_HALF 256                               // threads of a group (NG x HALF = 768: 80 registers, the 3 x 3 window of phase C stays in registers)
#endif
#ifndef SE_LF_ROWS
#define SE_LF_ROWS 4                                 // cells of one column a thread relaxes: 2 or 4
#endif
#define SE_LF_TH (SE_LF_ROWS * (SE_LF_HALF / 64))    // tile height: 16 (24 with groups of 384 threads)
#define SE_LF_RH (SE_LF_TH + 2)                      // ring rows
#define SE_LF_RW (SE_LF_TW + 2)                      // ring columns
// TMA box rows are 64 bytes of light and 32 bytes of ids (16-byte rows -- a float4, four ids -- made the TMA unit the
// bottleneck: 1512 box rows per 1024 cells): the light box is 18 groups of four float4 and starts 4 columns left of the
// tile, the id box is 10 groups of 8 ids and starts 8 columns left of it.  Ring column j is element j + LCOL0 of a light
// row, j + ICOL0 of an id row and j + BCOL0 of a row of id bytes.
#define SE_LF_LSTRIDE (SE_LF_TW + 8)                 // light / term row stride (float4)
#define SE_LF_LCOL0 3
#define SE_LF_ISTRIDE (SE_LF_TW + 16)                // id row stride (u32)
#define SE_LF_ICOL0 7
#define SE_LF_BSTRIDE (SE_LF_TW + 8)                 // id byte row stride
#define SE_LF_BCOL0 3
#define SE_LF_LIGHT_BYTES (SE_LF_RH * SE_LF_LSTRIDE * 16)
#define SE_LF_IDS_BYTES (SE_LF_RH * SE_LF_ISTRIDE * 4)
#define SE_LF_IDS_OFFSET ((SE_LF_LIGHT_BYTES + 127) / 128 * 128)
#define SE_LF_BUF_BYTES ((SE_LF_IDS_OFFSET + SE_LF_IDS_BYTES + 127) / 128 * 128)
#define SE_LF_MISSING 0xFFu
#define SE_LF_THREADS (SE_LF_NG * SE_LF_HALF)
#ifndef SE_LF_NBUF
#define SE_LF_NBUF 2                                 // TMA buffers per half: the loads of tile k + NBUF - 1 are issued when tile k's phase C starts
#endif
#define SE_LF_RING_CELLS (SE_LF_RH * SE_LF_RW)
#define SE_LF_MAX_BLOCKS ((SE_LF_TW / 2 + 1) * (SE_LF_TH / 2 + 1))
static_assert(SE_LF_MAX_BLOCKS - SE_LF_HALF <= 64 && SE_LF_HALF - SE_LF_RING_CELLS % SE_LF_HALF - 64 >= 0, "se_step_lit: work split of phases A and B");

struct alignas(64) SeTensorMap { unsigned long long opaque[16]; };   // CUtensorMap (cuda.h), encoded by the host

struct SeLitParams {
    unsigned* new_cells;     // ids after the step (the other ping-pong buffer), local rows
    float4* light_out;
    int W, Hl, gy0, Hg;
    int frame;
    int n_mods;              // already cut at the first mod_size == 0
    const SeMod* mods;
    int table_bytes, pool_offset;
    const unsigned* lut;
    const unsigned* pool;
    int tiles_x, tiles_y;
    int buf_offset;          // byte offset of the first input buffer in dynamic shared memory (behind the staged table)
    unsigned tiles_x_magic;  // floor(2^32 / tiles_x) + 1: tile index -> (row, column) without a division
};

static __device__ __forceinline__ void se_mbar_init(unsigned mbar_sa, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar_sa), "r"(count) : "memory"); }
static __device__ __forceinline__ void se_mbar_expect_tx(unsigned mbar_sa, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar_sa), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void se_mbar_wait(unsigned mbar_sa, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" :: "r"(mbar_sa), "r"(parity) : "memory");
}
static __device__ __forceinline__ void se_tma_load_3d(unsigned dst_sa, const SeTensorMap* tm, int c0, int c1, int c2, unsigned mbar_sa) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(dst_sa), "l"(reinterpret_cast<unsigned long long>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(mbar_sa) : "memory");
}

// a term / light value as two packed pairs: (x, y) and (z, w), the registers an LDS.128 delivers
struct SeF4P { unsigned long long xy, zw; };
static __device__ __forceinline__ SeF4P se_lds_f4p(unsigned a) {
    SeF4P v;
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(v.xy), "=l"(v.zw) : "r"(a));
    return v;
}
static __device__ __forceinline__ unsigned long long se_add2(unsigned long long a, unsigned long long b) {   // two IEEE f32 adds (FADD2)
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
static __device__ __forceinline__ float se_lo(unsigned long long v) { return __uint_as_float((unsigned)v); }
static __device__ __forceinline__ float se_hi(unsigned long long v) { return __uint_as_float((unsigned)(v >> 32)); }
static __device__ __forceinline__ float se_max3(float a, float b, float c) {
#if (__CUDACC_VER_MAJOR__ > 12) || (__CUDACC_VER_MAJOR__ == 12 && __CUDACC_VER_MINOR__ >= 9)
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));   // FMNMX3 (the ptxas of CUDA 12.8 and older rejects the three-input form)
    return r;
#else
    return fmaxf(fmaxf(a, b), c);
#endif
}

// phase A for one ring / tile cell (row i, column j) of a tile that touches the edge of the grid or of the buffer
static __device__ __forceinline__ void se_lit_stage_cell(const SeLitParams& p, const unsigned* fat_sm, unsigned light_sa, unsigned ids_sa,
                                                         int x_org, int yl_org, int i, int j) {
    const unsigned c = (unsigned)(i * SE_LF_LSTRIDE + j + SE_LF_LCOL0);
    const int x = x_org + j, yl = yl_org + i, y = p.gy0 + yl;
    const bool local = x >= 0 && x < p.W && y >= 0 && y < p.Hg && yl >= 0 && yl < p.Hl;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (local) {
        const unsigned id = se_lds_u32(ids_sa + 4u * (unsigned)(i * SE_LF_ISTRIDE + j + SE_LF_ICOL0));
        float4 li;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(li.x), "=f"(li.y), "=f"(li.z), "=f"(li.w) : "r"(light_sa + 16u * c));
        const unsigned nf = fat_sm[id < 255u ? id : 255u];
        const float keep = (nf & SE_F_OBSTACLE) ? 0.0f : 1.0f;             // vec4(vec3(float(!obstacle)), 1.0), :498
        const float la = li.w;
        v = make_float4((li.x * keep) * la, (li.y * keep) * la, (li.z * keep) * la, la);
    }
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(light_sa + 16u * c), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// per-cell modification scan over the culled list (indices into the frame's records, order kept): se_mod_lookup
static __device__ __forceinline__ bool se_mod_lookup_culled(const SeMod* __restrict__ mods, const unsigned char* idx, int n, int x, int y, unsigned& mat_out) {
    bool got = false;
    int fin = 1;   // MAT_NULL
    for (int i = 0; i < n; ++i) {
        const SeMod m = mods[idx[i]];
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

// one neighbour term of the branch-free phase C: sums as two packed adds; a zero alpha counts as the running maximum of
// the alphas before it (math.glsl:168-176: adding the zero first is exact, so one predicated add corrects the sum), and the
// running maximum itself is max(previous, alpha) in both cases (it is never negative)
#define SE_LF_ACC(v)                                                                      \
    {                                                                                     \
        const float w_ = se_hi((v).zw);                                                   \
        sxy = se_add2(sxy, (v).xy); szw = se_add2(szw, (v).zw);                           \
        if (w_ == 0.0f) szw = ((unsigned long long)__float_as_uint(se_hi(szw) + mf) << 32) | (szw & 0xFFFFFFFFull); \
        mf = fmaxf(mf, w_);                                                               \
    }
#define SE_LF_MAX2(v1, v2)                                                                \
    {                                                                                     \
        mxx = se_max3(mxx, se_lo((v1).xy), se_lo((v2).xy)); mxy = se_max3(mxy, se_hi((v1).xy), se_hi((v2).xy)); \
        mxz = se_max3(mxz, se_lo((v1).zw), se_lo((v2).zw));                               \
    }
#define SE_LF_FINISH(out)                                                                 \
    {                                                                                     \
        const float ax = se_lo(sxy) * 0.125f, ay = se_hi(sxy) * 0.125f, az = se_lo(szw) * 0.125f, aw = se_hi(szw) * 0.125f;   /* num == 8: exact */ \
        out = make_float4(ax * 0.5f + mxx * 0.5f, ay * 0.5f + mxy * 0.5f, az * 0.5f + mxz * 0.5f, aw);   /* mix(avg.rgb, max.rgb, 0.5), :521 */ \
    }

extern "C" __global__ void __launch_bounds__(SE_LF_THREADS, 1) se_step_lit(const __grid_constant__ SeTensorMap tm_cells, const __grid_constant__ SeTensorMap tm_light,
                                                                              const SeLitParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(16) unsigned fat_sm[256 + 256 * 4];   // the fat-cell flags, then the emission (a float4 per material)
    __shared__ unsigned char cull_sm[SE_LF_NG][2][256];
    __shared__ int n_cull_sm[SE_LF_NG][2];
    __shared__ __align__(8) unsigned long long mbar[SE_LF_NG * SE_LF_NBUF];
    __shared__ __align__(16) unsigned char ids8[SE_LF_NG][2][SE_LF_RH * SE_LF_BSTRIDE];
    const int tid = threadIdx.x, lane = tid & 31, half = tid / SE_LF_HALF, ht = tid - half * SE_LF_HALF, hw = ht >> 5;   // `half`: the thread's group
    unsigned smem_sa;
    asm volatile("mov.u32 %0, %1;" : "=r"(smem_sa) : "r"((unsigned)__cvta_generic_to_shared(smem)));
    // The table is read where it lies (L1 / L2): phase B is off the critical path (its gather is in flight while phase A
    // runs), and the shared memory holds the TMA buffers instead -- the loads are what the kernel waits for.
    const SeTabG tab{p.lut, p.pool};
    if (tid < 256) fat_sm[tid] = se_fat_table[tid];
    for (int i = tid; i < 256 * 4; i += SE_LF_THREADS) fat_sm[256 + i] = __float_as_uint(se_emission_table[i]);
    // both tables are read with plain shared-space addresses (a C array access costs four uniform instructions per look-up
    // here: the kernel's shared memory is addressed through the cluster window)
    const unsigned fat_sa = (unsigned)__cvta_generic_to_shared(fat_sm);
    const unsigned mbar_sa = (unsigned)__cvta_generic_to_shared(mbar) + 8u * SE_LF_NBUF * (unsigned)half;   // this half's barriers
    if (tid == 0) {
        for (unsigned k = 0; k < (unsigned)(SE_LF_NG * SE_LF_NBUF); ++k) se_mbar_init((unsigned)__cvta_generic_to_shared(mbar) + 8u * k, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int n_tiles = p.tiles_x * p.tiles_y;
    const int first = SE_LF_NG * (int)blockIdx.x + half, stride = SE_LF_NG * (int)gridDim.x;
    // TMA destinations are 128-byte aligned whatever the base of dynamic shared memory is
    const unsigned buf0_sa = ((smem_sa + (unsigned)p.buf_offset + 127u) & ~127u) + (unsigned)half * (SE_LF_NBUF * SE_LF_BUF_BYTES);
    const unsigned ids8_base = (unsigned)__cvta_generic_to_shared(&ids8[half][0][0]);
    auto tile_xy = [&](int t, int& bx, int& by) {          // t = by * tiles_x + bx
        by = p.tiles_x > 1 ? (int)__umulhi((unsigned)t, p.tiles_x_magic) : t;
        bx = t - by * p.tiles_x;
        if (bx < 0) { --by; bx += p.tiles_x; }
    };
    auto issue = [&](int t, int buf) {                       // one thread: both boxes of tile t into buffer `buf`
        int bx, by;
        tile_xy(t, bx, by);
        const unsigned dst = buf0_sa + (unsigned)buf * SE_LF_BUF_BYTES, mb = mbar_sa + 8u * (unsigned)buf;
        se_mbar_expect_tx(mb, SE_LF_LIGHT_BYTES + SE_LF_IDS_BYTES);
        se_tma_load_3d(dst, &tm_light, 0, bx * (SE_LF_TW / 4) - 1, by * SE_LF_TH - 1, mb);
        se_tma_load_3d(dst + SE_LF_IDS_OFFSET, &tm_cells, 0, bx * (SE_LF_TW / 8) - 1, by * SE_LF_TH - 1, mb);
    };
    if (ht == 0) {
        for (int j = 0; j < SE_LF_NBUF - 1; ++j)
            if (first + j * stride < n_tiles) issue(first + j * stride, j);
    }

    int ox, oy;
    se_margolus_offset(p.frame, ox, oy);
    const int nbx = SE_LF_TW / 2 + ox, n_blocks = nbx * (SE_LF_TH / 2 + oy);
    // buffers: tile k sits in bufC, tile k + 1 is staged in bufS (mbarrier parity parS), tile k + NBUF - 1 is loaded into bufI
    int bufI = 0, bufC = SE_LF_NBUF - 1, bufS = 0;
    unsigned parS = 0u;
    unsigned tileC = 0u;                                     // tile k: bx | by << 10 | interior << 31 (set when it was staged)
    for (int k = -1, t = first - stride;; ++k, t += stride) {
        const bool has_next = t + stride < n_tiles;
        if (k >= 0) {
            const int par = k & 1;
            // every thread of the half is past phase C of tile k - 1: that tile's buffer takes the loads of tile k + NBUF - 1
            // (the proxy fence orders the generic writes of its phase A before the async ones)
            if (ht == 0 && t + (SE_LF_NBUF - 1) * stride < n_tiles) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(t + (SE_LF_NBUF - 1) * stride, bufI);
            }
            // ---- phase C of tile k: new id + light of every tile cell: column lane + 32 * (warp & 1), SE_LF_ROWS rows ----
            const int bx = (int)(tileC & 0x3FFu), by = (int)((tileC >> 10) & 0x1FFFFFu);
            const unsigned light_sa = buf0_sa + (unsigned)bufC * SE_LF_BUF_BYTES;
            const unsigned ids8_sa = ids8_base + (unsigned)par * (SE_LF_RH * SE_LF_BSTRIDE);
            const int n_cull = n_cull_sm[half][par];
            const unsigned char* const cull = cull_sm[half][par];
            const int tx = lane + 32 * (hw & 1);
            const int x = bx * SE_LF_TW + tx;
#define SE_LF_T(r, c) se_lds_f4p(tp + 16u * (unsigned)((r) * SE_LF_LSTRIDE + (c)))          /* ring row row0 + r, ring column tx + c */
            if (tileC >> 31) {
                // the 3 x 3 window slides down the thread's column: three terms are loaded per cell, six stay in registers
                const int row0 = (hw >> 1) * SE_LF_ROWS;
                const unsigned tp = light_sa + 16u * (unsigned)(row0 * SE_LF_LSTRIDE + tx + SE_LF_LCOL0);   // term (row0 - 1, tx - 1) of the tile
                const unsigned idp = ids8_sa + (unsigned)((row0 + 1) * SE_LF_BSTRIDE + tx + 1 + SE_LF_BCOL0);
                const int y0 = p.gy0 + by * SE_LF_TH + row0;
                const size_t idx0 = (size_t)(by * SE_LF_TH + row0) * p.W + x;
                unsigned* cell_out = p.new_cells + idx0;
                float4* light_out = p.light_out + idx0;
                SeF4P a0 = SE_LF_T(0, 0), a1 = SE_LF_T(0, 1), a2 = SE_LF_T(0, 2);
                SeF4P b0 = SE_LF_T(1, 0), b1 = SE_LF_T(1, 1), b2 = SE_LF_T(1, 2);
#pragma unroll
                for (int i = 0; i < SE_LF_ROWS; ++i) {
                    const SeF4P c0 = SE_LF_T(i + 2, 0), c1 = SE_LF_T(i + 2, 1), c2 = SE_LF_T(i + 2, 2);
                    unsigned id = se_lds_u8(idp + (unsigned)(i * SE_LF_BSTRIDE));
                    if (n_cull) {
                        unsigned m;
                        if (se_mod_lookup_culled(p.mods, cull, n_cull, x, y0 + i, m)) id = m;
                    }
                    *cell_out = id;
                    cell_out += p.W;
                    const unsigned me = id < 255u ? id : 255u;
                    float4 out;
                    {   // DOWN, UP, DOWNLEFT, UPLEFT, DOWNRIGHT, UPRIGHT, RIGHT, LEFT (math.glsl:154-166)
                        unsigned long long sxy = 0ull, szw = 0ull;
                        float mf = 0.0f, mxx = 0.0f, mxy = 0.0f, mxz = 0.0f;
                        SE_LF_ACC(c1) SE_LF_ACC(a1) SE_LF_MAX2(c1, a1)
                        SE_LF_ACC(c0) SE_LF_ACC(a0) SE_LF_MAX2(c0, a0)
                        SE_LF_ACC(c2) SE_LF_ACC(a2) SE_LF_MAX2(c2, a2)
                        SE_LF_ACC(b2) SE_LF_ACC(b0) SE_LF_MAX2(b2, b0)
                        SE_LF_FINISH(out)
                    }
                    if (se_lds_u32(fat_sa + 4u * me) & SE_F_EMISSIVE)                  // operations.glsl:126-127
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(out.x), "=f"(out.y), "=f"(out.z), "=f"(out.w) : "r"(fat_sa + 1024u + 16u * me));
                    *light_out = out;
                    light_out += p.W;
                    a0 = b0; a1 = b1; a2 = b2;
                    b0 = c0; b1 = c1; b2 = c2;
                }
            } else if (x < p.W) {
                const int row0 = (hw >> 1) * SE_LF_ROWS;
                const unsigned tp = light_sa + 16u * (unsigned)(row0 * SE_LF_LSTRIDE + tx + SE_LF_LCOL0);
                const unsigned idp = ids8_sa + (unsigned)((row0 + 1) * SE_LF_BSTRIDE + tx + 1 + SE_LF_BCOL0);
#define SE_LF_LDT(dst, off) asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(dst.x), "=f"(dst.y), "=f"(dst.z), "=f"(dst.w) : "r"(tp + 16u * (unsigned)(off)))
                float4 a0, a1, a2, b0, b1, b2;
                SE_LF_LDT(a0, 0); SE_LF_LDT(a1, 1); SE_LF_LDT(a2, 2);
                SE_LF_LDT(b0, SE_LF_LSTRIDE); SE_LF_LDT(b1, SE_LF_LSTRIDE + 1); SE_LF_LDT(b2, SE_LF_LSTRIDE + 2);
#pragma unroll 1
                for (int i = 0; i < SE_LF_ROWS; ++i) {
                    const int yl = by * SE_LF_TH + row0 + i;
                    if (yl >= p.Hl) break;
                    float4 c0, c1, c2;
                    SE_LF_LDT(c0, (i + 2) * SE_LF_LSTRIDE); SE_LF_LDT(c1, (i + 2) * SE_LF_LSTRIDE + 1); SE_LF_LDT(c2, (i + 2) * SE_LF_LSTRIDE + 2);
                    const int y = p.gy0 + yl;
                    const size_t idx = (size_t)yl * p.W + x;
                    unsigned id = se_lds_u8(idp + (unsigned)(i * SE_LF_BSTRIDE));
                    if (n_cull) {
                        unsigned m;
                        if (se_mod_lookup_culled(p.mods, cull, n_cull, x, y, m)) id = m;
                    }
                    p.new_cells[idx] = id;
                    const unsigned me = id < 255u ? id : 255u;
                    float4 light;
                    if (fat_sm[me] & SE_F_EMISSIVE) {                       // operations.glsl:126-127
                        light = make_float4(se_emission_table[me * 4 + 0], se_emission_table[me * 4 + 1], se_emission_table[me * 4 + 2], se_emission_table[me * 4 + 3]);
                    } else if (y == 0) {                                    // :128-129
                        light = make_float4(1.0f, 1.0f, 1.0f, 0.999999f);
                    } else {
                        float4 avg = make_float4(0.f, 0.f, 0.f, 0.f), mx = make_float4(0.f, 0.f, 0.f, 0.f);
                        float max_falloff = 0.0f;
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
                        if (num > 0) {                                      // :516-518
                            const float dn = (float)num;
                            avg.x = __fdiv_rn(avg.x, dn); avg.y = __fdiv_rn(avg.y, dn); avg.z = __fdiv_rn(avg.z, dn); avg.w = __fdiv_rn(avg.w, dn);
                        }
                        light = make_float4(avg.x * 0.5f + mx.x * 0.5f, avg.y * 0.5f + mx.y * 0.5f, avg.z * 0.5f + mx.z * 0.5f, avg.w);   // mix(avg.rgb, max.rgb, 0.5), :521
                    }
                    p.light_out[idx] = light;
                    a0 = b0; a1 = b1; a2 = b2;
                    b0 = c0; b1 = c1; b2 = c2;
                }
#undef SE_LF_LDT
            }
#undef SE_LF_T
        }
        if (!has_next) break;

        // ---- tile k + 1: its loads have had phase C to land ----
        {
            const int par = (k + 1) & 1;
            int bx, by;
            tile_xy(t + stride, bx, by);
            const int x_org = bx * SE_LF_TW - 1, yl_org = by * SE_LF_TH - 1;      // ring cell (0, 0): column / local row
            const bool interior = x_org >= 0 && x_org + SE_LF_TW + 1 < p.W && yl_org >= 0 && yl_org + SE_LF_TH + 1 < p.Hl &&
                                  p.gy0 + yl_org >= 0 && p.gy0 + yl_org + SE_LF_TH + 1 < p.Hg;
            tileC = (unsigned)bx | ((unsigned)by << 10) | (interior ? 0x80000000u : 0u);
            const unsigned light_sa = buf0_sa + (unsigned)bufS * SE_LF_BUF_BYTES, ids_sa = light_sa + SE_LF_IDS_OFFSET;
            const unsigned ids8_sa = ids8_base + (unsigned)par * (SE_LF_RH * SE_LF_BSTRIDE);
            se_mbar_wait(mbar_sa + 8u * (unsigned)bufS, parS);

            // ---- phase B, first part: the blocks that cover the tile, old ids from the TMA buffer, table read issued ----
            // one block per thread per pass (a second pass only when both block offsets are 1: 49 blocks)
            auto block_ids = [&](int b, unsigned& v, unsigned& seed, unsigned& q) {   // false: a row of the block is not in the local buffer
                const int bj = ox ? b / (SE_LF_TW / 2 + 1) : b / (SE_LF_TW / 2), bi = b - bj * nbx;
                const int cx = 2 * bi + 1 - ox, cy = 2 * bj + 1 - oy;      // ring coordinates of the block's top-left cell
                const unsigned e0 = (unsigned)(cy * SE_LF_ISTRIDE + cx + SE_LF_ICOL0);
                unsigned a = se_lds_u32(ids_sa + 4u * e0), bb = se_lds_u32(ids_sa + 4u * (e0 + 1u));
                unsigned c = se_lds_u32(ids_sa + 4u * (e0 + SE_LF_ISTRIDE)), d = se_lds_u32(ids_sa + 4u * (e0 + SE_LF_ISTRIDE + 1u));
                a = se_clamp_id(a); bb = se_clamp_id(bb); c = se_clamp_id(c); d = se_clamp_id(d);
                if (!interior) {                                           // WALL outside the grid (operations.glsl:45-51), MISSING outside the buffer
                    const int x0 = x_org + cx, yl0 = yl_org + cy, y0 = p.gy0 + yl0;
                    const bool xin0 = x0 >= 0 && x0 < p.W, xin1 = x0 + 1 >= 0 && x0 + 1 < p.W;
                    const bool yin0 = y0 >= 0 && y0 < p.Hg, yin1 = y0 + 1 >= 0 && y0 + 1 < p.Hg;
                    const bool loc0 = yl0 >= 0 && yl0 < p.Hl, loc1 = yl0 + 1 >= 0 && yl0 + 1 < p.Hl;
                    a = !(xin0 && yin0) ? 2u : (loc0 ? a : SE_LF_MISSING);
                    bb = !(xin1 && yin0) ? 2u : (loc0 ? bb : SE_LF_MISSING);
                    c = !(xin0 && yin1) ? 2u : (loc1 ? c : SE_LF_MISSING);
                    d = !(xin1 && yin1) ? 2u : (loc1 ? d : SE_LF_MISSING);
                }
                v = a | (bb << 8) | (c << 16) | (d << 24);
                seed = (unsigned)(x_org + cx) * 461u + (unsigned)(p.gy0 + yl_org + cy) * 2131u + (unsigned)p.frame * (2131u * 2131u);
                q = ids8_sa + (unsigned)(cy * SE_LF_BSTRIDE + cx + SE_LF_BCOL0);
                return !((a | bb | c | d) & 0x80u);
            };
            auto block_store = [&](unsigned q, unsigned nv) {
                se_sts_u8(q, nv & 0xFFu); se_sts_u8(q + 1u, (nv >> 8) & 0xFFu);
                se_sts_u8(q + SE_LF_BSTRIDE, (nv >> 16) & 0xFFu); se_sts_u8(q + SE_LF_BSTRIDE + 1u, nv >> 24);
            };
            unsigned bv = 0u, bseed = 0u, bq = 0u, be = 0u, bsel = 0x3210u;
            bool bok = false;
            const bool has_block = ht < n_blocks;
            if (has_block) {
                bok = block_ids(ht, bv, bseed, bq);
                if (bok) {
                    const bool mirror = se_hashi(bseed * 213u) <= SE_MIRROR_UMAX;
#if SE_LUT_TWO_TABLES
                    be = se_tab_entry(tab, se_idx4(bv) + (mirror ? (unsigned)SE_N4 : 0u));
#else
                    bsel = mirror ? 0x2301u : 0x3210u;
                    be = se_tab_entry(tab, se_idx4(__byte_perm(bv, 0u, bsel)));
#endif
                }
            }

            // ---- phase A: the neighbour term of every ring + tile cell, in place; the short last pass goes to the upper threads ----
            if (interior) {
#pragma unroll
                for (int n = 0; n <= SE_LF_RING_CELLS / SE_LF_HALF; ++n) {
                    const bool last = n == SE_LF_RING_CELLS / SE_LF_HALF;
                    const int c = last ? ht + (SE_LF_RING_CELLS - SE_LF_HALF) : ht + n * SE_LF_HALF;
                    if (!last || c >= (SE_LF_RING_CELLS / SE_LF_HALF) * SE_LF_HALF) {
                        const int i = c / SE_LF_RW;                          // c = i * RW + j
                        const unsigned e = (unsigned)(c + i * (SE_LF_ISTRIDE - SE_LF_RW) + SE_LF_ICOL0);   // i * ISTRIDE + j + ICOL0
                        const unsigned la_sa = light_sa + 16u * (unsigned)(c + i * (SE_LF_LSTRIDE - SE_LF_RW) + SE_LF_LCOL0);
                        const unsigned id = se_lds_u32(ids_sa + 4u * e);
                        const SeF4P li = se_lds_f4p(la_sa);
                        const unsigned nf = se_lds_u32(fat_sa + 4u * (id < 255u ? id : 255u));
                        // (rgb * keep) * a with keep in {0, 1} is rgb * (keep ? a : 0) for finite light
                        const float la = se_hi(li.zw), lk = (nf & SE_F_OBSTACLE) ? 0.0f : la;
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(la_sa),
                                     "f"(se_lo(li.xy) * lk), "f"(se_hi(li.xy) * lk), "f"(se_lo(li.zw) * lk), "f"(la) : "memory");
                    }
                }
            } else {
                for (int c = ht; c < SE_LF_RING_CELLS; c += SE_LF_HALF) {
                    const int i = c / SE_LF_RW, j = c - i * SE_LF_RW;
                    se_lit_stage_cell(p, fat_sm, light_sa, ids_sa, x_org, yl_org, i, j);
                }
            }

            // ---- phase B, second part: the table entry has arrived; rand.y-dependent and slow states, new ids as bytes ----
            if (has_block) {
                unsigned nv = bv;
                if (bok) {
                    if (be & SE_E_SPECIAL) be = se_block_special(be, bv, bseed, bsel, 0, 0, 0, tab, fat_sm);
                    nv = SE_LUT_TWO_TABLES ? be : __byte_perm(be, 0u, bsel);
                }
                block_store(bq, nv);
            }
            // the second pass (at most 41 blocks) goes to threads that had no cell in the short last pass of phase A
            if (ht >= SE_LF_HALF - SE_LF_RING_CELLS % SE_LF_HALF - 64 && ht + 64 + SE_LF_RING_CELLS % SE_LF_HALF < n_blocks) {
                unsigned v, seed, q;
                const bool ok = block_ids(ht + 64 + SE_LF_RING_CELLS % SE_LF_HALF, v, seed, q);
                block_store(q, ok ? se_block_lut(v, seed, 0, 0, 0, tab, fat_sm) : v);
            }
            // the last warp culls the modifications against the tile (order kept: last match wins, falling_sand.glsl:764-773)
            if (hw == SE_LF_HALF / 32 - 1) {
                int n = 0;
                const int x_lo = bx * SE_LF_TW, y_lo = p.gy0 + by * SE_LF_TH;
                for (int m0 = 0; m0 < p.n_mods; m0 += 32) {
                    bool keep = false;
                    if (m0 + lane < p.n_mods) keep = se_mod_touches(p.mods[m0 + lane], x_lo, x_lo + SE_LF_TW - 1, y_lo, y_lo + SE_LF_TH - 1);
                    const unsigned ballot = __ballot_sync(0xFFFFFFFFu, keep);
                    if (keep) cull_sm[half][par][n + __popc(ballot & ((1u << lane) - 1u))] = (unsigned char)(m0 + lane);
                    n += __popc(ballot);
                }
                if (lane == 0) n_cull_sm[half][par] = n;
            }
        }
        asm volatile("bar.sync %0, %1;" :: "r"(half + 1), "n"(SE_LF_HALF) : "memory");
        bufI = bufC; bufC = bufS;
        if (++bufS == SE_LF_NBUF) { bufS = 0; parS ^= 1u; }
    }
}
#undef SE_LF_ACC
#undef SE_LF_MAX2
#undef SE_LF_FINISH
#endif  // SE_HOST_EMU
#endif  // SE_LUT_ELIGIBLE
