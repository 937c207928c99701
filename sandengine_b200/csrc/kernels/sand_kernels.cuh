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
// hash43(uvec3(pos_rounded, frame)) -- only the lanes the rule set consumes are evaluated
static __device__ __forceinline__ void se_hash43(int px, int py, int frame, SeRand& rnd) {
    const unsigned x = (unsigned)px * 461u + (unsigned)py * 2131u + (unsigned)frame * (2131u * 2131u);
    rnd.u[0] = se_hashi(x * 213u);
    rnd.u[1] = (SE_RAND_LANES & 2u) ? se_hashi(x * 2131u) : 0u;
    rnd.u[2] = (SE_RAND_LANES & 4u) ? se_hashi(x * 21313u) : 0u;
    rnd.u[3] = (SE_RAND_LANES & 8u) ? se_hashi(x * 213132u) : 0u;
}

// falling_sand.glsl:698-718.  The caller has already taken the all-EMPTY early-out (:692-694).
static __device__ __forceinline__ void se_block(unsigned& s, unsigned& r, unsigned& d, unsigned& dr, int px, int py, int frame) {
    SeRand rnd;
    se_hash43(px, py, frame, rnd);
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

// ---------------------------------------------------------------------------------------------
// K1a: one Margolus step, one thread per 2x2 block, straight from/to global memory.
//   IN_PLACE : out == in, only cells whose id changed are stored (blocks partition the grid, so the
//              update is race-free in place -- SURVEY.md section 7)
//   HAS_MODS : apply the modification override per cell
// Thread (tx, ty) of the grid handles block column bx and BPT block rows.
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

#define SE_DEFINE_STEP_GLOBAL(NAME, IN_PLACE, HAS_MODS)                                  \
    extern "C" __global__ void __launch_bounds__(256) NAME(const SeStepParams p) {        \
        __shared__ unsigned fat_sm[256];                                                  \
        const int t = threadIdx.y * blockDim.x + threadIdx.x;                             \
        if (t < 256) fat_sm[t] = se_fat_table[t];                                         \
        __syncthreads();                                                                  \
        se_step_global_impl<IN_PLACE, HAS_MODS>(p, fat_sm);                               \
    }
SE_DEFINE_STEP_GLOBAL(se_step_inplace, true, false)
SE_DEFINE_STEP_GLOBAL(se_step_inplace_mods, true, true)
SE_DEFINE_STEP_GLOBAL(se_step_pingpong, false, false)
SE_DEFINE_STEP_GLOBAL(se_step_pingpong_mods, false, true)

// ---------------------------------------------------------------------------------------------
// K3: lighting relaxation (operations.glsl:114-169), one thread per cell.
//   old_cells : material ids BEFORE this step (input_data)    new_cells : ids AFTER this step
//   light_in / light_out : float4 per cell, ping-pong (simulation.rs:239)
// The module is compiled with -fmad=false: every sum below is evaluated as written (no FMA), in the
// shader's neighbour order DOWN, UP, DOWNLEFT, UPLEFT, DOWNRIGHT, UPRIGHT, RIGHT, LEFT (math.glsl:154-166).
// ---------------------------------------------------------------------------------------------
struct SeLightParams {
    const unsigned* old_cells;
    const unsigned* new_cells;
    const float4* light_in;
    float4* light_out;
    int W, Hl, gy0, Hg;
};

extern "C" __global__ void __launch_bounds__(256) se_light(const SeLightParams p) {
    __shared__ unsigned fat_sm[256];
    {
        const int t = threadIdx.y * blockDim.x + threadIdx.x;
        if (t < 256) fat_sm[t] = se_fat_table[t];
    }
    __syncthreads();
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int yl = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= p.W || yl >= p.Hl) return;
    const int y = p.gy0 + yl;                     // global row
    const size_t idx = (size_t)yl * p.W + x;
    const unsigned me = min(p.new_cells[idx], 255u);
    float4 light;
    if (fat_sm[me] & SE_F_EMISSIVE) {             // :126-127
        light = make_float4(se_emission_table[me * 4 + 0], se_emission_table[me * 4 + 1], se_emission_table[me * 4 + 2], se_emission_table[me * 4 + 3]);
    } else if (y == 0) {                          // :128-129
        light = make_float4(1.0f, 1.0f, 1.0f, 0.999999f);
    } else {
        const int NX[8] = {0, 0, -1, -1, 1, 1, 1, -1};
        const int NY[8] = {1, -1, 1, -1, 1, -1, 0, 0};
        float4 avg = make_float4(0.f, 0.f, 0.f, 0.f), mx = make_float4(0.f, 0.f, 0.f, 0.f);
        float max_falloff = 0.0f;
        int num = 0;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const int nx = x + NX[n], ny = y + NY[n];
            if (nx < 0 || nx >= p.W || ny < 0 || ny >= p.Hg) continue;          // outOfBounds, :139-141
            const int nyl = ny - p.gy0;
            if (nyl < 0 || nyl >= p.Hl) continue;   // missing ghost row: result row is stale by schedule
            const size_t nidx = (size_t)nyl * p.W + nx;
            const unsigned nf = fat_sm[min(p.old_cells[nidx], 255u)];
            const float keep = (nf & SE_F_OBSTACLE) ? 0.0f : 1.0f;
            const float4 t = p.light_in[nidx];
            const float lr = t.x * keep, lg = t.y * keep, lb = t.z * keep, la = t.w * 1.0f;
            const float falloff = (la == 0.0f) ? max_falloff : la;
            const float vr = lr * la, vg = lg * la, vb = lb * la;
            avg.x += vr; avg.y += vg; avg.z += vb; avg.w += falloff;
            max_falloff = fmaxf(falloff, max_falloff);
            num += 1;
            mx.x = fmaxf(mx.x, vr); mx.y = fmaxf(mx.y, vg); mx.z = fmaxf(mx.z, vb); mx.w = fmaxf(mx.w, falloff);
        }
        if (num > 0) {
            const float dn = (float)num;
            avg.x = __fdiv_rn(avg.x, dn); avg.y = __fdiv_rn(avg.y, dn); avg.z = __fdiv_rn(avg.z, dn); avg.w = __fdiv_rn(avg.w, dn);
        }
        // mix(avg.rgb, max.rgb, 0.5) = avg*(1-0.5) + max*0.5
        light = make_float4(avg.x * 0.5f + mx.x * 0.5f, avg.y * 0.5f + mx.y * 0.5f, avg.z * 0.5f + mx.z * 0.5f, avg.w);
    }
    p.light_out[idx] = light;
}

// frame == 1: every cell becomes EMPTY (falling_sand.glsl:743-746); lighting (if on) then runs with
// new_cells == all-EMPTY through se_light.
extern "C" __global__ void __launch_bounds__(256) se_fill_cells(unsigned* cells, size_t n, unsigned value) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) cells[i] = value;
}
