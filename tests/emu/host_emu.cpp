// Test infrastructure: compiles the GENERATED CUDA rule header (rules_gen.cuh) and the pure device
// functions of kernels/sand_kernels.cuh (hash, block transition, transition-table build + lookup) on the
// HOST (g++), so that code-generation and table bugs can be found without a GPU.  Device intrinsics
// are shimmed with their exact host equivalents; the __global__ kernels are compiled out (SE_HOST_EMU).
// This is not a product path and not the oracle; tests/test_codegen_host_emulation.py compares it with
// the oracle.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#define SE_HOST_EMU 1
#define __device__
#define __forceinline__ inline
#define __restrict__
struct uint4 { unsigned x, y, z, w; };
struct float4 { float x, y, z, w; };
struct float2 { float x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float __uint2float_rn(unsigned u) { return (float)u; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return std::sqrt(a); }
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel) {
    unsigned long long v = ((unsigned long long)b << 32) | a;
    unsigned r = 0;
    for (int k = 0; k < 4; ++k) r |= (unsigned)((v >> (8 * ((sel >> (4 * k)) & 7))) & 0xFF) << (8 * k);
    return r;
}
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p += v; return o; }
static inline int atomicAdd(int* p, int v) { int o = *p; *p += v; return o; }
static inline unsigned atomicOr(unsigned* p, unsigned v) { unsigned o = *p; *p |= v; return o; }
using std::abs;

#include "sand_kernels.cuh"

// one in-place Margolus step with the generic (generated-code) block transition
extern "C" void emu_step_inplace(uint32_t* cells, int W, int H, int frame) {
    int ox, oy;
    se_margolus_offset(frame, ox, oy);
    for (int y0 = -oy; y0 < H; y0 += 2)
        for (int x0 = -ox; x0 < W; x0 += 2) {
            unsigned raw[4];
            for (int k = 0; k < 4; ++k) {
                int x = x0 + (k & 1), y = y0 + (k >> 1);
                raw[k] = (x < 0 || x >= W || y < 0 || y >= H) ? 2u : cells[(size_t)y * W + x];
            }
            unsigned q[4];
            for (int k = 0; k < 4; ++k) q[k] = se_fat_table[raw[k] < 255u ? raw[k] : 255u];
            if ((raw[0] | raw[1] | raw[2] | raw[3]) != 0u) se_block(q[0], q[1], q[2], q[3], x0, y0, frame);
            for (int k = 0; k < 4; ++k) {
                int x = x0 + (k & 1), y = y0 + (k >> 1);
                if (!(x < 0 || x >= W || y < 0 || y >= H)) cells[(size_t)y * W + x] = SE_ID(q[k]);
            }
        }
}

// D2: the per-cell modification override exactly as the *_mods kernels decide it (se_mod_lookup on the records that
// se_sim_step staged).  out[y * W + x] = overriding material id, or 0xFFFFFFFF when the cell is left to simulate().
extern "C" void emu_mod_override(const void* mods, int n_mods, int W, int H, uint32_t* out) {
    const SeMod* m = (const SeMod*)mods;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            unsigned mat = 0;
            out[(size_t)y * W + x] = se_mod_lookup(m, n_mods, x, y, mat) ? mat : 0xFFFFFFFFu;
        }
}

// K5 (colour shading): se_shade_cell over the grid (glibc sinf here, CUDA sinf on the device)
extern "C" void emu_shade(const uint32_t* cells, int W, int H, float* rgba) {
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const uint32_t raw = cells[(size_t)y * W + x];
            const float4 c = se_shade_cell(raw < 255u ? raw : 255u, x, y);
            float* o = rgba + 4 * ((size_t)y * W + x);
            o[0] = c.x; o[1] = c.y; o[2] = c.z; o[3] = c.w;
        }
}

// K3 (lighting): the two per-thread phases of se_light, run CTA by CTA with a "__syncthreads" between them
extern "C" void emu_light(const uint32_t* old_cells, const uint32_t* new_cells, const float* light_in, float* light_out,
                          int W, int Hl, int gy0, int Hg, int* n_interior_ctas) {
    SeLightParams p{old_cells, new_cells, reinterpret_cast<const float4*>(light_in), reinterpret_cast<float4*>(light_out), W, Hl, gy0, Hg};
    std::vector<float4> term(SE_LT_TERMS);
    int n_int = 0;
    for (int by = 0; by < (Hl + SE_LT_H - 1) / SE_LT_H; ++by)
        for (int bx = 0; bx < (W + SE_LT_W - 1) / SE_LT_W; ++bx) {
            for (auto& t : term) t = make_float4(NAN, NAN, NAN, NAN);      // poison: every term read must have been staged
            if (se_light_tile_is_interior(p, bx, by)) {
                ++n_int;
                for (int tid = 0; tid < 256; ++tid) se_light_stage<true>(p, se_fat_table, term.data(), bx, by, tid);
                for (int tid = 0; tid < 256; ++tid) se_light_compute<true>(p, se_fat_table, term.data(), bx, by, tid, SeNewIdFromGlobal{p.new_cells});
            } else {
                for (int tid = 0; tid < 256; ++tid) se_light_stage<false>(p, se_fat_table, term.data(), bx, by, tid);
                for (int tid = 0; tid < 256; ++tid) se_light_compute<false>(p, se_fat_table, term.data(), bx, by, tid, SeNewIdFromGlobal{p.new_cells});
            }
        }
    if (n_interior_ctas) *n_interior_ctas = n_int;
}

extern "C" int emu_lut_eligible(void) { return SE_LUT_ELIGIBLE; }

#if SE_LUT_ELIGIBLE
// table image as the kernels see it: base[N^4 * tables] u32 entries and pool[] of {thr, A, B} triples
static std::vector<unsigned> g_base, g_pool;
static unsigned g_pool_entries = 0;
static SeTab g_tab{nullptr, nullptr};
static unsigned g_flag_pop = 0;      // 1: build the census variant of the table (population-changing outcomes flagged)

// returns the number of pool entries (two passes like se_sim_create in mode 2: count, then fill)
extern "C" int emu_build_lut(void) {
    g_base.assign((size_t)SE_LUT_ENTRIES, 0u);
    unsigned counter = 0;
    for (unsigned idx = 0; idx < (unsigned)SE_LUT_ENTRIES; ++idx) se_build_lut_entry(idx, g_base.data(), nullptr, &counter, 0u, g_flag_pop);
    g_pool_entries = counter;
    g_pool.assign((size_t)4 * (counter + 1), 0u);
    counter = 0;
    for (unsigned idx = 0; idx < (unsigned)SE_LUT_ENTRIES; ++idx) se_build_lut_entry(idx, g_base.data(), g_pool.data(), &counter, g_pool_entries, g_flag_pop);
    g_tab = SeTab{g_base.data(), g_pool.data()};
    return (int)g_pool_entries;
}
extern "C" int emu_lut_mode(void) { return SE_LUT_MODE; }
extern "C" int emu_lut_two_tables(void) { return SE_LUT_TWO_TABLES; }

// one in-place Margolus step through the transition table (se_block_lut), ids packed as in the tile kernel
extern "C" void emu_step_lut_inplace(uint32_t* cells, int W, int H, int frame) {
    int ox, oy;
    se_margolus_offset(frame, ox, oy);
    for (int y0 = -oy; y0 < H; y0 += 2)
        for (int x0 = -ox; x0 < W; x0 += 2) {
            unsigned v = 0;
            for (int k = 0; k < 4; ++k) {
                int x = x0 + (k & 1), y = y0 + (k >> 1);
                unsigned id = (x < 0 || x >= W || y < 0 || y >= H) ? 2u : cells[(size_t)y * W + x];
                if (id >= SE_N_MATERIALS) id = 1u;
                v |= id << (8 * k);
            }
            unsigned nv = v;
            {   // like the tile kernel: no all-EMPTY branch, T0[0] == 0
                const unsigned seed = (unsigned)x0 * 461u + (unsigned)y0 * 2131u + (unsigned)frame * (2131u * 2131u);
                nv = se_block_lut(v, seed, x0, y0, frame, g_tab, se_fat_table);
            }
            for (int k = 0; k < 4; ++k) {
                int x = x0 + (k & 1), y = y0 + (k >> 1);
                if (!(x < 0 || x >= W || y < 0 || y >= H)) cells[(size_t)y * W + x] = (nv >> (8 * k)) & 0xFFu;
            }
        }
}

#if !SE_LUT_TWO_TABLES
// ---- running census: flagged table + per-block deltas, decisions exactly as in se_k1c_body<true> ----
// (re)builds the table with SE_E_POPFLAG on the outcomes that are not permutations; returns how many plain entries carry it
extern "C" int emu_build_census_lut(void) {
    g_flag_pop = 1;
    emu_build_lut();
    g_flag_pop = 0;
    int n = 0;
    for (unsigned e : g_base) n += (!(e & SE_E_SPECIAL) && (e & SE_E_POPFLAG)) ? 1 : 0;
    return n;
}

// one in-place table step of a full grid (gy0 = 0) that also maintains census256 for the owned rows [own_y0, own_y1);
// n_filtered counts the changed, fully counted blocks that the popbits filter let skip the per-cell comparison
extern "C" void emu_step_lut_census(uint32_t* cells, int W, int H, int frame, int own_y0, int own_y1, long long* census256, long long* n_filtered) {
    int ox, oy;
    se_margolus_offset(frame, ox, oy);
    int hist[256] = {0};
    for (int y0 = -oy; y0 < H; y0 += 2)
        for (int x0 = -ox; x0 < W; x0 += 2) {
            const int y1 = y0 + 1;
            const int st0 = (y0 < 0) ? 1 : 0, st1 = (y1 >= H) ? 1 : 0;
            const bool cx0 = x0 >= 0, cx1 = (x0 + 1) < W;
            unsigned raw[4] = {2u, 2u, 2u, 2u};
            if (st0 == 0 && cx0) raw[0] = cells[(size_t)y0 * W + x0];
            if (st0 == 0 && cx1) raw[1] = cells[(size_t)y0 * W + x0 + 1];
            if (st1 == 0 && cx0) raw[2] = cells[(size_t)y1 * W + x0];
            if (st1 == 0 && cx1) raw[3] = cells[(size_t)y1 * W + x0 + 1];
            unsigned v = 0;
            for (int k = 0; k < 4; ++k) v |= (raw[k] < SE_N_MATERIALS ? raw[k] : 1u) << (8 * k);
            const unsigned seed = (unsigned)x0 * 461u + (unsigned)y0 * 2131u + (unsigned)frame * (2131u * 2131u);
            // as in the kernel: the entry's flag decides whether the block is looked at; generated-code blocks always are
            const bool mirror = se_hashi(seed * 213u) <= SE_MIRROR_UMAX;
            const unsigned sel = mirror ? 0x2301u : 0x3210u;
            unsigned e = se_tab_entry(g_tab, se_idx4(__byte_perm(v, 0u, sel)));
            bool look = false;
            for (int k = 0; k < 4; ++k) look = look || raw[k] >= (unsigned)SE_N_MATERIALS;
            if (e & SE_E_SPECIAL) {
                if (e & SE_E_SLOW) look = true;
                e = se_block_special(e, v, seed, sel, 0, 0, 0, g_tab, se_fat_table);
            }
            if (e & SE_E_POPFLAG) look = true;
            e &= ~SE_E_POPFLAG;
            const unsigned nv = __byte_perm(e, 0u, sel);
            const unsigned cm_rows = ((st0 == 0 && y0 >= own_y0 && y0 < own_y1) ? 3u : 0u) | ((st1 == 0 && y1 >= own_y0 && y1 < own_y1) ? 12u : 0u);
            const unsigned cm_cols = (cx0 ? 5u : 0u) | (cx1 ? 10u : 0u);
            const unsigned cm = cm_rows & cm_cols;
            if (n_filtered && cm == 0xFu && nv != v && !look) ++*n_filtered;
            se_census_block(hist, look, raw[0], raw[1], raw[2], raw[3], nv, cm);
            const unsigned nn[4] = {nv & 0xFFu, (nv >> 8) & 0xFFu, (nv >> 16) & 0xFFu, nv >> 24};
            if (st0 == 0 && cx0 && nn[0] != raw[0]) cells[(size_t)y0 * W + x0] = nn[0];
            if (st0 == 0 && cx1 && nn[1] != raw[1]) cells[(size_t)y0 * W + x0 + 1] = nn[1];
            if (st1 == 0 && cx0 && nn[2] != raw[2]) cells[(size_t)y1 * W + x0] = nn[2];
            if (st1 == 0 && cx1 && nn[3] != raw[3]) cells[(size_t)y1 * W + x0 + 1] = nn[3];
        }
    for (int i = 0; i < 256; ++i) census256[i] += hist[i];
}
#else
extern "C" int emu_build_census_lut(void) { return -2; }
extern "C" void emu_step_lut_census(uint32_t*, int, int, int, int, int, long long*, long long*) {}
#endif
#else
extern "C" int emu_lut_mode(void) { return 0; }
extern "C" int emu_lut_two_tables(void) { return 0; }
extern "C" int emu_build_census_lut(void) { return -2; }
extern "C" void emu_step_lut_census(uint32_t*, int, int, int, int, int, long long*, long long*) {}
extern "C" int emu_build_lut(void) { return -2; }
extern "C" void emu_step_lut_inplace(uint32_t*, int, int, int) {}
#endif
