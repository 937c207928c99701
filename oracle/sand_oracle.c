/* ORACLE -- test infrastructure, NOT product code.
 *
 * Plain-C restatement of the reference's compute shader
 *     /root/reference/shaders/compute/gen/falling_sand.glsl
 * (the single self-contained shader `Simulation::new` compiles, sandengine-core/src/simulation.rs:130).
 * Every function cites the shader lines it follows.  The rule functions and material tables are
 * NOT in this file: they are generated from the YAML by oracle/oracle_lang.py::emit_c_rules() into
 * "rules_gen.h", the same split the reference has between hand-written GLSL and gen/{materials,rules}.glsl.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this.  Build: oracle/build_oracle.py (gcc -O2 -march=x86-64-v3 -ffp-contract=off -fopenmp).  -ffp-contract=off is
 * REQUIRED: the lighting sums must be evaluated as written (no FMA), SURVEY.md section 7 "hard parts".
 *
 * Parity status: PINNED by the reference's own shader.  The reference ships no state-level tests or fixtures
 *  and its GLSL cannot run as GLSL in this image, but oracle/build_ref.py compiles the shader
 * text itself for the CPU (oracle/_ref/, GLSL-types shim + syntactic translation); this file agrees with it bit for bit
 * on cell ids and with error 0 on light for every case of tests/ref_cases.py (tests/test_ref_shader.py), and with the
 * committed outputs of that shader (tests/golden/ref_shader_goldens.json) where /root/reference is absent.  Also kept:
 * the survey's cross-check vectors (tests/test_oracle_kat.py) and the independently written pure-Python restatement
 * oracle/pyoracle.py.  The one exception: LEFT-rule semantics are a definition (oracle_lang.py docstring) -- the
 * reference cannot compile such rules -- i.e. "parity unpinned" for that row only.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* GLSL vec4 with its component aliases, so that rule text such as `self.mat.emission.g` compiles as C */
typedef union {
    struct { float x, y, z, w; };
    struct { float r, g, b, a; };
    float v[4];
} vec4;

typedef struct {            /* falling_sand.glsl:211-218 */
    int id;
    vec4 color;
    float density;
    vec4 emission;
    int type;
} Material;

typedef struct { int x, y; } ivec2;

typedef struct {            /* falling_sand.glsl:221-224 */
    Material mat;
    ivec2 pos;
} Cell;

typedef struct {            /* falling_sand.glsl:346-351; std140 stride 32 B == simulation.rs:45-56 */
    int32_t position[2];
    int32_t mod_shape;
    int32_t mod_size;
    int32_t mod_matID;
    int32_t _pad4[3];
} SimModification;

#define MODSHAPE_CIRCLE 0   /* falling_sand.glsl:343-344 */
#define MODSHAPE_SQUARE 1
#define MAX_MODIFICATIONS 256   /* simulation.rs:43; falling_sand.glsl:354 */

static inline Cell newCell(Material mat, ivec2 pos) {   /* falling_sand.glsl:226-228 */
    Cell c; c.mat = mat; c.pos = pos; return c;
}

static void swap_cells(Cell* a, Cell* b);

/* `uniform int frame` of the shader (falling_sand.glsl:334): visible to the generated rule text; set by the
 * step drivers below before the parallel loops (read-only inside them). */
static int frame = 0;

/* Scalar GLSL built-ins for rule conditions (GLSL 4.30 section 8.3 definitions; oracle_lang._to_c renames the calls).
 * Type-generic like GLSL's overloads: an int argument list stays int, anything with a float is float. */
static inline float se_minf(float a, float b) { return b < a ? b : a; }
static inline int se_mini(int a, int b) { return b < a ? b : a; }
static inline float se_maxf(float a, float b) { return a < b ? b : a; }
static inline int se_maxi(int a, int b) { return a < b ? b : a; }
static inline float se_signf(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
static inline int se_signi(int x) { return x > 0 ? 1 : (x < 0 ? -1 : 0); }
static inline int se_absi(int x) { return x < 0 ? -x : x; }
static inline float se_mod(float x, float y) { return x - y * floorf(x / y); }
static inline float se_fract(float x) { return x - floorf(x); }
static inline float se_step(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
#define se_abs(x) _Generic((x), float: fabsf, double: fabsf, default: se_absi)(x)
#define se_sign(x) _Generic((x), float: se_signf, double: se_signf, default: se_signi)(x)
#define se_min(a, b) _Generic((a) + (b), float: se_minf, double: se_minf, default: se_mini)((a), (b))
#define se_max(a, b) _Generic((a) + (b), float: se_maxf, double: se_maxf, default: se_maxi)((a), (b))
#define se_clamp(x, lo, hi) se_min(se_max((x), (lo)), (hi))
#define se_floor(x) floorf(x)
#define se_ceil(x) ceilf(x)
#define se_sqrt(x) sqrtf(x)
#define se_float(x) ((float)(x))
#define se_int(x) ((int)(x))

#include "rules_gen.h"   /* TYPE_*, isType_*, MATERIALS[], MAT_*, rule_*, apply{Mirrored,Left,Right}Rules */

/* falling_sand.glsl:371-378 -- guarded swap: no-op when either side is WALL or NULL typed. */
static void swap_cells(Cell* a, Cell* b) {
    if (a->mat.type == TYPE_WALL || b->mat.type == TYPE_WALL || a->mat.type == TYPE_NULL || b->mat.type == TYPE_NULL) {
        return;
    }
    Cell tmp = *a; *a = *b; *b = tmp;
}

/* falling_sand.glsl:310-317 -- linear search over materials(); unknown id => MAT_NULL.
 * ids are 0..N-1 in table order (materials.rs:91,196), so the search is a bounds check. */
static inline Material getMaterialFromID(int id) {
    if (id >= 0 && id < N_MATERIALS) return MATERIALS[id];
    return MAT_NULL;
}

/* ---- integer hash, falling_sand.glsl:58-79, 115-120 ---- */
uint32_t so_hashi(uint32_t x) {                 /* :58-66 */
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
static inline void hash4i(uint32_t y, uint32_t out[4]) {   /* :69-79 */
    static const uint32_t mul[4] = {213u, 2131u, 21313u, 213132u};
    for (int i = 0; i < 4; i++) out[i] = so_hashi(y * mul[i]);
}
/* :115-120.  float(0xffffffffU) rounds to 2^32, so the division is an exact scaling; the only
 * rounding is the uint->float conversion (round-to-nearest-even). */
static inline void hash43(int px, int py, int pz, float out[4], uint32_t lanes[4]) {
    uint32_t x = (uint32_t)px * 461u + (uint32_t)py * 2131u + (uint32_t)pz * 2131u * 2131u;
    hash4i(x, lanes);
    const float denom = (float)0xffffffffU;
    for (int i = 0; i < 4; i++) out[i] = (float)lanes[i] / denom;
}
/* exported for the hash known-answer tests */
void so_hash43(int px, int py, int pz, uint32_t* seed, uint32_t lanes[4], float r[4]) {
    *seed = (uint32_t)px * 461u + (uint32_t)py * 2131u + (uint32_t)pz * 2131u * 2131u;
    hash43(px, py, pz, r, lanes);
}

/* ---- simulation context (uniforms + bound textures of the shader) ---- */
typedef struct {
    const uint32_t* input_data;   /* W*H material ids (the .r channel of the RGBA32F texture) */
    const float* input_light;     /* W*H*4 or NULL */
    int W, H, frame;
    const SimModification* mods;  /* MAX_MODIFICATIONS entries */
} Ctx;

static inline int outOfBounds(const Ctx* c, int x, int y) {   /* :362-364 */
    return x >= c->W || x < 0 || y >= c->H || y < 0;
}

static inline void getMargolusOffset(int frame, int off[2]) {   /* :380-389 */
    frame = frame % 4;
    if (frame == 1) { off[0] = 1; off[1] = 1; }
    else if (frame == 2) { off[0] = 0; off[1] = 1; }
    else if (frame == 3) { off[0] = 1; off[1] = 0; }
    else { off[0] = 0; off[1] = 0; }
}

static inline Cell getCell(const Ctx* c, int x, int y) {   /* :400-412, SCREEN_IS_BORDER defined (:4) */
    ivec2 pos = {x, y};
    if (outOfBounds(c, x, y)) return newCell(MAT_WALL, pos);
    int matID = (int)c->input_data[(size_t)y * c->W + x];
    return newCell(getMaterialFromID(matID), pos);
}

static inline int emission_rgb_zero(const Material* m) {
    return m->emission.v[0] == 0.0f && m->emission.v[1] == 0.0f && m->emission.v[2] == 0.0f;
}
static inline int isLightObstacle(const Cell* cell) {   /* :423-425 */
    return emission_rgb_zero(&cell->mat) && !isType_EMPTY(*cell);
}

/* Margolus block transition shared by the per-cell and per-block drivers: falling_sand.glsl:692-718.
 * cells = self,right,down,downright at pos_rounded+{(0,0),(1,0),(0,1),(1,1)}.  Returns 0 on the
 * all-EMPTY early-out (:692-694), in which case the caller's result is MAT_EMPTY. */
static inline int block_transition(Cell* self, Cell* right, Cell* down, Cell* downright, ivec2 pos_rounded, int frame) {
    if (self->mat.id == MAT_EMPTY.id && right->mat.id == MAT_EMPTY.id && down->mat.id == MAT_EMPTY.id && downright->mat.id == MAT_EMPTY.id) {
        return 0;
    }
    float rand[4]; uint32_t lanes[4];
    hash43(pos_rounded.x, pos_rounded.y, frame, rand, lanes);   /* :698 (rand2, up, upright are unused) */
    int shouldMirror = rand[0] < 0.5f;                            /* :701 */
    if (shouldMirror) { swap_cells(self, right); swap_cells(down, downright); }   /* :702-705 */
    applyMirroredRules(self, right, down, downright, rand, pos_rounded);           /* :707 */
#if HAVE_LEFT_RULES
    /* Left rules run in the mirrored view, before the un-mirror (definition). */
    if (shouldMirror) applyLeftRules(self, right, down, downright, rand, pos_rounded);
#endif
    if (shouldMirror) { swap_cells(self, right); swap_cells(down, downright); }   /* :709-712 */
    if (!shouldMirror) applyRightRules(self, right, down, downright, rand, pos_rounded);   /* :714-718 */
    return 1;
}

/* falling_sand.glsl:676-733 */
static Cell simulate(const Ctx* c, int gx, int gy) {
    int off[2]; getMargolusOffset(c->frame, off);
    int pos[2] = {gx + off[0], gy + off[1]};
    ivec2 pos_rounded = {(pos[0] / 2) * 2, (pos[1] / 2) * 2};
    int marg_idx = (pos[0] & 1) + (pos[1] & 1) * 2;
    pos_rounded.x -= off[0]; pos_rounded.y -= off[1];

    Cell self = getCell(c, pos_rounded.x, pos_rounded.y);
    Cell right = getCell(c, pos_rounded.x + 1, pos_rounded.y);
    Cell down = getCell(c, pos_rounded.x, pos_rounded.y + 1);
    Cell downright = getCell(c, pos_rounded.x + 1, pos_rounded.y + 1);

    if (!block_transition(&self, &right, &down, &downright, pos_rounded, c->frame)) {
        return newCell(MAT_EMPTY, pos_rounded);
    }
    switch (marg_idx) {   /* :720-729 */
        case 0: return self;
        case 1: return right;
        case 2: return down;
        default: return downright;
    }
}

/* ---- colour shading, falling_sand.glsl:81-85, 123-148, 455-463 (render product; tolerance only) ---- */
static inline float fractf(float x) { return x - floorf(x); }
static void old_hash2(float px, float py, float out[2]) {   /* :81-85 */
    float a = px * 127.1f + py * 311.7f;
    float b = px * 269.5f + py * 183.3f;
    out[0] = -1.0f + 2.0f * fractf(sinf(a) * 43758.5453123f);
    out[1] = -1.0f + 2.0f * fractf(sinf(b) * 43758.5453123f);
}
static float _noise(float px, float py) {   /* :123-137 */
    const float K1 = 0.366025404f, K2 = 0.211324865f;
    float s = (px + py) * K1;
    float ix = floorf(px + s), iy = floorf(py + s);
    float t = (ix + iy) * K2;
    float ax = px - ix + t, ay = py - iy + t;
    float m = (ax < ay) ? 0.0f : 1.0f;          /* step(a.y, a.x) */
    float ox = m, oy = 1.0f - m;
    float bx = ax - ox + K2, by = ay - oy + K2;
    float cx = ax - 1.0f + 2.0f * K2, cy = ay - 1.0f + 2.0f * K2;
    float h0 = fmaxf(0.5f - (ax * ax + ay * ay), 0.0f);
    float h1 = fmaxf(0.5f - (bx * bx + by * by), 0.0f);
    float h2 = fmaxf(0.5f - (cx * cx + cy * cy), 0.0f);
    float g0[2], g1[2], g2[2];
    old_hash2(ix + 0.0f, iy + 0.0f, g0);
    old_hash2(ix + ox, iy + oy, g1);
    old_hash2(ix + 1.0f, iy + 1.0f, g2);
    float n0 = h0 * h0 * h0 * h0 * (ax * g0[0] + ay * g0[1]);
    float n1 = h1 * h1 * h1 * h1 * (bx * g1[0] + by * g1[1]);
    float n2 = h2 * h2 * h2 * h2 * (cx * g2[0] + cy * g2[1]);
    return 0.25f + 0.5f * (n0 * 70.0f + n1 * 70.0f + n2 * 70.0f);
}
static float noise(float px, float py, int octaves, float lacunarity, float frequency) {   /* :139-148 */
    float f = 0.0f;
    for (int o = 1; o < octaves + 1; o++) {
        f += 1.0f / (float)o * _noise(px * frequency, py * frequency);
        px *= lacunarity; py *= lacunarity;
    }
    return f;
}
static inline float clamp01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }

/* ---- outputs of one dispatch ---- */
typedef struct {
    uint32_t* output_data;   /* W*H */
    float* output_light;     /* W*H*4 or NULL (lighting not evaluated) */
    float* output_color;     /* W*H*4 or NULL */
} Out;

/* falling_sand.glsl:454-526 */
static void setCell(const Ctx* c, const Out* o, int x, int y, const Material* mat) {
    size_t idx = (size_t)y * c->W + x;
    if (o->output_color) {   /* :455-463, :525 */
        float col[4] = {mat->color.v[0], mat->color.v[1], mat->color.v[2], mat->color.v[3]};
        if (mat->id != MAT_EMPTY.id) {
            float rnd = noise((float)x, (float)y, 3, 2.0f, 0.25f) * 0.25f;
            col[0] = clamp01(col[0] - rnd); col[1] = clamp01(col[1] - rnd); col[2] = clamp01(col[2] - rnd);
        }
        memcpy(o->output_color + idx * 4, col, sizeof col);
    }
    o->output_data[idx] = (uint32_t)mat->id;   /* :466-467 */
    if (!o->output_light) return;

    /* :195-207 neighbour order: DOWN, UP, DOWNLEFT, UPLEFT, DOWNRIGHT, UPRIGHT, RIGHT, LEFT */
    static const int NX[8] = {0, 0, -1, -1, 1, 1, 1, -1};
    static const int NY[8] = {1, -1, 1, -1, 1, -1, 0, 0};
    float light[4];
    if (!emission_rgb_zero(mat)) {                      /* :481-482 */
        memcpy(light, mat->emission.v, sizeof light);
    } else if (y == 0) {                                /* :483-484 */
        light[0] = light[1] = light[2] = 1.0f; light[3] = 0.999999f;
    } else {                                            /* :485-523 */
        float avg[4] = {0, 0, 0, 0}, mx[4] = {0, 0, 0, 0};
        float max_falloff = 0.0f;
        int num = 0;
        for (int n = 0; n < 8; n++) {
            int nx = x + NX[n], ny = y + NY[n];
            if (outOfBounds(c, nx, ny)) continue;       /* :494-496 */
            Cell neigh = getCell(c, nx, ny);            /* old materials (input_data), :469-473 */
            int obst = isLightObstacle(&neigh);
            const float* tl = c->input_light + ((size_t)ny * c->W + nx) * 4;
            float keep = obst ? 0.0f : 1.0f;            /* vec4(vec3(float(!obst)), 1.0), :498 */
            float ld[4] = {tl[0] * keep, tl[1] * keep, tl[2] * keep, tl[3] * 1.0f};
            float falloff = (ld[3] == 0.0f) ? max_falloff : ld[3];   /* :500-505 */
            float l[4] = {ld[0] * ld[3], ld[1] * ld[3], ld[2] * ld[3], falloff};   /* :506 */
            for (int k = 0; k < 4; k++) avg[k] += l[k];               /* :507 */
            max_falloff = fmaxf(falloff, max_falloff);                /* :508 */
            num += 1;
            for (int k = 0; k < 4; k++) mx[k] = fmaxf(mx[k], l[k]);   /* :512-513 */
        }
        if (num > 0) { float d = (float)num; for (int k = 0; k < 4; k++) avg[k] /= d; }   /* :516-518 */
        /* mix(a, b, 0.5) = a*(1-0.5) + b*0.5, :521 */
        for (int k = 0; k < 3; k++) light[k] = avg[k] * 0.5f + mx[k] * 0.5f;
        light[3] = avg[3];
    }
    memcpy(o->output_light + idx * 4, light, sizeof light);   /* :524 */
}

/* falling_sand.glsl:737-799 -- one invocation of main() for cell (x, y). */
static void shader_main(const Ctx* c, const Out* o, int x, int y) {
    if (x >= c->W || x < 0 || y >= c->H || y < 0) return;   /* :739-741 */
    if (c->frame == 1) {                                     /* :743-746 */
        setCell(c, o, x, y, &MAT_EMPTY);
        return;
    }
    int got_modified = 0;
    Material final_mat = MAT_NULL;
    for (int i = 0; i < MAX_MODIFICATIONS; i++) {            /* :752-777 */
        const SimModification* mod = &c->mods[i];
        if (mod->mod_size == 0) break;
        int dx = abs(mod->position[0] - x), dy = abs(mod->position[1] - y);
        Material mat = getMaterialFromID(mod->mod_matID);
        switch (mod->mod_shape) {
            case MODSHAPE_CIRCLE: {
                /* sqrt(pow(dx,2) + pow(dy,2)) <= size ; pow(d,2) taken as d*d */
                float fx = (float)dx, fy = (float)dy;
                float dist = sqrtf(fx * fx + fy * fy);
                if (dist <= (float)mod->mod_size) { got_modified = 1; final_mat = mat; }
                break;
            }
            case MODSHAPE_SQUARE:
                if (dx <= mod->mod_size && dy <= mod->mod_size) { got_modified = 1; final_mat = mat; }
                break;
            default: break;
        }
    }
    if (got_modified && final_mat.id != MAT_NULL.id) {      /* :791-794 */
        setCell(c, o, x, y, &final_mat);
        return;
    }
    Cell result = simulate(c, x, y);                         /* :797-798 */
    setCell(c, o, x, y, &result.mat);
}

/* One dispatch of the shader over the whole grid (simulation.rs:220-234).  `mods` holds n_mods
 * entries; like the host (simulation.rs:203-208) only the first 256 are used, the remaining UBO
 * slots have mod_size == 0.  `frame` is the value AFTER the host's increment (simulation.rs:201). */
void so_step_cells(const uint32_t* in_cells, uint32_t* out_cells, const float* in_light, float* out_light,
                   float* out_color, int W, int H, int frame_, const SimModification* mods, int n_mods) {
    SimModification ubo[MAX_MODIFICATIONS];
    memset(ubo, 0, sizeof ubo);
    if (n_mods > MAX_MODIFICATIONS) n_mods = MAX_MODIFICATIONS;
    if (n_mods > 0) memcpy(ubo, mods, (size_t)n_mods * sizeof(SimModification));
    Ctx c = {in_cells, in_light, W, H, frame_, ubo};
    frame = frame_;
    Out o = {out_cells, (in_light && out_light) ? out_light : NULL, out_color};
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) shader_main(&c, &o, x, y);
}

/* Per-block, in-place form of the same update (lighting off, no modifications, frame != 1): each
 * Margolus block is evaluated once and its in-grid cells are written back.  Equivalent to
 * so_step_cells because a cell's result depends only on its own block (falling_sand.glsl:687-690)
 * and blocks partition the grid; tests/test_oracle_kat.py checks the two drivers against each other.
 * This is the form timed as the CPU baseline (it does a quarter of the shader's redundant work). */
void so_step_blocks_inplace(uint32_t* cells, int W, int H, int frame_) {
    frame = frame_;
    int off[2]; getMargolusOffset(frame, off);
    Ctx c = {cells, NULL, W, H, frame, NULL};
    int nby = (H + off[1] + 1) / 2, nbx = (W + off[0] + 1) / 2;
#pragma omp parallel for schedule(static)
    for (int by = 0; by < nby; by++) {
        for (int bx = 0; bx < nbx; bx++) {
            ivec2 pr = {bx * 2 - off[0], by * 2 - off[1]};
            Cell q[4];
            q[0] = getCell(&c, pr.x, pr.y);     q[1] = getCell(&c, pr.x + 1, pr.y);
            q[2] = getCell(&c, pr.x, pr.y + 1); q[3] = getCell(&c, pr.x + 1, pr.y + 1);
            if (!block_transition(&q[0], &q[1], &q[2], &q[3], pr, frame)) continue;   /* all EMPTY stays EMPTY */
            for (int k = 0; k < 4; k++) {
                int x = pr.x + (k & 1), y = pr.y + (k >> 1);
                if (!outOfBounds(&c, x, y)) cells[(size_t)y * W + x] = (uint32_t)q[k].mat.id;
            }
        }
    }
}

/* n_steps of the host loop (simulation.rs:195-253) without modifications or lighting:
 * frame += 1, dispatch, (ping-pong is implicit in the in-place form). Returns the final frame. */
int so_run_blocks(uint32_t* cells, int W, int H, int frame0, int n_steps) {
    for (int s = 0; s < n_steps; s++) {
        frame0 += 1;
        if (frame0 == 1) { memset(cells, 0, (size_t)W * H * sizeof(uint32_t)); continue; }
        so_step_blocks_inplace(cells, W, H, frame0);
    }
    return frame0;
}

/* Strip form of the per-block update, used ONLY by the CPU (gloo) emulation of the multi-GPU ghost-row
 * schedule (tests/test_distributed_gloo.py): `cells` holds local rows [gy0, gy0 + Hl) of a grid of Hg rows.
 * Rows outside the global grid read as WALL; a block that needs a row inside the grid but outside the
 * local buffer is skipped (its in-buffer row goes stale -- exactly what the ghost-row schedule accounts for).
 * Positions fed to the hash are GLOBAL (falling_sand.glsl:698). */
void so_step_blocks_strip(uint32_t* cells, int W, int Hl, int gy0, int Hg, int frame_) {
    frame = frame_;
    int off[2]; getMargolusOffset(frame, off);
    int jb0 = (gy0 + off[1]) / 2;
    int y_end = (gy0 + Hl < Hg) ? gy0 + Hl : Hg;
    int nby = (y_end + off[1] + 1) / 2 - jb0, nbx = (W + off[0] + 1) / 2;
#pragma omp parallel for schedule(static)
    for (int jl = 0; jl < nby; jl++) {
        int y0 = 2 * (jb0 + jl) - off[1], y1 = y0 + 1;
        int st0 = (y0 < 0) ? 1 : (y0 < gy0 ? 2 : 0);
        int st1 = (y1 >= Hg) ? 1 : (y1 >= gy0 + Hl ? 2 : 0);
        if (st0 == 2 || st1 == 2 || y0 >= y_end) continue;
        for (int bx = 0; bx < nbx; bx++) {
            ivec2 pr = {bx * 2 - off[0], y0};
            Cell q[4];
            for (int k = 0; k < 4; k++) {
                int x = pr.x + (k & 1), y = pr.y + (k >> 1);
                ivec2 pos = {x, y};
                if (x < 0 || x >= W || y < 0 || y >= Hg) q[k] = newCell(MAT_WALL, pos);
                else q[k] = newCell(getMaterialFromID((int)cells[(size_t)(y - gy0) * W + x]), pos);
            }
            if (!block_transition(&q[0], &q[1], &q[2], &q[3], pr, frame)) continue;
            for (int k = 0; k < 4; k++) {
                int x = pr.x + (k & 1), y = pr.y + (k >> 1);
                if (!(x < 0 || x >= W || y < 0 || y >= Hg)) cells[(size_t)(y - gy0) * W + x] = (uint32_t)q[k].mat.id;
            }
        }
    }
}

int so_n_materials(void) { return N_MATERIALS; }
