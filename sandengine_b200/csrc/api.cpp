// C ABI implementation (include/sandengine_b200.h): rule compilation (front end + NVRTC) and the
// simulation host that replaces sandengine-core's `Simulation` (/root/reference/sandengine-core/src/simulation.rs).
//
// Device work is launched through the CUDA driver API on functions loaded from the NVRTC-built cubin;
// the driver entry points are resolved with cudaGetDriverEntryPoint so the library does not link
// libcuda and can be loaded (and can compile rules) on a host without a GPU.
#include "../../include/sandengine_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <nvrtc.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "lang/codegen.h"
#include "lang/lang.h"
#include "static_kernels.h"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

// No C++ exception may cross the C ABI (std::bad_alloc from the host-side containers, anything unexpected from the
// front end): every extern "C" entry point is a function-try-block that ends in SE_ABI_CATCH.
int fail_nothrow(int code, const char* where, const char* what) noexcept {
    try {
        g_err = std::string(where) + ": " + what;
    } catch (...) {
        g_err.clear();
    }
    return code;
}
#define SE_ABI_CATCH(NAME)                                                                                  \
    catch (const std::bad_alloc&) { return fail_nothrow(SE_ERR_INTERNAL, NAME, "out of host memory"); }     \
    catch (const std::exception& e_) { return fail_nothrow(SE_ERR_INTERNAL, NAME, e_.what()); }             \
    catch (...) { return fail_nothrow(SE_ERR_INTERNAL, NAME, "unknown C++ exception"); }

int map_kind(se::ErrKind k) {
    switch (k) {
        case se::ErrKind::Yaml: return SE_ERR_YAML;
        case se::ErrKind::MissingField: return SE_ERR_MISSING_FIELD;
        case se::ErrKind::InvalidType: return SE_ERR_INVALID_TYPE;
        case se::ErrKind::NotFound: return SE_ERR_NOT_FOUND;
        case se::ErrKind::NotRecognized: return SE_ERR_NOT_RECOGNIZED;
        case se::ErrKind::Unsupported: return SE_ERR_UNSUPPORTED;
    }
    return SE_ERR_INVALID_ARG;
}

const char* const KERNEL_SOURCE =
#include "_gen/sand_kernels_embed.inc"
    ;

// ---- driver API through the runtime (no libcuda link dependency) -----------------------------
struct Driver {
    decltype(&cuModuleLoadData) ModuleLoadData = nullptr;
    decltype(&cuModuleUnload) ModuleUnload = nullptr;
    decltype(&cuModuleGetFunction) ModuleGetFunction = nullptr;
    decltype(&cuLaunchKernel) LaunchKernel = nullptr;
    decltype(&cuLaunchCooperativeKernel) LaunchCooperativeKernel = nullptr;
    decltype(&cuFuncSetAttribute) FuncSetAttribute = nullptr;
    decltype(&cuGetErrorString) GetErrorString = nullptr;
    decltype(&cuStreamWriteValue32) StreamWriteValue32 = nullptr;
    decltype(&cuStreamWaitValue32) StreamWaitValue32 = nullptr;
    decltype(&cuOccupancyMaxActiveBlocksPerMultiprocessor) OccupancyMaxActiveBlocks = nullptr;
    bool ok = false;
    std::string why;
};

Driver& driver() {
    static Driver d;
    static std::once_flag once;
    std::call_once(once, [] {
        auto get = [&](const char* name, void** fn) {
            cudaDriverEntryPointQueryResult q;
            cudaError_t e = cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q);
            if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !*fn) {
                d.why = std::string("cudaGetDriverEntryPoint(") + name + ") failed: " + cudaGetErrorString(e);
                (void)cudaGetLastError();
                return false;
            }
            return true;
        };
        d.ok = get("cuModuleLoadData", (void**)&d.ModuleLoadData) && get("cuModuleUnload", (void**)&d.ModuleUnload) &&
               get("cuModuleGetFunction", (void**)&d.ModuleGetFunction) && get("cuLaunchKernel", (void**)&d.LaunchKernel) &&
               get("cuLaunchCooperativeKernel", (void**)&d.LaunchCooperativeKernel) &&
               get("cuFuncSetAttribute", (void**)&d.FuncSetAttribute) && get("cuGetErrorString", (void**)&d.GetErrorString) &&
               get("cuStreamWriteValue32", (void**)&d.StreamWriteValue32) && get("cuStreamWaitValue32", (void**)&d.StreamWaitValue32) &&
               get("cuOccupancyMaxActiveBlocksPerMultiprocessor", (void**)&d.OccupancyMaxActiveBlocks);
    });
    return d;
}

std::string cu_err(CUresult r) {
    const char* s = nullptr;
    if (driver().GetErrorString) driver().GetErrorString(r, &s);
    return s ? s : ("CUresult " + std::to_string((int)r));
}

#define SE_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) return fail(SE_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)
#define SE_CU(call)                                                                   \
    do {                                                                              \
        CUresult r_ = (call);                                                         \
        if (r_ != CUDA_SUCCESS) return fail(SE_ERR_CUDA, std::string(#call) + ": " + cu_err(r_)); \
    } while (0)

// mirrors the device structs of kernels/sand_kernels.cuh
struct SeMod { int px, py, shape, size, mat, pad0, pad1, pad2; };
struct SeStepParams {
    const unsigned* in;
    unsigned* out;
    int W, Hl, gy0, Hg;
    int frame;
    int n_mods;
    const SeMod* mods;
};
struct SeTileParams {
    unsigned* buf0;
    unsigned* buf1;
    int W, Hl, gy0, Hg;
    int frame0, nblk, tsteps, nsub_last;
    unsigned seq_base;
    unsigned* done;
    int HY, HX, PH;
    int tiles_x, tiles_y;
    int lut_words, pool_offset, tile_offset;
    const unsigned* lut;
};
struct SeShadeParams {
    const unsigned* cells;
    float4* rgba_f32;
    unsigned* rgba8;
    int W, rows, y0;
};
struct SeLutStepParams {
    unsigned* cells;
    int W, Hl, gy0, Hg;
    int frame;
    int lut_words, pool_offset;
    const unsigned* lut;
};
struct SeLutCensusParams {
    unsigned long long* census;
    const unsigned* popbits;
    int pop_words, pop_offset;
    int own_y0, own_y1;
};
struct SeLightParams {
    const unsigned* old_cells;
    const unsigned* new_cells;
    const float4* light_in;
    float4* light_out;
    int W, Hl, gy0, Hg;
};
struct SeFusedParams {
    SeLightParams lp;
    int frame;
    int n_mods;
    const SeMod* mods;
    int lut_words, pool_offset;
    int tile_offset;
    const unsigned* lut;
    int tiles_x, tiles_y;
};

}  // namespace

struct se_rules {
    se::CompiledRules cr;
    std::string glsl_materials, glsl_rules;
    std::vector<char> cubin;
    std::string nvrtc_log;
    bool compiled = false;
    int tile_threads = 1024;   // CTA size of the tile kernel (compile-time launch bound; tunable: env SE_TILE_THREADS)
    bool experimental_kernels = false;   // compiled with env SE_EXPERIMENTAL_KERNELS=1
    int light_rows = 4;        // rows per thread of se_light => tile height 8 * light_rows (tunable: env SE_LT_ROWS)
};

struct Neighbour {
    bool attached = false;
    bool ipc = false;
    unsigned* cells[2] = {nullptr, nullptr};
    float4* light[2] = {nullptr, nullptr};   // lit strips only
    bool light_ipc = false;
    unsigned* flags = nullptr;   // the neighbour's flag words (inside its cells[0] allocation)
    uint64_t local_rows = 0, ghost_top = 0, ghost_bottom = 0;
};

struct se_sim {
    const se_rules* rules = nullptr;
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    CUmodule mod = nullptr;
    CUfunction f_inplace = nullptr, f_inplace_mods = nullptr, f_pingpong = nullptr, f_pingpong_mods = nullptr, f_light = nullptr, f_fill = nullptr, f_shade = nullptr;
    int W = 0, Hg = 0;
    int row_begin = 0, row_end = 0, ghost_top = 0, ghost_bottom = 0;
    int gy0 = 0, Hl = 0;   // local buffer: rows [gy0, gy0 + Hl) of the global grid
    bool lighting = false;
    unsigned* cells[2] = {nullptr, nullptr};
    int cur = 0;
    float4* light[2] = {nullptr, nullptr};
    int lcur = 0;
    SeMod* d_mods = nullptr;
    // pinned staging ring for the modification UBO: SE_MOD_SLOTS x SE_MAX_MODIFICATIONS records.  A slot is
    // reused only after the async copy that read it has completed (event per slot), so se_sim_step never
    // has to synchronise the stream.
    SeMod* h_mods = nullptr;
    cudaEvent_t mod_events[16] = {};
    unsigned mod_slot = 0;
    std::vector<se_modification> pending;
    int frame = 0;
    uint64_t launches = 0;
    unsigned long long* d_census = nullptr;   // 4 slots of 256 bins (slot 0 is also used by the synchronous census)
    // asynchronous census: side stream + events; census_done[b] guards cell buffer b against being overwritten
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_main = nullptr, census_done[2] = {nullptr, nullptr};
    bool census_pending[2] = {false, false};
    unsigned census_slot = 0;
    // transition-table tile kernel (K1b)
    bool tiled = false;
    CUfunction f_tiles = nullptr, f_build_lut = nullptr, f_lut_global = nullptr;
    unsigned* d_lut = nullptr;
    unsigned* d_tile_done = nullptr;   // per-tile sequence numbers (dataflow between the T-blocks of one launch)
    unsigned tile_seq = 0;             // T-blocks completed so far
    // se_step_tiles spins on flags written by other CTAs of the same grid, so all of its CTAs must be resident
    // together: it is launched cooperatively (the driver then places the whole grid at once, even when another
    // stream's kernels share the device).  False only on a device / context that reports no support.
    bool coop = false;
    int T = 0, HY = 0, HX = 0, PH = 0, tiles_x = 0, tiles_y = 0, lut_words = 0, pool_offset = 0, tile_offset = 0, tile_smem = 0, tile_grid = 0, tile_grid_max = 0, k1c_grid = 0;
    // fused step + modifications + lighting (SE_FLAG_FUSED_LIGHT_EXPERIMENTAL)
    bool fused_light = false;
    CUfunction f_light_fused = nullptr;
    int lf_smem = 0, lf_grid = 0, lf_tiles_x = 0, lf_tiles_y = 0;
    // running census (SE_FLAG_RUNNING_CENSUS, experimental): d_running is valid only between K1c-census steps
    bool running = false, running_valid = false, running_copy_pending = false;
    CUfunction f_lut_global_census = nullptr, f_build_popbits = nullptr;
    unsigned* d_popbits = nullptr;
    unsigned long long* d_running = nullptr;
    int pop_words = 0, pop_offset = 0;
    cudaEvent_t running_copy_done = nullptr;
    Neighbour nb[2];
    // Device-side exchange protocol: 4 flag words live right behind cells[0] (same allocation, so that one
    // IPC handle maps both): [0]/[1] = "done computing" epoch of the strip above/below, [2]/[3] = "ghost rows
    // delivered" epoch from above/below.  Written by the neighbours, waited on by this sim's stream.
    unsigned* flags = nullptr;
    unsigned epoch = 0;
    static size_t flags_offset(size_t W_, size_t Hl_) { return ((W_ * Hl_ * sizeof(unsigned)) + 255) / 256 * 256; }
    size_t cells_bytes() const { return (size_t)W * Hl * sizeof(unsigned); }
    size_t owned_offset() const { return (size_t)ghost_top * W; }
    size_t owned_cells() const { return (size_t)W * (row_end - row_begin); }
};

namespace {

int compile_front(const char* yaml, size_t len, se_rules** out, bool with_nvrtc) {
    if (!yaml || !out) return fail(SE_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    se_rules* r = new se_rules();
    try {
        se::ParsingResult pr = se::parse_string(std::string(yaml, len));
        r->glsl_materials = se::emit_glsl_materials(pr);
        r->glsl_rules = se::emit_glsl_rules(pr);
        if (with_nvrtc) r->cr = se::compile_rules(pr);   // typed expression check + CUDA C emission
        else r->cr.parsed = pr;                          // parse_string only, like the reference crate
    } catch (const se::ParseError& e) {
        int code = map_kind(e.kind);
        std::string msg = e.what();
        delete r;
        return fail(code, msg);
    } catch (const std::exception& e) {
        std::string msg = e.what();
        delete r;
        return fail(SE_ERR_INVALID_ARG, msg);
    }
    if (with_nvrtc) {
        nvrtcProgram prog;
        const char* hdr_src[1] = {r->cr.cuda_header.c_str()};
        const char* hdr_name[1] = {"rules_gen.cuh"};
        if (nvrtcCreateProgram(&prog, KERNEL_SOURCE, "sand_kernels.cu", 1, hdr_src, hdr_name) != NVRTC_SUCCESS) {
            delete r;
            return fail(SE_ERR_COMPILE, "nvrtcCreateProgram failed");
        }
        // -fmad=false: the lighting sums and any float arithmetic in rule conditions are evaluated as
        // written (no FMA contraction), matching the reference expression tree.
        if (const char* tt = std::getenv("SE_TILE_THREADS")) {
            int v = std::atoi(tt);
            if (v == 256 || v == 512 || v == 768 || v == 1024) r->tile_threads = v;
        }
        const std::string def_threads = "-DSE_TILE_THREADS=" + std::to_string(r->tile_threads);
        std::vector<std::string> extra;                       // experiments only: SE_NVRTC_DEFS="-DX=1 -DY=2"
        if (const char* ek = std::getenv("SE_EXPERIMENTAL_KERNELS")) {   // the kernels behind the experimental SE_FLAG_*s
            if (std::string(ek) == "1") { extra.push_back("-DSE_EXPERIMENTAL_KERNELS=1"); r->experimental_kernels = true; }
        }
        if (const char* lr = std::getenv("SE_LT_ROWS")) {     // experiments only: se_light tile height
            const int v = std::atoi(lr);
            if (v == 2 || v == 4 || v == 8) { r->light_rows = v; extra.push_back("-DSE_LT_ROWS=" + std::to_string(v)); }
        }
        if (const char* mc = std::getenv("SE_LT_MINCTAS")) {
            const int v = std::atoi(mc);
            if (v >= 1 && v <= 8) extra.push_back("-DSE_LT_MINCTAS=" + std::to_string(v));
        }
        if (const char* defs = std::getenv("SE_NVRTC_DEFS")) {
            std::string d(defs), tok;
            for (size_t i = 0; i <= d.size(); ++i) {
                if (i == d.size() || d[i] == ' ') { if (!tok.empty()) extra.push_back(tok); tok.clear(); }
                else tok.push_back(d[i]);
            }
        }
        std::vector<const char*> opts = {"-arch=sm_100a", "-std=c++17", "-lineinfo", "-fmad=false", def_threads.c_str()};
        for (auto& e : extra) opts.push_back(e.c_str());
        nvrtcResult res = nvrtcCompileProgram(prog, (int)opts.size(), opts.data());
        size_t log_size = 0;
        nvrtcGetProgramLogSize(prog, &log_size);
        r->nvrtc_log.resize(log_size ? log_size - 1 : 0);
        if (log_size > 1) {
            std::vector<char> log(log_size);
            nvrtcGetProgramLog(prog, log.data());
            r->nvrtc_log.assign(log.data());
        }
        if (res != NVRTC_SUCCESS) {
            std::string msg = "NVRTC: " + std::string(nvrtcGetErrorString(res)) + "\n" + r->nvrtc_log;
            nvrtcDestroyProgram(&prog);
            delete r;
            return fail(SE_ERR_COMPILE, msg);
        }
        size_t cubin_size = 0;
        if (nvrtcGetCUBINSize(prog, &cubin_size) != NVRTC_SUCCESS || cubin_size == 0) {
            nvrtcDestroyProgram(&prog);
            delete r;
            return fail(SE_ERR_COMPILE, "NVRTC produced no cubin");
        }
        r->cubin.resize(cubin_size);
        nvrtcGetCUBIN(prog, r->cubin.data());
        nvrtcDestroyProgram(&prog);
        r->compiled = true;
    }
    *out = r;
    return SE_OK;
}

// A cell buffer that an asynchronous census is still reading must not be overwritten: make the main stream
// wait (on the device) for that census before the next writer of the buffer.
int guard_buffer_write(se_sim* s, int buf) {
    if (s->census_pending[buf]) {
        SE_CUDA(cudaStreamWaitEvent(s->stream, s->census_done[buf], 0));
        s->census_pending[buf] = false;
    }
    return SE_OK;
}

int launch(se_sim* s, CUfunction f, dim3 grid, dim3 block, void** args, unsigned smem = 0, bool cooperative = false) {
    if (cooperative) {
        CUresult r = driver().LaunchCooperativeKernel(f, grid.x, grid.y, grid.z, block.x, block.y, block.z, smem, (CUstream)s->stream, args);
        if (r != CUDA_SUCCESS) {
            // e.g. CUDA_ERROR_COOPERATIVE_LAUNCH_TOO_LARGE when the context can hold fewer CTAs than the occupancy
            // query promised (an MPS partition), or a driver without support: the grid is still <= occupancy x SMs,
            // so the same grid is launched normally from now on; a real launch error then surfaces from that call.
            s->coop = false;
            cooperative = false;
        }
    }
    if (!cooperative)
        SE_CU(driver().LaunchKernel(f, grid.x, grid.y, grid.z, block.x, block.y, block.z, smem, (CUstream)s->stream, args, nullptr));
    s->launches++;
    return SE_OK;
}

int one_step(se_sim* s, bool use_mods, int n_mods) {
    { int rc = guard_buffer_write(s, 0); if (rc) return rc; rc = guard_buffer_write(s, 1); if (rc) return rc; }
    s->frame += 1;
    const int frame = s->frame;
    const int ox = ((frame & 3) == 1 || (frame & 3) == 3) ? 1 : 0;
    const int oy = ((frame & 3) == 1 || (frame & 3) == 2) ? 1 : 0;
    if (frame == 1) {
        // falling_sand.glsl:743-746: every cell becomes EMPTY; modifications are ignored this frame.
        size_t n = (size_t)s->W * s->Hl;
        unsigned zero = 0;
        if (!s->lighting) {
            unsigned* buf = s->cells[s->cur];
            void* args[] = {&buf, &n, &zero};
            return launch(s, s->f_fill, dim3(148 * 8), dim3(256), args);
        }
        unsigned* outb = s->cells[s->cur ^ 1];
        void* args[] = {&outb, &n, &zero};
        int rc = launch(s, s->f_fill, dim3(148 * 8), dim3(256), args);
        if (rc) return rc;
        SeLightParams lp{s->cells[s->cur], outb, s->light[s->lcur], s->light[s->lcur ^ 1], s->W, s->Hl, s->gy0, s->Hg};
        void* largs[] = {&lp};
        rc = launch(s, s->f_light, dim3((s->W + 31) / 32, (s->Hl + 8 * s->rules->light_rows - 1) / (8 * s->rules->light_rows)), dim3(256), largs);
        if (rc) return rc;
        s->cur ^= 1;
        s->lcur ^= 1;
        return SE_OK;
    }
    const int jb0 = (s->gy0 + oy) >> 1;
    const int y_end = std::min(s->Hg, s->gy0 + s->Hl);
    const int nby = ((y_end + oy + 1) >> 1) - jb0;
    const int nbx = (s->W + ox + 1) >> 1;
    dim3 block(64, 4), grid((nbx + 63) / 64, (nby + 3) / 4);
    SeStepParams p;
    p.W = s->W; p.Hl = s->Hl; p.gy0 = s->gy0; p.Hg = s->Hg; p.frame = frame;
    p.n_mods = use_mods ? n_mods : 0;
    p.mods = s->d_mods;
    void* args[] = {&p};
    if (!s->lighting) {
        p.in = s->cells[s->cur];
        p.out = s->cells[s->cur];
        return launch(s, p.n_mods ? s->f_inplace_mods : s->f_inplace, grid, block, args);
    }
    if (s->fused_light) {
        SeFusedParams fp;
        fp.lp = SeLightParams{s->cells[s->cur], s->cells[s->cur ^ 1], s->light[s->lcur], s->light[s->lcur ^ 1], s->W, s->Hl, s->gy0, s->Hg};
        fp.frame = frame; fp.n_mods = p.n_mods; fp.mods = s->d_mods;
        fp.lut_words = s->lut_words; fp.pool_offset = s->pool_offset; fp.tile_offset = s->tile_offset; fp.lut = s->d_lut;
        fp.tiles_x = s->lf_tiles_x; fp.tiles_y = s->lf_tiles_y;
        void* fargs[] = {&fp};
        int rcf = launch(s, s->f_light_fused, dim3(s->lf_grid), dim3(256), fargs, (unsigned)s->lf_smem);
        if (rcf) return rcf;
        s->cur ^= 1;
        s->lcur ^= 1;
        return SE_OK;
    }
    p.in = s->cells[s->cur];
    p.out = s->cells[s->cur ^ 1];
    int rc = launch(s, p.n_mods ? s->f_pingpong_mods : s->f_pingpong, grid, block, args);
    if (rc) return rc;
    SeLightParams lp{p.in, p.out, s->light[s->lcur], s->light[s->lcur ^ 1], s->W, s->Hl, s->gy0, s->Hg};
    void* largs[] = {&lp};
    rc = launch(s, s->f_light, dim3((s->W + 31) / 32, (s->Hl + 8 * s->rules->light_rows - 1) / (8 * s->rules->light_rows)), dim3(256), largs);
    if (rc) return rc;
    s->cur ^= 1;
    s->lcur ^= 1;
    return SE_OK;
}

}  // namespace

extern "C" {

const char* se_last_error(void) { return g_err.c_str(); }
const char* se_version(void) { return "sandengine_b200 0.1 (sm_100a)"; }

int se_rules_compile_yaml(const char* yaml, size_t len, se_rules** out) try {
    return compile_front(yaml, len, out, true);
} SE_ABI_CATCH("se_rules_compile_yaml")
int se_rules_parse_only(const char* yaml, size_t len, se_rules** out) try {
    return compile_front(yaml, len, out, false);
} SE_ABI_CATCH("se_rules_parse_only")

int se_rules_destroy(se_rules* r) try {
    delete r;
    return SE_OK;
} SE_ABI_CATCH("se_rules_destroy")

int se_rules_text(const se_rules* r, int which, const char** text, size_t* len) try {
    if (!r || !text || !len) return fail(SE_ERR_INVALID_ARG, "null argument");
    const std::string* s = nullptr;
    switch (which) {
        case 0: s = &r->glsl_materials; break;
        case 1: s = &r->glsl_rules; break;
        case 2: s = &r->cr.cuda_header; break;
        case 3: s = &r->nvrtc_log; break;
        default: return fail(SE_ERR_INVALID_ARG, "which must be 0..3");
    }
    *text = s->c_str();
    *len = s->size();
    return SE_OK;
} SE_ABI_CATCH("se_rules_text")

int se_rules_cubin(const se_rules* r, const void** data, size_t* len) try {
    if (!r || !data || !len) return fail(SE_ERR_INVALID_ARG, "null argument");
    if (!r->compiled) return fail(SE_ERR_INVALID_ARG, "rules were parsed without NVRTC compilation");
    *data = r->cubin.data();
    *len = r->cubin.size();
    return SE_OK;
} SE_ABI_CATCH("se_rules_cubin")

int se_rules_counts(const se_rules* r, int32_t* n_rules, int32_t* n_types, int32_t* n_materials) try {
    if (!r) return fail(SE_ERR_INVALID_ARG, "null argument");
    if (n_rules) *n_rules = (int32_t)r->cr.parsed.rules.size();
    if (n_types) *n_types = (int32_t)r->cr.parsed.types.size();
    if (n_materials) *n_materials = (int32_t)r->cr.parsed.materials.size();
    return SE_OK;
} SE_ABI_CATCH("se_rules_counts")

int se_rules_material(const se_rules* r, int32_t id, const char** name, const char** type_name, float* density, float* color4,
                      float* emission4, int32_t* selectable) try {
    if (!r) return fail(SE_ERR_INVALID_ARG, "null argument");
    if (id < 0 || id >= (int32_t)r->cr.parsed.materials.size()) return fail(SE_ERR_INVALID_ARG, "material id out of range");
    const se::SandMaterial& m = r->cr.parsed.materials[id];
    if (name) *name = m.name.c_str();
    if (type_name) *type_name = m.mattype.c_str();
    if (density) *density = m.density;
    if (color4) std::memcpy(color4, m.color, sizeof m.color);
    if (emission4) std::memcpy(emission4, m.emission, sizeof m.emission);
    if (selectable) *selectable = m.selectable ? 1 : 0;
    return SE_OK;
} SE_ABI_CATCH("se_rules_material")

int se_rules_material_id(const se_rules* r, const char* name, int32_t* id) try {
    if (!r || !name || !id) return fail(SE_ERR_INVALID_ARG, "null argument");
    for (auto& m : r->cr.parsed.materials)
        if (m.name == name) { *id = m.id; return SE_OK; }
    return fail(SE_ERR_NOT_FOUND, std::string("(NotFound) material '") + name + "'");
} SE_ABI_CATCH("se_rules_material_id")

int se_rules_rule(const se_rules* r, int32_t index, const char** name, int32_t* used, int32_t* kind, const char** precondition) try {
    if (!r) return fail(SE_ERR_INVALID_ARG, "null argument");
    if (index < 0 || index >= (int32_t)r->cr.parsed.rules.size()) return fail(SE_ERR_INVALID_ARG, "rule index out of range");
    const se::SandRule& ru = r->cr.parsed.rules[index];
    if (name) *name = ru.name.c_str();
    if (used) *used = ru.used ? 1 : 0;
    if (kind) *kind = ru.effective_type() == se::SandRuleType::Mirrored ? 0 : ru.effective_type() == se::SandRuleType::Left ? 1 : 2;
    if (precondition) *precondition = ru.has_precondition ? ru.precondition.c_str() : nullptr;
    return SE_OK;
} SE_ABI_CATCH("se_rules_rule")

// ---------------------------------------------------------------------------------------------
int se_sim_destroy(se_sim* s) try {
    if (!s) return SE_OK;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    for (int w = 0; w < 2; ++w) {
        if (s->nb[w].ipc)
            for (int b = 0; b < 2; ++b)
                if (s->nb[w].cells[b]) cudaIpcCloseMemHandle(s->nb[w].cells[b]);
        if (s->nb[w].light_ipc)
            for (int b = 0; b < 2; ++b)
                if (s->nb[w].light[b]) cudaIpcCloseMemHandle(s->nb[w].light[b]);
    }
    for (int b = 0; b < 2; ++b) {
        if (s->cells[b]) cudaFree(s->cells[b]);
        if (s->light[b]) cudaFree(s->light[b]);
    }
    if (s->d_mods) cudaFree(s->d_mods);
    if (s->h_mods) cudaFreeHost(s->h_mods);
    for (auto& ev : s->mod_events) if (ev) cudaEventDestroy(ev);
    if (s->d_census) cudaFree(s->d_census);
    if (s->aux_stream) { cudaStreamSynchronize(s->aux_stream); cudaStreamDestroy(s->aux_stream); }
    if (s->ev_main) cudaEventDestroy(s->ev_main);
    for (auto& ev : s->census_done) if (ev) cudaEventDestroy(ev);
    if (s->d_lut) cudaFree(s->d_lut);
    if (s->d_tile_done) cudaFree(s->d_tile_done);
    if (s->d_popbits) cudaFree(s->d_popbits);
    if (s->d_running) cudaFree(s->d_running);
    if (s->running_copy_done) cudaEventDestroy(s->running_copy_done);
    if (s->mod && driver().ok) driver().ModuleUnload(s->mod);
    if (s->own_stream) cudaStreamDestroy(s->own_stream);
    delete s;
    return SE_OK;
} SE_ABI_CATCH("se_sim_destroy")

int se_sim_create(const se_rules* rules, const se_create_params* prm, se_sim** out) try {
    if (!rules || !prm || !out) return fail(SE_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    if (!rules->compiled) return fail(SE_ERR_INVALID_ARG, "rules were not compiled (use se_rules_compile_yaml)");
    if (prm->width == 0 || prm->height == 0 || prm->width > (1u << 30) || prm->height > (1u << 30))
        return fail(SE_ERR_INVALID_ARG, "bad grid size");
    uint32_t rb = prm->row_begin, re = prm->row_end ? prm->row_end : prm->height;
    if (rb >= re || re > prm->height) return fail(SE_ERR_INVALID_ARG, "bad row range");
    if ((rb & 1u) || ((re & 1u) && re != prm->height)) return fail(SE_ERR_INVALID_ARG, "strip boundaries must be even rows");
    // Strips exchange ghost rows of the id buffer only; the light field of a strip would need its own ghost rows
    // (one row per step, operations.glsl:114-160).  Not implemented: refuse instead of relaxing stale light.
    if ((prm->flags & SE_FLAG_LIGHTING) && (rb > 0 || re < prm->height) && !(prm->flags & SE_FLAG_LIT_STRIP_EXPERIMENTAL))
        return fail(SE_ERR_UNSUPPORTED, "(Unsupported) lighting on a strip (row_begin/row_end) is not implemented: run lighting on one device");
    // per-step kernels index block rows with gridDim.y (<= 65535 CTAs of 4 block rows / 8 light rows)
    if ((uint64_t)(re - rb) + 2ull * prm->halo_rows > 65535ull * 8ull)
        return fail(SE_ERR_INVALID_ARG, "more than 524280 rows per device are not supported (shard the grid into strips)");
    if ((prm->flags & (SE_FLAG_RUNNING_CENSUS | SE_FLAG_FUSED_LIGHT_EXPERIMENTAL)) && !rules->experimental_kernels)
        return fail(SE_ERR_INVALID_ARG, "experimental flag: compile the rules with env SE_EXPERIMENTAL_KERNELS=1 (the kernels are not built by default)");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        (void)cudaGetLastError();
        return fail(SE_ERR_CUDA, std::string("no CUDA device available (there is no CPU fallback): ") + cudaGetErrorString(e));
    }
    if (prm->device < 0 || prm->device >= ndev) return fail(SE_ERR_INVALID_ARG, "bad device ordinal");
    SE_CUDA(cudaSetDevice(prm->device));
    SE_CUDA(cudaFree(nullptr));   // make sure the primary context exists
    if (!driver().ok) return fail(SE_ERR_CUDA, driver().why);

    se_sim* s = new se_sim();
    s->rules = rules;
    s->device = prm->device;
    s->W = (int)prm->width;
    s->Hg = (int)prm->height;
    s->row_begin = (int)rb;
    s->row_end = (int)re;
    s->ghost_top = rb > 0 ? (int)prm->halo_rows : 0;
    s->ghost_bottom = re < prm->height ? (int)prm->halo_rows : 0;
    s->ghost_top = std::min(s->ghost_top, s->row_begin);
    s->ghost_bottom = std::min(s->ghost_bottom, s->Hg - s->row_end);
    s->gy0 = s->row_begin - s->ghost_top;
    s->Hl = (s->row_end - s->row_begin) + s->ghost_top + s->ghost_bottom;
    s->lighting = (prm->flags & SE_FLAG_LIGHTING) != 0;

    auto bail = [&](int code) { se_sim_destroy(s); return code; };
#define SE_TRY(expr) do { int rc_ = (expr); if (rc_ != SE_OK) return bail(rc_); } while (0)
#define SE_CUDA_S(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fail(SE_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); return bail(SE_ERR_CUDA); } } while (0)
#define SE_CU_S(call) do { CUresult r_ = (call); if (r_ != CUDA_SUCCESS) { fail(SE_ERR_CUDA, std::string(#call) + ": " + cu_err(r_)); return bail(SE_ERR_CUDA); } } while (0)

    SE_CUDA_S(cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking));
    s->stream = s->own_stream;
    SE_CU_S(driver().ModuleLoadData(&s->mod, rules->cubin.data()));
    SE_CU_S(driver().ModuleGetFunction(&s->f_inplace, s->mod, "se_step_inplace"));
    SE_CU_S(driver().ModuleGetFunction(&s->f_inplace_mods, s->mod, "se_step_inplace_mods"));
    SE_CU_S(driver().ModuleGetFunction(&s->f_pingpong, s->mod, "se_step_pingpong"));
    SE_CU_S(driver().ModuleGetFunction(&s->f_pingpong_mods, s->mod, "se_step_pingpong_mods"));
    SE_CU_S(driver().ModuleGetFunction(&s->f_light, s->mod, "se_light"));
    SE_CU_S(driver().ModuleGetFunction(&s->f_fill, s->mod, "se_fill_cells"));
    SE_CU_S(driver().ModuleGetFunction(&s->f_shade, s->mod, "se_shade"));

    // Simulation::new allocates zero-filled textures (simulation.rs:145,177-181)
    const bool two = s->lighting;
    const size_t flags_off = se_sim::flags_offset((size_t)s->W, (size_t)s->Hl);
    SE_CUDA_S(cudaMalloc(&s->cells[0], flags_off + 256));
    SE_CUDA_S(cudaMemsetAsync(s->cells[0], 0, flags_off + 256, s->stream));
    s->flags = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(s->cells[0]) + flags_off);
    if (two) {
        SE_CUDA_S(cudaMalloc(&s->cells[1], s->cells_bytes()));
        SE_CUDA_S(cudaMemsetAsync(s->cells[1], 0, s->cells_bytes(), s->stream));
        for (int b = 0; b < 2; ++b) {
            SE_CUDA_S(cudaMalloc(&s->light[b], (size_t)s->W * s->Hl * sizeof(float4)));
            SE_CUDA_S(cudaMemsetAsync(s->light[b], 0, (size_t)s->W * s->Hl * sizeof(float4), s->stream));
        }
    }
    SE_CUDA_S(cudaMalloc(&s->d_mods, SE_MAX_MODIFICATIONS * sizeof(SeMod)));
    SE_CUDA_S(cudaMallocHost(&s->h_mods, 16 * SE_MAX_MODIFICATIONS * sizeof(SeMod)));
    for (auto& ev : s->mod_events) SE_CUDA_S(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    SE_CUDA_S(cudaMalloc(&s->d_census, 4 * 256 * sizeof(unsigned long long)));
    SE_CUDA_S(cudaStreamCreateWithFlags(&s->aux_stream, cudaStreamNonBlocking));
    SE_CUDA_S(cudaEventCreateWithFlags(&s->ev_main, cudaEventDisableTiming));
    for (auto& ev : s->census_done) SE_CUDA_S(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));

    // ---- K1b: transition table + shared-memory tiles with temporal blocking -------------------------
    // Used for runs of steps without modifications when lighting is off, the rule set is table-eligible
    // (codegen.h) and rows are 16-byte aligned.  temporal_block == 1 forces the per-step kernel K1a.
    if (rules->cr.lut_eligible && !s->lighting && (s->W % 4) == 0 && prm->temporal_block != 1) {
        const int N = rules->cr.tables.n_materials, NCLS = (int)rules->cr.lut_thresholds.size() + 1;
        const int N4 = N * N * N * N;
        const int NE = N4 * rules->cr.lut_tables;                 // table entries (two tables for Left/Right rule sets, experimental)
        const int POOL_MAX = 4095;
        (void)NCLS;
        const size_t pool_off = ((size_t)NE * 2 + 7) / 8 * 8;     // pool entries are 8 bytes {thr, A, B}
        const size_t lut_cap = pool_off + (size_t)POOL_MAX * 8 + 16;
        unsigned* d_counter = nullptr;
        SE_CU_S(driver().ModuleGetFunction(&s->f_tiles, s->mod, "se_step_tiles"));
        SE_CU_S(driver().ModuleGetFunction(&s->f_build_lut, s->mod, "se_build_lut"));
        SE_CU_S(driver().ModuleGetFunction(&s->f_lut_global, s->mod, "se_step_lut_global"));
        SE_CUDA_S(cudaMalloc(&s->d_lut, lut_cap));
        SE_CUDA_S(cudaMalloc(&d_counter, sizeof(unsigned)));
        SE_CUDA_S(cudaMemsetAsync(s->d_lut, 0, lut_cap, s->stream));
        SE_CUDA_S(cudaMemsetAsync(d_counter, 0, sizeof(unsigned), s->stream));
        unsigned short* base = reinterpret_cast<unsigned short*>(s->d_lut);
        void* pool = reinterpret_cast<char*>(s->d_lut) + pool_off;
        void* bargs[] = {&base, &pool, &d_counter};
        SE_TRY(launch(s, s->f_build_lut, dim3((NE + 255) / 256), dim3(256), bargs));
        unsigned n_pool = 0;
        SE_CUDA_S(cudaMemcpyAsync(&n_pool, d_counter, sizeof(unsigned), cudaMemcpyDeviceToHost, s->stream));
        SE_CUDA_S(cudaStreamSynchronize(s->stream));
        cudaFree(d_counter);
        int smem_sm = 0, smem_optin = 0;
        SE_CUDA_S(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, s->device));
        SE_CUDA_S(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, s->device));
        if ((int)n_pool <= POOL_MAX) {
            const size_t lut_bytes = pool_off + (size_t)n_pool * 8;
            s->pool_offset = (int)pool_off;
            s->lut_words = (int)((lut_bytes + 3) / 4);
            s->tile_offset = (int)((lut_bytes + 15) / 16 * 16);
            // two CTAs per SM: each gets half of the SM's shared memory minus the per-CTA reservation (1 KB)
            // and the kernel's static shared memory (1 KB fat table)
            const int budget = std::min(smem_optin, smem_sm / 2 - 2048 - 256);
            int T = prm->temporal_block ? (int)prm->temporal_block : 8;
            T = std::max(2, T + (T & 1));
            int PH_max = ((budget - s->tile_offset) / 256) & ~1;
            PH_max = std::min(PH_max, 256);   // measured: taller tiles (up to the 276 rows that fit) are slower (less load/compute overlap)
            // rows: the Margolus row offset changes every other frame, so T fused steps need only T/2+1 halo rows
            const int HY = ((T / 2 + 1) + 1) & ~1;
            if (PH_max >= 4 * T + 16) {
                int n_sm = 0;
                SE_CUDA_S(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, s->device));
                const int grid_max = 2 * n_sm;                       // persistent: 2 CTAs per SM
                s->T = T;
                s->HY = HY;
                s->HX = (T + 3) & ~3;
                s->tiles_x = (s->W + (256 - 2 * s->HX) - 1) / (256 - 2 * s->HX);
                // Tile height: every CTA processes ceil(tiles / grid) tiles of PH rows, so the launch costs
                // rounds * (PH + c).  Pick the PH that minimises it (avoids a nearly empty last round: at
                // 16384 x 2176 rows per GPU a fixed PH = 256 would spend 3 rounds on 2.3 rounds of work).
                long best_cost = -1;
                int best_PH = PH_max;
                for (int PH = PH_max; PH >= 4 * T + 16; PH -= 2) {
                    const int ty = (s->Hl + (PH - 2 * HY) - 1) / (PH - 2 * HY);
                    const long tiles = (long)s->tiles_x * ty;
                    // launches usually carry several T-blocks whose tiles are dealt to the CTAs as one sequence
                    // (dataflow inside the kernel), so the round count is taken over a typical 8-T-block launch
                    const long rounds = (8 * tiles + grid_max - 1) / grid_max;
                    // useful rows per tile shrink with PH: account for the halo rows recomputed by every tile
                    // + 24: per-tile fixed work (table/phase set-up, barriers, exposed load/store) expressed in rows;
                    // fitted on B200 (16384 x 2116 rows: 3 rounds of PH 190 beat 4 rounds of PH 138)
                    const long cost = rounds * (PH + 24);
                    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_PH = PH; }
                }
                int PH = best_PH;
                if (const char* ov = std::getenv("SE_TILE_PH")) {      // experiments only: override the search
                    const int v = std::atoi(ov) & ~1;
                    if (v >= 4 * T + 16 && v <= PH_max) PH = v;
                }
                s->PH = PH;
                s->tile_smem = s->tile_offset + 256 * PH;
                s->tiles_y = (s->Hl + (PH - 2 * HY) - 1) / (PH - 2 * HY);
                s->tile_grid = std::min(grid_max, s->tiles_x * s->tiles_y);
                s->tile_grid_max = grid_max;
                s->k1c_grid = grid_max;
                if (const char* kg = std::getenv("SE_K1C_GRID")) s->k1c_grid = std::max(1, std::atoi(kg));   // experiments only
                SE_CU_S(driver().FuncSetAttribute(s->f_tiles, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, s->tile_smem));
                SE_CU_S(driver().FuncSetAttribute(s->f_lut_global, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, s->tile_offset));
                // the dataflow between T-blocks needs every CTA of the grid to be resident: size the grid from the
                // occupancy the driver reports for this kernel / block size / shared-memory footprint
                int occ = 0;
                SE_CU_S(driver().OccupancyMaxActiveBlocks(&occ, s->f_tiles, rules->tile_threads, (size_t)s->tile_smem));
                if (occ < 1) { fail(SE_ERR_CUDA, "se_step_tiles does not fit on an SM"); return bail(SE_ERR_CUDA); }
                s->tile_grid_max = std::min(grid_max, occ * n_sm);
                int coop_attr = 0;
                SE_CUDA_S(cudaDeviceGetAttribute(&coop_attr, cudaDevAttrCooperativeLaunch, s->device));
                s->coop = coop_attr != 0 && !std::getenv("SE_NO_COOP_LAUNCH");   // env: experiments only
                SE_CUDA_S(cudaMalloc(&s->cells[1], s->cells_bytes()));
                SE_CUDA_S(cudaMemsetAsync(s->cells[1], 0, s->cells_bytes(), s->stream));
                SE_CUDA_S(cudaMalloc(&s->d_tile_done, (size_t)s->tiles_x * s->tiles_y * sizeof(unsigned)));
                SE_CUDA_S(cudaMemsetAsync(s->d_tile_done, 0, (size_t)s->tiles_x * s->tiles_y * sizeof(unsigned), s->stream));
                s->tiled = true;
                if ((prm->flags & SE_FLAG_RUNNING_CENSUS) && rules->cr.lut_tables == 1) {
                    SE_CU_S(driver().ModuleGetFunction(&s->f_lut_global_census, s->mod, "se_step_lut_global_census"));
                    SE_CU_S(driver().ModuleGetFunction(&s->f_build_popbits, s->mod, "se_build_popbits"));
                    s->pop_words = (N4 + 31) / 32;
                    s->pop_offset = s->tile_offset;                       // behind the staged table (16-aligned)
                    SE_CUDA_S(cudaMalloc(&s->d_popbits, (size_t)s->pop_words * sizeof(unsigned)));
                    SE_CUDA_S(cudaMemsetAsync(s->d_popbits, 0, (size_t)s->pop_words * sizeof(unsigned), s->stream));
                    SE_CUDA_S(cudaMalloc(&s->d_running, 256 * sizeof(unsigned long long)));
                    SE_CUDA_S(cudaEventCreateWithFlags(&s->running_copy_done, cudaEventDisableTiming));
                    void* pargs[] = {&s->d_popbits};
                    SE_TRY(launch(s, s->f_build_popbits, dim3((N4 + 255) / 256), dim3(256), pargs));
                    SE_CU_S(driver().FuncSetAttribute(s->f_lut_global_census, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES,
                                                      s->pop_offset + ((s->pop_words * 4 + 15) & ~15)));
                    s->running = true;
                }
            }
        }
    }
    // ---- K3f (experimental): table step + override + lighting in one kernel --------------------------------
    if ((prm->flags & SE_FLAG_FUSED_LIGHT_EXPERIMENTAL) && s->lighting && rules->cr.lut_eligible) {
        // the table is built exactly as for K1b above (kept separate so that the default path is untouched)
        const int N = rules->cr.tables.n_materials;
        const int N4 = N * N * N * N * rules->cr.lut_tables;
        const int POOL_MAX = 4095;
        const size_t pool_off = ((size_t)N4 * 2 + 7) / 8 * 8;
        const size_t lut_cap = pool_off + (size_t)POOL_MAX * 8 + 16;
        unsigned* d_counter = nullptr;
        CUfunction f_build = nullptr;
        SE_CU_S(driver().ModuleGetFunction(&f_build, s->mod, "se_build_lut"));
        SE_CU_S(driver().ModuleGetFunction(&s->f_light_fused, s->mod, "se_light_fused"));
        SE_CUDA_S(cudaMalloc(&s->d_lut, lut_cap));
        SE_CUDA_S(cudaMalloc(&d_counter, sizeof(unsigned)));
        SE_CUDA_S(cudaMemsetAsync(s->d_lut, 0, lut_cap, s->stream));
        SE_CUDA_S(cudaMemsetAsync(d_counter, 0, sizeof(unsigned), s->stream));
        unsigned short* base = reinterpret_cast<unsigned short*>(s->d_lut);
        void* pool = reinterpret_cast<char*>(s->d_lut) + pool_off;
        void* bargs[] = {&base, &pool, &d_counter};
        SE_TRY(launch(s, f_build, dim3((N4 + 255) / 256), dim3(256), bargs));
        unsigned n_pool = 0;
        SE_CUDA_S(cudaMemcpyAsync(&n_pool, d_counter, sizeof(unsigned), cudaMemcpyDeviceToHost, s->stream));
        SE_CUDA_S(cudaStreamSynchronize(s->stream));
        cudaFree(d_counter);
        if ((int)n_pool <= POOL_MAX) {
            const size_t lut_bytes = pool_off + (size_t)n_pool * 8;
            const int tile_h = 8 * rules->light_rows;
            s->pool_offset = (int)pool_off;
            s->lut_words = (int)((lut_bytes + 3) / 4);
            s->tile_offset = (int)((lut_bytes + 15) / 16 * 16);
            s->lf_smem = s->tile_offset + (tile_h + 2) * 34 * 16 + (tile_h + 2) * 36;     // table + term[] + ids[] (SE_LT_TERMS, SE_LF_IDS_BYTES)
            s->lf_tiles_x = (s->W + 31) / 32;
            s->lf_tiles_y = (s->Hl + tile_h - 1) / tile_h;
            int n_sm = 0, occ = 0, smem_optin = 0;
            SE_CUDA_S(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, s->device));
            SE_CUDA_S(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, s->device));
            if (s->lf_smem <= smem_optin) {
                SE_CU_S(driver().FuncSetAttribute(s->f_light_fused, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, s->lf_smem));
                SE_CU_S(driver().OccupancyMaxActiveBlocks(&occ, s->f_light_fused, 256, (size_t)s->lf_smem));
                if (occ >= 1) {
                    s->lf_grid = (int)std::min<long long>((long long)occ * n_sm, (long long)s->lf_tiles_x * s->lf_tiles_y);
                    s->fused_light = true;
                }
            }
        }
    }
    SE_CUDA_S(cudaStreamSynchronize(s->stream));
    *out = s;
    return SE_OK;
} SE_ABI_CATCH("se_sim_create")

int se_sim_step(se_sim* s, uint32_t n_steps) try {
    if (!s) return fail(SE_ERR_INVALID_ARG, "null sim");
    if (n_steps == 0) return SE_OK;
    SE_CUDA(cudaSetDevice(s->device));
    // simulation.rs:203-208: the first min(len, 256) pending modifications go into the UBO; the shader
    // stops at the first entry with mod_size == 0 (falling_sand.glsl:754-756).
    int n_mods = 0, n_staged = 0;
    if (!s->pending.empty()) {
        // Stage the UBO the way the host of the reference does (all of the first min(len, 256) records are
        // copied, simulation.rs:205-207); the kernels scan only up to the first mod_size == 0 record.
        n_staged = (int)std::min<size_t>(s->pending.size(), SE_MAX_MODIFICATIONS);
        const unsigned slot = s->mod_slot++ % 16u;
        SE_CUDA(cudaEventSynchronize(s->mod_events[slot]));   // returns at once unless 16 steps are still in flight
        SeMod* stage = s->h_mods + (size_t)slot * SE_MAX_MODIFICATIONS;
        bool open = true;
        for (int i = 0; i < n_staged; ++i) {
            const se_modification& m = s->pending[i];
            if (m.mod_size == 0) open = false;
            SeMod d;
            d.px = m.position[0]; d.py = m.position[1]; d.shape = m.mod_shape;
            // Sizes above 2^30 are staged as 2^30: for every position within +-2^29 of the grid the record still covers
            // exactly the cells the shader's test covers (all of them), and the per-CTA culling (px +- size, 32-bit) cannot
            // overflow.  mod_size = INT_MAX ("fill everything") fills the grid in the reference, too.
            d.size = std::min(m.mod_size, 1 << 30);
            // getMaterialFromID: unknown id => MAT_NULL (id 1), gen/materials.glsl:79-86
            d.mat = (m.mod_matID >= 0 && m.mod_matID < s->rules->cr.tables.n_materials) ? m.mod_matID : 1;
            d.pad0 = d.pad1 = d.pad2 = 0;
            stage[i] = d;
            if (open) n_mods = i + 1;
        }
        SE_CUDA(cudaMemcpyAsync(s->d_mods, stage, (size_t)n_staged * sizeof(SeMod), cudaMemcpyHostToDevice, s->stream));
        SE_CUDA(cudaEventRecord(s->mod_events[slot], s->stream));
    }
    for (uint32_t k = 0; k < n_steps;) {
        const bool mods_now = (k == 0 && n_mods > 0);
        if (s->tiled && !mods_now && s->frame + 1 != 1) {
            // a run of plain steps: fuse up to T of them per launch (ping-pong buffers)
            const int nsub = (int)std::min<uint32_t>((uint32_t)s->T, n_steps - k);
            if (nsub == 1) {
                // a lone step (the per-frame path): K1c, table transitions straight from global memory, in place
                { int rc = guard_buffer_write(s, s->cur); if (rc) return rc; }
                SeLutStepParams lp;
                lp.cells = s->cells[s->cur];
                lp.W = s->W; lp.Hl = s->Hl; lp.gy0 = s->gy0; lp.Hg = s->Hg; lp.frame = s->frame + 1;
                lp.lut_words = s->lut_words; lp.pool_offset = s->pool_offset; lp.lut = s->d_lut;
                void* largs[] = {&lp};
                int rc;
                if (s->running && s->running_valid) {
                    SeLutCensusParams cx{s->d_running, s->d_popbits, s->pop_words, s->pop_offset, s->row_begin, s->row_end};
                    void* cargs[] = {&lp, &cx};
                    rc = launch(s, s->f_lut_global_census, dim3(s->k1c_grid), dim3(512), cargs, (unsigned)(s->pop_offset + ((s->pop_words * 4 + 15) & ~15)));
                } else {
                    rc = launch(s, s->f_lut_global, dim3(s->k1c_grid), dim3(512), largs, (unsigned)s->tile_offset);   // SE_K1C_THREADS
                }
                if (rc) return rc;
                s->frame += 1;
                k += 1;
                continue;
            }
            // a run of plain steps: ONE launch of ceil(run / T) T-blocks (dataflow between them inside the kernel)
            s->running_valid = false;
            { int rc = guard_buffer_write(s, 0); if (rc) return rc; rc = guard_buffer_write(s, 1); if (rc) return rc; }
            const uint32_t run = std::min<uint32_t>(n_steps - k, 64u * (uint32_t)s->T);     // bound the kernel duration
            const int nblk = (int)((run + (uint32_t)s->T - 1) / (uint32_t)s->T);
            SeTileParams tp;
            tp.buf0 = s->cells[s->cur]; tp.buf1 = s->cells[s->cur ^ 1];
            tp.W = s->W; tp.Hl = s->Hl; tp.gy0 = s->gy0; tp.Hg = s->Hg;
            tp.frame0 = s->frame + 1; tp.nblk = nblk; tp.tsteps = s->T; tp.nsub_last = (int)(run - (uint32_t)(nblk - 1) * (uint32_t)s->T);
            tp.seq_base = s->tile_seq; tp.done = s->d_tile_done;
            tp.HY = s->HY; tp.HX = s->HX; tp.PH = s->PH;
            tp.tiles_x = s->tiles_x; tp.tiles_y = s->tiles_y; tp.lut_words = s->lut_words; tp.pool_offset = s->pool_offset; tp.tile_offset = s->tile_offset;
            tp.lut = s->d_lut;
            void* targs[] = {&tp};
            const long long items = (long long)nblk * s->tiles_x * s->tiles_y;
            const int grid = (int)std::min<long long>((long long)s->tile_grid_max, items);
            int rc = launch(s, s->f_tiles, dim3(grid), dim3(s->rules->tile_threads), targs, (unsigned)s->tile_smem, s->coop);
            if (rc) return rc;
            s->frame += (int)run;
            s->tile_seq += (unsigned)nblk;
            s->cur ^= (nblk & 1);
            k += run;
            continue;
        }
        s->running_valid = false;
        int rc = one_step(s, k == 0, n_mods);
        if (rc) return rc;
        ++k;
    }
    s->pending.clear();   // simulation.rs:252
    return SE_OK;
} SE_ABI_CATCH("se_sim_step")

int se_sim_push_modifications(se_sim* s, const se_modification* mods, uint32_t n) try {
    if (!s || (!mods && n)) return fail(SE_ERR_INVALID_ARG, "null argument");
    s->pending.insert(s->pending.end(), mods, mods + n);
    return SE_OK;
} SE_ABI_CATCH("se_sim_push_modifications")

int se_sim_set_frame(se_sim* s, int32_t frame) try {
    if (!s) return fail(SE_ERR_INVALID_ARG, "null sim");
    if (frame < 0) return fail(SE_ERR_INVALID_ARG, "frame must be >= 0");
    s->frame = frame;
    return SE_OK;
} SE_ABI_CATCH("se_sim_set_frame")

int se_sim_get_frame(const se_sim* s, int32_t* frame) try {
    if (!s || !frame) return fail(SE_ERR_INVALID_ARG, "null argument");
    *frame = s->frame;
    return SE_OK;
} SE_ABI_CATCH("se_sim_get_frame")

int se_sim_upload_cells(se_sim* s, const uint32_t* host) try {
    if (!s || !host) return fail(SE_ERR_INVALID_ARG, "null argument");
    SE_CUDA(cudaSetDevice(s->device));
    { int rc = guard_buffer_write(s, s->cur); if (rc) return rc; }
    s->running_valid = false;
    SE_CUDA(cudaMemcpyAsync(s->cells[s->cur] + s->owned_offset(), host, s->owned_cells() * sizeof(unsigned), cudaMemcpyHostToDevice, s->stream));
    SE_CUDA(cudaStreamSynchronize(s->stream));
    return SE_OK;
} SE_ABI_CATCH("se_sim_upload_cells")

int se_sim_download_cells(se_sim* s, uint32_t* host) try {
    if (!s || !host) return fail(SE_ERR_INVALID_ARG, "null argument");
    SE_CUDA(cudaSetDevice(s->device));
    SE_CUDA(cudaMemcpyAsync(host, s->cells[s->cur] + s->owned_offset(), s->owned_cells() * sizeof(unsigned), cudaMemcpyDeviceToHost, s->stream));
    SE_CUDA(cudaStreamSynchronize(s->stream));
    return SE_OK;
} SE_ABI_CATCH("se_sim_download_cells")

int se_sim_upload_light(se_sim* s, const float* host) try {
    if (!s || !host) return fail(SE_ERR_INVALID_ARG, "null argument");
    if (!s->lighting) return fail(SE_ERR_INVALID_ARG, "sim was created without SE_FLAG_LIGHTING");
    SE_CUDA(cudaSetDevice(s->device));
    SE_CUDA(cudaMemcpyAsync(s->light[s->lcur] + s->owned_offset(), host, s->owned_cells() * sizeof(float4), cudaMemcpyHostToDevice, s->stream));
    SE_CUDA(cudaStreamSynchronize(s->stream));
    return SE_OK;
} SE_ABI_CATCH("se_sim_upload_light")

int se_sim_download_light(se_sim* s, float* host) try {
    if (!s || !host) return fail(SE_ERR_INVALID_ARG, "null argument");
    if (!s->lighting) return fail(SE_ERR_INVALID_ARG, "sim was created without SE_FLAG_LIGHTING");
    SE_CUDA(cudaSetDevice(s->device));
    SE_CUDA(cudaMemcpyAsync(host, s->light[s->lcur] + s->owned_offset(), s->owned_cells() * sizeof(float4), cudaMemcpyDeviceToHost, s->stream));
    SE_CUDA(cudaStreamSynchronize(s->stream));
    return SE_OK;
} SE_ABI_CATCH("se_sim_download_light")

int se_sim_download_color(se_sim* s, float* host_f32, uint32_t* host_rgba8) try {
    if (!s || (!host_f32 && !host_rgba8)) return fail(SE_ERR_INVALID_ARG, "null argument");
    SE_CUDA(cudaSetDevice(s->device));
    const int rows = s->row_end - s->row_begin;
    const size_t n = (size_t)s->W * rows;
    float4* d_f32 = nullptr;
    unsigned* d_u8 = nullptr;
    if (host_f32) SE_CUDA(cudaMalloc(&d_f32, n * sizeof(float4)));
    if (host_rgba8) { cudaError_t e = cudaMalloc(&d_u8, n * sizeof(unsigned)); if (e != cudaSuccess) { cudaFree(d_f32); return fail(SE_ERR_CUDA, cudaGetErrorString(e)); } }
    SeShadeParams sp{s->cells[s->cur] + s->owned_offset(), d_f32, d_u8, s->W, rows, s->row_begin};
    void* args[] = {&sp};
    int rc = launch(s, s->f_shade, dim3((s->W + 63) / 64, (rows + 3) / 4), dim3(64, 4), args);
    cudaError_t e = cudaSuccess;
    if (rc == SE_OK && host_f32) e = cudaMemcpyAsync(host_f32, d_f32, n * sizeof(float4), cudaMemcpyDeviceToHost, s->stream);
    if (rc == SE_OK && e == cudaSuccess && host_rgba8) e = cudaMemcpyAsync(host_rgba8, d_u8, n * sizeof(unsigned), cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(d_f32);
    cudaFree(d_u8);
    if (rc) return rc;
    if (e != cudaSuccess) return fail(SE_ERR_CUDA, cudaGetErrorString(e));
    return SE_OK;
} SE_ABI_CATCH("se_sim_download_color")

int se_sim_device_cells(se_sim* s, void** dptr, size_t* pitch) try {
    if (!s || !dptr) return fail(SE_ERR_INVALID_ARG, "null argument");
    *dptr = s->cells[s->cur] + s->owned_offset();
    if (pitch) *pitch = (size_t)s->W * sizeof(unsigned);
    return SE_OK;
} SE_ABI_CATCH("se_sim_device_cells")

// running census: (re)count into d_running on the main stream when a step other than K1c-census ran since
static int ensure_running_census(se_sim* s) {
    if (s->running_valid) return SE_OK;
    SE_CUDA(cudaMemsetAsync(s->d_running, 0, 256 * sizeof(unsigned long long), s->stream));
    se_static::launch_census(s->cells[s->cur] + s->owned_offset(), s->owned_cells(), s->d_running, s->stream);
    s->launches++;
    SE_CUDA(cudaGetLastError());
    s->running_valid = true;
    return SE_OK;
}

int se_sim_census(se_sim* s, uint64_t* counts256) try {
    if (!s || !counts256) return fail(SE_ERR_INVALID_ARG, "null argument");
    SE_CUDA(cudaSetDevice(s->device));
    if (s->running) {
        int rc = ensure_running_census(s);
        if (rc) return rc;
        SE_CUDA(cudaMemcpyAsync(counts256, s->d_running, 256 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
        SE_CUDA(cudaStreamSynchronize(s->stream));
        return SE_OK;
    }
    SE_CUDA(cudaMemsetAsync(s->d_census, 0, 256 * sizeof(unsigned long long), s->stream));
    se_static::launch_census(s->cells[s->cur] + s->owned_offset(), s->owned_cells(), s->d_census, s->stream);
    s->launches++;
    SE_CUDA(cudaGetLastError());
    SE_CUDA(cudaMemcpyAsync(counts256, s->d_census, 256 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
    SE_CUDA(cudaStreamSynchronize(s->stream));
    return SE_OK;
} SE_ABI_CATCH("se_sim_census")

int se_sim_census_async(se_sim* s, uint64_t* host_counts256) try {
    if (!s || !host_counts256) return fail(SE_ERR_INVALID_ARG, "null argument");
    SE_CUDA(cudaSetDevice(s->device));
    if (s->running) {
        // the running census lives on the main stream: a 2 KB copy in stream order, no pass over the grid
        int rc = ensure_running_census(s);
        if (rc) return rc;
        SE_CUDA(cudaMemcpyAsync(host_counts256, s->d_running, 256 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
        SE_CUDA(cudaEventRecord(s->running_copy_done, s->stream));
        s->running_copy_pending = true;
        return SE_OK;
    }
    const int buf = s->cur;
    if (s->census_pending[buf]) {
        // the previous census of this buffer is still being tracked by the same event: finish it first
        SE_CUDA(cudaEventSynchronize(s->census_done[buf]));
        s->census_pending[buf] = false;
    }
    unsigned long long* slot = s->d_census + 256 * (size_t)(s->census_slot++ % 4u);
    SE_CUDA(cudaEventRecord(s->ev_main, s->stream));
    SE_CUDA(cudaStreamWaitEvent(s->aux_stream, s->ev_main, 0));
    SE_CUDA(cudaMemsetAsync(slot, 0, 256 * sizeof(unsigned long long), s->aux_stream));
    se_static::launch_census(s->cells[buf] + s->owned_offset(), s->owned_cells(), slot, s->aux_stream);
    s->launches++;
    SE_CUDA(cudaGetLastError());
    SE_CUDA(cudaMemcpyAsync(host_counts256, slot, 256 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->aux_stream));
    SE_CUDA(cudaEventRecord(s->census_done[buf], s->aux_stream));
    s->census_pending[buf] = true;
    return SE_OK;
} SE_ABI_CATCH("se_sim_census_async")

int se_sim_census_wait(se_sim* s) try {
    if (!s) return fail(SE_ERR_INVALID_ARG, "null sim");
    SE_CUDA(cudaSetDevice(s->device));
    SE_CUDA(cudaStreamSynchronize(s->aux_stream));
    if (s->running_copy_pending) {
        SE_CUDA(cudaEventSynchronize(s->running_copy_done));
        s->running_copy_pending = false;
    }
    return SE_OK;
} SE_ABI_CATCH("se_sim_census_wait")

int se_sim_set_stream(se_sim* s, void* stream) try {
    if (!s) return fail(SE_ERR_INVALID_ARG, "null sim");
    SE_CUDA(cudaSetDevice(s->device));
    SE_CUDA(cudaStreamSynchronize(s->stream));
    s->stream = stream ? (cudaStream_t)stream : s->own_stream;
    return SE_OK;
} SE_ABI_CATCH("se_sim_set_stream")

int se_sim_synchronize(se_sim* s) try {
    if (!s) return fail(SE_ERR_INVALID_ARG, "null sim");
    SE_CUDA(cudaSetDevice(s->device));
    SE_CUDA(cudaStreamSynchronize(s->stream));
    return SE_OK;
} SE_ABI_CATCH("se_sim_synchronize")

int se_sim_launch_count(const se_sim* s, uint64_t* n) try {
    if (!s || !n) return fail(SE_ERR_INVALID_ARG, "null argument");
    *n = s->launches;
    return SE_OK;
} SE_ABI_CATCH("se_sim_launch_count")

// ---- strips ----------------------------------------------------------------------------------
int se_sim_ipc_export(se_sim* s, void* handles, uint64_t* local_rows, uint64_t* ghost_top, uint64_t* ghost_bottom) try {
    if (!s || !handles) return fail(SE_ERR_INVALID_ARG, "null argument");
    SE_CUDA(cudaSetDevice(s->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    std::memset(handles, 0, 128);
    for (int b = 0; b < 2; ++b) {
        if (!s->cells[b]) continue;
        cudaIpcMemHandle_t h;
        SE_CUDA(cudaIpcGetMemHandle(&h, s->cells[b]));
        std::memcpy((char*)handles + 64 * b, &h, 64);
    }
    if (local_rows) *local_rows = (uint64_t)s->Hl;
    if (ghost_top) *ghost_top = (uint64_t)s->ghost_top;
    if (ghost_bottom) *ghost_bottom = (uint64_t)s->ghost_bottom;
    return SE_OK;
} SE_ABI_CATCH("se_sim_ipc_export")

int se_sim_ipc_attach(se_sim* s, int which, const void* handles, uint64_t nb_local_rows, uint64_t nb_ghost_top, uint64_t nb_ghost_bottom) try {
    if (!s || !handles || which < 0 || which > 1) return fail(SE_ERR_INVALID_ARG, "bad argument");
    SE_CUDA(cudaSetDevice(s->device));
    const uint64_t owned = (uint64_t)(s->row_end - s->row_begin);
    if ((which == 0 ? nb_ghost_bottom : nb_ghost_top) > owned || nb_ghost_top + nb_ghost_bottom >= nb_local_rows)
        return fail(SE_ERR_INVALID_ARG, "neighbour's ghost rows exceed the rows this strip owns");
    Neighbour& nb = s->nb[which];
    if (nb.ipc)
        for (int b = 0; b < 2; ++b)
            if (nb.cells[b]) { cudaIpcCloseMemHandle(nb.cells[b]); nb.cells[b] = nullptr; }
    nb.attached = false;
    nb.local_rows = nb_local_rows;
    nb.ghost_top = nb_ghost_top;
    nb.ghost_bottom = nb_ghost_bottom;
    nb.ipc = true;
    for (int b = 0; b < 2; ++b) {
        if (!s->cells[b]) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, (const char*)handles + 64 * b, 64);
        void* p = nullptr;
        SE_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        nb.cells[b] = (unsigned*)p;
    }
    nb.flags = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(nb.cells[0]) + se_sim::flags_offset((size_t)s->W, (size_t)nb_local_rows));
    nb.attached = true;
    return SE_OK;
} SE_ABI_CATCH("se_sim_ipc_attach")

int se_sim_attach_local(se_sim* s, int which, se_sim* other) try {
    if (!s || !other || which < 0 || which > 1) return fail(SE_ERR_INVALID_ARG, "bad argument");
    if (s->W != other->W || s->Hg != other->Hg) return fail(SE_ERR_INVALID_ARG, "neighbour belongs to a different grid");
    SE_CUDA(cudaSetDevice(s->device));
    if (other->device != s->device) {
        int can = 0;
        SE_CUDA(cudaDeviceCanAccessPeer(&can, s->device, other->device));
        if (!can) return fail(SE_ERR_CUDA, "devices cannot access each other's memory");
        cudaError_t e = cudaDeviceEnablePeerAccess(other->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(SE_ERR_CUDA, cudaGetErrorString(e));
        (void)cudaGetLastError();
    }
    if ((which == 0 ? other->ghost_bottom : other->ghost_top) > s->row_end - s->row_begin)
        return fail(SE_ERR_INVALID_ARG, "neighbour's ghost rows exceed the rows this strip owns");
    Neighbour& nb = s->nb[which];
    if (nb.ipc)
        for (int b = 0; b < 2; ++b)
            if (nb.cells[b]) { cudaIpcCloseMemHandle(nb.cells[b]); nb.cells[b] = nullptr; }
    nb.local_rows = (uint64_t)other->Hl;
    nb.ghost_top = (uint64_t)other->ghost_top;
    nb.ghost_bottom = (uint64_t)other->ghost_bottom;
    nb.ipc = false;
    nb.cells[0] = other->cells[0];
    nb.cells[1] = other->cells[1];
    if (s->lighting != other->lighting) return fail(SE_ERR_INVALID_ARG, "neighbour differs in SE_FLAG_LIGHTING");
    nb.light[0] = other->light[0];
    nb.light[1] = other->light[1];
    nb.light_ipc = false;
    nb.flags = other->flags;
    nb.attached = true;
    return SE_OK;
} SE_ABI_CATCH("se_sim_attach_local")

int se_sim_halo_push(se_sim* s) try {
    if (!s) return fail(SE_ERR_INVALID_ARG, "null sim");
    SE_CUDA(cudaSetDevice(s->device));
    const size_t rowb = (size_t)s->W * sizeof(unsigned);
    // strip above (which = 0): my first `nb.ghost_bottom` owned rows become its bottom ghost rows
    if (s->nb[0].attached && s->nb[0].ghost_bottom > 0) {
        const Neighbour& nb = s->nb[0];
        const unsigned* src = s->cells[s->cur] + s->owned_offset();
        unsigned* dst = nb.cells[s->cur] + (size_t)(nb.local_rows - nb.ghost_bottom) * s->W;
        SE_CUDA(cudaMemcpyAsync(dst, src, rowb * nb.ghost_bottom, cudaMemcpyDeviceToDevice, s->stream));
    }
    // strip below (which = 1): my last `nb.ghost_top` owned rows become its top ghost rows
    if (s->nb[1].attached && s->nb[1].ghost_top > 0) {
        const Neighbour& nb = s->nb[1];
        const unsigned* src = s->cells[s->cur] + s->owned_offset() + (size_t)(s->row_end - s->row_begin - (int)nb.ghost_top) * s->W;
        unsigned* dst = nb.cells[s->cur];
        SE_CUDA(cudaMemcpyAsync(dst, src, rowb * nb.ghost_top, cudaMemcpyDeviceToDevice, s->stream));
    }
    if (s->lighting) {
        // lit strips: the same rows of the current light buffer (the neighbours step in lock-step, so lcur agrees)
        const size_t lrowb = (size_t)s->W * sizeof(float4);
        for (int w = 0; w < 2; ++w) {
            const Neighbour& nb = s->nb[w];
            const uint64_t g = (w == 0) ? nb.ghost_bottom : nb.ghost_top;
            if (!nb.attached || g == 0) continue;
            if (!nb.light[s->lcur]) return fail(SE_ERR_INVALID_ARG, "lit strip: the neighbour's light buffers are not attached (se_sim_ipc_attach_light)");
            const float4* src = s->light[s->lcur] + s->owned_offset() + (w == 0 ? 0 : (size_t)(s->row_end - s->row_begin - (int)g) * s->W);
            float4* dst = nb.light[s->lcur] + (w == 0 ? (size_t)(nb.local_rows - g) * s->W : 0);
            SE_CUDA(cudaMemcpyAsync(dst, src, lrowb * g, cudaMemcpyDeviceToDevice, s->stream));
        }
    }
    return SE_OK;
} SE_ABI_CATCH("se_sim_halo_push")

int se_sim_ipc_export_light(se_sim* s, void* handles) try {
    if (!s || !handles) return fail(SE_ERR_INVALID_ARG, "null argument");
    if (!s->lighting) return fail(SE_ERR_INVALID_ARG, "sim was created without SE_FLAG_LIGHTING");
    SE_CUDA(cudaSetDevice(s->device));
    std::memset(handles, 0, 128);
    for (int b = 0; b < 2; ++b) {
        cudaIpcMemHandle_t h;
        SE_CUDA(cudaIpcGetMemHandle(&h, s->light[b]));
        std::memcpy((char*)handles + 64 * b, &h, 64);
    }
    return SE_OK;
} SE_ABI_CATCH("se_sim_ipc_export_light")

int se_sim_ipc_attach_light(se_sim* s, int which, const void* handles) try {
    if (!s || !handles || which < 0 || which > 1) return fail(SE_ERR_INVALID_ARG, "bad argument");
    if (!s->lighting) return fail(SE_ERR_INVALID_ARG, "sim was created without SE_FLAG_LIGHTING");
    SE_CUDA(cudaSetDevice(s->device));
    Neighbour& nb = s->nb[which];
    if (nb.light_ipc)
        for (int b = 0; b < 2; ++b)
            if (nb.light[b]) { cudaIpcCloseMemHandle(nb.light[b]); nb.light[b] = nullptr; }
    nb.light_ipc = true;
    for (int b = 0; b < 2; ++b) {
        cudaIpcMemHandle_t h;
        std::memcpy(&h, (const char*)handles + 64 * b, 64);
        void* p = nullptr;
        SE_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        nb.light[b] = (float4*)p;
    }
    return SE_OK;
} SE_ABI_CATCH("se_sim_ipc_attach_light")

int se_sim_halo_exchange_async(se_sim* s) try {
    if (!s) return fail(SE_ERR_INVALID_ARG, "null sim");
    SE_CUDA(cudaSetDevice(s->device));
    if (!driver().ok) return fail(SE_ERR_CUDA, driver().why);
    const unsigned e = ++s->epoch;
    CUstream st = (CUstream)s->stream;
    // (a) tell both neighbours that this strip has finished every step enqueued so far
    for (int w = 0; w < 2; ++w)
        if (s->nb[w].attached) SE_CU(driver().StreamWriteValue32(st, (CUdeviceptr)(s->nb[w].flags + (w == 0 ? 1 : 0)), e, CU_STREAM_WRITE_VALUE_DEFAULT));
    // (b) wait until the neighbours have finished theirs: only then may their ghost rows be overwritten
    for (int w = 0; w < 2; ++w)
        if (s->nb[w].attached) SE_CU(driver().StreamWaitValue32(st, (CUdeviceptr)(s->flags + (w == 0 ? 0 : 1)), e, CU_STREAM_WAIT_VALUE_GEQ));
    // (c) push boundary rows over NVLink
    int rc = se_sim_halo_push(s);
    if (rc) return rc;
    // (d) publish "delivered"
    for (int w = 0; w < 2; ++w)
        if (s->nb[w].attached) SE_CU(driver().StreamWriteValue32(st, (CUdeviceptr)(s->nb[w].flags + (w == 0 ? 3 : 2)), e, CU_STREAM_WRITE_VALUE_DEFAULT));
    // (e) later work of this stream starts once this strip's own ghost rows have been delivered
    for (int w = 0; w < 2; ++w)
        if (s->nb[w].attached) SE_CU(driver().StreamWaitValue32(st, (CUdeviceptr)(s->flags + (w == 0 ? 2 : 3)), e, CU_STREAM_WAIT_VALUE_GEQ));
    return SE_OK;
} SE_ABI_CATCH("se_sim_halo_exchange_async")

}  // extern "C"
