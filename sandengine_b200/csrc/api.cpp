// C ABI implementation (include/sandengine_b200.h): rule compilation (front end + NVRTC) and the
// simulation host that replaces sandengine-core's `Simulation` (/root/reference/sandengine-core/src/simulation.rs).
//
// Device work is launched through the CUDA driver API on functions loaded from the NVRTC-built cubin;
// the driver entry points are resolved with cudaGetDriverEntryPoint so the library does not link
// libcuda and can be loaded (and can compile rules) on a host without a GPU.
#include "../../include/sandengine_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <type_traits>
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

// ---- NVRTC, loaded by path ---------------------------------------------------------------------
// The kernels are compiled by the NVRTC of the toolkit this library was built against (CUDA 12.9: its ptxas knows the
// three-input max and the packed f32 adds the lit kernel uses).  A process that imported PyTorch first already has torch's
// own, older libnvrtc.so.12 mapped, and a plain link dependency would silently bind to that one -- so the library is
// opened by absolute path (a path never matches an already loaded SONAME), with the default search as the fallback; the
// kernel source compiles with either (it checks __CUDACC_VER_MINOR__).  SE_NVRTC_LIB overrides the choice.
struct Nvrtc {
    decltype(&nvrtcCreateProgram) CreateProgram = nullptr;
    decltype(&nvrtcCompileProgram) CompileProgram = nullptr;
    decltype(&nvrtcGetProgramLogSize) GetProgramLogSize = nullptr;
    decltype(&nvrtcGetProgramLog) GetProgramLog = nullptr;
    decltype(&nvrtcGetErrorString) GetErrorString = nullptr;
    decltype(&nvrtcGetCUBINSize) GetCUBINSize = nullptr;
    decltype(&nvrtcGetCUBIN) GetCUBIN = nullptr;
    decltype(&nvrtcDestroyProgram) DestroyProgram = nullptr;
    decltype(&nvrtcVersion) Version = nullptr;
    bool ok = false;
    std::string why, path;
    Nvrtc() {
        std::vector<std::string> candidates;
        if (const char* e = std::getenv("SE_NVRTC_LIB")) candidates.push_back(e);
#ifdef SE_NVRTC_DIR
        candidates.push_back(std::string(SE_NVRTC_DIR) + "/libnvrtc.so.12");
#endif
        candidates.push_back("/usr/local/cuda/lib64/libnvrtc.so.12");
        candidates.push_back("libnvrtc.so.12");
        candidates.push_back("libnvrtc.so");
        void* h = nullptr;
        for (const auto& c : candidates) {
            h = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL | RTLD_DEEPBIND);
            if (h) { path = c; break; }
            why += c + ": " + (dlerror() ? "not loadable" : "?") + "; ";
        }
        if (!h) return;
        auto get = [&](auto& fn, const char* name) {
            fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(dlsym(h, name));
            if (!fn) why += std::string("missing ") + name + "; ";
            return fn != nullptr;
        };
        ok = get(CreateProgram, "nvrtcCreateProgram") & get(CompileProgram, "nvrtcCompileProgram") & get(GetProgramLogSize, "nvrtcGetProgramLogSize") &
             get(GetProgramLog, "nvrtcGetProgramLog") & get(GetErrorString, "nvrtcGetErrorString") & get(GetCUBINSize, "nvrtcGetCUBINSize") &
             get(GetCUBIN, "nvrtcGetCUBIN") & get(DestroyProgram, "nvrtcDestroyProgram") & get(Version, "nvrtcVersion");
    }
};
Nvrtc& nvrtc() {
    static Nvrtc n;
    return n;
}

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
    decltype(&cuTensorMapEncodeTiled) TensorMapEncodeTiled = nullptr;
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
               get("cuOccupancyMaxActiveBlocksPerMultiprocessor", (void**)&d.OccupancyMaxActiveBlocks) &&
               get("cuTensorMapEncodeTiled", (void**)&d.TensorMapEncodeTiled);
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
    unsigned* nbr_buf0[2];
    unsigned* nbr_buf1[2];
    unsigned* nbr_flags[2];
    const unsigned* in_flags[2];
    int nbr_gy0[2];
    int W, Hl, gy0, Hg;
    int own_y0, own_y1;
    int frame0, nblk, tsteps, nsub_last;
    unsigned seq_base;
    unsigned* done;
    unsigned* status;
    unsigned* queue;
    unsigned queue_base;
    int HY, HX, PH, THo;
    int tiles_x, tiles_y;
    int push_rows;
    int table_bytes, pool_offset, tile_offset, tile_stride;
    const unsigned* lut;
    const unsigned* pool;
    unsigned spin_limit;
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
    int table_bytes, pool_offset;
    const unsigned* lut;
    const unsigned* pool;
    int n_mods;
    const SeMod* mods;
};
struct SeLutCensusParams {
    unsigned long long* census;
    int own_y0, own_y1;
};
struct SeLightParams {
    const unsigned* old_cells;
    const unsigned* new_cells;
    const float4* light_in;
    float4* light_out;
    int W, Hl, gy0, Hg;
};

struct SeLitParams {
    unsigned* new_cells;
    float4* light_out;
    int W, Hl, gy0, Hg;
    int frame;
    int n_mods;
    const SeMod* mods;
    int table_bytes, pool_offset;
    const unsigned* lut;
    const unsigned* pool;
    int tiles_x, tiles_y;
    int buf_offset;
    unsigned tiles_x_magic;
};
// geometry of se_step_lit (kernels/sand_kernels.cuh: SE_LF_*).  Threads per half, rows per thread and buffer count can be overridden
// for experiments (SE_LF_NG = 2..4 groups, SE_LF_HALF = threads per group, SE_LF_ROWS = 2 | 4, SE_LF_NBUF = 2 | 3 in the environment: passed to NVRTC and used for the
// launch and the tensor maps alike).
constexpr int LF_TW = 64;
int lf_env(const char* name, int dflt, int lo, int hi) {
    const char* e = std::getenv(name);
    if (!e) return dflt;
    const int v = std::atoi(e);
    return v >= lo && v <= hi ? v : dflt;
}
int lf_ng() { return lf_env("SE_LF_NG", 3, 2, 4); }
int lf_half() { const int v = lf_env("SE_LF_HALF", 256, 256, 512); return v / 64 * 64; }
int lf_rows() { return lf_env("SE_LF_ROWS", 4, 2, 4) == 2 ? 2 : 4; }
int lf_th() { return lf_rows() * (lf_half() / 64); }
int lf_nbuf() { return lf_env("SE_LF_NBUF", 2, 2, 3); }
int lf_buf_bytes() {
    const int light = (lf_th() + 2) * (LF_TW + 8) * 16, ids = (lf_th() + 2) * (LF_TW + 16) * 4;
    return ((light + 127) / 128 * 128 + ids + 127) / 128 * 128;
}

// words of the per-sim flag block that sits behind cells[0] in the same allocation (one IPC handle maps both)
enum : int {
    FLAG_EPOCH = 0,       // [0]/[1] "done computing" epoch of the strip above/below, [2]/[3] "ghost rows delivered" from above/below
    FLAG_STATUS = 8,      // != 0: a tile of se_step_tiles gave up waiting for a flag (results invalid)
    FLAG_QUEUE = 9,       // work queue of se_step_tiles (never reset: the host knows its value at every launch)
    FLAG_TILE_IN = 64,    // [64, 64 + cap): sequence numbers of the boundary tiles of the strip ABOVE, [64 + cap, 64 + 2 cap): of the strip BELOW
};

}  // namespace

struct se_rules {
    se::CompiledRules cr;
    std::string glsl_materials, glsl_rules;
    std::vector<char> cubin;
    std::string nvrtc_log;
    bool compiled = false;
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
    // transition-table kernels (K1b tiles, K1c per frame)
    bool tiled = false;
    CUfunction f_tiles = nullptr, f_build_lut = nullptr, f_lut_global = nullptr, f_lut_global_mods = nullptr;
    unsigned* d_lut = nullptr;         // table image: N^4 * tables entries, then (mode 1, same allocation) the pool
    unsigned* d_pool = nullptr;        // mode 2: pool entries {thr, A, B} in global memory
    unsigned* d_tile_done = nullptr;   // per-tile sequence numbers (dataflow between the T-blocks of one launch)
    int tile_done_cap = 0;
    unsigned tile_seq = 0;             // T-blocks completed so far
    unsigned tile_queue = 0;           // value of the device-side work queue counter when the next launch starts
    // se_step_tiles spins on flags written by other CTAs of the same grid, so all of its CTAs must be resident
    // together: it is launched cooperatively (the driver then places the whole grid at once, even when another
    // stream's kernels share the device).  Without cooperative launch support every launch carries ONE T-block
    // (its tiles wait on no flag of the same grid), so partial residency cannot deadlock.
    bool coop = false;
    int lut_mode = 0;
    int T = 0, HY = 0, HX = 0, PH_max = 0, tiles_x = 0, table_bytes = 0, pool_offset = 0, tile_offset = 0, tile_grid = 0, k1c_grid = 0, k1c_smem = 0;
    int smem_budget = 0;
    int device_share = 1;
    // ghost rows (strips): rows of each ghost zone, counted from the owned rows outwards, that hold the neighbour's
    // current state.  The per-step kernels recompute the ghost rows redundantly and use the zone up from the outside
    // (one row every other frame; one row per frame with lighting); the tile kernel refreshes `push_rows` of them
    // in every T-block.  se_sim_step exchanges (device-ordered) whenever the next launch needs more than there is.
    int ghost_valid = 0;
    // fused step + modifications + lighting (K3f, se_step_lit): tensor maps of the two id buffers and the two light buffers
    bool fused_light = false;
    CUfunction f_step_lit = nullptr;
    CUtensorMap tm_cells[2], tm_light[2];
    int lf_smem = 0, lf_grid = 0, lf_tiles_x = 0, lf_tiles_y = 0, lf_buf_offset = 0;
    // running census (SE_FLAG_RUNNING_CENSUS, experimental): d_running is valid only between K1c-census steps
    bool running = false, running_valid = false, running_copy_pending = false;
    CUfunction f_lut_global_census = nullptr, f_lut_global_census_mods = nullptr;
    unsigned* d_lut_census = nullptr;      // copy of the table image whose outcomes carry SE_E_POPFLAG (kernels/sand_kernels.cuh)
    unsigned long long* d_running = nullptr;
    cudaEvent_t running_copy_done = nullptr;
    Neighbour nb[2];
    // Device-side exchange protocol: 4 flag words live right behind cells[0] (same allocation, so that one
    // IPC handle maps both): [0]/[1] = "done computing" epoch of the strip above/below, [2]/[3] = "ghost rows
    // delivered" epoch from above/below.  Written by the neighbours, waited on by this sim's stream.
    unsigned* flags = nullptr;
    unsigned epoch = 0;
    static size_t flags_offset(size_t W_, size_t Hl_) { return ((W_ * Hl_ * sizeof(unsigned)) + 255) / 256 * 256; }
    static size_t tile_flag_cap(size_t W_) { return W_ / 192 + 2; }             // tiles per row for the widest halo (HX <= 32)
    static size_t flags_bytes(size_t W_) { return (((size_t)FLAG_TILE_IN + 2 * tile_flag_cap(W_)) * sizeof(unsigned) + 255) / 256 * 256; }
    bool is_strip() const { return ghost_top > 0 || ghost_bottom > 0; }
    bool has_neighbours() const { return nb[0].attached || nb[1].attached; }
    int ghost_rows() const { return ghost_top > 0 && ghost_bottom > 0 ? std::min(ghost_top, ghost_bottom) : std::max(ghost_top, ghost_bottom); }
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
        if (!nvrtc().ok) {
            delete r;
            return fail(SE_ERR_COMPILE, "NVRTC not available: " + nvrtc().why);
        }
        nvrtcProgram prog;
        const char* hdr_src[1] = {r->cr.cuda_header.c_str()};
        const char* hdr_name[1] = {"rules_gen.cuh"};
        // experiments only: SE_KERNEL_SOURCE_FILE=<path> compiles that file instead of the embedded kernel source (A/B runs of two
        // kernel versions on the same GPU box)
        std::string source_override;
        if (const char* sf = std::getenv("SE_KERNEL_SOURCE_FILE")) {
            if (FILE* f = std::fopen(sf, "rb")) {
                char buf[65536];
                size_t n;
                while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) source_override.append(buf, n);
                std::fclose(f);
            }
        }
        if (nvrtc().CreateProgram(&prog, source_override.empty() ? KERNEL_SOURCE : source_override.c_str(), "sand_kernels.cu", 1, hdr_src, hdr_name) != NVRTC_SUCCESS) {
            delete r;
            return fail(SE_ERR_COMPILE, "nvrtcCreateProgram failed");
        }
        // -fmad=false: the lighting sums and any float arithmetic in rule conditions are evaluated as
        // written (no FMA contraction), matching the reference expression tree.
        std::vector<std::string> extra;                       // experiments only: SE_NVRTC_DEFS="-DX=1 -DY=2"
        if (const char* lr = std::getenv("SE_LT_ROWS")) {     // experiments only: se_light tile height
            const int v = std::atoi(lr);
            if (v == 2 || v == 4 || v == 8) { r->light_rows = v; extra.push_back("-DSE_LT_ROWS=" + std::to_string(v)); }
        }
        if (const char* mc = std::getenv("SE_LT_MINCTAS")) {
            const int v = std::atoi(mc);
            if (v >= 1 && v <= 8) extra.push_back("-DSE_LT_MINCTAS=" + std::to_string(v));
        }
        extra.push_back("-DSE_LF_NG=" + std::to_string(lf_ng()));
        extra.push_back("-DSE_LF_HALF=" + std::to_string(lf_half()));
        extra.push_back("-DSE_LF_ROWS=" + std::to_string(lf_rows()));
        extra.push_back("-DSE_LF_NBUF=" + std::to_string(lf_nbuf()));
        if (const char* defs = std::getenv("SE_NVRTC_DEFS")) {
            std::string d(defs), tok;
            for (size_t i = 0; i <= d.size(); ++i) {
                if (i == d.size() || d[i] == ' ') { if (!tok.empty()) extra.push_back(tok); tok.clear(); }
                else tok.push_back(d[i]);
            }
        }
        std::vector<const char*> opts = {"-arch=sm_100a", "-std=c++17", "-lineinfo", "-fmad=false"};
        for (auto& e : extra) opts.push_back(e.c_str());
        nvrtcResult res = nvrtc().CompileProgram(prog, (int)opts.size(), opts.data());
        size_t log_size = 0;
        nvrtc().GetProgramLogSize(prog, &log_size);
        r->nvrtc_log.clear();
        if (log_size > 1) {
            std::vector<char> log(log_size);
            nvrtc().GetProgramLog(prog, log.data());
            r->nvrtc_log.assign(log.data());
        }
        {
            int vmaj = 0, vmin = 0;
            nvrtc().Version(&vmaj, &vmin);
            r->nvrtc_log += "[NVRTC " + std::to_string(vmaj) + "." + std::to_string(vmin) + " from " + nvrtc().path + "]\n";
        }
        if (res != NVRTC_SUCCESS) {
            std::string msg = "NVRTC: " + std::string(nvrtc().GetErrorString(res)) + "\n" + r->nvrtc_log;
            nvrtc().DestroyProgram(&prog);
            delete r;
            return fail(SE_ERR_COMPILE, msg);
        }
        size_t cubin_size = 0;
        if (nvrtc().GetCUBINSize(prog, &cubin_size) != NVRTC_SUCCESS || cubin_size == 0) {
            nvrtc().DestroyProgram(&prog);
            delete r;
            return fail(SE_ERR_COMPILE, "NVRTC produced no cubin");
        }
        r->cubin.resize(cubin_size);
        nvrtc().GetCUBIN(prog, r->cubin.data());
        nvrtc().DestroyProgram(&prog);
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
    if (cooperative)
        SE_CU(driver().LaunchCooperativeKernel(f, grid.x, grid.y, grid.z, block.x, block.y, block.z, smem, (CUstream)s->stream, args));
    else
        SE_CU(driver().LaunchKernel(f, grid.x, grid.y, grid.z, block.x, block.y, block.z, smem, (CUstream)s->stream, args, nullptr));
    s->launches++;
    return SE_OK;
}

int halo_exchange_async(se_sim* s);

// Strips: make sure at least `need` ghost rows per side hold the neighbours' current rows before the next launch
// (device-ordered exchange, no host synchronisation).  A strip without attached neighbours (a stand-alone probe of one
// strip) is left alone: its ghost rows are then simply the caller's business.
// A tile of se_step_tiles that gave up waiting for a neighbour strip's flag leaves a mark instead of hanging the device.
// Call with the stream idle.
int check_status(se_sim* s) {
    if (!s->tiled || !s->is_strip()) return SE_OK;
    unsigned st = 0;
    SE_CUDA(cudaMemcpy(&st, s->flags + FLAG_STATUS, sizeof st, cudaMemcpyDeviceToHost));
    if (st == 0) return SE_OK;
    SE_CUDA(cudaMemset(s->flags + FLAG_STATUS, 0, sizeof st));
    return fail(SE_ERR_CUDA, "a strip waited in vain for a neighbour strip (step every strip of the grid before synchronising any of them); the cell state is invalid");
}

int ensure_ghosts(se_sim* s, int need) {
    if (!s->is_strip() || !s->has_neighbours() || s->ghost_valid >= need) return SE_OK;
    return halo_exchange_async(s);
}

int one_step(se_sim* s, bool use_mods, int n_mods) {
    { int rc = guard_buffer_write(s, 0); if (rc) return rc; rc = guard_buffer_write(s, 1); if (rc) return rc; }
    const int frame = s->frame + 1;
    const int ox = ((frame & 3) == 1 || (frame & 3) == 3) ? 1 : 0;
    const int oy = ((frame & 3) == 1 || (frame & 3) == 2) ? 1 : 0;
    // Ghost rows: a step keeps the outermost valid ghost row valid only when that row's block lies inside the valid
    // zone (the strip boundaries are even rows, so the row is a block's inner row iff ghost_valid + oy is even); the
    // 8-neighbour light stencil (operations.glsl:114-160) uses up one more row every frame.
    if (s->is_strip() && frame != 1) {
        const int after = s->lighting ? s->ghost_valid - 1 : s->ghost_valid - ((s->ghost_valid + oy) & 1);
        if (after < 0) { int rc = ensure_ghosts(s, s->ghost_rows()); if (rc) return rc; }
    } else if (s->is_strip() && s->lighting) {
        int rc = ensure_ghosts(s, 1); if (rc) return rc;
    }
    s->frame = frame;
    if (frame == 1) {
        // falling_sand.glsl:743-746: every cell becomes EMPTY; modifications are ignored this frame.
        size_t n = (size_t)s->W * s->Hl;
        unsigned zero = 0;
        if (!s->lighting) {
            unsigned* buf = s->cells[s->cur];
            void* args[] = {&buf, &n, &zero};
            s->ghost_valid = s->ghost_rows();               // the ghost rows are EMPTY like everything else
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
        if (s->is_strip()) s->ghost_valid = std::max(0, s->ghost_valid - 1);
        return SE_OK;
    }
    if (s->is_strip()) s->ghost_valid = std::max(0, s->lighting ? s->ghost_valid - 1 : s->ghost_valid - ((s->ghost_valid + oy) & 1));
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
        // K3f: one TMA-fed pass reads ids + light of the current buffers and writes the other pair
        SeLitParams lp;
        lp.new_cells = s->cells[s->cur ^ 1]; lp.light_out = s->light[s->lcur ^ 1];
        lp.W = s->W; lp.Hl = s->Hl; lp.gy0 = s->gy0; lp.Hg = s->Hg; lp.frame = frame;
        lp.n_mods = p.n_mods; lp.mods = s->d_mods;
        lp.table_bytes = s->table_bytes; lp.pool_offset = s->pool_offset; lp.lut = s->d_lut;
        lp.pool = s->d_pool ? s->d_pool : s->d_lut + s->pool_offset / 4;      // mode 1: the pool sits behind the table in one image
        lp.tiles_x = s->lf_tiles_x; lp.tiles_y = s->lf_tiles_y; lp.buf_offset = s->lf_buf_offset;
        lp.tiles_x_magic = s->lf_tiles_x > 1 ? (unsigned)((1ull << 32) / (unsigned long long)s->lf_tiles_x + 1ull) : 0u;
        void* fargs[] = {&s->tm_cells[s->cur], &s->tm_light[s->lcur], &lp};
        int rcf = launch(s, s->f_step_lit, dim3(s->lf_grid), dim3(lf_ng() * lf_half()), fargs, (unsigned)s->lf_smem);
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

// ---- K1b geometry: tile rows of one launch ------------------------------------------------------------------
// The tiles' interiors partition the OWNED rows into tiles_y rows of THo rows (the last one may be shorter).  Every
// half-CTA processes ceil(items / workers) tiles of PH = THo + 2 HY rows, so the launch costs about
// rounds * (PH + c): pick the tiles_y that minimises it (avoids a nearly empty last round, which is what costs
// strong scaling: at 16384 x 2048 rows per GPU a fixed PH would spend 9 rounds on 8.4 rounds of work).
// The halo is sized for the sub-steps a T-block of THIS launch really has (20 steps with T = 12 are two T-blocks of 10: a halo
// of 12 columns and 6 rows, not the 8 rows a T-block of 12 would need): HX = sub-steps rounded up to 4, HY = even >= n/2 + 1.
struct TileGeom { int tiles_y, THo, PH, HX, HY, tiles_x; };

TileGeom choose_tile_geometry(const se_sim* s, int nblk, int tsteps, int owned) {
    const int workers = 2 * s->tile_grid;
    const int HY = ((tsteps / 2 + 1) + 1) & ~1, HX = (tsteps + 3) & ~3;
    const int tiles_x = (s->W + (256 - 2 * HX) - 1) / (256 - 2 * HX);
    const int push_min = s->is_strip() ? s->HY : 2;
    TileGeom best{0, 0, 0};
    double best_cost = -1;
    const int tho_max = (s->PH_max - 2 * HY) & ~1;
    const int ty_min = std::max(1, (owned + tho_max - 1) / tho_max);
    for (int ty = ty_min; ty <= ty_min + 4 * workers; ++ty) {
        int tho = ((owned + ty - 1) / ty + 1) & ~1;
        if (tho > tho_max) continue;
        if (tho < 2 * tsteps + 8 && ty > ty_min) break;              // tiles this flat are mostly halo
        const int last = owned - (ty - 1) * tho;
        if (last <= 0 || (ty > 1 && last < push_min)) continue;
        const long long items = (long long)nblk * tiles_x * ty;
        // Measured on B200 (scripts/geom_sweep.py, profiles/r2_geom_sweep.txt): a tile costs what its sub-steps cost, and a sub-step
        // is ceil(PH / 32) block-row iterations of the half's 16 warps -- with a super-linear tail above six iterations (the two
        // halves of an SM overlap each other's load / store phases less well with tall tiles).  kIter[i]: relative cost of a tile of
        // i iterations; the work queue is dynamic, so the makespan is items / workers + about half a tile, not whole rounds.
        static const double kIter[] = {0.030, 0.030, 0.035, 0.0437, 0.0559, 0.0685, 0.0800, 0.0992, 0.1131, 0.1270, 0.1410};
        const int iters = (tho + 2 * HY + 31) / 32;
        const double tile_cost = iters <= 10 ? kIter[iters] : kIter[10] + 0.014 * (iters - 10);
        const double cost = ((double)items / workers + 0.5) * tile_cost;
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = TileGeom{ty, tho, tho + 2 * HY, HX, HY, tiles_x}; }
    }
    if (best.tiles_y == 0) {                                          // cannot happen for owned >= 2; keep a safe answer anyway
        const int tho = std::min(tho_max, (owned + 1) & ~1);
        best = TileGeom{(owned + tho - 1) / tho, tho, tho + 2 * HY, HX, HY, tiles_x};
    }
    if (const char* ov = std::getenv("SE_TILE_PH")) {                // experiments only: fix the tile height
        const int v = (std::atoi(ov) & ~1) - 2 * HY;
        if (v >= 2 && v <= tho_max) best = TileGeom{(owned + v - 1) / v, v, v + 2 * HY, HX, HY, tiles_x};
    }
    return best;
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
    if (s->d_pool) cudaFree(s->d_pool);
    if (s->d_tile_done) cudaFree(s->d_tile_done);
    if (s->d_lut_census) cudaFree(s->d_lut_census);
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
    // per-step kernels index block rows with gridDim.y (<= 65535 CTAs of 4 block rows / 8 light rows)
    if ((uint64_t)(re - rb) + 2ull * prm->halo_rows > 65535ull * 8ull)
        return fail(SE_ERR_INVALID_ARG, "more than 524280 rows per device are not supported (shard the grid into strips)");
    if ((prm->halo_rows & 1u) != 0u) return fail(SE_ERR_INVALID_ARG, "halo_rows must be even (strip buffers start on even rows)");
    if (prm->device_share > 64u) return fail(SE_ERR_INVALID_ARG, "device_share out of range");
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
    const size_t flags_len = se_sim::flags_bytes((size_t)s->W);
    SE_CUDA_S(cudaMalloc(&s->cells[0], flags_off + flags_len));
    SE_CUDA_S(cudaMemsetAsync(s->cells[0], 0, flags_off + flags_len, s->stream));
    s->flags = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(s->cells[0]) + flags_off);
    s->device_share = prm->device_share ? (int)prm->device_share : 1;
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

    // ---- transition-table kernels: K1b (tiles + temporal blocking) and K1c (one step, per frame) ----------------
    // Used for steps without modifications when lighting is off, the rule set is table-eligible (codegen.h) and rows
    // are 16-byte aligned.  temporal_block == 1 forces the per-step generated-code kernel K1a.
    if (rules->cr.lut_eligible && (s->W % 4) == 0 && prm->temporal_block != 1) {
        const int N = rules->cr.tables.n_materials;
        const size_t N4 = (size_t)N * N * N * N;
        const size_t NE = N4 * (size_t)rules->cr.lut_tables;            // table entries (one table per view for Left/Right rule sets and in mode 2)
        s->lut_mode = rules->cr.lut_mode;
        int smem_optin = 0, n_sm = 0, coop_attr = 0;
        SE_CUDA_S(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, s->device));
        SE_CUDA_S(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, s->device));
        SE_CUDA_S(cudaDeviceGetAttribute(&coop_attr, cudaDevAttrCooperativeLaunch, s->device));
        s->smem_budget = smem_optin - 1024 - 64;                        // static shared memory of the kernels: 1 KB fat table
        // default 12: load + store of a tile cost about three sub-steps, the halo grows with T; measured at 16384^2 over 20
        // steps: T = 6 / 8 / 10 -> 1110 / 1212 / 1308 Gcell/s, and by that model 12 .. 16 are best for long runs
        int T = prm->temporal_block ? (int)prm->temporal_block : 12;
        T = std::min(32, std::max(2, T + (T & 1)));
        // rows: the Margolus row offset changes every other frame, so T fused steps need only T/2+1 halo rows
        s->T = T;
        s->HY = ((T / 2 + 1) + 1) & ~1;
        s->HX = (T + 3) & ~3;
        const int min_tile = 2 * 256 * (4 * T + 16);                    // both halves must hold a tile of at least 4T+16 rows
        SE_CU_S(driver().ModuleGetFunction(&s->f_tiles, s->mod, "se_step_tiles"));
        SE_CU_S(driver().ModuleGetFunction(&s->f_build_lut, s->mod, "se_build_lut"));
        SE_CU_S(driver().ModuleGetFunction(&s->f_lut_global, s->mod, "se_step_lut_global"));
        SE_CU_S(driver().ModuleGetFunction(&s->f_lut_global_mods, s->mod, "se_step_lut_global_mods"));
        unsigned* d_counter = nullptr;
        SE_CUDA_S(cudaMalloc(&d_counter, sizeof(unsigned)));
        SE_CUDA_S(cudaMemsetAsync(d_counter, 0, sizeof(unsigned), s->stream));
        unsigned n_pool = 0;
        bool ok = false;
        if (s->lut_mode == 1) {
            // mode 1: table + pool in one image that every CTA stages into shared memory
            const size_t pool_off = (NE * 4 + 15) / 16 * 16;                // pool entries are read with 128-bit loads
            const long long cap = ((long long)s->smem_budget - (long long)pool_off - min_tile) / 16;
            if (cap >= 0) {
                const size_t image = pool_off + (size_t)cap * 16 + 16;
                SE_CUDA_S(cudaMalloc(&s->d_lut, image));
                SE_CUDA_S(cudaMemsetAsync(s->d_lut, 0, image, s->stream));
                unsigned* base = s->d_lut;
                unsigned* pool = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(s->d_lut) + pool_off);
                unsigned ucap = (unsigned)cap, flag_pop = 0;
                void* bargs[] = {&base, &pool, &d_counter, &ucap, &flag_pop};
                SE_TRY(launch(s, s->f_build_lut, dim3((unsigned)((NE + 255) / 256)), dim3(256), bargs));
                SE_CUDA_S(cudaMemcpyAsync(&n_pool, d_counter, sizeof(unsigned), cudaMemcpyDeviceToHost, s->stream));
                SE_CUDA_S(cudaStreamSynchronize(s->stream));
                if ((long long)n_pool <= cap) {
                    s->pool_offset = (int)pool_off;
                    s->table_bytes = (int)(pool_off + (size_t)n_pool * 16);
                    s->tile_offset = (s->table_bytes + 15) / 16 * 16;
                    ok = true;
                }
            }
        } else {
            // mode 2: the table stays in global memory.  Pass 1 counts the pool entries, pass 2 fills them.
            SE_CUDA_S(cudaMalloc(&s->d_lut, NE * 4));
            unsigned* base = s->d_lut;
            unsigned* pool = nullptr;
            unsigned ucap = 0, flag_pop = 0;
            void* bargs[] = {&base, &pool, &d_counter, &ucap, &flag_pop};
            SE_TRY(launch(s, s->f_build_lut, dim3((unsigned)((NE + 255) / 256)), dim3(256), bargs));
            SE_CUDA_S(cudaMemcpyAsync(&n_pool, d_counter, sizeof(unsigned), cudaMemcpyDeviceToHost, s->stream));
            SE_CUDA_S(cudaStreamSynchronize(s->stream));
            if (n_pool < 0x3FFFFFFFu) {
                SE_CUDA_S(cudaMalloc(&s->d_pool, (size_t)std::max(1u, n_pool) * 16));
                SE_CUDA_S(cudaMemsetAsync(d_counter, 0, sizeof(unsigned), s->stream));
                pool = s->d_pool;
                ucap = n_pool;
                SE_TRY(launch(s, s->f_build_lut, dim3((unsigned)((NE + 255) / 256)), dim3(256), bargs));
                SE_CUDA_S(cudaStreamSynchronize(s->stream));
                s->pool_offset = 0;
                s->table_bytes = 0;
                s->tile_offset = 0;
                ok = true;
            }
        }
        cudaFree(d_counter);
        const int owned = s->row_end - s->row_begin;
        // ---- K3f: with lighting on, step + override + lighting in one TMA-fed kernel (se_step_lit) ----
        if (ok && s->lighting && (s->W % 8) == 0 && !std::getenv("SE_NO_FUSED_LIT")) {   // W % 8: the id box moves in groups of 8;     // env: A/B against the two-kernel path
            // one CTA of two halves per SM, LF_NBUF input buffers per half (+ 128: the kernel aligns its buffers itself); the table
            // is read where it lies in global memory
            s->lf_buf_offset = 0;
            s->lf_smem = lf_ng() * lf_nbuf() * lf_buf_bytes() + 128;
            if (s->lf_smem + 16 * 1024 <= smem_optin) {
                SE_CU_S(driver().ModuleGetFunction(&s->f_step_lit, s->mod, "se_step_lit"));
                SE_CU_S(driver().FuncSetAttribute(s->f_step_lit, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, s->lf_smem));
                s->lf_tiles_x = (s->W + LF_TW - 1) / LF_TW;
                s->lf_tiles_y = (s->Hl + lf_th() - 1) / lf_th();
                s->lf_grid = (int)std::min<long long>((long long)std::max(1, n_sm / s->device_share), ((long long)s->lf_tiles_x * s->lf_tiles_y + lf_ng() - 1) / lf_ng());
                bool maps_ok = true;
                for (int b = 0; b < 2 && maps_ok; ++b) {
                    // Box rows are 64 bytes of light and 32 bytes of ids: 16-byte rows (a float4, four ids) made the TMA unit the
                    // bottleneck.  (A 2-D box with a 272-byte row is accepted by the encoder and then faults as an illegal instruction.)
                    // ids: 3-D [Hl][W/8][8] u32, box 8 x 10 x (TH + 2): the box starts 8 columns left of the tile; out-of-range
                    // elements read 0
                    const cuuint64_t cdim[3] = {8, (cuuint64_t)s->W / 8, (cuuint64_t)s->Hl}, cstr[2] = {32, (cuuint64_t)s->W * 4};
                    const cuuint32_t cbox[3] = {8, (LF_TW + 16) / 8, (cuuint32_t)lf_th() + 2}, cel[3] = {1, 1, 1};
                    maps_ok = driver().TensorMapEncodeTiled(&s->tm_cells[b], CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, s->cells[b], cdim, cstr, cbox, cel,
                                                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
                    // light: 3-D [Hl][W/4][16] f32 (four cells), box 16 x 18 x (TH + 2): starts 4 columns left of the tile
                    const cuuint64_t ldim[3] = {16, (cuuint64_t)s->W / 4, (cuuint64_t)s->Hl}, lstr[2] = {64, (cuuint64_t)s->W * 16};
                    const cuuint32_t lbox[3] = {16, (LF_TW + 8) / 4, (cuuint32_t)lf_th() + 2}, lel[3] = {1, 1, 1};
                    maps_ok = maps_ok && driver().TensorMapEncodeTiled(&s->tm_light[b], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, s->light[b], ldim, lstr, lbox, lel,
                                                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
                }
                s->fused_light = maps_ok;
            }
        }
        // strips: the tile kernel reads HY ghost rows and pushes as many into the neighbours
        if (ok && s->is_strip() && ((int)prm->halo_rows < s->HY || owned < 2 * s->HY)) ok = false;
        if (ok && !s->lighting) {
            int PH_max = (((s->smem_budget - s->tile_offset) / 2) / 256) & ~1;
            PH_max = std::min(PH_max, 320);                            // taller tiles: fewer, coarser work items for the same bytes
            if (const char* pm = std::getenv("SE_TILE_PH_MAX")) PH_max = std::min(PH_max, std::max(4 * T + 16, std::atoi(pm) & ~1));   // experiments only
            s->PH_max = PH_max;
            s->tiles_x = (s->W + (256 - 2 * 4) - 1) / (256 - 2 * 4);      // the most tile columns any launch can have (HX >= 4)
            s->tile_grid = std::max(1, n_sm / s->device_share);        // persistent: one CTA (two halves) per SM
            s->k1c_grid = std::max(1, 2 * n_sm / s->device_share);
            if (const char* kg = std::getenv("SE_K1C_GRID")) s->k1c_grid = std::max(1, std::atoi(kg));   // experiments only
            s->k1c_smem = s->tile_offset;
            const int tile_smem_max = s->tile_offset + 2 * 256 * PH_max;
            SE_CU_S(driver().FuncSetAttribute(s->f_tiles, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, tile_smem_max));
            if (s->k1c_smem > 0) SE_CU_S(driver().FuncSetAttribute(s->f_lut_global, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, s->k1c_smem));
            if (s->k1c_smem > 0) SE_CU_S(driver().FuncSetAttribute(s->f_lut_global_mods, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, s->k1c_smem));
            int occ = 0;
            SE_CU_S(driver().OccupancyMaxActiveBlocks(&occ, s->f_tiles, 1024, (size_t)tile_smem_max));
            if (occ < 1) { fail(SE_ERR_CUDA, "se_step_tiles does not fit on an SM"); return bail(SE_ERR_CUDA); }
            s->coop = coop_attr != 0 && !std::getenv("SE_NO_COOP_LAUNCH");   // env: experiments only
            SE_CUDA_S(cudaMalloc(&s->cells[1], s->cells_bytes()));
            SE_CUDA_S(cudaMemsetAsync(s->cells[1], 0, s->cells_bytes(), s->stream));
            s->tile_done_cap = s->tiles_x * (owned / 8 + 2);
            SE_CUDA_S(cudaMalloc(&s->d_tile_done, (size_t)s->tile_done_cap * sizeof(unsigned)));
            SE_CUDA_S(cudaMemsetAsync(s->d_tile_done, 0, (size_t)s->tile_done_cap * sizeof(unsigned), s->stream));
            s->tiled = true;
            if ((prm->flags & SE_FLAG_RUNNING_CENSUS) && rules->cr.lut_tables == 1 && s->lut_mode == 1) {
                // the census variant of K1c stages a second image of the table: same entries, population-changing outcomes flagged
                SE_CU_S(driver().ModuleGetFunction(&s->f_lut_global_census, s->mod, "se_step_lut_global_census"));
                SE_CU_S(driver().ModuleGetFunction(&s->f_lut_global_census_mods, s->mod, "se_step_lut_global_census_mods"));
                const size_t image = (size_t)s->tile_offset + 16;
                SE_CUDA_S(cudaMalloc(&s->d_lut_census, image));
                SE_CUDA_S(cudaMemsetAsync(s->d_lut_census, 0, image, s->stream));
                unsigned* d_cnt = nullptr;
                SE_CUDA_S(cudaMalloc(&d_cnt, sizeof(unsigned)));
                SE_CUDA_S(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned), s->stream));
                unsigned* base = s->d_lut_census;
                unsigned* pool = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(s->d_lut_census) + s->pool_offset);
                unsigned ucap = (unsigned)((s->table_bytes - s->pool_offset) / 16), flag_pop = 1;
                void* cargs[] = {&base, &pool, &d_cnt, &ucap, &flag_pop};
                SE_TRY(launch(s, s->f_build_lut, dim3((unsigned)((NE + 255) / 256)), dim3(256), cargs));
                SE_CUDA_S(cudaStreamSynchronize(s->stream));
                cudaFree(d_cnt);
                SE_CUDA_S(cudaMalloc(&s->d_running, 256 * sizeof(unsigned long long)));
                SE_CUDA_S(cudaEventCreateWithFlags(&s->running_copy_done, cudaEventDisableTiming));
                SE_CU_S(driver().FuncSetAttribute(s->f_lut_global_census, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, s->k1c_smem));
                SE_CU_S(driver().FuncSetAttribute(s->f_lut_global_census_mods, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, s->k1c_smem));
                s->running = true;
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
        if (s->tiled && s->frame + 1 != 1) {
            if (n_steps - k == 1 || mods_now) {
                // a lone step (the per-frame path) or the step that consumes the frame's modifications (a brush held down must
                // not drop the per-frame path to the generated-code kernel): K1c, table transitions straight from global memory,
                // in place, the override applied per cell where a record can touch the warp's work item
                { int rc = guard_buffer_write(s, s->cur); if (rc) return rc; }
                const int frame = s->frame + 1;
                const int oy = ((frame & 3) == 1 || (frame & 3) == 2) ? 1 : 0;
                if (s->is_strip()) {
                    if (s->ghost_valid - ((s->ghost_valid + oy) & 1) < 0) { int rc = ensure_ghosts(s, s->ghost_rows()); if (rc) return rc; }
                    s->ghost_valid = std::max(0, s->ghost_valid - ((s->ghost_valid + oy) & 1));
                }
                SeLutStepParams lp;
                lp.cells = s->cells[s->cur];
                lp.W = s->W; lp.Hl = s->Hl; lp.gy0 = s->gy0; lp.Hg = s->Hg; lp.frame = frame;
                lp.table_bytes = s->table_bytes; lp.pool_offset = s->pool_offset; lp.lut = s->d_lut; lp.pool = s->d_pool;
                lp.n_mods = mods_now ? n_mods : 0; lp.mods = s->d_mods;
                void* largs[] = {&lp};
                int rc;
                if (s->running && s->running_valid) {
                    SeLutCensusParams cx{s->d_running, s->row_begin, s->row_end};
                    lp.lut = s->d_lut_census;
                    void* cargs[] = {&lp, &cx};
                    rc = launch(s, mods_now ? s->f_lut_global_census_mods : s->f_lut_global_census, dim3(s->k1c_grid), dim3(512), cargs, (unsigned)s->k1c_smem);
                } else {
                    rc = launch(s, mods_now ? s->f_lut_global_mods : s->f_lut_global, dim3(s->k1c_grid), dim3(512), largs, (unsigned)s->k1c_smem);   // SE_K1C_THREADS
                }
                if (rc) return rc;
                s->frame += 1;
                k += 1;
                continue;
            }
            // a run of plain steps: ONE launch of nblk T-blocks of (nearly) equal length -- 20 steps are 7 + 7 + 6, not
            // 8 + 8 + 4 -- with dataflow between them inside the kernel.  Without cooperative launch: one T-block per launch.
            s->running_valid = false;
            { int rc = guard_buffer_write(s, 0); if (rc) return rc; rc = guard_buffer_write(s, 1); if (rc) return rc; }
            uint32_t run = std::min<uint32_t>(n_steps - k, 64u * (uint32_t)s->T);     // bound the kernel duration
            if (!s->coop) run = std::min<uint32_t>(run, (uint32_t)s->T);
            const int nblk = (int)((run + (uint32_t)s->T - 1) / (uint32_t)s->T);
            const int tsteps = (int)((run + (uint32_t)nblk - 1) / (uint32_t)nblk);
            { int rc = ensure_ghosts(s, s->HY); if (rc) return rc; }
            const TileGeom g = choose_tile_geometry(s, nblk, tsteps, s->row_end - s->row_begin);
            if ((long long)g.tiles_x * g.tiles_y > s->tile_done_cap) return fail(SE_ERR_INTERNAL, "tile flag array too small");
            SeTileParams tp;
            std::memset(&tp, 0, sizeof tp);
            tp.buf0 = s->cells[s->cur]; tp.buf1 = s->cells[s->cur ^ 1];
            const size_t cap = se_sim::tile_flag_cap((size_t)s->W);
            int push_rows = 0;
            for (int w = 0; w < 2; ++w) {
                const Neighbour& nb = s->nb[w];
                const int nb_ghost = (int)(w == 0 ? nb.ghost_bottom : nb.ghost_top);      // the neighbour's ghost zone that my rows fill
                if (!nb.attached || nb_ghost == 0) continue;
                tp.nbr_buf0[w] = nb.cells[s->cur]; tp.nbr_buf1[w] = nb.cells[s->cur ^ 1];
                // my tiles publish in the neighbour's "from below" (w == 0: I am below it) / "from above" array
                tp.nbr_flags[w] = nb.flags + FLAG_TILE_IN + (w == 0 ? cap : 0);
                tp.in_flags[w] = s->flags + FLAG_TILE_IN + (w == 0 ? 0 : cap);
                tp.nbr_gy0[w] = (w == 0) ? (s->row_begin - ((int)nb.local_rows - (int)nb.ghost_bottom)) : (s->row_end - (int)nb.ghost_top);
                push_rows = push_rows ? std::min(push_rows, nb_ghost) : nb_ghost;
            }
            // rows pushed per boundary tile: the whole ghost zone of the neighbour when the boundary tile rows hold that many
            const int last_rows = (s->row_end - s->row_begin) - (g.tiles_y - 1) * g.THo;
            push_rows = std::min(push_rows, std::min(g.THo, last_rows));
            tp.W = s->W; tp.Hl = s->Hl; tp.gy0 = s->gy0; tp.Hg = s->Hg;
            tp.own_y0 = s->row_begin; tp.own_y1 = s->row_end;
            tp.frame0 = s->frame + 1; tp.nblk = nblk; tp.tsteps = tsteps; tp.nsub_last = (int)(run - (uint32_t)(nblk - 1) * (uint32_t)tsteps);
            tp.seq_base = s->tile_seq; tp.done = s->d_tile_done; tp.status = s->flags + FLAG_STATUS;
            tp.queue = s->flags + FLAG_QUEUE; tp.queue_base = s->tile_queue;
            tp.HY = g.HY; tp.HX = g.HX; tp.PH = g.PH; tp.THo = g.THo;
            tp.tiles_x = g.tiles_x; tp.tiles_y = g.tiles_y; tp.push_rows = push_rows;
            tp.table_bytes = s->table_bytes; tp.pool_offset = s->pool_offset; tp.tile_offset = s->tile_offset; tp.tile_stride = 256 * g.PH;
            tp.lut = s->d_lut; tp.pool = s->d_pool;
            tp.spin_limit = 4000000u;                                   // ~2 s of polling before a tile gives up on a flag
            void* targs[] = {&tp};
            const long long items = (long long)nblk * g.tiles_x * g.tiles_y;
            const int grid = (int)std::min<long long>((long long)s->tile_grid, (items + 1) / 2);
            const unsigned smem_bytes = (unsigned)(s->tile_offset + 2 * 256 * g.PH);
            if (s->coop) {
                // A cooperative launch can be refused (CUDA_ERROR_COOPERATIVE_LAUNCH_TOO_LARGE under an MPS partition that holds
                // fewer CTAs than the device).  The same grid launched normally could deadlock on its own flags, so from then
                // on every launch carries ONE T-block: its tiles wait on no flag of the same grid.
                const CUresult cr = driver().LaunchCooperativeKernel(s->f_tiles, (unsigned)grid, 1, 1, 1024, 1, 1, smem_bytes, (CUstream)s->stream, targs);
                if (cr != CUDA_SUCCESS) {
                    s->coop = false;
                    (void)cudaGetLastError();
                    continue;                                           // plan this run again, one T-block per launch
                }
                s->launches++;
            } else {
                int rc = launch(s, s->f_tiles, dim3(grid), dim3(1024), targs, smem_bytes, false);
                if (rc) return rc;
            }
            s->frame += (int)run;
            s->tile_seq += (unsigned)nblk;
            s->tile_queue += (unsigned)items + 2u * (unsigned)grid;      // every half draws exactly one number past the end
            s->cur ^= (nblk & 1);
            // Every neighbour pushed at least HY rows (its boundary tile rows are at least that high).  HY and not the
            // actual count: the figure must be the same on every strip of the grid, because it decides when the
            // stream-ordered exchange of the per-step kernels happens, and that exchange runs in lock-step epochs.
            if (s->is_strip()) s->ghost_valid = s->has_neighbours() ? s->HY : 0;
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
    s->ghost_valid = 0;                                   // the neighbours' copies of my rows and mine of theirs are stale now
    SE_CUDA(cudaMemcpyAsync(s->cells[s->cur] + s->owned_offset(), host, s->owned_cells() * sizeof(unsigned), cudaMemcpyHostToDevice, s->stream));
    SE_CUDA(cudaStreamSynchronize(s->stream));
    return SE_OK;
} SE_ABI_CATCH("se_sim_upload_cells")

int se_sim_download_cells(se_sim* s, uint32_t* host) try {
    if (!s || !host) return fail(SE_ERR_INVALID_ARG, "null argument");
    SE_CUDA(cudaSetDevice(s->device));
    SE_CUDA(cudaMemcpyAsync(host, s->cells[s->cur] + s->owned_offset(), s->owned_cells() * sizeof(unsigned), cudaMemcpyDeviceToHost, s->stream));
    SE_CUDA(cudaStreamSynchronize(s->stream));
    return check_status(s);
} SE_ABI_CATCH("se_sim_download_cells")

int se_sim_upload_light(se_sim* s, const float* host) try {
    if (!s || !host) return fail(SE_ERR_INVALID_ARG, "null argument");
    if (!s->lighting) return fail(SE_ERR_INVALID_ARG, "sim was created without SE_FLAG_LIGHTING");
    SE_CUDA(cudaSetDevice(s->device));
    s->ghost_valid = 0;
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

int se_sim_checksum(se_sim* s, uint64_t* sum) try {
    if (!s || !sum) return fail(SE_ERR_INVALID_ARG, "null argument");
    SE_CUDA(cudaSetDevice(s->device));
    SE_CUDA(cudaMemsetAsync(s->d_census, 0, sizeof(unsigned long long), s->stream));
    se_static::launch_checksum(s->cells[s->cur] + s->owned_offset(), s->owned_cells(), (unsigned long long)s->row_begin * (unsigned long long)s->W, s->d_census, s->stream);
    s->launches++;
    SE_CUDA(cudaGetLastError());
    unsigned long long v = 0;
    SE_CUDA(cudaMemcpyAsync(&v, s->d_census, sizeof v, cudaMemcpyDeviceToHost, s->stream));
    SE_CUDA(cudaStreamSynchronize(s->stream));
    *sum = v;
    return check_status(s);
} SE_ABI_CATCH("se_sim_checksum")

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
    return check_status(s);
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
    s->ghost_valid = s->ghost_rows();   // every strip pushes in this protocol: the neighbours refresh this strip's ghost rows alike
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
    return halo_exchange_async(s);
} SE_ABI_CATCH("se_sim_halo_exchange_async")

}  // extern "C"

namespace {
int halo_exchange_async(se_sim* s) {
    SE_CUDA(cudaSetDevice(s->device));
    if (!driver().ok) return fail(SE_ERR_CUDA, driver().why);
    const unsigned e = ++s->epoch;
    CUstream st = (CUstream)s->stream;
    // (a) tell both neighbours that this strip has finished every step enqueued so far
    for (int w = 0; w < 2; ++w)
        if (s->nb[w].attached) SE_CU(driver().StreamWriteValue32(st, (CUdeviceptr)(s->nb[w].flags + FLAG_EPOCH + (w == 0 ? 1 : 0)), e, CU_STREAM_WRITE_VALUE_DEFAULT));
    // (b) wait until the neighbours have finished theirs: only then may their ghost rows be overwritten
    for (int w = 0; w < 2; ++w)
        if (s->nb[w].attached) SE_CU(driver().StreamWaitValue32(st, (CUdeviceptr)(s->flags + FLAG_EPOCH + (w == 0 ? 0 : 1)), e, CU_STREAM_WAIT_VALUE_GEQ));
    // (c) push boundary rows over NVLink
    int rc = se_sim_halo_push(s);
    if (rc) return rc;
    // (d) publish "delivered"
    for (int w = 0; w < 2; ++w)
        if (s->nb[w].attached) SE_CU(driver().StreamWriteValue32(st, (CUdeviceptr)(s->nb[w].flags + FLAG_EPOCH + (w == 0 ? 3 : 2)), e, CU_STREAM_WRITE_VALUE_DEFAULT));
    // (e) later work of this stream starts once this strip's own ghost rows have been delivered
    for (int w = 0; w < 2; ++w)
        if (s->nb[w].attached) SE_CU(driver().StreamWaitValue32(st, (CUdeviceptr)(s->flags + FLAG_EPOCH + (w == 0 ? 2 : 3)), e, CU_STREAM_WAIT_VALUE_GEQ));
    s->ghost_valid = s->ghost_rows();
    return SE_OK;
}
}  // namespace

