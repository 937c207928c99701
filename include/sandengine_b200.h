/* sandengine_b200 -- C ABI of the B200-native falling-sand simulation core.
 *
 * This is the drop-in boundary for the ONE hot path of ARez2/sandengine: the Margolus 2x2 block update
 * generated from materials.yaml, with its two riders (modification override, flood-fill lighting).
 * The reference has no FFI; its seam is two in-process Rust APIs.  Each entry point below names the
 * reference interface it replaces (file:line under /root/reference).  A Rust `extern "C"` binding
 * and the `Simulation` shim that would sit on top of it are shown in INTEGRATION.md.
 *
 * Conventions: every function returns 0 (SE_OK) or a negative se_status; no exceptions or panics
 * cross the ABI; se_last_error() returns a thread-local message for the last failure.  Plain pointers
 * and sizes only.  A se_sim is single-owner and not thread-safe (mirrors the GL-context rule of the
 * reference, simulation.rs:97-126); distinct sims are independent.  se_sim_step() is asynchronous on
 * the sim's stream; uploads/downloads synchronise that stream.
 *
 * There is NO CPU fallback: without a CUDA device se_sim_create() fails with SE_ERR_CUDA.
 * se_rules_compile_yaml() needs only NVRTC (no device) and so also works on a build host.
 */
#ifndef SANDENGINE_B200_H
#define SANDENGINE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum se_status {
    SE_OK = 0,
    SE_ERR_YAML = -1,            /* serde_yaml error surfaced by parse_string, parser.rs:95-98 */
    SE_ERR_MISSING_FIELD = -2,   /* ParsingErr::MissingField   parser.rs:46-50 */
    SE_ERR_INVALID_TYPE = -3,    /* ParsingErr::InvalidType    parser.rs:53-58 */
    SE_ERR_NOT_FOUND = -4,       /* ParsingErr::NotFound       parser.rs:61-65 */
    SE_ERR_NOT_RECOGNIZED = -5,  /* ParsingErr::NotRecognized  parser.rs:68-72 */
    SE_ERR_UNSUPPORTED = -6,     /* valid in the reference's grammar but outside this build's limits */
    SE_ERR_COMPILE = -7,         /* NVRTC failure (the reference panics on a GLSL compile error, simulation.rs:133-137) */
    SE_ERR_CUDA = -8,
    SE_ERR_INVALID_ARG = -9,
    SE_ERR_INTERNAL = -10        /* host allocation failure or an unexpected C++ exception: caught at the boundary, never thrown across it */
} se_status;

typedef struct se_rules se_rules;   /* parsed rule set + generated CUDA C + sm_100a cubin */
typedef struct se_sim se_sim;       /* opaque; owns device buffers of one grid (or one strip of it) */

/* == simulation.rs:45-56 `SimModification` (std140 stride 32 B of the shader's UBO, falling_sand.glsl:43-55) */
typedef struct se_modification {
    int32_t position[2];
    int32_t mod_shape;   /* SE_MODSHAPE_* */
    int32_t mod_size;
    int32_t mod_matID;
    int32_t _pad4[3];
} se_modification;
#define SE_MODSHAPE_CIRCLE 0     /* simulation.rs:41 */
#define SE_MODSHAPE_SQUARE 1     /* simulation.rs:42 */
#define SE_MAX_MODIFICATIONS 256 /* simulation.rs:43 */

#define SE_FLAG_LIGHTING 1u      /* evaluate the lighting relaxation every step (operations.glsl:114-169) */
/* Keep the per-material census of the owned rows up to date inside the per-frame step kernel (population deltas of the
 * blocks where a SET fired) so that se_sim_census[_async] after a se_sim_step(sim, 1) needs no pass over the grid.
 * Results are identical to the recount; only rule sets with a single shared-memory transition table and lighting off
 * use it, elsewhere the flag is ignored. */
#define SE_FLAG_RUNNING_CENSUS 2u
/* Accepted for compatibility (round 1 gated lighting on a strip behind it): lit strips are always allowed now.  The
 * light field gets ghost rows too: attach the neighbours' light buffers (se_sim_ipc_export_light / _attach_light, or
 * se_sim_attach_local); the light stencil uses up one ghost row per step and se_sim_step exchanges when they run out. */
#define SE_FLAG_LIT_STRIP_EXPERIMENTAL 4u
/* Accepted for compatibility (round 1 gated the one-kernel lit path behind it): with SE_FLAG_LIGHTING, a table-eligible
 * rule set and a width that is a multiple of 8, the Margolus step, the modification override and the lighting relaxation
 * always run as ONE kernel (se_step_lit); elsewhere as two.  Same results either way. */
#define SE_FLAG_FUSED_LIGHT_EXPERIMENTAL 8u

typedef struct se_create_params {
    uint32_t width;          /* simSize.x */
    uint32_t height;         /* simSize.y of the WHOLE grid */
    uint32_t flags;          /* SE_FLAG_* */
    int32_t device;          /* CUDA device ordinal */
    /* Strip decomposition (multi-GPU): this sim owns global rows [row_begin, row_end) and keeps
     * `halo_rows` ghost rows towards each neighbour.  Single GPU: row_begin = 0, row_end = 0 (= height),
     * halo_rows = 0.  row_begin/row_end and halo_rows must be even (Margolus blocks are 2 rows).  With neighbours
     * attached se_sim_step keeps the ghost rows current by itself: the tile kernel pushes its boundary rows into the
     * neighbours' ghost rows as part of its store phase, the per-step kernels exchange on the stream when the ghost rows
     * are used up.  Every strip of a grid must be given the same sequence of se_sim_step calls. */
    uint32_t row_begin, row_end, halo_rows;
    uint32_t temporal_block; /* Margolus steps fused per launch by the tiled kernel; 0 = library default */
    /* Number of sims that share this CUDA device AND are stepped concurrently (several strips of one grid on one GPU,
     * the single-process test set-up): the persistent tile kernel then takes 1/device_share of the SMs so that the
     * strips' kernels, which wait for each other's boundary tiles, are co-resident.  0 or 1: the device is this sim's. */
    uint32_t device_share;
} se_create_params;

/* ---- codegen seam: replaces sandengine_lang::parse_string + create_glsl_from_parser ------------
 * (parser.rs:93, sandengine-lang/src/lib.rs:17).  Parses the YAML rule language, generates CUDA C and
 * compiles it for sm_100a.  Parser errors map 1:1 onto the four ParsingErr classes. */
int se_rules_compile_yaml(const char* yaml, size_t len, se_rules** out);
int se_rules_parse_only(const char* yaml, size_t len, se_rules** out);   /* front end only: no NVRTC */
int se_rules_destroy(se_rules* r);
/* Known-answer text: what the reference's emitter writes to gen/materials.glsl (which=0) and
 * gen/rules.glsl (which=1); which=2: the generated CUDA header; which=3: NVRTC log. Pointer valid until destroy. */
int se_rules_text(const se_rules* r, int which, const char** text, size_t* len);
int se_rules_cubin(const se_rules* r, const void** data, size_t* len);
int se_rules_counts(const se_rules* r, int32_t* n_rules, int32_t* n_types, int32_t* n_materials);
/* ParsingResult.materials[i]: id == index (materials.rs:91,196). color/emission: 4 floats each. */
int se_rules_material(const se_rules* r, int32_t id, const char** name, const char** type_name, float* density,
                      float* color4, float* emission4, int32_t* selectable);
int se_rules_material_id(const se_rules* r, const char* name, int32_t* id);
/* ParsingResult.rules[i] (rules.rs:25-44): kind 0 Mirrored / 1 Left / 2 Right as this build executes it. */
int se_rules_rule(const se_rules* r, int32_t index, const char** name, int32_t* used, int32_t* kind, const char** precondition);

/* ---- simulation seam: replaces sandengine_core::simulation::Simulation ---------------------- */
int se_sim_create(const se_rules* rules, const se_create_params* params, se_sim** out);   /* Simulation::new, simulation.rs:128-192 */
int se_sim_destroy(se_sim* s);
/* One call == n_steps calls of Simulation::run (simulation.rs:195-253): frame += 1 before each step;
 * pending modifications are consumed by the FIRST step only and then cleared (:246-252). */
int se_sim_step(se_sim* s, uint32_t n_steps);
/* sim.modifications.push(..) (sandengine-core/src/lib.rs:59-67).  Appends; only the first 256 pending
 * entries are used by the next step (simulation.rs:205), like the reference.  mod_size values above 2^30 are applied as
 * 2^30 (same cells as the shader's test for any position within +-2^29 of the grid: all of them). */
int se_sim_push_modifications(se_sim* s, const se_modification* mods, uint32_t n);
int se_sim_set_frame(se_sim* s, int32_t frame);     /* params.frame, simulation.rs:78 */
int se_sim_get_frame(const se_sim* s, int32_t* frame);
/* Cell state (the reference's input_data texture, .r channel): packed uint32 material ids, row-major,
 * y down, OWNED rows only: width * (row_end - row_begin) entries. */
int se_sim_upload_cells(se_sim* s, const uint32_t* host_cells);
int se_sim_download_cells(se_sim* s, uint32_t* host_cells);
/* Light state (input_light texture): 4 floats per cell, owned rows.  Needs SE_FLAG_LIGHTING. */
int se_sim_upload_light(se_sim* s, const float* host_rgba);
int se_sim_download_light(se_sim* s, float* host_rgba);
/* Colour shading of the owned rows on the device (the reference's output_color texture, operations.glsl:100-108,170:
 * material colour minus simplex noise of the position, clamped).  host_rgba_f32: width*rows*4 floats, or NULL;
 * host_rgba8: width*rows packed R|G<<8|B<<16|A<<24 bytes for headless readback, or NULL. */
int se_sim_download_color(se_sim* s, float* host_rgba_f32, uint32_t* host_rgba8);
/* Device pointer of the current cell buffer's first OWNED row (for CUDA-GL interop / zero-copy readers);
 * pitch in bytes.  Valid until the next se_sim_step. */
int se_sim_device_cells(se_sim* s, void** dptr, size_t* pitch_bytes);
/* Per-material population of the owned rows (256 bins), computed on the device. */
int se_sim_census(se_sim* s, uint64_t* counts256);
/* Checksum of the owned rows: sum over cells of mix(global cell index, id) modulo 2^64 (mix: see static_kernels.cu).
 * Order- and sharding-independent: the values of the strips of a grid add up (mod 2^64) to the value of the whole grid,
 * which is how a sharded run is compared with an unsharded one without gathering the grid. */
int se_sim_checksum(se_sim* s, uint64_t* sum);
/* Asynchronous census: enqueued after everything issued so far, runs on the sim's side stream CONCURRENTLY with
 * later steps (a later step that would overwrite the buffer being counted waits for it on the device).
 * host_counts256 (256 x uint64; pinned memory for a truly asynchronous copy) is valid after
 * se_sim_census_wait().  Up to 4 censuses may be in flight. */
int se_sim_census_async(se_sim* s, uint64_t* host_counts256);
int se_sim_census_wait(se_sim* s);
/* Run on a caller-provided CUDA stream (cudaStream_t as void*); NULL restores the sim's own stream. */
int se_sim_set_stream(se_sim* s, void* cuda_stream);
int se_sim_synchronize(se_sim* s);
/* Number of kernel launches issued by this sim so far (for the bench's gpu_launches claim). */
int se_sim_launch_count(const se_sim* s, uint64_t* n);

/* ---- strips: ghost-row exchange between neighbouring sims ---------------------- */
/* Export a CUDA IPC handle (64 bytes) for each of this sim's two cell buffers so that a neighbour
 * process can map them; buffer geometry is returned alongside. */
int se_sim_ipc_export(se_sim* s, void* handles_2x64, uint64_t* local_rows, uint64_t* ghost_top, uint64_t* ghost_bottom);
/* Attach a neighbour living in ANOTHER process (which: 0 = the strip above, 1 = the strip below) from its
 * exported handles. */
int se_sim_ipc_attach(se_sim* s, int which, const void* handles_2x64, uint64_t nb_local_rows,
                      uint64_t nb_ghost_top, uint64_t nb_ghost_bottom);
/* Same pair for the two light buffers of a lit strip; call after se_sim_ipc_attach. */
int se_sim_ipc_export_light(se_sim* s, void* handles_2x64);
int se_sim_ipc_attach_light(se_sim* s, int which, const void* handles_2x64);
/* Attach a neighbour that lives in THIS process (one process driving several strips / devices, the
 * reference's single-process model); peer access is enabled when the devices differ. */
int se_sim_attach_local(se_sim* s, int which, se_sim* neighbour);
/* Push this sim's boundary rows into the attached neighbours' ghost rows (device-to-device over NVLink).
 * The caller is responsible for ordering against the neighbours' kernels (host barriers).  se_sim_step does its own
 * exchanges; these two entry points remain for callers that want to drive the exchange themselves. */
int se_sim_halo_push(se_sim* s);
/* The same exchange with the ordering done ON THE DEVICE, no host synchronisation: stream-ordered flag
 * writes / waits on peer-mapped words (cuStreamWriteValue32 / cuStreamWaitValue32):
 *   signal "done computing" -> wait for the neighbours' "done" -> push rows -> signal "delivered" ->
 *   make this stream wait for the neighbours' "delivered".
 * Every strip must call it the same number of times (lock-step epochs). */
int se_sim_halo_exchange_async(se_sim* s);

const char* se_last_error(void);
const char* se_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SANDENGINE_B200_H */
