// C++17 host-side mirror of the reference's two Rust APIs, header-only, on top of the C ABI (sandengine_b200.h).
//
// The reference is compiled code (Rust) and no Rust toolchain exists in this image, so the host side above the ABI is
// C++ (here) and Python (sandengine_b200/*.py); bindings/rust/ holds the same shim as Rust source.  Names, argument
// meaning and error behaviour follow the reference so that code written against it reads the same:
//
//   sandengine_lang::parse_string / parse_path            sandengine-lang/src/parser.rs:93, src/lib.rs:10
//   sandengine_lang::ParsingResult {rules, types, materials}                 parser.rs:84-89
//   sandengine_lang::ParsingErr::{MissingField, InvalidType, NotFound, NotRecognized}   parser.rs:42-73
//   sandengine_lang::create_cuda_from_parser              replaces create_glsl_from_parser, src/lib.rs:17
//   sandengine_core::Simulation::{new_, run}, .params, .modifications        sandengine-core/src/simulation.rs:97-253
//   sandengine_core::Params, SimModification, MODSHAPE_*, MAX_MODIFICATIONS  simulation.rs:41-92
//
// Differences that follow from the boundary: `Simulation::new_` takes the compiled rule set instead of a glium display
// (the GL context is gone), a failure that makes the reference panic (shader compile, allocation: simulation.rs:133-138)
// throws `SandEngineError`, and the textures `output_color` / `output_light` are read with download_*().
#pragma once

#include <cstdint>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "sandengine_b200.h"

namespace sandengine_b200 {

// Every non-zero se_status.  what() is se_last_error(): for parser errors the reference's text, which contains the
// ParsingErr class name ("(MissingField) Mandatory field 'rules' is missing in ...", parser.rs:44-72), so that
// `err.to_string().contains("MissingField")` of the reference's tests carries over as what() / kind().
class SandEngineError : public std::runtime_error {
public:
    SandEngineError(int status, const std::string& msg) : std::runtime_error(msg), status_(status) {}
    int status() const { return status_; }
    const char* kind() const {
        switch (status_) {
            case SE_ERR_YAML: return "Yaml";
            case SE_ERR_MISSING_FIELD: return "MissingField";
            case SE_ERR_INVALID_TYPE: return "InvalidType";
            case SE_ERR_NOT_FOUND: return "NotFound";
            case SE_ERR_NOT_RECOGNIZED: return "NotRecognized";
            case SE_ERR_UNSUPPORTED: return "Unsupported";
            case SE_ERR_COMPILE: return "Compile";
            case SE_ERR_CUDA: return "Cuda";
            case SE_ERR_INVALID_ARG: return "InvalidArg";
            case SE_ERR_INTERNAL: return "Internal";
        }
        return "Unknown";
    }

private:
    int status_;
};

inline void check(int status) {
    if (status != SE_OK) throw SandEngineError(status, se_last_error());
}

}  // namespace sandengine_b200

namespace sandengine_lang {

using ParsingErr = sandengine_b200::SandEngineError;

struct SandMaterial {   // parser/materials.rs:10-29
    int32_t id = 0;
    std::string name, type;
    float color[4] = {0, 0, 0, 0}, emission[4] = {0, 0, 0, 0};
    bool selectable = true;
    float density = 0.0f;
};

struct SandRule {       // parser/rules.rs:25-44 (what the ABI exposes of it)
    std::string name;
    bool used = false;
    enum class Type { Mirrored = 0, Left = 1, Right = 2 } ruletype = Type::Mirrored;   // rules.rs:14-19
    bool has_precondition = false;
    std::string precondition;
};

// parser.rs:84-89.  Owns the native rule set (parsed text, generated CUDA C, sm_100a cubin when compiled).
class ParsingResult {
public:
    std::vector<SandRule> rules;
    int32_t n_types = 0;
    std::vector<SandMaterial> materials;

    ParsingResult() = default;
    ParsingResult(se_rules* h, bool compiled) : h_(h), compiled_(compiled) { load(); }
    ParsingResult(ParsingResult&& o) noexcept { *this = std::move(o); }
    ParsingResult& operator=(ParsingResult&& o) noexcept {
        if (this != &o) {
            reset();
            rules = std::move(o.rules); n_types = o.n_types; materials = std::move(o.materials);
            h_ = o.h_; compiled_ = o.compiled_;
            o.h_ = nullptr;
        }
        return *this;
    }
    ParsingResult(const ParsingResult&) = delete;
    ParsingResult& operator=(const ParsingResult&) = delete;
    ~ParsingResult() { reset(); }

    const se_rules* handle() const { return h_; }
    bool compiled() const { return compiled_; }
    // known-answer text: what the reference's emitter writes to gen/materials.glsl and gen/rules.glsl (lib.rs:17-148)
    std::string glsl_materials() const { return text(0); }
    std::string glsl_rules() const { return text(1); }
    std::string cuda_header() const { return text(2); }      // the generated CUDA C that replaces them
    std::string nvrtc_log() const { return text(3); }
    int32_t material_id(const std::string& name) const {
        int32_t id = 0;
        sandengine_b200::check(se_rules_material_id(h_, name.c_str(), &id));
        return id;
    }
    std::vector<SandMaterial> selectable_materials() const {  // sandengine-core/src/lib.rs:21-27
        std::vector<SandMaterial> out;
        for (const auto& m : materials)
            if (m.selectable) out.push_back(m);
        return out;
    }

private:
    se_rules* h_ = nullptr;
    bool compiled_ = false;

    void reset() {
        if (h_) se_rules_destroy(h_);
        h_ = nullptr;
    }
    std::string text(int which) const {
        const char* p = nullptr;
        size_t n = 0;
        sandengine_b200::check(se_rules_text(h_, which, &p, &n));
        return std::string(p ? p : "", n);
    }
    void load() {
        int32_t nr = 0, nt = 0, nm = 0;
        sandengine_b200::check(se_rules_counts(h_, &nr, &nt, &nm));
        n_types = nt;
        for (int32_t i = 0; i < nm; ++i) {
            SandMaterial m;
            const char *name = nullptr, *type = nullptr;
            int32_t sel = 0;
            sandengine_b200::check(se_rules_material(h_, i, &name, &type, &m.density, m.color, m.emission, &sel));
            m.id = i; m.name = name ? name : ""; m.type = type ? type : ""; m.selectable = sel != 0;
            materials.push_back(m);
        }
        for (int32_t i = 0; i < nr; ++i) {
            SandRule r;
            const char *name = nullptr, *pre = nullptr;
            int32_t used = 0, kind = 0;
            sandengine_b200::check(se_rules_rule(h_, i, &name, &used, &kind, &pre));
            r.name = name ? name : ""; r.used = used != 0; r.ruletype = static_cast<SandRule::Type>(kind);
            r.has_precondition = pre != nullptr; r.precondition = pre ? pre : "";
            rules.push_back(r);
        }
    }
};

// parser.rs:93.  compile = true also generates CUDA C and compiles it for sm_100a (NVRTC; works without a GPU) --
// the job `create_glsl_from_parser` + the GL driver's shader compiler do in the reference.
inline ParsingResult parse_string(const std::string& text, bool compile = true) {
    se_rules* h = nullptr;
    sandengine_b200::check(compile ? se_rules_compile_yaml(text.data(), text.size(), &h) : se_rules_parse_only(text.data(), text.size(), &h));
    return ParsingResult(h, compile);
}

// sandengine-lang/src/lib.rs:10-13
inline ParsingResult parse_path(const std::string& path, bool compile = true) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot read " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return parse_string(ss.str(), compile);
}

// Replaces create_glsl_from_parser (lib.rs:17-148, which writes gen/materials.glsl + gen/rules.glsl under the CWD):
// writes the generated CUDA header next to the two known-answer GLSL files.
inline void create_cuda_from_parser(const ParsingResult& res, const std::string& out_dir) {
    auto put = [&](const char* name, const std::string& s) {
        std::ofstream f(out_dir + "/" + name, std::ios::binary);
        if (!f) throw std::runtime_error("cannot write " + out_dir + "/" + name);
        f << s;
    };
    put("rules_gen.cuh", res.cuda_header());
    put("materials.glsl", res.glsl_materials());
    put("rules.glsl", res.glsl_rules());
}

}  // namespace sandengine_lang

namespace sandengine_core {

constexpr int32_t MODSHAPE_CIRCLE = SE_MODSHAPE_CIRCLE;        // simulation.rs:41
constexpr int32_t MODSHAPE_SQUARE = SE_MODSHAPE_SQUARE;        // simulation.rs:42
constexpr size_t MAX_MODIFICATIONS = SE_MAX_MODIFICATIONS;     // simulation.rs:43

using SimModification = se_modification;                       // simulation.rs:45-56, same field names, 32 bytes
static_assert(sizeof(SimModification) == 32, "std140 stride of the shader's uniform block");

struct Params {                                                // simulation.rs:70-92
    bool moveRight = true;
    float mousePos[2] = {0.0f, 0.0f};
    bool mousePressed = false;
    uint32_t brushSize = 5;
    int32_t brushMaterial = 0;      // the reference keeps the SandMaterial; only its id reaches the shader
    float time = 0.0f;
    int32_t frame = 0;
};

class Simulation {                                             // simulation.rs:97-126
public:
    Params params;
    std::vector<SimModification> modifications;

    // Simulation::new(display, size), simulation.rs:128-192.  flags: SE_FLAG_LIGHTING ...
    static Simulation new_(const sandengine_lang::ParsingResult& rules, std::pair<uint32_t, uint32_t> size, uint32_t flags = 0, int32_t device = 0) {
        return Simulation(rules, size, flags, device);
    }
    Simulation(const sandengine_lang::ParsingResult& rules, std::pair<uint32_t, uint32_t> size, uint32_t flags = 0, int32_t device = 0)
        : size_(size) {
        se_create_params p{};
        p.width = size.first; p.height = size.second; p.flags = flags; p.device = device;
        sandengine_b200::check(se_sim_create(rules.handle(), &p, &h_));
    }
    Simulation(Simulation&& o) noexcept : params(o.params), modifications(std::move(o.modifications)), h_(o.h_), size_(o.size_) { o.h_ = nullptr; }
    Simulation(const Simulation&) = delete;
    Simulation& operator=(const Simulation&) = delete;
    ~Simulation() {
        if (h_) se_sim_destroy(h_);
    }

    // Simulation::run, simulation.rs:195-253: frame += 1, the first min(len, 256) modifications are applied by this
    // step only, then the list is cleared.
    void run() { step(1); }
    // n x run() in one call (the library fuses runs of plain steps on the device)
    void step(uint32_t n_steps) {
        sandengine_b200::check(se_sim_set_frame(h_, params.frame));
        if (!modifications.empty()) {
            sandengine_b200::check(se_sim_push_modifications(h_, modifications.data(), static_cast<uint32_t>(modifications.size())));
            modifications.clear();
        }
        sandengine_b200::check(se_sim_step(h_, n_steps));
        sandengine_b200::check(se_sim_get_frame(h_, &params.frame));
    }
    // sandengine-core/src/lib.rs:59-67: one CIRCLE of brushSize / brushMaterial at mousePos * size
    void push_brush() {
        SimModification m{};
        m.position[0] = static_cast<int32_t>(params.mousePos[0] * size_.first);
        m.position[1] = static_cast<int32_t>(params.mousePos[1] * size_.second);
        m.mod_shape = MODSHAPE_CIRCLE; m.mod_size = static_cast<int32_t>(params.brushSize); m.mod_matID = params.brushMaterial;
        modifications.push_back(m);
    }

    std::pair<uint32_t, uint32_t> size() const { return size_; }
    size_t cells() const { return static_cast<size_t>(size_.first) * size_.second; }
    void upload_cells(const std::vector<uint32_t>& ids) { expect(ids.size() == cells()); sandengine_b200::check(se_sim_upload_cells(h_, ids.data())); }
    std::vector<uint32_t> download_cells() {
        std::vector<uint32_t> out(cells());
        sandengine_b200::check(se_sim_download_cells(h_, out.data()));
        return out;
    }
    void upload_light(const std::vector<float>& rgba) { expect(rgba.size() == 4 * cells()); sandengine_b200::check(se_sim_upload_light(h_, rgba.data())); }
    std::vector<float> download_light() {                      // the reference's `output_light`
        std::vector<float> out(4 * cells());
        sandengine_b200::check(se_sim_download_light(h_, out.data()));
        return out;
    }
    std::vector<float> download_color() {                      // the reference's `output_color`
        std::vector<float> out(4 * cells());
        sandengine_b200::check(se_sim_download_color(h_, out.data(), nullptr));
        return out;
    }
    std::vector<uint64_t> census() {
        std::vector<uint64_t> out(256);
        sandengine_b200::check(se_sim_census(h_, out.data()));
        return out;
    }
    se_sim* handle() { return h_; }

private:
    se_sim* h_ = nullptr;
    std::pair<uint32_t, uint32_t> size_;
    static void expect(bool ok) {
        if (!ok) throw std::invalid_argument("buffer size does not match the grid");
    }
};

}  // namespace sandengine_core
