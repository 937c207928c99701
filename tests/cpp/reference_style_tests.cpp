// Tests of the C++ host-side mirror (include/sandengine_b200.hpp), written the way the reference tests its parser
// (/root/reference/tests/test_sandengine-lang.rs: one malformed document per ParsingErr class, checked through the
// error text) plus the simulation seam of sandengine-core/src/simulation.rs.  Own YAML text; plain asserts; exit code
// = number of failed checks.  Built and run by tests/test_cpp_mirror.py.
//   argv[1] = path of data/materials.yaml      argv[2] = "gpu" to also run the device part
#include <cstdio>
#include <cstring>
#include <string>

#include "sandengine_b200.hpp"

using sandengine_lang::parse_path;
using sandengine_lang::parse_string;
using sandengine_lang::ParsingErr;

static int g_failed = 0;
#define CHECK(cond)                                                              \
    do {                                                                         \
        if (!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); ++g_failed; } \
    } while (0)

// `assert!(res.err().unwrap().to_string().contains("<class>"))` of the reference's tests
static void expect_err(const char* what, const std::string& yaml, const char* klass, const char* fragment = nullptr) {
    try {
        parse_string(yaml, /*compile=*/false);
        std::printf("FAILED %s: no error\n", what);
        ++g_failed;
    } catch (const ParsingErr& e) {
        const bool ok = std::strstr(e.what(), klass) && std::string(e.kind()) == klass && (!fragment || std::strstr(e.what(), fragment));
        if (!ok) { std::printf("FAILED %s: got (%s) %s\n", what, e.kind(), e.what()); ++g_failed; }
    }
}

static const char* RULES =
    "rules:\n"
    "  sink:\n"
    "    if: DOWN.mat.density < SELF.mat.density\n"
    "    do: SWAP SELF DOWN\n"
    "  slip:\n"
    "    if: DOWNRIGHT.mat.density < SELF.mat.density\n"
    "    do: SWAP SELF DOWNRIGHT\n"
    "    probability: 0.25\n";
static const char* TYPES =
    "types:\n"
    "  grain:\n"
    "    base_rules: [sink, slip]\n";
static const char* MATERIALS =
    "materials:\n"
    "  grit:\n"
    "    type: grain\n"
    "    color: [200, 180, 90]\n"
    "    density: 1.7\n";

static void parser_tests() {
    // the three sections are mandatory (parser.rs:121-147)
    expect_err("missing_rules", std::string(TYPES) + MATERIALS, "MissingField", "'rules'");
    expect_err("missing_types", std::string(RULES) + MATERIALS, "MissingField", "'types'");
    expect_err("missing_materials", std::string(RULES) + TYPES, "MissingField", "'materials'");
    // a rule name must be a string (parser.rs:156-169)
    expect_err("invalid_name", std::string("rules:\n  1.0:\n    if: DOWN.mat.density < SELF.mat.density\n    do: SWAP SELF DOWN\n") + TYPES + MATERIALS, "InvalidType");
    // `if` is mandatory at the top level of a rule (rules.rs:220-224); `color` in a material (materials.rs)
    expect_err("missing_if", std::string("rules:\n  sink:\n    do: SWAP SELF DOWN\n  slip:\n    if: SELF.mat == EMPTY\n    do: SWAP SELF DOWN\n") + TYPES + MATERIALS,
               "MissingField", "'if'");
    expect_err("missing_color", std::string(RULES) + TYPES + "materials:\n  grit:\n    type: grain\n    density: 1.7\n", "MissingField", "'color'");
    // a base rule / a material type / a compared material that does not exist (types.rs:121-147, materials.rs, rules.rs:246-260)
    expect_err("unknown_base_rule", std::string(RULES) + "types:\n  grain:\n    base_rules: [sink, hover]\n" + MATERIALS, "NotFound", "hover");
    expect_err("unknown_type", std::string(RULES) + TYPES + "materials:\n  grit:\n    type: fluid\n    color: [1, 2, 3]\n    density: 1.0\n", "NotFound", "fluid");
    expect_err("unknown_material_in_condition",
               std::string("rules:\n  sink:\n    if: DOWN.mat == lava\n    do: SWAP SELF DOWN\n  slip:\n    if: SELF.mat == EMPTY\n    do: SWAP SELF DOWN\n") + TYPES + MATERIALS,
               "NotFound", "lava");
    // an action that is neither SWAP nor SET (rules.rs:355-420)
    expect_err("unknown_action", std::string("rules:\n  sink:\n    if: SELF.mat == EMPTY\n    do: MOVE SELF DOWN\n  slip:\n    if: SELF.mat == EMPTY\n    do: SWAP SELF DOWN\n") + TYPES + MATERIALS,
               "NotRecognized");

    // a well-formed document
    auto res = parse_string(std::string(RULES) + TYPES + MATERIALS, /*compile=*/false);
    CHECK(res.rules.size() == 2 && res.rules[0].name == "sink" && res.rules[0].used && res.rules[1].used);
    CHECK(res.n_types == 4);                                                   // EMPTY, NULL, WALL + grain (types.rs:53-69)
    CHECK(res.materials.size() == 4 && res.materials[3].name == "grit" && res.materials[3].id == 3 && res.materials[3].type == "grain");
    CHECK(res.materials[0].name == "EMPTY" && res.materials[1].name == "NULL" && res.materials[2].name == "WALL");   // materials.rs:52-83
    CHECK(res.materials[3].color[0] == 200.0f / 255.0f && res.materials[3].color[3] == 1.0f);                      // parser.rs:219-235
    CHECK(res.material_id("grit") == 3);
    CHECK(res.glsl_rules().find("void rule_sink") != std::string::npos && res.glsl_rules().find("rand.y <= 0.25") != std::string::npos);
}

static void default_rule_set(const char* path) {
    auto res = parse_path(path);                                               // + CUDA codegen + NVRTC for sm_100a, no GPU needed
    CHECK(res.compiled());
    CHECK(res.materials.size() == 11 && res.materials[3].name == "sand" && res.materials[10].name == "dirt");      // gen/materials.glsl:50-60
    CHECK(res.cuda_header().find("se_apply_mirrored") != std::string::npos);
    CHECK(res.selectable_materials().size() >= 8);
    int used = 0;
    for (const auto& r : res.rules) used += r.used;
    CHECK(used == 8);                                                          // gen/rules.glsl:102-113
}

static void simulation_without_a_device(const char* path) {
    auto res = parse_path(path);
    try {
        auto sim = sandengine_core::Simulation::new_(res, {64, 64});
        std::printf("FAILED: Simulation::new_ succeeded without a device\n");
        ++g_failed;
    } catch (const sandengine_b200::SandEngineError& e) {
        CHECK(std::string(e.kind()) == "Cuda" && std::strstr(e.what(), "no CPU fallback"));
    }
}

static void simulation_on_the_device(const char* path) {
    using namespace sandengine_core;
    auto res = parse_path(path);
    // the survey's 16 x 16 state KAT: g[y][x] = pal[(7x + 13y + (x y mod 5)) mod 8], 40 steps from frame 1
    const uint32_t pal[8] = {0, 3, 5, 7, 4, 10, 0, 0};
    std::vector<uint32_t> g(16 * 16);
    for (int y = 0; y < 16; ++y)
        for (int x = 0; x < 16; ++x) g[y * 16 + x] = pal[(7 * x + 13 * y + ((x * y) % 5)) % 8];
    auto sim = Simulation::new_(res, {16, 16});
    sim.upload_cells(g);
    sim.params.frame = 1;
    for (int i = 0; i < 40; ++i) sim.run();
    CHECK(sim.params.frame == 41);
    const uint64_t want[11] = {99, 0, 0, 32, 31, 30, 0, 26, 0, 1, 37};
    auto hist = sim.census();
    for (int i = 0; i < 11; ++i) CHECK(hist[i] == want[i]);
    // a brush stamp is consumed by exactly one run() (simulation.rs:246-252)
    sim.params.mousePos[0] = 0.5f; sim.params.mousePos[1] = 0.5f; sim.params.brushSize = 2; sim.params.brushMaterial = res.material_id("rock");
    sim.push_brush();
    CHECK(sim.modifications.size() == 1 && sim.modifications[0].position[0] == 8);
    sim.run();
    CHECK(sim.modifications.empty());
    CHECK(sim.download_cells()[8 * 16 + 8] == 4u);
    // frame 1 clears the grid (falling_sand.glsl:743-746)
    sim.params.frame = 0;
    sim.run();
    auto cleared = sim.download_cells();
    bool all_empty = true;
    for (uint32_t v : cleared) all_empty = all_empty && v == 0;
    CHECK(all_empty);
}

int main(int argc, char** argv) {
    if (argc < 2) { std::printf("usage: %s data/materials.yaml [gpu]\n", argv[0]); return 99; }
    parser_tests();
    default_rule_set(argv[1]);
    if (argc > 2 && std::string(argv[2]) == "gpu") simulation_on_the_device(argv[1]);
    else simulation_without_a_device(argv[1]);
    std::printf("%s (%d failed)\n", g_failed ? "FAILED" : "ok", g_failed);
    return g_failed;
}
