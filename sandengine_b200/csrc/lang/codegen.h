// CUDA C back end of the rule language: the B200 build's replacement for the reference's GLSL
// emitter (`create_glsl_from_parser`, /root/reference/sandengine-lang/src/lib.rs:17-148).
//
// Input: a ParsingResult whose rule text is still GLSL-shaped (exactly what the reference would have
// written into gen/rules.glsl).  That text is parsed into a typed expression tree and re-emitted as
// CUDA C over *fat cells* -- one 32-bit register per cell:
//     bits  0..7   material id            bits  8..15  type id
//     bits 16..23  flags (SE_F_*)         bits 24..31  density rank (dense rank of the f32 density)
// so that `a.mat.density < b.mat.density` is one integer compare, `isType_T(c)` is a bit test against
// a compile-time mask (T plus all transitive children, types.rs:186-200) and `swap` is a predicated
// register exchange (guard = SE_F_NOSWAP, operations.glsl:16-23).
// `rand.y <= p` is lowered to an integer compare on the raw hash lane: uint->float conversion is
// monotone, so {u : float_rn(u)/2^32 <= p} is a prefix [0, U]; U is found here by bisection with the
// very float expression the shader evaluates (math.glsl:74-79), see rand_threshold().
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "lang.h"

namespace se {

enum : uint32_t {
    SE_F_NOSWAP = 1u << 16,      // type is WALL or NULL: swap() refuses (operations.glsl:16-23)
    SE_F_EMISSIVE = 1u << 17,    // emission.rgb != 0 (operations.glsl:126)
    SE_F_OBSTACLE = 1u << 18,    // isLightObstacle: emission.rgb == 0 && type != EMPTY (operations.glsl:68-70)
};

struct MaterialTables {
    int n_materials = 0;
    int n_types = 0;
    uint32_t fat[256];          // fat cell for every possible id byte; unknown ids -> NULL's fat cell
    float density[256];
    float emission[256][4];
    float color[256][4];
    uint8_t selectable[256];
    std::vector<std::string> material_names;
    std::vector<std::string> type_names;
    std::vector<uint64_t> type_masks;   // per type id: bit t set for the type itself and every transitive child
};

struct CompiledRules {
    ParsingResult parsed;
    MaterialTables tables;
    std::string cuda_header;     // the generated "rules_gen.cuh" (tables + rule functions + callers)
    uint32_t rand_lanes = 1;     // bit k: hash lane k (x,y,z,w) is consumed by the rule set (bit 0: mirror)
    bool have_left = false, have_right = false;
    int n_used_rules = 0;
    // Transition-table (LUT) mode: possible when the block transition is a pure function of
    // (4 material ids, mirror bit, which of the sorted rand.y thresholds the hash lane clears):
    // no pos / frame / other rand lanes / generic float use of rand, and few enough materials.
    bool lut_eligible = false;
    int lut_mode = 0;            // 0: generated code only; 1: table in shared memory; 2: table in global memory (up to 127 materials)
    int lut_tables = 1;          // 2: separate tables for the unmirrored and the mirrored view (rule sets with Left/Right rules, and always in mode 2)
    std::vector<uint32_t> lut_thresholds;   // sorted distinct U: condition i is `u1 <= U_i`
};

// Largest u in [0, 2^32) with  float_rn(u) / 2^32 <op> p  (op is "<=" or "<"); returns false when no u
// satisfies it; *all is set when every u does.
bool rand_threshold(float p, bool strict, uint32_t* U, bool* all);

// Throws ParseError (NotFound / NotRecognized / Unsupported) for rule text the reference would have
// passed on to the GLSL compiler and failed there.
CompiledRules compile_rules(const ParsingResult& parsed);

}  // namespace se
