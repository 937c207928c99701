// Rule-language front end (host side, C++): the B200 build's replacement for `sandengine-lang`.
//
// Mirrors the reference crate's public surface (file:line in /root/reference):
//   parse_string            sandengine-lang/src/parser.rs:93-152
//   ParsingResult           parser.rs:84-89
//   SandRule/SandRuleType   parser/rules.rs:14-44
//   SandType                parser/types.rs:12-26
//   SandMaterial            parser/materials.rs:10-29
//   ParsingErr              parser.rs:42-73  (MissingField / InvalidType / NotFound / NotRecognized)
//   GLSLConvertible + create_glsl_from_parser   parser.rs:77-79, sandengine-lang/src/lib.rs:17-148
// The GLSL emitter is kept ONLY as a known-answer check (its output for data/materials.yaml must
// hash to the reference's checked-in gen/*.glsl); the product back end is emit_cuda() in codegen.h.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

namespace se {

enum class ErrKind { Yaml = 1, MissingField, InvalidType, NotFound, NotRecognized, Unsupported };

struct ParseError : std::runtime_error {
    ErrKind kind;
    ParseError(ErrKind k, const std::string& msg) : std::runtime_error(msg), kind(k) {}
};

enum class SandRuleType { Mirrored, Left, Right };

struct SandRule {
    std::string name;
    SandRuleType ruletype = SandRuleType::Mirrored;   // the reference's (dead-code) classification, rules.rs:154-163
    std::vector<std::string> if_conds;                // GLSL-shaped text, exactly as the reference builds it
    std::vector<std::string> do_actions;
    std::vector<float> probabilities;
    bool mirror = true;
    bool has_precondition = true;                     // Option<String>::is_some()
    std::string precondition;
    bool used = false;
    bool mentions_left = false;                       // LEFT / DOWNLEFT appears in the raw if/do text
    bool mentions_right = false;
    std::string left_conflict;                        // non-empty: LEFT in a mirrored rule or mixed with RIGHT (an error once the rule is used)
    // Classification this build actually executes: Left iff !mirror && mentions_left.
    SandRuleType effective_type() const {
        if (mirror) return SandRuleType::Mirrored;
        return mentions_left ? SandRuleType::Left : SandRuleType::Right;
    }
};

struct SandType {
    int id = 0;
    std::string name;
    std::string inherits;
    std::vector<std::string> children;
    std::vector<std::string> base_rules;
};

struct SandMaterial {
    int id = 0;
    std::string name;
    std::string mattype;
    float color[4] = {0, 0, 0, 0};
    float emission[4] = {0, 0, 0, 0};
    bool selectable = true;
    float density = 0.f;
    std::vector<std::string> extra_rules;
};

struct ParsingResult {
    std::vector<SandRule> rules;
    std::vector<SandType> types;
    std::vector<SandMaterial> materials;
};

// Throws ParseError.
ParsingResult parse_string(const std::string& yaml_text);

// Rust `Display` of an f32: shortest digits that round-trip, positional notation.
std::string f32_display(float v);

// Known-answer emitters (byte-exact restatement of create_glsl_from_parser's two outputs).
std::string emit_glsl_materials(const ParsingResult& r);
std::string emit_glsl_rules(const ParsingResult& r);

}  // namespace se
