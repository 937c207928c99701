// Minimal YAML reader for the sandengine rule language (materials.yaml).
//
// The reference front end reads its input with serde_yaml 0.9 (sandengine-lang/src/parser.rs:95).
// This is a from-scratch reader for the subset that rule files use: block mappings, block and flow
// sequences, flow mappings, plain / quoted scalars, comments.  Scalars are resolved with the YAML 1.2
// core schema the way serde_yaml does (true/false only -- `yes`/`on` stay strings; ints incl. 0x/0o;
// floats incl. .inf/.nan; ~/null/empty => null).  Mapping order is preserved (rule order is semantic:
// sandengine-lang/src/lib.rs:83-100).  Anchors, tags, block scalars and multi-line plain scalars are
// rejected with an error rather than mis-read.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace se {

struct YamlValue {
    enum Kind { Null, Bool, Int, Float, String, Seq, Map };
    Kind kind = Null;
    bool b = false;
    int64_t i = 0;
    double f = 0.0;
    std::string s;                                      // String payload; also the source text of any scalar
    std::vector<YamlValue> seq;
    std::vector<std::pair<YamlValue, YamlValue>> map;   // insertion ordered

    bool is_map() const { return kind == Map; }
    bool is_seq() const { return kind == Seq; }
    bool is_str() const { return kind == String; }
    bool is_bool() const { return kind == Bool; }
    // serde_yaml `Value::as_f64`: any number.   `as_u64`: non-negative integer.
    bool as_f64(double* out) const;
    bool as_u64(uint64_t* out) const;
    // `Value::get(key)`: only mappings have members; missing => nullptr.
    const YamlValue* get(const std::string& key) const;
    // Debug-ish rendering used in InvalidType messages.
    std::string repr() const;
};

// Throws std::runtime_error("yaml: line N: ...") on malformed / unsupported input.
YamlValue yaml_parse(const std::string& text);

}  // namespace se
