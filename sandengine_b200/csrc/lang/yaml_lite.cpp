#include "yaml_lite.h"

#include <cctype>
#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace se {

namespace {

[[noreturn]] void fail(int line, const std::string& msg) {
    throw std::runtime_error("yaml: line " + std::to_string(line) + ": " + msg);
}

struct Line {
    int indent;
    std::string text;  // content without indentation / trailing comment / trailing blanks
    int no;
};

bool all_digits(const std::string& s, size_t from, int base) {
    if (from >= s.size()) return false;
    for (size_t k = from; k < s.size(); ++k) {
        char c = s[k];
        bool ok = base == 10 ? std::isdigit((unsigned char)c) : base == 8 ? (c >= '0' && c <= '7') : base == 2 ? (c == '0' || c == '1') : std::isxdigit((unsigned char)c);
        if (!ok) return false;
    }
    return true;
}

bool looks_float(const std::string& s) {
    // [-+]?(\.[0-9]+|[0-9]+(\.[0-9]*)?)([eE][-+]?[0-9]+)?
    size_t k = 0, n = s.size();
    if (k < n && (s[k] == '-' || s[k] == '+')) ++k;
    size_t digits = 0;
    if (k < n && s[k] == '.') {
        ++k;
        while (k < n && std::isdigit((unsigned char)s[k])) { ++k; ++digits; }
        if (!digits) return false;
    } else {
        while (k < n && std::isdigit((unsigned char)s[k])) { ++k; ++digits; }
        if (!digits) return false;
        if (k < n && s[k] == '.') {
            ++k;
            while (k < n && std::isdigit((unsigned char)s[k])) ++k;
        }
    }
    if (k < n && (s[k] == 'e' || s[k] == 'E')) {
        ++k;
        if (k < n && (s[k] == '-' || s[k] == '+')) ++k;
        size_t ed = 0;
        while (k < n && std::isdigit((unsigned char)s[k])) { ++k; ++ed; }
        if (!ed) return false;
    }
    return k == n;
}

// YAML 1.2 core-schema resolution of a plain (unquoted) scalar.
YamlValue resolve_plain(const std::string& s) {
    YamlValue v;
    v.s = s;
    if (s.empty() || s == "~" || s == "null" || s == "Null" || s == "NULL") { v.kind = YamlValue::Null; return v; }
    if (s == "true" || s == "True" || s == "TRUE") { v.kind = YamlValue::Bool; v.b = true; return v; }
    if (s == "false" || s == "False" || s == "FALSE") { v.kind = YamlValue::Bool; v.b = false; return v; }
    // Integers as serde_yaml 0.9.34 resolves them (the reference's Cargo.lock; de.rs parse_unsigned_int / parse_negative_int /
    // digits_but_not_number): optional sign, then decimal, 0x, 0o or 0b digits; a decimal with a leading zero ("007", "-012") is a
    // string -- and stays one, it is not retried as a float.
    size_t sign = (s[0] == '-' || s[0] == '+') ? 1 : 0;
    if (all_digits(s, sign, 10)) {
        if (s.size() - sign > 1 && s[sign] == '0') { v.kind = YamlValue::String; return v; }
        errno = 0;
        long long x = std::strtoll(s.c_str(), nullptr, 10);
        if (errno == 0) { v.kind = YamlValue::Int; v.i = x; return v; }
        v.kind = YamlValue::Float; v.f = std::strtod(s.c_str(), nullptr); return v;
    }
    if (s.size() > sign + 2 && s[sign] == '0') {
        const int radix = s[sign + 1] == 'x' ? 16 : s[sign + 1] == 'o' ? 8 : s[sign + 1] == 'b' ? 2 : 0;
        if (radix && all_digits(s, sign + 2, radix)) {
            const int64_t mag = (int64_t)std::strtoull(s.c_str() + sign + 2, nullptr, radix);
            v.kind = YamlValue::Int;
            v.i = s[0] == '-' ? -mag : mag;
            return v;
        }
    }
    if (looks_float(s)) { v.kind = YamlValue::Float; v.f = std::strtod(s.c_str(), nullptr); return v; }
    {
        std::string t = s;
        size_t o = (t[0] == '-' || t[0] == '+') ? 1 : 0;
        std::string body = t.substr(o);
        if (body == ".inf" || body == ".Inf" || body == ".INF") { v.kind = YamlValue::Float; v.f = (t[0] == '-') ? -INFINITY : INFINITY; return v; }
        if (t == ".nan" || t == ".NaN" || t == ".NAN") { v.kind = YamlValue::Float; v.f = NAN; return v; }
    }
    v.kind = YamlValue::String;
    return v;
}

std::string rtrim(std::string s) {
    while (!s.empty() && (s.back() == ' ' || s.back() == '\t' || s.back() == '\r')) s.pop_back();
    return s;
}
std::string trim(const std::string& s) {
    size_t a = 0;
    while (a < s.size() && (s[a] == ' ' || s[a] == '\t')) ++a;
    return rtrim(s.substr(a));
}

// Remove a trailing comment (`#` at line start or after whitespace, outside quotes).
std::string strip_comment(const std::string& s) {
    char q = 0;
    for (size_t k = 0; k < s.size(); ++k) {
        char c = s[k];
        if (q) {
            if (q == '"' && c == '\\') { ++k; continue; }
            if (c == q) q = 0;
        } else if (c == '"' || c == '\'') {
            // a quote only opens a quoted scalar at a token start
            if (k == 0 || s[k - 1] == ' ' || s[k - 1] == '[' || s[k - 1] == '{' || s[k - 1] == ',' || s[k - 1] == ':') q = c;
        } else if (c == '#' && (k == 0 || s[k - 1] == ' ' || s[k - 1] == '\t')) {
            return s.substr(0, k);
        }
    }
    return s;
}

int flow_balance(const std::string& s) {
    int depth = 0;
    char q = 0;
    for (size_t k = 0; k < s.size(); ++k) {
        char c = s[k];
        if (q) {
            if (q == '"' && c == '\\') { ++k; continue; }
            if (c == q) q = 0;
        } else if (c == '"' || c == '\'') q = c;
        else if (c == '[' || c == '{') ++depth;
        else if (c == ']' || c == '}') --depth;
    }
    return depth;
}

std::string unquote(const std::string& s, int line) {
    char q = s[0];
    std::string out;
    size_t k = 1;
    for (; k < s.size(); ++k) {
        char c = s[k];
        if (q == '\'') {
            if (c == '\'') {
                if (k + 1 < s.size() && s[k + 1] == '\'') { out.push_back('\''); ++k; continue; }
                break;
            }
            out.push_back(c);
        } else {
            if (c == '\\' && k + 1 < s.size()) {
                char e = s[++k];
                switch (e) {
                    case 'n': out.push_back('\n'); break;
                    case 't': out.push_back('\t'); break;
                    case '0': out.push_back('\0'); break;
                    case '\\': out.push_back('\\'); break;
                    case '"': out.push_back('"'); break;
                    case '/': out.push_back('/'); break;
                    default: fail(line, std::string("unsupported escape \\") + e);
                }
                continue;
            }
            if (c == '"') break;
            out.push_back(c);
        }
    }
    if (k >= s.size()) fail(line, "unterminated quoted scalar");
    if (!trim(s.substr(k + 1)).empty()) fail(line, "unexpected text after quoted scalar");
    return out;
}

YamlValue scalar_from_text(const std::string& raw, int line) {
    std::string s = trim(raw);
    if (!s.empty() && (s[0] == '"' || s[0] == '\'')) {
        YamlValue v;
        v.kind = YamlValue::String;
        v.s = unquote(s, line);
        return v;
    }
    if (!s.empty() && (s[0] == '&' || s[0] == '*' || s[0] == '!' || s[0] == '|' || s[0] == '>' || s[0] == '%' || s[0] == '@' || s[0] == '`'))
        fail(line, "unsupported YAML feature (anchor/alias/tag/block scalar) near '" + s + "'");
    return resolve_plain(s);
}

// ---- flow collections ------------------------------------------------------------------------
struct FlowParser {
    const std::string& s;
    size_t p = 0;
    int line;
    FlowParser(const std::string& str, int ln) : s(str), line(ln) {}
    void ws() { while (p < s.size() && (s[p] == ' ' || s[p] == '\t' || s[p] == '\n')) ++p; }
    std::string token(bool in_map_key) {
        ws();
        size_t start = p;
        if (p < s.size() && (s[p] == '"' || s[p] == '\'')) {
            char q = s[p++];
            while (p < s.size()) {
                if (q == '"' && s[p] == '\\') { p += 2; continue; }
                if (s[p] == q) {
                    if (q == '\'' && p + 1 < s.size() && s[p + 1] == '\'') { p += 2; continue; }
                    ++p;
                    break;
                }
                ++p;
            }
            return s.substr(start, p - start);
        }
        while (p < s.size()) {
            char c = s[p];
            if (c == ',' || c == ']' || c == '}') break;
            if (in_map_key && c == ':' && (p + 1 >= s.size() || s[p + 1] == ' ' || s[p + 1] == ',' || s[p + 1] == '}')) break;
            if (c == '[' || c == '{') fail(line, "unexpected '" + std::string(1, c) + "' inside flow scalar");
            ++p;
        }
        return s.substr(start, p - start);
    }
    YamlValue value() {
        ws();
        if (p >= s.size()) fail(line, "unexpected end of flow collection");
        if (s[p] == '[') {
            ++p;
            YamlValue v;
            v.kind = YamlValue::Seq;
            ws();
            if (p < s.size() && s[p] == ']') { ++p; return v; }
            while (true) {
                ws();
                if (p < s.size() && s[p] == ',') fail(line, "empty entry in flow sequence");   // libyaml: "did not find expected node content"
                v.seq.push_back(value());
                ws();
                if (p >= s.size()) fail(line, "unterminated flow sequence");
                if (s[p] == ',') {
                    ++p;
                    ws();
                    if (p < s.size() && s[p] == ']') { ++p; return v; }  // trailing comma
                    continue;
                }
                if (s[p] == ']') { ++p; return v; }
                fail(line, "expected ',' or ']' in flow sequence");
            }
        }
        if (s[p] == '{') {
            ++p;
            YamlValue v;
            v.kind = YamlValue::Map;
            ws();
            if (p < s.size() && s[p] == '}') { ++p; return v; }
            while (true) {
                ws();
                if (p < s.size() && (s[p] == ':' || s[p] == ',')) fail(line, "empty key in flow mapping");
                YamlValue key = scalar_from_text(token(true), line);
                ws();
                YamlValue val;
                if (p < s.size() && s[p] == ':') { ++p; val = value(); }
                for (auto& kv : v.map)
                    if (kv.first.kind == key.kind && kv.first.s == key.s) fail(line, "duplicate entry with key '" + key.s + "'");
                v.map.emplace_back(std::move(key), std::move(val));
                ws();
                if (p >= s.size()) fail(line, "unterminated flow mapping");
                if (s[p] == ',') { ++p; ws(); if (p < s.size() && s[p] == '}') { ++p; return v; } continue; }
                if (s[p] == '}') { ++p; return v; }
                fail(line, "expected ',' or '}' in flow mapping");
            }
        }
        return scalar_from_text(token(false), line);
    }
};

YamlValue parse_inline(const std::string& text, int line) {
    std::string t = trim(text);
    if (!t.empty() && (t[0] == '[' || t[0] == '{')) {
        FlowParser fp(t, line);
        YamlValue v = fp.value();
        fp.ws();
        if (fp.p != t.size()) fail(line, "unexpected text after flow collection");
        return v;
    }
    return scalar_from_text(t, line);
}

// Position of the ':' that separates key and value in a block-mapping line, or npos.
size_t find_key_sep(const std::string& t) {
    if (t.empty()) return std::string::npos;
    size_t k = 0;
    if (t[0] == '"' || t[0] == '\'') {
        char q = t[0];
        for (k = 1; k < t.size(); ++k) {
            if (q == '"' && t[k] == '\\') { ++k; continue; }
            if (t[k] == q) {
                if (q == '\'' && k + 1 < t.size() && t[k + 1] == '\'') { ++k; continue; }
                break;
            }
        }
        ++k;
        while (k < t.size() && t[k] == ' ') ++k;
        if (k < t.size() && t[k] == ':' && (k + 1 == t.size() || t[k + 1] == ' ')) return k;
        return std::string::npos;
    }
    if (t[0] == '[' || t[0] == '{') return std::string::npos;
    for (k = 0; k < t.size(); ++k)
        if (t[k] == ':' && (k + 1 == t.size() || t[k + 1] == ' ' || t[k + 1] == '\t')) return k;
    return std::string::npos;
}

struct BlockParser {
    std::vector<Line> lines;
    size_t pos = 0;

    bool is_seq_item(const Line& l) const { return l.text == "-" || (l.text.size() >= 2 && l.text[0] == '-' && l.text[1] == ' '); }

    // Multi-line plain scalar: `text` (first line, already consumed) continues on the following lines as long as they
    // are indented deeper than `parent_indent`; YAML folds each line break into one space (libyaml / serde_yaml do, e.g.
    // a long `if:` condition wrapped over two lines).  Quoted / flow / tagged values do not continue here, and a
    // continuation line that itself looks like `key: value` is an error in libyaml too.
    void fold_plain(std::string& text, int parent_indent, int first_no) {
        const char c0 = text.empty() ? ' ' : text[0];
        const bool plain = !(c0 == '"' || c0 == '\'' || c0 == '[' || c0 == '{' || c0 == '&' || c0 == '*' || c0 == '!' || c0 == '|' ||
                             c0 == '>' || c0 == '%' || c0 == '@' || c0 == '`');
        int last_no = first_no;
        while (pos < lines.size() && lines[pos].indent > parent_indent) {
            const Line& c = lines[pos];
            if (!plain) fail(c.no, "unexpected indented line after a complete value");
            if (c.no != last_no + 1) fail(c.no, "blank or comment line inside a multi-line plain scalar is not supported");
            if (find_key_sep(c.text) != std::string::npos) fail(c.no, "mapping values are not allowed in this context");
            text += " " + c.text;
            last_no = c.no;
            ++pos;
        }
    }

    YamlValue parse_block(int indent) {
        const Line& l = lines[pos];
        if (is_seq_item(l)) return parse_seq(indent);
        if (find_key_sep(l.text) != std::string::npos) return parse_map(indent);
        // single scalar / flow value on its own line
        YamlValue v = parse_inline(l.text, l.no);
        ++pos;
        if (pos < lines.size() && lines[pos].indent >= indent && !(v.kind == YamlValue::Seq || v.kind == YamlValue::Map))
            fail(lines[pos].no, "multi-line plain scalars are not supported");
        return v;
    }

    YamlValue parse_map(int indent) {
        YamlValue m;
        m.kind = YamlValue::Map;
        while (pos < lines.size()) {
            const Line l = lines[pos];
            if (l.indent < indent) break;
            if (l.indent > indent) fail(l.no, "unexpected indentation");
            if (is_seq_item(l)) fail(l.no, "sequence item where a mapping key was expected");
            size_t sep = find_key_sep(l.text);
            if (sep == std::string::npos) fail(l.no, "expected 'key: value'");
            YamlValue key = scalar_from_text(l.text.substr(0, sep), l.no);
            if (key.kind == YamlValue::Seq || key.kind == YamlValue::Map) fail(l.no, "complex mapping keys are not supported");
            std::string rest = trim(l.text.substr(sep + 1));
            ++pos;
            YamlValue val;
            if (rest.empty()) {
                if (pos < lines.size() && lines[pos].indent > indent) val = parse_block(lines[pos].indent);
                else if (pos < lines.size() && lines[pos].indent == indent && is_seq_item(lines[pos])) val = parse_seq(indent);
                // else: null
            } else {
                {   // a plain scalar cannot contain ": " (libyaml: "mapping values are not allowed in this context"),
                    // e.g. an unquoted `a ? b : c` condition
                    const char c0 = rest[0];
                    if (c0 != '"' && c0 != '\'' && c0 != '[' && c0 != '{' && find_key_sep(rest) != std::string::npos)
                        fail(l.no, "mapping values are not allowed in this context (quote the value)");
                }
                fold_plain(rest, indent, l.no);
                val = parse_inline(rest, l.no);
            }
            for (auto& kv : m.map)
                if (kv.first.kind == key.kind && kv.first.s == key.s) fail(l.no, "duplicate entry with key '" + key.s + "'");
            m.map.emplace_back(std::move(key), std::move(val));
        }
        return m;
    }

    YamlValue parse_seq(int indent) {
        YamlValue s;
        s.kind = YamlValue::Seq;
        while (pos < lines.size()) {
            Line& l = lines[pos];
            if (l.indent < indent) break;
            if (l.indent > indent) fail(l.no, "unexpected indentation in sequence");
            if (!is_seq_item(l)) break;
            std::string rest = l.text.size() > 1 ? l.text.substr(2) : std::string();
            size_t lead = 0;
            while (lead < rest.size() && rest[lead] == ' ') ++lead;
            std::string item = trim(rest);
            if (item.empty()) {
                ++pos;
                if (pos < lines.size() && lines[pos].indent > indent) s.seq.push_back(parse_block(lines[pos].indent));
                else s.seq.push_back(YamlValue());
            } else if (find_key_sep(item) != std::string::npos || (item == "-" || (item.size() > 1 && item[0] == '-' && item[1] == ' '))) {
                // "- key: value" (compact mapping) or nested "- - x": re-read the remainder as a block
                // that starts at the column of its first character.
                l.indent = indent + 2 + (int)lead;
                l.text = item;
                s.seq.push_back(parse_block(l.indent));
            } else {
                const int item_no = l.no;
                ++pos;                          // (l is not used below: fold_plain may walk past it)
                fold_plain(item, indent, item_no);
                s.seq.push_back(parse_inline(item, item_no));
            }
        }
        return s;
    }
};

}  // namespace

bool YamlValue::as_f64(double* out) const {
    if (kind == Int) { *out = (double)i; return true; }
    if (kind == Float) { *out = f; return true; }
    return false;
}
bool YamlValue::as_u64(uint64_t* out) const {
    if (kind == Int && i >= 0) { *out = (uint64_t)i; return true; }
    return false;
}
const YamlValue* YamlValue::get(const std::string& key) const {
    if (kind != Map) return nullptr;
    for (auto& kv : map)
        if (kv.first.kind == String && kv.first.s == key) return &kv.second;
    return nullptr;
}
std::string YamlValue::repr() const {
    switch (kind) {
        case Null: return "Null";
        case Bool: return b ? "Bool(true)" : "Bool(false)";
        case Int: return "Number(" + std::to_string(i) + ")";
        case Float: return "Number(" + s + ")";
        case String: return "String(\"" + s + "\")";
        case Seq: return "Sequence [..]";
        case Map: return "Mapping {..}";
    }
    return "?";
}

YamlValue yaml_parse(const std::string& text) {
    BlockParser bp;
    // 1. physical lines -> logical lines (comments stripped, multi-line flow collections joined)
    std::vector<std::pair<std::string, int>> phys;
    {
        size_t start = 0;
        int no = 1;
        while (start <= text.size()) {
            size_t end = text.find('\n', start);
            if (end == std::string::npos) end = text.size();
            phys.emplace_back(text.substr(start, end - start), no++);
            start = end + 1;
        }
    }
    for (size_t k = 0; k < phys.size(); ++k) {
        std::string raw = rtrim(strip_comment(phys[k].first));
        int no = phys[k].second;
        size_t ind = 0;
        while (ind < raw.size() && raw[ind] == ' ') ++ind;
        if (ind < raw.size() && raw[ind] == '\t') fail(no, "tab characters are not allowed for indentation");
        std::string content = raw.substr(ind);
        if (content.empty()) continue;
        if (content == "---" && ind == 0) {
            if (!bp.lines.empty()) fail(no, "multiple documents are not supported");
            continue;
        }
        if (content == "..." && ind == 0) break;
        // Brackets are flow syntax only when the VALUE of the line starts with one (`key: [a,` / `- {x: 1,` / `[a,`); inside a
        // plain scalar (`if: a[0] > b`) they are ordinary characters, as in libyaml.
        size_t vstart = 0;
        {
            std::string rest = content;
            while (rest.size() >= 2 && rest[0] == '-' && rest[1] == ' ') {        // "- - [..." / "- key: [..."
                size_t skip = 2;
                while (skip < rest.size() && rest[skip] == ' ') ++skip;
                vstart += skip;
                rest = rest.substr(skip);
            }
            if (!rest.empty() && rest[0] != '[' && rest[0] != '{') {
                size_t sep = find_key_sep(rest);
                if (sep != std::string::npos) {
                    size_t v = sep + 1;
                    while (v < rest.size() && rest[v] == ' ') ++v;
                    vstart += v;
                    rest = rest.substr(v);
                }
            }
            if (rest.empty() || (rest[0] != '[' && rest[0] != '{')) vstart = std::string::npos;
        }
        if (vstart != std::string::npos) {
            int bal = flow_balance(content.substr(vstart));
            while (bal > 0) {
                if (++k >= phys.size()) fail(no, "unterminated flow collection");
                std::string more = trim(strip_comment(phys[k].first));
                content += " " + more;
                bal = flow_balance(content.substr(vstart));
            }
            if (bal < 0) fail(no, "unbalanced ']' or '}'");
        }
        bp.lines.push_back(Line{(int)ind, content, no});
    }
    if (bp.lines.empty()) return YamlValue();
    YamlValue v = bp.parse_block(bp.lines[0].indent);
    if (bp.pos < bp.lines.size()) fail(bp.lines[bp.pos].no, "unexpected content (bad indentation?)");
    return v;
}

}  // namespace se
