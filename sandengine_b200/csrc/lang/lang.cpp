#include "lang.h"

#include <charconv>
#include <cmath>
#include <regex>

#include "yaml_lite.h"

namespace se {

namespace {

// parser.rs:27-34
const char* const GLOBAL_CELLNAMES[6] = {"SELF", "LEFT", "RIGHT", "DOWN", "DOWNRIGHT", "DOWNLEFT"};

// does the raw if/do text of a rule name the LEFT / RIGHT column as a cell (whole identifiers)?
void mentions_cells(const std::string& text, bool& left, bool& right) {
    left = right = false;
    auto ident = [](char c) { return (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || (c >= '0' && c <= '9') || c == '_'; };
    for (size_t i = 0; i < text.size();) {
        if (!ident(text[i])) { ++i; continue; }
        size_t j = i;
        while (j < text.size() && ident(text[j])) ++j;
        const std::string tok = text.substr(i, j - i);
        if (tok == "LEFT" || tok == "DOWNLEFT") left = true;
        if (tok == "RIGHT" || tok == "DOWNRIGHT") right = true;
        i = j;
    }
}
bool is_cellname(const std::string& s) {
    for (auto* c : GLOBAL_CELLNAMES)
        if (s == c) return true;
    return false;
}
constexpr float DEFAULT_VAL_PROBABILITY = 1.0f;   // parser.rs:39

// Error constructors: message layout follows the thiserror strings of parser.rs:46-72 (without ANSI colours).
ParseError missing_field(const std::string& field, const std::string& in) {
    return ParseError(ErrKind::MissingField, "(MissingField) Mandatory field '" + field + "' is missing in '" + in + "'");
}
ParseError invalid_type(const std::string& wrong, const std::string& in, const std::string& expected) {
    return ParseError(ErrKind::InvalidType, "(InvalidType) The type of the field '" + wrong + "' inside of '" + in + "' is invalid. Expected: '" + expected + "'");
}
ParseError not_found(const std::string& missing, const std::string& in) {
    return ParseError(ErrKind::NotFound, "(NotFound) The name '" + missing + "' (in '" + in + "') was not found. Make sure it was defined before referencing it.");
}
ParseError not_recognized(const std::string& what, const std::string& in) {
    return ParseError(ErrKind::NotRecognized, "(NotRecognized) The expression '" + what + "' (in '" + in + "') was not recognized as valid syntax. Please check it is valid.");
}

const char* TYPE_HINT_STRING = "string";
const char* TYPE_HINT_BOOL = "bool (true/false)";
const char* TYPE_HINT_FLOAT = "float (0.0 to 1.0)";
const char* TYPE_HINT_SEQUENCE = "sequence (array, '[...]')";
const char* TYPE_HINT_COLOR = "sequence (array, '[...]') of 3-4 floats (range 0.0-1.0) OR integers (range 0-255). (With 3 elements, the alpha channel defaults to 1.0)";
const char* TYPE_HINT_MAPPING = "mapping (dictionary-like)";

void replace_all(std::string& s, const std::string& from, const std::string& to) {
    if (from.empty()) return;
    std::string out;
    size_t pos = 0, hit;
    while ((hit = s.find(from, pos)) != std::string::npos) {
        out.append(s, pos, hit - pos);
        out += to;
        pos = hit + from.size();
    }
    out.append(s, pos, std::string::npos);
    s.swap(out);
}

std::string trim_end(std::string s) {
    while (!s.empty() && std::isspace((unsigned char)s.back())) s.pop_back();
    return s;
}

bool contains(const std::vector<std::string>& v, const std::string& s) {
    for (auto& e : v)
        if (e == s) return true;
    return false;
}

// rules.rs:338-351 -- ordered literal replaces
void parse_global_scope(std::string& s) {
    replace_all(s, " or ", " || ");
    replace_all(s, " and ", " && ");
    replace_all(s, "not ", " !");
    replace_all(s, "empty", "MAT_EMPTY");
    replace_all(s, "SELF", "self");
    replace_all(s, "RIGHT", "right");
    replace_all(s, "LEFT", "left");
    replace_all(s, "DOWN", "down");
    replace_all(s, "DOWNRIGHT", "downright");
    replace_all(s, "DOWNLEFT", "downleft");
}

// rules.rs:355-420
std::string parse_do(const std::string& parent, const std::string& do_str) {
    static const std::regex swap_re("SWAP (\\w+) (\\w+)");
    static const std::regex set_re("SET (\\w+) (\\w+)");
    std::string out;
    bool found = false;
    std::smatch m;
    if (std::regex_search(do_str, m, swap_re)) {
        found = true;
        if (!is_cellname(m[1].str())) throw not_found(m[1].str(), parent);
        if (!is_cellname(m[2].str())) throw not_found(m[2].str(), parent);
        out += "swap(" + m[1].str() + ", " + m[2].str() + ");\n";
    }
    if (std::regex_search(do_str, m, set_re)) {
        found = true;
        if (!is_cellname(m[1].str())) throw not_found(m[1].str(), parent);
        out += m[1].str() + " = newCell(MAT_" + m[2].str() + ", pos);\n";
    }
    if (!found) throw not_recognized(do_str, parent);
    return out;
}

// rules.rs:207-334
void parse_conditionals(const YamlValue& parent, bool parent_is_else, std::string parent_path, SandRule& rule,
                        const std::vector<std::string>& type_names, const std::vector<std::string>& material_names,
                        std::string& raw_text) {
    static const std::regex mat_re("\\w*.mat\\s*(==|!=)\\s*(\\w*)");
    static const std::regex type_re("isType_(\\w*)\\(\\w*\\)");
    const YamlValue* if_cond = parent.get("if");
    if (!if_cond && !parent_is_else) {
        throw missing_field("if", parent_path);
    } else if (if_cond) {
        parent_path += "/if";
        if (!if_cond->is_str()) throw invalid_type("if", parent_path, TYPE_HINT_STRING);
        std::string cond = if_cond->s;
        raw_text += cond + " ";
        parse_global_scope(cond);
        {
            std::string snapshot = cond;
            for (std::sregex_iterator it(snapshot.begin(), snapshot.end(), mat_re), end; it != end; ++it) {
                std::string cap = (*it)[2].str();
                bool ok = false;
                for (auto& mname : material_names) {
                    if (mname == cap) {
                        replace_all(cond, mname, "MAT_" + mname);   // global substring replace, rules.rs:252
                        ok = true;
                        break;
                    }
                }
                if (!ok) throw not_found(cap, parent_path);
            }
        }
        {
            std::string snapshot = cond;
            for (std::sregex_iterator it(snapshot.begin(), snapshot.end(), type_re), end; it != end; ++it) {
                std::string cap = (*it)[1].str();
                if (!contains(type_names, cap)) throw not_found(cap, parent_path + " -> isType_");
            }
        }
        rule.if_conds.push_back(cond);
    }

    const YamlValue* do_action = parent.get("do");
    if (!do_action) throw missing_field("do", parent_path);
    std::string do_parent_path = parent_path + "/do";
    std::string do_string;
    if (do_action->is_str()) {
        raw_text += do_action->s + " ";
        do_string = parse_do(do_parent_path, do_action->s);
        parse_global_scope(do_string);
    }
    if (do_action->is_seq()) {
        for (auto& item : do_action->seq) {
            if (item.is_str()) {
                raw_text += item.s + " ";
                std::string part = parse_do(do_parent_path, item.s);
                parse_global_scope(part);
                do_string += part;
            }
        }
    }
    rule.do_actions.push_back(trim_end(do_string));

    const YamlValue* prob = parent.get("probability");
    if (prob) {
        double p;
        if (!prob->as_f64(&p)) throw invalid_type("probability", parent_path, TYPE_HINT_FLOAT);
        rule.probabilities.push_back((float)p);
    } else {
        rule.probabilities.push_back(DEFAULT_VAL_PROBABILITY);
    }

    const YamlValue* else_ = parent.get("else");
    if (else_) parse_conditionals(*else_, true, parent_path + "/else", rule, type_names, material_names, raw_text);
}

// rules.rs:105-203
std::vector<SandRule> parse_rules(const YamlValue& rules, const std::vector<std::string>& type_names,
                                  const std::vector<std::string>& material_names) {
    std::vector<SandRule> out;
    for (auto& kv : rules.map) {
        if (!kv.first.is_str()) throw invalid_type(kv.first.repr(), "rules", TYPE_HINT_STRING);
        SandRule rule;
        rule.name = kv.first.s;
        std::string raw_text;
        parse_conditionals(kv.second, false, "rules/" + rule.name, rule, type_names, material_names, raw_text);

        if (const YamlValue* m = kv.second.get("mirrored")) {
            if (!m->is_bool()) throw invalid_type("mirrored", "rules/" + rule.name, TYPE_HINT_BOOL);
            rule.mirror = m->b;
        }
        // rules.rs:154-163 -- the text is already lower-case here, so "LEFT" never matches (kept for the record).
        if (rule.mirror) rule.ruletype = SandRuleType::Mirrored;
        else rule.ruletype = rule.do_actions[0].find("LEFT") != std::string::npos ? SandRuleType::Left : SandRuleType::Right;

        bool do_precondition = true;
        if (const YamlValue* p = kv.second.get("precondition")) {
            if (!p->is_bool()) throw invalid_type("precondition", "rules/" + rule.name, TYPE_HINT_BOOL);
            do_precondition = p->b;
        }
        rule.has_precondition = do_precondition;

        // whole identifiers only (a material LEFTOVER is not a cell name); a conflicting rule is an error only once a type or a
        // material uses it -- the reference emits used rules only, so an unused draft rule must not reject the file
        mentions_cells(raw_text, rule.mentions_left, rule.mentions_right);
        if (rule.mentions_left && (rule.mirror || rule.mentions_right)) rule.left_conflict = trim_end(raw_text);
        out.push_back(std::move(rule));
    }
    return out;
}

// a rule that a type or a material refers to must be executable
void require_usable(const SandRule& rule) {
    if (!rule.left_conflict.empty())
        throw not_recognized(rule.left_conflict, "rules/" + rule.name + " (LEFT in a mirrored rule, or LEFT mixed with RIGHT)");
}

// types.rs:78-88
void update_rule_precondition(SandRule& rule, const std::string& clause) {
    if (!rule.has_precondition) return;
    if (rule.precondition.empty()) rule.precondition = clause;
    else rule.precondition += " || " + clause;
}

// types.rs:186-200.  The reference recurses without a guard and overflows the stack on an inheritance cycle
// (`a: {inherits: a}`); here a cycle is an InvalidType error.
void add_child_to_type(const std::string& parent_name, const std::string& child, std::vector<SandType>& types, size_t depth = 0) {
    if (depth > types.size() + 1) throw invalid_type("inherits", "types/" + child, "an acyclic chain of parent types");
    std::string pp;
    for (auto& t : types) {
        if (t.name == parent_name) {
            t.children.push_back(child);
            pp = t.inherits;
            break;
        }
    }
    if (!pp.empty()) add_child_to_type(pp, child, types, depth + 1);
}

// types.rs:202-210
std::vector<std::string> get_parents_rules(const std::vector<SandType>& all, const SandType& cur, size_t depth = 0) {
    if (cur.inherits.empty()) return {};
    if (depth > all.size() + 1) throw invalid_type("inherits", "types/" + cur.name, "an acyclic chain of parent types");
    const SandType* parent = nullptr;
    for (auto& t : all)
        if (t.name == cur.inherits) { parent = &t; break; }
    if (!parent) throw not_found(cur.inherits, "types/" + cur.name + "/inherits");   // reference: unwrap() panic
    std::vector<std::string> rules = parent->base_rules;
    auto more = get_parents_rules(all, *parent, depth + 1);
    rules.insert(rules.end(), more.begin(), more.end());
    return rules;
}

SandRule* find_rule(std::vector<SandRule>& rules, const std::string& name) {
    for (auto& r : rules)
        if (r.name == name) return &r;
    return nullptr;
}

// types.rs:51-182
std::vector<SandType> parse_types(const YamlValue& types, std::vector<SandRule>& rules, const std::vector<std::string>& rule_names,
                                  const std::vector<std::string>& type_names) {
    std::vector<SandType> out(3);
    out[0].id = 0; out[0].name = "EMPTY";
    out[1].id = 1; out[1].name = "NULL";
    out[2].id = 2; out[2].name = "WALL";
    int idx = 3;
    for (auto& kv : types.map) {
        if (!kv.first.is_str()) throw invalid_type(kv.first.repr(), "types", TYPE_HINT_STRING);
        SandType t;
        t.id = idx;
        t.name = kv.first.s;
        if (const YamlValue* inh = kv.second.get("inherits")) {
            if (!inh->is_str()) throw invalid_type("inherits", "types/" + t.name, TYPE_HINT_STRING);
            t.inherits = inh->s;
            for (auto& tn : type_names)
                if (tn == t.inherits) add_child_to_type(t.inherits, t.name, out);
        }
        if (const YamlValue* br = kv.second.get("base_rules")) {
            if (!br->is_seq()) throw invalid_type("base_rules", "types/" + t.name, TYPE_HINT_SEQUENCE);
            for (auto& item : br->seq) {
                if (!item.is_str()) throw invalid_type("base_rules", "types/" + t.name, TYPE_HINT_STRING);
                if (!contains(rule_names, item.s)) throw not_found(item.s, "types/" + t.name + "/base_rules");
                SandRule* rule = find_rule(rules, item.s);
                require_usable(*rule);
                rule->used = true;
                update_rule_precondition(*rule, "isType_" + t.name + "(self)");
                t.base_rules.push_back(item.s);
            }
        }
        out.push_back(std::move(t));
        ++idx;
    }
    std::vector<SandType> snapshot = out;
    for (auto& st : snapshot) {
        if (st.inherits.empty()) continue;
        for (auto& rn : get_parents_rules(out, st)) {
            SandRule* rule = find_rule(rules, rn);
            update_rule_precondition(*rule, "isType_" + st.name + "(self)");
        }
    }
    return out;
}

// parser.rs:192-251
void extract_vec4(const YamlValue& data, const std::string& parent_name, const char* field, const float def[4], bool mandatory, float out[4]) {
    std::string in = "materials/" + parent_name + "/" + field;
    for (int k = 0; k < 4; ++k) out[k] = def[k];
    const YamlValue* v = data.get(field);
    if (!v) {
        if (mandatory) throw missing_field(field, in);
        return;
    }
    if (!v->is_seq()) throw invalid_type(field, in, TYPE_HINT_COLOR);
    if (v->seq.empty() || v->seq.size() > 4 || v->seq.size() < 3) throw invalid_type(field, in, TYPE_HINT_COLOR);
    for (size_t k = 0; k < v->seq.size(); ++k) {
        const YamlValue& c = v->seq[k];
        uint64_t u;
        if (c.as_u64(&u) && u > 0 && u <= 255) { out[k] = (float)u / 255.0f; continue; }
        double f;
        if (c.as_f64(&f)) {
            if (f > 1.0 && f <= 255.0) { out[k] = (float)f / 255.0f; continue; }
            if (f >= 0.0 && f <= 1.0) { out[k] = (float)f; continue; }
        }
        throw invalid_type(field, in, TYPE_HINT_COLOR);
    }
}

// materials.rs:51-201
std::vector<SandMaterial> parse_materials(const YamlValue& materials, std::vector<SandRule>& rules, const std::vector<std::string>& type_names) {
    std::vector<SandMaterial> out(3);
    out[0].id = 0; out[0].name = out[0].mattype = "EMPTY"; out[0].selectable = true; out[0].density = 1.0f;
    out[1].id = 1; out[1].name = out[1].mattype = "NULL"; out[1].selectable = false; out[1].density = 0.0f;
    out[1].color[0] = 1.f; out[1].color[2] = 1.f; out[1].color[3] = 1.f;
    out[2].id = 2; out[2].name = out[2].mattype = "WALL"; out[2].selectable = false; out[2].density = 9999.0f;
    out[2].color[0] = 0.1f; out[2].color[1] = 0.2f; out[2].color[2] = 0.3f; out[2].color[3] = 1.0f;
    int idx = 3;
    for (auto& kv : materials.map) {
        if (!kv.first.is_str()) throw invalid_type(kv.first.repr(), "materials", TYPE_HINT_STRING);
        SandMaterial m;
        m.id = idx;
        m.name = kv.first.s;
        const YamlValue* ty = kv.second.get("type");
        if (!ty) throw missing_field("type", "materials/" + m.name);
        if (!ty->is_str()) throw invalid_type("type", "materials/" + m.name, TYPE_HINT_STRING);
        m.mattype = ty->s;
        if (!contains(type_names, m.mattype)) throw not_found(m.mattype, "materials/" + m.name + "/type");
        const float def_color[4] = {1.f, 0.f, 1.f, 1.f}, def_em[4] = {0.f, 0.f, 0.f, 0.f};
        extract_vec4(kv.second, m.name, "color", def_color, true, m.color);
        extract_vec4(kv.second, m.name, "emission", def_em, false, m.emission);
        if (const YamlValue* sel = kv.second.get("selectable")) m.selectable = sel->is_bool() ? sel->b : false;
        const YamlValue* dens = kv.second.get("density");
        if (!dens) throw missing_field("density", "materials/" + m.name);
        double d;
        if (!dens->as_f64(&d)) throw invalid_type("density", "materials/" + m.name + "/density", TYPE_HINT_FLOAT);
        m.density = (float)d;
        if (const YamlValue* ex = kv.second.get("extra_rules")) {
            if (!ex->is_seq()) throw invalid_type("extra_rules", "materials/" + m.name, TYPE_HINT_SEQUENCE);
            for (auto& item : ex->seq) {
                if (!item.is_str()) throw invalid_type("extra_rules", "materials/" + m.name, TYPE_HINT_STRING);
                for (auto& r : rules) {   // unknown names are silently ignored (materials.rs:153-165)
                    if (r.name == item.s) {
                        require_usable(r);
                        r.used = true;
                        update_rule_precondition(r, "self.mat == MAT_" + m.name);
                        m.extra_rules.push_back(item.s);
                    }
                }
            }
        }
        out.push_back(std::move(m));
        ++idx;
    }
    return out;
}

// parser.rs:174-187
const YamlValue& check_mapping(const YamlValue& root, const char* key) {
    const YamlValue* v = root.get(key);
    if (!v) throw missing_field(key, "Root/ Base level of YAML file");
    if (!v->is_map()) throw invalid_type(v->repr(), "Root/ Base level of YAML file", TYPE_HINT_MAPPING);
    return *v;
}

// parser.rs:156-169
std::vector<std::string> preparse_keys(const YamlValue& map, const char* err_name) {
    std::vector<std::string> names;
    for (auto& kv : map.map) {
        if (!kv.first.is_str()) throw invalid_type(kv.first.repr(), err_name, TYPE_HINT_STRING);
        names.push_back(kv.first.s);
    }
    return names;
}

// rules.rs:47-73
std::string func_logic(const SandRule& r, size_t i_if, size_t i_do, size_t i_p, size_t indent_lvl) {
    bool no_if = i_if >= r.if_conds.size(), no_do = i_do >= r.do_actions.size();
    if (no_if && no_do) return "";
    if (no_if && !no_do) return r.do_actions[i_do];
    if (no_do) return "";
    std::string ind1(indent_lvl * 4, ' '), ind2((indent_lvl + 1) * 4, ' ');
    float p = r.probabilities[i_p];
    std::string prob = (p == DEFAULT_VAL_PROBABILITY) ? "" : "rand.y <= " + f32_display(p) + " && ";
    return ind1 + "if (" + prob + r.if_conds[i_if] + ") {\n" + ind2 + r.do_actions[i_do] + "\n" + ind1 + "} else {\n" +
           func_logic(r, i_if + 1, i_do + 1, i_p + 1, indent_lvl + 1) + "\n" + ind1 + "}";
}

// rules.rs:75-98
std::string rule_glsl(const SandRule& r) {
    std::string celldir = r.ruletype == SandRuleType::Left ? "left" : "right";
    std::string precond = r.has_precondition ? "    if (!(" + r.precondition + ")) {\n        return;\n    }\n" : "";
    return "void rule_" + r.name + " (inout Cell self, inout Cell " + celldir +
           ", inout Cell down, inout Cell downright, vec4 rand, ivec2 pos) {\n" + precond + func_logic(r, 0, 0, 0, 1) + "\n}";
}

}  // namespace

std::string f32_display(float v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
    char buf[128];
    auto res = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);
    return std::string(buf, res.ptr);
}

ParsingResult parse_string(const std::string& text) {
    YamlValue data;
    try {
        data = yaml_parse(text);
    } catch (const std::runtime_error& e) {
        throw ParseError(ErrKind::Yaml, e.what());
    }
    const YamlValue& raw_rules = check_mapping(data, "rules");
    const YamlValue& raw_types = check_mapping(data, "types");
    const YamlValue& raw_materials = check_mapping(data, "materials");
    std::vector<std::string> rule_names = preparse_keys(raw_rules, "rules");
    std::vector<std::string> type_names = preparse_keys(raw_types, "types");
    type_names.push_back("EMPTY");
    std::vector<std::string> material_names = preparse_keys(raw_materials, "materials");
    material_names.push_back("EMPTY");

    ParsingResult res;
    try {
        res.rules = parse_rules(raw_rules, type_names, material_names);
    } catch (const ParseError& e) {
        throw ParseError(e.kind, std::string("Error while parsing rules: '") + e.what() + "'");
    }
    try {
        res.types = parse_types(raw_types, res.rules, rule_names, type_names);
    } catch (const ParseError& e) {
        throw ParseError(e.kind, std::string("Error while parsing types: '") + e.what() + "'");
    }
    try {
        res.materials = parse_materials(raw_materials, res.rules, type_names);
    } catch (const ParseError& e) {
        throw ParseError(e.kind, std::string("Error while parsing materials: '") + e.what() + "'");
    }
    return res;
}

// sandengine-lang/src/lib.rs:21-65
std::string emit_glsl_materials(const ParsingResult& r) {
    std::string s;
    for (auto& t : r.types) s += "#define TYPE_" + t.name + " " + std::to_string(t.id) + "\n\n";
    for (auto& t : r.types) {
        std::string tc = "return cell.mat.type == TYPE_" + t.name;
        for (auto& c : t.children) tc += " || cell.mat.type == TYPE_" + c;
        s += "bool isType_" + t.name + "(Cell cell) {\n    " + tc + ";\n}\n\n";
    }
    s += "\n";
    std::string list;
    size_t n = r.materials.size();
    for (size_t k = 0; k < n; ++k) {
        const SandMaterial& m = r.materials[k];
        s += "#define MAT_" + m.name + " Material(" + std::to_string(m.id) + ", vec4(" + f32_display(m.color[0]) + ", " +
             f32_display(m.color[1]) + ", " + f32_display(m.color[2]) + ", " + f32_display(m.color[3]) + "), " + f32_display(m.density) +
             ", vec4(" + f32_display(m.emission[0]) + ", " + f32_display(m.emission[1]) + ", " + f32_display(m.emission[2]) + ", " +
             f32_display(m.emission[3]) + "), TYPE_" + m.mattype + ")\n";
        list += "        MAT_" + m.name + (k == n - 1 ? "" : ",\n");
    }
    std::string ns = std::to_string(n);
    s += "\nMaterial[" + ns + "] materials() {\n    Material allMaterials[" + ns + "] = {\n" + list +
         "\n    };\n    return allMaterials;\n}\n\n"
         "Material getMaterialFromID(int id) {\n    for (int i = 0; i < materials().length(); i++) {\n"
         "        if (id == materials()[i].id) {\n            return materials()[i];\n        };\n    };\n"
         "    return MAT_NULL;\n}\n\n";
    return s;
}

// sandengine-lang/src/lib.rs:78-142
std::string emit_glsl_rules(const ParsingResult& r) {
    std::string funcs, mir, left, right;
    for (auto& rule : r.rules) {
        if (!rule.used) continue;
        funcs += rule_glsl(rule) + "\n\n";
        if (rule.ruletype == SandRuleType::Mirrored) mir += "rule_" + rule.name + "(self, right, down, downright, rand, pos);\n";
        else if (rule.ruletype == SandRuleType::Left) left += "rule_" + rule.name + "(self, left, down, downright, rand, pos);\n";
        else right += "rule_" + rule.name + "(self, right, down, downright, rand, pos);\n";
    }
    const std::string hdr =
        "(\n    inout Cell self,\n    inout Cell right,\n    inout Cell down,\n    inout Cell downright,\n    vec4 rand,\n    ivec2 pos) {\n    ";
    return "\n// =============== RULES ===============\n" + funcs + "\n\n\n// =============== CALLERS ===============\n" +
           "void applyMirroredRules" + hdr + trim_end(mir) + "\n}\n\n\nvoid applyLeftRules" + hdr + trim_end(left) +
           "\n}\n\nvoid applyRightRules" + hdr + trim_end(right) + "\n}";
}

}  // namespace se
