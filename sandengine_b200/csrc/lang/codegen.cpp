#include "codegen.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <sstream>

namespace se {

namespace {

ParseError not_found(const std::string& missing, const std::string& in) {
    return ParseError(ErrKind::NotFound, "(NotFound) The name '" + missing + "' (in '" + in + "') was not found. Make sure it was defined before referencing it.");
}
ParseError not_recognized(const std::string& what, const std::string& in) {
    return ParseError(ErrKind::NotRecognized, "(NotRecognized) The expression '" + what + "' (in '" + in + "') was not recognized as valid syntax. Please check it is valid.");
}
ParseError unsupported(const std::string& what) { return ParseError(ErrKind::Unsupported, "(Unsupported) " + what); }

std::string hex32(uint32_t v) {
    char b[16];
    std::snprintf(b, sizeof b, "0x%08xu", v);
    return b;
}
std::string f32_lit(float v) {   // exact: hex-float (or bit pattern for inf/nan), never re-parsed as decimal
    char b[64];
    if (std::isfinite(v)) {
        std::snprintf(b, sizeof b, "%af", (double)v);
        return b;
    }
    uint32_t bits;
    std::memcpy(&bits, &v, 4);
    std::snprintf(b, sizeof b, "__uint_as_float(0x%08xu)", bits);
    return b;
}

// ---------------------------------------------------------------------------------------------
// Typed values produced while parsing a condition
// ---------------------------------------------------------------------------------------------
enum class VT { Bool, Int, Float, Mat, Cell, Vec4, RandVec, PosVec, VecN };

struct Val {
    VT type = VT::Bool;
    std::string code;          // C expression (Bool/Int/Float), or register name (Cell / Mat-of-cell)
    // special forms kept for strength reduction
    bool is_lit = false;       // Float/Int literal
    double lit = 0;            // literal value (float literals are already rounded to f32)
    bool is_density = false;   // Float: density of cell register `cell`
    bool is_rand = false;      // Float: hash lane `lane`
    int lane = 0;
    std::string cell;          // Mat: owning cell register ("" => constant material `mat_id`)
    int mat_id = -1;
    std::string vec_field;     // Vec4: "color" | "emission"
    std::vector<std::string> comps;   // VecN: float code of each component (2..4): a swizzle of color / emission, or vecN(...)
};

struct Ctx {
    const MaterialTables& tb;
    const ParsingResult& pr;
    const std::vector<float>& rank_values;   // density value of each rank
    bool left_view;                          // rule is a Left rule: `left`/`downleft` bind to r/dr
    std::string where;                       // for messages
    uint32_t* rand_lanes;
    std::vector<uint32_t>* y_thresholds;     // integer thresholds taken on hash lane 1 (rand.y)
    bool* lut_ok;                            // cleared by anything a transition table cannot capture
};

// ---------------------------------------------------------------------------------------------
struct Lexer {
    std::string s;
    size_t p = 0;
    std::string tok;   // current token text
    enum K { End, Ident, Number, Op } kind = End;
    explicit Lexer(const std::string& str) : s(str) { next(); }
    void next() {
        while (p < s.size() && std::isspace((unsigned char)s[p])) ++p;
        if (p >= s.size()) { kind = End; tok.clear(); return; }
        char c = s[p];
        if (std::isalpha((unsigned char)c) || c == '_') {
            size_t st = p;
            while (p < s.size() && (std::isalnum((unsigned char)s[p]) || s[p] == '_')) ++p;
            tok = s.substr(st, p - st);
            kind = Ident;
            return;
        }
        if (std::isdigit((unsigned char)c) || (c == '.' && p + 1 < s.size() && std::isdigit((unsigned char)s[p + 1]))) {
            size_t st = p;
            while (p < s.size() && (std::isalnum((unsigned char)s[p]) || s[p] == '.' ||
                                    ((s[p] == '+' || s[p] == '-') && p > st && (s[p - 1] == 'e' || s[p - 1] == 'E') &&
                                     !(st + 1 < s.size() && s[st] == '0' && (s[st + 1] == 'x' || s[st + 1] == 'X')))))
                ++p;
            tok = s.substr(st, p - st);
            kind = Number;
            return;
        }
        static const char* two[] = {"||", "&&", "==", "!=", "<=", ">=", "<<", ">>"};
        for (auto* t : two)
            if (s.compare(p, 2, t) == 0) { tok = t; p += 2; kind = Op; return; }
        tok = std::string(1, c);
        ++p;
        kind = Op;
    }
    bool is_op(const char* o) const { return kind == Op && tok == o; }
};

struct Parser {
    Lexer lx;
    Ctx& cx;
    std::string text;
    Parser(const std::string& t, Ctx& c) : lx(t), cx(c), text(t) {}

    [[noreturn]] void bad(const std::string& why) { throw not_recognized(text + "  [" + why + "]", cx.where); }

    // ---- helpers ----
    static Val mk_bool(const std::string& c) { Val v; v.type = VT::Bool; v.code = c; return v; }
    static Val mk_int(const std::string& c) { Val v; v.type = VT::Int; v.code = c; return v; }
    static Val mk_float(const std::string& c) { Val v; v.type = VT::Float; v.code = c; return v; }
    static Val mk_flit(float f) { Val v = mk_float(f32_lit(f)); v.is_lit = true; v.lit = f; return v; }
    static Val mk_ilit(long long i) { Val v = mk_int(std::to_string(i)); v.is_lit = true; v.lit = (double)i; return v; }

    std::string as_float_code(const Val& v) {
        if (v.type == VT::Float) {
            if (v.is_rand) *cx.lut_ok = false;   // rand used as a plain float, not against a literal threshold
            return v.code;
        }
        if (v.type == VT::Int) {
            if (v.is_lit) return f32_lit((float)v.lit);
            return "((float)(" + v.code + "))";
        }
        bad("expected a number");
    }
    Val to_bool(const Val& v) {
        if (v.type != VT::Bool) bad("expected a boolean expression");
        return v;
    }

    std::string cell_reg(const std::string& name) {
        if (name == "self") return "s";
        if (name == "down") return "d";
        if (!cx.left_view) {
            if (name == "right") return "r";
            if (name == "downright") return "dr";
            if (name == "left" || name == "downleft")
                throw not_recognized(text, cx.where + " (LEFT/DOWNLEFT is only available in `mirrored: false` rules)");
        } else {
            if (name == "left") return "r";
            if (name == "downleft") return "dr";
            if (name == "right" || name == "downright")
                throw not_recognized(text, cx.where + " (a LEFT rule cannot also use RIGHT/DOWNRIGHT)");
        }
        return "";
    }

    int find_material(const std::string& name) {
        for (int k = 0; k < cx.tb.n_materials; ++k)
            if (cx.tb.material_names[k] == name) return k;
        return -1;
    }
    int find_type(const std::string& name) {
        for (int k = 0; k < cx.tb.n_types; ++k)
            if (cx.tb.type_names[k] == name) return k;
        return -1;
    }

    // density of a cell register against a literal: decided on ranks at code-generation time
    Val density_vs_lit(const std::string& cell, const std::string& op, float lit) {
        // pred(r) = rank_values[r] op lit
        int n = (int)cx.rank_values.size();
        auto pred = [&](int r) {
            float v = cx.rank_values[r];
            if (op == "<") return v < lit;
            if (op == "<=") return v <= lit;
            if (op == ">") return v > lit;
            if (op == ">=") return v >= lit;
            if (op == "==") return v == lit;
            return v != lit;
        };
        // enumerate the (<=256) ranks that satisfy the predicate and emit a compact test
        std::vector<int> yes;
        for (int r = 0; r < n; ++r)
            if (pred(r)) yes.push_back(r);
        if (yes.empty()) return mk_bool("false");
        if ((int)yes.size() == n) return mk_bool("true");
        bool contiguous = yes.back() - yes.front() + 1 == (int)yes.size();
        std::string rk = "SE_RANK(" + cell + ")";
        if (contiguous && yes.front() == 0) return mk_bool("(" + rk + " <= " + std::to_string(yes.back()) + "u)");
        if (contiguous && yes.back() == n - 1) return mk_bool("(" + rk + " >= " + std::to_string(yes.front()) + "u)");
        if (contiguous) return mk_bool("(" + rk + " >= " + std::to_string(yes.front()) + "u && " + rk + " <= " + std::to_string(yes.back()) + "u)");
        // "!=" with a hit in the middle: complement of a single rank
        std::vector<int> no;
        for (int r = 0; r < n; ++r)
            if (!pred(r)) no.push_back(r);
        if (no.size() == 1) return mk_bool("(" + rk + " != " + std::to_string(no[0]) + "u)");
        bad("unsupported density comparison");
    }

    Val rand_vs_lit(int lane, const std::string& op, float lit) {
        *cx.rand_lanes |= 1u << lane;
        if (std::isnan(lit)) return mk_bool(op == "!=" ? "true" : "false");
        bool negate = (op == ">" || op == ">=");
        bool strict = (op == "<" || op == ">=");   // x >= p  ==  !(x < p)
        if (op == "==" || op == "!=") {
            *cx.lut_ok = false;
            return mk_bool("(SE_RANDF(" + std::to_string(lane) + ") " + op + " " + f32_lit(lit) + ")");
        }
        uint32_t U;
        bool all;
        bool any = rand_threshold(lit, strict, &U, &all);
        if (lane != 1) *cx.lut_ok = false;
        else if (any && !all) cx.y_thresholds->push_back(U);
        std::string base;
        if (!any) base = "false";
        else if (all) base = "true";
        else base = "(rnd.u[" + std::to_string(lane) + "] <= " + hex32(U) + ")";
        if (!negate) return mk_bool(base);
        if (base == "false") return mk_bool("true");
        if (base == "true") return mk_bool("false");
        return mk_bool("(!" + base + ")");
    }

    static std::string flip(const std::string& op) {
        if (op == "<") return ">";
        if (op == ">") return "<";
        if (op == "<=") return ">=";
        if (op == ">=") return "<=";
        return op;
    }

    Val compare(const Val& a, const std::string& op, const Val& b) {
        bool eqop = (op == "==" || op == "!=");
        if (a.type == VT::VecN || b.type == VT::VecN) {
            // GLSL: == / != on vectors compare the whole value and give one bool (e.g. `mat.emission.rgb != vec3(0.0)`,
            // the reference's own isLightObstacle test, operations.glsl:68-70)
            if (a.type != VT::VecN || b.type != VT::VecN || !eqop || a.comps.size() != b.comps.size())
                bad("vectors can only be compared with == / != against a vector of the same size");
            std::string all;
            for (size_t k = 0; k < a.comps.size(); ++k) all += (k ? " && " : "") + std::string("(") + a.comps[k] + " == " + b.comps[k] + ")";
            return mk_bool(op == "==" ? "(" + all + ")" : "(!(" + all + "))");
        }
        if (a.type == VT::Mat || b.type == VT::Mat) {
            if (a.type != VT::Mat || b.type != VT::Mat || !eqop) bad("materials can only be compared with == / != against materials");
            auto idc = [&](const Val& v) { return v.cell.empty() ? std::to_string(v.mat_id) + "u" : "SE_ID(" + v.cell + ")"; };
            return mk_bool("(" + idc(a) + " " + op + " " + idc(b) + ")");
        }
        if (a.type == VT::Bool && b.type == VT::Bool && eqop) return mk_bool("((" + a.code + ") " + op + " (" + b.code + "))");
        if (a.type == VT::Int && b.type == VT::Int) return mk_bool("(" + a.code + " " + op + " " + b.code + ")");
        if ((a.type != VT::Int && a.type != VT::Float) || (b.type != VT::Int && b.type != VT::Float)) bad("operands of '" + op + "' are not comparable");
        // float comparison (ints are converted, GLSL implicit conversion)
        if (a.is_density && b.is_density) {
            const char* m = op == "<" ? "SE_DENS_LT" : op == ">" ? "SE_DENS_GT" : op == "<=" ? "SE_DENS_LE" : op == ">=" ? "SE_DENS_GE" : op == "==" ? "SE_DENS_EQ" : "SE_DENS_NE";
            return mk_bool(std::string(m) + "(" + a.cell + ", " + b.cell + ")");
        }
        if (a.is_density && b.is_lit) return density_vs_lit(a.cell, op, (float)b.lit);
        if (b.is_density && a.is_lit) return density_vs_lit(b.cell, flip(op), (float)a.lit);
        if (a.is_rand && b.is_lit) return rand_vs_lit(a.lane, op, (float)b.lit);
        if (b.is_rand && a.is_lit) return rand_vs_lit(b.lane, flip(op), (float)a.lit);
        if (a.is_lit && b.is_lit) {
            float x = (float)a.lit, y = (float)b.lit;
            bool r = op == "<" ? x < y : op == ">" ? x > y : op == "<=" ? x <= y : op == ">=" ? x >= y : op == "==" ? x == y : x != y;
            return mk_bool(r ? "true" : "false");
        }
        return mk_bool("(" + as_float_code(a) + " " + op + " " + as_float_code(b) + ")");
    }

    Val arith(const Val& a, const std::string& op, const Val& b) {
        if ((a.type != VT::Int && a.type != VT::Float) || (b.type != VT::Int && b.type != VT::Float)) bad("arithmetic on non-numbers");
        if (a.type == VT::Int && b.type == VT::Int) return mk_int("(" + a.code + " " + op + " " + b.code + ")");
        if (op == "%") bad("'%' needs integer operands");
        // explicit round-to-nearest intrinsics: never contracted into an FMA (oracle is built -ffp-contract=off)
        const char* fn = op == "+" ? "__fadd_rn" : op == "-" ? "__fsub_rn" : op == "*" ? "__fmul_rn" : "__fdiv_rn";
        return mk_float(std::string(fn) + "(" + as_float_code(a) + ", " + as_float_code(b) + ")");
    }

    // ---- grammar (precedence of GLSL 4.30, lowest first): ?:  ||  &&  |  ^  &  == !=  < > <= >=  << >>  + -  * / %  unary ----
    // The reference passes conditions to the GLSL compiler verbatim, so anything a scalar GLSL expression may contain
    // can appear in a rule file; what has no bit-exact meaning (transcendentals, vectors) is refused with a clean error.
    Val parse_ternary() {
        Val c = parse_or();
        if (!lx.is_op("?")) return c;
        lx.next();
        Val a = parse_ternary();
        if (!lx.is_op(":")) bad("missing ':' of the conditional operator");
        lx.next();
        Val b = parse_ternary();
        const std::string cc = to_bool(c).code;
        if (a.type == VT::Bool && b.type == VT::Bool) return mk_bool("(" + cc + " ? " + a.code + " : " + b.code + ")");
        if (a.type == VT::Int && b.type == VT::Int) return mk_int("(" + cc + " ? " + a.code + " : " + b.code + ")");
        if (is_num(a) && is_num(b)) return mk_float("(" + cc + " ? " + as_float_code(a) + " : " + as_float_code(b) + ")");
        bad("the two branches of '?:' must both be booleans or both be numbers");
    }
    Val parse_or() {
        Val a = parse_and();
        while (lx.is_op("||")) { lx.next(); Val b = parse_and(); a = mk_bool("(" + to_bool(a).code + " || " + to_bool(b).code + ")"); }
        return a;
    }
    Val parse_and() {
        Val a = parse_bitor();
        while (lx.is_op("&&")) { lx.next(); Val b = parse_bitor(); a = mk_bool("(" + to_bool(a).code + " && " + to_bool(b).code + ")"); }
        return a;
    }
    Val int_op(const Val& a, const std::string& op, const Val& b) {
        if (a.type != VT::Int || b.type != VT::Int) bad("'" + op + "' needs integer operands");
        return mk_int("(" + a.code + " " + op + " " + b.code + ")");
    }
    Val parse_bitor() {
        Val a = parse_bitxor();
        while (lx.is_op("|")) { lx.next(); Val b = parse_bitxor(); a = int_op(a, "|", b); }
        return a;
    }
    Val parse_bitxor() {
        Val a = parse_bitand();
        while (lx.is_op("^")) { lx.next(); Val b = parse_bitand(); a = int_op(a, "^", b); }
        return a;
    }
    Val parse_bitand() {
        Val a = parse_eq();
        while (lx.is_op("&")) { lx.next(); Val b = parse_eq(); a = int_op(a, "&", b); }
        return a;
    }
    Val parse_eq() {
        Val a = parse_rel();
        while (lx.is_op("==") || lx.is_op("!=")) { std::string op = lx.tok; lx.next(); Val b = parse_rel(); a = compare(a, op, b); }
        return a;
    }
    Val parse_rel() {
        Val a = parse_shift();
        while (lx.is_op("<") || lx.is_op(">") || lx.is_op("<=") || lx.is_op(">=")) { std::string op = lx.tok; lx.next(); Val b = parse_shift(); a = compare(a, op, b); }
        return a;
    }
    Val parse_shift() {
        Val a = parse_add();
        while (lx.is_op("<<") || lx.is_op(">>")) { std::string op = lx.tok; lx.next(); Val b = parse_add(); a = int_op(a, op, b); }
        return a;
    }
    Val parse_add() {
        Val a = parse_mul();
        while (lx.is_op("+") || lx.is_op("-")) { std::string op = lx.tok; lx.next(); Val b = parse_mul(); a = arith(a, op, b); }
        return a;
    }
    Val parse_mul() {
        Val a = parse_unary();
        while (lx.is_op("*") || lx.is_op("/") || lx.is_op("%")) { std::string op = lx.tok; lx.next(); Val b = parse_unary(); a = arith(a, op, b); }
        return a;
    }
    Val parse_unary() {
        if (lx.is_op("!")) { lx.next(); Val a = parse_unary(); return mk_bool("(!" + to_bool(a).code + ")"); }
        if (lx.is_op("-")) {
            lx.next();
            Val a = parse_unary();
            if (a.is_lit && a.type == VT::Float) return mk_flit(-(float)a.lit);
            if (a.is_lit && a.type == VT::Int) return mk_ilit(-(long long)a.lit);
            if (a.type == VT::Int) return mk_int("(-" + a.code + ")");
            if (a.type == VT::Float) return mk_float("(-" + a.code + ")");
            bad("unary '-' on a non-number");
        }
        if (lx.is_op("+")) { lx.next(); return parse_unary(); }
        if (lx.is_op("~")) {
            lx.next();
            Val a = parse_unary();
            if (a.type != VT::Int) bad("'~' needs an integer operand");
            return mk_int("(~" + a.code + ")");
        }
        return parse_postfix();
    }
    Val parse_postfix() {
        Val v = parse_primary();
        while (lx.is_op(".")) {
            lx.next();
            if (lx.kind != Lexer::Ident) bad("expected a member name after '.'");
            std::string m = lx.tok;
            lx.next();
            v = member(v, m);
        }
        return v;
    }
    static int comp_index(const std::string& m) {
        if (m == "x" || m == "r") return 0;
        if (m == "y" || m == "g") return 1;
        if (m == "z" || m == "b") return 2;
        if (m == "w" || m == "a") return 3;
        return -1;
    }
    Val member(const Val& v, const std::string& m) {
        if (v.type == VT::Cell) {
            if (m == "mat") { Val r; r.type = VT::Mat; r.cell = v.code; return r; }
            bad("cells only have a '.mat' member here");
        }
        if (v.type == VT::Mat) {
            bool cst = v.cell.empty();
            if (m == "density") {
                if (cst) return mk_flit(cx.tb.density[v.mat_id]);
                Val r = mk_float("__ldg(&se_density_table[SE_ID(" + v.cell + ")])");
                r.is_density = true;
                r.cell = v.cell;
                return r;
            }
            if (m == "id") return cst ? mk_ilit(v.mat_id) : mk_int("((int)SE_ID(" + v.cell + "))");
            if (m == "type") return cst ? mk_ilit((cx.tb.fat[v.mat_id] >> 8) & 0xFF) : mk_int("((int)SE_TYPE(" + v.cell + "))");
            if (m == "color" || m == "emission") { Val r = v; r.type = VT::Vec4; r.vec_field = m; return r; }
            bad("unknown material member '" + m + "'");
        }
        if (v.type == VT::Vec4 && m.size() >= 2 && m.size() <= 4) {
            Val r;
            r.type = VT::VecN;
            for (char c : m) {
                int k = comp_index(std::string(1, c));
                if (k < 0) bad("unknown vector component '" + m + "'");
                if (v.cell.empty()) r.comps.push_back(f32_lit(v.vec_field == "color" ? cx.tb.color[v.mat_id][k] : cx.tb.emission[v.mat_id][k]));
                else r.comps.push_back("__ldg(&se_" + v.vec_field + "_table[SE_ID(" + v.cell + ") * 4 + " + std::to_string(k) + "])");
            }
            return r;
        }
        if (v.type == VT::Vec4) {
            int k = comp_index(m);
            if (k < 0) bad("unknown vector component '" + m + "'");
            if (v.cell.empty()) return mk_flit(v.vec_field == "color" ? cx.tb.color[v.mat_id][k] : cx.tb.emission[v.mat_id][k]);
            return mk_float("__ldg(&se_" + v.vec_field + "_table[SE_ID(" + v.cell + ") * 4 + " + std::to_string(k) + "])");
        }
        if (v.type == VT::RandVec) {
            int k = comp_index(m);
            if (k < 0) bad("unknown component of rand");
            *cx.rand_lanes |= 1u << k;
            Val r = mk_float("SE_RANDF(" + std::to_string(k) + ")");
            r.is_rand = true;
            r.lane = k;
            return r;
        }
        if (v.type == VT::PosVec) {
            *cx.lut_ok = false;
            if (m == "x") return mk_int("px");
            if (m == "y") return mk_int("py");
            bad("unknown component of pos");
        }
        bad("'." + m + "' applied to a value without members");
    }
    static bool is_num(const Val& v) { return v.type == VT::Int || v.type == VT::Float; }

    // Scalar GLSL built-ins with an exact meaning (GLSL 4.30 section 8.3 definitions, evaluated in f32 without
    // contraction, or on ints).  min(x,y) = y < x ? y : x;  max(x,y) = x < y ? y : x;  clamp = min(max(x,lo),hi);
    // mod(x,y) = x - y*floor(x/y);  fract(x) = x - floor(x);  step(e,x) = x < e ? 0 : 1;  int() truncates.
    static bool is_builtin(const std::string& id) {
        static const char* names[] = {"abs", "min", "max", "clamp", "mod", "floor", "ceil", "fract", "sign", "step", "sqrt", "float", "int"};
        for (auto* n : names)
            if (id == n) return true;
        return false;
    }
    Val call_builtin(const std::string& id, const std::vector<Val>& a) {
        auto need = [&](size_t n) { if (a.size() != n) bad(id + "() takes " + std::to_string(n) + " argument(s)"); };
        auto nums = [&]() { for (auto& v : a) if (!is_num(v)) bad(id + "() needs numeric arguments"); };
        auto all_int = [&]() { for (auto& v : a) if (v.type != VT::Int) return false; return true; };
        auto pick = [&](const Val& x, const Val& y, bool take_y_if_less_than_x) {   // min: y < x ? y : x   max: x < y ? y : x
            const bool ints = x.type == VT::Int && y.type == VT::Int;
            const std::string X = ints ? x.code : as_float_code(x), Y = ints ? y.code : as_float_code(y);
            const std::string c = take_y_if_less_than_x ? "(" + Y + " < " + X + " ? " + Y + " : " + X + ")" : "(" + X + " < " + Y + " ? " + Y + " : " + X + ")";
            return ints ? mk_int(c) : mk_float(c);
        };
        if (id == "float") {
            need(1);
            if (a[0].type == VT::Bool) return mk_float("(" + a[0].code + " ? 1.0f : 0.0f)");
            nums();
            return mk_float(as_float_code(a[0]));
        }
        if (id == "int") {
            need(1);
            if (a[0].type == VT::Bool) return mk_int("(" + a[0].code + " ? 1 : 0)");
            nums();
            return a[0].type == VT::Int ? mk_int(a[0].code) : mk_int("((int)(" + as_float_code(a[0]) + "))");
        }
        nums();
        if (id == "abs") {
            need(1);
            if (all_int()) return mk_int("(" + a[0].code + " < 0 ? (-" + a[0].code + ") : " + a[0].code + ")");
            return mk_float("fabsf(" + as_float_code(a[0]) + ")");
        }
        if (id == "min") { need(2); return pick(a[0], a[1], true); }
        if (id == "max") { need(2); return pick(a[0], a[1], false); }
        if (id == "clamp") { need(3); return pick(pick(a[0], a[1], false), a[2], true); }
        if (id == "sign") {
            need(1);
            if (all_int()) return mk_int("(" + a[0].code + " > 0 ? 1 : (" + a[0].code + " < 0 ? -1 : 0))");
            const std::string X = as_float_code(a[0]);
            return mk_float("(" + X + " > 0.0f ? 1.0f : (" + X + " < 0.0f ? -1.0f : 0.0f))");
        }
        if (id == "floor") { need(1); return mk_float("floorf(" + as_float_code(a[0]) + ")"); }
        if (id == "ceil") { need(1); return mk_float("ceilf(" + as_float_code(a[0]) + ")"); }
        if (id == "sqrt") { need(1); return mk_float("__fsqrt_rn(" + as_float_code(a[0]) + ")"); }
        if (id == "fract") { need(1); const std::string X = as_float_code(a[0]); return mk_float("__fsub_rn(" + X + ", floorf(" + X + "))"); }
        if (id == "step") { need(2); return mk_float("(" + as_float_code(a[1]) + " < " + as_float_code(a[0]) + " ? 0.0f : 1.0f)"); }
        if (id == "mod") {
            need(2);
            const std::string X = as_float_code(a[0]), Y = as_float_code(a[1]);
            return mk_float("__fsub_rn(" + X + ", __fmul_rn(" + Y + ", floorf(__fdiv_rn(" + X + ", " + Y + "))))");
        }
        bad("unknown function '" + id + "'");
    }

    Val parse_primary() {
        if (lx.is_op("(")) {
            lx.next();
            Val v = parse_ternary();
            if (!lx.is_op(")")) bad("missing ')'");
            lx.next();
            if (v.type == VT::Bool || v.type == VT::Int || v.type == VT::Float) {
                // keep the special forms (density / rand / literal) through redundant parentheses
                return v;
            }
            return v;
        }
        if (lx.kind == Lexer::Number) {
            std::string t = lx.tok;
            lx.next();
            bool is_hex = t.size() > 2 && t[0] == '0' && (t[1] == 'x' || t[1] == 'X');
            bool is_float = !is_hex && (t.find('.') != std::string::npos || t.find('e') != std::string::npos || t.find('E') != std::string::npos);
            std::string body = t;
            while (!body.empty() && (body.back() == 'f' || body.back() == 'F' || body.back() == 'u' || body.back() == 'U') && !is_hex) {
                if (body.back() == 'f' || body.back() == 'F') is_float = true;
                body.pop_back();
            }
            char* end = nullptr;
            if (is_float) {
                float f = std::strtof(body.c_str(), &end);
                if (!end || *end) bad("malformed number '" + t + "'");
                return mk_flit(f);
            }
            long long i = std::strtoll(body.c_str(), &end, 0);
            if (!end || (*end && *end != 'u' && *end != 'U')) bad("malformed number '" + t + "'");
            return mk_ilit(i);
        }
        if (lx.kind == Lexer::Ident) {
            std::string id = lx.tok;
            lx.next();
            if (id == "true" || id == "false") return mk_bool(id);
            if ((id == "vec2" || id == "vec3" || id == "vec4") && lx.is_op("(")) {
                const size_t n = (size_t)(id[3] - '0');
                lx.next();
                Val r;
                r.type = VT::VecN;
                size_t n_args = 0;
                if (!lx.is_op(")")) {
                    for (;;) {
                        Val a = parse_ternary();
                        ++n_args;
                        if (a.type == VT::VecN) r.comps.insert(r.comps.end(), a.comps.begin(), a.comps.end());
                        else if (is_num(a)) r.comps.push_back(as_float_code(a));
                        else bad(id + "() needs numbers or vectors");
                        if (lx.is_op(",")) { lx.next(); continue; }
                        break;
                    }
                }
                if (!lx.is_op(")")) bad("missing ')' after the arguments of " + id + "()");
                lx.next();
                if (n_args == 1 && r.comps.size() == 1) r.comps.assign(n, r.comps[0]);       // vec3(0.0): one scalar fills every component
                if (r.comps.size() != n) bad(id + "() needs 1 or " + std::to_string(n) + " components");
                return r;
            }
            if (is_builtin(id) && lx.is_op("(")) {
                lx.next();
                std::vector<Val> args;
                if (!lx.is_op(")")) {
                    for (;;) {
                        args.push_back(parse_ternary());
                        if (lx.is_op(",")) { lx.next(); continue; }
                        break;
                    }
                }
                if (!lx.is_op(")")) bad("missing ')' after the arguments of " + id + "()");
                lx.next();
                return call_builtin(id, args);
            }
            if (id == "rand") { Val v; v.type = VT::RandVec; return v; }
            if (id == "pos") { Val v; v.type = VT::PosVec; return v; }
            if (id == "frame") { *cx.lut_ok = false; return mk_int("frame"); }
            std::string reg = cell_reg(id);
            if (!reg.empty()) { Val v; v.type = VT::Cell; v.code = reg; return v; }
            if (id.rfind("MAT_", 0) == 0) {
                int m = find_material(id.substr(4));
                if (m < 0) throw not_found(id, cx.where);
                Val v; v.type = VT::Mat; v.mat_id = m; return v;
            }
            if (id.rfind("TYPE_", 0) == 0) {
                int t = find_type(id.substr(5));
                if (t < 0) throw not_found(id, cx.where);
                return mk_ilit(t);
            }
            if (id.rfind("isType_", 0) == 0) {
                int t = find_type(id.substr(7));
                if (t < 0) throw not_found(id.substr(7), cx.where + " -> isType_");
                if (!lx.is_op("(")) bad("isType_* needs a cell argument");
                lx.next();
                Val arg = parse_ternary();
                if (arg.type != VT::Cell) bad("isType_* needs a cell argument");
                if (!lx.is_op(")")) bad("missing ')'");
                lx.next();
                uint64_t mask = cx.tb.type_masks[t];
                char b[64];
                if (cx.tb.n_types <= 32) std::snprintf(b, sizeof b, "SE_ISTYPE32(%s, 0x%08xu)", arg.code.c_str(), (uint32_t)mask);
                else std::snprintf(b, sizeof b, "SE_ISTYPE64(%s, 0x%016llxull)", arg.code.c_str(), (unsigned long long)mask);
                return mk_bool(b);
            }
            throw not_found(id, cx.where);
        }
        bad("unexpected token '" + lx.tok + "'");
    }

    Val parse_all() {
        Val v = parse_ternary();
        if (lx.kind != Lexer::End) bad("unexpected trailing '" + lx.tok + "'");
        return v;
    }
};

std::string compile_condition(const std::string& text, Ctx& cx) {
    Parser p(text, cx);
    Val v = p.parse_all();
    if (v.type != VT::Bool) throw not_recognized(text + "  [condition is not boolean]", cx.where);
    return v.code;
}

// "swap(self, down);\nself = newCell(MAT_x, pos);"  ->  CUDA statements
std::string compile_actions(const std::string& text, Ctx& cx) {
    std::string out;
    size_t start = 0;
    Parser names("", cx);
    while (start < text.size()) {
        size_t end = text.find(';', start);
        if (end == std::string::npos) end = text.size();
        std::string st = text.substr(start, end - start);
        start = end + 1;
        size_t a = st.find_first_not_of(" \t\n\r");
        if (a == std::string::npos) continue;
        st = st.substr(a);
        names.text = st;
        if (st.rfind("swap(", 0) == 0) {
            size_t comma = st.find(", "), close = st.rfind(')');
            if (comma == std::string::npos || close == std::string::npos) throw not_recognized(st, cx.where);
            std::string c1 = names.cell_reg(st.substr(5, comma - 5)), c2 = names.cell_reg(st.substr(comma + 2, close - comma - 2));
            if (c1.empty() || c2.empty()) throw not_found(st, cx.where);
            out += "SE_SWAP(" + c1 + ", " + c2 + "); ";
        } else {
            static const std::string kSet = " = newCell(MAT_";
            size_t eq = st.find(kSet);
            size_t close = st.rfind(", pos)");
            if (eq == std::string::npos || close == std::string::npos) throw not_recognized(st, cx.where);
            std::string c1 = names.cell_reg(st.substr(0, eq));
            if (c1.empty()) throw not_found(st.substr(0, eq), cx.where);
            std::string mname = st.substr(eq + kSet.size(), close - (eq + kSet.size()));
            int m = names.find_material(mname);
            if (m < 0) throw not_found("MAT_" + mname, cx.where);
            out += c1 + " = " + hex32(cx.tb.fat[m]) + "; ";
        }
    }
    return out;
}

std::string emit_rule(const SandRule& r, Ctx& cx) {
    std::ostringstream o;
    o << "static __device__ __forceinline__ void se_rule_" << r.name
      << "(unsigned& s, unsigned& r, unsigned& d, unsigned& dr, const SeRand& rnd, const int px, const int py, const int frame) {\n";
    if (r.has_precondition) {
        cx.where = "rules/" + r.name + "/precondition";
        o << "    if (!(" << compile_condition(r.precondition, cx) << ")) return;\n";
    }
    // rules.rs:47-73: conditions, actions and probabilities are consumed from the front in lockstep;
    // an action left over without a condition is an unconditional trailing else.
    size_t n_if = r.if_conds.size(), n_do = r.do_actions.size();
    size_t depth = 0;
    std::string indent = "    ";
    for (size_t k = 0; k < n_do; ++k) {
        cx.where = "rules/" + r.name + "/do[" + std::to_string(k) + "]";
        std::string act = compile_actions(r.do_actions[k], cx);
        if (k < n_if) {
            cx.where = "rules/" + r.name + "/if[" + std::to_string(k) + "]";
            // The reference pastes the probability test in front of the condition WITHOUT parentheses
            // (rules.rs:56-68: "rand.y <= {p} && {cond}"), so with a top-level `||` in the condition the
            // probability only guards the first disjunct.  Compile the very same text to keep that.
            float p = r.probabilities[k];
            std::string text = r.if_conds[k];
            if (p != 1.0f) text = "rand.y <= " + f32_display(p) + " && " + text;
            std::string cond = compile_condition(text, cx);
            o << indent << "if (" << cond << ") { " << act << "} else {\n";
            indent += "    ";
            ++depth;
        } else {
            o << indent << act << "\n";
            break;   // get_func_logic returns after the first condition-less action
        }
    }
    for (size_t k = 0; k < depth; ++k) {
        indent.resize(indent.size() - 4);
        o << indent << "}\n";
    }
    o << "    (void)px; (void)py; (void)frame; (void)rnd;\n}\n\n";
    return o.str();
}

}  // namespace

bool rand_threshold(float p, bool strict, uint32_t* U, bool* all) {
    auto pred = [&](uint32_t u) {
        float r = (float)u / 4294967296.0f;   // vec4(hash4i(x)) / float(0xffffffffU), math.glsl:74-79
        return strict ? (r < p) : (r <= p);
    };
    *all = false;
    *U = 0;
    if (std::isnan(p) || !pred(0u)) return false;
    if (pred(0xFFFFFFFFu)) { *all = true; *U = 0xFFFFFFFFu; return true; }
    uint64_t lo = 0, hi = 0xFFFFFFFFull;   // pred(lo) true, pred(hi) false
    while (hi - lo > 1) {
        uint64_t mid = (lo + hi) / 2;
        if (pred((uint32_t)mid)) lo = mid; else hi = mid;
    }
    *U = (uint32_t)lo;
    return true;
}

CompiledRules compile_rules(const ParsingResult& parsed) {
    CompiledRules out;
    out.parsed = parsed;
    MaterialTables& tb = out.tables;
    tb.n_materials = (int)parsed.materials.size();
    tb.n_types = (int)parsed.types.size();
    if (tb.n_materials > 255) throw unsupported("at most 255 materials are supported (material ids are one byte; 255 is reserved for unknown ids)");
    if (tb.n_types > 64) throw unsupported("at most 64 types are supported");

    for (auto& t : parsed.types) tb.type_names.push_back(t.name);
    tb.type_masks.assign(tb.n_types, 0);
    for (auto& t : parsed.types) {
        uint64_t mask = 1ull << t.id;
        for (auto& c : t.children)
            for (auto& t2 : parsed.types)
                if (t2.name == c) { mask |= 1ull << t2.id; break; }
        tb.type_masks[t.id] = mask;
    }
    // density ranks
    std::vector<float> ranks;
    for (auto& m : parsed.materials) {
        if (!std::isfinite(m.density)) throw unsupported("material '" + m.name + "' has a non-finite density");
        bool seen = false;
        for (float v : ranks) if (v == m.density) { seen = true; break; }
        if (!seen) ranks.push_back(m.density);
    }
    std::sort(ranks.begin(), ranks.end());
    auto rank_of = [&](float d) { for (size_t k = 0; k < ranks.size(); ++k) if (ranks[k] == d) return (uint32_t)k; return 0u; };

    std::memset(tb.fat, 0, sizeof tb.fat);
    std::memset(tb.density, 0, sizeof tb.density);
    std::memset(tb.emission, 0, sizeof tb.emission);
    std::memset(tb.color, 0, sizeof tb.color);
    std::memset(tb.selectable, 0, sizeof tb.selectable);
    for (auto& m : parsed.materials) {
        int type_id = -1;
        for (auto& t : parsed.types) if (t.name == m.mattype) { type_id = t.id; break; }
        if (type_id < 0) {
            if (m.mattype == "EMPTY") type_id = 0;   // "EMPTY" is always a known type name (parser.rs:116)
            else throw not_found(m.mattype, "materials/" + m.name + "/type");
        }
        bool em = m.emission[0] != 0.f || m.emission[1] != 0.f || m.emission[2] != 0.f;
        bool is_empty_type = (tb.type_masks[0] >> type_id) & 1ull;
        uint32_t flags = 0;
        if (type_id == 1 || type_id == 2) flags |= SE_F_NOSWAP;
        if (em) flags |= SE_F_EMISSIVE;
        if (!em && !is_empty_type) flags |= SE_F_OBSTACLE;
        tb.fat[m.id] = (uint32_t)m.id | ((uint32_t)type_id << 8) | flags | (rank_of(m.density) << 24);
        tb.density[m.id] = m.density;
        for (int k = 0; k < 4; ++k) { tb.emission[m.id][k] = m.emission[k]; tb.color[m.id][k] = m.color[k]; }
        tb.selectable[m.id] = m.selectable ? 1 : 0;
        tb.material_names.push_back(m.name);
    }
    for (int id = tb.n_materials; id < 256; ++id) {   // unknown id => MAT_NULL (gen/materials.glsl:79-86)
        tb.fat[id] = tb.fat[1];
        tb.density[id] = tb.density[1];
        for (int k = 0; k < 4; ++k) { tb.emission[id][k] = tb.emission[1][k]; tb.color[id][k] = tb.color[1][k]; }
    }

    // ---- rule functions ----
    std::ostringstream fn, mir, left, right;
    out.rand_lanes = 1;   // lane x: mirror decision (falling_sand.glsl:86)
    bool lut_ok = true;
    std::vector<uint32_t> y_thr;
    for (auto& r : parsed.rules) {
        if (!r.used) continue;
        SandRuleType et = r.effective_type();
        Ctx cx{tb, parsed, ranks, et == SandRuleType::Left, "rules/" + r.name, &out.rand_lanes, &y_thr, &lut_ok};
        fn << emit_rule(r, cx);
        std::string call = "    se_rule_" + r.name + "(s, r, d, dr, rnd, px, py, frame);\n";
        if (et == SandRuleType::Mirrored) mir << call;
        else if (et == SandRuleType::Left) { left << call; out.have_left = true; }
        else { right << call; out.have_right = true; }
        ++out.n_used_rules;
    }

    std::sort(y_thr.begin(), y_thr.end());
    y_thr.erase(std::unique(y_thr.begin(), y_thr.end()), y_thr.end());
    // Transition-table modes (kernels/sand_kernels.cuh, "transition tables"):
    //   1  the table (N^4 32-bit entries per view) fits beside the tiles in shared memory
    //   2  the table lives in global memory (L2 / HBM resident): rule sets with up to 127 materials
    // Left/Right rules break the mirror symmetry a single table relies on: such sets get one table per view
    // (unmirrored / mirrored evaluation); the global-memory mode always keeps both views (no slow path at all).
    const bool lr = out.have_left || out.have_right;
    const size_t n4 = (size_t)tb.n_materials * tb.n_materials * tb.n_materials * tb.n_materials;
    int mode = 0, tables = lr ? 2 : 1;
    if (lut_ok && tb.n_materials <= 127 && y_thr.size() <= 15) {
        if ((size_t)tables * n4 * 4 <= 60u * 1024u) mode = 1;
        else { mode = 2; tables = 2; }
    }
    if (const char* fm = std::getenv("SE_LUT_FORCE_MODE")) {       // experiments / tests: "0" (generated code only) or "2"
        const int v = std::atoi(fm);
        if (v == 0) mode = 0;
        if (v == 2 && mode == 1) { mode = 2; tables = 2; }
    }
    out.lut_eligible = mode != 0;
    out.lut_mode = mode;
    out.lut_tables = out.lut_eligible ? tables : 1;
    out.lut_thresholds = y_thr;

    std::ostringstream h;
    h << "// GENERATED by sandengine_b200 (CUDA C back end of the rule language). Do not edit.\n";
    h << "#define SE_LUT_ELIGIBLE " << (out.lut_eligible ? 1 : 0) << "\n";
    h << "#define SE_LUT_MODE " << out.lut_mode << "\n";
    h << "#define SE_LUT_TWO_TABLES " << (out.lut_tables == 2 ? 1 : 0) << "\n";
    h << "#define SE_LUT_NCLS " << (y_thr.size() + 1) << "\n";
    h << "// class c of a block = number of thresholds its rand.y hash lane exceeds (u1 > U_i)\n";
    h << "#define SE_LUT_CLASS(u1) (0u";
    for (uint32_t t : y_thr) h << " + ((u1) > " << hex32(t) << " ? 1u : 0u)";
    h << ")\n";
    h << "__device__ const unsigned se_lut_thresholds[16] = {";
    for (size_t k = 0; k < 16; ++k) h << (k < y_thr.size() ? hex32(y_thr[k]) : std::string("0xffffffffu")) << (k < 15 ? ", " : "");
    h << "};\n";
    h << "#define SE_N_MATERIALS " << tb.n_materials << "\n#define SE_N_TYPES " << tb.n_types << "\n";
    h << "#define SE_RAND_LANES " << out.rand_lanes << "u\n";
    h << "#define SE_HAVE_LEFT_RULES " << (out.have_left ? 1 : 0) << "\n#define SE_HAVE_RIGHT_RULES " << (out.have_right ? 1 : 0) << "\n";
    {
        uint32_t U; bool all;
        rand_threshold(0.5f, true, &U, &all);   // shouldMirror = rand.x < 0.5 (falling_sand.glsl:86)
        h << "#define SE_MIRROR_UMAX " << hex32(U) << "\n";
    }
    h << "#define SE_FAT_EMPTY " << hex32(tb.fat[0]) << "\n#define SE_FAT_NULL " << hex32(tb.fat[1]) << "\n#define SE_FAT_WALL " << hex32(tb.fat[2]) << "\n";
    for (int k = 0; k < tb.n_materials; ++k) h << "// material " << k << ": " << tb.material_names[k] << "  fat=" << hex32(tb.fat[k]) << "\n";
    h << "__device__ const unsigned se_fat_table[256] = {";
    for (int k = 0; k < 256; ++k) h << (k % 8 == 0 ? "\n    " : " ") << hex32(tb.fat[k]) << ",";
    h << "\n};\n__device__ const float se_density_table[256] = {";
    for (int k = 0; k < 256; ++k) h << (k % 4 == 0 ? "\n    " : " ") << f32_lit(tb.density[k]) << ",";
    h << "\n};\n__device__ const float se_emission_table[256 * 4] = {";
    for (int k = 0; k < 256; ++k) { h << "\n   "; for (int c = 0; c < 4; ++c) h << " " << f32_lit(tb.emission[k][c]) << ","; }
    h << "\n};\n__device__ const float se_color_table[256 * 4] = {";
    for (int k = 0; k < 256; ++k) { h << "\n   "; for (int c = 0; c < 4; ++c) h << " " << f32_lit(tb.color[k][c]) << ","; }
    h << "\n};\n\n";
    h << fn.str();
    const char* sig = "(unsigned& s, unsigned& r, unsigned& d, unsigned& dr, const SeRand& rnd, const int px, const int py, const int frame) {\n";
    h << "static __device__ __forceinline__ void se_apply_mirrored" << sig << mir.str() << "    (void)s; (void)r; (void)d; (void)dr; (void)rnd; (void)px; (void)py; (void)frame;\n}\n";
    h << "// Left rules are called in the mirrored view: r = LEFT, dr = DOWNLEFT.\n";
    h << "static __device__ __forceinline__ void se_apply_left" << sig << left.str() << "    (void)s; (void)r; (void)d; (void)dr; (void)rnd; (void)px; (void)py; (void)frame;\n}\n";
    h << "static __device__ __forceinline__ void se_apply_right" << sig << right.str() << "    (void)s; (void)r; (void)d; (void)dr; (void)rnd; (void)px; (void)py; (void)frame;\n}\n";
    out.cuda_header = h.str();
    return out;
}

}  // namespace se
