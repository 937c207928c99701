"""ORACLE (test infrastructure, not product code) -- Python restatement of `sandengine-lang`.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` leg may
import this module.  The product front end is the C++ code under `sandengine_b200/csrc/lang/`; it
shares no code with this file.

What is restated (all paths relative to /root/reference):
  * `parse_string`                      sandengine-lang/src/parser.rs:93-152
  * `preparse_keys`, mapping checks     parser.rs:156-187
  * `extract_vec4`                      parser.rs:192-251
  * `parse_rules`/`parse_conditionals`  parser/rules.rs:105-334
  * `parse_global_scope`, `parse_do`    parser/rules.rs:338-420
  * `parse_types` (+ child / inherited-rule propagation)   parser/types.rs:51-210
  * `parse_materials`                   parser/materials.rs:51-201
  * GLSL text templates                 rules.rs:47-98, types.rs:30-45, materials.rs:30-47,
                                        sandengine-lang/src/lib.rs:17-148
The GLSL emitter exists only to PIN this restatement: its output for `data/materials.yaml` must be
byte-identical to the reference's checked-in `shaders/compute/gen/{materials,rules}.glsl`
(tests/test_oracle_lang.py; SHA-256 of those files is committed under tests/golden/).

On top of the restated parser this file adds `emit_c_rules()`, which rewrites the same (GLSL-shaped)
rule text into plain C for `oracle/sand_oracle.c` -- the CPU restatement of the compute shader.

LEFT rules: the reference's Left/Right classification is dead code
(`do_actions[0].contains("LEFT")` runs after the text was lower-cased, rules.rs:157 vs :292) and a
Left rule could not compile as GLSL.  Definition used by this oracle AND by the product (parity for
it is unpinned by the reference): a `mirrored: false` rule whose if/do text mentions LEFT or DOWNLEFT
is a Left rule; it runs only when shouldMirror is true, in the mirrored view (guarded swaps as in
falling_sand.glsl:86-90, rule with left = 2nd cell and downleft = 4th cell, guarded un-swap).  A rule
that mixes LEFT* with RIGHT*, or a mirrored rule that mentions LEFT*, is rejected (NotRecognized).
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import yaml

GLOBAL_CELLNAMES = ["SELF", "LEFT", "RIGHT", "DOWN", "DOWNRIGHT", "DOWNLEFT"]  # parser.rs:27-34
DEFAULT_VAL_MIRRORED = True          # parser.rs:37
DEFAULT_VAL_PRECONDITION = True      # parser.rs:38
DEFAULT_VAL_PROBABILITY = np.float32(1.0)  # parser.rs:39


class ParsingErr(Exception):
    """parser.rs:42-73.  `kind` is one of MissingField / InvalidType / NotFound / NotRecognized."""

    def __init__(self, kind: str, msg: str):
        super().__init__(f"({kind}) {msg}")
        self.kind = kind


def _missing(field_name, missing_in):
    return ParsingErr("MissingField", f"Mandatory field '{field_name}' is missing in '{missing_in}'")


def _invalid(wrong, missing_in, expected):
    return ParsingErr("InvalidType", f"The type of the field '{wrong!r}' inside of '{missing_in}' is invalid. Expected: '{expected}'")


def _notfound(missing, missing_in):
    return ParsingErr("NotFound", f"The name '{missing}' (in '{missing_in}') was not found. Make sure it was defined before referencing it.")


def _notrecog(unrecog, missing_in):
    return ParsingErr("NotRecognized", f"The expression '{unrecog}' (in '{missing_in}') was not recognized as valid syntax. Please check it is valid.")


def _require_usable(rule):
    """A rule that a type or a material refers to must be executable (module docstring: the Left-rule definition)."""
    if rule.left_conflict is not None:
        raise _notrecog(rule.left_conflict, f"rules/{rule.name} (LEFT in a mirrored rule, or LEFT mixed with RIGHT)")


# ----------------------------------------------------------------------------------------------
# YAML loading with serde_yaml-0.9-like (YAML 1.2 core schema) scalar resolution.  PyYAML is
# YAML 1.1 (yes/no/on/off are bools, 1e3 is a string); override the implicit resolvers.
# ----------------------------------------------------------------------------------------------
class _Loader(yaml.SafeLoader):
    pass


_Loader.yaml_implicit_resolvers = {}
_Loader.add_implicit_resolver("tag:yaml.org,2002:bool", re.compile(r"^(?:true|True|TRUE|false|False|FALSE)$"), list("tTfF"))
_Loader.add_implicit_resolver("tag:yaml.org,2002:null", re.compile(r"^(?:~|null|Null|NULL|)$"), ["~", "n", "N", ""])
# serde_yaml 0.9.34 (the reference's Cargo.lock), de.rs parse_unsigned_int / parse_negative_int / digits_but_not_number: signed decimal,
# hex, octal and binary integers; a decimal with a leading zero ("007", "-012") is a STRING (and not a float either)
_Loader.add_implicit_resolver("tag:yaml.org,2002:int",
                              re.compile(r"^[-+]?(?:0|[1-9][0-9]*|0o[0-7]+|0x[0-9a-fA-F]+|0b[01]+)$"), list("-+0123456789"))
_Loader.add_implicit_resolver(
    "tag:yaml.org,2002:float",
    re.compile(r"^(?![-+]?0[0-9]+$)(?:[-+]?(?:\.[0-9]+|[0-9]+(?:\.[0-9]*)?)(?:[eE][-+]?[0-9]+)?|[-+]?\.(?:inf|Inf|INF)|\.(?:nan|NaN|NAN))$"),
    list("-+0123456789."),
)


def _construct_int(loader, node):
    s = loader.construct_scalar(node)
    sign = -1 if s.startswith("-") else 1
    body = s.lstrip("+-")
    for prefix, radix in (("0o", 8), ("0x", 16), ("0b", 2)):
        if body.startswith(prefix):
            return sign * int(body[2:], radix)
    return int(s)


def _construct_float(loader, node):
    s = loader.construct_scalar(node).lower()
    if s.endswith(".inf"):
        return float("-inf") if s.startswith("-") else float("inf")
    if s == ".nan":
        return float("nan")
    return float(s)


_Loader.add_constructor("tag:yaml.org,2002:int", _construct_int)
_Loader.add_constructor("tag:yaml.org,2002:float", _construct_float)


def _construct_mapping_no_duplicates(loader, node):
    """serde_yaml 0.9 refuses a mapping with a repeated key ("duplicate entry with key ...") when it builds the `Value`
    that parser.rs:95-98 deserialises into; PyYAML would silently keep the last one.  Found by scripts/diff_frontends.py."""
    loader.flatten_mapping(node)
    seen = set()
    for key_node, _ in node.value:
        key = loader.construct_object(key_node, deep=True)
        try:
            if key in seen:
                raise yaml.constructor.ConstructorError(None, None, f"duplicate entry with key {key!r}", key_node.start_mark)
            seen.add(key)
        except TypeError:      # unhashable key (a sequence or mapping as key): left to PyYAML
            pass
    return loader.construct_mapping(node, deep=True)


_Loader.add_constructor("tag:yaml.org,2002:map", _construct_mapping_no_duplicates)


def _as_str(v):
    return v if isinstance(v, str) else None


def _as_bool(v):
    return v if isinstance(v, bool) else None


def _as_f64(v):  # serde_yaml Value::as_f64: any number
    if isinstance(v, bool):
        return None
    if isinstance(v, (int, float)):
        return float(v)
    return None


def _as_u64(v):
    if isinstance(v, bool):
        return None
    if isinstance(v, int) and v >= 0:
        return v
    return None


def _get(v, key):  # serde_yaml Value::get on non-mappings returns None
    if isinstance(v, dict):
        return v.get(key)
    return None


def _has(v, key):
    return isinstance(v, dict) and key in v


def f32_display(x) -> str:
    """Rust `Display` for f32 (shortest round-trip digits, never scientific): materials.rs:32-45, rules.rs:60."""
    x = np.float32(x)
    if np.isnan(x):
        return "NaN"
    if np.isinf(x):
        return "inf" if x > 0 else "-inf"
    s = np.format_float_positional(x, unique=True, trim="-")
    if s == "-0":
        return "-0"
    return s


# ----------------------------------------------------------------------------------------------
@dataclass
class SandRule:  # rules.rs:25-44
    name: str
    ruletype: str  # "Mirrored" | "Left" | "Right"  (reference classification, rules.rs:154-163)
    if_conds: List[str]
    do_actions: List[str]
    probabilities: List[np.float32]
    mirror: bool
    precondition: Optional[str]
    used: bool = False
    mentions_left: bool = False   # oracle/product definition of a Left rule (module docstring)
    mentions_right: bool = False
    left_conflict: Optional[str] = None   # raw text when LEFT sits in a mirrored rule or is mixed with RIGHT: an error once the rule is used

    @staticmethod
    def _func_logic(if_conds, do_actions, probabilities, indent_lvl):  # rules.rs:47-73
        if not if_conds and not do_actions:
            return ""
        if not if_conds and do_actions:
            return do_actions[0]
        ind1 = " " * (indent_lvl * 4)
        ind2 = " " * ((indent_lvl + 1) * 4)
        p = probabilities[0]
        prob = "" if p == DEFAULT_VAL_PROBABILITY else f"rand.y <= {f32_display(p)} && "
        inner = SandRule._func_logic(if_conds[1:], do_actions[1:], probabilities[1:], indent_lvl + 1)
        return f"{ind1}if ({prob}{if_conds[0]}) {{\n{ind2}{do_actions[0]}\n{ind1}}} else {{\n{inner}\n{ind1}}}"

    def get_glsl_code(self) -> str:  # rules.rs:75-98
        celldir = "left" if self.ruletype == "Left" else "right"
        precond = "" if self.precondition is None else f"    if (!({self.precondition})) {{\n        return;\n    }}\n"
        body = SandRule._func_logic(list(self.if_conds), list(self.do_actions), list(self.probabilities), 1)
        return (f"void rule_{self.name} (inout Cell self, inout Cell {celldir}, inout Cell down, inout Cell downright, "
                f"vec4 rand, ivec2 pos) {{\n{precond}{body}\n}}")

    @property
    def effective_type(self) -> str:
        """Mirrored / Left / Right under the definition in the module docstring."""
        if self.mirror:
            return "Mirrored"
        return "Left" if self.mentions_left else "Right"


@dataclass
class SandType:  # types.rs:12-26
    id: int
    name: str
    inherits: str = ""
    children: List[str] = field(default_factory=list)
    base_rules: List[str] = field(default_factory=list)

    def get_checker_func(self) -> str:  # types.rs:30-39
        tc = f"return cell.mat.type == TYPE_{self.name}"
        for c in self.children:
            tc += f" || cell.mat.type == TYPE_{c}"
        return f"bool isType_{self.name}(Cell cell) {{\n    {tc};\n}}\n\n"

    def get_glsl_code(self) -> str:  # types.rs:41-45
        return f"#define TYPE_{self.name} {self.id}\n\n"


@dataclass
class SandMaterial:  # materials.rs:10-29
    id: int
    name: str
    mattype: str
    color: List[np.float32]
    emission: List[np.float32]
    selectable: bool
    density: np.float32
    extra_rules: List[str] = field(default_factory=list)

    def get_glsl_code(self) -> str:  # materials.rs:30-47
        c, e = self.color, self.emission
        return (f"#define MAT_{self.name} Material({self.id}, vec4({f32_display(c[0])}, {f32_display(c[1])}, "
                f"{f32_display(c[2])}, {f32_display(c[3])}), {f32_display(self.density)}, vec4({f32_display(e[0])}, "
                f"{f32_display(e[1])}, {f32_display(e[2])}, {f32_display(e[3])}), TYPE_{self.mattype})\n")


@dataclass
class ParsingResult:  # parser.rs:84-89
    rules: List[SandRule]
    types: List[SandType]
    materials: List[SandMaterial]


# ----------------------------------------------------------------------------------------------
def parse_global_scope(s: str) -> str:  # rules.rs:338-351 (ordered literal replaces)
    s = s.replace(" or ", " || ")
    s = s.replace(" and ", " && ")
    s = s.replace("not ", " !")
    s = s.replace("empty", "MAT_EMPTY")
    s = s.replace("SELF", "self")
    s = s.replace("RIGHT", "right")
    s = s.replace("LEFT", "left")
    s = s.replace("DOWN", "down")
    s = s.replace("DOWNRIGHT", "downright")
    s = s.replace("DOWNLEFT", "downleft")
    return s


_SWAP_RE = re.compile(r"SWAP (\w+) (\w+)")
_SET_RE = re.compile(r"SET (\w+) (\w+)")
_MAT_RE = re.compile(r"\w*.mat\s*(==|!=)\s*(\w*)")
_TYPE_RE = re.compile(r"isType_(\w*)\(\w*\)")


def parse_do(parent: str, do_str: str) -> str:  # rules.rs:355-420
    out = ""
    found = False
    m = _SWAP_RE.search(do_str)
    if m:
        found = True
        for g in (m.group(1), m.group(2)):
            if g not in GLOBAL_CELLNAMES:
                raise _notfound(g, parent)
        out += f"swap({m.group(1)}, {m.group(2)});\n"
    m = _SET_RE.search(do_str)
    if m:
        found = True
        if m.group(1) not in GLOBAL_CELLNAMES:
            raise _notfound(m.group(1), parent)
        out += f"{m.group(1)} = newCell(MAT_{m.group(2)}, pos);\n"
    if not found:
        raise _notrecog(do_str, parent)
    return out


def _parse_conditionals(parent, parent_is_else, parent_path, if_conds, do_actions, probabilities, type_names, material_names, raw_text):
    # rules.rs:207-334
    if_cond = _get(parent, "if") if _has(parent, "if") else None
    has_if = _has(parent, "if")
    if not has_if and not parent_is_else:
        raise _missing("if", parent_path)
    elif has_if:
        parent_path = f"{parent_path}/if"
        s = _as_str(if_cond)
        if s is None:
            raise _invalid("if", parent_path, "string")
        raw_text.append(s)
        s = parse_global_scope(s)
        for cap in [m.group(2) for m in _MAT_RE.finditer(s)]:
            for mname in material_names:
                if mname == cap:
                    s = s.replace(mname, f"MAT_{mname}")
                    break
            else:
                raise _notfound(cap, parent_path)
        for cap in [m.group(1) for m in _TYPE_RE.finditer(s)]:
            if cap not in type_names:
                raise _notfound(cap, f"{parent_path} -> isType_")
        if_conds.append(s)

    if not _has(parent, "do"):
        raise _missing("do", parent_path)
    do_action = _get(parent, "do")
    do_parent_path = f"{parent_path}/do"
    do_string = ""
    if isinstance(do_action, str):
        raw_text.append(do_action)
        do_string = parse_global_scope(parse_do(do_parent_path, do_action))
    if isinstance(do_action, list):
        for item in do_action:
            if isinstance(item, str):
                raw_text.append(item)
                do_string += parse_global_scope(parse_do(do_parent_path, item))
    do_string = do_string.rstrip()
    do_actions.append(do_string)

    if _has(parent, "probability"):
        p = _as_f64(_get(parent, "probability"))
        if p is None:
            raise _invalid("probability", parent_path, "float (0.0 to 1.0)")
        probabilities.append(np.float32(p))
    else:
        probabilities.append(DEFAULT_VAL_PROBABILITY)

    if _has(parent, "else"):
        _parse_conditionals(_get(parent, "else"), True, f"{parent_path}/else", if_conds, do_actions, probabilities,
                            type_names, material_names, raw_text)


def _parse_rules(rules, type_names, material_names):  # rules.rs:105-203
    out = []
    for key, val in rules.items():
        name = _as_str(key)
        if name is None:
            raise _invalid(key, "rules", "string")
        if_conds, do_actions, probs, raw_text = [], [], [], []
        _parse_conditionals(val, False, f"rules/{name}", if_conds, do_actions, probs, type_names, material_names, raw_text)
        if _has(val, "mirrored"):
            is_mirrored = _as_bool(_get(val, "mirrored"))
            if is_mirrored is None:
                raise _invalid("mirrored", f"rules/{name}", "bool (true/false)")
        else:
            is_mirrored = DEFAULT_VAL_MIRRORED
        if is_mirrored:
            ruletype = "Mirrored"
        else:
            ruletype = "Left" if "LEFT" in do_actions[0] else "Right"   # always Right: text is lower-case by now
        if _has(val, "precondition"):
            pre = _as_bool(_get(val, "precondition"))
            if pre is None:
                raise _invalid("precondition", f"rules/{name}", "bool (true/false)")
        else:
            pre = DEFAULT_VAL_PRECONDITION
        joined = " ".join(raw_text)
        idents = set(re.findall(r"[A-Za-z0-9_]+", joined))      # whole identifiers: a material LEFTOVER is not a cell name
        mentions_left = bool(idents & {"LEFT", "DOWNLEFT"})
        mentions_right = bool(idents & {"RIGHT", "DOWNRIGHT"})
        rule = SandRule(name, ruletype, if_conds, do_actions, probs, is_mirrored, "" if pre else None,
                        False, mentions_left, mentions_right)
        # an error only once a type or a material uses the rule (the reference emits used rules only)
        rule.left_conflict = joined if mentions_left and (is_mirrored or mentions_right) else None
        out.append(rule)
    return out


def _update_rule_precondition(rule: SandRule, typename: str):  # types.rs:78-88
    if rule.precondition is not None:
        if rule.precondition == "":
            rule.precondition = f"isType_{typename}(self)"
        else:
            rule.precondition = f"{rule.precondition} || isType_{typename}(self)"


def _add_child_to_type(parent_name, childname, types, depth=0):  # types.rs:186-200
    if depth > len(types) + 1:   # the reference overflows its stack on an inheritance cycle; a clean error here
        raise _invalid("inherits", f"types/{childname}", "an acyclic chain of parent types")
    pp = ""
    for t in types:
        if t.name == parent_name:
            t.children.append(childname)
            pp = t.inherits
            break
    if pp:
        _add_child_to_type(pp, childname, types, depth + 1)


def _get_parents_rules(all_types, cur, depth=0):  # types.rs:202-210
    if not cur.inherits:
        return []
    if depth > len(all_types) + 1:
        raise _invalid("inherits", f"types/{cur.name}", "an acyclic chain of parent types")
    parent = next((t for t in all_types if t.name == cur.inherits), None)
    if parent is None:
        # reference: `.unwrap()` on None panics (types.rs:206); surfaced as NotFound here.
        raise _notfound(cur.inherits, f"types/{cur.name}/inherits")
    return list(parent.base_rules) + _get_parents_rules(all_types, parent, depth + 1)


def _parse_types(types, rules, rule_names, type_names):  # types.rs:51-182
    structs = [SandType(0, "EMPTY"), SandType(1, "NULL"), SandType(2, "WALL")]
    idx = len(structs)
    for key, val in types.items():
        name = _as_str(key)
        if name is None:
            raise _invalid(key, "types", "string")
        parent = ""
        if _has(val, "inherits"):
            parent = _as_str(_get(val, "inherits"))
            if parent is None:
                raise _invalid("inherits", f"types/{name}", "string")
            for tn in type_names:
                if tn == parent:
                    _add_child_to_type(parent, name, structs)
        base_rules = []
        if _has(val, "base_rules"):
            b = _get(val, "base_rules")
            if not isinstance(b, list):
                raise _invalid("base_rules", f"types/{name}", "sequence (array, '[...]')")
            for br in b:
                rn = _as_str(br)
                if rn is None:
                    raise _invalid("base_rules", f"types/{name}", "string")
                if rn not in rule_names:
                    raise _notfound(rn, f"types/{name}/base_rules")
                rule = next(r for r in rules if r.name == rn)
                _require_usable(rule)
                rule.used = True
                _update_rule_precondition(rule, name)
                base_rules.append(rn)
        structs.append(SandType(idx, name, parent, [], base_rules))
        idx += 1
    for st in list(structs):
        if not st.inherits:
            continue
        for rn in _get_parents_rules(structs, st):
            rule = next(r for r in rules if r.name == rn)
            _update_rule_precondition(rule, st.name)
    return structs


def _extract_vec4(data, parent_name, field_name, default, mandatory):  # parser.rs:192-251
    missing_in = f"materials/{parent_name}/{field_name}"
    if not _has(data, field_name):
        if mandatory:
            raise _missing(field_name, missing_in)
        return [np.float32(x) for x in default]
    v = _get(data, field_name)
    vec = [np.float32(x) for x in default]
    if not isinstance(v, list):
        raise _invalid(field_name, missing_in, "color")
    if len(v) == 0 or len(v) > 4 or len(v) < 3:
        raise _invalid(field_name, missing_in, "color")
    for i, comp in enumerate(v):
        u = _as_u64(comp)
        if u is not None and 0 < u <= 255:
            vec[i] = np.float32(np.float32(u) / np.float32(255.0))
            continue
        f = _as_f64(comp)
        if f is not None:
            if 1.0 < f <= 255.0:
                vec[i] = np.float32(np.float32(f) / np.float32(255.0))
                continue
            if 0.0 <= f <= 1.0:
                vec[i] = np.float32(f)
                continue
        raise _invalid(field_name, missing_in, "color")
    return vec


def _parse_materials(materials, rules, type_names):  # materials.rs:51-201
    f = np.float32
    structs = [
        SandMaterial(0, "EMPTY", "EMPTY", [f(0), f(0), f(0), f(0)], [f(0)] * 4, True, f(1.0)),
        SandMaterial(1, "NULL", "NULL", [f(1), f(0), f(1), f(1)], [f(0)] * 4, False, f(0.0)),
        SandMaterial(2, "WALL", "WALL", [f(0.1), f(0.2), f(0.3), f(1.0)], [f(0)] * 4, False, f(9999.0)),
    ]
    idx = len(structs)
    for key, val in materials.items():
        name = _as_str(key)
        if name is None:
            raise _invalid(key, "materials", "string")
        if not _has(val, "type"):
            raise _missing("type", f"materials/{name}")
        mattype = _as_str(_get(val, "type"))
        if mattype is None:
            raise _invalid("type", f"materials/{name}", "string")
        if mattype not in type_names:
            raise _notfound(mattype, f"materials/{name}/type")
        color = _extract_vec4(val, name, "color", [1.0, 0.0, 1.0, 1.0], True)
        emission = _extract_vec4(val, name, "emission", [0.0, 0.0, 0.0, 0.0], False)
        if _has(val, "selectable"):
            sel = _as_bool(_get(val, "selectable"))
            selectable = sel if sel is not None else False
        else:
            selectable = True
        if not _has(val, "density"):
            raise _missing("density", f"materials/{name}")
        d = _as_f64(_get(val, "density"))
        if d is None:
            raise _invalid("density", f"materials/{name}/density", "float (0.0 to 1.0)")
        density = np.float32(d)
        extra_rules = []
        if _has(val, "extra_rules"):
            ex = _get(val, "extra_rules")
            if not isinstance(ex, list):
                raise _invalid("extra_rules", f"materials/{name}", "sequence (array, '[...]')")
            for er in ex:
                ern = _as_str(er)
                if ern is None:
                    raise _invalid("extra_rules", f"materials/{name}", "string")
                for r in rules:
                    if r.name == ern:
                        _require_usable(r)
                        r.used = True
                        if r.precondition is not None:
                            if r.precondition == "":
                                r.precondition = f"self.mat == MAT_{name}"
                            else:
                                r.precondition = f"{r.precondition} || self.mat == MAT_{name}"
                        extra_rules.append(ern)
        structs.append(SandMaterial(idx, name, mattype, color, emission, selectable, density, extra_rules))
        idx += 1
    return structs


def _check_mapping(data, keyname):  # parser.rs:174-187
    if not _has(data, keyname):
        raise _missing(keyname, "Root/ Base level of YAML file")
    v = data[keyname]
    if not isinstance(v, dict):
        raise _invalid(v, "Root/ Base level of YAML file", "mapping (dictionary-like)")
    return v


def _preparse_keys(mapping, err_name):  # parser.rs:156-169
    names = []
    for k in mapping.keys():
        if not isinstance(k, str):
            raise _invalid(k, err_name, "string")
        names.append(k)
    return names


def parse_string(text: str) -> ParsingResult:  # parser.rs:93-152
    try:
        data = yaml.load(text, Loader=_Loader)
    except yaml.YAMLError as e:
        raise ParsingErr("Yaml", str(e))
    raw_rules = _check_mapping(data, "rules")
    raw_types = _check_mapping(data, "types")
    raw_materials = _check_mapping(data, "materials")
    rule_names = _preparse_keys(raw_rules, "rules")
    type_names = _preparse_keys(raw_types, "types") + ["EMPTY"]
    material_names = _preparse_keys(raw_materials, "materials") + ["EMPTY"]
    try:
        rules = _parse_rules(raw_rules, type_names, material_names)
    except ParsingErr as e:
        raise ParsingErr(e.kind, f"Error while parsing rules: '{e}'")
    try:
        types = _parse_types(raw_types, rules, rule_names, type_names)
    except ParsingErr as e:
        raise ParsingErr(e.kind, f"Error while parsing types: '{e}'")
    try:
        materials = _parse_materials(raw_materials, rules, type_names)
    except ParsingErr as e:
        raise ParsingErr(e.kind, f"Error while parsing materials: '{e}'")
    return ParsingResult(rules, types, materials)


def parse_path(path) -> ParsingResult:  # sandengine-lang/src/lib.rs:10-13
    with open(path, "r") as fh:
        return parse_string(fh.read())


# ----------------------------------------------------------------------------------------------
# GLSL emitter (golden check only)                     sandengine-lang/src/lib.rs:17-148
# ----------------------------------------------------------------------------------------------
def emit_glsl_materials(res: ParsingResult) -> str:
    s = ""
    for t in res.types:
        s += t.get_glsl_code()
    for t in res.types:
        s += t.get_checker_func()
    s += "\n"
    lst = ""
    n = len(res.materials)
    for i, m in enumerate(res.materials):
        s += m.get_glsl_code()
        lst += f"        MAT_{m.name}" + ("" if i == n - 1 else ",\n")
    s += (f"\nMaterial[{n}] materials() {{\n    Material allMaterials[{n}] = {{\n{lst}\n    }};\n    return allMaterials;\n}}\n\n"
          "Material getMaterialFromID(int id) {\n    for (int i = 0; i < materials().length(); i++) {\n"
          "        if (id == materials()[i].id) {\n            return materials()[i];\n        };\n    };\n"
          "    return MAT_NULL;\n}\n\n")
    return s


def emit_glsl_rules(res: ParsingResult, patched_left: bool = False) -> str:
    """gen/rules.glsl.  patched_left=False reproduces the reference byte for byte (incl. its broken Left rules, which do
    not compile as GLSL).  patched_left=True is the MINIMAL PATCH of the reference's emitter that realises the LEFT
    definition of the module docstring inside the reference's unmodified shader template:
      * a Left rule (effective_type) gets the parameters (self, left, down, downleft) instead of (self, right, down,
        downright)                                                             [rules.rs:77-80]
      * applyLeftRules -- which simulate() calls with the UN-mirrored cells when shouldMirror (falling_sand.glsl:99-103)
        -- re-enters the mirrored view with the same guarded swaps as :86-90, calls the rules, and leaves it again
        (a guarded un-swap followed by the same guarded swap is the identity, so this equals "before the un-mirror")
                                                                               [sandengine-lang/src/lib.rs:88-98]"""
    funcs, mir, left, right = "", "", "", ""
    for r in res.rules:
        if not r.used:
            continue
        kind = r.effective_type if patched_left else r.ruletype
        code = r.get_glsl_code()
        if patched_left and kind == "Left":
            code = code.replace("(inout Cell self, inout Cell right, inout Cell down, inout Cell downright,",
                                "(inout Cell self, inout Cell left, inout Cell down, inout Cell downleft,", 1)
        funcs += f"{code}\n\n"
        if kind == "Mirrored":
            mir += f"rule_{r.name}(self, right, down, downright, rand, pos);\n"
        elif kind == "Left":
            left += (f"rule_{r.name}(self, right, down, downright, rand, pos);\n" if patched_left
                     else f"rule_{r.name}(self, left, down, downright, rand, pos);\n")
        else:
            right += f"rule_{r.name}(self, right, down, downright, rand, pos);\n"
    if patched_left and left:
        left = "swap(self, right);\nswap(down, downright);\n" + left + "swap(self, right);\nswap(down, downright);\n"
    hdr = ("(\n    inout Cell self,\n    inout Cell right,\n    inout Cell down,\n    inout Cell downright,\n"
           "    vec4 rand,\n    ivec2 pos) {\n    ")
    return (f"\n// =============== RULES ===============\n{funcs}\n\n\n// =============== CALLERS ===============\n"
            f"void applyMirroredRules{hdr}{mir.rstrip()}\n}}\n\n\nvoid applyLeftRules{hdr}{left.rstrip()}\n}}\n\n"
            f"void applyRightRules{hdr}{right.rstrip()}\n}}")


# ----------------------------------------------------------------------------------------------
# C emitter for oracle/sand_oracle.c -- the same rule text, mechanically rewritten:
#   `X.mat == MAT_y`      -> `(X.mat.id == MAT_y.id)`         (GLSL struct equality; ids are unique)
#   `swap(a, b);`         -> `swap_cells(&a, &b);`            (guarded swap, operations.glsl:16-23)
#   `a = newCell(M, pos)` -> unchanged (C function returning a Cell)
# Cells are passed by pointer and re-bound to local lvalues through macros so that the rule text
# itself stays what the reference would have emitted.
# ----------------------------------------------------------------------------------------------
_MATCMP_RE = re.compile(r"(\w+)\.mat\s*(==|!=)\s*(MAT_\w+)")
_SWAPCALL_RE = re.compile(r"swap\((\w+), (\w+)\);")
# GLSL float literals are f32; a bare `0.1` in C would be a double and change `rand.y <= 0.1`.
_FLOATLIT_RE = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)(?![\w.])")


# GLSL built-ins a condition may call (the reference hands conditions to the GLSL compiler verbatim): renamed to the
# type-generic se_* helpers of sand_oracle.c; `float(x)` / `int(x)` constructor syntax is not C.
_BUILTIN_RE = re.compile(r"\b(abs|min|max|clamp|mod|floor|ceil|fract|sign|step|sqrt|float|int)\(")


# `X.mat.emission.rgb != vec3(0.0)`: GLSL compares vectors as a whole (one bool); written out per component for C
_VECCMP_RE = re.compile(r"(\w+)\.mat\.(color|emission)\.([rgbaxyzw]{2,4})\s*(==|!=)\s*vec([234])\(([^()]*)\)")


def _vec_compare_to_c(m) -> str:
    cell, fld, swz, op, n, args = m.group(1), m.group(2), m.group(3), m.group(4), int(m.group(5)), [a.strip() for a in m.group(6).split(",")]
    if len(args) == 1:
        args = args * n
    if len(args) != n or len(swz) != n:
        return m.group(0)                      # left as it is: the C compiler rejects it, like the GLSL compiler would
    body = " && ".join(f"({cell}.mat.{fld}.{c} == {a})" for c, a in zip(swz, args))
    return f"({body})" if op == "==" else f"(!({body}))"


def _to_c(text: str) -> str:
    text = _VECCMP_RE.sub(_vec_compare_to_c, text)
    text = _BUILTIN_RE.sub(lambda m: f"se_{m.group(1)}(", text)
    text = _MATCMP_RE.sub(lambda m: f"({m.group(1)}.mat.id {m.group(2)} {m.group(3)}.id)", text)
    text = _SWAPCALL_RE.sub(lambda m: f"swap_cells(&{m.group(1)}, &{m.group(2)});", text)
    text = _FLOATLIT_RE.sub(lambda m: m.group(1) + "f", text)
    return text


def emit_c_rules(res: ParsingResult) -> str:
    """Generated half of the C oracle (plays the role gen/materials.glsl + gen/rules.glsl play for the shader)."""
    o = ["/* GENERATED by oracle/oracle_lang.py -- oracle (test infrastructure), do not edit. */\n"]
    for t in res.types:
        o.append(f"#define TYPE_{t.name} {t.id}\n")
    for t in res.types:
        tc = f"cell.mat.type == TYPE_{t.name}" + "".join(f" || cell.mat.type == TYPE_{c}" for c in t.children)
        o.append(f"static inline int isType_{t.name}(Cell cell) {{ return {tc}; }}\n")
    o.append(f"#define N_MATERIALS {len(res.materials)}\n")
    o.append("static const Material MATERIALS[N_MATERIALS] = {\n")
    for m in res.materials:
        c, e = m.color, m.emission
        fl = lambda x: repr(float(np.float32(x))) + "f" if "." in repr(float(np.float32(x))) or "e" in repr(float(np.float32(x))) else repr(float(np.float32(x))) + ".0f"
        o.append(f"  {{ {m.id}, {{{{{fl(c[0])}, {fl(c[1])}, {fl(c[2])}, {fl(c[3])}}}}}, {fl(m.density)}, "
                 f"{{{{{fl(e[0])}, {fl(e[1])}, {fl(e[2])}, {fl(e[3])}}}}}, TYPE_{m.mattype} }},\n")
    o.append("};\n")
    for m in res.materials:
        o.append(f"#define MAT_{m.name} (MATERIALS[{m.id}])\n")
    o.append("\n")
    mir, left, right = [], [], []
    for r in res.rules:
        if not r.used:
            continue
        et = r.effective_type
        second, fourth = ("left", "downleft") if et == "Left" else ("right", "downright")
        pre = "" if r.precondition is None else f"    if (!({_to_c(r.precondition)})) {{ return; }}\n"
        body = _to_c(SandRule._func_logic(list(r.if_conds), list(r.do_actions), list(r.probabilities), 1))
        o.append(f"static void rule_{r.name}(Cell* p_self, Cell* p_second, Cell* p_down, Cell* p_fourth, const float* rand4, ivec2 pos) {{\n"
                 f"#define self (*p_self)\n#define {second} (*p_second)\n#define down (*p_down)\n#define {fourth} (*p_fourth)\n"
                 f"    const struct {{ float x, y, z, w; }} rand = {{ rand4[0], rand4[1], rand4[2], rand4[3] }}; (void)rand;\n"
                 f"{pre}{body}\n"
                 f"#undef self\n#undef {second}\n#undef down\n#undef {fourth}\n}}\n\n")
        call = f"    rule_{r.name}(self, right, down, downright, rand, pos);\n"
        (mir if et == "Mirrored" else left if et == "Left" else right).append(call)
    sig = "(Cell* self, Cell* right, Cell* down, Cell* downright, const float* rand, ivec2 pos)"
    o.append(f"static void applyMirroredRules{sig} {{\n{''.join(mir)}    (void)self; (void)right; (void)down; (void)downright; (void)rand; (void)pos;\n}}\n")
    o.append(f"/* Left rules: called in the MIRRORED view (see oracle_lang.py docstring, SURVEY 8a P3). */\n"
             f"static void applyLeftRules{sig} {{\n{''.join(left)}    (void)self; (void)right; (void)down; (void)downright; (void)rand; (void)pos;\n}}\n")
    o.append(f"static void applyRightRules{sig} {{\n{''.join(right)}    (void)self; (void)right; (void)down; (void)downright; (void)rand; (void)pos;\n}}\n")
    o.append(f"#define HAVE_LEFT_RULES {1 if left else 0}\n")
    return "".join(o)
