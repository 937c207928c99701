"""Generator of the synthetic 64-material rule set of BASELINE.json configs[4]:
fixed-width collision-free names (m03.., t03.., r00..), a type tree of depth >= 6 with parents defined
before children, >= 24 rules mixing mirrored, non-mirrored RIGHT and non-mirrored LEFT rules, several
probabilities, nested else chains, SET and SWAP.  Deterministic in `seed`.

Names are chosen so that the reference's global substring replaces cannot corrupt them
(rules.rs:246-260, 338-351): no name is a substring of another identifier, none contains
SELF/RIGHT/LEFT/DOWN/empty/" or "/" and "/"not ", and a condition never compares the same material twice.
"""
from __future__ import annotations

from .grids import hashi
import numpy as np


def _rng(seed: int):
    state = {"c": np.uint32(seed * 2654435761 & 0xFFFFFFFF)}

    def nxt(n: int) -> int:
        state["c"] = np.uint32((int(state["c"]) + 0x9E3779B9) & 0xFFFFFFFF)
        return int(hashi(np.array([state["c"]], dtype=np.uint32))[0]) % n
    return nxt


def synthetic_rule_set(n_materials: int = 64, n_rules: int = 28, seed: int = 5, kinds=("mirrored", "mirrored", "right", "left")):
    """Returns (yaml_text, ids, mix): ids maps material names to ids, mix is a DEFAULT_MIX-style tuple for grids.
    `kinds` is the cycle of rule kinds; ("mirrored",) with n_materials <= 12 gives a transition-table-eligible set."""
    rnd = _rng(seed)
    n_user_mats = n_materials - 3
    n_types = 14
    # type tree: t03..t08 is a chain of depth 6 (t03 <- t04 <- ... <- t08); the rest hang off random earlier types
    types = []
    for k in range(n_types):
        name = f"t{k + 3:02d}"
        if k == 0:
            parent = None
        elif k < 6:
            parent = f"t{k + 2:02d}"
        else:
            parent = f"t{3 + rnd(k):02d}" if rnd(3) else None
        types.append({"name": name, "parent": parent, "base_rules": []})
    mats = []
    for k in range(n_user_mats):
        name = f"m{k + 3:02d}"
        t = types[rnd(n_types)]["name"]
        density = [0.05, 0.3, 0.8, 1.1, 1.3, 1.5, 1.5, 2.0, 2.5, 3.0, 4.0, 6.0][rnd(12)] + 0.001 * rnd(5)
        em = None
        if rnd(10) == 0:
            em = [round(0.1 * (1 + rnd(9)), 1), round(0.1 * rnd(10), 1), round(0.1 * rnd(10), 1), 0.9]
        mats.append({"name": name, "type": t, "density": density, "color": [20 + rnd(230), 20 + rnd(230), 20 + rnd(230)],
                     "emission": em, "extra_rules": [], "selectable": rnd(8) != 0})
    probs = [1.0, 1.0, 0.5, 0.25, 0.1, 0.03, 0.004, 0.75]
    rules = []
    for k in range(n_rules):
        name = f"r{k:02d}"
        kind = kinds[k % len(kinds)]
        side, dside = ("LEFT", "DOWNLEFT") if kind == "left" else ("RIGHT", "DOWNRIGHT")
        tA, tB = types[rnd(n_types)]["name"], types[rnd(n_types)]["name"]
        mA, mB = mats[rnd(n_user_mats)]["name"], mats[rnd(n_user_mats)]["name"]
        if mA == mB:
            mB = mats[(int(mA[1:]) - 3 + 1) % n_user_mats]["name"]
        form = k % 7
        r = {"name": name, "kind": kind, "precondition": (k % 5 != 0)}
        if form == 0:     # fall / slide chain
            r["if"] = "DOWN.mat.density < SELF.mat.density"
            r["do"] = "SWAP SELF DOWN"
            r["else"] = {"if": f"{side}.mat.density < SELF.mat.density and {dside}.mat.density < SELF.mat.density",
                         "probability": probs[rnd(len(probs))], "do": f"SWAP SELF {dside}"}
        elif form == 1:   # sideways flow
            r["if"] = f"isType_{tA}(SELF) and {side}.mat.density < SELF.mat.density"
            r["do"] = f"SWAP SELF {side}"
            r["else"] = {"if": f"isType_{tA}(DOWN) and {dside}.mat.density < DOWN.mat.density", "do": f"SWAP DOWN {dside}",
                         "else": {"if": f"isType_EMPTY({side}) and not isType_{tB}(SELF)", "probability": 0.25, "do": f"SWAP SELF {side}"}}
        elif form == 2:   # rise
            r["if"] = f"isType_{tA}(DOWN) and not isType_{tB}(SELF) and DOWN.mat.density < SELF.mat.density"
            r["do"] = "SWAP DOWN SELF"
        elif form == 3:   # reaction: SET
            r["if"] = f"SELF.mat == {mA} and {side}.mat == {mB}"
            r["probability"] = probs[2 + rnd(6)]
            r["do"] = f"SET SELF {mB}"
            r["precondition"] = False
        elif form == 4:   # growth into empty space
            r["if"] = f"isType_EMPTY(SELF) and DOWN.mat == {mA} and {dside}.mat != {mB}"
            r["probability"] = probs[4 + rnd(3)]
            r["do"] = f"SET SELF {mA}"
            r["precondition"] = False
        elif form == 5:   # decay with trailing unconditional else
            r["if"] = f"isType_{tA}(SELF) and isType_EMPTY(DOWN)"
            r["probability"] = probs[3 + rnd(4)]
            r["do"] = ["SET SELF EMPTY"]
            r["else"] = {"if": f"SELF.mat.density > 1.2 and {side}.mat.density <= 1.0", "probability": 0.5,
                         "do": f"SWAP SELF {side}"}
        else:             # density literal + type id compare
            r["if"] = f"SELF.mat.type != TYPE_{tB} and DOWN.mat.density >= 1.0 and DOWN.mat.density < SELF.mat.density"
            r["do"] = "SWAP SELF DOWN"
            r["probability"] = probs[rnd(len(probs))]
        rules.append(r)
        # attach: alternately as a base rule of a type or an extra rule of a material
        if k % 2 == 0:
            types[rnd(n_types)]["base_rules"].append(name)
        else:
            mats[rnd(n_user_mats)]["extra_rules"].append(name)
    # make sure every rule is used at least once (already true) and each top type has something that moves
    types[0]["base_rules"].append(rules[0]["name"]) if rules[0]["name"] not in types[0]["base_rules"] else None

    def emit_cond(d, ind):
        out = []
        for key in ("if", "probability"):
            if key in d:
                out.append(f"{ind}{key}: {d[key]}")
        if isinstance(d["do"], list):
            out.append(f"{ind}do:")
            out += [f"{ind}  - {a}" for a in d["do"]]
        else:
            out.append(f"{ind}do: {d['do']}")
        if "else" in d:
            out.append(f"{ind}else:")
            out += emit_cond(d["else"], ind + "  ")
        return out

    y = ["# synthetic rule set generated by sandengine_b200.synth_rules (seed %d)" % seed, "rules:"]
    for r in rules:
        y.append(f"  {r['name']}:")
        y.append(f"    mirrored: {'true' if r['kind'] == 'mirrored' else 'false'}")
        if not r["precondition"]:
            y.append("    precondition: false")
        y += emit_cond(r, "    ")
    y.append("types:")
    for t in types:
        y.append(f"  {t['name']}:")
        if t["parent"]:
            y.append(f"    inherits: {t['parent']}")
        if t["base_rules"]:
            y.append(f"    base_rules: [{', '.join(t['base_rules'])}]")
    y.append("materials:")
    for m in mats:
        y.append(f"  {m['name']}:")
        y.append(f"    type: {m['type']}")
        y.append(f"    density: {m['density']:.3f}")
        y.append(f"    color: [{m['color'][0]}, {m['color'][1]}, {m['color'][2]}]")
        if m["emission"]:
            y.append(f"    emission: [{m['emission'][0]}, {m['emission'][1]}, {m['emission'][2]}, {m['emission'][3]}]")
        if not m["selectable"]:
            y.append("    selectable: false")
        if m["extra_rules"]:
            y.append(f"    extra_rules: [{', '.join(m['extra_rules'])}]")
    text = "\n".join(y) + "\n"
    ids = {"EMPTY": 0}
    ids.update({m["name"]: k + 3 for k, m in enumerate(mats)})
    frac = 0.55 / n_user_mats
    mix = (("EMPTY", 0.45),) + tuple((m["name"], frac) for m in mats)
    return text, ids, mix
