"""ORACLE (test infrastructure) -- an independent pure-Python restatement of the block update, for SMALL grids.

Purpose: triangulate oracle/sand_oracle.c.  The C oracle executes rule text that oracle_lang.emit_c_rules()
rewrote into C; this module never sees that C: it evaluates the reference-shaped GLSL condition strings directly
(`||`/`&&`/`!` mapped to Python, `eval` over Cell/Material objects with f32 densities) and interprets the
if/do/else structure from the parsed fields (rules.rs:47-73).  Two differently built evaluators agreeing on
the default, the "rich" (Left/Right) and the synthetic 64-material rule sets is the strongest pin available,
given that the reference ships no state-level fixture and its GLSL cannot run here.

Follows /root/reference/shaders/compute/gen/falling_sand.glsl: hash :58-79,115-120; swap :371-378;
getMargolusOffset :380-389; getCell :400-412; simulate :676-733.  Lighting / modifications are not restated
here (the C oracle's lighting is pinned by the survey's lighting KAT).
"""
from __future__ import annotations

import re
from dataclasses import dataclass

import numpy as np

from . import oracle_lang as L

M32 = 0xFFFFFFFF


def hashi(x: int) -> int:  # :58-66
    x &= M32
    x ^= x >> 16
    x = (x * 0x7FEB352D) & M32
    x ^= x >> 15
    x = (x * 0x846CA68B) & M32
    x ^= x >> 16
    return x


def hash43(px: int, py: int, pz: int):  # :115-120 (uvec3(ivec) keeps the bits of negative coordinates)
    x = ((px & M32) * 461 + (py & M32) * 2131 + (pz & M32) * 2131 * 2131) & M32
    lanes = [hashi(x * m) for m in (213, 2131, 21313, 213132)]
    denom = np.float32(0xFFFFFFFF)     # rounds to 2^32
    return [np.float32(u) / denom for u in lanes]


@dataclass(frozen=True, eq=False)
class Material:  # :211-218 -- struct equality compares every member; ids are unique, so identity of the id
    id: int
    color: object
    density: np.float32
    emission: object
    type: int

    def __eq__(self, other):
        return self.id == other.id

    def __ne__(self, other):
        return self.id != other.id

    def __hash__(self):
        return self.id


class Cell:  # :221-224
    __slots__ = ("mat",)

    def __init__(self, mat):
        self.mat = mat


class _Vec4:
    def __init__(self, v):
        self.x, self.y, self.z, self.w = v
        self.r, self.g, self.b, self.a = v


class _Pos:
    def __init__(self, x, y):
        self.x, self.y = x, y


def _to_python(cond: str) -> str:
    s = cond.replace("||", " or ").replace("&&", " and ")
    s = re.sub(r"!(?!=)", " not ", s)
    return s.strip()


class PyOracle:
    def __init__(self, yaml_text: str):
        res = L.parse_string(yaml_text)
        self.types = {t.name: t.id for t in res.types}
        self.mats = [Material(m.id, _Vec4([np.float32(x) for x in m.color]), np.float32(m.density), _Vec4([np.float32(x) for x in m.emission]),
                              self.types.get(m.mattype, 0)) for m in res.materials]
        self.env = {f"MAT_{m.name}": self.mats[m.id] for m in res.materials}
        self.env.update({f"TYPE_{t.name}": t.id for t in res.types})
        for t in res.types:
            ids = frozenset([t.id] + [self.types[c] for c in t.children])   # types.rs:30-39
            self.env[f"isType_{t.name}"] = (lambda ids: (lambda cell: cell.mat.type in ids))(ids)
        self.TYPE_WALL, self.TYPE_NULL = self.types["WALL"], self.types["NULL"]
        self.rules = {"Mirrored": [], "Left": [], "Right": []}
        for r in res.rules:
            if not r.used:
                continue
            pre = compile(_to_python(r.precondition), f"<pre {r.name}>", "eval") if r.precondition is not None else None
            conds = [compile(_to_python(c), f"<if {r.name}>", "eval") for c in r.if_conds]
            self.rules[r.effective_type].append((r, pre, conds))

    # :371-378
    def _swap(self, q, a, b):
        ta, tb = q[a].mat.type, q[b].mat.type
        if ta in (self.TYPE_WALL, self.TYPE_NULL) or tb in (self.TYPE_WALL, self.TYPE_NULL):
            return
        q[a], q[b] = q[b], q[a]

    def _run_actions(self, text, q, names):
        for st in text.split(";"):
            st = st.strip()
            if not st:
                continue
            m = re.fullmatch(r"swap\((\w+), (\w+)\)", st)
            if m:
                self._swap(q, names[m.group(1)], names[m.group(2)])
                continue
            m = re.fullmatch(r"(\w+) = newCell\((MAT_\w+), pos\)", st)
            assert m, st
            q[names[m.group(1)]] = Cell(self.env[m.group(2)])

    def _apply(self, kind, q, rand, pos):
        names = {"self": 0, "down": 2}
        names.update({"left": 1, "downleft": 3} if kind == "Left" else {"right": 1, "downright": 3})
        for r, pre, conds in self.rules[kind]:
            def scope():
                e = dict(self.env)
                e.update({n: q[i] for n, i in names.items()})
                e["rand"], e["pos"], e["frame"] = rand, pos, self._frame
                return e
            if pre is not None and not eval(pre, {}, scope()):
                continue
            # rules.rs:47-73: conditions/actions/probabilities consumed from the front in lock-step
            for k, action in enumerate(r.do_actions):
                if k < len(conds):
                    p = r.probabilities[k]
                    # the reference pastes `rand.y <= p && cond` without parentheses: `and` binds tighter than `or`
                    text = _to_python(r.if_conds[k])
                    if p != np.float32(1.0):
                        text = f"rand.y <= _p and {text}"
                    e = scope()
                    e["_p"] = np.float32(L.f32_display(p))
                    if eval(text, {}, e):
                        self._run_actions(action, q, names)
                        break
                else:
                    self._run_actions(action, q, names)
                    break

    def step(self, cells: np.ndarray, frame: int) -> np.ndarray:
        """One dispatch (frame = value after the host increment), per-block form; returns the new grid."""
        H, W = cells.shape
        out = cells.copy()
        self._frame = frame
        f = frame % 4
        ox, oy = {1: (1, 1), 2: (0, 1), 3: (1, 0), 0: (0, 0)}[f]   # :380-389
        for y0 in range(-oy, H, 2):
            for x0 in range(-ox, W, 2):
                q = []
                for dy, dx in ((0, 0), (0, 1), (1, 0), (1, 1)):
                    x, y = x0 + dx, y0 + dy
                    if x < 0 or x >= W or y < 0 or y >= H:
                        q.append(Cell(self.mats[2]))                 # WALL, :400-404
                    else:
                        i = int(cells[y, x])
                        q.append(Cell(self.mats[i] if 0 <= i < len(self.mats) else self.mats[1]))
                if all(c.mat.id == 0 for c in q):                    # :692-694
                    continue
                r = hash43(x0, y0, frame)
                rand, pos = _Vec4(r), _Pos(x0, y0)
                mirror = bool(r[0] < np.float32(0.5))                # :701
                if mirror:
                    self._swap(q, 0, 1); self._swap(q, 2, 3)
                self._apply("Mirrored", q, rand, pos)
                if mirror:
                    self._apply("Left", q, rand, pos)                # definition: oracle_lang.py docstring
                    self._swap(q, 0, 1); self._swap(q, 2, 3)
                else:
                    self._apply("Right", q, rand, pos)
                for k, (dy, dx) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
                    x, y = x0 + dx, y0 + dy
                    if 0 <= x < W and 0 <= y < H:
                        out[y, x] = q[k].mat.id
        return out

    def run(self, cells, frame, n_steps):
        cells = np.ascontiguousarray(cells, dtype=np.uint32).copy()
        for _ in range(n_steps):
            frame += 1
            cells = np.zeros_like(cells) if frame == 1 else self.step(cells, frame)
        return cells, frame
