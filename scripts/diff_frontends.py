"""Tooling (CPU): differential fuzz of the two independent restatements of the reference's parser -- the C++ front end
behind the ABI and oracle/oracle_lang.py -- over line-level mutants of the rule files.  For every mutant both must
either fail with the same ParsingErr class or succeed with byte-identical gen/materials.glsl + gen/rules.glsl text.
   python scripts/diff_frontends.py <seed> <n_mutants> [indicators]
`indicators` adds a mutation that drops YAML indicator characters ({ } [ ] : ? # & * ! | > % @ `) into the middle of values.
That mode is exploratory, not a gate: ~0.3 % of its mutants get a different ERROR CLASS from the two YAML readers (`? ` /
`- ` / `, ` indicators inside flow collections, a bare `!` tag, `]#x`) -- malformed documents all, no valid rule file among
them; see DESIGN.md section 7.
Mutations keep the YAML shape simple (the C++ reader is a YAML-1.2-core subset): delete / duplicate / swap lines, replace a
scalar by another token of the same file, rename an identifier, tweak a number, toggle a bool."""
import random
import re
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
import sandengine_b200 as se  # noqa: E402
import yaml_cases as Y  # noqa: E402
from oracle import oracle_lang  # noqa: E402

INDICATORS = "indicators" in sys.argv[3:]
SOURCES = [(REPO / "data" / "materials.yaml").read_text(), Y.BASE_OK, Y.RICH_YAML, Y.EXPR_YAML, Y.FUNC_YAML]
TOKENS = ["SELF", "DOWN", "RIGHT", "LEFT", "DOWNRIGHT", "DOWNLEFT", "SWAP", "SET", "EMPTY", "empty", "and", "or", "not ", "true", "false",
          "0.5", "1", "1.0", "255", "-1", "mirrored", "probability", "precondition", "if", "do", "else", "inherits", "base_rules",
          "extra_rules", "type", "color", "density", "emission", "selectable"]


def mutate(text, rng):
    lines = text.split("\n")
    for _ in range(rng.randint(1, 3)):
        k = rng.randrange(15)
        if k == 13 and not INDICATORS:
            k = 14
        i = rng.randrange(len(lines))
        if k == 0 and len(lines) > 3:
            del lines[i]
        elif k == 1:
            lines.insert(i, lines[i])
        elif k == 2:
            j = rng.randrange(len(lines)); lines[i], lines[j] = lines[j], lines[i]
        elif k == 3:                                   # replace the value after "key: "
            m = re.match(r"^(\s*[\w-]+:\s*)(\S.*)$", lines[i])
            if m:
                other = rng.choice([ln for ln in lines if ":" in ln] or [lines[i]])
                lines[i] = m.group(1) + other.split(":", 1)[1].strip()
        elif k == 4:                                   # swap one word for a token
            words = re.findall(r"[A-Za-z_]\w*", lines[i])
            if words:
                lines[i] = lines[i].replace(rng.choice(words), rng.choice(TOKENS), 1)
        elif k == 5:                                   # tweak a number
            lines[i] = re.sub(r"\d+(\.\d+)?", lambda m: rng.choice(["0", "1", "1.0", "0.001", "300", "2.5", "-3", m.group(0) + "0"]), lines[i], count=1)
        elif k == 6:                                   # rename a key
            m = re.match(r"^(\s*)([\w-]+)(:.*)$", lines[i])
            if m:
                lines[i] = m.group(1) + rng.choice(TOKENS + [m.group(2) + "x"]) + m.group(3)
        elif k == 7:                                   # toggle a bool
            lines[i] = lines[i].replace("true", "false") if "true" in lines[i] else lines[i].replace("false", "true")
        elif k == 8:                                   # quote the value (single / double)
            m = re.match(r"^(\s*[\w-]+:\s*)(\S.*)$", lines[i])
            if m and not m.group(2).startswith(("[", "{", "'", '"')):
                q = rng.choice("'\"")
                if q not in m.group(2) and "\\" not in m.group(2):
                    lines[i] = m.group(1) + q + m.group(2) + q
        elif k == 9:                                   # a YAML 1.2 core-schema special scalar as value
            m = re.match(r"^(\s*[\w-]+:\s*)(\S.*)$", lines[i])
            if m:
                lines[i] = m.group(1) + rng.choice(["~", "null", "Null", "yes", "no", "on", "off", "y", "n", "True", "FALSE", "0x10", "0o17", "010",
                                                    "1e3", "1E-2", ".5", "5.", "+3", "-0.0", ".inf", "-.INF", ".nan", "1_000", "0b11", "''", '""'])
        elif k == 10:                                  # flow list -> block list
            m = re.match(r"^(\s*)([\w-]+):\s*\[(.*)\]\s*$", lines[i])
            if m and "[" not in m.group(3):
                items = [x.strip() for x in m.group(3).split(",") if x.strip()]
                lines[i:i + 1] = [f"{m.group(1)}{m.group(2)}:"] + [f"{m.group(1)}  - {x}" for x in items]
        elif k == 11:                                  # trailing comment / blank line / comment line
            c = rng.randrange(3)
            if c == 0 and lines[i].strip():
                lines[i] = lines[i] + "   # note: x"
            elif c == 1:
                lines.insert(i, "")
            else:
                lines.insert(i, " " * rng.choice([0, 2, 4, 7]) + "# comment: [not, a, list]")
        elif k == 12:                                  # shift the indentation of one line
            lines[i] = (" " * rng.choice([1, 2])) + lines[i] if rng.random() < 0.5 else lines[i][min(2, len(lines[i]) - len(lines[i].lstrip())):]
        elif k == 13:                                  # YAML indicator characters inside a value
            m = re.match(r"^(\s*[\w-]+:\s*)(\S.*)$", lines[i])
            if m:
                v = m.group(2)
                cut = rng.randrange(len(v) + 1)
                lines[i] = m.group(1) + v[:cut] + rng.choice([" : ", ": ", " :", " ? 1 : 2", " #x", "#x", " - ", ", ", " [", "] ", " {", "} ", " & ", " * ", " ! ", " | ", " > ", "%", "@", "`"]) + v[cut:]
        else:                                          # shift a whole block (a line and everything deeper below it) by two columns
            ind = len(lines[i]) - len(lines[i].lstrip())
            j = i + 1
            while j < len(lines) and (not lines[j].strip() or len(lines[j]) - len(lines[j].lstrip()) > ind):
                j += 1
            for t in range(i + 1, j):
                if lines[t].strip():
                    lines[t] = "  " + lines[t]
    return "\n".join(lines)


def run(text):
    try:
        r = se.parse_string(text, compile=False)
        nat = ("ok", r.glsl_materials, r.glsl_rules)
    except se.SandEngineError as e:
        nat = (e.kind,)
    try:
        o = oracle_lang.parse_string(text)
        orc = ("ok", oracle_lang.emit_glsl_materials(o), oracle_lang.emit_glsl_rules(o))
    except oracle_lang.ParsingErr as e:
        orc = (e.kind,)
    except Exception as e:  # noqa: BLE001  (PyYAML refusing the document: the YAML level, class "Yaml")
        orc = ("Yaml",) if type(e).__module__.startswith("yaml") else ("PYTHON-EXCEPTION " + type(e).__name__ + ": " + str(e)[:100],)
    return nat, orc


if __name__ == "__main__":
    seed, n = int(sys.argv[1]), int(sys.argv[2])
    rng = random.Random(seed)
    stats, bad = {}, 0
    for it in range(n):
        text = mutate(rng.choice(SOURCES), rng)
        nat, orc = run(text)
        stats[nat[0]] = stats.get(nat[0], 0) + 1
        if nat != orc:
            bad += 1
            if bad <= 12:
                print(f"--- DISAGREE #{it}: native {nat[0]!r} vs oracle {orc[0]!r}")
                if nat[0] == orc[0] == "ok":
                    import difflib
                    for a, b in ((nat[1], orc[1]), (nat[2], orc[2])):
                        print("\n".join(list(difflib.unified_diff(a.splitlines(), b.splitlines(), lineterm="", n=0))[:8]))
                (REPO / "gpurun_out" / f"frontend_disagree_{seed}_{it}.yaml").write_text(text)
    print("mutants", n, "outcomes", stats, "disagreements", bad)
