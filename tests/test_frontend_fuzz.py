"""Mutation fuzzing of the native front end under AddressSanitizer + UBSan (CPU only, ~25 s)."""
import subprocess

import yaml_cases as Y
from conftest import DEFAULT_YAML, REPO


def test_front_end_survives_mutated_yaml(tmp_path):
    seeds = []
    for k, text in enumerate([Y.BASE_OK, Y.RICH_YAML, Y.EXPR_YAML, DEFAULT_YAML.read_text()]):
        p = tmp_path / f"seed{k}.yaml"
        p.write_text(text)
        seeds.append(str(p))
    exe = tmp_path / "fuzz_frontend"
    csrc = REPO / "sandengine_b200" / "csrc"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                           "-fno-omit-frame-pointer", "-I", str(csrc), str(REPO / "tests" / "fuzz" / "fuzz_frontend.cpp"),
                           str(csrc / "lang" / "yaml_lite.cpp"), str(csrc / "lang" / "lang.cpp"), str(csrc / "lang" / "codegen.cpp"),
                           "-o", str(exe)])
    r = subprocess.run([str(exe), "4000", *seeds], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    last = r.stdout.strip().splitlines()[-1]
    fields = dict(kv.split("=") for kv in last.split())
    assert int(fields["other"]) == 0 and int(fields["ok"]) > 0 and int(fields["parse_errors"]) > 0, last
