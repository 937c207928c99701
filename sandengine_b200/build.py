"""Builds libsandengine_b200.so in-tree (nvcc, sm_100a) -- `python -m sandengine_b200.build`.

Steps
  1. embed kernels/sand_kernels.cuh as a C++ raw string (csrc/_gen/sand_kernels_embed.inc): it is the
     NVRTC translation unit that is compiled together with the generated rules header at rule-compile time.
  2. nvcc -gencode arch=compute_100a,code=sm_100a: api.cpp + lang/*.cpp + static_kernels.cu -> .so
     (static cudart; libnvrtc opened at run time by path; no libcuda link dependency).
  3. (inspect=True) run the front end on data/materials.yaml, dump the generated CUDA header and
     compile the rule kernels ahead of time with `-Xptxas -v` into build/ so registers/spills/SASS can be
     checked without a GPU (cuobjdump -sass build/sand_kernels_default.cubin).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
GEN = CSRC / "_gen"
REPO = PKG.parent
LIB = PKG / "libsandengine_b200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]

SOURCES = ["api.cpp", "lang/yaml_lite.cpp", "lang/lang.cpp", "lang/codegen.cpp", "static_kernels.cu"]
DEPS = SOURCES + ["lang/yaml_lite.h", "lang/lang.h", "lang/codegen.h", "static_kernels.h", "kernels/sand_kernels.cuh",
                  "../../include/sandengine_b200.h"]


def _embed_kernels() -> None:
    GEN.mkdir(parents=True, exist_ok=True)
    src = (CSRC / "kernels" / "sand_kernels.cuh").read_text()
    delim = "SEKERN"
    assert f"){delim}\"" not in src
    # split into chunks: some compilers cap a single string literal; adjacent literals concatenate
    chunks, step = [], 12000
    for i in range(0, len(src), step):
        chunks.append(f'R"{delim}({src[i:i + step]}){delim}"')
    out = "\n".join(chunks) + "\n"
    path = GEN / "sand_kernels_embed.inc"
    if not path.exists() or path.read_text() != out:
        path.write_text(out)


def _stamp() -> str:
    h = hashlib.sha256()
    for d in DEPS:
        h.update((CSRC / d).read_bytes())
    h.update(Path(__file__).read_bytes())
    return h.hexdigest()


def build(force: bool = False, inspect: bool = False, verbose: bool = True) -> Path:
    _embed_kernels()
    stamp_file = GEN / "build.stamp"
    stamp = _stamp()
    if force or not LIB.exists() or not stamp_file.exists() or stamp_file.read_text() != stamp:
        cmd = [NVCC, *ARCH, "-O3", "-std=c++17", "-lineinfo", "-shared", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
               "-cudart", "static", "-I", str(CSRC), "-I", str(REPO / "include")]
        cmd += [str(CSRC / s) for s in SOURCES]
        # NVRTC is opened at run time by path (csrc/api.cpp: struct Nvrtc): the toolkit's own, next to this nvcc
        nvrtc_dir = Path(NVCC).resolve().parent.parent / "lib64"
        cmd += [f'-DSE_NVRTC_DIR="{nvrtc_dir}"', "-o", str(LIB), "-ldl"]
        if verbose:
            print("[build]", " ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed building libsandengine_b200.so")
        if verbose and (r.stdout or r.stderr):
            print(r.stdout + r.stderr)
        stamp_file.write_text(stamp)
    if inspect:
        inspect_default()
    return LIB


def inspect_default() -> Path:
    """AOT-compile the NVRTC translation unit for data/materials.yaml with -Xptxas -v (no GPU needed)."""
    from . import lang

    out_dir = REPO / "build"
    out_dir.mkdir(exist_ok=True)
    rules = lang.parse_path(REPO / "data" / "materials.yaml")
    (out_dir / "rules_gen.cuh").write_text(rules.cuda_header)
    cu = out_dir / "sand_kernels_default.cu"
    cu.write_text((CSRC / "kernels" / "sand_kernels.cuh").read_text())
    cubin = out_dir / "sand_kernels_default.cubin"
    # same options as the NVRTC compile in api.cpp::compile_front
    cmd = [NVCC, *ARCH, "-O3", "-std=c++17", "-lineinfo", "-fmad=false", "-Xptxas", "-v", "-cubin",
           "-I", str(out_dir), str(cu), "-o", str(cubin)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    (out_dir / "ptxas_default.log").write_text(r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed on the generated rule kernels")
    # the cubin NVRTC produced for the same rule set (what actually runs)
    (out_dir / "sand_kernels_default.nvrtc.cubin").write_bytes(rules.cubin)
    return cubin


if __name__ == "__main__":
    build(force="--force" in sys.argv, inspect="--inspect" in sys.argv)
    print(LIB)
