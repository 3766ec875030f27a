"""In-tree build of libyolo2_b200.so (CUDA kernels + C host runtime + C++ Detector).

nvcc cross-compiles for sm_100a without a GPU.  The library is written next to this
file so it travels with the repo snapshot to the GPU box; nothing is JIT-cached.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "libyolo2_b200.so"
OBJ = PKG / "build"

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
GENCODE = ["-gencode", "arch=compute_100a,code=sm_100a"]
CUDA_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
              "-I", str(ROOT / "include"), "-I", str(CSRC / "cuda")]
# decode / NMS must reproduce the reference's IEEE float expressions bit for bit
EXACT_FLAGS = ["-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false"]
EXACT_FILES = {"region.cu", "nms.cu"}
C_FLAGS = ["-O2", "-std=gnu11", "-fPIC", "-Wall", "-Wno-unused-result", "-ffp-contract=off",
           "-I", str(ROOT / "include"), "-I", str(CSRC / "host")]
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-Wall", "-I", str(ROOT / "include"), "-I", str(CSRC / "host")]


def _sources():
    cu = sorted((CSRC / "cuda").glob("*.cu"))
    c = sorted((CSRC / "host").glob("*.c"))
    cpp = sorted((CSRC / "host").glob("*.cpp"))
    return cu, c, cpp


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def _run(cmd, verbose):
    if verbose:
        print(" ".join(str(c) for c in cmd), flush=True)
    r = subprocess.run([str(c) for c in cmd], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"build step failed: {' '.join(str(c) for c in cmd[:4])} ...")
    if verbose and (r.stdout or r.stderr):
        sys.stderr.write(r.stdout + r.stderr)


def build(force: bool = False, verbose: bool = False, ptxas_info: bool = False) -> Path:
    cu, c, cpp = _sources()
    headers = list((ROOT / "include").glob("*.h")) + list((ROOT / "include").glob("*.hpp")) + \
        list((CSRC / "cuda").glob("*.cuh")) + list((CSRC / "host").glob("*.h"))
    stamp = OBJ / "stamp.txt"
    digest = _digest(cu + c + cpp + headers + [Path(__file__)])
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB
    OBJ.mkdir(exist_ok=True)
    objs = []
    for src in cu:
        o = OBJ / (src.stem + ".cu.o")
        flags = list(CUDA_FLAGS)
        if src.name in EXACT_FILES:
            flags += EXACT_FLAGS
        if ptxas_info:
            flags += ["-Xptxas", "-v"]
        _run([NVCC, *GENCODE, *flags, "-c", src, "-o", o], verbose or ptxas_info)
        objs.append(o)
    for src in c:
        o = OBJ / (src.stem + ".c.o")
        _run(["gcc", *C_FLAGS, "-c", src, "-o", o], verbose)
        objs.append(o)
    for src in cpp:
        o = OBJ / (src.stem + ".cpp.o")
        _run(["g++", *CXX_FLAGS, "-c", src, "-o", o], verbose)
        objs.append(o)
    # -Bsymbolic-functions: the drop-in API keeps the reference's unprefixed names (error, strip,
    # free_list ...); calls between our own functions must not be interposed by a same-named
    # symbol of the host process (glibc exports error(3)).
    _run([NVCC, *GENCODE, "-shared", "-Xlinker", "-Bsymbolic-functions", "-o", LIB, *objs, "-lcudart", "-lm",
          "-lpthread", "-lstdc++"], verbose)
    stamp.write_text(digest)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True, ptxas_info="--ptxas" in sys.argv)
    print(p)
