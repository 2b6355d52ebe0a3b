"""Build recipe of libdxmcb200.so (CUDA kernels + C ABI + drop-in C++ host classes) for sm_100a.

Everything is compiled in-tree with explicit nvcc / g++ commands (no JIT cache), so the built
library travels with the repo snapshot to the GPU box. nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, "libdxmcb200.so")
OBJ = os.path.join(HERE, "build")

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
CXX = os.environ.get("CXX", "g++")

# host code: -ffp-contract=off so table builders round like the reference oracle build; device code: FMA contraction on,
# the bit-exact parts (geometry, LUT index and exponent) use explicit round-to-nearest intrinsics
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
CXX_FLAGS = ["-O2", "-std=c++20", "-fPIC", "-ffp-contract=off", "-pthread", "-Wno-narrowing", "-fno-gnu-unique"]
INCLUDES = [f"-I{ROOT}/include", f"-I{HERE}/include", f"-I{HERE}/host", f"-I{HERE}/csrc"]

CUDA_SOURCES = ["csrc/transport.cu"]
HOST_SOURCES = ["host/xrl_lite.cpp", "host/matdb.cpp", "host/scene_capi.cpp"]


def _xraylib():
    """Compile and link flags of a real xraylib (the reference's element data library, CMakeLists.txt:47-55), when there is
    one: DXMCB200_XRAYLIB=<prefix> (with include/xraylib/xraylib.h and lib/libxrl.so), or `pkg-config libxrl`. DXMCB200_XRAYLIB=0
    forces the in-repo approximate xrl_lite. Returns (cflags, ldflags) or None."""
    prefix = os.environ.get("DXMCB200_XRAYLIB", "")
    if prefix == "0":
        return None
    if prefix:
        for inc in (os.path.join(prefix, "include", "xraylib"), os.path.join(prefix, "include")):
            if os.path.exists(os.path.join(inc, "xraylib.h")):
                lib = os.path.join(prefix, "lib")
                return [f"-I{inc}"], [f"-L{lib}", "-lxrl", "-Xlinker", f"-rpath={lib}"]
        raise RuntimeError(f"DXMCB200_XRAYLIB={prefix}: no xraylib.h under it")
    pc = shutil.which("pkg-config")
    if pc and subprocess.run([pc, "--exists", "libxrl"]).returncode == 0:
        cflags = subprocess.check_output([pc, "--cflags", "libxrl"], text=True).split()
        libs = subprocess.check_output([pc, "--libs", "libxrl"], text=True).split()
        return cflags, libs
    return None


def _newer(src_files, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(f) > t for f in src_files)


def _deps():
    deps = []
    for base in (os.path.join(HERE, "include"), os.path.join(HERE, "host"), os.path.join(HERE, "csrc"), os.path.join(ROOT, "include")):
        for d, _, files in os.walk(base):
            deps += [os.path.join(d, f) for f in files if f.endswith((".h", ".hpp", ".cuh", ".cu", ".cpp"))]
    return deps


def _run(cmd, log=None):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log is not None:
        log.append(p.stdout)
    if p.returncode != 0:
        sys.stderr.write(p.stdout)
        raise RuntimeError("build step failed: " + " ".join(cmd))
    return p.stdout


def build(force: bool = False, verbose: bool = False) -> str:
    deps = _deps()
    if not force and not _newer(deps, OUT):
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    objs = []
    log: list[str] = []
    for s in CUDA_SOURCES:
        o = os.path.join(OBJ, os.path.basename(s) + ".o")
        _run([NVCC, *NVCC_FLAGS, *INCLUDES, "-c", os.path.join(HERE, s), "-o", o], log)
        objs.append(o)
    procs = []
    xrl = _xraylib()
    xrl_cflags = ["-DDXMCB200_USE_XRAYLIB", *xrl[0]] if xrl else []
    xrl_ldflags = xrl[1] if xrl else []
    for s in HOST_SOURCES:
        o = os.path.join(OBJ, os.path.basename(s) + ".o")
        procs.append((s, subprocess.Popen([CXX, *CXX_FLAGS, *xrl_cflags, *INCLUDES, "-c", os.path.join(HERE, s), "-o", o], stdout=subprocess.PIPE,
                                          stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f"compiling {s} failed")
    _run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT, *objs, "-ldl", "-lpthread", "-Xlinker", "-Bsymbolic", *xrl_ldflags], log)
    with open(os.path.join(OBJ, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
