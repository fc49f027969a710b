"""Build libelastic_b200.so (hand-written CUDA for sm_100a + C-ABI) in-tree with nvcc.

    python -m elasticreconstruction_b200.build [--force] [--verbose]

The shared object lands next to this file so that it travels to the GPU box with the repo snapshot.
nvcc cross-compiles without a GPU.  There is deliberately no fallback: if nvcc is missing or the build
fails this raises.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libelastic_b200.so")
BIN_DIR = os.path.join(HERE, "bin")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "--fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off,-fopenmp",
              "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found; libelastic_b200 cannot be built (there is no CPU fallback)")


def _sources():
    out = []
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cpp")) and not f.startswith("app_"):
            out.append(os.path.join(CSRC, f))
    return out


def _stale(target, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    srcs = _sources()
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    deps.append(os.path.join(ROOT, "include", "elastic_b200.h"))
    if force or _stale(LIB, deps):
        objs = []
        os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
        for s in srcs:
            o = os.path.join(HERE, "build", os.path.basename(s) + ".o")
            if force or _stale(o, deps):
                cmd = [_nvcc()] + ARCH + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
                if verbose:
                    print(" ".join(cmd), flush=True)
                subprocess.check_call(cmd)
            objs.append(o)
        cmd = [_nvcc()] + ARCH + ["-shared", "-Xcompiler", "-fopenmp", "-o", LIB] + objs + ["-lcudart", "-ldl"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB


def build_apps(force: bool = False, verbose: bool = False):
    """The drop-in executables (apps/Integrate, apps/BuildCorrespondence, apps/FragmentOptimizer): C++ hosts over the C ABI."""
    apps_dir = os.path.join(ROOT, "apps")
    outs = []
    if not os.path.isdir(apps_dir):
        return outs
    os.makedirs(BIN_DIR, exist_ok=True)
    hdrs = [os.path.join(apps_dir, f) for f in os.listdir(apps_dir) if f.endswith(".h")]
    for name in ("Integrate", "BuildCorrespondence", "FragmentOptimizer"):
        src = os.path.join(apps_dir, name + ".cpp")
        if not os.path.exists(src):
            continue
        out = os.path.join(BIN_DIR, name)
        if force or _stale(out, [src, LIB] + hdrs):
            cmd = ["g++", "-O2", "-std=c++17", "-fopenmp", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"),
                   "-I", apps_dir, src, "-o", out, "-L", HERE, "-lelastic_b200", "-Wl,-rpath,$ORIGIN/..",
                   "-Wl,-rpath-link," + HERE, "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd)
        outs.append(out)
    tool_src = os.path.join(apps_dir, "er_io_tool.cpp")
    if os.path.exists(tool_src):
        out = os.path.join(BIN_DIR, "er_io_tool")
        if force or _stale(out, [tool_src] + hdrs):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", apps_dir, tool_src, "-o", out])
        outs.append(out)
    return outs


if __name__ == "__main__":
    force = "--force" in sys.argv
    verbose = "--verbose" in sys.argv
    print(build_library(force, verbose))
    for a in build_apps(force, verbose):
        print(a)
