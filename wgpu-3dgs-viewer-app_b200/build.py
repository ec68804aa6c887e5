"""Builds libb200gs.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python wgpu-3dgs-viewer-app_b200/build.py [--force] [--verbose] [--synccheck]

--synccheck builds the variant compute-sanitizer's synccheck can follow (the compositor's two barrier helpers kept out
of line, csrc/composite.cu); it forces a rebuild of composite.cu — run plain `build.py --force` afterwards to go back.

Cross-compiles without a GPU.  preprocess.cu and aux.cu are built with -fmad=false (one
rounding per float operation: bit parity with the CPU oracle); host code with
-ffp-contract=off for the same reason.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, "libb200gs.so")
OBJ = os.path.join(HERE, "_build")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off,-Wall",
               "-I", os.path.join(ROOT, "include")]
CU = {
    "preprocess.cu": ["-fmad=false"],
    "aux.cu": ["-fmad=false"],
    "sort.cu": [],
    "sort_wide.cu": [],
    "bin.cu": [],
    "composite.cu": [],
    "api.cu": [],
}
CPP = ["host/host.cpp"]
HEADERS = ["csrc/common.cuh", "csrc/host_api.h", "../include/b200gs.h"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, synccheck=False):
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(HERE, h) for h in HEADERS] + [os.path.abspath(__file__)]
    jobs, objs = [], []
    for src, extra in CU.items():
        s = os.path.join(HERE, "csrc", src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        sc = synccheck and src == "composite.cu"
        if force or sc or _stale(o, [s] + hdrs):
            jobs.append([_nvcc()] + ARCH + NVCC_COMMON + extra + (["-DB200GS_SYNCCHECK_BUILD"] if sc else []) +
                        (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o])
    for src in CPP:
        s = os.path.join(HERE, src)
        o = os.path.join(OBJ, os.path.basename(src).replace(".cpp", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            jobs.append(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-ffp-contract=off", "-Wall",
                         "-pthread", "-I", os.path.join(ROOT, "include"), "-c", s, "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or verbose:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("build failed: " + " ".join(cmd))

    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(run, jobs))
    if jobs or force or _stale(OUT, objs):
        run([_nvcc()] + ARCH + ["-shared", "-o", OUT] + objs + ["-lpthread"])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, synccheck="--synccheck" in sys.argv))
