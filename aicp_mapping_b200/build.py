"""Builds libaicp_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m aicp_mapping_b200.build [--force] [--verbose]

-fmad=false is part of the numerical contract (DESIGN.md "Arithmetic"): the float geometry and the float64 solve must not
be contracted into FMAs.  -lineinfo keeps ncu's source page usable.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libaicp_b200.so")
OBJ_DIR = os.path.join(HERE, "build")
SOURCES = ["api.cu", "index.cu", "sort.cu", "normals.cu", "icp.cu", "append.cu", "overlap.cu", "crop.cu", "prefilter.cu", "alignability.cu", "ingest.cu", "svm.cu", "comm.cu", "config_yaml.cpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
         "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall,-Wno-unused-function", "-ccbin", "/usr/bin/g++"]


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".hpp"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "aicp_b200.h"))
    return hdrs


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, variant=None, defines=()):
    """variant / defines: build an experiment copy (lib/libaicp_b200_<variant>.so with -D<define>...) next to the product
    library; used by tools/ab_bench.py only."""
    global LIB
    lib_path = LIB if not variant else os.path.join(LIB_DIR, "libaicp_b200_%s.so" % variant)
    obj_dir = OBJ_DIR if not variant else os.path.join(OBJ_DIR, variant)
    return _build(force or bool(variant), verbose, lib_path, obj_dir, ["-D" + d for d in defines])


def _build(force, verbose, LIB, OBJ_DIR, extra):
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdrs = _deps()
    jobs = []
    objs = []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        op = os.path.join(OBJ_DIR, src.rsplit(".", 1)[0] + ".o")
        objs.append(op)
        if force or _stale(op, [sp] + hdrs + [os.path.abspath(__file__)]):
            cmd = [NVCC] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu", "-c", sp, "-o", op]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for cmd, r in ex.map(run, jobs):
                if verbose or r.returncode != 0:
                    sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr + "\n")
                if r.returncode != 0:
                    raise RuntimeError("nvcc failed for " + cmd[-3])
    if jobs or force or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
                                                     "-Xlinker", "--no-undefined", "-ldl", "-lpthread", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr + "\n")
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
