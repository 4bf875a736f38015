"""Ahead-of-time build of libgcanet_b200.so (sm_100a only, no import-time JIT).

    python -m gcanet_b200.build [--force] [--verbose]

Each csrc/*.cu is compiled with nvcc to an object file (in parallel) and linked into
gcanet_b200/lib/libgcanet_b200.so.  The library has no PyTorch dependency; it is
loaded with ctypes (gcanet_b200/_cabi.py).  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import argparse
import concurrent.futures
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "lib", "obj")
LIB = os.path.join(HERE, "lib", "libgcanet_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
    "-I", INCLUDE,
]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or "nvcc"
    for c in (cand, "/usr/local/cuda/bin/nvcc"):
        try:
            subprocess.run([c, "--version"], check=True, capture_output=True)
            return c
        except Exception:
            continue
    raise RuntimeError("nvcc not found; cannot build libgcanet_b200.so")


def _newer(src_list, target) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_list)


def build_library(force: bool = False, verbose: bool = False, extra_flags=()) -> str:
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    headers = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + sorted(glob.glob(os.path.join(INCLUDE, "*.h")))
    if not sources:
        raise RuntimeError(f"no CUDA sources under {CSRC}")
    os.makedirs(OBJ, exist_ok=True)
    # objects built with other flags (e.g. --aids) are stale whatever their age
    stamp = os.path.join(OBJ, ".flags")
    flags_now = " ".join([*NVCC_FLAGS, *[f for f in extra_flags if f not in ("-Xptxas", "-v")]])
    try:
        with open(stamp) as f:
            force = force or f.read() != flags_now
    except OSError:
        force = True
    with open(stamp, "w") as f:
        f.write(flags_now)
    nvcc = None
    jobs = []
    for src in sources:
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        if force or _newer([src] + headers, obj):
            nvcc = nvcc or _nvcc()
            jobs.append((src, obj, [nvcc, *NVCC_FLAGS, *extra_flags, "-c", src, "-o", obj]))

    def run(job):
        src, obj, cmd = job
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose and r.stderr.strip():
            print(r.stderr, flush=True)
        return obj

    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + ".o") for s in sources]
    if force or jobs or _newer(objs, LIB):
        nvcc = nvcc or _nvcc()
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--ptxas-v", action="store_true", help="print registers / spills / shared memory per kernel")
    ap.add_argument("--aids", action="store_true",
                    help="compile the measurement knobs in (GCANET_* environment variables, see DESIGN.md section 7)")
    a = ap.parse_args()
    extra = ["-Xptxas", "-v"] if a.ptxas_v else []
    if a.aids:
        extra.append("-DGCANET_MEASUREMENT_AIDS")
    print(build_library(force=a.force or a.ptxas_v or a.aids, verbose=a.verbose or a.ptxas_v, extra_flags=extra))
