"""Ahead-of-time build of libgcanet_b200.so (sm_100a only, no import-time JIT).

    python -m gcanet_b200.build [--force] [--verbose]

Each csrc/*.cu is compiled with nvcc to an object file (in parallel) and linked into
gcanet_b200/lib/libgcanet_b200.so.  The library has no PyTorch dependency; it is
loaded with ctypes (gcanet_b200/_cabi.py).  nvcc cross-compiles without a GPU.

csrc_ext/torch_ext.cpp -- the PyTorch C++ extension over that C-ABI, the counterpart of the reference's pybind modules --
is compiled with g++ against the installed torch's headers into gcanet_b200/lib/gcanet_b200_ext.so (build_extension,
loaded by gcanet_b200/native_ext.py).
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


EXT_SRC = os.path.join(HERE, "csrc_ext", "torch_ext.cpp")
EXT = os.path.join(HERE, "lib", "gcanet_b200_ext.so")
EXT_NAME = "gcanet_b200_ext"


def build_extension(force: bool = False, verbose: bool = False) -> str:
    """g++ build of the PyTorch C++ extension (pybind module + TORCH_LIBRARY registration) that wraps the C-ABI.  It links
    against libgcanet_b200.so (found next to it through $ORIGIN) and the torch libraries of the running interpreter,
    which are already loaded whenever the module is imported (native_ext.py imports torch first)."""
    import sysconfig

    import torch
    from torch.utils import cpp_extension as ce

    if not os.path.exists(LIB):
        raise RuntimeError("build libgcanet_b200.so first (build_library)")
    stamp = EXT + ".torch"                              # rebuilt when the torch it was compiled against changes
    same_torch = os.path.exists(stamp) and open(stamp).read() == torch.__version__
    if not force and same_torch and not _newer([EXT_SRC, os.path.join(INCLUDE, "gcanet_b200.h"), LIB], EXT):
        return EXT
    cuda_home = os.environ.get("CUDA_HOME") or "/usr/local/cuda"
    args = ["-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden",
            f"-DTORCH_EXTENSION_NAME={EXT_NAME}", "-DTORCH_API_INCLUDE_EXTENSION_H",
            f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    for inc in [*ce.include_paths(), os.path.join(cuda_home, "include"), sysconfig.get_paths()["include"], INCLUDE]:
        args += ["-I", inc]
    args += [EXT_SRC, "-o", EXT]
    for lp in ce.library_paths():
        args += ["-L", lp]
    args += ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python",
             "-L", os.path.dirname(LIB), "-lgcanet_b200", "-Wl,-rpath,$ORIGIN"]
    # The module must share the process's libstdc++.so.6 with torch: a toolchain whose libstdc++.so is missing links
    # libstdc++.a into the module instead, and that private copy (its own, never initialised locale / iostream state)
    # crashes the first time an error message formats a number.  So: the system compiler first, and the result is checked.
    problems = []
    for cxx in dict.fromkeys(["/usr/bin/g++", "g++", os.environ.get("CXX") or "g++"]):
        cmd = [cxx, *args]
        if verbose:
            print(" ".join(cmd), flush=True)
        try:
            r = subprocess.run(cmd, capture_output=True, text=True)
        except OSError as exc:
            problems.append(f"{cxx}: {exc}")
            continue
        if r.returncode != 0:
            problems.append(f"{cxx} failed on {EXT_SRC}:\n{r.stdout}\n{r.stderr}")
            continue
        with open(EXT, "rb") as f:
            if b"libstdc++.so.6" in f.read():             # DT_NEEDED entry present
                break
        problems.append(f"{cxx} linked libstdc++ statically into {EXT}")
        os.remove(EXT)
    else:
        raise RuntimeError("could not build the PyTorch extension:\n" + "\n".join(problems))
    with open(stamp, "w") as f:
        f.write(torch.__version__)
    return EXT


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
    print(build_extension(force=a.force, verbose=a.verbose))
