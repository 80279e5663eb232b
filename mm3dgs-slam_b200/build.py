#!/usr/bin/env python
"""Build libgsrast_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python mm3dgs-slam_b200/build.py [--force] [--verbose]

Output: mm3dgs-slam_b200/lib/libgsrast_b200.so (git-ignored, travels to the GPU box with gpurun).
No torch involved: the library is plain CUDA C++ behind the extern "C" API of include/gsrast_b200.h.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
SO = os.path.join(LIBDIR, "libgsrast_b200.so")
SOURCES = ["preprocess.cu", "preprocess_bwd.cu", "binning.cu", "render.cu", "slam_ops.cu", "map_ops.cu", "comm.cu", "c_api.cu"]
# per-file extra flags.  (--use_fast_math on preprocess_bwd.cu — legal there, nothing in the backward feeds an
# integer output — was measured SLOWER on B200: 74 vs 64 registers, 106 vs 95 us; so everything uses nvcc defaults.)
EXTRA_FLAGS = {"preprocess_bwd.cu": ["--use_fast_math"] if os.environ.get("GSR_FAST_BWD") else []}
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# nvcc floating-point defaults on purpose (-fmad=true, IEEE div/sqrt, no fast-math): the integer
# outputs (radii, tile rectangles, sort keys) must match the reference build bit for bit.
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
          "--expt-relaxed-constexpr", "-DGSR_BUILD"]


def _deps():
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    d += [os.path.join(HERE, "..", "include", h) for h in ("gsrast_b200.h", "gsloss_b200.h", "gscomm_b200.h")]
    return d


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(p) > t for p in _deps())


def _run(cmd, verbose):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 or verbose:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed")


def build(force=False, verbose=False, ptxas_info=False, defines=(), out=None):
    """defines/out: build an experimental variant (e.g. defines=["MY_SWITCH=1"], out=".../libvariant.so") for A/B runs."""
    global SO
    if out is None and not force and not needs_build():
        return SO
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj" if out is None else "obj_" + os.path.basename(out))
    so = SO if out is None else out
    os.makedirs(objdir, exist_ok=True)
    extra = ["-Xptxas", "-v"] if ptxas_info else []
    jobs, objs = [], []
    for s in SOURCES:
        o = os.path.join(objdir, s + ".o")
        objs.append(o)
        jobs.append([NVCC, "-c", os.path.join(CSRC, s), "-o", o] + ARCH + CFLAGS + extra + EXTRA_FLAGS.get(s, [])
                    + ["-D" + d for d in defines])
    with ThreadPoolExecutor(len(jobs)) as ex:
        list(ex.map(lambda c: _run(c, verbose or ptxas_info), jobs))
    _run([NVCC, "-shared", "-o", so] + objs + ARCH + ["-cudart", "shared", "-Xcompiler", "-fPIC"], verbose)
    return so


if __name__ == "__main__":
    defs = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--define=")]
    outs = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, ptxas_info="--ptxas" in sys.argv,
                defines=defs, out=outs[0] if outs else None))
