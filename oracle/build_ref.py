#!/usr/bin/env python
"""Recipe: compile the UNMODIFIED reference rasterizer into oracle/_ref/ (test infrastructure).

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported by the product package
(`mm3dgs-slam_b200/`); only tests/, __graft_entry__.smoke() and bench.py's baseline legs use it.

What it does
------------
Compiles the reference's own five translation units *where they lie* under
/root/reference/submodules/diff-gaussian-rasterization (DGR/):

    DGR/ext.cpp, DGR/rasterize_points.cu,
    DGR/cuda_rasterizer/{forward,backward,rasterizer_impl}.cu

for sm_100a into `oracle/_ref/ref_dgr_C.so`, a torch extension whose pybind surface is the
reference's `_C` module (DGR/ext.cpp:15-19: rasterize_gaussians,
rasterize_gaussians_backward, mark_visible).  No reference source is copied into this repo;
only the built .so lands in oracle/_ref/ (git-ignored, NOT gpurun-ignored, so it travels to
the GPU box where /root/reference does not exist).

The only deviation from the reference's setup.py (DGR/setup.py:21-29) is two command-line
flags: `-include cstdint` (DGR/cuda_rasterizer/rasterizer_impl.h:24,40-41 use uintptr_t /
uint32_t / uint64_t without including <cstdint>, which gcc 13 rejects) and an explicit
`-gencode arch=compute_100a,code=sm_100a` (the reference carries no arch list).  Default
nvcc floating-point flags are kept (-fmad=true, no --use_fast_math), as in the reference.

The reference's Python wrapper (DGR/diff_gaussian_rasterization/__init__.py) does
`from . import _C`; oracle/ref_api.py re-states that thin wrapper around `ref_dgr_C` so the
compiled reference can be imported side by side with the B200 drop-in.
"""
import os
import shutil
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
DGR = "/root/reference/submodules/diff-gaussian-rasterization"
MOD = "ref_dgr_C"


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("reference build step failed")
    return r


def so_path():
    return os.path.join(OUT, MOD + ".so")


def build(force=False, verbose=False):
    """Build oracle/_ref/ref_dgr_C.so if the reference sources are present. Returns path or None."""
    if os.path.exists(so_path()) and not force:
        return so_path()
    if not os.path.isdir(DGR):
        return None  # GPU box: only the prebuilt .so is used
    import torch
    from torch.utils import cpp_extension as ce

    os.makedirs(OUT, exist_ok=True)
    obj = os.path.join(OUT, "obj")
    os.makedirs(obj, exist_ok=True)
    inc = []
    for p in ce.include_paths():
        inc += ["-isystem", p]
    inc += ["-isystem", sysconfig.get_paths()["include"]]
    inc += ["-I", os.path.join(DGR, "third_party/glm"), "-I", DGR]
    defs = ["-DTORCH_EXTENSION_NAME=" + MOD, "-DTORCH_API_INCLUDE_EXTENSION_H",
            "-D_GLIBCXX_USE_CXX11_ABI=" + str(int(torch.compiled_with_cxx11_abi()))]
    nvcc = os.path.join(ce.CUDA_HOME or "/usr/local/cuda", "bin", "nvcc")
    cu = ["rasterize_points.cu", "cuda_rasterizer/forward.cu", "cuda_rasterizer/backward.cu",
          "cuda_rasterizer/rasterizer_impl.cu"]
    jobs = []
    objs = []
    for s in cu:
        o = os.path.join(obj, os.path.basename(s) + ".o")
        objs.append(o)
        jobs.append([nvcc, "-c", os.path.join(DGR, s), "-o", o, "-std=c++17", "-O3",
                     "-gencode", "arch=compute_100a,code=sm_100a", "-include", "cstdint",
                     "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-w"] + defs + inc)
    o = os.path.join(obj, "ext.cpp.o")
    objs.append(o)
    jobs.append(["g++", "-c", os.path.join(DGR, "ext.cpp"), "-o", o, "-std=c++17", "-O2", "-fPIC",
                 "-include", "cstdint", "-w", "-isystem",
                 os.path.join(ce.CUDA_HOME or "/usr/local/cuda", "include")] + defs + inc)
    with ThreadPoolExecutor(5) as ex:
        list(ex.map(_run, jobs))
    libdirs = ce.library_paths(device_type="cuda")
    link = ["g++", "-shared", "-o", so_path()] + objs
    for d in libdirs:
        link += ["-L", d, "-Wl,-rpath," + d]
    link += ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart"]
    _run(link)
    shutil.rmtree(obj, ignore_errors=True)
    if verbose:
        print("built", so_path())
    return so_path()


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print(p or "reference sources not present and no prebuilt .so")
