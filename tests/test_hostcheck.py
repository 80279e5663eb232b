"""The product's per-Gaussian math on the CPU: tests/hostcheck/hostcheck.cpp compiles the GSR_HD functions of
mm3dgs-slam_b200/csrc/gsr_math.cuh (the very functions k_preprocess_fwd / k_preprocess_bwd call) with g++ and this test
compares them with the oracle (itself pinned to the reference by tests/golden) — radii, tile rectangles and tile counts
exactly, floating-point results within 1e-4.  No GPU involved; the kernels' orchestration around these functions is
covered by the -m gpu parity tests."""
import ctypes
import os
import shutil
import subprocess

import pytest
import torch

import gsr_synth as S
from oracle import gs_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hc(tmp_path_factory):
    if not shutil.which("g++"):
        pytest.skip("no C++ compiler")
    so = str(tmp_path_factory.mktemp("hostcheck") / "libhostcheck.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "mm3dgs-slam_b200", "csrc"),
                    "-x", "c++", os.path.join(ROOT, "tests", "hostcheck", "hostcheck.cpp"), "-o", so], check=True)
    return ctypes.CDLL(so)


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))


def _scene(P, W, H, deg, seed, cam_index):
    gs, cam, _, _ = S.make_scene(P, W, H, seed, deg)
    if cam_index is not None:
        cam = S.orbit_cameras(W, H, 3, (0.0, 0.0, 4.0), 0.5)[cam_index]
    return gs, cam


CASES = [(4000, 160, 120, 0, 1, None), (3000, 128, 96, 3, 2, 1), (2500, 200, 88, 2, 7, 2)]


@pytest.mark.parametrize("P,W,H,deg,seed,cam_index", CASES)
def test_forward_math_matches_oracle(hc, P, W, H, deg, seed, cam_index):
    gs, cam = _scene(P, W, H, deg, seed, cam_index)
    M = gs["shs"].shape[1]
    view, proj = cam.viewmatrix.reshape(16).contiguous(), cam.projmatrix.reshape(16).contiguous()
    radii, tiles, bits = (torch.zeros(P, dtype=torch.int32) for _ in range(3))
    rect = torch.zeros(P, 4, dtype=torch.int32)
    depth, xy, conic, rgb = torch.zeros(P), torch.zeros(P, 2), torch.zeros(P, 3), torch.zeros(P, 3)
    hc.hc_forward(P, deg, M, _p(gs["means3D"]), _p(gs["scales"]), _p(gs["rotations"]), ctypes.c_float(1.0), _p(gs["shs"]),
                  _p(view), _p(proj), _p(cam.campos), W, H, ctypes.c_float(cam.tanfovx), ctypes.c_float(cam.tanfovy),
                  _p(radii), _p(tiles), _p(rect), _p(depth), _p(xy), _p(conic), _p(rgb), _p(bits))
    pre = O.preprocess(gs["means3D"], gs["opacities"], cam.viewmatrix, cam.projmatrix, cam.campos, W, H, cam.tanfovx,
                       cam.tanfovy, gs["scales"], gs["rotations"], 1.0, None, gs["shs"], deg)
    vis = pre["radii"] > 0
    assert int(vis.sum()) > P // 4
    # integer outputs: exact (host FMA contraction may differ from nvcc's on a knife-edge value; allow 1 in 1000)
    bad = (radii != pre["radii"]) | (tiles.long() != pre["tiles_touched"]) | ((rect.long() != pre["rect"]).any(-1) & vis)
    assert int(bad.sum()) <= P // 1000, int(bad.sum())
    ok = vis & ~bad
    assert rel(depth[ok], pre["depth"][ok]) < 1e-6 and rel(xy[ok], pre["xy"][ok]) < 1e-5
    assert rel(conic[ok], pre["conic_opacity"][ok, :3]) < 1e-4
    assert rel(rgb[ok], pre["rgb"][ok]) < 1e-5
    want_bits = (pre["clamped"].long() * torch.tensor([1, 2, 4])).sum(-1)
    assert int((bits.long()[ok] != want_bits[ok]).sum()) <= 2


@pytest.mark.parametrize("P,W,H,deg,seed,cam_index", CASES)
def test_backward_math_matches_oracle(hc, P, W, H, deg, seed, cam_index):
    gs, cam = _scene(P, W, H, deg, seed, cam_index)
    M = gs["shs"].shape[1]
    view = cam.viewmatrix.reshape(16).contiguous()
    pre = O.preprocess(gs["means3D"], gs["opacities"], cam.viewmatrix, cam.projmatrix, cam.campos, W, H, cam.tanfovx,
                       cam.tanfovy, gs["scales"], gs["rotations"], 1.0, None, gs["shs"], deg)
    radii = pre["radii"].contiguous()
    g = torch.Generator().manual_seed(seed)
    zero2, zero3 = torch.zeros(P, 2), torch.zeros(P, 3)
    kw = dict(scales=gs["scales"], rotations=gs["rotations"], scale_modifier=1.0, shs=gs["shs"], sh_degree=deg)

    # covariance path: dL/dconic -> dL/dcov3D, dL/dmean, dL/dscale, dL/drot
    dconic = torch.randn(P, 3, generator=g)
    want = O.preprocess_backward(gs["means3D"], cam.viewmatrix, cam.projmatrix, cam.campos, W, H, cam.tanfovx, cam.tanfovy,
                                 pre, dict(dL_dconic=dconic, dL_dmean2D=zero2, dL_dcolors=zero3), **kw)
    dcov, dmean, dscale, drot = torch.zeros(P, 6), torch.zeros(P, 3), torch.zeros(P, 3), torch.zeros(P, 4)
    hc.hc_cov_backward(P, _p(radii), _p(gs["means3D"]), _p(gs["scales"]), _p(gs["rotations"]), ctypes.c_float(1.0), _p(view),
                       W, H, ctypes.c_float(cam.tanfovx), ctypes.c_float(cam.tanfovy), _p(dconic), _p(dcov), _p(dmean),
                       _p(dscale), _p(drot))
    assert rel(dcov, want["dL_dcov3D"]) < 1e-4
    assert rel(dmean, want["dL_dmeans3D"]) < 1e-4
    assert rel(dscale, want["dL_dscales"]) < 1e-4
    assert rel(drot, want["dL_drotations"]) < 1e-4

    # SH path: clamp-masked dL/dRGB -> dL/dsh and the view-direction part of dL/dmean
    dcol = torch.randn(P, 3, generator=g)
    want = O.preprocess_backward(gs["means3D"], cam.viewmatrix, cam.projmatrix, cam.campos, W, H, cam.tanfovx, cam.tanfovy,
                                 pre, dict(dL_dconic=zero3, dL_dmean2D=zero2, dL_dcolors=dcol), **kw)
    bits = (pre["clamped"].long() * torch.tensor([1, 2, 4])).sum(-1).to(torch.int32).contiguous()
    dsh, dmean = torch.zeros(P, M, 3), torch.zeros(P, 3)
    hc.hc_sh_backward(P, deg, M, _p(radii), _p(bits), _p(gs["means3D"]), _p(gs["shs"]), _p(cam.campos), _p(dcol), _p(dsh),
                      _p(dmean))
    assert rel(dsh, want["dL_dsh"]) < 1e-5
    if deg > 0:
        assert rel(dmean, want["dL_dmeans3D"]) < 1e-4
    else:
        assert float(dmean.abs().max()) == 0.0 and float(want["dL_dmeans3D"].abs().max()) < 1e-12
