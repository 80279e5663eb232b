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


def rms_rel(a, b):
    """Per-element differences against the RMS magnitude of the reference tensor (a wrong class of small-magnitude
    entries shows up here even when the global-max norm hides it)."""
    a, b = a.double(), b.double()
    return float(torch.sqrt(((a - b) ** 2).mean()) / (torch.sqrt((b ** 2).mean()) + 1e-30))


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
    depth, xy, conic, rgb, lam = torch.zeros(P), torch.zeros(P, 2), torch.zeros(P, 3), torch.zeros(P, 3), torch.zeros(P)
    hc.hc_forward(P, deg, M, _p(gs["means3D"]), _p(gs["scales"]), _p(gs["rotations"]), ctypes.c_float(1.0), _p(gs["shs"]),
                  _p(view), _p(proj), _p(cam.campos), W, H, ctypes.c_float(cam.tanfovx), ctypes.c_float(cam.tanfovy),
                  _p(radii), _p(tiles), _p(rect), _p(depth), _p(xy), _p(conic), _p(rgb), _p(bits), _p(lam))
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
    dmean_pre = torch.zeros(P, 3)
    proj = cam.projmatrix.reshape(16).contiguous()
    hc.hc_cov_backward(P, _p(radii), _p(gs["means3D"]), _p(gs["scales"]), _p(gs["rotations"]), ctypes.c_float(1.0), _p(view),
                       _p(proj), W, H, ctypes.c_float(cam.tanfovx), ctypes.c_float(cam.tanfovy), _p(dconic), _p(dcov),
                       _p(dmean), _p(dscale), _p(drot), _p(dmean_pre))
    for got, key in ((dcov, "dL_dcov3D"), (dmean, "dL_dmeans3D"), (dmean_pre, "dL_dmeans3D"), (dscale, "dL_dscales"),
                     (drot, "dL_drotations")):
        assert rel(got, want[key]) < 1e-4, key
        assert rms_rel(got, want[key]) < 1e-4, key

    # screen-position path: dL/dmean2D -> dL/dmean
    dm2 = torch.randn(P, 2, generator=g)
    want = O.preprocess_backward(gs["means3D"], cam.viewmatrix, cam.projmatrix, cam.campos, W, H, cam.tanfovx, cam.tanfovy,
                                 pre, dict(dL_dconic=zero3, dL_dmean2D=dm2, dL_dcolors=zero3), **kw)
    dmean = torch.zeros(P, 3)
    hc.hc_ndc_backward(P, _p(radii), _p(gs["means3D"]), _p(proj), _p(dm2.contiguous()), _p(dmean))
    assert rel(dmean, want["dL_dmeans3D"]) < 1e-5 and rms_rel(dmean, want["dL_dmeans3D"]) < 1e-5

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


def _cull_check(hc, xy, conic, opac, lam, W, H):
    n = xy.shape[0]
    stats = (ctypes.c_longlong * 4)()
    hc.hc_cull_check.restype = ctypes.c_longlong
    bad = hc.hc_cull_check(n, _p(xy.contiguous()), _p(conic.contiguous()), _p(opac.contiguous()), _p(lam.contiguous()), W, H, stats)
    return int(bad), [int(v) for v in stats]


def test_culling_never_drops_a_contributing_pair(hc):
    """The sub-tile tests of the forward blend (csrc/gsr_cull.cuh) only skip work: for every splat of a projected scene
    and for adversarial synthetic splats (needle-thin, huge, nearly transparent, centred on sub-tile corners), a
    sub-tile containing ANY pixel that passes the reference's per-pixel rule is kept."""
    W, H = 160, 120
    gs, cam = _scene(6000, W, H, 0, 3, 1)
    pre = O.preprocess(gs["means3D"], gs["opacities"], cam.viewmatrix, cam.projmatrix, cam.campos, W, H, cam.tanfovx,
                       cam.tanfovy, gs["scales"], gs["rotations"], 1.0, None, gs["shs"], 0)
    vis = pre["radii"] > 0
    a, b, c = pre["conic_opacity"][vis, 0], pre["conic_opacity"][vis, 1], pre["conic_opacity"][vis, 2]
    # lam_max of the 2-D covariance = 1 / smaller eigenvalue of the conic
    mid = 0.5 * (a + c)
    lam_min_conic = mid - torch.sqrt(torch.clamp(mid * mid - (a * c - b * b), min=0))
    lam = (1.0 / lam_min_conic).float()
    bad, st = _cull_check(hc, pre["xy"][vis].float(), pre["conic_opacity"][vis, :3].float(),
                          pre["conic_opacity"][vis, 3].float(), lam, W, H)
    assert bad == 0, (bad, st)
    assert st[0] > 0 and st[2] < 0.25 * st[3]            # and it does skip: fewer than a quarter of all pairs are kept

    g = torch.Generator().manual_seed(11)
    n = 4000
    # covariances with extreme anisotropy / size, random orientation; centres snapped near sub-tile corners half the time
    l1 = torch.exp(torch.rand(n, generator=g) * 9.0 - 1.2)          # 0.3 .. 2400 px^2
    l2 = l1 * torch.exp(-torch.rand(n, generator=g) * 7.0)          # ratio down to 1e-3
    l2 = torch.clamp(l2, min=0.3)
    th = torch.rand(n, generator=g) * 3.14159265
    cs, sn = torch.cos(th), torch.sin(th)
    sxx, syy, sxy = l1 * cs * cs + l2 * sn * sn, l1 * sn * sn + l2 * cs * cs, (l1 - l2) * cs * sn
    det = sxx * syy - sxy * sxy
    conic = torch.stack([syy / det, -sxy / det, sxx / det], -1).float()
    lam = torch.maximum(l1, l2).float()
    xy = torch.stack([torch.rand(n, generator=g) * (W + 40) - 20, torch.rand(n, generator=g) * (H + 40) - 20], -1)
    snap = torch.rand(n, generator=g) < 0.5
    corner = torch.stack([torch.round(xy[:, 0] / 8) * 8, torch.round(xy[:, 1] / 4) * 4], -1) + (torch.rand(n, 2, generator=g) - 0.5) * 1e-3
    xy = torch.where(snap[:, None], corner, xy).float()
    opac = torch.cat([torch.rand(n // 2, generator=g), 1.0 / 255.0 + torch.rand(n - n // 2, generator=g) * 0.02]).float()
    bad, st = _cull_check(hc, xy, conic, opac, lam, W, H)
    assert bad == 0, (bad, st)


def test_tile_partition_division_is_exact(hc):
    """k / w by multiplication (binning.cu's instance -> (row, column) unflattening) for every width and instance index
    the API admits: tile grids are at most 1023 x 1023 (gsr_forward_preprocess rejects larger images)."""
    hc.hc_magic_div_check.restype = ctypes.c_longlong
    assert hc.hc_magic_div_check(1023, 1023) == 0
