"""CPU tests of the oracle itself: its analytic backward (the reference's hand-written derivative,
restated) against fp64 autograd through a differentiable restatement of the forward, and internal
consistency of binning.  (The oracle-vs-reference pin lives in test_oracle_golden.py.)"""
import torch

import gsr_synth as S
from oracle import gs_oracle as O


def _scene(P=300, W=64, H=48, deg=2, seed=3):
    gs, _, dL, _ = S.make_scene(P, W, H, seed=seed, sh_degree=deg)
    gs["opacities"] = gs["opacities"] * 0.6   # stay below the alpha clamp: the reference ignores it in backward
    cam = S.orbit_cameras(W, H, 3, (0.0, 0.0, 4.0), 0.5)[1]
    return gs, cam, dL, torch.tensor([0.2, 0.5, 0.7])


def test_analytic_backward_matches_autograd_fp64():
    W, H, deg = 64, 48, 2
    gs, cam, dL, bg = _scene(300, W, H, deg)
    dt = torch.float64
    leaf = {k: v.to(dt).clone().requires_grad_(True) for k, v in gs.items()}
    img = O.differentiable_render(leaf["means3D"], leaf["opacities"], leaf["scales"], leaf["rotations"], leaf["shs"],
                                  deg, cam.viewmatrix, cam.projmatrix, cam.campos, bg, W, H, cam.tanfovx, cam.tanfovy)
    (img * dL.to(dt)).sum().backward()
    pre, binning, fwd = O.rasterize_forward(gs["means3D"], gs["opacities"], cam.viewmatrix, cam.projmatrix, cam.campos,
                                            bg, W, H, cam.tanfovx, cam.tanfovy, gs["scales"], gs["rotations"], 1.0,
                                            None, gs["shs"], deg, dtype=dt)
    assert float((img.detach() - fwd["color"]).abs().max()) < 1e-12
    g = O.rasterize_backward(dL, pre, binning, fwd, gs["means3D"], cam.viewmatrix, cam.projmatrix, cam.campos, bg, W,
                             H, cam.tanfovx, cam.tanfovy, gs["scales"], gs["rotations"], 1.0, None, gs["shs"], deg,
                             dtype=dt)
    for name, key in (("means3D", "dL_dmeans3D"), ("scales", "dL_dscales"), ("rotations", "dL_drotations"),
                      ("opacities", "dL_dopacity"), ("shs", "dL_dsh")):
        a, b = g[key], leaf[name].grad
        assert float((a - b).abs().max() / b.abs().max()) < 1e-6, name


def test_reference_gradient_conventions_under_both_clamps():
    """Where the reference's hand-written backward is NOT the true derivative — alpha clamped at 0.99 (gradient flows
    straight through, CR/backward.cu:489-490) and x/z, y/z clamped at 1.3 tanfov (clamped t.x / t.y held constant,
    CR/backward.cu:166-167,259-264) — differentiable_render(clamp_straight_through=True) follows the reference's
    convention, so autograd through it equals the analytic restatement on a scene that hits both clamps.  This is
    what makes it a valid oracle for the camera gradients (row a17), which the reference itself does not produce."""
    P, W, H, deg = 4000, 96, 64, 1
    gs, _, dL, _ = S.make_scene(P, W, H, seed=13, sh_degree=deg)
    gs["opacities"] = torch.clamp(gs["opacities"] * 1.6, max=1.0)
    cam = S.orbit_cameras(W, H, 3, (0.0, 0.0, 4.0), 0.5)[2]
    bg, dt = torch.tensor([0.2, 0.5, 0.7]), torch.float64
    leaf = {k: v.to(dt).clone().requires_grad_(True) for k, v in gs.items()}
    img = O.differentiable_render(leaf["means3D"], leaf["opacities"], leaf["scales"], leaf["rotations"], leaf["shs"],
                                  deg, cam.viewmatrix, cam.projmatrix, cam.campos, bg, W, H, cam.tanfovx, cam.tanfovy,
                                  clamp_straight_through=True)
    (img * dL.to(dt)).sum().backward()
    pre, binning, fwd = O.rasterize_forward(gs["means3D"], gs["opacities"], cam.viewmatrix, cam.projmatrix, cam.campos,
                                            bg, W, H, cam.tanfovx, cam.tanfovy, gs["scales"], gs["rotations"], 1.0,
                                            None, gs["shs"], deg, dtype=dt)
    g = O.rasterize_backward(dL, pre, binning, fwd, gs["means3D"], cam.viewmatrix, cam.projmatrix, cam.campos, bg, W,
                             H, cam.tanfovx, cam.tanfovy, gs["scales"], gs["rotations"], 1.0, None, gs["shs"], deg,
                             dtype=dt)
    cp = O._cov2d_parts(gs["means3D"].to(dt), O.cov3d_from_scale_rot(gs["scales"].to(dt), 1.0, gs["rotations"].to(dt)),
                        cam.viewmatrix.to(dt).reshape(16), W, H, cam.tanfovx, cam.tanfovy)
    vis = pre["radii"] > 0
    assert int(((cp["txtz"].abs() > cp["limx"]) | (cp["tytz"].abs() > cp["limy"]))[vis].sum()) > 0   # frustum clamp hit
    for name, key in (("means3D", "dL_dmeans3D"), ("scales", "dL_dscales"), ("rotations", "dL_drotations"),
                      ("opacities", "dL_dopacity"), ("shs", "dL_dsh")):
        a, b = g[key], leaf[name].grad
        assert float((a - b).abs().max() / b.abs().max()) < 1e-6, name


def test_fp32_oracle_close_to_fp64():
    W, H, deg = 64, 48, 1
    gs, cam, dL, bg = _scene(400, W, H, deg, seed=8)
    outs = []
    for dt in (torch.float32, torch.float64):
        pre, binning, fwd = O.rasterize_forward(gs["means3D"], gs["opacities"], cam.viewmatrix, cam.projmatrix,
                                                cam.campos, bg, W, H, cam.tanfovx, cam.tanfovy, gs["scales"],
                                                gs["rotations"], 1.0, None, gs["shs"], deg, dtype=dt)
        outs.append((pre, binning, fwd))
    assert torch.equal(outs[0][0]["radii"], outs[1][0]["radii"])
    assert torch.equal(outs[0][1]["point_list"], outs[1][1]["point_list"])
    assert float((outs[0][2]["color"].double() - outs[1][2]["color"]).abs().max()) < 1e-4


def test_binning_invariants():
    W, H = 100, 70                      # ragged tile grid 7 x 5
    gs, cam, dL, bg = S.make_scene(2000, W, H, seed=4)
    pre = O.preprocess(gs["means3D"], gs["opacities"], cam.viewmatrix, cam.projmatrix, cam.campos, W, H, cam.tanfovx,
                       cam.tanfovy, gs["scales"], gs["rotations"], 1.0, None, gs["shs"], 0)
    b = O.bin_and_sort(pre, W, H)
    keys = b["keys"]
    assert bool((keys[1:] >= keys[:-1]).all())                       # sortedness
    assert b["num_rendered"] == int(pre["tiles_touched"].sum())
    assert int((b["ranges"][:, 1] - b["ranges"][:, 0]).sum()) == b["num_rendered"]
    # ties keep ascending Gaussian index (stable sort)
    same = keys[1:] == keys[:-1]
    assert bool((b["point_list"][1:][same] > b["point_list"][:-1][same]).all())
    # culled Gaussians (behind the near plane) never appear
    behind = gs["means3D"][:, 2] <= 0.1
    assert int(pre["radii"][behind].abs().sum()) == 0
    assert not bool(torch.isin(b["point_list"], torch.nonzero(behind).reshape(-1)).any())


def test_pearson_restatement_matches_scipy():
    """oracle/loss_oracle.py::pearson_corrcoef restates torchmetrics' definition (not installed, no version pinned by
    the reference); scipy.stats.pearsonr is an independent implementation of the same published coefficient."""
    from scipy import stats

    from oracle import loss_oracle as LO
    g = torch.Generator().manual_seed(3)
    x = torch.randn(5000, generator=g, dtype=torch.float64) * 1.5 + 3.0
    for y in (0.4 * x + torch.randn(5000, generator=g, dtype=torch.float64), -x + 0.1, 1 / (x.abs() + 200.0)):
        r = float(LO.pearson_corrcoef(x, y))
        assert abs(r - stats.pearsonr(x.numpy(), y.numpy())[0]) < 1e-12
    # the loss wrapper: invert_estimate picks the smaller of the two variants (R/utils/loss_utils.py:54-58)
    est = 1.0 / x.abs()
    a = 1 - LO.pearson_corrcoef(-est, x)
    b = 1 - LO.pearson_corrcoef(1 / (est + 200.0), x)
    assert float(LO.pearson_loss(x, est, invert_estimate=True)) == float(min(a, b))


def test_pearson_unmasked_is_per_column():
    """The reference's unmasked pearson_loss (R/utils/loss_utils.py:52-53,60 called with mask=None from
    R/slam/mapper.py:862-868) hands torchmetrics the 2-D [H, W] images: H samples of W outputs, one coefficient per
    image column, averaged by the trailing .mean().  The oracle restates exactly that; scipy per column is the check."""
    from scipy import stats

    from oracle import loss_oracle as LO
    g = torch.Generator().manual_seed(4)
    H, W = 24, 9
    x = torch.randn(H, W, generator=g, dtype=torch.float64)
    y = 0.5 * x + torch.randn(H, W, generator=g, dtype=torch.float64)
    r = LO.pearson_corrcoef(y, x)
    assert r.shape == (W,)
    want = [stats.pearsonr(y[:, c].numpy(), x[:, c].numpy())[0] for c in range(W)]
    assert float((r - torch.tensor(want)).abs().max()) < 1e-12
    loss = float(LO.pearson_loss(x, y, mask=None, invert_estimate=False))
    assert abs(loss - (1 - sum(want) / W)) < 1e-12
    # ... and it is NOT the coefficient of the flattened images
    flat = 1 - stats.pearsonr(y.reshape(-1).numpy(), x.reshape(-1).numpy())[0]
    assert abs(loss - flat) > 1e-4
    cfg = LO.mapper_default(use_gt_depth=False)
    assert cfg["depth_mode"] == LO.DEPTH_PEARSON_COLS and cfg["depth_mask"] == 0

