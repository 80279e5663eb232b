"""GPU parity tests: the B200 library (through its C-ABI, via the drop-in Python package) against
(1) the CPU oracle (oracle/gs_oracle.py) and (2) the compiled reference (oracle/_ref) on identical
seeded inputs.  Tolerances: integer / index outputs bit-exact; images and gradients within 1e-4
relative (BASELINE.json north_star)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-4


@pytest.fixture(scope="module")
def dgr(built_lib):
    import diff_gaussian_rasterization as m
    return m


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_api
    if not ref_api.available():
        pytest.skip("compiled reference (oracle/_ref) not present")
    ref_api.load()
    return ref_api


def _run(mod_rasterize, gs, rs, dL, mode):
    """One forward+backward through an autograd rasterizer; returns image, radii and grads."""
    leaves = {k: v.clone().requires_grad_(True) for k, v in gs.items()}
    P = gs["means3D"].shape[0]
    means2D = torch.zeros(P, 3, device=gs["means3D"].device, requires_grad=True)
    kw = dict(scales=leaves["scales"], rotations=leaves["rotations"])
    if mode == "sh":
        kw["shs"] = leaves["shs"]
    elif mode == "colors":
        leaves["colors"] = (gs["shs"][:, 0] * 0.28209479177387814 + 0.5).clamp(min=0).clone().requires_grad_(True)
        kw["colors_precomp"] = leaves["colors"]
    color, radii = mod_rasterize(leaves["means3D"], means2D, leaves["opacities"], rs, **kw)
    (color * dL).sum().backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    grads["means2D"] = means2D.grad
    return color.detach(), radii.detach(), grads


def _ours(dgr):
    def f(means3D, means2D, opacities, rs, **kw):
        return dgr.GaussianRasterizer(rs)(means3D=means3D, means2D=means2D, opacities=opacities, **kw)
    return f


def _theirs(ref):
    def f(means3D, means2D, opacities, rs, **kw):
        return ref.rasterize(means3D, means2D, opacities, rs, **kw)
    return f


CASES = [
    # P, W, H, sh_degree, mode, seed
    (300, 64, 48, 0, "sh", 1),
    (1000, 64, 48, 3, "sh", 2),
    (2000, 128, 96, 1, "sh", 3),
    (5000, 200, 120, 0, "colors", 4),     # ragged: neither dimension a multiple of 16
    (20000, 320, 240, 0, "sh", 5),
    (100000, 320, 240, 0, "sh", 0),       # BASELINE config 2
]


@pytest.mark.parametrize("P,W,H,deg,mode,seed", CASES)
def test_vs_reference(dgr, ref, P, W, H, deg, mode, seed):
    from tests.util import rel_err, scene_on, settings_for
    dev = torch.device("cuda:0")
    gs, cam, dL, bg = scene_on(dev, P, W, H, seed, deg)
    bg = torch.tensor([0.1, 0.3, 0.6], device=dev) if seed % 2 else bg
    rs = settings_for(dgr, cam, bg, deg, dev)
    c1, r1, g1 = _run(_ours(dgr), gs, rs, dL, mode)
    c2, r2, g2 = _run(_theirs(ref), gs, rs, dL, mode)
    assert torch.equal(r1, r2), f"radii differ in {(r1 != r2).sum().item()} of {P}"
    assert rel_err(c1, c2) < TOL
    for k in g2:
        assert torch.isfinite(g1[k]).all(), k
        assert rel_err(g1[k], g2[k]) < TOL, (k, rel_err(g1[k], g2[k]))


@pytest.mark.parametrize("P,W,H,deg,seed", [(3000, 128, 96, 0, 7), (100000, 320, 240, 0, 0), (50000, 640, 480, 2, 9)])
def test_intermediate_state_bit_exact(dgr, ref, P, W, H, deg, seed):
    """tiles_touched, sort keys, sorted Gaussian lists, tile ranges and n_contrib match bit for bit;
    per-Gaussian 2-D quantities match to float rounding."""
    from tests import ws_decode
    from tests.util import rel_err, scene_on, settings_for
    dev = torch.device("cuda:0")
    gs, cam, dL, bg = scene_on(dev, P, W, H, seed, deg)
    rs = settings_for(dgr, cam, bg, deg, dev)
    with torch.no_grad():
        R, color, radii, geom, binning, img = dgr._forward_native(
            gs["means3D"], gs["shs"], None, gs["opacities"], gs["scales"], gs["rotations"], None, rs,
            rs.viewmatrix, rs.projmatrix, rs.campos, rs.bg)
        f = ref.forward(gs["means3D"], gs["opacities"], rs.viewmatrix, rs.projmatrix, rs.campos, rs.bg, W, H,
                        cam.tanfovx, cam.tanfovy, gs["scales"], gs["rotations"], 1.0, None, gs["shs"], deg)
    torch.cuda.synchronize()
    assert R == f["num_rendered"]
    assert torch.equal(radii, f["radii"])
    mg, rg = ws_decode.decode_geom(geom, P, W, H), ref.decode_geom(f["geom"], P)
    vis = (radii > 0).cpu()
    assert torch.equal(mg["tiles_touched"], rg["tiles_touched"])
    for k in ("means2D", "depths", "conic_opacity", "rgb"):
        a, b = mg[k][vis], rg[k][vis]
        assert torch.equal(a, b) or rel_err(a, b) < 1e-6, (k, rel_err(a, b))
    assert torch.equal(mg["depths"][vis], rg["depths"][vis]), "depth bits feed the sort key"
    assert torch.equal(mg["clamped"][vis], rg["clamped"][vis].bool())
    mi, ri = ws_decode.decode_img(img, W, H), ref.decode_img(f["img"], W, H)
    T = mi["ranges"].shape[0]
    assert torch.equal(mi["ranges"], ri["ranges"][:T])
    mb, rb = ws_decode.decode_binning(binning, R, mg, mi), ref.decode_binning(f["binning"], R)
    assert torch.equal(mb["point_list"], rb["point_list"])
    assert torch.equal(mb["keys"], rb["keys"])
    # depth order of the visible Gaussians themselves: ascending (depth bits, index); culled ones are dropped
    nvis = int(vis.sum())
    sid = mg["sorted_ids"][:nvis]
    k = mg["depth_keys"][sid]
    assert bool((k[1:] >= k[:-1]).all())
    tie = k[1:] == k[:-1]
    assert bool((sid[1:][tie] > sid[:-1][tie]).all())
    assert torch.equal(torch.sort(sid).values, torch.nonzero(vis).reshape(-1))
    assert torch.equal(mi["n_contrib"], ri["n_contrib"])
    assert rel_err(mi["final_T"], ri["final_T"]) < 1e-6
    assert rel_err(color, f["color"]) < TOL


@pytest.mark.parametrize("deg,mode", [(0, "sh"), (2, "sh"), (0, "colors")])
def test_vs_cpu_oracle(dgr, deg, mode):
    from oracle import gs_oracle as O
    from tests.util import rel_err, scene_on, settings_for
    dev = torch.device("cuda:0")
    P, W, H = 800, 80, 56
    gs, cam, dL, bg = scene_on(dev, P, W, H, 11, deg)
    bg = torch.tensor([0.2, 0.5, 0.7], device=dev)
    rs = settings_for(dgr, cam, bg, deg, dev)
    c1, r1, g1 = _run(_ours(dgr), gs, rs, dL, mode)
    cpu = {k: v.cpu() for k, v in gs.items()}
    colors = (cpu["shs"][:, 0] * 0.28209479177387814 + 0.5).clamp(min=0) if mode == "colors" else None
    pre, binning, fwd = O.rasterize_forward(cpu["means3D"], cpu["opacities"], cam.viewmatrix, cam.projmatrix,
                                            cam.campos, bg.cpu(), W, H, cam.tanfovx, cam.tanfovy, cpu["scales"],
                                            cpu["rotations"], 1.0, None, None if colors is not None else cpu["shs"],
                                            deg, colors)
    g = O.rasterize_backward(dL.cpu(), pre, binning, fwd, cpu["means3D"], cam.viewmatrix, cam.projmatrix, cam.campos,
                             bg.cpu(), W, H, cam.tanfovx, cam.tanfovy, cpu["scales"], cpu["rotations"], 1.0, None,
                             None if colors is not None else cpu["shs"], deg)
    assert torch.equal(r1.cpu(), pre["radii"])
    assert rel_err(c1, fwd["color"]) < TOL
    assert rel_err(g1["means3D"], g["dL_dmeans3D"]) < 2e-4
    assert rel_err(g1["scales"], g["dL_dscales"]) < 2e-4
    assert rel_err(g1["rotations"], g["dL_drotations"]) < 2e-4
    assert rel_err(g1["opacities"], g["dL_dopacity"]) < 2e-4
    assert rel_err(g1["means2D"], g["dL_dmeans2D"]) < 2e-4
    if mode == "sh":
        assert rel_err(g1["shs"], g["dL_dsh"]) < 2e-4
    else:
        assert rel_err(g1["colors"], g["dL_dcolors"]) < 2e-4


def test_cov3d_precomp_path(dgr, ref):
    from oracle import gs_oracle as O
    from tests.util import rel_err, scene_on, settings_for
    dev = torch.device("cuda:0")
    P, W, H = 4000, 160, 112
    gs, cam, dL, bg = scene_on(dev, P, W, H, 21, 0)
    rs = settings_for(dgr, cam, bg, 0, dev)
    cov = O.cov3d_from_scale_rot(gs["scales"].cpu(), 1.0, gs["rotations"].cpu()).to(dev)
    outs = []
    for fn in (_ours(dgr), _theirs(ref)):
        m = gs["means3D"].clone().requires_grad_(True)
        c = cov.clone().requires_grad_(True)
        o = gs["opacities"].clone().requires_grad_(True)
        s = gs["shs"].clone().requires_grad_(True)
        m2 = torch.zeros(P, 3, device=dev, requires_grad=True)
        color, radii = fn(m, m2, o, rs, shs=s, cov3D_precomp=c)
        (color * dL).sum().backward()
        outs.append((color.detach(), radii, m.grad, c.grad, o.grad, s.grad))
    assert torch.equal(outs[0][1], outs[1][1])
    for a, b in zip(outs[0][2:], outs[1][2:]):
        assert rel_err(a, b) < TOL
    assert rel_err(outs[0][0], outs[1][0]) < TOL


def test_rotated_cameras_and_scale_modifier(dgr, ref):
    import gsr_synth as S
    from tests.util import rel_err, scene_on, settings_for
    dev = torch.device("cuda:0")
    P, W, H = 30000, 320, 240
    gs, _, dL, bg = scene_on(dev, P, W, H, 31, 1)
    for cam in S.orbit_cameras(W, H, 3, (0.0, 0.0, 4.0), 0.5):
        rs = settings_for(dgr, cam, bg, 1, dev, scale_modifier=0.7)
        c1, r1, g1 = _run(_ours(dgr), gs, rs, dL, "sh")
        c2, r2, g2 = _run(_theirs(ref), gs, rs, dL, "sh")
        assert torch.equal(r1, r2)
        assert rel_err(c1, c2) < TOL
        for k in g2:
            assert rel_err(g1[k], g2[k]) < TOL, k


def test_edge_cases(dgr):
    from tests.util import scene_on, settings_for
    dev = torch.device("cuda:0")
    W, H = 48, 32
    gs, cam, dL, bg = scene_on(dev, 64, W, H, 5, 0)
    bg = torch.tensor([0.25, 0.5, 0.75], device=dev)
    rs = settings_for(dgr, cam, bg, 0, dev)
    rast = dgr.GaussianRasterizer(rs)
    # P == 0 -> zero image (reference: DGR/rasterize_points.cu:81), empty radii
    e = {k: v[:0] for k, v in gs.items()}
    color, radii = rast(means3D=e["means3D"], means2D=torch.zeros(0, 3, device=dev), opacities=e["opacities"],
                        shs=e["shs"], scales=e["scales"], rotations=e["rotations"])
    assert color.shape == (3, H, W) and float(color.abs().max()) == 0.0 and radii.numel() == 0
    # everything behind the camera -> background only, all radii 0, zero grads
    m = gs["means3D"].clone()
    m[:, 2] = -1.0
    m.requires_grad_(True)
    color, radii = rast(means3D=m, means2D=torch.zeros(64, 3, device=dev), opacities=gs["opacities"], shs=gs["shs"],
                        scales=gs["scales"], rotations=gs["rotations"])
    assert int(radii.abs().sum()) == 0
    assert torch.allclose(color, bg[:, None, None].expand_as(color))
    color.sum().backward()
    assert float(m.grad.abs().max()) == 0.0
    # argument validation mirrors the reference
    with pytest.raises(Exception):
        rast(means3D=gs["means3D"], means2D=None, opacities=gs["opacities"], scales=gs["scales"], rotations=gs["rotations"])
    with pytest.raises(Exception):
        rast(means3D=gs["means3D"], means2D=None, opacities=gs["opacities"], shs=gs["shs"], scales=gs["scales"])
    with pytest.raises(RuntimeError):
        rast(means3D=gs["means3D"].reshape(-1), means2D=None, opacities=gs["opacities"], shs=gs["shs"],
             scales=gs["scales"], rotations=gs["rotations"])
    # markVisible == near-plane test
    vis = rast.markVisible(gs["means3D"])
    assert torch.equal(vis, gs["means3D"][:, 2] > 0.1)


def test_camera_gradients(dgr):
    """Extension a17: dL/d{viewmatrix, projmatrix, campos} against fp64 autograd through the oracle."""
    import gsr_synth as S
    from oracle import gs_oracle as O
    from tests.util import rel_err, settings_for
    dev = torch.device("cuda:0")
    W, H, deg = 64, 48, 2
    gs, _, dL, _ = S.make_scene(300, W, H, seed=3, sh_degree=deg)
    gs["opacities"] = gs["opacities"] * 0.6          # keep alpha below the 0.99 clamp (see oracle docstring)
    cam = S.orbit_cameras(W, H, 3, (0.0, 0.0, 4.0), 0.5)[1]
    bg = torch.tensor([0.2, 0.5, 0.7])
    dt = torch.float64
    view = cam.viewmatrix.to(dt).clone().requires_grad_(True)
    proj = cam.projmatrix.to(dt).clone().requires_grad_(True)
    campos = cam.campos.to(dt).clone().requires_grad_(True)
    img = O.differentiable_render(gs["means3D"].to(dt), gs["opacities"].to(dt), gs["scales"].to(dt),
                                  gs["rotations"].to(dt), gs["shs"].to(dt), deg, view, proj, campos, bg, W, H,
                                  cam.tanfovx, cam.tanfovy)
    (img * dL.to(dt)).sum().backward()
    v = cam.viewmatrix.to(dev).requires_grad_(True)
    p = cam.projmatrix.to(dev).requires_grad_(True)
    c = cam.campos.to(dev).requires_grad_(True)
    rs = dgr.GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, bg.to(dev), 1.0, v, p, deg, c, False, False)
    d = {k: t.to(dev) for k, t in gs.items()}
    color, _ = dgr.GaussianRasterizer(rs)(means3D=d["means3D"], means2D=torch.zeros(300, 3, device=dev),
                                          opacities=d["opacities"], shs=d["shs"], scales=d["scales"],
                                          rotations=d["rotations"])
    (color * dL.to(dev)).sum().backward()
    assert rel_err(color, img) < TOL
    assert rel_err(v.grad, view.grad) < 1e-3
    assert rel_err(p.grad, proj.grad) < 1e-3
    assert rel_err(c.grad, campos.grad) < 1e-3


def test_grad_targets_accumulate(dgr):
    """Extension: gradients added straight into caller-provided accumulators == autograd accumulation."""
    import gsr_synth as S
    from tests.util import rel_err, scene_on, settings_for
    dev = torch.device("cuda:0")
    P, W, H, deg = 20000, 160, 120, 1
    gs, _, dL, bg = scene_on(dev, P, W, H, 41, deg)
    cams = S.orbit_cameras(W, H, 2, (0.0, 0.0, 4.0), 0.4)
    names = ["means3D", "shs", "opacities", "scales", "rotations"]

    def run(use_targets, overwrite_first=False):
        p = {k: gs[k].clone().requires_grad_(True) for k in names}
        tg = {k: torch.zeros_like(v) for k, v in p.items()} if use_targets else None
        if overwrite_first:     # stale contents everywhere except the opacity accumulator
            for k in names:
                if k != "opacities":
                    tg[k].fill_(123.0)
        outs = []
        for ci, cam in enumerate(cams):
            rs = settings_for(dgr, cam, bg, deg, dev)
            m2 = torch.zeros(P, 3, device=dev, requires_grad=True)
            color, _ = dgr.GaussianRasterizer(rs)(means3D=p["means3D"], means2D=m2, opacities=p["opacities"],
                                                  shs=p["shs"], scales=p["scales"], rotations=p["rotations"],
                                                  grad_targets=dict(tg, _overwrite=True) if (overwrite_first and ci == 0) else tg)
            outs.append(color)
        if overwrite_first:     # the overwriting frame must run first in the backward: do them one by one
            for o in outs:
                o.backward(dL)
        else:
            torch.autograd.backward(outs, [dL] * len(outs))
        if use_targets:
            assert all(v.grad is None for v in p.values())
            return tg
        return {k: v.grad for k, v in p.items()}

    a, b, c = run(False), run(True), run(True, overwrite_first=True)
    for k in names:
        assert rel_err(b[k], a[k]) < 1e-5, k
        assert rel_err(c[k], a[k]) < 1e-5, k


def _slam_case(dev, P=30000, W=320, H=240, deg=0, seed=51):
    import gsr_synth as S
    from tests.util import scene_on
    gs, _, dL, bg = scene_on(dev, P, W, H, seed, deg)
    g = torch.Generator().manual_seed(seed)
    dL2 = torch.randn(3, H, W, generator=g).to(dev)
    w2c = S.look_at_w2c((0.3, -0.1, 0.2), (0.0, 0.0, 4.0)).to(dev)
    return gs, dL, dL2, torch.tensor([0.0, 0.0, 0.0], device=dev), w2c, W, H


def _slam_backward(rasterize, settings_cls, gs, dL, dL2, bg, w2c, W, H, deg=0):
    from tests import slam_glue
    dev = w2c.device
    p = {k: v.clone().requires_grad_(True) for k, v in gs.items()}
    pose = w2c.clone().requires_grad_(True)
    rs = slam_glue.settings(settings_cls, W, H, bg, deg, dev)
    rgb, depth, radii, m2 = slam_glue.render_two_pass(rasterize, rs, p, pose)
    ((rgb * dL).sum() + (depth * dL2).sum()).backward()
    g = {k: v.grad for k, v in p.items()}
    g["pose"], g["means2D"] = pose.grad, m2.grad
    return rgb.detach(), depth.detach(), radii, g


def test_slam_render_pattern_vs_reference(dgr, ref):
    """The reference renderer's call pattern (two passes, shared means2D leaf, python-side pose transform)
    through the drop-in and through the compiled reference: images, all parameter gradients, the pose
    gradient and the accumulated viewspace gradient agree."""
    from collections import namedtuple
    from tests.util import rel_err
    dev = torch.device("cuda:0")
    gs, dL, dL2, bg, w2c, W, H = _slam_case(dev)
    RS = namedtuple("RS", dgr.GaussianRasterizationSettings._fields)
    a = _slam_backward(_ours(dgr), dgr.GaussianRasterizationSettings, gs, dL, dL2, bg, w2c, W, H)
    b = _slam_backward(_theirs(ref), RS, gs, dL, dL2, bg, w2c, W, H)
    assert torch.equal(a[2], b[2])
    assert rel_err(a[0], b[0]) < TOL and rel_err(a[1], b[1]) < TOL
    for k in b[3]:
        assert rel_err(a[3][k], b[3][k]) < TOL, (k, rel_err(a[3][k], b[3][k]))


@pytest.mark.parametrize("deg", [0, 2])
def test_fused_rgb_depth_equals_two_passes(dgr, deg):
    """Extension: one call with extra_colors == the two-pass render, images and gradients."""
    from tests import slam_glue
    from tests.util import rel_err
    dev = torch.device("cuda:0")
    gs, dL, dL2, bg, w2c, W, H = _slam_case(dev, deg=deg, seed=52)
    bg = torch.tensor([0.2, 0.4, 0.1], device=dev)
    a = _slam_backward(_ours(dgr), dgr.GaussianRasterizationSettings, gs, dL, dL2, bg, w2c, W, H, deg)
    p = {k: v.clone().requires_grad_(True) for k, v in gs.items()}
    pose = w2c.clone().requires_grad_(True)
    rs = slam_glue.settings(dgr.GaussianRasterizationSettings, W, H, bg, deg, dev)
    means_cam = slam_glue.camera_frame(p, pose)
    m2 = torch.zeros_like(means_cam, requires_grad=True)
    rgb, depth, radii = dgr.GaussianRasterizer(rs)(means3D=means_cam, means2D=m2, opacities=p["opacities"], shs=p["shs"],
                                                   scales=p["scales"], rotations=p["rotations"],
                                                   extra_colors=slam_glue.depth_silhouette(means_cam))
    ((rgb * dL).sum() + (depth * dL2).sum()).backward()
    assert torch.equal(radii, a[2])
    assert rel_err(rgb, a[0]) < 1e-6 and rel_err(depth, a[1]) < 1e-6
    g = {k: v.grad for k, v in p.items()}
    g["pose"], g["means2D"] = pose.grad, m2.grad
    for k in g:
        assert rel_err(g[k], a[3][k]) < 2e-5, (k, rel_err(g[k], a[3][k]))


@pytest.mark.parametrize("P,W,H,deg,seed", [
    (150000, 1920, 1080, 0, 61),   # 8160 tiles: partition with 8 warps/CTA and ~160 KB of per-warp counters
    (60000, 2560, 1440, 1, 62),    # 14400 tiles: 4 warps/CTA
    (3000, 16, 16, 0, 63),         # a single tile
    (257, 33, 17, 2, 64),          # ragged everything, P just over one block
])
def test_image_size_extremes_vs_reference(dgr, ref, P, W, H, deg, seed):
    from tests import ws_decode
    from tests.util import rel_err, scene_on, settings_for
    dev = torch.device("cuda:0")
    gs, cam, dL, bg = scene_on(dev, P, W, H, seed, deg)
    rs = settings_for(dgr, cam, bg, deg, dev)
    c1, r1, g1 = _run(_ours(dgr), gs, rs, dL, "sh")
    c2, r2, g2 = _run(_theirs(ref), gs, rs, dL, "sh")
    assert torch.equal(r1, r2)
    assert rel_err(c1, c2) < TOL
    for k in g2:
        assert rel_err(g1[k], g2[k]) < TOL, (k, rel_err(g1[k], g2[k]))
    # and the instance lists, bit for bit
    with torch.no_grad():
        R, _, radii, geom, binning, img = dgr._forward_native(
            gs["means3D"], gs["shs"], None, gs["opacities"], gs["scales"], gs["rotations"], None, rs,
            rs.viewmatrix, rs.projmatrix, rs.campos, rs.bg)
        f = ref.forward(gs["means3D"], gs["opacities"], rs.viewmatrix, rs.projmatrix, rs.campos, rs.bg, W, H,
                        cam.tanfovx, cam.tanfovy, gs["scales"], gs["rotations"], 1.0, None, gs["shs"], deg)
    torch.cuda.synchronize()
    assert R == f["num_rendered"]
    mg, mi = ws_decode.decode_geom(geom, P, W, H), ws_decode.decode_img(img, W, H)
    mb, rb = ws_decode.decode_binning(binning, R, mg, mi), ref.decode_binning(f["binning"], R)
    assert torch.equal(mb["point_list"], rb["point_list"]) and torch.equal(mb["keys"], rb["keys"])
    assert torch.equal(mi["n_contrib"], ref.decode_img(f["img"], W, H)["n_contrib"])


def test_huge_splats_and_duplicates(dgr, ref):
    """Splats covering the whole tile grid, many exactly equal depths (ties must keep index order) and
    duplicated Gaussians."""
    from tests.util import rel_err, scene_on, settings_for
    dev = torch.device("cuda:0")
    P, W, H = 4000, 320, 240
    gs, cam, dL, bg = scene_on(dev, P, W, H, 71, 0)
    gs["scales"][:200] *= 400.0                       # whole-screen splats
    gs["means3D"][200:1200, 2] = 2.5                   # a thousand identical depths
    gs["means3D"][1200:1400] = gs["means3D"][1400:1600]   # exact duplicates
    gs["scales"][1200:1400] = gs["scales"][1400:1600]
    rs = settings_for(dgr, cam, bg, 0, dev)
    c1, r1, g1 = _run(_ours(dgr), gs, rs, dL, "sh")
    c2, r2, g2 = _run(_theirs(ref), gs, rs, dL, "sh")
    assert torch.equal(r1, r2)
    assert rel_err(c1, c2) < TOL
    for k in g2:
        assert rel_err(g1[k], g2[k]) < TOL, (k, rel_err(g1[k], g2[k]))
    with torch.no_grad():
        R, _, _, geom, binning, img = dgr._forward_native(
            gs["means3D"], gs["shs"], None, gs["opacities"], gs["scales"], gs["rotations"], None, rs,
            rs.viewmatrix, rs.projmatrix, rs.campos, rs.bg)
        f = ref.forward(gs["means3D"], gs["opacities"], rs.viewmatrix, rs.projmatrix, rs.campos, rs.bg, W, H,
                        cam.tanfovx, cam.tanfovy, gs["scales"], gs["rotations"], 1.0, None, gs["shs"], 0)
    from tests import ws_decode
    torch.cuda.synchronize()
    mg, mi = ws_decode.decode_geom(geom, P, W, H), ws_decode.decode_img(img, W, H)
    mb, rb = ws_decode.decode_binning(binning, R, mg, mi), ref.decode_binning(f["binning"], R)
    assert R == f["num_rendered"] and torch.equal(mb["point_list"], rb["point_list"])


def test_prepared_forward_equals_plain(dgr):
    """prepare_forward() + forward(prepared=...) == plain forward, for several frames prepared up front."""
    import gsr_synth as S
    from tests.util import scene_on, settings_for
    dev = torch.device("cuda:0")
    P, W, H = 20000, 160, 120
    gs, _, dL, bg = scene_on(dev, P, W, H, 81, 0)
    cams = S.orbit_cameras(W, H, 3, (0.0, 0.0, 4.0), 0.4)
    rss = [settings_for(dgr, c, bg, 0, dev) for c in cams]
    kw = dict(shs=gs["shs"], scales=gs["scales"], rotations=gs["rotations"])
    handles = [dgr.prepare_forward(gs["means3D"], gs["opacities"], rs, **kw) for rs in rss]
    for rs, h in zip(rss, handles):
        m = gs["means3D"].detach().clone().requires_grad_(True)
        a, ra = dgr.GaussianRasterizer(rs)(means3D=m, means2D=torch.zeros(P, 3, device=dev), opacities=gs["opacities"], **kw)
        (a * dL).sum().backward()
        m2 = gs["means3D"].requires_grad_(True)
        m2.grad = None
        b, rb = dgr.GaussianRasterizer(rs)(means3D=m2, means2D=torch.zeros(P, 3, device=dev), opacities=gs["opacities"],
                                           prepared=h, **kw)
        (b * dL).sum().backward()
        assert torch.equal(ra, rb) and torch.equal(a, b)
        assert torch.allclose(m.grad, m2.grad, rtol=1e-4, atol=1e-6 * float(m.grad.abs().max()))
    with pytest.raises(RuntimeError):
        dgr.GaussianRasterizer(rss[0])(means3D=gs["means3D"].detach().clone(), means2D=None, opacities=gs["opacities"],
                                       prepared=handles[1], **kw)


def test_native_pose_and_generated_depth_colours(dgr):
    """Extension (§8f-2, first step): the camera pose goes in as viewmatrix/projmatrix (gradients through the
    library's camera gradients) and the depth/silhouette colours are generated inside the library — no
    per-Gaussian host op.  Against the reference-style path (python-side pose transform + explicit [z,1,z^2])
    on an isotropic scene (the two modes differ for anisotropic splats by the reference's own quirk: in the
    python-transform mode the rotations are not moved to the camera frame)."""
    import gsr_synth as S
    from tests import slam_glue
    from tests.util import rel_err, scene_on
    dev = torch.device("cuda:0")
    P, W, H = 20000, 320, 240
    gs, _, dL, _ = scene_on(dev, P, W, H, 91, 0)
    gs["scales"] = gs["scales"][:, :1].repeat(1, 3).contiguous()
    g = torch.Generator().manual_seed(5)
    dL2 = torch.randn(3, H, W, generator=g).to(dev)
    bg = torch.zeros(3, device=dev)
    from oracle.gs_oracle import quat_to_rot

    def pose_from(qt):          # (quaternion, translation) -> 4x4 world-to-camera, like the SLAM pose tensor
        q = qt[:4] / qt[:4].norm()
        top = torch.cat([quat_to_rot(q), qt[4:, None]], 1)
        return torch.cat([top, torch.tensor([[0.0, 0.0, 0.0, 1.0]], device=qt.device)], 0)

    qt0 = torch.tensor([0.98, 0.05, -0.12, 0.08, 0.1, -0.05, 0.2], device=dev)

    # reference-style
    p = {k: v.clone().requires_grad_(True) for k, v in gs.items()}
    qt1 = qt0.clone().requires_grad_(True)
    pose = pose_from(qt1)
    rs = slam_glue.settings(dgr.GaussianRasterizationSettings, W, H, bg, 0, dev)
    mc = slam_glue.camera_frame(p, pose)
    rgb, depth, radii = dgr.GaussianRasterizer(rs)(means3D=mc, means2D=torch.zeros_like(mc, requires_grad=True),
                                                   opacities=p["opacities"], shs=p["shs"], scales=p["scales"],
                                                   rotations=p["rotations"], extra_colors=slam_glue.depth_silhouette(mc))
    ((rgb * dL).sum() + (depth * dL2).sum()).backward()

    # native
    q = {k: v.clone().requires_grad_(True) for k, v in gs.items()}
    qt2 = qt0.clone().requires_grad_(True)
    pose2 = pose_from(qt2)
    cam = S.make_camera(W, H)
    view = pose2.t()
    proj = view @ S.projection_matrix(*S.intrinsics(W, H), W, H).t().to(dev)
    campos = torch.linalg.inv(view)[3, :3]
    rs2 = dgr.GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, bg, 1.0, view, proj, 0, campos, False, False)
    rgb2, depth2, radii2 = dgr.GaussianRasterizer(rs2)(means3D=q["means3D"], means2D=torch.zeros(P, 3, device=dev),
                                                       opacities=q["opacities"], shs=q["shs"], scales=q["scales"],
                                                       rotations=q["rotations"], extra_colors=dgr.DEPTH_SILHOUETTE)
    ((rgb2 * dL).sum() + (depth2 * dL2).sum()).backward()
    assert (radii != radii2).float().mean() < 1e-3          # fp order of the pose transform differs between the modes
    # (a handful of splats land on a different radius / tile rectangle in the two modes, hence the loose bars)
    assert rel_err(rgb2, rgb) < 5e-3 and rel_err(depth2, depth) < 5e-3
    for k in ("means3D", "opacities", "shs"):
        assert rel_err(q[k].grad, p[k].grad) < 5e-3, (k, rel_err(q[k].grad, p[k].grad))
    # per-axis scale gradients depend on the splat's orientation relative to the camera (which the python-transform
    # mode does not rotate); the gradient w.r.t. the shared isotropic scale is the invariant quantity
    assert rel_err(q["scales"].grad.sum(1), p["scales"].grad.sum(1)) < 5e-3
    # pose gradient in the SLAM parameterisation (unit quaternion + translation): the unconstrained matrix
    # gradients differ (the native mode also differentiates cov2D w.r.t. the view rotation), their projections
    # onto valid poses must agree
    assert rel_err(qt2.grad, qt1.grad) < 5e-3, (qt2.grad, qt1.grad)


def test_scaling_config_5m_1080p(dgr, ref):
    """BASELINE config 5 geometry (5M Gaussians, 1920x1080): instance count, radii, the whole per-tile
    depth-ordered list (compared on the device), contributor counts and the image against the reference."""
    from tests.util import rel_err, scene_on, settings_for
    dev = torch.device("cuda:0")
    P, W, H = 5_000_000, 1920, 1080
    gs, cam, dL, bg = scene_on(dev, P, W, H, 0, 0)
    rs = settings_for(dgr, cam, bg, 0, dev)
    with torch.no_grad():
        R, color, radii, geom, binning, img = dgr._forward_native(
            gs["means3D"], gs["shs"], None, gs["opacities"], gs["scales"], gs["rotations"], None, rs,
            rs.viewmatrix, rs.projmatrix, rs.campos, rs.bg)
        f = ref.forward(gs["means3D"], gs["opacities"], rs.viewmatrix, rs.projmatrix, rs.campos, rs.bg, W, H,
                        cam.tanfovx, cam.tanfovy, gs["scales"], gs["rotations"], 1.0, None, gs["shs"], 0)
    torch.cuda.synchronize()
    assert R == f["num_rendered"] and R > 10_000_000
    assert torch.equal(radii, f["radii"])
    # both libraries keep the sorted Gaussian-id list first in their binning buffer
    mine = binning[: 4 * R].view(torch.int32)
    theirs = f["binning"][: 4 * R].view(torch.int32)
    assert torch.equal(mine, theirs)
    assert rel_err(color, f["color"]) < TOL
    # backward runs and stays finite at this size
    m = gs["means3D"].clone().requires_grad_(True)
    c, _ = dgr.GaussianRasterizer(rs)(means3D=m, means2D=torch.zeros(P, 3, device=dev), opacities=gs["opacities"],
                                      shs=gs["shs"], scales=gs["scales"], rotations=gs["rotations"])
    (c * dL).sum().backward()
    assert bool(torch.isfinite(m.grad).all()) and float(m.grad.abs().max()) > 0
