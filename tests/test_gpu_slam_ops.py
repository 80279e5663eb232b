"""GPU parity of the steps either side of the rasterizer (include/gsloss_b200.h):
gsr_slam_loss against (1) the reference's own l1_loss / ssim outputs committed under tests/golden/loss_*.pt and
(2) the fp64 evaluation of oracle/loss_oracle.py for every loss composition of the mapper and the tracker;
gsr_adam_step against torch.optim.Adam (the reference's optimizer) on the CPU.
Tolerance: 1e-4 relative (BASELINE.json north_star), measured against the largest gradient magnitude."""
import glob
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-4
HERE = os.path.dirname(os.path.abspath(__file__))
LOSS_FILES = sorted(glob.glob(os.path.join(HERE, "golden", "loss_*.pt")))


@pytest.fixture(scope="module")
def ops(built_lib):
    import gsr_slam_ops as m
    return m


def rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).abs().max() / (b.double().abs().max() + 1e-30))


def _cuda(d):
    return {k: v.cuda() for k, v in d.items()}


@pytest.mark.parametrize("path", LOSS_FILES, ids=[os.path.basename(f)[:-3] for f in LOSS_FILES])
def test_l1_ssim_matches_reference_golden(ops, path):
    import gsr_synth as S
    gold = torch.load(path, weights_only=False)
    c = gold["case"]
    d = _cuda(S.make_loss_inputs(c["W"], c["H"], c["seed"]))
    cfg = dict(color_mode=ops.COLOR_L1_SSIM, lambda_dssim=c["lam"])
    losses, g_img, g_dep = ops.slam_loss_and_grads(cfg, d["image"], None, d["gt_color"])
    assert g_dep is None
    losses = losses.cpu()
    assert abs(float(losses[0]) - float(gold["loss"])) < TOL * abs(float(gold["loss"]))
    assert abs(float(losses[3]) - float(gold["ssim"])) < TOL * abs(float(gold["ssim"]))
    assert rel(g_img, gold["dL_dimage"]) < TOL
    cfg = dict(color_mode=ops.COLOR_MASKED_L1_MEAN, color_mask=ops.MASK_SILHOUETTE, sil_threshold=0.99)
    losses, g_img, _ = ops.slam_loss_and_grads(cfg, d["image"], d["depth_image"], d["gt_color"])
    assert abs(float(losses[0]) - float(gold["masked_l1"])) < TOL * abs(float(gold["masked_l1"]))
    assert rel(g_img, gold["masked_dL_dimage"]) < TOL


def _oracle64(cfg, d, target_key):
    from oracle import loss_oracle as LO
    img = d["image"].double().cpu().requires_grad_(True)
    dep = d["depth_image"].double().cpu().requires_grad_(True)
    total, color, depth = LO.slam_loss(cfg, img, dep, d["gt_color"].double().cpu(), d[target_key].double().cpu(),
                                       d["gt_depth"].double().cpu())
    total.backward()
    gi = img.grad if img.grad is not None else torch.zeros_like(img)
    gd = dep.grad if dep.grad is not None else torch.zeros_like(dep)
    return total.detach(), color.detach(), depth.detach(), gi, gd


COMPOSITIONS = [
    # name, builder kwargs, depth target, inputs may hold NaN depths (only the SplaTAM masks exclude them)
    ("mapper_splatam", {}, "gt_depth", True),
    ("mapper_default", dict(use_gt_depth=False), "est_depth", False),
    ("mapper_default", dict(use_gt_depth=True, pearson_weight=0.001), "gt_depth", False),
    ("tracker_splatam", {}, "gt_depth", True),
    ("tracker_default", dict(use_gt_depth=False), "est_depth", False),
    ("tracker_default", dict(use_gt_depth=True, pearson_weight=0.001), "gt_depth", False),
]


@pytest.mark.parametrize("W,H,seed", [(37, 21, 5), (80, 56, 3), (640, 480, 11)])
@pytest.mark.parametrize("name,kw,target,with_nan", COMPOSITIONS, ids=[f"{c[0]}-{c[2]}" for c in COMPOSITIONS])
def test_slam_loss_compositions_match_oracle(ops, name, kw, target, with_nan, W, H, seed):
    """Every loss the mapper (R/slam/mapper.py:839-885) and the tracker (R/slam/tracker.py:110-144) form."""
    import gsr_synth as S
    cfg = getattr(ops, name)(**kw)
    d = _cuda(S.make_loss_inputs(W, H, seed, nan_frac=0.002 if with_nan else 0.0))
    losses, g_img, g_dep = ops.slam_loss_and_grads(cfg, d["image"], d["depth_image"], d["gt_color"], d[target], d["gt_depth"])
    total, color, depth, gi, gd = _oracle64(cfg, d, target)
    losses = losses.cpu().double()
    assert abs(float(losses[0]) - float(total)) <= TOL * abs(float(total))
    assert abs(float(losses[1]) - float(color)) <= TOL * abs(float(color))
    assert abs(float(losses[2]) - float(depth)) <= TOL * abs(float(depth)) + 1e-7
    assert rel(g_img, gi) < TOL
    assert rel(g_dep, gd) < TOL
    assert float(g_dep[1:].abs().max()) == 0.0            # silhouette and depth^2 only feed detached masks


def test_slam_loss_autograd_and_grad_scale(ops):
    """The autograd wrapper scales the stored gradients by the upstream gradient; grad_scale does the same inside
    the kernels; a value-only call (no tensor requires grad) skips the gradient pass."""
    import gsr_synth as S
    d = _cuda(S.make_loss_inputs(96, 64, 9))
    cfg = ops.mapper_splatam()
    img = d["image"].clone().requires_grad_(True)
    dep = d["depth_image"].clone().requires_grad_(True)
    total, terms = ops.slam_loss(cfg, img, dep, d["gt_color"], d["gt_depth"], d["gt_depth"], return_terms=True)
    (3.0 * total).backward()
    _, g_img, g_dep = ops.slam_loss_and_grads(cfg, d["image"], d["depth_image"], d["gt_color"], d["gt_depth"], d["gt_depth"],
                                              grad_scale=3.0)
    assert rel(img.grad, g_img) < 1e-6 and rel(dep.grad, g_dep) < 1e-6
    assert abs(float(terms[0]) - (0.5 * float(terms[1]) + float(terms[2]))) < 1e-6
    before = ops._lib.gsr_launch_count()
    v = ops.slam_loss(cfg, d["image"], d["depth_image"], d["gt_color"], d["gt_depth"], d["gt_depth"])
    assert ops._lib.gsr_launch_count() - before == 2 and abs(float(v) - float(total.detach())) < 1e-7


def test_slam_loss_is_deterministic(ops):
    import gsr_synth as S
    d = _cuda(S.make_loss_inputs(640, 480, 2))
    cfg = ops.mapper_default(use_gt_depth=True)
    a = ops.slam_loss_and_grads(cfg, d["image"], d["depth_image"].nan_to_num(2.0), d["gt_color"], d["gt_depth"], d["gt_depth"])
    b = ops.slam_loss_and_grads(cfg, d["image"], d["depth_image"].nan_to_num(2.0), d["gt_color"], d["gt_depth"], d["gt_depth"])
    assert all(torch.equal(x, y) for x, y in zip(a, b))


def test_slam_loss_empty_mask_is_nan_like_torch(ops):
    """mean over an empty selection is NaN in the reference (torch.mean of nothing); same here, no crash."""
    import gsr_synth as S
    d = _cuda(S.make_loss_inputs(40, 24, 1))
    cfg = ops.mapper_splatam()
    losses, g_img, g_dep = ops.slam_loss_and_grads(cfg, d["image"], d["depth_image"], d["gt_color"],
                                                   torch.zeros_like(d["gt_depth"]), torch.zeros_like(d["gt_depth"]))
    assert torch.isnan(losses[2]) and torch.isfinite(losses[1]) and torch.isfinite(g_img).all()
    assert float(g_dep.nan_to_num(0).abs().max()) == 0.0


@pytest.mark.parametrize("sizes", [(1000 * 3, 1000 * 3, 1000, 1000 * 3, 1000 * 4), (7, 13, 1, 250_001, 3)])
def test_flat_adam_matches_torch_adam(ops, sizes):
    """Five steps of gsr_adam_step against torch.optim.Adam(l, lr=0.0, eps=1e-15) with per-group learning rates
    (R/slam/gaussian_model.py:151-189); segment boundaries not aligned to the kernel's 4-float vectors."""
    g = torch.Generator().manual_seed(4)
    names = ["xyz", "f_dc", "opacity", "scaling", "rotation"]
    lrs = dict(zip(names, [1.6e-4, 2.5e-3, 5e-2, 1e-3, 1e-3]))
    cpu = {k: torch.randn(n, generator=g).requires_grad_(True) for k, n in zip(names, sizes)}
    ref_opt = torch.optim.Adam([{"params": [cpu[k]], "lr": lrs[k], "name": k} for k in names], lr=0.0, eps=1e-15)
    opt = ops.FlatAdam({k: v.detach().cuda() for k, v in cpu.items()}, lrs, eps=1e-15)
    n = sum(sizes)
    for step in range(5):
        grads = torch.randn(n, generator=g) * (torch.rand(n, generator=g) < 0.7)     # 30 % exact zeros (invisible Gaussians)
        if step == 2:
            lrs["xyz"] = 1.0e-4                                                       # schedule (gaussian_model.py:196-202)
            opt.lrs["xyz"] = 1.0e-4
            ref_opt.param_groups[0]["lr"] = 1.0e-4
        off = 0
        for k, sz in zip(names, sizes):
            cpu[k].grad = grads[off: off + sz].clone()
            off += sz
        ref_opt.step()
        gg = grads.cuda()
        opt.step(gg, zero_grads=(step == 4))
    assert float(gg.abs().max()) == 0.0                                               # zero_grads cleared the bucket
    want = torch.cat([cpu[k].detach() for k in names])
    torch.testing.assert_close(opt.flat.cpu(), want, rtol=2e-5, atol=1e-7)
    st = ref_opt.state[cpu["opacity"]]
    off = sizes[0] + sizes[1]
    torch.testing.assert_close(opt.exp_avg[off: off + sizes[2]].cpu(), st["exp_avg"], rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(opt.exp_avg_sq[off: off + sizes[2]].cpu(), st["exp_avg_sq"], rtol=1e-5, atol=1e-12)


def test_flat_adam_grad_scale_is_mean_of_keyframes(ops):
    """grad_scale = 1/K turns the summed keyframe gradients of the sharded map step into their mean."""
    g = torch.Generator().manual_seed(8)
    p0 = torch.randn(4099, generator=g)
    grads = torch.randn(4099, generator=g)
    a = ops.FlatAdam({"p": p0.cuda()}, {"p": 1e-2})
    b = ops.FlatAdam({"p": p0.cuda()}, {"p": 1e-2})
    a.step((grads * 8).cuda(), grad_scale=0.125)
    b.step(grads.cuda())
    torch.testing.assert_close(a.flat, b.flat, rtol=1e-6, atol=1e-8)


def test_map_iteration_render_loss_adam(ops, built_lib):
    """One mapping iteration end to end on the library alone: fused RGB + depth render -> gsr_slam_loss gradients ->
    rasterizer backward into the flat gradient bucket -> gsr_adam_step; the loss must go down over a few iterations."""
    import diff_gaussian_rasterization as dgr
    import gsr_synth as S
    from gsr_mapstep import GradBucket
    W, H, P = 160, 120, 4000
    gs, cam, _, bg = S.make_scene(P, W, H, seed=5)
    dev = torch.device("cuda:0")
    target = {k: v.to(dev) for k, v in gs.items()}
    rs = dgr.GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, bg.to(dev), 1.0, cam.viewmatrix.to(dev),
                                           cam.projmatrix.to(dev), 0, cam.campos.to(dev), False, False)

    def render(p):
        return dgr.GaussianRasterizer(rs)(means3D=p["means3D"], means2D=torch.zeros(P, 3, device=dev), opacities=p["opacities"],
                                          shs=p["shs"], scales=p["scales"], rotations=p["rotations"],
                                          extra_colors=dgr.DEPTH_SILHOUETTE)
    with torch.no_grad():
        gt_color, gt_depth_img, _ = render(target)
    gt_depth = gt_depth_img[0].contiguous()
    g = torch.Generator().manual_seed(1)
    start = {k: (v.cpu() + 0.02 * torch.randn(v.shape, generator=g)).to(dev) for k, v in target.items()}
    start["opacities"] = start["opacities"].clamp(0.01, 0.99)
    start["scales"] = start["scales"].abs() + 1e-4
    opt = ops.FlatAdam(start, {"means3D": 1e-3, "scales": 1e-4, "rotations": 1e-3, "opacities": 1e-3, "shs": 2e-3})
    params = {k: v.requires_grad_(True) for k, v in opt.views.items()}
    bucket = GradBucket(params)
    cfg = ops.mapper_splatam()
    hist = []
    for _ in range(12):
        bucket.zero_()
        bucket.attach()
        image, depth_img, _ = render(params)
        losses, g_img, g_dep = ops.slam_loss_and_grads(cfg, image, depth_img, gt_color, gt_depth, gt_depth)
        torch.autograd.backward([image, depth_img], [g_img, g_dep])
        opt.step(bucket.flat)
        with torch.no_grad():
            params["opacities"].clamp_(0.01, 0.99)
            params["scales"].clamp_(min=1e-4)
        hist.append(float(losses[0]))
    assert hist[-1] < 0.95 * hist[0] and all(h == h for h in hist), hist     # measured ratio ~0.5; wide margin on purpose


def _surgery_case(P, dev, seed=0):
    import gsr_slam_ops as ops
    g = torch.Generator().manual_seed(seed)
    shapes = {"xyz": (3,), "f_dc": (1, 3), "opacity": (1,), "scaling": (3,), "rotation": (4,)}
    params = {k: torch.randn(P, *sh, generator=g).to(dev) for k, sh in shapes.items()}
    opt = ops.FlatAdam(params, {k: 1e-3 for k in shapes})
    opt.steps = 7
    opt.exp_avg.copy_(torch.randn(opt.flat.numel(), generator=g))
    opt.exp_avg_sq.copy_(torch.rand(opt.flat.numel(), generator=g))
    m0, v0 = ({k: t.clone() for k, t in opt._group_views(buf).items()} for buf in (opt.exp_avg, opt.exp_avg_sq))
    return opt, params, shapes, m0, v0, g


@pytest.mark.gpu
@pytest.mark.parametrize("P,frac", [(50, 0.6), (4096, 0.5), (4097, 0.999), (100003, 0.0), (100003, 1.0), (1000000, 0.7)])
def test_flat_adam_surgery_matches_reference_optimizer_surgery(P, frac):
    """FlatAdam.prune / .extend (one mask scan + one gather pass over parameters and both moments) against the
    reference's _prune_optimizer / cat_tensors_to_optimizer semantics (R/slam/gaussian_model.py:380-451) carried out
    with torch indexing / torch.cat: bit-identical rows, zeros for the moments of appended rows, step count untouched.
    Edge cases: nothing kept, everything kept, row counts off the 4096-row scan chunk."""
    dev = torch.device("cuda:0")
    opt, params, shapes, m0, v0, g = _surgery_case(P, dev)
    keep = (torch.rand(P, generator=g) < frac).to(dev)
    views = opt.prune(keep)
    n1 = int(keep.sum())
    assert opt.steps == 7 and opt.flat.numel() == sum(v.numel() for v in views.values()) == 14 * n1
    for k in shapes:
        assert torch.equal(views[k], params[k][keep]) and views[k].shape == (n1,) + shapes[k]
        assert torch.equal(opt._group_views(opt.exp_avg)[k], m0[k][keep])
        assert torch.equal(opt._group_views(opt.exp_avg_sq)[k], v0[k][keep])
        assert views[k].is_contiguous() and (n1 == 0 or views[k].data_ptr() >= opt.flat.data_ptr())
    new = {k: torch.randn(9, *sh, generator=g).to(dev) for k, sh in shapes.items()}
    views = opt.extend(new)
    for k in shapes:
        assert torch.equal(views[k], torch.cat((params[k][keep], new[k]), 0))
        assert torch.equal(opt._group_views(opt.exp_avg)[k], torch.cat((m0[k][keep], torch.zeros_like(new[k])), 0))
        assert torch.equal(opt._group_views(opt.exp_avg_sq)[k], torch.cat((v0[k][keep], torch.zeros_like(new[k])), 0))
    assert list(opt.seg_end) == [(n1 + 9) * c for c in (3, 6, 7, 10, 14)]
    # the optimizer keeps working on the re-laid-out buffers: one step equals torch.optim.Adam with the same state
    grads = torch.randn(opt.flat.numel(), generator=g).to(dev)
    ref_p = opt.flat.clone().requires_grad_(True)
    ref_p.grad = grads.clone()
    ref = torch.optim.Adam([ref_p], lr=1e-3, eps=1e-15)
    ref.state[ref_p] = {"step": torch.tensor(7.0), "exp_avg": opt.exp_avg.clone(), "exp_avg_sq": opt.exp_avg_sq.clone()}
    ref.step()
    opt.step(grads)
    assert float((opt.flat - ref_p.detach()).abs().max()) < 1e-6
