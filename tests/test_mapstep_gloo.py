"""Multi-rank host logic of the keyframe-sharded map step on CPU (gloo, world_size 2):
sharded step == single-rank gradient accumulation over the same keyframes, and the bucket is what
the collective runs on (no gather copy).  The per-keyframe render here is the CPU oracle's
differentiable forward (test infrastructure) — the product rasterizer has no CPU path."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import gsr_synth as S
from gsr_mapstep import GradBucket, ShardedMapStep, shard_keyframes


def _scene():
    W, H = 32, 32
    gs, _, dL, _ = S.make_scene(60, W, H, seed=5, sh_degree=0)
    cams = S.orbit_cameras(W, H, 4, (0.0, 0.0, 4.0), 0.3)
    return gs, cams, dL, W, H


def _frame_fn(dL, W, H):
    from oracle import gs_oracle as O

    def f(p, cam):
        img = O.differentiable_render(p["means3D"], p["opacities"], p["scales"], p["rotations"], p["shs"], 0,
                                      cam.viewmatrix, cam.projmatrix, cam.campos, torch.zeros(3), W, H, cam.tanfovx,
                                      cam.tanfovy)
        (img * dL.double()).sum().backward()
        return img.detach()
    return f


def _params(gs):
    return {k: gs[k].clone().requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gs, cams, dL, W, H = _scene()
    step = ShardedMapStep(_params(gs), _frame_fn(dL, W, H))
    step.step(cams)
    if rank == 0:
        torch.save(step.bucket.flat.clone(), out)
    dist.destroy_process_group()


def test_shard_assignment():
    assert shard_keyframes(8, 0, 8) == [0] and shard_keyframes(8, 3, 4) == [3, 7]
    assert shard_keyframes(3, 3, 4) == []          # idle rank still joins the all-reduce
    assert sorted(sum((shard_keyframes(11, r, 4) for r in range(4)), [])) == list(range(11))


def test_bucket_views_alias_flat():
    gs, *_ = _scene()
    p = _params(gs)
    b = GradBucket(p)
    assert b.flat.numel() == sum(v.numel() for v in p.values()) == 60 * 14     # 56 B per Gaussian at SH-0
    p["scales"].grad.add_(1.0)
    assert float(b.flat.sum()) == 60 * 3
    assert p["means3D"].grad.data_ptr() == b.flat.data_ptr()


def test_sharded_equals_accumulated(tmp_path):
    gs, cams, dL, W, H = _scene()
    single = ShardedMapStep(_params(gs), _frame_fn(dL, W, H))
    single.step(cams)
    ref = single.bucket.flat.clone()
    assert float(ref.abs().max()) > 0
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "bucket.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    assert torch.allclose(got, ref, rtol=1e-5, atol=1e-6 * float(ref.abs().max()))


def test_two_phase_schedule_matches_per_frame():
    """forward_fn mode (all forwards, then one autograd.backward) == frame_fn mode."""
    from oracle import gs_oracle as O
    gs, cams, dL, W, H = _scene()

    def fwd(p, cam):
        img = O.differentiable_render(p["means3D"], p["opacities"], p["scales"], p["rotations"], p["shs"], 0,
                                      cam.viewmatrix, cam.projmatrix, cam.campos, torch.zeros(3), W, H, cam.tanfovx,
                                      cam.tanfovy)
        return img, dL.double()

    a = ShardedMapStep(_params(gs), _frame_fn(dL, W, H))
    a.step(cams)
    b = ShardedMapStep(_params(gs), forward_fn=fwd)
    out = b.step(cams)
    assert len(out) == len(cams)
    assert torch.allclose(a.bucket.flat, b.bucket.flat, rtol=1e-6, atol=1e-7 * float(a.bucket.flat.abs().max()))


def _stats_worker(rank, world, port, out):
    from gsr_mapstep import DensificationStats
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    st = DensificationStats(50, "cpu")
    for kf in shard_keyframes(4, rank, world):
        g = torch.Generator().manual_seed(100 + kf)
        st.add(torch.randint(0, 5, (50,), generator=g), torch.randn(50, 3, generator=g))
    res = st.reduce()
    if rank == 0:
        torch.save(res, out)
    dist.destroy_process_group()


def test_densification_stats_reduction(tmp_path):
    from gsr_mapstep import DensificationStats
    st = DensificationStats(50, "cpu")
    for kf in range(4):
        g = torch.Generator().manual_seed(100 + kf)
        st.add(torch.randint(0, 5, (50,), generator=g), torch.randn(50, 3, generator=g))
    want = st.reduce()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "stats.pt")
    mp.spawn(_stats_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    for a, b in zip(got, want):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-6)
    assert float(want[1].max()) <= 4 and float(want[2].max()) <= 4


def _targets_forward_fn(dL, W, H, scale):
    """A CPU stand-in for the library's grad_targets contract (rasterize_gaussians docstring): the gradients of a frame
    are ADDED into the accumulators the step hands out, or — when the dict carries _overwrite — stored, except for the
    opacity accumulator, which is always added into (the blend backward reduces into it) and which the step zeroes."""
    from oracle import gs_oracle as O

    def fwd(p, cam, targets):
        leaves = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}

        def sink(k, store):
            def hook(g):
                with torch.no_grad():
                    (targets[k].copy_ if store else targets[k].add_)(g.to(targets[k].dtype))
            return hook
        for k, leaf in leaves.items():
            leaf.register_hook(sink(k, bool(targets.get("_overwrite")) and k != "opacities"))
        img = O.differentiable_render(leaves["means3D"], leaves["opacities"], leaves["scales"], leaves["rotations"],
                                      leaves["shs"], 0, cam.viewmatrix, cam.projmatrix, cam.campos, torch.zeros(3), W, H,
                                      cam.tanfovx, cam.tanfovy)
        return img, dL.double() * scale[0]
    return fwd


def _targets_worker(rank, world, port, out, K):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gs, cams, dL, W, H = _scene()
    scale = [3.0]
    step = ShardedMapStep(_params(gs), forward_fn=_targets_forward_fn(dL, W, H, scale), direct_targets=True)
    step.step(cams[:K])            # a first step with other gradients: nothing of it may survive into the second
    scale[0] = 1.0
    step.step(cams[:K])
    if rank == 0:
        torch.save((step.bucket.flat.clone(), step._single), out)
    dist.destroy_process_group()


def test_direct_targets_single_keyframe_overwrites(tmp_path):
    """direct_targets: a rank that owns ONE keyframe runs in overwrite mode (only the opacity slice of the bucket is
    zeroed), a rank with several accumulates into a zeroed bucket; either way two consecutive steps at world_size 2
    equal single-rank accumulation of the second step's gradients."""
    gs, cams, dL, W, H = _scene()
    for K, single in ((2, True), (4, False)):
        ref = ShardedMapStep(_params(gs), _frame_fn(dL, W, H))
        ref.step(cams[:K])
        want = ref.bucket.flat.clone()
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        out = str(tmp_path / f"bucket_{K}.pt")
        mp.spawn(_targets_worker, args=(2, port, out, K), nprocs=2, join=True)
        got, was_single = torch.load(out)
        assert was_single == single
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-6 * float(want.abs().max())), K
