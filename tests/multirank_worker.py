"""Worker of tests/test_gpu_multirank.py (one process per GPU under torchrun): the library's peer-memory exchange
against NCCL, and the keyframe-sharded map step at world_size N against single-rank gradient accumulation."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mm3dgs-slam_b200")):
    sys.path.insert(0, p)

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import diff_gaussian_rasterization as dgr  # noqa: E402
import gsr_synth as S  # noqa: E402
from gsr_mapstep import PeerExchange, ShardedMapStep  # noqa: E402

modes = ["peer"] + (["nvls"] if os.environ.get("GSR_TEST_NVLS", "1") == "1" else [])
for mode in modes:
    for n in (1, 5, 1000, 14 * 100003, 14_000_000):
        try:
            ex = PeerExchange(n, dev, None, mode)
        except RuntimeError as e:
            if mode == "nvls":
                if rank == 0:
                    print("nvls unavailable:", e)
                break
            raise
        g = torch.Generator().manual_seed(100 * rank + 7)
        for rep in range(3):          # repeated calls: epochs advance, flags are never reset
            x = torch.randn(n, generator=g).to(dev)
            ex.flat.copy_(x)
            want = x.clone()
            dist.all_reduce(want)
            ex.all_reduce()
            torch.cuda.synchronize()
            tol = 1e-5 * float(want.abs().max()) + 1e-6
            assert float((ex.flat - want).abs().max()) <= tol, (mode, n, rep, float((ex.flat - want).abs().max()))
            # every rank holds the SAME bits (each element is reduced once, by its owner)
            ref = ex.flat.clone()
            dist.broadcast(ref, 0)
            assert torch.equal(ref, ex.flat), (mode, n, rep, "ranks differ")
            # all-gather: every rank contributes its slice; the result is the concatenation of the owners' slices
            mine = torch.randn(n, generator=g).to(dev)
            ex.flat.copy_(mine)
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            want = mine.clone()
            for r in range(world):
                lo, hi = ex.slice_range(r)
                want[lo:hi] = parts[r][lo:hi]
            ex.all_gather()
            torch.cuda.synchronize()
            assert torch.equal(ex.flat, want), (mode, n, rep, "all_gather")
        del ex
if rank == 0:
    print("peer exchange ok:", modes)

# keyframe-sharded map step: world ranks == single-rank accumulation over the same keyframes
P, W, H = 20000, 160, 120
gs = S.make_gaussians(P, W, H, seed=0, sh_degree=0)
dL = torch.randn(3, H, W, generator=torch.Generator().manual_seed(3)).to(dev)
bg = torch.zeros(3, device=dev)
names = ["means3D", "shs", "opacities", "scales", "rotations"]
params = {k: gs[k].to(dev).requires_grad_(True) for k in names}


def forward_fn(p, rs, targets=None):
    m2 = torch.zeros(P, 3, device=dev, requires_grad=True)
    color, _ = dgr.GaussianRasterizer(rs)(means3D=p["means3D"], means2D=m2, opacities=p["opacities"], shs=p["shs"],
                                          scales=p["scales"], rotations=p["rotations"], grad_targets=targets)
    return color, dL


# K keyframes: 2 per rank (several frames add into the bucket) and 1 per rank (overwrite mode, no zero fill)
for K in (2 * world, world):
    cams = S.orbit_cameras(W, H, K, (0.0, 0.0, 4.0), radius=0.5)
    kfs = [dgr.GaussianRasterizationSettings(c.H, c.W, c.tanfovx, c.tanfovy, bg, 1.0, c.viewmatrix.to(dev),
                                             c.projmatrix.to(dev), 0, c.campos.to(dev), False, False) for c in cams]
    solo = ShardedMapStep(params, forward_fn=forward_fn, streams=1, direct_targets=True, exchange="nccl")
    solo.world, solo.rank = 1, 0      # single-rank ground truth: accumulate all K keyframes locally, no exchange
    solo.step(kfs)
    torch.cuda.synchronize()
    want = solo.bucket.flat.clone()
    for exch in ("nccl", "auto"):
        st = ShardedMapStep(params, forward_fn=forward_fn, streams=2, direct_targets=True, exchange=exch)
        for _ in range(3):
            st.step(kfs)
        torch.cuda.synchronize()
        got = st.bucket.flat.clone()
        err = float((got - want).abs().max() / want.abs().max())
        assert err < 1e-4, (K, exch, st.exchange, err)
        ref = got.clone()
        dist.broadcast(ref, 0)
        assert exch == "nccl" or torch.equal(ref, got), "ranks hold different bits"
        if rank == 0:
            print(f"map step K={K} exchange={st.exchange}: world {world} vs single-rank accumulation rel err {err:.2e}")
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("MULTIRANK OK")
