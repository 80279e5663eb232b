#!/usr/bin/env python
"""Workload for ncu captures of the rasterizer kernels: N sequential forward + backward calls (one stream, stock API)
of one C-main keyframe (1M Gaussians, 640x480, SH degree 0).

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python profiles/frame_profile.py 3
    ncu --set full --clock-control none --import-source on -k regex:k_ -s <launches of the warm-up frames> -c <one frame> \
        -o gpurun_out/frame python profiles/frame_profile.py 3

Times under ncu are cold-cache and serialised; bench.py holds the CUDA-event numbers."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mm3dgs-slam_b200")):
    sys.path.insert(0, p)
import diff_gaussian_rasterization as dgr  # noqa: E402
import gsr_synth as S  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
P, W, H = (int(os.environ.get(k, d)) for k, d in (("GSR_P", 1_000_000), ("GSR_W", 640), ("GSR_H", 480)))
dev = torch.device("cuda:0")
gs = S.make_gaussians(P, W, H, seed=0, sh_degree=0)
centroid = gs["means3D"][gs["means3D"][:, 2] > 0.1].mean(0).tolist()
cam = S.orbit_cameras(W, H, 8, centroid, radius=0.5)[0]
bg = torch.zeros(3, device=dev)
rs = dgr.GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, bg, 1.0, cam.viewmatrix.to(dev), cam.projmatrix.to(dev),
                                       0, cam.campos.to(dev), False, False)
p = {k: gs[k].to(dev).requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
dL = torch.randn(3, H, W, generator=torch.Generator().manual_seed(12345)).to(dev)
for _ in range(n):
    m2 = torch.zeros(P, 3, device=dev, requires_grad=True)
    color, _ = dgr.GaussianRasterizer(rs)(means3D=p["means3D"], means2D=m2, opacities=p["opacities"], shs=p["shs"],
                                          scales=p["scales"], rotations=p["rotations"])
    color.backward(dL)
torch.cuda.synchronize()
print("launches", dgr._lib.gsr_launch_count())
