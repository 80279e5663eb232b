#!/usr/bin/env python
"""Workload for the ncu launch list of the loss / optimizer kernels (include/gsloss_b200.h):

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --csv --log-file gpurun_out/ops_launches.csv python profiles/ops_profile.py

Three gsr_slam_loss calls (mapper composition, 640x480, value + gradients) and three gsr_adam_step calls over the
14M-float bucket of 1M Gaussians.  Times under ncu are cold-cache and serialised; bench.py's `iteration_ops` holds
the CUDA-event numbers."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mm3dgs-slam_b200")):
    sys.path.insert(0, p)
import gsr_slam_ops as ops  # noqa: E402
import gsr_synth as S  # noqa: E402

dev = torch.device("cuda:0")
d = {k: v.to(dev) for k, v in S.make_loss_inputs(640, 480, 3).items()}
cfg = ops.mapper_splatam()
for _ in range(3):
    ops.slam_loss_and_grads(cfg, d["image"], d["depth_image"], d["gt_color"], d["gt_depth"], d["gt_depth"])
n = 14_000_000
opt = ops.FlatAdam({"p": torch.randn(n, device=dev)}, {"p": 1e-4})
g = torch.randn(n, device=dev)
for _ in range(3):
    opt.step(g)
torch.cuda.synchronize()
