#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native Gaussian rasterizer (BASELINE.json metric).

Metric: forward+backward frames/s at 1M Gaussians, 640x480 (config "C-main").
A *step* is one 8-keyframe mapping step: every keyframe gets one rasterizer forward + backward
(SH path, fixed random dL/dpix), parameter gradients accumulate into one flat bucket, and with
N > 1 GPUs the keyframes are sharded r, r+N, ... with a single NCCL all-reduce of the bucket
(strong scaling: 8 keyframes per step whatever N).  value = 8*steps / time = whole-job frames/s;
at N = 1 this is exactly the single-GPU fwd+bwd frames/s of the metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

`--impl reference` times the UNMODIFIED reference rasterizer compiled from /root/reference
(oracle/_ref, its own CUDA code path through its own pybind API) on the same workload; the
reference has no multi-GPU path, so with N > 1 rank 0 alone runs all keyframes.  If the compiled
reference is absent the CPU oracle port is timed on a bounded sample instead.

Timing rules followed: >= 3 warm-up steps, CUDA events on the launching stream, barrier +
synchronize on both sides, max over ranks, nvidia-smi clocks sampled during the timed region.
Each step's working set (params + per-frame workspaces + gradients, several hundred MB) exceeds
the 126 MB L2, so no explicit flush is needed between iterations (stated in config.l2).
"""
import argparse
import ctypes
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "mm3dgs-slam_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import gsr_synth as S  # noqa: E402

METRIC = "fwd+bwd frames/s at 1M Gaussians 640x480 (8-keyframe map step)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--P", type=int, default=1_000_000)
    ap.add_argument("--W", type=int, default=640)
    ap.add_argument("--H", type=int, default=480)
    ap.add_argument("--keyframes", type=int, default=8)
    ap.add_argument("--sh-degree", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip m1/m2/m3/iteration_ops/parity (scaling sweeps)")
    ap.add_argument("--repeats", type=int, default=5, help="timed blocks of --steps steps; the line reports the median")
    ap.add_argument("--config", default="cmain", choices=["cmain", "c5", "custom"],
                    help="cmain: 1M / 640x480 / K=8 (the metric's configuration); c5: BASELINE config 5, 5M / 1920x1080 / "
                         "K=32; custom: take --P --W --H --keyframes as given")
    a = ap.parse_args()
    if a.config == "c5":
        a.P, a.W, a.H, a.keyframes = 5_000_000, 1920, 1080, 32
    elif (a.P, a.W, a.H, a.keyframes) != (1_000_000, 640, 480, 8):
        a.config = "custom"
    return a


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for n, v in zip(names, r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(n)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# algorithmic bytes per stage: P Gaussians, R tile instances, N pixels, M SH coeffs
def stage_bytes(P, R, N, M=1):
    g_in = 44 + 12 * M
    return {
        "preprocess_fwd": (g_in + 48) * P,
        "depth_sort": 8 * P,                       # SURVEY's "scan" row: the per-Gaussian ordering pass
        "tile_partition": (12 + 24 + 8) * R,       # duplicate + ideal sort + range rows (tile_totals is folded into it)
        "render_fwd": 40 * R + 20 * N,
        "render_bwd": 76 * R + 20 * N,
        "preprocess_bwd": (g_in + 36 + 48 + g_in) * P,
    }


# ------------------------------------------------------------------------------------------------
def build_workload(args, device):
    gs = S.make_gaussians(args.P, args.W, args.H, seed=0, sh_degree=args.sh_degree)
    centroid = gs["means3D"][gs["means3D"][:, 2] > 0.1].mean(0).tolist()
    cams = S.orbit_cameras(args.W, args.H, args.keyframes, centroid, radius=0.5)
    g = torch.Generator().manual_seed(12345)
    dL = torch.randn(3, args.H, args.W, generator=g)
    return gs, cams, dL


def settings_list(mod_settings, cams, bg, sh_degree, device):
    return [mod_settings(image_height=c.H, image_width=c.W, tanfovx=c.tanfovx, tanfovy=c.tanfovy, bg=bg,
                         scale_modifier=1.0, viewmatrix=c.viewmatrix.to(device), projmatrix=c.projmatrix.to(device),
                         sh_degree=sh_degree, campos=c.campos.to(device), prefiltered=False, debug=False)
            for c in cams]


def cpu_baseline(args, gs, cam, dL, tile_stride=12):
    """Oracle port (pure PyTorch, CPU) on a bounded sample of the same workload: full per-Gaussian
    stages + sort, blend fwd/bwd on every `tile_stride`-th tile, scaled to the full tile count."""
    from oracle import gs_oracle as O
    bg = torch.zeros(3)
    t0 = time.perf_counter()
    pre = O.preprocess(gs["means3D"], gs["opacities"], cam.viewmatrix, cam.projmatrix, cam.campos, cam.W, cam.H,
                       cam.tanfovx, cam.tanfovy, gs["scales"], gs["rotations"], 1.0, None, gs["shs"], args.sh_degree)
    binning = O.bin_and_sort(pre, cam.W, cam.H)
    t1 = time.perf_counter()
    T = pre["grid"][0] * pre["grid"][1]
    tiles = list(range(0, T, tile_stride))
    fwd = O.render_forward(pre, binning, pre["rgb"], bg, cam.W, cam.H, tiles=tiles)
    rb = O.render_backward(pre, binning, pre["rgb"], bg, fwd, dL, cam.W, cam.H, tiles=tiles)
    t2 = time.perf_counter()
    O.preprocess_backward(gs["means3D"], cam.viewmatrix, cam.projmatrix, cam.campos, cam.W, cam.H, cam.tanfovx,
                          cam.tanfovy, pre, rb, gs["scales"], gs["rotations"], 1.0, None, gs["shs"], args.sh_degree)
    t3 = time.perf_counter()
    est = (t1 - t0) + (t2 - t1) * (T / len(tiles)) + (t3 - t2)
    return {"value": 1.0 / est, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"1 keyframe of the same scene: all {args.P} Gaussians through preprocess/sort/backward-preprocess, "
                      f"blend fwd+bwd on {len(tiles)} of {T} tiles (every {tile_stride}th) scaled x{T / len(tiles):.1f}; "
                      f"{t3 - t0:.1f} s of CPU work, os.cpu_count()={os.cpu_count()}",
            "est_seconds_per_frame": est}


def measure_m2(rasterize, settings_cls, params, dL, device, args, fused=None, iters=6):
    """Secondary metric M2: one full SLAM render — the reference renderer's call pattern
    restated in tests/slam_glue.py (python-side pose transform, RGB pass + depth/silhouette pass sharing one
    means2D leaf) — plus backward of a scalar loss, frames/s.  `fused`: GaussianRasterizer class to also time
    the single-call extra_colors path."""
    from tests import slam_glue
    P, W, H = args.P, args.W, args.H
    bg = torch.zeros(3, device=device)
    rs = slam_glue.settings(settings_cls, W, H, bg, args.sh_degree, device)
    w2c = S.look_at_w2c((0.3, -0.1, 0.2), (0.0, 0.0, 4.0)).to(device)
    p = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}

    def two_pass():
        pose = w2c.clone().requires_grad_(True)
        rgb, depth, _, _ = slam_glue.render_two_pass(rasterize, rs, p, pose)
        ((rgb * dL).sum() + (depth * dL).sum()).backward()

    def one_pass():
        pose = w2c.clone().requires_grad_(True)
        mc = slam_glue.camera_frame(p, pose)
        m2 = torch.zeros_like(mc, requires_grad=True)
        rgb, depth, _ = fused(rs)(means3D=mc, means2D=m2, opacities=p["opacities"], shs=p["shs"], scales=p["scales"],
                                  rotations=p["rotations"], extra_colors=slam_glue.depth_silhouette(mc))
        ((rgb * dL).sum() + (depth * dL).sum()).backward()

    def native():   # pose as view/proj matrices (library camera gradients), depth colours generated in the library
        import diff_gaussian_rasterization as dgr_
        pose = w2c.clone().requires_grad_(True)
        view = pose.t()
        proj = view @ S.projection_matrix(*S.intrinsics(W, H), W, H).t().to(device)
        rs2 = settings_cls(image_height=H, image_width=W, tanfovx=rs.tanfovx, tanfovy=rs.tanfovy, bg=bg,
                           scale_modifier=1.0, viewmatrix=view, projmatrix=proj, sh_degree=args.sh_degree,
                           campos=torch.linalg.inv(view)[3, :3], prefiltered=False, debug=False)
        m2 = torch.zeros(P, 3, device=device, requires_grad=True)
        rgb, depth, _ = fused(rs2)(means3D=p["means3D"], means2D=m2, opacities=p["opacities"], shs=p["shs"],
                                   scales=p["scales"], rotations=p["rotations"], extra_colors=dgr_.DEPTH_SILHOUETTE)
        ((rgb * dL).sum() + (depth * dL).sum()).backward()

    out = {}
    for name, fn in (("two_pass", two_pass),) + ((("fused_rgbd", one_pass), ("fused_native_pose", native))
                                                 if fused is not None else ()):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out[name + "_fps"] = iters / (e0.elapsed_time(e1) / 1e3)
    out["what"] = "full SLAM render (pose transform + RGB + depth/silhouette) fwd+bwd, 1 GPU, same scene"
    return out


def measure_m1(dgr, params, rs, dL, device, iters=50, warm=10):
    """M1 as SURVEY.md §8(d) states it: ONE rasterizer call forward + backward through the stock API — one stream, no
    prepare_forward, no grad_targets — median of `iters` CUDA-event-timed iterations after `warm` warm-ups.  `synced`
    adds a host synchronisation after every backward, which is what a tracker iteration that reads its loss does
    (R/slam/tracker.py:99-167)."""
    P = params["means3D"].shape[0]
    p = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}

    def call():
        m2 = torch.zeros(P, 3, device=device, requires_grad=True)
        color, _ = dgr.GaussianRasterizer(rs)(means3D=p["means3D"], means2D=m2, opacities=p["opacities"], shs=p["shs"],
                                              scales=p["scales"], rotations=p["rotations"])
        color.backward(dL)
        for t in p.values():
            t.grad = None

    out = {}
    for name, sync in (("fps", False), ("synced_fps", True)):
        for _ in range(warm):
            call()
        torch.cuda.synchronize()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
        evs[0].record()
        for i in range(iters):
            call()
            evs[i + 1].record()
            if sync:
                evs[i + 1].synchronize()
        torch.cuda.synchronize()
        per = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(iters))
        out["m1_single_call_" + name] = 1e3 / per[len(per) // 2]
        out["ms_" + name.replace("_fps", "").replace("fps", "median")] = per[len(per) // 2]
    out["what"] = ("one GaussianRasterizer forward + backward of keyframe 0, stock API, one stream, median of "
                   f"{iters} event-timed calls; synced = host waits for every call like a tracker iteration")
    return out


def measure_blend(dgr, params, rs, stages, device):
    """(pixel, Gaussian) pairs/s of the two blend kernels and the contributing fraction, keyframe 0.
    Counts come from gsr_blend_stats (a plain per-pixel replay of the finished forward)."""
    lib = dgr._lib
    lib.gsr_blend_stats.restype = ctypes.c_int
    lib.gsr_blend_stats.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int64,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    P = params["means3D"].shape[0]
    with torch.no_grad():
        R, _, _, geom, binning, img = dgr._forward_native(params["means3D"], params["shs"], None, params["opacities"],
                                                          params["scales"], params["rotations"], None, rs, rs.viewmatrix,
                                                          rs.projmatrix, rs.campos, rs.bg)
    out = torch.zeros(4, dtype=torch.int64, device=device)
    rc = lib.gsr_blend_stats(torch.cuda.current_stream().cuda_stream, P, int(rs.image_width), int(rs.image_height),
                             int(R.cap), geom.data_ptr(), binning.data_ptr(), img.data_ptr(), out.data_ptr())
    if rc != 0:
        raise RuntimeError(lib.gsr_last_error().decode())
    upper, walked, blended, bwd_pairs = (int(x) for x in out.tolist())
    d = {"pairs_upper_bound": upper, "pairs_walk_to_last_contributor": walked, "pairs_contributing": blended,
         "pairs_evaluated_bwd": bwd_pairs, "contributing_fraction_of_upper_bound": blended / max(upper, 1),
         "contributing_fraction_of_evaluated_bwd": blended / max(bwd_pairs, 1)}
    if "render_fwd" in stages:
        d["fwd_contributing_gpairs_per_s"] = blended / (stages["render_fwd"]["ms"] * 1e-3) / 1e9
        d["fwd_upper_bound_gpairs_per_s"] = upper / (stages["render_fwd"]["ms"] * 1e-3) / 1e9
    if "render_bwd" in stages:
        d["bwd_contributing_gpairs_per_s"] = blended / (stages["render_bwd"]["ms"] * 1e-3) / 1e9
        d["bwd_evaluated_gpairs_per_s"] = bwd_pairs / (stages["render_bwd"]["ms"] * 1e-3) / 1e9
    d["what"] = ("keyframe 0: upper bound = sum over tiles of 256 x list length (SURVEY 8d); contributing = (pixel, splat) "
                 "pairs actually blended; evaluated_bwd = 32 x (warp, entry) pairs the backward visits")
    return d


def parity_check(dgr, params, rs, dL, device):
    """Once, outside every timed region: image and gradients of keyframe 0 against the unmodified compiled reference
    (oracle/_ref) on the same inputs, max-norm relative error; the bench refuses to report a number otherwise."""
    from oracle import ref_api
    if not ref_api.available():
        return {"checked": False, "why": "oracle/_ref/ref_dgr_C.so not present"}
    P = params["means3D"].shape[0]
    res = []
    for fn in ("ours", "ref"):
        p = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
        m2 = torch.zeros(P, 3, device=device, requires_grad=True)
        if fn == "ours":
            color, _ = dgr.GaussianRasterizer(rs)(means3D=p["means3D"], means2D=m2, opacities=p["opacities"], shs=p["shs"],
                                                  scales=p["scales"], rotations=p["rotations"])
        else:
            color, _ = ref_api.rasterize(p["means3D"], m2, p["opacities"], rs, shs=p["shs"], scales=p["scales"],
                                         rotations=p["rotations"])
        color.backward(dL)
        res.append((color.detach(), {k: v.grad for k, v in p.items()}))
    torch.cuda.synchronize()

    def rel(a, b):
        return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))
    errs = {"color": rel(res[0][0], res[1][0])}
    errs.update({"d_" + k: rel(res[0][1][k], res[1][1][k]) for k in res[0][1]})
    ok = all(e < 1e-4 for e in errs.values())
    return {"checked": True, "ok": ok, "tolerance": 1e-4, "max_rel_err": {k: float(f"{v:.3g}") for k, v in errs.items()},
            "against": "oracle/_ref (unmodified reference compiled for sm_100a), keyframe 0 of the timed workload"}


def ssim_window(device):
    g1 = torch.tensor([math.exp(-((x - 5) ** 2) / (2 * 1.5 ** 2)) for x in range(11)], device=device)
    g1 = (g1 / g1.sum()).unsqueeze(1)
    return g1.mm(g1.t()).unsqueeze(0).unsqueeze(0).expand(3, 1, 11, 11).contiguous()


def torch_mapping_loss(img, dep, gt, gt_depth, window):
    """BASELINE leg only: the torch composition the reference's mapper runs per iteration, restated inline —
    l1_loss + ssim (R/utils/loss_utils.py:64-68,114-154) and the masked depth L1 of R/slam/mapper.py:839-860."""
    import torch.nn.functional as F
    mu1, mu2 = F.conv2d(img, window, padding=5, groups=3), F.conv2d(gt, window, padding=5, groups=3)
    mu1_sq, mu2_sq, mu12 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = F.conv2d(img * img, window, padding=5, groups=3) - mu1_sq
    s2 = F.conv2d(gt * gt, window, padding=5, groups=3) - mu2_sq
    s12 = F.conv2d(img * gt, window, padding=5, groups=3) - mu12
    ssim = (((2 * mu12 + 1e-4) * (2 * s12 + 9e-4)) / ((mu1_sq + mu2_sq + 1e-4) * (s1 + s2 + 9e-4))).mean()
    im = 0.8 * torch.abs(img - gt).mean() + 0.2 * (1.0 - ssim)
    depth = dep[0]
    unc = (dep[2] - depth ** 2).detach()
    mask = ((gt_depth > 0) & (~torch.isnan(depth)) & (~torch.isnan(unc))).detach()
    return torch.abs(gt_depth - depth)[mask].mean() + 0.5 * im


LRS = {"means3D": 1.6e-4, "shs": 2.5e-3, "opacities": 5e-2, "scales": 1e-3, "rotations": 1e-3}   # R/configs/TUM.yml mapping


def measure_m3(impl, rasterize, settings_cls, params, device, args, iters=10):
    """M3-style figure, one mapping iteration as the reference runs it (R/slam/mapper.py:797-939:
    one keyframe per Adam step): full render (RGB + depth/silhouette) -> L1 + SSIM + masked depth L1 -> backward ->
    Adam step -> zero_grad, iterations/s on one GPU.
      impl "reference": the reference's call pattern — python pose transform, two rasterizer calls (tests/slam_glue.py),
                        the torch loss composition, torch.optim.Adam(lr=0.0, eps=1e-15) over the parameter groups;
      impl "b200":      one fused RGB+depth rasterizer call (library camera path, generated depth colours), gsr_slam_loss,
                        backward straight into the flat gradient bucket, gsr_adam_step with in-pass zero_grad."""
    from tests import slam_glue
    P, W, H = args.P, args.W, args.H
    bg = torch.zeros(3, device=device)
    w2c = S.look_at_w2c((0.3, -0.1, 0.2), (0.0, 0.0, 4.0)).to(device)
    g = torch.Generator().manual_seed(5)
    gt_color = torch.rand(3, H, W, generator=g).to(device)
    gt_depth = (torch.rand(H, W, generator=g) * 4 + 0.5).to(device)
    names = [k for k in LRS if k in params]

    if impl == "reference":
        rs = slam_glue.settings(settings_cls, W, H, bg, args.sh_degree, device)
        p = {k: params[k].detach().clone().requires_grad_(True) for k in names}
        opt = torch.optim.Adam([{"params": [p[k]], "lr": LRS[k], "name": k} for k in names], lr=0.0, eps=1e-15)
        window = ssim_window(device)

        def iteration():
            rgb, depth, _, _ = slam_glue.render_two_pass(rasterize, rs, p, w2c)
            torch_mapping_loss(rgb, depth, gt_color, gt_depth, window).backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
    else:
        import diff_gaussian_rasterization as dgr_
        import gsr_slam_ops as ops
        from gsr_mapstep import GradBucket
        opt = ops.FlatAdam({k: params[k].detach() for k in names}, LRS, eps=1e-15)
        p = {k: v.requires_grad_(True) for k, v in opt.views.items()}
        bucket = GradBucket(p)
        view = w2c.t().contiguous()
        proj = (view @ S.projection_matrix(*S.intrinsics(W, H), W, H).t().to(device)).contiguous()
        cam = S.make_camera(W, H)
        rs = settings_cls(image_height=H, image_width=W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=bg, scale_modifier=1.0,
                          viewmatrix=view, projmatrix=proj, sh_degree=args.sh_degree,
                          campos=torch.linalg.inv(view)[3, :3].contiguous(), prefiltered=False, debug=False)
        cfg = ops.mapper_splatam()
        m2 = torch.zeros(P, 3, device=device, requires_grad=True)

        def iteration():
            rgb, depth, _ = rasterize(rs)(means3D=p["means3D"], means2D=m2, opacities=p["opacities"], shs=p["shs"],
                                          scales=p["scales"], rotations=p["rotations"], extra_colors=dgr_.DEPTH_SILHOUETTE,
                                          grad_targets=bucket.views)
            _, g_img, g_dep = ops.slam_loss_and_grads(cfg, rgb, depth, gt_color, gt_depth, gt_depth)
            torch.autograd.backward([rgb, depth], [g_img, g_dep])
            opt.step(bucket.flat, zero_grads=True)
            m2.grad = None

    for _ in range(3):
        iteration()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        iteration()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return {"iterations_per_s": 1e3 / ms, "ms_per_iteration": ms,
            "what": "one mapping iteration (render RGB+depth, L1+SSIM+depth loss, backward, Adam step, zero_grad), 1 GPU, "
                    "1 keyframe per step as in the reference mapper"}


def measure_iteration_ops(device, args, iters=20):
    """the image loss (value + gradient) and the optimizer step of one mapping iteration,
    library call vs the reference's torch composition (R/utils/loss_utils.py l1_loss + ssim + the masked depth L1 of
    R/slam/mapper.py:839-860, restated inline; torch.optim.Adam as R/slam/gaussian_model.py:189) on the same GPU."""
    import torch.nn.functional as F

    import gsr_slam_ops as ops
    W, H, P = args.W, args.H, args.P
    d = {k: v.to(device) for k, v in S.make_loss_inputs(W, H, 3).items()}
    cfg = ops.mapper_splatam()

    def fused_loss():
        ops.slam_loss_and_grads(cfg, d["image"], d["depth_image"], d["gt_color"], d["gt_depth"], d["gt_depth"])

    window = ssim_window(device)

    def torch_loss():
        img = d["image"].clone().requires_grad_(True)
        dep = d["depth_image"].clone().requires_grad_(True)
        torch_mapping_loss(img, dep, d["gt_color"], d["gt_depth"], window).backward()

    n = 14 * P
    flat = {"p": torch.randn(n, device=device)}
    grads = torch.randn(n, device=device)
    fa = ops.FlatAdam(flat, {"p": 1e-4})
    tp = torch.randn(n, device=device, requires_grad=True)
    tp.grad = grads.clone()
    topt = torch.optim.Adam([tp], lr=1e-4, eps=1e-15)

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e3

    out = {"loss_us": timed(fused_loss), "loss_torch_us": timed(torch_loss),
           "adam_us": timed(lambda: fa.step(grads)), "adam_torch_us": timed(topt.step)}
    out["adam_gbs"] = 28.0 * n / (out["adam_us"] * 1e-6) / 1e9      # read p, g, m, v; write p, m, v
    out["what"] = (f"mapping-iteration loss (L1 + SSIM + masked depth L1, value and gradient, {W}x{H}) and Adam step over the "
                   f"{n}-float bucket: library call vs the torch composition the reference uses, microseconds per call")
    return out


# ------------------------------------------------------------------------------------------------
_T0 = time.perf_counter()


def phase(msg):
    """Progress marks on stderr (wall clock since start): where a run's time goes outside the timed region."""
    if int(os.environ.get("RANK", "0")) == 0:
        print(f"[bench +{time.perf_counter() - _T0:7.1f}s] {msg}", file=sys.stderr, flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    distributed = world > 1
    have_cuda = torch.cuda.is_available()

    if args.impl == "reference":
        from oracle import ref_api
        if rank != 0:
            return 0       # the reference has no multi-GPU path: rank 0 alone runs it
        if not (have_cuda and ref_api.available()):
            return reference_cpu_port(args)
    if not have_cuda:
        print(json.dumps({"error": "no CUDA device: the B200 path has no CPU fallback"}))
        return 1

    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if distributed and args.impl == "b200":
        # keep stdout for the one JSON line: NCCL's own banner / debug output goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=device)

    phase("building the synthetic scene")
    gs_cpu, cams, dL_cpu = build_workload(args, device)
    phase("scene built")
    bg = torch.zeros(3, device=device)
    dL = dL_cpu.to(device)
    names = ["means3D", "shs", "opacities", "scales", "rotations"]   # optimizer-group order of the reference
    params = {k: gs_cpu[k].to(device).requires_grad_(True) for k in names}
    P, N = args.P, args.W * args.H

    if args.impl == "b200":
        import diff_gaussian_rasterization as dgr
        from gsr_mapstep import ShardedMapStep
        kfs = settings_list(dgr.GaussianRasterizationSettings, cams, bg, args.sh_degree, device)

        def prepare_fn(p, rs):   # phase 1 of every keyframe is enqueued before the first host hand-off
            return dgr.prepare_forward(p["means3D"], p["opacities"], rs, shs=p["shs"], scales=p["scales"],
                                       rotations=p["rotations"])

        def forward_fn(p, rs, targets=None, handle=None):
            m2 = torch.zeros(P, 3, device=device, requires_grad=True)
            color, _ = dgr.GaussianRasterizer(rs)(means3D=p["means3D"], means2D=m2, opacities=p["opacities"],
                                                  shs=p["shs"], scales=p["scales"], rotations=p["rotations"],
                                                  grad_targets=targets, prepared=handle)
            return color, dL          # back-propagated with the fixed dL/dpix by ShardedMapStep.step

        # the rasterizer inputs ARE the optimised tensors here, so the backward kernels add straight into the
        # flat gradient bucket instead of returning tensors for autograd to accumulate;
        # consecutive keyframes alternate between two CUDA streams (binning of one overlaps blending of another)
        stepper = ShardedMapStep(params, forward_fn=forward_fn, streams=int(os.environ.get("GSR_BENCH_STREAMS", "4")),
                                 direct_targets=not os.environ.get("GSR_BENCH_NO_TARGETS"),
                                 # (overwrite_first measured slower here: its per-frame backward calls make autograd
                                 #  re-join the streams after every frame: 1705 vs 1861 frames/s)
                                 overwrite_first=bool(os.environ.get("GSR_BENCH_OVERWRITE")),
                                 prepare_fn=None if os.environ.get("GSR_BENCH_NO_PREPARE") else prepare_fn,
                                 exchange=os.environ.get("GSR_EXCHANGE", "auto"))
        step = lambda: stepper.step(kfs)  # noqa: E731
        launch_count = dgr._lib.gsr_launch_count
        launch_count.restype = ctypes.c_longlong
    else:
        from collections import namedtuple
        RS = namedtuple("RS", "image_height image_width tanfovx tanfovy bg scale_modifier viewmatrix projmatrix "
                              "sh_degree campos prefiltered debug")
        kfs = settings_list(RS, cams, bg, args.sh_degree, device)

        def step():
            for p in params.values():
                p.grad = None
            out = []
            for rs in kfs:
                m2 = torch.zeros(P, 3, device=device, requires_grad=True)
                color, _ = ref_api.rasterize(params["means3D"], m2, params["opacities"], rs, shs=params["shs"],
                                             scales=params["scales"], rotations=params["rotations"])
                color.backward(dL)
                out.append(color.detach())
            return out
        launch_count = None

    phase("instance counts")
    # instance counts for the byte accounting (one untimed forward per keyframe)
    R_list, vis_list = [], []
    if args.impl == "b200":
        with torch.no_grad():
            for rs in kfs:
                R, _, radii, _, _, _ = dgr._forward_native(params["means3D"], params["shs"], None, params["opacities"],
                                                           params["scales"], params["rotations"], None, rs,
                                                           rs.viewmatrix, rs.projmatrix, rs.campos, rs.bg)
                R_list.append(R)
                vis_list.append(int((radii > 0).sum()))

    def barrier():
        if distributed and args.impl == "b200":
            dist.barrier()
        torch.cuda.synchronize()

    # nvidia-smi needs ~100 ms to start and the timed region can be a few tens of ms: start sampling before the
    # warm-up (same load) so that the record covers warm-up + timed region
    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else local_rank)
    if rank == 0:
        sampler.start()
    phase("warm-up")
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    phase("timed region")
    n0 = launch_count() if launch_count else 0
    # `repeats` timed blocks of EXACTLY --steps steps, each bracketed by barrier + synchronize on both sides and timed
    # with CUDA events on the launching stream; a block's time is the max over ranks, the line reports the median block.
    block_ms = []
    for _ in range(max(1, args.repeats)):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            step()
        ev1.record()
        barrier()
        block_ms.append(ev0.elapsed_time(ev1))
    n1 = launch_count() if launch_count else 0
    if distributed and args.impl == "b200":
        t = torch.tensor(block_ms, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        block_ms = [float(x) for x in t.tolist()]
    ms = statistics.median(block_ms)
    if sum(block_ms) < 400.0:   # short timed region: keep the same load going (untimed, same count on every rank) until
        for _ in range(min(400, int(400.0 / max(ms / args.steps, 1e-3)) + 1)):   # the 100 ms sampler has a few samples
            step()
        barrier()
    clocks = sampler.stop() if rank == 0 else None
    phase("timed region done")
    frames = args.keyframes * args.steps
    value = frames / (ms / 1e3)

    workloads = {"cmain": "C-main: 1M-Gaussian synthetic scene, 640x480, SH degree 0, 8 orbit keyframes/step",
                 "c5": "C5 (BASELINE config 5): 5M-Gaussian synthetic scene, 1920x1080, SH degree 0, 32 orbit keyframes/step",
                 "custom": "custom"}
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # identical in both arms (the driver compares it): only what defines the workload
        "config": {"workload": workloads[args.config], "P": args.P, "W": args.W, "H": args.H,
                   "keyframes_per_step": args.keyframes, "sh_degree": args.sh_degree,
                   "l2": "per-step working set (56 B/G params + 116 B/G grads + per-frame 48 B/G records and "
                         "24+ B/instance binning, > 400 MB) exceeds the 126 MB L2; no explicit flush"},
        "run": {"parallelism": f"keyframe-sharded dp{world}" if args.impl == "b200" else "1 gpu (the reference has no multi-GPU path)",
                "streams_per_rank": int(os.environ.get("GSR_BENCH_STREAMS", "4")) if args.impl == "b200" else 1,
                "repeats": len(block_ms), "block_ms": [round(x, 3) for x in block_ms], "statistic": "median block"},
        "clocks": clocks,
    }
    if args.impl == "reference":
        line["impl"] = "reference"
        line["cpu_baseline"] = {"value": value, "unit": "frames/s", "cores": 1, "kind": "reference",
                                "sample": "the reference's only implementation is CUDA: compiled unmodified from "
                                          "/root/reference for sm_100a (oracle/_ref) and run on this GPU, full workload"}
        line["e2e"] = {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        try:
            line["m2"] = measure_m2(lambda m3, m2, op, rs_, **kw: ref_api.rasterize(m3, m2, op, rs_, **kw), RS, params,
                                    dL, device, args)
        except Exception as ex:
            line["m2"] = {"error": repr(ex)}
        try:
            line["m3"] = measure_m3("reference", lambda m3, m2, op, rs_, **kw: ref_api.rasterize(m3, m2, op, rs_, **kw), RS,
                                    params, device, args)
        except Exception as ex:
            line["m3"] = {"error": repr(ex)}
        print(json.dumps(line), flush=True)
        return 0

    line["gpu_launches"] = int(n1 - n0) // max(1, len(block_ms))      # per timed block of --steps steps
    line["run"]["exchange"] = stepper.exchange if world > 1 else "none (1 rank)"
    if stepper.exchange_error:
        line["run"]["exchange_fallback_reason"] = stepper.exchange_error
    line["run"]["R_mean"] = sum(R_list) / max(len(R_list), 1)
    line["run"]["visible_mean"] = sum(vis_list) / max(len(vis_list), 1)

    # ---- where a step's time goes on a rank: kernels vs the gradient exchange (CUDA events inside the steps, which run
    # back to back exactly like the timed region; read after the last one) ----
    stepper.timing = True
    for _ in range(12):
        step()
    tm = stepper.timings()[2:]
    stepper.timing = False
    barrier()
    t = torch.tensor([statistics.median(x["kernel_ms"] for x in tm), statistics.median(x["comm_ms"] for x in tm)], device=device)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    line["kernel_ms"], line["comm_ms"] = round(float(t[0]), 4), round(float(t[1]), 4)
    line["run"]["timing_note"] = ("kernel_ms: step start -> local gradient bucket complete; comm_ms: the exchange after it "
                                  "(includes waiting for the slowest rank); steps back to back as in the timed region, "
                                  "medians of 10 steps, max over ranks")

    phase("stage profile")
    # ---- per-stage profile (untimed pass with the library's stage events on) --------------------
    lib = dgr._lib
    nst = lib.gsr_profile_num_stages()
    lib.gsr_profile_stage_name.restype = ctypes.c_char_p
    lib.gsr_profile_enable(1)
    torch.cuda.synchronize()
    ms_arr, cnt_arr = (ctypes.c_float * nst)(), (ctypes.c_int * nst)()
    lib.gsr_profile_collect(ms_arr, cnt_arr)
    saved_streams, stepper.nstreams = stepper.nstreams, 1   # stage events need one in-order stream
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    stepper.nstreams = saved_streams
    lib.gsr_profile_collect(ms_arr, cnt_arr)
    lib.gsr_profile_enable(0)
    my_kfs = len(range(rank, args.keyframes, world))
    R_mine = [R_list[i] for i in range(rank, args.keyframes, world)]
    Rm = sum(R_mine) / max(len(R_mine), 1)
    sb = stage_bytes(P, Rm, N, (args.sh_degree + 1) ** 2)
    peak, peak_src = measured_peak()
    stages, avg_ms = {}, {}
    for i in range(nst):
        if cnt_arr[i]:
            avg_ms[lib.gsr_profile_stage_name(i).decode()] = ms_arr[i] / cnt_arr[i]
    if "tile_totals" in avg_ms:   # the per-tile count kernel belongs to the tile partition (it runs ahead of the sort)
        avg_ms["tile_partition"] = avg_ms.get("tile_partition", 0.0) + avg_ms.pop("tile_totals")
    for nm, avg in avg_ms.items():
        stages[nm] = {"ms": avg, "alg_bytes": sb[nm], "gbs": sb[nm] / (avg * 1e-3) / 1e9}
    total_ms = sum(s["ms"] for s in stages.values())
    dom = max(stages, key=lambda k: stages[k]["ms"])
    line["stages"] = {k: {"ms": round(v["ms"], 4), "share": round(v["ms"] / total_ms, 3), "gbs": round(v["gbs"], 1)}
                      for k, v in stages.items()}
    traffic = None
    try:   # DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get(dom)
    except Exception:
        pass
    line["roofline"] = {"bound": "hbm", "kernel": dom, "achieved": stages[dom]["gbs"], "peak": peak, "unit": "GB/s",
                        "frac": stages[dom]["gbs"] / peak, "traffic": traffic, "peak_source": peak_src,
                        "alg_bytes_per_launch": stages[dom]["alg_bytes"], "ms_per_launch": stages[dom]["ms"],
                        "call_alg_bytes": sum(sb.values()), "call_ms_sum_of_stages": total_ms,
                        "call_frac": sum(sb.values()) / (total_ms * 1e-3) / 1e9 / peak}

    if rank == 0 and not args.no_extras:
        try:
            line["blend"] = measure_blend(dgr, params, kfs[0], stages, device)
        except Exception as ex:
            line["blend"] = {"error": repr(ex)}
        try:
            line["parity_check"] = parity_check(dgr, params, kfs[0], dL, device)
        except Exception as ex:   # the checker itself failed (e.g. the compiled reference could not be loaded): say so
            line["parity_check"] = {"checked": False, "why": repr(ex)[:300]}
    parity_failed = bool(rank == 0 and line.get("parity_check", {}).get("checked") and not line["parity_check"]["ok"])
    if distributed:   # every rank learns the verdict (this is also the barrier behind rank 0's extra work)
        flag = torch.tensor([1 if parity_failed else 0], device=device)
        dist.broadcast(flag, 0)
        parity_failed = bool(int(flag.item()))
    if parity_failed:
        if rank == 0:
            print(json.dumps({"error": "parity check against oracle/_ref failed", "parity_check": line["parity_check"]}),
                  flush=True)
        return 1

    phase("end-to-end leg")
    # ---- end to end through the public API with HOST buffers ------------------------------------
    if not args.no_e2e:
        host_in = {k: gs_cpu[k].contiguous().pin_memory() for k in names}
        host_grad = torch.empty(stepper.bucket.flat.numel(), dtype=torch.float32).pin_memory()
        host_img = torch.empty(my_kfs, 3, args.H, args.W).pin_memory()

        # Pipelined like a real consumer would: every step still uploads its inputs from pinned host memory and
        # downloads its results (gradient bucket + images), but on a copy stream with double-buffered device
        # parameters / result staging, so the PCIe traffic of step i+1 / i-1 overlaps the kernels of step i.
        copy_stream = torch.cuda.Stream(device)       # uploads
        down_stream = torch.cuda.Stream(device)       # downloads (PCIe is full duplex: keep the directions apart)
        main_stream = torch.cuda.current_stream(device)
        # one flat pinned host block / one flat device block per buffer: a single copy per direction per step.
        # With N > 1 ranks the host copies are SHARDED: rank r uploads only its 1/N slice of the parameters and the
        # library's all-gather kernel replicates it over NVLink (include/gscomm_b200.h); after the exchange every rank
        # holds the reduced bucket and rank r downloads only its 1/N slice — host traffic does not grow with N.
        sizes = [params[k].numel() for k in names]
        total = sum(sizes)
        host_flat = torch.cat([host_in[k].reshape(-1) for k in names]).pin_memory()
        shard = distributed and stepper.peer is not None
        if shard:
            from gsr_mapstep import PeerExchange
            pex = [PeerExchange(total, device, None, "auto") for _ in range(2)]
            pflat = [p.flat for p in pex]
            ulo, uhi = pex[0].slice_range()
            glo, ghi = stepper.peer.slice_range()
        else:
            pex = None
            pflat = [torch.empty(total, dtype=torch.float32, device=device) for _ in range(2)]
            ulo, uhi, glo, ghi = 0, total, 0, total
        h2d = (uhi - ulo) * 4
        d2h = (ghi - glo) * 4 + host_img.numel() * 4
        pbuf = []
        for fl in pflat:
            d, off = {}, 0
            for k, n_ in zip(names, sizes):
                d[k] = fl[off: off + n_].view_as(params[k]).requires_grad_(True)
                off += n_
            pbuf.append(d)
        up_done = [torch.cuda.Event() for _ in range(2)]
        free_in = [torch.cuda.Event() for _ in range(2)]      # step that read pbuf[b] has finished
        stage_grad = [torch.empty(ghi - glo, dtype=torch.float32, device=device) for _ in range(2)]
        stage_img = [torch.empty(max(my_kfs, 1), 3, args.H, args.W, device=device) for _ in range(2)]
        free_out = [torch.cuda.Event() for _ in range(2)]     # download of stage[b] has finished
        state = {"i": 0}

        def upload(b):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(free_in[b])
                pflat[b][ulo:uhi].copy_(host_flat[ulo:uhi], non_blocking=True)
                if shard:
                    pex[b].all_gather()
                up_done[b].record(copy_stream)

        for b in range(2):
            free_in[b].record(main_stream)
            free_out[b].record(main_stream)
        upload(0)

        def e2e_step():
            i = state["i"]
            b = i & 1
            upload(b ^ 1)                                   # next step's inputs, overlapping this step's kernels
            main_stream.wait_event(up_done[b])
            imgs = stepper.step(kfs, params=pbuf[b]) if stepper.direct_targets else step()
            main_stream.wait_event(free_out[b])
            stage_grad[b].copy_(stepper.bucket.flat[glo:ghi], non_blocking=True)
            if imgs:
                torch.stack(imgs, out=stage_img[b][:len(imgs)])
            free_in[b].record(main_stream)
            done = torch.cuda.Event()
            done.record(main_stream)
            with torch.cuda.stream(down_stream):
                down_stream.wait_event(done)
                host_grad[glo:ghi].copy_(stage_grad[b], non_blocking=True)
                host_img[:len(imgs)].copy_(stage_img[b][:len(imgs)], non_blocking=True)
                free_out[b].record(down_stream)
            state["i"] = i + 1

        for _ in range(2):
            e2e_step()
        main_stream.wait_stream(copy_stream)
        main_stream.wait_stream(down_stream)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_e2e = max(5, args.steps)      # long enough to amortise the un-overlapped last download
        e0.record()
        for _ in range(n_e2e):
            e2e_step()
        main_stream.wait_stream(copy_stream)
        main_stream.wait_stream(down_stream)               # the last download is inside the timed region
        e1.record()
        barrier()
        ems = e0.elapsed_time(e1)
        if distributed:
            t = torch.tensor([ems], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        line["e2e"] = {"value": args.keyframes * n_e2e / (ems / 1e3), "unit": "frames/s", "h2d_bytes_per_step": h2d,
                       "d2h_bytes_per_step": d2h, "steps": n_e2e,
                       "sharded_host_copies": bool(shard),
                       "what": "every step: Gaussian parameters uploaded from pinned host memory, 8-keyframe step through "
                               "GaussianRasterizer, gradient bucket + rendered images downloaded to pinned host memory; "
                               "copies run on a copy stream with double-buffered device parameters / result staging so "
                               "they overlap the neighbouring steps' kernels.  With N > 1 every rank moves 1/N of the "
                               "parameters / of the reduced bucket over PCIe (bytes_per_step are per rank) and the "
                               "library's all-gather replicates the parameters over NVLink"}

    phase("secondary metrics")
    extras = rank == 0 and world == 1 and not args.no_extras
    if extras:
        try:
            line["m1"] = measure_m1(dgr, params, kfs[0], dL, device)
        except Exception as ex:
            line["m1"] = {"error": repr(ex)}
    if extras:
        try:
            line["m2"] = measure_m2(
                lambda m3, m2, op, rs_, **kw: dgr.GaussianRasterizer(rs_)(means3D=m3, means2D=m2, opacities=op, **kw),
                dgr.GaussianRasterizationSettings, params, dL, device, args, fused=dgr.GaussianRasterizer)
        except Exception as ex:
            line["m2"] = {"error": repr(ex)}
    if extras:
        try:
            line["m3"] = measure_m3("b200", dgr.GaussianRasterizer, dgr.GaussianRasterizationSettings, params, device, args)
        except Exception as ex:
            line["m3"] = {"error": repr(ex)}
        try:
            line["iteration_ops"] = measure_iteration_ops(device, args)
        except Exception as ex:
            line["iteration_ops"] = {"error": repr(ex)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline(args, gs_cpu, cams[0], dL_cpu)
        except Exception as ex:  # the baseline must never take the bench line down
            line["cpu_baseline"] = {"error": repr(ex)}
    phase("done")
    if rank == 0:
        print(json.dumps(line), flush=True)
    if distributed:
        try:
            dist.barrier()
            dist.destroy_process_group()
        except Exception:
            pass
    return 0


def reference_cpu_port(args):
    """Fallback reference arm when the compiled reference is unavailable: the CPU oracle port."""
    gs, cams, dL = build_workload(args, None)
    t = time.perf_counter()
    cb = cpu_baseline(args, gs, cams[0], dL)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "frames/s", "n_gpus": args.gpus,
            "steps": 1, "warmup": 0, "ms_per_step": cb["est_seconds_per_frame"] * 1e3 * args.keyframes,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C-main (CPU oracle port, bounded sample)", "P": args.P, "W": args.W, "H": args.H},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "frames/s", "h2d_bytes_per_step": 0,
                                        "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t}
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
