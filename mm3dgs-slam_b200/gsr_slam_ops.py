"""Host side of include/gsloss_b200.h: the image losses and the optimizer step that sit either side of the
rasterizer in one MM3DGS-SLAM optimisation iteration.

    slam_loss(cfg, image, depth_image, gt_color, depth_target, gt_depth) -> scalar (autograd)
    slam_loss_and_grads(...) -> (losses[4], dL/dimage, dL/ddepth_image)    no autograd graph, no host sync
    mapper_splatam / mapper_default / tracker_splatam / tracker_default    the reference's four compositions
        (R/slam/mapper.py:839-885, R/slam/tracker.py:110-144) as configurations
    FlatAdam       torch.optim.Adam(l, lr=0.0, eps=1e-15) of R/slam/gaussian_model.py:151-189 over one flat
                   parameter / gradient / moment bucket, one launch per step

Plumbing only (allocation + ctypes marshalling); all compute is in libgsrast_b200.so.  CUDA tensors only —
there is no CPU path.
"""
from __future__ import annotations

import ctypes
from typing import Dict

import torch

from diff_gaussian_rasterization import _check, _lib

COLOR_NONE, COLOR_L1_SSIM, COLOR_MASKED_L1_MEAN, COLOR_MASKED_L1_SUM = 0, 1, 2, 3
DEPTH_NONE, DEPTH_L1_MEAN, DEPTH_L1_SUM, DEPTH_PEARSON, DEPTH_PEARSON_INV, DEPTH_PEARSON_COLS = 0, 1, 2, 3, 4, 5
MASK_GT_DEPTH_POS, MASK_NOT_NAN, MASK_SILHOUETTE = 1, 2, 4
ADAM_MAX_SEGMENTS = 16


class _LossConfig(ctypes.Structure):   # struct gsr_loss_config
    _fields_ = [("width", ctypes.c_int32), ("height", ctypes.c_int32), ("color_mode", ctypes.c_int32),
                ("depth_mode", ctypes.c_int32), ("color_mask", ctypes.c_int32), ("depth_mask", ctypes.c_int32),
                ("lambda_dssim", ctypes.c_float), ("sil_threshold", ctypes.c_float), ("color_weight", ctypes.c_float),
                ("depth_weight", ctypes.c_float), ("grad_scale", ctypes.c_float), ("_pad", ctypes.c_int32)]


def _bind():
    vp, i32, i64, sz, f32, f64 = (ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_size_t, ctypes.c_float,
                                 ctypes.c_double)
    _lib.gsr_slam_loss_ws_bytes.restype = sz
    _lib.gsr_slam_loss_ws_bytes.argtypes = [i32, i32]
    _lib.gsr_slam_loss.restype = ctypes.c_int
    _lib.gsr_slam_loss.argtypes = [vp, ctypes.POINTER(_LossConfig), vp, vp, vp, vp, vp, vp, sz, vp, vp, vp]
    _lib.gsr_adam_step.restype = ctypes.c_int
    _lib.gsr_adam_step.argtypes = [vp, vp, vp, vp, vp, i64, i32, ctypes.POINTER(i64), ctypes.POINTER(f64), f64, f64, f64,
                                   i64, f32, i32]
    _lib.gsr_compact_ws_bytes.restype = sz
    _lib.gsr_compact_ws_bytes.argtypes = [i64]
    _lib.gsr_compact_scan.restype = ctypes.c_int
    _lib.gsr_compact_scan.argtypes = [vp, i64, vp, vp, vp, vp, sz]
    _lib.gsr_compact_gather.restype = ctypes.c_int
    _lib.gsr_compact_gather.argtypes = [vp, i64, i64, i32, ctypes.POINTER(i32), vp, vp, i32, ctypes.POINTER(vp),
                                        ctypes.POINTER(vp)]


_bind()


# ---- the reference's loss compositions -----------------------------------------------------------------------
def mapper_splatam(lambda_dssim=0.2):
    """losses["depth"] + 0.5 * losses["im"]   (R/slam/mapper.py:839-860)"""
    return dict(color_mode=COLOR_L1_SSIM, lambda_dssim=lambda_dssim, depth_mode=DEPTH_L1_MEAN,
                depth_mask=MASK_GT_DEPTH_POS | MASK_NOT_NAN, color_weight=0.5, depth_weight=1.0)


def mapper_default(lambda_dssim=0.2, pearson_weight=0.05, use_gt_depth=False):
    """(1-l) L1 + l (1 - SSIM) + w * pearson_loss(depth, est or gt, invert_estimate=False)   (R/slam/mapper.py:862-885).
    Without ground-truth depth the reference passes mask=None, so torchmetrics sees the 2-D [H, W] images and returns one
    coefficient per image column (averaged by the trailing .mean()): DEPTH_PEARSON_COLS.  With ground-truth depth the
    masked, flattened pixels give one global coefficient: DEPTH_PEARSON."""
    return dict(color_mode=COLOR_L1_SSIM, lambda_dssim=lambda_dssim,
                depth_mode=DEPTH_PEARSON if use_gt_depth else DEPTH_PEARSON_COLS,
                depth_mask=MASK_GT_DEPTH_POS if use_gt_depth else 0, color_weight=1.0, depth_weight=pearson_weight)


def tracker_splatam():
    """sum|gt_depth - depth|[mask] + 0.5 * sum|gt_color - image|[mask]   (R/slam/tracker.py:110-126)"""
    m = MASK_GT_DEPTH_POS | MASK_NOT_NAN | MASK_SILHOUETTE
    return dict(color_mode=COLOR_MASKED_L1_SUM, color_mask=m, depth_mode=DEPTH_L1_SUM, depth_mask=m,
                sil_threshold=0.99, color_weight=0.5, depth_weight=1.0)


def tracker_default(pearson_weight=0.05, use_gt_depth=False):
    """mean|image - gt|[:, sil] + w * pearson_loss(depth, est or gt, mask, invert_estimate=True)   (R/slam/tracker.py:127-144)"""
    return dict(color_mode=COLOR_MASKED_L1_MEAN, color_mask=MASK_SILHOUETTE, depth_mode=DEPTH_PEARSON_INV,
                depth_mask=MASK_SILHOUETTE | (MASK_GT_DEPTH_POS if use_gt_depth else 0), sil_threshold=0.99,
                color_weight=1.0, depth_weight=pearson_weight)


def _f32c(t, name, shape=None):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"gsr_slam_ops: {name} must be a CUDA tensor (there is no CPU path)")
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.to(torch.float32).contiguous()
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise RuntimeError(f"gsr_slam_ops: {name} has shape {tuple(t.shape)}, expected {tuple(shape)}")
    return t


def slam_loss_and_grads(cfg: dict, image, depth_image, gt_color, depth_target=None, gt_depth=None,
                        want_grad: bool = True, grad_scale: float = 1.0):
    """One gsr_slam_loss call on torch's current stream.  Returns (losses, dL_dimage, dL_ddepth_image):
    losses = device tensor [total, colour term, depth term, mean SSIM]; the gradient tensors are None for an
    image the configuration does not use (or when want_grad is False).  Nothing here waits for the GPU."""
    ref = image if image is not None else depth_image
    if ref is None:
        raise RuntimeError("gsr_slam_ops: image or depth_image required")
    H, W = int(ref.shape[-2]), int(ref.shape[-1])
    image = _f32c(image, "image", (3, H, W))
    depth_image = _f32c(depth_image, "depth_image", (3, H, W))
    gt_color = _f32c(gt_color, "gt_color", (3, H, W))
    depth_target = _f32c(depth_target, "depth_target", (H, W))
    gt_depth = _f32c(gt_depth, "gt_depth", (H, W))
    dev = ref.device
    c = _LossConfig(W, H, int(cfg.get("color_mode", 0)), int(cfg.get("depth_mode", 0)), int(cfg.get("color_mask", 0)),
                    int(cfg.get("depth_mask", 0)), float(cfg.get("lambda_dssim", 0.2)), float(cfg.get("sil_threshold", 0.5)),
                    float(cfg.get("color_weight", 1.0)), float(cfg.get("depth_weight", 1.0)), float(grad_scale), 0)
    with torch.cuda.device(dev):
        ws = torch.empty(_lib.gsr_slam_loss_ws_bytes(W, H), dtype=torch.uint8, device=dev)
        losses = torch.empty(4, dtype=torch.float32, device=dev)
        d_img = torch.empty_like(image) if (want_grad and c.color_mode != COLOR_NONE) else None
        d_dep = torch.empty_like(depth_image) if (want_grad and c.depth_mode != DEPTH_NONE) else None
        p = lambda t: None if t is None else t.data_ptr()   # noqa: E731
        _check(_lib.gsr_slam_loss(torch.cuda.current_stream(dev).cuda_stream, ctypes.byref(c), p(image), p(depth_image),
                                  p(gt_color), p(depth_target), p(gt_depth), ws.data_ptr(), ws.numel(), losses.data_ptr(),
                                  p(d_img), p(d_dep)))
    return losses, d_img, d_dep


class _SlamLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, depth_image, gt_color, depth_target, gt_depth, cfg):
        need = (image is not None and image.requires_grad) or (depth_image is not None and depth_image.requires_grad)
        losses, d_img, d_dep = slam_loss_and_grads(cfg, image, depth_image, gt_color, depth_target, gt_depth, want_grad=need)
        ctx.grads = (d_img, d_dep)
        ctx.mark_non_differentiable(losses)
        return losses[0].clone(), losses

    @staticmethod
    def backward(ctx, g_total, _g_losses):
        d_img, d_dep = ctx.grads
        return (None if d_img is None else d_img * g_total, None if d_dep is None else d_dep * g_total,
                None, None, None, None)


def slam_loss(cfg: dict, image, depth_image, gt_color, depth_target=None, gt_depth=None, return_terms: bool = False):
    """Differentiable scalar loss (w.r.t. image and depth_image).  return_terms: also the [total, colour, depth, ssim]
    device tensor."""
    total, losses = _SlamLoss.apply(image, depth_image, gt_color, depth_target, gt_depth, cfg)
    return (total, losses) if return_terms else total


# ---- Adam over a flat bucket ---------------------------------------------------------------------------------
class FlatAdam:
    """`params`: dict name -> tensor; the tensors are re-homed as views of ONE flat fp32 buffer (`self.flat`, same
    order), so that parameters, gradients (a GradBucket-style flat tensor with the same layout) and both Adam moments
    are four parallel arrays and a step is one kernel launch.  `prune` / `extend` are the reference's optimizer surgery
    (_prune_optimizer / cat_tensors_to_optimizer) on those arrays.  `lrs`: dict name -> learning rate (host floats; update
    `self.lrs[name]` between steps for a schedule, R/slam/gaussian_model.py:196-202)."""

    def __init__(self, params: Dict[str, torch.Tensor], lrs: Dict[str, float], betas=(0.9, 0.999), eps: float = 1e-15):
        if len(params) > ADAM_MAX_SEGMENTS:
            raise ValueError(f"at most {ADAM_MAX_SEGMENTS} parameter groups")
        self.names = list(params.keys())
        self.lrs = dict(lrs)
        self.beta1, self.beta2 = float(betas[0]), float(betas[1])
        self.eps = float(eps)
        self.steps = 0
        self._rehome({k: p.detach() for k, p in params.items()}, None, None)

    def _rehome(self, params, exp_avg, exp_avg_sq):
        """(Re)build the three flat buffers from per-group tensors (construction and map surgery; bookkeeping only —
        the optimizer arithmetic is gsr_adam_step's)."""
        first = next(iter(params.values()))
        n = sum(p.numel() for p in params.values())
        self.flat = torch.empty(n, dtype=torch.float32, device=first.device)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=first.device)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=first.device)
        self.views = {}
        ends, off = [], 0
        for k in self.names:
            p = params[k]
            sl = slice(off, off + p.numel())
            self.views[k] = self.flat[sl].view(p.shape)
            self.views[k].copy_(p)
            if exp_avg is not None:
                self.exp_avg[sl].view(p.shape).copy_(exp_avg[k])
                self.exp_avg_sq[sl].view(p.shape).copy_(exp_avg_sq[k])
            off += p.numel()
            ends.append(off)
        self.seg_end = (ctypes.c_int64 * len(ends))(*ends)
        return self.views

    def _group_views(self, flat):
        out, off = {}, 0
        for k in self.names:
            v = self.views[k]
            out[k] = flat[off: off + v.numel()].view(v.shape)
            off += v.numel()
        return out

    def _relayout(self, keep, rows_out):
        """One gsr_compact_gather pass: the kept rows of every group of parameters and both moments land in fresh flat
        buffers laid out for `rows_out` rows per group.  Returns the new (flat, exp_avg, exp_avg_sq), the row widths
        and the number of kept rows."""
        if not self.flat.is_cuda:
            raise RuntimeError("FlatAdam: CUDA tensors only (there is no CPU path)")
        dev = self.flat.device
        P = int(next(iter(self.views.values())).shape[0])
        widths = []
        for k in self.names:
            w = 1
            for d in self.views[k].shape[1:]:
                w *= int(d)
            widths.append(w)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            new_index, kept = None, P
            if keep is not None:
                if keep.shape != (P,) or keep.device != dev:
                    raise RuntimeError("FlatAdam.prune: keep must be a [P] mask on the parameters' device")
                k8 = keep.to(torch.uint8).contiguous() if keep.dtype != torch.uint8 else keep.contiguous()
                new_index = torch.empty(max(P, 1), dtype=torch.int32, device=dev)
                count = torch.empty(1, dtype=torch.int64, device=dev)
                ws = torch.empty(_lib.gsr_compact_ws_bytes(P), dtype=torch.uint8, device=dev)
                _check(_lib.gsr_compact_scan(stream, P, k8.data_ptr(), new_index.data_ptr(), count.data_ptr(),
                                             ws.data_ptr(), ws.numel()))
                kept = int(count.item())       # the one host hand-off: the new buffers are sized by it
                if rows_out is None:
                    rows_out = kept
            wtot = sum(widths)
            outs = [torch.empty(rows_out * wtot, dtype=torch.float32, device=dev) for _ in range(3)]
            srcs = [self.flat, self.exp_avg, self.exp_avg_sq]
            if P:
                w_arr = (ctypes.c_int32 * len(widths))(*widths)
                s_arr = (ctypes.c_void_p * 3)(*[t.data_ptr() for t in srcs])
                d_arr = (ctypes.c_void_p * 3)(*[t.data_ptr() for t in outs])
                _check(_lib.gsr_compact_gather(stream, P, rows_out, len(widths), w_arr,
                                               k8.data_ptr() if keep is not None else None,
                                               new_index.data_ptr() if keep is not None else None, 3, s_arr, d_arr))
        return outs, widths, kept

    def _adopt(self, outs, widths, rows):
        self.flat, self.exp_avg, self.exp_avg_sq = outs
        shapes = {k: (rows,) + tuple(self.views[k].shape[1:]) for k in self.names}
        self.views, ends, off = {}, [], 0
        for k, w in zip(self.names, widths):
            self.views[k] = self.flat[off: off + rows * w].view(shapes[k])
            off += rows * w
            ends.append(off)
        self.seg_end = (ctypes.c_int64 * len(ends))(*ends)
        return self.views

    def prune(self, keep: torch.Tensor):
        """Keep the rows (dim 0 of every group) selected by the bool mask `keep`, parameters and both moments alike
        (_prune_optimizer, R/slam/gaussian_model.py:380-399; the mapper calls it with ~prune_mask): one mask scan + ONE
        gather pass over the three flat buffers (gsr_compact_scan / gsr_compact_gather) instead of 21 boolean-index
        kernels.  The step count is kept.  Returns the new parameter views — the old ones are stale, as the
        reference's old Parameters are."""
        outs, widths, kept = self._relayout(keep, None)
        return self._adopt(outs, widths, kept)

    def extend(self, new: Dict[str, torch.Tensor]):
        """Append rows to every group; their moments start at zero and they share the running step count
        (cat_tensors_to_optimizer, R/slam/gaussian_model.py:418-451).  The existing rows move to their place in the
        longer layout in one gather pass; the appended rows are copied behind them.  Returns the new parameter views."""
        P = int(next(iter(self.views.values())).shape[0])
        A = int(next(iter(new.values())).shape[0])
        outs, widths, _ = self._relayout(None, P + A)
        views = self._adopt(outs, widths, P + A)
        m, v = self._group_views(self.exp_avg), self._group_views(self.exp_avg_sq)
        for k in self.names:
            views[k][P:].copy_(new[k].detach().to(self.flat).reshape(views[k][P:].shape))
            m[k][P:].zero_()
            v[k][P:].zero_()
        return views

    def step(self, flat_grads: torch.Tensor, grad_scale: float = 1.0, zero_grads: bool = False):
        if not self.flat.is_cuda:
            raise RuntimeError("FlatAdam.step: CUDA tensors only (there is no CPU path)")
        if flat_grads.numel() != self.flat.numel() or flat_grads.dtype != torch.float32 or not flat_grads.is_contiguous() \
                or flat_grads.device != self.flat.device:
            raise RuntimeError("FlatAdam.step: gradient bucket must be a contiguous fp32 CUDA tensor of the parameter size")
        self.steps += 1
        lr = (ctypes.c_double * len(self.names))(*[float(self.lrs[k]) for k in self.names])
        with torch.cuda.device(self.flat.device):
            _check(_lib.gsr_adam_step(torch.cuda.current_stream(self.flat.device).cuda_stream, self.flat.data_ptr(),
                                      flat_grads.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                      self.flat.numel(), len(self.names), self.seg_end, lr, self.beta1, self.beta2, self.eps,
                                      self.steps, float(grad_scale), int(bool(zero_grads))))
