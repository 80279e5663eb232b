"""Drop-in replacement for the reference's `diff_gaussian_rasterization` package, backed by the
B200-native C-ABI library `libgsrast_b200.so` (include/gsrast_b200.h).

Mirrors the public surface of DGR/diff_gaussian_rasterization/__init__.py
(DGR = /root/reference/submodules/diff-gaussian-rasterization):

    GaussianRasterizationSettings   (__init__.py:157-169)  same fields, same order
    GaussianRasterizer              (__init__.py:171-220)  .forward(...) -> (color[3,H,W], radii[P]); .markVisible
    rasterize_gaussians             (__init__.py:21-42)    same positional signature

so `slam/renderer.py` (R/slam/renderer.py:15-18,140,196-214) imports and calls it unchanged.
Same error behaviour: plain `Exception` for the SH/colour and scale/covariance exclusivity checks
(__init__.py:191-195), `RuntimeError` for a mis-shaped means3D (DGR/rasterize_points.cu:57-59),
argument snapshot on failure when `debug` is set (__init__.py:83-90,132-139).

Differences, all additive:
  * work is enqueued on torch's CURRENT stream (the reference uses the legacy default stream);
  * `viewmatrix`, `projmatrix` and `campos` are differentiable: if the tensors inside
    `raster_settings` require grad they receive gradients;
  * there is no CPU path and no fallback: importing this module without the built CUDA library
    raises ImportError, and calling it with non-CUDA tensors raises RuntimeError.

Host code here is plumbing only (allocation + ctypes marshalling); all compute is in the library.
"""
from __future__ import annotations

import ctypes
import os
from typing import NamedTuple

import torch
import torch.nn as nn

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians", "prepare_forward",
           "DEPTH_SILHOUETTE"]

# ------------------------------------------------------------------------------------------------
# Library loading
# ------------------------------------------------------------------------------------------------
_LIB_PATH = os.environ.get(
    "GSRAST_B200_LIB",
    os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "lib", "libgsrast_b200.so"),
)


class _Gaussians(ctypes.Structure):
    _fields_ = [
        ("P", ctypes.c_int32), ("sh_degree", ctypes.c_int32), ("sh_coeffs", ctypes.c_int32), ("raw_params", ctypes.c_int32),
        ("means3D", ctypes.c_void_p), ("shs", ctypes.c_void_p), ("colors_precomp", ctypes.c_void_p),
        ("opacities", ctypes.c_void_p), ("scales", ctypes.c_void_p), ("rotations", ctypes.c_void_p),
        ("cov3D_precomp", ctypes.c_void_p), ("scale_modifier", ctypes.c_float), ("extra_mode", ctypes.c_int32),
        ("extra_colors", ctypes.c_void_p),
    ]


class _Camera(ctypes.Structure):
    _fields_ = [
        ("width", ctypes.c_int32), ("height", ctypes.c_int32), ("tanfovx", ctypes.c_float), ("tanfovy", ctypes.c_float),
        ("viewmatrix", ctypes.c_void_p), ("projmatrix", ctypes.c_void_p), ("campos", ctypes.c_void_p),
        ("background", ctypes.c_void_p), ("prefiltered", ctypes.c_int32), ("debug", ctypes.c_int32),
    ]


class _Grads(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in (
        "dL_dmeans2D", "dL_dconic", "dL_dopacity", "dL_dcolors", "dL_dmeans3D", "dL_dcov3D", "dL_dsh",
        "dL_dscales", "dL_drotations", "dL_dviewmatrix", "dL_dprojmatrix", "dL_dcampos")] + [
        ("accumulate", ctypes.c_int32), ("_pad", ctypes.c_int32), ("dL_dextra", ctypes.c_void_p),
        ("dL_dopacity_raw", ctypes.c_void_p)]


def _load():
    if not os.path.exists(_LIB_PATH):
        raise ImportError(
            f"diff_gaussian_rasterization (B200): CUDA library not found at {_LIB_PATH}. "
            "Build it with `python mm3dgs-slam_b200/build.py` (or __graft_entry__.build()). "
            "There is no CPU fallback.")
    lib = ctypes.CDLL(_LIB_PATH)
    vp, i32, i64, sz = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_size_t
    lib.gsr_abi_version.restype = ctypes.c_int
    lib.gsr_last_error.restype = ctypes.c_char_p
    lib.gsr_geom_ws_bytes.restype = sz
    lib.gsr_geom_ws_bytes.argtypes = [i32, i32, i32]
    lib.gsr_img_ws_bytes.restype = sz
    lib.gsr_img_ws_bytes.argtypes = [i32, i32]
    lib.gsr_binning_ws_bytes.restype = sz
    lib.gsr_binning_ws_bytes.argtypes = [i64]
    lib.gsr_forward_preprocess.restype = ctypes.c_int
    lib.gsr_forward_preprocess.argtypes = [vp, ctypes.POINTER(_Gaussians), ctypes.POINTER(_Camera), vp, vp, sz, vp, sz,
                                           ctypes.POINTER(i32)]
    lib.gsr_forward_render.restype = ctypes.c_int
    lib.gsr_forward_render.argtypes = [vp, ctypes.POINTER(_Gaussians), ctypes.POINTER(_Camera), vp, i64, vp, vp, sz, vp, vp, vp]
    lib.gsr_backward.restype = ctypes.c_int
    lib.gsr_backward.argtypes = [vp, ctypes.POINTER(_Gaussians), ctypes.POINTER(_Camera), vp, i64, vp, vp, vp, vp, vp,
                                 ctypes.POINTER(_Grads)]
    lib.gsr_mark_visible.restype = ctypes.c_int
    lib.gsr_mark_visible.argtypes = [vp, i32, vp, vp, vp, vp]
    if lib.gsr_abi_version() != 6:
        raise ImportError("libgsrast_b200.so ABI version mismatch")
    return lib


_lib = _load()


def _check(rc):
    if rc != 0:
        msg = _lib.gsr_last_error()
        raise RuntimeError("gsrast_b200: " + (msg.decode() if msg else f"error {rc}"))


def _ptr(t):
    """Device pointer of a tensor, or NULL for the reference's 'empty tensor = absent' convention."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _prep(t, device):
    """fp32, contiguous, on `device` (no copy when already so)."""
    if t is None or t.numel() == 0:
        return t
    if t.dtype != torch.float32 or t.device != device or not t.is_contiguous():
        t = t.to(device=device, dtype=torch.float32).contiguous()
    return t


def _snapshot(args, path):
    """Argument dump on failure (debug mode), like the reference's: tensors and plain scalars only — handles that wrap
    ctypes structs / streams / events (PreparedFrame) cannot be pickled and would hide the original error."""
    def keep(a):
        if isinstance(a, torch.Tensor):
            return a.detach().cpu().clone()
        if isinstance(a, tuple) and hasattr(a, "_fields"):     # raster settings
            return type(a)(*[keep(x) for x in a])
        return a if isinstance(a, (int, float, bool, str, type(None))) else None
    torch.save(tuple(keep(a) for a in args), path)


# ------------------------------------------------------------------------------------------------
# Native calls
# ------------------------------------------------------------------------------------------------
DEPTH_SILHOUETTE = "depth_silhouette"   # extra_colors=DEPTH_SILHOUETTE: (z, 1, z^2) generated inside the library


def _structs(means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, rs, view, proj, campos, bg,
             extra=None, raw=False):
    """extra: None, a [P,3] tensor, or DEPTH_SILHOUETTE.  raw: opacities / scales / rotations are raw parameters."""
    P = means3D.shape[0]
    M = sh.shape[1] if (sh is not None and sh.numel() != 0) else 0
    gen = isinstance(extra, str)
    g = _Gaussians(P, int(rs.sh_degree), M, 1 if raw else 0, _ptr(means3D), _ptr(sh), _ptr(colors_precomp), _ptr(opacities),
                   _ptr(scales), _ptr(rotations), _ptr(cov3Ds_precomp), float(rs.scale_modifier), 1 if gen else 0,
                   None if gen else _ptr(extra))
    c = _Camera(int(rs.image_width), int(rs.image_height), float(rs.tanfovx), float(rs.tanfovy), _ptr(view), _ptr(proj),
                _ptr(campos), _ptr(bg), int(bool(rs.prefiltered)), int(bool(rs.debug)))
    return g, c


class _GeomLayout(ctypes.Structure):
    _fields_ = [(n, ctypes.c_size_t) for n in ("rec", "rects", "depth_keys", "sorted_ids", "counters", "total")]


class _Count(int):
    """Instance count of a forward: the true number of tile instances (compares like an int), with the capacity the
    binning workspace was carved with in `.cap` (what gsr_backward must be given; cap >= count)."""
    cap = 0


class _SlotPool:
    """Pinned int32 slots for the device->host hand-off of the instance count R (cudaHostAlloc is far too slow to
    call per frame).  A slot is handed out once and comes back when its consumer has read it, so any number of
    frames can be in flight without two of them sharing a slot; the pool grows in blocks of 64."""

    def __init__(self):
        self.free = []
        self.blocks = []

    def get(self):
        if not self.free:
            blk = torch.empty(64, dtype=torch.int32).pin_memory()
            self.blocks.append(blk)
            self.free.extend(blk[i: i + 1] for i in range(64))
        return self.free.pop()

    def put(self, slot):
        self.free.append(slot)


_slots = _SlotPool()
# last seen instance counts per (device, P, W, H): capacity estimate of the next forward's instance list
_r_seen = {}


def _validate(means3D, extra_colors):
    """Argument checks shared by the autograd forward and prepare_forward (same messages as the reference where it
    has them: DGR/rasterize_points.cu:57-59)."""
    if means3D.dim() != 2 or means3D.shape[1] != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    if not means3D.is_cuda:
        raise RuntimeError("gsrast_b200: tensors must live on a CUDA device (there is no CPU path)")
    if isinstance(extra_colors, str):
        if extra_colors != DEPTH_SILHOUETTE:
            raise ValueError("extra_colors must be a [P,3] tensor or diff_gaussian_rasterization.DEPTH_SILHOUETTE")
    elif extra_colors is not None and extra_colors.numel() != 0:
        if extra_colors.dim() != 2 or tuple(extra_colors.shape) != (means3D.shape[0], 3):
            raise RuntimeError("extra_colors must have dimensions (num_points, 3)")


def _extra_key(extra):
    """Identity of the extra-colour argument a frame was prepared / rendered with."""
    if isinstance(extra, str):
        return extra
    if extra is None or extra.numel() == 0:
        return None
    return extra.data_ptr()


class PreparedFrame:
    """Phase 1 of a forward (projection + depth sort) already enqueued; see prepare_forward()."""
    __slots__ = ("tensors", "rs", "radii", "geom", "img", "r_host", "event", "stream", "g", "c", "key", "raw")

    def release(self):
        if getattr(self, "r_host", None) is not None:
            _slots.put(self.r_host)
            self.r_host = None


def _input_key(t, extra):
    """Device pointers of everything phase 1 read: a prepared handle is only valid for exactly these inputs."""
    return tuple(None if (x is None or x.numel() == 0) else x.data_ptr() for x in t) + (_extra_key(extra),)


def _enqueue_r_copy(geom, P, W, H, stream):
    """Asynchronous copy of the device-side instance count into a pinned slot + an event after it."""
    lay = _GeomLayout()
    _lib.gsr_geom_layout_of(ctypes.c_int32(P), ctypes.c_int32(W), ctypes.c_int32(H), ctypes.byref(lay))
    slot = _slots.get()
    slot.copy_(geom[lay.counters: lay.counters + 4].view(torch.int32), non_blocking=True)
    ev = torch.cuda.Event()
    ev.record(stream)
    return slot, ev


def prepare_forward(means3D, opacities, raster_settings, shs=None, colors_precomp=None, scales=None, rotations=None,
                    cov3D_precomp=None, extra_colors=None, raw_params=False):
    """Extension.  Enqueue phase 1 of a forward (projection, tile rectangles, instance count R, depth sort)
    without waiting for R, and return a handle to pass as `prepared=` to GaussianRasterizer.forward with the
    SAME tensors and settings.  Issuing phase 1 of several frames before the first forward() means every
    frame's R is already on the host when its forward() needs it — the host never stalls on the hand-off."""
    rs = raster_settings
    _validate(means3D, extra_colors)
    if (shs is None) == (colors_precomp is None):
        raise Exception("Please provide excatly one of either SHs or precomputed colors!")
    if ((scales is None or rotations is None) and cov3D_precomp is None) or (
            (scales is not None or rotations is not None) and cov3D_precomp is not None):
        raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
    dev = means3D.device
    e = torch.Tensor([])
    with torch.cuda.device(dev):
        t = [_prep(x if x is not None else e, dev) for x in (means3D, shs, colors_precomp, opacities, scales, rotations,
                                                             cov3D_precomp, rs.viewmatrix, rs.projmatrix, rs.campos,
                                                             rs.bg)]
        extra = extra_colors if isinstance(extra_colors, str) else _prep(extra_colors, dev)
        P = means3D.shape[0]
        H, W = int(rs.image_height), int(rs.image_width)
        pf = PreparedFrame()
        pf.tensors, pf.rs = tuple(t) + (extra,), rs
        pf.key = _input_key(t, extra)
        pf.raw = bool(raw_params)
        pf.stream = torch.cuda.current_stream(dev)
        pf.radii = torch.empty((P,), dtype=torch.int32, device=dev)
        pf.r_host = None
        if P == 0:
            pf.geom = pf.img = pf.event = None
            return pf
        u8 = dict(dtype=torch.uint8, device=dev)
        pf.geom = torch.empty(_lib.gsr_geom_ws_bytes(P, W, H), **u8)
        pf.img = torch.empty(_lib.gsr_img_ws_bytes(W, H), **u8)
        pf.g, pf.c = _structs(*t[:7], rs, *t[7:11],
                              extra if (isinstance(extra, str) or (extra is not None and extra.numel())) else None,
                              raw=pf.raw)
        _check(_lib.gsr_forward_preprocess(pf.stream.cuda_stream, ctypes.byref(pf.g), ctypes.byref(pf.c),
                                           pf.radii.data_ptr(), pf.geom.data_ptr(), pf.geom.numel(), pf.img.data_ptr(),
                                           pf.img.numel(), None))
        pf.r_host, pf.event = _enqueue_r_copy(pf.geom, P, W, H, pf.stream)
    return pf


def _forward_native(means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, rs, view, proj, campos, bg,
                    extra=None, prepared=None, raw=False):
    """Returns (R, color, radii, geom, binning, img); with `extra` ([P,3] colours) the 7th element is the
    extra [3,H,W] image blended in the same pass.  R is a _Count: the true instance count, with the capacity the
    binning workspace was carved with (what gsr_backward needs) in R.cap.

    The host never waits for the GPU to drain: phase 2 is enqueued against a capacity ESTIMATE of the instance list
    (the largest count seen for this problem size, plus a margin) right behind phase 1, and only then does the host
    look at the true count — by which time the projection kernel that produces it has long finished while the rest of
    the forward is still queued.  If the estimate was too small (the kernels then leave the outputs untouched) phase 2
    is simply enqueued again with the exact size."""
    dev = means3D.device
    P = means3D.shape[0]
    H, W = int(rs.image_height), int(rs.image_width)
    cur = torch.cuda.current_stream(dev)
    stream = cur.cuda_stream
    u8 = dict(dtype=torch.uint8, device=dev)
    color = torch.empty((3, H, W), dtype=torch.float32, device=dev)
    has_extra = isinstance(extra, str) or (extra is not None and extra.numel() != 0)
    extra_img = torch.empty((3, H, W), dtype=torch.float32, device=dev) if has_extra else None
    if P == 0:
        color.zero_()
        e = torch.empty(0, **u8)
        radii = torch.empty((0,), dtype=torch.int32, device=dev)
        return (_Count(0), color, radii, e, e, e) + ((torch.zeros_like(color),) if extra is not None else ())
    R = None
    if prepared is not None:
        key = _input_key((means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, view, proj, campos,
                          bg), extra)
        if prepared.key != key or prepared.radii.shape[0] != P or prepared.rs is not rs or prepared.raw != bool(raw):
            raise RuntimeError("gsrast_b200: `prepared` was made for different inputs / settings")
        radii, geom, img, g, c = prepared.radii, prepared.geom, prepared.img, prepared.g, prepared.c
        if cur != prepared.stream:
            cur.wait_stream(prepared.stream)
        slot, event = prepared.r_host, prepared.event
        prepared.r_host = None                     # the slot goes back to the pool below
        if event.query():                          # usually true: phase 1 was enqueued frames ago
            R = int(slot.item())
    else:
        radii = torch.empty((P,), dtype=torch.int32, device=dev)
        geom = torch.empty(_lib.gsr_geom_ws_bytes(P, W, H), **u8)
        img = torch.empty(_lib.gsr_img_ws_bytes(W, H), **u8)
        g, c = _structs(means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, rs, view, proj, campos,
                        bg, extra if has_extra else None, raw=raw)
        _check(_lib.gsr_forward_preprocess(stream, ctypes.byref(g), ctypes.byref(c), radii.data_ptr(), geom.data_ptr(),
                                           geom.numel(), img.data_ptr(), img.numel(), None))
        slot, event = _enqueue_r_copy(geom, P, W, H, cur)
    ckey = (dev.index, P, W, H)
    seen = _r_seen.get(ckey)

    def render(cap):
        binning = torch.empty(_lib.gsr_binning_ws_bytes(cap), **u8)
        _check(_lib.gsr_forward_render(stream, ctypes.byref(g), ctypes.byref(c), radii.data_ptr(), cap, geom.data_ptr(),
                                       binning.data_ptr(), binning.numel(), img.data_ptr(), color.data_ptr(),
                                       _ptr(extra_img)))
        return binning

    if R is None and seen is not None:
        cap = seen + (seen >> 2) + 65536           # estimate: phase 2 goes out before the host has seen R
        binning = render(cap)
        event.synchronize()
        R = int(slot.item())
    else:
        if R is None:                              # first call for this problem size: wait for the count
            event.synchronize()
            R = int(slot.item())
        cap, binning = -1, None
    _slots.put(slot)
    if R < 0:
        raise RuntimeError("gsrast_b200: more than 2^31-1 tile instances")
    _r_seen[ckey] = R if seen is None else max(R, seen - (seen >> 6))   # running maximum with a slow decay
    if R > cap:
        cap = R
        binning = render(cap)
    R = _Count(R)
    R.cap = cap
    return (R, color, radii, geom, binning, img) + ((extra_img,) if has_extra else ())


def _backward_native(grad_color, means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, rs, view,
                     proj, campos, bg, radii, R, geom, binning, img, want_cam, targets=None, extra=None,
                     grad_extra_img=None, raw=False):
    """targets: optional dict of gradient accumulators (means3D, shs, opacities, scales, rotations) the
    kernels add into directly (gsr_grads.accumulate)."""
    dev = means3D.device
    P = means3D.shape[0]
    M = sh.shape[1] if (sh is not None and sh.numel() != 0) else 0
    f32 = dict(dtype=torch.float32, device=dev)
    has_sr = scales is not None and scales.numel() != 0
    has_cov = cov3Ds_precomp is not None and cov3Ds_precomp.numel() != 0
    # accumulated-into buffers share one zero fill: [means2D 3 | conic 4 | colors 3 | opacity 1] per Gaussian
    # (raw parameters: the blend backward's opacity sums are scratch — the chain through the sigmoid is applied by the
    #  per-Gaussian backward, which writes / adds dL/d(raw opacity))
    acc = torch.zeros(P * (10 if (targets and not raw) else 11), **f32)
    g_means2D = acc[: 3 * P].view(P, 3)
    g_conic = acc[3 * P: 7 * P].view(P, 4)
    g_colors = acc[7 * P: 10 * P].view(P, 3)
    g_opacity_raw = None
    if targets:
        g_opacity, g_means3D = targets["opacities"], targets["means3D"]   # atomics / += land in the accumulators
        if raw:
            g_opacity_raw, g_opacity = targets["opacities"], acc[10 * P: 11 * P].view(P, 1)
        g_sh = targets["shs"] if M > 0 else None
        g_scales = targets["scales"] if has_sr else None
        g_rots = targets["rotations"] if has_sr else None
        g_cov3D = None
    else:
        g_opacity = acc[10 * P: 11 * P].view(P, 1)
        if raw:
            g_opacity_raw = torch.empty((P, 1), **f32)
        g_means3D = torch.empty((P, 3), **f32)
        g_cov3D = torch.empty((P, 6), **f32) if has_cov else None
        g_sh = torch.empty((P, M, 3), **f32) if M > 0 else None
        g_scales = torch.empty((P, 3), **f32) if has_sr else None
        g_rots = torch.empty((P, 4), **f32) if has_sr else None
    g_cam = torch.zeros(35, **f32) if want_cam else None
    has_extra = isinstance(extra, str) or (extra is not None and extra.numel() != 0)
    g_extra = torch.zeros((P, 3), **f32) if has_extra else None
    if P == 0:
        return (g_means2D, g_colors, g_opacity_raw if raw else g_opacity, g_means3D, g_cov3D, g_sh, g_scales, g_rots, g_cam,
                g_extra)
    stream = torch.cuda.current_stream(dev).cuda_stream
    g, c = _structs(means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, rs, view, proj, campos, bg,
                    extra if has_extra else None, raw=raw)
    gr = _Grads(g_means2D.data_ptr(), g_conic.data_ptr(), g_opacity.data_ptr(), g_colors.data_ptr(),
                g_means3D.data_ptr(), _ptr(g_cov3D), _ptr(g_sh), _ptr(g_scales), _ptr(g_rots),
                g_cam.data_ptr() if want_cam else None,
                g_cam.data_ptr() + 64 if want_cam else None,
                g_cam.data_ptr() + 128 if want_cam else None,
                1 if (targets and not targets.get("_overwrite")) else 0, 0, _ptr(g_extra), _ptr(g_opacity_raw))
    _check(_lib.gsr_backward(stream, ctypes.byref(g), ctypes.byref(c), radii.data_ptr(), R, _ptr(geom), _ptr(binning),
                             _ptr(img), grad_color.data_ptr(), _ptr(grad_extra_img) if has_extra else None,
                             ctypes.byref(gr)))
    return (g_means2D, g_colors, g_opacity_raw if raw else g_opacity, g_means3D, g_cov3D, g_sh, g_scales, g_rots, g_cam,
            g_extra)


# ------------------------------------------------------------------------------------------------
# Public API (same names and argument meaning as the reference)
# ------------------------------------------------------------------------------------------------
def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings, grad_targets=None, extra_colors=None, prepared=None, raw_params=False):
    """`grad_targets` (extension, optional): dict with fp32 contiguous accumulators for means3D, shs,
    opacities, scales, rotations.  The backward kernels then ADD this call's gradients straight into them
    (e.g. views of the map step's flat bucket) and autograd receives no gradient for those inputs — for
    the case where the rasterizer inputs ARE the optimised tensors.  With grad_targets["_overwrite"] = True
    the kernels OVERWRITE means3D / shs / scales / rotations instead (first frame into a fresh accumulator:
    no zero fill, no read-modify-write); the opacity accumulator must still be zero on entry."""
    rs = raster_settings
    out = _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                    cov3Ds_precomp, rs, rs.viewmatrix, rs.projmatrix, rs.campos, grad_targets,
                                    extra_colors, prepared, bool(raw_params))
    # reference contract: (color, radii); with extra_colors: (color, extra_image, radii)
    return (out[0], out[1]) if extra_colors is None else (out[0], out[2], out[1])


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings, viewmatrix, projmatrix, campos, grad_targets=None, extra_colors=None, prepared=None,
                raw_params=False):
        rs = raster_settings
        if raw_params and (scales is None or scales.numel() == 0 or rotations is None or rotations.numel() == 0):
            raise ValueError("raw_params needs scales and rotations (not cov3D_precomp)")
        ctx.raw = bool(raw_params)
        if grad_targets:
            if cov3Ds_precomp is not None and cov3Ds_precomp.numel() != 0:
                raise ValueError("grad_targets is not supported with cov3D_precomp")
            need = ["means3D", "opacities"] + (["shs"] if sh is not None and sh.numel() else []) + (
                ["scales", "rotations"] if scales is not None and scales.numel() else [])
            src = dict(means3D=means3D, opacities=opacities, shs=sh, scales=scales, rotations=rotations)
            for k in need:
                t = grad_targets.get(k)
                if (t is None or t.dtype != torch.float32 or not t.is_contiguous() or t.device != means3D.device
                        or t.numel() != src[k].numel()):
                    raise ValueError(f"grad_targets['{k}'] must be a contiguous fp32 CUDA tensor shaped like the input")
        ctx.grad_targets = grad_targets if grad_targets else None
        _validate(means3D, extra_colors)
        dev = means3D.device
        with torch.cuda.device(dev):
            t = [_prep(x, dev) for x in (means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                                         viewmatrix, projmatrix, campos, rs.bg)]
            means3D_, sh_, colors_, opac_, scales_, rots_, cov_, view_, proj_, campos_, bg_ = t
            extra_ = extra_colors if isinstance(extra_colors, str) else _prep(extra_colors, dev)
            args = (means3D_, sh_, colors_, opac_, scales_, rots_, cov_, rs, view_, proj_, campos_, bg_, extra_, prepared,
                    ctx.raw)
            if rs.debug:
                try:
                    out = _forward_native(*args)
                except Exception:
                    _snapshot(args, "snapshot_fw.dump")
                    print("\ngsrast_b200: forward failed; arguments written to snapshot_fw.dump")
                    raise
            else:
                out = _forward_native(*args)
        num_rendered, color, radii, geom, binning, img = out[:6]
        extra_img = out[6] if len(out) > 6 else None
        ctx.raster_settings = rs
        ctx.num_rendered = int(num_rendered.cap)     # the capacity the binning workspace was carved with
        ctx.has_extra = extra_img is not None
        ctx.extra_gen = isinstance(extra_, str)
        ctx.save_for_backward(means3D_, sh_, colors_, opac_, scales_, rots_, cov_, view_, proj_, campos_, bg_, radii,
                              geom, binning, img,
                              extra_ if (ctx.has_extra and not ctx.extra_gen) else torch.empty(0))
        ctx.mark_non_differentiable(radii)
        if ctx.has_extra:
            return color, radii, extra_img
        return color, radii

    @staticmethod
    def backward(ctx, grad_out_color, _grad_radii, grad_out_extra=None):
        rs = ctx.raster_settings
        (means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, view, proj, campos, bg, radii,
         geom, binning, img, extra) = ctx.saved_tensors
        dev = means3D.device
        want_cam = any(ctx.needs_input_grad[9:12])
        with torch.cuda.device(dev):
            if grad_out_color is None:
                grad_out_color = torch.zeros((3, int(rs.image_height), int(rs.image_width)), device=dev)
            grad = _prep(grad_out_color, dev)
            grad_extra = None
            if ctx.has_extra:
                grad_extra = _prep(grad_out_extra, dev) if grad_out_extra is not None else torch.zeros_like(grad)
            args = (grad, means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, rs, view, proj,
                    campos, bg, radii, ctx.num_rendered, geom, binning, img, want_cam, ctx.grad_targets,
                    (DEPTH_SILHOUETTE if ctx.extra_gen else extra) if ctx.has_extra else None, grad_extra, ctx.raw)
            if rs.debug:
                try:
                    res = _backward_native(*args)
                except Exception:
                    _snapshot(args, "snapshot_bw.dump")
                    print("\ngsrast_b200: backward failed; arguments written to snapshot_bw.dump")
                    raise
            else:
                res = _backward_native(*args)
        g_means2D, g_colors, g_opacity, g_means3D, g_cov3D, g_sh, g_scales, g_rots, g_cam, g_extra = res
        if ctx.extra_gen:
            g_extra = None          # the chain through z was applied inside the library
        has_colors = colors_precomp is not None and colors_precomp.numel() != 0
        g_view = g_proj = g_campos = None
        if g_cam is not None:
            g_view = g_cam[0:16].view_as(view) if ctx.needs_input_grad[9] else None
            g_proj = g_cam[16:32].view_as(proj) if ctx.needs_input_grad[10] else None
            g_campos = g_cam[32:35].view_as(campos) if ctx.needs_input_grad[11] else None
        if ctx.grad_targets:   # already added into the accumulators by the kernels
            return (None, g_means2D, None, g_colors if has_colors else None, None, None, None, None,
                    None, g_view, g_proj, g_campos, None, g_extra, None, None)
        if opacities.dim() == 1:
            g_opacity = g_opacity.view(-1)
        return (g_means3D, g_means2D, g_sh, g_colors if has_colors else None, g_opacity, g_scales, g_rots, g_cov3D,
                None, g_view, g_proj, g_campos, None, g_extra, None, None)


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        """bool[P]: Gaussians that pass the near-plane frustum test for this camera."""
        rs = self.raster_settings
        with torch.no_grad():
            if not positions.is_cuda:
                raise RuntimeError("gsrast_b200: tensors must live on a CUDA device (there is no CPU path)")
            dev = positions.device
            pos = _prep(positions, dev)
            P = pos.shape[0]
            present = torch.zeros((P,), dtype=torch.uint8, device=dev)
            if P:
                with torch.cuda.device(dev):
                    view, proj = _prep(rs.viewmatrix, dev), _prep(rs.projmatrix, dev)
                    _check(_lib.gsr_mark_visible(torch.cuda.current_stream(dev).cuda_stream, P, pos.data_ptr(),
                                                 view.data_ptr(), proj.data_ptr(), present.data_ptr()))
            return present.bool()

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, grad_targets=None, extra_colors=None, prepared=None, raw_params=False):
        """Reference signature and return value (color[3,H,W], radii[P]).  Extensions (keyword-only in spirit):
        `extra_colors` [P,3] -> returns (color, extra_image[3,H,W], radii): the extra colours are blended in the
        same pass (the SLAM renderer's second, depth/silhouette call fused into the first); pass
        extra_colors=DEPTH_SILHOUETTE to have the library generate (z, 1, z^2) from the view-space depth itself;
        `grad_targets`: see rasterize_gaussians; `prepared`: handle from prepare_forward();
        `raw_params=True`: opacities / scales / rotations are the model's RAW parameters (_opacity, _scaling,
        _rotation) — the library applies sigmoid / exp / normalize itself (R/slam/gaussian_model.py:108-132) and the
        gradients come back w.r.t. the raw parameters, so the three activation ops and their autograd graph drop out
        of every render."""
        rs = self.raster_settings
        if (shs is None) == (colors_precomp is None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3D_precomp is None) or (
                (scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        empty = torch.Tensor([])
        return rasterize_gaussians(
            means3D, means2D,
            empty if shs is None else shs,
            empty if colors_precomp is None else colors_precomp,
            opacities,
            empty if scales is None else scales,
            empty if rotations is None else rotations,
            empty if cov3D_precomp is None else cov3D_precomp,
            rs, grad_targets, extra_colors, prepared, raw_params)
