"""Thin access layer to the COMPILED, UNMODIFIED reference rasterizer (oracle/_ref/ref_dgr_C.so).

TEST INFRASTRUCTURE ONLY (see oracle/build_ref.py).  Used by the GPU parity tests, by
tests/golden/make_golden.py and by `bench.py --impl reference`.  Never imported by the product.

The reference's own Python wrapper (DGR/diff_gaussian_rasterization/__init__.py) cannot be
imported from here on the GPU box (/root/reference does not exist there) and does
`from . import _C`; this module re-states the little it does — argument order of the two pybind
entry points (DGR/ext.cpp:15-19, DGR/rasterize_points.h:18-67) and the mapping of the 8 native
gradients onto the autograd inputs (DGR/diff_gaussian_rasterization/__init__.py:143-153) — and adds
decoders for the reference's opaque scratch buffers so that intermediate state (sorted keys,
tile ranges, per-Gaussian 2-D quantities) can be compared bit for bit.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_C = None


def available():
    return os.path.exists(os.path.join(_HERE, "_ref", "ref_dgr_C.so"))


def load():
    global _C
    if _C is None:
        if not available():
            raise RuntimeError("oracle/_ref/ref_dgr_C.so missing: run `python oracle/build_ref.py` where /root/reference exists")
        p = os.path.join(_HERE, "_ref")
        if p not in sys.path:
            sys.path.insert(0, p)
        import ref_dgr_C  # noqa
        _C = ref_dgr_C
    return _C


def _e(t):
    return torch.Tensor([]) if t is None else t


def forward(means3D, opacities, viewmatrix, projmatrix, campos, bg, W, H, tanfovx, tanfovy, scales=None,
            rotations=None, scale_modifier=1.0, cov3D_precomp=None, shs=None, sh_degree=0, colors_precomp=None,
            prefiltered=False, debug=False):
    """_C.rasterize_gaussians -> dict(num_rendered, color, radii, geom, binning, img)."""
    C = load()
    R, color, radii, geom, binning, img = C.rasterize_gaussians(
        bg, means3D, _e(colors_precomp), opacities, _e(scales), _e(rotations), float(scale_modifier),
        _e(cov3D_precomp), viewmatrix, projmatrix, float(tanfovx), float(tanfovy), int(H), int(W), _e(shs),
        int(sh_degree), campos, bool(prefiltered), bool(debug))
    return dict(num_rendered=R, color=color, radii=radii, geom=geom, binning=binning, img=img)


def backward(fwd, dL_dcolor, means3D, viewmatrix, projmatrix, campos, bg, tanfovx, tanfovy, scales=None,
             rotations=None, scale_modifier=1.0, cov3D_precomp=None, shs=None, sh_degree=0, colors_precomp=None,
             debug=False):
    """_C.rasterize_gaussians_backward -> dict of the 8 reference gradients."""
    C = load()
    g = C.rasterize_gaussians_backward(
        bg, means3D, fwd["radii"], _e(colors_precomp), _e(scales), _e(rotations), float(scale_modifier),
        _e(cov3D_precomp), viewmatrix, projmatrix, float(tanfovx), float(tanfovy), dL_dcolor, _e(shs),
        int(sh_degree), campos, fwd["geom"], int(fwd["num_rendered"]), fwd["binning"], fwd["img"], bool(debug))
    names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales",
             "dL_drotations"]
    return dict(zip(names, g))


class RefRasterize(torch.autograd.Function):
    """Autograd wrapper over the compiled reference with the reference's own input order
    (means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, settings)."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, rs):
        f = forward(means3D, opacities, rs.viewmatrix, rs.projmatrix, rs.campos, rs.bg, rs.image_width,
                    rs.image_height, rs.tanfovx, rs.tanfovy, scales, rotations, rs.scale_modifier, cov3Ds_precomp, sh,
                    rs.sh_degree, colors_precomp, rs.prefiltered, rs.debug)
        ctx.rs = rs
        ctx.R = f["num_rendered"]
        ctx.save_for_backward(means3D, sh, colors_precomp, scales, rotations, cov3Ds_precomp, f["radii"], f["geom"],
                              f["binning"], f["img"])
        ctx.mark_non_differentiable(f["radii"])
        return f["color"], f["radii"]

    @staticmethod
    def backward(ctx, g_color, _):
        rs = ctx.rs
        means3D, sh, colors_precomp, scales, rotations, cov3Ds_precomp, radii, geom, binning, img = ctx.saved_tensors
        f = dict(num_rendered=ctx.R, radii=radii, geom=geom, binning=binning, img=img)
        g = backward(f, g_color.contiguous(), means3D, rs.viewmatrix, rs.projmatrix, rs.campos, rs.bg, rs.tanfovx,
                     rs.tanfovy, scales, rotations, rs.scale_modifier, cov3Ds_precomp, sh, rs.sh_degree,
                     colors_precomp, rs.debug)
        return (g["dL_dmeans3D"], g["dL_dmeans2D"], g["dL_dsh"], g["dL_dcolors"], g["dL_dopacity"], g["dL_dscales"],
                g["dL_drotations"], g["dL_dcov3D"], None)


def rasterize(means3D, means2D, opacities, rs, shs=None, colors_precomp=None, scales=None, rotations=None,
              cov3D_precomp=None):
    return RefRasterize.apply(means3D, means2D, _e(shs), _e(colors_precomp), opacities, _e(scales), _e(rotations),
                              _e(cov3D_precomp), rs)


# ---- decoders of the reference's scratch buffers (layouts: CR/rasterizer_impl.cu:155-194) -------
def _walk(buf, fields):
    """fields: list of (name, numpy dtype, count). 128-byte aligned bump allocation."""
    raw = buf.detach().cpu().numpy()
    base = buf.data_ptr()
    off = 0
    out = {}
    for name, dt, count in fields:
        a = (base + off + 127) // 128 * 128 - base
        nbytes = np.dtype(dt).itemsize * count
        arr = raw[a:a + nbytes].view(dt).copy()
        if np.dtype(dt).kind == "u" and np.dtype(dt).itemsize > 1:
            arr = arr.astype(np.int64)      # torch has only partial unsigned support
        out[name] = torch.from_numpy(arr)
        off = a + nbytes
    return out


def decode_geom(geom, P):
    d = _walk(geom, [("depths", np.float32, P), ("clamped", np.uint8, 3 * P), ("internal_radii", np.int32, P),
                     ("means2D", np.float32, 2 * P), ("cov3D", np.float32, 6 * P), ("conic_opacity", np.float32, 4 * P),
                     ("rgb", np.float32, 3 * P), ("tiles_touched", np.uint32, P)])
    d["means2D"] = d["means2D"].view(P, 2)
    d["cov3D"] = d["cov3D"].view(P, 6)
    d["conic_opacity"] = d["conic_opacity"].view(P, 4)
    d["rgb"] = d["rgb"].view(P, 3)
    d["clamped"] = d["clamped"].view(P, 3)
    return d


def decode_binning(binning, R):
    d = _walk(binning, [("point_list", np.uint32, R), ("point_list_unsorted", np.uint32, R), ("keys", np.uint64, R),
                        ("keys_unsorted", np.uint64, R)])
    return d


def decode_img(img, W, H):
    N = W * H
    d = _walk(img, [("final_T", np.float32, N), ("n_contrib", np.uint32, N), ("ranges", np.uint32, 2 * N)])
    d["final_T"] = d["final_T"].view(H, W)
    d["n_contrib"] = d["n_contrib"].view(H, W)
    d["ranges"] = d["ranges"].view(N, 2)
    return d
