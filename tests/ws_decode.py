"""Decode the B200 library's opaque workspaces (layout introspection entry points of the C-ABI)."""
import ctypes

import numpy as np
import torch

import diff_gaussian_rasterization as dgr


class _GeomLayout(ctypes.Structure):
    _fields_ = [(n, ctypes.c_size_t) for n in ("rec", "rects", "depth_keys", "sorted_ids", "counters", "total")]


class _ImgLayout(ctypes.Structure):
    _fields_ = [(n, ctypes.c_size_t) for n in ("final_T", "n_contrib", "ranges", "total")]


class _BinLayout(ctypes.Structure):
    _fields_ = [(n, ctypes.c_size_t) for n in ("point_list", "total")]


def _arr(buf, off, dt, count):
    raw = buf.detach().cpu().numpy()
    a = raw[off:off + np.dtype(dt).itemsize * count].view(dt).copy()
    if np.dtype(dt).kind == "u" and np.dtype(dt).itemsize > 1:
        a = a.astype(np.int64)
    return torch.from_numpy(a)


def decode_geom(geom, P, W, H):
    lay = _GeomLayout()
    dgr._lib.gsr_geom_layout_of(ctypes.c_int32(P), ctypes.c_int32(W), ctypes.c_int32(H), ctypes.byref(lay))
    rec = _arr(geom, lay.rec, np.float32, 12 * P).view(P, 12)
    bits = rec[:, 11].contiguous().view(torch.int32)
    rects = _arr(geom, lay.rects, np.uint16, 4 * P).view(P, 4)
    return dict(means2D=rec[:, 0:2].contiguous(), depths=rec[:, 2].contiguous(), cull_r2=rec[:, 3].contiguous(),
                conic_opacity=rec[:, 4:8].contiguous(), rgb=rec[:, 8:11].contiguous(),
                clamped=torch.stack([(bits & 1) != 0, (bits & 2) != 0, (bits & 4) != 0], -1),
                rects=rects, tiles_touched=(rects[:, 2] - rects[:, 0]) * (rects[:, 3] - rects[:, 1]),
                depth_keys=_arr(geom, lay.depth_keys, np.uint32, P),
                sorted_ids=_arr(geom, lay.sorted_ids, np.uint32, 2 * P).view(P, 2)[:, 1].contiguous())


def decode_img(img, W, H):
    lay = _ImgLayout()
    dgr._lib.gsr_img_layout_of(ctypes.c_int32(W), ctypes.c_int32(H), ctypes.byref(lay))
    T = ((W + 15) // 16) * ((H + 15) // 16)
    return dict(final_T=_arr(img, lay.final_T, np.float32, W * H).view(H, W),
                n_contrib=_arr(img, lay.n_contrib, np.uint32, W * H).view(H, W),
                ranges=_arr(img, lay.ranges, np.uint32, 2 * T).view(T, 2))


def decode_binning(binning, R, geom_dec, img_dec):
    """point_list plus the (tile << 32 | depth bits) key of every instance, reconstructed from the tile
    ranges and the per-Gaussian depth keys (the library never materialises 64-bit keys)."""
    lay = _BinLayout()
    dgr._lib.gsr_binning_layout_of(ctypes.c_int64(R), ctypes.byref(lay))
    pl = _arr(binning, lay.point_list, np.uint32, R)
    rng = img_dec["ranges"]
    tile_of = torch.zeros(R, dtype=torch.int64)
    for t in torch.nonzero(rng[:, 1] > rng[:, 0]).reshape(-1).tolist():
        tile_of[int(rng[t, 0]):int(rng[t, 1])] = t
    keys = (tile_of << 32) | geom_dec["depth_keys"][pl]
    return dict(point_list=pl, keys=keys)
