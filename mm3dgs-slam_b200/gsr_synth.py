"""Seeded synthetic scenes for parity tests and benchmarks.

Host-side utility; pure PyTorch on CPU (generated on CPU then moved, so the same seed gives the
same scene on every box).  Camera conventions follow R/slam/renderer.py:47-83,117-124:
`viewmatrix` is the transposed world-to-camera matrix, `projmatrix = viewmatrix @ P^T` with P from
getProjectionMatrix2 (R/utils/graphics_utils.py:85-94), `campos` the camera centre.
"""
from __future__ import annotations

import math
from typing import NamedTuple, Optional

import torch

# TUM fr1 intrinsics at 640x480 (R/configs/TUM.yml:84-87)
_FX, _FY, _CX, _CY, _W0, _H0 = 517.3, 516.5, 318.6, 255.3, 640, 480
ZNEAR, ZFAR = 0.01, 100.0  # R/slam/renderer.py:51-52


class Camera(NamedTuple):
    W: int
    H: int
    fx: float
    fy: float
    cx: float
    cy: float
    tanfovx: float
    tanfovy: float
    viewmatrix: torch.Tensor   # [4,4] = w2c^T
    projmatrix: torch.Tensor   # [4,4] = w2c^T @ P^T
    campos: torch.Tensor       # [3]


def intrinsics(W, H):
    sx, sy = W / _W0, H / _H0
    return _FX * sx, _FY * sy, _CX * sx, _CY * sy


def projection_matrix(fx, fy, cx, cy, W, H, znear=ZNEAR, zfar=ZFAR):
    """OpenGL-style projection from pinhole intrinsics (restates getProjectionMatrix2)."""
    return torch.tensor([
        [2 * fx / W, 0.0, -(W - 2 * cx) / W, 0.0],
        [0.0, 2 * fy / H, -(H - 2 * cy) / H, 0.0],
        [0.0, 0.0, zfar / (zfar - znear), -(zfar * znear) / (zfar - znear)],
        [0.0, 0.0, 1.0, 0.0]], dtype=torch.float32)


def make_camera(W, H, w2c: Optional[torch.Tensor] = None) -> Camera:
    fx, fy, cx, cy = intrinsics(W, H)
    if w2c is None:
        w2c = torch.eye(4)
    w2c = w2c.to(torch.float32)
    view = w2c.t().contiguous()
    proj = (view @ projection_matrix(fx, fy, cx, cy, W, H).t()).contiguous()
    campos = torch.linalg.inv(view)[3, :3].contiguous()
    return Camera(W, H, fx, fy, cx, cy, W / (2 * fx), H / (2 * fy), view, proj, campos)


def look_at_w2c(eye, target, up=(0.0, -1.0, 0.0)):
    """World-to-camera with +z forward, +x right, +y down (the SLAM/OpenCV frame)."""
    eye = torch.as_tensor(eye, dtype=torch.float64)
    target = torch.as_tensor(target, dtype=torch.float64)
    f = target - eye
    f = f / f.norm()
    upv = torch.as_tensor(up, dtype=torch.float64)
    r = torch.linalg.cross(f, -upv)
    r = r / r.norm()
    d = torch.linalg.cross(f, r)
    Rm = torch.stack([r, d, f], 0)
    w2c = torch.eye(4, dtype=torch.float64)
    w2c[:3, :3] = Rm
    w2c[:3, 3] = -Rm @ eye
    return w2c.to(torch.float32)


def orbit_cameras(W, H, K, centroid, radius=0.5):
    """K poses on a circle of `radius` metres around the origin, looking at the cloud centroid."""
    cams = []
    for k in range(K):
        a = 2 * math.pi * k / K
        eye = (radius * math.cos(a), radius * math.sin(a), 0.0)
        cams.append(make_camera(W, H, look_at_w2c(eye, centroid)))
    return cams


def make_gaussians(P, W, H, seed=0, sh_degree=0, behind_frac=0.02, big_frac=0.01):
    """Seeded Gaussian cloud in the camera frame of the identity pose."""
    g = torch.Generator().manual_seed(seed)
    fx, fy, cx, cy = intrinsics(W, H)
    z = torch.rand(P, generator=g) * 7.5 + 0.5
    behind = torch.rand(P, generator=g) < behind_frac
    z = torch.where(behind, torch.rand(P, generator=g) * 0.2 - 0.1, z)   # z in (-0.1, 0.1): near-culled
    u = (torch.rand(P, generator=g) * 1.2 - 0.1) * W
    v = (torch.rand(P, generator=g) * 1.2 - 0.1) * H
    zz = torch.where(behind, torch.ones_like(z), z)
    means = torch.stack([(u - cx) / fx * zz, (v - cy) / fy * zz, z], -1)
    logs = torch.log(1.5 * zz / fx)[:, None] + 0.5 * torch.randn(P, 3, generator=g)
    big = torch.rand(P, generator=g) < big_frac
    scales = torch.exp(logs) * torch.where(big, 20.0, 1.0)[:, None]
    q = torch.randn(P, 4, generator=g)
    rotations = q / q.norm(dim=1, keepdim=True)
    opacities = torch.sigmoid(1.5 * torch.randn(P, 1, generator=g))
    M = (sh_degree + 1) ** 2
    shs = torch.zeros(P, M, 3)
    shs[:, 0] = (torch.rand(P, 3, generator=g) - 0.5) / 0.28209479177387814
    if M > 1:
        shs[:, 1:] = 0.1 * torch.randn(P, M - 1, 3, generator=g)
    return dict(means3D=means.float().contiguous(), scales=scales.float().contiguous(),
                rotations=rotations.float().contiguous(), opacities=opacities.float().contiguous(),
                shs=shs.float().contiguous())


def make_scene(P, W, H, seed=0, sh_degree=0):
    """(gaussians dict, identity-pose Camera, fixed dL/dpix [3,H,W], bg [3])."""
    gs = make_gaussians(P, W, H, seed, sh_degree)
    cam = make_camera(W, H)
    g = torch.Generator().manual_seed(seed + 12345)
    dL = torch.randn(3, H, W, generator=g)
    bg = torch.zeros(3)
    return gs, cam, dL, bg


def make_loss_inputs(W, H, seed=0, nan_frac=0.002):
    """Seeded synthetic inputs of the image losses (CPU tensors): a rendered colour image and a ground truth that is
    a blurred, noisy copy of it (so SSIM is away from both 0 and 1), a rendered (depth, silhouette, depth^2) image with
    a band of low silhouette, a few NaN depths, and a ground-truth / estimated depth pair with invalid (zero) pixels."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
    base = torch.stack([0.5 + 0.4 * torch.sin(7 * xx + 3 * yy), 0.5 + 0.4 * torch.cos(5 * yy - 2 * xx), 0.3 + 0.5 * xx * yy])
    image = (base + 0.08 * torch.randn(3, H, W, generator=g)).clamp(0, 1)
    gt = torch.nn.functional.avg_pool2d(image.unsqueeze(0), 3, 1, 1).squeeze(0) + 0.05 * torch.randn(3, H, W, generator=g)
    gt = gt.clamp(0, 1)
    depth = 2.0 + 1.5 * torch.sin(4 * xx) * torch.cos(3 * yy) + 0.05 * torch.randn(H, W, generator=g)
    sil = (0.6 + 0.6 * torch.rand(H, W, generator=g)).clamp(max=1.0)
    sil[:, : max(1, W // 10)] = 0.3 * torch.rand(H, max(1, W // 10), generator=g)      # unobserved band
    depth_sq = depth ** 2 + 0.02 * torch.rand(H, W, generator=g)
    nan = torch.rand(H, W, generator=g) < nan_frac
    depth = torch.where(nan, torch.full_like(depth, float("nan")), depth)
    gt_depth = (depth.nan_to_num(2.0) + 0.1 * torch.randn(H, W, generator=g)).clamp(min=0.2)
    gt_depth = torch.where(torch.rand(H, W, generator=g) < 0.1, torch.zeros_like(gt_depth), gt_depth)   # invalid pixels
    est_depth = 0.5 * gt_depth + 0.3 + 0.05 * torch.randn(H, W, generator=g)           # monocular: affine, noisy
    return dict(image=image.float().contiguous(), gt_color=gt.float().contiguous(),
                depth_image=torch.stack([depth, sil, depth_sq]).float().contiguous(),
                gt_depth=gt_depth.float().contiguous(), est_depth=est_depth.float().contiguous())
