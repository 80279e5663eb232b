"""TEST INFRASTRUCTURE — restatement of the reference's image losses in plain torch ops (differentiable
by autograd, runs on CPU or on a CUDA device), used only as the checker for gsr_slam_loss
(tests/, __graft_entry__.smoke()).  Never imported by the product; bench.py restates the composition it times inline.

Follows (R = /root/reference):
  l1_loss, l2_loss          R/utils/loss_utils.py:64-72
  gaussian / create_window  R/utils/loss_utils.py:98-112
  ssim / _ssim              R/utils/loss_utils.py:114-154
  pearson_loss              R/utils/loss_utils.py:43-61
  mapper composition        R/slam/mapper.py:832-887
  tracker composition       R/slam/tracker.py:104-144

Pinned against the reference's own l1_loss / ssim by tests/golden/make_loss_golden.py (the reference module is
imported in the build container and its outputs are committed as tests/golden/loss_*.pt).
`pearson_corrcoef` comes from torchmetrics (R/utils/loss_utils.py:16), which is neither vendored in the reference
nor installed in this image (the reference pins no version: R/environment.yml lists it without one) — PARITY
UNPINNED for that one function: `pearson_corrcoef` below restates the published definition
(torchmetrics.functional.regression.pearson: mean/var/cov accumulated over the batch, corr = cov / sqrt(var_x var_y),
clamped to [-1, 1]); tests/test_oracle_cpu.py checks it against scipy.stats.pearsonr, an independent implementation of
the same coefficient.
"""
from math import exp

import torch
import torch.nn.functional as F


def l1_loss(network_output, gt, mask=None):
    if mask is None:
        return torch.abs(network_output - gt).mean()
    return torch.abs(network_output - gt)[:, mask].mean()


def gaussian(window_size, sigma):
    g = torch.tensor([exp(-((x - window_size // 2) ** 2) / float(2 * sigma ** 2)) for x in range(window_size)],
                     dtype=torch.float32)
    return g / g.sum()


def create_window(window_size, channel):
    w1 = gaussian(window_size, 1.5).unsqueeze(1)
    w2 = w1.mm(w1.t()).float().unsqueeze(0).unsqueeze(0)
    return w2.expand(channel, 1, window_size, window_size).contiguous()


def ssim(img1, img2, window_size=11, size_average=True):
    channel = img1.size(-3)
    window = create_window(window_size, channel).to(img1.device).type_as(img1)
    pad = window_size // 2
    mu1 = F.conv2d(img1, window, padding=pad, groups=channel)
    mu2 = F.conv2d(img2, window, padding=pad, groups=channel)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq = F.conv2d(img1 * img1, window, padding=pad, groups=channel) - mu1_sq
    sigma2_sq = F.conv2d(img2 * img2, window, padding=pad, groups=channel) - mu2_sq
    sigma12 = F.conv2d(img1 * img2, window, padding=pad, groups=channel) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim_map = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    return ssim_map.mean() if size_average else ssim_map.mean(1).mean(1).mean(1)


def pearson_corrcoef(preds, target):
    """Published definition (see the module docstring); 1-D inputs."""
    n = preds.numel()
    mx, my = preds.mean(), target.mean()
    var_x = ((preds - mx) ** 2).sum() / (n - 1)
    var_y = ((target - my) ** 2).sum() / (n - 1)
    cov = ((preds - mx) * (target - my)).sum() / (n - 1)
    return torch.clamp(cov / (var_x * var_y).sqrt(), -1.0, 1.0)


def pearson_loss(render, estimate, mask=None, invert_estimate=True):
    r = render[mask] if mask is not None else render.reshape(-1)
    e = estimate[mask] if mask is not None else estimate.reshape(-1)
    if invert_estimate:
        a = (1 - pearson_corrcoef(-e, r)).mean()
        b = (1 - pearson_corrcoef(1 / (e + 200.0), r)).mean()
        return min(a, b)
    return (1 - pearson_corrcoef(e, r)).mean()


# ---- the C-ABI's configuration space, restated with the functions above ------------------------------------
COLOR_NONE, COLOR_L1_SSIM, COLOR_MASKED_L1_MEAN, COLOR_MASKED_L1_SUM = 0, 1, 2, 3
DEPTH_NONE, DEPTH_L1_MEAN, DEPTH_L1_SUM, DEPTH_PEARSON, DEPTH_PEARSON_INV = 0, 1, 2, 3, 4
MASK_GT_DEPTH_POS, MASK_NOT_NAN, MASK_SILHOUETTE = 1, 2, 4


def build_mask(flags, depth_image, gt_depth, sil_threshold):
    ref = depth_image if depth_image is not None else gt_depth
    m = torch.ones(ref.shape[-2:], dtype=torch.bool, device=ref.device)
    if flags & MASK_GT_DEPTH_POS:
        m = m & (gt_depth > 0)
    if flags & MASK_NOT_NAN:
        depth = depth_image[0]
        unc = (depth_image[2] - depth ** 2).detach()
        m = m & (~torch.isnan(depth)) & (~torch.isnan(unc))
    if flags & MASK_SILHOUETTE:
        m = m & (depth_image[1] > sil_threshold)
    return m.detach()


def slam_loss(cfg, image, depth_image, gt_color, depth_target, gt_depth):
    """cfg: dict with the fields of gsr_loss_config.  Returns (total, colour term, depth term)."""
    zero = torch.zeros((), dtype=torch.float32, device=(image if image is not None else depth_image).device)
    color, depth = zero, zero
    cm, dm = cfg.get("color_mode", 0), cfg.get("depth_mode", 0)
    thr = cfg.get("sil_threshold", 0.5)
    if cm == COLOR_L1_SSIM:
        lam = cfg["lambda_dssim"]
        color = (1 - lam) * l1_loss(image, gt_color) + lam * (1.0 - ssim(image, gt_color))
    elif cm == COLOR_MASKED_L1_MEAN:
        color = torch.abs(image - gt_color)[:, build_mask(cfg.get("color_mask", 0), depth_image, gt_depth, thr)].mean()
    elif cm == COLOR_MASKED_L1_SUM:
        mask3 = torch.tile(build_mask(cfg.get("color_mask", 0), depth_image, gt_depth, thr), (3, 1, 1))
        color = torch.abs(gt_color - image)[mask3].sum()
    if dm != DEPTH_NONE:
        mask = build_mask(cfg.get("depth_mask", 0), depth_image, gt_depth, thr)
        x = depth_image[0]
        if dm == DEPTH_L1_MEAN:
            depth = torch.abs(depth_target - x)[mask].mean()
        elif dm == DEPTH_L1_SUM:
            depth = torch.abs(depth_target - x)[mask].sum()
        else:
            depth = pearson_loss(x, depth_target, mask=mask, invert_estimate=(dm == DEPTH_PEARSON_INV))
    total = cfg.get("color_weight", 1.0) * color + cfg.get("depth_weight", 1.0) * depth
    return total, color, depth


# The reference's four loss compositions as configurations (lambda / weights from R/configs/*.yml).
def mapper_splatam(lambda_dssim=0.2):          # R/slam/mapper.py:839-860
    return dict(color_mode=COLOR_L1_SSIM, lambda_dssim=lambda_dssim, depth_mode=DEPTH_L1_MEAN,
                depth_mask=MASK_GT_DEPTH_POS | MASK_NOT_NAN, color_weight=0.5, depth_weight=1.0)


def mapper_default(lambda_dssim=0.2, pearson_weight=0.05, use_gt_depth=False):   # R/slam/mapper.py:862-885
    return dict(color_mode=COLOR_L1_SSIM, lambda_dssim=lambda_dssim, depth_mode=DEPTH_PEARSON,
                depth_mask=MASK_GT_DEPTH_POS if use_gt_depth else 0, color_weight=1.0, depth_weight=pearson_weight)


def tracker_splatam():                         # R/slam/tracker.py:110-126
    m = MASK_GT_DEPTH_POS | MASK_NOT_NAN | MASK_SILHOUETTE
    return dict(color_mode=COLOR_MASKED_L1_SUM, color_mask=m, depth_mode=DEPTH_L1_SUM, depth_mask=m,
                sil_threshold=0.99, color_weight=0.5, depth_weight=1.0)


def tracker_default(pearson_weight=0.05, use_gt_depth=False):                    # R/slam/tracker.py:127-144
    return dict(color_mode=COLOR_MASKED_L1_MEAN, color_mask=MASK_SILHOUETTE, depth_mode=DEPTH_PEARSON_INV,
                depth_mask=MASK_SILHOUETTE | (MASK_GT_DEPTH_POS if use_gt_depth else 0), sil_threshold=0.99,
                color_weight=1.0, depth_weight=pearson_weight)
