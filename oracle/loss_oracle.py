"""TEST INFRASTRUCTURE — restatement of the reference's image losses in plain torch ops (differentiable
by autograd, runs on CPU or on a CUDA device), used only as the checker for gsr_slam_loss
(tests/, __graft_entry__.smoke()).  Never imported by the product; bench.py restates the composition it times inline.

Follows (R = /root/reference):
  l1_loss                   R/utils/loss_utils.py:64-68
  window_1d                 R/utils/loss_utils.py:98-112  (gaussian / create_window)
  ssim_map / ssim           R/utils/loss_utils.py:114-154 (evaluated separably: the reference's 2-D window is an outer
                            product, so two 1-D passes give the same moments up to float rounding)
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
import torch
import torch.nn.functional as F


def l1_loss(rendered, target, mask=None):
    """Mean absolute difference, optionally over the pixels selected by a [H,W] mask in every channel."""
    diff = (rendered - target).abs()
    return diff.mean() if mask is None else diff[:, mask].mean()


def window_1d(size=11, sigma=1.5):
    """The reference's 11-tap window: exp(-(x - 5)^2 / (2 sigma^2)) evaluated in double, stored as float32 and divided
    by its float32 sum; its 2-D window is the outer product of this vector with itself."""
    x = torch.arange(size, dtype=torch.float64) - size // 2
    g = torch.exp(-(x * x) / (2.0 * sigma * sigma)).to(torch.float32)
    return g / g.sum()


def _blur(x, w):
    """Zero-padded 'same' filtering of every channel with the separable window w (x) w."""
    lead = x.dim() == 3
    x4 = x.unsqueeze(0) if lead else x
    c, r = x4.shape[1], w.numel() // 2
    k = w.to(device=x4.device, dtype=x4.dtype)
    y = F.conv2d(x4, k.view(1, 1, 1, -1).repeat(c, 1, 1, 1), padding=(0, r), groups=c)
    y = F.conv2d(y, k.view(1, 1, -1, 1).repeat(c, 1, 1, 1), padding=(r, 0), groups=c)
    return y.squeeze(0) if lead else y


def ssim_map(a, b, size=11):
    """Per-pixel structural similarity of two images from their windowed first and second moments."""
    w = window_1d(size)
    m_a, m_b = _blur(a, w), _blur(b, w)
    var_a = _blur(a * a, w) - m_a * m_a
    var_b = _blur(b * b, w) - m_b * m_b
    cov = _blur(a * b, w) - m_a * m_b
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    return ((2 * m_a * m_b + c1) * (2 * cov + c2)) / ((m_a * m_a + m_b * m_b + c1) * (var_a + var_b + c2))


def ssim(a, b, window_size=11):
    return ssim_map(a, b, window_size).mean()


def pearson_corrcoef(preds, target):
    """Published definition (see the module docstring).  1-D inputs: one coefficient.  2-D inputs [N, d]: torchmetrics
    treats dim 0 as the N samples and dim 1 as d independent outputs and returns d coefficients — which is what the
    reference's UNMASKED call gets, because it passes the [H, W] depth images as they are (R/utils/loss_utils.py:52-53,
    called with mask=None from R/slam/mapper.py:862-868): one coefficient per image column."""
    n = preds.shape[0]
    mx, my = preds.mean(0), target.mean(0)
    var_x = ((preds - mx) ** 2).sum(0) / (n - 1)
    var_y = ((target - my) ** 2).sum(0) / (n - 1)
    cov = ((preds - mx) * (target - my)).sum(0) / (n - 1)
    return torch.clamp(cov / (var_x * var_y).sqrt(), -1.0, 1.0)


def pearson_loss(render, estimate, mask=None, invert_estimate=True):
    """mask=None keeps the 2-D images (per-column coefficients, averaged by the trailing .mean()), like the reference."""
    r = render[mask] if mask is not None else render
    e = estimate[mask] if mask is not None else estimate
    if invert_estimate:
        a = (1 - pearson_corrcoef(-e, r)).mean()
        b = (1 - pearson_corrcoef(1 / (e + 200.0), r)).mean()
        return min(a, b)
    return (1 - pearson_corrcoef(e, r)).mean()


# ---- the C-ABI's configuration space, restated with the functions above ------------------------------------
COLOR_NONE, COLOR_L1_SSIM, COLOR_MASKED_L1_MEAN, COLOR_MASKED_L1_SUM = 0, 1, 2, 3
DEPTH_NONE, DEPTH_L1_MEAN, DEPTH_L1_SUM, DEPTH_PEARSON, DEPTH_PEARSON_INV, DEPTH_PEARSON_COLS = 0, 1, 2, 3, 4, 5
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
        elif dm == DEPTH_PEARSON_COLS:     # the reference's mask=None call: 2-D images, one coefficient per column
            depth = pearson_loss(x, depth_target, mask=None, invert_estimate=False)
        else:
            depth = pearson_loss(x, depth_target, mask=mask, invert_estimate=(dm == DEPTH_PEARSON_INV))
    total = cfg.get("color_weight", 1.0) * color + cfg.get("depth_weight", 1.0) * depth
    return total, color, depth


# The reference's four loss compositions as configurations (lambda / weights from R/configs/*.yml).
def mapper_splatam(lambda_dssim=0.2):          # R/slam/mapper.py:839-860
    return dict(color_mode=COLOR_L1_SSIM, lambda_dssim=lambda_dssim, depth_mode=DEPTH_L1_MEAN,
                depth_mask=MASK_GT_DEPTH_POS | MASK_NOT_NAN, color_weight=0.5, depth_weight=1.0)


def mapper_default(lambda_dssim=0.2, pearson_weight=0.05, use_gt_depth=False):   # R/slam/mapper.py:862-885
    return dict(color_mode=COLOR_L1_SSIM, lambda_dssim=lambda_dssim,
                depth_mode=DEPTH_PEARSON if use_gt_depth else DEPTH_PEARSON_COLS,
                depth_mask=MASK_GT_DEPTH_POS if use_gt_depth else 0, color_weight=1.0, depth_weight=pearson_weight)


def tracker_splatam():                         # R/slam/tracker.py:110-126
    m = MASK_GT_DEPTH_POS | MASK_NOT_NAN | MASK_SILHOUETTE
    return dict(color_mode=COLOR_MASKED_L1_SUM, color_mask=m, depth_mode=DEPTH_L1_SUM, depth_mask=m,
                sil_threshold=0.99, color_weight=0.5, depth_weight=1.0)


def tracker_default(pearson_weight=0.05, use_gt_depth=False):                    # R/slam/tracker.py:127-144
    return dict(color_mode=COLOR_MASKED_L1_MEAN, color_mask=MASK_SILHOUETTE, depth_mode=DEPTH_PEARSON_INV,
                depth_mask=MASK_SILHOUETTE | (MASK_GT_DEPTH_POS if use_gt_depth else 0), sil_threshold=0.99,
                color_weight=1.0, depth_weight=pearson_weight)
