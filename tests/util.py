"""Shared helpers for the parity tests."""
import torch

import gsr_synth as S


def rel_err(a, b):
    """max |a-b| / max |b|  — the 'within 1e-4 relative' metric of BASELINE.json's north_star."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    if a.numel() == 0 and b.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def rms_rel(a, b):
    """RMS of the element-wise differences over the RMS of the reference tensor.  Next to the max-norm metric above:
    that one is blind to a wrong class of small-magnitude entries (anything well below max|b| passes), this one is
    not dominated by the single largest element."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    if a.numel() == 0 and b.numel() == 0:
        return 0.0
    return float(torch.sqrt(((a - b) ** 2).mean()) / (torch.sqrt((b ** 2).mean()) + 1e-30))


def settings_for(mod, cam, bg, sh_degree, device, scale_modifier=1.0, debug=False):
    return mod.GaussianRasterizationSettings(
        image_height=cam.H, image_width=cam.W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=bg.to(device),
        scale_modifier=scale_modifier, viewmatrix=cam.viewmatrix.to(device), projmatrix=cam.projmatrix.to(device),
        sh_degree=sh_degree, campos=cam.campos.to(device), prefiltered=False, debug=debug)


def scene_on(device, P, W, H, seed=0, sh_degree=0):
    gs, cam, dL, bg = S.make_scene(P, W, H, seed, sh_degree)
    gs = {k: v.to(device) for k, v in gs.items()}
    return gs, cam, dL.to(device), bg.to(device)
