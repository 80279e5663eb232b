#!/usr/bin/env python
"""Generate golden vectors for the image losses by importing the UNMODIFIED reference module
R/utils/loss_utils.py in the build container (CPU torch):

    python tests/golden/make_loss_golden.py            # writes tests/golden/loss_*.pt

`torchmetrics` (imported at R/utils/loss_utils.py:16 for pearson_corrcoef) is not installed here, so the import is
satisfied with a stub and only the reference's own functions l1_loss / ssim are exercised: value and autograd
gradient of  (1 - l) * l1_loss(image, gt) + l * (1 - ssim(image, gt))  (R/slam/mapper.py:856-860) and of the masked
tracker term l1 over [:, mask] (R/slam/tracker.py:129).  Inputs are seed-addressed (gsr_synth.make_loss_inputs), so
a fixture stores the case description and the reference's outputs only.
tests/test_oracle_golden.py replays the cases through oracle/loss_oracle.py.
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("GSR_REFERENCE", "/root/reference")
for p in (ROOT, os.path.join(ROOT, "mm3dgs-slam_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

CASES = {
    # name: W, H, seed, lambda_dssim
    "loss_ragged_80x56": dict(W=80, H=56, seed=3, lam=0.2),
    "loss_small_37x21": dict(W=37, H=21, seed=5, lam=0.2),
    "loss_160x120_lam05": dict(W=160, H=120, seed=7, lam=0.5),
}


def import_reference_losses():
    tm = types.ModuleType("torchmetrics")
    tmf = types.ModuleType("torchmetrics.functional")
    tmr = types.ModuleType("torchmetrics.functional.regression")

    def _absent(*a, **k):
        raise RuntimeError("torchmetrics is not installed; pearson is not part of the golden vectors")
    tmr.pearson_corrcoef = _absent
    sys.modules.update({"torchmetrics": tm, "torchmetrics.functional": tmf, "torchmetrics.functional.regression": tmr})
    sys.path.insert(0, REF)
    import importlib
    return importlib.import_module("utils.loss_utils")


def main():
    import gsr_synth as S
    L = import_reference_losses()
    out = os.path.dirname(os.path.abspath(__file__))
    for name, c in CASES.items():
        d = S.make_loss_inputs(c["W"], c["H"], c["seed"])
        img = d["image"].clone().requires_grad_(True)
        l1 = L.l1_loss(img, d["gt_color"])
        s = L.ssim(img, d["gt_color"])
        loss = (1 - c["lam"]) * l1 + c["lam"] * (1.0 - s)
        loss.backward()
        mask = d["depth_image"][1] > 0.99
        img2 = d["image"].clone().requires_grad_(True)
        lm = L.l1_loss(img2, d["gt_color"], mask)
        lm.backward()
        torch.save(dict(case=c, l1=l1.detach(), ssim=s.detach(), loss=loss.detach(), dL_dimage=img.grad.clone(),
                        masked_l1=lm.detach(), masked_dL_dimage=img2.grad.clone()), os.path.join(out, name + ".pt"))
        print(name, float(l1), float(s), float(loss), float(lm))


if __name__ == "__main__":
    main()
