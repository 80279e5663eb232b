#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED compiled reference (oracle/_ref) on a B200.

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'   # then copy *.pt into tests/golden/

Inputs are seed-addressed (gsr_synth.make_scene / orbit_cameras), so each fixture stores only the
case description and the reference's outputs: image, radii, the 8 gradients, num_rendered, the
sorted (key, Gaussian-id) list, tile ranges and per-pixel n_contrib / final_T.
tests/test_oracle_golden.py replays the cases through the CPU oracle.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "mm3dgs-slam_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import gsr_synth as S  # noqa: E402
from oracle import ref_api  # noqa: E402

CASES = {
    # name: P, W, H, sh_degree, mode, seed, camera (None = identity, int = orbit index of 3), bg, scale_modifier
    "c0_sh0_identity": dict(P=300, W=64, H=48, deg=0, mode="sh", seed=1, cam=None, bg=(0.0, 0.0, 0.0), mod=1.0),
    "c1_sh3_orbit_bg": dict(P=1000, W=64, H=48, deg=3, mode="sh", seed=2, cam=1, bg=(0.1, 0.3, 0.6), mod=1.0),
    "c2_colors_ragged": dict(P=800, W=80, H=56, deg=0, mode="colors", seed=11, cam=2, bg=(0.2, 0.5, 0.7), mod=1.0),
    "c3_cov3d_scalemod": dict(P=600, W=64, H=48, deg=1, mode="cov3d", seed=21, cam=None, bg=(0.0, 0.0, 0.0), mod=0.7),
}


def case_inputs(c):
    gs, cam, dL, _ = S.make_scene(c["P"], c["W"], c["H"], c["seed"], c["deg"])
    if c["cam"] is not None:
        cam = S.orbit_cameras(c["W"], c["H"], 3, (0.0, 0.0, 4.0), 0.5)[c["cam"]]
    extra = {}
    if c["mode"] == "colors":
        extra["colors_precomp"] = (gs["shs"][:, 0] * 0.28209479177387814 + 0.5).clamp(min=0).contiguous()
    if c["mode"] == "cov3d":
        from oracle import gs_oracle as O
        extra["cov3D_precomp"] = O.cov3d_from_scale_rot(gs["scales"], c["mod"], gs["rotations"]).contiguous()
    return gs, cam, dL, torch.tensor(c["bg"]), extra


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    dev = torch.device("cuda:0")
    ref_api.load()
    for name, c in CASES.items():
        gs, cam, dL, bg, extra = case_inputs(c)
        d = {k: v.to(dev) for k, v in gs.items()}
        kw = dict(sh_degree=c["deg"], scale_modifier=c["mod"])
        if c["mode"] == "colors":
            kw["colors_precomp"] = extra["colors_precomp"].to(dev)
        else:
            kw["shs"] = d["shs"]
        if c["mode"] == "cov3d":
            kw["cov3D_precomp"] = extra["cov3D_precomp"].to(dev)
        else:
            kw["scales"], kw["rotations"] = d["scales"], d["rotations"]
        view, proj, campos = cam.viewmatrix.to(dev), cam.projmatrix.to(dev), cam.campos.to(dev)
        f = ref_api.forward(d["means3D"], d["opacities"], view, proj, campos, bg.to(dev), c["W"], c["H"], cam.tanfovx,
                            cam.tanfovy, **kw)
        g = ref_api.backward(f, dL.to(dev), d["means3D"], view, proj, campos, bg.to(dev), cam.tanfovx, cam.tanfovy,
                             **kw)
        torch.cuda.synchronize()
        R = f["num_rendered"]
        b = ref_api.decode_binning(f["binning"], R)
        im = ref_api.decode_img(f["img"], c["W"], c["H"])
        T = ((c["W"] + 15) // 16) * ((c["H"] + 15) // 16)
        out = dict(case=c, num_rendered=R, color=f["color"].cpu(), radii=f["radii"].cpu(),
                   keys=b["keys"], point_list=b["point_list"].to(torch.int32), ranges=im["ranges"][:T].to(torch.int32),
                   n_contrib=im["n_contrib"].to(torch.int32), final_T=im["final_T"],
                   grads={k: v.cpu() for k, v in g.items()})
        torch.save(out, os.path.join(outdir, name + ".pt"))
        print(name, "R", R, "visible", int((f["radii"] > 0).sum()))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
