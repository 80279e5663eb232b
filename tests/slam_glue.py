"""Restatement (for tests) of what the reference's Renderer.render() does around the rasterizer in the
`transform_means_python` mode both shipped configs use (R/slam/renderer.py:85-224, R/configs/TUM.yml:28):
means are moved to the camera frame in PyTorch (pose gradient via autograd), the rasterizer gets an
identity view matrix, and EVERY render is two rasterizer calls sharing one `means2D` leaf — the RGB pass
(SH colours) and the depth/silhouette pass with colors_precomp = [z, 1, z^2] (renderer.py:26-43,196-214)."""
import torch

import gsr_synth as S


def camera_frame(params, w2c):
    xyz1 = torch.cat([params["means3D"], torch.ones_like(params["means3D"][:, :1])], 1)
    return (w2c @ xyz1.T).T[:, :3]


def depth_silhouette(means_cam):
    z = means_cam[:, 2:3]
    return torch.cat([z, torch.ones_like(z), z * z], 1)


def settings(settings_cls, W, H, bg, sh_degree, device):
    cam = S.make_camera(W, H)            # identity view; projmatrix = P^T
    return settings_cls(image_height=H, image_width=W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=bg.to(device),
                        scale_modifier=1.0, viewmatrix=torch.eye(4, device=device),
                        projmatrix=cam.projmatrix.to(device), sh_degree=sh_degree,
                        campos=torch.zeros(3, device=device), prefiltered=False, debug=False)


def render_two_pass(rasterize, rs, params, w2c):
    """rasterize(means3D, means2D, opacities, rs, **kw) -> (image, radii).  Returns rgb, depth, radii, means2D."""
    means_cam = camera_frame(params, w2c)
    means2D = torch.zeros_like(means_cam, requires_grad=True)
    rgb, radii = rasterize(means_cam, means2D, params["opacities"], rs, shs=params["shs"], scales=params["scales"],
                           rotations=params["rotations"])
    depth, _ = rasterize(means_cam, means2D, params["opacities"], rs, colors_precomp=depth_silhouette(means_cam),
                         scales=params["scales"], rotations=params["rotations"])
    return rgb, depth, radii, means2D
