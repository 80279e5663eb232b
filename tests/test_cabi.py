"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/gsrast_b200.h declares, the ctypes struct mirrors have the C layout, and the Python package
exposes the reference's public names.  No kernels are launched."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    syms = set()
    for h in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if h.endswith(".h"):
            src = open(os.path.join(ROOT, "include", h)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            syms |= set(re.findall(r"\b(gsr_[a-z_0-9]+)\s*\(", src))
    return sorted(syms)


def test_header_symbols_exported(built_lib):
    lib = ctypes.CDLL(built_lib)
    syms = _declared_symbols()
    assert len(syms) >= 15 and "gsr_slam_loss" in syms and "gsr_adam_step" in syms
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/*.h but not exported"
    lib.gsr_abi_version.restype = ctypes.c_int
    assert lib.gsr_abi_version() == 6


def test_headers_are_plain_c(tmp_path):
    """include/*.h is the drop-in boundary: it must compile as strict C99 (and as C++) with no torch / CUDA types."""
    import shutil
    import subprocess
    src = tmp_path / "hdr_check.c"
    src.write_text('#include "gsrast_b200.h"\n#include "gsloss_b200.h"\n'
                   "int main(void) { gsr_loss_config c; gsr_gaussians g; gsr_camera k; gsr_grads d;\n"
                   "  (void)c; (void)g; (void)k; (void)d; return GSR_ABI_VERSION == 6 ? 0 : 1; }\n")
    inc = os.path.join(ROOT, "include")
    if shutil.which("gcc"):
        subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", inc, "-fsyntax-only", str(src)],
                       check=True)
    if shutil.which("g++"):
        subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-I", inc, "-fsyntax-only", "-x", "c++", str(src)], check=True)


def test_c_client_links_and_runs(built_lib, tmp_path):
    """A plain C host (tests/c_client.c) links the library and drives the entry points that need no GPU: the
    boundary is usable without Python or torch."""
    import shutil
    import subprocess
    import pytest
    if not shutil.which("gcc"):
        pytest.skip("no C compiler")
    exe = str(tmp_path / "c_client")
    libdir = os.path.dirname(built_lib)
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c_client.c"), "-o", exe, "-L", libdir, "-lgsrast_b200",
                    "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert out.startswith("abi 6 ")


def test_workspace_layouts_are_sane(built_lib):
    """Workspace sizing is host arithmetic: offsets 256-byte aligned and increasing, sizes cover the documented
    per-Gaussian / per-pixel / per-instance arrays, nothing overflows at the largest supported problem sizes."""
    import diff_gaussian_rasterization as dgr
    lib = dgr._lib
    lib.gsr_geom_layout_of.restype = None
    lib.gsr_img_layout_of.restype = None
    lib.gsr_binning_layout_of.restype = None

    class Img(ctypes.Structure):
        _fields_ = [(k, ctypes.c_size_t) for k in ("final_T", "n_contrib", "ranges", "total")]

    class Bin(ctypes.Structure):
        _fields_ = [(k, ctypes.c_size_t) for k in ("point_list", "total")]

    prev = 0
    for P, W, H in [(0, 16, 16), (1, 1, 1), (1000, 320, 240), (1_000_000, 640, 480), (5_000_000, 1920, 1080),
                    (50_000_000, 3840, 2160)]:
        g = dgr._GeomLayout()
        lib.gsr_geom_layout_of(ctypes.c_int32(P), ctypes.c_int32(W), ctypes.c_int32(H), ctypes.byref(g))
        offs = [g.rec, g.rects, g.depth_keys, g.counters, g.sorted_ids]
        assert offs == sorted(offs) and all(o % 256 == 0 for o in offs) and g.total % 256 == 0
        assert g.total == lib.gsr_geom_ws_bytes(P, W, H) >= 48 * P + 8 * P + 4 * P + 16 * P    # record, rect, key, sort buffers
        assert g.rects - g.rec >= 48 * P and g.depth_keys - g.rects >= 8 * P
        assert g.total >= prev
        prev = g.total
        im = Img()
        lib.gsr_img_layout_of(ctypes.c_int32(W), ctypes.c_int32(H), ctypes.byref(im))
        tiles = ((W + 15) // 16) * ((H + 15) // 16)
        assert im.n_contrib - im.final_T >= 4 * W * H and im.ranges - im.n_contrib >= 4 * W * H
        assert im.total == lib.gsr_img_ws_bytes(W, H) >= 8 * W * H + 8 * tiles
    for R in (0, 1, 4_500_000, 3_000_000_000):          # R is 64-bit: 3e9 instances = 15 GB of lists, no wrap-around
        b = Bin()
        lib.gsr_binning_layout_of(ctypes.c_int64(R), ctypes.byref(b))
        assert b.point_list == 0 and b.total == lib.gsr_binning_ws_bytes(R) >= (5 * R if R else 0)    # 4-byte list entry + 1 contribution byte per instance


def test_struct_mirrors_match_c_layout(built_lib):
    import diff_gaussian_rasterization as dgr
    # 4 x int32 + 7 pointers + float + int32
    assert ctypes.sizeof(dgr._Gaussians) == 16 + 7 * 8 + 8 + 8 and dgr._Gaussians.extra_colors.offset == 80
    assert dgr._Gaussians.means3D.offset == 16 and dgr._Gaussians.scale_modifier.offset == 72
    assert ctypes.sizeof(dgr._Camera) == 16 + 4 * 8 + 8
    assert dgr._Camera.viewmatrix.offset == 16 and dgr._Camera.prefiltered.offset == 48
    assert ctypes.sizeof(dgr._Grads) == 12 * 8 + 8 + 8 + 8 and dgr._Grads.accumulate.offset == 96 and dgr._Grads.dL_dextra.offset == 104
    assert dgr._Grads.dL_dopacity_raw.offset == 112 and dgr._Gaussians.raw_params.offset == 12
    import gsr_slam_ops as ops
    assert ctypes.sizeof(ops._LossConfig) == 48 and ops._LossConfig.lambda_dssim.offset == 24   # 6 x int32, 5 x float, pad


def test_slam_ops_reject_cpu_tensors_and_bad_arguments(built_lib):
    """No CPU path for the loss / optimizer either; argument checks answer before any launch."""
    import pytest
    import torch

    import gsr_slam_ops as ops
    img = torch.rand(3, 8, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.slam_loss(ops.mapper_splatam(), img, img, img, img[0], img[0])
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.FlatAdam({"a": torch.zeros(4)}, {"a": 1e-3}).step(torch.zeros(4))      # bookkeeping may live anywhere, the step not
    lib = ctypes.CDLL(built_lib)
    lib.gsr_slam_loss_ws_bytes.restype = ctypes.c_size_t
    assert lib.gsr_slam_loss_ws_bytes(640, 480) >= 9 * 640 * 480 * 4
    assert lib.gsr_slam_loss_ws_bytes(0, 480) == 0
    lib.gsr_last_error.restype = ctypes.c_char_p
    assert lib.gsr_slam_loss(None, None, None, None, None, None, None, None, 0, None, None, None) == -1
    assert b"null" in lib.gsr_last_error()
    lib.gsr_adam_step.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                                          ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int64,
                                                          ctypes.c_float, ctypes.c_int32]
    assert lib.gsr_adam_step(None, None, None, None, None, 8, 1, None, None, 0.9, 0.999, 1e-15, 0, 1.0, 0) == -1   # step < 1


def test_python_surface_matches_reference(built_lib):
    import inspect

    import diff_gaussian_rasterization as dgr
    assert dgr.GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
        "sh_degree", "campos", "prefiltered", "debug")
    sig = inspect.signature(dgr.GaussianRasterizer.forward)
    assert list(sig.parameters)[:9] == ["self", "means3D", "means2D", "opacities", "shs", "colors_precomp", "scales",
                                        "rotations", "cov3D_precomp"]     # + optional additive extensions
    assert list(inspect.signature(dgr.rasterize_gaussians).parameters)[:9] == [
        "means3D", "means2D", "sh", "colors_precomp", "opacities", "scales", "rotations", "cov3Ds_precomp",
        "raster_settings"]
    assert hasattr(dgr.GaussianRasterizer, "markVisible")


def test_error_behaviour_matches_reference(built_lib):
    """Same exceptions, raised before any device work: plain Exception for the SH / colour and the scale+rotation /
    covariance exclusivity checks (DGR/diff_gaussian_rasterization/__init__.py:191-195), RuntimeError for a mis-shaped
    means3D (DGR/rasterize_points.cu:57-59)."""
    import pytest
    import torch

    import diff_gaussian_rasterization as dgr
    import gsr_synth as S
    gs, cam, _, bg = S.make_scene(16, 32, 32)
    rs = dgr.GaussianRasterizationSettings(32, 32, cam.tanfovx, cam.tanfovy, bg, 1.0, cam.viewmatrix, cam.projmatrix,
                                           0, cam.campos, False, False)
    ras, m2 = dgr.GaussianRasterizer(rs), torch.zeros(16, 3)
    sr = dict(scales=gs["scales"], rotations=gs["rotations"])
    for kw in (dict(shs=gs["shs"], colors_precomp=torch.zeros(16, 3), **sr), dict(**sr)):
        with pytest.raises(Exception, match="SHs or precomputed colors") as ei:
            ras(means3D=gs["means3D"], means2D=m2, opacities=gs["opacities"], **kw)
        assert type(ei.value) is Exception
    for kw in (dict(shs=gs["shs"], scales=gs["scales"]), dict(shs=gs["shs"], cov3D_precomp=torch.zeros(16, 6), **sr)):
        with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance") as ei:
            ras(means3D=gs["means3D"], means2D=m2, opacities=gs["opacities"], **kw)
        assert type(ei.value) is Exception
    with pytest.raises(RuntimeError, match=r"means3D must have dimensions \(num_points, 3\)"):
        ras(means3D=torch.zeros(16, 4), means2D=m2, opacities=gs["opacities"], shs=gs["shs"], **sr)


def test_no_cpu_fallback(built_lib):
    """CPU tensors are rejected loudly, never silently rendered on the host."""
    import pytest
    import torch

    import diff_gaussian_rasterization as dgr
    import gsr_synth as S
    gs, cam, dL, bg = S.make_scene(16, 32, 32)
    rs = dgr.GaussianRasterizationSettings(32, 32, cam.tanfovx, cam.tanfovy, bg, 1.0, cam.viewmatrix, cam.projmatrix,
                                           0, cam.campos, False, False)
    with pytest.raises(RuntimeError, match="CUDA"):
        dgr.GaussianRasterizer(rs)(means3D=gs["means3D"], means2D=torch.zeros(16, 3), opacities=gs["opacities"],
                                   shs=gs["shs"], scales=gs["scales"], rotations=gs["rotations"])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mm3dgs-slam_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                assert "from oracle" not in txt and "import oracle" not in txt, os.path.join(d, f)


def test_flat_adam_surgery_has_no_cpu_path(built_lib):
    """FlatAdam.prune / .extend run the library's compaction kernels (gsr_compact_scan / gsr_compact_gather); like the
    rest of the product they refuse CPU tensors instead of falling back (the parity test against the reference's
    _prune_optimizer / cat_tensors_to_optimizer is tests/test_gpu_slam_ops.py::test_flat_adam_surgery_*)."""
    import pytest
    import torch

    import gsr_slam_ops as ops
    params = {"xyz": torch.randn(10, 3), "opacity": torch.randn(10, 1)}
    opt = ops.FlatAdam(params, {"xyz": 1e-3, "opacity": 1e-3})
    with pytest.raises(RuntimeError):
        opt.prune(torch.ones(10, dtype=torch.bool))
    with pytest.raises(RuntimeError):
        opt.extend({k: v[:2] for k, v in params.items()})


def test_tile_partition_plan_fits_the_hardware(built_lib):
    """gsr_partition_plan_of: for every problem size the library accepts the one-kernel tile partition must fit a B200
    SM (<= 220 KB dynamic shared memory, <= 1024 threads), its chunks must cover the case that every Gaussian is visible,
    a chunk must stay within the 16-bit CTA-relative ranks, and its look-back rows must fit the geometry workspace."""
    lib = ctypes.CDLL(built_lib)

    class Plan(ctypes.Structure):
        _fields_ = [("ctas", ctypes.c_int32), ("chunk_capacity", ctypes.c_int32), ("warps", ctypes.c_int32),
                    ("_pad", ctypes.c_int32), ("smem_bytes", ctypes.c_size_t)]
    lib.gsr_geom_ws_bytes.restype = ctypes.c_size_t
    lib.gsr_geom_ws_bytes.argtypes = [ctypes.c_int32] * 3
    for W, H in ((16, 16), (128, 96), (320, 240), (640, 330), (640, 480), (1920, 1080), (2560, 1440), (3840, 1600), (100, 70)):
        T = ((W + 15) // 16) * ((H + 15) // 16)
        for P in (1, 300, 2048, 4097, 100_000, 1_000_000, 5_000_000, 50_000_000):
            p = Plan()
            lib.gsr_partition_plan_of(ctypes.c_int32(P), ctypes.c_int32(W), ctypes.c_int32(H), ctypes.byref(p))
            assert 1 <= p.warps <= 32, (P, W, H, p.warps)
            assert p.smem_bytes <= 220 * 1024
            assert 1024 <= p.chunk_capacity <= 65535 and p.chunk_capacity % 32 == 0
            assert p.ctas >= 1 and p.ctas * p.chunk_capacity >= P                  # every Gaussian may be visible
            assert p.smem_bytes >= 12 * p.chunk_capacity + 4 * T + 3 * T * p.warps  # rectangles + bases + counters + tags
            assert lib.gsr_geom_ws_bytes(P, W, H) >= 4 * p.ctas * T                # look-back rows live in the geometry ws
    # a tile grid too large for one warp's counters (more than ~26k tiles, e.g. 3840x2160 = 32400) is reported as such
    # (warps == 0 -> gsr_forward_render returns GSR_ERR_INVALID), not mis-planned
    for W, H in ((3840, 2160), (16000, 16000)):
        p = Plan()
        lib.gsr_partition_plan_of(ctypes.c_int32(1000), ctypes.c_int32(W), ctypes.c_int32(H), ctypes.byref(p))
        assert p.warps == 0
