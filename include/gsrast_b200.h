/*
 * gsrast_b200.h — C-ABI of the B200-native differentiable 3D-Gaussian rasterizer.
 *
 * This is the drop-in boundary for the hot path of VITA-Group/MM3DGS-SLAM: every entry point
 * below replaces one native entry of the reference's `diff_gaussian_rasterization._C`
 * extension (DGR = /root/reference/submodules/diff-gaussian-rasterization):
 *
 *   gsr_forward_preprocess + gsr_forward_render
 *        <- CudaRasterizer::Rasterizer::forward      DGR/cuda_rasterizer/rasterizer.h:35-58,
 *           bound as _C.rasterize_gaussians           DGR/ext.cpp:16, DGR/rasterize_points.cu:35-115
 *   gsr_backward
 *        <- CudaRasterizer::Rasterizer::backward     DGR/cuda_rasterizer/rasterizer.h:60-84,
 *           bound as _C.rasterize_gaussians_backward  DGR/ext.cpp:17, DGR/rasterize_points.cu:117-196
 *   gsr_mark_visible
 *        <- CudaRasterizer::Rasterizer::markVisible  DGR/cuda_rasterizer/rasterizer.h:27-33,
 *           bound as _C.mark_visible                  DGR/ext.cpp:18, DGR/rasterize_points.cu:198-217
 *   gsr_*_ws_bytes
 *        <- required<GeometryState/ImageState/BinningState>()  DGR/cuda_rasterizer/rasterizer_impl.h:67-73
 *
 * Plain C: raw device pointers + sizes, no torch / C++ types.  The caller owns every buffer
 * (including the three opaque scratch workspaces that carry state from forward to backward,
 * as in the reference); the library keeps no state between calls.  All work is enqueued on
 * the caller's CUDA stream (the reference uses the legacy default stream; DGR/cuda_rasterizer/
 * rasterizer_impl.cu:148,289,314).
 *
 * Conventions shared with the reference:
 *   - all arrays are fp32, contiguous, row-major; a NULL pointer means "input absent"
 *     (the reference's empty-tensor convention, DGR/diff_gaussian_rasterization/__init__.py:197-207);
 *   - viewmatrix / projmatrix are 16 floats indexed column-major: x' = m[0]x + m[4]y + m[8]z + m[12]
 *     (DGR/cuda_rasterizer/auxiliary.h:58-77);
 *   - out_color is planar [3,H,W]; radii is int32 [P]; tiles are 16x16 pixels.
 *
 * Every function returns 0 on success or a negative gsr_status; gsr_last_error() gives the message
 * for the calling thread.  There is no CPU fallback: without a CUDA device every launch fails.
 */
#ifndef GSRAST_B200_H_
#define GSRAST_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSR_ABI_VERSION 6

typedef void* gsr_stream_t; /* cudaStream_t */

enum gsr_status {
    GSR_OK = 0,
    GSR_ERR_INVALID = -1, /* bad argument (NULL required pointer, negative size, ...) */
    GSR_ERR_CUDA = -2,    /* CUDA runtime error; message in gsr_last_error() */
    GSR_ERR_WORKSPACE = -3 /* workspace too small */
};

/* Gaussian scene inputs (reference: arguments of Rasterizer::forward, rasterizer.h:39-51). */
typedef struct gsr_gaussians {
    int32_t P;                 /* number of Gaussians */
    int32_t sh_degree;         /* active SH degree D (0..3) */
    int32_t sh_coeffs;         /* M = coefficients per Gaussian in `shs` (0 if absent) */
    int32_t raw_params;        /* 0: opacities / scales / rotations hold the values the rasterizer works with (the
                                * reference's convention).  1 (extension): they hold the RAW optimizer
                                * parameters and the library applies the reference's activations itself — opacity =
                                * sigmoid(x), scale = exp(s), rotation = q / max(|q|, 1e-12) (R/slam/gaussian_model.py:
                                * 108-132, torch.sigmoid / torch.exp / F.normalize) — and gsr_backward returns the
                                * gradients w.r.t. the raw parameters (dL_dopacity_raw, dL_dscales, dL_drotations).
                                * Requires scales + rotations (not cov3D_precomp). */
    const float* means3D;      /* [P,3] */
    const float* shs;          /* [P,M,3] or NULL */
    const float* colors_precomp; /* [P,3] or NULL (exactly one of shs / colors_precomp) */
    const float* opacities;    /* [P] */
    const float* scales;       /* [P,3] or NULL */
    const float* rotations;    /* [P,4] (r,x,y,z; NOT normalised by the kernel) or NULL */
    const float* cov3D_precomp; /* [P,6] or NULL (exactly one of scales+rotations / cov3D_precomp) */
    float scale_modifier;
    int32_t extra_mode;        /* 0: extra colours (if any) come from extra_colors; 1: the library GENERATES the SLAM
                                * depth/silhouette colours (z, 1, z^2) from the view-space depth itself (extra_colors
                                * must be NULL) and chains dL/dz into dL_dmeans3D / dL_dviewmatrix in the backward —
                                * no per-Gaussian host-side op at all (R/slam/renderer.py:26-43) */
    /* Extension: [P,3] extra per-Gaussian colours blended in the SAME pass into a second
     * [3,H,W] image (the SLAM renderer's depth / silhouette colours [z, 1, z^2], for which the reference
     * runs the whole rasterizer a second time, R/slam/renderer.py:207-214).  NULL = absent. */
    const float* extra_colors;
} gsr_gaussians;

/* Camera + raster settings (reference: GaussianRasterizationSettings, __init__.py:157-169). */
typedef struct gsr_camera {
    int32_t width, height;
    float tanfovx, tanfovy;
    const float* viewmatrix;   /* 16 floats, device */
    const float* projmatrix;   /* 16 floats, device */
    const float* campos;       /* 3 floats, device */
    const float* background;   /* 3 floats, device */
    int32_t prefiltered;       /* if set, a culled Gaussian traps the kernel (auxiliary.h:154-161) */
    int32_t debug;             /* if set, synchronise + check after every stage (auxiliary.h:166-173) */
} gsr_camera;

/* Gradient outputs of gsr_backward.  Shapes as the reference returns them
 * (DGR/rasterize_points.cu:151-159).  Buffers marked (acc) must be zero-filled by the caller:
 * the blend backward accumulates into them; the others are fully written by the kernels
 * (zeros for culled Gaussians), so the caller may leave them uninitialised. */
typedef struct gsr_grads {
    float* dL_dmeans2D;   /* [P,3] (acc)  NDC-scaled screen gradient, z always 0 */
    float* dL_dconic;     /* [P,4] (acc)  internal: d/d(conic) in .x .y .w */
    float* dL_dopacity;   /* [P]   (acc) */
    float* dL_dcolors;    /* [P,3] (acc)  d/d(colors_precomp) or internal d/d(rgb) on the SH path */
    float* dL_dmeans3D;   /* [P,3] */
    float* dL_dcov3D;     /* [P,6]; may be NULL when cov3D_precomp is absent (not written then) */
    float* dL_dsh;        /* [P,M,3] or NULL when M == 0 */
    float* dL_dscales;    /* [P,3] or NULL when scales absent */
    float* dL_drotations; /* [P,4] or NULL when rotations absent */
    /* Extension over the reference: gradients w.r.t. the camera.  Each may be
     * NULL (skipped).  (acc): zero-filled by the caller. */
    float* dL_dviewmatrix; /* [16] (acc) same column-major indexing as the input */
    float* dL_dprojmatrix; /* [16] (acc) */
    float* dL_dcampos;     /* [3]  (acc) */
    /* If non-zero, dL_dmeans3D / dL_dcov3D / dL_dsh / dL_dscales / dL_drotations are ADDED to the
     * buffers' current contents instead of overwriting them (culled Gaussians add nothing).  Lets a
     * caller point them straight at the parameters' gradient accumulators (e.g. the flat bucket of
     * the keyframe-sharded map step) and skip one read-add-write pass per parameter per frame. */
    int32_t accumulate;
    int32_t _pad;
    float* dL_dextra;      /* [P,3] (acc) gradient w.r.t. the extra colours; required iff extra colours are used
                            * (with extra_mode 1 it is scratch: the chain to the means is applied in the library) */
    float* dL_dopacity_raw; /* [P] required iff raw_params: gradient w.r.t. the raw opacity (written, or added with
                             * `accumulate`); dL_dopacity is then scratch for the blend backward (acc) */
} gsr_grads;

int gsr_abi_version(void);
const char* gsr_last_error(void);

/* Workspace sizes in bytes.  geom: per-Gaussian state; img: per-pixel + per-tile state;
 * binning: the per-tile, depth-ordered Gaussian-id list of the R tile instances (R = num_rendered). */
size_t gsr_geom_ws_bytes(int32_t P, int32_t width, int32_t height);
size_t gsr_img_ws_bytes(int32_t width, int32_t height);
size_t gsr_binning_ws_bytes(int64_t R);

/* Forward, phase 1: per-Gaussian projection / covariance / SH colour / tile rectangles, the
 * instance count R, and the depth sort of the Gaussians.  Writes radii[P].  If num_rendered is not
 * NULL the call waits (on an event recorded right after the projection kernel, not on the whole
 * stream) and stores R there (host memory) so the caller can size the binning workspace — the one
 * host hand-off the reference also has (rasterizer_impl.cu:280-281).  With num_rendered == NULL nothing
 * is waited for: the caller fetches R itself later from geom_ws + gsr_geom_layout.counters (e.g. with an
 * asynchronous copy + event), which lets it enqueue phase 1 of many frames before the first hand-off. */
int gsr_forward_preprocess(gsr_stream_t stream, const gsr_gaussians* g, const gsr_camera* cam,
                           int32_t* radii, void* geom_ws, size_t geom_ws_bytes,
                           void* img_ws, size_t img_ws_bytes, int32_t* num_rendered);

/* Forward, phase 2: stable partition of the tile instances by tile in depth order (== the
 * reference's (tile|depth)-sorted list), tile ranges, per-tile front-to-back alpha compositing.
 * R is the CAPACITY of binning_ws in tile instances: either the value phase 1 produced, or — to keep the host from
 * ever waiting on the GPU — any upper estimate chosen before phase 1 has finished (e.g. the previous frame's count
 * plus a margin).  The kernels read the true count from the device; if it exceeds R they leave the outputs untouched
 * (no out-of-bounds access), and the caller, who learns the true count from geom_ws + gsr_geom_layout.counters
 * afterwards, repeats this call with a large enough workspace.  gsr_backward must be given the same R. */
int gsr_forward_render(gsr_stream_t stream, const gsr_gaussians* g, const gsr_camera* cam,
                       const int32_t* radii, int64_t R,
                       void* geom_ws, void* binning_ws, size_t binning_ws_bytes, void* img_ws,
                       float* out_color, float* out_extra /* [3,H,W]; required iff extra colours are used */);

/* Backward of the whole call. */
int gsr_backward(gsr_stream_t stream, const gsr_gaussians* g, const gsr_camera* cam,
                 const int32_t* radii, int64_t R,
                 const void* geom_ws, const void* binning_ws, const void* img_ws,
                 const float* dL_dpixels /* [3,H,W] */, const float* dL_dpixels_extra /* [3,H,W] or NULL */,
                 const gsr_grads* grads);

/* present[i] = 1 if Gaussian i passes the near-plane test (z_view > 0.1). */
int gsr_mark_visible(gsr_stream_t stream, int32_t P, const float* means3D, const float* viewmatrix,
                     const float* projmatrix, uint8_t* present);

/* Diagnostic (bench / tests; not on the hot path): work counts of the blend stage of a FINISHED forward, from a plain
 * per-pixel replay of the tile lists (SURVEY.md §8d asks for (pixel, Gaussian) pairs/s and the contributing fraction).
 * out (device, 4 x uint64): [0] sum over tiles of 256 * list length — the pairs before early termination;
 * [1] sum over pixels of the position of the last contributor — pairs a per-pixel walk evaluates (the reference's
 * forward, CR/forward.cu:306-357, walks at least these); [2] pairs actually blended (alpha >= 1/255, power <= 0);
 * [3] 32 x the (warp, list entry) pairs the forward recorded — the pairs this library's backward evaluates. */
int gsr_blend_stats(gsr_stream_t stream, int32_t P, int32_t width, int32_t height, int64_t R, const void* geom_ws,
                    const void* binning_ws, const void* img_ws, uint64_t* out);

/* Optional stage profiler (off by default).  When enabled, every stage boundary records a CUDA
 * event on the caller's stream; gsr_profile_collect() synchronises, sums the elapsed time per stage
 * since the previous collect into ms[gsr_profile_num_stages()] / counts[...] and resets.
 * gsr_launch_count(): number of this library's own kernels launched so far in this process. */
void gsr_profile_enable(int on);
int gsr_profile_num_stages(void);
const char* gsr_profile_stage_name(int stage);
int gsr_profile_collect(float* ms, int* counts);
long long gsr_launch_count(void);

/* Introspection for tests (sub-buffers of the opaque workspaces; byte offsets from the base). */
typedef struct gsr_geom_layout {
    size_t rec;        /* float4[3P]: (px,py,depth,cull_r2) (conic.x,conic.y,conic.z,opacity) (r,g,b,clamp bits) */
    size_t rects;      /* ushort4[P]: tile rectangle {x0,y0,x1,y1}, empty for culled Gaussians */
    size_t depth_keys; /* uint32[P]: float bits of the view depth, 0xFFFFFFFF for culled Gaussians */
    size_t sorted_ids; /* uint2[P]: {depth key, id} of the VISIBLE Gaussians in (depth, index) order (counters[2] entries) */
    size_t counters;   /* uint32[>=3]: [0] = R, the number of tile instances; [2] = number of visible Gaussians
                        * (valid after gsr_forward_preprocess) */
    size_t total;
} gsr_geom_layout;
typedef struct gsr_img_layout { size_t final_T, n_contrib, ranges, total; } gsr_img_layout;
typedef struct gsr_binning_layout { size_t point_list, total; } gsr_binning_layout;
/* How the one-kernel tile partition is laid out for a problem size (host-side planning only; tests check that it fits
 * the hardware for every size the library accepts): CTAs launched, Gaussians a CTA's shared memory is sized for,
 * warps per CTA (0 = the tile grid does not fit: gsr_forward_render fails), dynamic shared memory per CTA. */
typedef struct gsr_partition_plan { int32_t ctas, chunk_capacity, warps, _pad; size_t smem_bytes; } gsr_partition_plan;
void gsr_partition_plan_of(int32_t P, int32_t width, int32_t height, gsr_partition_plan* out);
void gsr_geom_layout_of(int32_t P, int32_t width, int32_t height, gsr_geom_layout* out);
void gsr_img_layout_of(int32_t width, int32_t height, gsr_img_layout* out);
void gsr_binning_layout_of(int64_t R, gsr_binning_layout* out);

#ifdef __cplusplus
}
#endif
#endif /* GSRAST_B200_H_ */
