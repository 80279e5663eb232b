/*
 * gsloss_b200.h — C-ABI of the steps either side of the rasterizer in an MM3DGS-SLAM optimisation
 * iteration, exported by the same libgsrast_b200.so:
 *
 *   gsr_slam_loss   the per-iteration image losses of the mapper and the tracker, value AND gradient
 *                   w.r.t. the two rendered images in three launches, no host synchronisation:
 *        <- l1_loss / ssim / _ssim / pearson_loss        R/utils/loss_utils.py:43-68,114-154
 *           as composed by Mapper.optimize_map           R/slam/mapper.py:832-887
 *           and Tracker.optimize_cam                     R/slam/tracker.py:104-144
 *   gsr_adam_step   one torch.optim.Adam step over a flat parameter / gradient bucket with per-segment
 *                   learning rates (the reference's optimizer: R/slam/gaussian_model.py:151-189,
 *                   torch.optim.Adam(l, lr=0.0, eps=1e-15), stepped at R/slam/mapper.py:938)
 *
 * (R = /root/reference.)  Same conventions as gsrast_b200.h: raw fp32 device pointers, planar [C,H,W]
 * images, work enqueued on the caller's stream, 0 / negative gsr_status return, gsr_last_error().
 */
#ifndef GSLOSS_B200_H_
#define GSLOSS_B200_H_

#include "gsrast_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* How the colour term is formed (image, gt_color are [3,H,W]). */
enum gsr_color_loss {
    GSR_COLOR_NONE = 0,
    GSR_COLOR_L1_SSIM = 1,        /* (1-l)*mean|I-G| + l*(1 - mean SSIM(I,G)); unmasked   mapper.py:856-860,863-865 */
    GSR_COLOR_MASKED_L1_MEAN = 2, /* mean of |I-G| over [:, mask]                         tracker.py:129 */
    GSR_COLOR_MASKED_L1_SUM = 3   /* sum of |G-I| over the mask tiled to 3 channels        tracker.py:123-125 */
};

/* How the depth term is formed.  x = depth_image[0] (rendered depth), y = depth_target ([H,W]). */
enum gsr_depth_loss {
    GSR_DEPTH_NONE = 0,
    GSR_DEPTH_L1_MEAN = 1,     /* mean |y - x| over the mask                              mapper.py:853 */
    GSR_DEPTH_L1_SUM = 2,      /* sum  |y - x| over the mask                              tracker.py:121 */
    GSR_DEPTH_PEARSON = 3,     /* 1 - corr(x, y) over the mask                            loss_utils.py:60 */
    GSR_DEPTH_PEARSON_INV = 4, /* min(1 - corr(-y, x), 1 - corr(1/(y+200), x))            loss_utils.py:54-58 */
    GSR_DEPTH_PEARSON_COLS = 5 /* mean over the W image columns c of 1 - corr(x[:, c], y[:, c]): what the reference's
                                * UNMASKED call computes, because torchmetrics reads a 2-D [H,W] input as H samples of W
                                * outputs (loss_utils.py:52-53,60 with mask=None, mapper.py:862-868).  depth_mask must be 0 */
};

/* Mask terms, AND-ed; a zero flag set means "all pixels".  The masks carry no gradient (the reference
 * detaches them). */
enum gsr_mask_flags {
    GSR_MASK_GT_DEPTH_POS = 1, /* gt_depth > 0                                            mapper.py:849, tracker.py:115 */
    GSR_MASK_NOT_NAN = 2,      /* !isnan(depth) && !isnan(depth_sq - depth^2)             mapper.py:848 */
    GSR_MASK_SILHOUETTE = 4    /* depth_image[1] > sil_threshold                          tracker.py:107 (0.99) */
};

typedef struct gsr_loss_config {
    int32_t width, height;
    int32_t color_mode;        /* gsr_color_loss */
    int32_t depth_mode;        /* gsr_depth_loss */
    int32_t color_mask;        /* gsr_mask_flags, used by the MASKED colour modes */
    int32_t depth_mask;        /* gsr_mask_flags */
    float lambda_dssim;        /* l of GSR_COLOR_L1_SSIM */
    float sil_threshold;
    float color_weight;        /* total = color_weight * colour term + depth_weight * depth term */
    float depth_weight;
    float grad_scale;          /* the gradients written are grad_scale * d total / d image */
    int32_t _pad;
} gsr_loss_config;

/* Scratch the loss needs (3 derivative maps of [3,H,W] for the SSIM backward + per-CTA partial sums). */
size_t gsr_slam_loss_ws_bytes(int32_t width, int32_t height);

/* image        [3,H,W] rendered colour                  depth_image  [3,H,W] rendered (depth, silhouette, depth^2),
 * gt_color     [3,H,W]                                               NULL if no mask / depth term needs it
 * depth_target [H,W]   y of the depth term (gt or estimated depth), NULL with GSR_DEPTH_NONE
 * gt_depth     [H,W]   only for GSR_MASK_GT_DEPTH_POS (may alias depth_target)
 * losses       [4]     device: total, colour term, depth term, mean SSIM (0 when not computed)
 * dL_dimage    [3,H,W] written (not accumulated); NULL to skip all gradient work
 * dL_ddepth_image [3,H,W] written (channels 1 and 2 are zero); NULL when depth_mode is NONE */
int gsr_slam_loss(gsr_stream_t stream, const gsr_loss_config* cfg, const float* image, const float* depth_image,
                  const float* gt_color, const float* depth_target, const float* gt_depth, void* ws, size_t ws_bytes,
                  float* losses, float* dL_dimage, float* dL_ddepth_image);

/* One Adam step over flat buffers of n floats: params, grads, exp_avg, exp_avg_sq (torch.optim.Adam,
 * amsgrad=False, weight_decay=0, maximize=False).  The bucket is divided into num_segments consecutive
 * segments; segment s covers [seg_end[s-1], seg_end[s]) and uses learning rate seg_lr[s] (both HOST arrays,
 * at most GSR_ADAM_MAX_SEGMENTS).  `step` is the 1-based step count used for the bias corrections.
 * Hyper-parameters are doubles because torch derives 1 - beta, the bias corrections and lr / bias_correction1 from
 * Python floats before rounding to fp32 (1 - 0.999f would already differ from 1 - 0.999 by 1.3e-5 relative).
 * grad_scale multiplies the gradient on read (e.g. 1/K to turn the summed keyframe gradients into a mean).
 * zero_grads != 0 clears the gradient bucket in the same pass (optimizer.zero_grad). */
#define GSR_ADAM_MAX_SEGMENTS 16
int gsr_adam_step(gsr_stream_t stream, float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                  int32_t num_segments, const int64_t* seg_end, const double* seg_lr, double beta1, double beta2, double eps,
                  int64_t step, float grad_scale, int32_t zero_grads);

/* Map surgery over the flat parameter / optimizer buffers: the reference's _prune_optimizer and
 * cat_tensors_to_optimizer (R/slam/gaussian_model.py:380-399, :418-451; called from prune_points :401-416 and
 * densification_postfix :453-485).  A flat buffer is ngroups slabs [P, widths[g]] fp32 one after the other (the
 * optimizer-group order); the parameters and both Adam moments are nbuf parallel buffers with that layout.
 *
 * gsr_compact_scan:   new_index[i] = number of rows j < i with keep[j] != 0 (keep == NULL keeps every row);
 *                     *count (device) = number of kept rows.  ws: gsr_compact_ws_bytes(P) of device scratch.
 * gsr_compact_gather: for every kept row i, every group g and every buffer b:
 *                     dst[b][group g of a rows_out-row layout][new_index[i]][:] = src[b][group g of a P-row layout][i][:].
 *                     rows_out >= kept rows; rows past the kept ones (the ones an extension appends) are left untouched.
 *                     With keep == NULL and new_index == NULL it re-lays the P rows out for rows_out rows (extension).
 * Exact copies: the result equals torch's boolean-mask indexing / torch.cat bit for bit. */
#define GSR_COMPACT_MAX_GROUPS 16
#define GSR_COMPACT_MAX_BUFFERS 4
size_t gsr_compact_ws_bytes(int64_t P);
int gsr_compact_scan(gsr_stream_t stream, int64_t P, const uint8_t* keep, uint32_t* new_index, int64_t* count, void* ws,
                     size_t ws_bytes);
int gsr_compact_gather(gsr_stream_t stream, int64_t P, int64_t rows_out, int32_t ngroups, const int32_t* widths,
                       const uint8_t* keep, const uint32_t* new_index, int32_t nbuf, const float* const* src,
                       float* const* dst);

#ifdef __cplusplus
}
#endif
#endif /* GSLOSS_B200_H_ */
