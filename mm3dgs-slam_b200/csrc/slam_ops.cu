// slam_ops.cu — the steps either side of the rasterizer in one optimisation iteration (include/gsloss_b200.h):
//
//   gsr_slam_loss : image losses of the mapper / tracker, value and gradient, three launches
//                   (tile pass -> one-CTA finalize -> tile pass), no host synchronisation.
//                   Follows R/utils/loss_utils.py:43-68 (pearson_loss, l1_loss), :98-154 (gaussian window, ssim, _ssim)
//                   as composed by R/slam/mapper.py:832-887 and R/slam/tracker.py:104-144.
//   gsr_adam_step : torch.optim.Adam over a flat bucket (R/slam/gaussian_model.py:189, R/slam/mapper.py:938).
//
// Both are HBM-streaming, image- or bucket-sized passes; the point of fusing them is launch count
// (the torch composition of L1 + SSIM + masked depth loss is ~90 launches forward + backward).
#include "gsr_internal.cuh"
#include "../../include/gsloss_b200.h"
#include <math.h>

namespace gsr {

constexpr int kLT = 256;                 // threads per loss CTA
constexpr int kTW = 32, kTH = 16;        // output tile of a loss CTA (pixels)
constexpr int kWR = 5;                   // SSIM window radius (11 taps)
constexpr int kPW = kTW + 2 * kWR, kPH = kTH + 2 * kWR;
constexpr int kNPart = 10;               // doubles per CTA partial record
constexpr int kNStat = 16;               // floats of finalized statistics

struct SsimWin { float w[2 * kWR + 1]; };

struct LossArgs {
    gsr_loss_config c;
    const float *image, *dimg, *gt_color, *dtarget, *gt_depth;
    float* dmaps;        // [3 maps][3][H][W]: A = dS/dmu1 (total), B = dS/dE11, C = dS/dE12
    double* partials;    // [CTAs][kNPart]
    float* stats;        // [kNStat]
    float* colstats;     // [W][4] per image column: xbar, ybar, ka, kb (GSR_DEPTH_PEARSON_COLS)
    double* col_loss;    // [W]    per image column: 1 - r
    float* losses;       // [4]
    float *dL_dimage, *dL_ddimg;
    SsimWin win;
};

// partial slots: colour CTAs {0: sum |x-y|, 1: sum SSIM, 2: count}; depth CTAs {0: sum |y-x|, 1: n, 2: Sx, 3: Sy,
// 4: Sxx, 5: Syy, 6: Sxy, 7: Sy2, 8: Sy2y2, 9: Sxy2} with y2 = 1/(y+200) (the second PEARSON_INV variant)
// stats slots:
enum { kStL1 = 0, kStSsim, kStDepthL1, kStXbar, kStYbar, kStKa, kStKb, kStVariant };

__device__ __forceinline__ bool loss_mask(int flags, const LossArgs& a, int pix, int HW)
{
    bool m = true;
    if (flags & GSR_MASK_GT_DEPTH_POS) m = m && (a.gt_depth[pix] > 0.f);
    if (flags & GSR_MASK_NOT_NAN) {
        const float d = a.dimg[pix], d2 = a.dimg[2 * HW + pix];
        const float u = d2 - __fmul_rn(d, d);
        m = m && !isnan(d) && !isnan(u);
    }
    if (flags & GSR_MASK_SILHOUETTE) m = m && (a.dimg[HW + pix] > a.c.sil_threshold);
    return m;
}

__device__ __forceinline__ float sgn(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); }

// Sum acc[0..N) over a CTA of WARPS warps in a fixed order and store to out[0..N).
template <int N, int WARPS>
__device__ __forceinline__ void block_sum_store(const double (&acc)[N], double* out, double (*s_red)[WARPS])
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < N; j++) {
        double v = acc[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_red[j][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < N) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < WARPS; w++) v += s_red[threadIdx.x][w];
        out[threadIdx.x] = v;
    }
}

// ---- pass 1: per-tile sums (+ the three SSIM derivative maps) ------------------------------------------
__global__ void __launch_bounds__(kLT, 5) k_loss_fwd(const LossArgs a)
{
    __shared__ float s_x[kPH][kPW], s_y[kPH][kPW];
    __shared__ float s_h[5][kPH][kTW];
    __shared__ double s_red[kNPart][kLT / 32];
    const int W = a.c.width, H = a.c.height, HW = W * H;
    const int tid = threadIdx.x, z = blockIdx.z;
    const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH;
    const int cta = (z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;

    if (z < 3) {
        // a thread sees at most two pixels: plain floats here, fp64 from the CTA reduction on
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
        const float* X = a.image + (size_t)z * HW;
        const float* Y = a.gt_color + (size_t)z * HW;
        if (a.c.color_mode == GSR_COLOR_L1_SSIM) {
            for (int i = tid; i < kPH * kPW; i += kLT) {
                const int py = i / kPW, px = i - py * kPW;
                const int gx = x0 + px - kWR, gy = y0 + py - kWR;
                const bool in = gx >= 0 && gx < W && gy >= 0 && gy < H;   // conv2d zero padding
                s_x[py][px] = in ? X[gy * W + gx] : 0.f;
                s_y[py][px] = in ? Y[gy * W + gx] : 0.f;
            }
            __syncthreads();
            for (int i = tid; i < kPH * kTW; i += kLT) {
                const int r = i / kTW, c = i - r * kTW;
                float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
                for (int k = 0; k <= 2 * kWR; k++) {
                    const float w = a.win.w[k], x = s_x[r][c + k], y = s_y[r][c + k];
                    m1 += w * x; m2 += w * y; e11 += w * (x * x); e22 += w * (y * y); e12 += w * (x * y);
                }
                s_h[0][r][c] = m1; s_h[1][r][c] = m2; s_h[2][r][c] = e11; s_h[3][r][c] = e22; s_h[4][r][c] = e12;
            }
            __syncthreads();
            // vertical pass: a thread owns one column and two ADJACENT rows, so the 12 staged rows it reads serve both
            static_assert(kTH * kTW == 2 * kLT, "two output pixels per thread");
            const int c = tid & (kTW - 1), r0 = (tid / kTW) * 2;
            float mo[2][5];
#pragma unroll
            for (int j = 0; j < 2; j++)
#pragma unroll
                for (int q = 0; q < 5; q++) mo[j][q] = 0.f;
#pragma unroll
            for (int k = 0; k <= 2 * kWR + 1; k++) {
#pragma unroll
                for (int q = 0; q < 5; q++) {
                    const float h = s_h[q][r0 + k][c];
                    if (k <= 2 * kWR) mo[0][q] += a.win.w[k] * h;
                    if (k >= 1) mo[1][q] += a.win.w[k - 1] * h;
                }
            }
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int r = r0 + j;
                const int gx = x0 + c, gy = y0 + r;
                if (gx >= W || gy >= H) continue;
                const float mu1 = mo[j][0], mu2 = mo[j][1], e11 = mo[j][2], e22 = mo[j][3], e12 = mo[j][4];
                // R/utils/loss_utils.py:126-148
                const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
                const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu1_mu2 = mu1 * mu2;
                const float sig1 = e11 - mu1_sq, sig2 = e22 - mu2_sq, sig12 = e12 - mu1_mu2;
                const float ta = 2.f * mu1_mu2 + C1, tb = 2.f * sig12 + C2;
                const float tc = mu1_sq + mu2_sq + C1, td = sig1 + sig2 + C2;
                const float inv_cd = 1.f / (tc * td);
                const float S = ta * tb * inv_cd;
                const float dS_dsig1 = -S / td;
                const float dS_dsig12 = 2.f * ta * inv_cd;
                const float dS_dmu1 = 2.f * mu2 * tb * inv_cd - 2.f * mu1 * S / tc;
                const int pix = gy * W + gx;
                if (a.dL_dimage) {
                    // S as a function of (mu1, E11, E12): sigma1^2 = E11 - mu1^2, sigma12 = E12 - mu1 mu2
                    a.dmaps[(size_t)(0 * 3 + z) * HW + pix] = dS_dmu1 - 2.f * mu1 * dS_dsig1 - mu2 * dS_dsig12;
                    a.dmaps[(size_t)(1 * 3 + z) * HW + pix] = dS_dsig1;
                    a.dmaps[(size_t)(2 * 3 + z) * HW + pix] = dS_dsig12;
                }
                acc0 += fabsf(s_x[r + kWR][c + kWR] - s_y[r + kWR][c + kWR]);
                acc1 += S;
            }
        } else if (a.c.color_mode != GSR_COLOR_NONE) {
            for (int i = tid; i < kTH * kTW; i += kLT) {
                const int r = i / kTW, c = i - r * kTW;
                const int gx = x0 + c, gy = y0 + r;
                if (gx >= W || gy >= H) continue;
                const int pix = gy * W + gx;
                if (loss_mask(a.c.color_mask, a, pix, HW)) {
                    acc0 += fabsf(X[pix] - Y[pix]);
                    acc2 += 1.f;
                }
            }
        }
        const double acc[3] = {(double)acc0, (double)acc1, (double)acc2};
        block_sum_store<3, kLT / 32>(acc, a.partials + (size_t)cta * kNPart, s_red);
    } else {
        double acc[kNPart];
#pragma unroll
        for (int j = 0; j < kNPart; j++) acc[j] = 0.0;
        if (a.c.depth_mode != GSR_DEPTH_NONE)
        for (int i = tid; i < kTH * kTW; i += kLT) {
            const int r = i / kTW, c = i - r * kTW;
            const int gx = x0 + c, gy = y0 + r;
            if (gx >= W || gy >= H) continue;
            const int pix = gy * W + gx;
            if (!loss_mask(a.c.depth_mask, a, pix, HW)) continue;
            const float xf = a.dimg[pix], yf = a.dtarget[pix];
            const double x = xf, y = yf;
            acc[0] += (double)fabsf(yf - xf);
            acc[1] += 1.0;
            acc[2] += x; acc[3] += y; acc[4] += x * x; acc[5] += y * y; acc[6] += x * y;
            if (a.c.depth_mode == GSR_DEPTH_PEARSON_INV) {
                const double y2 = (double)(1.f / (yf + 200.f));
                acc[7] += y2; acc[8] += y2 * y2; acc[9] += x * y2;
            }
        }
        block_sum_store<kNPart, kLT / 32>(acc, a.partials + (size_t)cta * kNPart, s_red);
    }
}

// ---- GSR_DEPTH_PEARSON_COLS: one correlation coefficient per image column -----------------------------------
// A CTA owns 32 adjacent columns; lane = column (128-byte row segments), the 8 warps stride over the rows.  Sums in
// double, folded over the warps in a fixed order (bit-reproducible).  Writes per column: 1 - r and the four scalars
// the gradient needs:  d(1 - r_c)/dx_i = ka_c (y_i - ybar_c) + kb_c (x_i - xbar_c), already scaled by
// grad_scale * depth_weight / W.
__global__ void __launch_bounds__(256) k_pearson_cols(const LossArgs a)
{
    __shared__ double s_acc[8][5][32];
    const int W = a.c.width, H = a.c.height;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = blockIdx.x * 32 + lane;
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (x < W)
        for (int y = warp; y < H; y += 8) {
            const double xv = a.dimg[(size_t)y * W + x], yv = a.dtarget[(size_t)y * W + x];
            acc[0] += xv; acc[1] += yv; acc[2] += xv * xv; acc[3] += yv * yv; acc[4] += xv * yv;
        }
#pragma unroll
    for (int j = 0; j < 5; j++) s_acc[warp][j][lane] = acc[j];
    __syncthreads();
    if (warp != 0 || x >= W) return;
    double d[5];
#pragma unroll
    for (int j = 0; j < 5; j++) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < 8; w++) v += s_acc[w][j][lane];
        d[j] = v;
    }
    const double n = (double)H;
    const double xbar = d[0] / n, ybar = d[1] / n;
    const double sxx = d[2] - d[0] * d[0] / n, syy = d[3] - d[1] * d[1] / n, sxy = d[4] - d[0] * d[1] / n;
    const double r = sxy / sqrt(sxx * syy);
    const bool clamped = (r > 1.0) || (r < -1.0);
    a.col_loss[x] = 1.0 - fmin(fmax(r, -1.0), 1.0);
    const double k = clamped ? 0.0 : (double)a.c.grad_scale * (double)a.c.depth_weight / (double)W;
    a.colstats[4 * x + 0] = (float)xbar;
    a.colstats[4 * x + 1] = (float)ybar;
    a.colstats[4 * x + 2] = (float)(-k / sqrt(sxx * syy));
    a.colstats[4 * x + 3] = (float)(k * r / sxx);
}

// ---- pass 2: one CTA folds the partials in a fixed order and derives every scalar the gradient needs ------
constexpr int kFinT = 1024;   // the finalize CTA: latency-bound walk over the partial records, so as wide as a CTA gets
__global__ void __launch_bounds__(kFinT) k_loss_finalize(const LossArgs a, int n_tiles)
{
    __shared__ double s_red[kNPart][kFinT / 32];
    __shared__ double s_sum[2][kNPart];
    {   // colour records: 3 slots in use
        double acc[3] = {0.0, 0.0, 0.0};
        for (int i = threadIdx.x; i < 3 * n_tiles; i += kFinT) {
            const double* p = a.partials + (size_t)i * kNPart;
            acc[0] += p[0]; acc[1] += p[1]; acc[2] += p[2];
        }
        block_sum_store<3, kFinT / 32>(acc, s_sum[0], s_red);
        __syncthreads();
    }
    {   // depth records
        double acc[kNPart];
#pragma unroll
        for (int j = 0; j < kNPart; j++) acc[j] = 0.0;
        for (int i = threadIdx.x; i < n_tiles; i += kFinT) {
            const double* p = a.partials + (size_t)(3 * n_tiles + i) * kNPart;
#pragma unroll
            for (int j = 0; j < kNPart; j++) acc[j] += p[j];
        }
        block_sum_store<kNPart, kFinT / 32>(acc, s_sum[1], s_red);
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    const gsr_loss_config& c = a.c;
    const double N = 3.0 * c.width * c.height;
    const double gs = c.grad_scale;
    double color = 0.0, depth = 0.0, ssim_mean = 0.0;
    float st[kNStat];
    for (int j = 0; j < kNStat; j++) st[j] = 0.f;

    if (c.color_mode == GSR_COLOR_L1_SSIM) {
        const double l1 = s_sum[0][0] / N;
        ssim_mean = s_sum[0][1] / N;
        color = (1.0 - (double)c.lambda_dssim) * l1 + (double)c.lambda_dssim * (1.0 - ssim_mean);
        st[kStL1] = (float)(gs * c.color_weight * (1.0 - (double)c.lambda_dssim) / N);
        st[kStSsim] = (float)(-gs * c.color_weight * (double)c.lambda_dssim / N);
    } else if (c.color_mode == GSR_COLOR_MASKED_L1_MEAN) {
        color = s_sum[0][0] / s_sum[0][2];            // empty mask -> 0/0 = NaN, like torch's mean of nothing
        st[kStL1] = (float)(gs * c.color_weight / s_sum[0][2]);
    } else if (c.color_mode == GSR_COLOR_MASKED_L1_SUM) {
        color = s_sum[0][0];
        st[kStL1] = (float)(gs * c.color_weight);
    }

    const double* d = s_sum[1];
    const double n = d[1];
    if (c.depth_mode == GSR_DEPTH_L1_MEAN) {
        depth = d[0] / n;
        st[kStDepthL1] = (float)(gs * c.depth_weight / n);
    } else if (c.depth_mode == GSR_DEPTH_L1_SUM) {
        depth = d[0];
        st[kStDepthL1] = (float)(gs * c.depth_weight);
    } else if (c.depth_mode == GSR_DEPTH_PEARSON || c.depth_mode == GSR_DEPTH_PEARSON_INV) {
        // corr = cov / sqrt(var_x var_y) on centred sums (the (n-1) normalisers cancel), clamped to [-1, 1]
        const double xbar = d[2] / n, sxx = d[4] - d[2] * d[2] / n;
        double ybar = d[3] / n, syy = d[5] - d[3] * d[3] / n, sxy = d[6] - d[2] * d[3] / n;
        double sign = 1.0;
        int variant = 0;
        double r = sxy / sqrt(sxx * syy);
        if (c.depth_mode == GSR_DEPTH_PEARSON_INV) {
            const double r1 = -r;                                           // corr(-y, x)
            const double ybar2 = d[7] / n, syy2 = d[8] - d[7] * d[7] / n, sxy2 = d[9] - d[2] * d[7] / n;
            const double r2 = sxy2 / sqrt(sxx * syy2);                      // corr(1/(y+200), x)
            const double l1 = 1.0 - fmin(fmax(r1, -1.0), 1.0), l2 = 1.0 - fmin(fmax(r2, -1.0), 1.0);
            if (l2 < l1) { variant = 2; r = r2; ybar = ybar2; syy = syy2; }   // python min(a, b): b only if b < a
            else { variant = 1; r = r1; sign = -1.0; }                       // y -> -y: centred y flips sign
        }
        const bool clamped = (r > 1.0) || (r < -1.0);
        depth = 1.0 - fmin(fmax(r, -1.0), 1.0);
        // d(1 - r)/dx_i = -[(y_i - ybar) / sqrt(sxx syy) - r (x_i - xbar) / sxx]
        const double k = clamped ? 0.0 : gs * c.depth_weight;
        st[kStXbar] = (float)xbar;
        st[kStYbar] = (float)ybar;
        st[kStKa] = (float)(-k * sign / sqrt(sxx * syy));
        st[kStKb] = (float)(k * r / sxx);
        st[kStVariant] = (float)variant;
    }
    if (c.depth_mode == GSR_DEPTH_PEARSON_COLS) {   // mean over the columns, in a fixed order
        double sum = 0.0;
        for (int x = 0; x < c.width; x++) sum += a.col_loss[x];
        depth = sum / (double)c.width;
    }
    for (int j = 0; j < kNStat; j++) a.stats[j] = st[j];
    a.losses[0] = (float)((double)c.color_weight * color + (double)c.depth_weight * depth);
    a.losses[1] = (float)color;
    a.losses[2] = (float)depth;
    a.losses[3] = (float)ssim_mean;
}

// ---- pass 3: gradients w.r.t. the two rendered images ------------------------------------------------------
__global__ void __launch_bounds__(kLT) k_loss_bwd(const LossArgs a)
{
    __shared__ float s_m[3][kPH][kPW];
    __shared__ float s_h[3][kPH][kTW];
    const int W = a.c.width, H = a.c.height, HW = W * H;
    const int tid = threadIdx.x, z = blockIdx.z;
    const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH;

    if (z < 3) {
        if (!a.dL_dimage) return;
        const float* X = a.image + (size_t)z * HW;
        const float* Y = a.gt_color + (size_t)z * HW;
        float* G = a.dL_dimage + (size_t)z * HW;
        const float k_l1 = a.stats[kStL1];
        if (a.c.color_mode == GSR_COLOR_L1_SSIM) {
            const float k_ssim = a.stats[kStSsim];
            for (int i = tid; i < kPH * kPW; i += kLT) {
                const int py = i / kPW, px = i - py * kPW;
                const int gx = x0 + px - kWR, gy = y0 + py - kWR;
                const bool in = gx >= 0 && gx < W && gy >= 0 && gy < H;
#pragma unroll
                for (int m = 0; m < 3; m++) s_m[m][py][px] = in ? a.dmaps[(size_t)(m * 3 + z) * HW + gy * W + gx] : 0.f;
            }
            __syncthreads();
            for (int i = tid; i < kPH * kTW; i += kLT) {
                const int r = i / kTW, c = i - r * kTW;
                float h0 = 0.f, h1 = 0.f, h2 = 0.f;
#pragma unroll
                for (int k = 0; k <= 2 * kWR; k++) {
                    const float w = a.win.w[k];
                    h0 += w * s_m[0][r][c + k]; h1 += w * s_m[1][r][c + k]; h2 += w * s_m[2][r][c + k];
                }
                s_h[0][r][c] = h0; s_h[1][r][c] = h1; s_h[2][r][c] = h2;
            }
            __syncthreads();
            const int c = tid & (kTW - 1), r0 = (tid / kTW) * 2;   // one column, two adjacent rows (as in pass 1)
            float co[2][3];
#pragma unroll
            for (int j = 0; j < 2; j++)
#pragma unroll
                for (int q = 0; q < 3; q++) co[j][q] = 0.f;
#pragma unroll
            for (int k = 0; k <= 2 * kWR + 1; k++) {
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    const float h = s_h[q][r0 + k][c];
                    if (k <= 2 * kWR) co[0][q] += a.win.w[k] * h;
                    if (k >= 1) co[1][q] += a.win.w[k - 1] * h;
                }
            }
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int gx = x0 + c, gy = y0 + r0 + j;
                if (gx >= W || gy >= H) continue;
                const int pix = gy * W + gx;
                const float x = X[pix], y = Y[pix];
                G[pix] = k_ssim * (co[j][0] + 2.f * x * co[j][1] + y * co[j][2]) + k_l1 * sgn(x - y);
            }
        } else {
            for (int i = tid; i < kTH * kTW; i += kLT) {
                const int r = i / kTW, c = i - r * kTW;
                const int gx = x0 + c, gy = y0 + r;
                if (gx >= W || gy >= H) continue;
                const int pix = gy * W + gx;
                float g = 0.f;
                if (a.c.color_mode != GSR_COLOR_NONE && loss_mask(a.c.color_mask, a, pix, HW)) g = k_l1 * sgn(X[pix] - Y[pix]);
                G[pix] = g;
            }
        }
    } else if (a.dL_ddimg) {
        const float k_l1 = a.stats[kStDepthL1], xbar = a.stats[kStXbar], ybar = a.stats[kStYbar];
        const float ka = a.stats[kStKa], kb = a.stats[kStKb];
        const int variant = (int)a.stats[kStVariant];
        for (int i = tid; i < kTH * kTW; i += kLT) {
            const int r = i / kTW, c = i - r * kTW;
            const int gx = x0 + c, gy = y0 + r;
            if (gx >= W || gy >= H) continue;
            const int pix = gy * W + gx;
            float g = 0.f;
            if (a.c.depth_mode != GSR_DEPTH_NONE && loss_mask(a.c.depth_mask, a, pix, HW)) {
                const float x = a.dimg[pix], y = a.dtarget[pix];
                if (a.c.depth_mode == GSR_DEPTH_L1_MEAN || a.c.depth_mode == GSR_DEPTH_L1_SUM) {
                    g = k_l1 * sgn(x - y);
                } else if (a.c.depth_mode == GSR_DEPTH_PEARSON_COLS) {
                    const float4 cs = reinterpret_cast<const float4*>(a.colstats)[gx];
                    g = cs.z * (y - cs.y) + cs.w * (x - cs.x);
                } else {
                    const float yv = variant == 2 ? 1.f / (y + 200.f) : y;   // variant 1's sign lives in ka
                    g = ka * (yv - ybar) + kb * (x - xbar);
                }
            }
            a.dL_ddimg[pix] = g;
            a.dL_ddimg[HW + pix] = 0.f;
            a.dL_ddimg[2 * HW + pix] = 0.f;
        }
    }
}

// ---- Adam -------------------------------------------------------------------------------------------------
struct AdamSegs {
    int n;
    long long end[GSR_ADAM_MAX_SEGMENTS];
    float neg_step_size[GSR_ADAM_MAX_SEGMENTS];   // -lr / bias_correction1
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float nss, float w1, float beta2,
                                         float w2, float inv_bc2_sqrt_is_div, float eps)
{
    // torch/optim/adam.py _single_tensor_adam: lerp_, mul_().addcmul_(), (sqrt / bc2_sqrt).add_(eps), addcdiv_
    m = m + w1 * (g - m);
    v = v * beta2 + (w2 * g) * g;
    const float denom = sqrtf(v) / inv_bc2_sqrt_is_div + eps;
    p = p + nss * (m / denom);
}

__global__ void __launch_bounds__(256) k_adam(float* __restrict__ params, float* __restrict__ grads, float* __restrict__ exp_avg,
                                              float* __restrict__ exp_avg_sq, long long n, const AdamSegs segs, float w1,
                                              float beta2, float w2, float bc2_sqrt, float eps, float grad_scale, int zero_grads)
{
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const long long e0 = i * 4;
        int s = 0;
        while (s < segs.n - 1 && e0 >= segs.end[s]) s++;
        float4 p = reinterpret_cast<float4*>(params)[i];
        float4 g = __ldcs(reinterpret_cast<const float4*>(grads) + i);
        float4 m = reinterpret_cast<float4*>(exp_avg)[i];
        float4 v = reinterpret_cast<float4*>(exp_avg_sq)[i];
        float nss[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int sk = s;
            while (sk < segs.n - 1 && e0 + k >= segs.end[sk]) sk++;
            nss[k] = segs.neg_step_size[sk];
        }
        adam_one(p.x, g.x * grad_scale, m.x, v.x, nss[0], w1, beta2, w2, bc2_sqrt, eps);
        adam_one(p.y, g.y * grad_scale, m.y, v.y, nss[1], w1, beta2, w2, bc2_sqrt, eps);
        adam_one(p.z, g.z * grad_scale, m.z, v.z, nss[2], w1, beta2, w2, bc2_sqrt, eps);
        adam_one(p.w, g.w * grad_scale, m.w, v.w, nss[3], w1, beta2, w2, bc2_sqrt, eps);
        reinterpret_cast<float4*>(params)[i] = p;
        reinterpret_cast<float4*>(exp_avg)[i] = m;
        reinterpret_cast<float4*>(exp_avg_sq)[i] = v;
        if (zero_grads) reinterpret_cast<float4*>(grads)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // tail (n not a multiple of 4)
    const long long t = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) {
        int s = 0;
        while (s < segs.n - 1 && t >= segs.end[s]) s++;
        float p = params[t], m = exp_avg[t], v = exp_avg_sq[t];
        adam_one(p, grads[t] * grad_scale, m, v, segs.neg_step_size[s], w1, beta2, w2, bc2_sqrt, eps);
        params[t] = p; exp_avg[t] = m; exp_avg_sq[t] = v;
        if (zero_grads) grads[t] = 0.f;
    }
}

struct LossWS { float* dmaps; double* partials; float* stats; float* colstats; double* col_loss; size_t total; };
static LossWS loss_ws_carve(char* base, int W, int H)
{
    LossWS w;
    size_t off = 0;
    const size_t tiles = (size_t)((W + kTW - 1) / kTW) * ((H + kTH - 1) / kTH);
    w.dmaps = (float*)(base + off);      off = align_up(off + (size_t)9 * W * H * sizeof(float), 256);
    w.partials = (double*)(base + off);  off = align_up(off + tiles * 4 * kNPart * sizeof(double), 256);
    w.stats = (float*)(base + off);      off = align_up(off + kNStat * sizeof(float), 256);
    w.colstats = (float*)(base + off);   off = align_up(off + (size_t)W * 4 * sizeof(float), 256);
    w.col_loss = (double*)(base + off);  off = align_up(off + (size_t)W * sizeof(double), 256);
    w.total = off;
    return w;
}

}  // namespace gsr

using namespace gsr;

extern "C" {

size_t gsr_slam_loss_ws_bytes(int32_t width, int32_t height)
{
    if (width <= 0 || height <= 0) return 0;
    return loss_ws_carve(nullptr, width, height).total;
}

int gsr_slam_loss(gsr_stream_t stream_, const gsr_loss_config* cfg, const float* image, const float* depth_image,
                  const float* gt_color, const float* depth_target, const float* gt_depth, void* ws, size_t ws_bytes,
                  float* losses, float* dL_dimage, float* dL_ddepth_image)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!cfg || !losses) return api_fail(GSR_ERR_INVALID, "null config / losses");
    if (cfg->width <= 0 || cfg->height <= 0) return api_fail(GSR_ERR_INVALID, "image size must be positive");
    if (cfg->color_mode < GSR_COLOR_NONE || cfg->color_mode > GSR_COLOR_MASKED_L1_SUM ||
        cfg->depth_mode < GSR_DEPTH_NONE || cfg->depth_mode > GSR_DEPTH_PEARSON_COLS)
        return api_fail(GSR_ERR_INVALID, "unknown colour / depth loss mode");
    if (cfg->depth_mode == GSR_DEPTH_PEARSON_COLS && cfg->depth_mask != 0)
        return api_fail(GSR_ERR_INVALID, "GSR_DEPTH_PEARSON_COLS is the unmasked form: depth_mask must be 0");
    if ((cfg->color_mask | cfg->depth_mask) & ~7) return api_fail(GSR_ERR_INVALID, "unknown mask flag");
    const bool masked_color = cfg->color_mode == GSR_COLOR_MASKED_L1_MEAN || cfg->color_mode == GSR_COLOR_MASKED_L1_SUM;
    const int used_masks = (masked_color ? cfg->color_mask : 0) | (cfg->depth_mode != GSR_DEPTH_NONE ? cfg->depth_mask : 0);
    if (cfg->color_mode != GSR_COLOR_NONE && (!image || !gt_color)) return api_fail(GSR_ERR_INVALID, "image / gt_color required");
    if (cfg->depth_mode != GSR_DEPTH_NONE && (!depth_image || !depth_target))
        return api_fail(GSR_ERR_INVALID, "depth_image / depth_target required by the depth term");
    if ((used_masks & (GSR_MASK_NOT_NAN | GSR_MASK_SILHOUETTE)) && !depth_image)
        return api_fail(GSR_ERR_INVALID, "depth_image required by the mask");
    if ((used_masks & GSR_MASK_GT_DEPTH_POS) && !gt_depth) return api_fail(GSR_ERR_INVALID, "gt_depth required by the mask");
    if (cfg->color_mode == GSR_COLOR_NONE && dL_dimage && !image) return api_fail(GSR_ERR_INVALID, "image required");
    if (!ws || ws_bytes < gsr_slam_loss_ws_bytes(cfg->width, cfg->height)) return api_fail(GSR_ERR_WORKSPACE, "loss workspace too small");
    const bool want_grad = dL_dimage != nullptr || dL_ddepth_image != nullptr;
    if (want_grad && cfg->color_mode != GSR_COLOR_NONE && !dL_dimage) return api_fail(GSR_ERR_INVALID, "dL_dimage required");
    if (want_grad && cfg->depth_mode != GSR_DEPTH_NONE && !dL_ddepth_image) return api_fail(GSR_ERR_INVALID, "dL_ddepth_image required");

    LossWS w = loss_ws_carve((char*)ws, cfg->width, cfg->height);
    LossArgs a;
    a.c = *cfg;
    a.image = image; a.dimg = depth_image; a.gt_color = gt_color; a.dtarget = depth_target; a.gt_depth = gt_depth;
    a.dmaps = w.dmaps; a.partials = w.partials; a.stats = w.stats; a.losses = losses;
    a.colstats = w.colstats; a.col_loss = w.col_loss;
    a.dL_dimage = dL_dimage; a.dL_ddimg = dL_ddepth_image;
    {   // R/utils/loss_utils.py:98-105: float32 tensor of exp(-(x-5)^2 / (2 sigma^2)), divided by its float32 sum
        float g[2 * kWR + 1], sum = 0.f;
        for (int i = 0; i <= 2 * kWR; i++) { g[i] = (float)exp(-(double)((i - kWR) * (i - kWR)) / (2.0 * 1.5 * 1.5)); sum += g[i]; }
        for (int i = 0; i <= 2 * kWR; i++) a.win.w[i] = g[i] / sum;
    }
    const int tx = (cfg->width + kTW - 1) / kTW, ty = (cfg->height + kTH - 1) / kTH;
    const dim3 grid(tx, ty, 4);
    k_loss_fwd<<<grid, kLT, 0, stream>>>(a);
    int launches = 2;
    if (cfg->depth_mode == GSR_DEPTH_PEARSON_COLS) {
        k_pearson_cols<<<(cfg->width + 31) / 32, 256, 0, stream>>>(a);
        launches++;
    }
    k_loss_finalize<<<1, kFinT, 0, stream>>>(a, tx * ty);
    if (want_grad) {
        k_loss_bwd<<<grid, kLT, 0, stream>>>(a);   // CTAs of an image without a gradient buffer return at once
        launches++;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return api_fail(GSR_ERR_CUDA, "slam loss launch", e);
    api_count_launches(launches);
    return GSR_OK;
}

int gsr_adam_step(gsr_stream_t stream_, float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                  int32_t num_segments, const int64_t* seg_end, const double* seg_lr, double beta1, double beta2, double eps,
                  int64_t step, float grad_scale, int32_t zero_grads)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || step < 1) return api_fail(GSR_ERR_INVALID, "n < 0 or step < 1");
    if (n == 0) return GSR_OK;
    if (!params || !grads || !exp_avg || !exp_avg_sq) return api_fail(GSR_ERR_INVALID, "null bucket pointer");
    if (num_segments < 1 || num_segments > GSR_ADAM_MAX_SEGMENTS || !seg_end || !seg_lr)
        return api_fail(GSR_ERR_INVALID, "1..GSR_ADAM_MAX_SEGMENTS segments required");
    if ((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) != 0)
        return api_fail(GSR_ERR_INVALID, "buckets must be 16-byte aligned");
    AdamSegs segs;
    segs.n = num_segments;
    // torch/optim/adam.py: python-double scalars, rounded to fp32 when they meet the tensors
    const double bc1 = 1.0 - pow(beta1, (double)step);
    const double bc2 = 1.0 - pow(beta2, (double)step);
    long long prev = 0;
    for (int s = 0; s < num_segments; s++) {
        if (seg_end[s] < prev) return api_fail(GSR_ERR_INVALID, "segment ends must be non-decreasing");
        prev = seg_end[s];
        segs.end[s] = seg_end[s];
        segs.neg_step_size[s] = (float)(-(seg_lr[s] / bc1));
    }
    if (prev != n) return api_fail(GSR_ERR_INVALID, "last segment must end at n");
    for (int s = num_segments; s < GSR_ADAM_MAX_SEGMENTS; s++) { segs.end[s] = n; segs.neg_step_size[s] = 0.f; }
    const float w1 = (float)(1.0 - beta1), w2 = (float)(1.0 - beta2);
    const long long n4 = n >> 2;
    int blocks = (int)((n4 + 255) / 256);
    const int cap = 148 * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    k_adam<<<blocks, 256, 0, stream>>>(params, grads, exp_avg, exp_avg_sq, n, segs, w1, (float)beta2, w2, (float)sqrt(bc2), (float)eps,
                                       grad_scale, zero_grads);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return api_fail(GSR_ERR_CUDA, "adam launch", e);
    api_count_launches(1);
    return GSR_OK;
}

}  // extern "C"
