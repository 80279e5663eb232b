// hostcheck.cpp — runs the product's per-Gaussian math (mm3dgs-slam_b200/csrc/gsr_math.cuh, the GSR_HD functions the
// CUDA kernels call) on the CPU, so that tests/test_hostcheck.py can compare it with the oracle without a GPU.
// Test infrastructure: compiled by the test with g++, never linked into the library.
#include "gsr_cull.cuh"
#include "gsr_math.cuh"
#include "gsr_bwd_math.cuh"
#include <string.h>

using namespace gsr;

extern "C" {

// Forward of k_preprocess_fwd's per-Gaussian part: covariance from (scale, raw quaternion), projection, conic,
// radius, tile rectangle; SH colour with the kernel's direction normalisation.  Outputs are zero for culled Gaussians.
void hc_forward(int P, int deg, int M, const float* means, const float* scales, const float* rots, float mod,
                const float* shs, const float* view, const float* proj, const float* campos, int W, int H, float tanfovx,
                float tanfovy, int* radii, int* tiles, int* rect, float* depth, float* xy, float* conic, float* rgb,
                int* clamp_bits, float* lam_max)
{
    const float focal_x = W / (2.0f * tanfovx), focal_y = H / (2.0f * tanfovy);
    const int gx = (W + GSR_TILE - 1) / GSR_TILE, gy = (H + GSR_TILE - 1) / GSR_TILE;
    for (int i = 0; i < P; i++) {
        const V3 p = {means[3 * i], means[3 * i + 1], means[3 * i + 2]};
        const V3 sc = {scales[3 * i], scales[3 * i + 1], scales[3 * i + 2]};
        const V4 q = {rots[4 * i], rots[4 * i + 1], rots[4 * i + 2], rots[4 * i + 3]};
        float cov6[6];
        cov3d_from_scale_rot(sc, mod, q, cov6);
        const PreOut o = preprocess_one(p, cov6, view, proj, W, H, tanfovx, tanfovy, focal_x, focal_y, gx, gy);
        radii[i] = o.radius;
        tiles[i] = o.tiles;
        rect[4 * i] = o.x0; rect[4 * i + 1] = o.y0; rect[4 * i + 2] = o.x1; rect[4 * i + 3] = o.y1;
        depth[i] = o.depth;
        xy[2 * i] = o.px; xy[2 * i + 1] = o.py;
        conic[3 * i] = o.cx; conic[3 * i + 1] = o.cy; conic[3 * i + 2] = o.cz;
        rgb[3 * i] = rgb[3 * i + 1] = rgb[3 * i + 2] = 0.f;
        clamp_bits[i] = 0;
        lam_max[i] = o.lam_max;
        if (o.radius > 0) {
            V3 dir = {p.x - campos[0], p.y - campos[1], p.z - campos[2]};
            const float len = sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
            dir.x = dir.x / len; dir.y = dir.y / len; dir.z = dir.z / len;
            const V3 c = sh_to_rgb(deg, shs + (size_t)i * M * 3, dir);
            clamp_bits[i] = (c.x < 0 ? 1 : 0) | (c.y < 0 ? 2 : 0) | (c.z < 0 ? 4 : 0);
            rgb[3 * i] = fmaxf(c.x, 0.f); rgb[3 * i + 1] = fmaxf(c.y, 0.f); rgb[3 * i + 2] = fmaxf(c.z, 0.f);
        }
    }
}

// Covariance path of k_preprocess_bwd (csrc/gsr_bwd_math.cuh): dL/dconic -> dL/dmean (through J and the view
// transform), dL/dscale, dL/d(raw quaternion) on the (scale, rotation) path, and dL/dcov3D + dL/dmean on the
// precomputed-covariance path (fed with the covariance the forward builds from the same scale / rotation, so both
// paths must agree on dL/dmean).  The conic is the forward's, as in the kernel.  Only Gaussians with radii > 0 are touched.
void hc_cov_backward(int P, const int* radii, const float* means, const float* scales, const float* rots, float mod,
                     const float* view, const float* proj, int W, int H, float tanfovx, float tanfovy, const float* dconic,
                     float* dcov, float* dmean, float* dscale, float* drot, float* dmean_precomp)
{
    const float focal_x = W / (2.0f * tanfovx), focal_y = H / (2.0f * tanfovy);
    const int gx = (W + GSR_TILE - 1) / GSR_TILE, gy = (H + GSR_TILE - 1) / GSR_TILE;
    for (int i = 0; i < P; i++) {
        for (int k = 0; k < 6; k++) dcov[6 * i + k] = 0.f;
        for (int k = 0; k < 3; k++) dmean[3 * i + k] = dscale[3 * i + k] = dmean_precomp[3 * i + k] = 0.f;
        for (int k = 0; k < 4; k++) drot[4 * i + k] = 0.f;
        if (radii[i] <= 0) continue;
        const V3 p = {means[3 * i], means[3 * i + 1], means[3 * i + 2]};
        const V3 sc = {scales[3 * i], scales[3 * i + 1], scales[3 * i + 2]};
        const V4 q = {rots[4 * i], rots[4 * i + 1], rots[4 * i + 2], rots[4 * i + 3]};
        float cov6[6];
        cov3d_from_scale_rot(sc, mod, q, cov6);
        const PreOut o = preprocess_one(p, cov6, view, proj, W, H, tanfovx, tanfovy, focal_x, focal_y, gx, gy);
        const ViewFrame f = view_frame(p, view, focal_x, focal_y, tanfovx, tanfovy);
        const Sym2 Hq = cov2d_grad_from_conic(o.cx, o.cy, o.cz, dconic[3 * i], dconic[3 * i + 1], dconic[3 * i + 2]);
        const V3 s_eff = {mod * sc.x, mod * sc.y, mod * sc.z};
        V3 ds, dM0, dM1;
        V4 dq;
        cov_chain_scale_rot(f, Hq, s_eff, q, ds, dq, dM0, dM1);
        V3 dm = view_t_mul(f, view_chain(f, dM0, dM1, focal_x, focal_y));
        dmean[3 * i] = dm.x; dmean[3 * i + 1] = dm.y; dmean[3 * i + 2] = dm.z;
        dscale[3 * i] = ds.x; dscale[3 * i + 1] = ds.y; dscale[3 * i + 2] = ds.z;
        drot[4 * i] = dq.x; drot[4 * i + 1] = dq.y; drot[4 * i + 2] = dq.z; drot[4 * i + 3] = dq.w;
        cov_chain_precomp(f, Hq, cov6, dcov + 6 * i, dM0, dM1);
        dm = view_t_mul(f, view_chain(f, dM0, dM1, focal_x, focal_y));
        dmean_precomp[3 * i] = dm.x; dmean_precomp[3 * i + 1] = dm.y; dmean_precomp[3 * i + 2] = dm.z;
    }
}

// Screen-position path: dL/dmean2D (NDC units) -> dL/dmean.
void hc_ndc_backward(int P, const int* radii, const float* means, const float* proj, const float* dmean2d, float* dmean)
{
    for (int i = 0; i < P; i++) {
        for (int k = 0; k < 3; k++) dmean[3 * i + k] = 0.f;
        if (radii[i] <= 0) continue;
        const V3 p = {means[3 * i], means[3 * i + 1], means[3 * i + 2]};
        V3 dh;
        const V3 dm = ndc_chain(p, proj, dmean2d[2 * i], dmean2d[2 * i + 1], dh);
        dmean[3 * i] = dm.x; dmean[3 * i + 1] = dm.y; dmean[3 * i + 2] = dm.z;
    }
}

// SH path of k_preprocess_bwd: clamp-masked dL/dRGB -> dL/dsh and the view-direction part of dL/dmean.
void hc_sh_backward(int P, int deg, int M, const int* radii, const int* clamp_bits, const float* means, const float* shs,
                    const float* campos, const float* dcolors, float* dsh, float* dmean)
{
    for (int i = 0; i < P; i++) {
        for (int k = 0; k < 3 * M; k++) dsh[(size_t)i * M * 3 + k] = 0.f;
        for (int k = 0; k < 3; k++) dmean[3 * i + k] = 0.f;
        if (radii[i] <= 0) continue;
        const V3 d0 = {means[3 * i] - campos[0], means[3 * i + 1] - campos[1], means[3 * i + 2] - campos[2]};
        const float inv_len = GSR_RSQRT(d0.x * d0.x + d0.y * d0.y + d0.z * d0.z);
        const V3 dir = {d0.x * inv_len, d0.y * inv_len, d0.z * inv_len};
        float dRGB[3];
        for (int ch = 0; ch < 3; ch++) dRGB[ch] = ((clamp_bits[i] >> ch) & 1) ? 0.f : dcolors[3 * i + ch];
        const V3 ddir = sh_grad(deg, shs + (size_t)i * M * 3, dir, dRGB, dsh + (size_t)i * M * 3);
        const V3 dm = unit_vector_grad(dir, inv_len, ddir);
        dmean[3 * i] = dm.x; dmean[3 * i + 1] = dm.y; dmean[3 * i + 2] = dm.z;
    }
}

// The work-skipping tests of the blend kernels (gsr_cull.cuh) against the reference's per-pixel rule
// (CR/forward.cu:332-345): for every splat and every 8x4 sub-tile of the W x H image, if ANY pixel of the sub-tile
// would blend the splat (power <= 0 and min(0.99, o exp(power)) >= 1/255), both tests must keep the pair.
// Thresholds are built exactly as k_preprocess_fwd stores them (low 3 bits of the power threshold replaced by flags).
// Returns the number of violations; stats[0] = pairs that contribute, stats[1] = pairs kept by the circle test,
// stats[2] = pairs kept by circle + ellipse tests, stats[3] = all pairs examined.
long long hc_cull_check(int n, const float* xy, const float* conic, const float* opacity, const float* lam_max, int W, int H,
                        long long* stats)
{
    long long bad = 0;
    stats[0] = stats[1] = stats[2] = stats[3] = 0;
    for (int i = 0; i < n; i++) {
        const float x = xy[2 * i], y = xy[2 * i + 1];
        const float A = conic[3 * i], B = conic[3 * i + 1], C = conic[3 * i + 2], o = opacity[i];
        const float rc2 = cull_radius2(lam_max[i], o);
        float lim = cull_power(lam_max[i], o);
        {   // the kernel stores SH clamp flags in the low 3 mantissa bits; worst case for the test is all flags clear
            unsigned u;
            memcpy(&u, &lim, 4);
            u &= ~7u;
            memcpy(&lim, &u, 4);
        }
        for (int sy = 0; sy < H; sy += 4)
            for (int sx = 0; sx < W; sx += 8) {
                bool contributes = false;
                for (int py = sy; py < sy + 4 && !contributes; py++)
                    for (int px = sx; px < sx + 8; px++) {
                        const float dx = x - (float)px, dy = y - (float)py;
                        const float power = -0.5f * (A * dx * dx + C * dy * dy) - B * dx * dy;
                        if (power > 0.0f) continue;
                        const float alpha = fminf(0.99f, o * expf(power));
                        if (alpha < 1.0f / 255.0f) continue;
                        contributes = true;
                        break;
                    }
                const float sx0 = (float)sx, sx1 = (float)(sx + 7), sy0 = (float)sy, sy1 = (float)(sy + 3);
                const bool h1 = subtile_hit(sx0, sx1, sy0, sy1, x, y, rc2);
                const bool h2 = h1 && subtile_hit_ellipse(sx0, sx1, sy0, sy1, x, y, A, B, C, lim);
                stats[3]++;
                stats[0] += contributes;
                stats[1] += h1;
                stats[2] += h2;
                if (contributes && !h2) bad++;
            }
    }
    return bad;
}

// Exhaustive check of the tile partition's division-by-multiplication (div_magic / div_by_magic): every rectangle
// width w the API admits (tile grids up to 1023 wide) and every instance index k < w * max_h.
long long hc_magic_div_check(int max_w, int max_h)
{
    long long bad = 0;
    for (uint32_t w = 1; w <= (uint32_t)max_w; w++) {
        const uint32_t magic = div_magic(w);
        const uint32_t kmax = w * (uint32_t)max_h;
        for (uint32_t k = 0; k < kmax; k++) bad += (div_by_magic(k, w, magic) != k / w);
    }
    return bad;
}

}  // extern "C"
