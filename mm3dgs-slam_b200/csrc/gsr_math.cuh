// gsr_math.cuh — per-Gaussian math shared by the forward and backward preprocessing kernels.
//
// Numerical contract: every value that feeds an INTEGER output of the reference (radius, tile
// rectangle, tiles_touched, the depth bits of the sort key) is evaluated with the same
// association order as the reference's expressions (CR = /root/reference/submodules/
// diff-gaussian-rasterization/cuda_rasterizer), including glm's column-major 3x3 product
// (sum over k left to right, zero terms kept), so that nvcc's default FMA contraction yields the
// same bits.  Both builds use nvcc defaults (-fmad=true, IEEE div/sqrt, no fast-math).
//
// All functions are GSR_HD so a host-side harness (tests/hostcheck) can run the same math on CPU.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define GSR_HD __host__ __device__ __forceinline__
#else
#define GSR_HD inline
#endif

#define GSR_TILE 16

namespace gsr {

// Real spherical-harmonics constants (same values as the reference, CR/auxiliary.h:21-39).
#define GSR_SH_C0 0.28209479177387814f
#define GSR_SH_C1 0.4886025119029199f
#define GSR_SH_C2_0 1.0925484305920792f
#define GSR_SH_C2_1 -1.0925484305920792f
#define GSR_SH_C2_2 0.31539156525252005f
#define GSR_SH_C2_3 -1.0925484305920792f
#define GSR_SH_C2_4 0.5462742152960396f
#define GSR_SH_C3_0 -0.5900435899266435f
#define GSR_SH_C3_1 2.890611442640554f
#define GSR_SH_C3_2 -0.4570457994644658f
#define GSR_SH_C3_3 0.3731763325901154f
#define GSR_SH_C3_4 -0.4570457994644658f
#define GSR_SH_C3_5 1.445305721320277f
#define GSR_SH_C3_6 -0.5900435899266435f

struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };

// Column-major 3x3 (c[col][row]) with glm's product order: R[j][i] = sum_k A[k][i]*B[j][k].
struct M3 {
    float c[3][3];
};

GSR_HD M3 m3_mul(const M3& A, const M3& B)
{
    M3 R;
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 3; i++)
            R.c[j][i] = A.c[0][i] * B.c[j][0] + A.c[1][i] * B.c[j][1] + A.c[2][i] * B.c[j][2];
    return R;
}

GSR_HD M3 m3_transpose(const M3& A)
{
    M3 R;
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 3; i++) R.c[j][i] = A.c[i][j];
    return R;
}

GSR_HD M3 m3_cols(float a0, float a1, float a2, float b0, float b1, float b2, float c0, float c1, float c2)
{
    M3 R;
    R.c[0][0] = a0; R.c[0][1] = a1; R.c[0][2] = a2;
    R.c[1][0] = b0; R.c[1][1] = b1; R.c[1][2] = b2;
    R.c[2][0] = c0; R.c[2][1] = c1; R.c[2][2] = c2;
    return R;
}

// x' = m[0]x + m[4]y + m[8]z + m[12]  (CR/auxiliary.h:58-77)
GSR_HD V3 xform4x3(const V3& p, const float* m)
{
    V3 r;
    r.x = m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12];
    r.y = m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13];
    r.z = m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14];
    return r;
}
GSR_HD V4 xform4x4(const V3& p, const float* m)
{
    V4 r;
    r.x = m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12];
    r.y = m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13];
    r.z = m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14];
    r.w = m[3] * p.x + m[7] * p.y + m[11] * p.z + m[15];
    return r;
}

// NDC -> pixel; the reference evaluates this in double (double literals), CR/auxiliary.h:41-44.
// Exact k / w by multiplication, used to unflatten an instance index k into (row, column) of a w-wide tile
// rectangle: q = umulhi(k, ceil(2^32 / w)) equals k / w whenever k * w < 2^32 (here k < 2^20 instances per splat,
// w < 2^12 tiles).  w == 1 has no 32-bit magic (2^32) and is handled by the caller's fast path.
GSR_HD uint32_t div_magic(uint32_t w) { return 0xffffffffu / w + 1u; }
GSR_HD uint32_t div_by_magic(uint32_t k, uint32_t w, uint32_t magic)
{
#if defined(__CUDA_ARCH__)
    return (w == 1u) ? k : __umulhi(k, magic);
#else
    return (w == 1u) ? k : (uint32_t)(((unsigned long long)k * magic) >> 32);
#endif
}

GSR_HD float ndc2pix(float v, int S) { return (float)(((v + 1.0) * S - 1.0) * 0.5); }

// Tile rectangle of a splat of integer radius r centred at p (CR/auxiliary.h:46-56).
GSR_HD void tile_rect(float px, float py, int r, int gx, int gy, int& x0, int& y0, int& x1, int& y1)
{
    int a = (int)((px - r) / GSR_TILE);
    int b = (int)((py - r) / GSR_TILE);
    int c = (int)((px + r + GSR_TILE - 1) / GSR_TILE);
    int d = (int)((py + r + GSR_TILE - 1) / GSR_TILE);
    x0 = a < 0 ? 0 : (a > gx ? gx : a);
    y0 = b < 0 ? 0 : (b > gy ? gy : b);
    x1 = c < 0 ? 0 : (c > gx ? gx : c);
    y1 = d < 0 ? 0 : (d > gy ? gy : d);
}

// World-space covariance (upper triangle) from scale and the RAW quaternion (r,x,y,z).
// Sigma = (S R)^T (S R) in glm's column-major convention == R_std S^2 R_std^T (CR/forward.cu:118-152).
GSR_HD void cov3d_from_scale_rot(const V3& scale, float mod, const V4& q, float* cov6)
{
    M3 S = m3_cols(mod * scale.x, 0.f, 0.f, 0.f, mod * scale.y, 0.f, 0.f, 0.f, mod * scale.z);
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    M3 R = m3_cols(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
                   2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
                   2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
    M3 M = m3_mul(S, R);
    M3 Sg = m3_mul(m3_transpose(M), M);
    cov6[0] = Sg.c[0][0];
    cov6[1] = Sg.c[0][1];
    cov6[2] = Sg.c[0][2];
    cov6[3] = Sg.c[1][1];
    cov6[4] = Sg.c[1][2];
    cov6[5] = Sg.c[2][2];
}

// Intermediates of the EWA projection (CR/forward.cu:74-113; reused by CR/backward.cu:144-198).
struct Cov2D {
    M3 T;         // T = W * J (glm order); T.c[0][k], T.c[1][k] are the rows of d(u,v)/d(world)
    float a, b, c; // 2-D covariance with the +0.3 low-pass on the diagonal
    V3 t;         // view-space mean with x/z, y/z clamped to the 1.3*tanfov guard band
    float txtz, tytz, limx, limy;
};

GSR_HD Cov2D cov2d_project(const V3& mean, float focal_x, float focal_y, float tanfovx, float tanfovy,
                           const float* cov6, const float* view)
{
    Cov2D o;
    V3 t = xform4x3(mean, view);
    o.limx = 1.3f * tanfovx;
    o.limy = 1.3f * tanfovy;
    o.txtz = t.x / t.z;
    o.tytz = t.y / t.z;
    t.x = fminf(o.limx, fmaxf(-o.limx, o.txtz)) * t.z;
    t.y = fminf(o.limy, fmaxf(-o.limy, o.tytz)) * t.z;
    o.t = t;
    M3 J = m3_cols(focal_x / t.z, 0.0f, -(focal_x * t.x) / (t.z * t.z),
                   0.0f, focal_y / t.z, -(focal_y * t.y) / (t.z * t.z),
                   0.f, 0.f, 0.f);
    M3 W = m3_cols(view[0], view[4], view[8], view[1], view[5], view[9], view[2], view[6], view[10]);
    o.T = m3_mul(W, J);
    M3 Vrk = m3_cols(cov6[0], cov6[1], cov6[2], cov6[1], cov6[3], cov6[4], cov6[2], cov6[4], cov6[5]);
    M3 cov = m3_mul(m3_mul(m3_transpose(o.T), m3_transpose(Vrk)), o.T);
    o.a = cov.c[0][0] + 0.3f;
    o.b = cov.c[0][1];
    o.c = cov.c[1][1] + 0.3f;
    return o;
}

// SH basis evaluation (degree <= 3) for one Gaussian; sh points at M float3 coefficients.
// Returns the un-clamped colour + 0.5 (CR/forward.cu:20-71).
GSR_HD V3 sh_to_rgb(int deg, const float* sh, const V3& dir)
{
#define SHV(k, ch) sh[3 * (k) + (ch)]
    float res[3];
    const float x = dir.x, y = dir.y, z = dir.z;
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        float r = GSR_SH_C0 * SHV(0, ch);
        if (deg > 0) {
            r = r - GSR_SH_C1 * y * SHV(1, ch) + GSR_SH_C1 * z * SHV(2, ch) - GSR_SH_C1 * x * SHV(3, ch);
            if (deg > 1) {
                float xx = x * x, yy = y * y, zz = z * z;
                float xy = x * y, yz = y * z, xz = x * z;
                r = r + GSR_SH_C2_0 * xy * SHV(4, ch) + GSR_SH_C2_1 * yz * SHV(5, ch) +
                    GSR_SH_C2_2 * (2.0f * zz - xx - yy) * SHV(6, ch) + GSR_SH_C2_3 * xz * SHV(7, ch) +
                    GSR_SH_C2_4 * (xx - yy) * SHV(8, ch);
                if (deg > 2) {
                    r = r + GSR_SH_C3_0 * y * (3.0f * xx - yy) * SHV(9, ch) + GSR_SH_C3_1 * xy * z * SHV(10, ch) +
                        GSR_SH_C3_2 * y * (4.0f * zz - xx - yy) * SHV(11, ch) +
                        GSR_SH_C3_3 * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * SHV(12, ch) +
                        GSR_SH_C3_4 * x * (4.0f * zz - xx - yy) * SHV(13, ch) +
                        GSR_SH_C3_5 * z * (xx - yy) * SHV(14, ch) + GSR_SH_C3_6 * x * (xx - 3.0f * yy) * SHV(15, ch);
                }
            }
        }
        res[ch] = r + 0.5f;
    }
#undef SHV
    V3 o = {res[0], res[1], res[2]};
    return o;
}

// ------------------------------------------------------------------------------------------
// Forward preprocessing of one Gaussian (CR/forward.cu:155-256).
// ------------------------------------------------------------------------------------------
struct PreOut {
    int radius;      // 0 => culled
    int tiles;       // tiles touched
    float depth;     // view-space z
    float px, py;    // pixel-space mean
    float cx, cy, cz; // conic (inverse 2-D covariance)
    float lam_max;   // larger eigenvalue of the 2-D covariance (for the blend kernels' cull radius)
    int x0, y0, x1, y1; // tile rectangle
};

GSR_HD PreOut preprocess_one(const V3& p, const float* cov6, const float* view, const float* proj, int W, int H,
                             float tanfovx, float tanfovy, float focal_x, float focal_y, int gx, int gy)
{
    PreOut o;
    o.radius = 0;
    o.tiles = 0;
    o.depth = 0.f;
    o.px = o.py = o.cx = o.cy = o.cz = o.lam_max = 0.f;
    o.x0 = o.y0 = o.x1 = o.y1 = 0;
    V4 p_hom = xform4x4(p, proj);
    float p_w = 1.0f / (p_hom.w + 0.0000001f);
    float projx = p_hom.x * p_w, projy = p_hom.y * p_w;
    V3 p_view = xform4x3(p, view);
    if (p_view.z <= 0.1f) return o;  // near cull only (CR/auxiliary.h:154)

    Cov2D c2 = cov2d_project(p, focal_x, focal_y, tanfovx, tanfovy, cov6, view);
    float det = (c2.a * c2.c - c2.b * c2.b);
    if (det == 0.0f) return o;
    float det_inv = 1.f / det;
    o.cx = c2.c * det_inv;
    o.cy = -c2.b * det_inv;
    o.cz = c2.a * det_inv;
    float mid = 0.5f * (c2.a + c2.c);
    float lambda1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
    float lambda2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
    float my_radius = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));
    o.px = ndc2pix(projx, W);
    o.py = ndc2pix(projy, H);
    int x0, y0, x1, y1;
    tile_rect(o.px, o.py, (int)my_radius, gx, gy, x0, y0, x1, y1);
    int n = (x1 - x0) * (y1 - y0);
    if (n == 0) return o;
    o.radius = (int)my_radius;
    o.tiles = n;
    o.x0 = x0; o.y0 = y0; o.x1 = x1; o.y1 = y1;
    o.depth = p_view.z;
    o.lam_max = fmaxf(lambda1, lambda2);
    return o;
}

// ------------------------------------------------------------------------------------------
// Backward preprocessing of one Gaussian (CR/backward.cu:144-274 cov2D, :346-396 projection,
// :20-139 SH, :278-341 scale/rotation).
// ------------------------------------------------------------------------------------------

// d/d(cov2D entries) -> dL/dcov3D[6] and the covariance-path part of dL/dmean.
// Also returns dL/dT (2x3) and dL/dt for the camera-gradient extension.
struct Cov2DGrad {
    float dcov[6];
    V3 dmean;      // W^T dL/dt
    V3 dt;         // dL/dt (view-space mean)
    float dT[2][3]; // dL/dT00..02, dL/dT10..12
};

GSR_HD Cov2DGrad cov2d_backward(const V3& mean, const float* cov6, float h_x, float h_y, float tanfovx,
                                float tanfovy, const float* view, float dconx, float dcony, float dconz)
{
    Cov2DGrad g;
    Cov2D c2 = cov2d_project(mean, h_x, h_y, tanfovx, tanfovy, cov6, view);
    const M3& T = c2.T;
    const V3 t = c2.t;
    const float x_grad_mul = (c2.txtz < -c2.limx || c2.txtz > c2.limx) ? 0.f : 1.f;
    const float y_grad_mul = (c2.tytz < -c2.limy || c2.tytz > c2.limy) ? 0.f : 1.f;
    const float a = c2.a, b = c2.b, c = c2.c;
    float denom = a * c - b * b;
    float dL_da = 0, dL_db = 0, dL_dc = 0;
    float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
    if (denom2inv != 0) {
        dL_da = denom2inv * (-c * c * dconx + 2 * b * c * dcony + (denom - a * c) * dconz);
        dL_dc = denom2inv * (-a * a * dconz + 2 * a * b * dcony + (denom - a * c) * dconx);
        dL_db = denom2inv * 2 * (b * c * dconx - (denom + 2 * b * b) * dcony + a * b * dconz);
        g.dcov[0] = (T.c[0][0] * T.c[0][0] * dL_da + T.c[0][0] * T.c[1][0] * dL_db + T.c[1][0] * T.c[1][0] * dL_dc);
        g.dcov[3] = (T.c[0][1] * T.c[0][1] * dL_da + T.c[0][1] * T.c[1][1] * dL_db + T.c[1][1] * T.c[1][1] * dL_dc);
        g.dcov[5] = (T.c[0][2] * T.c[0][2] * dL_da + T.c[0][2] * T.c[1][2] * dL_db + T.c[1][2] * T.c[1][2] * dL_dc);
        g.dcov[1] = 2 * T.c[0][0] * T.c[0][1] * dL_da + (T.c[0][0] * T.c[1][1] + T.c[0][1] * T.c[1][0]) * dL_db + 2 * T.c[1][0] * T.c[1][1] * dL_dc;
        g.dcov[2] = 2 * T.c[0][0] * T.c[0][2] * dL_da + (T.c[0][0] * T.c[1][2] + T.c[0][2] * T.c[1][0]) * dL_db + 2 * T.c[1][0] * T.c[1][2] * dL_dc;
        g.dcov[4] = 2 * T.c[0][2] * T.c[0][1] * dL_da + (T.c[0][1] * T.c[1][2] + T.c[0][2] * T.c[1][1]) * dL_db + 2 * T.c[1][1] * T.c[1][2] * dL_dc;
    } else {
        for (int i = 0; i < 6; i++) g.dcov[i] = 0;
    }
    // Sigma * rows of T
    const float V00 = cov6[0], V01 = cov6[1], V02 = cov6[2], V11 = cov6[3], V12 = cov6[4], V22 = cov6[5];
    const float s0x = T.c[0][0] * V00 + T.c[0][1] * V01 + T.c[0][2] * V02;
    const float s0y = T.c[0][0] * V01 + T.c[0][1] * V11 + T.c[0][2] * V12;
    const float s0z = T.c[0][0] * V02 + T.c[0][1] * V12 + T.c[0][2] * V22;
    const float s1x = T.c[1][0] * V00 + T.c[1][1] * V01 + T.c[1][2] * V02;
    const float s1y = T.c[1][0] * V01 + T.c[1][1] * V11 + T.c[1][2] * V12;
    const float s1z = T.c[1][0] * V02 + T.c[1][1] * V12 + T.c[1][2] * V22;
    g.dT[0][0] = 2 * s0x * dL_da + s1x * dL_db;
    g.dT[0][1] = 2 * s0y * dL_da + s1y * dL_db;
    g.dT[0][2] = 2 * s0z * dL_da + s1z * dL_db;
    g.dT[1][0] = 2 * s1x * dL_dc + s0x * dL_db;
    g.dT[1][1] = 2 * s1y * dL_dc + s0y * dL_db;
    g.dT[1][2] = 2 * s1z * dL_dc + s0z * dL_db;
    // T = W * J  ->  dL/dJ (non-zero entries); W.c[k][j] = view[4*j + k]
    float dL_dJ00 = view[0] * g.dT[0][0] + view[4] * g.dT[0][1] + view[8] * g.dT[0][2];
    float dL_dJ02 = view[2] * g.dT[0][0] + view[6] * g.dT[0][1] + view[10] * g.dT[0][2];
    float dL_dJ11 = view[1] * g.dT[1][0] + view[5] * g.dT[1][1] + view[9] * g.dT[1][2];
    float dL_dJ12 = view[2] * g.dT[1][0] + view[6] * g.dT[1][1] + view[10] * g.dT[1][2];
    float tz = 1.f / t.z;
    float tz2 = tz * tz;
    float tz3 = tz2 * tz;
    g.dt.x = x_grad_mul * -h_x * tz2 * dL_dJ02;
    g.dt.y = y_grad_mul * -h_y * tz2 * dL_dJ12;
    g.dt.z = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * t.x) * tz3 * dL_dJ02 + (2 * h_y * t.y) * tz3 * dL_dJ12;
    // mean -> t is the 4x3 view transform: dL/dmean = W^T dL/dt
    g.dmean.x = view[0] * g.dt.x + view[1] * g.dt.y + view[2] * g.dt.z;
    g.dmean.y = view[4] * g.dt.x + view[5] * g.dt.y + view[6] * g.dt.z;
    g.dmean.z = view[8] * g.dt.x + view[9] * g.dt.y + view[10] * g.dt.z;
    return g;
}

// dL/d(scale), dL/d(raw quaternion) from dL/dcov3D[6] (CR/backward.cu:278-341).
GSR_HD void cov3d_backward(const V3& scale, float mod, const V4& q, const float* dcov, V3& dscale, V4& dq)
{
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    // R in glm column-major: Rg.c[col][row]
    M3 R = m3_cols(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
                   2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
                   2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
    const float sx = mod * scale.x, sy = mod * scale.y, sz = mod * scale.z;
    M3 S = m3_cols(sx, 0.f, 0.f, 0.f, sy, 0.f, 0.f, 0.f, sz);
    M3 M = m3_mul(S, R);
    M3 dSig = m3_cols(dcov[0], 0.5f * dcov[1], 0.5f * dcov[2],
                      0.5f * dcov[1], dcov[3], 0.5f * dcov[4],
                      0.5f * dcov[2], 0.5f * dcov[4], dcov[5]);
    M3 dM = m3_mul(M, dSig);
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 3; i++) dM.c[j][i] *= 2.0f;
    M3 Rt = m3_transpose(R);
    M3 dMt = m3_transpose(dM);
    dscale.x = Rt.c[0][0] * dMt.c[0][0] + Rt.c[0][1] * dMt.c[0][1] + Rt.c[0][2] * dMt.c[0][2];
    dscale.y = Rt.c[1][0] * dMt.c[1][0] + Rt.c[1][1] * dMt.c[1][1] + Rt.c[1][2] * dMt.c[1][2];
    dscale.z = Rt.c[2][0] * dMt.c[2][0] + Rt.c[2][1] * dMt.c[2][1] + Rt.c[2][2] * dMt.c[2][2];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        dMt.c[0][i] *= sx;
        dMt.c[1][i] *= sy;
        dMt.c[2][i] *= sz;
    }
    dq.x = 2 * z * (dMt.c[0][1] - dMt.c[1][0]) + 2 * y * (dMt.c[2][0] - dMt.c[0][2]) + 2 * x * (dMt.c[1][2] - dMt.c[2][1]);
    dq.y = 2 * y * (dMt.c[1][0] + dMt.c[0][1]) + 2 * z * (dMt.c[2][0] + dMt.c[0][2]) + 2 * r * (dMt.c[1][2] - dMt.c[2][1]) - 4 * x * (dMt.c[2][2] + dMt.c[1][1]);
    dq.z = 2 * x * (dMt.c[1][0] + dMt.c[0][1]) + 2 * r * (dMt.c[2][0] - dMt.c[0][2]) + 2 * z * (dMt.c[1][2] + dMt.c[2][1]) - 4 * y * (dMt.c[2][2] + dMt.c[0][0]);
    dq.w = 2 * r * (dMt.c[0][1] - dMt.c[1][0]) + 2 * x * (dMt.c[2][0] + dMt.c[0][2]) + 2 * y * (dMt.c[1][2] + dMt.c[2][1]) - 4 * z * (dMt.c[1][1] + dMt.c[0][0]);
}

// d(normalize(v))/dv applied to dv (CR/auxiliary.h:106-116).
GSR_HD V3 dnormvdv(const V3& v, const V3& dv)
{
    float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
    float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
    V3 o;
    o.x = ((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * invsum32;
    o.y = (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * invsum32;
    o.z = (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * invsum32;
    return o;
}

// SH backward for one Gaussian: writes dL/dsh (M float3) and returns dL/d(dir) (pre-normalisation
// chain is applied by the caller).  dRGB must already be clamp-masked.  (CR/backward.cu:20-139)
GSR_HD V3 sh_backward(int deg, const float* sh, const V3& dir, const float* dRGB, float* dsh)
{
#define SHV(k, ch) sh[3 * (k) + (ch)]
#define DSH(k, ch) dsh[3 * (k) + (ch)]
    const float x = dir.x, y = dir.y, z = dir.z;
    float ddx = 0.f, ddy = 0.f, ddz = 0.f;
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        const float g = dRGB[ch];
        float dx = 0.f, dy = 0.f, dz = 0.f;
        DSH(0, ch) = GSR_SH_C0 * g;
        if (deg > 0) {
            DSH(1, ch) = (-GSR_SH_C1 * y) * g;
            DSH(2, ch) = (GSR_SH_C1 * z) * g;
            DSH(3, ch) = (-GSR_SH_C1 * x) * g;
            dx = -GSR_SH_C1 * SHV(3, ch);
            dy = -GSR_SH_C1 * SHV(1, ch);
            dz = GSR_SH_C1 * SHV(2, ch);
            if (deg > 1) {
                float xx = x * x, yy = y * y, zz = z * z;
                float xy = x * y, yz = y * z, xz = x * z;
                DSH(4, ch) = (GSR_SH_C2_0 * xy) * g;
                DSH(5, ch) = (GSR_SH_C2_1 * yz) * g;
                DSH(6, ch) = (GSR_SH_C2_2 * (2.f * zz - xx - yy)) * g;
                DSH(7, ch) = (GSR_SH_C2_3 * xz) * g;
                DSH(8, ch) = (GSR_SH_C2_4 * (xx - yy)) * g;
                dx += GSR_SH_C2_0 * y * SHV(4, ch) + GSR_SH_C2_2 * 2.f * -x * SHV(6, ch) + GSR_SH_C2_3 * z * SHV(7, ch) + GSR_SH_C2_4 * 2.f * x * SHV(8, ch);
                dy += GSR_SH_C2_0 * x * SHV(4, ch) + GSR_SH_C2_1 * z * SHV(5, ch) + GSR_SH_C2_2 * 2.f * -y * SHV(6, ch) + GSR_SH_C2_4 * 2.f * -y * SHV(8, ch);
                dz += GSR_SH_C2_1 * y * SHV(5, ch) + GSR_SH_C2_2 * 2.f * 2.f * z * SHV(6, ch) + GSR_SH_C2_3 * x * SHV(7, ch);
                if (deg > 2) {
                    DSH(9, ch) = (GSR_SH_C3_0 * y * (3.f * xx - yy)) * g;
                    DSH(10, ch) = (GSR_SH_C3_1 * xy * z) * g;
                    DSH(11, ch) = (GSR_SH_C3_2 * y * (4.f * zz - xx - yy)) * g;
                    DSH(12, ch) = (GSR_SH_C3_3 * z * (2.f * zz - 3.f * xx - 3.f * yy)) * g;
                    DSH(13, ch) = (GSR_SH_C3_4 * x * (4.f * zz - xx - yy)) * g;
                    DSH(14, ch) = (GSR_SH_C3_5 * z * (xx - yy)) * g;
                    DSH(15, ch) = (GSR_SH_C3_6 * x * (xx - 3.f * yy)) * g;
                    dx += (GSR_SH_C3_0 * SHV(9, ch) * 3.f * 2.f * xy + GSR_SH_C3_1 * SHV(10, ch) * yz +
                           GSR_SH_C3_2 * SHV(11, ch) * -2.f * xy + GSR_SH_C3_3 * SHV(12, ch) * -3.f * 2.f * xz +
                           GSR_SH_C3_4 * SHV(13, ch) * (-3.f * xx + 4.f * zz - yy) + GSR_SH_C3_5 * SHV(14, ch) * 2.f * xz +
                           GSR_SH_C3_6 * SHV(15, ch) * 3.f * (xx - yy));
                    dy += (GSR_SH_C3_0 * SHV(9, ch) * 3.f * (xx - yy) + GSR_SH_C3_1 * SHV(10, ch) * xz +
                           GSR_SH_C3_2 * SHV(11, ch) * (-3.f * yy + 4.f * zz - xx) + GSR_SH_C3_3 * SHV(12, ch) * -3.f * 2.f * yz +
                           GSR_SH_C3_4 * SHV(13, ch) * -2.f * xy + GSR_SH_C3_5 * SHV(14, ch) * -2.f * yz +
                           GSR_SH_C3_6 * SHV(15, ch) * -3.f * 2.f * xy);
                    dz += (GSR_SH_C3_1 * SHV(10, ch) * xy + GSR_SH_C3_2 * SHV(11, ch) * 4.f * 2.f * yz +
                           GSR_SH_C3_3 * SHV(12, ch) * 3.f * (2.f * zz - xx - yy) + GSR_SH_C3_4 * SHV(13, ch) * 4.f * 2.f * xz +
                           GSR_SH_C3_5 * SHV(14, ch) * (xx - yy));
                }
            }
        }
        ddx += dx * g;
        ddy += dy * g;
        ddz += dz * g;
    }
#undef SHV
#undef DSH
    V3 o = {ddx, ddy, ddz};
    return o;
}

}  // namespace gsr
