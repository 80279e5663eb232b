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

}  // namespace gsr
