// gsr_bwd_math.cuh — per-Gaussian BACKWARD math of k_preprocess_bwd, derived in matrix form.
//
// What it computes is what the reference's computeCov2DCUDA + preprocessCUDA backward compute
// (CR/backward.cu:144-274, :346-396, :20-139, :278-341; gradient conventions in the comments below), but
// none of it is evaluated the reference's way.  Nothing here feeds an integer output, so the bit-exactness
// contract of gsr_math.cuh does not apply and the structure of the problem is used instead:
//
//   * the conic K = Q^-1 is already in the forward's record, so dL/dQ = -K G K is two 2x2 products on
//     stored values — the 2-D covariance (a, b, c), its determinant and the 3x3 sandwich that produces
//     them are never recomputed;
//   * J has four non-zeros, so M = J W (rows m0, m1: the screen-space axes in world space) is two
//     scaled row sums of the view rotation;
//   * Sigma = L L^T with L = R diag(s): with r_i = R^T m_i (the screen axes in the Gaussian's frame),
//     y_i = s o r_i and z = H y, everything downstream is rank two:
//         dL/ds_j = 2 (z0_j r0_j + z1_j r1_j)
//         dL/dR   = 2 (m0 (s o z0)^T + m1 (s o z1)^T)       (never materialised beyond its 9 entries)
//         dL/dM_i = 2 R (s o z_i)                            (Sigma u_i = L L^T u_i = L z_i)
//     so neither Sigma nor dL/dSigma is formed on the (scale, rotation) path;
//   * the quaternion gradient is written from u^T R(q) w = u.w - 2[...] + 2[...] - 2 r q_v.(u x w)
//     in terms of the diagonal, the symmetric part and the axial vector of dL/dR;
//   * SH: the basis gradients are evaluated once per coefficient (not once per channel):
//     ddir = sum_k grad(b_k) * (sh_k . dRGB); degree 0 has no direction dependence at all.
//
// All functions are GSR_HD: tests/hostcheck runs them on the CPU against the oracle.
#pragma once
#include "gsr_math.cuh"

#if defined(__CUDA_ARCH__)
#define GSR_RCP(x) __frcp_rn(x)
#define GSR_RSQRT(x) rsqrtf(x)
#else
#define GSR_RCP(x) (1.0f / (x))
#define GSR_RSQRT(x) (1.0f / sqrtf(x))
#endif

namespace gsr {

GSR_HD float dot3(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
GSR_HD V3 lin2(float a, const V3& u, float b, const V3& v)
{
    V3 r = {a * u.x + b * v.x, a * u.y + b * v.y, a * u.z + b * v.z};
    return r;
}
GSR_HD V3 had(const V3& a, const V3& b)
{
    V3 r = {a.x * b.x, a.y * b.y, a.z * b.z};
    return r;
}

// The camera-side frame of one Gaussian: view-space mean (guard-band clamped like the forward,
// CR/forward.cu:82-88), the four non-zero Jacobian entries and the two rows of M = J W.
struct ViewFrame {
    V3 w0, w1, w2;            // rows of the view rotation W (W[k][c] = view[4c + k])
    V3 m0, m1;                // rows of M = J W
    float j00, j02, j11, j12;
    float tx, ty, iz;         // clamped view-space x, y; 1 / z
    float mask_x, mask_y;     // 0 where the guard band clamped (no gradient flows to t.x / t.y through J)
};

GSR_HD ViewFrame view_frame(const V3& p, const float* view, float fx, float fy, float tanfovx, float tanfovy)
{
    ViewFrame f;
    f.w0.x = view[0]; f.w0.y = view[4]; f.w0.z = view[8];
    f.w1.x = view[1]; f.w1.y = view[5]; f.w1.z = view[9];
    f.w2.x = view[2]; f.w2.y = view[6]; f.w2.z = view[10];
    const float tx = dot3(f.w0, p) + view[12], ty = dot3(f.w1, p) + view[13], tz = dot3(f.w2, p) + view[14];
    const float limx = 1.3f * tanfovx, limy = 1.3f * tanfovy;
    const float ux = tx / tz, uy = ty / tz;          // exact quotients: they decide the clamp masks
    f.mask_x = (ux < -limx || ux > limx) ? 0.f : 1.f;
    f.mask_y = (uy < -limy || uy > limy) ? 0.f : 1.f;
    f.tx = fminf(limx, fmaxf(-limx, ux)) * tz;
    f.ty = fminf(limy, fmaxf(-limy, uy)) * tz;
    f.iz = GSR_RCP(tz);
    f.j00 = fx * f.iz;
    f.j11 = fy * f.iz;
    f.j02 = -f.j00 * f.tx * f.iz;
    f.j12 = -f.j11 * f.ty * f.iz;
    f.m0 = lin2(f.j00, f.w0, f.j02, f.w2);
    f.m1 = lin2(f.j11, f.w1, f.j12, f.w2);
    return f;
}

// dL/dQ for the symmetric 2x2 covariance Q = [[a, b], [b, c]] from the gradient of its inverse (the conic):
// H = -K G K * D^2 / (D^2 + 1e-7), K = conic, G = [[gA, gB], [gB, gC]].  The reference's dL_dconic.y is half of
// the true off-diagonal gradient (CR/backward.cu:536-538), i.e. exactly the symmetric split G needs, and its
// 1 / (D^2 + 1e-7) regularisation (CR/backward.cu:205) is reproduced through det K = 1 / D.
// Returns H00 = dL/da, H11 = dL/dc and H01 = dL/db / 2.
struct Sym2 { float xx, xy, yy; };
GSR_HD Sym2 cov2d_grad_from_conic(float kx, float ky, float kz, float gA, float gB, float gC)
{
    const float n00 = kx * gA + ky * gB, n01 = kx * gB + ky * gC;
    const float n10 = ky * gA + kz * gB, n11 = ky * gB + kz * gC;
    const float detk = kx * kz - ky * ky;
    const float f = -GSR_RCP(1.0f + 0.0000001f * detk * detk);
    Sym2 h;
    h.xx = f * (n00 * kx + n01 * ky);
    h.xy = f * (n00 * ky + n01 * kz);
    h.yy = f * (n10 * ky + n11 * kz);
    return h;
}

// Columns-as-rows view of R(q) for the raw quaternion q = (r, x, y, z) (CR/forward.cu:129-137 builds the same matrix).
struct Rot3 { V3 r0, r1, r2; };   // rows of the textbook R
GSR_HD Rot3 rot_from_quat(const V4& q)
{
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z, rx = r * x, ry = r * y, rz = r * z;
    Rot3 R;
    R.r0.x = 1.f - 2.f * (yy + zz); R.r0.y = 2.f * (xy - rz);       R.r0.z = 2.f * (xz + ry);
    R.r1.x = 2.f * (xy + rz);       R.r1.y = 1.f - 2.f * (xx + zz); R.r1.z = 2.f * (yz - rx);
    R.r2.x = 2.f * (xz - ry);       R.r2.y = 2.f * (yz + rx);       R.r2.z = 1.f - 2.f * (xx + yy);
    return R;
}
GSR_HD V3 rot_t_mul(const Rot3& R, const V3& v)   // R^T v
{
    V3 o = {R.r0.x * v.x + R.r1.x * v.y + R.r2.x * v.z, R.r0.y * v.x + R.r1.y * v.y + R.r2.y * v.z,
            R.r0.z * v.x + R.r1.z * v.y + R.r2.z * v.z};
    return o;
}
GSR_HD V3 rot_mul(const Rot3& R, const V3& v)     // R v
{
    V3 o = {dot3(R.r0, v), dot3(R.r1, v), dot3(R.r2, v)};
    return o;
}

// (scale, rotation) path: dL/ds_eff (the reference differentiates with respect to mod*scale and does NOT
// multiply by mod, CR/backward.cu:319-322), dL/d(raw quaternion), and dL/dM rows for the view chain.
GSR_HD void cov_chain_scale_rot(const ViewFrame& f, const Sym2& H, const V3& s_eff, const V4& q, V3& dscale, V4& dq,
                                V3& dM0, V3& dM1)
{
    const Rot3 R = rot_from_quat(q);
    const V3 r0 = rot_t_mul(R, f.m0), r1 = rot_t_mul(R, f.m1);     // screen axes in the Gaussian's frame
    const V3 y0 = had(s_eff, r0), y1 = had(s_eff, r1);
    const V3 z0 = lin2(H.xx, y0, H.xy, y1), z1 = lin2(H.xy, y0, H.yy, y1);
    dscale.x = 2.f * (z0.x * r0.x + z1.x * r1.x);
    dscale.y = 2.f * (z0.y * r0.y + z1.y * r1.y);
    dscale.z = 2.f * (z0.z * r0.z + z1.z * r1.z);
    const V3 v0 = had(s_eff, z0), v1 = had(s_eff, z1);
    dM0 = rot_mul(R, v0);
    dM1 = rot_mul(R, v1);
    dM0.x *= 2.f; dM0.y *= 2.f; dM0.z *= 2.f;
    dM1.x *= 2.f; dM1.y *= 2.f; dM1.z *= 2.f;
    // A = m0 v0^T + m1 v1^T = dL/dR / 2
    const float a00 = f.m0.x * v0.x + f.m1.x * v1.x, a01 = f.m0.x * v0.y + f.m1.x * v1.y, a02 = f.m0.x * v0.z + f.m1.x * v1.z;
    const float a10 = f.m0.y * v0.x + f.m1.y * v1.x, a11 = f.m0.y * v0.y + f.m1.y * v1.y, a12 = f.m0.y * v0.z + f.m1.y * v1.z;
    const float a20 = f.m0.z * v0.x + f.m1.z * v1.x, a21 = f.m0.z * v0.y + f.m1.z * v1.y, a22 = f.m0.z * v0.z + f.m1.z * v1.z;
    const float p01 = a01 + a10, p02 = a02 + a20, p12 = a12 + a21;     // symmetric part
    const float cx = a12 - a21, cy = a20 - a02, cz = a01 - a10;        // axial vector of the antisymmetric part
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    dq.x = -4.f * (x * cx + y * cy + z * cz);
    dq.y = 4.f * (y * p01 + z * p02 - r * cx) - 8.f * x * (a11 + a22);
    dq.z = 4.f * (x * p01 + z * p12 - r * cy) - 8.f * y * (a00 + a22);
    dq.w = 4.f * (x * p02 + y * p12 - r * cz) - 8.f * z * (a00 + a11);
}

// Precomputed-covariance path: dL/dcov3D in the reference's 6-vector convention (diagonal entries plain,
// off-diagonal entries doubled, CR/backward.cu:216-228) and dL/dM rows.
GSR_HD void cov_chain_precomp(const ViewFrame& f, const Sym2& H, const float* cov6, float* dcov, V3& dM0, V3& dM1)
{
    const V3 u0 = lin2(H.xx, f.m0, H.xy, f.m1), u1 = lin2(H.xy, f.m0, H.yy, f.m1);   // U = H M
    dcov[0] = f.m0.x * u0.x + f.m1.x * u1.x;
    dcov[3] = f.m0.y * u0.y + f.m1.y * u1.y;
    dcov[5] = f.m0.z * u0.z + f.m1.z * u1.z;
    dcov[1] = 2.f * (f.m0.x * u0.y + f.m1.x * u1.y);
    dcov[2] = 2.f * (f.m0.x * u0.z + f.m1.x * u1.z);
    dcov[4] = 2.f * (f.m0.y * u0.z + f.m1.y * u1.z);
    const float s00 = cov6[0], s01 = cov6[1], s02 = cov6[2], s11 = cov6[3], s12 = cov6[4], s22 = cov6[5];
    dM0.x = 2.f * (s00 * u0.x + s01 * u0.y + s02 * u0.z);
    dM0.y = 2.f * (s01 * u0.x + s11 * u0.y + s12 * u0.z);
    dM0.z = 2.f * (s02 * u0.x + s12 * u0.y + s22 * u0.z);
    dM1.x = 2.f * (s00 * u1.x + s01 * u1.y + s02 * u1.z);
    dM1.y = 2.f * (s01 * u1.x + s11 * u1.y + s12 * u1.z);
    dM1.z = 2.f * (s02 * u1.x + s12 * u1.y + s22 * u1.z);
}

// M = J W -> dL/dJ (four entries) -> dL/dt (view-space mean, through J only).
GSR_HD V3 view_chain(const ViewFrame& f, const V3& dM0, const V3& dM1, float fx, float fy)
{
    const float dj00 = dot3(dM0, f.w0), dj02 = dot3(dM0, f.w2);
    const float dj11 = dot3(dM1, f.w1), dj12 = dot3(dM1, f.w2);
    const float iz2 = f.iz * f.iz;
    const float ax = -fx * iz2, ay = -fy * iz2;     // dJ02/dtx, dJ12/dty (and dJ00/dtz, dJ11/dtz)
    V3 dt;
    dt.x = f.mask_x * ax * dj02;
    dt.y = f.mask_y * ay * dj12;
    dt.z = ax * dj00 + ay * dj11 - 2.f * f.iz * (ax * f.tx * dj02 + ay * f.ty * dj12);
    return dt;
}

// W^T dt: gradient of the world-space mean through t = V [p; 1].
GSR_HD V3 view_t_mul(const ViewFrame& f, const V3& dt)
{
    V3 o = {f.w0.x * dt.x + f.w1.x * dt.y + f.w2.x * dt.z, f.w0.y * dt.x + f.w1.y * dt.y + f.w2.y * dt.z,
            f.w0.z * dt.x + f.w1.z * dt.y + f.w2.z * dt.z};
    return o;
}

// Screen position: ndc = hom.xy / (hom.w + 1e-7), hom = Proj [p; 1].  (gx, gy) = dL/d(ndc) (the reference's
// dL_dmean2D is already scaled to NDC units, CR/backward.cu:528-529).  Returns dL/dp; dh = dL/d(hom.x, hom.y, hom.w).
GSR_HD V3 ndc_chain(const V3& p, const float* proj, float gx, float gy, V3& dh)
{
    const float hx = proj[0] * p.x + proj[4] * p.y + proj[8] * p.z + proj[12];
    const float hy = proj[1] * p.x + proj[5] * p.y + proj[9] * p.z + proj[13];
    const float hw = proj[3] * p.x + proj[7] * p.y + proj[11] * p.z + proj[15];
    const float iw = GSR_RCP(hw + 0.0000001f);
    dh.x = iw * gx;
    dh.y = iw * gy;
    dh.z = -(hx * dh.x + hy * dh.y) * iw;
    V3 o = {proj[0] * dh.x + proj[1] * dh.y + proj[3] * dh.z, proj[4] * dh.x + proj[5] * dh.y + proj[7] * dh.z,
            proj[8] * dh.x + proj[9] * dh.y + proj[11] * dh.z};
    return o;
}

// SH colour backward.  dsh[k][ch] = b_k(dir) dRGB[ch]; returns dL/d(dir) = sum_k grad b_k(dir) (sh_k . dRGB).
// Basis functions b_k as in CR/auxiliary.h:21-39 / CR/forward.cu:20-71; dRGB is already clamp-masked.
GSR_HD V3 sh_grad(int deg, const float* sh, const V3& d, const float* g, float* dsh)
{
    V3 dd = {0.f, 0.f, 0.f};
    // b_k and grad b_k, one coefficient at a time; c = sh_k . dRGB couples it to the direction
#define GSR_SH_TERM(k, bk, gxk, gyk, gzk)                                                       \
    {                                                                                           \
        const float b_ = (bk);                                                                  \
        dsh[3 * (k)] = b_ * g[0]; dsh[3 * (k) + 1] = b_ * g[1]; dsh[3 * (k) + 2] = b_ * g[2];   \
        const float c_ = sh[3 * (k)] * g[0] + sh[3 * (k) + 1] * g[1] + sh[3 * (k) + 2] * g[2];  \
        dd.x += c_ * (gxk); dd.y += c_ * (gyk); dd.z += c_ * (gzk);                             \
    }
    dsh[0] = GSR_SH_C0 * g[0]; dsh[1] = GSR_SH_C0 * g[1]; dsh[2] = GSR_SH_C0 * g[2];
    if (deg < 1) return dd;
    const float x = d.x, y = d.y, z = d.z;
    GSR_SH_TERM(1, -GSR_SH_C1 * y, 0.f, -GSR_SH_C1, 0.f)
    GSR_SH_TERM(2, GSR_SH_C1 * z, 0.f, 0.f, GSR_SH_C1)
    GSR_SH_TERM(3, -GSR_SH_C1 * x, -GSR_SH_C1, 0.f, 0.f)
    if (deg < 2) return dd;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    GSR_SH_TERM(4, GSR_SH_C2_0 * xy, GSR_SH_C2_0 * y, GSR_SH_C2_0 * x, 0.f)
    GSR_SH_TERM(5, GSR_SH_C2_1 * yz, 0.f, GSR_SH_C2_1 * z, GSR_SH_C2_1 * y)
    GSR_SH_TERM(6, GSR_SH_C2_2 * (2.f * zz - xx - yy), GSR_SH_C2_2 * -2.f * x, GSR_SH_C2_2 * -2.f * y, GSR_SH_C2_2 * 4.f * z)
    GSR_SH_TERM(7, GSR_SH_C2_3 * xz, GSR_SH_C2_3 * z, 0.f, GSR_SH_C2_3 * x)
    GSR_SH_TERM(8, GSR_SH_C2_4 * (xx - yy), GSR_SH_C2_4 * 2.f * x, GSR_SH_C2_4 * -2.f * y, 0.f)
    if (deg < 3) return dd;
    const float q4 = 4.f * zz - xx - yy;
    GSR_SH_TERM(9, GSR_SH_C3_0 * y * (3.f * xx - yy), GSR_SH_C3_0 * 6.f * xy, GSR_SH_C3_0 * 3.f * (xx - yy), 0.f)
    GSR_SH_TERM(10, GSR_SH_C3_1 * xy * z, GSR_SH_C3_1 * yz, GSR_SH_C3_1 * xz, GSR_SH_C3_1 * xy)
    GSR_SH_TERM(11, GSR_SH_C3_2 * y * q4, GSR_SH_C3_2 * -2.f * xy, GSR_SH_C3_2 * (q4 - 2.f * yy), GSR_SH_C3_2 * 8.f * yz)
    GSR_SH_TERM(12, GSR_SH_C3_3 * z * (2.f * zz - 3.f * xx - 3.f * yy), GSR_SH_C3_3 * -6.f * xz, GSR_SH_C3_3 * -6.f * yz,
                GSR_SH_C3_3 * 3.f * (2.f * zz - xx - yy))
    GSR_SH_TERM(13, GSR_SH_C3_4 * x * q4, GSR_SH_C3_4 * (q4 - 2.f * xx), GSR_SH_C3_4 * -2.f * xy, GSR_SH_C3_4 * 8.f * xz)
    GSR_SH_TERM(14, GSR_SH_C3_5 * z * (xx - yy), GSR_SH_C3_5 * 2.f * xz, GSR_SH_C3_5 * -2.f * yz, GSR_SH_C3_5 * (xx - yy))
    GSR_SH_TERM(15, GSR_SH_C3_6 * x * (xx - 3.f * yy), GSR_SH_C3_6 * 3.f * (xx - yy), GSR_SH_C3_6 * -6.f * xy, 0.f)
#undef GSR_SH_TERM
    return dd;
}

// Gradient through dir = v / |v|: the component of ddir tangent to the unit sphere, divided by |v|.
GSR_HD V3 unit_vector_grad(const V3& dir, float inv_len, const V3& ddir)
{
    const float along = dot3(dir, ddir);
    V3 o = {(ddir.x - dir.x * along) * inv_len, (ddir.y - dir.y * along) * inv_len, (ddir.z - dir.z * along) * inv_len};
    return o;
}

}  // namespace gsr
