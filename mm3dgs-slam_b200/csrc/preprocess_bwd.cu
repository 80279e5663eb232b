// preprocess_bwd.cu — per-Gaussian backward kernel (cov2D backward + projection + SH + scale/rotation,
// fused; replaces computeCov2DCUDA + backward preprocessCUDA, CR/backward.cu:144-274, :346-396).
//
// Nothing in the backward feeds an integer output, so the bit-exactness contract of gsr_math.cuh does not
// apply here; the kernel nevertheless reuses the forward's helpers (one source of truth for the math).
#include "preprocess_common.cuh"

namespace gsr {

// ----------------------------------------------------------------------------------------------
// Backward (cov2D backward + projection + SH + scale/rotation, one pass over the Gaussians)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <bool CAM, bool ACC>
__global__ void __launch_bounds__(kPB) k_preprocess_bwd(PreBwdArgs a)
{
    __shared__ __align__(16) float s_means[kPB * 3];   // in: means      out: dL/dmeans3D
    __shared__ __align__(16) float s_scales[kPB * 3];  // in: scales     out: dL/dscales
    __shared__ __align__(16) float s_rots[kPB * 4];    // in: rotations  out: dL/drotations
    __shared__ __align__(16) float s_sh[kPB * 3];      // in: sh (M==1)  out: dL/dsh (M==1)
    __shared__ __align__(16) float s_g2[kPB * 3];      // in: dL/dmean2D [.,3]
    __shared__ __align__(16) float s_gc[kPB * 4];      // in: dL/dconic  [.,4]
    __shared__ __align__(16) float s_gcol[kPB * 3];    // in: dL/dcolor  [.,3]
    __shared__ float s_cam[35];
    __shared__ float s_red[CAM ? (kPB / 32) * 35 : 1];
    // ACC: previous contents of the gradient accumulators
    __shared__ __align__(16) float s_o_means[ACC ? kPB * 3 : 4];
    __shared__ __align__(16) float s_o_scales[ACC ? kPB * 3 : 4];
    __shared__ __align__(16) float s_o_rots[ACC ? kPB * 4 : 4];
    __shared__ __align__(16) float s_o_sh[ACC ? kPB * 3 : 4];

    const int tid = threadIdx.x;
    const int base = blockIdx.x * kPB;
    const int nb = min(kPB, a.P - base);
    const bool has_sr = (a.scales != nullptr);
    const bool sh_path = (a.shs != nullptr);

    stage_in(a.means + (size_t)base * 3, s_means, nb * 3, tid);
    if (has_sr) {
        stage_in(a.scales + (size_t)base * 3, s_scales, nb * 3, tid);
        stage_in(a.rots + (size_t)base * 4, s_rots, nb * 4, tid);
    }
    if (sh_path && a.M == 1) stage_in(a.shs + (size_t)base * 3, s_sh, nb * 3, tid);
    stage_in(a.dL_dmean2D + (size_t)base * 3, s_g2, nb * 3, tid);
    stage_in(a.dL_dconic + (size_t)base * 4, s_gc, nb * 4, tid);
    if (sh_path) stage_in(a.dL_dcolors + (size_t)base * 3, s_gcol, nb * 3, tid);
    if (ACC) {
        stage_in(a.dL_dmeans3D + (size_t)base * 3, s_o_means, nb * 3, tid);
        if (has_sr) {
            stage_in(a.dL_dscales + (size_t)base * 3, s_o_scales, nb * 3, tid);
            stage_in(a.dL_drots + (size_t)base * 4, s_o_rots, nb * 4, tid);
        }
        if (sh_path && a.M == 1) stage_in(a.dL_dsh + (size_t)base * 3, s_o_sh, nb * 3, tid);
    }
    if (tid < 16) s_cam[tid] = a.view[tid];
    else if (tid < 32) s_cam[tid] = a.proj[tid - 16];
    else if (tid < 35) s_cam[tid] = a.campos[tid - 32];
    __syncthreads();

    const int idx = base + tid;
    V3 dmean = {0.f, 0.f, 0.f}, dscale = {0.f, 0.f, 0.f};
    V4 dq = {0.f, 0.f, 0.f, 0.f};
    float dsh0[3] = {0.f, 0.f, 0.f};
    float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float cam[CAM ? 35 : 1];
    if (CAM) {
#pragma unroll
        for (int k = 0; k < 35; k++) cam[k] = 0.f;
    }
    bool vis = false;
    if (tid < nb) vis = a.radii[idx] > 0;
    if (vis) {
        const float* view = s_cam;
        const float* proj = s_cam + 16;
        V3 m = {s_means[3 * tid], s_means[3 * tid + 1], s_means[3 * tid + 2]};
        V3 sc = {0.f, 0.f, 0.f};
        V4 q = {0.f, 0.f, 0.f, 0.f};
        float cov6[6];
        if (has_sr) {
            sc.x = s_scales[3 * tid]; sc.y = s_scales[3 * tid + 1]; sc.z = s_scales[3 * tid + 2];
            q.x = s_rots[4 * tid]; q.y = s_rots[4 * tid + 1]; q.z = s_rots[4 * tid + 2]; q.w = s_rots[4 * tid + 3];
            cov3d_from_scale_rot(sc, a.scale_mod, q, cov6);  // same bits as the forward
        } else {
#pragma unroll
            for (int k = 0; k < 6; k++) cov6[k] = a.cov3D_pre[(size_t)idx * 6 + k];
        }
        // --- cov2D path
        Cov2DGrad cg = cov2d_backward(m, cov6, a.focal_x, a.focal_y, a.tanfovx, a.tanfovy, view, s_gc[4 * tid],
                                      s_gc[4 * tid + 1], s_gc[4 * tid + 3]);
#pragma unroll
        for (int k = 0; k < 6; k++) dcov[k] = cg.dcov[k];
        dmean = cg.dmean;
        // --- projection path (CR/backward.cu:369-387)
        V4 m_hom = xform4x4(m, proj);
        float m_w = 1.0f / (m_hom.w + 0.0000001f);
        float mul1 = (proj[0] * m.x + proj[4] * m.y + proj[8] * m.z + proj[12]) * m_w * m_w;
        float mul2 = (proj[1] * m.x + proj[5] * m.y + proj[9] * m.z + proj[13]) * m_w * m_w;
        const float g2x = s_g2[3 * tid], g2y = s_g2[3 * tid + 1];
        V3 dproj;
        dproj.x = (proj[0] * m_w - proj[3] * mul1) * g2x + (proj[1] * m_w - proj[3] * mul2) * g2y;
        dproj.y = (proj[4] * m_w - proj[7] * mul1) * g2x + (proj[5] * m_w - proj[7] * mul2) * g2y;
        dproj.z = (proj[8] * m_w - proj[11] * mul1) * g2x + (proj[9] * m_w - proj[11] * mul2) * g2y;
        dmean.x += dproj.x;
        dmean.y += dproj.y;
        dmean.z += dproj.z;
        float dz_extra = 0.f;
        if (a.dL_dextra_gen != nullptr) {   // generated colours (z, 1, z^2): dL/dz = dE0 + 2 z dE2, z = row 2 of view * [p;1]
            const float z = a.rec[(size_t)idx * 3].z;
            dz_extra = a.dL_dextra_gen[(size_t)idx * 3] + 2.f * z * a.dL_dextra_gen[(size_t)idx * 3 + 2];
            dmean.x += view[2] * dz_extra;
            dmean.y += view[6] * dz_extra;
            dmean.z += view[10] * dz_extra;
        }
        if (CAM) {
            // t = V[p;1]: dV[r][c] += dt[r] p[c]; flat index c*4 + r
            const float pc[4] = {m.x, m.y, m.z, 1.f};
            const float dtv[3] = {cg.dt.x, cg.dt.y, cg.dt.z};
#pragma unroll
            for (int c = 0; c < 4; c++)
#pragma unroll
                for (int r = 0; r < 3; r++) cam[c * 4 + r] += (dtv[r] + (r == 2 ? dz_extra : 0.f)) * pc[c];
            // A = J V3: dV3[r][c] = sum_i J[i][r] dA[i][c]
            const V3 t = cov2d_project(m, a.focal_x, a.focal_y, a.tanfovx, a.tanfovy, cov6, view).t;
            const float j00 = a.focal_x / t.z, j02 = -(a.focal_x * t.x) / (t.z * t.z);
            const float j11 = a.focal_y / t.z, j12 = -(a.focal_y * t.y) / (t.z * t.z);
#pragma unroll
            for (int c = 0; c < 3; c++) {
                cam[c * 4 + 0] += j00 * cg.dT[0][c];
                cam[c * 4 + 1] += j11 * cg.dT[1][c];
                cam[c * 4 + 2] += j02 * cg.dT[0][c] + j12 * cg.dT[1][c];
            }
            // hom = Proj[p;1]; ndc = hom.xy * m_w
            const float dhx = m_w * g2x, dhy = m_w * g2y;
            const float dhw = -(m_hom.x * g2x + m_hom.y * g2y) * m_w * m_w;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                cam[16 + c * 4 + 0] += dhx * pc[c];
                cam[16 + c * 4 + 1] += dhy * pc[c];
                cam[16 + c * 4 + 3] += dhw * pc[c];
            }
        }
        // --- SH path
        if (sh_path) {
            const int bits = __float_as_int(a.rec[(size_t)idx * 3 + 2].w) & 7;
            float dRGB[3] = {s_gcol[3 * tid], s_gcol[3 * tid + 1], s_gcol[3 * tid + 2]};
            dRGB[0] *= (bits & 1) ? 0.f : 1.f;
            dRGB[1] *= (bits & 2) ? 0.f : 1.f;
            dRGB[2] *= (bits & 4) ? 0.f : 1.f;
            V3 dir_orig = {m.x - s_cam[32], m.y - s_cam[33], m.z - s_cam[34]};
            float len = sqrtf(dir_orig.x * dir_orig.x + dir_orig.y * dir_orig.y + dir_orig.z * dir_orig.z);
            V3 dir = {dir_orig.x / len, dir_orig.y / len, dir_orig.z / len};
            V3 ddir;
            if (a.M == 1) {
                ddir = sh_backward(a.D, s_sh + 3 * tid, dir, dRGB, dsh0);
            } else if (!ACC) {
                ddir = sh_backward(a.D, a.shs + (size_t)idx * a.M * 3, dir, dRGB, a.dL_dsh + (size_t)idx * a.M * 3);
                // coefficients above the active degree keep a zero gradient
                for (int k = (a.D + 1) * (a.D + 1) * 3; k < a.M * 3; k++) a.dL_dsh[(size_t)idx * a.M * 3 + k] = 0.f;
            } else {
                float tmp[48];
                ddir = sh_backward(a.D, a.shs + (size_t)idx * a.M * 3, dir, dRGB, tmp);
                for (int k = 0; k < (a.D + 1) * (a.D + 1) * 3; k++) a.dL_dsh[(size_t)idx * a.M * 3 + k] += tmp[k];
            }
            V3 dm = dnormvdv(dir_orig, ddir);
            dmean.x += dm.x;
            dmean.y += dm.y;
            dmean.z += dm.z;
            if (CAM) {
                cam[32] -= dm.x;
                cam[33] -= dm.y;
                cam[34] -= dm.z;
            }
        }
        // --- scale / rotation
        if (has_sr) cov3d_backward(sc, a.scale_mod, q, dcov, dscale, dq);
    } else if (!ACC && tid < nb && sh_path && a.M != 1) {
        for (int k = 0; k < a.M * 3; k++) a.dL_dsh[(size_t)idx * a.M * 3 + k] = 0.f;
    }
    __syncthreads();  // everyone is done reading the staged inputs; reuse them for the outputs
    s_means[3 * tid] = dmean.x; s_means[3 * tid + 1] = dmean.y; s_means[3 * tid + 2] = dmean.z;
    if (has_sr) {
        s_scales[3 * tid] = dscale.x; s_scales[3 * tid + 1] = dscale.y; s_scales[3 * tid + 2] = dscale.z;
        s_rots[4 * tid] = dq.x; s_rots[4 * tid + 1] = dq.y; s_rots[4 * tid + 2] = dq.z; s_rots[4 * tid + 3] = dq.w;
    }
    if (sh_path && a.M == 1) {
        s_sh[3 * tid] = dsh0[0]; s_sh[3 * tid + 1] = dsh0[1]; s_sh[3 * tid + 2] = dsh0[2];
    }
    if (a.dL_dcov3D != nullptr && tid < nb) {
#pragma unroll
        for (int k = 0; k < 6; k++) {
            if (ACC) a.dL_dcov3D[(size_t)idx * 6 + k] += dcov[k];
            else a.dL_dcov3D[(size_t)idx * 6 + k] = dcov[k];
        }
    }
    if (CAM) {
        const int w = tid >> 5, l = tid & 31;
#pragma unroll
        for (int k = 0; k < 35; k++) {
            float v = warp_sum(cam[k]);
            if (l == 0) s_red[w * 35 + k] = v;
        }
    }
    __syncthreads();
    stage_out<ACC>(a.dL_dmeans3D + (size_t)base * 3, s_means, s_o_means, nb * 3, tid);
    if (has_sr) {
        stage_out<ACC>(a.dL_dscales + (size_t)base * 3, s_scales, s_o_scales, nb * 3, tid);
        stage_out<ACC>(a.dL_drots + (size_t)base * 4, s_rots, s_o_rots, nb * 4, tid);
    }
    if (sh_path && a.M == 1) stage_out<ACC>(a.dL_dsh + (size_t)base * 3, s_sh, s_o_sh, nb * 3, tid);
    if (CAM && tid < 35) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < kPB / 32; w++) v += s_red[w * 35 + tid];
        float* dst = tid < 16 ? (a.dL_dview ? a.dL_dview + tid : nullptr)
                   : tid < 32 ? (a.dL_dproj ? a.dL_dproj + (tid - 16) : nullptr)
                              : (a.dL_dcampos ? a.dL_dcampos + (tid - 32) : nullptr);
        if (dst != nullptr && v != 0.f) atomicAdd(dst, v);
    }
}

void launch_preprocess_bwd(const PreBwdArgs& a, cudaStream_t s)
{
    if (a.P <= 0) return;
    const bool cam = a.dL_dview || a.dL_dproj || a.dL_dcampos;
    const int grid = (a.P + kPB - 1) / kPB;
    if (cam && a.accumulate) k_preprocess_bwd<true, true><<<grid, kPB, 0, s>>>(a);
    else if (cam) k_preprocess_bwd<true, false><<<grid, kPB, 0, s>>>(a);
    else if (a.accumulate) k_preprocess_bwd<false, true><<<grid, kPB, 0, s>>>(a);
    else k_preprocess_bwd<false, false><<<grid, kPB, 0, s>>>(a);
}

}  // namespace gsr
