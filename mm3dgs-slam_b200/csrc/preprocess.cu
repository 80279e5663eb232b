// preprocess.cu — per-Gaussian forward and backward kernels (HBM-streaming stages).
//
// Replaces the reference's preprocessCUDA (CR/forward.cu:155-256), checkFrustum
// (CR/rasterizer_impl.cu:54-66), computeCov2DCUDA + backward preprocessCUDA
// (CR/backward.cu:144-274, :346-396) — the last two fused into one kernel.
//
// B200 design: these kernels are pure HBM streams (~100-150 B per Gaussian, ~300 FLOP), so the
// only thing that matters is moving each byte once with full-width transactions.  The AoS inputs
// ([P,3] / [P,4] fp32) are staged block-wide through shared memory with 128-bit loads
// (ld.global.nc, L1 no-allocate) and the outputs leave through shared memory with 128-bit stores;
// every thread then works on its own Gaussian out of conflict-free shared memory (stride-3 /
// stride-12 words).  One 48-byte record per Gaussian replaces the reference's five separate
// arrays, and cov3D is never written (the backward recomputes it from scale/rotation).
#include "gsr_internal.cuh"
#include "gsr_math.cuh"
#include <stdio.h>

namespace gsr {

constexpr int kPB = 256;  // Gaussians (threads) per block

__device__ __forceinline__ float4 ld_stream4(const float4* p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_stream1(const float* p)
{
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

// Copy n contiguous floats global -> shared with the widest aligned transactions available.
__device__ __forceinline__ void stage_in(const float* __restrict__ src, float* dst, int n, int tid)
{
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const int n4 = n >> 2;
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int i = tid; i < n4; i += kPB) d4[i] = ld_stream4(s4 + i);
        for (int i = (n4 << 2) + tid; i < n; i += kPB) dst[i] = ld_stream1(src + i);
    } else {
        for (int i = tid; i < n; i += kPB) dst[i] = ld_stream1(src + i);
    }
}
// Copy n contiguous floats shared -> global; ACC: add `old` (the accumulator's previous contents, staged into
// shared memory at kernel start so that the read overlaps the math instead of sitting in front of the store).
template <bool ACC>
__device__ __forceinline__ void stage_out(float* __restrict__ dst, const float* src, const float* old, int n, int tid)
{
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        const int n4 = n >> 2;
        const float4* s4 = reinterpret_cast<const float4*>(src);
        const float4* o4 = reinterpret_cast<const float4*>(old);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int i = tid; i < n4; i += kPB) {
            float4 v = s4[i];
            if (ACC) {
                const float4 o = o4[i];
                v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
            }
            d4[i] = v;
        }
        for (int i = (n4 << 2) + tid; i < n; i += kPB) dst[i] = ACC ? old[i] + src[i] : src[i];
    } else {
        for (int i = tid; i < n; i += kPB) dst[i] = ACC ? old[i] + src[i] : src[i];
    }
}

// Conservative squared radius (pixels^2) outside of which the splat cannot reach alpha >= 1/255:
//   alpha = o*exp(power) >= 1/255  =>  power >= -ln(255 o),   power <= -|d|^2 / (2 lam_max)
//   =>  |d|^2 <= 2 lam_max ln(255 o).
// 3% + 0.5 px^2 slack covers the rounding of the conic and of the power evaluation; for very large
// splats (lam_max > 1000 px^2) the relative error of the evaluated power is no longer negligible
// against that slack, so they are never culled (+inf).  Used only to SKIP work in the blend kernels;
// it never changes which (pixel, splat) pairs contribute.
__device__ __forceinline__ float cull_radius2(float lam_max, float opacity)
{
    if (!(lam_max <= 1000.f)) return __int_as_float(0x7f800000);
    const float L = fmaxf(logf(255.f * opacity), 0.f);
    return 2.06f * lam_max * L + 0.5f;   // NaN opacity -> NaN -> never culled (tests are !(d2 > r2))
}
// Threshold for the exact ellipse-vs-sub-tile test of the forward blend: the splat can only contribute
// where q(d) = -power(d) <= ln(255 o); 3% + 0.02 slack covers the rounding of the conic and of the
// power evaluation inside the circle above (|d|^2 <= 2.06*1000*5.6, conic entries <= 1/0.3).  The low
// 3 mantissa bits are overwritten with the SH clamp flags by the caller, hence the extra 1e-5.
__device__ __forceinline__ float cull_power(float lam_max, float opacity)
{
    if (!(lam_max <= 1000.f)) return __int_as_float(0x7f800000);
    const float L = fmaxf(logf(255.f * opacity), 0.f);
    return (1.03f * L + 0.02f) * 1.00001f;
}

// ----------------------------------------------------------------------------------------------
// Forward
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPB) k_preprocess_fwd(PreArgs a)
{
    __shared__ __align__(16) float s_means[kPB * 3];
    __shared__ __align__(16) float s_scales[kPB * 3];
    __shared__ __align__(16) float s_rots[kPB * 4];
    __shared__ __align__(16) float s_col[kPB * 3];   // SH (M==1) or precomputed colours
    __shared__ __align__(16) float4 s_rec[kPB * 3];
    __shared__ float s_cam[35];
    __shared__ uint32_t s_tiles;

    const int tid = threadIdx.x;
    const int base = blockIdx.x * kPB;
    const int nb = min(kPB, a.P - base);
    const bool has_sr = (a.cov3D_pre == nullptr);
    const bool sh_path = (a.colors == nullptr);
    if (tid == 0) s_tiles = 0;

    stage_in(a.means + (size_t)base * 3, s_means, nb * 3, tid);
    if (has_sr) {
        stage_in(a.scales + (size_t)base * 3, s_scales, nb * 3, tid);
        stage_in(a.rots + (size_t)base * 4, s_rots, nb * 4, tid);
    }
    if (!sh_path)
        stage_in(a.colors + (size_t)base * 3, s_col, nb * 3, tid);
    else if (a.M == 1)
        stage_in(a.shs + (size_t)base * 3, s_col, nb * 3, tid);
    if (tid < 16) s_cam[tid] = a.view[tid];
    else if (tid < 32) s_cam[tid] = a.proj[tid - 16];
    else if (tid < 35) s_cam[tid] = a.campos[tid - 32];
    __syncthreads();

    const int idx = base + tid;
    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0;
    int radius = 0, tiles = 0;
    if (tid < nb) {
        const float* view = s_cam;
        const float* proj = s_cam + 16;
        V3 p = {s_means[3 * tid], s_means[3 * tid + 1], s_means[3 * tid + 2]};
        float cov6[6];
        if (has_sr) {
            V3 sc = {s_scales[3 * tid], s_scales[3 * tid + 1], s_scales[3 * tid + 2]};
            V4 q = {s_rots[4 * tid], s_rots[4 * tid + 1], s_rots[4 * tid + 2], s_rots[4 * tid + 3]};
            cov3d_from_scale_rot(sc, a.scale_mod, q, cov6);
        } else {
#pragma unroll
            for (int k = 0; k < 6; k++) cov6[k] = a.cov3D_pre[(size_t)idx * 6 + k];
        }
        PreOut o = preprocess_one(p, cov6, view, proj, a.W, a.H, a.tanfovx, a.tanfovy, a.focal_x, a.focal_y,
                                  a.gx, a.gy);
        if (a.prefiltered) {
            V3 pv = xform4x3(p, view);
            if (pv.z <= 0.1f) {
                printf("gsrast_b200: Gaussian %d culled although prefiltered is set\n", idx);
                __trap();
            }
        }
        radius = o.radius;
        tiles = o.tiles;
        if (radius > 0) {
            float cr, cg, cb;
            int bits = 0;
            if (sh_path) {
                V3 dir = {p.x - s_cam[32], p.y - s_cam[33], p.z - s_cam[34]};
                float len = sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
                dir.x = dir.x / len;
                dir.y = dir.y / len;
                dir.z = dir.z / len;
                const float* sh = (a.M == 1) ? (s_col + 3 * tid) : (a.shs + (size_t)idx * a.M * 3);
                V3 c = sh_to_rgb(a.D, sh, dir);
                bits = (c.x < 0 ? 1 : 0) | (c.y < 0 ? 2 : 0) | (c.z < 0 ? 4 : 0);
                cr = fmaxf(c.x, 0.0f);
                cg = fmaxf(c.y, 0.0f);
                cb = fmaxf(c.z, 0.0f);
            } else {
                cr = s_col[3 * tid];
                cg = s_col[3 * tid + 1];
                cb = s_col[3 * tid + 2];
            }
            const float opacity = a.opac[idx];
            r0 = make_float4(o.px, o.py, o.depth, cull_radius2(o.lam_max, opacity));
            r1 = make_float4(o.cx, o.cy, o.cz, opacity);
            // r2.w: power threshold for the sub-tile test, low 3 bits = SH clamp flags (read by the backward)
            r2 = make_float4(cr, cg, cb, __int_as_float((__float_as_int(cull_power(o.lam_max, opacity)) & ~7) | bits));
        }
        a.radii[idx] = radius;
        a.rects[idx] = radius > 0 ? make_ushort4((unsigned short)o.x0, (unsigned short)o.y0, (unsigned short)o.x1,
                                                 (unsigned short)o.y1)
                                  : make_ushort4(0, 0, 0, 0);
        a.depth_keys[idx] = radius > 0 ? __float_as_uint(o.depth) : 0xffffffffu;
    }
    {   // R = total number of tile instances: warp reduce, one shared atomic per warp, one global per CTA
        const uint32_t wsum = __reduce_add_sync(0xffffffffu, (uint32_t)tiles);
        if ((tid & 31) == 0 && wsum) atomicAdd(&s_tiles, wsum);
    }
    s_rec[3 * tid + 0] = r0;
    s_rec[3 * tid + 1] = r1;
    s_rec[3 * tid + 2] = r2;
    __syncthreads();
    float4* out = a.rec + (size_t)base * 3;
    for (int i = tid; i < nb * 3; i += kPB) out[i] = s_rec[i];
    if (tid == 0 && s_tiles) atomicAdd(a.num_rendered, s_tiles);
}

void launch_preprocess_fwd(const PreArgs& a, cudaStream_t s)
{
    if (a.P <= 0) return;
    k_preprocess_fwd<<<(a.P + kPB - 1) / kPB, kPB, 0, s>>>(a);
}

// present[i] = near-plane test only (CR/auxiliary.h:139-164 with prefiltered=false).
__global__ void __launch_bounds__(kPB) k_mark_visible(int P, const float* __restrict__ means,
                                                      const float* __restrict__ view, uint8_t* __restrict__ present)
{
    __shared__ __align__(16) float s_means[kPB * 3];
    const int tid = threadIdx.x, base = blockIdx.x * kPB;
    const int nb = min(kPB, P - base);
    stage_in(means + (size_t)base * 3, s_means, nb * 3, tid);
    __syncthreads();
    if (tid < nb) {
        V3 p = {s_means[3 * tid], s_means[3 * tid + 1], s_means[3 * tid + 2]};
        float z = view[2] * p.x + view[6] * p.y + view[10] * p.z + view[14];
        present[base + tid] = (z <= 0.1f) ? 0 : 1;
    }
}
void launch_mark_visible(int P, const float* means, const float* view, const float* /*proj*/, uint8_t* present,
                         cudaStream_t s)
{
    if (P <= 0) return;
    k_mark_visible<<<(P + kPB - 1) / kPB, kPB, 0, s>>>(P, means, view, present);
}

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
        if (CAM) {
            // t = V[p;1]: dV[r][c] += dt[r] p[c]; flat index c*4 + r
            const float pc[4] = {m.x, m.y, m.z, 1.f};
            const float dtv[3] = {cg.dt.x, cg.dt.y, cg.dt.z};
#pragma unroll
            for (int c = 0; c < 4; c++)
#pragma unroll
                for (int r = 0; r < 3; r++) cam[c * 4 + r] += dtv[r] * pc[c];
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
