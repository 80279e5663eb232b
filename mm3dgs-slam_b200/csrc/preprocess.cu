// preprocess.cu — per-Gaussian forward kernels (the backward lives in preprocess_bwd.cu).
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
#include "preprocess_common.cuh"
#include "gsr_cull.cuh"

namespace gsr {

// ----------------------------------------------------------------------------------------------
// Forward
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPB, 5) k_preprocess_fwd(PreArgs a)
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
        if (a.extra_gen != nullptr) {   // SLAM depth / silhouette colours (z, 1, z^2); R/slam/renderer.py:26-43
            const float z = radius > 0 ? o.depth : 0.f;
            a.extra_gen[(size_t)idx * 3 + 0] = z;
            a.extra_gen[(size_t)idx * 3 + 1] = radius > 0 ? 1.f : 0.f;
            a.extra_gen[(size_t)idx * 3 + 2] = z * z;
        }
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

}  // namespace gsr
