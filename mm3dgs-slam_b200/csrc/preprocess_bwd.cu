// preprocess_bwd.cu — per-Gaussian backward kernel: conic / screen-position / colour gradients of the blend
// backward -> gradients of means, scales, rotations, SH (and optionally of the camera matrices), one pass over the
// Gaussians.  Replaces computeCov2DCUDA + backward preprocessCUDA (CR/backward.cu:144-274, :346-396).
//
// The math lives in gsr_bwd_math.cuh (derived in matrix form; it reads the conic the forward stored instead of
// re-deriving the 2-D covariance).  The kernel is a persistent, TMA-fed stream: every CTA walks 256-Gaussian chunks;
// one thread asks the TMA unit for the NEXT chunk's eight contiguous input slices (cp.async.bulk into the other
// shared-memory stage, completion counted on an mbarrier) before the CTA starts the math of the current one, and the
// four gradient slices leave as bulk stores — or, when accumulating into a gradient bucket, as bulk f32 reduce-adds
// executed by the memory system, so the accumulators are never read by the SM.
#include "async_copy.cuh"
#include "gsr_bwd_math.cuh"
#include "preprocess_common.cuh"

namespace gsr {

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct __align__(128) BwdIn {   // one stage of staged inputs; every member is a multiple of 16 bytes
    float4 rec[kPB * 3];        // the forward's records: (px, py, depth, cull r^2) (conic, opacity) (rgb, bits)
    float means[kPB * 3];
    float scales[kPB * 3];
    float rots[kPB * 4];
    float sh[kPB * 3];          // M == 1 only
    float g2[kPB * 3];          // dL/dmean2D
    float gc[kPB * 4];          // dL/dconic
    float gcol[kPB * 3];        // dL/dcolor
};
struct __align__(128) BwdOut {
    float means[kPB * 3];
    float scales[kPB * 3];
    float rots[kPB * 4];
    float sh[kPB * 3];
};
constexpr size_t kBwdSmem = 2 * sizeof(BwdIn) + 2 * sizeof(BwdOut) + 64;

template <bool CAM, bool ACC>
__global__ void __launch_bounds__(kPB, 2) k_preprocess_bwd(PreBwdArgs a, int nchunks, int bulk_ok)
{
    extern __shared__ __align__(128) unsigned char s_raw[];
    BwdIn* s_in = reinterpret_cast<BwdIn*>(s_raw);
    BwdOut* s_out = reinterpret_cast<BwdOut*>(s_raw + 2 * sizeof(BwdIn));
    uint64_t* s_full = reinterpret_cast<uint64_t*>(s_raw + 2 * sizeof(BwdIn) + 2 * sizeof(BwdOut));
    __shared__ float s_cam[35];
    __shared__ int s_chunk[2];   // chunk index of each stage (tickets: dynamic scheduling)

    const int tid = threadIdx.x;
    const bool has_sr = (a.scales != nullptr);
    const bool sh_path = (a.shs != nullptr);
    const bool sh1 = sh_path && a.M == 1;

    if (tid == 0) {
        mbar_init(&s_full[0], 1);
        mbar_init(&s_full[1], 1);
        fence_mbar_init();
    }
    if (tid < 16) s_cam[tid] = a.view[tid];
    else if (tid < 32) s_cam[tid] = a.proj[tid - 16];
    else if (tid < 35) s_cam[tid] = a.campos[tid - 32];
    __syncthreads();

    const uint32_t stage_bytes = (uint32_t)(sizeof(float4) * kPB * 3 + sizeof(float) * kPB * (3 + 3 + 4)) +
                                 (has_sr ? (uint32_t)sizeof(float) * kPB * 7 : 0u) +
                                 (sh1 ? (uint32_t)sizeof(float) * kPB * 3 : 0u) +
                                 (sh_path ? (uint32_t)sizeof(float) * kPB * 3 : 0u);
    // a chunk goes through the TMA unit when it is full and every array is 16-byte aligned (bulk_ok)
    auto is_bulk = [&](int c) { return bulk_ok && (c + 1) * kPB <= a.P; };
    auto issue = [&](int c, int st) {   // tid 0 only
        BwdIn& in = s_in[st];
        const size_t b = (size_t)c * kPB;
        mbar_arrive_expect_tx(&s_full[st], stage_bytes);
        bulk_g2s(in.rec, a.rec + b * 3, sizeof(float4) * kPB * 3, &s_full[st]);
        bulk_g2s(in.means, a.means + b * 3, sizeof(float) * kPB * 3, &s_full[st]);
        bulk_g2s(in.g2, a.dL_dmean2D + b * 3, sizeof(float) * kPB * 3, &s_full[st]);
        bulk_g2s(in.gc, a.dL_dconic + b * 4, sizeof(float) * kPB * 4, &s_full[st]);
        if (has_sr) {
            bulk_g2s(in.scales, a.scales + b * 3, sizeof(float) * kPB * 3, &s_full[st]);
            bulk_g2s(in.rots, a.rots + b * 4, sizeof(float) * kPB * 4, &s_full[st]);
        }
        if (sh1) bulk_g2s(in.sh, a.shs + b * 3, sizeof(float) * kPB * 3, &s_full[st]);
        if (sh_path) bulk_g2s(in.gcol, a.dL_dcolors + b * 3, sizeof(float) * kPB * 3, &s_full[st]);
    };

    float cam[CAM ? 35 : 1];
    if (CAM) {
#pragma unroll
        for (int k = 0; k < 35; k++) cam[k] = 0.f;
    }

    // Chunks are handed out by a ticket counter, one chunk ahead of the math.  Every CTA draws exactly one ticket past
    // the end, so a launch draws nchunks + gridDim.x tickets: atomicInc with that period leaves the counter at zero.
    const unsigned ticket_wrap = (unsigned)nchunks + gridDim.x - 1u;
    if (tid == 0) {
        const int c0 = (int)atomicInc(a.chunk_ticket, ticket_wrap);
        s_chunk[0] = c0;
        if (c0 < nchunks && is_bulk(c0)) issue(c0, 0);
    }
    __syncthreads();

    for (int it = 0;; it++) {
        const int st = it & 1;
        const int c = s_chunk[st];
        if (c >= nchunks) break;
        BwdIn& in = s_in[st];
        BwdOut& out = s_out[st];
        const int base = c * kPB;
        const int nb = min(kPB, a.P - base);
        if (tid == 0) {   // the other stage was last read in iteration it-1, which every thread left through two barriers
            const int nxt = (int)atomicInc(a.chunk_ticket, ticket_wrap);
            s_chunk[st ^ 1] = nxt;
            if (nxt < nchunks && is_bulk(nxt)) issue(nxt, st ^ 1);
        }
        if (is_bulk(c)) {
            mbar_wait(&s_full[st], (uint32_t)((it >> 1) & 1));
        } else {   // ragged last chunk / unaligned arrays: synchronous staging
            stage_in(reinterpret_cast<const float*>(a.rec + (size_t)base * 3), reinterpret_cast<float*>(in.rec), nb * 12, tid);
            stage_in(a.means + (size_t)base * 3, in.means, nb * 3, tid);
            stage_in(a.dL_dmean2D + (size_t)base * 3, in.g2, nb * 3, tid);
            stage_in(a.dL_dconic + (size_t)base * 4, in.gc, nb * 4, tid);
            if (has_sr) {
                stage_in(a.scales + (size_t)base * 3, in.scales, nb * 3, tid);
                stage_in(a.rots + (size_t)base * 4, in.rots, nb * 4, tid);
            }
            if (sh1) stage_in(a.shs + (size_t)base * 3, in.sh, nb * 3, tid);
            if (sh_path) stage_in(a.dL_dcolors + (size_t)base * 3, in.gcol, nb * 3, tid);
            __syncthreads();
        }

        const int idx = base + tid;
        V3 dmean = {0.f, 0.f, 0.f}, dscale = {0.f, 0.f, 0.f};
        V4 dq = {0.f, 0.f, 0.f, 0.f};
        float dsh0[3] = {0.f, 0.f, 0.f};
        float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        // a visible Gaussian carries a cull radius >= 0.5 (or inf / NaN); culled ones an all-zero record
        const float4 r0 = in.rec[3 * tid];
        const bool vis = (tid < nb) && (r0.w != 0.f);
        if (vis) {
            const float* view = s_cam;
            const float* proj = s_cam + 16;
            const V3 p = {in.means[3 * tid], in.means[3 * tid + 1], in.means[3 * tid + 2]};
            const float4 r1 = in.rec[3 * tid + 1];
            const ViewFrame f = view_frame(p, view, a.focal_x, a.focal_y, a.tanfovx, a.tanfovy);
            const Sym2 H = cov2d_grad_from_conic(r1.x, r1.y, r1.z, in.gc[4 * tid], in.gc[4 * tid + 1], in.gc[4 * tid + 3]);
            V3 dM0, dM1;
            if (has_sr) {
                V3 sc = {in.scales[3 * tid], in.scales[3 * tid + 1], in.scales[3 * tid + 2]};
                V4 q = {in.rots[4 * tid], in.rots[4 * tid + 1], in.rots[4 * tid + 2], in.rots[4 * tid + 3]};
                float inv_norm = 1.f;
                if (a.raw) {
                    sc = act_exp3(sc);
                    q = act_normalize4(q, inv_norm);
                }
                const V3 s_eff = {a.scale_mod * sc.x, a.scale_mod * sc.y, a.scale_mod * sc.z};
                cov_chain_scale_rot(f, H, s_eff, q, dscale, dq, dM0, dM1);
                if (a.raw) {
                    // s = exp(r): ds/dr = s;  q = r / |r|: dq/dr = (I - q q^T) / |r|
                    dscale.x *= sc.x; dscale.y *= sc.y; dscale.z *= sc.z;
                    const float qd = q.x * dq.x + q.y * dq.y + q.z * dq.z + q.w * dq.w;
                    dq.x = (dq.x - q.x * qd) * inv_norm; dq.y = (dq.y - q.y * qd) * inv_norm;
                    dq.z = (dq.z - q.z * qd) * inv_norm; dq.w = (dq.w - q.w * qd) * inv_norm;
                }
            } else {
                float cov6[6];
#pragma unroll
                for (int k = 0; k < 6; k++) cov6[k] = a.cov3D_pre[(size_t)idx * 6 + k];
                cov_chain_precomp(f, H, cov6, dcov, dM0, dM1);
            }
            V3 dt = view_chain(f, dM0, dM1, a.focal_x, a.focal_y);
            if (a.dL_dextra_gen != nullptr)   // generated colours (z, 1, z^2): dL/dz = dE0 + 2 z dE2, z = t.z
                dt.z += a.dL_dextra_gen[(size_t)idx * 3] + 2.f * r0.z * a.dL_dextra_gen[(size_t)idx * 3 + 2];
            dmean = view_t_mul(f, dt);
            V3 dh;
            const V3 dmp = ndc_chain(p, proj, in.g2[3 * tid], in.g2[3 * tid + 1], dh);
            dmean.x += dmp.x;
            dmean.y += dmp.y;
            dmean.z += dmp.z;
            if (CAM) {
                // t = V [p;1]: dV[r][c] += dt[r] pc[c] (flat c*4 + r);  M = J W: dW[k][c] = sum_i J[i][k] dM_i[c]
                const float pc[4] = {p.x, p.y, p.z, 1.f};
                const float dtv[3] = {dt.x, dt.y, dt.z};
                const float dm0[3] = {dM0.x, dM0.y, dM0.z}, dm1[3] = {dM1.x, dM1.y, dM1.z};
                const float dhv[3] = {dh.x, dh.y, dh.z};
#pragma unroll
                for (int cc = 0; cc < 4; cc++) {
#pragma unroll
                    for (int r = 0; r < 3; r++) cam[cc * 4 + r] += dtv[r] * pc[cc];
                    cam[16 + cc * 4 + 0] += dhv[0] * pc[cc];
                    cam[16 + cc * 4 + 1] += dhv[1] * pc[cc];
                    cam[16 + cc * 4 + 3] += dhv[2] * pc[cc];
                }
#pragma unroll
                for (int cc = 0; cc < 3; cc++) {
                    cam[cc * 4 + 0] += f.j00 * dm0[cc];
                    cam[cc * 4 + 1] += f.j11 * dm1[cc];
                    cam[cc * 4 + 2] += f.j02 * dm0[cc] + f.j12 * dm1[cc];
                }
            }
            if (sh_path) {
                const int bits = __float_as_int(in.rec[3 * tid + 2].w) & 7;
                float dRGB[3] = {in.gcol[3 * tid], in.gcol[3 * tid + 1], in.gcol[3 * tid + 2]};
                if (bits & 1) dRGB[0] = 0.f;
                if (bits & 2) dRGB[1] = 0.f;
                if (bits & 4) dRGB[2] = 0.f;
                if (a.D == 0) {   // degree 0: colour does not depend on the view direction
                    if (a.M == 1) {
                        dsh0[0] = GSR_SH_C0 * dRGB[0]; dsh0[1] = GSR_SH_C0 * dRGB[1]; dsh0[2] = GSR_SH_C0 * dRGB[2];
                    } else {
                        float* o = a.dL_dsh + (size_t)idx * a.M * 3;
#pragma unroll
                        for (int ch = 0; ch < 3; ch++) {
                            if (ACC) atomicAdd(&o[ch], GSR_SH_C0 * dRGB[ch]);
                            else o[ch] = GSR_SH_C0 * dRGB[ch];
                        }
                        if (!ACC)
                            for (int k = 3; k < a.M * 3; k++) o[k] = 0.f;
                    }
                } else {
                    const V3 d0 = {p.x - s_cam[32], p.y - s_cam[33], p.z - s_cam[34]};
                    const float inv_len = GSR_RSQRT(dot3(d0, d0));
                    const V3 dir = {d0.x * inv_len, d0.y * inv_len, d0.z * inv_len};
                    const float* sh = a.shs + (size_t)idx * a.M * 3;
                    const int used = (a.D + 1) * (a.D + 1) * 3;
                    float tmp[48];
                    const V3 ddir = sh_grad(a.D, sh, dir, dRGB, tmp);
                    float* o = a.dL_dsh + (size_t)idx * a.M * 3;
                    for (int k = 0; k < used; k++) {
                        if (ACC) atomicAdd(&o[k], tmp[k]);
                        else o[k] = tmp[k];
                    }
                    if (!ACC)
                        for (int k = used; k < a.M * 3; k++) o[k] = 0.f;   // coefficients above the active degree
                    const V3 dm = unit_vector_grad(dir, inv_len, ddir);
                    dmean.x += dm.x;
                    dmean.y += dm.y;
                    dmean.z += dm.z;
                    if (CAM) {
                        cam[32] -= dm.x;
                        cam[33] -= dm.y;
                        cam[34] -= dm.z;
                    }
                }
            }
        } else if (!ACC && tid < nb && sh_path && a.M != 1) {
            for (int k = 0; k < a.M * 3; k++) a.dL_dsh[(size_t)idx * a.M * 3 + k] = 0.f;
        }
        if (a.dL_dcov3D != nullptr && tid < nb) {
#pragma unroll
            for (int k = 0; k < 6; k++) {
                if (ACC) atomicAdd(&a.dL_dcov3D[(size_t)idx * 6 + k], dcov[k]);
                else a.dL_dcov3D[(size_t)idx * 6 + k] = dcov[k];
            }
        }
        if (a.raw && tid < nb) {   // o = sigmoid(x): do/dx = o (1 - o); culled Gaussians received no opacity gradient
            const float o = act_sigmoid(a.opac[idx]);
            const float g = a.dL_dopacity[idx] * o * (1.f - o);
            if (ACC) atomicAdd(&a.dL_dopacity_raw[idx], g);
            else a.dL_dopacity_raw[idx] = g;
        }

        // the bulk stores issued two iterations ago have finished reading this output stage
        if (tid == 0) bulk_wait_read<1>();
        __syncthreads();   // (also: every thread is done with the input stage)
        out.means[3 * tid] = dmean.x; out.means[3 * tid + 1] = dmean.y; out.means[3 * tid + 2] = dmean.z;
        if (has_sr) {
            out.scales[3 * tid] = dscale.x; out.scales[3 * tid + 1] = dscale.y; out.scales[3 * tid + 2] = dscale.z;
            out.rots[4 * tid] = dq.x; out.rots[4 * tid + 1] = dq.y; out.rots[4 * tid + 2] = dq.z; out.rots[4 * tid + 3] = dq.w;
        }
        if (sh1) {
            out.sh[3 * tid] = dsh0[0]; out.sh[3 * tid + 1] = dsh0[1]; out.sh[3 * tid + 2] = dsh0[2];
        }
        if (is_bulk(c)) {
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
                const size_t b = (size_t)base;
                if (ACC) {
                    bulk_s2g_add_f32(a.dL_dmeans3D + b * 3, out.means, sizeof(float) * kPB * 3);
                    if (has_sr) {
                        bulk_s2g_add_f32(a.dL_dscales + b * 3, out.scales, sizeof(float) * kPB * 3);
                        bulk_s2g_add_f32(a.dL_drots + b * 4, out.rots, sizeof(float) * kPB * 4);
                    }
                    if (sh1) bulk_s2g_add_f32(a.dL_dsh + b * 3, out.sh, sizeof(float) * kPB * 3);
                } else {
                    bulk_s2g(a.dL_dmeans3D + b * 3, out.means, sizeof(float) * kPB * 3);
                    if (has_sr) {
                        bulk_s2g(a.dL_dscales + b * 3, out.scales, sizeof(float) * kPB * 3);
                        bulk_s2g(a.dL_drots + b * 4, out.rots, sizeof(float) * kPB * 4);
                    }
                    if (sh1) bulk_s2g(a.dL_dsh + b * 3, out.sh, sizeof(float) * kPB * 3);
                }
                bulk_commit();
            }
        } else {
            __syncthreads();
            stage_out<ACC>(a.dL_dmeans3D + (size_t)base * 3, out.means, nb * 3, tid);
            if (has_sr) {
                stage_out<ACC>(a.dL_dscales + (size_t)base * 3, out.scales, nb * 3, tid);
                stage_out<ACC>(a.dL_drots + (size_t)base * 4, out.rots, nb * 4, tid);
            }
            if (sh1) stage_out<ACC>(a.dL_dsh + (size_t)base * 3, out.sh, nb * 3, tid);
        }
    }
    if (tid == 0) bulk_wait_read<0>();   // shared memory must outlive the stores that read it

    if (CAM) {
        // Camera gradients: 35 sums over all Gaussians, with heavy cancellation (the rotation terms).  Per-thread fp32
        // partials (a few Gaussians each) are folded in fp64: warp -> CTA -> one record per CTA, and the last CTA to
        // finish adds the records up in index order, so the result does not depend on the order CTAs ran in.
        double (*s_red)[35] = reinterpret_cast<double (*)[35]>(s_raw);     // the staging buffers are free now
        __syncthreads();
        const int w = tid >> 5, l = tid & 31;
#pragma unroll
        for (int k = 0; k < 35; k++) {
            double v = (double)cam[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (l == 0) s_red[w][k] = v;
        }
        __syncthreads();
        if (tid < 35) {
            double v = 0.0;
#pragma unroll
            for (int w2 = 0; w2 < kPB / 32; w2++) v += s_red[w2][tid];
            a.cam_partials[(size_t)blockIdx.x * 35 + tid] = v;
        }
        __threadfence();
        __syncthreads();
        __shared__ unsigned s_last;
        if (tid == 0) s_last = (atomicAdd(a.cam_done, 1u) == gridDim.x - 1) ? 1u : 0u;
        __syncthreads();
        if (s_last && tid < 35) {
            double v = 0.0;
            for (unsigned c = 0; c < gridDim.x; c++) v += __ldcg(&a.cam_partials[(size_t)c * 35 + tid]);
            float* dst = tid < 16 ? (a.dL_dview ? a.dL_dview + tid : nullptr)
                       : tid < 32 ? (a.dL_dproj ? a.dL_dproj + (tid - 16) : nullptr)
                                  : (a.dL_dcampos ? a.dL_dcampos + (tid - 32) : nullptr);
            if (dst != nullptr) *dst += (float)v;     // (acc) buffer: the caller's contents are kept
            if (tid == 0) *a.cam_done = 0u;           // ready for the next launch on this workspace
        }
    }
}

template <bool CAM, bool ACC>
static void launch_variant(const PreBwdArgs& a, int nchunks, int bulk_ok, int grid, cudaStream_t s)
{
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaFuncSetAttribute(k_preprocess_bwd<CAM, ACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem);
        configured[dev] = true;
    }
    k_preprocess_bwd<CAM, ACC><<<grid, kPB, kBwdSmem, s>>>(a, nchunks, bulk_ok);
}

void launch_preprocess_bwd(const PreBwdArgs& a, cudaStream_t s)
{
    if (a.P <= 0) return;
    const bool cam = a.dL_dview || a.dL_dproj || a.dL_dcampos;
    const int nchunks = (a.P + kPB - 1) / kPB;
    const int grid = min(min(nchunks, 2 * device_sm_count()), kCamPartialRows);
    auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const bool sh1 = a.shs != nullptr && a.M == 1;
    const int bulk_ok = al(a.rec) && al(a.means) && al(a.dL_dmean2D) && al(a.dL_dconic) && al(a.dL_dmeans3D) &&
                        (!a.scales || (al(a.scales) && al(a.rots) && al(a.dL_dscales) && al(a.dL_drots))) &&
                        (!sh1 || (al(a.shs) && al(a.dL_dsh))) && (!a.shs || al(a.dL_dcolors));
    if (cam && a.accumulate) launch_variant<true, true>(a, nchunks, bulk_ok, grid, s);
    else if (cam) launch_variant<true, false>(a, nchunks, bulk_ok, grid, s);
    else if (a.accumulate) launch_variant<false, true>(a, nchunks, bulk_ok, grid, s);
    else launch_variant<false, false>(a, nchunks, bulk_ok, grid, s);
}

}  // namespace gsr
