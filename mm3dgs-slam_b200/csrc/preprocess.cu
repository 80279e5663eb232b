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
#include "async_copy.cuh"
#include "preprocess_common.cuh"
#include "gsr_cull.cuh"

namespace gsr {

// ----------------------------------------------------------------------------------------------
// Forward
// ----------------------------------------------------------------------------------------------
// Persistent, TMA-fed: a CTA walks 256-Gaussian chunks; one thread asks the TMA unit for the next chunk's input
// slices (cp.async.bulk -> the other shared-memory stage, completion on an mbarrier) before the CTA starts the math
// of the current chunk, and the 12 KB of records leave as one bulk store.
struct __align__(128) FwdIn {    // every member a multiple of 16 bytes
    float means[kPB * 3];
    float scales[kPB * 3];
    float rots[kPB * 4];
    float col[kPB * 3];          // SH (M == 1) or precomputed colours
    float opac[kPB];
};
struct __align__(128) FwdOut {
    float4 rec[kPB * 3];
};
constexpr size_t kFwdSmem = 2 * sizeof(FwdIn) + 2 * sizeof(FwdOut) + 64;

__global__ void __launch_bounds__(kPB, 4) k_preprocess_fwd(PreArgs a, int nchunks, int bulk_ok)
{
    extern __shared__ __align__(128) unsigned char s_raw[];
    FwdIn* s_in = reinterpret_cast<FwdIn*>(s_raw);
    FwdOut* s_out = reinterpret_cast<FwdOut*>(s_raw + 2 * sizeof(FwdIn));
    uint64_t* s_full = reinterpret_cast<uint64_t*>(s_raw + 2 * sizeof(FwdIn) + 2 * sizeof(FwdOut));
    __shared__ float s_cam[35];
    __shared__ uint32_t s_tiles;
    __shared__ int s_chunk[2];                            // chunk index of each stage (tickets: dynamic scheduling)
    __shared__ uint32_t s_hist[kSortBins];                // this CTA's share of the depth sort's first digit histogram

    const int tid = threadIdx.x;
    const bool has_sr = (a.cov3D_pre == nullptr);
    const bool sh_path = (a.colors == nullptr);
    const bool col_staged = !sh_path || a.M == 1;
    if (tid == 0) {
        s_tiles = 0;
        mbar_init(&s_full[0], 1);
        mbar_init(&s_full[1], 1);
        fence_mbar_init();
    }
    if (tid < 16) s_cam[tid] = a.view[tid];
    else if (tid < 32) s_cam[tid] = a.proj[tid - 16];
    else if (tid < 35) s_cam[tid] = a.campos[tid - 32];
    for (int i = tid; i < kSortBins; i += kPB) s_hist[i] = 0u;

    const uint32_t stage_bytes = (uint32_t)sizeof(float) * kPB * (3 + 1) + (has_sr ? (uint32_t)sizeof(float) * kPB * 7 : 0u) +
                                 (col_staged ? (uint32_t)sizeof(float) * kPB * 3 : 0u);
    auto is_bulk = [&](int c) { return bulk_ok && (c + 1) * kPB <= a.P; };
    auto issue = [&](int c, int st) {   // tid 0 only
        FwdIn& in = s_in[st];
        const size_t b = (size_t)c * kPB;
        mbar_arrive_expect_tx(&s_full[st], stage_bytes);
        bulk_g2s(in.means, a.means + b * 3, sizeof(float) * kPB * 3, &s_full[st]);
        bulk_g2s(in.opac, a.opac + b, sizeof(float) * kPB, &s_full[st]);
        if (has_sr) {
            bulk_g2s(in.scales, a.scales + b * 3, sizeof(float) * kPB * 3, &s_full[st]);
            bulk_g2s(in.rots, a.rots + b * 4, sizeof(float) * kPB * 4, &s_full[st]);
        }
        if (col_staged) bulk_g2s(in.col, (sh_path ? a.shs : a.colors) + b * 3, sizeof(float) * kPB * 3, &s_full[st]);
    };
    // chunks are handed out by a ticket counter, one chunk ahead of the math: no CTA is left with a longer static share
    if (tid == 0) {
        const int c0 = (int)atomicAdd(a.chunk_ticket, 1u);
        s_chunk[0] = c0;
        if (c0 < nchunks && is_bulk(c0)) issue(c0, 0);
    }
    __syncthreads();
    {   // zero this CTA's slice of the depth sort's look-back state (consumed by the kernels that follow)
        uint4* z = reinterpret_cast<uint4*>(a.status);
        const size_t n4 = a.status_words >> 2;
        for (size_t i = (size_t)blockIdx.x * kPB + tid; i < n4; i += (size_t)gridDim.x * kPB) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }

    uint32_t my_tiles = 0;
    for (int it = 0;; it++) {
        const int st = it & 1;
        const int c = s_chunk[st];
        if (c >= nchunks) break;
        FwdIn& in = s_in[st];
        FwdOut& out = s_out[st];
        const int base = c * kPB;
        const int nb = min(kPB, a.P - base);
        if (tid == 0) {   // the other stage was last read in iteration it-1, which every thread left through two barriers
            const int nxt = (int)atomicAdd(a.chunk_ticket, 1u);
            s_chunk[st ^ 1] = nxt;
            if (nxt < nchunks && is_bulk(nxt)) issue(nxt, st ^ 1);
        }
        if (is_bulk(c)) {
            mbar_wait(&s_full[st], (uint32_t)((it >> 1) & 1));
        } else {   // ragged last chunk / unaligned arrays
            stage_in(a.means + (size_t)base * 3, in.means, nb * 3, tid);
            stage_in(a.opac + (size_t)base, in.opac, nb, tid);
            if (has_sr) {
                stage_in(a.scales + (size_t)base * 3, in.scales, nb * 3, tid);
                stage_in(a.rots + (size_t)base * 4, in.rots, nb * 4, tid);
            }
            if (col_staged) stage_in((sh_path ? a.shs : a.colors) + (size_t)base * 3, in.col, nb * 3, tid);
            __syncthreads();
        }

        const int idx = base + tid;
        float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0;
        uint32_t dkey = 0xffffffffu;   // depth key of a visible Gaussian
        if (a.raw && tid < nb) {
            // fused activations (gsr_gaussians.raw_params): exp / normalize / sigmoid applied to this thread's staged
            // parameters IN PLACE in shared memory, so the code below — whose floating-point association order is
            // pinned bit for bit against the reference — is the same instruction stream for both kinds of input
            float inv_norm;
            const V3 sc = act_exp3(V3{in.scales[3 * tid], in.scales[3 * tid + 1], in.scales[3 * tid + 2]});
            const V4 q = act_normalize4(V4{in.rots[4 * tid], in.rots[4 * tid + 1], in.rots[4 * tid + 2], in.rots[4 * tid + 3]},
                                        inv_norm);
            in.scales[3 * tid] = sc.x; in.scales[3 * tid + 1] = sc.y; in.scales[3 * tid + 2] = sc.z;
            in.rots[4 * tid] = q.x; in.rots[4 * tid + 1] = q.y; in.rots[4 * tid + 2] = q.z; in.rots[4 * tid + 3] = q.w;
            in.opac[tid] = act_sigmoid(in.opac[tid]);
        }
        if (tid < nb) {
            const float* view = s_cam;
            const float* proj = s_cam + 16;
            V3 p = {in.means[3 * tid], in.means[3 * tid + 1], in.means[3 * tid + 2]};
            float cov6[6];
            if (has_sr) {
                V3 sc = {in.scales[3 * tid], in.scales[3 * tid + 1], in.scales[3 * tid + 2]};
                V4 q = {in.rots[4 * tid], in.rots[4 * tid + 1], in.rots[4 * tid + 2], in.rots[4 * tid + 3]};
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
            const int radius = o.radius;
            my_tiles += (uint32_t)o.tiles;
            if (radius > 0) {
                float cr, cg, cb;
                int bits = 0;
                if (sh_path) {
                    V3 dir = {0.f, 0.f, 1.f};
                    if (a.D > 0) {   // degree 0 has no view dependence: skip the normalisation (three IEEE divisions + a root)
                        dir.x = p.x - s_cam[32];
                        dir.y = p.y - s_cam[33];
                        dir.z = p.z - s_cam[34];
                        float len = sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
                        dir.x = dir.x / len;
                        dir.y = dir.y / len;
                        dir.z = dir.z / len;
                    }
                    const float* sh = (a.M == 1) ? (in.col + 3 * tid) : (a.shs + (size_t)idx * a.M * 3);
                    V3 cc = sh_to_rgb(a.D, sh, dir);
                    bits = (cc.x < 0 ? 1 : 0) | (cc.y < 0 ? 2 : 0) | (cc.z < 0 ? 4 : 0);
                    cr = fmaxf(cc.x, 0.0f);
                    cg = fmaxf(cc.y, 0.0f);
                    cb = fmaxf(cc.z, 0.0f);
                } else {
                    cr = in.col[3 * tid];
                    cg = in.col[3 * tid + 1];
                    cb = in.col[3 * tid + 2];
                }
                const float opacity = in.opac[tid];
                r0 = make_float4(o.px, o.py, o.depth, cull_radius2(o.lam_max, opacity));
                r1 = make_float4(o.cx, o.cy, o.cz, opacity);
                // r2.w: power threshold for the sub-tile test, low 3 bits = SH clamp flags (read by the backward)
                r2 = make_float4(cr, cg, cb, __int_as_float((__float_as_int(cull_power(o.lam_max, opacity)) & ~7) | bits));
            }
            a.radii[idx] = radius;
            a.rects[idx] = radius > 0 ? make_ushort4((unsigned short)o.x0, (unsigned short)o.y0, (unsigned short)o.x1,
                                                     (unsigned short)o.y1)
                                      : make_ushort4(0, 0, 0, 0);
            if (radius > 0) dkey = __float_as_uint(o.depth);
            a.depth_keys[idx] = dkey;
            if (a.extra_gen != nullptr) {   // SLAM depth / silhouette colours (z, 1, z^2); R/slam/renderer.py:26-43
                const float z = radius > 0 ? o.depth : 0.f;
                a.extra_gen[(size_t)idx * 3 + 0] = z;
                a.extra_gen[(size_t)idx * 3 + 1] = radius > 0 ? 1.f : 0.f;
                a.extra_gen[(size_t)idx * 3 + 2] = z * z;
            }
        }
        // histogram of the depth sort's first digit (the low 9 bits of the normalised key: mantissa bits, spread over the
        // 512 counters), while the key is in a register; every sort pass builds the histogram of the NEXT digit itself
        if (dkey != 0xffffffffu) atomicAdd(&s_hist[(dkey - kSortKeyBase) & (uint32_t)(kSortBins - 1)], 1u);
        // the bulk store issued two iterations ago has finished reading this output stage
        if (tid == 0) bulk_wait_read<1>();
        __syncthreads();   // (also: every thread is done with the input stage)
        out.rec[3 * tid + 0] = r0;
        out.rec[3 * tid + 1] = r1;
        out.rec[3 * tid + 2] = r2;
        float4* dst = a.rec + (size_t)base * 3;
        if (is_bulk(c)) {
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
                bulk_s2g(dst, out.rec, sizeof(float4) * kPB * 3);
                bulk_commit();
            }
        } else {
            __syncthreads();
            for (int i = tid; i < nb * 3; i += kPB) dst[i] = out.rec[i];
        }
    }
    if (tid == 0) bulk_wait_read<0>();
    {   // R = total number of tile instances: warp reduce, one shared atomic per warp, one global atomic per CTA
        const uint32_t wsum = __reduce_add_sync(0xffffffffu, my_tiles);
        if ((tid & 31) == 0 && wsum) atomicAdd(&s_tiles, wsum);
        __syncthreads();
        if (tid == 0 && s_tiles) atomicAdd(a.num_rendered, s_tiles);
        for (int i = tid; i < kSortBins; i += kPB) {
            const uint32_t h = s_hist[i];
            if (h) atomicAdd(&a.ghist[i], h);
        }
    }
}

void launch_preprocess_fwd(const PreArgs& a, cudaStream_t s)
{
    if (a.P <= 0) return;
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaFuncSetAttribute(k_preprocess_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem);
        configured[dev] = true;
    }
    const int nchunks = (a.P + kPB - 1) / kPB;
    const int grid = min(nchunks, 4 * device_sm_count());
    auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const bool col_staged = a.colors != nullptr || a.M == 1;
    const int bulk_ok = al(a.means) && al(a.opac) && al(a.rec) && (a.cov3D_pre || (al(a.scales) && al(a.rots))) &&
                        (!col_staged || al(a.colors ? a.colors : a.shs));
    k_preprocess_fwd<<<grid, kPB, kFwdSmem, s>>>(a, nchunks, bulk_ok);
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
