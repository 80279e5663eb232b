// render.cu — per-tile alpha compositing, forward and backward.
//
// Replaces renderCUDA forward (CR/forward.cu:261-374) and backward (CR/backward.cu:399-557).
// Per-pixel semantics are identical (same skip rules, same sequential blend order, same
// `contributor` bookkeeping); what changes is how the work is organised for sm_100a:
//
//  * one CTA per 16x16 tile, 8 warps, each warp owns an 8x4-pixel sub-tile so that the set of
//    splats touching a warp is compact;
//  * the tile's splat list is staged 256 records at a time into shared memory — one 48-byte
//    record per splat (position, conic, opacity, colour) instead of the reference's three global
//    arrays plus a per-(pixel,splat) colour gather from global memory;
//  * each warp first culls the staged batch against its sub-tile with a conservative circle test
//    (one splat per lane, __ballot_sync), then walks only the surviving bits — splats that cannot
//    reach alpha >= 1/255 anywhere in the sub-tile are never evaluated.  The test is conservative,
//    so exactly the same (pixel, splat) pairs contribute as in the reference;
//  * early termination is per warp (__all_sync on the running transmittance), the CTA leaves when
//    all 8 warps are done;
//  * the forward records, per list entry, which of the tile's 8 warps blended it for at least one
//    pixel (one byte per entry);
//  * backward: the traversal starts at the CTA-wide last contributor instead of the end of the
//    list and each warp visits exactly the entries it blended in the forward (no culling test, no
//    wasted evaluation); the 9 per-splat gradient terms are reduced over the warp with a transposing
//    butterfly (14 shuffles instead of 45) and leave as ONE fire-and-forget RED.ADD.F32 per
//    (warp, splat, term) instead of one atomicAdd per (pixel, splat, term).
#include "async_copy.cuh"
#include "gsr_internal.cuh"
#include "gsr_cull.cuh"

namespace gsr {

constexpr int kFwdMinCtas = 8;   // resident CTAs per SM the forward blend is compiled for (6 with the extra channels)
constexpr int kBwdMinCtas = 5;   // backward blend (4 with the extra channels); 6 measured slower (spills)
constexpr int kBatch = 256;
constexpr unsigned kFull = 0xffffffffu;

struct TileGeom {
    int px, py;       // this lane's pixel
    float sx0, sx1;   // sub-tile pixel-centre extent
    float sy0, sy1;
    bool inside;
};

__device__ __forceinline__ TileGeom tile_geom(int tile, int gx, int W, int H)
{
    TileGeom g;
    const int tx = tile % gx, ty = tile / gx;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ox = tx * kTile + (warp & 1) * 8, oy = ty * kTile + (warp >> 1) * 4;
    g.px = ox + (lane & 7);
    g.py = oy + (lane >> 3);
    g.sx0 = (float)ox;
    g.sx1 = (float)(ox + 7);
    g.sy0 = (float)oy;
    g.sy1 = (float)(oy + 3);
    g.inside = g.px < W && g.py < H;
    return g;
}

// ----------------------------------------------------------------------------------------------
// Forward
// ----------------------------------------------------------------------------------------------
// EXTRA: three more colour channels per splat (`extra` [P,3], e.g. the SLAM renderer's depth /
// silhouette colours [z, 1, z^2]) are blended in the same pass into `out_extra` [3,H,W] — one
// preprocess + one sort + one traversal instead of the reference's second full rasterizer call
// (R/slam/renderer.py:196-214).
template <bool EXTRA>
__global__ void __launch_bounds__(256, EXTRA ? 6 : kFwdMinCtas) k_render_fwd(
    int W, int H, int gx, const uint2* __restrict__ ranges,
    const uint32_t* __restrict__ point_list, const float4* __restrict__ rec, const float* __restrict__ bg,
    float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, float* __restrict__ out_color,
    uint8_t* __restrict__ contrib, const float* __restrict__ extra, float* __restrict__ out_extra,
    const uint32_t* __restrict__ counters, uint32_t cap)
{
    // two stages of 256 staged records: batch k+1 is gathered by cp.async (LDGSTS, 3 x 16 B per record) while the
    // warps blend batch k; one block barrier per batch
    __shared__ float4 s_r0[2][kBatch];  // px, py, depth, cull r^2
    __shared__ float4 s_r1[2][kBatch];  // conic xyz, opacity
    __shared__ float4 s_r2[2][kBatch];  // rgb, bits
    __shared__ float s_ex[2][EXTRA ? kBatch * 3 : 1];
    __shared__ uint32_t s_mask[2][kBatch / 32][8];   // [stage][32-entry group][warp]: entries this warp blended

    if (counters[kCntR] > cap) return;   // the instance list did not fit the caller's workspace: the host re-runs the call
    const int tile = ((int)blockIdx.x);
    const TileGeom g = tile_geom(tile, gx, W, H);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float pixx = (float)g.px, pixy = (float)g.py;
    const uint2 range = ranges[tile];
    const int n = (int)(range.y - range.x);

    float T = 1.0f;
    float C0 = 0.f, C1 = 0.f, C2 = 0.f;
    float E0 = 0.f, E1 = 0.f, E2 = 0.f;
    uint32_t last_contributor = 0;
    bool done = !g.inside;
    bool warp_done = __all_sync(kFull, done);

    // gather of one batch into stage `st`: every thread copies the record of one list entry
    auto gather = [&](int i, int st, uint32_t id) {
        if (i < n) {
            const float4* r = rec + (size_t)id * 3;
            cp_async16(&s_r0[st][threadIdx.x], r);
            cp_async16(&s_r1[st][threadIdx.x], r + 1);
            cp_async16(&s_r2[st][threadIdx.x], r + 2);
            if (EXTRA) {
                const float* e = extra + (size_t)id * 3;
                cp_async4(&s_ex[st][3 * threadIdx.x + 0], e);
                cp_async4(&s_ex[st][3 * threadIdx.x + 1], e + 1);
                cp_async4(&s_ex[st][3 * threadIdx.x + 2], e + 2);
            }
        }
        cp_async_commit();
    };
    const uint32_t* my_list = point_list + range.x + threadIdx.x;
    gather((int)threadIdx.x, 0, (int)threadIdx.x < n ? my_list[0] : 0u);
    uint32_t id_next = (kBatch + (int)threadIdx.x < n) ? my_list[kBatch] : 0u;   // ids run one batch ahead of the gathers

    int buf = 0;
    for (int base = 0;; base += kBatch) {
        cp_async_wait<0>();
        // this thread's copies of batch `base` have landed; after the barrier everybody's have, and every warp has
        // finished batch base-kBatch (stage buf^1 is free)
        const bool all_done = __syncthreads_and(warp_done);
        if (base > 0) {
            // contribution byte of the previous batch's entry t: bit w set <=> warp w blended it for some pixel.
            // The backward visits exactly these (warp, entry) pairs and nothing else.
            const int t = threadIdx.x, pbase = base - kBatch;
            if (pbase + t < n) {
                const uint32_t* m = s_mask[buf ^ 1][t >> 5];
                uint32_t byte = 0;
#pragma unroll
                for (int w = 0; w < 8; w++) byte |= ((m[w] >> (t & 31)) & 1u) << w;
                contrib[range.x + pbase + t] = (uint8_t)byte;
            }
        }
        if (base >= n || all_done) break;
        if (base + kBatch < n) {
            gather(base + kBatch + (int)threadIdx.x, buf ^ 1, id_next);
            id_next = (base + 2 * kBatch + (int)threadIdx.x < n) ? my_list[base + 2 * kBatch] : 0u;
        }
        const int cur = buf;
        buf ^= 1;
        if (lane < kBatch / 32) s_mask[cur][lane][warp] = 0u;   // groups this warp does not reach stay "not blended"
        __syncwarp();
        if (warp_done) continue;
        const float4* r0s = s_r0[cur];
        const float4* r1s = s_r1[cur];
        const float4* r2s = s_r2[cur];
        const int cnt = min(kBatch, n - base);
        for (int k = 0; k < cnt; k += 32) {
            bool hit = false;
            if (k + lane < cnt) {
                const float4 a = r0s[k + lane];
                hit = subtile_hit(g.sx0, g.sx1, g.sy0, g.sy1, a.x, a.y, a.w);
                if (hit) {
                    const float4 co = r1s[k + lane];
                    hit = subtile_hit_ellipse(g.sx0, g.sx1, g.sy0, g.sy1, a.x, a.y, co.x, co.y, co.z, r2s[k + lane].w);
                }
            }
            unsigned m = __ballot_sync(kFull, hit);
            unsigned blended = 0u;
            while (m) {
                const int bit = __ffs(m) - 1;
                const int j = k + bit;
                m &= m - 1;
                const float4 a = r0s[j];
                const float4 co = r1s[j];
                const float dx = a.x - pixx, dy = a.y - pixy;
                const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
                const float alpha = fminf(0.99f, co.w * expf(power));
                const float test_T = T * (1 - alpha);
                const bool live = !done && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
                if (live) {
                    if (test_T < 0.0001f) {
                        done = true;
                    } else {
                        const float4 c = r2s[j];
                        const float w = alpha * T;
                        C0 += c.x * w;
                        C1 += c.y * w;
                        C2 += c.z * w;
                        if (EXTRA) {
                            E0 += s_ex[cur][3 * j + 0] * w;
                            E1 += s_ex[cur][3 * j + 1] * w;
                            E2 += s_ex[cur][3 * j + 2] * w;
                        }
                        T = test_T;
                        last_contributor = (uint32_t)(base + j + 1);
                        blended |= 1u << bit;
                    }
                }
            }
            blended = __reduce_or_sync(kFull, blended);
            if (lane == 0) s_mask[cur][k >> 5][warp] = blended;
            if (__all_sync(kFull, done)) {
                warp_done = true;
                break;
            }
        }
    }
    if (g.inside) {
        const int pix = g.py * W + g.px;
        final_T[pix] = T;
        n_contrib[pix] = last_contributor;
        const size_t HW = (size_t)H * W;
        out_color[pix] = C0 + T * bg[0];
        out_color[HW + pix] = C1 + T * bg[1];
        out_color[2 * HW + pix] = C2 + T * bg[2];
        if (EXTRA) {
            out_extra[pix] = E0 + T * bg[0];
            out_extra[HW + pix] = E1 + T * bg[1];
            out_extra[2 * HW + pix] = E2 + T * bg[2];
        }
    }
}

void launch_render_fwd(int W, int H, int gx, int gy, const uint2* ranges, const uint32_t* point_list,
                       const float4* rec, const float* bg, float* final_T, uint32_t* n_contrib, float* out_color,
                       uint8_t* contrib, const float* extra, float* out_extra, const uint32_t* counters, uint32_t cap,
                       cudaStream_t s)
{
    if (extra != nullptr)
        k_render_fwd<true><<<gx * gy, 256, 0, s>>>(W, H, gx, ranges, point_list, rec, bg, final_T, n_contrib,
                                                   out_color, contrib, extra, out_extra, counters, cap);
    else
        k_render_fwd<false><<<gx * gy, 256, 0, s>>>(W, H, gx, ranges, point_list, rec, bg, final_T, n_contrib,
                                                    out_color, contrib, nullptr, nullptr, counters, cap);
}

// ----------------------------------------------------------------------------------------------
// Backward
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

template <bool EXTRA>
__global__ void __launch_bounds__(256, EXTRA ? 4 : kBwdMinCtas) k_render_bwd(int W, int H, int gx, const uint2* __restrict__ ranges,
                                                    const uint32_t* __restrict__ point_list,
                                                    const float4* __restrict__ rec, const float* __restrict__ bg,
                                                    const float* __restrict__ final_T,
                                                    const uint32_t* __restrict__ n_contrib,
                                                    const uint8_t* __restrict__ contrib,
                                                    const float* __restrict__ dL_dpix, float* __restrict__ dL_dmean2D,
                                                    float* __restrict__ dL_dconic, float* __restrict__ dL_dopacity,
                                                    float* __restrict__ dL_dcolors, const float* __restrict__ extra,
                                                    const float* __restrict__ dL_dpix_extra,
                                                    float* __restrict__ dL_dextra)
{
    // two stages: batch k+1 (the entries somebody blended) is gathered by cp.async while the warps work on batch k
    __shared__ float4 s_r0[2][kBatch];
    __shared__ float4 s_r1[2][kBatch];
    __shared__ float4 s_r2[2][kBatch];
    __shared__ float s_ex[2][EXTRA ? kBatch * 3 : 1];
    __shared__ uint32_t s_id[2][kBatch];
    __shared__ uint8_t s_cb[2][kBatch];     // forward's contribution byte: bit w <=> warp w blended this entry
    __shared__ uint32_t s_max;

    const int tile = ((int)blockIdx.x);
    const TileGeom g = tile_geom(tile, gx, W, H);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float pixx = (float)g.px, pixy = (float)g.py;
    const uint2 range = ranges[tile];
    const int pix = g.py * W + g.px;
    const size_t HW = (size_t)H * W;

    const float T_final = g.inside ? final_T[pix] : 0.f;
    float T = T_final;
    const uint32_t last_contributor = g.inside ? n_contrib[pix] : 0u;
    float dLp0 = 0.f, dLp1 = 0.f, dLp2 = 0.f;
    if (g.inside) {
        dLp0 = dL_dpix[pix];
        dLp1 = dL_dpix[HW + pix];
        dLp2 = dL_dpix[2 * HW + pix];
    }
    float dLe0 = 0.f, dLe1 = 0.f, dLe2 = 0.f;
    if (EXTRA && g.inside) {
        dLe0 = dL_dpix_extra[pix];
        dLe1 = dL_dpix_extra[HW + pix];
        dLe2 = dL_dpix_extra[2 * HW + pix];
    }
    const float bg_dot_dpixel = bg[0] * (dLp0 + dLe0) + bg[1] * (dLp1 + dLe1) + bg[2] * (dLp2 + dLe2);
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;          // accum_rec
    float lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, last_alpha = 0.f;  // last colour / alpha
    float acc3 = 0.f, acc4 = 0.f, acc5 = 0.f, lc3 = 0.f, lc4 = 0.f, lc5 = 0.f;   // same for the extra channels
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
    // role of this lane after the butterfly: lane 4k holds term k (k < 8), lane 1 holds term 8
    float* red_base = nullptr;
    int red_stride = 0;
    {
        const int k = (lane == 1) ? 8 : (((lane & 3) == 0) ? (lane >> 2) : -1);
        if (k == 0 || k == 1) { red_base = dL_dmean2D + k; red_stride = 3; }
        else if (k >= 2 && k <= 4) { red_base = dL_dconic + (k == 4 ? 3 : k - 2); red_stride = 4; }
        else if (k == 5) { red_base = dL_dopacity; red_stride = 1; }
        else if (k >= 6) { red_base = dL_dcolors + (k - 6); red_stride = 3; }
    }
    // EXTRA: second 4-term butterfly (colour b + 3 extra channels) ends in lanes 0, 8, 16, 24
    float* red2_base = nullptr;
    if (EXTRA && (lane & 7) == 0) red2_base = (lane == 0) ? dL_dcolors + 2 : dL_dextra + ((lane >> 3) - 1);

    const uint32_t wmax = __reduce_max_sync(kFull, last_contributor);
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
    if (lane == 0) atomicMax(&s_max, wmax);
    __syncthreads();
    const int bmax = (int)s_max;   // CTA-wide last contributor: nothing behind it matters

    // staged slot t of the batch that ends at list position `hi` holds position hi-1-t (back to front).  The
    // (contribution byte, Gaussian id) pair of a slot is fetched one batch ahead of its record gather.
    auto fetch = [&](int hi, uint8_t& cb, uint32_t& id) {
        cb = 0;
        id = 0;
        const int t = threadIdx.x;
        if (hi > 0 && t < min(kBatch, hi)) {
            const uint32_t p = range.x + (uint32_t)(hi - 1 - t);
            cb = contrib[p];
            if (cb) id = point_list[p];   // entries nobody blended are never visited: skip their gather
        }
    };
    auto gather = [&](int st, uint8_t cb, uint32_t id) {
        const int t = threadIdx.x;
        s_cb[st][t] = cb;
        if (cb) {
            const float4* r = rec + (size_t)id * 3;
            s_id[st][t] = id;
            cp_async16(&s_r0[st][t], r);
            cp_async16(&s_r1[st][t], r + 1);
            cp_async16(&s_r2[st][t], r + 2);
            if (EXTRA) {
                const float* e = extra + (size_t)id * 3;
                cp_async4(&s_ex[st][3 * t + 0], e);
                cp_async4(&s_ex[st][3 * t + 1], e + 1);
                cp_async4(&s_ex[st][3 * t + 2], e + 2);
            }
        }
        cp_async_commit();
    };
    uint8_t cb_next;
    uint32_t id_next;
    fetch(bmax, cb_next, id_next);
    gather(0, cb_next, id_next);
    fetch(bmax - kBatch, cb_next, id_next);

    int buf = 0;
    for (int hi = bmax; hi > 0; hi -= kBatch) {
        const int cnt = min(kBatch, hi);
        cp_async_wait<0>();
        __syncthreads();   // batch `hi` has landed for everybody; every warp is done with the other stage
        if (hi - kBatch > 0) {
            gather(buf ^ 1, cb_next, id_next);
            fetch(hi - 2 * kBatch, cb_next, id_next);
        }
        const int cur = buf;
        buf ^= 1;
        const float4* r0s = s_r0[cur];
        const float4* r1s = s_r1[cur];
        const float4* r2s = s_r2[cur];
        // first slot this warp cares about: position < wmax  <=>  t > hi-1-wmax
        int t0 = hi - (int)wmax;
        if (t0 < 0) t0 = 0;
        for (int k = (t0 & ~31); k < cnt; k += 32) {
            const int tt = k + lane;
            const bool hit = (tt < cnt) && ((s_cb[cur][tt] >> warp) & 1u);   // this warp blended it in the forward
            unsigned m = __ballot_sync(kFull, hit);
            while (m) {
                const int b = __ffs(m) - 1;
                const int j = k + b;
                m &= m - 1;
                const uint32_t pos = (uint32_t)(hi - 1 - j);   // 0-based list position
                const float4 a = r0s[j];
                const float4 co = r1s[j];
                const float dx = a.x - pixx, dy = a.y - pixy;
                const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
                // The backward only has to meet the 1e-4 gradient band (the forward stays exact: n_contrib / final_T are
                // compared bit for bit), so it takes the two-instruction exponential and the one-instruction reciprocal
                // (ex2.approx / rcp.approx, ~2 ulp) instead of the accurate sequences: 16 of ~125 instructions per visit.
                const float G = __expf(power);
                const float alpha = fminf(0.99f, co.w * G);
                const bool live = (pos < last_contributor) && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
                float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f, v5 = 0.f, v6 = 0.f, v7 = 0.f, v8 = 0.f;
                float v9 = 0.f, v10 = 0.f, v11 = 0.f;
                if (live) {
                    const float4 c = r2s[j];
                    const float inv_1ma = __fdividef(1.f, 1.f - alpha);
                    T = T * inv_1ma;
                    const float dchannel_dcolor = alpha * T;
                    float dL_dalpha = 0.0f;
                    acc0 = last_alpha * lc0 + (1.f - last_alpha) * acc0;
                    lc0 = c.x;
                    dL_dalpha += (c.x - acc0) * dLp0;
                    acc1 = last_alpha * lc1 + (1.f - last_alpha) * acc1;
                    lc1 = c.y;
                    dL_dalpha += (c.y - acc1) * dLp1;
                    acc2 = last_alpha * lc2 + (1.f - last_alpha) * acc2;
                    lc2 = c.z;
                    dL_dalpha += (c.z - acc2) * dLp2;
                    v6 = dchannel_dcolor * dLp0;
                    v7 = dchannel_dcolor * dLp1;
                    v8 = dchannel_dcolor * dLp2;
                    if (EXTRA) {
                        const float e0c = s_ex[cur][3 * j + 0], e1c = s_ex[cur][3 * j + 1], e2c = s_ex[cur][3 * j + 2];
                        acc3 = last_alpha * lc3 + (1.f - last_alpha) * acc3;
                        lc3 = e0c;
                        dL_dalpha += (e0c - acc3) * dLe0;
                        acc4 = last_alpha * lc4 + (1.f - last_alpha) * acc4;
                        lc4 = e1c;
                        dL_dalpha += (e1c - acc4) * dLe1;
                        acc5 = last_alpha * lc5 + (1.f - last_alpha) * acc5;
                        lc5 = e2c;
                        dL_dalpha += (e2c - acc5) * dLe2;
                        v9 = dchannel_dcolor * dLe0;
                        v10 = dchannel_dcolor * dLe1;
                        v11 = dchannel_dcolor * dLe2;
                    }
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    dL_dalpha += (-T_final * inv_1ma) * bg_dot_dpixel;
                    const float dL_dG = co.w * dL_dalpha;
                    const float gdx = G * dx, gdy = G * dy;
                    const float dG_ddelx = -gdx * co.x - gdy * co.y;
                    const float dG_ddely = -gdy * co.z - gdx * co.y;
                    v0 = dL_dG * dG_ddelx * ddelx_dx;
                    v1 = dL_dG * dG_ddely * ddely_dy;
                    v2 = -0.5f * gdx * dx * dL_dG;
                    v3 = -0.5f * gdx * dy * dL_dG;
                    v4 = -0.5f * gdy * dy * dL_dG;
                    v5 = G * dL_dalpha;
                }
                {   // every visited (warp, entry) pair has a live lane (forward's contribution mask)
                    // Transposing butterfly: 8 terms reduced over 32 lanes with 4+2+1+1+1 shuffles
                    // (instead of 8x5); lane 4*k ends up holding the warp total of term k.
                    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
                    float a0 = (b4 ? v4 : v0) + __shfl_xor_sync(kFull, b4 ? v0 : v4, 16);
                    float a1 = (b4 ? v5 : v1) + __shfl_xor_sync(kFull, b4 ? v1 : v5, 16);
                    float a2 = (b4 ? v6 : v2) + __shfl_xor_sync(kFull, b4 ? v2 : v6, 16);
                    float a3 = (b4 ? v7 : v3) + __shfl_xor_sync(kFull, b4 ? v3 : v7, 16);
                    float c0 = (b3 ? a2 : a0) + __shfl_xor_sync(kFull, b3 ? a0 : a2, 8);
                    float c1 = (b3 ? a3 : a1) + __shfl_xor_sync(kFull, b3 ? a1 : a3, 8);
                    float e0 = (b2 ? c1 : c0) + __shfl_xor_sync(kFull, b2 ? c0 : c1, 4);
                    e0 += __shfl_xor_sync(kFull, e0, 2);
                    e0 += __shfl_xor_sync(kFull, e0, 1);
                    // term index held by this lane group: 4*b4 + 2*b3 + b2
                    // one fire-and-forget RED.ADD.F32 per (warp, splat, term), issued by the lanes that hold a total
                    if (EXTRA) {
                        float f0 = (b4 ? v10 : v8) + __shfl_xor_sync(kFull, b4 ? v8 : v10, 16);
                        float f1 = (b4 ? v11 : v9) + __shfl_xor_sync(kFull, b4 ? v9 : v11, 16);
                        float h0 = (b3 ? f1 : f0) + __shfl_xor_sync(kFull, b3 ? f0 : f1, 8);
                        h0 += __shfl_xor_sync(kFull, h0, 4);
                        h0 += __shfl_xor_sync(kFull, h0, 2);
                        h0 += __shfl_xor_sync(kFull, h0, 1);   // lanes with (b4, b3): term 8 + 2*b4 + b3
                        if (red_base != nullptr && lane != 1) atomicAdd(red_base + (size_t)s_id[cur][j] * red_stride, e0);
                        if (red2_base != nullptr) atomicAdd(red2_base + (size_t)s_id[cur][j] * 3, h0);
                    } else {
                        v8 = warp_sum(v8);
                        if (red_base != nullptr) atomicAdd(red_base + (size_t)s_id[cur][j] * red_stride, lane == 1 ? v8 : e0);
                    }
                }
            }
        }
    }
}

void launch_render_bwd(int W, int H, int gx, int gy, const uint2* ranges, const uint32_t* point_list,
                       const float4* rec, const float* bg, const float* final_T, const uint32_t* n_contrib,
                       const uint8_t* contrib, const float* dL_dpix, float* dL_dmean2D, float* dL_dconic,
                       float* dL_dopacity, float* dL_dcolors, const float* extra, const float* dL_dpix_extra,
                       float* dL_dextra, cudaStream_t s)
{
    if (extra != nullptr)
        k_render_bwd<true><<<gx * gy, 256, 0, s>>>(W, H, gx, ranges, point_list, rec, bg, final_T, n_contrib,
                                                   contrib, dL_dpix, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolors, extra,
                                                   dL_dpix_extra, dL_dextra);
    else
        k_render_bwd<false><<<gx * gy, 256, 0, s>>>(W, H, gx, ranges, point_list, rec, bg, final_T, n_contrib,
                                                    contrib, dL_dpix, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolors, nullptr,
                                                    nullptr, nullptr);
}

// ----------------------------------------------------------------------------------------------
// Diagnostic: work counts of the blend stage of a finished forward (gsr_blend_stats; bench / tests only).
// A plain per-pixel replay of the tile's list — no culling, no staging — so that the counts do not depend on
// any of the shortcuts the product kernels take.
//   out[0] += 256 * list length                     (pixel, splat) pairs before early termination
//   out[1] += per pixel, entries up to its last contributor   (what a per-pixel walk has to evaluate)
//   out[2] += per pixel, entries actually blended              (contributing pairs)
//   out[3] += 32 * (warp, entry) pairs recorded by k_render_fwd (pairs the backward evaluates)
__global__ void __launch_bounds__(256) k_blend_stats(int W, int H, int gx, const uint2* __restrict__ ranges,
                                                     const uint32_t* __restrict__ point_list,
                                                     const float4* __restrict__ rec, const uint32_t* __restrict__ n_contrib,
                                                     const uint8_t* __restrict__ contrib, unsigned long long* out)
{
    const int tile = (int)blockIdx.x;
    const TileGeom g = tile_geom(tile, gx, W, H);
    const uint2 range = ranges[tile];
    const uint32_t n = range.y - range.x;
    unsigned long long walked = 0, blended = 0, visits = 0;
    if (g.inside) {
        const uint32_t last = n_contrib[g.py * W + g.px];
        walked = last;
        const float pixx = (float)g.px, pixy = (float)g.py;
        for (uint32_t i = 0; i < last; i++) {
            const float4* r = rec + (size_t)point_list[range.x + i] * 3;
            const float4 a = r[0], co = r[1];
            const float dx = a.x - pixx, dy = a.y - pixy;
            const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
            const float alpha = fminf(0.99f, co.w * expf(power));
            if (!(power > 0.0f) && !(alpha < 1.0f / 255.0f)) blended++;
        }
    }
    for (uint32_t i = threadIdx.x; i < n; i += 256) visits += 32ull * (unsigned)__popc((unsigned)contrib[range.x + i]);
    __shared__ unsigned long long s_acc[3];
    if (threadIdx.x < 3) s_acc[threadIdx.x] = 0ull;
    __syncthreads();
    atomicAdd(&s_acc[0], walked);
    atomicAdd(&s_acc[1], blended);
    atomicAdd(&s_acc[2], visits);
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd(out + 0, 256ull * n);
        atomicAdd(out + 1, s_acc[0]);
        atomicAdd(out + 2, s_acc[1]);
        atomicAdd(out + 3, s_acc[2]);
    }
}

void launch_blend_stats(int W, int H, int gx, int gy, const uint2* ranges, const uint32_t* point_list, const float4* rec,
                        const uint32_t* n_contrib, const uint8_t* contrib, unsigned long long* out, cudaStream_t s)
{
    k_blend_stats<<<gx * gy, 256, 0, s>>>(W, H, gx, ranges, point_list, rec, n_contrib, contrib, out);
}

}  // namespace gsr
