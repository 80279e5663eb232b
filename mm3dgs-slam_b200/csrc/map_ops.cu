// map_ops.cu — map surgery on the flat parameter / optimizer buffers (second half).
//
// Replaces the reference's _prune_optimizer / cat_tensors_to_optimizer (R/slam/gaussian_model.py:380-399, :418-451):
// boolean-mask indexing of every parameter group and both Adam moments (7 groups x 3 tensors = 21 index kernels, each
// with its own mask scan and allocation) and torch.cat of every group.  Here the keep mask is scanned ONCE and a
// single pass moves every kept row of every group of every buffer to its place in the re-laid-out flat buffers
// (the groups are [P, w_g] slabs one after the other, so a change of P moves every group).
//
//   gsr_compact_scan   : new_index[i] = number of kept rows before row i; count = number of kept rows
//   gsr_compact_gather : dst_b[group g][new_index[i]][:] = src_b[group g][i][:] for every kept row, all b, all g
//
// Pure HBM streaming: 4 B/row of mask + index traffic, then every source byte read once and every kept byte written
// once with coalesced accesses.
#include "gsr_internal.cuh"
#include "../../include/gsloss_b200.h"

namespace gsr {

constexpr int kCT = 1024;            // threads per CTA of the flag scan
constexpr int kCRows = 4;            // rows per thread
constexpr int kCChunk = kCT * kCRows;

__device__ __forceinline__ uint32_t block_sum_u32(uint32_t v, uint32_t* s_warp)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = __reduce_add_sync(0xffffffffu, v);
    if (lane == 0) s_warp[warp] = v;
    __syncthreads();
    uint32_t t = (threadIdx.x < (blockDim.x >> 5)) ? s_warp[threadIdx.x] : 0u;
    if (warp == 0) {
        t = __reduce_add_sync(0xffffffffu, t);
        if (lane == 0) s_warp[0] = t;
    }
    __syncthreads();
    const uint32_t r = s_warp[0];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kCT) k_flag_count(long long P, const uint8_t* __restrict__ keep, uint32_t* __restrict__ partial)
{
    __shared__ uint32_t s_warp[32];
    const long long base = (long long)blockIdx.x * kCChunk + (long long)threadIdx.x * kCRows;
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < kCRows; j++)
        if (base + j < P) c += (keep == nullptr || keep[base + j]) ? 1u : 0u;
    const uint32_t tot = block_sum_u32(c, s_warp);
    if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kCT) k_flag_scan(long long P, const uint8_t* __restrict__ keep,
                                                   const uint32_t* __restrict__ partial, uint32_t* __restrict__ new_index,
                                                   long long* __restrict__ count)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_wsum[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t before = 0;                         // kept rows in the chunks before this one
    for (int c = tid; c < (int)blockIdx.x; c += kCT) before += partial[c];
    before = block_sum_u32(before, s_warp);
    const long long base = (long long)blockIdx.x * kCChunk + (long long)tid * kCRows;
    uint32_t f[kCRows], c = 0;
#pragma unroll
    for (int j = 0; j < kCRows; j++) {
        f[j] = (base + j < P && (keep == nullptr || keep[base + j])) ? 1u : 0u;
        c += f[j];
    }
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    uint32_t run = before + incl - c;
    for (int w = 0; w < warp; w++) run += s_wsum[w];
#pragma unroll
    for (int j = 0; j < kCRows; j++) {
        if (base + j < P) new_index[base + j] = run;
        run += f[j];
    }
    if (blockIdx.x == gridDim.x - 1 && tid == kCT - 1) *count = (long long)run;
}

struct CompactArgs {
    long long P, rows_out;
    int ngroups, nbuf;
    int width[GSR_COMPACT_MAX_GROUPS];
    long long src_off[GSR_COMPACT_MAX_GROUPS], dst_off[GSR_COMPACT_MAX_GROUPS];   // in floats
    const float* src[GSR_COMPACT_MAX_BUFFERS];
    float* dst[GSR_COMPACT_MAX_BUFFERS];
};

// grid.y = group; a CTA grid-strides over the P * w elements of its group: consecutive threads read consecutive
// floats, and (kept rows keep their relative order) write nearly consecutive floats.
__global__ void __launch_bounds__(256) k_compact_gather(CompactArgs a, const uint8_t* __restrict__ keep,
                                                        const uint32_t* __restrict__ new_index)
{
    const int g = blockIdx.y;
    const uint32_t w = (uint32_t)a.width[g];
    const long long n = a.P * w;
    const long long so = a.src_off[g], doff = a.dst_off[g];
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const long long i = e / w;
        const uint32_t c = (uint32_t)(e - i * w);
        if (keep != nullptr && !keep[i]) continue;
        const long long o = doff + (long long)(new_index ? new_index[i] : (uint32_t)i) * w + c;
#pragma unroll
        for (int b = 0; b < GSR_COMPACT_MAX_BUFFERS; b++)
            if (b < a.nbuf) a.dst[b][o] = __ldcs(a.src[b] + so + e);
    }
}

}  // namespace gsr

using namespace gsr;

extern "C" {

size_t gsr_compact_ws_bytes(int64_t P)
{
    const int64_t chunks = P > 0 ? (P + kCChunk - 1) / kCChunk : 1;
    return align_up((size_t)chunks * sizeof(uint32_t));
}

int gsr_compact_scan(gsr_stream_t stream_, int64_t P, const uint8_t* keep, uint32_t* new_index, int64_t* count, void* ws,
                     size_t ws_bytes)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (P < 0 || P > 0xffffffffLL) return api_fail(GSR_ERR_INVALID, "P out of range");
    if (!count) return api_fail(GSR_ERR_INVALID, "count required");
    if (P == 0) {
        if (cudaMemsetAsync(count, 0, sizeof(int64_t), stream) != cudaSuccess) return api_fail(GSR_ERR_CUDA, "memset");
        return GSR_OK;
    }
    if (!new_index || !ws) return api_fail(GSR_ERR_INVALID, "new_index / workspace required");
    if (ws_bytes < gsr_compact_ws_bytes(P)) return api_fail(GSR_ERR_WORKSPACE, "compaction workspace too small");
    const int chunks = (int)((P + kCChunk - 1) / kCChunk);
    k_flag_count<<<chunks, kCT, 0, stream>>>((long long)P, keep, (uint32_t*)ws);
    k_flag_scan<<<chunks, kCT, 0, stream>>>((long long)P, keep, (const uint32_t*)ws, new_index, (long long*)count);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return api_fail(GSR_ERR_CUDA, "compact scan launch", e);
    api_count_launches(2);
    return GSR_OK;
}

int gsr_compact_gather(gsr_stream_t stream_, int64_t P, int64_t rows_out, int32_t ngroups, const int32_t* widths,
                       const uint8_t* keep, const uint32_t* new_index, int32_t nbuf, const float* const* src,
                       float* const* dst)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (P < 0 || rows_out < 0 || P > 0xffffffffLL) return api_fail(GSR_ERR_INVALID, "row counts out of range");
    if (ngroups < 1 || ngroups > GSR_COMPACT_MAX_GROUPS || !widths) return api_fail(GSR_ERR_INVALID, "1..GSR_COMPACT_MAX_GROUPS groups");
    if (nbuf < 1 || nbuf > GSR_COMPACT_MAX_BUFFERS || !src || !dst) return api_fail(GSR_ERR_INVALID, "1..GSR_COMPACT_MAX_BUFFERS buffers");
    if (keep != nullptr && new_index == nullptr) return api_fail(GSR_ERR_INVALID, "new_index required with a keep mask");
    CompactArgs a;
    a.P = P; a.rows_out = rows_out; a.ngroups = ngroups; a.nbuf = nbuf;
    long long so = 0, dof = 0;
    int wmax = 1;
    for (int g = 0; g < GSR_COMPACT_MAX_GROUPS; g++) {
        const int w = g < ngroups ? widths[g] : 1;
        if (w < 1) return api_fail(GSR_ERR_INVALID, "group width must be positive");
        a.width[g] = w; a.src_off[g] = so; a.dst_off[g] = dof;
        if (g < ngroups) { so += (long long)P * w; dof += (long long)rows_out * w; if (w > wmax) wmax = w; }
    }
    for (int b = 0; b < GSR_COMPACT_MAX_BUFFERS; b++) {
        a.src[b] = b < nbuf ? src[b] : nullptr;
        a.dst[b] = b < nbuf ? dst[b] : nullptr;
        if (b < nbuf && P > 0 && (!src[b] || (!dst[b] && rows_out > 0))) return api_fail(GSR_ERR_INVALID, "null buffer");
    }
    if (P == 0 || rows_out == 0) return GSR_OK;
    long long bx = ((long long)P * wmax + 255) / 256;
    const long long cap = (long long)device_sm_count() * 8;
    if (bx > cap) bx = cap;
    k_compact_gather<<<dim3((unsigned)bx, (unsigned)ngroups), 256, 0, stream>>>(a, keep, new_index);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return api_fail(GSR_ERR_CUDA, "compact gather launch", e);
    api_count_launches(1);
    return GSR_OK;
}

}  // extern "C"
