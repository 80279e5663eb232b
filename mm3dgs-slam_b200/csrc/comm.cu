// comm.cu — SUM all-reduce of the gradient bucket over peer memory (include/gscomm_b200.h): one kernel, two shots
// (owner pulls + reduces its slice, then broadcasts it), flag handshakes instead of host synchronisation.
#include "gsr_internal.cuh"
#include "../../include/gscomm_b200.h"
#include <stdlib.h>

namespace gsr {

constexpr int kArThreads = 512;
constexpr int kArMaxBlocks = 148;           // upper bound on the CTAs of the exchange kernel (sizes the flag arrays)
static int ar_blocks()                      // CTAs actually launched: GSR_AR_BLOCKS overrides the default (tuning)
{
    const char* e = getenv("GSR_AR_BLOCKS");
    int v = e ? atoi(e) : 48;   // measured at 8 ranks / 56 MB: 32..96 CTAs all give 146-150 us (switch-bound)
    if (v < 1) v = 1;
    if (v > kArMaxBlocks) v = kArMaxBlocks;
    return v;
}
constexpr int kArFlagWords = kArMaxBlocks * 2 * GSR_COMM_MAX_RANKS;

struct ArArgs {
    float* bucket[GSR_COMM_MAX_RANKS];
    uint32_t* flags[GSR_COMM_MAX_RANKS];
    float* multicast;
    long long n;
    int nranks, rank;
    uint32_t epoch;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Every CTA shakes hands with the same CTA of every other rank: thread t < nranks tells rank t "I am here" (a store
// into ITS flag array, slot [phase][my rank]) and waits until rank t has said the same to us.  Epochs only grow, so
// the flags are never reset.
__device__ __forceinline__ void handshake(const ArArgs& a, int phase)
{
    __syncthreads();
    if ((int)threadIdx.x < a.nranks) {
        const int peer = threadIdx.x;
        const size_t slot = ((size_t)blockIdx.x * 2 + phase) * GSR_COMM_MAX_RANKS;
        __threadfence_system();
        st_release_sys(a.flags[peer] + slot + a.rank, a.epoch);
        const uint32_t* mine = a.flags[a.rank] + slot + peer;
        while ((int32_t)(ld_acquire_sys(mine) - a.epoch) < 0) __nanosleep(64);
    }
    __syncthreads();
}

__device__ __forceinline__ float4 ld_peer4(const float4* p)
{
    float4 r;
    asm volatile("ld.global.relaxed.sys.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_peer4(float4* p, const float4& v)
{
    asm volatile("st.global.relaxed.sys.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 multimem_ld_reduce4(const float4* p)
{
    float4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void multimem_st4(float4* p, const float4& v)
{
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

template <int NR, bool NVLS>
__global__ void __launch_bounds__(kArThreads) k_allreduce(ArArgs a)
{
    handshake(a, 0);                                  // every rank's bucket is complete and visible
    const long long n4 = (a.n + 3) >> 2;              // the allocation is padded to 16 bytes
    const long long per = (n4 + NR - 1) / NR;
    const long long lo = per * a.rank, hi = min(lo + per, n4);
    const long long stride = (long long)gridDim.x * kArThreads;
    for (long long i = lo + (long long)blockIdx.x * kArThreads + threadIdx.x; i < hi; i += 2 * stride) {
        const long long j = i + stride;
        const bool two = j < hi;
        if (NVLS) {
            float4 s0 = multimem_ld_reduce4(reinterpret_cast<const float4*>(a.multicast) + i);
            float4 s1 = s0;
            if (two) s1 = multimem_ld_reduce4(reinterpret_cast<const float4*>(a.multicast) + j);
            multimem_st4(reinterpret_cast<float4*>(a.multicast) + i, s0);
            if (two) multimem_st4(reinterpret_cast<float4*>(a.multicast) + j, s1);
        } else {
            float4 v0[NR], v1[NR];
#pragma unroll
            for (int r = 0; r < NR; r++) {
                v0[r] = ld_peer4(reinterpret_cast<const float4*>(a.bucket[r]) + i);
                if (two) v1[r] = ld_peer4(reinterpret_cast<const float4*>(a.bucket[r]) + j);
            }
            float4 s0 = v0[0], s1 = v1[0];
#pragma unroll
            for (int r = 1; r < NR; r++) {            // fixed rank order: the same bits whatever rank owns the slice
                s0.x += v0[r].x; s0.y += v0[r].y; s0.z += v0[r].z; s0.w += v0[r].w;
                if (two) { s1.x += v1[r].x; s1.y += v1[r].y; s1.z += v1[r].z; s1.w += v1[r].w; }
            }
#pragma unroll
            for (int r = 0; r < NR; r++) {
                st_peer4(reinterpret_cast<float4*>(a.bucket[r]) + i, s0);
                if (two) st_peer4(reinterpret_cast<float4*>(a.bucket[r]) + j, s1);
            }
        }
    }
    handshake(a, 1);                                  // every rank's stores have landed everywhere
}

// All-gather: rank r's slice of the buffer (same slicing as the all-reduce) is copied into every other rank's buffer.
template <int NR, bool NVLS>
__global__ void __launch_bounds__(kArThreads) k_allgather(ArArgs a)
{
    handshake(a, 0);                                  // every rank is done reading the buffer's previous contents
    const long long n4 = (a.n + 3) >> 2;
    const long long per = (n4 + NR - 1) / NR;
    const long long lo = per * a.rank, hi = min(lo + per, n4);
    const long long stride = (long long)gridDim.x * kArThreads;
    const float4* mine = reinterpret_cast<const float4*>(a.bucket[a.rank]);
    for (long long i = lo + (long long)blockIdx.x * kArThreads + threadIdx.x; i < hi; i += stride) {
        const float4 v = mine[i];
        if (NVLS) {
            multimem_st4(reinterpret_cast<float4*>(a.multicast) + i, v);
        } else {
#pragma unroll
            for (int r = 0; r < NR; r++)
                if (r != a.rank) st_peer4(reinterpret_cast<float4*>(a.bucket[r]) + i, v);
        }
    }
    handshake(a, 1);                                  // every rank's slice has landed everywhere
}

template <int NR>
static void launch_ag(const ArArgs& a, int blocks, cudaStream_t s)
{
    if (a.multicast) k_allgather<NR, true><<<blocks, kArThreads, 0, s>>>(a);
    else k_allgather<NR, false><<<blocks, kArThreads, 0, s>>>(a);
}

template <int NR>
static void launch_ar(const ArArgs& a, int blocks, cudaStream_t s)
{
    if (a.multicast) k_allreduce<NR, true><<<blocks, kArThreads, 0, s>>>(a);
    else k_allreduce<NR, false><<<blocks, kArThreads, 0, s>>>(a);
}

}  // namespace gsr

using namespace gsr;

extern "C" {

size_t gsr_allreduce_flag_words(void) { return (size_t)kArFlagWords; }

static int exchange(bool gather, gsr_stream_t stream_, int32_t nranks, int32_t rank, float* const* bucket, float* multicast,
                    uint32_t* const* flags, int64_t n, uint32_t epoch)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (nranks < 1 || nranks > GSR_COMM_MAX_RANKS || rank < 0 || rank >= nranks) return api_fail(GSR_ERR_INVALID, "bad rank / nranks");
    if (nranks != 1 && nranks != 2 && nranks != 4 && nranks != 8) return api_fail(GSR_ERR_INVALID, "nranks must be 1, 2, 4 or 8");
    if (n < 0 || !bucket || !flags || epoch == 0) return api_fail(GSR_ERR_INVALID, "bad argument");
    if (n == 0 || nranks == 1) return GSR_OK;
    ArArgs a;
    for (int r = 0; r < GSR_COMM_MAX_RANKS; r++) {
        a.bucket[r] = r < nranks ? bucket[r] : nullptr;
        a.flags[r] = r < nranks ? flags[r] : nullptr;
        if (r < nranks && (!bucket[r] || !flags[r] || ((uintptr_t)bucket[r] & 15))) return api_fail(GSR_ERR_INVALID, "null / misaligned peer pointer");
    }
    a.multicast = multicast; a.n = n; a.nranks = nranks; a.rank = rank; a.epoch = epoch;
    const long long n4 = (n + 3) >> 2, per = (n4 + nranks - 1) / nranks;
    long long blocks = (per + 2 * kArThreads - 1) / (2 * kArThreads);
    if (blocks > ar_blocks()) blocks = ar_blocks();
    if (blocks < 1) blocks = 1;
    if (gather) {
        if (blocks > 32) blocks = 32;       // a copy: few CTAs saturate the links and the rest of the GPU keeps computing
        switch (nranks) {
            case 2: launch_ag<2>(a, (int)blocks, stream); break;
            case 4: launch_ag<4>(a, (int)blocks, stream); break;
            default: launch_ag<8>(a, (int)blocks, stream); break;
        }
    } else {
        switch (nranks) {
            case 2: launch_ar<2>(a, (int)blocks, stream); break;
            case 4: launch_ar<4>(a, (int)blocks, stream); break;
            default: launch_ar<8>(a, (int)blocks, stream); break;
        }
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return api_fail(GSR_ERR_CUDA, "exchange launch", e);
    api_count_launches(1);
    return GSR_OK;
}

int gsr_allreduce_sum_f32(gsr_stream_t stream, int32_t nranks, int32_t rank, float* const* bucket, float* multicast,
                          uint32_t* const* flags, int64_t n, uint32_t epoch)
{
    return exchange(false, stream, nranks, rank, bucket, multicast, flags, n, epoch);
}

int gsr_allgather_f32(gsr_stream_t stream, int32_t nranks, int32_t rank, float* const* buffer, float* multicast,
                      uint32_t* const* flags, int64_t n, uint32_t epoch)
{
    return exchange(true, stream, nranks, rank, buffer, multicast, flags, n, epoch);
}

}  // extern "C"
