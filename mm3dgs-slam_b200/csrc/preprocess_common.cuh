// preprocess_common.cuh — block-wide staging helpers shared by the per-Gaussian kernels.
#pragma once
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
// Copy n contiguous floats shared -> global (ACC: add to what is there — with float atomics, like the TMA reduce-adds
// of the full chunks, so that frames back-propagated on different streams may share one accumulator).  Only the
// ragged / unaligned chunks take this path; full aligned chunks leave through the TMA unit (async_copy.cuh).
template <bool ACC>
__device__ __forceinline__ void stage_out(float* __restrict__ dst, const float* src, int n, int tid)
{
    if (!ACC && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        const int n4 = n >> 2;
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int i = tid; i < n4; i += kPB) d4[i] = s4[i];
        for (int i = (n4 << 2) + tid; i < n; i += kPB) dst[i] = src[i];
    } else {
        for (int i = tid; i < n; i += kPB) {
            if (ACC) atomicAdd(dst + i, src[i]);
            else dst[i] = src[i];
        }
    }
}

// Raw optimizer parameters -> the values the rasterizer works with (gsr_gaussians.raw_params): the reference applies
// torch.exp / torch.sigmoid / F.normalize(eps=1e-12) to its parameters on every render (R/slam/gaussian_model.py:108-132).
__device__ __forceinline__ float act_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ V3 act_exp3(const V3& s) { return V3{expf(s.x), expf(s.y), expf(s.z)}; }
__device__ __forceinline__ V4 act_normalize4(const V4& q, float& inv_norm)
{
    const float n = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    inv_norm = 1.0f / fmaxf(n, 1e-12f);
    return V4{q.x * inv_norm, q.y * inv_norm, q.z * inv_norm, q.w * inv_norm};
}

}  // namespace gsr
