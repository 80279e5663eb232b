// binning.cu — tile-instance generation, (tile|depth) ordering and per-tile ranges.
//
// Replaces cub::DeviceScan::InclusiveSum (CR/rasterizer_impl.cu:277), duplicateWithKeys
// (:70-111), cub::DeviceRadixSort::SortPairs (:303-308) and identifyTileRanges (:116-138).
// The produced (key, value) list is bit-identical to the reference's: key = tile << 32 | depth
// bits, ascending, ties in ascending Gaussian index (stable LSD order).
#include "gsr_internal.cuh"
#include "gsr_math.cuh"
#include <cub/cub.cuh>

namespace gsr {

size_t scan_temp_bytes(int P)
{
    size_t n = 0;
    cub::DeviceScan::InclusiveSum(nullptr, n, (uint32_t*)nullptr, (uint32_t*)nullptr, P > 0 ? P : 1);
    return n;
}
size_t sort_temp_bytes(int64_t R)
{
    size_t n = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, n, (uint64_t*)nullptr, (uint64_t*)nullptr, (uint32_t*)nullptr,
                                    (uint32_t*)nullptr, R > 0 ? (int)R : 1);
    return n;
}

void launch_scan(const uint32_t* in, uint32_t* out, int P, char* temp, size_t temp_bytes, cudaStream_t s)
{
    if (P <= 0) return;
    cub::DeviceScan::InclusiveSum(temp, temp_bytes, in, out, P, s);
}

// One warp per 32 Gaussians; each lane owns one Gaussian and walks its tile rectangle.
__global__ void __launch_bounds__(256) k_duplicate(int P, const float4* __restrict__ rec, const int* __restrict__ radii,
                                                   const uint32_t* __restrict__ offsets, uint64_t* __restrict__ keys,
                                                   uint32_t* __restrict__ vals, int gx, int gy)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const int r = radii[idx];
    if (r <= 0) return;
    const float4 r0 = rec[(size_t)idx * 3];
    uint32_t off = (idx == 0) ? 0u : offsets[idx - 1];
    int x0, y0, x1, y1;
    tile_rect(r0.x, r0.y, r, gx, gy, x0, y0, x1, y1);
    const uint64_t depth_bits = (uint64_t)__float_as_uint(r0.z);
    for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++) {
            uint64_t key = (uint64_t)(uint32_t)(y * gx + x);
            key = (key << 32) | depth_bits;
            keys[off] = key;
            vals[off] = (uint32_t)idx;
            off++;
        }
}

void launch_duplicate(int P, const float4* rec, const int* radii, const uint32_t* offsets, uint64_t* keys,
                      uint32_t* vals, int gx, int gy, cudaStream_t s)
{
    if (P <= 0) return;
    k_duplicate<<<(P + 255) / 256, 256, 0, s>>>(P, rec, radii, offsets, keys, vals, gx, gy);
}

void launch_sort(BinWS& b, int64_t R, int end_bit, cudaStream_t s)
{
    if (R <= 0) return;
    cub::DeviceRadixSort::SortPairs(b.sort_temp, b.sort_temp_bytes, b.keys_unsorted, b.keys, b.point_list_unsorted,
                                    b.point_list, (int)R, 0, end_bit, s);
}

__global__ void __launch_bounds__(256) k_tile_ranges(int64_t R, const uint64_t* __restrict__ keys, uint2* ranges)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    const uint32_t cur = (uint32_t)(keys[i] >> 32);
    if (i == 0)
        ranges[cur].x = 0;
    else {
        const uint32_t prev = (uint32_t)(keys[i - 1] >> 32);
        if (cur != prev) {
            ranges[prev].y = (uint32_t)i;
            ranges[cur].x = (uint32_t)i;
        }
    }
    if (i == R - 1) ranges[cur].y = (uint32_t)R;
}

void launch_tile_ranges(int64_t R, const uint64_t* keys, uint2* ranges, int tiles, cudaStream_t s)
{
    cudaMemsetAsync(ranges, 0, (size_t)tiles * sizeof(uint2), s);
    if (R <= 0) return;
    k_tile_ranges<<<(unsigned)((R + 255) / 256), 256, 0, s>>>(R, keys, ranges);
}

}  // namespace gsr
