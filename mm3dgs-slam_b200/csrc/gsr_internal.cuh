// gsr_internal.cuh — workspace layouts and launcher prototypes shared by the translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/gsrast_b200.h"

namespace gsr {

constexpr int kTile = 16;
constexpr size_t kAlign = 256;

inline size_t align_up(size_t x, size_t a = kAlign) { return (x + a - 1) / a * a; }

// ---- per-Gaussian state ---------------------------------------------------------------------
// One 48-byte record per Gaussian so that the blend kernels gather a single contiguous chunk:
//   rec[3i+0] = (px, py, depth, cull_r2)        pixel mean, view depth, conservative cull radius^2
//   rec[3i+1] = (conic.x, conic.y, conic.z, opacity)
//   rec[3i+2] = (r, g, b, bits)                 bits: SH clamp flags (bit ch set => channel clamped)
// The reference keeps these in separate arrays (GeometryState, CR/rasterizer_impl.h:29-44) and
// additionally stores cov3D (24 B) which the backward here recomputes from scale/rotation.
// counters[] slots (device side; zeroed at the start of every forward together with ghist / status)
constexpr int kCntR = 0;         // number of tile instances (num_rendered)
constexpr int kCntClaim = 1;     // tile partition: next free slot of the instance stream
constexpr int kCntVisible = 2;   // number of visible Gaussians (= length of the depth-sorted list)
constexpr int kCntTicket = 3;    // [4] block tickets of the four depth-sort passes (slots 3..6)
constexpr int kCntChunkFwd = 8;  // chunk ticket of k_preprocess_fwd
constexpr int kCntChunkBwd = 9;  // chunk ticket of k_preprocess_bwd (wraps to zero by itself)
constexpr int kCntTotalsDone = 10; // blocks of depth-sort pass 0 that have finished (the last one builds the tile ranges)
constexpr int kCntPartTicket = 11; // block ticket of the tile partition
// Depth sort: the keys are the float bits of view depths > 0.1, so `bits - kSortKeyBase` is a small non-negative number
// (< 2^27 for every depth below 6553) whose order — ties included — is the order of the keys: three 9-bit digits sort
// it, and a fourth pass over the top 5 bits runs only when a deeper Gaussian exists.
constexpr int kSortDigits = 4;
constexpr int kSortBits = 9;
constexpr int kSortBins = 1 << kSortBits;
constexpr uint32_t kSortKeyBase = 0x3DCCCCCDu;   // float bits of 0.1f: every visible depth is larger (CR/auxiliary.h:154)
constexpr int kTileDiffReplicas = 8;   // copies of the tile-count difference array (spreads the flush REDs)

struct SortWS {
    uint2 *pairs_a, *pairs_b;                      // depth-sort ping-pong of {depth key, Gaussian id} (P each); result in pairs_a
    uint32_t* ghist;                               // [4][512] global digit histograms of the visible depth keys
    uint32_t* status;                              // [sort_chunks(P)][256] look-back state of the sort passes
    size_t status_words;
    int* tile_diff;                                // [kTileDiffReplicas][(gy+1)*(gx+1)] difference arrays of the per-tile instance counts
    uint32_t* tile_status;                         // [partition CTAs][T] look-back state of the tile partition
    uint32_t* tile_starts;                         // [T] first list position of every tile
};
constexpr int kCamPartialRows = 1024;   // >= the grid of k_preprocess_bwd (2 CTAs per SM)
struct GeomWS {
    double* cam_partials;  // [kCamPartialRows][35] scratch of the camera-gradient reduction
    float4* rec;
    ushort4* rects;        // tile rectangle {x0, y0, x1, y1}; empty for culled Gaussians
    uint32_t* depth_keys;  // float bits of the view depth; 0xFFFFFFFF for culled Gaussians
    uint32_t* counters;    // kCnt* slots
    size_t zero_bytes;     // bytes from `counters` that must be zero when a forward starts (counters, ghist, status)
    float* extra_gen;      // [P,3] generated (z, 1, z^2) colours (extra_mode 1)
    SortWS sort;
    size_t total;
};
struct ImgWS {
    float* final_T;
    uint32_t* n_contrib;
    uint2* ranges;
    size_t total;
};
struct BinWS {
    uint32_t* point_list;   // [R] final per-tile, depth-ordered Gaussian ids
    uint8_t* contrib;       // [R] per list entry: which of the tile's 8 warps blended it in the forward
    size_t total;
};

GeomWS geom_ws_carve(char* base, int P, int W, int H);
ImgWS img_ws_carve(char* base, int W, int H);
BinWS bin_ws_carve(char* base, int64_t R);
int sort_chunks(int P);
void tile_partition_plan(int P, int T, int& ctas, int& per_cta, int& warps, size_t& smem);

// ---- launchers (each enqueues on `s`) ---------------------------------------------------------
struct PreArgs {
    int P, D, M, W, H, gx, gy, prefiltered;
    int raw;                  // opacities / scales / rotations are raw optimizer parameters (fused activations)
    const float *means, *scales, *rots, *opac, *shs, *colors, *cov3D_pre, *view, *proj, *campos;
    float scale_mod, tanfovx, tanfovy, focal_x, focal_y;
    int* radii;
    float4* rec;
    ushort4* rects;
    uint32_t* depth_keys;
    uint32_t* num_rendered;   // device counter, zeroed by the caller
    uint32_t* ghist;          // [4][512] digit histograms of the depth sort; this kernel fills digit 0, zeroed by the caller
    uint32_t* status;         // look-back state of the depth sort: zeroed by this kernel (status_words u32)
    size_t status_words;
    uint32_t* chunk_ticket;   // chunk scheduler of the persistent kernel, zeroed by the caller
    float* extra_gen;         // [P,3] generated depth/silhouette colours (z, 1, z^2), or nullptr
};
void launch_preprocess_fwd(const PreArgs& a, cudaStream_t s);
void launch_mark_visible(int P, const float* means, const float* view, const float* proj, uint8_t* present,
                         cudaStream_t s);

int launch_depth_sort(const uint32_t* depth_keys, const ushort4* rects, int P, int gx, int gy, SortWS& w, uint2* ranges,
                      uint32_t* counters, cudaStream_t s);
int launch_tile_partition(int P, const ushort4* rects, int gx, int gy, SortWS& w,
                          uint32_t* counters, uint32_t cap, uint32_t* point_list, cudaStream_t s);

void launch_render_fwd(int W, int H, int gx, int gy, const uint2* ranges, const uint32_t* point_list,
                       const float4* rec, const float* bg, float* final_T, uint32_t* n_contrib, float* out_color,
                       uint8_t* contrib, const float* extra, float* out_extra, const uint32_t* counters, uint32_t cap,
                       cudaStream_t s);
void launch_render_bwd(int W, int H, int gx, int gy, const uint2* ranges, const uint32_t* point_list,
                       const float4* rec, const float* bg, const float* final_T, const uint32_t* n_contrib,
                       const uint8_t* contrib, const float* dL_dpix, float* dL_dmean2D /*[P,3]*/, float* dL_dconic /*[P,4]*/,
                       float* dL_dopacity, float* dL_dcolors /*[P,3]*/, const float* extra, const float* dL_dpix_extra,
                       float* dL_dextra, cudaStream_t s);

void launch_blend_stats(int W, int H, int gx, int gy, const uint2* ranges, const uint32_t* point_list, const float4* rec,
                        const uint32_t* n_contrib, const uint8_t* contrib, unsigned long long* out, cudaStream_t s);

struct PreBwdArgs {
    int P, D, M, W, H;
    int raw;                      // see PreArgs::raw: gradients are chained back to the raw parameters
    const float* opac;            // [P] raw opacities (raw mode)
    const float* dL_dopacity;     // [P] dL/d(activated opacity) accumulated by the blend backward (raw mode)
    float* dL_dopacity_raw;       // [P] out: dL/d(raw opacity) (raw mode)
    const float *means, *scales, *rots, *shs, *cov3D_pre, *view, *proj, *campos;
    float scale_mod, tanfovx, tanfovy, focal_x, focal_y;
    const int* radii;
    const float4* rec;
    const float* dL_dmean2D;  // [P,3]
    const float* dL_dconic;   // [P,4]
    const float* dL_dcolors;  // [P,3]
    float *dL_dmeans3D, *dL_dcov3D, *dL_dsh, *dL_dscales, *dL_drots;
    float *dL_dview, *dL_dproj, *dL_dcampos;
    double* cam_partials;     // [kCamPartialRows][35] per-CTA fp64 partial sums of the camera gradients (geometry workspace)
    unsigned* cam_done;       // CTAs that have written their record (zero on entry, reset by the kernel)
    unsigned* chunk_ticket;   // chunk scheduler of the persistent kernel (zero on entry; wraps back to zero by itself)
    int accumulate;
    const float* dL_dextra_gen;   // [P,3] gradient of the generated (z, 1, z^2) colours, or nullptr
};
void launch_preprocess_bwd(const PreBwdArgs& a, cudaStream_t s);

int device_sm_count();   // SM count of the current device (cached per device)

// ---- shared by the translation units that define extern "C" entry points (c_api.cu owns the state) ----
int api_fail(int code, const char* what, cudaError_t e = cudaSuccess);   // records gsr_last_error(), returns code
void api_count_launches(int n);                                          // feeds gsr_launch_count()

}  // namespace gsr
