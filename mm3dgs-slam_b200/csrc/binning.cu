// binning.cu — builds the per-tile, depth-ordered splat lists.
//
// Replaces cub::DeviceScan::InclusiveSum (CR/rasterizer_impl.cu:277), duplicateWithKeys (:70-111),
// the 64-bit cub::DeviceRadixSort::SortPairs over all R tile instances (:303-308) and
// identifyTileRanges (:116-138).  The result is the same list the reference produces — instances
// ordered by (tile, depth bits, Gaussian index) — but it is built B200-first, without ever
// materialising or sorting 64-bit (tile|depth) keys:
//
//   1. depth sort of the P Gaussians (not of the R >> P instances): stable LSD radix sort of the
//      32-bit depth keys in three 11/11/10-bit passes (culled Gaussians carry key 0xFFFFFFFF and sink
//      to the end).  Ties keep ascending Gaussian index — the reference's tie order, because its
//      stable sort starts from instances emitted in index order.
//   2. ONE stable multi-bin partition of the instances by tile id, walking the Gaussians in depth
//      order: within a tile, instances therefore appear in (depth, index) order.
//
// Both steps are the same three-kernel pattern (count -> scan -> scatter) built around per-warp
// 16-bit histograms in shared memory — the wide digits / thousands of tile bins are only possible
// because a B200 CTA can hold 32-160 KB of counters next to its data:
//   count   : CTA-local histogram of a contiguous chunk (shared-memory atomics), one row of H[c][bin]
//   scan    : per bin, exclusive prefix over the chunks (+ per-bin totals)
//   scatter : exclusive scan over the bins (in shared memory), then each warp ranks its contiguous
//             sub-chunk, in order, against its private counters (same-bin lanes of a 32-element step are
//             matched with MATCH.ANY for digits / one ballot per bit for tile ids).
// The tile partition additionally materialises the instance stream once ({tile, id} records, written
// by warps that each own an equal share of their CTA's instances) so that the scatter pass is two
// coalesced sweeps.  HBM traffic: 3 x 16 B per Gaussian for the depth sort + ~20 B per instance,
// instead of the reference's ~150 B per instance (6 onesweep passes over 12-byte pairs).
#include "gsr_internal.cuh"
#include "gsr_math.cuh"
#include <algorithm>
#include <stdlib.h>

namespace gsr {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kSortThreads = 512;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortItems = 8;                // keys per thread (4 measured slower)
constexpr int kSortChunk = kSortThreads * kSortItems;   // 4096 keys per CTA

int sort_chunks(int P) { return (P + kSortChunk - 1) / kSortChunk; }
__device__ __forceinline__ int sort_chunks_dev(int n) { return (n + kSortChunk - 1) / kSortChunk; }

__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// Lanes holding the same `nbits`-bit value as this lane (among lanes with valid == true), built from
// one ballot per bit (VOTE is cheap; the hardware MATCH.ANY serialises over the distinct values,
// and the values here — 11-bit digits of random keys, tile ids of neighbouring splats — are
// almost all distinct within a warp).
__device__ __forceinline__ unsigned match_any_bits(uint32_t v, bool valid, int nbits)
{
    unsigned peers = __ballot_sync(kFullMask, valid);
    for (int bit = 0; bit < nbits; bit++) {
        const bool set = (v >> bit) & 1u;
        const unsigned b = __ballot_sync(kFullMask, set);
        peers &= set ? b : ~b;
    }
    return valid ? peers : 0u;
}

// Exclusive scan of `n` u32 values in shared memory (n a multiple of blockDim.x, in place), all threads call.
// Each thread owns n/blockDim consecutive values.  `s_warp` is scratch for blockDim/32 partials.
template <int PER_THREAD>
__device__ __forceinline__ void block_exclusive_scan(uint32_t* s_data, uint32_t* s_warp)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t v[PER_THREAD];
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < PER_THREAD; i++) {
        v[i] = s_data[tid * PER_THREAD + i];
        sum += v[i];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(kFullMask, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0;
    for (int w = 0; w < warp; w++) wbase += s_warp[w];
    uint32_t run = wbase + incl - sum;
#pragma unroll
    for (int i = 0; i < PER_THREAD; i++) {
        s_data[tid * PER_THREAD + i] = run;
        run += v[i];
    }
    __syncthreads();
}

// ================================================================================================
// 1. Depth sort (visible Gaussians only)
// ================================================================================================
// Four single-kernel LSD passes over 8-bit digits in the "onesweep" form.  The four global digit histograms are
// accumulated by k_preprocess_fwd while the key is still in a register, so a pass only has to
//   (1) rank its 4096-key block locally (MATCH.ANY + per-warp 16-bit counters, one dependent sweep),
//   (2) learn how many keys with the same digit precede the block: decoupled look-back over the blocks' published
//       per-digit counts, one digit per thread, four predecessors in flight per step.  Blocks take their index from
//       a ticket counter, so every block a look-back waits on is already running;
//   (3) reorder the block in shared memory by digit and write it out as runs (about 16 pairs = 128 B per digit and
//       block), so the scatter costs a few store wavefronts per warp instead of one per key.
// 8-bit digits rather than the 11/11/10 split of a count / scan / scatter sort: with 256 digits a 4096-key block
// holds runs, the per-warp counters are 8 KB instead of 64 KB, and a look-back step is one 1 KB row.
// Pass 0 drops the culled Gaussians (key 0xFFFFFFFF): later passes and the tile partition only see the visible ones.
// Keys travel with their Gaussian id as one 8-byte pair; pass 0 reads the bare depth keys (the id is the index).
//
// status[block][digit]: bits 31..28 tag, bits 27..0 count.  For pass p, tag 2p+1 = the block's own count, tag 2p+2 =
// inclusive prefix over blocks 0..block; anything else = not published yet for this pass (the array is zeroed once per
// frame and reused by the four passes).
constexpr uint32_t kStatMask = 0x0fffffffu;
constexpr int kRadixBits = kSortBits;
constexpr int kRBins = kSortBins;
static_assert(kRBins == kSortThreads, "one thread per digit in the prefix / look-back sections");
__device__ __forceinline__ uint32_t sort_digit(uint32_t key, int pass)
{
    return ((key - kSortKeyBase) >> (kRadixBits * pass)) & (uint32_t)(kRBins - 1);
}

__device__ __forceinline__ uint32_t ld_status(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
__device__ __forceinline__ void st_status(uint32_t* p, uint32_t v) { *reinterpret_cast<volatile uint32_t*>(p) = v; }
__device__ __forceinline__ uint4 ld_status4(const uint32_t* p)
{
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status4(uint32_t* p, const uint4& v)
{
    asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Pass 0 additionally derives the per-tile instance counts from the tile rectangles (TileCount below): a splat adds one
// instance to every tile of its rectangle [x0, x1) x [y0, y1), which as a 2-D difference array is four corner updates
// per SPLAT (+1, -1, -1, +1) instead of one count per instance.  Every block accumulates them in shared memory (the
// pass is a latency-bound walk whose load/store unit is otherwise idle), adds its non-zero cells to one of
// kTileDiffReplicas global copies, and the block that finishes last sums the copies, integrates along x and y
// (= instances per tile) and scans the tiles into ranges[t] = {start, end} and tile_starts[t] — ready long before the
// tile partition runs, which can therefore place every instance at its final position in one kernel.
struct TileCount {
    const ushort4* rects;     // [P] tile rectangles
    int gx, gy;
    int* tile_diff;           // [kTileDiffReplicas][(gy+1)*(gx+1)], zero on entry
    uint2* ranges;            // [T] out
    uint32_t* tile_starts;    // [T] out
};

__device__ __forceinline__ void tile_ranges_from_diff(const TileCount& tc, int* cell, uint32_t* s_wsum, uint32_t* s_carry)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x;
    const int pitch = tc.gx + 1, rows = tc.gy + 1, cells = pitch * rows, T = tc.gx * tc.gy;
    for (int c = tid; c < cells; c += nthreads) {
        int acc = 0;
#pragma unroll
        for (int k = 0; k < kTileDiffReplicas; k++) acc += __ldcg(tc.tile_diff + (size_t)k * cells + c);
        cell[c] = acc;
    }
    __syncthreads();
    for (int r = tid; r < rows; r += nthreads) {              // along x
        int run = 0;
        for (int x = 0; x < pitch; x++) { run += cell[r * pitch + x]; cell[r * pitch + x] = run; }
    }
    __syncthreads();
    for (int x = tid; x < pitch; x += nthreads) {             // along y
        int run = 0;
        for (int r = 0; r < rows; r++) { run += cell[r * pitch + x]; cell[r * pitch + x] = run; }
    }
    if (tid == 0) *s_carry = 0;
    __syncthreads();
    for (int t0 = 0; t0 < T; t0 += nthreads) {
        const int t = t0 + tid;
        const uint32_t v = (t < T) ? (uint32_t)cell[(t / tc.gx) * pitch + (t % tc.gx)] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t x = __shfl_up_sync(kFullMask, incl, o);
            if (lane >= o) incl += x;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        uint32_t wb = 0;
        for (int w = 0; w < warp; w++) wb += s_wsum[w];
        const uint32_t carry = *s_carry;
        const uint32_t start = carry + wb + incl - v;
        if (t < T) {
            tc.tile_starts[t] = start;
            tc.ranges[t] = v ? make_uint2(start, start + v) : make_uint2(0u, 0u);   // empty tiles: {0,0} like the reference's memset
        }
        __syncthreads();
        if (tid == nthreads - 1) *s_carry = carry + wb + incl;
        __syncthreads();
    }
}

template <int PASS>
__global__ void __launch_bounds__(kSortThreads, 2) k_sort_pass(const uint32_t* __restrict__ keys_in,
                                                               const uint2* __restrict__ pairs_in, int n_in,
                                                               const uint32_t* __restrict__ ghist, uint32_t* status,
                                                               uint32_t* counters, uint2* __restrict__ pairs_out,
                                                               TileCount tc)
{
    extern __shared__ __align__(16) unsigned char s_dynamic[];
    uint2* s_pairs = reinterpret_cast<uint2*>(s_dynamic);                      // the block, reordered by digit (32 KB)
    int* s_tdiff = reinterpret_cast<int*>(s_dynamic + sizeof(uint2) * kSortChunk);   // pass 0 only: tile-count difference array
    __shared__ uint16_t s_cnt[kSortWarps][kRBins];        // per-warp digit counters -> exclusive prefix over the warps
    __shared__ uint32_t s_lstart[kRBins];                 // first local slot of each digit
    __shared__ uint32_t s_gdst[kRBins];                   // global destination of local slot 0 of each digit's run
    __shared__ uint32_t s_warp[kRBins / 32];
    __shared__ uint32_t s_next[PASS < kSortDigits - 1 ? kRBins : 1];   // this block's histogram of the NEXT digit
    __shared__ uint32_t s_block, s_total;
    constexpr uint32_t tag_agg = (uint32_t)(2 * PASS + 1) << 28, tag_incl = (uint32_t)(2 * PASS + 2) << 28;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // the fourth pass has work only if some key does not fit 27 bits (a view depth beyond 6553): otherwise every key
    // carries digit 0 and the list pass 2 wrote is the result (k_tile_partition makes the same test to pick the buffer)
    if (PASS == kSortDigits - 1 && ghist[PASS * kRBins] == counters[kCntVisible]) return;
    if (tid == 0) s_block = atomicAdd(&counters[kCntTicket + PASS], 1u);
    if (PASS < kSortDigits - 1) s_next[tid] = 0u;
    __syncthreads();
    const int block = (int)s_block;
    const int n = PASS == 0 ? n_in : (int)counters[kCntVisible];
    if (block * kSortChunk >= n) return;
    const int tcells = PASS == 0 ? (tc.gx + 1) * (tc.gy + 1) : 0;
    if (PASS == 0)
        for (int i = tid; i < tcells; i += kSortThreads) s_tdiff[i] = 0;

    // each warp owns a contiguous sub-chunk; lanes hold consecutive elements of each 32-element batch
    const int wbase = block * kSortChunk + warp * (kSortChunk / kSortWarps);
    uint32_t key[kSortItems], val[kSortItems];
#pragma unroll
    for (int j = 0; j < kSortItems; j++) {
        const int i = wbase + j * 32 + lane;
        if (PASS == 0) {
            key[j] = (i < n) ? keys_in[i] : 0xffffffffu;
            val[j] = (uint32_t)i;
        } else {
            const uint2 kv = (i < n) ? pairs_in[i] : make_uint2(0xffffffffu, 0u);
            key[j] = kv.x;
            val[j] = kv.y;
        }
    }
    {   // this warp's 512 counters
        uint4* z = reinterpret_cast<uint4*>(&s_cnt[warp][0]);
        z[lane] = make_uint4(0u, 0u, 0u, 0u);
        z[lane + 32] = make_uint4(0u, 0u, 0u, 0u);
    }
    ushort4 rect[PASS == 0 ? kSortItems : 1];
    if (PASS == 0) {
#pragma unroll
        for (int j = 0; j < kSortItems; j++) {
            const int i = wbase + j * 32 + lane;
            rect[j] = (i < n && key[j] != 0xffffffffu) ? tc.rects[i] : make_ushort4(0, 0, 0, 0);
        }
    }
    // exclusive scan of the global digit histogram: where each digit's keys start in the output (threads < 256)
    uint32_t gstart = 0;
    if (tid < kRBins) {
        const uint32_t v = ghist[PASS * kRBins + tid];
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(kFullMask, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        gstart = incl - v;
    }
    __syncthreads();
    if (tid < kRBins) {
        for (int w = 0; w < warp; w++) gstart += s_warp[w];
        if (PASS == 0 && block == 0 && tid == kRBins - 1)          // number of visible Gaussians, for the later passes
            counters[kCntVisible] = gstart + ghist[kRBins - 1];
    }
    if (PASS == 0) {   // corner updates of the tile-count difference array (shared-memory atomics; s_tdiff was zeroed above)
        const int pitch = tc.gx + 1;
#pragma unroll
        for (int j = 0; j < kSortItems; j++) {
            const ushort4 r = rect[j];
            if (r.z > r.x && r.w > r.y) {
                atomicAdd(&s_tdiff[r.y * pitch + r.x], 1);
                atomicAdd(&s_tdiff[r.y * pitch + r.z], -1);
                atomicAdd(&s_tdiff[r.w * pitch + r.x], -1);
                atomicAdd(&s_tdiff[r.w * pitch + r.z], 1);
            }
        }
    }

    // rank inside the warp's sub-chunk: all matches first (independent: eight MATCH.ANY in flight), then ONE dependent
    // sweep over the warp's counters.  Pass 0 ranks only the visible keys (culled ones are dropped here).
    unsigned peers[kSortItems];
    bool valid[kSortItems];
#pragma unroll
    for (int j = 0; j < kSortItems; j++) {
        const int i = wbase + j * 32 + lane;
        valid[j] = (i < n) && (PASS != 0 || key[j] != 0xffffffffu);
        const uint32_t d = valid[j] ? sort_digit(key[j], PASS) : 0xffffffffu;
        peers[j] = __match_any_sync(kFullMask, d);
        if (PASS < kSortDigits - 1) {   // histogram of the next pass's digit (its keys are in registers here)
            const uint32_t dn = valid[j] ? sort_digit(key[j], PASS + 1) : 0xffffffffu;
            if (PASS == kSortDigits - 2) {          // the top digit takes one or two values: aggregate over the warp first
                const unsigned pn = __match_any_sync(kFullMask, dn);
                if (valid[j] && lane == __ffs(pn) - 1) atomicAdd(&s_next[dn], (uint32_t)__popc(pn));
            } else if (valid[j]) {
                atomicAdd(&s_next[dn], 1u);
            }
        }
    }
    __syncwarp();
    uint32_t ofs[kSortItems];
#pragma unroll
    for (int j = 0; j < kSortItems; j++) {
        const uint32_t d = sort_digit(key[j], PASS);
        const int leader = __ffs(peers[j]) - 1;
        uint32_t old = 0;
        if (valid[j] && lane == leader) {
            old = s_cnt[warp][d];
            s_cnt[warp][d] = (uint16_t)(old + __popc(peers[j]));
        }
        ofs[j] = __shfl_sync(kFullMask, old, leader) + __popc(peers[j] & lanemask_lt());
        __syncwarp();
    }
    __syncthreads();
    // one digit per thread: prefix over the warps, the block's count -> publish -> local run starts -> look back
    uint32_t agg = 0;
    uint32_t* my_status = status + (size_t)block * kRBins;
    if (tid < kRBins) {
#pragma unroll
        for (int w = 0; w < kSortWarps; w++) {
            const uint32_t t = s_cnt[w][tid];
            s_cnt[w][tid] = (uint16_t)agg;
            agg += t;
        }
        st_status(my_status + tid, (block == 0 ? tag_incl : tag_agg) | agg);
        uint32_t incl = agg;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(kFullMask, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        s_lstart[tid] = incl - agg;     // warp-local for now
    }
    __syncthreads();
    if (tid < kRBins) {
        uint32_t wb = 0;
        for (int w = 0; w < warp; w++) wb += s_warp[w];
        const uint32_t lstart = s_lstart[tid] + wb;
        if (tid == kRBins - 1) s_total = lstart + agg;
        uint32_t excl = 0;
        int p = block - 1;
        while (p >= 0) {   // four predecessors per step; block 0 always publishes an inclusive value
            uint32_t v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) v[u] = (p - u >= 0) ? ld_status(status + (size_t)(p - u) * kRBins + tid) : tag_incl;
            int used = 0;
            bool done = false, stalled = false;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (done || stalled) continue;
                const uint32_t tag = v[u] & ~kStatMask;
                if (tag == tag_incl) {
                    excl += v[u] & kStatMask;
                    done = true;
                } else if (tag == tag_agg) {
                    excl += v[u] & kStatMask;
                    used++;
                } else {
                    stalled = true;
                }
            }
            if (done) break;
            p -= used;
            if (stalled) __nanosleep(20);
        }
        if (block > 0) st_status(my_status + tid, tag_incl | (excl + agg));
        s_lstart[tid] = lstart;
        s_gdst[tid] = gstart + excl - lstart;   // (mod 2^32) global position = s_gdst[digit] + local slot
    }
    __syncthreads();
    // reorder in shared memory: local slot = run start of the digit + keys of earlier warps + rank inside the warp
#pragma unroll
    for (int j = 0; j < kSortItems; j++) {
        if (valid[j]) {
            const uint32_t d = sort_digit(key[j], PASS);
            s_pairs[s_lstart[d] + s_cnt[warp][d] + ofs[j]] = make_uint2(key[j], val[j]);
        }
    }
    __syncthreads();
    const uint32_t total = s_total;
    for (uint32_t i = tid; i < total; i += kSortThreads) {
        const uint2 kv = s_pairs[i];
        pairs_out[s_gdst[sort_digit(kv.x, PASS)] + i] = kv;
    }
    if (PASS < kSortDigits - 1) {   // (every thread passed block barriers after the last s_next update)
        const uint32_t h = s_next[tid];
        if (h) atomicAdd(const_cast<uint32_t*>(ghist) + (PASS + 1) * kRBins + tid, h);
    }
    if (PASS == 0) {
        // flush this block's difference array; the block that finishes last turns the sums into the tile ranges
        int* mine = tc.tile_diff + (size_t)(block % kTileDiffReplicas) * tcells;
        for (int c = tid; c < tcells; c += kSortThreads) {
            const int v = s_tdiff[c];
            if (v != 0) atomicAdd(mine + c, v);
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) s_block = (atomicAdd(&counters[kCntTotalsDone], 1u) == (uint32_t)(sort_chunks_dev(n) - 1)) ? 1u : 0u;
        __syncthreads();
        if (s_block) {
            __threadfence();
            tile_ranges_from_diff(tc, s_tdiff, s_lstart, &s_total);   // (s_lstart: 256 free words for the 16 warp sums)
        }
    }
}

template <int PASS>
static void launch_sort_pass(const uint32_t* kin, const uint2* pin, int P, SortWS& w, uint32_t* counters, uint2* pout,
                             const TileCount& tc, cudaStream_t s)
{
    const size_t smem = sizeof(uint2) * kSortChunk + (PASS == 0 ? sizeof(int) * (size_t)(tc.gx + 1) * (tc.gy + 1) : 0);
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {   // 22 KB static + 32 KB (+ the difference array) dynamic: opt in
        cudaFuncSetAttribute(k_sort_pass<PASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        configured[dev] = true;
    }
    k_sort_pass<PASS><<<sort_chunks(P), kSortThreads, smem, s>>>(kin, pin, P, w.ghist, w.status, counters, pout, tc);
}

// Requires counters / ghist / status / tile_diff zeroed and digit 0 of ghist filled (k_preprocess_fwd).  Result:
// w.pairs_a (w.pairs_b if the fourth pass had to run, see sorted_list()) holds the counters[kCntVisible] visible Gaussians
// as {depth key, id} in (depth, index) order; ranges / w.tile_starts hold every tile's place in the instance list.
int launch_depth_sort(const uint32_t* depth_keys, const ushort4* rects, int P, int gx, int gy, SortWS& w, uint2* ranges,
                      uint32_t* counters, cudaStream_t s)
{
    if (P <= 0) return 0;
    if (sizeof(int) * (size_t)(gx + 1) * (gy + 1) > 140 * 1024) return -1;   // tile grid too large for pass 0's shared memory
    TileCount tc{rects, gx, gy, w.tile_diff, ranges, w.tile_starts};
    launch_sort_pass<0>(depth_keys, nullptr, P, w, counters, w.pairs_a, tc, s);
    launch_sort_pass<1>(nullptr, w.pairs_a, P, w, counters, w.pairs_b, tc, s);
    launch_sort_pass<2>(nullptr, w.pairs_b, P, w, counters, w.pairs_a, tc, s);
    launch_sort_pass<3>(nullptr, w.pairs_a, P, w, counters, w.pairs_b, tc, s);   // leaves at once unless a depth > 6553 exists
    return 0;
}

// ================================================================================================
// 2. Tile partition (R instances, generated on the fly from the depth-ordered Gaussians)
// ================================================================================================
// Walks the flattened instances of the 32 Gaussians held by the lanes of a warp (lane l owns a rect
// of n_l = w_l*h_l tiles), 32 instances per step in lane-major order, and calls f(tile, src_lane, valid)
// for every lane in every step.  `packed` = x0 | y0 << 10 | w << 20 per lane (x0,y0 < 1024, w < 4096).
// Non-empty lanes must precede empty ones (true for depth-sorted Gaussians: culled ones sort last),
// so the exclusive offsets of the non-empty lanes are strictly increasing and the source lane of
// instance i is found with one __reduce_or_sync + popc instead of a shuffle binary search.
// `kb` = index of the lane's first instance inside its splat (a warp's range may start mid-splat).
template <typename F>
__device__ __forceinline__ void warp_for_each_instance(uint32_t packed, uint32_t kb, uint32_t n_l, int gx, F&& f)
{
    const int lane = threadIdx.x & 31;
    uint32_t incl = n_l;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(kFullMask, incl, o);
        if (lane >= o) incl += t;
    }
    const uint32_t total = __shfl_sync(kFullMask, incl, 31);
    const uint32_t excl = incl - n_l;
    const uint32_t magic = div_magic(packed >> 20);   // exact k / w by multiplication (gsr_math.cuh)
    const unsigned le_mask = 0xffffffffu >> (31 - lane);
    for (uint32_t b = 0; b < total; b += 32) {
        const uint32_t i = b + lane;
        const bool valid = i < total;
        // lanes whose run starts inside this window mark their start; lanes that started before count as base
        const unsigned started_before = __ballot_sync(kFullMask, n_l != 0 && excl < b);
        const unsigned starts = __reduce_or_sync(kFullMask, (n_l != 0 && (excl - b) < 32u) ? (1u << (excl - b)) : 0u);
        const int src = max(__popc(started_before) + __popc(starts & le_mask) - 1, 0);
        const uint32_t s_excl = __shfl_sync(kFullMask, excl - kb, src);   // (i - s_excl) = index inside the splat
        const uint32_t s_pk = __shfl_sync(kFullMask, packed, src);
        const uint32_t s_magic = __shfl_sync(kFullMask, magic, src);
        uint32_t tile = 0xffffffffu;
        if (valid) {
            const uint32_t k = i - s_excl;
            const uint32_t w = s_pk >> 20, x0 = s_pk & 1023u, y0 = (s_pk >> 10) & 1023u;
            const uint32_t ry = div_by_magic(k, w, s_magic);
            const uint32_t rx = k - ry * w;
            tile = (y0 + ry) * (uint32_t)gx + x0 + rx;
        }
        f(tile, src, valid, i);
    }
}

__device__ __forceinline__ void load_rect(const uint2* __restrict__ perm, const ushort4* __restrict__ rects, int i,
                                          int n, uint32_t& id, uint32_t& packed, uint32_t& cnt)
{
    id = 0;
    packed = 1u << 20;
    cnt = 0;
    if (i < n) {
        id = perm[i].y;
        const ushort4 r = rects[id];
        const uint32_t w = r.z - r.x, h = r.w - r.y;   // culled Gaussians store an empty rect
        cnt = w * h;
        packed = (uint32_t)r.x | ((uint32_t)r.y << 10) | ((w ? w : 1u) << 20);
    }
}

// Rank of this lane's instance among the instances of the same tile this warp has seen so far, and count it.
// `cnt` / `tag`: the warp's private per-tile counters (16 bit) and scratch bytes in shared memory.  The instances
// of a step come from a handful of neighbouring splats and a splat's tiles are distinct, so two lanes rarely
// hold the same tile: every lane writes its lane id to tag[tile] and reads it back — if nobody lost that race all
// tiles are distinct and the rank is a plain (non-atomic) read-increment-write; only otherwise are same-tile lanes
// matched (one ballot per tile-id bit) and ranked in lane order.  No shared-memory atomics (2 cycles per lane).
// `delta` (the same for every lane of a call): what each instance adds to its counter — 1 for ranking, +-1 (mod 2^16)
// for the corner updates of a difference array.
__device__ __forceinline__ uint32_t warp_rank_tile(uint16_t* cnt, uint8_t* tag, uint32_t tile, bool valid, int tbits,
                                                   uint32_t delta = 1u)
{
    const int lane = threadIdx.x & 31;
    if (valid) tag[tile] = (uint8_t)lane;
    __syncwarp();
    const bool lost = valid && tag[tile] != (uint8_t)lane;
    const uint32_t old = valid ? cnt[tile] : 0u;
    uint32_t rank = old;
    if (!__any_sync(kFullMask, lost)) {
        if (valid) cnt[tile] = (uint16_t)(old + delta);
    } else {
        const unsigned peers = match_any_bits(tile, valid, tbits);
        __syncwarp();
        if (valid && lane == __ffs(peers) - 1) cnt[tile] = (uint16_t)(old + delta * (uint32_t)__popc(peers));
        rank = old + __popc(peers & lanemask_lt());
    }
    __syncwarp();
    return rank;
}

// The tile partition.  ONE kernel: the per-tile instance counts (and with them every tile's range in the final list)
// are already known — k_preprocess_fwd accumulates them as a 2-D difference array and its last CTA integrates it —
// so the only global information a CTA lacks is how many instances of each tile belong to the CTAs before it, and
// that arrives by decoupled look-back over per-tile rows (the depth sort's scheme with T bins instead of 256).
//
// A CTA takes a ticket c and owns the c-th contiguous chunk of depth-ordered Gaussians.  It loads their tile
// rectangles, scans the instance counts in shared memory and gives every warp an EQUAL share of the chunk's
// instances (a share may start and end in the middle of a splat — a splat's tiles are distinct, so any
// cut keeps the per-tile order intact).  Splats covering hundreds of tiles would otherwise serialise
// the one warp that owns them.  Each warp enumerates its share twice:
//   phase 1 counts its instances per tile (private 16-bit counters for all T tiles in shared memory — what the
//           B200's 227 KB buys); per tile, the counts are then prefixed over the CTA's warps, their sum is published
//           as row c of the look-back state and the look-back returns the number of earlier instances of the tile;
//   phase 2 ranks every instance among the CTA's earlier instances of the same tile (prefix + running count) and
//           stores its Gaussian id at  point_list[tile start + earlier CTAs + rank]  — the final position.
// No intermediate instance stream, no histogram matrix scan, no second kernel.
// Dynamic shared memory: uint32 s_id[per_cta], s_pk[per_cta], s_ex[per_cta + 1], s_base[T]; uint16 s_cnt[nwarps][T];
// uint8 s_tag[nwarps][T].
template <typename F>
__device__ __forceinline__ void warp_share_for_each(const uint32_t* s_id, const uint32_t* s_pk, const uint32_t* s_ex,
                                                    int per_cta, int first, uint32_t a, uint32_t b, int gx, F&& f)
{
    const int lane = threadIdx.x & 31;
    uint32_t done = 0;
    for (int g0 = first; g0 < per_cta && s_ex[g0] < b; g0 += 32) {
        const int g = g0 + lane;
        uint32_t id = 0, packed = 1u << 20, kb = 0, cnt = 0;
        if (g < per_cta) {
            const uint32_t ex = s_ex[g], full = s_ex[g + 1] - ex;
            if (ex < b && full != 0) {
                kb = a > ex ? a - ex : 0u;
                const uint32_t ke = min(full, b - ex);
                cnt = ke - kb;
                id = s_id[g];
                packed = s_pk[g];
            }
        }
        const uint32_t round_total = __reduce_add_sync(kFullMask, cnt);
        warp_for_each_instance(packed, kb, cnt, gx, [&](uint32_t tile, int src, bool valid, uint32_t i) {
            f(tile, __shfl_sync(kFullMask, id, src), valid, done + i);
        });
        done += round_total;
    }
}

constexpr uint32_t kTileAgg = 1u << 30, kTileIncl = 2u << 30, kTileVal = (1u << 30) - 1u;
constexpr int kTileLook = 8;     // predecessors per look-back step

__global__ void __launch_bounds__(1024) k_tile_partition(const uint2* __restrict__ perm_a, const uint2* __restrict__ perm_b,
                                                        const uint32_t* __restrict__ ghist, int cap_cta,
                                                        const ushort4* __restrict__ rects, int gx, int T,
                                                        uint32_t* status, const uint32_t* __restrict__ tile_starts,
                                                        uint32_t* __restrict__ point_list, uint32_t* __restrict__ counters,
                                                        uint32_t cap)
{
    if (counters[kCntR] > cap) return;   // instance list does not fit the caller's workspace: nothing is built
    const int n = (int)counters[kCntVisible];          // the depth sort kept only the visible Gaussians
    // the sorted list is where the last sort pass that ran left it (the fourth one runs only for depths beyond 6553)
    const uint2* __restrict__ perm = (ghist[(kSortDigits - 1) * kSortBins] == (uint32_t)n) ? perm_a : perm_b;
    extern __shared__ uint32_t s_dyn[];
    __shared__ uint32_t s_wsum[32];
    __shared__ int s_block;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    if (tid == 0) s_block = (int)atomicAdd(&counters[kCntPartTicket], 1u);
    __syncthreads();
    const int block = s_block;
    // chunk size: the visible Gaussians divided evenly over the grid (the host sized the grid as whole waves of CTAs
    // without knowing how many Gaussians are visible), at most what the shared-memory layout was sized for
    const int per_cta = min(cap_cta, ((n + (int)gridDim.x - 1) / (int)gridDim.x + 31) & ~31);
    const int c0 = block * per_cta;
    if (c0 >= n) return;                   // (so is every later ticket: nobody will look back at this row)
    uint32_t* s_id = s_dyn;
    uint32_t* s_pk = s_id + cap_cta;
    uint32_t* s_ex = s_pk + cap_cta;       // [per_cta + 1] exclusive instance offsets inside the chunk
    uint32_t* s_base = s_ex + ((cap_cta + 1 + 3) & ~3);
    const int Tp = (T + 7) & ~7;           // row pitch: keeps every row 16-byte aligned
    uint16_t* s_cnt = reinterpret_cast<uint16_t*>(s_base + Tp);
    uint8_t* s_tag = reinterpret_cast<uint8_t*>(s_cnt + (size_t)nwarps * Tp);
    {   // zero the counters (the tags need no initial value)
        uint4* z = reinterpret_cast<uint4*>(s_cnt);
        const int n16 = nwarps * Tp / 8;
        for (int i = tid; i < n16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    const int m = min(per_cta, n - c0);    // Gaussians in this chunk
    // load + block-wide exclusive scan of the counts (thread t owns the contiguous slice [t*q, (t+1)*q))
    const int q = (per_cta + blockDim.x - 1) / blockDim.x;
    uint32_t run = 0;
    for (int j = 0; j < q; j++) {
        const int g = tid * q + j;
        if (g < per_cta) {
            uint32_t id, packed, cnt;
            load_rect(perm, rects, g < m ? c0 + g : n, n, id, packed, cnt);
            s_id[g] = id;
            s_pk[g] = packed;
            s_ex[g] = run;                 // thread-local exclusive offset, fixed up below
            run += cnt;
        }
    }
    uint32_t incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(kFullMask, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    uint32_t tbase = incl - run;
    for (int w = 0; w < warp; w++) tbase += s_wsum[w];
    uint32_t S = 0;
    for (int w = 0; w < nwarps; w++) S += s_wsum[w];
    for (int j = 0; j < q; j++) {
        const int g = tid * q + j;
        if (g < per_cta) s_ex[g] += tbase;
    }
    if (tid == 0) s_ex[per_cta] = S;
    __syncthreads();

    // this warp's share [a, b) of the chunk's S instances
    const uint32_t a = (uint32_t)(((uint64_t)S * warp) / nwarps), b = (uint32_t)(((uint64_t)S * (warp + 1)) / nwarps);
    const uint32_t len = b - a;
    int first = 0;
    if (len != 0) {
        // first splat of the share: last g with s_ex[g] <= a  (s_ex is non-decreasing; zero-count splats only at the end)
        int lo = 0, hi = per_cta;          // invariant: s_ex[lo] <= a < s_ex[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (s_ex[mid] <= a) lo = mid; else hi = mid;
        }
        first = lo;
    }
    uint16_t* my = s_cnt + (size_t)warp * Tp;
    uint8_t* my_tag = s_tag + (size_t)warp * Tp;
    int tbits = 1;
    while ((1 << tbits) < T) tbits++;
    if (len != 0)
        warp_share_for_each(s_id, s_pk, s_ex, per_cta, first, a, b, gx, [&](uint32_t tile, uint32_t, bool valid, uint32_t) {
            warp_rank_tile(my, my_tag, tile, valid, tbits);
        });
    __syncthreads();
    // per tile: exclusive prefix over the warps; the sum is this CTA's count
    for (int t = tid; t < T; t += blockDim.x) {
        uint32_t acc = 0;
#pragma unroll 8
        for (int w = 0; w < nwarps; w++) {
            const uint32_t c = s_cnt[(size_t)w * Tp + t];
            s_cnt[(size_t)w * Tp + t] = (uint16_t)acc;
            acc += c;
        }
        s_base[t] = acc;
    }
    __syncthreads();
    // publish, look back, final base — FOUR tiles per thread with 16-byte accesses: with thousands of tiles and few
    // warps (1920x1080: 8160 tiles, 4 warps) the look-back is the CTA's longest phase, and its cost is the number of
    // dependent L2 round trips per thread.  Every 32-bit word carries its own tag, so a torn 16-byte store is harmless.
    const int pitch = (T + 3) & ~3;
    uint32_t* my_status = status + (size_t)block * pitch;
    for (int q4 = tid; q4 * 4 < T; q4 += blockDim.x) {
        const int t0 = q4 * 4;
        uint32_t agg[4], excl[4] = {0u, 0u, 0u, 0u};
        bool fin[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            agg[c] = (t0 + c < T) ? s_base[t0 + c] : 0u;
            fin[c] = false;
        }
        const uint32_t tag0 = block == 0 ? kTileIncl : kTileAgg;
        st_status4(my_status + t0, make_uint4(tag0 | agg[0], tag0 | agg[1], tag0 | agg[2], tag0 | agg[3]));
        int p = block - 1;
        while (p >= 0 && !(fin[0] && fin[1] && fin[2] && fin[3])) {   // kTileLook predecessors per step
            uint4 v[kTileLook];
#pragma unroll
            for (int u = 0; u < kTileLook; u++)
                v[u] = (p - u >= 0) ? ld_status4(status + (size_t)(p - u) * pitch + t0)
                                    : make_uint4(kTileIncl, kTileIncl, kTileIncl, kTileIncl);   // before block 0: nothing
            int used = 0;
            bool stalled = false;
#pragma unroll
            for (int u = 0; u < kTileLook; u++) {
                if (stalled) continue;
                const uint32_t w4[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
                if (!(w4[0] & ~kTileVal) || !(w4[1] & ~kTileVal) || !(w4[2] & ~kTileVal) || !(w4[3] & ~kTileVal)) {
                    stalled = true;          // not (completely) published yet
                    continue;
                }
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    if (!fin[c]) {
                        excl[c] += w4[c] & kTileVal;
                        if ((w4[c] & ~kTileVal) == kTileIncl) fin[c] = true;
                    }
                }
                used++;
            }
            p -= used;
            if (stalled) __nanosleep(40);
        }
        if (block > 0)
            st_status4(my_status + t0, make_uint4(kTileIncl | (excl[0] + agg[0]), kTileIncl | (excl[1] + agg[1]),
                                                  kTileIncl | (excl[2] + agg[2]), kTileIncl | (excl[3] + agg[3])));
#pragma unroll
        for (int c = 0; c < 4; c++)
            if (t0 + c < T) s_base[t0 + c] = tile_starts[t0 + c] + excl[c];
    }
    __syncthreads();
    if (len != 0)
        warp_share_for_each(s_id, s_pk, s_ex, per_cta, first, a, b, gx, [&](uint32_t tile, uint32_t gid, bool valid, uint32_t) {
            const uint32_t rank = warp_rank_tile(my, my_tag, tile, valid, tbits);
            if (valid) point_list[s_base[tile] + rank] = gid;
        });
}

// Chunking of the tile partition: ranks are 16 bit and relative to the CTA, so a CTA may own at most 65535
// Gaussians; the CTA count is capped so that the look-back state stays small.
static size_t tile_partition_smem(int T, int per_cta, int warps)
{
    const size_t Tp = (size_t)((T + 7) & ~7);
    return (size_t)(2 * per_cta + ((per_cta + 1 + 3) & ~3)) * 4 + Tp * 4 + Tp * 3 * (size_t)warps;
}
void tile_partition_plan(int P, int T, int& ctas, int& per_cta, int& warps, size_t& smem)
{
    // as many warps per CTA as their private per-tile counters (3 bytes per tile) allow in ~200 KB of shared memory
    // next to the chunk's rectangles, at most 32: the kernel is a latency-bound walk, so short per-warp shares matter
    // more than anything else.  `per_cta` is the CAPACITY of a chunk (what shared memory is laid out for); the kernel
    // divides the visible Gaussians evenly over the grid, and the grid is a whole number of waves of CTAs (one CTA per SM)
    // so that no partial last wave leaves most SMs idle: 346 chunks of 2048 on 148 SMs were 2.3 waves.
    static int waves = 0;
    if (waves == 0) {
        const char* e = getenv("GSR_PART_WAVES");
        waves = e ? atoi(e) : 2;
        if (waves < 1) waves = 1;
    }
    warps = 32;
    for (;;) {
        per_cta = 4096;
        while (tile_partition_smem(T, per_cta, warps) > 200 * 1024 && per_cta > 1024) per_cta -= 256;
        if (tile_partition_smem(T, per_cta, warps) <= 200 * 1024 || warps == 1) break;
        warps >>= 1;
    }
    if (tile_partition_smem(T, per_cta, warps) > 200 * 1024 || T > 65535) warps = 0;   // does not fit at all
    const int need = P > 0 ? (P + per_cta - 1) / per_cta : 0;                 // every Gaussian may be visible
    const int want = P > 0 ? std::min(waves * device_sm_count(), (P + 1023) / 1024) : 0;
    ctas = std::max(need, want);
    smem = tile_partition_smem(T, per_cta, warps > 0 ? warps : 1);
}

// `cap`: number of instances the caller's point_list can hold.  If the scene has more (counters[kCntR], known on the
// device only), the kernel returns without touching it.
int launch_tile_partition(int P, const ushort4* rects, int gx, int gy, SortWS& w,
                          uint32_t* counters, uint32_t cap, uint32_t* point_list, cudaStream_t s)
{
    const int T = gx * gy;
    int ctas, per_cta, warps;
    size_t smem;
    tile_partition_plan(P, T, ctas, per_cta, warps, smem);
    if (warps == 0 || smem > 220 * 1024 || per_cta > 65535) return -1;   // image / scene too large
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {   // opt in to the large B200 carve-out once per device
        cudaFuncSetAttribute(k_tile_partition, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        configured[dev] = true;
    }
    k_tile_partition<<<ctas, warps * 32, smem, s>>>(w.pairs_a, w.pairs_b, w.ghist, per_cta, rects, gx, T, w.tile_status,
                                                    w.tile_starts, point_list, counters, cap);
    return 0;
}

}  // namespace gsr
