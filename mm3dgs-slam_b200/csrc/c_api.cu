// c_api.cu — the extern "C" boundary declared in include/gsrast_b200.h.
//
// Host-side orchestration only: carve the caller's workspaces, enqueue the kernels on the
// caller's stream in the order of CudaRasterizer::Rasterizer::forward / backward
// (CR/rasterizer_impl.cu:198-336, :340-434).  No state is kept between calls.
#include "gsr_internal.cuh"
#include <atomic>
#include <mutex>
#include <string>
#include <vector>
#include <stdio.h>
#include <string.h>

namespace gsr {

static thread_local std::string g_err;

// ---- optional stage profiler (off by default; CUDA events on the caller's stream) -----------------
// Stage ids are stable and named by gsr_profile_stage_name().
enum Stage { ST_BEGIN = -1, ST_PREPROCESS = 0, ST_DEPTH_SORT, ST_TILE_PARTITION, ST_RENDER, ST_RENDER_BWD,
             ST_PREPROCESS_BWD, ST_COUNT };
static const char* kStageNames[ST_COUNT] = {"preprocess_fwd", "depth_sort", "tile_partition", "render_fwd",
                                            "render_bwd", "preprocess_bwd"};
static std::atomic<int> g_prof_on{0};
static std::atomic<long long> g_launches{0};
static std::mutex g_prof_mu;
struct Mark { int stage; cudaEvent_t ev; };
static std::vector<Mark> g_marks;
static std::vector<cudaEvent_t> g_pool;

static void prof_mark(int stage, cudaStream_t s)
{
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    cudaEvent_t ev;
    if (!g_pool.empty()) { ev = g_pool.back(); g_pool.pop_back(); }
    else if (cudaEventCreate(&ev) != cudaSuccess) return;
    cudaEventRecord(ev, s);
    g_marks.push_back({stage, ev});
}

static int fail(int code, const char* what, cudaError_t e = cudaSuccess)
{
    g_err = what;
    if (e != cudaSuccess) {
        g_err += ": ";
        g_err += cudaGetErrorString(e);
    }
    return code;
}

int api_fail(int code, const char* what, cudaError_t e) { return fail(code, what, e); }

int device_sm_count()
{
    static std::atomic<int> cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    int n = cached[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}
void api_count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define GSR_CUDA(expr)                                                     \
    do {                                                                   \
        cudaError_t _e = (expr);                                           \
        if (_e != cudaSuccess) return fail(GSR_ERR_CUDA, #expr, _e);       \
    } while (0)

// After each stage: always catch launch-configuration errors; in debug mode also synchronise so
// that execution errors surface at the stage that caused them (CR/auxiliary.h:166-173).
#define GSR_STAGE(name, debug, stream)                                                   \
    do {                                                                                 \
        cudaError_t _e = cudaGetLastError();                                             \
        if (_e == cudaSuccess && (debug)) _e = cudaStreamSynchronize(stream);            \
        if (_e != cudaSuccess) return fail(GSR_ERR_CUDA, "stage " name, _e);             \
    } while (0)
#define GSR_MARK(stage, stream, nkernels)                                                \
    do {                                                                                 \
        g_launches.fetch_add((nkernels), std::memory_order_relaxed);                     \
        prof_mark((stage), (stream));                                                    \
    } while (0)

template <typename T>
static T* carve(char*& p, size_t count)
{
    T* r = reinterpret_cast<T*>(p);
    p += align_up(count * sizeof(T));
    return r;
}

GeomWS geom_ws_carve(char* base, int P, int W, int H)
{
    GeomWS w;
    char* p = base;
    const size_t n = P > 0 ? (size_t)P : 1;
    const int T = ((W + kTile - 1) / kTile) * ((H + kTile - 1) / kTile);
    w.rec = carve<float4>(p, n * 3);
    w.rects = carve<ushort4>(p, n);
    w.depth_keys = carve<uint32_t>(p, n);
    // counters | ghist | tile_diff are contiguous: one small memset zeroes them at the start of a forward
    int ctas, per_cta, warps;
    size_t smem;
    tile_partition_plan((int)n, T, ctas, per_cta, warps, smem);
    w.counters = carve<uint32_t>(p, 64);          // [63] = CTA counter of the camera-gradient reduction
    w.sort.ghist = carve<uint32_t>(p, kSortDigits * kSortBins);
    w.sort.tile_diff = carve<int>(p, (size_t)kTileDiffReplicas * ((W + kTile - 1) / kTile + 1) * ((H + kTile - 1) / kTile + 1));
    w.zero_bytes = (size_t)(p - reinterpret_cast<char*>(w.counters));
    // look-back state of the depth sort and of the tile partition, contiguous: zeroed by the projection kernel itself
    // (a few MB of stores spread over its CTAs instead of a separate memset in front of it)
    w.sort.status = reinterpret_cast<uint32_t*>(p);
    carve<uint32_t>(p, (size_t)sort_chunks((int)n) * kSortBins);
    w.sort.tile_status = carve<uint32_t>(p, (size_t)ctas * ((T + 3) & ~3));   // rows padded to 16 bytes
    w.sort.status_words = (size_t)(p - reinterpret_cast<char*>(w.sort.status)) / sizeof(uint32_t);
    w.extra_gen = carve<float>(p, n * 3);
    w.sort.pairs_a = carve<uint2>(p, n);
    w.sort.pairs_b = carve<uint2>(p, n);
    w.cam_partials = carve<double>(p, (size_t)kCamPartialRows * 35);
    w.sort.tile_starts = carve<uint32_t>(p, (size_t)T);
    w.total = (size_t)(p - base);
    return w;
}

ImgWS img_ws_carve(char* base, int W, int H)
{
    ImgWS w;
    char* p = base;
    const size_t N = (size_t)W * H;
    const size_t tiles = (size_t)((W + kTile - 1) / kTile) * ((H + kTile - 1) / kTile);
    w.final_T = carve<float>(p, N);
    w.n_contrib = carve<uint32_t>(p, N);
    w.ranges = carve<uint2>(p, tiles);
    w.total = (size_t)(p - base);
    return w;
}

BinWS bin_ws_carve(char* base, int64_t R)
{
    BinWS w;
    char* p = base;
    w.point_list = carve<uint32_t>(p, R > 0 ? (size_t)R : 1);
    w.contrib = carve<uint8_t>(p, R > 0 ? (size_t)R : 1);
    w.total = (size_t)(p - base);
    return w;
}

// Pinned 4-byte slot + event per (host thread, device) for the one device->host hand-off of R.
struct HostSlot {
    int32_t* pinned = nullptr;
    cudaEvent_t ev = nullptr;
};
static HostSlot& host_slot()
{
    static thread_local HostSlot slots[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    HostSlot& s = slots[dev];
    if (!s.pinned) {
        if (cudaHostAlloc((void**)&s.pinned, 64, cudaHostAllocDefault) != cudaSuccess) s.pinned = nullptr;
        if (cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming) != cudaSuccess) s.ev = nullptr;
    }
    return s;
}

static int check_common(const gsr_gaussians* g, const gsr_camera* cam)
{
    if (!g || !cam) return fail(GSR_ERR_INVALID, "null gaussians/camera struct");
    if (g->P < 0) return fail(GSR_ERR_INVALID, "P < 0");
    if (cam->width <= 0 || cam->height <= 0) return fail(GSR_ERR_INVALID, "image size must be positive");
    if (g->P > 0) {
        if (!g->means3D || !g->opacities) return fail(GSR_ERR_INVALID, "means3D / opacities required");
        if ((g->shs == nullptr) == (g->colors_precomp == nullptr))
            return fail(GSR_ERR_INVALID, "provide exactly one of shs / colors_precomp");
        const bool sr = g->scales != nullptr && g->rotations != nullptr;
        if (sr == (g->cov3D_precomp != nullptr) || ((g->scales != nullptr) != (g->rotations != nullptr)))
            return fail(GSR_ERR_INVALID, "provide exactly one of (scales, rotations) / cov3D_precomp");
        if (g->shs && (g->sh_coeffs < 1 || g->sh_degree < 0 || g->sh_degree > 3 ||
                       (g->sh_degree + 1) * (g->sh_degree + 1) > g->sh_coeffs))
            return fail(GSR_ERR_INVALID, "sh_degree / sh_coeffs inconsistent");
        if (!cam->viewmatrix || !cam->projmatrix || !cam->campos || !cam->background)
            return fail(GSR_ERR_INVALID, "camera pointers required");
        if (g->extra_mode != 0 && (g->extra_mode != 1 || g->extra_colors))
            return fail(GSR_ERR_INVALID, "extra_mode must be 0, or 1 with extra_colors == NULL");
        if (g->raw_params != 0 && (g->raw_params != 1 || !sr))
            return fail(GSR_ERR_INVALID, "raw_params must be 0, or 1 with scales + rotations");
    }
    return GSR_OK;
}

}  // namespace gsr

using namespace gsr;

extern "C" {

int gsr_abi_version(void) { return GSR_ABI_VERSION; }

long long gsr_launch_count(void) { return g_launches.load(); }
int gsr_profile_num_stages(void) { return ST_COUNT; }
const char* gsr_profile_stage_name(int i) { return (i >= 0 && i < ST_COUNT) ? kStageNames[i] : ""; }
void gsr_profile_enable(int on) { g_prof_on.store(on ? 1 : 0); }
int gsr_profile_collect(float* ms, int* counts)
{
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (int i = 0; i < ST_COUNT; i++) { ms[i] = 0.f; counts[i] = 0; }
    if (!g_marks.empty()) {
        cudaError_t e = cudaEventSynchronize(g_marks.back().ev);
        if (e != cudaSuccess) return fail(GSR_ERR_CUDA, "profile sync", e);
    }
    for (size_t i = 1; i < g_marks.size(); i++) {
        const int st = g_marks[i].stage;
        if (st < 0) continue;
        float t = 0.f;
        if (cudaEventElapsedTime(&t, g_marks[i - 1].ev, g_marks[i].ev) == cudaSuccess) { ms[st] += t; counts[st]++; }
    }
    for (auto& m : g_marks) g_pool.push_back(m.ev);
    g_marks.clear();
    return GSR_OK;
}
const char* gsr_last_error(void) { return g_err.c_str(); }

size_t gsr_geom_ws_bytes(int32_t P, int32_t W, int32_t H) { return geom_ws_carve(nullptr, P, W, H).total; }
size_t gsr_img_ws_bytes(int32_t W, int32_t H) { return img_ws_carve(nullptr, W, H).total; }
size_t gsr_binning_ws_bytes(int64_t R) { return bin_ws_carve(nullptr, R).total; }

void gsr_partition_plan_of(int32_t P, int32_t W, int32_t H, gsr_partition_plan* o)
{
    int ctas = 0, per_cta = 0, warps = 0;
    size_t smem = 0;
    const int T = ((W + kTile - 1) / kTile) * ((H + kTile - 1) / kTile);
    tile_partition_plan(P > 0 ? P : 1, T, ctas, per_cta, warps, smem);
    o->ctas = ctas; o->chunk_capacity = per_cta; o->warps = warps; o->_pad = 0; o->smem_bytes = smem;
}

void gsr_geom_layout_of(int32_t P, int32_t W, int32_t H, gsr_geom_layout* o)
{
    GeomWS w = geom_ws_carve(nullptr, P, W, H);
    o->rec = (size_t)w.rec;
    o->rects = (size_t)w.rects;
    o->depth_keys = (size_t)w.depth_keys;
    o->sorted_ids = (size_t)w.sort.pairs_a;
    o->counters = (size_t)w.counters;
    o->total = w.total;
}
void gsr_img_layout_of(int32_t W, int32_t H, gsr_img_layout* o)
{
    ImgWS w = img_ws_carve(nullptr, W, H);
    o->final_T = (size_t)w.final_T;
    o->n_contrib = (size_t)w.n_contrib;
    o->ranges = (size_t)w.ranges;
    o->total = w.total;
}
void gsr_binning_layout_of(int64_t R, gsr_binning_layout* o)
{
    BinWS w = bin_ws_carve(nullptr, R);
    o->point_list = (size_t)w.point_list;
    o->total = w.total;
}

int gsr_forward_preprocess(gsr_stream_t stream_, const gsr_gaussians* g, const gsr_camera* cam, int32_t* radii,
                           void* geom_ws, size_t geom_ws_bytes, void* img_ws, size_t img_ws_bytes,
                           int32_t* num_rendered)
{
    if (int rc = check_common(g, cam)) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int P = g->P, W = cam->width, H = cam->height;
    if (num_rendered) *num_rendered = 0;
    if (P == 0) return GSR_OK;
    if (!radii || !geom_ws || !img_ws) return fail(GSR_ERR_INVALID, "radii / workspaces required");
    if (geom_ws_bytes < gsr_geom_ws_bytes(P, W, H) || img_ws_bytes < gsr_img_ws_bytes(W, H))
        return fail(GSR_ERR_WORKSPACE, "geometry / image workspace too small");
    if ((W + kTile - 1) / kTile > 1023 || (H + kTile - 1) / kTile > 1023)
        return fail(GSR_ERR_INVALID, "image larger than 16368 pixels in one dimension");
    GeomWS gw = geom_ws_carve((char*)geom_ws, P, W, H);

    PreArgs a;
    a.P = P; a.D = g->sh_degree; a.M = g->shs ? g->sh_coeffs : 0; a.W = W; a.H = H;
    a.gx = (W + kTile - 1) / kTile; a.gy = (H + kTile - 1) / kTile; a.prefiltered = cam->prefiltered;
    a.raw = g->raw_params;
    a.means = g->means3D; a.scales = g->scales; a.rots = g->rotations; a.opac = g->opacities; a.shs = g->shs;
    a.colors = g->colors_precomp; a.cov3D_pre = g->cov3D_precomp;
    a.view = cam->viewmatrix; a.proj = cam->projmatrix; a.campos = cam->campos;
    a.scale_mod = g->scale_modifier; a.tanfovx = cam->tanfovx; a.tanfovy = cam->tanfovy;
    a.focal_y = H / (2.0f * cam->tanfovy);
    a.focal_x = W / (2.0f * cam->tanfovx);
    a.radii = radii; a.rec = gw.rec; a.rects = gw.rects; a.depth_keys = gw.depth_keys; a.num_rendered = gw.counters + kCntR;
    a.ghist = gw.sort.ghist; a.status = gw.sort.status; a.status_words = gw.sort.status_words;
    a.chunk_ticket = gw.counters + kCntChunkFwd;
    a.extra_gen = g->extra_mode == 1 ? gw.extra_gen : nullptr;
    prof_mark(ST_BEGIN, stream);
    // counters, digit histograms, tile-count difference array (~50 KB)
    GSR_CUDA(cudaMemsetAsync(gw.counters, 0, gw.zero_bytes, stream));
    launch_preprocess_fwd(a, stream);
    GSR_STAGE("preprocess", cam->debug, stream);
    GSR_MARK(ST_PREPROCESS, stream, 1);
    // R goes to the host through a pinned slot + event, so the depth sort (which does not need R) is
    // already running while the caller waits for R and allocates the instance list.
    HostSlot& hs = host_slot();
    const bool async_r = num_rendered && hs.pinned && hs.ev;
    if (async_r) {
        GSR_CUDA(cudaMemcpyAsync(hs.pinned, gw.counters, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        GSR_CUDA(cudaEventRecord(hs.ev, stream));
    }
    // (pass 0 also derives the per-tile instance counts and tile ranges from the rectangles, so that the tile
    // partition can place every instance at its final position in one kernel)
    if (launch_depth_sort(gw.depth_keys, gw.rects, P, a.gx, a.gy, gw.sort, img_ws_carve((char*)img_ws, W, H).ranges,
                          gw.counters, stream) != 0)
        return fail(GSR_ERR_INVALID, "tile grid too large (more than ~45k tiles)");
    GSR_STAGE("depth_sort", cam->debug, stream);
    GSR_MARK(ST_DEPTH_SORT, stream, 4);
    if (async_r) {
        GSR_CUDA(cudaEventSynchronize(hs.ev));
        *num_rendered = *hs.pinned;
    } else if (num_rendered) {
        uint32_t r = 0;
        GSR_CUDA(cudaMemcpyAsync(&r, gw.counters, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        GSR_CUDA(cudaStreamSynchronize(stream));
        *num_rendered = (int32_t)r;
    }
    if (num_rendered && *num_rendered < 0) return fail(GSR_ERR_INVALID, "more than 2^31-1 tile instances");
    return GSR_OK;
}

int gsr_forward_render(gsr_stream_t stream_, const gsr_gaussians* g, const gsr_camera* cam, const int32_t* radii,
                       int64_t R, void* geom_ws, void* binning_ws, size_t binning_ws_bytes, void* img_ws,
                       float* out_color, float* out_extra)
{
    if (int rc = check_common(g, cam)) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int P = g->P, W = cam->width, H = cam->height;
    if (!out_color) return fail(GSR_ERR_INVALID, "out_color required");
    if ((g->extra_colors || g->extra_mode == 1) && !out_extra)
        return fail(GSR_ERR_INVALID, "out_extra required with extra colours");
    if (R < 0 || R > 0x7fffffffLL) return fail(GSR_ERR_INVALID, "num_rendered out of range");
    if (P == 0) {  // reference: zero image, nothing else (DGR/rasterize_points.cu:81)
        GSR_CUDA(cudaMemsetAsync(out_color, 0, sizeof(float) * 3 * (size_t)W * H, stream));
        if (out_extra) GSR_CUDA(cudaMemsetAsync(out_extra, 0, sizeof(float) * 3 * (size_t)W * H, stream));
        return GSR_OK;
    }
    if (!radii || !geom_ws || !img_ws || (R > 0 && !binning_ws))
        return fail(GSR_ERR_INVALID, "radii / workspaces required");
    if (binning_ws_bytes < gsr_binning_ws_bytes(R)) return fail(GSR_ERR_WORKSPACE, "binning workspace too small");
    GeomWS gw = geom_ws_carve((char*)geom_ws, P, W, H);
    ImgWS iw = img_ws_carve((char*)img_ws, W, H);
    BinWS bw = bin_ws_carve((char*)binning_ws, R);
    const int gx = (W + kTile - 1) / kTile, gy = (H + kTile - 1) / kTile;

    prof_mark(ST_BEGIN, stream);
    if (launch_tile_partition(P, gw.rects, gx, gy, gw.sort, gw.counters, (uint32_t)R, bw.point_list, stream) != 0)
        return fail(GSR_ERR_INVALID, "tile grid too large for the shared-memory tile partition (more than ~26k tiles of 16x16 pixels)");
    GSR_STAGE("tile_partition", cam->debug, stream);
    GSR_MARK(ST_TILE_PARTITION, stream, 1);
    launch_render_fwd(W, H, gx, gy, iw.ranges, bw.point_list, gw.rec, cam->background, iw.final_T, iw.n_contrib,
                      out_color, bw.contrib, g->extra_mode == 1 ? gw.extra_gen : g->extra_colors, out_extra,
                      gw.counters, (uint32_t)R, stream);
    GSR_STAGE("render", cam->debug, stream);
    GSR_MARK(ST_RENDER, stream, 1);
    return GSR_OK;
}

int gsr_backward(gsr_stream_t stream_, const gsr_gaussians* g, const gsr_camera* cam, const int32_t* radii, int64_t R,
                 const void* geom_ws, const void* binning_ws, const void* img_ws, const float* dL_dpixels,
                 const float* dL_dpixels_extra, const gsr_grads* gr)
{
    if (int rc = check_common(g, cam)) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int P = g->P, W = cam->width, H = cam->height;
    if (P == 0) return GSR_OK;
    if (!gr || !dL_dpixels || !radii || !geom_ws || !img_ws) return fail(GSR_ERR_INVALID, "null argument");
    if (!gr->dL_dmeans2D || !gr->dL_dconic || !gr->dL_dopacity || !gr->dL_dcolors || !gr->dL_dmeans3D)
        return fail(GSR_ERR_INVALID, "required gradient buffer missing");
    const float* extra = g->extra_mode == 1 ? geom_ws_carve((char*)geom_ws, P, W, H).extra_gen : g->extra_colors;
    if (extra && (!dL_dpixels_extra || !gr->dL_dextra))
        return fail(GSR_ERR_INVALID, "dL_dpixels_extra / dL_dextra required with extra colours");
    const int M = g->shs ? g->sh_coeffs : 0;
    if ((M > 0 && !gr->dL_dsh) || (g->scales && (!gr->dL_dscales || !gr->dL_drotations)))
        return fail(GSR_ERR_INVALID, "gradient buffer for a provided input missing");
    if (g->raw_params && !gr->dL_dopacity_raw) return fail(GSR_ERR_INVALID, "dL_dopacity_raw required with raw_params");
    GeomWS gw = geom_ws_carve((char*)geom_ws, P, W, H);
    ImgWS iw = img_ws_carve((char*)img_ws, W, H);
    const int gx = (W + kTile - 1) / kTile, gy = (H + kTile - 1) / kTile;

    prof_mark(ST_BEGIN, stream);
    if (R > 0) {
        if (!binning_ws) return fail(GSR_ERR_INVALID, "binning workspace required");
        BinWS bw = bin_ws_carve((char*)binning_ws, R);
        launch_render_bwd(W, H, gx, gy, iw.ranges, bw.point_list, gw.rec, cam->background, iw.final_T, iw.n_contrib,
                          bw.contrib, dL_dpixels, gr->dL_dmeans2D, gr->dL_dconic, gr->dL_dopacity, gr->dL_dcolors,
                          extra, dL_dpixels_extra, gr->dL_dextra, stream);
        GSR_STAGE("render_backward", cam->debug, stream);
        GSR_MARK(ST_RENDER_BWD, stream, 1);
    }
    PreBwdArgs a;
    a.P = P; a.D = g->sh_degree; a.M = M; a.W = W; a.H = H;
    a.raw = g->raw_params; a.opac = g->opacities; a.dL_dopacity = gr->dL_dopacity; a.dL_dopacity_raw = gr->dL_dopacity_raw;
    a.means = g->means3D; a.scales = g->scales; a.rots = g->rotations; a.shs = g->shs; a.cov3D_pre = g->cov3D_precomp;
    a.view = cam->viewmatrix; a.proj = cam->projmatrix; a.campos = cam->campos;
    a.scale_mod = g->scale_modifier; a.tanfovx = cam->tanfovx; a.tanfovy = cam->tanfovy;
    a.focal_y = H / (2.0f * cam->tanfovy);
    a.focal_x = W / (2.0f * cam->tanfovx);
    a.radii = radii; a.rec = gw.rec;
    a.dL_dmean2D = gr->dL_dmeans2D; a.dL_dconic = gr->dL_dconic; a.dL_dcolors = gr->dL_dcolors;
    a.dL_dmeans3D = gr->dL_dmeans3D; a.dL_dcov3D = gr->dL_dcov3D; a.dL_dsh = gr->dL_dsh;
    a.dL_dscales = gr->dL_dscales; a.dL_drots = gr->dL_drotations;
    a.dL_dview = gr->dL_dviewmatrix; a.dL_dproj = gr->dL_dprojmatrix; a.dL_dcampos = gr->dL_dcampos;
    a.cam_partials = gw.cam_partials; a.cam_done = gw.counters + 63; a.chunk_ticket = gw.counters + kCntChunkBwd;
    a.accumulate = gr->accumulate;
    a.dL_dextra_gen = g->extra_mode == 1 ? gr->dL_dextra : nullptr;
    launch_preprocess_bwd(a, stream);
    GSR_STAGE("preprocess_backward", cam->debug, stream);
    GSR_MARK(ST_PREPROCESS_BWD, stream, 1);
    return GSR_OK;
}

int gsr_blend_stats(gsr_stream_t stream_, int32_t P, int32_t W, int32_t H, int64_t R, const void* geom_ws,
                    const void* binning_ws, const void* img_ws, uint64_t* out)
{
    if (P <= 0 || W <= 0 || H <= 0 || R < 0) return fail(GSR_ERR_INVALID, "bad size");
    if (!geom_ws || !img_ws || !out || (R > 0 && !binning_ws)) return fail(GSR_ERR_INVALID, "null argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    GeomWS gw = geom_ws_carve((char*)geom_ws, P, W, H);
    ImgWS iw = img_ws_carve((char*)img_ws, W, H);
    BinWS bw = bin_ws_carve((char*)binning_ws, R);
    GSR_CUDA(cudaMemsetAsync(out, 0, 4 * sizeof(uint64_t), stream));
    if (R == 0) return GSR_OK;
    launch_blend_stats(W, H, (W + kTile - 1) / kTile, (H + kTile - 1) / kTile, iw.ranges, bw.point_list, gw.rec,
                       iw.n_contrib, bw.contrib, reinterpret_cast<unsigned long long*>(out), stream);
    GSR_STAGE("blend_stats", 0, stream);
    return GSR_OK;
}

int gsr_mark_visible(gsr_stream_t stream_, int32_t P, const float* means3D, const float* viewmatrix,
                     const float* projmatrix, uint8_t* present)
{
    if (P < 0) return fail(GSR_ERR_INVALID, "P < 0");
    if (P == 0) return GSR_OK;
    if (!means3D || !viewmatrix || !present) return fail(GSR_ERR_INVALID, "null argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    launch_mark_visible(P, means3D, viewmatrix, projmatrix, present, stream);
    GSR_STAGE("mark_visible", 0, stream);
    return GSR_OK;
}

}  // extern "C"
