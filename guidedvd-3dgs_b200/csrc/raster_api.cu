// raster_api.cu -- C-ABI entry points (include/gvd_raster.h): buffer carving and stage sequencing.
// Mirrors the control flow of DGR/cuda_rasterizer/rasterizer_impl.cu:197-447 on an explicit stream;
// no torch types, no allocation, no library kernels (the reference's CUB scan and radix sort are
// replaced by the compaction / depth-sort / counting-sort kernels of raster_forward.cu).
#ifndef GVD_HOST_EMU
#include <nvtx3/nvToolsExt.h>
#else  // host emulation (tests/cuda_emu): no profiler tools there
static inline void nvtxRangePushA(const char*) {}
static inline void nvtxRangePop() {}
#endif
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <string>

#include "../../include/gvd_raster.h"
#include "raster_common.cuh"

thread_local std::string g_raster_err;  // shared with grad_exchange.cu; hidden visibility keeps it out of the ABI
#define g_err g_raster_err

namespace {

int fail(const char* what, cudaError_t e) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return 1;
}
int fail_msg(const std::string& s) {
    g_err = s;
    return 2;
}

#define GVD_CHECK(expr, what)                         \
    do {                                              \
        cudaError_t _e = (expr);                      \
        if (_e != cudaSuccess) return fail(what, _e); \
    } while (0)

// after a launch: always catch launch-config errors; in debug mode also sync (auxiliary.h:166-173)
#define GVD_STAGE(what)                                                   \
    do {                                                                  \
        cudaError_t _e = cudaGetLastError();                              \
        if (_e != cudaSuccess) return fail(what, _e);                     \
        if (debug) {                                                      \
            _e = cudaStreamSynchronize(stream);                           \
            if (_e != cudaSuccess) return fail(what " (debug sync)", _e); \
        }                                                                 \
    } while (0)

// ---- optional per-stage timing -------------------------------------------------------------
struct StageProfiler {
    bool on = false;
    static const int kMax = 4096;
    cudaEvent_t beg[kMax], end[kMax];
    int stage[kMax];
    int used = 0, created = 0;
    GvdRasterStageTimes acc{};
    void flush() {
        for (int i = 0; i < used; ++i) {
            if (cudaEventSynchronize(end[i]) != cudaSuccess) continue;
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, beg[i], end[i]) == cudaSuccess) {
                acc.ms[stage[i]] += ms;
                acc.calls[stage[i]] += 1;
            }
        }
        used = 0;
    }
    int begin(int st, cudaStream_t s) {
        if (!on) return -1;
        if (used == kMax) flush();
        if (used == created) {
            cudaEventCreate(&beg[created]);
            cudaEventCreate(&end[created]);
            ++created;
        }
        stage[used] = st;
        cudaEventRecord(beg[used], s);
        return used++;
    }
    void finish(int h, cudaStream_t s) {
        if (h >= 0) cudaEventRecord(end[h], s);
    }
};
StageProfiler g_prof;

// NVTX ranges around the same stages (GVD_NVTX=1; header-only nvtx3, resolved by the tools at run time): an nsys / ncu
// timeline of train_*.py then shows "gvd:preprocess", "gvd:render_bwd", ... instead of anonymous kernel runs.
static const char* const kStageNames[] = {"gvd:preprocess", "gvd:bin_count", "gvd:bin_fill", "gvd:depth_sort", "gvd:export_keys",
                                          "gvd:render_fwd", "gvd:render_bwd", "gvd:gaussian_bwd"};
static bool nvtx_on() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("GVD_NVTX");
        on = (e && e[0] == '1') ? 1 : 0;
    }
    return on == 1;
}

struct StageScope {
    int h;
    cudaStream_t s;
    bool nv;
    StageScope(int st, cudaStream_t s_) : h(g_prof.begin(st, s_)), s(s_), nv(nvtx_on()) {
        if (nv) nvtxRangePushA(kStageNames[st < 8 ? st : 0]);
    }
    ~StageScope() {
        if (nv) nvtxRangePop();
        g_prof.finish(h, s);
    }
};

template <typename T>
void obtain(char*& chunk, T*& ptr, size_t count, size_t alignment = 128) {
    size_t offset = (reinterpret_cast<uintptr_t>(chunk) + alignment - 1) & ~(alignment - 1);
    ptr = reinterpret_cast<T*>(offset);
    chunk = reinterpret_cast<char*>(ptr + count);
}

// The carve_* functions define the scratch layouts; calling them on a null base yields sizes
// (same trick as CudaRasterizer::required<T>, rasterizer_impl.h:63-69).
RasterGeomPtrs carve_geom(char*& chunk, size_t P) {
    RasterGeomPtrs g;
    obtain(chunk, g.splat, P);
    obtain(chunk, g.clamped, P);
    obtain(chunk, g.tiles_touched, P);
    obtain(chunk, g.vis_id, P);
    obtain(chunk, g.counts, 8);
    return g;
}

RasterSortPtrs carve_sort(char*& chunk, size_t P) {
    RasterSortPtrs so;
    so.nb = (P + GVD_PRE_BLOCK - 1) / GVD_PRE_BLOCK;
    so.nt = (P + GVD_SORT_TILE - 1) / GVD_SORT_TILE;
    so.ns = (so.nt + GVD_SORT_SUPER - 1) / GVD_SORT_SUPER;
    obtain(chunk, so.depth_key, P);
    obtain(chunk, so.blk_vis, so.nb);
    obtain(chunk, so.blk_tiles, so.nb);
    obtain(chunk, so.key[0], P);
    obtain(chunk, so.key[1], P);
    obtain(chunk, so.val[0], P);
    obtain(chunk, so.val[1], P);
    so.zeroed_words = GVD_GHIST_COPIES * 4 * 256 + 4 * so.nt * 256 + 4 * so.ns * 256 + 4;
    obtain(chunk, so.zeroed, so.zeroed_words);
    so.ghist = so.zeroed;
    so.agg = so.ghist + GVD_GHIST_COPIES * 4 * 256;
    so.incl = so.agg + 4 * so.nt * 256;
    so.ticket = so.incl + 4 * so.ns * 256;
    return so;
}

RasterHistPtrs carve_hist(char*& chunk, size_t rows, size_t T) {
    RasterHistPtrs h;
    h.rows = rows;
    obtain(chunk, h.hist, rows * T);
    obtain(chunk, h.tile_total, T);
    return h;
}

RasterBinPtrs carve_binning(char*& chunk, size_t R, bool with_keys) {
    RasterBinPtrs b;
    obtain(chunk, b.point_list, R + 4);  // +4: the TMA id copies round their length up to 16 bytes
    b.keys = nullptr;
    if (with_keys) obtain(chunk, b.keys, R);
    return b;
}

RasterImgPtrs carve_img(char*& chunk, size_t tiles, size_t pixels) {
    RasterImgPtrs im;
    obtain(chunk, im.ranges, tiles);
    obtain(chunk, im.n_contrib, pixels);
    return im;
}

inline dim3 tile_grid(int width, int height) {
    return dim3((width + GVD_TILE_X - 1) / GVD_TILE_X, (height + GVD_TILE_Y - 1) / GVD_TILE_Y, 1);
}

}  // namespace

extern "C" {

int gvd_raster_abi_version(void) { return GVD_RASTER_ABI_VERSION; }

int gvd_raster_profile_enable(int on) {
    g_prof.flush();
    g_prof.acc = GvdRasterStageTimes{};
    g_prof.on = on != 0;
    return 0;
}
int gvd_raster_profile_read(GvdRasterStageTimes* out) {
    if (!out) return fail_msg("gvd_raster_profile_read: null out");
    g_prof.flush();
    *out = g_prof.acc;
    return 0;
}
const char* gvd_last_error(void) { return g_err.c_str(); }

size_t gvd_raster_geom_bytes(int P, int width, int height) {
    (void)width;
    (void)height;
    char* p = nullptr;
    carve_geom(p, (size_t)P);
    return (size_t)p + 128;
}
size_t gvd_raster_sort_bytes(int P) {
    char* p = nullptr;
    carve_sort(p, (size_t)P);
    return (size_t)p + 128;
}
size_t gvd_raster_hist_bytes(int num_visible, int width, int height) {
    char* p = nullptr;
    dim3 g = tile_grid(width, height);
    carve_hist(p, ((size_t)num_visible + GVD_BIN_CHUNK - 1) / GVD_BIN_CHUNK, (size_t)g.x * g.y);
    return (size_t)p + 128;
}
size_t gvd_raster_binning_bytes(int R, int export_keys) {
    char* p = nullptr;
    carve_binning(p, (size_t)R, export_keys != 0);
    return (size_t)p + 128;
}
size_t gvd_raster_img_bytes(int width, int height) {
    char* p = nullptr;
    dim3 g = tile_grid(width, height);
    carve_img(p, (size_t)g.x * g.y, (size_t)width * height);
    return (size_t)p + 128;
}
size_t gvd_raster_backward_scratch_bytes(int P) { return (size_t)P * GVD_ACC_STRIDE * sizeof(float) + 128; }

int gvd_raster_layout(int P, int R, int width, int height, GvdRasterLayout* out) {
    if (!out) return fail_msg("gvd_raster_layout: null out");
    dim3 tg = tile_grid(width, height);
    char* p = nullptr;
    RasterGeomPtrs g = carve_geom(p, (size_t)P);
    out->geom_splat = (size_t)g.splat;
    out->geom_clamped = (size_t)g.clamped;
    out->geom_tiles_touched = (size_t)g.tiles_touched;
    out->geom_visible_ids = (size_t)g.vis_id;
    out->geom_counts = (size_t)g.counts;
    p = nullptr;
    RasterBinPtrs b = carve_binning(p, (size_t)R, true);
    out->bin_point_list = (size_t)b.point_list;
    out->bin_point_list_keys = (size_t)b.keys;
    p = nullptr;
    RasterImgPtrs im = carve_img(p, (size_t)tg.x * tg.y, (size_t)width * height);
    out->img_ranges = (size_t)im.ranges;
    out->img_n_contrib = (size_t)im.n_contrib;
    return 0;
}

// Event the exact path waits on for R when the caller supplied none: one per host thread and device, created lazily.
static cudaEvent_t internal_r_event() {
    thread_local cudaEvent_t ev[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (!ev[dev] && cudaEventCreateWithFlags(&ev[dev], cudaEventDisableTiming) != cudaSuccess) ev[dev] = nullptr;
    return ev[dev];
}

int gvd_raster_forward(GvdRasterForwardArgs* a, gvd_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (!a) return fail_msg("gvd_raster_forward: null args");
    const bool debug = a->debug != 0;
    a->num_rendered = 0;
    a->num_visible = 0;
    if (a->P <= 0) return 0;
    if (a->width <= 0 || a->height <= 0) return fail_msg("gvd_raster_forward: bad image size");
    if ((a->shs == nullptr) == (a->colors_precomp == nullptr))
        return fail_msg("Please provide excatly one of either SHs or precomputed colors!");
    if (((a->scales == nullptr || a->rotations == nullptr) && a->cov3D_precomp == nullptr) ||
        ((a->scales != nullptr || a->rotations != nullptr) && a->cov3D_precomp != nullptr))
        return fail_msg("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!");
    const bool spec = a->spec_binning_buffer != nullptr && !debug;
    if ((!a->geom_alloc && !a->geom_buffer) || (!a->img_alloc && !a->img_buffer) || (!a->temp_alloc && !a->sort_buffer))
        return fail_msg("gvd_raster_forward: null allocator");
    if (!spec && (!a->binning_alloc || !a->temp_alloc)) return fail_msg("gvd_raster_forward: null binning / temp allocator");
    if (spec && (!a->spec_hist_buffer || !a->num_rendered_pinned))
        return fail_msg("gvd_raster_forward: the speculative path needs spec_hist_buffer and num_rendered_pinned");
    if (a->shs && (a->D < 0 || a->D > 3 || (a->D + 1) * (a->D + 1) > a->M))
        return fail_msg("gvd_raster_forward: SH degree / coefficient count mismatch");
    if (a->raw_params && (!a->shs || !a->scales || !a->rotations || a->cov3D_precomp || a->colors_precomp || (a->M > 1 && !a->shs_rest)))
        return fail_msg("gvd_raster_forward: raw_params needs shs (dc) + shs_rest, scales and rotations, and no precomputed covariance / colours");

    const int P = a->P;
    const float focal_y = a->height / (2.0f * a->tan_fovy);
    const float focal_x = a->width / (2.0f * a->tan_fovx);
    const dim3 grid = tile_grid(a->width, a->height);
    const size_t tiles = (size_t)grid.x * grid.y;
    if (grid.x > 0xffff || grid.y > 0xffff || tiles > GVD_MAX_TILES)
        return fail_msg("gvd_raster_forward: image too large (more than 49152 tiles of 16x16)");

    const size_t geom_need = gvd_raster_geom_bytes(P, a->width, a->height);
    if (a->geom_buffer && a->geom_bytes < geom_need) return fail_msg("gvd_raster_forward: geom_buffer too small");
    char* gp = a->geom_buffer ? (char*)a->geom_buffer : (char*)a->geom_alloc(a->alloc_user, geom_need);
    if (!gp) return fail_msg("gvd_raster_forward: geometry allocator returned null");
    RasterGeomPtrs g = carve_geom(gp, (size_t)P);
    const size_t img_need = gvd_raster_img_bytes(a->width, a->height);
    if (a->img_buffer && a->img_bytes < img_need) return fail_msg("gvd_raster_forward: img_buffer too small");
    char* ip = a->img_buffer ? (char*)a->img_buffer : (char*)a->img_alloc(a->alloc_user, img_need);
    if (!ip) return fail_msg("gvd_raster_forward: image allocator returned null");
    RasterImgPtrs im = carve_img(ip, tiles, (size_t)a->width * a->height);
    const size_t sort_need = gvd_raster_sort_bytes(P);
    if (a->sort_buffer && a->sort_bytes < sort_need) return fail_msg("gvd_raster_forward: sort_buffer too small");
    char* sp = a->sort_buffer ? (char*)a->sort_buffer : (char*)a->temp_alloc(a->alloc_user, sort_need);
    if (!sp) return fail_msg("gvd_raster_forward: temp allocator returned null (sort scratch)");
    RasterSortPtrs so = carve_sort(sp, (size_t)P);

    {
        StageScope t(GVD_STAGE_PREPROCESS, stream);
        gvd_launch_preprocess(*a, g, so, focal_x, focal_y, grid, stream);
    }
    GVD_STAGE("preprocess");
    // R and V leave the GPU ~40 us into the frame, right behind the compaction kernel -- long before the instance list
    // is needed -- so the exact path below can size its buffers without ever letting the GPU run dry: the depth sort
    // is already queued behind the event the host waits on.
    cudaEvent_t r_event = a->r_ready_event ? reinterpret_cast<cudaEvent_t>(a->r_ready_event) : nullptr;
    {
        StageScope t(GVD_STAGE_SORT, stream);
        gvd_launch_compact(P, g, so, a->num_rendered_pinned, stream);
        if (a->num_rendered_pinned && !debug) {
            if (!r_event && !spec) r_event = internal_r_event();
            if (r_event) GVD_CHECK(cudaEventRecord(r_event, stream), "record R event");
        }
        gvd_launch_depth_sort(P, g, so, stream);
    }
    GVD_STAGE("compact + depth sort");

    RasterBinPtrs b;
    RasterHistPtrs h;
    uint32_t capacity = 0xffffffffu;  // instance slots available in the binning buffer
    bool have_instances = true;
    if (spec) {
        // speculative path: no host round trip; the caller validates R and V afterwards (both were in pinned memory
        // when r_ready_event fired)
        a->num_rendered = -1;
        a->num_visible = -1;
        const size_t per = sizeof(uint32_t) + (a->export_keys ? sizeof(uint64_t) : 0);
        const size_t fixed = gvd_raster_binning_bytes(0, a->export_keys);
        const size_t room = a->spec_binning_bytes > fixed + 256 ? a->spec_binning_bytes - fixed - 256 : 0;
        capacity = (uint32_t)std::min<size_t>(room / per, 0x7fffffffu);
        char* bp = (char*)a->spec_binning_buffer;
        b = carve_binning(bp, (size_t)capacity, a->export_keys != 0);
        // largest number of chunk rows whose layout fits the caller's hist buffer
        const size_t hfixed = gvd_raster_hist_bytes(0, a->width, a->height);
        const size_t hroom = a->spec_hist_bytes > hfixed + 256 ? a->spec_hist_bytes - hfixed - 256 : 0;
        const size_t rows = std::min<size_t>(hroom / (tiles * sizeof(uint32_t)), ((size_t)P + GVD_BIN_CHUNK - 1) / GVD_BIN_CHUNK);
        char* hp = (char*)a->spec_hist_buffer;
        h = carve_hist(hp, rows, tiles);
    } else {
        // Sizes of the instance list and of the chunk histogram. Like the reference (rasterizer_impl.cu:281-282) the
        // buffers are caller-owned and sized from R; unlike it, the wait is for the first two kernels only.
        int rv[2] = {0, 0};
        if (a->num_rendered_pinned && r_event && !debug) {
            GVD_CHECK(cudaEventSynchronize(r_event), "wait for R");
            rv[0] = reinterpret_cast<volatile int*>(a->num_rendered_pinned)[0];
            rv[1] = reinterpret_cast<volatile int*>(a->num_rendered_pinned)[1];
        } else {
            uint32_t c[2] = {0, 0};
            GVD_CHECK(cudaMemcpyAsync(c, g.counts, sizeof(c), cudaMemcpyDeviceToHost, stream), "copy num_rendered");
            GVD_CHECK(cudaStreamSynchronize(stream), "sync num_rendered");
            rv[0] = (int)c[1];
            rv[1] = (int)c[0];
        }
        a->num_rendered = rv[0];
        a->num_visible = rv[1];
        char* bp = (char*)a->binning_alloc(a->alloc_user, gvd_raster_binning_bytes(rv[0], a->export_keys));
        if (!bp) return fail_msg("gvd_raster_forward: binning allocator returned null");
        b = carve_binning(bp, (size_t)rv[0], a->export_keys != 0);
        char* hp = (char*)a->temp_alloc(a->alloc_user, gvd_raster_hist_bytes(rv[1], a->width, a->height));
        if (!hp) return fail_msg("gvd_raster_forward: temp allocator returned null (chunk histogram)");
        h = carve_hist(hp, ((size_t)rv[1] + GVD_BIN_CHUNK - 1) / GVD_BIN_CHUNK, tiles);
        have_instances = rv[0] > 0;
    }

    {
        // level 2a: per-chunk tile histograms -> per-tile prefixes -> ranges
        StageScope t(GVD_STAGE_SCAN, stream);
        GVD_CHECK(gvd_launch_bin_count(g, so, h, im, grid, stream), "bin_count");
    }
    GVD_STAGE("bin_count");
    if (have_instances) {
        {
            // level 2b: stable scatter of the Gaussian ids into the tile lists
            StageScope t(GVD_STAGE_EMIT, stream);
            GVD_CHECK(gvd_launch_bin_fill(g, so, h, b, im, grid, capacity, stream), "bin_fill");
        }
        GVD_STAGE("bin_fill");
        if (a->export_keys) {
            StageScope t(GVD_STAGE_PACK, stream);
            gvd_launch_export_keys(capacity, so, b, im, grid, stream);
        }
        GVD_STAGE("export_keys");
    }
    {
        StageScope t(GVD_STAGE_RENDER_FWD, stream);
        gvd_launch_render_forward(*a, g, b, im, grid, capacity, stream);
    }
    GVD_STAGE("render_forward");
    return 0;
}

int gvd_raster_backward(const GvdRasterBackwardArgs* a, gvd_stream_t stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (!a) return fail_msg("gvd_raster_backward: null args");
    const bool debug = a->debug != 0;
    if (a->P <= 0) return 0;
    if (!a->geom_buffer || !a->img_buffer || !a->scratch) return fail_msg("gvd_raster_backward: null scratch buffer");
    if (a->R > 0 && !a->binning_buffer) return fail_msg("gvd_raster_backward: null binning buffer");
    if (!a->dL_dmeans2D || !a->dL_dmeans3D || !a->dL_dopacity) return fail_msg("gvd_raster_backward: null output");
    if (a->shs && !a->dL_dsh) return fail_msg("gvd_raster_backward: shs given but dL_dsh is null");
    if (a->scales && (!a->dL_dscales || !a->dL_drotations || !a->rotations))
        return fail_msg("gvd_raster_backward: scales given but dL_dscales/dL_drotations is null");
    if (!a->scales && !a->cov3D_precomp) return fail_msg("gvd_raster_backward: neither scales nor cov3D_precomp");
    if (a->raw_params && (!a->shs || !a->scales || !a->opacities || a->cov3D_precomp || a->colors_precomp ||
                          (a->M > 1 && (!a->shs_rest || !a->dL_dsh_rest))))
        return fail_msg("gvd_raster_backward: raw_params needs shs (dc) + shs_rest + dL_dsh_rest, scales, rotations and the raw opacities");

    const int P = a->P;
    const float focal_y = a->height / (2.0f * a->tan_fovy);
    const float focal_x = a->width / (2.0f * a->tan_fovx);
    const dim3 grid = tile_grid(a->width, a->height);
    const size_t tiles = (size_t)grid.x * grid.y;

    char* gp = (char*)a->geom_buffer;
    RasterGeomPtrs g = carve_geom(gp, (size_t)P);
    char* bp = (char*)a->binning_buffer;
    RasterBinPtrs b = carve_binning(bp, (size_t)a->R, false);
    char* ip = (char*)a->img_buffer;
    RasterImgPtrs im = carve_img(ip, tiles, (size_t)a->width * a->height);

    float* acc = reinterpret_cast<float*>(((uintptr_t)a->scratch + 127) & ~(uintptr_t)127);
    gvd_launch_zero_fill(acc, (size_t)P * GVD_ACC_STRIDE, stream);
    GVD_STAGE("zero accumulator");

    // The dense gradient outputs are zero everywhere except on the V visible Gaussians (the reference zero-fills them
    // with torch::zeros, rasterize_points.cu:158-167). When the caller says they all lie in one region, the render CTAs
    // clear it on the side -- that kernel is issue-bound and leaves the memory system idle; otherwise each output is
    // cleared by its own launch.
    float4* zero = nullptr;
    size_t zero_n4 = 0;
    if (a->zero_region && a->zero_region_bytes >= 16 && ((uintptr_t)a->zero_region & 15) == 0 && a->R > 0) {
        zero = reinterpret_cast<float4*>(a->zero_region);
        zero_n4 = a->zero_region_bytes / 16;
        if (a->zero_region_bytes & 15)
            gvd_launch_zero_bytes((char*)a->zero_region + zero_n4 * 16, a->zero_region_bytes & 15, stream);
    } else if (a->zero_region) {
        gvd_launch_zero_bytes(a->zero_region, a->zero_region_bytes, stream);
    } else {
        const size_t M3 = (size_t)a->M * 3;
        gvd_launch_zero_bytes(a->dL_dmeans2D, (size_t)P * 3 * 4, stream);
        gvd_launch_zero_bytes(a->dL_dmeans3D, (size_t)P * 3 * 4, stream);
        gvd_launch_zero_bytes(a->dL_dopacity, (size_t)P * 4, stream);
        if (a->dL_dcolors) gvd_launch_zero_bytes(a->dL_dcolors, (size_t)P * 3 * 4, stream);
        if (a->dL_dcov3D) gvd_launch_zero_bytes(a->dL_dcov3D, (size_t)P * 6 * 4, stream);
        if (a->dL_dsh && a->raw_params) {  // two tensors: d/d_features_dc [P,1,3] and d/d_features_rest [P,M-1,3]
            gvd_launch_zero_bytes(a->dL_dsh, (size_t)P * 3 * 4, stream);
            if (a->dL_dsh_rest) gvd_launch_zero_bytes(a->dL_dsh_rest, (size_t)P * (M3 - 3) * 4, stream);
        } else if (a->dL_dsh) {
            gvd_launch_zero_bytes(a->dL_dsh, (size_t)P * M3 * 4, stream);
        }
        if (a->dL_dscales) gvd_launch_zero_bytes(a->dL_dscales, (size_t)P * 3 * 4, stream);
        if (a->dL_drotations) gvd_launch_zero_bytes(a->dL_drotations, (size_t)P * 4 * 4, stream);
    }
    GVD_STAGE("zero gradients");

    if (a->R > 0) {
        {
            StageScope t(GVD_STAGE_RENDER_BWD, stream);
            gvd_launch_render_backward(*a, g, b, im, acc, zero, zero_n4, grid, stream);
        }
        GVD_STAGE("render_backward");
    }
    {
        StageScope t(GVD_STAGE_GAUSSIAN_BWD, stream);
        gvd_launch_gaussian_backward(*a, g, acc, focal_x, focal_y, a->num_visible, stream);
    }
    GVD_STAGE("gaussian_backward");
    return 0;
}

int gvd_raster_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                            uint8_t* present, gvd_stream_t stream_) {
    (void)projmatrix;  // the reference passes it but only the view-space z test is live (auxiliary.h:154)
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (P <= 0) return 0;
    if (!means3D || !viewmatrix || !present) return fail_msg("gvd_raster_mark_visible: null pointer");
    gvd_launch_mark_visible(P, means3D, viewmatrix, present, stream);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail("mark_visible", e);
    return 0;
}

}  // extern "C"
