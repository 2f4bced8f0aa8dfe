// point_project.cu -- z-buffer projection of a coloured point cloud (include/gvd_points.h; reference: scene/pcd2img.py:4-70).
//
// Integer / byte work bound by HBM: 27 bytes read per point (3 doubles + 3 colour bytes), one 64-bit atomic per point
// that lands in the image, 4 bytes written per pixel.  Three stream-ordered launches:
//   1. pp_depth_kernel   -- every point: camera transform, filters, pixel; atomicMin of its ordered depth key.
//   2. pp_winner_kernel  -- every point again (the transform is cheaper than storing 12 bytes per point): a point whose
//                           key equals its pixel's minimum bids its index with atomicMin (ties -> lowest index).
//   3. pp_paint_kernel   -- every pixel: copy the winner's colour, write the mask.
// The per-pixel tables are cleared by a kernel (no copy-engine work in the stream).  Arithmetic is float64 with the
// reference's operation order (row-times-vector sums left to right, no FMA contraction: __dmul_rn / __dadd_rn), so the
// rounded pixel and the depth comparisons agree with numpy's.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/gvd_points.h"

namespace {

thread_local std::string g_points_err;

struct PpCamera {
    double K[9];
    double E[16];
    double near_plane, far_plane;
    int width, height;
};

// a float64 as an unsigned integer with the same order (negative values included)
__device__ __forceinline__ unsigned long long ordered_key(double z) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(z);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__device__ __forceinline__ double dot3_rn(double a0, double b0, double a1, double b1, double a2, double b2) {
    return __dadd_rn(__dadd_rn(__dmul_rn(a0, b0), __dmul_rn(a1, b1)), __dmul_rn(a2, b2));
}

// pixel index of point i, or -1 when it is filtered out; *key = ordered camera depth       (pcd2img.py:26-53)
__device__ __forceinline__ long long pp_project(const double* __restrict__ pts, long long i, const PpCamera& c,
                                                unsigned long long* key) {
    const double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    // extrinsics @ [x, y, z, 1]: four products summed left to right
    const double cx = __dadd_rn(dot3_rn(c.E[0], x, c.E[1], y, c.E[2], z), c.E[3]);
    const double cy = __dadd_rn(dot3_rn(c.E[4], x, c.E[5], y, c.E[6], z), c.E[7]);
    const double cz = __dadd_rn(dot3_rn(c.E[8], x, c.E[9], y, c.E[10], z), c.E[11]);
    if (!(cz > c.near_plane && cz < c.far_plane)) return -1;
    const double ix = dot3_rn(c.K[0], cx, c.K[1], cy, c.K[2], cz);
    const double iy = dot3_rn(c.K[3], cx, c.K[4], cy, c.K[5], cz);
    const double iw = dot3_rn(c.K[6], cx, c.K[7], cy, c.K[8], cz);
    const double u = rint(__ddiv_rn(ix, iw)), v = rint(__ddiv_rn(iy, iw));   // np.round: half to even
    if (!(u >= 0.0 && u < (double)c.width && v >= 0.0 && v < (double)c.height)) return -1;   // also rejects NaN / inf
    *key = ordered_key(cz);
    return (long long)v * c.width + (long long)u;
}

__global__ void __launch_bounds__(256) pp_clear_kernel(unsigned long long* __restrict__ depth, int* __restrict__ winner,
                                                       long long pixels) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < pixels; p += (long long)gridDim.x * blockDim.x) {
        depth[p] = ~0ull;
        winner[p] = 0x7fffffff;
    }
}

__global__ void __launch_bounds__(256) pp_depth_kernel(const double* __restrict__ pts, long long n, PpCamera c,
                                                       unsigned long long* __restrict__ depth) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        unsigned long long key;
        const long long pix = pp_project(pts, i, c, &key);
        if (pix >= 0) atomicMin(&depth[pix], key);
    }
}

__global__ void __launch_bounds__(256) pp_winner_kernel(const double* __restrict__ pts, long long n, PpCamera c,
                                                        const unsigned long long* __restrict__ depth, int* __restrict__ winner) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        unsigned long long key;
        const long long pix = pp_project(pts, i, c, &key);
        if (pix >= 0 && depth[pix] == key) atomicMin(&winner[pix], (int)i);
    }
}

__global__ void __launch_bounds__(256) pp_paint_kernel(const uint8_t* __restrict__ colors, const int* __restrict__ winner,
                                                       long long pixels, uint8_t* __restrict__ image, uint8_t* __restrict__ mask) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < pixels; p += (long long)gridDim.x * blockDim.x) {
        const int w = winner[p];
        uint8_t r = 0, g = 0, b = 0, m = 0;
        if (w != 0x7fffffff) {
            r = colors[3 * (long long)w];
            g = colors[3 * (long long)w + 1];
            b = colors[3 * (long long)w + 2];
            m = 1;
        }
        image[3 * p] = r;
        image[3 * p + 1] = g;
        image[3 * p + 2] = b;
        mask[p] = m;
    }
}

int pp_grid(long long n) {
    long long g = (n + 255) / 256;
    const long long cap = 148 * 8;  // 8 resident 256-thread CTAs per SM, grid-stride beyond that
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" {

const char* gvd_points_last_error(void) { return g_points_err.c_str(); }

size_t gvd_point_project_scratch_bytes(int width, int height) {
    if (width <= 0 || height <= 0) return 0;
    return (size_t)width * height * (sizeof(unsigned long long) + sizeof(int));
}

int gvd_point_project(const double* points, const uint8_t* colors, long long n, const double* intrinsics,
                      const double* extrinsics, int width, int height, double near_plane, double far_plane, uint8_t* image,
                      uint8_t* mask, void* scratch, size_t scratch_bytes, gvd_points_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (width <= 0 || height <= 0 || n < 0) { g_points_err = "gvd_point_project: width, height must be positive, n >= 0"; return 2; }
    if (n > 0x7ffffffell) { g_points_err = "gvd_point_project: at most 2^31 - 2 points"; return 2; }
    if (!image || !mask || !scratch || !intrinsics || !extrinsics || (n > 0 && (!points || !colors))) {
        g_points_err = "gvd_point_project: null pointer";
        return 2;
    }
    if (scratch_bytes < gvd_point_project_scratch_bytes(width, height) || (reinterpret_cast<uintptr_t>(scratch) & 7)) {
        g_points_err = "gvd_point_project: scratch too small or not 8-byte aligned";
        return 2;
    }
    PpCamera c;  // the camera matrices are HOST memory (25 doubles, as the reference's numpy arrays): passed by value
    for (int i = 0; i < 9; ++i) c.K[i] = intrinsics[i];
    for (int i = 0; i < 16; ++i) c.E[i] = extrinsics[i];
    c.near_plane = near_plane;
    c.far_plane = far_plane;
    c.width = width;
    c.height = height;
    const long long pixels = (long long)width * height;
    unsigned long long* depth = reinterpret_cast<unsigned long long*>(scratch);
    int* winner = reinterpret_cast<int*>(depth + pixels);
    pp_clear_kernel<<<pp_grid(pixels), 256, 0, s>>>(depth, winner, pixels);
    if (n > 0) {
        pp_depth_kernel<<<pp_grid(n), 256, 0, s>>>(points, n, c, depth);
        pp_winner_kernel<<<pp_grid(n), 256, 0, s>>>(points, n, c, depth, winner);
    }
    pp_paint_kernel<<<pp_grid(pixels), 256, 0, s>>>(colors, winner, pixels, image, mask);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_points_err = std::string("gvd_point_project: ") + cudaGetErrorString(e); return 1; }
    return 0;
}

}  // extern "C"
