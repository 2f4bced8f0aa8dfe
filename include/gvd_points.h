/*
 * gvd_points.h -- C ABI of the point-cloud z-buffer projection (BASELINE.json configs[0]).
 *
 * Replaces the numpy routine /root/reference/scene/pcd2img.py::project_point_cloud_to_image (:4-70), which the
 * reference calls from scene/dataset_readers.py:34 and tools/get_replica_dust3r_project_2d.py:7 to splat a coloured
 * point cloud into an image: world -> camera (4x4 extrinsics), near/far filter, 3x3 intrinsics with a divide by the
 * third row, round-half-even to a pixel, nearest point per pixel wins.  All arithmetic is float64 like the reference;
 * exact-depth ties go to the lowest point index.
 *
 * points, colors, image, mask and scratch are CUDA device pointers (caller-owned); the two camera matrices are HOST
 * pointers read during the call; the call is stream-ordered and never synchronises; returns 0 on success
 * (2 = bad argument, 1 = CUDA error; message: gvd_points_last_error()).
 */
#ifndef GVD_POINTS_H_
#define GVD_POINTS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define GVD_POINTS_API __attribute__((visibility("default")))
#else
#define GVD_POINTS_API
#endif

typedef struct CUstream_st* gvd_points_stream_t; /* == cudaStream_t */

/* bytes of scratch for an image of width x height (a 64-bit depth key and a 32-bit winner per pixel) */
GVD_POINTS_API size_t gvd_point_project_scratch_bytes(int width, int height);

/* points  [n, 3] float64 (row-major), colors [n, 3] uint8 (device); intrinsics [3, 3], extrinsics [4, 4] float64
 * (HOST, row-major, as numpy holds them) -> image [height, width, 3] uint8 and mask [height, width] uint8 (both fully
 * written: pixels no point lands on are 0).  n may be 0. */
GVD_POINTS_API int gvd_point_project(const double* points, const uint8_t* colors, long long n, const double* intrinsics,
                                     const double* extrinsics, int width, int height, double near_plane, double far_plane,
                                     uint8_t* image, uint8_t* mask, void* scratch, size_t scratch_bytes,
                                     gvd_points_stream_t stream);

GVD_POINTS_API const char* gvd_points_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* GVD_POINTS_H_ */
