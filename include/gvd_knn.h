/*
 * gvd_knn.h -- C ABI of the B200-native 3-nearest-neighbour op.
 *
 * Drop-in boundary for the reference's native extension submodules/simple-knn:
 *   distCUDA2(points) (simple-knn/spatial.cu:15-27) -> SimpleKNN::knn (simple_knn.cu:192-228).
 * For every point: the mean of the squared distances to its 3 nearest OTHER points and the
 * indices of those points (nearest first).  Exact (not approximate) like the reference.
 * All pointers are CUDA device pointers; scratch is caller-owned; no host synchronisation
 * (the reference copies the bounding box to the host twice and cudaMallocs per call).
 */
#ifndef GVD_KNN_H_
#define GVD_KNN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define GVD_KNN_API __attribute__((visibility("default")))
#else
#define GVD_KNN_API
#endif

typedef struct CUstream_st* gvd_knn_stream_t; /* == cudaStream_t */

/* bytes of scratch needed by gvd_knn3 for P points */
GVD_KNN_API size_t gvd_knn3_tmp_bytes(int P);

/* xyz: float[P,3]; mean_d2: float[P]; idx3: int32[P,3] (original indices, nearest first).
 * With fewer than 4 points the missing neighbours have distance FLT_MAX and index 0, as in the
 * reference (simple_knn.cu:157-158,187-190).  Returns 0 on success. */
GVD_KNN_API int gvd_knn3(int P, const float* xyz, float* mean_d2, int32_t* idx3, void* tmp, size_t tmp_bytes,
                         gvd_knn_stream_t stream);

GVD_KNN_API const char* gvd_knn_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* GVD_KNN_H_ */
