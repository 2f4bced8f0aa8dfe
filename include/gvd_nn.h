/*
 * gvd_nn.h -- C ABI of the B200-native building blocks of the ViewCrafter U-Net denoiser and DDIM sampler.
 *
 * The reference runs these layers through PyTorch library kernels (cuBLAS / cuDNN / ATen) from
 * third_party/ViewCrafter/lvdm/modules/{attention.py,networks/openaimodel3d.py} and
 * lvdm/models/samplers/ddim.py; there is no reference FFI for them, so the boundary is the operator level:
 * each entry point names the reference module/lines it serves.  All pointers are CUDA device pointers, all
 * buffers caller-owned, every call takes the stream; return 0 on success (message: gvd_nn_last_error()).
 */
#ifndef GVD_NN_H_
#define GVD_NN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define GVD_NN_API __attribute__((visibility("default")))
#else
#define GVD_NN_API
#endif

typedef struct CUstream_st* gvd_nn_stream_t; /* == cudaStream_t */

enum { GVD_ACT_NONE = 0, GVD_ACT_SILU = 1, GVD_ACT_GELU = 2 };

/* Strided-batched bf16 GEMM on tcgen05 tensor cores (fp32 accumulation in TMEM):
 *     C[b,h,m,n] = act( alpha * sum_k A[b,h,m,k] * B[b,h,n,k] + bias[n] ) + residual[b,h,m,n]
 * A and B are K-major (k contiguous); every row/batch stride is in ELEMENTS and must be a multiple of 8.
 * Serves nn.Linear / 1x1 conv / im2col'ed 3x3 and (3,1,1) convs (openaimodel3d.py:155-236,255-279), the
 * q/k/v/out projections, GEGLU/FF linears and the QK^T / PV products of CrossAttention (attention.py:81-144). */
typedef struct GvdGemmArgs {
    int M, N, K;
    int batch_h, batch_b;              /* two batch levels (e.g. heads, batch); use 1 when unused */
    const void* A; long long lda, a_stride_h, a_stride_b;   /* bf16 */
    const void* B; long long ldb, b_stride_h, b_stride_b;   /* bf16 */
    void* C;       long long ldc, c_stride_h, c_stride_b;   /* bf16, or fp32 when out_fp32 */
    const float* bias;                 /* [N] fp32 or NULL */
    const void* residual;              /* same layout/dtype as C, or NULL */
    float alpha;
    int act;                           /* GVD_ACT_* applied before the residual add */
    int out_fp32;
} GvdGemmArgs;
GVD_NN_API int gvd_gemm_bf16(const GvdGemmArgs* args, gvd_nn_stream_t stream);

GVD_NN_API const char* gvd_nn_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* GVD_NN_H_ */
