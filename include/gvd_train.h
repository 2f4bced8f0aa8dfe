/*
 * gvd_train.h -- C ABI of the per-iteration 3DGS training-step pieces that sit around the rasterizer
 * (SURVEY.md section 8f row f3).  The reference runs them as ~150 small PyTorch launches per iteration:
 *   utils/loss_utils.py:18-28,46-82   l1_loss / l1_loss_mask and ssim (five 11x11 depthwise convolutions + ~15 elementwise
 *                                     kernels forward, as many again in autograd's backward), combined in
 *                                     train_baseline.py:82-83 as (1 - lambda) * L1 + lambda * (1 - SSIM);
 *   scene/gaussian_model.py:524-527   add_densification_stats (boolean-mask indexing: a nonzero() + host sync each)
 *   train_baseline.py:109             max_radii2D[visibility] = max(...)
 *   scene/gaussian_model.py:179-190   torch.optim.Adam(eps=1e-15) over the six parameter groups.
 * All pointers are CUDA device pointers (fp32 unless stated), buffers caller-owned, every call stream-ordered without
 * synchronisation; return 0 on success, 2 = bad argument, 1 = CUDA error (message: gvd_train_last_error()).
 */
#ifndef GVD_TRAIN_H_
#define GVD_TRAIN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define GVD_TRAIN_API __attribute__((visibility("default")))
#else
#define GVD_TRAIN_API
#endif

typedef struct CUstream_st* gvd_train_stream_t; /* == cudaStream_t */

/* Fused photometric loss, forward.  img, gt: [C, H, W] (the rasterizer's output layout).  Writes
 *   out[0] = mean |img - gt|   (l1_loss, loss_utils.py:18-20)
 *   out[1] = mean SSIM map     (ssim with the 11x11 Gaussian window, sigma 1.5, zero padding; loss_utils.py:46-82)
 * and, when dmaps is not NULL, the three per-pixel partial derivatives of the SSIM map with respect to the window
 * means of img, img^2 and img*gt (dmaps [3, C, H, W]) that the backward convolves.  scratch: *_scratch_bytes(). */
GVD_TRAIN_API size_t gvd_photometric_loss_scratch_bytes(int C, int H, int W);
GVD_TRAIN_API int gvd_photometric_loss_forward(const float* img, const float* gt, int C, int H, int W, float* out,
                                               float* dmaps, void* scratch, size_t scratch_bytes,
                                               gvd_train_stream_t stream);

/* Backward: dimg [C, H, W] = coef[0] * d(mean|img-gt|)/dimg + coef[1] * d(mean SSIM)/dimg, coef = two DEVICE floats
 * (the upstream gradients of the two means, e.g. {(1 - lambda) g, -lambda g}), so no host value is needed. */
GVD_TRAIN_API int gvd_photometric_loss_backward(const float* img, const float* gt, const float* dmaps, const float* coef,
                                                int C, int H, int W, float* dimg, gvd_train_stream_t stream);

/* Densification bookkeeping of one iteration, fused and free of host round trips (gaussian_model.py:524-527 and
 * train_baseline.py:109): for every Gaussian with radii[i] > 0 (the visibility filter):
 *   xyz_gradient_accum[i] += ||means2D_grad[i, 0:2]||;  denom[i] += 1;  max_radii2D[i] = max(max_radii2D[i], radii[i]).
 * means2D_grad [P, 3] fp32, radii [P] int32, the three accumulators [P] fp32. */
GVD_TRAIN_API int gvd_densification_stats(const float* means2D_grad, const int* radii, long long P, float* xyz_gradient_accum,
                                          float* denom, float* max_radii2D, gvd_train_stream_t stream);

/* One Adam update of a flat parameter tensor with torch.optim.Adam's arithmetic (no weight decay, no amsgrad;
 * torch/optim/adam.py `_single_tensor_adam`): m = lerp(m, g, 1 - b1); v = b2 v + (1 - b2) g^2;
 * p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps).  `step` = t >= 1 (host integer);
 * the hyper-parameters are doubles, as torch holds them (1 - beta2 must be formed in double: 1 - 0.999f is off by 1e-5). */
GVD_TRAIN_API int gvd_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, double lr,
                                double beta1, double beta2, double eps, int step, gvd_train_stream_t stream);

/* Binary erosion (dilate = 0) / dilation (dilate = 1) of N float masks [N, H, W] (non-zero = set) with a rectangular
 * structuring element covering the offsets [lo_y, hi_y] x [lo_x, hi_x] around each pixel, outside the image = unset:
 * scipy.ndimage.binary_erosion / binary_dilation(structure=np.ones((k, k))) as utils/viewcrafter_wrapper.py:602-647 applies
 * them on the CPU, one mask at a time, to the rendered guidance masks of every diffusion round.  out != in. */
GVD_TRAIN_API int gvd_mask_morphology(const float* in, float* out, long long N, int H, int W, int lo_y, int hi_y, int lo_x,
                                      int hi_x, int dilate, gvd_train_stream_t stream);

GVD_TRAIN_API const char* gvd_train_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* GVD_TRAIN_H_ */
