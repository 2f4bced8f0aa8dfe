// train_kernels.cu -- fused photometric loss (L1 + SSIM) forward/backward, densification statistics and the Adam update
// of the 3DGS training step (include/gvd_train.h).  All HBM-bound fp32 work on small tensors (a 640x480 image is 3.7 MB):
// what is bought is launch count -- one kernel where the reference issues dozens -- and the absence of host round trips.
//
// SSIM (utils/loss_utils.py:46-82): with the 11x11 Gaussian window w (sigma 1.5, zero padding) and per channel
//   mu1 = w*x, mu2 = w*y, e11 = w*(x x), e22 = w*(y y), e12 = w*(x y),
//   A1 = 2 mu1 mu2 + C1, A2 = 2 (e12 - mu1 mu2) + C2, B1 = mu1^2 + mu2^2 + C1, B2 = (e11 - mu1^2) + (e22 - mu2^2) + C2,
//   S = A1 A2 / (B1 B2).
// The window is separable; a 16x16 tile stages its 26x26 halo of x and y in shared memory, runs the horizontal pass for
// the five products, then the vertical one.  The backward needs d(sum S)/dx(p) = sum_q w(q-p) [ dS/dmu1(q) +
// 2 x(p) dS/de11(q) + y(p) dS/de12(q) ]: three more separable convolutions of maps the forward stores.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/gvd_train.h"

namespace {

thread_local std::string g_train_err;

constexpr int TILE = 16, R = 5, HALO = TILE + 2 * R;  // 26
constexpr float kC1 = 0.01f * 0.01f, kC2 = 0.03f * 0.03f;

struct Window { float w[2 * R + 1]; };

// gaussian(11, 1.5) / sum, in float32 like torch.Tensor([...]) / gauss.sum()   (loss_utils.py:36-38)
Window make_window() {
    Window g;
    float sum = 0.f;
    for (int i = 0; i < 2 * R + 1; ++i) {
        g.w[i] = (float)exp(-(double)((i - R) * (i - R)) / (2.0 * 1.5 * 1.5));
        sum += g.w[i];
    }
    for (int i = 0; i < 2 * R + 1; ++i) g.w[i] /= sum;
    return g;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// grid (ceil(W/16), ceil(H/16), C), 256 threads.  partial[block] = (sum |x-y|, sum S) of the block's pixels.
__global__ void __launch_bounds__(256) ssim_l1_forward_kernel(const float* __restrict__ img, const float* __restrict__ gt, int H, int W,
                                                              Window g, float* __restrict__ dmaps, double* __restrict__ partial) {
    __shared__ float sx[HALO][HALO + 1], sy[HALO][HALO + 1];
    __shared__ float hz[5][HALO][TILE + 1];
    __shared__ float red[2][8];
    const int c = blockIdx.z;
    const int x0 = blockIdx.x * TILE, y0 = blockIdx.y * TILE;
    const size_t plane = (size_t)H * W;
    const float* px = img + c * plane;
    const float* py = gt + c * plane;
    for (int i = threadIdx.x; i < HALO * HALO; i += 256) {
        const int r = i / HALO, q = i % HALO;
        const int yy = y0 + r - R, xx = x0 + q - R;
        const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
        sx[r][q] = in ? px[(size_t)yy * W + xx] : 0.f;   // zero padding of F.conv2d(padding=5)
        sy[r][q] = in ? py[(size_t)yy * W + xx] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < HALO * TILE; i += 256) {  // horizontal pass: 26 rows x 16 columns x 5 products
        const int r = i / TILE, q = i % TILE;
        float a = 0.f, b = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
        for (int k = 0; k < 2 * R + 1; ++k) {
            const float xv = sx[r][q + k], yv = sy[r][q + k], wk = g.w[k];
            a = fmaf(wk, xv, a);
            b = fmaf(wk, yv, b);
            aa = fmaf(wk, xv * xv, aa);
            bb = fmaf(wk, yv * yv, bb);
            ab = fmaf(wk, xv * yv, ab);
        }
        hz[0][r][q] = a; hz[1][r][q] = b; hz[2][r][q] = aa; hz[3][r][q] = bb; hz[4][r][q] = ab;
    }
    __syncthreads();
    const int tx = threadIdx.x % TILE, ty = threadIdx.x / TILE;
    const int X = x0 + tx, Y = y0 + ty;
    float l1 = 0.f, s = 0.f;
    if (X < W && Y < H) {
        float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
        for (int k = 0; k < 2 * R + 1; ++k) {
            const float wk = g.w[k];
            mu1 = fmaf(wk, hz[0][ty + k][tx], mu1);
            mu2 = fmaf(wk, hz[1][ty + k][tx], mu2);
            e11 = fmaf(wk, hz[2][ty + k][tx], e11);
            e22 = fmaf(wk, hz[3][ty + k][tx], e22);
            e12 = fmaf(wk, hz[4][ty + k][tx], e12);
        }
        const float m11 = mu1 * mu1, m22 = mu2 * mu2, m12 = mu1 * mu2;
        const float A1 = 2.f * m12 + kC1, A2 = 2.f * (e12 - m12) + kC2;
        const float B1 = m11 + m22 + kC1, B2 = (e11 - m11) + (e22 - m22) + kC2;
        const float inv = 1.0f / (B1 * B2);
        s = A1 * A2 * inv;
        l1 = fabsf(sx[ty + R][tx + R] - sy[ty + R][tx + R]);
        if (dmaps) {
            const size_t o = c * plane + (size_t)Y * W + X, n = (size_t)gridDim.z * plane;
            // dS/dmu1 = 2 mu2 (A2 - A1) / (B1 B2) - 2 mu1 S (B2 - B1) / (B1 B2);  dS/de11 = -S / B2;  dS/de12 = 2 A1 / (B1 B2)
            dmaps[o] = 2.f * mu2 * (A2 - A1) * inv - 2.f * mu1 * s * (B2 - B1) * inv;
            dmaps[n + o] = -s / B2;
            dmaps[2 * n + o] = 2.f * A1 * inv;
        }
    }
    l1 = warp_sum(l1);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = l1; red[1][threadIdx.x >> 5] = s; }
    __syncthreads();
    if (threadIdx.x < 2) {
        double t = 0.0;
        for (int w2 = 0; w2 < 8; ++w2) t += red[threadIdx.x][w2];
        const size_t b = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        partial[2 * b + threadIdx.x] = t;
    }
}

// one CTA: fixed-order sum of the per-block partials -> out[0] = mean |x-y|, out[1] = mean S (deterministic)
__global__ void __launch_bounds__(256) loss_finalize_kernel(const double* __restrict__ partial, long long blocks, double inv_n,
                                                            float* __restrict__ out) {
    __shared__ double sh[2][256];
    double a = 0.0, b = 0.0;
    for (long long i = threadIdx.x; i < blocks; i += 256) { a += partial[2 * i]; b += partial[2 * i + 1]; }
    sh[0][threadIdx.x] = a;
    sh[1][threadIdx.x] = b;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { sh[0][threadIdx.x] += sh[0][threadIdx.x + o]; sh[1][threadIdx.x] += sh[1][threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = (float)(sh[0][0] * inv_n); out[1] = (float)(sh[1][0] * inv_n); }
}

__global__ void __launch_bounds__(256) ssim_l1_backward_kernel(const float* __restrict__ img, const float* __restrict__ gt,
                                                               const float* __restrict__ dmaps, const float* __restrict__ coef,
                                                               int H, int W, Window g, float inv_n, float* __restrict__ dimg) {
    __shared__ float sm[3][HALO][HALO + 1];
    __shared__ float hz[3][HALO][TILE + 1];
    const int c = blockIdx.z;
    const int x0 = blockIdx.x * TILE, y0 = blockIdx.y * TILE;
    const size_t plane = (size_t)H * W, n = (size_t)gridDim.z * plane;
    for (int i = threadIdx.x; i < HALO * HALO; i += 256) {
        const int r = i / HALO, q = i % HALO;
        const int yy = y0 + r - R, xx = x0 + q - R;
        const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
        const size_t o = c * plane + (size_t)yy * W + xx;
#pragma unroll
        for (int m = 0; m < 3; ++m) sm[m][r][q] = in ? dmaps[m * n + o] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < HALO * TILE; i += 256) {
        const int r = i / TILE, q = i % TILE;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int k = 0; k < 2 * R + 1; ++k) {
            const float wk = g.w[k];
            a0 = fmaf(wk, sm[0][r][q + k], a0);
            a1 = fmaf(wk, sm[1][r][q + k], a1);
            a2 = fmaf(wk, sm[2][r][q + k], a2);
        }
        hz[0][r][q] = a0; hz[1][r][q] = a1; hz[2][r][q] = a2;
    }
    __syncthreads();
    const int tx = threadIdx.x % TILE, ty = threadIdx.x / TILE;
    const int X = x0 + tx, Y = y0 + ty;
    if (X >= W || Y >= H) return;
    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
#pragma unroll
    for (int k = 0; k < 2 * R + 1; ++k) {
        const float wk = g.w[k];
        v0 = fmaf(wk, hz[0][ty + k][tx], v0);
        v1 = fmaf(wk, hz[1][ty + k][tx], v1);
        v2 = fmaf(wk, hz[2][ty + k][tx], v2);
    }
    const size_t o = c * plane + (size_t)Y * W + X;
    const float xv = img[o], yv = gt[o];
    const float d = xv - yv;
    const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);   // torch.abs backward: sign(x), 0 at 0
    dimg[o] = inv_n * (coef[0] * sgn + coef[1] * (v0 + 2.f * xv * v1 + yv * v2));
}

__global__ void __launch_bounds__(256) densification_stats_kernel(const float* __restrict__ g2d, const int* __restrict__ radii, long long P,
                                                                  float* __restrict__ accum, float* __restrict__ denom,
                                                                  float* __restrict__ max_radii) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (long long)gridDim.x * blockDim.x) {
        const int r = radii[i];
        if (r <= 0) continue;
        const float gx = g2d[3 * i], gy = g2d[3 * i + 1];
        accum[i] += sqrtf(gx * gx + gy * gy);
        denom[i] += 1.0f;
        max_radii[i] = fmaxf(max_radii[i], (float)r);
    }
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long long n, float one_minus_b1, float b2, float one_minus_b2,
                                                   float step_size, float bias2_sqrt, float eps) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gi = g[i];
        const float mi = m[i] + one_minus_b1 * (gi - m[i]);              // exp_avg.lerp_(grad, 1 - beta1)
        const float vi = v[i] * b2 + one_minus_b2 * (gi * gi);           // mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bias2_sqrt + eps;                // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
        p[i] = p[i] - step_size * (mi / denom);                          // param.addcdiv_(exp_avg, denom, value=-step_size)
    }
}

// Binary erosion / dilation of N masks [N, H, W] with a rectangular structuring element given as the window offsets
// [lo, hi] it covers around a pixel on each axis (scipy.ndimage.binary_erosion / binary_dilation with np.ones((k, k)),
// border_value = 0: utils/viewcrafter_wrapper.py:602-647).  Non-zero input = set; output is 0 / 1 in float.
__global__ void __launch_bounds__(256) morph_rect_kernel(const float* __restrict__ in, float* __restrict__ out, long long N, int H, int W,
                                                         int lo_y, int hi_y, int lo_x, int hi_x, int dilate) {
    const long long total = N * H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        const int y = (int)((i / W) % H);
        const float* img = in + (i / ((long long)H * W)) * (long long)H * W;
        bool r = !dilate;  // erosion: all set (outside the image counts as unset); dilation: any set
        for (int dy = lo_y; dy <= hi_y && r != (bool)dilate; ++dy) {
            const int yy = y + dy;
            for (int dx = lo_x; dx <= hi_x; ++dx) {
                const int xx = x + dx;
                const bool set = yy >= 0 && yy < H && xx >= 0 && xx < W && img[(long long)yy * W + xx] != 0.f;
                if (dilate ? set : !set) { r = dilate; break; }
            }
        }
        out[i] = r ? 1.f : 0.f;
    }
}

int grid_for(long long n) {
    long long g = (n + 255) / 256;
    const long long cap = 148 * 8;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

long long loss_blocks(int C, int H, int W) { return (long long)((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE) * C; }

}  // namespace

extern "C" {

const char* gvd_train_last_error(void) { return g_train_err.c_str(); }

size_t gvd_photometric_loss_scratch_bytes(int C, int H, int W) {
    if (C <= 0 || H <= 0 || W <= 0) return 0;
    return (size_t)loss_blocks(C, H, W) * 2 * sizeof(double);
}

int gvd_photometric_loss_forward(const float* img, const float* gt, int C, int H, int W, float* out, float* dmaps, void* scratch,
                                 size_t scratch_bytes, gvd_train_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (C <= 0 || H <= 0 || W <= 0 || C > 65535) { g_train_err = "gvd_photometric_loss_forward: needs 0 < C <= 65535, H > 0, W > 0"; return 2; }
    if (!img || !gt || !out || !scratch) { g_train_err = "gvd_photometric_loss_forward: null pointer"; return 2; }
    if (scratch_bytes < gvd_photometric_loss_scratch_bytes(C, H, W) || (reinterpret_cast<uintptr_t>(scratch) & 7)) {
        g_train_err = "gvd_photometric_loss_forward: scratch too small or not 8-byte aligned";
        return 2;
    }
    static const Window g = make_window();
    dim3 grid((W + TILE - 1) / TILE, (H + TILE - 1) / TILE, C);
    double* partial = reinterpret_cast<double*>(scratch);
    ssim_l1_forward_kernel<<<grid, 256, 0, s>>>(img, gt, H, W, g, dmaps, partial);
    loss_finalize_kernel<<<1, 256, 0, s>>>(partial, loss_blocks(C, H, W), 1.0 / ((double)C * H * W), out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_train_err = std::string("gvd_photometric_loss_forward: ") + cudaGetErrorString(e); return 1; }
    return 0;
}

int gvd_photometric_loss_backward(const float* img, const float* gt, const float* dmaps, const float* coef, int C, int H, int W,
                                  float* dimg, gvd_train_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (C <= 0 || H <= 0 || W <= 0 || C > 65535) { g_train_err = "gvd_photometric_loss_backward: needs 0 < C <= 65535, H > 0, W > 0"; return 2; }
    if (!img || !gt || !dmaps || !coef || !dimg) { g_train_err = "gvd_photometric_loss_backward: null pointer"; return 2; }
    static const Window g = make_window();
    dim3 grid((W + TILE - 1) / TILE, (H + TILE - 1) / TILE, C);
    ssim_l1_backward_kernel<<<grid, 256, 0, s>>>(img, gt, dmaps, coef, H, W, g, (float)(1.0 / ((double)C * H * W)), dimg);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_train_err = std::string("gvd_photometric_loss_backward: ") + cudaGetErrorString(e); return 1; }
    return 0;
}

int gvd_densification_stats(const float* means2D_grad, const int* radii, long long P, float* xyz_gradient_accum, float* denom,
                            float* max_radii2D, gvd_train_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (P < 0) { g_train_err = "gvd_densification_stats: P must be >= 0"; return 2; }
    if (P == 0) return 0;
    if (!means2D_grad || !radii || !xyz_gradient_accum || !denom || !max_radii2D) { g_train_err = "gvd_densification_stats: null pointer"; return 2; }
    densification_stats_kernel<<<grid_for(P), 256, 0, s>>>(means2D_grad, radii, P, xyz_gradient_accum, denom, max_radii2D);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int gvd_mask_morphology(const float* in, float* out, long long N, int H, int W, int lo_y, int hi_y, int lo_x, int hi_x, int dilate,
                        gvd_train_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (N < 0 || H <= 0 || W <= 0 || lo_y > hi_y || lo_x > hi_x) { g_train_err = "gvd_mask_morphology: needs N >= 0, H, W > 0, lo <= hi"; return 2; }
    if (N == 0) return 0;
    if (!in || !out || in == out) { g_train_err = "gvd_mask_morphology: null pointer or in-place call (the window reads neighbours)"; return 2; }
    morph_rect_kernel<<<grid_for(N * H * W), 256, 0, s>>>(in, out, N, H, W, lo_y, hi_y, lo_x, hi_x, dilate ? 1 : 0);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int gvd_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, double lr, double beta1, double beta2,
                  double eps, int step, gvd_train_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (n < 0 || step < 1) { g_train_err = "gvd_adam_step: needs n >= 0 and step >= 1"; return 2; }
    if (n == 0) return 0;
    if (!param || !grad || !exp_avg || !exp_avg_sq) { g_train_err = "gvd_adam_step: null pointer"; return 2; }
    // scalar prologue of _single_tensor_adam in double, like the Python floats it is written with; every scalar is
    // rounded to fp32 once, where torch hands it to a kernel
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    const float step_size = (float)(lr / bc1), bias2_sqrt = (float)sqrt(bc2);
    adam_kernel<<<grid_for(n), 256, 0, s>>>(param, grad, exp_avg, exp_avg_sq, n, (float)(1.0 - beta1), (float)beta2,
                                            (float)(1.0 - beta2), step_size, bias2_sqrt, (float)eps);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // extern "C"
