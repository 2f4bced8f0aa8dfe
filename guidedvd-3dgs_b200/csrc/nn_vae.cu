// nn_vae.cu -- the one layer geometry of the VAE encoder that the denoiser's kernels do not cover (include/gvd_nn.h):
// Downsample of lvdm/modules/networks/ae_modules.py:93-106 pads the image on the RIGHT and BOTTOM only
// (F.pad(x, (0,1,0,1))) and then runs a stride-2 3x3 convolution WITHOUT padding, so tap (ky, kx) of output pixel
// (oy, ox) reads input pixel (2 oy + ky, 2 ox + kx) -- not (2 oy + ky - 1, ...) as the symmetric pad-1 convolutions do.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/gvd_nn.h"

extern thread_local std::string g_nn_err_ext;

namespace {

// x [F, H, W, C] bf16 -> col [F, Ho, Wo, 9, C], Ho = (H - 2) / 2 + 1, K order (ky, kx, c); 16-byte vectors of 8 channels
__global__ void __launch_bounds__(256) im2col3x3_down_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ col,
                                                             int F, int H, int W, int C, int Ho, int Wo) {
    const int vec = C / 8;
    const long long total = (long long)F * Ho * Wo * 9 * vec;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % vec);
        long long t = i / vec;
        const int tap = (int)(t % 9);
        t /= 9;
        const int ox = (int)(t % Wo);
        t /= Wo;
        const int oy = (int)(t % Ho);
        const int f = (int)(t / Ho);
        const int iy = 2 * oy + tap / 3, ix = 2 * ox + tap % 3;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (iy < H && ix < W) val = __ldg(reinterpret_cast<const uint4*>(x + (((size_t)f * H + iy) * W + ix) * C) + v);
        reinterpret_cast<uint4*>(col)[i] = val;
    }
}

}  // namespace

extern "C" {

int gvd_im2col3x3_down_cl(const void* x, void* col, int F, int H, int W, int C, gvd_nn_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (!x || !col) { g_nn_err_ext = "gvd_im2col3x3_down_cl: null pointer"; return 2; }
    if (C % 8 || H < 2 || W < 2) { g_nn_err_ext = "gvd_im2col3x3_down_cl: needs C % 8 == 0, H >= 2, W >= 2"; return 2; }
    const int Ho = (H - 2) / 2 + 1, Wo = (W - 2) / 2 + 1;
    const long long total = (long long)F * Ho * Wo * 9 * (C / 8);
    if (total <= 0) return 0;
    long long g = (total + 255) / 256;
    const long long cap = 148 * 16;
    im2col3x3_down_kernel<<<(unsigned)(g > cap ? cap : g), 256, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)col, F, H, W, C, Ho, Wo);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // extern "C"
