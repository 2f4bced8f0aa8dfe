// nn_backward.cu -- input-gradient (dX) kernels of the memory-bound U-Net layers, and the vector-Jacobian product of the
// DDIM `pred_x0` arithmetic: what the guided sampler (lvdm/models/samplers/ddim_guidance.py:259-337,
// `pred_x0.backward(gradient=..., inputs=x)`) needs beside the tensor-core GEMMs.  Only activations receive gradients;
// the network's parameters are frozen in that loop.
//
// Same layout as nn_kernels.cu: channels-last bf16 activations [frames, pixels, channels], fp32 arithmetic inside a
// kernel, one bf16 rounding at the output.  bf16 roundings of the forward are treated as identities by the backward
// (what autograd does for `.to(bfloat16)` under autocast).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string>

#include "../../include/gvd_nn.h"

extern thread_local std::string g_nn_err_ext;

#ifndef GVD_HOST_EMU
bool gvd_mma_temporal_attention_bwd(const void* q, const void* k, const void* v, const void* dout, void* dq, void* dk, void* dv, int B, int T,
                                    long long S, int H, float scale, cudaStream_t s);
#endif

namespace {

__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }  // approximate reciprocal, <= 2 ulp
// d/dx [x * sigmoid(x)]
__device__ __forceinline__ float dsilu(float x) {
    const float s = sigmoid_f(x);
    return s * (1.0f + x * (1.0f - s));
}
__device__ __forceinline__ float gelu(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
// d/dx [x * Phi(x)] = Phi(x) + x * phi(x)
__device__ __forceinline__ float dgelu(float x) {
    return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.39894228040143268f * __expf(-0.5f * x * x);
}
__device__ __forceinline__ float round_bf16(float x) { return __bfloat162float(__float2bfloat16(x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float2 t = __bfloat1622float2(h2[e]);
        f[2 * e] = t.x;
        f[2 * e + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
    uint4 u;
    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
    return u;
}

int grid_for(long long n, int block = 256, int cap = 148 * 16) {
    long long g = (n + block - 1) / block;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// One wave of CTAs: the backward kernels are resident two per SM (launch bounds below), so floor(296 / F) chunks per frame
// keeps every CTA on the machine at once -- the rule it replaces (~592 CTAs) ran 2.03 waves at 25 frames.  16..4096 rows
// per CTA.
// frames per launch pair, as in the forward (nn_kernels.cu::gn_group_frames): x and dy of a group stay in the L2 between
// the sums pass and the dx pass.  GVD_GN_GROUP_MB bounds 3 tensors x frames x S x C x 2 bytes; default 0 = one group (measured
// slower with groups, see nn_kernels.cu).
int gn_bwd_group_frames(int F, long long S, int C) {
    static long long budget = -1;
    if (budget < 0) {
        const char* e = getenv("GVD_GN_GROUP_MB");
        budget = (e ? atoll(e) : 0) << 20;
    }
    if (budget <= 0 || F <= 1) return F;
    const long long per_frame = 3ll * S * C * 2;
    long long fmax = budget / (per_frame > 0 ? per_frame : 1);
    if (fmax < 1) fmax = 1;
    if (fmax >= F) return F;
    const long long ngroups = (F + fmax - 1) / fmax;
    return (int)((F + ngroups - 1) / ngroups);
}

int gn_bwd_chunks(int F, long long S) {
    long long want = 296 / (F > 0 ? F : 1);
    long long maxc = (S + 15) / 16;
    long long minc = (S + 4095) / 4096;
    long long c = want < minc ? minc : want;
    if (c > maxc) c = maxc;
    if (c < 1) c = 1;
    if (c > 2048) c = 2048;
    return (int)c;
}

// ---- a thread-private ring of 16-byte cp.async copies: GN_RING row steps in flight per thread whatever the register
// budget (the first version kept two rows in registers: 16 warps per SM x 64 bytes could not cover the HBM latency --
// profiles/r02_norm_bwd_bench_before.txt: 1.7 TB/s).  No slot is shared between threads, so no barrier is needed:
// cp.async.wait_group orders a thread's own copies. ----
constexpr int GN_RING = 4;
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
#ifdef GVD_HOST_EMU
    *reinterpret_cast<uint4*>(smem_dst) = *reinterpret_cast<const uint4*>(gsrc);
#else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit() {
#ifndef GVD_HOST_EMU
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
#ifndef GVD_HOST_EMU
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}

// ---------------- GroupNorm backward (channels-last) ----------------
// y = act(xh * gamma + beta), xh = (x - mean) * rstd over S x cpg per (frame, group).  With g = dy * act'(.) * gamma:
//     dx = rstd * (g - mean(g) - xh * mean(g * xh))            (means over the same S x cpg)
// Pass 1 (this kernel): per (frame, chunk, group) partial sums of g and g * xh.  `stats` = (sum x, sum x^2) per
// (frame, group) from gvd_groupnorm_cl_stats.  Thread layout as in the forward: a thread owns 8 channels, walks rows.
// SILU: 0 none, 1 SiLU saw the bf16-rounded norm output (GroupNormSpecific), 2 SiLU in fp32 (plain nn.GroupNorm).
template <int SILU>
__device__ __forceinline__ float gn_dact(float xv, float sc, float sf) {
    const float z = fmaf(xv, sc, sf);
    return dsilu(SILU == 1 ? round_bf16(z) : z);
}

// merged-per-thread sums land in one of GN_COPIES shared copies (by row slot), so the shared atomics of the 6-32
// threads that own the same channels do not serialise on one address
constexpr int GN_COPIES = 8;

template <int SILU>
__global__ void __launch_bounds__(256, 2) gn_bwd_partial_kernel(const __nv_bfloat16* __restrict__ x,
                                                                const __nv_bfloat16* __restrict__ dy,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                const float* __restrict__ stats, int S, int C, int groups,
                                                                int rows_per_chunk, float eps, long long stat_rows,
                                                                double* __restrict__ partial) {
    extern __shared__ double gn_sh[];  // acc[GN_COPIES][groups*2] doubles, mean[groups], rstd[groups] floats, then the ring
    double* acc = gn_sh;
    float* smean = reinterpret_cast<float*>(acc + GN_COPIES * groups * 2);
    float* srstd = smean + groups;
    uint4* ring = reinterpret_cast<uint4*>(gn_sh) + (GN_COPIES * groups * 2 * sizeof(double) + groups * 2 * sizeof(float) + 15) / 16;
    const int f = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
    const int cpg = C / groups, vecs = C / 8;
    for (int i = threadIdx.x; i < GN_COPIES * groups * 2; i += blockDim.x) acc[i] = 0.0;
    if (threadIdx.x < groups) {
        const double n = (double)stat_rows * cpg;  // rows behind the statistics (> S when they were summed across shards)
        const double s = stats[((size_t)f * groups + threadIdx.x) * 2], q = stats[((size_t)f * groups + threadIdx.x) * 2 + 1];
        const double mean = s / n;
        const double var = fmax(q / n - mean * mean, 0.0);
        smean[threadIdx.x] = (float)mean;
        srstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
    }
    __syncthreads();
    const int r0 = chunk * rows_per_chunk, r1 = min(S, r0 + rows_per_chunk);
    const int vper = vecs <= 256 ? vecs : 256;
    const int rows_par = vecs <= 256 ? 256 / vecs : 1;
    const int rsub = threadIdx.x / vper;
    if (rsub < rows_par)
        for (int v = threadIdx.x % vper; v < vecs; v += vper) {
            float sc[8], sf[8], gm[8], mu[8], a1[8], a2[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int c = 8 * v + e, g = c / cpg;
                mu[e] = smean[g];
                gm[e] = gamma[c];
                sc[e] = srstd[g] * gm[e];
                sf[e] = beta[c] - mu[e] * sc[e];
                a1[e] = a2[e] = 0.f;
            }
            const uint4* xin = reinterpret_cast<const uint4*>(x + (size_t)f * S * C) + v;
            const uint4* din = reinterpret_cast<const uint4*>(dy + (size_t)f * S * C) + v;
            const int first = r0 + rsub;
            const int n = first < r1 ? (r1 - first + rows_par - 1) / rows_par : 0;
#pragma unroll
            for (int k = 0; k < GN_RING; ++k) {
                if (k < n) {
                    cp_async16(&ring[(2 * k) * 256 + threadIdx.x], xin + (size_t)(first + k * rows_par) * vecs);
                    cp_async16(&ring[(2 * k + 1) * 256 + threadIdx.x], din + (size_t)(first + k * rows_par) * vecs);
                }
                cp_async_commit();
            }
            for (int k = 0; k < n; ++k) {
                cp_async_wait<GN_RING - 1>();
                const int slot = k % GN_RING;
                const uint4 ux = ring[(2 * slot) * 256 + threadIdx.x], ud = ring[(2 * slot + 1) * 256 + threadIdx.x];
                if (k + GN_RING < n) {
                    cp_async16(&ring[(2 * slot) * 256 + threadIdx.x], xin + (size_t)(first + (k + GN_RING) * rows_par) * vecs);
                    cp_async16(&ring[(2 * slot + 1) * 256 + threadIdx.x], din + (size_t)(first + (k + GN_RING) * rows_par) * vecs);
                }
                cp_async_commit();
                float xv[8], dv[8];
                unpack8(ux, xv);
                unpack8(ud, dv);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    // g / gamma: the factor gamma is applied once, to the sums
                    const float g = SILU ? dv[e] * gn_dact<SILU>(xv[e], sc[e], sf[e]) : dv[e];
                    a1[e] += g;
                    a2[e] = fmaf(g, xv[e] - mu[e], a2[e]);
                }
            }
            // channels of one group are summed in the thread first (cpg >= 8: at most two groups per thread)
            double* my = acc + (rsub % GN_COPIES) * groups * 2;
            int gcur = (8 * v) / cpg;
            double s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int g = (8 * v + e) / cpg;
                if (g != gcur) {
                    atomicAdd(&my[2 * gcur], s1);
                    atomicAdd(&my[2 * gcur + 1], s2 * (double)srstd[gcur]);
                    s1 = s2 = 0.0;
                    gcur = g;
                }
                s1 += (double)a1[e] * (double)gm[e];
                s2 += (double)a2[e] * (double)gm[e];
            }
            atomicAdd(&my[2 * gcur], s1);
            atomicAdd(&my[2 * gcur + 1], s2 * (double)srstd[gcur]);
        }
    __syncthreads();
    double* out = partial + ((size_t)f * nchunks + chunk) * groups * 2;
    for (int i = threadIdx.x; i < groups * 2; i += blockDim.x) {
        double t = 0.0;
#pragma unroll
        for (int cpy = 0; cpy < GN_COPIES; ++cpy) t += acc[cpy * groups * 2 + i];
        out[i] = t;
    }
}

// Pass 2: dx from the folded sums.  With gr = gamma rstd, A = -rstd^2 m2, B = -rstd m1 - mean A (per channel):
//     dx = dy act'(.) gr + x A + B
template <int SILU>
__global__ void __launch_bounds__(256, 2) gn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ x,
                                                              const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              const float* __restrict__ stats, const double* __restrict__ partial,
                                                              int S, int C, int groups, int nchunks, int rows_per_cta, float eps,
                                                              long long stat_rows) {
    extern __shared__ double gn_shd[];  // fold[256] doubles, then mean, rstd, m1, m2: [groups] floats each, then the ring
    double* fold = gn_shd;
    float* smean = reinterpret_cast<float*>(fold + 256);
    float* srstd = smean + groups;
    float* sm1 = srstd + groups;
    float* sm2 = sm1 + groups;
    uint4* ring = reinterpret_cast<uint4*>(gn_shd) + (256 * sizeof(double) + groups * 4 * sizeof(float) + 15) / 16;
    const int f = blockIdx.y;
    const int cpg = C / groups, vecs = C / 8;
    {  // the chunks' partial sums, folded by all threads: entry i = tid % (2 groups), every (256 / (2 groups))-th chunk
        const int width = groups * 2, parts = 256 / width;
        const int i = threadIdx.x % width, part = threadIdx.x / width;
        double t = 0.0;
        if (part < parts)
            for (int c = part; c < nchunks; c += parts) t += partial[((size_t)f * nchunks + c) * width + i];
        fold[threadIdx.x] = t;
        __syncthreads();
        if (threadIdx.x < groups) {
            const double n = (double)stat_rows * cpg;
            const double s = stats[((size_t)f * groups + threadIdx.x) * 2], q = stats[((size_t)f * groups + threadIdx.x) * 2 + 1];
            const double mean = s / n;
            const double var = fmax(q / n - mean * mean, 0.0);
            smean[threadIdx.x] = (float)mean;
            srstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
            double b1 = 0.0, b2 = 0.0;
            for (int pp = 0; pp < parts; ++pp) {
                b1 += fold[pp * width + 2 * threadIdx.x];
                b2 += fold[pp * width + 2 * threadIdx.x + 1];
            }
            sm1[threadIdx.x] = (float)(b1 / n);
            sm2[threadIdx.x] = (float)(b2 / n);
        }
        __syncthreads();
    }
    const int vper = vecs <= 256 ? vecs : 256;
    const int rows_par = vecs <= 256 ? 256 / vecs : 1;
    const int rsub = threadIdx.x / vper;
    if (rsub >= rows_par) return;
    const int r0 = blockIdx.x * rows_per_cta, r1 = min(S, r0 + rows_per_cta);
    for (int v = threadIdx.x % vper; v < vecs; v += vper) {
        float sc[8], sf[8], gr[8], ca[8], cb[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int c = 8 * v + e, g = c / cpg;
            const float mu = smean[g], rs = srstd[g], gm = gamma[c];
            sc[e] = rs * gm;
            sf[e] = beta[c] - mu * sc[e];
            gr[e] = gm * rs;
            ca[e] = -rs * rs * sm2[g];
            cb[e] = -rs * sm1[g] - mu * ca[e];
        }
        const uint4* xin = reinterpret_cast<const uint4*>(x + (size_t)f * S * C) + v;
        const uint4* din = reinterpret_cast<const uint4*>(dy + (size_t)f * S * C) + v;
        uint4* dout = reinterpret_cast<uint4*>(dx + (size_t)f * S * C) + v;
        const int first = r0 + rsub;
        const int n = first < r1 ? (r1 - first + rows_par - 1) / rows_par : 0;
#pragma unroll
        for (int k = 0; k < GN_RING; ++k) {
            if (k < n) {
                cp_async16(&ring[(2 * k) * 256 + threadIdx.x], xin + (size_t)(first + k * rows_par) * vecs);
                cp_async16(&ring[(2 * k + 1) * 256 + threadIdx.x], din + (size_t)(first + k * rows_par) * vecs);
            }
            cp_async_commit();
        }
        for (int k = 0; k < n; ++k) {
            cp_async_wait<GN_RING - 1>();
            const int slot = k % GN_RING;
            const uint4 ux = ring[(2 * slot) * 256 + threadIdx.x], ud = ring[(2 * slot + 1) * 256 + threadIdx.x];
            if (k + GN_RING < n) {
                cp_async16(&ring[(2 * slot) * 256 + threadIdx.x], xin + (size_t)(first + (k + GN_RING) * rows_par) * vecs);
                cp_async16(&ring[(2 * slot + 1) * 256 + threadIdx.x], din + (size_t)(first + (k + GN_RING) * rows_par) * vecs);
            }
            cp_async_commit();
            float xv[8], dv[8], o[8];
            unpack8(ux, xv);
            unpack8(ud, dv);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float g = SILU ? dv[e] * gn_dact<SILU>(xv[e], sc[e], sf[e]) : dv[e];
                o[e] = fmaf(g, gr[e], fmaf(xv[e], ca[e], cb[e]));
            }
            dout[(size_t)(first + k * rows_par) * vecs] = pack8(o);
        }
    }
}

// Fold the per-chunk partial sums of one frame into (sum g, sum g xh) per group: the exchange unit when the rows of a
// group are spread over several GPUs (the mirror image of gn_fold_kernel in the forward).
__global__ void __launch_bounds__(256) gn_bwd_fold_kernel(const double* __restrict__ partial, double* __restrict__ sums, int nchunks,
                                                          int groups) {
    const int f = blockIdx.x;
    for (int i = threadIdx.x; i < groups * 2; i += blockDim.x) {
        double a = 0.0;
        for (int c = 0; c < nchunks; ++c) a += partial[((size_t)f * nchunks + c) * groups * 2 + i];
        sums[(size_t)f * groups * 2 + i] = a;
    }
}

// dynamic shared memory of the two kernels above (group arrays + the cp.async ring)
inline size_t gn_bwd_partial_smem(int groups) {
    return (GN_COPIES * groups * 2 * sizeof(double) + groups * 2 * sizeof(float) + 15) / 16 * 16 + (size_t)GN_RING * 2 * 256 * 16;
}
inline size_t gn_bwd_apply_smem(int groups) { return (256 * sizeof(double) + groups * 4 * sizeof(float) + 15) / 16 * 16 + (size_t)GN_RING * 2 * 256 * 16; }

void launch_gn_bwd_partial(int do_silu, dim3 grid, cudaStream_t s, const __nv_bfloat16* x, const __nv_bfloat16* dy, const float* gamma,
                           const float* beta, const float* stats, int S, int C, int groups, int rows_per_chunk, float eps, long long stat_rows,
                           double* partial) {
    const size_t sm = gn_bwd_partial_smem(groups);
    if (do_silu == 0) gn_bwd_partial_kernel<0><<<grid, 256, sm, s>>>(x, dy, gamma, beta, stats, S, C, groups, rows_per_chunk, eps, stat_rows, partial);
    else if (do_silu == 1) gn_bwd_partial_kernel<1><<<grid, 256, sm, s>>>(x, dy, gamma, beta, stats, S, C, groups, rows_per_chunk, eps, stat_rows, partial);
    else gn_bwd_partial_kernel<2><<<grid, 256, sm, s>>>(x, dy, gamma, beta, stats, S, C, groups, rows_per_chunk, eps, stat_rows, partial);
}
void launch_gn_bwd_apply(int do_silu, dim3 grid, cudaStream_t s, const __nv_bfloat16* x, const __nv_bfloat16* dy, __nv_bfloat16* dx,
                         const float* gamma, const float* beta, const float* stats, const double* partial, int S, int C, int groups, int nchunks,
                         int rows_per_cta, float eps, long long stat_rows) {
    const size_t sm = gn_bwd_apply_smem(groups);
    if (do_silu == 0) gn_bwd_apply_kernel<0><<<grid, 256, sm, s>>>(x, dy, dx, gamma, beta, stats, partial, S, C, groups, nchunks, rows_per_cta, eps, stat_rows);
    else if (do_silu == 1) gn_bwd_apply_kernel<1><<<grid, 256, sm, s>>>(x, dy, dx, gamma, beta, stats, partial, S, C, groups, nchunks, rows_per_cta, eps, stat_rows);
    else gn_bwd_apply_kernel<2><<<grid, 256, sm, s>>>(x, dy, dx, gamma, beta, stats, partial, S, C, groups, nchunks, rows_per_cta, eps, stat_rows);
}

// ---------------- adjoint of the nearest-neighbour 2x upsampling: dx[h, w] = sum of the 2 x 2 block of dy ----------------
// fp32 sum of the four bf16 gradients, one rounding (what autograd's upsample_nearest2d_backward does under autocast)
__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, int F,
                                                             int H, int W, int C) {
    const int vec = C / 8;
    const long long total = (long long)F * H * W * vec;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % vec);
        long long t = i / vec;
        const int ix = (int)(t % W);
        t /= W;
        const int iy = (int)(t % H);
        const int f = (int)(t / H);
        const uint4* in = reinterpret_cast<const uint4*>(dy) + ((((size_t)f * 2 * H + 2 * iy) * 2 * W + 2 * ix) * vec + v);
        const uint4 u0 = __ldg(in), u1 = __ldg(in + vec), u2 = __ldg(in + (size_t)2 * W * vec), u3 = __ldg(in + (size_t)2 * W * vec + vec);
        float a[8], b[8], c[8], d[8], o[8];
        unpack8(u0, a);
        unpack8(u1, b);
        unpack8(u2, c);
        unpack8(u3, d);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = (a[e] + b[e]) + (c[e] + d[e]);
        reinterpret_cast<uint4*>(dx)[i] = pack8(o);
    }
}

// ---------------- LayerNorm backward over the last dim (one warp per row) ----------------
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                            const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx,
                                                            const float* __restrict__ gamma, long long rows, int C, float eps) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const __nv_bfloat162* xin = reinterpret_cast<const __nv_bfloat162*>(x + row * C);
    const __nv_bfloat162* din = reinterpret_cast<const __nv_bfloat162*>(dy + row * C);
    const int pairs = C / 2;
    float s = 0.f;
    for (int i = lane; i < pairs; i += 32) {
        const float2 v = __bfloat1622float2(xin[i]);
        s += v.x + v.y;
    }
    const float mean = warp_sum(s) / C;
    float q = 0.f;
    for (int i = lane; i < pairs; i += 32) {
        const float2 v = __bfloat1622float2(xin[i]);
        q += (v.x - mean) * (v.x - mean) + (v.y - mean) * (v.y - mean);
    }
    const float rstd = rsqrtf(warp_sum(q) / C + eps);
    float s1 = 0.f, s2 = 0.f;
    for (int i = lane; i < pairs; i += 32) {
        const float2 v = __bfloat1622float2(xin[i]);
        const float2 d = __bfloat1622float2(din[i]);
        const float g0 = d.x * gamma[2 * i], g1 = d.y * gamma[2 * i + 1];
        s1 += g0 + g1;
        s2 += g0 * (v.x - mean) * rstd + g1 * (v.y - mean) * rstd;
    }
    const float m1 = warp_sum(s1) / C, m2 = warp_sum(s2) / C;
    __nv_bfloat162* dout = reinterpret_cast<__nv_bfloat162*>(dx + row * C);
    for (int i = lane; i < pairs; i += 32) {
        const float2 v = __bfloat1622float2(xin[i]);
        const float2 d = __bfloat1622float2(din[i]);
        const float g0 = d.x * gamma[2 * i], g1 = d.y * gamma[2 * i + 1];
        dout[i] = __floats2bfloat162_rn(rstd * (g0 - m1 - (v.x - mean) * rstd * m2), rstd * (g1 - m1 - (v.y - mean) * rstd * m2));
    }
}

// ---------------- GEGLU backward: out = a * bf16(gelu(g)), h = [a | g] ----------------
__global__ void __launch_bounds__(256) geglu_bwd_kernel(const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ dout,
                                                        __nv_bfloat16* __restrict__ dh, long long rows, int D) {
    const long long pairs = rows * (D / 2);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / (D / 2);
        const int j = (int)(i % (D / 2));
        const __nv_bfloat162* row = reinterpret_cast<const __nv_bfloat162*>(h + r * 2 * D);
        const float2 a = __bfloat1622float2(row[j]);
        const float2 g = __bfloat1622float2(row[D / 2 + j]);
        const float2 d = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(dout + r * D)[j]);
        __nv_bfloat162* orow = reinterpret_cast<__nv_bfloat162*>(dh + r * 2 * D);
        orow[j] = __floats2bfloat162_rn(d.x * round_bf16(gelu(g.x)), d.y * round_bf16(gelu(g.y)));
        orow[D / 2 + j] = __floats2bfloat162_rn(d.x * a.x * dgelu(g.x), d.y * a.y * dgelu(g.y));
    }
}

// ---------------- softmax backward per row: ds = p * (dp - sum_j p_j dp_j) (one warp per row) ----------------
// ds may alias dp (a lane reads and writes only its own columns), so neither is __restrict__.
__global__ void __launch_bounds__(256) softmax_bwd_rows_kernel(const __nv_bfloat16* __restrict__ p, const __nv_bfloat16* dp,
                                                               __nv_bfloat16* ds, long long ld, long long rows, int cols) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const __nv_bfloat16* pr = p + row * ld;
    const __nv_bfloat16* dr = dp + row * ld;
    float dot = 0.f;
    for (int i = lane; i < cols; i += 32) dot = fmaf(__bfloat162float(pr[i]), __bfloat162float(dr[i]), dot);
    dot = warp_sum(dot);
    __nv_bfloat16* out = ds + row * ld;
    for (int i = lane; i < cols; i += 32) out[i] = __float2bfloat16(__bfloat162float(pr[i]) * (__bfloat162float(dr[i]) - dot));
    for (long long i = cols + lane; i < ld; i += 32) out[i] = __float2bfloat16(0.f);  // zero the K padding of the next GEMM
}

// ---------------- col2im: adjoint of im2col3x3_kernel (gather form, no atomics) ----------------
// dcol [F, Ho, Wo, 9, C] -> dx [F, H, W, C]: every input pixel sums the (output pixel, tap) pairs that read it.
__global__ void __launch_bounds__(256) col2im3x3_kernel(const __nv_bfloat16* __restrict__ dcol, __nv_bfloat16* __restrict__ dx,
                                                        int F, int H, int W, int C, int Ho, int Wo, int stride, int up) {
    const int vec = C / 8;
    const long long total = (long long)F * H * W * vec;
    const int sub = up ? 2 : 1;  // an input pixel stands for sub x sub pixels of the upsampled image the taps walk
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % vec);
        long long t = i / vec;
        const int ix = (int)(t % W);
        t /= W;
        const int iy = (int)(t % H);
        const int f = (int)(t / H);
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
        for (int a = 0; a < sub; ++a)
            for (int ky = 0; ky < 3; ++ky) {
                const int ty = iy * sub + a + 1 - ky;  // = oy * stride
                if (ty < 0 || ty % stride) continue;
                const int oy = ty / stride;
                if (oy >= Ho) continue;
                for (int b = 0; b < sub; ++b)
                    for (int kx = 0; kx < 3; ++kx) {
                        const int tx = ix * sub + b + 1 - kx;
                        if (tx < 0 || tx % stride) continue;
                        const int ox = tx / stride;
                        if (ox >= Wo) continue;
                        float d[8];
                        unpack8(__ldg(reinterpret_cast<const uint4*>(dcol + ((((size_t)f * Ho + oy) * Wo + ox) * 9 + ky * 3 + kx) * C) + v), d);
#pragma unroll
                        for (int e = 0; e < 8; ++e) acc[e] += d[e];
                    }
            }
        reinterpret_cast<uint4*>(dx)[i] = pack8(acc);
    }
}

// adjoint of im2col_t3_kernel: dcol [B, T, S, 3, C] -> dx [B, T, S, C]
__global__ void __launch_bounds__(256) col2im_t3_kernel(const __nv_bfloat16* __restrict__ dcol, __nv_bfloat16* __restrict__ dx,
                                                        int B, int T, long long S, int C) {
    const int vec = C / 8;
    const long long total = (long long)B * T * S * vec;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % vec);
        long long t = i / vec;
        const long long s = t % S;
        t /= S;
        const int it = (int)(t % T);
        const int b = (int)(t / T);
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
        for (int tap = 0; tap < 3; ++tap) {
            const int tt = it - tap + 1;  // the output frame whose tap `tap` read frame `it`
            if (tt < 0 || tt >= T) continue;
            float d[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(dcol + ((((size_t)b * T + tt) * S + s) * 3 + tap) * C) + v), d);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] += d[e];
        }
        reinterpret_cast<uint4*>(dx)[i] = pack8(acc);
    }
}

// ---------------- temporal self-attention backward: T <= 32 frames per (pixel, head), d = 64 ----------------
// One warp per (b, s, h) sequence, like the forward.  Phase 1: lane i owns query frame i -- recomputes its softmax
// row with the forward's rounding points, dP = dO V^T, dS = P o (dP - rowsum(P o dP)), dQ = scale dS K; P (as the bf16
// values the PV product consumed) and scale*dS go to shared memory.  Phase 2: lane j owns key frame j --
// dK_j = sum_i dS_ij Q_i, dV_j = sum_i P_ij dO_i, the Q / dO rows read as warp-wide broadcasts.
#define TAB_WARPS 4
#define TAB_LD 33  // row stride of the T x T matrices: conflict-free by row (phase 1) and by column (phase 2)
struct TabSmem {
    __nv_bfloat16 q[32 * 64], k[32 * 64], v[32 * 64], d[32 * 64];
    float p[32 * TAB_LD], ds[32 * TAB_LD];
};

__device__ __forceinline__ float dot64(const float* a, const __nv_bfloat16* row) {
    const uint4* rp = reinterpret_cast<const uint4*>(row);
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float t[8];
        unpack8(rp[c], t);
#pragma unroll
        for (int e = 0; e < 8; ++e) dot = fmaf(a[c * 8 + e], t[e], dot);
    }
    return dot;
}
__device__ __forceinline__ void axpy64(float* acc, float w, const __nv_bfloat16* row) {
    const uint4* rp = reinterpret_cast<const uint4*>(row);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float t[8];
        unpack8(rp[c], t);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[c * 8 + e] = fmaf(w, t[e], acc[c * 8 + e]);
    }
}
__device__ __forceinline__ void load_row64(float* dst, const __nv_bfloat16* row) {
    const uint4* rp = reinterpret_cast<const uint4*>(row);
#pragma unroll
    for (int c = 0; c < 8; ++c) unpack8(__ldg(rp + c), dst + c * 8);
}
__device__ __forceinline__ void store_row64(__nv_bfloat16* row, const float* src) {
    uint4* rp = reinterpret_cast<uint4*>(row);
#pragma unroll
    for (int c = 0; c < 8; ++c) rp[c] = pack8(src + c * 8);
}

__global__ void __launch_bounds__(TAB_WARPS * 32) temporal_attn_bwd_kernel(
    const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
    const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ dq, __nv_bfloat16* __restrict__ dk,
    __nv_bfloat16* __restrict__ dv, int B, int T, long long S, int H, float scale) {
    extern __shared__ __align__(16) unsigned char tab_raw[];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TabSmem& sm = reinterpret_cast<TabSmem*>(tab_raw)[wib];
    const long long w = (long long)blockIdx.x * TAB_WARPS + wib;
    const long long total = (long long)B * S * H;
    if (w >= total) return;
    const int h = (int)(w % H);
    const long long s = (w / H) % S;
    const int b = (int)(w / (H * S));
    const long long tstride = S * H * 64;
    const size_t base = ((size_t)b * T * S + s) * H * 64 + (size_t)h * 64;
    for (int r = lane >> 3; r < T; r += 4) {  // 8 lanes x 16 B cover one 128-byte row; 4 rows per pass
        const int c = lane & 7;
        reinterpret_cast<uint4*>(&sm.q[r * 64])[c] = __ldg(reinterpret_cast<const uint4*>(q + base + r * tstride) + c);
        reinterpret_cast<uint4*>(&sm.k[r * 64])[c] = __ldg(reinterpret_cast<const uint4*>(k + base + r * tstride) + c);
        reinterpret_cast<uint4*>(&sm.v[r * 64])[c] = __ldg(reinterpret_cast<const uint4*>(v + base + r * tstride) + c);
        reinterpret_cast<uint4*>(&sm.d[r * 64])[c] = __ldg(reinterpret_cast<const uint4*>(dout + base + r * tstride) + c);
    }
    __syncwarp();
    if (lane < T) {  // ---- phase 1: query frame `lane`
        float* prow = &sm.p[lane * TAB_LD];
        float* dsrow = &sm.ds[lane * TAB_LD];
        float reg[64];
        load_row64(reg, q + base + lane * tstride);
        float m = -INFINITY;
#pragma unroll 1
        for (int j = 0; j < T; ++j) {
            // einsum output rounded to bf16, scaled in bf16, then the fp32 softmax (attention.py:103)
            const float xs = round_bf16(round_bf16(dot64(reg, &sm.k[j * 64])) * scale);
            prow[j] = xs;
            m = fmaxf(m, xs);
        }
        float l = 0.f;
#pragma unroll 1
        for (int j = 0; j < T; ++j) {
            const float e = __expf(prow[j] - m);
            prow[j] = e;
            l += e;
        }
        const float inv = 1.0f / l;
        load_row64(reg, dout + base + lane * tstride);
        float rowdot = 0.f;
#pragma unroll 1
        for (int j = 0; j < T; ++j) {
            const float pj = prow[j] * inv;
            const float dpj = dot64(reg, &sm.v[j * 64]);
            prow[j] = pj;
            dsrow[j] = dpj;
            rowdot = fmaf(pj, dpj, rowdot);
        }
#pragma unroll
        for (int c = 0; c < 64; ++c) reg[c] = 0.f;
#pragma unroll 1
        for (int j = 0; j < T; ++j) {
            const float pj = prow[j];
            const float dsj = pj * (dsrow[j] - rowdot) * scale;
            dsrow[j] = dsj;
            prow[j] = round_bf16(pj);  // what the PV product multiplied V with
            axpy64(reg, dsj, &sm.k[j * 64]);
        }
        store_row64(dq + base + lane * tstride, reg);
    }
    __syncwarp();
    if (lane < T) {  // ---- phase 2: key frame `lane`
        float ak[64], av[64];
#pragma unroll
        for (int c = 0; c < 64; ++c) ak[c] = av[c] = 0.f;
#pragma unroll 1
        for (int i = 0; i < T; ++i) {
            axpy64(ak, sm.ds[i * TAB_LD + lane], &sm.q[i * 64]);
            axpy64(av, sm.p[i * TAB_LD + lane], &sm.d[i * 64]);
        }
        store_row64(dk + base + lane * tstride, ak);
        store_row64(dv + base + lane * tstride, av);
    }
}

// ---------------- vector-Jacobian product of the guided DDIM step's pred_x0 (ddim_guidance.py:263-278) ----------------
// forward:  mo = e_u + s (e_c - e_u);  v = mo * (phi * std(e_c)/std(mo) + 1 - phi)   (rescale_noise_cfg, unbiased stds
//           over the whole latent, utils_diffusion.py:147-158);  pred_x0 = r * (sqrt_ac * x - sqrt_1mac * v).
// Given G = dL/dpred_x0 this produces dL/dx through the explicit x term and dL/de_c, dL/de_u (the cotangents of the two
// U-Net forwards).  Pass 1: sums of e_c, e_c^2, mo, mo^2 and A = sum(Gv * mo) with Gv = -r * sqrt_1mac * G.
__global__ void __launch_bounds__(256) ddim_vjp_reduce_kernel(const float* __restrict__ e_c, const float* __restrict__ e_u,
                                                              const float* __restrict__ G, long long n, float cfg, float gv_coef,
                                                              double* __restrict__ stats) {
    double acc[5] = {0, 0, 0, 0, 0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float c = e_c[i], u = e_u[i];
        const float m = u + cfg * (c - u);
        acc[0] += c;
        acc[1] += (double)c * c;
        acc[2] += m;
        acc[3] += (double)m * m;
        acc[4] += (double)(gv_coef * G[i]) * m;
    }
    __shared__ double sh[5][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < 5; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[a] += __shfl_xor_sync(0xffffffffu, acc[a], o);
        if (lane == 0) sh[a][warp] = acc[a];
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double t = 0;
        for (int w2 = 0; w2 < 8; ++w2) t += sh[threadIdx.x][w2];
        atomicAdd(&stats[threadIdx.x], t);
    }
}

__global__ void __launch_bounds__(256) ddim_vjp_apply_kernel(const float* __restrict__ e_c, const float* __restrict__ e_u,
                                                             const float* __restrict__ G, float* __restrict__ dx,
                                                             float* __restrict__ de_c, float* __restrict__ de_u, long long n,
                                                             float cfg, float phi, float gv_coef, float gx_coef,
                                                             const double* __restrict__ stats) {
    const double nn = (double)n;
    const double mean_c = stats[0] / nn, mean_m = stats[2] / nn;
    float factor = 1.0f, kc = 0.f, km = 0.f;
    if (phi > 0.f) {
        const double var_c = (stats[1] - stats[0] * stats[0] / nn) / (nn - 1.0);
        const double var_m = (stats[3] - stats[2] * stats[2] / nn) / (nn - 1.0);
        const double sd_c = sqrt(var_c), sd_m = sqrt(var_m);
        factor = (float)(phi * sd_c / sd_m + (1.0 - phi));
        // d ratio / d e_c_i =  (e_c_i - mean_c) / ((n-1) sd_c sd_m);  d ratio / d mo_i = -sd_c (mo_i - mean_m) / ((n-1) sd_m^3)
        kc = (float)(stats[4] * phi / ((nn - 1.0) * sd_c * sd_m));
        km = (float)(-stats[4] * phi * sd_c / ((nn - 1.0) * sd_m * sd_m * sd_m));
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float c = e_c[i], u = e_u[i], g = G[i];
        const float m = u + cfg * (c - u);
        const float dmo = gv_coef * g * factor + km * (m - (float)mean_m);
        de_c[i] = cfg * dmo + kc * (c - (float)mean_c);
        if (de_u) de_u[i] = (1.0f - cfg) * dmo;
        dx[i] = gx_coef * g;
    }
}

}  // namespace

extern "C" {

size_t gvd_groupnorm_bwd_tmp_bytes(int F, long long S, int groups) {
    // the frames may run in groups (gn_bwd_group_frames), each with its own chunking: the largest of them all
    size_t most = (size_t)F * (gn_bwd_chunks(F, S) + 1);
    for (int fg = 1; fg < F; ++fg) {
        const size_t n = (size_t)fg * (gn_bwd_chunks(fg, S) + 1);
        if (n > most) most = n;
    }
    return most * groups * 2 * sizeof(double);
}

static int gn_bwd_check(const char* who, int C, int groups, int do_silu) {
    if (C % groups != 0 || C % 8 != 0 || groups > 128 || do_silu < 0 || do_silu > 2) {
        g_nn_err_ext = std::string(who) + ": needs C % groups == 0, C % 8 == 0, groups <= 128, do_silu in 0..2";
        return 2;
    }
    return 0;
}

int gvd_groupnorm_cl_bwd(const void* x, const void* dy, void* dx, const float* gamma, const float* beta, const float* stats, int F,
                         long long S, int C, int groups, float eps, int do_silu, void* tmp, size_t tmp_bytes,
                         gvd_nn_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (F <= 0 || S <= 0) return 0;
    if (!x || !dy || !dx || !gamma || !beta || !stats || !tmp) { g_nn_err_ext = "gvd_groupnorm_cl_bwd: null pointer"; return 2; }
    if (gn_bwd_check("gvd_groupnorm_cl_bwd", C, groups, do_silu)) return 2;
    if (reinterpret_cast<uintptr_t>(tmp) & 7) { g_nn_err_ext = "gvd_groupnorm_cl_bwd: scratch must be 8-byte aligned"; return 2; }
    double* partial = reinterpret_cast<double*>(tmp);
    const int fg = gn_bwd_group_frames(F, S, C);
    for (int f0 = 0; f0 < F; f0 += fg) {
        const int Fg = F - f0 < fg ? F - f0 : fg;
        int nchunks = gn_bwd_chunks(Fg, S);
        const int rows_per_chunk = (int)((S + nchunks - 1) / nchunks);
        nchunks = (int)((S + rows_per_chunk - 1) / rows_per_chunk);
        if (tmp_bytes < (size_t)Fg * nchunks * groups * 2 * sizeof(double)) { g_nn_err_ext = "gvd_groupnorm_cl_bwd: scratch too small"; return 2; }
        const size_t off = (size_t)f0 * S * C;
        const float* st = stats + (size_t)f0 * groups * 2;
        launch_gn_bwd_partial(do_silu, dim3(nchunks, Fg), s, (const __nv_bfloat16*)x + off, (const __nv_bfloat16*)dy + off, gamma, beta, st, (int)S, C,
                              groups, rows_per_chunk, eps, S, partial);
        launch_gn_bwd_apply(do_silu, dim3(nchunks, Fg), s, (const __nv_bfloat16*)x + off, (const __nv_bfloat16*)dy + off, (__nv_bfloat16*)dx + off, gamma,
                            beta, st, partial, (int)S, C, groups, nchunks, rows_per_chunk, eps, S);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_nn_err_ext = std::string("gvd_groupnorm_cl_bwd: ") + cudaGetErrorString(e); return 1; }
    return 0;
}

int gvd_groupnorm_cl_bwd_sums(const void* x, const void* dy, const float* gamma, const float* beta, const float* stats, double* sums,
                              int F, long long S, long long stat_rows, int C, int groups, float eps, int do_silu, void* tmp,
                              size_t tmp_bytes, gvd_nn_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (F <= 0) return 0;
    if (!sums || !stats) { g_nn_err_ext = "gvd_groupnorm_cl_bwd_sums: null pointer"; return 2; }
    if (gn_bwd_check("gvd_groupnorm_cl_bwd_sums", C, groups, do_silu)) return 2;
    if (S <= 0) return cudaMemsetAsync(sums, 0, sizeof(double) * F * groups * 2, s) == cudaSuccess ? 0 : 1;  // a shard without rows
    if (!x || !dy || !gamma || !beta || !tmp || stat_rows < S) { g_nn_err_ext = "gvd_groupnorm_cl_bwd_sums: null pointer or stat_rows < S"; return 2; }
    if (reinterpret_cast<uintptr_t>(tmp) & 7) { g_nn_err_ext = "gvd_groupnorm_cl_bwd_sums: scratch must be 8-byte aligned"; return 2; }
    int nchunks = gn_bwd_chunks(F, S);
    const int rows_per_chunk = (int)((S + nchunks - 1) / nchunks);
    nchunks = (int)((S + rows_per_chunk - 1) / rows_per_chunk);
    if (tmp_bytes < (size_t)F * nchunks * groups * 2 * sizeof(double)) { g_nn_err_ext = "gvd_groupnorm_cl_bwd_sums: scratch too small"; return 2; }
    double* partial = reinterpret_cast<double*>(tmp);
    launch_gn_bwd_partial(do_silu, dim3(nchunks, F), s, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, gamma, beta, stats, (int)S, C, groups,
                          rows_per_chunk, eps, stat_rows, partial);
    gn_bwd_fold_kernel<<<F, 256, 0, s>>>(partial, sums, nchunks, groups);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_nn_err_ext = std::string("gvd_groupnorm_cl_bwd_sums: ") + cudaGetErrorString(e); return 1; }
    return 0;
}

int gvd_groupnorm_cl_bwd_apply(const void* x, const void* dy, void* dx, const float* gamma, const float* beta, const float* stats,
                               const double* sums, int F, long long S, long long stat_rows, int C, int groups, float eps, int do_silu,
                               gvd_nn_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (F <= 0 || S <= 0) return 0;
    if (!x || !dy || !dx || !gamma || !beta || !stats || !sums || stat_rows < S) {
        g_nn_err_ext = "gvd_groupnorm_cl_bwd_apply: null pointer or stat_rows < S";
        return 2;
    }
    if (gn_bwd_check("gvd_groupnorm_cl_bwd_apply", C, groups, do_silu)) return 2;
    int nchunks = gn_bwd_chunks(F, S);
    const int rows_per_cta = (int)((S + nchunks - 1) / nchunks);
    nchunks = (int)((S + rows_per_cta - 1) / rows_per_cta);
    // the folded sums stand in for a one-chunk partial array
    launch_gn_bwd_apply(do_silu, dim3(nchunks, F), s, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, gamma, beta, stats,
                        sums, (int)S, C, groups, 1, rows_per_cta, eps, stat_rows);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_nn_err_ext = std::string("gvd_groupnorm_cl_bwd_apply: ") + cudaGetErrorString(e); return 1; }
    return 0;
}

int gvd_layernorm_bwd(const void* x, const void* dy, void* dx, const float* gamma, long long rows, int C, float eps,
                      gvd_nn_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (rows <= 0) return 0;
    if (!x || !dy || !dx || !gamma) { g_nn_err_ext = "gvd_layernorm_bwd: null pointer"; return 2; }
    if (C % 2) { g_nn_err_ext = "gvd_layernorm_bwd: C must be even"; return 2; }
    layernorm_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy,
                                                                  (__nv_bfloat16*)dx, gamma, rows, C, eps);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int gvd_geglu_bwd(const void* h, const void* dout, void* dh, long long rows, int D, gvd_nn_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (rows <= 0) return 0;
    if (!h || !dout || !dh) { g_nn_err_ext = "gvd_geglu_bwd: null pointer"; return 2; }
    if (D % 2) { g_nn_err_ext = "gvd_geglu_bwd: D must be even"; return 2; }
    geglu_bwd_kernel<<<grid_for(rows * (D / 2)), 256, 0, s>>>((const __nv_bfloat16*)h, (const __nv_bfloat16*)dout, (__nv_bfloat16*)dh, rows, D);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int gvd_softmax_bwd_rows(const void* p, const void* dp, void* ds, long long ld, long long rows, int cols, gvd_nn_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (rows <= 0) return 0;
    if (!p || !dp || !ds) { g_nn_err_ext = "gvd_softmax_bwd_rows: null pointer"; return 2; }
    if (cols <= 0 || cols > ld) { g_nn_err_ext = "gvd_softmax_bwd_rows: needs 0 < cols <= ld"; return 2; }
    softmax_bwd_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>((const __nv_bfloat16*)p, (const __nv_bfloat16*)dp, (__nv_bfloat16*)ds, ld,
                                                                     rows, cols);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int gvd_col2im3x3_cl(const void* dcol, void* dx, int F, int H, int W, int C, int stride, int upsample, gvd_nn_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (!dcol || !dx) { g_nn_err_ext = "gvd_col2im3x3_cl: null pointer"; return 2; }
    if (C % 8 || (stride != 1 && stride != 2) || (upsample && stride != 1)) {
        g_nn_err_ext = "gvd_col2im3x3_cl: needs C % 8 == 0, stride 1 or 2, stride 1 with upsample";
        return 2;
    }
    const int Hin = upsample ? 2 * H : H, Win = upsample ? 2 * W : W;
    const int Ho = (Hin + 2 - 3) / stride + 1, Wo = (Win + 2 - 3) / stride + 1;
    const long long total = (long long)F * H * W * (C / 8);
    if (total <= 0) return 0;
    col2im3x3_kernel<<<grid_for(total), 256, 0, s>>>((const __nv_bfloat16*)dcol, (__nv_bfloat16*)dx, F, H, W, C, Ho, Wo, stride, upsample ? 1 : 0);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int gvd_upsample2x_bwd_cl(const void* dy, void* dx, int F, int H, int W, int C, gvd_nn_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (!dy || !dx) { g_nn_err_ext = "gvd_upsample2x_bwd_cl: null pointer"; return 2; }
    if (C % 8) { g_nn_err_ext = "gvd_upsample2x_bwd_cl: C must be a multiple of 8"; return 2; }
    const long long total = (long long)F * H * W * (C / 8);
    if (total <= 0) return 0;
    upsample2x_bwd_kernel<<<grid_for(total), 256, 0, s>>>((const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, F, H, W, C);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int gvd_col2im_t3_cl(const void* dcol, void* dx, int B, int T, long long S, int C, gvd_nn_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (!dcol || !dx) { g_nn_err_ext = "gvd_col2im_t3_cl: null pointer"; return 2; }
    if (C % 8) { g_nn_err_ext = "gvd_col2im_t3_cl: C must be a multiple of 8"; return 2; }
    const long long total = (long long)B * T * S * (C / 8);
    if (total <= 0) return 0;
    col2im_t3_kernel<<<grid_for(total), 256, 0, s>>>((const __nv_bfloat16*)dcol, (__nv_bfloat16*)dx, B, T, S, C);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int gvd_temporal_attention_bwd(const void* q, const void* k, const void* v, const void* dout, void* dq, void* dk, void* dv, int B,
                               int T, long long S, int H, float scale, gvd_nn_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (!q || !k || !v || !dout || !dq || !dk || !dv) { g_nn_err_ext = "gvd_temporal_attention_bwd: null pointer"; return 2; }
    if (T > 32 || T <= 0) { g_nn_err_ext = "gvd_temporal_attention_bwd: needs 1 <= T <= 32"; return 2; }
    const long long warps = (long long)B * S * H;
    if (warps <= 0) return 0;
#ifndef GVD_HOST_EMU
    {  // tattn_mma.cu: the five products on mma.sync tiles (default; GVD_TATTN_MMA=0 selects the first kernel for A/B timing)
        static int on = -1;
        if (on < 0) {
            const char* e = getenv("GVD_TATTN_MMA");
            on = (e && e[0] == '0') ? 0 : 1;
        }
        if (on == 1 && gvd_mma_temporal_attention_bwd(q, k, v, dout, dq, dk, dv, B, T, S, H, scale, s))
            return cudaGetLastError() == cudaSuccess ? 0 : 1;
    }
#endif
    const int smem = (int)(TAB_WARPS * sizeof(TabSmem));
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(temporal_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) { g_nn_err_ext = std::string("gvd_temporal_attention_bwd attr: ") + cudaGetErrorString(e); return 1; }
        attr_set = true;
    }
    temporal_attn_bwd_kernel<<<(unsigned)((warps + TAB_WARPS - 1) / TAB_WARPS), TAB_WARPS * 32, smem, s>>>(
        (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, (const __nv_bfloat16*)dout, (__nv_bfloat16*)dq,
        (__nv_bfloat16*)dk, (__nv_bfloat16*)dv, B, T, S, H, scale);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int gvd_ddim_pred_x0_vjp(const GvdDdimVjpArgs* a, gvd_nn_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (!a || !a->e_cond || !a->grad_pred_x0 || !a->dx || !a->de_cond || !a->scratch) { g_nn_err_ext = "gvd_ddim_pred_x0_vjp: null pointer"; return 2; }
    if (a->e_uncond && !a->de_uncond) { g_nn_err_ext = "gvd_ddim_pred_x0_vjp: de_uncond missing"; return 2; }
    const long long n = a->n;
    if (n <= 0) return 0;
    if (reinterpret_cast<uintptr_t>(a->scratch) & 7) { g_nn_err_ext = "gvd_ddim_pred_x0_vjp: scratch must be 8-byte aligned"; return 2; }
    double* stats = reinterpret_cast<double*>(a->scratch);
    const float r = a->use_dynamic_rescale ? a->scale_prev / a->scale_t : 1.0f;
    const float gv = -r * a->sqrt_one_minus_alphas_cumprod_t, gx = r * a->sqrt_alphas_cumprod_t;
    // without an unconditional branch: mo = e_c, no rescale (ddim.py:222-232)
    const float* e_u = a->e_uncond ? a->e_uncond : a->e_cond;
    const float cfg = a->e_uncond ? a->cfg_scale : 1.0f;
    const float phi = a->e_uncond ? a->guidance_rescale : 0.f;
    float* de_u = a->e_uncond ? a->de_uncond : nullptr;
    if (a->scratch_bytes < 64) { g_nn_err_ext = "gvd_ddim_pred_x0_vjp: scratch too small (64 bytes)"; return 2; }
    cudaMemsetAsync(stats, 0, 8 * sizeof(double), s);
    ddim_vjp_reduce_kernel<<<grid_for(n, 256, 148 * 4), 256, 0, s>>>(a->e_cond, e_u, a->grad_pred_x0, n, cfg, gv, stats);
    ddim_vjp_apply_kernel<<<grid_for(n), 256, 0, s>>>(a->e_cond, e_u, a->grad_pred_x0, a->dx, a->de_cond, de_u, n, cfg, phi, gv, gx, stats);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_nn_err_ext = std::string("gvd_ddim_pred_x0_vjp: ") + cudaGetErrorString(e); return 1; }
    return 0;
}

}  // extern "C"
