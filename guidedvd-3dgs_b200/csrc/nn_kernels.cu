// nn_kernels.cu -- memory-bound layers of the U-Net denoiser on channels-last bf16 activations, and the fused
// DDIM update (include/gvd_nn.h).  Activations live in HBM as [frames, pixels, channels] bf16 (channels contiguous =
// K-major for the tensor-core GEMMs); statistics and softmax are computed in fp32 like the reference under autocast
// (GroupNormSpecific casts to float, lvdm/basics.py:76-78; softmax/LayerNorm are autocast-to-fp32 ops).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string>

#include "../../include/gvd_nn.h"

extern thread_local std::string g_nn_err_ext;

// nn_fast.cu: re-indexed variants. Measured on B200 (profiles/r02_first_hw_run.txt): GEGLU 722 -> 379 us, im2col 3x3
// 1110 -> 564 us, temporal im2col 168 -> 127 us at the C3 shapes (4.6-5.2 TB/s), so level 1 is the default; the
// temporal-attention variant measured 10 % SLOWER (583 -> 640 us) and stays behind level 2.
bool gvd_fast_geglu(const void* h, void* out, long long rows, int D, cudaStream_t s);
bool gvd_fast_im2col3x3(const void* x, void* col, int F, int H, int W, int C, int Ho, int Wo, int stride, int up, cudaStream_t s);
bool gvd_fast_im2col_t3(const void* x, void* col, int B, int T, long long S, int C, cudaStream_t s);
bool gvd_fast_temporal_attention(const void* q, const void* k, const void* v, void* out, int B, int T, long long S, int H, float scale,
                                 cudaStream_t s);
#ifndef GVD_HOST_EMU
// tattn_mma.cu: temporal attention on mma.sync tiles (default; GVD_TATTN_MMA=0 selects the first kernel for A/B timing)
bool gvd_mma_temporal_attention(const void* q, const void* k, const void* v, void* out, int B, int T, long long S, int H, float scale,
                                cudaStream_t s);
static bool tattn_mma_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("GVD_TATTN_MMA");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}
#endif
static int g_nn_fast = -1;  // -1: not decided yet (GVD_NN_FAST), 0 / 1 / 2 afterwards or through gvd_nn_set_fast
static int nn_fast_level() {
    if (g_nn_fast < 0) {
        const char* e = getenv("GVD_NN_FAST");
        g_nn_fast = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
    }
    return g_nn_fast;
}
static bool nn_fast_enabled() { return nn_fast_level() >= 1; }

namespace {

// x * sigmoid(x) with the approximate reciprocal (MUFU.RCP + FMUL, <= 2 ulp in fp32, far below the bf16 rounding that
// follows): the IEEE division it replaces was ~10 of the ~22 instructions per element of gn_apply_kernel.
__device__ __forceinline__ float silu(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
__device__ __forceinline__ float gelu(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------- GroupNorm (channels-last) ----------------
// x [F, S, C]; group g = channels [g*cpg, (g+1)*cpg); statistics over S x cpg per (frame, group).
// Thread layout: a thread owns one 16-byte vector of 8 channels and walks rows; the CTA covers
// (256 / (C/8)) rows per step, so every load is a full coalesced 16-byte access.
//
// Rows reach the thread through a thread-private ring of GN_RING 16-byte cp.async copies (no slot is shared, so
// cp.async.wait_group is the only ordering needed): the bytes in flight no longer depend on the register budget.  What the
// first version lost (profiles/r02_norm_bwd_bench_before.txt: 2.1-2.9 TB/s over the three passes) was (1) the merge of
// the per-thread sums -- 16 shared atomics per thread on 64 addresses, serialised, longer than the CTA's whole stream --
// (2) a 24-deep serial chain of L2 loads in front of every apply CTA (the fold of the chunk partials by 32 threads) and
// (3) 600 CTAs on 592 slots: two waves.
constexpr int GN_RING = 4;
constexpr int GN_COPIES = 8;  // shared copies of the per-group sums, picked by row slot
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

// Pass 1: grid (chunks, F): partial (sum, sumsq) per (frame, chunk, group).
__global__ void __launch_bounds__(256, 4) gn_partial_kernel(const __nv_bfloat16* __restrict__ x, int S, int C, int groups,
                                                            int rows_per_chunk, float* __restrict__ partial) {
    // Per-thread fp32 sums are merged in 2^-20 fixed point: integer adds commute, so the statistics (and with them the
    // whole denoiser) are bit-reproducible run to run, whatever order the atomics land in.
    extern __shared__ unsigned long long sh_fix[];  // [GN_COPIES][groups*2], then the ring
    unsigned long long* sh = sh_fix;
    uint4* ring = reinterpret_cast<uint4*>(sh_fix) + (GN_COPIES * groups * 2 * sizeof(unsigned long long) + 15) / 16;
    const int f = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
    const int cpg = C / groups, vecs = C / 8;
    for (int i = threadIdx.x; i < GN_COPIES * groups * 2; i += blockDim.x) sh[i] = 0ull;
    __syncthreads();
    const int r0 = chunk * rows_per_chunk, r1 = min(S, r0 + rows_per_chunk);
    // vecs <= 256: 256/vecs rows in flight per step; wider rows: one row per step, threads stride over the vectors
    const int vper = vecs <= 256 ? vecs : 256;
    const int rows_par = vecs <= 256 ? 256 / vecs : 1;
    const int rsub = threadIdx.x / vper;
    if (rsub < rows_par)
    for (int v = threadIdx.x % vper; v < vecs; v += vper) {
        float sm[8], sq[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) sm[e] = sq[e] = 0.f;
        const uint4* base = reinterpret_cast<const uint4*>(x + (size_t)f * S * C) + v;
        const int first = r0 + rsub;
        const int n = first < r1 ? (r1 - first + rows_par - 1) / rows_par : 0;
#pragma unroll
        for (int k = 0; k < GN_RING; ++k) {
            if (k < n) cp_async16(&ring[k * 256 + threadIdx.x], base + (size_t)(first + k * rows_par) * vecs);
            cp_async_commit();
        }
        for (int k = 0; k < n; ++k) {
            cp_async_wait<GN_RING - 1>();
            const int slot = k % GN_RING;
            const uint4 u = ring[slot * 256 + threadIdx.x];
            if (k + GN_RING < n) cp_async16(&ring[slot * 256 + threadIdx.x], base + (size_t)(first + (k + GN_RING) * rows_par) * vecs);
            cp_async_commit();
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 t = __bfloat1622float2(h2[e]);
                sm[2 * e] += t.x; sq[2 * e] += t.x * t.x;
                sm[2 * e + 1] += t.y; sq[2 * e + 1] += t.y * t.y;
            }
        }
        // the channels of one group are added up in the thread first -- as integers, so nothing about the result changes
        unsigned long long* my = sh + (rsub % GN_COPIES) * groups * 2;
        int gcur = (8 * v) / cpg;
        unsigned long long f1 = 0ull, f2 = 0ull;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int g = (8 * v + e) / cpg;
            if (g != gcur) {
                atomicAdd(&my[2 * gcur], f1);
                atomicAdd(&my[2 * gcur + 1], f2);
                f1 = f2 = 0ull;
                gcur = g;
            }
            f1 += (unsigned long long)__double2ll_rn((double)sm[e] * 1048576.0);
            f2 += (unsigned long long)__double2ll_rn((double)sq[e] * 1048576.0);
        }
        atomicAdd(&my[2 * gcur], f1);
        atomicAdd(&my[2 * gcur + 1], f2);
    }
    __syncthreads();
    float* out = partial + ((size_t)f * nchunks + chunk) * groups * 2;
    for (int i = threadIdx.x; i < groups * 2; i += blockDim.x) {
        unsigned long long t = 0ull;
#pragma unroll
        for (int cpy = 0; cpy < GN_COPIES; ++cpy) t += sh[cpy * groups * 2 + i];
        out[i] = (float)((double)(long long)t * (1.0 / 1048576.0));
    }
}

// The chunk partials of frame f folded by a whole 256-thread CTA: entry i = tid % (2 groups) of every (256 / (2 groups))-th
// chunk, then the parts in fixed order.  Afterwards fold[i], i < 2 groups, holds (sum, sumsq) interleaved per group.
// (Both callers use this one decomposition, so the fused and the split entry points see the same bits.)
__device__ __forceinline__ void gn_fold_frame(const float* __restrict__ partial, int nchunks, int groups, int f, double* fold) {
    const int width = groups * 2, parts = 256 / width;
    const int i = threadIdx.x % width, part = threadIdx.x / width;
    double t = 0.0;
    if (part < parts)
        for (int c = part; c < nchunks; c += parts) t += partial[((size_t)f * nchunks + c) * width + i];
    fold[threadIdx.x] = t;
    __syncthreads();
    double a = 0.0;
    if (threadIdx.x < width)
        for (int pp = 0; pp < parts; ++pp) a += fold[pp * width + threadIdx.x];
    __syncthreads();
    if (threadIdx.x < width) fold[threadIdx.x] = a;
    __syncthreads();
}

// Fold the per-chunk partials of one frame into (sum, sumsq) per group: the exchange unit when the rows of a group are
// spread over several GPUs.
__global__ void __launch_bounds__(256) gn_fold_kernel(const float* __restrict__ partial, float* __restrict__ stats, int nchunks,
                                                      int groups) {
    __shared__ double fold[256];
    const int f = blockIdx.x;
    gn_fold_frame(partial, nchunks, groups, f, fold);
    if (threadIdx.x < groups * 2) stats[(size_t)f * groups * 2 + threadIdx.x] = (float)fold[threadIdx.x];
}

// Pass 2: y = (x - mean) * rstd * gamma + beta, optional SiLU.
template <int SILU>
__global__ void __launch_bounds__(256, 4) gn_apply_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          const float* __restrict__ partial, int S, int C, int groups,
                                                          int nchunks, int rows_per_cta, float eps, long long stat_rows,
                                                          float* __restrict__ stats_out) {
    extern __shared__ double gn_shd[];  // fold[256] doubles, mean[groups], rstd[groups] floats, then the ring
    double* fold = gn_shd;
    float* sh = reinterpret_cast<float*>(fold + 256);
    uint4* ring = reinterpret_cast<uint4*>(gn_shd) + (256 * sizeof(double) + groups * 2 * sizeof(float) + 15) / 16;
    const int f = blockIdx.y;
    const int cpg = C / groups, vecs = C / 8;
    gn_fold_frame(partial, nchunks, groups, f, fold);
    // gvd_groupnorm_cl_keep_stats: the folded sums are also the backward's input -- what gn_fold_kernel would have written
    if (stats_out != nullptr && blockIdx.x == 0 && threadIdx.x < groups * 2) stats_out[(size_t)f * groups * 2 + threadIdx.x] = (float)fold[threadIdx.x];
    if (threadIdx.x < groups) {
        // the split entry points (gvd_groupnorm_cl_stats -> _apply) hand the folded sums over as floats: round here too, so
        // the fused call and the split one normalise with the same bits (the guided tape must not change the forward)
        const double s = (double)(float)fold[2 * threadIdx.x];
        const double q = (double)(float)fold[2 * threadIdx.x + 1];
        const double n = (double)stat_rows * cpg;  // rows behind the sums (> S when the sums were added up across shards)
        const double mean = s / n;
        const double var = fmax(q / n - mean * mean, 0.0);
        sh[threadIdx.x] = (float)mean;
        sh[groups + threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
    }
    __syncthreads();
    const int vper = vecs <= 256 ? vecs : 256;
    const int rows_par = vecs <= 256 ? 256 / vecs : 1;
    const int rsub = threadIdx.x / vper;
    if (rsub >= rows_par) return;
    for (int v = threadIdx.x % vper; v < vecs; v += vper) {
    // per-thread constants for its 8 channels: scale = rstd*gamma, shift = beta - mean*rstd*gamma
    float sc[8], sf[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int c = 8 * v + e, g = c / cpg;
        sc[e] = sh[groups + g] * gamma[c];
        sf[e] = beta[c] - sh[g] * sc[e];
    }
    const int r0 = blockIdx.x * rows_per_cta, r1 = min(S, r0 + rows_per_cta);
    const uint4* xin = reinterpret_cast<const uint4*>(x + (size_t)f * S * C) + v;
    uint4* yout = reinterpret_cast<uint4*>(y + (size_t)f * S * C) + v;
    auto norm8 = [&](uint4 u) {
        __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 t = __bfloat1622float2(h2[e]);
            float a = fmaf(t.x, sc[2 * e], sf[2 * e]), b = fmaf(t.y, sc[2 * e + 1], sf[2 * e + 1]);
            if (SILU == 1) {  // GroupNormSpecific casts its output to bf16 before nn.SiLU sees it (basics.py:76-78)
                a = silu(__bfloat162float(__float2bfloat16(a)));
                b = silu(__bfloat162float(__float2bfloat16(b)));
            } else if (SILU == 2) {  // plain nn.GroupNorm returns fp32 under autocast: SiLU in fp32, one rounding at the end
                a = silu(a);
                b = silu(b);
            }
            h2[e] = __floats2bfloat162_rn(a, b);
        }
        return u;
    };
    const int first = r0 + rsub;
    const int n = first < r1 ? (r1 - first + rows_par - 1) / rows_par : 0;
#pragma unroll
    for (int k = 0; k < GN_RING; ++k) {
        if (k < n) cp_async16(&ring[k * 256 + threadIdx.x], xin + (size_t)(first + k * rows_par) * vecs);
        cp_async_commit();
    }
    for (int k = 0; k < n; ++k) {
        cp_async_wait<GN_RING - 1>();
        const int slot = k % GN_RING;
        const uint4 u = ring[slot * 256 + threadIdx.x];
        if (k + GN_RING < n) cp_async16(&ring[slot * 256 + threadIdx.x], xin + (size_t)(first + (k + GN_RING) * rows_par) * vecs);
        cp_async_commit();
        yout[(size_t)(first + k * rows_par) * vecs] = norm8(u);
    }
    }
}

// ---------------- LayerNorm over the last dim (one warp per row) ----------------
__global__ void __launch_bounds__(256) layernorm_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        long long rows, int C, float eps) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const __nv_bfloat162* xin = reinterpret_cast<const __nv_bfloat162*>(x + row * C);
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
    __nv_bfloat162* yout = reinterpret_cast<__nv_bfloat162*>(y + row * C);
    for (int i = lane; i < pairs; i += 32) {
        const float2 v = __bfloat1622float2(xin[i]);
        yout[i] = __floats2bfloat162_rn((v.x - mean) * rstd * gamma[2 * i] + beta[2 * i],
                                        (v.y - mean) * rstd * gamma[2 * i + 1] + beta[2 * i + 1]);
    }
}

// The same normalisation with the row held in registers: one warp per row, NV 16-byte vectors per lane (C <= 256 * NV,
// C % 8 == 0), x read once and y written once with full-sector accesses (the kernel above reads the row three times
// with 4-byte accesses).  Same two-pass mean / variance arithmetic; only the order of the lane sums differs.
template <int NV>
__global__ void __launch_bounds__(256) layernorm_vec_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            long long rows, int C, float eps) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31, vecs = C >> 3;
    const uint4* xin = reinterpret_cast<const uint4*>(x + row * C);
    float v[NV][8];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int i = lane + 32 * k;
        uint4 u = make_uint4(0, 0, 0, 0);
        if (i < vecs) u = __ldg(xin + i);
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 t = __bfloat1622float2(h2[e]);
            v[k][2 * e] = t.x;
            v[k][2 * e + 1] = t.y;
            s += t.x + t.y;
        }
    }
    const float mean = warp_sum(s) / C;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k)
        if (lane + 32 * k < vecs) {
#pragma unroll
            for (int e = 0; e < 8; e += 2) q += (v[k][e] - mean) * (v[k][e] - mean) + (v[k][e + 1] - mean) * (v[k][e + 1] - mean);
        }
    const float rstd = rsqrtf(warp_sum(q) / C + eps);
    uint4* yout = reinterpret_cast<uint4*>(y + row * C);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int i = lane + 32 * k;
        if (i < vecs) {
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * i), g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * i + 1);
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * i), b1 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * i + 1);
            const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            uint4 u;
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e)
                h2[e] = __floats2bfloat162_rn((v[k][2 * e] - mean) * rstd * gg[2 * e] + bb[2 * e],
                                              (v[k][2 * e + 1] - mean) * rstd * gg[2 * e + 1] + bb[2 * e + 1]);
            yout[i] = u;
        }
    }
}

// ---------------- GEGLU: out[r, j] = h[r, j] * gelu(h[r, D + j]) ----------------
__global__ void __launch_bounds__(256) geglu_kernel(const __nv_bfloat16* __restrict__ h, __nv_bfloat16* __restrict__ out,
                                                    long long rows, int D) {
    const long long pairs = rows * (D / 2);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / (D / 2);
        const int j = (int)(i % (D / 2));
        const __nv_bfloat162* row = reinterpret_cast<const __nv_bfloat162*>(h + r * 2 * D);
        const float2 a = __bfloat1622float2(row[j]);
        const float2 g = __bfloat1622float2(row[D / 2 + j]);
        // F.gelu(gate) is materialised in bf16 before the product (attention.py:422-423)
        const float gx = __bfloat162float(__float2bfloat16(gelu(g.x))), gy = __bfloat162float(__float2bfloat16(gelu(g.y)));
        reinterpret_cast<__nv_bfloat162*>(out + r * D)[j] = __floats2bfloat162_rn(a.x * gx, a.y * gy);
    }
}

// ---------------- row softmax: fp32 scores -> bf16 probabilities (one warp per row) ----------------
__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename TIn>
__global__ void __launch_bounds__(256) softmax_rows_kernel(const TIn* __restrict__ x, long long ldx,
                                                           __nv_bfloat16* __restrict__ y, long long ldy, long long rows,
                                                           int cols) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const TIn* xr = x + row * ldx;
    float m = -INFINITY;
    for (int i = lane; i < cols; i += 32) m = fmaxf(m, ldf(xr + i));
    m = warp_max(m);
    float s = 0.f;
    for (int i = lane; i < cols; i += 32) s += __expf(ldf(xr + i) - m);
    const float inv = 1.0f / warp_sum(s);
    __nv_bfloat16* yr = y + row * ldy;
    for (int i = lane; i < cols; i += 32) yr[i] = __float2bfloat16(__expf(ldf(xr + i) - m) * inv);
    for (int i = cols + lane; i < ldy; i += 32) yr[i] = __float2bfloat16(0.f);  // zero the K padding
}

// ---------------- im2col for 3x3 convs on [F, H, W, C]; K order = (ky, kx, c) ----------------
// stride 1|2, pad 1; up=1 reads a nearest-neighbour 2x upsampled view of x (Upsample + conv fused).
__global__ void __launch_bounds__(256) im2col3x3_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ col,
                                                        int F, int H, int W, int C, int Ho, int Wo, int stride, int up) {
    const int vec = C / 8;  // 16-byte vectors per pixel
    const long long total = (long long)F * Ho * Wo * 9 * vec;
    const int Hin = up ? 2 * H : H, Win = up ? 2 * W : W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % vec);
        long long t = i / vec;
        const int tap = (int)(t % 9);
        t /= 9;
        const int ox = (int)(t % Wo);
        t /= Wo;
        const int oy = (int)(t % Ho);
        const int f = (int)(t / Ho);
        int iy = oy * stride + tap / 3 - 1, ix = ox * stride + tap % 3 - 1;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (iy >= 0 && iy < Hin && ix >= 0 && ix < Win) {
            if (up) { iy >>= 1; ix >>= 1; }
            val = reinterpret_cast<const uint4*>(x + (((size_t)f * H + iy) * W + ix) * C)[v];
        }
        reinterpret_cast<uint4*>(col)[i] = val;
    }
}

// ---------------- nearest-neighbour 2x upsampling of [F, H, W, C] -> [F, 2H, 2W, C] ----------------
// (Upsample.forward, openaimodel3d.py / ae_modules.py: F.interpolate(scale_factor=2, mode="nearest") in front of a 3x3
// convolution.)  Materialising the 4x tensor and running the implicit-GEMM convolution over it moves 4 units + the
// convolution's own reads; the fused im2col it replaces wrote and re-read a 36x matrix (3.8 GB for the VAE decoder's
// last Upsample at 5 x 320 x 512 x 256).  One thread reads one 16-byte vector and writes it four times.
__global__ void __launch_bounds__(256) upsample2x_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int F, int H,
                                                         int W, int C) {
    const int vec = C / 8;
    const long long total = (long long)F * H * W * vec;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % vec);
        long long t = i / vec;
        const int ix = (int)(t % W);
        t /= W;
        const int iy = (int)(t % H);
        const int f = (int)(t / H);
        const uint4 val = __ldg(reinterpret_cast<const uint4*>(x) + i);
        uint4* out = reinterpret_cast<uint4*>(y) + ((((size_t)f * 2 * H + 2 * iy) * 2 * W + 2 * ix) * vec + v);
        out[0] = val;
        out[vec] = val;
        out[(size_t)2 * W * vec] = val;
        out[(size_t)2 * W * vec + vec] = val;
    }
}

// ---------------- im2col for (3,1,1) temporal convs on [B, T, S, C]; K order = (kt, c) ----------------
__global__ void __launch_bounds__(256) im2col_t3_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ col,
                                                        int B, int T, long long S, int C) {
    const int vec = C / 8;
    const long long total = (long long)B * T * S * 3 * vec;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % vec);
        long long t = i / vec;
        const int tap = (int)(t % 3);
        t /= 3;
        const long long s = t % S;
        t /= S;
        const int tt = (int)(t % T);
        const int b = (int)(t / T);
        const int it = tt + tap - 1;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (it >= 0 && it < T) val = reinterpret_cast<const uint4*>(x + (((size_t)b * T + it) * S + s) * C)[v];
        reinterpret_cast<uint4*>(col)[i] = val;
    }
}

// ---------------- temporal self-attention: sequences of T <= 32 frames per (pixel, head), d = 64 ----------------
// q,k,v,out: [B, T, S, H*64] bf16. One warp per (b, s, h): K and V of the sequence (T x 64 each) are staged in shared
// memory with coalesced 128-byte row loads; lane i < T owns query frame i: its q row and output row live in registers,
// K/V rows are read as warp-wide broadcasts.
#define TA_WARPS 4
__global__ void __launch_bounds__(TA_WARPS * 32) temporal_attn_kernel(const __nv_bfloat16* __restrict__ q,
                                                                      const __nv_bfloat16* __restrict__ k,
                                                                      const __nv_bfloat16* __restrict__ v,
                                                                      __nv_bfloat16* __restrict__ out, int B, int T,
                                                                      long long S, int H, float scale) {
    __shared__ __align__(16) __nv_bfloat16 sk[TA_WARPS][32 * 64];
    __shared__ __align__(16) __nv_bfloat16 sv[TA_WARPS][32 * 64];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * TA_WARPS + wib;
    const long long total = (long long)B * S * H;
    if (w >= total) return;
    const int h = (int)(w % H);
    const long long s = (w / H) % S;
    const int b = (int)(w / (H * S));
    const long long tstride = S * H * 64;
    const size_t base = ((size_t)b * T * S + s) * H * 64 + (size_t)h * 64;
    // stage K, V: 8 lanes x 16 B cover one 128-byte row; 4 rows per pass
    for (int r = lane >> 3; r < T; r += 4) {
        const int c = lane & 7;
        reinterpret_cast<uint4*>(&sk[wib][r * 64])[c] = __ldg(reinterpret_cast<const uint4*>(k + base + r * tstride) + c);
        reinterpret_cast<uint4*>(&sv[wib][r * 64])[c] = __ldg(reinterpret_cast<const uint4*>(v + base + r * tstride) + c);
    }
    __syncwarp();
    if (lane >= T) return;
    float qf[64];
    {
        const uint4* qp = reinterpret_cast<const uint4*>(q + base + lane * tstride);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint4 u = __ldg(qp + c);
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 t = __bfloat1622float2(h2[e]);
                qf[c * 8 + 2 * e] = t.x;
                qf[c * 8 + 2 * e + 1] = t.y;
            }
        }
    }
    float sc[32];
    float m = -INFINITY;
#pragma unroll 1
    for (int j = 0; j < T; ++j) {
        const uint4* kp = reinterpret_cast<const uint4*>(&sk[wib][j * 64]);
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint4 u = kp[c];
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 t = __bfloat1622float2(h2[e]);
                dot = fmaf(qf[c * 8 + 2 * e], t.x, dot);
                dot = fmaf(qf[c * 8 + 2 * e + 1], t.y, dot);
            }
        }
        // the reference rounds the einsum output to bf16 and scales in bf16 before the fp32 softmax (attention.py:103)
        const float x = __bfloat162float(__float2bfloat16(__bfloat162float(__float2bfloat16(dot)) * scale));
        sc[j] = x;
        m = fmaxf(m, x);
    }
    float l = 0.f;
#pragma unroll 1
    for (int j = 0; j < T; ++j) {
        sc[j] = __expf(sc[j] - m);
        l += sc[j];
    }
    const float inv = 1.0f / l;
    float o[64];
#pragma unroll
    for (int c = 0; c < 64; ++c) o[c] = 0.f;
#pragma unroll 1
    for (int j = 0; j < T; ++j) {
        const float p = __bfloat162float(__float2bfloat16(sc[j] * inv));  // softmax output enters the PV einsum as bf16
        const uint4* vp = reinterpret_cast<const uint4*>(&sv[wib][j * 64]);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint4 u = vp[c];
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 t = __bfloat1622float2(h2[e]);
                o[c * 8 + 2 * e] = fmaf(p, t.x, o[c * 8 + 2 * e]);
                o[c * 8 + 2 * e + 1] = fmaf(p, t.y, o[c * 8 + 2 * e + 1]);
            }
        }
    }
    uint4* op = reinterpret_cast<uint4*>(out + base + lane * tstride);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        uint4 u;
        __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(o[c * 8 + 2 * e], o[c * 8 + 2 * e + 1]);
        op[c] = u;
    }
}

// ---------------- fused DDIM step (ddim.py:206-280, v-prediction, CFG + rescale, dynamic rescale) ----------------
// pass 1: model_output = e_u + s (e_c - e_u); accumulate sum / sumsq of e_c and of model_output (for the two stds)
__global__ void __launch_bounds__(256) ddim_cfg_kernel(const float* __restrict__ e_c, const float* __restrict__ e_u,
                                                       float* __restrict__ mo, long long n, float cfg,
                                                       double* __restrict__ stats) {
    double s_c = 0, q_c = 0, s_m = 0, q_m = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float c = e_c[i], u = e_u[i];
        const float m = u + cfg * (c - u);
        mo[i] = m;
        s_c += c; q_c += (double)c * c;
        s_m += m; q_m += (double)m * m;
    }
    __shared__ double sh[4][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s_c += __shfl_xor_sync(0xffffffffu, s_c, o); q_c += __shfl_xor_sync(0xffffffffu, q_c, o);
        s_m += __shfl_xor_sync(0xffffffffu, s_m, o); q_m += __shfl_xor_sync(0xffffffffu, q_m, o);
    }
    if (lane == 0) { sh[0][warp] = s_c; sh[1][warp] = q_c; sh[2][warp] = s_m; sh[3][warp] = q_m; }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0;
        for (int w2 = 0; w2 < 8; ++w2) t += sh[threadIdx.x][w2];
        atomicAdd(&stats[threadIdx.x], t);
    }
}

struct DdimCoef {
    float guidance_rescale, sqrt_ac, sqrt_1mac, a_prev_sqrt, dir_coef, sigma, scale_ratio;
    int use_rescale;
};

__global__ void __launch_bounds__(256) ddim_update_kernel(const float* __restrict__ x, const float* __restrict__ mo,
                                                          const float* __restrict__ noise, float* __restrict__ x_prev,
                                                          float* __restrict__ pred_x0, long long n, DdimCoef c,
                                                          const double* __restrict__ stats) {
    float factor = 1.0f;
    if (c.guidance_rescale > 0.f) {
        // torch.std default: unbiased (N-1)
        const double nn = (double)n;
        const double var_c = (stats[1] - stats[0] * stats[0] / nn) / (nn - 1.0);
        const double var_m = (stats[3] - stats[2] * stats[2] / nn) / (nn - 1.0);
        const float ratio = (float)sqrt(var_c) / (float)sqrt(var_m);
        factor = c.guidance_rescale * ratio + (1.0f - c.guidance_rescale);  // gr*(m*ratio) + (1-gr)*m
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float xt = x[i];
        float v = mo[i];
        if (c.guidance_rescale > 0.f) {
            const float r = v * (factor - (1.0f - c.guidance_rescale)) / c.guidance_rescale;  // v * ratio
            v = c.guidance_rescale * r + (1.0f - c.guidance_rescale) * v;
        }
        const float e_t = c.sqrt_ac * v + c.sqrt_1mac * xt;        // predict_eps_from_z_and_v (ddpm3d.py:247-251)
        float p0 = c.sqrt_ac * xt - c.sqrt_1mac * v;                // predict_start_from_z_and_v (ddpm3d.py:239-245)
        if (c.use_rescale) p0 *= c.scale_ratio;
        pred_x0[i] = p0;
        x_prev[i] = c.a_prev_sqrt * p0 + c.dir_coef * e_t + c.sigma * noise[i];
    }
}

// Frames per launch pair: both passes of a GroupNorm read x, so a group of frames whose tensors fit the L2 (126 MB) is
// normalised before the next group is touched -- the second pass then finds its input on chip instead of in HBM.
// `units` = tensors of S x C bf16 a frame moves through the two passes (forward: x, y; backward: x, dy, dx).
// GVD_GN_GROUP_MB bounds units x frames x S x C x 2 bytes; DEFAULT 0 = one group, because it does not pay: graph-timed on B200
// (profiles/r02_gn_groups.txt) 25 x 2560 x 320 runs forward / backward in 34 / 53 us ungrouped and 45 / 73 us with 48 MB groups
// (71 / 99 us with 24 MB) -- the extra launches' ramps and tails cost more than the L2 hits return.  Kept as an A/B knob.
int gn_group_frames(int F, long long S, int C, int units) {
    static long long budget = -1;
    if (budget < 0) {
        const char* e = getenv("GVD_GN_GROUP_MB");
        budget = (e ? atoll(e) : 0) << 20;
    }
    if (budget <= 0 || F <= 1) return F;
    const long long per_frame = (long long)units * S * C * 2;
    long long fmax = budget / (per_frame > 0 ? per_frame : 1);
    if (fmax < 1) fmax = 1;
    if (fmax >= F) return F;
    const long long ngroups = (F + fmax - 1) / fmax;
    return (int)((F + ngroups - 1) / ngroups);
}

int gn_chunks(int F, long long S) {
    long long want = 592 / (F > 0 ? F : 1);        // one wave: the kernels are resident four per SM (148 x 4 slots), never more CTAs than slots
    long long maxc = (S + 15) / 16;                // at least 16 rows per CTA
    long long minc = (S + 4095) / 4096;            // at most 4096 rows per CTA
    long long c = want < minc ? minc : want;
    if (c > maxc) c = maxc;
    if (c < 1) c = 1;
    if (c > 2048) c = 2048;
    return (int)c;
}

inline size_t gn_partial_smem(int groups) {
    return (GN_COPIES * groups * 2 * sizeof(unsigned long long) + 15) / 16 * 16 + (size_t)GN_RING * 256 * 16;
}
void launch_gn_apply(int do_silu, dim3 grid, cudaStream_t s, const __nv_bfloat16* x, __nv_bfloat16* y, const float* gamma, const float* beta,
                     const float* partial, int S, int C, int groups, int nchunks, int rows_per_cta, float eps, long long stat_rows,
                     float* stats_out = nullptr) {
    const size_t sm = (256 * sizeof(double) + groups * 2 * sizeof(float) + 15) / 16 * 16 + (size_t)GN_RING * 256 * 16;
    if (do_silu == 1) gn_apply_kernel<1><<<grid, 256, sm, s>>>(x, y, gamma, beta, partial, S, C, groups, nchunks, rows_per_cta, eps, stat_rows, stats_out);
    else if (do_silu == 2) gn_apply_kernel<2><<<grid, 256, sm, s>>>(x, y, gamma, beta, partial, S, C, groups, nchunks, rows_per_cta, eps, stat_rows, stats_out);
    else gn_apply_kernel<0><<<grid, 256, sm, s>>>(x, y, gamma, beta, partial, S, C, groups, nchunks, rows_per_cta, eps, stat_rows, stats_out);
}

int grid_for(long long n, int block = 256, int cap = 148 * 16) {
    long long g = (n + block - 1) / block;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" {

int gvd_nn_set_fast(int on) {
    const int was = nn_fast_level();
    if (on >= 0 && on <= 2) g_nn_fast = on;
    return was;
}

static int groupnorm_fused(const void* x, void* y, const float* gamma, const float* beta, float* stats_out, int F, long long S, int C,
                           int groups, float eps, int do_silu, float* tmp, size_t tmp_floats, gvd_nn_stream_t stream_);

int gvd_groupnorm_cl(const void* x, void* y, const float* gamma, const float* beta, int F, long long S, int C, int groups,
                     float eps, int do_silu, float* tmp, size_t tmp_floats, gvd_nn_stream_t stream_) {
    return groupnorm_fused(x, y, gamma, beta, nullptr, F, S, C, groups, eps, do_silu, tmp, tmp_floats, stream_);
}

int gvd_groupnorm_cl_keep_stats(const void* x, void* y, const float* gamma, const float* beta, float* stats, int F, long long S, int C,
                                int groups, float eps, int do_silu, float* tmp, size_t tmp_floats, gvd_nn_stream_t stream_) {
    if (!stats) { g_nn_err_ext = "gvd_groupnorm_cl_keep_stats: null stats"; return 2; }
    return groupnorm_fused(x, y, gamma, beta, stats, F, S, C, groups, eps, do_silu, tmp, tmp_floats, stream_);
}

static int groupnorm_fused(const void* x, void* y, const float* gamma, const float* beta, float* stats_out, int F, long long S, int C,
                           int groups, float eps, int do_silu, float* tmp, size_t tmp_floats, gvd_nn_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (F <= 0 || S <= 0) return 0;
    if (C % groups != 0 || C % 8 != 0 || groups > 128) {
        g_nn_err_ext = "gvd_groupnorm_cl: needs C % groups == 0, C % 8 == 0, groups <= 128";
        return 2;
    }
    const int fg = gn_group_frames(F, S, C, 2);
    for (int f0 = 0; f0 < F; f0 += fg) {
        const int Fg = F - f0 < fg ? F - f0 : fg;
        int nchunks = gn_chunks(Fg, S);  // one wave of CTAs per launch, >= 16 rows per CTA
        const int rows_per_chunk = (int)((S + nchunks - 1) / nchunks);
        nchunks = (int)((S + rows_per_chunk - 1) / rows_per_chunk);
        if (tmp_floats < (size_t)Fg * nchunks * groups * 2) { g_nn_err_ext = "gvd_groupnorm_cl: scratch too small"; return 2; }
        const __nv_bfloat16* xg = (const __nv_bfloat16*)x + (size_t)f0 * S * C;
        gn_partial_kernel<<<dim3(nchunks, Fg), 256, gn_partial_smem(groups), s>>>(xg, (int)S, C, groups, rows_per_chunk, tmp);
        launch_gn_apply(do_silu, dim3((unsigned)nchunks, Fg), s, xg, (__nv_bfloat16*)y + (size_t)f0 * S * C, gamma, beta, tmp, (int)S, C, groups,
                        nchunks, rows_per_chunk, eps, S, stats_out ? stats_out + (size_t)f0 * groups * 2 : nullptr);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_nn_err_ext = std::string("gvd_groupnorm_cl: ") + cudaGetErrorString(e); return 1; }
    return 0;
}

int gvd_groupnorm_cl_stats(const void* x, float* stats, int F, long long S, int C, int groups, float* tmp, size_t tmp_floats,
                           gvd_nn_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (F <= 0) return 0;
    if (C % groups != 0 || C % 8 != 0 || groups > 128) {
        g_nn_err_ext = "gvd_groupnorm_cl_stats: needs C % groups == 0, C % 8 == 0, groups <= 128";
        return 2;
    }
    if (S <= 0) return cudaMemsetAsync(stats, 0, sizeof(float) * F * groups * 2, s) == cudaSuccess ? 0 : 1;
    int nchunks = gn_chunks(F, S);
    const int rows_per_chunk = (int)((S + nchunks - 1) / nchunks);
    nchunks = (int)((S + rows_per_chunk - 1) / rows_per_chunk);
    if (tmp_floats < (size_t)F * nchunks * groups * 2) { g_nn_err_ext = "gvd_groupnorm_cl_stats: scratch too small"; return 2; }
    gn_partial_kernel<<<dim3(nchunks, F), 256, gn_partial_smem(groups), s>>>((const __nv_bfloat16*)x, (int)S, C, groups, rows_per_chunk, tmp);
    gn_fold_kernel<<<F, 256, 0, s>>>(tmp, stats, nchunks, groups);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_nn_err_ext = std::string("gvd_groupnorm_cl_stats: ") + cudaGetErrorString(e); return 1; }
    return 0;
}

int gvd_groupnorm_cl_apply(const void* x, void* y, const float* gamma, const float* beta, const float* stats, int F, long long S,
                           long long stat_rows, int C, int groups, float eps, int do_silu, gvd_nn_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (F <= 0 || S <= 0) return 0;
    if (C % groups != 0 || C % 8 != 0 || groups > 128 || stat_rows < S) {
        g_nn_err_ext = "gvd_groupnorm_cl_apply: needs C % groups == 0, C % 8 == 0, groups <= 128, stat_rows >= S";
        return 2;
    }
    int nchunks = gn_chunks(F, S);
    const int rows_per_cta = (int)((S + nchunks - 1) / nchunks);
    nchunks = (int)((S + rows_per_cta - 1) / rows_per_cta);
    launch_gn_apply(do_silu, dim3((unsigned)nchunks, F), s, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, gamma, beta, stats, (int)S, C, groups, 1,
                    rows_per_cta, eps, stat_rows);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_nn_err_ext = std::string("gvd_groupnorm_cl_apply: ") + cudaGetErrorString(e); return 1; }
    return 0;
}

size_t gvd_groupnorm_tmp_floats(int F, long long S, int groups) {
    // the fused call may run the frames in groups (gn_group_frames), each with its own chunking: the largest of them all
    size_t most = (size_t)F * (gn_chunks(F, S) + 1);
    for (int fg = 1; fg < F; ++fg) {
        const size_t n = (size_t)fg * (gn_chunks(fg, S) + 1);
        if (n > most) most = n;
    }
    return most * groups * 2;
}

int gvd_layernorm(const void* x, void* y, const float* gamma, const float* beta, long long rows, int C, float eps,
                  gvd_nn_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (rows <= 0) return 0;
    if (C % 2) { g_nn_err_ext = "gvd_layernorm: C must be even"; return 2; }
    const bool vec_ok = nn_fast_enabled() && (C % 8 == 0) && C <= 256 * 6 &&
                        ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(gamma) |
                          reinterpret_cast<uintptr_t>(beta)) & 15) == 0;
    const unsigned grid = (unsigned)((rows + 7) / 8);
    const __nv_bfloat16* xb = (const __nv_bfloat16*)x;
    __nv_bfloat16* yb = (__nv_bfloat16*)y;
    if (vec_ok && C <= 512) layernorm_vec_kernel<2><<<grid, 256, 0, s>>>(xb, yb, gamma, beta, rows, C, eps);
    else if (vec_ok && C <= 768) layernorm_vec_kernel<3><<<grid, 256, 0, s>>>(xb, yb, gamma, beta, rows, C, eps);
    else if (vec_ok && C <= 1280) layernorm_vec_kernel<5><<<grid, 256, 0, s>>>(xb, yb, gamma, beta, rows, C, eps);
    else if (vec_ok) layernorm_vec_kernel<6><<<grid, 256, 0, s>>>(xb, yb, gamma, beta, rows, C, eps);
    else
    layernorm_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, gamma, beta, rows, C, eps);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int gvd_geglu(const void* h, void* out, long long rows, int D, gvd_nn_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (rows <= 0) return 0;
    if (D % 2) { g_nn_err_ext = "gvd_geglu: D must be even"; return 2; }
    if (nn_fast_enabled() && gvd_fast_geglu(h, out, rows, D, s)) return cudaGetLastError() == cudaSuccess ? 0 : 1;
    geglu_kernel<<<grid_for(rows * (D / 2)), 256, 0, s>>>((const __nv_bfloat16*)h, (__nv_bfloat16*)out, rows, D);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int gvd_softmax_rows(const void* x, int x_is_bf16, long long ldx, void* y, long long ldy, long long rows, int cols,
                     gvd_nn_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (rows <= 0) return 0;
    if (x_is_bf16)
        softmax_rows_kernel<__nv_bfloat16><<<(unsigned)((rows + 7) / 8), 256, 0, s>>>((const __nv_bfloat16*)x, ldx, (__nv_bfloat16*)y, ldy, rows, cols);
    else
        softmax_rows_kernel<float><<<(unsigned)((rows + 7) / 8), 256, 0, s>>>((const float*)x, ldx, (__nv_bfloat16*)y, ldy, rows, cols);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int gvd_im2col3x3_cl(const void* x, void* col, int F, int H, int W, int C, int stride, int upsample, gvd_nn_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (C % 8) { g_nn_err_ext = "gvd_im2col3x3_cl: C must be a multiple of 8"; return 2; }
    const int Hin = upsample ? 2 * H : H, Win = upsample ? 2 * W : W;
    const int Ho = (Hin + 2 - 3) / stride + 1, Wo = (Win + 2 - 3) / stride + 1;
    const long long total = (long long)F * Ho * Wo * 9 * (C / 8);
    if (total <= 0) return 0;
    if (nn_fast_enabled() && gvd_fast_im2col3x3(x, col, F, H, W, C, Ho, Wo, stride, upsample, s)) return cudaGetLastError() == cudaSuccess ? 0 : 1;
    im2col3x3_kernel<<<grid_for(total), 256, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)col, F, H, W, C, Ho, Wo, stride, upsample);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int gvd_upsample2x_cl(const void* x, void* y, int F, int H, int W, int C, gvd_nn_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (!x || !y) { g_nn_err_ext = "gvd_upsample2x_cl: null pointer"; return 2; }
    if (C % 8) { g_nn_err_ext = "gvd_upsample2x_cl: C must be a multiple of 8"; return 2; }
    const long long total = (long long)F * H * W * (C / 8);
    if (total <= 0) return 0;
    upsample2x_kernel<<<grid_for(total), 256, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, F, H, W, C);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int gvd_im2col_t3_cl(const void* x, void* col, int B, int T, long long S, int C, gvd_nn_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (C % 8) { g_nn_err_ext = "gvd_im2col_t3_cl: C must be a multiple of 8"; return 2; }
    const long long total = (long long)B * T * S * 3 * (C / 8);
    if (total <= 0) return 0;
    if (nn_fast_enabled() && gvd_fast_im2col_t3(x, col, B, T, S, C, s)) return cudaGetLastError() == cudaSuccess ? 0 : 1;
    im2col_t3_kernel<<<grid_for(total), 256, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)col, B, T, S, C);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int gvd_temporal_attention(const void* q, const void* k, const void* v, void* out, int B, int T, long long S, int H,
                           float scale, gvd_nn_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (T > 32 || T <= 0) { g_nn_err_ext = "gvd_temporal_attention: needs 1 <= T <= 32"; return 2; }
    const long long warps = (long long)B * S * H;
    if (warps <= 0) return 0;
#ifndef GVD_HOST_EMU
    if (nn_fast_level() == 1 && tattn_mma_enabled() && gvd_mma_temporal_attention(q, k, v, out, B, T, S, H, scale, s))
        return cudaGetLastError() == cudaSuccess ? 0 : 1;
#endif
    if (nn_fast_level() >= 2 && gvd_fast_temporal_attention(q, k, v, out, B, T, S, H, scale, s)) return cudaGetLastError() == cudaSuccess ? 0 : 1;
    temporal_attn_kernel<<<(unsigned)((warps + TA_WARPS - 1) / TA_WARPS), TA_WARPS * 32, 0, s>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)k,
                                                                    (const __nv_bfloat16*)v, (__nv_bfloat16*)out, B, T, S, H, scale);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int gvd_ddim_step(const GvdDdimArgs* a, gvd_nn_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (!a || !a->x || !a->e_cond || !a->x_prev || !a->pred_x0 || !a->noise || !a->scratch) { g_nn_err_ext = "gvd_ddim_step: null pointer"; return 2; }
    const long long n = a->n;
    if (n <= 0) return 0;
    double* stats = reinterpret_cast<double*>(a->scratch);
    float* mo = reinterpret_cast<float*>(stats + 4);
    cudaMemsetAsync(stats, 0, 4 * sizeof(double), s);
    const float* e_u = a->e_uncond ? a->e_uncond : a->e_cond;
    const float cfg = a->e_uncond ? a->cfg_scale : 1.0f;
    ddim_cfg_kernel<<<grid_for(n, 256, 148 * 4), 256, 0, s>>>(a->e_cond, e_u, mo, n, cfg, stats);
    DdimCoef c;
    c.guidance_rescale = a->e_uncond ? a->guidance_rescale : 0.f;
    c.sqrt_ac = a->sqrt_alphas_cumprod_t;
    c.sqrt_1mac = a->sqrt_one_minus_alphas_cumprod_t;
    c.a_prev_sqrt = sqrtf(a->ddim_alpha_prev);
    c.dir_coef = sqrtf(1.0f - a->ddim_alpha_prev - a->ddim_sigma * a->ddim_sigma);
    c.sigma = a->ddim_sigma * a->temperature;
    c.scale_ratio = a->scale_prev / a->scale_t;
    c.use_rescale = a->use_dynamic_rescale;
    ddim_update_kernel<<<grid_for(n), 256, 0, s>>>(a->x, mo, a->noise, a->x_prev, a->pred_x0, n, c, stats);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // extern "C"
