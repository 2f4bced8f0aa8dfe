// nn_fast.cu -- variants of four memory-bound kernels of the denoiser (GEGLU, the 3x3 and temporal im2col: re-indexed;
// temporal attention: K / V staged in fp32),
// selected at run time with GVD_NN_FAST=1 (default off: they have been executed on the host emulation, bit-identical
// to the kernels they replace, but not yet timed on a GPU).
//
// Why: profiles/r01_unet_kernel_breakdown.txt puts GEGLU at 8.4 % and the two im2col kernels at 13 % of a C3 forward --
// 14 ms and 22 ms for ~33 GB and ~60 GB of algorithmic traffic, i.e. 35-40 % of the measured HBM rate.  Both kernels
// index with 64-bit divisions per element (GEGLU: one div + mod per 4-byte pair; im2col: four per 16-byte vector);
// nothing else in them is expensive.  Here: 16-byte vectors for GEGLU (8 outputs per div), 32-bit index arithmetic
// whenever the element count allows (it does for every layer of the 576x1024 network), same arithmetic per element.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__device__ __forceinline__ float gelu(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// out[r, j] = h[r, j] * bf16(gelu(h[r, D + j])); one thread = 8 consecutive j of one row (two 16-byte loads, one store)
template <typename I>
__global__ void __launch_bounds__(256) geglu_vec_kernel(const __nv_bfloat16* __restrict__ h, __nv_bfloat16* __restrict__ out, I total,
                                                        I vecs, int D) {
    for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
        const I r = i / vecs, v = i - r * vecs;
        const uint4* row = reinterpret_cast<const uint4*>(h + (size_t)r * 2 * D);
        const uint4 ua = __ldg(row + v), ug = __ldg(row + vecs + v);
        const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&ua);
        const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&ug);
        uint4 uo;
        __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&uo);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 a = __bfloat1622float2(a2[e]), g = __bfloat1622float2(g2[e]);
            // F.gelu(gate) is materialised in bf16 before the product (attention.py:422-423)
            const float gx = __bfloat162float(__float2bfloat16(gelu(g.x))), gy = __bfloat162float(__float2bfloat16(gelu(g.y)));
            o2[e] = __floats2bfloat162_rn(a.x * gx, a.y * gy);
        }
        reinterpret_cast<uint4*>(out + (size_t)r * D)[v] = uo;
    }
}

template <typename I>
__global__ void __launch_bounds__(256) im2col3x3_idx_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ col, I total,
                                                            int H, int W, int C, int Ho, int Wo, int stride, int up) {
    const I vec = (I)(C / 8);
    const int Hin = up ? 2 * H : H, Win = up ? 2 * W : W;
    for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
        I t = i / vec;
        const I v = i - t * vec;
        I q = t / 9;
        const int tap = (int)(t - q * 9);
        t = q / (I)Wo;
        const int ox = (int)(q - t * (I)Wo);
        const I f = t / (I)Ho;
        const int oy = (int)(t - f * (I)Ho);
        int iy = oy * stride + tap / 3 - 1, ix = ox * stride + tap % 3 - 1;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (iy >= 0 && iy < Hin && ix >= 0 && ix < Win) {
            if (up) { iy >>= 1; ix >>= 1; }
            val = reinterpret_cast<const uint4*>(x + (((size_t)f * H + iy) * W + ix) * C)[v];
        }
        reinterpret_cast<uint4*>(col)[i] = val;
    }
}

template <typename I>
__global__ void __launch_bounds__(256) im2col_t3_idx_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ col, I total,
                                                            int T, I S, int C) {
    const I vec = (I)(C / 8);
    for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
        I t = i / vec;
        const I v = i - t * vec;
        I q = t / 3;
        const int tap = (int)(t - q * 3);
        t = q / S;
        const I s = q - t * S;
        const I b = t / (I)T;
        const int tt = (int)(t - b * (I)T);
        const int it = tt + tap - 1;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (it >= 0 && it < T) val = reinterpret_cast<const uint4*>(x + (((size_t)b * T + it) * S + s) * C)[v];
        reinterpret_cast<uint4*>(col)[i] = val;
    }
}

// Temporal self-attention (nn_kernels.cu::temporal_attn_kernel) with K and V of the sequence converted to fp32 ONCE while
// they are staged in shared memory: the 25 query lanes of the warp then read them as float4 broadcasts instead of each
// converting the same bf16 values again (3 200 conversions per lane in the original, about as many as its FMAs).  The
// per-lane arithmetic -- order of the FMAs, rounding points -- is unchanged, so the results are bit-identical.
#define TAF_WARPS 4
__global__ void __launch_bounds__(TAF_WARPS * 32) temporal_attn_f32stage_kernel(const __nv_bfloat16* __restrict__ q,
                                                                                const __nv_bfloat16* __restrict__ k,
                                                                                const __nv_bfloat16* __restrict__ v,
                                                                                __nv_bfloat16* __restrict__ out, int B, int T,
                                                                                long long S, int H, float scale) {
    extern __shared__ __align__(16) float taf_smem[];  // per warp: K[32][64] then V[32][64] in fp32
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* sk = taf_smem + (size_t)wib * 2 * 32 * 64;
    float* sv = sk + 32 * 64;
    const long long w = (long long)blockIdx.x * TAF_WARPS + wib;
    const long long total = (long long)B * S * H;
    if (w >= total) return;
    const int h = (int)(w % H);
    const long long s = (w / H) % S;
    const int b = (int)(w / (H * S));
    const long long tstride = S * H * 64;
    const size_t base = ((size_t)b * T * S + s) * H * 64 + (size_t)h * 64;
    for (int r = lane >> 3; r < T; r += 4) {  // 8 lanes x 16 B cover one 128-byte row; 4 rows per pass
        const int c = lane & 7;
        const uint4 uk = __ldg(reinterpret_cast<const uint4*>(k + base + r * tstride) + c);
        const uint4 uv = __ldg(reinterpret_cast<const uint4*>(v + base + r * tstride) + c);
        const __nv_bfloat162* k2 = reinterpret_cast<const __nv_bfloat162*>(&uk);
        const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&uv);
        float4* dk = reinterpret_cast<float4*>(sk + r * 64 + c * 8);
        float4* dv = reinterpret_cast<float4*>(sv + r * 64 + c * 8);
        const float2 k0 = __bfloat1622float2(k2[0]), k1 = __bfloat1622float2(k2[1]), k2f = __bfloat1622float2(k2[2]), k3 = __bfloat1622float2(k2[3]);
        const float2 v0 = __bfloat1622float2(v2[0]), v1 = __bfloat1622float2(v2[1]), v2f = __bfloat1622float2(v2[2]), v3 = __bfloat1622float2(v2[3]);
        dk[0] = make_float4(k0.x, k0.y, k1.x, k1.y);
        dk[1] = make_float4(k2f.x, k2f.y, k3.x, k3.y);
        dv[0] = make_float4(v0.x, v0.y, v1.x, v1.y);
        dv[1] = make_float4(v2f.x, v2f.y, v3.x, v3.y);
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
        const float4* kp = reinterpret_cast<const float4*>(sk + j * 64);
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const float4 t = kp[c];
            dot = fmaf(qf[4 * c], t.x, dot);
            dot = fmaf(qf[4 * c + 1], t.y, dot);
            dot = fmaf(qf[4 * c + 2], t.z, dot);
            dot = fmaf(qf[4 * c + 3], t.w, dot);
        }
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
        const float p = __bfloat162float(__float2bfloat16(sc[j] * inv));
        const float4* vp = reinterpret_cast<const float4*>(sv + j * 64);
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const float4 t = vp[c];
            o[4 * c] = fmaf(p, t.x, o[4 * c]);
            o[4 * c + 1] = fmaf(p, t.y, o[4 * c + 1]);
            o[4 * c + 2] = fmaf(p, t.z, o[4 * c + 2]);
            o[4 * c + 3] = fmaf(p, t.w, o[4 * c + 3]);
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

int fast_grid(long long n) {
    long long g = (n + 255) / 256;
    const long long cap = 148 * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

// launchers used by nn_kernels.cu's entry points when GVD_NN_FAST=1; return false when the variant does not apply
bool gvd_fast_geglu(const void* h, void* out, long long rows, int D, cudaStream_t s) {
    if (D % 8 || (reinterpret_cast<uintptr_t>(h) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return false;
    const long long vecs = D / 8, total = rows * vecs;
    if (total < (1ll << 31))
        geglu_vec_kernel<uint32_t><<<fast_grid(total), 256, 0, s>>>((const __nv_bfloat16*)h, (__nv_bfloat16*)out, (uint32_t)total, (uint32_t)vecs, D);
    else
        geglu_vec_kernel<unsigned long long><<<fast_grid(total), 256, 0, s>>>((const __nv_bfloat16*)h, (__nv_bfloat16*)out, (unsigned long long)total,
                                                                            (unsigned long long)vecs, D);
    return true;
}

bool gvd_fast_im2col3x3(const void* x, void* col, int F, int H, int W, int C, int Ho, int Wo, int stride, int up, cudaStream_t s) {
    const long long total = (long long)F * Ho * Wo * 9 * (C / 8);
    if (total >= (1ll << 31)) return false;  // the 64-bit kernel of nn_kernels.cu stays in charge
    im2col3x3_idx_kernel<uint32_t><<<fast_grid(total), 256, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)col, (uint32_t)total, H, W, C, Ho, Wo,
                                                                   stride, up);
    return true;
}

bool gvd_fast_im2col_t3(const void* x, void* col, int B, int T, long long S, int C, cudaStream_t s) {
    const long long total = (long long)B * T * S * 3 * (C / 8);
    if (total >= (1ll << 31)) return false;
    im2col_t3_idx_kernel<uint32_t><<<fast_grid(total), 256, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)col, (uint32_t)total, T, (uint32_t)S, C);
    return true;
}

bool gvd_fast_temporal_attention(const void* q, const void* k, const void* v, void* out, int B, int T, long long S, int H, float scale,
                                 cudaStream_t s) {
    const long long warps = (long long)B * S * H;
    const int smem = TAF_WARPS * 2 * 32 * 64 * (int)sizeof(float);  // 64 KB
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(temporal_attn_f32stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return false;
        attr_set = true;
    }
    temporal_attn_f32stage_kernel<<<(unsigned)((warps + TAF_WARPS - 1) / TAF_WARPS), TAF_WARPS * 32, smem, s>>>(
        (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, (__nv_bfloat16*)out, B, T, S, H, scale);
    return true;
}
