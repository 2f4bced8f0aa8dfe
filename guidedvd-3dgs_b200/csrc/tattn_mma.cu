// tattn_mma.cu -- temporal self-attention over T <= 32 frames per (pixel, head), head dim 64, on warp-level tensor-core
// tiles (include/gvd_nn.h::gvd_temporal_attention; CrossAttention inside TemporalTransformer, attention.py:365-412 with
// '(b h w) t c' sequences).
//
// The first kernel (nn_kernels.cu::temporal_attn_kernel, lane = query frame, K/V in shared memory) spends ~5 000 warp
// instructions per (pixel, head) on scalar FMAs and bf16 -> fp32 conversions and runs at 13 % of the HBM rate
// (profiles/r02_unet_kernel_breakdown.txt).  Here one warp owns one (pixel, head) and the two 32 x 32 x 64 products are
// 64 mma.sync.m16n8k16 (bf16 in, fp32 accumulate) with NO shared memory:
//   * Q and K rows are read from global memory straight into A / B fragments with 16-byte loads: lane (g, c) takes the
//     16-byte chunks c and c + 4 of rows g, g + 8, g + 16, g + 24.  That is a permutation of the head dimension relative
//     to the canonical fragment layout, applied to Q and K alike, so the dot products are unchanged.
//   * S = Q K^T lands in the C layout, which IS the A layout of P for the second product (rows g / g + 8, column pairs).
//   * V is read the same way and turned into B fragments (pairs along the KEY axis) with 32 movmatrix.trans; the head
//     dimension comes out permuted such that every lane ends up with 8 consecutive output channels -> 16-byte stores.
// Rounding points are those of the reference under autocast (and of the first kernel): bf16(bf16(q.k) * scale), fp32
// softmax, bf16 probabilities, fp32 PV accumulation, bf16 output.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

#ifndef GVD_HOST_EMU
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t movm_trans(uint32_t x) {
    uint32_t y;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}
#else  // tests/cuda_emu: the two warp-collective tile instructions with the PTX fragment layouts spelled out
inline void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    emu_mma_m16n8k16_bf16(d, a0, a1, a2, a3, b0, b1);
}
inline uint32_t movm_trans(uint32_t x) { return emu_movmatrix_trans_b16(x); }
#endif
__device__ __forceinline__ float bf16r(float x) { return __bfloat162float(__float2bfloat16(x)); }
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

constexpr int TM_WARPS = 4;

// MT = number of 16-row query tiles (1: T <= 16, 2: T <= 32); key tiles of 8: NT = 2 * MT
template <int MT>
__global__ void __launch_bounds__(TM_WARPS * 32, 3) temporal_attn_mma_kernel(const __nv_bfloat16* __restrict__ q,
                                                                           const __nv_bfloat16* __restrict__ k,
                                                                           const __nv_bfloat16* __restrict__ v,
                                                                           __nv_bfloat16* __restrict__ out, int B, int T,
                                                                           long long S, int H, float scale) {
    constexpr int NT = 2 * MT, RJ = 2 * MT;  // RJ: rows per lane (g + 8 j)
    const int lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
    const long long w = (long long)blockIdx.x * TM_WARPS + (threadIdx.x >> 5);
    if (w >= (long long)B * S * H) return;
    const int h = (int)(w % H);
    const long long s = (w / H) % S;
    const int b = (int)(w / ((long long)H * S));
    const long long tstride = S * H * 64;
    const size_t base = ((size_t)b * T * S + s) * H * 64 + (size_t)h * 64 + (size_t)c * 8;

    uint32_t qr[RJ][8], kr[RJ][8], vr[RJ][8];
#pragma unroll
    for (int j = 0; j < RJ; ++j) {
        const int t = g + 8 * j;
        uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0, b0 = a0, b1 = a0, c0 = a0, c1 = a0;
        if (t < T) {
            const size_t off = base + (size_t)t * tstride;
            a0 = __ldg(reinterpret_cast<const uint4*>(q + off));
            a1 = __ldg(reinterpret_cast<const uint4*>(q + off + 32));
            b0 = __ldg(reinterpret_cast<const uint4*>(k + off));
            b1 = __ldg(reinterpret_cast<const uint4*>(k + off + 32));
            c0 = __ldg(reinterpret_cast<const uint4*>(v + off));
            c1 = __ldg(reinterpret_cast<const uint4*>(v + off + 32));
        }
        qr[j][0] = a0.x; qr[j][1] = a0.y; qr[j][2] = a0.z; qr[j][3] = a0.w;
        qr[j][4] = a1.x; qr[j][5] = a1.y; qr[j][6] = a1.z; qr[j][7] = a1.w;
        kr[j][0] = b0.x; kr[j][1] = b0.y; kr[j][2] = b0.z; kr[j][3] = b0.w;
        kr[j][4] = b1.x; kr[j][5] = b1.y; kr[j][6] = b1.z; kr[j][7] = b1.w;
        vr[j][0] = c0.x; vr[j][1] = c0.y; vr[j][2] = c0.z; vr[j][3] = c0.w;
        vr[j][4] = c1.x; vr[j][5] = c1.y; vr[j][6] = c1.z; vr[j][7] = c1.w;
    }

    // ---- S = Q K^T : query tile mt = rows (g, g + 8) + 16 mt = lane rows j = 2 mt, 2 mt + 1; key tile nt = lane row j = nt ----
    float sacc[MT][NT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) sacc[mt][nt][e] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
                mma16816(sacc[mt][nt], qr[2 * mt][2 * ks], qr[2 * mt + 1][2 * ks], qr[2 * mt][2 * ks + 1], qr[2 * mt + 1][2 * ks + 1],
                         kr[nt][2 * ks], kr[nt][2 * ks + 1]);
        }

    // ---- V -> B fragments: after the transposition lane (g, c) holds, in vr[j][r], the pair V[8 j + 2 c + {0, 1}][d(g, r)] ----
#pragma unroll
    for (int j = 0; j < RJ; ++j)
#pragma unroll
        for (int r = 0; r < 8; ++r) vr[j][r] = movm_trans(vr[j][r]);

    // ---- softmax per query row: a row's 8 * NT... (8 nt + 2 c + e) keys are spread over the 4 lanes of a quad ----
    uint32_t pa[MT][NT][2];  // bf16 pairs: [..][0] row g + 16 mt, [..][1] row g + 8 + 16 mt
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            float x[NT][2];
            float m = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int tk = 8 * nt + 2 * c + e;
                    // the reference rounds the einsum output to bf16 and scales in bf16 before the fp32 softmax (attention.py:103)
                    const float val = tk < T ? bf16r(bf16r(sacc[mt][nt][2 * hf + e]) * scale) : -INFINITY;
                    x[nt][e] = val;
                    m = fmaxf(m, val);
                }
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
            float l = 0.f;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    x[nt][e] = __expf(x[nt][e] - m);
                    l += x[nt][e];
                }
            l += __shfl_xor_sync(0xffffffffu, l, 1);
            l += __shfl_xor_sync(0xffffffffu, l, 2);
            const float inv = 1.0f / l;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) pa[mt][nt][hf] = pack2(x[nt][0] * inv, x[nt][1] * inv);  // bf16 probabilities
        }

    // ---- O = P V : k step ks2 = keys 16 ks2 .. + 15 (key tiles 2 ks2, 2 ks2 + 1); n tile r = 8 output channels ----
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        float oacc[8][4];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int e = 0; e < 4; ++e) oacc[r][e] = 0.f;
#pragma unroll
            for (int ks2 = 0; ks2 < MT; ++ks2)
                mma16816(oacc[r], pa[mt][2 * ks2][0], pa[mt][2 * ks2][1], pa[mt][2 * ks2 + 1][0], pa[mt][2 * ks2 + 1][1],
                         vr[2 * ks2][r], vr[2 * ks2 + 1][r]);
        }
        // lane (g, c) holds, over r = 0..3 (4..7), the 8 consecutive channels of chunk c (c + 4) of rows g and g + 8
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            const int t = g + 8 * hf + 16 * mt;
            if (t < T) {
                __nv_bfloat16* dst = out + base + (size_t)t * tstride;
                *reinterpret_cast<uint4*>(dst) = make_uint4(pack2(oacc[0][2 * hf], oacc[0][2 * hf + 1]), pack2(oacc[1][2 * hf], oacc[1][2 * hf + 1]),
                                                            pack2(oacc[2][2 * hf], oacc[2][2 * hf + 1]), pack2(oacc[3][2 * hf], oacc[3][2 * hf + 1]));
                *reinterpret_cast<uint4*>(dst + 32) = make_uint4(pack2(oacc[4][2 * hf], oacc[4][2 * hf + 1]), pack2(oacc[5][2 * hf], oacc[5][2 * hf + 1]),
                                                                 pack2(oacc[6][2 * hf], oacc[6][2 * hf + 1]), pack2(oacc[7][2 * hf], oacc[7][2 * hf + 1]));
            }
        }
    }
}

// ---- backward: the same tiles, five products per (pixel, head) ----
//   S = Q K^T, dP = dO V^T          (A = Q / dO fragments, B = K / V fragments: all straight from global memory)
//   P = softmax(S), dS = P o (dP - rowsum(P o dP)) * scale      (C layout = the A layout of the next products)
//   dQ = dS K        (B = K with the KEY axis as k: the fragments of K after movmatrix.trans)
//   dV = P^T dO,  dK = dS^T Q   (A = the 8 x 8 blocks of P / dS transposed by movmatrix; B = dO / Q after movmatrix.trans)
// The first backward (nn_backward.cu::temporal_attn_bwd_kernel) does the same with scalar FMAs out of shared memory
// (lane = frame): 16.8 ms of a guided step.  Rounding points: those of the forward for S and P; dS enters its two
// products as bf16 (the first backward kept it in fp32 -- a difference far below the bf16 rounding of dQ / dK).
template <int MT>
__global__ void __launch_bounds__(TM_WARPS * 32, 2) temporal_attn_bwd_mma_kernel(
    const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
    const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ dq, __nv_bfloat16* __restrict__ dk,
    __nv_bfloat16* __restrict__ dv, int B, int T, long long S, int H, float scale) {
    constexpr int NT = 2 * MT, RJ = 2 * MT;
    const int lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
    const long long w = (long long)blockIdx.x * TM_WARPS + (threadIdx.x >> 5);
    if (w >= (long long)B * S * H) return;
    const int h = (int)(w % H);
    const long long s = (w / H) % S;
    const int b = (int)(w / ((long long)H * S));
    const long long tstride = S * H * 64;
    const size_t base = ((size_t)b * T * S + s) * H * 64 + (size_t)h * 64 + (size_t)c * 8;

    auto load_frag = [&](const __nv_bfloat16* src, uint32_t (&r)[RJ][8]) {
#pragma unroll
        for (int j = 0; j < RJ; ++j) {
            const int t = g + 8 * j;
            uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
            if (t < T) {
                const size_t off = base + (size_t)t * tstride;
                a0 = __ldg(reinterpret_cast<const uint4*>(src + off));
                a1 = __ldg(reinterpret_cast<const uint4*>(src + off + 32));
            }
            r[j][0] = a0.x; r[j][1] = a0.y; r[j][2] = a0.z; r[j][3] = a0.w;
            r[j][4] = a1.x; r[j][5] = a1.y; r[j][6] = a1.z; r[j][7] = a1.w;
        }
    };
    // rows (g, g + 8) + 16 mt of `a` against rows 8 nt + g of `bm`, contracted over the head dimension
    auto qk_product = [&](const uint32_t (&a)[RJ][8], const uint32_t (&bm)[RJ][8], float (&acc)[MT][NT][4]) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    mma16816(acc[mt][nt], a[2 * mt][2 * ks], a[2 * mt + 1][2 * ks], a[2 * mt][2 * ks + 1], a[2 * mt + 1][2 * ks + 1],
                             bm[nt][2 * ks], bm[nt][2 * ks + 1]);
            }
    };
    // out rows (g, g + 8) + 16 mt = sum over 16-row k steps of A (C-layout pairs a[mt][nt][hf]) x bt (transposed fragments)
    auto store_rows = [&](__nv_bfloat16* dst_base, const float (&o)[8][4], int mt) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            const int t = g + 8 * hf + 16 * mt;
            if (t < T) {
                __nv_bfloat16* dst = dst_base + base + (size_t)t * tstride;
                *reinterpret_cast<uint4*>(dst) = make_uint4(pack2(o[0][2 * hf], o[0][2 * hf + 1]), pack2(o[1][2 * hf], o[1][2 * hf + 1]),
                                                            pack2(o[2][2 * hf], o[2][2 * hf + 1]), pack2(o[3][2 * hf], o[3][2 * hf + 1]));
                *reinterpret_cast<uint4*>(dst + 32) = make_uint4(pack2(o[4][2 * hf], o[4][2 * hf + 1]), pack2(o[5][2 * hf], o[5][2 * hf + 1]),
                                                                 pack2(o[6][2 * hf], o[6][2 * hf + 1]), pack2(o[7][2 * hf], o[7][2 * hf + 1]));
            }
        }
    };

    uint32_t qr[RJ][8], kr[RJ][8], dor[RJ][8];
    float pf[MT][NT][4];  // S, then P (fp32, normalised)
    load_frag(q, qr);
    load_frag(k, kr);
    qk_product(qr, kr, pf);
    load_frag(dout, dor);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            float m = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int tk = 8 * nt + 2 * c + e;
                    const float val = tk < T ? bf16r(bf16r(pf[mt][nt][2 * hf + e]) * scale) : -INFINITY;  // attention.py:103
                    pf[mt][nt][2 * hf + e] = val;
                    m = fmaxf(m, val);
                }
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
            float l = 0.f;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float x = __expf(pf[mt][nt][2 * hf + e] - m);
                    pf[mt][nt][2 * hf + e] = x;
                    l += x;
                }
            l += __shfl_xor_sync(0xffffffffu, l, 1);
            l += __shfl_xor_sync(0xffffffffu, l, 2);
            const float inv = 1.0f / l;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) pf[mt][nt][2 * hf + e] *= inv;
        }

    uint32_t pa[MT][NT][2], dsa[MT][NT][2];  // bf16 pairs in the C layout: [..][0] row g + 16 mt, [..][1] row g + 8 + 16 mt
    {
        uint32_t vr[RJ][8];
        float dp[MT][NT][4];
        load_frag(v, vr);
        qk_product(dor, vr, dp);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                float rowdot = 0.f;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) rowdot = fmaf(pf[mt][nt][2 * hf + e], dp[mt][nt][2 * hf + e], rowdot);
                rowdot += __shfl_xor_sync(0xffffffffu, rowdot, 1);
                rowdot += __shfl_xor_sync(0xffffffffu, rowdot, 2);
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const float p0 = pf[mt][nt][2 * hf], p1 = pf[mt][nt][2 * hf + 1];
                    pa[mt][nt][hf] = pack2(p0, p1);
                    dsa[mt][nt][hf] = pack2(p0 * (dp[mt][nt][2 * hf] - rowdot) * scale, p1 * (dp[mt][nt][2 * hf + 1] - rowdot) * scale);
                }
            }
    }

    // ---- dQ = dS K : k step ks2 = keys 16 ks2 .. + 15; B = K fragments with the key axis as k ----
#pragma unroll
    for (int j = 0; j < RJ; ++j)
#pragma unroll
        for (int r = 0; r < 8; ++r) kr[j][r] = movm_trans(kr[j][r]);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        float o[8][4];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int e = 0; e < 4; ++e) o[r][e] = 0.f;
#pragma unroll
            for (int ks2 = 0; ks2 < MT; ++ks2)
                mma16816(o[r], dsa[mt][2 * ks2][0], dsa[mt][2 * ks2][1], dsa[mt][2 * ks2 + 1][0], dsa[mt][2 * ks2 + 1][1], kr[2 * ks2][r],
                         kr[2 * ks2 + 1][r]);
        }
        store_rows(dq, o, mt);
    }

    // ---- dV = P^T dO and dK = dS^T Q : rows are keys (tile kt), k step mt = queries 16 mt .. + 15 ----
#pragma unroll
    for (int j = 0; j < RJ; ++j)
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            dor[j][r] = movm_trans(dor[j][r]);
            qr[j][r] = movm_trans(qr[j][r]);
        }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {  // block (queries 16 mt + 8 hf .., keys 8 nt ..) -> (keys 8 nt .., queries 16 mt + 8 hf ..)
                pa[mt][nt][hf] = movm_trans(pa[mt][nt][hf]);
                dsa[mt][nt][hf] = movm_trans(dsa[mt][nt][hf]);
            }
#pragma unroll
    for (int kt = 0; kt < MT; ++kt) {
        float o[8][4];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int e = 0; e < 4; ++e) o[r][e] = 0.f;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
                mma16816(o[r], pa[mt][2 * kt][0], pa[mt][2 * kt + 1][0], pa[mt][2 * kt][1], pa[mt][2 * kt + 1][1], dor[2 * mt][r], dor[2 * mt + 1][r]);
        }
        store_rows(dv, o, kt);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int e = 0; e < 4; ++e) o[r][e] = 0.f;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
                mma16816(o[r], dsa[mt][2 * kt][0], dsa[mt][2 * kt + 1][0], dsa[mt][2 * kt][1], dsa[mt][2 * kt + 1][1], qr[2 * mt][r], qr[2 * mt + 1][r]);
        }
        store_rows(dk, o, kt);
    }
}

}  // namespace

bool gvd_mma_temporal_attention_bwd(const void* q, const void* k, const void* v, const void* dout, void* dq, void* dk, void* dv, int B, int T,
                                    long long S, int H, float scale, cudaStream_t s) {
    if (T <= 0 || T > 32) return false;
    if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(dout) |
         reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dk) | reinterpret_cast<uintptr_t>(dv)) & 15)
        return false;
    const long long warps = (long long)B * S * H;
    const unsigned grid = (unsigned)((warps + TM_WARPS - 1) / TM_WARPS);
    auto args = [&](auto kern) {
        kern<<<grid, TM_WARPS * 32, 0, s>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, (const __nv_bfloat16*)dout,
                                            (__nv_bfloat16*)dq, (__nv_bfloat16*)dk, (__nv_bfloat16*)dv, B, T, S, H, scale);
    };
    if (T <= 16) args(temporal_attn_bwd_mma_kernel<1>);
    else args(temporal_attn_bwd_mma_kernel<2>);
    return true;
}

// Returns false when the geometry is not served (the caller launches the first kernel instead).
bool gvd_mma_temporal_attention(const void* q, const void* k, const void* v, void* out, int B, int T, long long S, int H, float scale,
                                cudaStream_t s) {
    if (T <= 0 || T > 32) return false;
    if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(out)) & 15)
        return false;
    const long long warps = (long long)B * S * H;
    const unsigned grid = (unsigned)((warps + TM_WARPS - 1) / TM_WARPS);
    if (T <= 16)
        temporal_attn_mma_kernel<1><<<grid, TM_WARPS * 32, 0, s>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v,
                                                                   (__nv_bfloat16*)out, B, T, S, H, scale);
    else
        temporal_attn_mma_kernel<2><<<grid, TM_WARPS * 32, 0, s>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v,
                                                                   (__nv_bfloat16*)out, B, T, S, H, scale);
    return true;
}

#ifdef GVD_HOST_EMU
// C entry points of the host build (tests/test_tattn_mma_emu_cpu.py); on the GPU the two functions above are reached through
// gvd_temporal_attention / gvd_temporal_attention_bwd
extern "C" int gvd_emu_mma_temporal_attention(const void* q, const void* k, const void* v, void* out, int B, int T, long long S, int H,
                                              float scale) {
    return gvd_mma_temporal_attention(q, k, v, out, B, T, S, H, scale, nullptr) ? 0 : 2;
}
extern "C" int gvd_emu_mma_temporal_attention_bwd(const void* q, const void* k, const void* v, const void* dout, void* dq, void* dk, void* dv,
                                                  int B, int T, long long S, int H, float scale) {
    return gvd_mma_temporal_attention_bwd(q, k, v, dout, dq, dk, dv, B, T, S, H, scale, nullptr) ? 0 : 2;
}
#endif
