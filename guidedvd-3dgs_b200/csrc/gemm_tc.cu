// gemm_tc.cu -- strided-batched bf16 GEMM on Blackwell tensor cores (include/gvd_nn.h::gvd_gemm_bf16).
//
// Two kernels:
//  * gemm_bf16_persistent_kernel<BN> (default): persistent CTAs (one per SM) loop over 128 x BN tiles (BN = 64/128/256,
//    picked per problem); the accumulator is DOUBLE-BUFFERED in TMEM so the MMA warp runs tile i+1 while the four
//    epilogue warps drain tile i; the epilogue stages bf16 rows in swizzled shared memory and writes them (and reads
//    the residual) with fully coalesced 16-byte accesses.
//  * gemm_bf16_kernel (fallback: fp32 output or unaligned destinations): one 128x128 tile per CTA.
//
// Fallback kernel -- one CTA computes a 128x128 tile of C:
//   warp 0   : TMA producer   -- cp.async.bulk.tensor (4-D tensor maps, 128-byte swizzle) into a 3-stage smem ring
//   warp 1   : MMA issuer     -- one elected lane issues tcgen05.mma (M=128, N=128, K=16, bf16 -> fp32 in TMEM),
//                                tcgen05.commit releases smem stages / signals the epilogue; also owns TMEM alloc
//   warps 2-5: epilogue       -- tcgen05.ld (32 lanes x 16 columns), alpha / bias / activation / residual, bf16 or
//                                fp32 stores straight to the (strided) destination
// Two CTAs fit per SM (96 KB of stages each, 128 TMEM columns each), so one CTA's epilogue overlaps the other's
// main loop.  Out-of-range rows/columns/k are zero-filled by TMA and masked in the epilogue.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cstdlib>
#include <string>

#include "../../include/gvd_nn.h"
#include "tc_common.cuh"

thread_local std::string g_nn_err_ext;
#define g_nn_err g_nn_err_ext

namespace {


constexpr int BM = 128, BN = 128, BK = 64, STAGES = 3;
constexpr int A_STAGE_BYTES = BM * BK * 2, B_STAGE_BYTES = BN * BK * 2;
constexpr int SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int NUM_THREADS = 192;

struct EpiParams {
    void* C;
    long long ldc, c_stride_h, c_stride_b;
    const float* bias;
    const float* bias2;
    const void* residual;
    float alpha;
    int act, out_fp32;
    int M, N, K, batch_h;
    int b_mn;  // B operand is MN-major: smem tile = BK rows of 64 n (128 bytes), tensor-map coordinates (n, k, h, b)
    // implicit-GEMM convolution (gvd_conv_bf16): 0 = plain GEMM; 1 = 3x3 over (x, y) of a (c, x, y, frame) map;
    // 2 = 3 taps over the frame axis of a (c, pixel, frame, batch) map.  K block kb = (tap, 64-channel block).
    int conv_kind, conv_w, conv_cin;
    int tma_group;  // k blocks requested per burst by the one-CTA kernel's producer (1 .. stages / 2)
};

__device__ __forceinline__ float bf16r(float x) { return __bfloat162float(__float2bfloat16(x)); }

__device__ __forceinline__ float apply_act(float x, int act) {
    if (act == GVD_ACT_SILU) return x / (1.0f + __expf(-x));
    if (act == GVD_ACT_GELU) return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
    return x;
}

__global__ void __launch_bounds__(NUM_THREADS, 2)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, EpiParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES));
    uint64_t* full = bars;               // [STAGES]
    uint64_t* empty = bars + STAGES;     // [STAGES]
    uint64_t* tmem_full = bars + 2 * STAGES;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
    const int bh = blockIdx.z % p.batch_h, bb = blockIdx.z / p.batch_h;
    const int num_kb = (p.K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmap_a);
        tc::prefetch_tmap(&tmap_b);
        for (int s = 0; s < STAGES; ++s) {
            tc::mbar_init(&full[s], 1);
            tc::mbar_init(&empty[s], 1);
        }
        tc::mbar_init(tmem_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tmem_ptr, BN);  // 128 fp32 columns
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (uint32_t)((kb / STAGES) & 1);
                tc::mbar_wait(&empty[s], ph ^ 1u);
                tc::mbar_expect_tx(&full[s], A_STAGE_BYTES + B_STAGE_BYTES);
                tc::tma_load_4d(smem_a + s * A_STAGE_BYTES, &tmap_a, &full[s], kb * BK, m0, bh, bb);
                tc::tma_load_4d(smem_b + s * B_STAGE_BYTES, &tmap_b, &full[s], kb * BK, n0, bh, bb);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = tc::make_idesc_bf16(BM, BN);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (uint32_t)((kb / STAGES) & 1);
                tc::mbar_wait(&full[s], ph);
                tc::fence_after_sync();
                const uint32_t a_addr = tc::smem_u32(smem_a + s * A_STAGE_BYTES);
                const uint32_t b_addr = tc::smem_u32(smem_b + s * B_STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    const uint64_t adesc = tc::make_desc_kmajor_sw128(a_addr + k * 32);
                    const uint64_t bdesc = tc::make_desc_kmajor_sw128(b_addr + k * 32);
                    tc::umma_bf16(tmem_base, adesc, bdesc, idesc, (kb | k) != 0);
                }
                tc::umma_commit(&empty[s]);  // frees this smem stage once the MMAs above have read it
            }
            tc::umma_commit(tmem_full);      // accumulator complete
        }
    } else {
        // ---- epilogue: warp w may only touch TMEM lanes 32*(w%4) .. +31 ----
        const int q = warp & 3;
        tc::mbar_wait(tmem_full, 0);
        tc::fence_after_sync();
        const int m = m0 + q * 32 + lane;
        const bool row_ok = m < p.M;
        const long long row_off = (long long)bb * p.c_stride_b + (long long)bh * p.c_stride_h + (long long)m * p.ldc;
        const bool vec_ok = ((p.ldc & 7) == 0) && ((p.c_stride_h & 7) == 0) && ((p.c_stride_b & 7) == 0) &&
                            ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                            (p.residual == nullptr || (reinterpret_cast<uintptr_t>(p.residual) & 15) == 0);
#pragma unroll 1
        for (int c = 0; c < BN; c += 16) {
            if (n0 + c >= p.N) break;  // warp-uniform
            uint32_t v[16];
            tc::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
            tc::tmem_ld_wait();
            if (!row_ok) continue;
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float x = __uint_as_float(v[j]);
                const int n = n0 + c + j;
                if (p.act == GVD_ACT_ROUND_SCALE) {
                    x = bf16r(bf16r(x) * p.alpha);  // einsum output rounded to bf16, then "* scale" in bf16 (attention.py:103)
                } else {
                    x *= p.alpha;
                    if (p.bias != nullptr && n < p.N) x += p.bias[n];
                    x = apply_act(x, p.act);
                }
                if (!p.out_fp32) {
                    // the reference materialises the layer output in bf16 before any following add: keep those rounding points
                    if (p.bias2 != nullptr && n < p.N) x = bf16r(x) + p.bias2[n];
                    if (p.residual != nullptr) x = bf16r(x);
                }
                f[j] = x;
            }
            const int n_first = n0 + c;
            const bool full16 = (n_first + 16 <= p.N);
            if (p.out_fp32) {
                float* dst = reinterpret_cast<float*>(p.C) + row_off + n_first;
                const float* res = p.residual ? reinterpret_cast<const float*>(p.residual) + row_off + n_first : nullptr;
                if (full16 && vec_ok) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        float4 o = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                        if (res) {
                            const float4 r = *reinterpret_cast<const float4*>(res + j);
                            o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                        }
                        *reinterpret_cast<float4*>(dst + j) = o;
                    }
                } else {
                    for (int j = 0; j < 16 && n_first + j < p.N; ++j) dst[j] = f[j] + (res ? res[j] : 0.f);
                }
            } else {
                __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.C) + row_off + n_first;
                const __nv_bfloat16* res =
                    p.residual ? reinterpret_cast<const __nv_bfloat16*>(p.residual) + row_off + n_first : nullptr;
                if (full16 && vec_ok) {
#pragma unroll
                    for (int j = 0; j < 16; j += 8) {
                        float r8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                        if (res) {
                            const uint4 rv = *reinterpret_cast<const uint4*>(res + j);
                            const __nv_bfloat162* rp = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                const float2 rf = __bfloat1622float2(rp[t]);
                                r8[2 * t] = rf.x;
                                r8[2 * t + 1] = rf.y;
                            }
                        }
                        uint4 ov;
                        __nv_bfloat162* op = reinterpret_cast<__nv_bfloat162*>(&ov);
#pragma unroll
                        for (int t = 0; t < 4; ++t)
                            op[t] = __floats2bfloat162_rn(f[j + 2 * t] + r8[2 * t], f[j + 2 * t + 1] + r8[2 * t + 1]);
                        *reinterpret_cast<uint4*>(dst + j) = ov;
                    }
                } else {
                    for (int j = 0; j < 16 && n_first + j < p.N; ++j)
                        dst[j] = __float2bfloat16(f[j] + (res ? __bfloat162float(res[j]) : 0.f));
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, BN);
}


// ====================================================================================================================
// Persistent kernel
// ====================================================================================================================
constexpr int P_THREADS = 320;  // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quadrant)

template <int BN_> struct PCfg {
    static constexpr int STAGES_ = BN_ == 256 ? 3 : (BN_ == 128 ? 6 : 8);
    static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN_ * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int HALF = BN_ / 2;                           // columns handled by one epilogue warp
    static constexpr int EPI_WARP_BYTES = 32 * HALF * 2;          // one warp's 32 rows x BN/2 bf16
    static constexpr int BAR_OFF = STAGES_ * STAGE_BYTES + 8 * EPI_WARP_BYTES;
    static constexpr int SMEM = BAR_OFF + 256 + 1024;
    static constexpr int TMEM_COLS = 2 * BN_ <= 128 ? 128 : (2 * BN_ <= 256 ? 256 : 512);
};

// 32 accumulator columns of one row -> alpha / bias / activation / bias2 with the reference's bf16 rounding points -> bf16
template <int MODE>  // 0: plain (alpha, optional bias), 1: general
__device__ __forceinline__ void epi_convert32(const uint32_t (&v)[32], uint32_t (&packed)[16], const EpiParams& p, int n_first) {
    if (MODE == 0) {
        float b[32];
        if (p.bias != nullptr) {
            if (n_first + 32 <= p.N && ((n_first & 3) == 0)) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + n_first + j));
                    b[j] = t.x; b[j + 1] = t.y; b[j + 2] = t.z; b[j + 3] = t.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) b[j] = (n_first + j < p.N) ? p.bias[n_first + j] : 0.f;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) b[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
            __nv_bfloat162 h2 = __floats2bfloat162_rn(fmaf(__uint_as_float(v[j]), p.alpha, b[j]),
                                                      fmaf(__uint_as_float(v[j + 1]), p.alpha, b[j + 1]));
            packed[j / 2] = *reinterpret_cast<uint32_t*>(&h2);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
            float x2[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float x = __uint_as_float(v[j + e]);
                const int n = n_first + j + e;
                if (p.act == GVD_ACT_ROUND_SCALE) {
                    x = bf16r(bf16r(x) * p.alpha);
                } else {
                    x *= p.alpha;
                    if (p.bias != nullptr && n < p.N) x += p.bias[n];
                    x = apply_act(x, p.act);
                }
                if (p.bias2 != nullptr && n < p.N) x = bf16r(x) + p.bias2[n];
                x2[e] = x;
            }
            __nv_bfloat162 h2 = __floats2bfloat162_rn(x2[0], x2[1]);
            packed[j / 2] = *reinterpret_cast<uint32_t*>(&h2);
        }
    }
}

// GELU for the fused GEGLU epilogue: erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, two MUFU + ~12 FMA-pipe
// instructions instead of erff's ~30), evaluated so that the negative tail keeps its relative accuracy:
//   gelu(g) = g - 0.5 g q  (g >= 0),  0.5 g q  (g < 0),  q = erfc(|g| / sqrt 2) = poly(t) exp(-g^2 / 2),  t = 1 / (1 + p |g| / sqrt 2)
__device__ __forceinline__ float gelu_fast(float g) {
    const float y = fabsf(g) * 0.70710678118654752f;
    const float t = __fdividef(1.0f, fmaf(0.3275911f, y, 1.0f));
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    const float q = poly * t * __expf(-y * y);
    const float hq = 0.5f * g * q;
    return g >= 0.0f ? g - hq : hq;
}

// One epilogue warp's share of a 128 x BN_ accumulator tile: its TMEM lane quadrant q (32 rows) x one half of the columns.
//   phase 1: TMEM -> registers -> alpha / bias / activation (reference rounding points) -> bf16 -> swizzled smem row `lane`
//   phase 2: coalesced 16-byte write-out (+ residual, prefetched before the accumulator was ready)
// GEGLU (act == GVD_ACT_GEGLU): the weight rows were interleaved in blocks of 16 ([16 value rows | 16 gate rows], see
// vc_b200.ops.geglu_weight), so every 32 accumulator columns hold 16 values and their 16 gates; the warp writes
// out[., n / 2 ..] = bf16(value) * bf16(gelu(bf16(gate))) -- GEGLU (attention.py:415-423) without the 2D-wide intermediate.
// PAIR: the accumulator buffer is handed back to the LEADER CTA's MMA thread (cluster-scope arrive).
template <int BN_, bool GEGLU, bool PAIR>
__device__ __forceinline__ void epilogue_tile(const EpiParams& p, uint8_t* stage, uint32_t tmem_base, uint64_t* tmem_full_bar,
                                              uint32_t full_parity, uint64_t* tmem_empty_bar, uint32_t tmem_empty_cluster, int m0,
                                              int n0, long long base_off, int buf, int q, int half, int lane, bool plain) {
    constexpr int HALF = BN_ / 2;                       // accumulator columns of this warp
    constexpr int OUT_HALF = GEGLU ? HALF / 2 : HALF;   // output columns of this warp
    constexpr int ROW_BYTES = OUT_HALF * 2, CHUNKS = ROW_BYTES / 16;  // 16-byte chunks per staged row
    constexpr int ROWS_PER_PASS = 32 / CHUNKS;
    constexpr int SWZ = CHUNKS >= 8 ? 7 : CHUNKS - 1;
    const int chunk = lane % CHUNKS, rsub = lane / CHUNKS;
    const int n_out0 = GEGLU ? n0 / 2 : n0, n_lim = GEGLU ? p.N / 2 : p.N;
    const int n = n_out0 + chunk * 8;
    uint4 resv[32 / ROWS_PER_PASS];
    if (!GEGLU && p.residual != nullptr) {
#pragma unroll
        for (int i = 0; i < 32 / ROWS_PER_PASS; ++i) {
            const int m = m0 + q * 32 + i * ROWS_PER_PASS + rsub;
            resv[i] = make_uint4(0, 0, 0, 0);
            if (m < p.M && n < n_lim)
                resv[i] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.residual) + base_off +
                                                               (long long)m * p.ldc + n));
        }
    }
    tc::mbar_wait(tmem_full_bar, full_parity);
    tc::fence_after_sync();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN_ + half * HALF);
#pragma unroll 1
    for (int c = 0; c < HALF; c += 32) {
        if (n0 + c >= p.N) break;
        uint32_t v[32], packed[16];
        tc::tmem_ld32(taddr + (uint32_t)c, v);
        tc::tmem_ld_wait();
        if (GEGLU) {
            const int nb = n0 + c;  // 32 | nb; p.N % 32 == 0, so the block is whole
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                float o2[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    float a = __uint_as_float(v[j + e]) * p.alpha, g = __uint_as_float(v[16 + j + e]) * p.alpha;
                    if (p.bias != nullptr) {
                        a += p.bias[nb + j + e];
                        g += p.bias[nb + 16 + j + e];
                    }
                    o2[e] = bf16r(a) * bf16r(gelu_fast(bf16r(g)));
                }
                __nv_bfloat162 h2 = __floats2bfloat162_rn(o2[0], o2[1]);
                packed[j / 2] = *reinterpret_cast<uint32_t*>(&h2);
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int ch = c / 16 + k;
                *reinterpret_cast<uint4*>(stage + lane * ROW_BYTES + ((ch ^ (lane & SWZ)) << 4)) =
                    make_uint4(packed[4 * k], packed[4 * k + 1], packed[4 * k + 2], packed[4 * k + 3]);
            }
        } else {
            if (plain) epi_convert32<0>(v, packed, p, n0 + c);
            else epi_convert32<1>(v, packed, p, n0 + c);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int ch = c / 8 + k;
                *reinterpret_cast<uint4*>(stage + lane * ROW_BYTES + ((ch ^ (lane & SWZ)) << 4)) =
                    make_uint4(packed[4 * k], packed[4 * k + 1], packed[4 * k + 2], packed[4 * k + 3]);
            }
        }
    }
    // all TMEM reads of this warp are done: hand the accumulator buffer back to the MMA thread
    tc::fence_before_sync();
    __syncwarp();
    if (lane == 0) {
        if (PAIR) tc::mbar_arrive_cluster(tmem_empty_cluster);
        else tc::mbar_arrive(tmem_empty_bar);
    }
#pragma unroll
    for (int i = 0; i < 32 / ROWS_PER_PASS; ++i) {
        const int rr = i * ROWS_PER_PASS + rsub;
        const int m = m0 + q * 32 + rr;
        if (m < p.M && n < n_lim) {
            uint4 val = *reinterpret_cast<const uint4*>(stage + rr * ROW_BYTES + ((chunk ^ (rr & SWZ)) << 4));
            const long long off = base_off + (long long)m * p.ldc + n;
            if (!GEGLU && p.residual != nullptr) {
                const uint4 rv = resv[i];
                __nv_bfloat162* a2 = reinterpret_cast<__nv_bfloat162*>(&val);
                const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 af = __bfloat1622float2(a2[e]), rf = __bfloat1622float2(r2[e]);
                    a2[e] = __floats2bfloat162_rn(af.x + rf.x, af.y + rf.y);
                }
            }
            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.C) + off) = val;
        }
    }
    __syncwarp();  // the staging rows are rewritten by the next tile
}

template <int BN_, bool GEGLU>
__global__ void __launch_bounds__(P_THREADS, 1)
gemm_bf16_persistent_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                            EpiParams p, int tiles_m, int tiles_n, int total_tiles) {
    using Cfg = PCfg<BN_>;
    constexpr int ST = Cfg::STAGES_;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_epi = smem + ST * Cfg::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
    uint64_t* full = bars;                   // [ST]   TMA -> MMA
    uint64_t* empty = bars + ST;             // [ST]   MMA -> TMA
    uint64_t* tmem_full = bars + 2 * ST;     // [2]    MMA -> epilogue
    uint64_t* tmem_empty = bars + 2 * ST + 2;  // [2]  epilogue -> MMA
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * ST + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = (p.K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmap_a);
        tc::prefetch_tmap(&tmap_b);
        for (int s = 0; s < ST; ++s) {
            tc::mbar_init(&full[s], 1);
            tc::mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            tc::mbar_init(&tmem_full[b], 1);
            tc::mbar_init(&tmem_empty[b], 8);  // one arrival per epilogue warp
        }
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;
    const int tiles_mn = tiles_m * tiles_n;

    if (warp == 0) {
        if (lane == 0) {
            int kc = 0;  // running k-block counter across tiles -> stage / phase
            const int G = ST >= 6 ? p.tma_group : 1;  // bursts only where the ring is deep enough to keep MMAs fed meanwhile
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int z = tile / tiles_mn, r = tile - z * tiles_mn;
                const int m0 = (r / tiles_n) * BM, n0 = (r % tiles_n) * BN_;
                const int bh = z % p.batch_h, bb = z / p.batch_h;
                if (p.conv_kind) {
                    // implicit-GEMM convolution: the A tile of tap (dy, dx) is the activation tile shifted by (dy, dx);
                    // negative / past-the-end coordinates are zero-filled by TMA, which is the zero padding.
                    // Coordinates advance incrementally: no division inside the k loop.
                    // kind 2: rows are (frame, pixel) flattened, a tap shifts them by conv_w = S rows; bh is the batch item
                    int ax = m0 - p.conv_w, ay = bh;
                    if (p.conv_kind == 1) {
                        ay = m0 / p.conv_w;
                        ax = m0 - ay * p.conv_w;
                    }
                    const int taps = p.conv_kind == 1 ? 9 : 3;
                    int dx = -1, dy = -1, kw = 0;  // kw: k coordinate in the weight matrix
                    for (int tap = 0; tap < taps; ++tap) {
                        const int cx = p.conv_kind == 1 ? ax + dx : ax + tap * p.conv_w;
                        const int cy = p.conv_kind == 1 ? ay + dy : ay;
                        for (int cc = 0; cc < p.conv_cin; cc += BK * G) {  // bursts of tma_group k blocks (see below)
                            const int g_n = min(G, (p.conv_cin - cc) / BK);
                            for (int g = 0; g < g_n; ++g) tc::mbar_wait(&empty[(kc + g) % ST], (uint32_t)((((kc + g) / ST) & 1) ^ 1));
                            for (int g = 0; g < g_n; ++g, ++kc, kw += BK) {
                                const int s = kc % ST;
                                tc::mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
                                uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
                                tc::tma_load_4d(sa, &tmap_a, &full[s], cc + g * BK, cx, cy, bb);
                                tc::tma_load_4d(sa + Cfg::A_BYTES, &tmap_b, &full[s], kw, n0, 0, 0);
                            }
                        }
                        if (++dx == 2) { dx = -1; ++dy; }
                    }
                    continue;
                }
                // k blocks are requested in bursts of G: the rows of an A tile are far apart in memory, and the
                // 128-byte pieces a row contributes to consecutive k blocks reach DRAM together instead of 256 clk apart
                for (int kb = 0; kb < num_kb; kb += G) {
                    const int g_n = min(G, num_kb - kb);
                    for (int g = 0; g < g_n; ++g) tc::mbar_wait(&empty[(kc + g) % ST], (uint32_t)((((kc + g) / ST) & 1) ^ 1));
                    for (int g = 0; g < g_n; ++g, ++kc) {
                        const int s = kc % ST, k0 = (kb + g) * BK;
                        tc::mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
                        uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
                        tc::tma_load_4d(sa, &tmap_a, &full[s], k0, m0, bh, bb);
                        if (p.b_mn) tc::tma_load_4d(sa + Cfg::A_BYTES, &tmap_b, &full[s], n0, k0, bh, bb);
                        else tc::tma_load_4d(sa + Cfg::A_BYTES, &tmap_b, &full[s], k0, n0, bh, bb);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // MN-major B (bit 16 of the instruction descriptor): same 8-row / 1024-byte swizzle atoms, but the rows are
            // k and one MMA (K = 16) consumes 16 of them -- the operand advances by 16 * 128 bytes per step
            const uint32_t idesc = tc::make_idesc_bf16(BM, BN_) | (p.b_mn ? (1u << 16) : 0u);
            const uint32_t b_step = p.b_mn ? 16u * 128u : 32u;
            int kc = 0, it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                tc::mbar_wait(&tmem_empty[buf], (uint32_t)(((it >> 1) & 1) ^ 1));  // epilogue has drained this buffer
                tc::fence_after_sync();
                const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN_);
                for (int kb = 0; kb < num_kb; ++kb, ++kc) {
                    const int s = kc % ST;
                    const uint32_t ph = (uint32_t)((kc / ST) & 1);
                    tc::mbar_wait(&full[s], ph);
                    tc::fence_after_sync();
                    const uint32_t a_addr = tc::smem_u32(smem + s * Cfg::STAGE_BYTES);
                    const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        tc::umma_bf16(tmem_d, tc::make_desc_kmajor_sw128(a_addr + k * 32),
                                      tc::make_desc_kmajor_sw128(b_addr + k * b_step), idesc, (kb | k) != 0);
                    tc::umma_commit(&empty[s]);
                }
                tc::umma_commit(&tmem_full[buf]);
            }
        }
    } else {
        const int ew = warp - 2;
        const int q = warp & 3;      // TMEM lane quadrant this warp may access (hardware rule: warp id % 4)
        const int half = ew >> 2;    // which half of the tile's columns
        uint8_t* stage = smem_epi + ew * Cfg::EPI_WARP_BYTES;
        const bool plain = (p.act == GVD_ACT_NONE) && (p.bias2 == nullptr);
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int z = tile / tiles_mn, r = tile - z * tiles_mn;
            const int m0 = (r / tiles_n) * BM, n0 = (r % tiles_n) * BN_ + half * Cfg::HALF;
            const int bh = z % p.batch_h, bb = z / p.batch_h;
            const int buf = it & 1;
            epilogue_tile<BN_, GEGLU, false>(p, stage, tmem_base, &tmem_full[buf], (uint32_t)((it >> 1) & 1), &tmem_empty[buf], 0u, m0, n0,
                                             (long long)bb * p.c_stride_b + (long long)bh * p.c_stride_h, buf, q, half, lane, plain);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}


// ====================================================================================================================
// CTA-pair kernel (cta_group::2): one cluster of two CTAs (the two SMs of a TPC) computes a 256 x BN_ tile.
// Each CTA stages ITS 128 rows of A and ITS half of the B tile (BN_/2 rows); the leader's single thread issues
// tcgen05.mma.cta_group::2 (M = 256), which reads A and B from both CTAs' shared memory and accumulates rows 0-127 in
// the leader's TMEM and rows 128-255 in the peer's.  Per SM and K = 16 step that is 4 KB of A + BN_/2 x 32 B of B from
// shared memory instead of 4 KB + BN_ x 32 B: at BN_ = 128 the one-CTA kernel needs the whole 128 B/clk of its SM's
// shared memory, the pair 96 B/clk; at BN_ = 256 96 -> 64 B/clk.  Same roles and epilogue as the one-CTA kernel.
// Barriers: full[s] lives in the leader (its producer arms it for the bytes of BOTH CTAs, both CTAs' TMA loads
// complete on it), empty[s] / tmem_full[b] are signalled in both CTAs by multicast commits, tmem_empty[b] collects
// the 16 epilogue warps of the pair in the leader.
// ====================================================================================================================
template <int BN_> struct P2Cfg {
    static constexpr int BH = BN_ / 2;  // B rows staged per CTA
    static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BH * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int HALF = BN_ / 2;
    static constexpr int EPI_WARP_BYTES = 32 * HALF * 2;
    static constexpr int STAGES_ = BN_ == 256 ? 5 : 7;
    static constexpr int BAR_OFF = STAGES_ * STAGE_BYTES + 8 * EPI_WARP_BYTES;
    static constexpr int SMEM = BAR_OFF + 256 + 1024;
    static constexpr int TMEM_COLS = 2 * BN_ <= 256 ? 256 : 512;
};

template <int BN_, bool GEGLU>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P_THREADS, 1)
gemm_bf16_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, EpiParams p,
                      int tiles_m, int tiles_n, int total_tiles) {
    using Cfg = P2Cfg<BN_>;
    constexpr int ST = Cfg::STAGES_;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_epi = smem + ST * Cfg::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
    uint64_t* full = bars;                     // [ST]  used in the leader only
    uint64_t* empty = bars + ST;               // [ST]  per CTA
    uint64_t* tmem_full = bars + 2 * ST;       // [2]   per CTA
    uint64_t* tmem_empty = bars + 2 * ST + 2;  // [2]   used in the leader only
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * ST + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = tc::cluster_ctarank();
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int num_kb = (p.K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmap_a);
        tc::prefetch_tmap(&tmap_b);
        for (int s = 0; s < ST; ++s) {
            tc::mbar_init(&full[s], 1);
            tc::mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            tc::mbar_init(&tmem_full[b], 1);
            tc::mbar_init(&tmem_empty[b], 16);  // the 8 epilogue warps of each CTA of the pair
        }
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc_2cta(tmem_ptr, Cfg::TMEM_COLS);
    tc::fence_before_sync();
    __syncthreads();
    tc::cluster_sync_all();  // the peer's barriers are initialised before anything signals them
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;
    const int tiles_mn = tiles_m * tiles_n;

    if (warp == 0) {
        if (lane == 0) {
            int kc = 0;
            const int G = ST >= 6 ? p.tma_group : 1;
            for (int tile = pair; tile < total_tiles; tile += npairs) {
                const int z = tile / tiles_mn, r = tile - z * tiles_mn;
                const int m0 = (r / tiles_n) * (2 * BM) + (int)rank * BM, nb0 = (r % tiles_n) * BN_ + (int)rank * Cfg::BH;
                const int bh = z % p.batch_h, bb = z / p.batch_h;
                int ax = m0 - p.conv_w, ay = bh;  // implicit-GEMM convolution coordinates (see the one-CTA kernel)
                if (p.conv_kind == 1) {
                    ay = m0 / p.conv_w;
                    ax = m0 - ay * p.conv_w;
                }
                const int taps = p.conv_kind == 1 ? 9 : (p.conv_kind == 2 ? 3 : 1);
                const int kper = p.conv_kind ? p.conv_cin : num_kb * BK;
                int dx = -1, dy = -1, kw = 0;
                for (int tap = 0; tap < taps; ++tap) {
                    const int cx = p.conv_kind == 1 ? ax + dx : (p.conv_kind == 2 ? ax + tap * p.conv_w : m0);
                    const int cy = p.conv_kind == 1 ? ay + dy : bh;
                    for (int cc = 0; cc < kper; cc += BK * G) {  // bursts of tma_group k blocks, as in the one-CTA kernel
                        const int g_n = min(G, (kper - cc + BK - 1) / BK);
                        for (int g = 0; g < g_n; ++g) tc::mbar_wait(&empty[(kc + g) % ST], (uint32_t)((((kc + g) / ST) & 1) ^ 1));
                        for (int g = 0; g < g_n; ++g, ++kc, kw += BK) {
                            const int s = kc % ST;
                            const uint32_t fb = tc::mapa(tc::smem_u32(&full[s]), 0);
                            if (rank == 0) tc::mbar_expect_tx(&full[s], 2 * Cfg::STAGE_BYTES);
                            uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
                            tc::tma_load_4d_2cta(sa, &tmap_a, fb, cc + g * BK, cx, cy, bb);
                            if (p.conv_kind) tc::tma_load_4d_2cta(sa + Cfg::A_BYTES, &tmap_b, fb, kw, nb0, 0, 0);
                            else tc::tma_load_4d_2cta(sa + Cfg::A_BYTES, &tmap_b, fb, kw, nb0, bh, bb);
                        }
                    }
                    if (++dx == 2) { dx = -1; ++dy; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            const uint32_t idesc = tc::make_idesc_bf16(2 * BM, BN_);
            int kc = 0, it = 0;
            for (int tile = pair; tile < total_tiles; tile += npairs, ++it) {
                const int buf = it & 1;
                tc::mbar_wait(&tmem_empty[buf], (uint32_t)(((it >> 1) & 1) ^ 1));
                tc::fence_after_sync();
                const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN_);
                for (int kb = 0; kb < num_kb; ++kb, ++kc) {
                    const int s = kc % ST;
                    tc::mbar_wait(&full[s], (uint32_t)((kc / ST) & 1));
                    tc::fence_after_sync();
                    const uint32_t a_addr = tc::smem_u32(smem + s * Cfg::STAGE_BYTES);
                    const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        tc::umma_bf16_2cta(tmem_d, tc::make_desc_kmajor_sw128(a_addr + k * 32),
                                           tc::make_desc_kmajor_sw128(b_addr + k * 32), idesc, (kb | k) != 0);
                    tc::umma_commit_2cta(&empty[s]);
                }
                tc::umma_commit_2cta(&tmem_full[buf]);
            }
        }
    } else {
        const int ew = warp - 2;
        const int q = warp & 3;
        const int half = ew >> 2;
        uint8_t* stage = smem_epi + ew * Cfg::EPI_WARP_BYTES;
        const bool plain = (p.act == GVD_ACT_NONE) && (p.bias2 == nullptr);
        const uint32_t te0 = tc::mapa(tc::smem_u32(&tmem_empty[0]), 0), te1 = tc::mapa(tc::smem_u32(&tmem_empty[1]), 0);
        int it = 0;
        for (int tile = pair; tile < total_tiles; tile += npairs, ++it) {
            const int z = tile / tiles_mn, r = tile - z * tiles_mn;
            const int m0 = (r / tiles_n) * (2 * BM) + (int)rank * BM, n0 = (r % tiles_n) * BN_ + half * Cfg::HALF;
            const int bh = z % p.batch_h, bb = z / p.batch_h;
            const int buf = it & 1;
            epilogue_tile<BN_, GEGLU, true>(p, stage, tmem_base, &tmem_full[buf], (uint32_t)((it >> 1) & 1), nullptr, buf ? te1 : te0, m0, n0,
                                            (long long)bb * p.c_stride_b + (long long)bh * p.c_stride_h, buf, q, half, lane, plain);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::cluster_sync_all();  // both CTAs are done with the pair's tensor memory and with each other's barriers
    if (warp == 1) tc::tmem_dealloc_2cta(tmem_base, Cfg::TMEM_COLS);
}

template <int BN_, bool GEGLU = false>
cudaError_t launch_pair(const CUtensorMap& ta, const CUtensorMap& tb, const EpiParams& p, int batch, cudaStream_t s) {
    using Cfg = P2Cfg<BN_>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_bf16_pair_kernel<BN_, GEGLU>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (num_sms <= 0) num_sms = 148;
    }
    const int tiles_m = (p.M + 2 * BM - 1) / (2 * BM), tiles_n = (p.N + BN_ - 1) / BN_;
    const long long total = (long long)tiles_m * tiles_n * batch;
    const int pairs = (int)(total < num_sms / 2 ? total : num_sms / 2);
    gemm_bf16_pair_kernel<BN_, GEGLU><<<2 * pairs, P_THREADS, Cfg::SMEM, s>>>(ta, tb, p, tiles_m, tiles_n, (int)total);
    return cudaGetLastError();
}

// GVD_GEMM_PAIR: 1 (default) = CTA-pair kernel where it applies, 0 = one-CTA kernels only (A/B timing)
bool pair_enabled() {
#ifdef GVD_HOST_EMU
    return false;  // tests/cuda_emu runs one block at a time: no CTA pairs there
#endif
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("GVD_GEMM_PAIR");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}

template <int BN_, bool GEGLU = false>
cudaError_t launch_persistent(const CUtensorMap& ta, const CUtensorMap& tb, const EpiParams& p, int batch, cudaStream_t s) {
    using Cfg = PCfg<BN_>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_bf16_persistent_kernel<BN_, GEGLU>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (num_sms <= 0) num_sms = 148;
    }
    const int tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN_ - 1) / BN_;
    const long long total = (long long)tiles_m * tiles_n * batch;
    const int grid = (int)(total < num_sms ? total : num_sms);
    gemm_bf16_persistent_kernel<BN_, GEGLU><<<grid, P_THREADS, Cfg::SMEM, s>>>(ta, tb, p, tiles_m, tiles_n, (int)total);
    return cudaGetLastError();
}

// ---- host side: tensor maps ----
PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
#ifdef GVD_HOST_EMU
    return emu_cuTensorMapEncodeTiled;
#else
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
#endif
}

// 4-D bf16 view: (K contiguous, rows with stride ld, h with stride sh, b with stride sb), all in elements
bool make_tmap(CUtensorMap* map, const void* base, long long K, long long rows, long long H, long long Bn, long long ld,
               long long sh, long long sb, int box_rows) {
    auto enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[4] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)H, (cuuint64_t)Bn};
    // strides of a size-1 dimension are never used to form an address but must still be valid (multiple of 16 bytes)
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)(H > 1 ? sh : ld) * 2, (cuuint64_t)(Bn > 1 ? sb : ld) * 2};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// MN-major B: 4-D bf16 view (N contiguous, K rows with stride ld, h, b); box = 64 n x BK k
bool make_tmap_mn(CUtensorMap* map, const void* base, long long N, long long K, long long H, long long Bn, long long ld,
                  long long sh, long long sb) {
    auto enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[4] = {(cuuint64_t)N, (cuuint64_t)K, (cuuint64_t)H, (cuuint64_t)Bn};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)(H > 1 ? sh : ld) * 2, (cuuint64_t)(Bn > 1 ? sb : ld) * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)BK, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// When the CTA-pair kernel pays (B200, tools/bench_gemm_pair.py, profiles/r02_gemm_pair_sweep.txt): 256-wide tiles with
// K >= 640 (+3 % at K = 640, +11-13 % at 1280-5120, +24 % at K = 11520); at K = 320 its per-tile cross-CTA hand-overs cost
// 13-17 %, its 128-wide form loses everywhere, and few tiles leave the second wave of 74 pairs empty.
// GVD_GEMM_TMA_GROUP: k blocks per producer burst.  Default 2: measured on B200 against 1 in the same process order
// (tools/bench_gemm_pair.py, A streamed from HBM): M = 230400, N = 320: K = 1280 0.26-0.29 -> 0.25 ms, K = 2880 0.59 -> 0.53 ms;
// M = 57600, N = 640, K = 5760 0.46-0.51 -> 0.44 ms; 3 is slightly behind 2.
int tma_group() {
    static int g = 0;
    if (!g) {
        const char* e = getenv("GVD_GEMM_TMA_GROUP");
        g = (e && e[0] >= '1' && e[0] <= '3') ? e[0] - '0' : 2;
    }
    return g;
}

// Column split for outputs that are neither narrow nor a multiple of 256 (N = 320, 640 in the U-Net): the leading
// floor(N / 256) * 256 columns run on 256-wide tiles (96 B/clk of shared-memory operand traffic per SM, pairs at 64) and
// only the remainder on a narrow tile, instead of everything on 128-wide tiles, which need the whole 128 B/clk and pad
// N = 320 to 384.  Two launches, same k order per output element -- bit-identical results.  GVD_GEMM_SPLIT=0 turns it off.
// Measured (profiles/r02_gemm_split_sweep.txt): implicit convolutions gain 16-23 % (25 x 72 x 128, 320 -> 320: 874 -> 1042
// TFLOP/s; 36 x 64, 640 -> 640: 1008 -> 1241), linears only from K ~ 2560 (+5 %); below that the narrow launch is a second
// HBM pass over A for a sliver of the flops and the split LOSES 15-22 % (K = 320: 590 -> 458), so plain GEMMs split from
// K = 2048 on; every 3 x 3 convolution of the U-Net qualifies (K = 9 Cin >= 2880), the temporal ones from Cin = 1280.
bool split_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("GVD_GEMM_SPLIT");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}
int split_columns(long long M, int N, int K, int batch, bool conv) {
    if (!split_enabled() || N <= 256 || (N % 256) == 0 || (N % 8) != 0) return 0;
    (void)conv;
    if (K < 2048) return 0;
    if (((M + BM - 1) / BM) * batch < 148) return 0;  // each launch should fill the machine on its own
    return N / 256 * 256;
}

bool use_pair(long long M, int N, int K, int batch, int bn) {
    if (!pair_enabled() || bn != 256 || K < 640) return false;
    const long long t128 = (M + 127) / 128, t256 = (M + 255) / 256, tn = (N + bn - 1) / bn;
    if (t256 * 2 * 8 > t128 * 9) return false;  // 256-row tiles would pad M by more than 12.5 % over 128-row tiles
    const long long tp = t256 * tn * batch, to = t128 * tn * batch;
    const double eff_pair = (double)tp / (double)((tp + 73) / 74 * 74), eff_one = (double)to / (double)((to + 147) / 148 * 148);
    return eff_pair >= 0.9 * eff_one;
}

}  // namespace

static int conv_columns(const GvdConvArgs* a, const CUtensorMap& ta, long long M, int batch_h, int batch_b, int taps, int kpt, int n_first,
                        int N, cudaStream_t s);

extern "C" {

const char* gvd_nn_last_error(void) { return g_nn_err.c_str(); }

int gvd_gemm_bf16(const GvdGemmArgs* a, gvd_nn_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (!a || !a->A || !a->B || !a->C) { g_nn_err = "gvd_gemm_bf16: null pointer"; return 2; }
    if (a->M <= 0 || a->N <= 0 || a->batch_h <= 0 || a->batch_b <= 0) return 0;
    if (a->K <= 0) { g_nn_err = "gvd_gemm_bf16: K must be positive"; return 2; }
    auto bad = [](long long v) { return (v & 7) != 0; };
    if (bad(a->lda) || bad(a->ldb) || (a->batch_h > 1 && (bad(a->a_stride_h) || bad(a->b_stride_h))) ||
        (a->batch_b > 1 && (bad(a->a_stride_b) || bad(a->b_stride_b))) ||
        (reinterpret_cast<uintptr_t>(a->A) & 15) || (reinterpret_cast<uintptr_t>(a->B) & 15)) {
        g_nn_err = "gvd_gemm_bf16: operand strides must be multiples of 8 elements and bases 16-byte aligned";
        return 2;
    }
    if (a->act == GVD_ACT_GEGLU) {
        const bool ok = !a->out_fp32 && !a->residual && !a->bias2 && !a->b_mn_major && (a->N % 32 == 0) && (a->ldc % 8 == 0) &&
                        (a->batch_h == 1 || a->c_stride_h % 8 == 0) && (a->batch_b == 1 || a->c_stride_b % 8 == 0) &&
                        (reinterpret_cast<uintptr_t>(a->C) % 16 == 0);
        if (!ok) {
            g_nn_err = "gvd_gemm_bf16: GVD_ACT_GEGLU needs N % 32 == 0, bf16 output with 16-byte aligned rows, no residual / bias2 / b_mn_major";
            return 2;
        }
        CUtensorMap ga, gb;
        EpiParams gp{a->C, a->ldc, a->c_stride_h, a->c_stride_b, a->bias, nullptr, nullptr, a->alpha, a->act, 0,
                     a->M, a->N, a->K, a->batch_h, 0, 0, 0, 0, tma_group()};
        const int batch = a->batch_h * a->batch_b;
        const bool pair = use_pair(a->M, a->N, a->K, batch, 256);
        if (!make_tmap(&ga, a->A, a->K, a->M, a->batch_h, a->batch_b, a->lda, a->a_stride_h, a->a_stride_b, BM) ||
            !make_tmap(&gb, a->B, a->K, a->N, a->batch_h, a->batch_b, a->ldb, a->b_stride_h, a->b_stride_b, pair ? 128 : 256)) {
            g_nn_err = "gvd_gemm_bf16: cuTensorMapEncodeTiled failed";
            return 1;
        }
        cudaError_t e = pair ? launch_pair<256, true>(ga, gb, gp, batch, s) : launch_persistent<256, true>(ga, gb, gp, batch, s);
        if (e != cudaSuccess) { g_nn_err = std::string("gvd_gemm_bf16 GEGLU launch: ") + cudaGetErrorString(e); return 1; }
        return 0;
    }
    CUtensorMap ta, tb;
    EpiParams p{a->C, a->ldc, a->c_stride_h, a->c_stride_b, a->bias, a->bias2, a->residual, a->alpha, a->act, a->out_fp32,
                a->M, a->N, a->K, a->batch_h, a->b_mn_major != 0, 0, 0, 0, tma_group()};
    const bool aligned_out = !a->out_fp32 && (a->N % 8 == 0) && (a->ldc % 8 == 0) &&
                             (a->batch_h == 1 || a->c_stride_h % 8 == 0) && (a->batch_b == 1 || a->c_stride_b % 8 == 0) &&
                             (reinterpret_cast<uintptr_t>(a->C) % 16 == 0) &&
                             (a->residual == nullptr || reinterpret_cast<uintptr_t>(a->residual) % 16 == 0);
    const long long total_tiles128 = (long long)((a->M + BM - 1) / BM) * ((a->N + 127) / 128) * a->batch_h * a->batch_b;
    if (a->b_mn_major) {
        if (!aligned_out) { g_nn_err = "gvd_gemm_bf16: b_mn_major needs bf16 output with 16-byte aligned rows"; return 2; }
        if (!make_tmap(&ta, a->A, a->K, a->M, a->batch_h, a->batch_b, a->lda, a->a_stride_h, a->a_stride_b, BM) ||
            !make_tmap_mn(&tb, a->B, a->N, a->K, a->batch_h, a->batch_b, a->ldb, a->b_stride_h, a->b_stride_b)) {
            g_nn_err = "gvd_gemm_bf16: cuTensorMapEncodeTiled failed";
            return 1;
        }
        cudaError_t e = launch_persistent<64>(ta, tb, p, a->batch_h * a->batch_b, s);
        if (e != cudaSuccess) { g_nn_err = std::string("gvd_gemm_bf16 persistent launch: ") + cudaGetErrorString(e); return 1; }
        return 0;
    }
    if (aligned_out && total_tiles128 < (1ll << 30)) {
        if (const int n1 = split_columns(a->M, a->N, a->K, a->batch_h * a->batch_b, false)) {
            GvdGemmArgs lo = *a, hi = *a;
            lo.N = n1;
            hi.N = a->N - n1;
            hi.B = reinterpret_cast<const __nv_bfloat16*>(a->B) + (long long)n1 * a->ldb;
            hi.C = reinterpret_cast<__nv_bfloat16*>(a->C) + n1;
            if (a->bias) hi.bias = a->bias + n1;
            if (a->bias2) hi.bias2 = a->bias2 + n1;
            if (a->residual) hi.residual = reinterpret_cast<const __nv_bfloat16*>(a->residual) + n1;
            const int r = gvd_gemm_bf16(&lo, stream_);
            return r ? r : gvd_gemm_bf16(&hi, stream_);
        }
        // tile width: least padding of N, ties to the wider tile (fewer A re-reads); 64 only for narrow outputs
        auto padded = [&](int bn) { return (long long)((a->N + bn - 1) / bn) * bn; };
        int bn = 128;
        if (a->N <= 64) bn = 64;
        else if (padded(256) <= padded(128)) bn = 256;
        const int batch = a->batch_h * a->batch_b;
        const bool pair = use_pair(a->M, a->N, a->K, batch, bn);
        if (!make_tmap(&ta, a->A, a->K, a->M, a->batch_h, a->batch_b, a->lda, a->a_stride_h, a->a_stride_b, BM) ||
            !make_tmap(&tb, a->B, a->K, a->N, a->batch_h, a->batch_b, a->ldb, a->b_stride_h, a->b_stride_b, pair ? bn / 2 : bn)) {
            g_nn_err = "gvd_gemm_bf16: cuTensorMapEncodeTiled failed";
            return 1;
        }
        if (pair) {
            cudaError_t e = launch_pair<256>(ta, tb, p, batch, s);
            if (e != cudaSuccess) { g_nn_err = std::string("gvd_gemm_bf16 pair launch: ") + cudaGetErrorString(e); return 1; }
            return 0;
        }
        cudaError_t e = bn == 256 ? launch_persistent<256>(ta, tb, p, batch, s)
                      : bn == 128 ? launch_persistent<128>(ta, tb, p, batch, s)
                                  : launch_persistent<64>(ta, tb, p, batch, s);
        if (e != cudaSuccess) { g_nn_err = std::string("gvd_gemm_bf16 persistent launch: ") + cudaGetErrorString(e); return 1; }
        return 0;
    }
    if (!make_tmap(&ta, a->A, a->K, a->M, a->batch_h, a->batch_b, a->lda, a->a_stride_h, a->a_stride_b, BM) ||
        !make_tmap(&tb, a->B, a->K, a->N, a->batch_h, a->batch_b, a->ldb, a->b_stride_h, a->b_stride_b, BN)) {
        g_nn_err = "gvd_gemm_bf16: cuTensorMapEncodeTiled failed";
        return 1;
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) { g_nn_err = std::string("gvd_gemm_bf16 attr: ") + cudaGetErrorString(e); return 1; }
        attr_set = true;
    }
    dim3 grid((a->N + BN - 1) / BN, (a->M + BM - 1) / BM, a->batch_h * a->batch_b);
    gemm_bf16_kernel<<<grid, NUM_THREADS, SMEM_BYTES, s>>>(ta, tb, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_nn_err = std::string("gvd_gemm_bf16 launch: ") + cudaGetErrorString(e); return 1; }
    return 0;
}

// ---- implicit-GEMM convolution -------------------------------------------------------------------------------
int gvd_conv_bf16_supported(int kind, int H, int W, int Cin, int Cout) {
    if (Cin <= 0 || Cout <= 0 || (Cin % 64) != 0 || (Cout % 8) != 0) return 0;
    if (kind == 2) return 1;
    if (kind != 1 || H <= 0 || W <= 0) return 0;
    if (!((W % 128 == 0) || (W <= 128 && 128 % W == 0))) return 0;
    // output tiles are 128 pixels of ONE frame: small frames would leave most of the last tile empty (9 x 16 pixels fill
    // 56 % of two tiles) -- those stay on the im2col route, whose rows run across frames
    const long long px = (long long)H * W, tiles = (px + 127) / 128;
    return tiles * 128 * 8 <= px * 9;
}

int gvd_conv_bf16(const GvdConvArgs* a, gvd_nn_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (!a || !a->x || !a->weight || !a->y) { g_nn_err = "gvd_conv_bf16: null pointer"; return 2; }
    if (!gvd_conv_bf16_supported(a->kind, a->H, a->W, a->Cin, a->Cout)) {
        g_nn_err = "gvd_conv_bf16: geometry not served by the implicit-GEMM route (see gvd_conv_bf16_supported)";
        return 2;
    }
    if ((reinterpret_cast<uintptr_t>(a->x) & 15) || (reinterpret_cast<uintptr_t>(a->weight) & 15) ||
        (reinterpret_cast<uintptr_t>(a->y) & 15) || (a->residual && (reinterpret_cast<uintptr_t>(a->residual) & 15))) {
        g_nn_err = "gvd_conv_bf16: tensors must be 16-byte aligned";
        return 2;
    }
    auto enc = get_encode();
    if (!enc) { g_nn_err = "gvd_conv_bf16: cuTensorMapEncodeTiled unavailable"; return 1; }
    const int taps = a->kind == 1 ? 9 : 3;
    const int kpt = a->Cin / BK;
    if (a->kind == 2 && a->S >= (1ll << 30)) { g_nn_err = "gvd_conv_bf16: S too large"; return 2; }
    const long long C = a->Cin;
    long long M;
    int batch_h, batch_b;
    cuuint64_t dims[4], strides[3];
    cuuint32_t box[4] = {(cuuint32_t)BK, 0, 0, 1}, estr[4] = {1, 1, 1, 1};
    if (a->kind == 1) {
        if (a->F <= 0) return 0;
        M = (long long)a->H * a->W;
        batch_h = 1;
        batch_b = a->F;
        dims[0] = (cuuint64_t)C; dims[1] = (cuuint64_t)a->W; dims[2] = (cuuint64_t)a->H; dims[3] = (cuuint64_t)a->F;
        strides[0] = (cuuint64_t)C * 2; strides[1] = (cuuint64_t)a->W * C * 2; strides[2] = (cuuint64_t)M * C * 2;
        box[1] = (cuuint32_t)(a->W >= 128 ? 128 : a->W);
        box[2] = (cuuint32_t)(128 / box[1]);
    } else {
        if (a->B <= 0 || a->T <= 0 || a->S <= 0) return 0;
        // rows = (frame, pixel) flattened per batch item: a temporal tap is a shift by S rows, and the rows before the
        // first / after the last frame are out of bounds of the map (zero-filled) -- tiles run across frames, no padding
        M = (long long)a->T * a->S;
        batch_h = a->B;
        batch_b = 1;
        dims[0] = (cuuint64_t)C; dims[1] = (cuuint64_t)M; dims[2] = (cuuint64_t)a->B; dims[3] = 1;
        strides[0] = (cuuint64_t)C * 2; strides[1] = (cuuint64_t)M * C * 2; strides[2] = (cuuint64_t)M * C * 2;
        box[1] = 128;
        box[2] = 1;
    }
    if (M >= (1ll << 31)) { g_nn_err = "gvd_conv_bf16: more than 2^31 pixels per frame"; return 2; }
    CUtensorMap ta;
    if (enc(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(a->x), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        g_nn_err = "gvd_conv_bf16: cuTensorMapEncodeTiled failed for the activation map";
        return 1;
    }
    const int N = a->Cout;
    if (const int n1 = split_columns(M, N, taps * a->Cin, batch_h * batch_b, true)) {
        // y keeps its row stride Cout: the narrow launch below must write with ldc = the full Cout, so the split lives in
        // conv_columns(), which takes the column window explicitly
        const int r = conv_columns(a, ta, M, batch_h, batch_b, taps, kpt, 0, n1, s);
        return r ? r : conv_columns(a, ta, M, batch_h, batch_b, taps, kpt, n1, N - n1, s);
    }
    return conv_columns(a, ta, M, batch_h, batch_b, taps, kpt, 0, N, s);
}

}  // extern "C"

// output columns [n_first, n_first + N) of the implicit-GEMM convolution whose activation map is `ta`
static int conv_columns(const GvdConvArgs* a, const CUtensorMap& ta, long long M, int batch_h, int batch_b, int taps, int kpt, int n_first,
                        int N, cudaStream_t s) {
    const long long C = a->Cin;
    const long long Kw = (long long)taps * C;
    CUtensorMap tb;
    auto padded = [&](int bn) { return (long long)((N + bn - 1) / bn) * bn; };
    int bn = 128;
    if (N <= 64) bn = 64;
    else if (padded(256) <= padded(128)) bn = 256;
    const bool pair = use_pair(M, N, taps * a->Cin, batch_h * batch_b, bn);
    const __nv_bfloat16* w = reinterpret_cast<const __nv_bfloat16*>(a->weight) + (long long)n_first * Kw;
    if (!make_tmap(&tb, w, Kw, N, 1, 1, Kw, 0, 0, pair ? bn / 2 : bn)) { g_nn_err = "gvd_conv_bf16: cuTensorMapEncodeTiled failed for the weight map"; return 1; }
    const long long ldc = a->Cout;
    EpiParams p{reinterpret_cast<__nv_bfloat16*>(a->y) + n_first, ldc, (long long)M * ldc, (long long)batch_h * M * ldc,
                a->bias ? a->bias + n_first : nullptr, a->bias2 ? a->bias2 + n_first : nullptr,
                a->residual ? reinterpret_cast<const __nv_bfloat16*>(a->residual) + n_first : nullptr, 1.0f, a->act, 0,
                (int)M, N, taps * kpt * BK, batch_h, 0, a->kind, a->kind == 1 ? a->W : (int)a->S, a->Cin, tma_group()};
    const int batch = batch_h * batch_b;
    cudaError_t e = pair ? launch_pair<256>(ta, tb, p, batch, s)
                  : bn == 256 ? launch_persistent<256>(ta, tb, p, batch, s)
                  : bn == 128 ? launch_persistent<128>(ta, tb, p, batch, s)
                              : launch_persistent<64>(ta, tb, p, batch, s);
    if (e != cudaSuccess) { g_nn_err = std::string("gvd_conv_bf16 launch: ") + cudaGetErrorString(e); return 1; }
    return 0;
}
