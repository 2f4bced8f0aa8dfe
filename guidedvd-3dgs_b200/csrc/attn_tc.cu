// attn_tc.cu -- fused softmax(Q K^T * scale) V on Blackwell tensor cores, head dim 64 (include/gvd_nn.h::gvd_flash_attention).
//
// Never materialises the score matrix (the reference's einsum path writes [b*h, N, N] scores: 21 GB per layer at
// 576x1024, attention.py:103,118).  One CTA = 128 query rows of one (batch, head); key/value blocks of 128 stream through
// a 2-stage TMA ring:
//   warp 0    : TMA producer  (Q once, then K_j / V_j tiles, 128-byte swizzle)
//   warp 1    : MMA issuer    S = Q K_j^T        (SS: both operands in smem, accumulator S in TMEM, 128 fp32 columns)
//                             O_j = P_j V_j      (TS: A = P_j read from TMEM as packed bf16, B = V_j smem tile, MN-major)
//   warps 2-5 : softmax       thread = query row. The whole S row (128 fp32) is pulled into registers with four
//                             back-to-back tcgen05.ld and ONE wait; row max, exp -> bf16 P written back to TMEM with
//                             tcgen05.st. O accumulates IN TMEM across key blocks (the PV MMA runs with accumulate on);
//                             the softmax thread rescales its O row in place only when its running max has grown by
//                             more than 2^8 (lazy rescale: P stays <= 256, far inside bf16/fp32 range, and the final
//                             1/l normalisation uses the same stale max, so the result is unchanged).  Optionally one
//                             exponential in four is evaluated on the FMA pipe (Cody-Waite split + cubic).
// TMEM: S [0,128) | P [128,192) | O [192,256)  -> 256 columns, two CTAs per SM.
// (flash_attn_v1_kernel below is the first version -- two passes over S, O folded in registers every block -- kept
// selectable with GVD_FLASH=v1 for A/B timing.)
// Logits stay in fp32 (acc * scale); the reference rounds them to bf16 twice before its fp32 softmax (attention.py:103),
// which this kernel deliberately does not emulate: the emulation costs three ALU ops per score in a loop that is
// MUFU/ALU-bound, and the difference is below the bf16 noise of the PV product (tests bound the error).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cstdlib>
#include <string>

#include "../../include/gvd_nn.h"
#include "tc_common.cuh"

extern thread_local std::string g_nn_err_ext;

namespace {

constexpr int FA_BM = 128, FA_BN = 128, FA_D = 64;
constexpr int FA_TILE_BYTES = 128 * 64 * 2;  // 16 KB: Q, K_j or V_j tile
constexpr int FA_STAGES = 2;
constexpr int FA_SMEM = FA_TILE_BYTES * (1 + 2 * FA_STAGES) + 1024 + 256;
constexpr int FA_THREADS = 192;
constexpr uint32_t FA_TMEM_COLS = 256, FA_S_COL = 0, FA_P_COL = 128, FA_O_COL = 192;

#ifndef GVD_HOST_EMU
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
#else  // tests/cuda_emu/tc_emu.h
inline float ex2(float x) { return exp2f(x); }
#endif

// MN-major B operand tile (V_j: 128 keys x 64 d, rows of 128 bytes, 128-byte swizzle): same geometry as the K-major
// tile (8-row atoms of 1024 bytes); the "major" bit lives in the instruction descriptor.
__device__ __forceinline__ uint32_t make_idesc_pv() {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) /* B is MN-major */ | ((uint32_t)(FA_D >> 3) << 17) |
           ((uint32_t)(FA_BM >> 4) << 24);
}

#ifndef GVD_HOST_EMU
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
#else
inline void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    tc::emu_umma_bf16_ts(tmem_d, tmem_a, bdesc, idesc, accumulate);
}
inline void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) { tc::emu_tmem_st16(taddr, v); }
inline void tmem_st_wait() {}
#endif

struct FaParams {
    __nv_bfloat16* out;
    long long ldo, o_stride_h, o_stride_b;
    int Nq, Nk, H;
    float scale;
    float* lse;  // optional [B, H, lse_ld]: log2 sum_j exp2(s_ij * scale * log2 e) per query row (generation 7 only)
    int lse_ld;
};

__global__ void __launch_bounds__(FA_THREADS, 2)
flash_attn_v1_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                  const __grid_constant__ CUtensorMap tmap_v, FaParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sq = smem;
    uint8_t* sk = smem + FA_TILE_BYTES;
    uint8_t* sv = smem + FA_TILE_BYTES * (1 + FA_STAGES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FA_TILE_BYTES * (1 + 2 * FA_STAGES));
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;              // [2]
    uint64_t* kv_empty = bars + 3;             // [2]
    uint64_t* s_full = bars + 5;
    uint64_t* p_full = bars + 6;
    uint64_t* o_full = bars + 7;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * FA_BM, h = blockIdx.y, b = blockIdx.z;
    const int nblk = (p.Nk + FA_BN - 1) / FA_BN;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmap_q);
        tc::prefetch_tmap(&tmap_k);
        tc::prefetch_tmap(&tmap_v);
        tc::mbar_init(q_full, 1);
        for (int s = 0; s < FA_STAGES; ++s) {
            tc::mbar_init(&kv_full[s], 1);
            tc::mbar_init(&kv_empty[s], 1);
        }
        tc::mbar_init(s_full, 1);
        tc::mbar_init(p_full, 4);  // one arrival per softmax warp
        tc::mbar_init(o_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tmem_ptr, FA_TMEM_COLS);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            tc::mbar_expect_tx(q_full, FA_TILE_BYTES);
            tc::tma_load_4d(sq, &tmap_q, q_full, 0, m0, h, b);
            for (int j = 0; j < nblk; ++j) {
                const int s = j % FA_STAGES;
                tc::mbar_wait(&kv_empty[s], (uint32_t)(((j / FA_STAGES) & 1) ^ 1));
                tc::mbar_expect_tx(&kv_full[s], 2 * FA_TILE_BYTES);
                tc::tma_load_4d(sk + s * FA_TILE_BYTES, &tmap_k, &kv_full[s], 0, j * FA_BN, h, b);
                tc::tma_load_4d(sv + s * FA_TILE_BYTES, &tmap_v, &kv_full[s], 0, j * FA_BN, h, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_qk = tc::make_idesc_bf16(FA_BM, FA_BN);
            const uint32_t idesc_pv = make_idesc_pv();
            const uint32_t q_addr = tc::smem_u32(sq);
            tc::mbar_wait(q_full, 0);
            for (int j = 0; j < nblk; ++j) {
                const int s = j % FA_STAGES;
                tc::mbar_wait(&kv_full[s], (uint32_t)((j / FA_STAGES) & 1));
                // S_j may overwrite S_{j-1} only after the softmax warps have read it: p_full(j-1) was waited on below
                tc::fence_after_sync();
                const uint32_t k_addr = tc::smem_u32(sk + s * FA_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < FA_D / 16; ++k)
                    tc::umma_bf16(tmem_base + FA_S_COL, tc::make_desc_kmajor_sw128(q_addr + k * 32),
                                  tc::make_desc_kmajor_sw128(k_addr + k * 32), idesc_qk, k != 0);
                tc::umma_commit(s_full);
                // O_j = P_j V_j once the softmax warps have written P_j
                tc::mbar_wait(p_full, (uint32_t)(j & 1));
                tc::fence_after_sync();
                const uint32_t v_addr = tc::smem_u32(sv + s * FA_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < FA_BN / 16; ++k)  // 16 keys per MMA: 8 packed-bf16 TMEM columns of P, 16 smem rows of V
                    umma_bf16_ts(tmem_base + FA_O_COL, tmem_base + FA_P_COL + k * 8,
                                 tc::make_desc_kmajor_sw128(v_addr + k * 16 * 128), idesc_pv, k != 0);
                tc::umma_commit(o_full);
                tc::umma_commit(&kv_empty[s]);
            }
        }
    } else {
        const int q = warp & 3;  // TMEM lane quadrant of this warp
        const int row = m0 + q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const float L2E = 1.4426950408889634f;
        float m_run = -INFINITY, l_run = 0.f;
        float o[FA_D];
#pragma unroll
        for (int c = 0; c < FA_D; ++c) o[c] = 0.f;
        for (int j = 0; j < nblk; ++j) {
            tc::mbar_wait(s_full, (uint32_t)(j & 1));
            tc::fence_after_sync();
            const int key0 = j * FA_BN;
            const bool tail = key0 + FA_BN > p.Nk;  // only the last key block can hold out-of-range keys
            // ---- pass 1: row max of the raw scores (scale > 0, so the max commutes with the scaling) ----
            float mx = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < FA_BN; c += 32) {
                uint32_t v[32];
                tc::tmem_ld32(tmem_base + lane_off + FA_S_COL + c, v);
                tc::tmem_ld_wait();
                if (!tail) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) mx = fmaxf(mx, __uint_as_float(v[e]));
                } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e) mx = fmaxf(mx, (key0 + c + e < p.Nk) ? __uint_as_float(v[e]) : -INFINITY);
                }
            }
            const float sl2 = p.scale * L2E;          // exp(x*scale - m) = exp2(x*sl2 - m*L2E), m kept in scaled units
            const float m_new = fmaxf(m_run, mx * p.scale);
            const float alpha = (m_run == -INFINITY) ? 0.f : ex2(( m_run - m_new) * L2E);
            const float mneg = -m_new * L2E;
            // ---- pass 2: P = exp(x - m_new) -> bf16 pairs -> TMEM ----
            float l_blk = 0.f;
#pragma unroll 1
            for (int c = 0; c < FA_BN; c += 32) {
                uint32_t v[32], pk[16];
                tc::tmem_ld32(tmem_base + lane_off + FA_S_COL + c, v);
                tc::tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 32; e += 2) {
                    float p0 = ex2(fmaf(__uint_as_float(v[e]), sl2, mneg));
                    float p1 = ex2(fmaf(__uint_as_float(v[e + 1]), sl2, mneg));
                    if (tail) {
                        if (key0 + c + e >= p.Nk) p0 = 0.f;
                        if (key0 + c + e + 1 >= p.Nk) p1 = 0.f;
                    }
                    l_blk += p0 + p1;
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);
                    pk[e / 2] = *reinterpret_cast<uint32_t*>(&h2);
                }
                tmem_st16(tmem_base + lane_off + FA_P_COL + c / 2, pk);
            }
            tmem_st_wait();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(p_full);
            l_run = l_run * alpha + l_blk;
            m_run = m_new;
            // ---- fold in O_j ----
            tc::mbar_wait(o_full, (uint32_t)(j & 1));
            tc::fence_after_sync();
#pragma unroll
            for (int c = 0; c < FA_D; c += 32) {
                uint32_t v[32];
                tc::tmem_ld32(tmem_base + lane_off + FA_O_COL + c, v);
                tc::tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 32; ++e) o[c + e] = fmaf(o[c + e], alpha, __uint_as_float(v[e]));
            }
        }
        if (row < p.Nq) {
            const float inv = 1.0f / l_run;
            __nv_bfloat16* dst = p.out + (long long)b * p.o_stride_b + (long long)h * p.o_stride_h + (long long)row * p.ldo;
#pragma unroll
            for (int c = 0; c < FA_D; c += 8) {
                uint4 u;
                __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(o[c + 2 * e] * inv, o[c + 2 * e + 1] * inv);
                *reinterpret_cast<uint4*>(dst + c) = u;
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, FA_TMEM_COLS);
}


// 2^x for x <= ~8 on the FMA pipe: x = n + f, n = round(x), f in [-0.5, 0.5]; 2^f by a cubic (max rel. error 1.1e-4,
// below the 2^-9 rounding of the bf16 P it feeds); 2^n by adding n to the exponent field.
__device__ __forceinline__ float ex2_fma(float x) {
    x = fmaxf(x, -126.0f);
    const float t = x + 12582912.0f;                 // 1.5 * 2^23: the low mantissa bits of t now hold n
    const float f = x - (t - 12582912.0f);
    float p = fmaf(f, 0.05550410866f, 0.24022650696f);
    p = fmaf(p, f, 0.69314718056f);
    p = fmaf(p, f, 1.0f);
    return __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));
}

// Softmax side of one 128-row query tile (4 warps, thread = query row): see the header comment.  s_col / p_col / o_col
// are the TMEM columns of this tile's S, P and O; `row` is the thread's global query row.
template <bool POLY>
__device__ __forceinline__ void softmax_tile(const FaParams& p, uint32_t tmem_base, uint32_t lane_off, uint32_t s_col,
                                             uint32_t p_col, uint32_t o_col, uint64_t* s_full, uint64_t* p_full,
                                             uint64_t* o_full, int nblk, int row, int lane, int b, int h) {
        const float sl2 = p.scale * 1.4426950408889634f;  // exp(x*scale - m) = exp2(x*sl2 - m2), m2 in log2 units
        float m_run = -INFINITY, l_run = 0.f;
        for (int j = 0; j < nblk; ++j) {
            tc::mbar_wait(s_full, (uint32_t)(j & 1));
            tc::fence_after_sync();
            const int key0 = j * FA_BN;
            uint32_t v[FA_BN];
#pragma unroll
            for (int c = 0; c < FA_BN; c += 32)
                tc::tmem_ld32(tmem_base + lane_off + s_col + c, *reinterpret_cast<uint32_t(*)[32]>(&v[c]));
            tc::tmem_ld_wait();
            if (key0 + FA_BN > p.Nk) {  // only the last key block can hold out-of-range keys
#pragma unroll
                for (int e = 0; e < FA_BN; ++e)
                    if (key0 + e >= p.Nk) v[e] = 0xff800000u;  // -inf
            }
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int e = 0; e < FA_BN; e += 4) {
                mx0 = fmaxf(mx0, __uint_as_float(v[e]));
                mx1 = fmaxf(mx1, __uint_as_float(v[e + 1]));
                mx2 = fmaxf(mx2, __uint_as_float(v[e + 2]));
                mx3 = fmaxf(mx3, __uint_as_float(v[e + 3]));
            }
            const float m_blk = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sl2;  // scale > 0: max commutes with it
            float m_new = m_run;
            if (j == 0) {
                m_new = m_blk;
            } else {
                const bool grow = m_blk > m_run + 8.0f;
                if (__any_sync(0xffffffffu, grow)) {  // warp-uniform: tcgen05.ld/st are warp-collective
                    float alpha = 1.0f;
                    if (grow) {
                        m_new = m_blk;
                        alpha = ex2(m_run - m_new);
                        l_run *= alpha;
                    }
#pragma unroll
                    for (int c = 0; c < FA_D; c += 16) {
                        uint32_t o[16];
                        tc::tmem_ld16(tmem_base + lane_off + o_col + c, o);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
                        tmem_st16(tmem_base + lane_off + o_col + c, o);
                    }
                }
            }
            const float mneg = -m_new;
            float l0 = 0.f, l1 = 0.f;
#pragma unroll
            for (int c = 0; c < FA_BN; c += 32) {
                uint32_t pk[16];
#pragma unroll
                for (int e = 0; e < 32; e += 2) {
                    const float x0 = fmaf(__uint_as_float(v[c + e]), sl2, mneg);
                    const float x1 = fmaf(__uint_as_float(v[c + e + 1]), sl2, mneg);
                    const float p0 = ex2(x0);
                    const float p1 = (POLY && (e & 2)) ? ex2_fma(x1) : ex2(x1);
                    l0 += p0;
                    l1 += p1;
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);
                    pk[e / 2] = *reinterpret_cast<uint32_t*>(&h2);
                }
                tmem_st16(tmem_base + lane_off + p_col + c / 2, pk);
            }
            tmem_st_wait();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(p_full);
            l_run += l0 + l1;
            m_run = m_new;
        }
        tc::mbar_wait(o_full, 0);
        tc::fence_after_sync();
        const float inv = 1.0f / l_run;
        __nv_bfloat16* dst = p.out + (long long)b * p.o_stride_b + (long long)h * p.o_stride_h + (long long)row * p.ldo;
#pragma unroll
        for (int c = 0; c < FA_D; c += 32) {
            uint32_t o[32];
            tc::tmem_ld32(tmem_base + lane_off + o_col + c, o);
            tc::tmem_ld_wait();
            if (row < p.Nq) {
#pragma unroll
                for (int e8 = 0; e8 < 32; e8 += 8) {
                    uint4 u;
                    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        h2[e] = __floats2bfloat162_rn(__uint_as_float(o[e8 + 2 * e]) * inv, __uint_as_float(o[e8 + 2 * e + 1]) * inv);
                    *reinterpret_cast<uint4*>(dst + c + e8) = u;
                }
            }
        }
    }

// Generation 6 softmax: ONE pass over S per key block, chunk-pipelined.
// ncu + tools/ubench_softmax_pipe.cu explain generations 2-5: a 128 x 128 block costs 1 024 clk of MUFU (16 ex2 / clk / SM)
// AND ~1 024 clk of TMEM reads (64 KB of fp32 scores at the 64 B / clk the tensor memory delivers) against 512 clk of
// MMA, and every earlier generation serialises the two: all of a thread's scores are loaded (and scanned for the row
// maximum) before its first exponential.  Here the exponentials of a block run against the STALE running maximum, so
// no separate maximum pass exists: chunk c + 1 (32 columns) is in flight from TMEM while chunk c is exponentiated, and the
// block's own maximum is tracked on the side.  Only if it exceeds the running maximum by more than 2^8 (the bound that
// keeps P inside bf16 / fp32 range; the first blocks of a row, then almost never) the warp redoes the block against the new
// maximum after rescaling O and l.  The first block has no running maximum and takes a maximum pass first.
__device__ __forceinline__ void softmax_tile_onepass(const FaParams& p, uint32_t tmem_base, uint32_t lane_off, uint32_t s_col,
                                                     uint32_t p_col, uint32_t o_col, uint64_t* s_full, uint64_t* p_full,
                                                     uint64_t* o_full, int nblk, int row, int lane, int b, int h) {
    const float sl2 = p.scale * 1.4426950408889634f;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nblk; ++j) {
        tc::mbar_wait(s_full, (uint32_t)(j & 1));
        tc::fence_after_sync();
        const int key0 = j * FA_BN;
        const bool tail = key0 + FA_BN > p.Nk;
        if (j == 0) {  // no running maximum yet: one maximum pass over the first block
            float mx = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < FA_BN; c += 32) {
                uint32_t v[32];
                tc::tmem_ld32(tmem_base + lane_off + s_col + c, v);
                tc::tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 32; ++e) mx = fmaxf(mx, (!tail || key0 + c + e < p.Nk) ? __uint_as_float(v[e]) : -INFINITY);
            }
            m_run = mx * sl2;
        }
        float m_blk, l_blk;
        // exponentials of the block against `mref`; returns the block's own maximum (log2 units) and row sum
        auto pass = [&](float mref) {
            const float mneg = -mref;
            float mx0 = -INFINITY, mx1 = -INFINITY, l0 = 0.f, l1 = 0.f;
            uint32_t va[32], vb[32];
            tc::tmem_ld32(tmem_base + lane_off + s_col, va);
#pragma unroll
            for (int c = 0; c < FA_BN; c += 32) {
                uint32_t(&cur)[32] = ((c >> 5) & 1) ? vb : va;
                uint32_t(&nxt)[32] = ((c >> 5) & 1) ? va : vb;
                tc::tmem_ld_wait();                                                        // chunk c has landed
                if (c + 32 < FA_BN) tc::tmem_ld32(tmem_base + lane_off + s_col + c + 32, nxt);  // chunk c + 1 in flight
                uint32_t pk[16];
#pragma unroll
                for (int e = 0; e < 32; e += 2) {
                    float x0 = __uint_as_float(cur[e]), x1 = __uint_as_float(cur[e + 1]);
                    if (tail) {
                        if (key0 + c + e >= p.Nk) x0 = -INFINITY;
                        if (key0 + c + e + 1 >= p.Nk) x1 = -INFINITY;
                    }
                    mx0 = fmaxf(mx0, x0);
                    mx1 = fmaxf(mx1, x1);
                    const float p0 = ex2(fmaf(x0, sl2, mneg));
                    const float p1 = ex2(fmaf(x1, sl2, mneg));
                    l0 += p0;
                    l1 += p1;
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);
                    pk[e / 2] = *reinterpret_cast<uint32_t*>(&h2);
                }
                tmem_st16(tmem_base + lane_off + p_col + c / 2, pk);
            }
            m_blk = fmaxf(mx0, mx1) * sl2;
            l_blk = l0 + l1;
        };
        pass(m_run);
        const bool grow = m_blk > m_run + 8.0f;
        if (__any_sync(0xffffffffu, grow)) {  // warp-uniform: tcgen05.ld / st are warp-collective
            float alpha = 1.0f;
            if (grow) {
                alpha = ex2(m_run - m_blk);
                m_run = m_blk;
                l_run *= alpha;
            }
            if (j > 0) {  // s_full(j) certified PV_{j-1}: O is quiescent
#pragma unroll
                for (int c = 0; c < FA_D; c += 16) {
                    uint32_t o[16];
                    tc::tmem_ld16(tmem_base + lane_off + o_col + c, o);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
                    tmem_st16(tmem_base + lane_off + o_col + c, o);
                }
            }
            pass(m_run);  // rows that did not grow recompute the same values
        }
        tmem_st_wait();
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(p_full);
        l_run += l_blk;
    }
    tc::mbar_wait(o_full, 0);
    tc::fence_after_sync();
    const float inv = 1.0f / l_run;
    __nv_bfloat16* dst = p.out + (long long)b * p.o_stride_b + (long long)h * p.o_stride_h + (long long)row * p.ldo;
#pragma unroll
    for (int c = 0; c < FA_D; c += 32) {
        uint32_t o[32];
        tc::tmem_ld32(tmem_base + lane_off + o_col + c, o);
        tc::tmem_ld_wait();
        if (row < p.Nq) {
#pragma unroll
            for (int e8 = 0; e8 < 32; e8 += 8) {
                uint4 u;
                __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    h2[e] = __floats2bfloat162_rn(__uint_as_float(o[e8 + 2 * e]) * inv, __uint_as_float(o[e8 + 2 * e + 1]) * inv);
                *reinterpret_cast<uint4*>(dst + c + e8) = u;
            }
        }
    }
}

template <int MODE>  // 0: generation 2; 1: with one exponential in four on the FMA pipe; 2: generation 6 (one pass)
__global__ void __launch_bounds__(FA_THREADS, 2)
flash_attn_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                  const __grid_constant__ CUtensorMap tmap_v, FaParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sq = smem;
    uint8_t* sk = smem + FA_TILE_BYTES;
    uint8_t* sv = smem + FA_TILE_BYTES * (1 + FA_STAGES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FA_TILE_BYTES * (1 + 2 * FA_STAGES));
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;              // [2]
    uint64_t* kv_empty = bars + 3;             // [2]
    uint64_t* s_full = bars + 5;
    uint64_t* p_full = bars + 6;
    uint64_t* o_full = bars + 7;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * FA_BM, h = blockIdx.y, b = blockIdx.z;
    const int nblk = (p.Nk + FA_BN - 1) / FA_BN;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmap_q);
        tc::prefetch_tmap(&tmap_k);
        tc::prefetch_tmap(&tmap_v);
        tc::mbar_init(q_full, 1);
        for (int s = 0; s < FA_STAGES; ++s) {
            tc::mbar_init(&kv_full[s], 1);
            tc::mbar_init(&kv_empty[s], 1);
        }
        tc::mbar_init(s_full, 1);
        tc::mbar_init(p_full, 4);  // one arrival per softmax warp
        tc::mbar_init(o_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tmem_ptr, FA_TMEM_COLS);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            tc::mbar_expect_tx(q_full, FA_TILE_BYTES);
            tc::tma_load_4d(sq, &tmap_q, q_full, 0, m0, h, b);
            for (int j = 0; j < nblk; ++j) {
                const int s = j % FA_STAGES;
                tc::mbar_wait(&kv_empty[s], (uint32_t)(((j / FA_STAGES) & 1) ^ 1));
                tc::mbar_expect_tx(&kv_full[s], 2 * FA_TILE_BYTES);
                tc::tma_load_4d(sk + s * FA_TILE_BYTES, &tmap_k, &kv_full[s], 0, j * FA_BN, h, b);
                tc::tma_load_4d(sv + s * FA_TILE_BYTES, &tmap_v, &kv_full[s], 0, j * FA_BN, h, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_qk = tc::make_idesc_bf16(FA_BM, FA_BN);
            const uint32_t idesc_pv = make_idesc_pv();
            const uint32_t q_addr = tc::smem_u32(sq);
            tc::mbar_wait(q_full, 0);
            for (int j = 0; j < nblk; ++j) {
                const int s = j % FA_STAGES;
                tc::mbar_wait(&kv_full[s], (uint32_t)((j / FA_STAGES) & 1));
                // S_j may overwrite S_{j-1}: the softmax warps read it before they arrived on p_full(j-1), waited below
                tc::fence_after_sync();
                const uint32_t k_addr = tc::smem_u32(sk + s * FA_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < FA_D / 16; ++k)
                    tc::umma_bf16(tmem_base + FA_S_COL, tc::make_desc_kmajor_sw128(q_addr + k * 32),
                                  tc::make_desc_kmajor_sw128(k_addr + k * 32), idesc_qk, k != 0);
                // s_full(j) also certifies that PV_{j-1} (issued earlier by this thread) is complete: O is quiescent
                // while the softmax warps rescale it
                tc::umma_commit(s_full);
                tc::mbar_wait(p_full, (uint32_t)(j & 1));
                tc::fence_after_sync();
                const uint32_t v_addr = tc::smem_u32(sv + s * FA_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < FA_BN / 16; ++k)  // 16 keys per MMA: 8 packed-bf16 TMEM columns of P, 16 smem rows of V
                    umma_bf16_ts(tmem_base + FA_O_COL, tmem_base + FA_P_COL + k * 8,
                                 tc::make_desc_kmajor_sw128(v_addr + k * 16 * 128), idesc_pv, (j | k) != 0);
                tc::umma_commit(&kv_empty[s]);
            }
            tc::umma_commit(o_full);
        }
    } else {
        const int q = warp & 3;  // TMEM lane quadrant of this warp
        if (MODE == 2)
            softmax_tile_onepass(p, tmem_base, (uint32_t)(q * 32) << 16, FA_S_COL, FA_P_COL, FA_O_COL, s_full, p_full, o_full, nblk,
                                 m0 + q * 32 + lane, lane, b, h);
        else
            softmax_tile<MODE == 1>(p, tmem_base, (uint32_t)(q * 32) << 16, FA_S_COL, FA_P_COL, FA_O_COL, s_full, p_full, o_full, nblk,
                                    m0 + q * 32 + lane, lane, b, h);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, FA_TMEM_COLS);
}


// ---- generation 7: key blocks split in two, S and P double-buffered, QK issued one sub-block ahead -------------------------
// ncu source view of generation 2 (profiles/r02_ncu_flash_v2_stalls.txt): the hottest line of the kernel, 26 % of all
// warp samples, is the softmax warps' wait for s_full.  The chain  softmax_j -> PV_j -> QK_{j+1} -> softmax_{j+1}  is
// serial, the MMAs' issue-to-commit latency sits on it twice per key block, and the second CTA of the SM does not fill the
// gaps (MUFU 59 % busy).  Here a 128-key tile is processed as two 64-key sub-blocks with two S buffers (64 columns each)
// and two P buffers: the MMA thread issues QK of sub-block i + 1 BEFORE it waits for the probabilities of sub-block i, so
// while the softmax warps exponentiate one half, the tensor pipe computes the scores of the next and the PV of the
// previous.  Same TMEM budget as generation 2 (S0 S1 | P0 P1 | O = 64 + 64 + 32 + 32 + 64 = 256 columns, two CTAs per SM), same
// shared-memory tiles (the sub-blocks are the two halves of the 128-key K / V tiles), one softmax thread per row.
// PV_{i-1} is no longer certified by s_full(i), so the lazy O rescale waits on o_done (only when a row's maximum grew).
constexpr uint32_t FA7_S_COL = 0, FA7_P_COL = 128, FA7_O_COL = 192;

__device__ __forceinline__ uint32_t make_idesc_pv_k64() { return make_idesc_pv(); }  // same shape: M 128, N 64 (d); K per MMA is 16

template <int POLY>  // 0: every exponential on the MUFU; n > 0: one in 2 n on the FMA pipe (Cody-Waite + cubic, ex2_fma)
__global__ void __launch_bounds__(FA_THREADS, 2)
flash_attn7_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                   const __grid_constant__ CUtensorMap tmap_v, FaParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sq = smem;
    uint8_t* sk = smem + FA_TILE_BYTES;
    uint8_t* sv = smem + FA_TILE_BYTES * (1 + FA_STAGES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FA_TILE_BYTES * (1 + 2 * FA_STAGES));
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;    // [2]
    uint64_t* kv_empty = bars + 3;   // [2]
    uint64_t* s_full = bars + 5;     // [2]
    uint64_t* p_full = bars + 7;     // [2]
    uint64_t* o_done = bars + 9;     // one completion per PV_i
    uint64_t* o_final = bars + 10;   // completes once, after the last PV
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 11);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * FA_BM, h = blockIdx.y, b = blockIdx.z;
    const int nblk = (p.Nk + FA_BN - 1) / FA_BN;   // 128-key tiles in shared memory
    const int nsub = (p.Nk + 63) / 64;             // 64-key sub-blocks that hold at least one key

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmap_q);
        tc::prefetch_tmap(&tmap_k);
        tc::prefetch_tmap(&tmap_v);
        tc::mbar_init(q_full, 1);
        for (int s = 0; s < FA_STAGES; ++s) {
            tc::mbar_init(&kv_full[s], 1);
            tc::mbar_init(&kv_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&s_full[s], 1);
            tc::mbar_init(&p_full[s], 4);  // one arrival per softmax warp
        }
        tc::mbar_init(o_done, 1);
        tc::mbar_init(o_final, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tmem_ptr, FA_TMEM_COLS);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            tc::mbar_expect_tx(q_full, FA_TILE_BYTES);
            tc::tma_load_4d(sq, &tmap_q, q_full, 0, m0, h, b);
            for (int j = 0; j < nblk; ++j) {
                const int s = j % FA_STAGES;
                tc::mbar_wait(&kv_empty[s], (uint32_t)(((j / FA_STAGES) & 1) ^ 1));
                tc::mbar_expect_tx(&kv_full[s], 2 * FA_TILE_BYTES);
                tc::tma_load_4d(sk + s * FA_TILE_BYTES, &tmap_k, &kv_full[s], 0, j * FA_BN, h, b);
                tc::tma_load_4d(sv + s * FA_TILE_BYTES, &tmap_v, &kv_full[s], 0, j * FA_BN, h, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_qk = tc::make_idesc_bf16(FA_BM, 64);
            const uint32_t idesc_pv = make_idesc_pv_k64();
            const uint32_t q_addr = tc::smem_u32(sq);
            auto issue_qk = [&](int i) {  // S[i & 1] = Q K_i^T, K_i = rows 64 (i & 1) .. + 63 of tile i / 2
                const int j = i >> 1, s = j % FA_STAGES;
                if ((i & 1) == 0) {
                    tc::mbar_wait(&kv_full[s], (uint32_t)((j / FA_STAGES) & 1));
                    tc::fence_after_sync();
                }
                const uint32_t k_addr = tc::smem_u32(sk + s * FA_TILE_BYTES) + (uint32_t)(i & 1) * 64 * 128;
#pragma unroll
                for (int k = 0; k < FA_D / 16; ++k)
                    tc::umma_bf16(tmem_base + FA7_S_COL + (uint32_t)(i & 1) * 64, tc::make_desc_kmajor_sw128(q_addr + k * 32),
                                  tc::make_desc_kmajor_sw128(k_addr + k * 32), idesc_qk, k != 0);
                tc::umma_commit(&s_full[i & 1]);
            };
            tc::mbar_wait(q_full, 0);
            issue_qk(0);
            for (int i = 0; i < nsub; ++i) {
                // S[(i+1)&1] was last read by the softmax of sub-block i-1, which finished before p_full(i-1) completed --
                // waited on in the previous iteration
                if (i + 1 < nsub) issue_qk(i + 1);
                tc::mbar_wait(&p_full[i & 1], (uint32_t)((i >> 1) & 1));
                tc::fence_after_sync();
                const int j = i >> 1, s = j % FA_STAGES;
                const uint32_t v_addr = tc::smem_u32(sv + s * FA_TILE_BYTES) + (uint32_t)(i & 1) * 64 * 128;
#pragma unroll
                for (int k = 0; k < 64 / 16; ++k)  // 16 keys per MMA: 8 packed-bf16 TMEM columns of P, 16 smem rows of V
                    umma_bf16_ts(tmem_base + FA7_O_COL, tmem_base + FA7_P_COL + (uint32_t)(i & 1) * 32 + k * 8,
                                 tc::make_desc_kmajor_sw128(v_addr + k * 16 * 128), idesc_pv, (i | k) != 0);
                tc::umma_commit(o_done);
                if (i + 1 == nsub) tc::umma_commit(o_final);
                if ((i & 1) || i + 1 == nsub) tc::umma_commit(&kv_empty[s]);  // the tile's last sub-block: K_j, V_j have no reader left
            }
        }
    } else {
        const int q = warp & 3;  // TMEM lane quadrant of this warp
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int row = m0 + q * 32 + lane;
        const float sl2 = p.scale * 1.4426950408889634f;
        float m_run = -INFINITY, l_run = 0.f;
        for (int i = 0; i < nsub; ++i) {
            tc::mbar_wait(&s_full[i & 1], (uint32_t)((i >> 1) & 1));
            tc::fence_after_sync();
            const int key0 = i * 64;
            uint32_t v[64];
            tc::tmem_ld32(tmem_base + lane_off + FA7_S_COL + (uint32_t)(i & 1) * 64, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
            tc::tmem_ld32(tmem_base + lane_off + FA7_S_COL + (uint32_t)(i & 1) * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
            tc::tmem_ld_wait();
            if (key0 + 64 > p.Nk) {  // only the last sub-block can hold out-of-range keys
#pragma unroll
                for (int e = 0; e < 64; ++e)
                    if (key0 + e >= p.Nk) v[e] = 0xff800000u;  // -inf
            }
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int e = 0; e < 64; e += 4) {
                mx0 = fmaxf(mx0, __uint_as_float(v[e]));
                mx1 = fmaxf(mx1, __uint_as_float(v[e + 1]));
                mx2 = fmaxf(mx2, __uint_as_float(v[e + 2]));
                mx3 = fmaxf(mx3, __uint_as_float(v[e + 3]));
            }
            const float m_blk = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sl2;  // scale > 0: max commutes with it
            float m_new = m_run;
            if (i == 0) {
                m_new = m_blk;
            } else {
                const bool grow = m_blk > m_run + 8.0f;
                if (__any_sync(0xffffffffu, grow)) {  // warp-uniform: tcgen05.ld / st are warp-collective
                    tc::mbar_wait(o_done, (uint32_t)((i - 1) & 1));  // PV_{i-1}: O is quiescent
                    tc::fence_after_sync();
                    float alpha = 1.0f;
                    if (grow) {
                        m_new = m_blk;
                        alpha = ex2(m_run - m_new);
                        l_run *= alpha;
                    }
#pragma unroll
                    for (int c = 0; c < FA_D; c += 16) {
                        uint32_t o[16];
                        tc::tmem_ld16(tmem_base + lane_off + FA7_O_COL + c, o);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
                        tmem_st16(tmem_base + lane_off + FA7_O_COL + c, o);
                    }
                }
            }
            const float mneg = -m_new;
            float l0 = 0.f, l1 = 0.f;
#pragma unroll
            for (int c = 0; c < 64; c += 32) {
                uint32_t pk[16];
#pragma unroll
                for (int e = 0; e < 32; e += 2) {
                    const float x0 = fmaf(__uint_as_float(v[c + e]), sl2, mneg), x1 = fmaf(__uint_as_float(v[c + e + 1]), sl2, mneg);
                    const float p0 = ex2(x0);
                    const float p1 = (POLY > 0 && (e >> 1) % (POLY > 0 ? POLY : 1) == 0) ? ex2_fma(x1) : ex2(x1);  // one in 2 * POLY
                    l0 += p0;
                    l1 += p1;
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);
                    pk[e / 2] = *reinterpret_cast<uint32_t*>(&h2);
                }
                tmem_st16(tmem_base + lane_off + FA7_P_COL + (uint32_t)(i & 1) * 32 + c / 2, pk);
            }
            tmem_st_wait();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&p_full[i & 1]);
            l_run += l0 + l1;
            m_run = m_new;
        }
        // The end of the accumulation has its own single-use barrier.  o_done completes once per sub-block, and a parity
        // wait on it is only meaningful when the waiter is at most one phase behind: the last thing this thread waited on,
        // s_full(nsub - 1), certifies PV up to nsub - 3 only, so "parity of phase nsub - 1" could be answered by phase
        // nsub - 3 while PV_{nsub-2} was still in flight.  (It never showed in a timed run -- the softmax of the last
        // sub-block outlasts that MMA -- but a compute-sanitizer run, with its different timing, produced one wrong tile.)
#ifdef GVD_EMU_OLD_EPILOGUE_WAIT  // host tests only (tests/test_attn_fwd_emu_cpu.py): the former wait, to show the checker sees it
        tc::mbar_wait(o_done, (uint32_t)((nsub - 1) & 1));
#else
        tc::mbar_wait(o_final, 0);
#endif
        tc::fence_after_sync();
        const float inv = 1.0f / l_run;
        // the statistic the backward needs (attn_bwd_tc.cu); m_run may be a stale maximum, m_run + log2 l_run is exact either way
        if (p.lse != nullptr && row < p.lse_ld) p.lse[((long long)b * p.H + h) * p.lse_ld + row] = m_run + __log2f(l_run);
        __nv_bfloat16* dst = p.out + (long long)b * p.o_stride_b + (long long)h * p.o_stride_h + (long long)row * p.ldo;
#pragma unroll
        for (int c = 0; c < FA_D; c += 32) {
            uint32_t o[32];
            tc::tmem_ld32(tmem_base + lane_off + FA7_O_COL + c, o);
            tc::tmem_ld_wait();
            if (row < p.Nq) {
#pragma unroll
                for (int e8 = 0; e8 < 32; e8 += 8) {
                    uint4 u;
                    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        h2[e] = __floats2bfloat162_rn(__uint_as_float(o[e8 + 2 * e]) * inv, __uint_as_float(o[e8 + 2 * e + 1]) * inv);
                    *reinterpret_cast<uint4*>(dst + c + e8) = u;
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, FA_TMEM_COLS);
}

// Generation 8 = generation 7 with TWO softmax threads per query row (8 softmax warps per CTA, four per scheduler with the
// two CTAs of the SM): warps w and w + 4 share a TMEM lane quadrant and split each 64-key sub-block -- 32 score columns,
// 16 packed P columns and 32 of the 64 O columns each -- and agree on the row maximum through two floats of shared memory
// and a 64-thread named barrier per sub-block; the row sums stay separate until the end.  (The same split was slower when
// the serial MMA chain bounded the kernel -- see the generation 7 comment.)  Measured on B200 (tools/gpu_check_flash8.sh,
// 25 x 5 heads x 9216 tokens): 3.73 ms against generation 7's 3.36 ms -- the per-sub-block pair barrier and the shared
// memory round trip cost more than the extra warps per scheduler win back, so this stays an A/B variant (GVD_FLASH=v8).
constexpr int FA8_THREADS = 64 + 8 * 32;
constexpr int FA8_SMEM = FA_SMEM + 2 * 2 * 128 * 4;  // + pair exchange: [parity][half][row] floats

#ifndef GVD_HOST_EMU
__device__ __forceinline__ void pair_bar(int q) { asm volatile("bar.sync %0, 64;" ::"r"(q + 1) : "memory"); }
#else
inline void pair_bar(int) { std::abort(); }  // named barriers have no host stand-in: generations 5 and 8 are not run there
#endif

__global__ void __launch_bounds__(FA8_THREADS, 2)
flash_attn8_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                   const __grid_constant__ CUtensorMap tmap_v, FaParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sq = smem;
    uint8_t* sk = smem + FA_TILE_BYTES;
    uint8_t* sv = smem + FA_TILE_BYTES * (1 + FA_STAGES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FA_TILE_BYTES * (1 + 2 * FA_STAGES));
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;    // [2]
    uint64_t* kv_empty = bars + 3;   // [2]
    uint64_t* s_full = bars + 5;     // [2]
    uint64_t* p_full = bars + 7;     // [2]
    uint64_t* o_done = bars + 9;     // one completion per PV_i
    uint64_t* o_final = bars + 10;   // completes once, after the last PV
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 11);
    float* xchg = reinterpret_cast<float*>(smem + FA_TILE_BYTES * (1 + 2 * FA_STAGES) + 256);  // [2][2][128]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * FA_BM, h = blockIdx.y, b = blockIdx.z;
    const int nblk = (p.Nk + FA_BN - 1) / FA_BN;   // 128-key tiles in shared memory
    const int nsub = (p.Nk + 63) / 64;             // 64-key sub-blocks that hold at least one key

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmap_q);
        tc::prefetch_tmap(&tmap_k);
        tc::prefetch_tmap(&tmap_v);
        tc::mbar_init(q_full, 1);
        for (int s = 0; s < FA_STAGES; ++s) {
            tc::mbar_init(&kv_full[s], 1);
            tc::mbar_init(&kv_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&s_full[s], 1);
            tc::mbar_init(&p_full[s], 8);  // one arrival per softmax warp
        }
        tc::mbar_init(o_done, 1);
        tc::mbar_init(o_final, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tmem_ptr, FA_TMEM_COLS);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            tc::mbar_expect_tx(q_full, FA_TILE_BYTES);
            tc::tma_load_4d(sq, &tmap_q, q_full, 0, m0, h, b);
            for (int j = 0; j < nblk; ++j) {
                const int s = j % FA_STAGES;
                tc::mbar_wait(&kv_empty[s], (uint32_t)(((j / FA_STAGES) & 1) ^ 1));
                tc::mbar_expect_tx(&kv_full[s], 2 * FA_TILE_BYTES);
                tc::tma_load_4d(sk + s * FA_TILE_BYTES, &tmap_k, &kv_full[s], 0, j * FA_BN, h, b);
                tc::tma_load_4d(sv + s * FA_TILE_BYTES, &tmap_v, &kv_full[s], 0, j * FA_BN, h, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_qk = tc::make_idesc_bf16(FA_BM, 64);
            const uint32_t idesc_pv = make_idesc_pv_k64();
            const uint32_t q_addr = tc::smem_u32(sq);
            auto issue_qk = [&](int i) {  // S[i & 1] = Q K_i^T, K_i = rows 64 (i & 1) .. + 63 of tile i / 2
                const int j = i >> 1, s = j % FA_STAGES;
                if ((i & 1) == 0) {
                    tc::mbar_wait(&kv_full[s], (uint32_t)((j / FA_STAGES) & 1));
                    tc::fence_after_sync();
                }
                const uint32_t k_addr = tc::smem_u32(sk + s * FA_TILE_BYTES) + (uint32_t)(i & 1) * 64 * 128;
#pragma unroll
                for (int k = 0; k < FA_D / 16; ++k)
                    tc::umma_bf16(tmem_base + FA7_S_COL + (uint32_t)(i & 1) * 64, tc::make_desc_kmajor_sw128(q_addr + k * 32),
                                  tc::make_desc_kmajor_sw128(k_addr + k * 32), idesc_qk, k != 0);
                tc::umma_commit(&s_full[i & 1]);
            };
            tc::mbar_wait(q_full, 0);
            issue_qk(0);
            for (int i = 0; i < nsub; ++i) {
                // S[(i+1)&1] was last read by the softmax of sub-block i-1, which finished before p_full(i-1) completed --
                // waited on in the previous iteration
                if (i + 1 < nsub) issue_qk(i + 1);
                tc::mbar_wait(&p_full[i & 1], (uint32_t)((i >> 1) & 1));
                tc::fence_after_sync();
                const int j = i >> 1, s = j % FA_STAGES;
                const uint32_t v_addr = tc::smem_u32(sv + s * FA_TILE_BYTES) + (uint32_t)(i & 1) * 64 * 128;
#pragma unroll
                for (int k = 0; k < 64 / 16; ++k)  // 16 keys per MMA: 8 packed-bf16 TMEM columns of P, 16 smem rows of V
                    umma_bf16_ts(tmem_base + FA7_O_COL, tmem_base + FA7_P_COL + (uint32_t)(i & 1) * 32 + k * 8,
                                 tc::make_desc_kmajor_sw128(v_addr + k * 16 * 128), idesc_pv, (i | k) != 0);
                tc::umma_commit(o_done);
                if (i + 1 == nsub) tc::umma_commit(o_final);
                if ((i & 1) || i + 1 == nsub) tc::umma_commit(&kv_empty[s]);  // the tile's last sub-block: K_j, V_j have no reader left
            }
        }
    } else {
        const int q = warp & 3;             // TMEM lane quadrant of this warp
        const int hf = (warp - 2) >> 2;     // which half of the sub-block's columns
        const int rloc = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int row = m0 + rloc;
        const float sl2 = p.scale * 1.4426950408889634f;
        float m_run = -INFINITY, l_run = 0.f;
        for (int i = 0; i < nsub; ++i) {
            tc::mbar_wait(&s_full[i & 1], (uint32_t)((i >> 1) & 1));
            tc::fence_after_sync();
            const int key0 = i * 64 + 32 * hf;
            uint32_t v[32];
            tc::tmem_ld32(tmem_base + lane_off + FA7_S_COL + (uint32_t)(i & 1) * 64 + 32 * hf, v);
            tc::tmem_ld_wait();
            if (key0 + 32 > p.Nk) {
#pragma unroll
                for (int e = 0; e < 32; ++e)
                    if (key0 + e >= p.Nk) v[e] = 0xff800000u;  // -inf
            }
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
                mx0 = fmaxf(mx0, __uint_as_float(v[e]));
                mx1 = fmaxf(mx1, __uint_as_float(v[e + 1]));
                mx2 = fmaxf(mx2, __uint_as_float(v[e + 2]));
                mx3 = fmaxf(mx3, __uint_as_float(v[e + 3]));
            }
            const float m_loc = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
            float* xb = xchg + (i & 1) * 256;
            xb[hf * 128 + rloc] = m_loc;
            pair_bar(q);
            const float m_blk = fmaxf(m_loc, xb[(hf ^ 1) * 128 + rloc]) * sl2;  // scale > 0: max commutes with it
            float m_new = m_run;
            if (i == 0) {
                m_new = m_blk;
            } else {
                const bool grow = m_blk > m_run + 8.0f;  // both threads of a row see the same m_blk and m_run
                if (__any_sync(0xffffffffu, grow)) {
                    tc::mbar_wait(o_done, (uint32_t)((i - 1) & 1));  // PV_{i-1}: O is quiescent
                    tc::fence_after_sync();
                    float alpha = 1.0f;
                    if (grow) {
                        m_new = m_blk;
                        alpha = ex2(m_run - m_new);
                        l_run *= alpha;
                    }
#pragma unroll
                    for (int c = 0; c < 32; c += 16) {
                        uint32_t o[16];
                        tc::tmem_ld16(tmem_base + lane_off + FA7_O_COL + 32 * hf + c, o);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
                        tmem_st16(tmem_base + lane_off + FA7_O_COL + 32 * hf + c, o);
                    }
                }
            }
            const float mneg = -m_new;
            float l0 = 0.f, l1 = 0.f;
            uint32_t pk[16];
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
                const float p0 = ex2(fmaf(__uint_as_float(v[e]), sl2, mneg));
                const float p1 = ex2(fmaf(__uint_as_float(v[e + 1]), sl2, mneg));
                l0 += p0;
                l1 += p1;
                __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);
                pk[e / 2] = *reinterpret_cast<uint32_t*>(&h2);
            }
            tmem_st16(tmem_base + lane_off + FA7_P_COL + (uint32_t)(i & 1) * 32 + 16 * hf, pk);
            tmem_st_wait();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&p_full[i & 1]);
            l_run += l0 + l1;
            m_run = m_new;
        }
        float* xb = xchg + (nsub & 1) * 256;  // the two halves of a row ran with the same maximum: their sums add
        xb[hf * 128 + rloc] = l_run;
        pair_bar(q);
        const float inv = 1.0f / (l_run + xb[(hf ^ 1) * 128 + rloc]);
        // The end of the accumulation has its own single-use barrier.  o_done completes once per sub-block, and a parity
        // wait on it is only meaningful when the waiter is at most one phase behind: the last thing this thread waited on,
        // s_full(nsub - 1), certifies PV up to nsub - 3 only, so "parity of phase nsub - 1" could be answered by phase
        // nsub - 3 while PV_{nsub-2} was still in flight.  (It never showed in a timed run -- the softmax of the last
        // sub-block outlasts that MMA -- but a compute-sanitizer run, with its different timing, produced one wrong tile.)
        tc::mbar_wait(o_final, 0);
        tc::fence_after_sync();
        __nv_bfloat16* dst = p.out + (long long)b * p.o_stride_b + (long long)h * p.o_stride_h + (long long)row * p.ldo + 32 * hf;
        uint32_t o[32];
        tc::tmem_ld32(tmem_base + lane_off + FA7_O_COL + 32 * hf, o);
        tc::tmem_ld_wait();
        if (row < p.Nq) {
#pragma unroll
            for (int e8 = 0; e8 < 32; e8 += 8) {
                uint4 u;
                __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    h2[e] = __floats2bfloat162_rn(__uint_as_float(o[e8 + 2 * e]) * inv, __uint_as_float(o[e8 + 2 * e + 1]) * inv);
                *reinterpret_cast<uint4*>(dst + e8) = u;
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, FA_TMEM_COLS);
}

// ---- generation 5: one CTA per SM, TWO query tiles in ping-pong, two softmax threads per row --------------------------
// tools/ubench_softmax_pipe.cu (profiles/r02_ubench_softmax.txt): MUFU.EX2 runs at 16 results / clk / SM, so a 128 x 128
// block holds 1 024 clk of exponentials against 512 clk of MMA -- with head dim 64 the tensor pipe cannot exceed 50 %
// unless exponentials leave the MUFU -- and the per-score instruction mix sustains 7 / 12 / 16 results per clock per SM
// with one / two / four softmax warps per scheduler.  Two intermediate designs were measured on B200 (N = 9216, 5 heads,
// 25 frames; generation 2: 4.13 ms) and removed: 2 CTAs x 8 softmax warps, two threads per row (4.45 ms), and one
// software-pipelined CTA with double-buffered S and 16 softmax warps, four threads per row (4.46 ms).  In both, all
// softmax warps of a query tile move in lockstep (wait for S, TMEM load, row maximum, exchange, exponentials, TMEM store,
// hand-over), so the latencies around the exponentials are exposed on every scheduler at once.  This generation is the
// ping-pong form of the same idea (the structure FlashAttention-4 uses at head dim 128): one CTA owns two 128-row query tiles A
// and B that share every K / V tile (fetched once, used by four MMAs): while A's warps are in their load / max / store
// phases, B's are in their exponentials and vice versa, and the tensor pipe works for one tile while the MUFU works for
// the other.  8 softmax warps per tile (two threads per row: 64 score columns each), so every scheduler holds two warps
// of A and two of B.  MMA order per key block: PV_A, QK_A(next), PV_B, QK_B(next) -- each tile's next scores are
// issued right behind its PV, as early as its S columns are free.
// TMEM: S_A S_B | P_A P_B | O_A O_B = 128 + 128 + 64 + 64 + 64 + 64 = 512 columns.
constexpr int FA5_THREADS = 64 + 16 * 32;
constexpr int FA5_STAGES = 3;
constexpr int FA5_BAR_OFF = FA_TILE_BYTES * (2 + 2 * FA5_STAGES);
constexpr int FA5_XCHG_OFF = FA5_BAR_OFF + 256;
constexpr int FA5_SMEM = FA5_XCHG_OFF + 2 * 2 * 2 * 128 * 4 + 1024;  // exchange: [tile][parity][half][row]

#ifndef GVD_HOST_EMU
__device__ __forceinline__ void pair_sync_id(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
#else
inline void pair_sync_id(int) { std::abort(); }
#endif

__global__ void __launch_bounds__(FA5_THREADS, 1)  // 18 warps are allocated as 20: 96 registers per thread
flash_attn5_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                   const __grid_constant__ CUtensorMap tmap_v, FaParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sq = smem;                                          // [2] query tiles
    uint8_t* sk = smem + 2 * FA_TILE_BYTES;                      // [3]
    uint8_t* sv = smem + FA_TILE_BYTES * (2 + FA5_STAGES);       // [3]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FA5_BAR_OFF);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;    // [3]
    uint64_t* kv_empty = bars + 4;   // [3]
    uint64_t* s_full = bars + 7;     // [2] per tile
    uint64_t* p_full = bars + 9;     // [2] per tile
    uint64_t* o_full = bars + 11;    // [2] per tile
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 13);
    float* xchg = reinterpret_cast<float*>(smem + FA5_XCHG_OFF);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * (2 * FA_BM), h = blockIdx.y, b = blockIdx.z;
    const int nblk = (p.Nk + FA_BN - 1) / FA_BN;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmap_q);
        tc::prefetch_tmap(&tmap_k);
        tc::prefetch_tmap(&tmap_v);
        tc::mbar_init(q_full, 1);
        for (int s = 0; s < FA5_STAGES; ++s) {
            tc::mbar_init(&kv_full[s], 1);
            tc::mbar_init(&kv_empty[s], 1);
        }
        for (int t = 0; t < 2; ++t) {
            tc::mbar_init(&s_full[t], 1);
            tc::mbar_init(&p_full[t], 8);  // the tile's 8 softmax warps
            tc::mbar_init(&o_full[t], 1);
        }
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tmem_ptr, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            tc::mbar_expect_tx(q_full, 2 * FA_TILE_BYTES);
            tc::tma_load_4d(sq, &tmap_q, q_full, 0, m0, h, b);
            tc::tma_load_4d(sq + FA_TILE_BYTES, &tmap_q, q_full, 0, m0 + FA_BM, h, b);
            for (int j = 0; j < nblk; ++j) {
                const int s = j % FA5_STAGES;
                tc::mbar_wait(&kv_empty[s], (uint32_t)(((j / FA5_STAGES) & 1) ^ 1));
                tc::mbar_expect_tx(&kv_full[s], 2 * FA_TILE_BYTES);
                tc::tma_load_4d(sk + s * FA_TILE_BYTES, &tmap_k, &kv_full[s], 0, j * FA_BN, h, b);
                tc::tma_load_4d(sv + s * FA_TILE_BYTES, &tmap_v, &kv_full[s], 0, j * FA_BN, h, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_qk = tc::make_idesc_bf16(FA_BM, FA_BN);
            const uint32_t idesc_pv = make_idesc_pv();
            const uint32_t q_addr = tc::smem_u32(sq);
            auto issue_qk = [&](int t, int j) {  // S_t = Q_t K_j^T
                const uint32_t k_addr = tc::smem_u32(sk + (j % FA5_STAGES) * FA_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < FA_D / 16; ++k)
                    tc::umma_bf16(tmem_base + (uint32_t)t * 128, tc::make_desc_kmajor_sw128(q_addr + t * FA_TILE_BYTES + k * 32),
                                  tc::make_desc_kmajor_sw128(k_addr + k * 32), idesc_qk, k != 0);
                tc::umma_commit(&s_full[t]);
            };
            tc::mbar_wait(q_full, 0);
            tc::mbar_wait(&kv_full[0], 0);
            tc::fence_after_sync();
            issue_qk(0, 0);
            issue_qk(1, 0);
            for (int j = 0; j < nblk; ++j) {
                const int s = j % FA5_STAGES;
                const uint32_t v_addr = tc::smem_u32(sv + s * FA_TILE_BYTES);
                for (int t = 0; t < 2; ++t) {
                    tc::mbar_wait(&p_full[t], (uint32_t)(j & 1));
                    tc::fence_after_sync();
#pragma unroll
                    for (int k = 0; k < FA_BN / 16; ++k)
                        umma_bf16_ts(tmem_base + 384 + (uint32_t)t * 64, tmem_base + 256 + (uint32_t)t * 64 + k * 8,
                                     tc::make_desc_kmajor_sw128(v_addr + k * 16 * 128), idesc_pv, (j | k) != 0);
                    if (t == 1) tc::umma_commit(&kv_empty[s]);  // K_j and V_j have no reader left
                    if (j + 1 < nblk) {
                        if (t == 0) {
                            tc::mbar_wait(&kv_full[(j + 1) % FA5_STAGES], (uint32_t)(((j + 1) / FA5_STAGES) & 1));
                            tc::fence_after_sync();
                        }
                        // S_t is free: the tile's softmax warps read it before they arrived on p_full(j); s_full(j+1) also
                        // certifies PV_t(j), so O_t is quiescent while they rescale it
                        issue_qk(t, j + 1);
                    } else {
                        tc::umma_commit(&o_full[t]);
                    }
                }
            }
        }
    } else {
        const int sw = warp - 2;
        const int t = sw >> 3;            // query tile
        const int q = warp & 3;           // TMEM lane quadrant (hardware rule: warp id % 4)
        const int hf = (sw >> 2) & 1;     // which half of the row's columns
        const int rloc = q * 32 + lane;
        const int row = m0 + t * FA_BM + rloc;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const uint32_t s_col = (uint32_t)t * 128 + 64 * hf, p_col = 256 + (uint32_t)t * 64 + 32 * hf, o_col = 384 + (uint32_t)t * 64 + 32 * hf;
        const int bar_id = 1 + t * 4 + q;
        float* xt = xchg + t * 512;
        const float sl2 = p.scale * 1.4426950408889634f;
        float m_run = -INFINITY, l_run = 0.f;
        for (int j = 0; j < nblk; ++j) {
            tc::mbar_wait(&s_full[t], (uint32_t)(j & 1));
            tc::fence_after_sync();
            const int key0 = j * FA_BN + 64 * hf;
            const bool tail = key0 + 64 > p.Nk;
            // pass 1: row maximum of this thread's 64 scores.  The scores are read from TMEM again for the exponentials
            // (S stays intact until p_full): holding all 64 would not fit the 96 registers 18 warps leave per thread.
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int c = 0; c < 64; c += 32) {
                uint32_t v[32];
                tc::tmem_ld32(tmem_base + lane_off + s_col + c, v);
                tc::tmem_ld_wait();
                if (tail) {
#pragma unroll
                    for (int e = 0; e < 32; ++e)
                        if (key0 + c + e >= p.Nk) v[e] = 0xff800000u;  // -inf
                }
#pragma unroll
                for (int e = 0; e < 32; e += 4) {
                    mx0 = fmaxf(mx0, __uint_as_float(v[e]));
                    mx1 = fmaxf(mx1, __uint_as_float(v[e + 1]));
                    mx2 = fmaxf(mx2, __uint_as_float(v[e + 2]));
                    mx3 = fmaxf(mx3, __uint_as_float(v[e + 3]));
                }
            }
            const float m_loc = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
            float* xb = xt + (j & 1) * 256;
            xb[hf * 128 + rloc] = m_loc;
            pair_sync_id(bar_id);
            const float m_blk = fmaxf(m_loc, xb[(hf ^ 1) * 128 + rloc]) * sl2;  // scale > 0: max commutes with it
            float m_new = m_run;
            if (j == 0) {
                m_new = m_blk;
            } else {
                const bool grow = m_blk > m_run + 8.0f;  // both threads of a row see the same m_blk and m_run
                if (__any_sync(0xffffffffu, grow)) {
                    float alpha = 1.0f;
                    if (grow) {
                        m_new = m_blk;
                        alpha = ex2(m_run - m_new);
                        l_run *= alpha;
                    }
#pragma unroll
                    for (int c = 0; c < 32; c += 16) {
                        uint32_t o[16];
                        tc::tmem_ld16(tmem_base + lane_off + o_col + c, o);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
                        tmem_st16(tmem_base + lane_off + o_col + c, o);
                    }
                }
            }
            const float mneg = -m_new;
            float l0 = 0.f, l1 = 0.f;
#pragma unroll
            for (int c = 0; c < 64; c += 32) {
                uint32_t v[32], pk[16];
                tc::tmem_ld32(tmem_base + lane_off + s_col + c, v);
                tc::tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 32; e += 2) {
                    float p0 = ex2(fmaf(__uint_as_float(v[e]), sl2, mneg));
                    float p1 = ex2(fmaf(__uint_as_float(v[e + 1]), sl2, mneg));
                    if (tail) {
                        if (key0 + c + e >= p.Nk) p0 = 0.f;
                        if (key0 + c + e + 1 >= p.Nk) p1 = 0.f;
                    }
                    l0 += p0;
                    l1 += p1;
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);
                    pk[e / 2] = *reinterpret_cast<uint32_t*>(&h2);
                }
                tmem_st16(tmem_base + lane_off + p_col + c / 2, pk);
            }
            tmem_st_wait();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&p_full[t]);
            l_run += l0 + l1;
            m_run = m_new;
        }
        float* xb = xt + (nblk & 1) * 256;
        xb[hf * 128 + rloc] = l_run;
        pair_sync_id(bar_id);
        const float inv = 1.0f / (l_run + xb[(hf ^ 1) * 128 + rloc]);
        tc::mbar_wait(&o_full[t], 0);
        tc::fence_after_sync();
        __nv_bfloat16* dst = p.out + (long long)b * p.o_stride_b + (long long)h * p.o_stride_h + (long long)row * p.ldo + 32 * hf;
        uint32_t o[32];
        tc::tmem_ld32(tmem_base + lane_off + o_col, o);
        tc::tmem_ld_wait();
        if (row < p.Nq) {
#pragma unroll
            for (int e8 = 0; e8 < 32; e8 += 8) {
                uint4 u;
                __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    h2[e] = __floats2bfloat162_rn(__uint_as_float(o[e8 + 2 * e]) * inv, __uint_as_float(o[e8 + 2 * e + 1]) * inv);
                *reinterpret_cast<uint4*>(dst + e8) = u;
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, 512);
}

#ifndef GVD_HOST_EMU
PFN_cuTensorMapEncodeTiled_v12000 fa_get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}
#else
PFN_cuTensorMapEncodeTiled_v12000 fa_get_encode() { return emu_cuTensorMapEncodeTiled; }
#endif

// [B, N, H*64] bf16 viewed as (d=64, rows=N with stride ld, heads with stride 64, batch with stride sb)
bool fa_make_tmap(CUtensorMap* map, const void* base, long long N, long long H, long long B, long long ld, long long sb) {
    auto enc = fa_get_encode();
    if (!enc) return false;
    cuuint64_t dims[4] = {64, (cuuint64_t)N, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, 64 * 2, (cuuint64_t)(B > 1 ? sb : ld) * 2};
    cuuint32_t box[4] = {64, 128, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

bool gvd_fa_make_tmap(CUtensorMap* map, const void* base, long long N, long long H, long long B, long long ld, long long sb) {
    return fa_make_tmap(map, base, N, H, B, ld, sb);
}

static int flash_forward(const void* q, const void* k, const void* v, void* out, float* lse, int B, int Nq, int Nk, int H,
                         long long q_batch_stride, long long kv_batch_stride, float scale, gvd_nn_stream_t stream_);

extern "C" int gvd_flash_attention(const void* q, const void* k, const void* v, void* out, int B, int Nq, int Nk, int H,
                                   long long q_batch_stride, long long kv_batch_stride, float scale,
                                   gvd_nn_stream_t stream_) {
    return flash_forward(q, k, v, out, nullptr, B, Nq, Nk, H, q_batch_stride, kv_batch_stride, scale, stream_);
}

extern "C" int gvd_flash_attention_lse(const void* q, const void* k, const void* v, void* out, float* lse, int B, int Nq, int Nk,
                                       int H, long long q_batch_stride, long long kv_batch_stride, float scale,
                                       gvd_nn_stream_t stream_) {
    if (!lse) { g_nn_err_ext = "gvd_flash_attention_lse: null lse"; return 2; }
    return flash_forward(q, k, v, out, lse, B, Nq, Nk, H, q_batch_stride, kv_batch_stride, scale, stream_);
}

static int flash_forward(const void* q, const void* k, const void* v, void* out, float* lse, int B, int Nq, int Nk, int H,
                         long long q_batch_stride, long long kv_batch_stride, float scale, gvd_nn_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (!q || !k || !v || !out) { g_nn_err_ext = "gvd_flash_attention: null pointer"; return 2; }
    if (B <= 0 || Nq <= 0 || H <= 0) return 0;
    if (Nk <= 0) { g_nn_err_ext = "gvd_flash_attention: Nk must be positive"; return 2; }
    const long long ld = (long long)H * 64;
    if ((q_batch_stride & 7) || (kv_batch_stride & 7)) { g_nn_err_ext = "gvd_flash_attention: batch strides must be multiples of 8"; return 2; }
    CUtensorMap tq, tk, tv;
    if (!fa_make_tmap(&tq, q, Nq, H, B, ld, q_batch_stride) || !fa_make_tmap(&tk, k, Nk, H, B, ld, kv_batch_stride) ||
        !fa_make_tmap(&tv, v, Nk, H, B, ld, kv_batch_stride)) {
        g_nn_err_ext = "gvd_flash_attention: cuTensorMapEncodeTiled failed";
        return 1;
    }
    // Default: generation 7 (64-key sub-blocks, S / P double-buffered, QK one sub-block ahead).  A/B timing knobs:
    // GVD_FLASH=v1 first generation, v2 second (one softmax thread per row, serial chain; GVD_FLASH_POLY=1: one exponential
    // in four on the FMA pipe), v5 two-tile ping-pong, v6 one-pass chunk-pipelined softmax.  Measured on B200 (N = 9216,
    // 5 heads, 25 frames): v1 5.03 ms, v2 4.13 ms (659 TFLOP/s), v2 + polynomial 4.46 ms, v5 4.54 ms, v6 4.15 ms,
    // v7 3.35 ms (810 TFLOP/s).  DESIGN.md section 7 has the accounting.
    static int variant = -1;
    if (variant < 0) {
        const char* v = getenv("GVD_FLASH");
        const char* pe = getenv("GVD_FLASH_POLY");
        int want = 7;
        if (v && v[0] == 'v' && v[1] == '1') want = 0;
        else if (v && v[0] == 'v' && v[1] == '2') want = (pe && pe[0] == '1') ? 2 : 1;
        else if (v && v[0] == 'v' && v[1] == '5') want = 5;
        else if (v && v[0] == 'v' && v[1] == '6') want = 6;
        else if (v && v[0] == 'v' && v[1] == '7') want = 7;
        else if (v && v[0] == 'v' && v[1] == '8') want = 11;
        if (want == 7 && pe && pe[0] >= '1' && pe[0] <= '3') want = 7 + (pe[0] - '0');  // 8: one exponential in 2 on the FMA pipe, 9: 1 in 4, 10: 1 in 8
        const void* fn = want == 0 ? (const void*)flash_attn_v1_kernel
                       : want == 1 ? (const void*)flash_attn_kernel<0>
                       : want == 2 ? (const void*)flash_attn_kernel<1>
                       : want == 5 ? (const void*)flash_attn5_kernel
                       : want == 6 ? (const void*)flash_attn_kernel<2>
                       : want == 8 ? (const void*)flash_attn7_kernel<1> : want == 9 ? (const void*)flash_attn7_kernel<2>
                       : want == 10 ? (const void*)flash_attn7_kernel<4>
                       : want == 11 ? (const void*)flash_attn8_kernel : (const void*)flash_attn7_kernel<0>;
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             want == 5 ? FA5_SMEM : (want == 11 ? FA8_SMEM : FA_SMEM));
        if (e != cudaSuccess) { g_nn_err_ext = std::string("gvd_flash_attention attr: ") + cudaGetErrorString(e); return 1; }
        variant = want;
    }
    FaParams p{reinterpret_cast<__nv_bfloat16*>(out), ld, 64, q_batch_stride, Nq, Nk, H, scale, lse, (Nq + 127) / 128 * 128};
    dim3 grid((Nq + FA_BM - 1) / FA_BM, H, B);
    if (lse != nullptr) {  // only generation 7 writes the row statistic, whatever GVD_FLASH selects for the plain forward
        static bool attr7 = false;
        if (!attr7) {
            cudaError_t e7 = cudaFuncSetAttribute(flash_attn7_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM);
            if (e7 != cudaSuccess) { g_nn_err_ext = std::string("gvd_flash_attention_lse attr: ") + cudaGetErrorString(e7); return 1; }
            attr7 = true;
        }
        flash_attn7_kernel<0><<<grid, FA_THREADS, FA_SMEM, s>>>(tq, tk, tv, p);
    } else if (variant == 0) flash_attn_v1_kernel<<<grid, FA_THREADS, FA_SMEM, s>>>(tq, tk, tv, p);
    else if (variant == 1) flash_attn_kernel<0><<<grid, FA_THREADS, FA_SMEM, s>>>(tq, tk, tv, p);
    else if (variant == 2) flash_attn_kernel<1><<<grid, FA_THREADS, FA_SMEM, s>>>(tq, tk, tv, p);
    else if (variant == 5) flash_attn5_kernel<<<dim3((Nq + 2 * FA_BM - 1) / (2 * FA_BM), H, B), FA5_THREADS, FA5_SMEM, s>>>(tq, tk, tv, p);
    else if (variant == 6) flash_attn_kernel<2><<<grid, FA_THREADS, FA_SMEM, s>>>(tq, tk, tv, p);
    else if (variant == 8) flash_attn7_kernel<1><<<grid, FA_THREADS, FA_SMEM, s>>>(tq, tk, tv, p);
    else if (variant == 9) flash_attn7_kernel<2><<<grid, FA_THREADS, FA_SMEM, s>>>(tq, tk, tv, p);
    else if (variant == 10) flash_attn7_kernel<4><<<grid, FA_THREADS, FA_SMEM, s>>>(tq, tk, tv, p);
    else if (variant == 11) flash_attn8_kernel<<<grid, FA8_THREADS, FA8_SMEM, s>>>(tq, tk, tv, p);
    else flash_attn7_kernel<0><<<grid, FA_THREADS, FA_SMEM, s>>>(tq, tk, tv, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_nn_err_ext = std::string("gvd_flash_attention launch: ") + cudaGetErrorString(e); return 1; }
    return 0;
}
