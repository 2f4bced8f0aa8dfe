// attn_bwd_tc.cu -- backward of softmax(Q K^T * scale) V on Blackwell tensor cores, head dim 64, WITHOUT materialising
// the score matrix (include/gvd_nn.h::gvd_flash_attention_bwd; the adjoint the guided sampler needs for
// CrossAttention.forward, attention.py:81-144, under torch.autograd in ddim_guidance.py:259-337).
//
// The first backward (vc_b200/ops.py::attention_bwd) recomputes S, P, dP and dS as [items, heads, Nq, Nk] bf16 matrices in
// HBM -- about ten score-sized passes per layer (1.6 GB each at 25 x 5 x 2560 x 2560) through five GEMM launches and
// two row kernels.  Here the forward keeps one fp32 per query row, L = log2 sum_j exp2(s_ij * scale * log2 e)
// (gvd_flash_attention_lse), a small kernel forms D_i = sum_d dO_id O_id, and ONE kernel template, run twice, does the
// rest on chip:
//
//   <false>  grid over 128-QUERY tiles; K / V tiles stream.   rows = queries
//            S = Q K^T, dP = dO V^T                (SS MMAs into TMEM)
//            dS = P o (dP - D_row),  P = exp2(S * scale * log2 e - L_row)   (softmax warps, registers)
//            dQ += dS K                            (TS MMA: A = dS as packed bf16 in TMEM, B = K tile, MN-major)
//   <true>   grid over 128-KEY tiles; Q / dO tiles stream.    rows = keys
//            S^T = K Q^T, dP^T = V dO^T
//            P^T, dS^T with the statistics now per COLUMN (read from global memory, one broadcast 16-byte load per four)
//            dV += P^T dO,  dK += dS^T Q
// Seven MMAs instead of the textbook five (S and dP are formed in both passes), in exchange for no atomics, no dQ
// accumulation in HBM and one code path.  The backward has NO row reductions inside the kernel -- the statistics are
// inputs -- so a sub-block's columns split freely over threads: 8 softmax warps, two per TMEM lane quadrant, each
// owning 32 of the 64 columns, with no exchange between them (unlike the forward, where two threads per row lose to
// the maximum exchange -- attn_tc.cu generation 8).
//
// FIRST FORM (GVD_FLASH_BWD_CTAS=1; the default is the second form further down, 1.24 vs 1.29 ms at 25 x 5 x 2560 x 2560).
// One CTA per SM (all 512 TMEM columns), 10 warps: warp 0 TMA, warp 1 MMA issue, warps 2-9 softmax.  The streamed
// operand arrives as 128-row tiles (3-stage ring) and is consumed as two 64-row sub-blocks; S, dP, P and dS are double
// buffered; as soon as the softmax of sub-block i is done, the S / dP products of sub-block i + 2 go into the tensor pipe
// AHEAD of sub-block i's accumulating MMAs (the first version issued them one ahead and behind the accumulators: ncu
// showed 19-30 % of the warp samples on the softmax warps' wait for S / dP, tensor pipe 24 %), so a separate barrier
// tells the softmax when P / dS of sub-block i - 2 have been consumed.  The streamed ring has three stages for the same
// reason: the next tile is needed one sub-block earlier.
//   TMEM: S 0,64 | dP 128,192 | P 256,288 | dS 320,352 | acc0 (dV) 384 | acc1 (dQ or dK) 448
// Rounding points: P and dS enter their MMAs as bf16 (as in the materialised backward); everything else fp32.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cstdlib>
#include <string>

#include "../../include/gvd_nn.h"
#include "tc_common.cuh"

extern thread_local std::string g_nn_err_ext;
// attn_tc.cu: [B, N, H*64] bf16 as a (64, N, H, B) tensor map with 64 x 128 boxes, 128-byte swizzle
#ifndef GVD_HOST_EMU
bool gvd_fa_make_tmap(CUtensorMap* map, const void* base, long long N, long long H, long long B, long long ld, long long sb);
#else  // the host build (tests/cuda_emu) holds this file alone: the same map, recorded by the driver-API stand-in
static bool gvd_fa_make_tmap(CUtensorMap* map, const void* base, long long N, long long H, long long B, long long ld, long long sb) {
    cuuint64_t dims[4] = {64, (cuuint64_t)N, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, 64 * 2, (cuuint64_t)(B > 1 ? sb : ld) * 2};
    cuuint32_t box[4] = {64, 128, 1, 1}, estr[4] = {1, 1, 1, 1};
    return emu_cuTensorMapEncodeTiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
#endif

namespace {

constexpr int FB_TILE = 128 * 64 * 2;  // 16 KB: one 128-row tile of Q, K, V or dO
constexpr int FB_STAGES = 3;
constexpr int FB_THREADS = 64 + 8 * 32;
constexpr int FB_SMEM = FB_TILE * (2 + 2 * FB_STAGES) + 1024 + 256 + FB_STAGES * 1024;  // + per-stage column statistics
constexpr uint32_t FB_S = 0, FB_DP = 128, FB_P = 256, FB_DS = 320, FB_ACC0 = 384, FB_ACC1 = 448;

#ifndef GVD_HOST_EMU
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// M 128, N 64, B operand MN-major (the 64 "n" are contiguous in shared memory, rows are the k index)
__device__ __forceinline__ uint32_t idesc_ts_mn() {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
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
#else  // tests/cuda_emu/tc_emu.h
inline float ex2(float x) { return exp2f(x); }
inline uint32_t idesc_ts_mn() { return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }
inline void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    tc::emu_umma_bf16_ts(tmem_d, tmem_a, bdesc, idesc, accumulate);
}
inline void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) { tc::emu_tmem_st16(taddr, v); }
inline void tmem_st_wait() {}
#endif
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

// global -> shared bulk copy (bytes % 16 == 0), completion counted on an mbarrier like the tensor-map loads
#ifndef GVD_HOST_EMU
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc::smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(tc::smem_u32(bar))
                 : "memory");
}
#else
inline void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) { tc::emu_bulk_g2s(smem_dst, gmem_src, bytes, bar); }
#endif

struct FbParams {
    const float* lse;    // [B, H, ldl]  base-2 log-sum-exp of the scaled logits
    const float* delta;  // [B, H, ldl]  sum_d dO O
    __nv_bfloat16* out0; // KV: dV   (rows = keys)
    __nv_bfloat16* out1; // KV: dK ; else dQ
    long long ld, stride_b;  // of the outputs: row stride, batch stride (head stride is 64)
    int nrow, ncol, H, ldl;  // rows of the resident operand, rows of the streamed one
    float scale;
};

// x0 / x1: the resident pair (Q, dO | K, V); y0 / y1: the streamed pair (K, V | Q, dO)
template <bool KV>
__global__ void __launch_bounds__(FB_THREADS, 1)
flash_bwd_kernel(const __grid_constant__ CUtensorMap tmap_x0, const __grid_constant__ CUtensorMap tmap_x1,
                 const __grid_constant__ CUtensorMap tmap_y0, const __grid_constant__ CUtensorMap tmap_y1, FbParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sx0 = smem;
    uint8_t* sx1 = smem + FB_TILE;
    uint8_t* sy0 = smem + 2 * FB_TILE;
    uint8_t* sy1 = smem + (2 + FB_STAGES) * FB_TILE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (2 + 2 * FB_STAGES) * FB_TILE);
    uint64_t* x_full = bars;
    uint64_t* y_full = bars + 1;     // [3]
    uint64_t* y_empty = bars + 4;    // [3]
    uint64_t* sdp_full = bars + 7;   // [2]  S and dP of a sub-block are in TMEM
    uint64_t* pds_full = bars + 9;   // [2]  P and dS of a sub-block are in TMEM (8 arrivals)
    uint64_t* pds_free = bars + 11;  // [2]  the accumulating MMAs have consumed P / dS of a sub-block
    uint64_t* acc_done = bars + 13;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 14);
    // KV: lse | delta of the streamed tile's 128 queries, per stage (the statistics are per COLUMN there)
    float* sstat = reinterpret_cast<float*>(smem + (2 + 2 * FB_STAGES) * FB_TILE + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
    const int nblk = (p.ncol + 127) / 128;
    const int nsub = (p.ncol + 63) / 64;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmap_x0);
        tc::prefetch_tmap(&tmap_x1);
        tc::prefetch_tmap(&tmap_y0);
        tc::prefetch_tmap(&tmap_y1);
        tc::mbar_init(x_full, 1);
        for (int s = 0; s < FB_STAGES; ++s) {
            tc::mbar_init(&y_full[s], 1);
            tc::mbar_init(&y_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&sdp_full[s], 1);
            tc::mbar_init(&pds_full[s], 8);  // one arrival per softmax warp
            tc::mbar_init(&pds_free[s], 1);
        }
        tc::mbar_init(acc_done, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tmem_ptr, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            tc::mbar_expect_tx(x_full, 2 * FB_TILE);
            tc::tma_load_4d(sx0, &tmap_x0, x_full, 0, m0, h, b);
            tc::tma_load_4d(sx1, &tmap_x1, x_full, 0, m0, h, b);
            for (int j = 0; j < nblk; ++j) {
                const int s = j % FB_STAGES;
                tc::mbar_wait(&y_empty[s], (uint32_t)(((j / FB_STAGES) & 1) ^ 1));
                tc::mbar_expect_tx(&y_full[s], 2 * FB_TILE + (KV ? 1024 : 0));
                tc::tma_load_4d(sy0 + s * FB_TILE, &tmap_y0, &y_full[s], 0, j * 128, h, b);
                tc::tma_load_4d(sy1 + s * FB_TILE, &tmap_y1, &y_full[s], 0, j * 128, h, b);
                if (KV) {
                    const long long so = ((long long)b * p.H + h) * p.ldl + (long long)j * 128;
                    bulk_g2s(sstat + s * 256, p.lse + so, 512, &y_full[s]);
                    bulk_g2s(sstat + s * 256 + 128, p.delta + so, 512, &y_full[s]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_ss = tc::make_idesc_bf16(128, 64);
            const uint32_t idesc_ts = idesc_ts_mn();
            const uint32_t x0_addr = tc::smem_u32(sx0), x1_addr = tc::smem_u32(sx1);
            auto issue_sdp = [&](int i) {  // S[i & 1] = X0 Y0sub^T, dP[i & 1] = X1 Y1sub^T; sub-block i = rows 64 (i & 1) .. + 63 of tile i / 2
                const int j = i >> 1, s = j % FB_STAGES;
                if ((i & 1) == 0) {
                    tc::mbar_wait(&y_full[s], (uint32_t)((j / FB_STAGES) & 1));
                    tc::fence_after_sync();
                }
                const uint32_t y0 = tc::smem_u32(sy0 + s * FB_TILE) + (uint32_t)(i & 1) * 64 * 128;
                const uint32_t y1 = tc::smem_u32(sy1 + s * FB_TILE) + (uint32_t)(i & 1) * 64 * 128;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    tc::umma_bf16(tmem_base + FB_S + (uint32_t)(i & 1) * 64, tc::make_desc_kmajor_sw128(x0_addr + k * 32),
                                  tc::make_desc_kmajor_sw128(y0 + k * 32), idesc_ss, k != 0);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    tc::umma_bf16(tmem_base + FB_DP + (uint32_t)(i & 1) * 64, tc::make_desc_kmajor_sw128(x1_addr + k * 32),
                                  tc::make_desc_kmajor_sw128(y1 + k * 32), idesc_ss, k != 0);
                tc::umma_commit(&sdp_full[i & 1]);
            };
            tc::mbar_wait(x_full, 0);
            tc::fence_after_sync();
            issue_sdp(0);
            if (nsub > 1) issue_sdp(1);
            for (int i = 0; i < nsub; ++i) {
                tc::mbar_wait(&pds_full[i & 1], (uint32_t)((i >> 1) & 1));  // softmax(i) is done: it has read S / dP[i & 1] and written P / dS[i & 1]
                tc::fence_after_sync();
                // S / dP two sub-blocks ahead go into the pipe FIRST: the softmax warps need them one softmax time from now,
                // the accumulators only at the end (issued after them, the profile showed the softmax warps waiting here)
                if (i + 2 < nsub) issue_sdp(i + 2);
                const int j = i >> 1, s = j % FB_STAGES;
                const uint32_t y0 = tc::smem_u32(sy0 + s * FB_TILE) + (uint32_t)(i & 1) * 64 * 128;
                const uint32_t y1 = tc::smem_u32(sy1 + s * FB_TILE) + (uint32_t)(i & 1) * 64 * 128;
                if (KV) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)  // 16 streamed rows per MMA: 8 packed-bf16 TMEM columns, 16 smem rows
                        umma_bf16_ts(tmem_base + FB_ACC0, tmem_base + FB_P + (uint32_t)(i & 1) * 32 + k * 8,
                                     tc::make_desc_kmajor_sw128(y1 + k * 16 * 128), idesc_ts, (i | k) != 0);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16_ts(tmem_base + FB_ACC1, tmem_base + FB_DS + (uint32_t)(i & 1) * 32 + k * 8,
                                 tc::make_desc_kmajor_sw128(y0 + k * 16 * 128), idesc_ts, (i | k) != 0);
                tc::umma_commit(&pds_free[i & 1]);  // softmax(i + 2) may overwrite P / dS[i & 1] once this completes
                if ((i & 1) || i + 1 == nsub) tc::umma_commit(&y_empty[s]);  // the tile's last sub-block
            }
            tc::umma_commit(acc_done);
        }
    } else {
        const int q = warp & 3;          // TMEM lane quadrant of this warp
        const int hf = (warp - 2) >> 2;  // which 32 of a sub-block's 64 columns
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int row = m0 + q * 32 + lane;
        const float sl2 = p.scale * 1.4426950408889634f;
        const float* lse = p.lse + ((long long)b * p.H + h) * p.ldl;
        const float* delta = p.delta + ((long long)b * p.H + h) * p.ldl;
        float nl_row = 0.f, d_row = 0.f;
        if (!KV) {  // rows are queries: m0 + 127 < ldl (ldl is a multiple of 128)
            nl_row = -lse[row];
            d_row = delta[row];
        }
        for (int i = 0; i < nsub; ++i) {
            const int col0 = i * 64 + 32 * hf;
            const float* st = sstat + ((i >> 1) % FB_STAGES) * 256 + (i & 1) * 64 + 32 * hf;  // KV: this thread's 32 columns
            if (KV && (i & 1) == 0) tc::mbar_wait(&y_full[(i >> 1) % FB_STAGES], (uint32_t)(((i >> 1) / FB_STAGES) & 1));  // the bulk copies have landed
            tc::mbar_wait(&sdp_full[i & 1], (uint32_t)((i >> 1) & 1));
            tc::fence_after_sync();
            uint32_t sv[32], dv[32];
            tc::tmem_ld32(tmem_base + lane_off + FB_S + (uint32_t)(i & 1) * 64 + 32 * hf, sv);
            tc::tmem_ld32(tmem_base + lane_off + FB_DP + (uint32_t)(i & 1) * 64 + 32 * hf, dv);
            tc::tmem_ld_wait();
            uint32_t pk[16], dk[16];
            auto element_pair = [&](int e, float& p0, float& p1, float& s0, float& s1) {
                float nl0 = nl_row, nl1 = nl_row, dd0 = d_row, dd1 = d_row;
                if constexpr (KV) {  // all lanes read the same shared address: a broadcast
                    const float4 a = *reinterpret_cast<const float4*>(st + (e & ~3)), c = *reinterpret_cast<const float4*>(st + 128 + (e & ~3));
                    nl0 = (e & 2) ? -a.z : -a.x;
                    nl1 = (e & 2) ? -a.w : -a.y;
                    dd0 = (e & 2) ? c.z : c.x;
                    dd1 = (e & 2) ? c.w : c.y;
                }
                p0 = ex2(fmaf(__uint_as_float(sv[e]), sl2, nl0));
                p1 = ex2(fmaf(__uint_as_float(sv[e + 1]), sl2, nl1));
                s0 = p0 * (__uint_as_float(dv[e]) - dd0);
                s1 = p1 * (__uint_as_float(dv[e + 1]) - dd1);
            };
            if (col0 + 32 <= p.ncol) {  // the common case carries no masking instructions at all
#pragma unroll
                for (int e = 0; e < 32; e += 2) {
                    float p0, p1, s0, s1;
                    element_pair(e, p0, p1, s0, s1);
                    if (KV) pk[e / 2] = pack_bf16(p0, p1);
                    dk[e / 2] = pack_bf16(s0, s1);
                }
            } else {  // only the last sub-block(s) can hold out-of-range columns
#pragma unroll
                for (int e = 0; e < 32; e += 2) {
                    float p0, p1, s0, s1;
                    element_pair(e, p0, p1, s0, s1);
                    if (col0 + e >= p.ncol) p0 = s0 = 0.f;
                    if (col0 + e + 1 >= p.ncol) p1 = s1 = 0.f;
                    if (KV) pk[e / 2] = pack_bf16(p0, p1);
                    dk[e / 2] = pack_bf16(s0, s1);
                }
            }
            if (i >= 2) {  // the accumulating MMAs of sub-block i - 2 have read P / dS[i & 1]
                tc::mbar_wait(&pds_free[i & 1], (uint32_t)(((i >> 1) - 1) & 1));
                tc::fence_after_sync();
            }
            if (KV) tmem_st16(tmem_base + lane_off + FB_P + (uint32_t)(i & 1) * 32 + 16 * hf, pk);
            tmem_st16(tmem_base + lane_off + FB_DS + (uint32_t)(i & 1) * 32 + 16 * hf, dk);
            tmem_st_wait();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&pds_full[i & 1]);
        }
        tc::mbar_wait(acc_done, 0);
        tc::fence_after_sync();
        // this warp's 32 of the 64 output channels, of dV (unscaled) and of dK / dQ (x scale)
        const long long off = (long long)b * p.stride_b + (long long)h * 64 + (long long)row * p.ld + 32 * hf;
#pragma unroll
        for (int which = KV ? 0 : 1; which < 2; ++which) {
            uint32_t o[32];
            tc::tmem_ld32(tmem_base + lane_off + (which ? FB_ACC1 : FB_ACC0) + 32 * hf, o);
            tc::tmem_ld_wait();
            const float f = which ? p.scale : 1.0f;
            __nv_bfloat16* dst = (which ? p.out1 : p.out0) + off;
            if (row < p.nrow) {
#pragma unroll
                for (int e8 = 0; e8 < 32; e8 += 8) {
                    uint4 u;
                    u.x = pack_bf16(__uint_as_float(o[e8]) * f, __uint_as_float(o[e8 + 1]) * f);
                    u.y = pack_bf16(__uint_as_float(o[e8 + 2]) * f, __uint_as_float(o[e8 + 3]) * f);
                    u.z = pack_bf16(__uint_as_float(o[e8 + 4]) * f, __uint_as_float(o[e8 + 5]) * f);
                    u.w = pack_bf16(__uint_as_float(o[e8 + 6]) * f, __uint_as_float(o[e8 + 7]) * f);
                    *reinterpret_cast<uint4*>(dst + e8) = u;
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, 512);
}

// ---- second form: TWO CTAs per SM, 256 TMEM columns each -------------------------------------------------------------------
// The kernel above is bound by its per-sub-block round trip (profiles/r02_ncu_flash_bwd.txt: tensor pipe 25 %, MUFU 30 %),
// which a second, independent CTA on the same SM hides -- the way the forward runs.  To fit 256 columns the sub-blocks are
// single-buffered and the bf16 operands overwrite the fp32 products they came from: P goes into S's columns and dS into
// dP's (a thread overwrites only columns it has already read: keys 32 c .. 32 c + 31 -> packed columns 32 c .. 32 c + 15).
// The next S / dP MMAs are issued right behind the accumulating MMAs that still read P / dS from those columns; tcgen05.mma
// instructions of one CTA execute in issue order, so the reads are over before the overwrite.
//   TMEM: S | P 0 | dP | dS 64 | acc0 (dV) 128 | acc1 (dQ or dK) 192 ; 6 warps: TMA, MMA, 4 softmax (thread = row, all 64 columns)
constexpr int FB2_STAGES = 2;
constexpr int FB2_THREADS = 192;
constexpr int FB2_SMEM = FB_TILE * (2 + 2 * FB2_STAGES) + 1024 + 256 + FB2_STAGES * 1024;
constexpr uint32_t FB2_S = 0, FB2_DP = 64, FB2_ACC0 = 128, FB2_ACC1 = 192;

template <bool KV>
__global__ void __launch_bounds__(FB2_THREADS, 2)
flash_bwd2_kernel(const __grid_constant__ CUtensorMap tmap_x0, const __grid_constant__ CUtensorMap tmap_x1,
                  const __grid_constant__ CUtensorMap tmap_y0, const __grid_constant__ CUtensorMap tmap_y1, FbParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sx0 = smem;
    uint8_t* sx1 = smem + FB_TILE;
    uint8_t* sy0 = smem + 2 * FB_TILE;
    uint8_t* sy1 = smem + (2 + FB2_STAGES) * FB_TILE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (2 + 2 * FB2_STAGES) * FB_TILE);
    uint64_t* x_full = bars;
    uint64_t* y_full = bars + 1;    // [2]
    uint64_t* y_empty = bars + 3;   // [2]
    uint64_t* sdp_full = bars + 5;
    uint64_t* pds_full = bars + 6;  // 4 arrivals
    uint64_t* acc_done = bars + 7;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);
    float* sstat = reinterpret_cast<float*>(smem + (2 + 2 * FB2_STAGES) * FB_TILE + 256);  // KV: lse | delta of the streamed tile, per stage

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
    const int nblk = (p.ncol + 127) / 128;
    const int nsub = (p.ncol + 63) / 64;

    if (warp == 0 && lane == 0) {
        tc::prefetch_tmap(&tmap_x0);
        tc::prefetch_tmap(&tmap_x1);
        tc::prefetch_tmap(&tmap_y0);
        tc::prefetch_tmap(&tmap_y1);
        tc::mbar_init(x_full, 1);
        for (int s = 0; s < FB2_STAGES; ++s) {
            tc::mbar_init(&y_full[s], 1);
            tc::mbar_init(&y_empty[s], 1);
        }
        tc::mbar_init(sdp_full, 1);
        tc::mbar_init(pds_full, 4);
        tc::mbar_init(acc_done, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc(tmem_ptr, 256);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            tc::mbar_expect_tx(x_full, 2 * FB_TILE);
            tc::tma_load_4d(sx0, &tmap_x0, x_full, 0, m0, h, b);
            tc::tma_load_4d(sx1, &tmap_x1, x_full, 0, m0, h, b);
            for (int j = 0; j < nblk; ++j) {
                const int s = j % FB2_STAGES;
                tc::mbar_wait(&y_empty[s], (uint32_t)(((j / FB2_STAGES) & 1) ^ 1));
                tc::mbar_expect_tx(&y_full[s], 2 * FB_TILE + (KV ? 1024 : 0));
                tc::tma_load_4d(sy0 + s * FB_TILE, &tmap_y0, &y_full[s], 0, j * 128, h, b);
                tc::tma_load_4d(sy1 + s * FB_TILE, &tmap_y1, &y_full[s], 0, j * 128, h, b);
                if (KV) {
                    const long long so = ((long long)b * p.H + h) * p.ldl + (long long)j * 128;
                    bulk_g2s(sstat + s * 256, p.lse + so, 512, &y_full[s]);
                    bulk_g2s(sstat + s * 256 + 128, p.delta + so, 512, &y_full[s]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_ss = tc::make_idesc_bf16(128, 64);
            const uint32_t idesc_ts = idesc_ts_mn();
            const uint32_t x0_addr = tc::smem_u32(sx0), x1_addr = tc::smem_u32(sx1);
            tc::mbar_wait(x_full, 0);
            tc::fence_after_sync();
            for (int i = 0; i < nsub; ++i) {
                const int j = i >> 1, s = j % FB2_STAGES;
                if ((i & 1) == 0) {
                    tc::mbar_wait(&y_full[s], (uint32_t)((j / FB2_STAGES) & 1));
                    tc::fence_after_sync();
                }
                const uint32_t y0 = tc::smem_u32(sy0 + s * FB_TILE) + (uint32_t)(i & 1) * 64 * 128;
                const uint32_t y1 = tc::smem_u32(sy1 + s * FB_TILE) + (uint32_t)(i & 1) * 64 * 128;
                // (the softmax of sub-block i - 1 has read S / dP: pds_full(i - 1) was waited on below; the MMAs of sub-block
                // i - 1 that read P / dS from these columns were issued before these and execute before them)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    tc::umma_bf16(tmem_base + FB2_S, tc::make_desc_kmajor_sw128(x0_addr + k * 32), tc::make_desc_kmajor_sw128(y0 + k * 32),
                                  idesc_ss, k != 0);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    tc::umma_bf16(tmem_base + FB2_DP, tc::make_desc_kmajor_sw128(x1_addr + k * 32), tc::make_desc_kmajor_sw128(y1 + k * 32),
                                  idesc_ss, k != 0);
                tc::umma_commit(sdp_full);
                tc::mbar_wait(pds_full, (uint32_t)(i & 1));
                tc::fence_after_sync();
                if (KV) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)  // keys 16 k .. + 15 of the sub-block: packed columns 32 (k / 2) + 8 (k % 2)
                        umma_bf16_ts(tmem_base + FB2_ACC0, tmem_base + FB2_S + (uint32_t)((k >> 1) * 32 + (k & 1) * 8),
                                     tc::make_desc_kmajor_sw128(y1 + k * 16 * 128), idesc_ts, (i | k) != 0);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16_ts(tmem_base + FB2_ACC1, tmem_base + FB2_DP + (uint32_t)((k >> 1) * 32 + (k & 1) * 8),
                                 tc::make_desc_kmajor_sw128(y0 + k * 16 * 128), idesc_ts, (i | k) != 0);
                if ((i & 1) || i + 1 == nsub) tc::umma_commit(&y_empty[s]);
            }
            tc::umma_commit(acc_done);
        }
    } else {
        const int q = warp & 3;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int row = m0 + q * 32 + lane;
        const float sl2 = p.scale * 1.4426950408889634f;
        const float* lse = p.lse + ((long long)b * p.H + h) * p.ldl;
        const float* delta = p.delta + ((long long)b * p.H + h) * p.ldl;
        float nl_row = 0.f, d_row = 0.f;
        if (!KV) {
            nl_row = -lse[row];
            d_row = delta[row];
        }
        for (int i = 0; i < nsub; ++i) {
            if (KV && (i & 1) == 0) tc::mbar_wait(&y_full[(i >> 1) % FB2_STAGES], (uint32_t)(((i >> 1) / FB2_STAGES) & 1));  // the bulk copies have landed
            tc::mbar_wait(sdp_full, (uint32_t)(i & 1));
            tc::fence_after_sync();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int col0 = i * 64 + 32 * c;
                const float* st = sstat + ((i >> 1) % FB2_STAGES) * 256 + (i & 1) * 64 + 32 * c;
                uint32_t sv[32], dv[32];
                tc::tmem_ld32(tmem_base + lane_off + FB2_S + 32 * c, sv);
                tc::tmem_ld32(tmem_base + lane_off + FB2_DP + 32 * c, dv);
                tc::tmem_ld_wait();
                uint32_t pk[16], dk[16];
                auto element_pair = [&](int e, float& p0, float& p1, float& s0, float& s1) {
                    float nl0 = nl_row, nl1 = nl_row, dd0 = d_row, dd1 = d_row;
                    if constexpr (KV) {  // all lanes read the same shared address: a broadcast
                        const float4 a = *reinterpret_cast<const float4*>(st + (e & ~3)), cc = *reinterpret_cast<const float4*>(st + 128 + (e & ~3));
                        nl0 = (e & 2) ? -a.z : -a.x;
                        nl1 = (e & 2) ? -a.w : -a.y;
                        dd0 = (e & 2) ? cc.z : cc.x;
                        dd1 = (e & 2) ? cc.w : cc.y;
                    }
                    p0 = ex2(fmaf(__uint_as_float(sv[e]), sl2, nl0));
                    p1 = ex2(fmaf(__uint_as_float(sv[e + 1]), sl2, nl1));
                    s0 = p0 * (__uint_as_float(dv[e]) - dd0);
                    s1 = p1 * (__uint_as_float(dv[e + 1]) - dd1);
                };
                if (col0 + 32 <= p.ncol) {
#pragma unroll
                    for (int e = 0; e < 32; e += 2) {
                        float p0, p1, s0, s1;
                        element_pair(e, p0, p1, s0, s1);
                        if (KV) pk[e / 2] = pack_bf16(p0, p1);
                        dk[e / 2] = pack_bf16(s0, s1);
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 32; e += 2) {
                        float p0, p1, s0, s1;
                        element_pair(e, p0, p1, s0, s1);
                        if (col0 + e >= p.ncol) p0 = s0 = 0.f;
                        if (col0 + e + 1 >= p.ncol) p1 = s1 = 0.f;
                        if (KV) pk[e / 2] = pack_bf16(p0, p1);
                        dk[e / 2] = pack_bf16(s0, s1);
                    }
                }
                // in place: these 16 packed columns lie inside the 32 fp32 columns just read
                if (KV) tmem_st16(tmem_base + lane_off + FB2_S + 32 * c, pk);
                tmem_st16(tmem_base + lane_off + FB2_DP + 32 * c, dk);
            }
            tmem_st_wait();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(pds_full);
        }
        tc::mbar_wait(acc_done, 0);
        tc::fence_after_sync();
        const long long off = (long long)b * p.stride_b + (long long)h * 64 + (long long)row * p.ld;
#pragma unroll
        for (int which = KV ? 0 : 1; which < 2; ++which) {
            const float f = which ? p.scale : 1.0f;
            __nv_bfloat16* dst = (which ? p.out1 : p.out0) + off;
#pragma unroll
            for (int c = 0; c < 64; c += 32) {
                uint32_t o[32];
                tc::tmem_ld32(tmem_base + lane_off + (which ? FB2_ACC1 : FB2_ACC0) + c, o);
                tc::tmem_ld_wait();
                if (row < p.nrow) {
#pragma unroll
                    for (int e8 = 0; e8 < 32; e8 += 8) {
                        uint4 u;
                        u.x = pack_bf16(__uint_as_float(o[e8]) * f, __uint_as_float(o[e8 + 1]) * f);
                        u.y = pack_bf16(__uint_as_float(o[e8 + 2]) * f, __uint_as_float(o[e8 + 3]) * f);
                        u.z = pack_bf16(__uint_as_float(o[e8 + 4]) * f, __uint_as_float(o[e8 + 5]) * f);
                        u.w = pack_bf16(__uint_as_float(o[e8 + 6]) * f, __uint_as_float(o[e8 + 7]) * f);
                        *reinterpret_cast<uint4*>(dst + c + e8) = u;
                    }
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, 256);
}

// D[b, h, i] = sum_d dO[b, i, h, d] O[b, i, h, d]; 8 lanes per (b, i, h), 16 bytes of each operand per lane.  Rows
// i in [Nq, ldl) get 0.
__global__ void flash_bwd_delta_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ out, float* __restrict__ delta,
                                       int B, int Nq, int H, int ldl, long long ld, long long stride_b) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long g = t >> 3;
    const int sub = (int)(t & 7);
    const long long total = (long long)B * H * ldl;
    if (g >= total) return;  // whole groups of 8 leave together (blockDim is a multiple of 8)
    const int i = (int)(g % ldl);
    const int h = (int)((g / ldl) % H);
    const int b = (int)(g / ((long long)ldl * H));
    float acc = 0.f;
    if (i < Nq) {
        const long long off = (long long)b * stride_b + (long long)i * ld + (long long)h * 64 + sub * 8;
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(dout + off));
        const uint4 c = __ldg(reinterpret_cast<const uint4*>(out + off));
        const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&a);
        const __nv_bfloat162* ch = reinterpret_cast<const __nv_bfloat162*>(&c);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 x = __bfloat1622float2(ah[e]), y = __bfloat1622float2(ch[e]);
            acc = fmaf(x.x, y.x, acc);
            acc = fmaf(x.y, y.y, acc);
        }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (sub == 0) delta[g] = acc;
}

}  // namespace

extern "C" int gvd_flash_attention_bwd(const GvdFlashBwdArgs* a, gvd_nn_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (!a) { g_nn_err_ext = "gvd_flash_attention_bwd: null args"; return 2; }
    if (!a->q || !a->k || !a->v || !a->out || !a->dout || !a->lse || !a->delta || !a->dq) {
        g_nn_err_ext = "gvd_flash_attention_bwd: null pointer";
        return 2;
    }
    if ((a->dk == nullptr) != (a->dv == nullptr)) { g_nn_err_ext = "gvd_flash_attention_bwd: dk and dv go together"; return 2; }
    if (a->B <= 0 || a->Nq <= 0 || a->H <= 0) return 0;
    if (a->Nk <= 0) { g_nn_err_ext = "gvd_flash_attention_bwd: Nk must be positive"; return 2; }
    const long long ld = (long long)a->H * 64;
    if ((a->q_batch_stride & 7) || (a->kv_batch_stride & 7)) { g_nn_err_ext = "gvd_flash_attention_bwd: batch strides must be multiples of 8"; return 2; }
    const int ldl = (a->Nq + 127) / 128 * 128;
    if ((reinterpret_cast<uintptr_t>(a->lse) | reinterpret_cast<uintptr_t>(a->delta)) & 15) {
        g_nn_err_ext = "gvd_flash_attention_bwd: lse / delta must be 16-byte aligned";
        return 2;
    }
    CUtensorMap tq, tk, tv, td;
    if (!gvd_fa_make_tmap(&tq, a->q, a->Nq, a->H, a->B, ld, a->q_batch_stride) ||
        !gvd_fa_make_tmap(&td, a->dout, a->Nq, a->H, a->B, ld, a->q_batch_stride) ||
        !gvd_fa_make_tmap(&tk, a->k, a->Nk, a->H, a->B, ld, a->kv_batch_stride) ||
        !gvd_fa_make_tmap(&tv, a->v, a->Nk, a->H, a->B, ld, a->kv_batch_stride)) {
        g_nn_err_ext = "gvd_flash_attention_bwd: cuTensorMapEncodeTiled failed";
        return 1;
    }
    // GVD_FLASH_BWD_CTAS=1: the first form (one CTA per SM, double-buffered sub-blocks); default 2: two CTAs per SM
    static int form = 0;
    if (!form) {
        const char* ev = getenv("GVD_FLASH_BWD_CTAS");
        const int want = (ev && ev[0] == '1') ? 1 : 2;
        cudaError_t e = cudaFuncSetAttribute(flash_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(flash_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(flash_bwd2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB2_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(flash_bwd2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB2_SMEM);
        if (e != cudaSuccess) { g_nn_err_ext = std::string("gvd_flash_attention_bwd attr: ") + cudaGetErrorString(e); return 1; }
        form = want;
    }
    {
        const long long threads = (long long)a->B * a->H * ldl * 8;
        flash_bwd_delta_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(
            reinterpret_cast<const __nv_bfloat16*>(a->dout), reinterpret_cast<const __nv_bfloat16*>(a->out), a->delta, a->B, a->Nq, a->H, ldl,
            ld, a->q_batch_stride);
    }
    {
        FbParams p{a->lse, a->delta, nullptr, reinterpret_cast<__nv_bfloat16*>(a->dq), ld, a->q_batch_stride, a->Nq, a->Nk, a->H, ldl, a->scale};
        dim3 grid((a->Nq + 127) / 128, a->H, a->B);
        if (form == 2) flash_bwd2_kernel<false><<<grid, FB2_THREADS, FB2_SMEM, s>>>(tq, td, tk, tv, p);
        else flash_bwd_kernel<false><<<grid, FB_THREADS, FB_SMEM, s>>>(tq, td, tk, tv, p);
    }
    if (a->dk) {
        FbParams p{a->lse, a->delta, reinterpret_cast<__nv_bfloat16*>(a->dv), reinterpret_cast<__nv_bfloat16*>(a->dk), ld, a->kv_batch_stride,
                   a->Nk, a->Nq, a->H, ldl, a->scale};
        dim3 grid((a->Nk + 127) / 128, a->H, a->B);
        if (form == 2) flash_bwd2_kernel<true><<<grid, FB2_THREADS, FB2_SMEM, s>>>(tk, tv, tq, td, p);
        else flash_bwd_kernel<true><<<grid, FB_THREADS, FB_SMEM, s>>>(tk, tv, tq, td, p);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_nn_err_ext = std::string("gvd_flash_attention_bwd launch: ") + cudaGetErrorString(e); return 1; }
    return 0;
}
