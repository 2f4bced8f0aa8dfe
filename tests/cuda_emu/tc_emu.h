// tests/cuda_emu: the namespace tc of csrc/tc_common.cuh on the host -- mbarrier, TMA tensor-map loads with the 128-byte
// swizzle, tensor memory, tcgen05.mma (both operands in shared memory, or A in tensor memory), tcgen05.ld / st / commit.
// Every CUDA thread is an OS thread (cuda_emu.h); blocks run one after another, so one tensor memory and one 256 KB
// shared-memory window serve.  What is modelled: addresses (descriptor start addresses, the swizzle as an XOR of address
// bits [4,7) with bits [7,10), box copies with zero fill outside the tensor, TMEM lane / column addressing, packed bf16
// A operands), barrier phases and transaction counts.  What is not: asynchrony -- an MMA or a copy completes at the call,
// so orderings that only asynchronous completion can break are the business of tests/test_barrier_protocol_cpu.py.
#pragma once
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <random>
#include <thread>
#include <vector>

#include "cuda.h"
#include "cuda_emu.h"

inline char* emu_smem_window = nullptr;          // 256 KB-aligned window holding the block's shared memory
inline uint32_t emu_tmem[128][512];             // lanes x 32-bit columns

// ---- optional asynchrony (GVD_EMU_ASYNC=<lag 0..99>): MMAs / commits go into an in-order queue, tensor-map and bulk copies
// into an unordered one, and are executed LATER -- by whichever thread is spinning in an mbarrier wait, with probability
// (100 - lag) % per spin.  The kernel then sees operands consumed and barriers completed at arbitrary later times, in the
// order the hardware guarantees and no other: an ordering that only holds "because the MMA is fast" breaks here.
struct EmuEngine {
    std::mutex mu;
    std::deque<std::function<void()>> pipe;      // tcgen05.mma / tcgen05.commit: executed in issue order
    std::vector<std::function<void()>> copies;   // TMA: any order
    std::mt19937 rng{12345};
    int lag = -1;                                // -1: synchronous (default)
};
inline EmuEngine emu_engine;
inline bool emu_async() {
    static int on = -2;
    if (on == -2) {
        const char* e = getenv("GVD_EMU_ASYNC");
        on = e ? atoi(e) : -1;
        emu_engine.lag = on;
    }
    return on >= 0;
}
inline void emu_engine_step(bool force) {  // execute one pending operation (maybe)
    std::lock_guard<std::mutex> g(emu_engine.mu);
    if (!force && (int)(emu_engine.rng() % 100) < emu_engine.lag) return;
    const bool take_copy = !emu_engine.copies.empty() && (emu_engine.pipe.empty() || (emu_engine.rng() & 1));
    if (take_copy) {
        const size_t i = emu_engine.rng() % emu_engine.copies.size();
        auto op = std::move(emu_engine.copies[i]);
        emu_engine.copies.erase(emu_engine.copies.begin() + (long)i);
        op();
    } else if (!emu_engine.pipe.empty()) {
        auto op = std::move(emu_engine.pipe.front());
        emu_engine.pipe.pop_front();
        op();
    }
}
inline void emu_engine_drain() {
    for (;;) {
        {
            std::lock_guard<std::mutex> g(emu_engine.mu);
            if (emu_engine.pipe.empty() && emu_engine.copies.empty()) return;
        }
        emu_engine_step(true);
    }
}
template <class F> inline void emu_issue_pipe(F&& f) {
    if (!emu_async()) { f(); return; }
    std::lock_guard<std::mutex> g(emu_engine.mu);
    emu_engine.pipe.emplace_back(std::forward<F>(f));
}
template <class F> inline void emu_issue_copy(F&& f) {
    if (!emu_async()) { f(); return; }
    std::lock_guard<std::mutex> g(emu_engine.mu);
    emu_engine.copies.emplace_back(std::forward<F>(f));
}

namespace tc {

inline uint32_t smem_u32(const void* p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    emu_smem_window = reinterpret_cast<char*>(a & ~uintptr_t(0x3FFFF));
    return (uint32_t)(a & 0x3FFFF);
}

// ---- mbarrier in 8 bytes: [63:32] signed transaction count | [31:20] pending arrivals | [19:8] arrival count | [0] phase ----
inline void emu_bar_update(uint64_t* bar, int arrivals, int tx) {
    std::atomic_ref<uint64_t> a(*bar);
    uint64_t old = a.load(), neu;
    do {
        int32_t t = (int32_t)(old >> 32) + tx;
        uint32_t pend = (uint32_t)((old >> 20) & 0xFFF) - (uint32_t)arrivals, init = (uint32_t)((old >> 8) & 0xFFF), phase = (uint32_t)(old & 1);
        if (pend == 0 && t == 0) {
            phase ^= 1;
            pend = init;
        }
        neu = ((uint64_t)(uint32_t)t << 32) | ((uint64_t)(pend & 0xFFF) << 20) | ((uint64_t)init << 8) | phase;
    } while (!a.compare_exchange_weak(old, neu));
}
inline void mbar_init(uint64_t* bar, uint32_t count) { std::atomic_ref<uint64_t>(*bar).store(((uint64_t)count << 20) | ((uint64_t)count << 8)); }
inline void fence_barrier_init() {}
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { emu_bar_update(bar, 1, (int)bytes); }  // arrive.expect_tx
inline void mbar_arrive(uint64_t* bar) { emu_bar_update(bar, 1, 0); }
inline void emu_complete_tx(uint64_t* bar, uint32_t bytes) { emu_bar_update(bar, 0, -(int)bytes); }
inline void mbar_wait(uint64_t* bar, uint32_t parity) {  // the phase of this parity has completed <=> the current phase's parity differs
    std::atomic_ref<uint64_t> a(*bar);
    while ((uint32_t)(a.load() & 1) == parity) {
        if (emu_async()) emu_engine_step(false);
        std::this_thread::yield();
    }
}

// ---- TMA ----
inline void prefetch_tmap(const CUtensorMap*) {}
inline uint32_t emu_swz(uint32_t off) { return off ^ (((off >> 7) & 7) << 4); }
inline void emu_tma_box_now(uint32_t dst, const CUtensorMap* m, uint64_t* bar, const int (&c)[5]);
inline void emu_tma_box(void* smem_dst, const CUtensorMap* m, uint64_t* bar, const int (&c)[5]) {
    const uint32_t dst = smem_u32(smem_dst);
    const CUtensorMap map = *m;  // the kernel parameter outlives the copy, but a by-value capture costs nothing
    const int c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3], c4 = c[4];
    emu_issue_copy([=]() { const int cc[5] = {c0, c1, c2, c3, c4}; emu_tma_box_now(dst, &map, bar, cc); });
}
inline void emu_tma_box_now(uint32_t dst, const CUtensorMap* m, uint64_t* bar, const int (&c)[5]) {
    const uint32_t row_bytes = m->box[0] * 2;  // the innermost box dimension is one swizzle row (128 bytes in every use here)
    uint32_t r = 0;
    for (uint32_t i4 = 0; i4 < m->box[4]; ++i4)
        for (uint32_t i3 = 0; i3 < m->box[3]; ++i3)
            for (uint32_t i2 = 0; i2 < m->box[2]; ++i2)
                for (uint32_t i1 = 0; i1 < m->box[1]; ++i1, ++r) {
                    const long long x[5] = {c[0], (long long)c[1] + i1, (long long)c[2] + i2, (long long)c[3] + i3, (long long)c[4] + i4};
                    bool inside = true;
                    for (int d = 1; d < 5; ++d) inside = inside && x[d] >= 0 && x[d] < (long long)m->dim[d];
                    for (uint32_t e = 0; e < m->box[0]; e += 8) {  // 16-byte chunks
                        char chunk[16] = {0};
                        for (uint32_t k = 0; k < 8; ++k) {
                            const long long x0 = x[0] + e + k;
                            if (inside && x0 >= 0 && x0 < (long long)m->dim[0]) {
                                const char* src = m->base + x0 * 2;
                                for (int d = 1; d < 5; ++d) src += x[d] * (long long)m->stride[d];
                                std::memcpy(chunk + 2 * k, src, 2);
                            }
                        }
                        const uint32_t lin = dst + r * row_bytes + e * 2;
                        std::memcpy(emu_smem_window + (m->swizzle == CU_TENSOR_MAP_SWIZZLE_128B ? emu_swz(lin) : lin), chunk, 16);
                    }
                }
    emu_complete_tx(bar, r * row_bytes);
}
inline void tma_load_3d(void* d, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) { emu_tma_box(d, m, bar, {c0, c1, c2, 0, 0}); }
inline void tma_load_4d(void* d, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) { emu_tma_box(d, m, bar, {c0, c1, c2, c3, 0}); }

// ---- tensor memory ----
inline void tmem_alloc(uint32_t* smem_result, uint32_t) { *smem_result = 0; }
inline void tmem_dealloc(uint32_t, uint32_t) { if (emu_async()) emu_engine_drain(); }  // the block is over: nothing may stay in flight
inline void fence_before_sync() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void fence_after_sync() { std::atomic_thread_fence(std::memory_order_seq_cst); }

inline uint64_t make_desc_kmajor_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
constexpr uint32_t make_idesc_bf16(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

inline float emu_smem_bf16(uint32_t off) {
    uint16_t h;
    std::memcpy(&h, emu_smem_window + emu_swz(off), 2);
    const uint32_t b = (uint32_t)h << 16;
    float f;
    std::memcpy(&f, &b, 4);
    return f;
}
inline float emu_b_element(uint32_t b_start, bool mn_major, int n, int k) {
    return mn_major ? emu_smem_bf16(b_start + (uint32_t)(k / 8) * 1024 + (uint32_t)(k % 8) * 128 + (uint32_t)n * 2)
                    : emu_smem_bf16(b_start + (uint32_t)(n / 8) * 1024 + (uint32_t)(n % 8) * 128 + (uint32_t)k * 2);
}
inline void emu_mma_store(uint32_t tmem_d, int m, int n, float sum, bool accumulate) {
    uint32_t& cell = emu_tmem[(tmem_d >> 16) + m][(tmem_d & 0xFFFF) + n];
    float old;
    std::memcpy(&old, &cell, 4);
    const float v = accumulate ? old + sum : sum;
    std::memcpy(&cell, &v, 4);
}
// D[tmem] (+)= A[smem] * B[smem]^T, K = 16 per instruction; M x N from the instruction descriptor, bit 16: B is MN-major
inline void emu_umma_ss_now(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate);
inline void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    emu_issue_pipe([=]() { emu_umma_ss_now(tmem_d, adesc, bdesc, idesc, accumulate); });
}
inline void emu_umma_ss_now(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    const uint32_t a_start = (uint32_t)(adesc & 0x3FFF) << 4, b_start = (uint32_t)(bdesc & 0x3FFF) << 4;
    const int M = (int)((idesc >> 24) & 0x1F) << 4, N = (int)((idesc >> 17) & 0x3F) << 3;
    const bool mn = (idesc >> 16) & 1;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float sum = 0.f;
            for (int k = 0; k < 16; ++k)
                sum += emu_smem_bf16(a_start + (uint32_t)(m / 8) * 1024 + (uint32_t)(m % 8) * 128 + (uint32_t)k * 2) * emu_b_element(b_start, mn, n, k);
            emu_mma_store(tmem_d, m, n, sum, accumulate);
        }
}
// the same with A read from tensor memory: row m = lane m, K = 16 bf16 packed in 8 consecutive 32-bit columns
inline void emu_umma_ts_now(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate);
inline void emu_umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    emu_issue_pipe([=]() { emu_umma_ts_now(tmem_d, tmem_a, bdesc, idesc, accumulate); });
}
inline void emu_umma_ts_now(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    const uint32_t b_start = (uint32_t)(bdesc & 0x3FFF) << 4;
    const int M = (int)((idesc >> 24) & 0x1F) << 4, N = (int)((idesc >> 17) & 0x3F) << 3;
    const bool mn = (idesc >> 16) & 1;
    for (int m = 0; m < M; ++m) {
        float a[16];
        for (int k = 0; k < 16; ++k) {
            const uint32_t w = emu_tmem[(tmem_a >> 16) + m][(tmem_a & 0xFFFF) + k / 2], b = ((k & 1) ? w >> 16 : w & 0xFFFFu) << 16;
            std::memcpy(&a[k], &b, 4);
        }
        for (int n = 0; n < N; ++n) {
            float sum = 0.f;
            for (int k = 0; k < 16; ++k) sum += a[k] * emu_b_element(b_start, mn, n, k);
            emu_mma_store(tmem_d, m, n, sum, accumulate);
        }
    }
}
inline void umma_commit(uint64_t* bar) { emu_issue_pipe([=]() { mbar_arrive(bar); }); }  // arrives once everything issued before it has executed

template <int NCOL>
inline void emu_tmem_ld(uint32_t taddr, uint32_t (&v)[NCOL]) {
    const uint32_t lane = (taddr >> 16) + (uint32_t)(emu_tid & 31), col = taddr & 0xFFFF;
    for (int j = 0; j < NCOL; ++j) v[j] = emu_tmem[lane][col + j];
}
inline void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) { emu_tmem_ld<16>(taddr, v); }
inline void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) { emu_tmem_ld<32>(taddr, v); }
inline void emu_tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    const uint32_t lane = (taddr >> 16) + (uint32_t)(emu_tid & 31), col = taddr & 0xFFFF;
    for (int j = 0; j < 16; ++j) emu_tmem[lane][col + j] = v[j];
}
inline void tmem_ld_wait() {}

// ---- CTA pairs (cta_group::2, clusters): not emulated -- one block runs at a time; the callers keep the pair kernels off ----
inline uint32_t cluster_ctarank() { return 0; }
inline void cluster_sync_all() { std::abort(); }
inline uint32_t mapa(uint32_t local, uint32_t) { return local; }
inline void mbar_arrive_cluster(uint32_t) { std::abort(); }
inline void tma_load_4d_2cta(void*, const CUtensorMap*, uint32_t, int, int, int, int) { std::abort(); }
inline void tmem_alloc_2cta(uint32_t*, uint32_t) { std::abort(); }
inline void tmem_dealloc_2cta(uint32_t, uint32_t) { std::abort(); }
inline void umma_bf16_2cta(uint32_t, uint64_t, uint64_t, uint32_t, bool) { std::abort(); }
inline void umma_commit_2cta(uint64_t*) { std::abort(); }

// global -> shared bulk copy counted on an mbarrier (no swizzle)
inline void emu_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    emu_issue_copy([=]() {
        std::memcpy(smem_dst, gmem_src, bytes);
        emu_complete_tx(bar, bytes);
    });
}

}  // namespace tc
