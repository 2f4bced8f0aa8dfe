// cuda_emu.h -- just enough of the CUDA execution model to RUN a .cu file's kernels on the host (test infrastructure).
//
// A kernel launch `k<<<grid, block, smem, stream>>>(args)` (rewritten to EMU_LAUNCH by build_emu.py) runs the blocks one
// after another; inside a block every CUDA thread is a real OS thread, so __syncthreads(), __syncwarp(), warp shuffles
// and shared-memory atomics keep their meaning: barriers are std::barrier objects (a thread that returns from the
// kernel drops out of them), a shuffle is an exchange through a per-block slot array between two warp barriers,
// `__shared__` arrays are static storage (blocks never overlap).  bf16 is emulated with round-to-nearest-even.
// What this does NOT model: memory alignment faults, register/shared-memory limits, intrinsic accuracy (__expf is expf).
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <functional>
#include <memory>
#include <string>
#include <thread>
#include <tuple>
#include <algorithm>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __grid_constant__
#define __cluster_dims__(...)
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint3_emu { unsigned x, y, z; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(8) int2 { int x, y; };
struct alignas(8) float2 { float x, y; };
struct float3 { float x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct uint3 { unsigned x, y, z; };
inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }
inline int2 make_int2(int a, int b) { return int2{a, b}; }
inline float2 make_float2(float a, float b) { return float2{a, b}; }
inline float3 make_float3(float a, float b, float c) { return float3{a, b, c}; }
inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }

inline thread_local uint3_emu threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

using std::max;
using std::min;
// CUDA's mixed-signedness overloads
inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }
inline unsigned min(int a, unsigned b) { return (unsigned)a < b ? (unsigned)a : b; }
inline unsigned max(unsigned a, int b) { return a > (unsigned)b ? a : (unsigned)b; }
inline unsigned max(int a, unsigned b) { return (unsigned)a > b ? (unsigned)a : b; }
inline long long min(long long a, int b) { return a < b ? a : b; }
inline long long max(long long a, int b) { return a > b ? a : b; }
inline unsigned long long min(unsigned long long a, unsigned b) { return a < b ? a : b; }
[[noreturn]] inline void __trap() { abort(); }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }

// ---- bf16 ----
struct __nv_bfloat16 { uint16_t v; };
struct __nv_bfloat162 { __nv_bfloat16 x, y; };
inline __nv_bfloat16 __float2bfloat16(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    __nv_bfloat16 r;
    if ((u & 0x7fffffffu) > 0x7f800000u) { r.v = 0x7fff; return r; }
    u += 0x7fffu + ((u >> 16) & 1u);
    r.v = (uint16_t)(u >> 16);
    return r;
}
inline float __bfloat162float(__nv_bfloat16 h) {
    uint32_t u = (uint32_t)h.v << 16;
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}
inline float2 __bfloat1622float2(__nv_bfloat162 h) { return float2{__bfloat162float(h.x), __bfloat162float(h.y)}; }
inline __nv_bfloat162 __floats2bfloat162_rn(float a, float b) { return __nv_bfloat162{__float2bfloat16(a), __float2bfloat16(b)}; }

// ---- math / memory intrinsics ----
inline void emu_yield() { std::this_thread::yield(); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }
template <class T> inline void __stcs(T* p, T v) { *p = v; }
template <class T> inline void __stcg(T* p, T v) { *p = v; }
template <class T> inline T __ldcs(const T* p) { return *p; }
template <class T> inline T __ldcg(const T* p) { return *p; }
inline float __fdividef(float a, float b) { return a / b; }
inline float __frcp_rn(float a) { return 1.0f / a; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline float __saturatef(float a) { return a < 0.f ? 0.f : (a > 1.f ? 1.f : a); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline int atomicAdd(int* p, int v) { return std::atomic_ref<int>(*p).fetch_add(v); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return std::atomic_ref<unsigned>(*p).fetch_add(v); }
inline unsigned atomicMax(unsigned* p, unsigned v) { std::atomic_ref<unsigned> a(*p); unsigned o = a.load(); while (v > o && !a.compare_exchange_weak(o, v)) {} return o; }
inline int atomicMax(int* p, int v) { std::atomic_ref<int> a(*p); int o = a.load(); while (v > o && !a.compare_exchange_weak(o, v)) {} return o; }
inline unsigned atomicOr(unsigned* p, unsigned v) { return std::atomic_ref<unsigned>(*p).fetch_or(v); }
inline unsigned atomicExch(unsigned* p, unsigned v) { return std::atomic_ref<unsigned>(*p).exchange(v); }
inline float __expf(float x) { return expf(x); }
inline float __log2f(float x) { return log2f(x); }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
template <class T> inline T __ldg(const T* p) { return *p; }
inline double atomicAdd(double* p, double v) { return std::atomic_ref<double>(*p).fetch_add(v); }
inline float atomicAdd(float* p, float v) { return std::atomic_ref<float>(*p).fetch_add(v); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return std::atomic_ref<unsigned long long>(*p).fetch_add(v); }
template <class T> inline T emu_atomic_min(T* p, T v) {
    std::atomic_ref<T> a(*p);
    T old = a.load();
    while (v < old && !a.compare_exchange_weak(old, v)) {}
    return old;
}
inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) { return emu_atomic_min(p, v); }
inline int atomicMin(int* p, int v) { return emu_atomic_min(p, v); }
inline long long __double2ll_rn(double d) { return llrint(d); }
inline long long __double_as_longlong(double d) { long long r; std::memcpy(&r, &d, 8); return r; }
// round-to-nearest double arithmetic that the compiler must not contract into FMAs (build_emu.py passes -ffp-contract=off)
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __ddiv_rn(double a, double b) { return a / b; }

// ---- block / warp machinery ----
struct EmuBlock {
    std::barrier<> block_bar;
    std::vector<std::unique_ptr<std::barrier<>>> warp_bars;
    std::vector<uint64_t> slots;
    std::atomic<int> vote[2] = {0, 0};
    EmuBlock(int nthreads) : block_bar(nthreads), slots(nthreads) {
        for (int w = 0; w * 32 < nthreads; ++w) warp_bars.emplace_back(new std::barrier<>(std::min(32, nthreads - w * 32)));
    }
};
inline thread_local EmuBlock* emu_blk = nullptr;
inline thread_local int emu_tid = 0;
inline thread_local unsigned emu_votes = 0;  // block-wide votes this thread has taken part in (selects the accumulator)

inline void __syncthreads() { emu_blk->block_bar.arrive_and_wait(); }
inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
// barrier + block-wide vote: every thread adds its predicate, reads the total after the barrier, and clears the
// accumulator after a second barrier (the next vote uses the other accumulator, so a fast thread cannot race the clear)
inline int __syncthreads_count(int pred) {
    std::atomic<int>& acc = emu_blk->vote[emu_votes++ & 1];
    if (pred) acc.fetch_add(1);
    emu_blk->block_bar.arrive_and_wait();
    const int r = acc.load();
    emu_blk->block_bar.arrive_and_wait();
    acc.store(0);
    return r;
}
inline int __syncthreads_or(int pred) { return __syncthreads_count(pred) != 0; }
inline int __syncthreads_and(int pred) { return __syncthreads_count(!pred) == 0; }
inline void __syncwarp(unsigned = 0xffffffffu) { emu_blk->warp_bars[emu_tid >> 5]->arrive_and_wait(); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    emu_blk->slots[emu_tid] = raw;
    __syncwarp();
    raw = emu_blk->slots[(emu_tid & ~31) | ((emu_tid ^ lane_mask) & 31)];
    __syncwarp();
    T out;
    std::memcpy(&out, &raw, sizeof(T));
    return out;
}

// generic "every lane publishes a value, then reads what it needs": all warp collectives below are built on it
template <class T, class F> inline auto emu_warp_exchange(T v, F&& reader) {
    static_assert(sizeof(T) <= 8, "warp payload");
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    emu_blk->slots[emu_tid] = raw;
    __syncwarp();
    const int base = emu_tid & ~31;
    auto get = [&](int lane) { T o; uint64_t r = emu_blk->slots[base + (lane & 31)]; std::memcpy(&o, &r, sizeof(T)); return o; };
    auto out = reader(get, emu_tid & 31);
    __syncwarp();
    return out;
}
inline int emu_warp_lanes() { return std::min(32, (int)(blockDim.x * blockDim.y * blockDim.z) - (emu_tid & ~31)); }
template <class T> inline T __shfl_sync(unsigned, T v, int src) {
    return emu_warp_exchange(v, [&](auto get, int) { return get(src); });
}
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned delta) {
    return emu_warp_exchange(v, [&](auto get, int lane) { return lane >= (int)delta ? get(lane - (int)delta) : v; });
}
template <class T> inline T __shfl_down_sync(unsigned, T v, unsigned delta) {
    return emu_warp_exchange(v, [&](auto get, int lane) { return lane + (int)delta < 32 ? get(lane + (int)delta) : v; });
}
inline unsigned __ballot_sync(unsigned, int pred) {
    return emu_warp_exchange((unsigned)(pred != 0), [&](auto get, int) { unsigned m = 0; for (int l = 0; l < emu_warp_lanes(); ++l) m |= get(l) << l; return m; });
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, !pred) == 0; }
inline unsigned __reduce_or_sync(unsigned, unsigned v) {
    return emu_warp_exchange(v, [&](auto get, int) { unsigned m = 0; for (int l = 0; l < emu_warp_lanes(); ++l) m |= get(l); return m; });
}
inline unsigned __reduce_max_sync(unsigned, unsigned v) {
    return emu_warp_exchange(v, [&](auto get, int) { unsigned m = 0; for (int l = 0; l < emu_warp_lanes(); ++l) m = std::max(m, get(l)); return m; });
}
inline unsigned __reduce_min_sync(unsigned, unsigned v) {
    return emu_warp_exchange(v, [&](auto get, int) { unsigned m = ~0u; for (int l = 0; l < emu_warp_lanes(); ++l) m = std::min(m, get(l)); return m; });
}
inline int __reduce_min_sync(unsigned, int v) {
    return emu_warp_exchange(v, [&](auto get, int) { int m = 0x7fffffff; for (int l = 0; l < emu_warp_lanes(); ++l) m = std::min(m, get(l)); return m; });
}
inline unsigned __reduce_add_sync(unsigned, unsigned v) {
    return emu_warp_exchange(v, [&](auto get, int) { unsigned m = 0; for (int l = 0; l < emu_warp_lanes(); ++l) m += get(l); return m; });
}
template <class T> inline unsigned __match_any_sync(unsigned, T v) {
    return emu_warp_exchange(v, [&](auto get, int) { unsigned m = 0; for (int l = 0; l < emu_warp_lanes(); ++l) m |= (unsigned)(get(l) == v) << l; return m; });
}
inline unsigned __activemask() { return 0xffffffffu; }

// ---- warp-level tensor-core tiles (csrc/tattn_mma.cu), fragment layouts of the PTX ISA ----
// mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32: lane (g = lane / 4, c = lane % 4) holds
//   A (16 x 16, row):  a0 = (row g, k 2c..2c+1)  a1 = (row g + 8, same k)  a2 = (row g, k 2c + 8..)  a3 = (row g + 8, k 2c + 8..)
//   B (16 x 8, col):   b0 = (k 2c..2c+1, n g)    b1 = (k 2c + 8.., n g)
//   C / D (16 x 8):    d0 = (row g, n 2c)  d1 = (row g, n 2c + 1)  d2 = (row g + 8, n 2c)  d3 = (row g + 8, n 2c + 1)
// All 32 lanes of the warp must call it (it is built on the warp exchange above); fp32 accumulation in k order.
inline float emu_bf16_bits(uint32_t pair, int hi) { uint32_t b = (hi ? pair >> 16 : pair & 0xffffu) << 16; float f; std::memcpy(&f, &b, 4); return f; }
inline void emu_mma_m16n8k16_bf16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    const int lane = emu_tid & 31, g = lane >> 2, c = lane & 3;
    float arow[2][16], bcol[2][16];  // rows g, g + 8 of A; columns 2c, 2c + 1 of B
    emu_warp_exchange((uint64_t)a0 | ((uint64_t)a1 << 32), [&](auto get, int) {
        for (int k = 0; k < 8; ++k) {
            const uint64_t r = get(g * 4 + k / 2);
            arow[0][k] = emu_bf16_bits((uint32_t)r, k & 1);
            arow[1][k] = emu_bf16_bits((uint32_t)(r >> 32), k & 1);
        }
        return 0;
    });
    emu_warp_exchange((uint64_t)a2 | ((uint64_t)a3 << 32), [&](auto get, int) {
        for (int k = 0; k < 8; ++k) {
            const uint64_t r = get(g * 4 + k / 2);
            arow[0][8 + k] = emu_bf16_bits((uint32_t)r, k & 1);
            arow[1][8 + k] = emu_bf16_bits((uint32_t)(r >> 32), k & 1);
        }
        return 0;
    });
    emu_warp_exchange((uint64_t)b0 | ((uint64_t)b1 << 32), [&](auto get, int) {
        for (int e = 0; e < 2; ++e)
            for (int k = 0; k < 8; ++k) {
                const uint64_t r = get((2 * c + e) * 4 + k / 2);
                bcol[e][k] = emu_bf16_bits((uint32_t)r, k & 1);
                bcol[e][8 + k] = emu_bf16_bits((uint32_t)(r >> 32), k & 1);
            }
        return 0;
    });
    for (int m = 0; m < 2; ++m)
        for (int e = 0; e < 2; ++e) {
            float acc = d[2 * m + e];
            for (int k = 0; k < 16; ++k) acc += arow[m][k] * bcol[e][k];
            d[2 * m + e] = acc;
        }
}
// movmatrix.sync.aligned.m8n8.trans.b16: lane (g, c) holds row g, columns 2c..2c+1 of an 8 x 8 matrix of 16-bit elements
// and receives the same positions of its transpose, i.e. M[2c][g] and M[2c + 1][g]
inline uint32_t emu_movmatrix_trans_b16(uint32_t x) {
    const int lane = emu_tid & 31, g = lane >> 2, c = lane & 3;
    return emu_warp_exchange(x, [&](auto get, int) {
        const uint32_t lo = get((2 * c) * 4 + (g >> 1)), hi = get((2 * c + 1) * 4 + (g >> 1));
        const uint32_t e0 = (g & 1) ? lo >> 16 : lo & 0xffffu, e1 = (g & 1) ? hi >> 16 : hi & 0xffffu;
        return e0 | (e1 << 16);
    });
}

// One set of OS threads per LAUNCH, walking the blocks in order (a thread per (block, CUDA thread) cost more in thread
// creation than in kernel work: the tiny U-Net of tests/test_unet_emu_cpu.py makes ~5 000 blocks of 192-320 threads).
// Every block gets a fresh EmuBlock -- a thread that returned early has dropped out of that block's barriers for good --
// put in place by thread 0 between two launch-wide barriers, so no thread ever sees another block's state.
inline void emu_launch(dim3 grid, dim3 block, const std::function<void()>& body) {
    const int nthreads = (int)(block.x * block.y * block.z);
    const unsigned long long nblocks = (unsigned long long)grid.x * grid.y * grid.z;
    if (nblocks == 0 || nthreads == 0) return;
    std::unique_ptr<EmuBlock> cur(new EmuBlock(nthreads));
    std::barrier<> launch_bar(nthreads);
    std::vector<std::thread> ts;
    ts.reserve(nthreads);
    for (int t = 0; t < nthreads; ++t)
        ts.emplace_back([&, t]() {
            for (unsigned long long b = 0; b < nblocks; ++b) {
                EmuBlock& blk = *cur;
                emu_blk = &blk;
                emu_tid = t;
                emu_votes = 0;
                threadIdx = uint3_emu{(unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)};
                blockIdx = uint3_emu{(unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((unsigned long long)grid.x * grid.y))};
                blockDim = block;
                gridDim = grid;
                body();
                blk.warp_bars[t >> 5]->arrive_and_drop();  // a finished thread no longer takes part in barriers
                blk.block_bar.arrive_and_drop();
                launch_bar.arrive_and_wait();              // everyone has left block b
                if (t == 0 && b + 1 < nblocks) cur.reset(new EmuBlock(nthreads));
                launch_bar.arrive_and_wait();              // block b + 1's barriers and slots are in place
            }
        });
    for (auto& th : ts) th.join();
}
#define EMU_LAUNCH(kernel, g, b, smem, stream, ...) emu_launch(dim3(g), dim3(b), [=]() { kernel(__VA_ARGS__); })

// ---- runtime API stubs ----
typedef struct CUstream_st* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
template <class F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
enum { cudaFuncAttributePreferredSharedMemoryCarveout = 9, cudaMemcpyDeviceToHost = 2, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToDevice = 3,
       cudaLaunchAttributeProgrammaticStreamSerialization = 4, cudaSharedmemCarveoutMaxShared = 100, cudaSharedmemCarveoutMaxL1 = 0 };
typedef struct CUevent_st* cudaEvent_t;
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, int) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
enum { cudaEventDisableTiming = 2 };
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = reinterpret_cast<cudaEvent_t>(1); return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
enum { cudaDevAttrMultiProcessorCount = 16 };
inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 148; return cudaSuccess; }
struct cudaLaunchAttribute { int id; struct { int programmaticStreamSerializationAllowed; } val; };
struct cudaLaunchConfig_t { dim3 gridDim, blockDim; size_t dynamicSmemBytes; cudaStream_t stream; cudaLaunchAttribute* attrs; unsigned numAttrs; };
template <class... Exp, class... Act> inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t* cfg, void (*kernel)(Exp...), Act&&... args) {
    auto tup = std::make_tuple(static_cast<Exp>(args)...);
    emu_launch(cfg->gridDim, cfg->blockDim, [&]() { std::apply(kernel, tup); });
    return cudaSuccess;
}
inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
inline cudaError_t cudaMalloc(void** p, size_t n) { return posix_memalign(p, 256, n) == 0 ? cudaSuccess : 2; }
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemset(void* p, int v, size_t n) { std::memset(p, v, n); return cudaSuccess; }
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
// no inter-process mapping on the host build: the multi-rank tests share memory themselves and pass the pointers
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { std::memset(h, 0, sizeof(*h)); std::memcpy(h, &p, sizeof(p)); return cudaSuccess; }
inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { std::memcpy(p, &h, sizeof(*p)); return cudaSuccess; }
inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
