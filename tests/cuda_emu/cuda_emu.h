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
#include <functional>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) alignas(n)

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint3_emu { unsigned x, y, z; };
struct uint4 { unsigned x, y, z, w; };
struct float2 { float x, y; };
inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }

inline thread_local uint3_emu threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

using std::max;
using std::min;

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
inline float __expf(float x) { return expf(x); }
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
    EmuBlock(int nthreads) : block_bar(nthreads), slots(nthreads) {
        for (int w = 0; w * 32 < nthreads; ++w) warp_bars.emplace_back(new std::barrier<>(std::min(32, nthreads - w * 32)));
    }
};
inline thread_local EmuBlock* emu_blk = nullptr;
inline thread_local int emu_tid = 0;

inline void __syncthreads() { emu_blk->block_bar.arrive_and_wait(); }
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

inline void emu_launch(dim3 grid, dim3 block, const std::function<void()>& body) {
    const int nthreads = (int)(block.x * block.y * block.z);
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                EmuBlock blk(nthreads);
                std::vector<std::thread> ts;
                ts.reserve(nthreads);
                for (int t = 0; t < nthreads; ++t)
                    ts.emplace_back([&, t]() {
                        emu_blk = &blk;
                        emu_tid = t;
                        threadIdx = uint3_emu{(unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)};
                        blockIdx = uint3_emu{bx, by, bz};
                        blockDim = block;
                        gridDim = grid;
                        body();
                        blk.warp_bars[t >> 5]->arrive_and_drop();  // a finished thread no longer takes part in barriers
                        blk.block_bar.arrive_and_drop();
                    });
                for (auto& th : ts) th.join();
            }
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
