// Cross-GPU gradient sum over NVLink peer memory (include/gvd_exchange.h).  sm_100a, one process per GPU.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <ctime>
#include <string>

#include "../../include/gvd_exchange.h"

extern thread_local std::string g_raster_err;  // raster_api.cu: text behind gvd_last_error()

namespace {

// Signal words of one rank (uint32 each): [0, 8) ready[q], [8, 16) done[q], [16] CTA ticket.
constexpr int kReady = 0, kDone = 8, kTicket = 16, kTimedOut = 17;
// A peer that never launches (crashed rank, mismatched call sequence) must not wedge this GPU: every flag wait gives
// up after this many nanoseconds of %globaltimer, records the epoch in kTimedOut (gvd_exchange_status reads it) and
// lets the kernel finish with an unspecified payload.
constexpr unsigned long long kWaitLimitNs = 20ull * 1000 * 1000 * 1000;

struct Params {
    float4* buf[GVD_EXCHANGE_MAX_RANKS];
    uint32_t* flag[GVD_EXCHANGE_MAX_RANKS];
    float4* mc;  // multicast (NVLS) mapping of the same buffer on all ranks, or null
    int world, rank;
    unsigned long long n_vec;  // float4 elements to reduce
    uint32_t epoch;
};

#ifndef GVD_HOST_EMU
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Payload accesses are plain (weak) 16-byte vector loads/stores -- system-scope accesses would each be ordered
// individually and run at a fraction of the NVLink rate. Correctness comes from the barriers around them: every address
// is read exactly once per call, after the entry barrier (acquire + bar.sync), so no stale L1 line can exist (L1 is
// invalidated at kernel start and L1::no_allocate keeps the peer lines out of it); the stores are published by the
// __threadfence_system() in front of the exit barrier's release.
__device__ __forceinline__ float4 ld_peer(const float4* p) {
    float4 v;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_peer(float4* p, float4 v) {
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// NVLS: one load returns the SUM of the addressed 16 bytes over every GPU bound to the multicast object (the NVSwitch
// fetches and adds the replicas), one store writes all replicas.  SASS: LDGMC.ADD / STG via the multicast address.
__device__ __forceinline__ float4 mc_ld_reduce(const float4* p) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st(float4* p, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#else  // GVD_HOST_EMU (tests/cuda_emu): ranks are processes sharing host memory; same ordering with C++ atomics
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
__device__ __forceinline__ float4 ld_peer(const float4* p) { return *p; }
__device__ __forceinline__ void st_peer(float4* p, float4 v) { *p = v; }
__device__ __forceinline__ float4 mc_ld_reduce(const float4* p) { return *p; }  // no multicast objects on the host build
__device__ __forceinline__ void mc_st(float4*, float4) {}
__device__ __forceinline__ unsigned long long global_ns() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
}
__device__ __forceinline__ void __nanosleep(unsigned) { emu_yield(); }
#endif
// Spins until *p has reached `epoch` (wrap-safe); false = gave up.
__device__ __forceinline__ bool wait_flag(const uint32_t* p, uint32_t epoch, uint32_t* timed_out_word) {
    const unsigned long long t0 = global_ns();
    unsigned spins = 0;
    while ((int32_t)(ld_acquire_sys(p) - epoch) < 0) {
        __nanosleep(64);
        if ((++spins & 1023u) == 0 && global_ns() - t0 > kWaitLimitNs) {
            *timed_out_word = epoch;
            return false;
        }
    }
    return true;
}

template <int WORLD>
__device__ __forceinline__ void reduce_slice(const Params& p, unsigned long long begin, unsigned long long end) {
    constexpr int U = WORLD <= 2 ? 8 : (WORLD <= 4 ? 2 : 1);  // independent 16-byte peer loads in flight per thread: U * WORLD
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long i = begin + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < end; i += U * stride) {
        float4 v[U][WORLD];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int q = 0; q < WORLD; ++q) v[u][q] = ld_peer(p.buf[q] + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float4 a = v[u][0];
#pragma unroll
            for (int q = 1; q < WORLD; ++q) {  // fixed order: the owner of a slice is the only one that adds it up
                a.x += v[u][q].x; a.y += v[u][q].y; a.z += v[u][q].z; a.w += v[u][q].w;
            }
#pragma unroll
            for (int q = 0; q < WORLD; ++q) st_peer(p.buf[q] + i + u * stride, a);
        }
    }
    for (; i < end; i += stride) {
        float4 a = ld_peer(p.buf[0] + i);
#pragma unroll
        for (int q = 1; q < WORLD; ++q) {
            const float4 b = ld_peer(p.buf[q] + i);
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
#pragma unroll
        for (int q = 0; q < WORLD; ++q) st_peer(p.buf[q] + i, a);
    }
}

// Slice r of the buffer, summed inside the NVSwitch and written back to every replica through the multicast mapping:
// per element ONE 16-byte request leaves this GPU in each direction, instead of WORLD-1 peer loads and WORLD-1 peer
// stores.  Link traffic per GPU: (WORLD-1)/WORLD of the buffer out (its replicas of the other ranks' slices, pulled by
// the switch) and the same amount in (the other ranks' sums, multicast by the switch) -- the two directions overlap.
__device__ __forceinline__ void reduce_slice_nvls(const Params& p, unsigned long long begin, unsigned long long end) {
    constexpr int U = 4;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long i = begin + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < end; i += U * stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = mc_ld_reduce(p.mc + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u) mc_st(p.mc + i + u * stride, v[u]);
    }
    for (; i < end; i += stride) mc_st(p.mc + i, mc_ld_reduce(p.mc + i));
}

template <int WORLD>
__global__ void __launch_bounds__(512) grad_allreduce_kernel(const Params p) {
    uint32_t* mine = p.flag[p.rank];
    // ---- entry: every peer has finished producing its gradients (its kernel is stream-ordered behind them) ----
    if (blockIdx.x == 0 && threadIdx.x < WORLD) st_release_sys(p.flag[threadIdx.x] + kReady + p.rank, p.epoch);
    if (threadIdx.x < WORLD) wait_flag(mine + kReady + threadIdx.x, p.epoch, mine + kTimedOut);
    __syncthreads();
    // ---- rank r sums slice r of every buffer and stores the sum into every buffer ----
    const unsigned long long per = (p.n_vec + WORLD - 1) / WORLD;
    const unsigned long long begin = per * p.rank;
    unsigned long long end = begin + per;
    if (end > p.n_vec) end = p.n_vec;
    if (begin < end) {
        if (p.mc != nullptr) reduce_slice_nvls(p, begin, end);
        else reduce_slice<WORLD>(p, begin, end);
    }
    // ---- exit: my stores have landed everywhere, and every peer's stores have landed here ----
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = atomicAdd(mine + kTicket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence_system();
    if (threadIdx.x == 0) mine[kTicket] = 0u;
    if (threadIdx.x < WORLD) {
        st_release_sys(p.flag[threadIdx.x] + kDone + p.rank, p.epoch);
        wait_flag(mine + kDone + threadIdx.x, p.epoch, mine + kTimedOut);
    }
}

int fail(const char* what, cudaError_t e) {
    g_raster_err = std::string(what) + ": " + cudaGetErrorString(e);
    return 1;
}

size_t padded(size_t payload) { return (payload + 15) / 16 * 16; }

}  // namespace

extern "C" {

int gvd_exchange_alloc(size_t payload_bytes, void** dev_ptr, unsigned char* handle) {
    if (!dev_ptr || !handle) { g_raster_err = "gvd_exchange_alloc: null argument"; return 2; }
    const size_t total = padded(payload_bytes) + GVD_EXCHANGE_FLAG_BYTES;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, total);
    if (e != cudaSuccess) return fail("gvd_exchange_alloc: cudaMalloc", e);
    if ((e = cudaMemset(p, 0, total)) != cudaSuccess) { cudaFree(p); return fail("gvd_exchange_alloc: cudaMemset", e); }
    cudaIpcMemHandle_t h;
    if ((e = cudaIpcGetMemHandle(&h, p)) != cudaSuccess) { cudaFree(p); return fail("gvd_exchange_alloc: cudaIpcGetMemHandle", e); }
    static_assert(sizeof(cudaIpcMemHandle_t) == GVD_EXCHANGE_HANDLE_BYTES, "handle size");
    std::memcpy(handle, &h, sizeof(h));
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) { cudaFree(p); return fail("gvd_exchange_alloc: sync", e); }
    *dev_ptr = p;
    return 0;
}

int gvd_exchange_free(void* dev_ptr) {
    if (!dev_ptr) return 0;
    cudaError_t e = cudaFree(dev_ptr);
    return e == cudaSuccess ? 0 : fail("gvd_exchange_free", e);
}

int gvd_exchange_open(const unsigned char* handle, void** peer_ptr) {
    if (!handle || !peer_ptr) { g_raster_err = "gvd_exchange_open: null argument"; return 2; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail("gvd_exchange_open: cudaIpcOpenMemHandle", e);
    *peer_ptr = p;
    return 0;
}

int gvd_exchange_close(void* peer_ptr) {
    if (!peer_ptr) return 0;
    cudaError_t e = cudaIpcCloseMemHandle(peer_ptr);
    return e == cudaSuccess ? 0 : fail("gvd_exchange_close", e);
}

int gvd_exchange_status(const void* own_ptr, size_t payload_bytes, uint32_t* timed_out_epoch) {
    if (!own_ptr || !timed_out_epoch) { g_raster_err = "gvd_exchange_status: null argument"; return 2; }
    const char* word = reinterpret_cast<const char*>(own_ptr) + padded(payload_bytes) + kTimedOut * sizeof(uint32_t);
    cudaError_t e = cudaMemcpy(timed_out_epoch, word, sizeof(uint32_t), cudaMemcpyDeviceToHost);
    return e == cudaSuccess ? 0 : fail("gvd_exchange_status", e);
}

int gvd_exchange_allreduce_sum(const GvdExchangeArgs* a, void* stream_) {
    if (!a) { g_raster_err = "gvd_exchange_allreduce_sum: null args"; return 2; }
    if (a->world < 2 || a->world > GVD_EXCHANGE_MAX_RANKS || a->rank < 0 || a->rank >= a->world) {
        g_raster_err = "gvd_exchange_allreduce_sum: world must be 2..8 and 0 <= rank < world";
        return 2;
    }
    if (a->n_floats % 4 != 0 || a->n_floats * sizeof(float) > padded(a->payload_bytes)) {
        g_raster_err = "gvd_exchange_allreduce_sum: n_floats must be a multiple of 4 and fit the payload";
        return 2;
    }
    Params p;
    std::memset(&p, 0, sizeof(p));
    for (int q = 0; q < a->world; ++q) {
        if (!a->bufs[q]) { g_raster_err = "gvd_exchange_allreduce_sum: null buffer pointer"; return 2; }
        p.buf[q] = reinterpret_cast<float4*>(a->bufs[q]);
        p.flag[q] = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(a->bufs[q]) + padded(a->payload_bytes));
    }
    p.mc = reinterpret_cast<float4*>(a->multicast);
    p.world = a->world;
    p.rank = a->rank;
    p.n_vec = a->n_floats / 4;
    p.epoch = a->epoch;
    if (p.n_vec == 0) return 0;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    // All CTAs spin in the entry barrier, so the grid must be co-resident: 2 CTAs of 512 threads per SM at most.
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned long long per = (p.n_vec + a->world - 1) / a->world;
    unsigned long long want = (per + 512 * 4 - 1) / (512 * 4);
    int grid = (int)(want < 1 ? 1 : (want > (unsigned long long)(2 * sms) ? 2 * sms : want));
    switch (a->world) {
#define GVD_CASE(W) case W: grad_allreduce_kernel<W><<<grid, 512, 0, s>>>(p); break;
        GVD_CASE(2) GVD_CASE(3) GVD_CASE(4) GVD_CASE(5) GVD_CASE(6) GVD_CASE(7) GVD_CASE(8)
#undef GVD_CASE
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail("gvd_exchange_allreduce_sum: launch", e);
}

}  // extern "C"
