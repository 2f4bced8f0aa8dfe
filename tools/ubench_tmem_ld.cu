// Micro-benchmark: what does reading fp32 scores out of tensor memory cost?  (flash attention:
// every 128 x 128 key block moves 64 KB of S from TMEM to registers before its exponentials can run.)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/ubench_tmem_ld.cu -o tools/ubench_tmem_ld.bin
// One CTA per SM allocates 512 TMEM columns; W warps (4 / 8 / 16: one, two, four per lane quadrant) loop over
// tcgen05.ld.32x32b.x32 (4 KB per warp instruction) with a wait every instruction (dependent) or every fourth
// (pipelined).  Prints bytes per clock per SM; the same for tcgen05.st.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int WARPS, int DEPTH, bool STORE>
__global__ void __launch_bounds__(WARPS * 32) k(long long* clocks, uint32_t* sink, int iters) {
    __shared__ uint32_t tmem_ptr;
    __shared__ long long dt[WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_ptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_ptr + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
    uint32_t acc = lane;
    uint32_t v[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] = lane + e;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            const uint32_t addr = base + (uint32_t)((d & 1) * 32);
            if (STORE) {
                asm volatile(
                    "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(addr),
                    "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
                    "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
                    "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
                    "r"(v[31])
                    : "memory");
            } else {
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                      "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                      "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                      "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(addr));
            }
        }
        if (STORE) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        else asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc ^= v[0] ^ v[31];  // static indices: a dynamic one would push v[] into local memory and time that instead
    }
    const long long t1 = clock64();
    if (lane == 0) dt[warp] = t1 - t0;
    if (acc == 0xdeadbeefu) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        long long m = 0;
        for (int w = 0; w < WARPS; ++w) m = dt[w] > m ? dt[w] : m;
        clocks[blockIdx.x] = m;
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_ptr) : "memory");
}

template <int WARPS, int DEPTH, bool STORE>
void run() {
    const int grid = 148, iters = 4096;
    long long* clk;
    uint32_t* sink;
    cudaMalloc(&clk, grid * sizeof(long long));
    cudaMalloc(&sink, 4);
    k<WARPS, DEPTH, STORE><<<grid, WARPS * 32>>>(clk, sink, iters);
    k<WARPS, DEPTH, STORE><<<grid, WARPS * 32>>>(clk, sink, iters);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), clk, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (long long v : h) avg += (double)v;
    avg /= grid;
    const double bytes = (double)iters * DEPTH * WARPS * 4096.0;
    printf("%s  warps %2d  %d per wait : %7.1f B/clk/SM  (%.0f clk per 4 KB warp access, %s)\n", STORE ? "tcgen05.st" : "tcgen05.ld", WARPS, DEPTH,
           bytes / avg, avg / ((double)iters * DEPTH), cudaGetErrorString(e));
    cudaFree(clk);
    cudaFree(sink);
}

int main() {
    run<4, 1, false>();
    run<4, 4, false>();
    run<8, 1, false>();
    run<8, 4, false>();
    run<16, 1, false>();
    run<16, 4, false>();
    run<4, 4, true>();
    run<8, 4, true>();
    run<16, 4, true>();
    return 0;
}
