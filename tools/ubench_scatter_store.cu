// Micro-benchmark for DESIGN.md section 8 item 1: what do 3.7 M scattered 4-byte stores cost on B200, and how much of
// bin_fill's 83 us (C2) do they explain?  Standalone:  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
// tools/ubench_scatter_store.cu -o /tmp/ubench && /tmp/ubench        (prints one line per pattern)
//
// The instance list of C2 is R = 3.67 M uint32 in T = 1200 tile lists; chunk c (64 depth-consecutive Gaussians) owns
// ~1.6 consecutive slots in every list it touches, so every 32-byte sector is written by ~8 different CTAs.
//   A  scattered:  slot(i) = the real pattern's shape -- entry i of "chunk" b goes to list (hash(i,b) % T), position b*k
//   B  same slots, but each CTA writes runs of 8 consecutive slots (full sectors) -- what a chunk-group staging would do
//   C  fully coalesced stream of the same 3.67 M words (lower bound)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__global__ void fill_indexed(const uint32_t* __restrict__ slot, uint32_t* __restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[slot[i]] = (uint32_t)i;
}
__global__ void fill_stream(uint32_t* __restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (uint32_t)i;
}

static float time_kernel(void (*launch)(void*), void* ctx, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) launch(ctx);
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) launch(ctx);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms * 1000.f / reps;
}

struct Ctx { const uint32_t* slot; uint32_t* dst; int n; };
static void launch_indexed(void* p) { Ctx* c = (Ctx*)p; fill_indexed<<<(c->n + 127) / 128, 128>>>(c->slot, c->dst, c->n); }
static void launch_stream(void* p) { Ctx* c = (Ctx*)p; fill_stream<<<(c->n + 127) / 128, 128>>>(c->dst, c->n); }

int main() {
    const int T = 1200, chunks = 1928, per_chunk = 1902;  // 3.67 M instances
    const int n = chunks * per_chunk;
    const int list_len = (n + T - 1) / T;                 // every list gets ~3056 entries
    std::vector<uint32_t> slotA(n), slotB(n);
    // pattern A: thread i = (chunk b, entry k). The chunk touches the lists in a pseudo-random order, ~1.6 slots per
    // list; in list t its slots are consecutive and come right after chunk b-1's slots of that list.
    std::vector<uint32_t> fill(T, 0);
    uint64_t s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 11); };
    for (int b = 0; b < chunks; ++b)
        for (int k = 0; k < per_chunk; ++k) {
            uint32_t t = rnd() % T;
            while (fill[t] >= (uint32_t)list_len) t = (t + 1) % T;
            slotA[(size_t)b * per_chunk + k] = t * list_len + fill[t]++;
        }
    // pattern B: same total, but a "group" of 8 chunks writes runs of 8 consecutive slots of one list
    std::fill(fill.begin(), fill.end(), 0);
    const int group = 8 * per_chunk;
    for (int g0 = 0, i = 0; g0 < n; g0 += group) {
        const int m = (n - g0 < group) ? n - g0 : group;
        for (int k = 0; k < m; k += 8, i += 8) {
            uint32_t t = rnd() % T;
            while (fill[t] + 8 > (uint32_t)list_len) t = (t + 1) % T;
            for (int j = 0; j < 8 && k + j < m; ++j) slotB[g0 + k + j] = t * list_len + fill[t] + j;
            fill[t] += 8;
        }
    }
    uint32_t *dA, *dB, *dst;
    cudaMalloc(&dA, n * 4ull);
    cudaMalloc(&dB, n * 4ull);
    cudaMalloc(&dst, (size_t)T * list_len * 4ull + 64);
    cudaMemcpy(dA, slotA.data(), n * 4ull, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, slotB.data(), n * 4ull, cudaMemcpyHostToDevice);
    Ctx a{dA, dst, n}, b{dB, dst, n}, c{nullptr, dst, n};
    printf("n = %d words (%.1f MB)\n", n, n * 4.0 / 1e6);
    printf("A scattered 4-byte stores, ~8 CTAs per sector : %7.1f us\n", time_kernel(launch_indexed, &a, 50));
    printf("B runs of 8 consecutive slots (full sectors)  : %7.1f us\n", time_kernel(launch_indexed, &b, 50));
    printf("C coalesced stream                            : %7.1f us\n", time_kernel(launch_stream, &c, 50));
    return 0;
}
