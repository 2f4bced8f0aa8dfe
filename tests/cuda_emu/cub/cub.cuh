// Host stand-in for the two CUB device-wide primitives the product calls (tests/cuda_emu): same call protocol
// (first call with d_temp_storage == nullptr returns the scratch size), results of a stable LSD radix sort / scan.
#pragma once
#include <algorithm>
#include <numeric>
#include <vector>

#include "../cuda_emu.h"

namespace cub {
struct DeviceRadixSort {
    template <class K, class V>
    static cudaError_t SortPairs(void* tmp, size_t& tmp_bytes, const K* kin, K* kout, const V* vin, V* vout, long long n, int begin_bit = 0,
                                 int end_bit = sizeof(K) * 8, cudaStream_t = nullptr) {
        if (!tmp) { tmp_bytes = 16; return cudaSuccess; }
        std::vector<long long> idx(n);
        std::iota(idx.begin(), idx.end(), 0ll);
        const K mask = end_bit - begin_bit >= (int)sizeof(K) * 8 ? ~K(0) : (((K(1) << (end_bit - begin_bit)) - 1) << begin_bit);
        std::stable_sort(idx.begin(), idx.end(), [&](long long a, long long b) { return (kin[a] & mask) < (kin[b] & mask); });
        std::vector<K> ks(n);
        std::vector<V> vs(n);
        for (long long i = 0; i < n; ++i) { ks[i] = kin[idx[i]]; vs[i] = vin[idx[i]]; }
        std::copy(ks.begin(), ks.end(), kout);
        std::copy(vs.begin(), vs.end(), vout);
        return cudaSuccess;
    }
};
struct DeviceScan {
    template <class I, class O>
    static cudaError_t InclusiveSum(void* tmp, size_t& tmp_bytes, I in, O out, long long n, cudaStream_t = nullptr) {
        if (!tmp) { tmp_bytes = 16; return cudaSuccess; }
        std::partial_sum(in, in + n, out);
        return cudaSuccess;
    }
};
}  // namespace cub
