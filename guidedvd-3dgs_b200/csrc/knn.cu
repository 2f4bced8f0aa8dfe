// knn.cu -- exact 3-nearest-neighbour search for sm_100a (include/gvd_knn.h).
//
// Replaces submodules/simple-knn (simple_knn.cu:192-228). Same result, different machine mapping:
//   1. bounding box by a grid-stride reduction (no host round trip; the reference copies min/max to the host)
//   2. 30-bit Morton codes + CUB radix sort (simple_knn.cu:45-70,217-220)
//   3. points gathered into Morton order as float4 (x, y, z, original index) and a two-level AABB tree:
//      leaves = 32 consecutive points (one coalesced 512-byte row per warp), nodes = 32 leaves
//      (the reference has one level of 1024-point boxes and brute-forces whole boxes, simple_knn.cu:78-117,150-190)
//   4. one warp per query point: the 32 lanes test 32 boxes, or evaluate 32 candidate points, per step, and a
//      warp-wide k-select (REDUX.MIN on the distance bits + ballot) keeps the three best.
#include <cub/cub.cuh>
#include <cfloat>
#include <cstdint>
#include <string>

#include "../../include/gvd_knn.h"

namespace {

thread_local std::string g_knn_err;

struct Box {
    float3 lo, hi;
};

template <typename T>
void obtain(char*& chunk, T*& ptr, size_t count, size_t alignment = 128) {
    size_t offset = (reinterpret_cast<uintptr_t>(chunk) + alignment - 1) & ~(alignment - 1);
    ptr = reinterpret_cast<T*>(offset);
    chunk = reinterpret_cast<char*>(ptr + count);
}

struct KnnTmp {
    float* bbox;  // [6] min xyz, max xyz (as ordered-int encoded floats during the reduction)
    uint32_t *codes, *codes_sorted, *idx, *idx_sorted;
    float4* spts;
    Box *leaf, *node;
    char* sort_temp;
    size_t sort_temp_bytes;
    size_t n_leaf, n_node;
};

KnnTmp carve(char*& p, size_t P) {
    KnnTmp t;
    obtain(p, t.bbox, 8);
    obtain(p, t.codes, P);
    obtain(p, t.codes_sorted, P);
    obtain(p, t.idx, P);
    obtain(p, t.idx_sorted, P);
    obtain(p, t.spts, P);
    t.n_leaf = (P + 31) / 32;
    t.n_node = (t.n_leaf + 31) / 32;
    obtain(p, t.leaf, t.n_leaf);
    obtain(p, t.node, t.n_node);
    t.sort_temp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t.sort_temp_bytes, t.codes, t.codes_sorted, t.idx, t.idx_sorted, (int)P);
    obtain(p, t.sort_temp, t.sort_temp_bytes);
    return t;
}

// order-preserving float <-> int mapping so that atomicMin/atomicMax on ints reduce floats
__device__ __forceinline__ int f2ord(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void bbox_init_kernel(int* bbox) {
    if (threadIdx.x < 3) bbox[threadIdx.x] = f2ord(FLT_MAX);
    else if (threadIdx.x < 6) bbox[threadIdx.x] = f2ord(-FLT_MAX);
}

__global__ void __launch_bounds__(256) bbox_kernel(int P, const float* __restrict__ xyz, int* __restrict__ bbox) {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float v = xyz[3 * (size_t)i + k];
            lo[k] = fminf(lo[k], v);
            hi[k] = fmaxf(hi[k], v);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            atomicMin(&bbox[k], f2ord(lo[k]));
            atomicMax(&bbox[3 + k], f2ord(hi[k]));
        }
    }
}

// simple_knn.cu:45-61
__device__ __forceinline__ uint32_t prep_morton(uint32_t x) {
    x = (x | (x << 16)) & 0x030000FF;
    x = (x | (x << 8)) & 0x0300F00F;
    x = (x | (x << 4)) & 0x030C30C3;
    x = (x | (x << 2)) & 0x09249249;
    return x;
}

__global__ void __launch_bounds__(256) morton_kernel(int P, const float* __restrict__ xyz, const int* __restrict__ bbox,
                                                     uint32_t* __restrict__ codes, uint32_t* __restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    uint32_t c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float lo = ord2f(bbox[k]), hi = ord2f(bbox[3 + k]);
        const float ext = hi - lo;
        // a flat axis (ext == 0) divides by zero in the reference; any fixed cell is as good there
        const float u = ext > 0.f ? (xyz[3 * (size_t)i + k] - lo) / ext : 0.f;
        c[k] = prep_morton((uint32_t)(fminf(fmaxf(u, 0.f), 1.f) * ((1 << 10) - 1)));
    }
    codes[i] = c[0] | (c[1] << 1) | (c[2] << 2);
    idx[i] = i;
}

// gather into Morton order + leaf boxes (one warp = one leaf of 32 points)
__global__ void __launch_bounds__(256) leaf_kernel(int P, const float* __restrict__ xyz,
                                                   const uint32_t* __restrict__ idx_sorted, float4* __restrict__ spts,
                                                   Box* __restrict__ leaf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float3 p = {0, 0, 0};
    const bool ok = i < P;
    if (ok) {
        const uint32_t o = idx_sorted[i];
        p = {xyz[3 * (size_t)o], xyz[3 * (size_t)o + 1], xyz[3 * (size_t)o + 2]};
        spts[i] = make_float4(p.x, p.y, p.z, __uint_as_float(o));
    }
    float3 lo = ok ? p : make_float3(FLT_MAX, FLT_MAX, FLT_MAX), hi = ok ? p : make_float3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo.x = fminf(lo.x, __shfl_xor_sync(0xffffffffu, lo.x, o));
        lo.y = fminf(lo.y, __shfl_xor_sync(0xffffffffu, lo.y, o));
        lo.z = fminf(lo.z, __shfl_xor_sync(0xffffffffu, lo.z, o));
        hi.x = fmaxf(hi.x, __shfl_xor_sync(0xffffffffu, hi.x, o));
        hi.y = fmaxf(hi.y, __shfl_xor_sync(0xffffffffu, hi.y, o));
        hi.z = fmaxf(hi.z, __shfl_xor_sync(0xffffffffu, hi.z, o));
    }
    if ((threadIdx.x & 31) == 0 && (i >> 5) < (P + 31) / 32) leaf[i >> 5] = {lo, hi};
}

__global__ void __launch_bounds__(256) node_kernel(int n_leaf, const Box* __restrict__ leaf, Box* __restrict__ node) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    Box b = {{FLT_MAX, FLT_MAX, FLT_MAX}, {-FLT_MAX, -FLT_MAX, -FLT_MAX}};
    if (i < n_leaf) b = leaf[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        b.lo.x = fminf(b.lo.x, __shfl_xor_sync(0xffffffffu, b.lo.x, o));
        b.lo.y = fminf(b.lo.y, __shfl_xor_sync(0xffffffffu, b.lo.y, o));
        b.lo.z = fminf(b.lo.z, __shfl_xor_sync(0xffffffffu, b.lo.z, o));
        b.hi.x = fmaxf(b.hi.x, __shfl_xor_sync(0xffffffffu, b.hi.x, o));
        b.hi.y = fmaxf(b.hi.y, __shfl_xor_sync(0xffffffffu, b.hi.y, o));
        b.hi.z = fmaxf(b.hi.z, __shfl_xor_sync(0xffffffffu, b.hi.z, o));
    }
    if ((threadIdx.x & 31) == 0 && (i >> 5) < (n_leaf + 31) / 32) node[i >> 5] = b;
}

// simple_knn.cu:119-129
__device__ __forceinline__ float box_dist2(const Box& b, const float3& p) {
    float3 d = {0, 0, 0};
    if (p.x < b.lo.x || p.x > b.hi.x) d.x = fminf(fabsf(p.x - b.lo.x), fabsf(p.x - b.hi.x));
    if (p.y < b.lo.y || p.y > b.hi.y) d.y = fminf(fabsf(p.y - b.lo.y), fabsf(p.y - b.hi.y));
    if (p.z < b.lo.z || p.z > b.hi.z) d.z = fminf(fabsf(p.z - b.lo.z), fabsf(p.z - b.hi.z));
    return d.x * d.x + d.y * d.y + d.z * d.z;
}

// Warp-wide 3-select. Every lane offers one candidate (dist, idx); best[] / bidx[] are warp-uniform and sorted
// ascending. Distances are >= 0, so their bit patterns order like the floats and REDUX.MIN (one instruction)
// finds the minimum; ties go to the lowest lane, and an equal distance never displaces an earlier one
// (strict '>' as in updateKBest, simple_knn.cu:131-148).
__device__ __forceinline__ void warp_select3(float d, uint32_t id, float (&best)[3], uint32_t (&bidx)[3]) {
    unsigned live = __ballot_sync(0xffffffffu, d < best[2]);
    while (live) {
        const uint32_t mbits = __reduce_min_sync(0xffffffffu, (d < best[2]) ? __float_as_uint(d) : 0x7f800000u);
        const float m = __uint_as_float(mbits);
        const unsigned who = __ballot_sync(0xffffffffu, (d < best[2]) && __float_as_uint(d) == mbits);
        const int src = __ffs(who) - 1;
        const uint32_t mid = __shfl_sync(0xffffffffu, id, src);
        if (m < best[0]) {
            best[2] = best[1]; bidx[2] = bidx[1];
            best[1] = best[0]; bidx[1] = bidx[0];
            best[0] = m; bidx[0] = mid;
        } else if (m < best[1]) {
            best[2] = best[1]; bidx[2] = bidx[1];
            best[1] = m; bidx[1] = mid;
        } else {
            best[2] = m; bidx[2] = mid;
        }
        if ((int)(threadIdx.x & 31) == src) d = FLT_MAX;
        live = __ballot_sync(0xffffffffu, d < best[2]);
    }
}

__device__ __forceinline__ float pt_dist2(const float4 c, const float3 q) {
    // same expression as updateKBest (simple_knn.cu:134-135): point - ref
    const float3 d = {c.x - q.x, c.y - q.y, c.z - q.z};
    return d.x * d.x + d.y * d.y + d.z * d.z;
}

__global__ void __launch_bounds__(256) knn3_kernel(int P, int n_leaf, int n_node, const float4* __restrict__ spts,
                                                   const Box* __restrict__ leaf, const Box* __restrict__ node,
                                                   float* __restrict__ mean_d2, int32_t* __restrict__ idx3) {
    const int lane = threadIdx.x & 31;
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // one warp per query (Morton position q)
    if (q >= P) return;
    const float4 me = spts[q];
    const float3 qp = {me.x, me.y, me.z};
    float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    uint32_t bidx[3] = {0, 0, 0};

    // seed: the query's own leaf (its Morton neighbours)
    const int my_leaf = q >> 5;
    {
        const int c = my_leaf * 32 + lane;
        float d = FLT_MAX;
        uint32_t id = 0;
        if (c < P && c != q) {
            const float4 cp = spts[c];
            d = pt_dist2(cp, qp);
            id = __float_as_uint(cp.w);
        }
        warp_select3(d, id, best, bidx);
    }
    // exact search: descend every node / leaf whose box may still hold something closer than the current third best
    for (int nb = 0; nb < n_node; nb += 32) {
        const int ni = nb + lane;
        const bool nvalid = ni < n_node;  // explicit validity: best[2] may still be FLT_MAX, a sentinel would pass "<="
        const float nd = nvalid ? box_dist2(node[ni], qp) : FLT_MAX;
        unsigned nm = __ballot_sync(0xffffffffu, nvalid && nd <= best[2]);
        while (nm) {
            const int n = nb + __ffs(nm) - 1;
            nm &= nm - 1;
            const int li = n * 32 + lane;
            const bool lvalid = li < n_leaf && li != my_leaf;
            const float ld = lvalid ? box_dist2(leaf[li], qp) : FLT_MAX;
            unsigned lm = __ballot_sync(0xffffffffu, lvalid && ld <= best[2]);
            while (lm) {
                const int l = n * 32 + __ffs(lm) - 1;
                lm &= lm - 1;
                // re-check against the (possibly tightened) bound; the box distance of leaf l sits in lane l%32
                const float ldl = __shfl_sync(0xffffffffu, ld, l & 31);
                if (ldl > best[2]) continue;
                const int c = l * 32 + lane;
                float d = FLT_MAX;
                uint32_t id = 0;
                if (c < P) {
                    const float4 cp = spts[c];
                    d = pt_dist2(cp, qp);
                    id = __float_as_uint(cp.w);
                }
                warp_select3(d, id, best, bidx);
            }
        }
    }
    if (lane == 0) {
        const uint32_t o = __float_as_uint(me.w);
        mean_d2[o] = (best[0] + best[1] + best[2]) / 3.0f;
        idx3[3 * (size_t)o + 0] = (int32_t)bidx[0];
        idx3[3 * (size_t)o + 1] = (int32_t)bidx[1];
        idx3[3 * (size_t)o + 2] = (int32_t)bidx[2];
    }
}

}  // namespace

extern "C" {

const char* gvd_knn_last_error(void) { return g_knn_err.c_str(); }

size_t gvd_knn3_tmp_bytes(int P) {
    char* p = nullptr;
    carve(p, (size_t)(P > 0 ? P : 1));
    return (size_t)p + 128;
}

int gvd_knn3(int P, const float* xyz, float* mean_d2, int32_t* idx3, void* tmp, size_t tmp_bytes,
             gvd_knn_stream_t stream_) {
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
    if (P <= 0) return 0;
    if (!xyz || !mean_d2 || !idx3 || !tmp) {
        g_knn_err = "gvd_knn3: null pointer";
        return 2;
    }
    if (tmp_bytes < gvd_knn3_tmp_bytes(P)) {
        g_knn_err = "gvd_knn3: scratch too small";
        return 2;
    }
    char* p = reinterpret_cast<char*>(tmp);
    KnnTmp t = carve(p, (size_t)P);
    int* bbox = reinterpret_cast<int*>(t.bbox);
    bbox_init_kernel<<<1, 32, 0, s>>>(bbox);
    bbox_kernel<<<std::min((P + 255) / 256, 148 * 8), 256, 0, s>>>(P, xyz, bbox);
    morton_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, xyz, bbox, t.codes, t.idx);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(t.sort_temp, t.sort_temp_bytes, t.codes, t.codes_sorted, t.idx,
                                                    t.idx_sorted, P, 0, 30, s);
    if (e != cudaSuccess) {
        g_knn_err = std::string("gvd_knn3 sort: ") + cudaGetErrorString(e);
        return 1;
    }
    const int n_leaf = (int)t.n_leaf, n_node = (int)t.n_node;
    leaf_kernel<<<(n_leaf * 32 + 255) / 256, 256, 0, s>>>(P, xyz, t.idx_sorted, t.spts, t.leaf);
    node_kernel<<<(n_node * 32 + 255) / 256, 256, 0, s>>>(n_leaf, t.leaf, t.node);
    const long long threads = (long long)P * 32;
    knn3_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(P, n_leaf, n_node, t.spts, t.leaf, t.node, mean_d2, idx3);
    e = cudaGetLastError();
    if (e != cudaSuccess) {
        g_knn_err = std::string("gvd_knn3: ") + cudaGetErrorString(e);
        return 1;
    }
    return 0;
}

}  // extern "C"
