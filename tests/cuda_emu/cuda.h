// tests/cuda_emu: host stand-in for the slice of the driver API the tensor-core kernels touch (tensor maps).
#pragma once
#include <cstdint>
typedef uint64_t cuuint64_t;
typedef uint32_t cuuint32_t;
typedef int CUresult;
enum { CUDA_SUCCESS = 0 };
enum { CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 = 9, CU_TENSOR_MAP_INTERLEAVE_NONE = 0, CU_TENSOR_MAP_SWIZZLE_128B = 3,
       CU_TENSOR_MAP_L2_PROMOTION_L2_256B = 3, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0 };
// what cuTensorMapEncodeTiled records: a strided view (innermost dimension contiguous) and the box one copy moves
struct alignas(64) CUtensorMap {
    const char* base;
    uint64_t dim[5];
    uint64_t stride[5];  // bytes; stride[0] = element size
    uint32_t box[5];
    int rank, elem_bytes, swizzle;
};
typedef CUtensorMap CUtensorMap_st;
inline CUresult emu_cuTensorMapEncodeTiled(CUtensorMap* m, int dtype, unsigned rank, void* base, const cuuint64_t* dims, const cuuint64_t* strides,
                                           const cuuint32_t* box, const cuuint32_t*, int, int swizzle, int, int) {
    if (dtype != CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 || rank < 1 || rank > 5) return 1;
    m->base = static_cast<const char*>(base);
    m->rank = (int)rank;
    m->elem_bytes = 2;
    m->swizzle = swizzle;
    for (unsigned i = 0; i < 5; ++i) {
        m->dim[i] = i < rank ? dims[i] : 1;
        m->box[i] = i < rank ? box[i] : 1;
        m->stride[i] = i == 0 ? 2 : (i < rank ? strides[i - 1] : 0);
    }
    return CUDA_SUCCESS;
}
