#pragma once
#include "cuda.h"
typedef CUresult (*PFN_cuTensorMapEncodeTiled_v12000)(CUtensorMap*, int, unsigned, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                                       const cuuint32_t*, int, int, int, int);
