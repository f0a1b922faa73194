// Host-side construction of TMA descriptors (CUtensorMap) for row-major bf16 matrices.
// cuTensorMapEncodeTiled is resolved through the runtime (cudaGetDriverEntryPoint), so the
// extension does not link libcuda.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "../runtime/common.h"

namespace m4t {

inline CUresult encode_tiled_2d(CUtensorMap* map, CUtensorMapDataType dt, void* base, const cuuint64_t* dims,
                                const cuuint64_t* strides, const cuuint32_t* box, CUtensorMapSwizzle swizzle) {
  using Fn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static Fn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    M4T_CHECK(e == cudaSuccess && q == cudaDriverEntryPointSuccess && p, "cuTensorMapEncodeTiled unavailable");
    return reinterpret_cast<Fn>(p);
  }();
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, dt, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

// Row-major [rows, cols] bf16 matrix with leading dimension `ld` (elements); boxes of
// box_cols contiguous elements (<= 64 = one 128-byte swizzle span) x box_rows rows.
inline CUtensorMap make_tmap_bf16_sw128(const void* base, int64_t rows, int64_t cols, int64_t ld, int box_cols,
                                        int box_rows) {
  CUtensorMap m;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  CUresult r = encode_tiled_2d(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, const_cast<void*>(base), dims, strides, box,
                               CU_TENSOR_MAP_SWIZZLE_128B);
  M4T_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code " << static_cast<int>(r));
  return m;
}

}  // namespace m4t
