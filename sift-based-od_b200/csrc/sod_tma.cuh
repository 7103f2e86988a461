// Host-side tensor-map construction shared by the matching translation units.  The encode entry
// point comes from the driver through the runtime (no link against libcuda).
#pragma once
#include <cudaTypedefs.h>

#include "sod_common.cuh"

namespace sod {

inline PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) !=
            cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }();
  return fn;
}

// Row-major [n_rows, row_elems] matrix, boxes of box_rows x box_elems (box_elems * elem_bytes must
// be 128: one swizzle row), 128-byte swizzle, out-of-bounds rows read as zero.
inline int make_rowmajor_map(CUtensorMap* m, CUtensorMapDataType dtype, int elem_bytes, const void* ptr,
                             int64_t n_rows, int row_elems, int box_elems, int box_rows) {
  auto fn = tensor_map_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return SOD_ERR_CUDA;
  }
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(row_elems), static_cast<cuuint64_t>(n_rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(row_elems) * elem_bytes};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(box_elems), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(m, dtype, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
    return SOD_ERR_CUDA;
  }
  return SOD_OK;
}

}  // namespace sod
