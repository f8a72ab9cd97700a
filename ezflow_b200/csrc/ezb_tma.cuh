// Host-side TMA tensor-map construction through the driver entry point (no libcuda link dependency).
#pragma once

#include "ezb_common.cuh"

namespace ezb {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

inline int make_tmap_ex(CUtensorMap* tm, CUtensorMapDataType dt, int rank, void* ptr, const cuuint64_t* dims,
                     const cuuint64_t* strides_bytes /*rank-1*/, const cuuint32_t* box, const cuuint32_t* elem_strides /*rank, or null*/,
                     CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(EZB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if (elem_strides)
    for (int i = 0; i < rank; ++i) estr[i] = elem_strides[i];
  CUresult r = fn(tm, dt, (cuuint32_t)rank, ptr, dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(EZB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return EZB_OK;
}

inline int make_tmap(CUtensorMap* tm, CUtensorMapDataType dt, int rank, void* ptr, const cuuint64_t* dims,
                     const cuuint64_t* strides_bytes /*rank-1*/, const cuuint32_t* box, CUtensorMapSwizzle swz) {
  return make_tmap_ex(tm, dt, rank, ptr, dims, strides_bytes, box, nullptr, swz);
}

}  // namespace ezb
