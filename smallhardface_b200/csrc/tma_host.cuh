// Host-side TMA tensor-map encoding shared by the tcgen05 conv kernels.
#pragma once
#include "common.cuh"

typedef CUresult (*ShfEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline ShfEncodeTiledFn shf_get_encode_fn() {
  static ShfEncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<ShfEncodeTiledFn>(p);
  }
  return fn;
}

// fp16 tensor, dims innermost-first, 128-byte swizzle, out-of-bounds reads return zero.  byte_strides (rank - 1 entries:
// the byte pitch of dims 1 .. rank-1) = nullptr means dense; a strided view (every s-th pixel of a map) passes its own.
inline int shf_encode_f16_map(CUtensorMap* map, void* base, int rank, const uint64_t* dims, const uint32_t* box,
                              const char* what, const uint64_t* byte_strides = nullptr) {
  ShfEncodeTiledFn fn = shf_get_encode_fn();
  SHF_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
  cuuint64_t gdim[5], gstride[4];
  cuuint32_t bdim[5], estride[5];
  uint64_t stride = 2;
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estride[i] = 1;
    stride *= dims[i];
    if (i < rank - 1) gstride[i] = byte_strides ? byte_strides[i] : stride;
  }
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, base, gdim, gstride, bdim, estride,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SHF_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
  return 0;
}
