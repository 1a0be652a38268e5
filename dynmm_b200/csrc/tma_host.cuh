// Host-side TMA descriptor helpers shared by the tensor-core kernels.
#pragma once

#include <stdlib.h>

#include <mutex>

#include "common.cuh"

namespace dynmm {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

inline int encode_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box, bool swizzle128 = true) {
#ifdef DYNMM_PLAN_DRYRUN
  // host-only planner report (tools/plan_report.cu): no driver, no descriptors -- only the tiling decisions matter
  (void)m; (void)base; (void)rank; (void)dims; (void)strides_bytes; (void)box; (void)swizzle128;
  return DYNMM_OK;
#endif
  EncodeTiledFn fn = get_encode_fn();
  // The encoder is a DRIVER entry point: it needs the primary context bound to the calling thread.
  // Threads that have not made a runtime call yet (e.g. an autograd worker running our backward)
  // get it bound here.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaFree(nullptr);
    ctx_bound = true;
  }
  if (!fn) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return DYNMM_ECUDA;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  // L2 promotion of the TMA requests: 128 B by default; DYNMM_TMA_L2=256 / 64 / 0 selects another granularity
  static const CUtensorMapL2promotion promo = [] {
    const char* e = getenv("DYNMM_TMA_L2");
    if (!e) return CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    switch (atoi(e)) {
      case 256: return CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
      case 64: return CU_TENSOR_MAP_L2_PROMOTION_L2_64B;
      case 0: return CU_TENSOR_MAP_L2_PROMOTION_NONE;
      default: return CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    }
  }();
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with %d (rank %d dims %llu,%llu,%llu,%llu box %u,%u,%u,%u)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
              (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], box[1], box[2], rank > 3 ? box[3] : 0);
    return DYNMM_ECUDA;
  }
  return DYNMM_OK;
}

}  // namespace dynmm
