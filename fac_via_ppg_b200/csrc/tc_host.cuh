// Host-side helpers shared by the tensor-core kernels (waveglow_tc.cu, waveglow_fused.cu): TMA tensor-map
// encoding through the driver entry point (the library links without libcuda), device properties per device.
#pragma once
#include "fac_common.cuh"
#include "tc_common.cuh"

namespace fac {
namespace tc {

constexpr int TC_BM = 128;       // time rows per tile and CTA (UMMA M per CTA)
constexpr int TC_NHALF = 256;    // N of one UMMA / columns of one accumulator region

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline CUtensorMapSwizzle tc_swizzle(int bk) {
  return bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      ptr = nullptr;
    return reinterpret_cast<EncodeTiledFn>(ptr);
  }();
  return fn;
}

// (B, T, C) channels-last 16-bit activation: box = bk channels x 128 rows of one utterance; rows outside [0, T)
// are zero-filled by the TMA unit (= Conv1d padding).
inline int make_act_map(CUtensorMap* m, const void* ptr, int B, int T, int C, int bk) {
  EncodeTiledFn fn = encode_fn();
  FAC_REQUIRE(fn != nullptr, "tensor-core path: cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)T * C * 2};
  cuuint32_t box[3] = {(cuuint32_t)bk, TC_BM, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, tc_swizzle(bk), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FAC_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation %dx%dx%d) failed: %d", B, T, C, (int)r);
  return 0;
}

// (N, K) row-major 16-bit weight: box = bk k x (min(256, N) / cta_group) rows.
inline int make_weight_map(CUtensorMap* m, const void* ptr, int N, int K, int cg, int bk) {
  EncodeTiledFn fn = encode_fn();
  FAC_REQUIRE(fn != nullptr, "tensor-core path: cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)((N < TC_NHALF ? N : TC_NHALF) / cg)};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, tc_swizzle(bk), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FAC_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weight %dx%d) failed: %d", N, K, (int)r);
  return 0;
}

inline int sm_count() {
  static int sms[FAC_MAX_DEVICES] = {};
  const int dev = current_device_slot();
  if (sms[dev] == 0) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
  return sms[dev];
}

}  // namespace tc
}  // namespace fac
