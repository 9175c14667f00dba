// TMA tensor-map construction shared by the tcgen05 kernels (host side).
#include <cuda.h>

#include "car_common.cuh"

namespace car {
namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

}  // namespace

// 2-D bf16 tensor [rows][K], row pitch ld elements; box = [box_rows][box_k]; swizzle = box_k * 2 bytes
int make_tmap_bf16(CUtensorMap *tm, const uint16_t *base, int rows, int K, int ld, int box_rows, int box_k) {
  static EncodeTiledFn enc = nullptr;
  if (!enc) {
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      enc = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return -10; }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_k, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapSwizzle sw = box_k * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : (box_k * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<uint16_t *>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: %d (rows=%d K=%d ld=%d box=%dx%d)", (int)r, rows, K, ld, box_rows, box_k); return -11; }
  return 0;
}

// 2-D tensor of 2- or 4-byte elements [rows][cols], row pitch ld elements; box = [box_rows][box_cols];
// swizzle = box_cols * elt_bytes (32 / 64 / 128 bytes)
int make_tmap_2d(CUtensorMap *tm, const void *base, int elt_bytes, int rows, int cols, int ld, int box_rows, int box_cols) {
  static EncodeTiledFn enc = nullptr;
  if (!enc) {
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      enc = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return -10; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * elt_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const int inner = box_cols * elt_bytes;
  CUtensorMapSwizzle sw = inner == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (inner == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = enc(tm, elt_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base),
                   dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: %d (rows=%d cols=%d ld=%d box=%dx%d elt=%d)", (int)r, rows, cols, ld, box_rows, box_cols, elt_bytes); return -11; }
  return 0;
}

}  // namespace car
