// Host side of tc05.cuh: tensor-map encoding through the driver entry point (no libcuda link).
#include "tc05.cuh"

namespace tc05 {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tensor_map_2d_b16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                           uint64_t row_stride_bytes, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return kg_fail(KG_ERR_CUDA, "tensor map: cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (row_stride_bytes & 15) || box_rows == 0 || box_rows > 256)
    return kg_fail(KG_ERR_INVALID, "tensor map: base/stride must be 16-byte aligned, box rows in 1..256");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {64, box_rows};            // 64 x 16-bit = one 128-byte swizzle row
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return kg_fail(KG_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return KG_OK;
}

}  // namespace tc05
