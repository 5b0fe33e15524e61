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

static int g_tc_terms = 3;
int tc_terms() { return g_tc_terms; }

}  // namespace tc05

// 3 (default): every tensor-core product of the library is the fp32-accurate three-term fp16 split.
// 1: single-product mode - operands rounded to 11 significant bits (per-row scaled fp16), one MMA per k-step
// instead of three; the "reduced-precision GEMM variant, reported separately" of north_star.  Affects
// kg_gemm_f32 (tensor-core path) and kg_distmult_rank; returns the previous setting.
extern "C" int kg_set_tc_terms(int terms) {
  const int old = tc05::g_tc_terms;
  if (terms == 1 || terms == 3) tc05::g_tc_terms = terms;
  return old;
}
