// Blackwell (sm_100a) building blocks used by the tensor-core kernels of this library:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 MMA / TMEM wrappers as inline PTX, the shared
// memory and instruction descriptors of tcgen05.mma, and the host-side tensor-map encoder.
// No CUTLASS/CuTe dependency: everything is spelled out here.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace tc05 {

// ------------------------------------------------------------------------------------------
// shared-memory addresses, mbarrier
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// make barrier initialisation visible to the async proxy (TMA, tcgen05.commit)
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}

// blocks until the phase with the given parity has completed
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 rx;\n\t"
      ".reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, px;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------------------------------
// TMA: 2-D tiled load global -> shared, completion on an mbarrier
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// ------------------------------------------------------------------------------------------
// TMEM allocation (one full warp), fences
// ------------------------------------------------------------------------------------------
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}

template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

__device__ __forceinline__ void fence_before_thread_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_after_thread_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ------------------------------------------------------------------------------------------
// tcgen05.mma (cta_group::1), operands in shared memory, accumulator in TMEM
// ------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor for a K-major operand tile stored as rows of 128 bytes with
// the 128-byte swizzle (what a TMA box {64 x 16-bit, rows} with CU_TENSOR_MAP_SWIZZLE_128B
// writes): 8-row groups are 1024 bytes apart (stride byte offset), the leading byte offset is
// unused, descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t smem_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address      bits [0,14)
  d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset bits [32,46)
  d |= (uint64_t)1 << 46;                             // version            bits [46,48)
  d |= (uint64_t)2 << 61;                             // SWIZZLE_128B       bits [61,64)
  return d;
}

// The same for an MN-major operand tile: what TMA boxes {64 x 16-bit along M/N, 64 rows along K} write, one
// 8 KB box per 64 elements of M/N.  A swizzle atom is 8 k-rows x 128 bytes (64 M/N elements); the stride byte
// offset is the distance between atoms along K (1024 bytes: the k-rows of a box are consecutive), the leading
// byte offset the distance between atoms along M/N (the box pitch).  One K = 16 MMA step covers two atoms along K:
// the start address advances by 2048 bytes per step.
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t smem_addr, uint32_t box_pitch_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address       bits [0,14)
  d |= (uint64_t)(box_pitch_bytes >> 4) << 16;        // leading byte offset bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset  bits [32,46)
  d |= (uint64_t)1 << 46;                             // version             bits [46,48)
  d |= (uint64_t)2 << 61;                             // SWIZZLE_128B        bits [61,64)
  return d;
}

// Instruction descriptor, kind::f16: A/B format 0 = F16, 1 = BF16; fp32 accumulate; a_mn / b_mn = 1 when that
// operand's tile is MN-major in shared memory (0 = K-major).
__host__ __device__ constexpr uint32_t instr_desc_f16(int ab_format, int M, int N, int a_mn = 0, int b_mn = 0) {
  return (1u << 4)                       // C format F32
         | ((uint32_t)ab_format << 7)    // A format
         | ((uint32_t)ab_format << 10)   // B format
         | ((uint32_t)a_mn << 15)        // A major
         | ((uint32_t)b_mn << 16)        // B major
         | ((uint32_t)(N >> 3) << 17)    // N
         | ((uint32_t)(M >> 4) << 24);   // M
}

__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// all previously issued MMAs of this thread arrive (once) on the barrier when they complete
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// TMEM -> registers: this thread's lane (row), 32 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// host: tensor map for a row-major 2-D array of 16-bit elements, box = {64 columns, box_rows},
// 128-byte swizzle, out-of-bounds elements read as zero
// ------------------------------------------------------------------------------------------
int make_tensor_map_2d_b16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                           uint64_t row_stride_bytes, uint32_t box_rows);
// number of split terms the tensor-core products use: 3 (fp32-accurate, default) or 1 (kg_set_tc_terms)
int tc_terms();

}  // namespace tc05
