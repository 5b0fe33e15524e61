// a3: RelGraphConv(regularizer="bdd") message passing, BLOCK-OWNER design (the reference model's
// 5x5 and 5x10 blocks; DGL RelGraphConv as constructed at kgvae/model.py:54-59): the shared pieces -
// bulk-async copy / reduction wrappers, the row source (one matrix or one block per rank over NVLink),
// register <-> shared memory helpers, eligibility.  The kernels are in rgcn_bdd_warp.cuh (the first,
// CTA-synchronised generation that lived here was superseded by them and has been removed).
//
// A thread owns TB whole diagonal blocks of the relation's weight - TB*si*so = 100 floats, read
// straight from the DGL-layout weight row [B][si][so] (they are contiguous) - and keeps them in
// REGISTERS while the relation lasts.  Per edge it reads its TB*si inputs with 64/128-bit shared
// memory loads, does TB*si*so FMAs and emits TB*so outputs with vector reductions: no per-edge
// weight traffic, no selects, 2/3 of the issued instructions are FMAs.  With B = 100 an edge needs
// 25 (5x5, TB = 4) or 50 (5x10, TB = 2) lanes: ONE or TWO warps per edge, so a 128-thread CTA runs
// 4 (2) independent slots with __syncwarp / a 64-thread named barrier as the only synchronisation.
//
// Data movement is bulk-asynchronous (TMA engine, no per-lane address arithmetic):
//   gather   one cp.async.bulk global -> shared per row (2 KB) into a 4-deep ring, completion on an
//            mbarrier per stage; L2 evict-first hint when the gathered matrix streams from HBM
//   scatter  the warp parks its scaled outputs in shared memory and ONE lane issues
//            cp.reduce.async.bulk.add.f32 shared -> global for the warp's contiguous 1-2 KB of the
//            destination row: the atomic adds happen in L2 on whole lines, the SM issues no REDs.
//
// Backward is fused: the warps of a slot split into an input-gradient role (same shape as the
// forward, transposed contraction) and a weight-gradient role (TB*si*so outer-product accumulators
// in the same registers, flushed with vector reductions when the relation changes); both read the
// gathered dagg[dst] row, the weight-gradient role also x[src].
#pragma once
#include "common.cuh"

namespace bddown {

constexpr int kCta = 128;     // threads per CTA (4 warps)
constexpr int kChunk = 256;   // consecutive relation-sorted edges per CTA

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}
__device__ __forceinline__ void red_add_v2(float* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy(bool evict_first) {
  uint64_t pol;
  if (evict_first) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_last(bool evict_last) {
  uint64_t pol;
  if (evict_last) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
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
// bulk copy global -> shared (bytes % 16 == 0, both 16-byte aligned), completes on `bar`
__device__ __forceinline__ void bulk_g2s(float* dst, const float* __restrict__ src, uint32_t bytes, uint64_t* bar,
                                         uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}
// bulk reduction shared -> global: dst[0..bytes/4) += src[0..bytes/4) (fp32), part of the thread's bulk group
__device__ __forceinline__ void bulk_red_add(float* dst, const float* src, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.L2::cache_hint.add.f32 [%0], [%1], %2, %3;"
               ::"l"(dst), "r"(smem_u32(src)), "r"(bytes), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// order this thread's generic-proxy shared-memory writes before later async-proxy (bulk) reads
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ const float* row_at(const float* base, int row, int width) {
  return base + (size_t)(unsigned)row * (unsigned)width;
}

// Where gathered rows live: one matrix, or (destination-partitioned training) one row block per
// rank, each in the owner's HBM and mapped into this process (CUDA IPC): row r is row r % part_rows
// of parts[r / part_rows].  A peer row is fetched by the same bulk copy, over NVLink - the
// "all-gather" of layer inputs is fused into the gather stage of the message-passing kernel and
// moves only the rows this rank's edges reference.
struct RowSource {
  const float* base;
  const float* const* parts;
  int part_rows;
  __device__ __forceinline__ const float* row(int r, int width) const {
    if (parts == nullptr) return row_at(base, r, width);
    const int owner = r / part_rows;
    return row_at(parts[owner], r - owner * part_rows, width);
  }
};

// N contiguous floats from shared memory with the widest loads the alignment of N allows
template <int N>
__device__ __forceinline__ void lds_vec(float (&v)[N], const float* p) {
  if (N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N; i += 4) {
      const float4 t = *reinterpret_cast<const float4*>(p + i);
      v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; i += 2) {
      const float2 t = *reinterpret_cast<const float2*>(p + i);
      v[i] = t.x; v[i + 1] = t.y;
    }
  }
}

// owner lanes per warp when `per` groups of N columns are spread over `warps` warps; even when N * 4
// is not a multiple of 16 so that every warp's share of a row is a whole number of 16-byte pieces
__device__ __host__ __forceinline__ int lanes_per_warp(int per, int warps, int n_cols) {
  int lpw = (per + warps - 1) / warps;
  if ((n_cols * 4) % 16 != 0 && (lpw & 1)) ++lpw;
  return lpw;
}

// shapes the block-owner kernels cover: 5x5 blocks (4 per thread) and 5x10 blocks (2 per thread)
// with at most 32 / 64 owner lanes per edge
inline bool eligible(int B, int si, int so) {
  if (si == 5 && so == 5) return B % 4 == 0 && B / 4 <= 32;
  if (si == 5 && so == 10) return B % 2 == 0 && B / 2 <= 64;
  return false;
}

}  // namespace bddown
