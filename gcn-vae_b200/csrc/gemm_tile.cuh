// fp32 SIMT GEMM main loop shared by kg_gemm_f32 and the rank kernel.
//
// CTA tile 128x128x16, 256 threads, 8x8 accumulators per thread laid out as two 4-wide strips
// (rows ty*4+{0..3} and 64+ty*4+{0..3}; same for columns) so that every shared-memory read is a
// conflict-free 128-bit load.  Global tiles are prefetched into registers while the previous
// tile is multiplied; shared memory is double-buffered, one __syncthreads per k-step.
//
// Accumulation order: every accumulator is a single fmaf chain over k = 0..K-1 in ascending
// order (zero-padded past K).  kg_distmult_rank relies on this to recompute a target's score
// with identical bits outside the tile kernel.
#pragma once
#include "common.cuh"

namespace kg_gemm {

constexpr int BM = 128, BN = 128, BK = 16, THREADS = 256, PAD = 4;

struct Smem {
  float a[2][BK][BM + PAD];
  float b[2][BK][BN + PAD];
};

// Operand view: element (r, k) where r is the tile's non-reduction index (m for A, n for B).
//   K_CONTIG = true : ptr[r * ld + k]    (A row-major [M,K]; B given as [N,K] i.e. trans_b)
//   K_CONTIG = false: ptr[k * ld + r]    (A given as [K,M] i.e. trans_a; B row-major [K,N])
template <bool K_CONTIG>
struct TileLoader {
  const float* ptr;
  int ld, rows, K;
  bool vec;  // 16-byte aligned base and ld % 4 == 0
  float4 reg[2];

  __device__ __forceinline__ float at(int r, int k) const {
    if (r < rows && k < K) return K_CONTIG ? __ldg(ptr + (size_t)r * ld + k) : __ldg(ptr + (size_t)k * ld + r);
    return 0.f;
  }

  // fetch this thread's 2 float4 of the (r0, k0) tile into registers
  __device__ __forceinline__ void fetch(int r0, int k0, int tid) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      int f = tid + q * THREADS;
      if (K_CONTIG) {
        int r = r0 + (f >> 2), k = k0 + (f & 3) * 4;
        if (vec && r < rows && k + 3 < K) {
          reg[q] = __ldg(reinterpret_cast<const float4*>(ptr + (size_t)r * ld + k));
        } else {
          reg[q] = make_float4(at(r, k), at(r, k + 1), at(r, k + 2), at(r, k + 3));
        }
      } else {
        int k = k0 + (f >> 5), r = r0 + (f & 31) * 4;
        if (vec && k < K && r + 3 < rows) {
          reg[q] = __ldg(reinterpret_cast<const float4*>(ptr + (size_t)k * ld + r));
        } else {
          reg[q] = make_float4(at(r, k), at(r + 1, k), at(r + 2, k), at(r + 3, k));
        }
      }
    }
  }

  // store the fetched registers into a [BK][128+PAD] shared tile
  __device__ __forceinline__ void stash(float (*tile)[BM + PAD], int tid) const {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      int f = tid + q * THREADS;
      if (K_CONTIG) {
        int r = f >> 2, k = (f & 3) * 4;
        tile[k + 0][r] = reg[q].x;
        tile[k + 1][r] = reg[q].y;
        tile[k + 2][r] = reg[q].z;
        tile[k + 3][r] = reg[q].w;
      } else {
        int k = f >> 5, r = (f & 31) * 4;
        *reinterpret_cast<float4*>(&tile[k][r]) = reg[q];
      }
    }
  }
};

// acc[i][j]: row m0 + (i<4 ? ty*4+i : 64+ty*4+i-4), col n0 + (j<4 ? tx*4+j : 64+tx*4+j-4)
template <bool A_KC, bool B_KC>
__device__ __forceinline__ void mainloop(TileLoader<A_KC>& la, TileLoader<B_KC>& lb, int m0, int n0,
                                         int k_begin, int k_end, Smem& sm, float (&acc)[8][8]) {
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  if (k_begin >= k_end) return;
  la.fetch(m0, k_begin, tid);
  lb.fetch(n0, k_begin, tid);
  la.stash(sm.a[0], tid);
  lb.stash(sm.b[0], tid);
  __syncthreads();

  int buf = 0;
  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    const bool more = k0 + BK < k_end;
    if (more) {
      la.fetch(m0, k0 + BK, tid);
      lb.fetch(n0, k0 + BK, tid);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&sm.a[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&sm.a[buf][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&sm.b[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&sm.b[buf][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) {
      la.stash(sm.a[buf ^ 1], tid);
      lb.stash(sm.b[buf ^ 1], tid);
      __syncthreads();
      buf ^= 1;
    }
  }
}

__device__ __forceinline__ int tile_row(int ty, int i) { return i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4); }
__device__ __forceinline__ int tile_col(int tx, int j) { return j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4); }

}  // namespace kg_gemm
