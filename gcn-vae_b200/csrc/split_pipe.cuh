// The fp32-accurate tensor-core product shared by the rank kernel (rank.cu) and the dense GEMM
// (gemm_tc.cu).
//
// Operands: every fp32 operand row x[0..K) is stored as two fp16 rows  hi | lo  side by side in
// one "cat" array of 2*Kp columns (Kp = K rounded up to 64, zero padded):
//     x * 2^s = hi + lo + O(2^-22 |x| 2^s),   s = per-row power of two with max|x| 2^s in [2^14, 2^15)
// and the tile accumulates, in fp32 in tensor memory,   lo*hi + hi*lo + hi*hi   (small terms
// first).  What is dropped (lo*lo, the split remainder) is ~2^-21 relative per product.
//
// Pipeline of one CTA (192 threads, persistent over a static round-robin tile list):
//   warp 0      TMA producer: {64 x 128} and {64 x 256} fp16 boxes, 128-byte swizzle, 4 stages
//   warp 1      one elected thread issues tcgen05.mma kind::f16 M=128 N=256 K=16, commits to mbarriers
//   warps 2..5  epilogue (caller-supplied): read the 128 x 256 fp32 accumulator from TMEM
//               (two accumulator buffers = all 512 TMEM columns, so MMA and epilogue overlap)
#pragma once
#include "tc05.cuh"

namespace splitpipe {

using namespace tc05;

constexpr int BM = 128, BN = 256, BK = 64;           // CTA tile; BK fp16 elements = one 128-byte row
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int THREADS = 192;
constexpr int PIPE_SMEM = STAGES * STAGE_BYTES + 128;  // stages + barriers; epilogue scratch follows
constexpr int EPI_THREADS = 128;

// tile -> (m0, n0, first k-step, number of k-steps); n fastest so that concurrently running CTAs
// share their A tile and stream B, which stays L2-resident
struct TileMap {
  int m_tiles, n_tiles, splits, k_steps, k_per_split;
  __host__ __device__ int total() const { return m_tiles * n_tiles * splits; }
  __device__ __forceinline__ void decode(int tile, int& m0, int& n0, int& split, int& ks0, int& nks) const {
    const int n_blk = tile % n_tiles;
    tile /= n_tiles;
    const int m_blk = tile % m_tiles;
    split = tile / m_tiles;
    m0 = m_blk * BM;
    n0 = n_blk * BN;
    ks0 = split * k_per_split;
    nks = min(k_per_split, k_steps - ks0);
  }
};

struct Pipe {
  uint8_t* smem;        // 1024-aligned; stage s at smem + s * STAGE_BYTES
  uint64_t *full, *empty, *tfull, *tempty;
  uint32_t tmem_base;
  uint8_t* scratch;     // epilogue scratch after the barriers
};

// all 192 threads; returns after TMEM is allocated and barriers are initialised
__device__ __forceinline__ Pipe pipe_setup(uint8_t* smem_raw, const CUtensorMap* tm_a, const CUtensorMap* tm_b,
                                           int epi_warps = 4) {
  Pipe P;
  P.smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(P.smem + STAGES * STAGE_BYTES);
  P.full = bars;                      // TMA -> MMA
  P.empty = bars + STAGES;            // MMA -> TMA
  P.tfull = bars + 2 * STAGES;        // MMA -> epilogue   [2]
  P.tempty = bars + 2 * STAGES + 2;   // epilogue -> MMA   [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  P.scratch = P.smem + PIPE_SMEM;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(tm_a);
    tma_prefetch_desc(tm_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(P.full + s, 1);
      mbar_init(P.empty + s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(P.tfull + i, 1);
      mbar_init(P.tempty + i, epi_warps);     // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  fence_before_thread_sync();
  __syncthreads();
  fence_after_thread_sync();
  P.tmem_base = *tmem_slot;
  return P;
}

// warp 0
// n_terms = 3: the fp32-accurate product lo*hi + hi*lo + hi*hi;  n_terms = 1: hi*hi only (single-product
// mode: operands rounded to 11 bits, a third of the MMAs - reported separately, never the default)
__device__ __forceinline__ void pipe_producer(const Pipe& P, const CUtensorMap* tm_a, const CUtensorMap* tm_b,
                                              const TileMap& tmap, int Kp, int n_terms = 3) {
  if (!elect_one()) return;
  int stage = 0;
  uint32_t phase = 0;
  const int total = tmap.total();
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    int m0, n0, split, ks0, nks;
    tmap.decode(tile, m0, n0, split, ks0, nks);
    for (int term = 3 - n_terms; term < 3; ++term) {
      // small terms first: lo*hi, hi*lo, then hi*hi   (hi at column 0, lo at column Kp)
      const int a_off = term == 0 ? Kp : 0, b_off = term == 1 ? Kp : 0;
      for (int ks = ks0; ks < ks0 + nks; ++ks) {
        mbar_wait(P.empty + stage, phase ^ 1);
        uint8_t* st = P.smem + stage * STAGE_BYTES;
        mbar_arrive_expect_tx(P.full + stage, STAGE_BYTES);
        tma_load_2d(st, tm_a, P.full + stage, a_off + ks * BK, m0);
        tma_load_2d(st + A_BYTES, tm_b, P.full + stage, b_off + ks * BK, n0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  }
}

// warp 1
__device__ __forceinline__ void pipe_mma(const Pipe& P, const TileMap& tmap, int n_terms = 3) {
  if (!elect_one()) return;
  constexpr uint32_t idesc = instr_desc_f16(0, BM, BN);
  int stage = 0;
  uint32_t phase = 0;
  int it = 0;
  const int total = tmap.total();
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
    int m0, n0, split, ks0, nks;
    tmap.decode(tile, m0, n0, split, ks0, nks);
    const int buf = it & 1;
    mbar_wait(P.tempty + buf, ((it >> 1) & 1) ^ 1);
    fence_after_thread_sync();
    const uint32_t tacc = P.tmem_base + buf * BN;
    for (int ks = 0; ks < n_terms * nks; ++ks) {
      mbar_wait(P.full + stage, phase);
      fence_after_thread_sync();
      const uint32_t sa = smem_u32(P.smem + stage * STAGE_BYTES);
      const uint64_t da = smem_desc_k_sw128(sa), db = smem_desc_k_sw128(sa + A_BYTES);
#pragma unroll
      for (int k = 0; k < BK / 16; ++k)   // 16 fp16 = 32 bytes per MMA: +2 in the (>>4) address field
        mma_f16_ss(tacc, da + 2 * k, db + 2 * k, idesc, (ks | k) != 0);
      mma_commit(P.empty + stage);        // frees the smem stage once these MMAs have read it
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    mma_commit(P.tfull + buf);            // accumulator complete
  }
}

// ------------------------------------------------------------------------------------------
// "Wide" variant of the producer / MMA pair (dense GEMM): one 96 KB stage per k-step holds
// A_hi | A_lo | B_hi | B_lo, so every operand tile crosses L2 -> shared memory ONCE per k-step and
// the three products  lo*hi + hi*lo + hi*hi  are issued from it (12 MMAs per stage).  The term-major
// pair above streams 6 tiles per k-step; at K = 500 that traffic, not the tensor pipe, bounded the
// GEMM.  Two stages fit the same shared memory as the four 48 KB ones (barriers full[0..1],
// empty[0..1] of the same Pipe).
//
// Either operand may be K-major (array rows = the operand's M/N index, contraction along the array's
// columns: one {64 x BM|BN} box per term) or MN-major (array rows = the contraction index: {64 x 64}
// boxes, one per 64 elements of M/N, k-rows beyond the array read as zero).  `cp` is the column at which
// the lo half of the operand's array starts.
// ------------------------------------------------------------------------------------------
constexpr int WIDE_STAGES = 2;
constexpr int WIDE_STAGE_BYTES = 2 * STAGE_BYTES;          // A_hi, A_lo, B_hi, B_lo
constexpr int MN_BOX_BYTES = 64 * BK * 2;                  // one {64 x 64} fp16 box
static_assert(WIDE_STAGES * WIDE_STAGE_BYTES <= STAGES * STAGE_BYTES, "wide stages must fit the pipe's shared memory");

struct WideOperands {
  int a_mn, b_mn;      // 1: MN-major
  int a_cp, b_cp;      // first column of the lo half
};

template <int TILE_MN>
__device__ __forceinline__ void wide_load(uint8_t* hi, uint8_t* lo, const CUtensorMap* tm, uint64_t* bar, int mn_major,
                                          int cp, int mn0, int ks, bool want_lo) {
  if (!mn_major) {
    tma_load_2d(hi, tm, bar, ks * BK, mn0);
    if (want_lo) tma_load_2d(lo, tm, bar, cp + ks * BK, mn0);
  } else {
#pragma unroll
    for (int j = 0; j < TILE_MN / 64; ++j) {
      const int c = mn0 + 64 * j;
      const bool inside = c < cp;                    // a box past the padded width reads zeros (fully out of bounds)
      tma_load_2d(hi + j * MN_BOX_BYTES, tm, bar, inside ? c : 2 * cp, ks * BK);
      if (want_lo) tma_load_2d(lo + j * MN_BOX_BYTES, tm, bar, inside ? cp + c : 2 * cp, ks * BK);
    }
  }
}

// warp 0
__device__ __forceinline__ void pipe_producer_wide(const Pipe& P, const CUtensorMap* tm_a, const CUtensorMap* tm_b,
                                                   const TileMap& tmap, const WideOperands& op, int n_terms = 3) {
  if (!elect_one()) return;
  int stage = 0;
  uint32_t phase = 0;
  const int total = tmap.total();
  const bool lo = n_terms == 3;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    int m0, n0, split, ks0, nks;
    tmap.decode(tile, m0, n0, split, ks0, nks);
    for (int ks = ks0; ks < ks0 + nks; ++ks) {
      mbar_wait(P.empty + stage, phase ^ 1);
      uint8_t* st = P.smem + stage * WIDE_STAGE_BYTES;
      mbar_arrive_expect_tx(P.full + stage, lo ? WIDE_STAGE_BYTES : STAGE_BYTES);
      wide_load<BM>(st, st + A_BYTES, tm_a, P.full + stage, op.a_mn, op.a_cp, m0, ks, lo);
      wide_load<BN>(st + 2 * A_BYTES, st + 2 * A_BYTES + B_BYTES, tm_b, P.full + stage, op.b_mn, op.b_cp, n0, ks, lo);
      if (++stage == WIDE_STAGES) { stage = 0; phase ^= 1; }
    }
  }
}

// warp 1
__device__ __forceinline__ void pipe_mma_wide(const Pipe& P, const TileMap& tmap, const WideOperands& op, int n_terms = 3) {
  if (!elect_one()) return;
  const uint32_t idesc = instr_desc_f16(0, BM, BN, op.a_mn, op.b_mn);
  // descriptor address field is in 16-byte units: a K = 16 step is 32 bytes along a K-major row, two 1024-byte
  // atoms of an MN-major tile
  const uint32_t a_step = op.a_mn ? 2048 >> 4 : 32 >> 4, b_step = op.b_mn ? 2048 >> 4 : 32 >> 4;
  int stage = 0;
  uint32_t phase = 0;
  int it = 0;
  const int total = tmap.total();
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
    int m0, n0, split, ks0, nks;
    tmap.decode(tile, m0, n0, split, ks0, nks);
    const int buf = it & 1;
    mbar_wait(P.tempty + buf, ((it >> 1) & 1) ^ 1);
    fence_after_thread_sync();
    const uint32_t tacc = P.tmem_base + buf * BN;
    for (int ks = 0; ks < nks; ++ks) {
      mbar_wait(P.full + stage, phase);
      fence_after_thread_sync();
      const uint32_t sa = smem_u32(P.smem + stage * WIDE_STAGE_BYTES);
      const uint32_t sb = sa + 2 * A_BYTES;
      const uint64_t a_hi = op.a_mn ? smem_desc_mn_sw128(sa, MN_BOX_BYTES) : smem_desc_k_sw128(sa);
      const uint64_t a_lo = op.a_mn ? smem_desc_mn_sw128(sa + A_BYTES, MN_BOX_BYTES) : smem_desc_k_sw128(sa + A_BYTES);
      const uint64_t b_hi = op.b_mn ? smem_desc_mn_sw128(sb, MN_BOX_BYTES) : smem_desc_k_sw128(sb);
      const uint64_t b_lo = op.b_mn ? smem_desc_mn_sw128(sb + B_BYTES, MN_BOX_BYTES) : smem_desc_k_sw128(sb + B_BYTES);
      if (n_terms == 3) {
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) mma_f16_ss(tacc, a_lo + k * a_step, b_hi + k * b_step, idesc, (ks | k) != 0);   // small terms first
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) mma_f16_ss(tacc, a_hi + k * a_step, b_lo + k * b_step, idesc, true);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) mma_f16_ss(tacc, a_hi + k * a_step, b_hi + k * b_step, idesc, true);
      } else {
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) mma_f16_ss(tacc, a_hi + k * a_step, b_hi + k * b_step, idesc, (ks | k) != 0);
      }
      mma_commit(P.empty + stage);        // frees the stage once these MMAs have read it
      if (++stage == WIDE_STAGES) { stage = 0; phase ^= 1; }
    }
    mma_commit(P.tfull + buf);            // accumulator complete
  }
}

// epilogue warps: wait for the accumulator of iteration `it`; returns this warp's TMEM address
__device__ __forceinline__ uint32_t epi_acquire(const Pipe& P, int it) {
  const int buf = it & 1, quad = (threadIdx.x >> 5) & 3;
  mbar_wait(P.tfull + buf, (it >> 1) & 1);
  fence_after_thread_sync();
  return P.tmem_base + ((uint32_t)(quad * 32) << 16) + buf * BN;
}

// epilogue warps: the accumulator buffer of iteration `it` may be overwritten
__device__ __forceinline__ void epi_release(const Pipe& P, int it) {
  fence_before_thread_sync();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(P.tempty + (it & 1));
}

__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// all threads, at the end
__device__ __forceinline__ void pipe_teardown(const Pipe& P) {
  fence_before_thread_sync();
  __syncthreads();
  if ((threadIdx.x >> 5) == 2) {
    fence_after_thread_sync();
    tmem_dealloc<512>(P.tmem_base);
  }
}

// ------------------------------------------------------------------------------------------
// operand split helpers (device)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// power-of-two scale that puts max|x| into [2^14, 2^15)
__device__ __forceinline__ float split_scale(float amax) {
  if (!(amax > 0.f) || !(amax < 3.0e38f)) return 1.f;
  int e;
  frexpf(amax, &e);
  e = 15 - e;
  e = e > 100 ? 100 : (e < -100 ? -100 : e);
  return ldexpf(1.f, e);
}

__device__ __forceinline__ void split_store(float x, float s, __half* hi_p, __half* lo_p) {
  const float xs = x * s;
  const __half hi = __float2half_rn(xs);
  *hi_p = hi;
  *lo_p = __float2half_rn(xs - __half2float(hi));
}

}  // namespace splitpipe
