// Dense fp32-accurate GEMM on the 5th-generation tensor cores (the large products of the path):
// the RelGraphConv self-loop  x @ loop_weight  (reference kgvae/model.py:55,58 via DGL) and its
// two backward products, and the MaskedLinear stacks of the IAF flow (kgvae/flow_network.py:15,
// 53-63; 90 products per forward at n_flows = 3) with theirs.
//
// Both operands are brought to K-major two-term fp16 "cat" rows (split_pipe.cuh) by a small
// conversion pass - a transposing one when the operand is given MN-major - and multiplied by
// the shared TMA/tcgen05 pipeline; the epilogue undoes the per-row power-of-two scales and
// applies bias / addend / ReLU / dropout mask / accumulate straight from tensor memory.
// Long-K, small-output products (the weight gradients: K = number of nodes) are split along K
// into a partials buffer and finished by a deterministic reduction kernel.
#include "split_pipe.cuh"

using namespace splitpipe;

namespace {

constexpr int EPI_WARPS = 8;                                    // two per 32-row quadrant: columns [0,128) and [128,256)
constexpr int GEMM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int EPI_PITCH = 20;                                   // padded row of a 32 x 16 transpose block (16-byte aligned)
constexpr int EPI_T = 32 * EPI_PITCH;
constexpr int SMEM_BYTES = PIPE_SMEM + 1024 + 2 * 2 * BN * 4 + EPI_WARPS * EPI_T * 4;   // + [2][BN] inv_sb, [2][BN] spare, transpose blocks

// ------------------------------------------------------------------------------------------
// operand conversion
// ------------------------------------------------------------------------------------------
// K-contiguous source: element (r, k) = src[r * ld + k]; one warp per row
__global__ void __launch_bounds__(256)
split_rows_kernel(const float* __restrict__ src, int ld, int rows, int K, int Kp, __half* __restrict__ cat,
                  float* __restrict__ inv_scale) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* x = src + (size_t)r * ld;
  float amax = 0.f;
  for (int k = lane; k < K; k += 32) amax = fmaxf(amax, fabsf(__ldg(x + k)));
  amax = warp_max(amax);
  const float s = split_scale(amax);
  __half* row = cat + (size_t)r * 2 * Kp;
  const bool vec2 = (ld & 1) == 0 && (reinterpret_cast<uintptr_t>(src) & 7) == 0;
  for (int k = 2 * lane; k < Kp; k += 64) {          // two columns per lane: half2 stores, 128 B per warp and term
    float x0 = 0.f, x1 = 0.f;
    if (vec2 && k + 1 < K) {
      const float2 t = __ldg(reinterpret_cast<const float2*>(x + k));
      x0 = t.x; x1 = t.y;
    } else {
      if (k < K) x0 = __ldg(x + k);
      if (k + 1 < K) x1 = __ldg(x + k + 1);
    }
    x0 *= s; x1 *= s;
    const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
    *reinterpret_cast<__half2*>(row + k) = __halves2half2(h0, h1);
    *reinterpret_cast<__half2*>(row + Kp + k) =
        __halves2half2(__float2half_rn(x0 - __half2float(h0)), __float2half_rn(x1 - __half2float(h1)));
  }
  if (lane == 0) inv_scale[r] = 1.f / s;
}

// MN-contiguous source: element (r, k) = src[k * ld + r]
// pass 1: amax_bits[r] = max_k |src[k, r]| (bit pattern of a non-negative float orders like an unsigned)
__global__ void __launch_bounds__(256)
colmax_kernel(const float* __restrict__ src, int ld, int K, int rows, int k_chunk, unsigned* __restrict__ amax_bits) {
  __shared__ float red[8][33];
  const int r = blockIdx.x * 32 + threadIdx.x;
  const int k0 = blockIdx.y * k_chunk, k1 = min(K, k0 + k_chunk);
  float m = 0.f;
  if (r < rows)
    for (int k = k0 + threadIdx.y; k < k1; k += 8) m = fmaxf(m, fabsf(__ldg(src + (size_t)k * ld + r)));
  red[threadIdx.y][threadIdx.x] = m;
  __syncthreads();
  if (threadIdx.y == 0 && r < rows) {
#pragma unroll
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i][threadIdx.x]);
    atomicMax(amax_bits + r, __float_as_uint(m));
  }
}

// pass 2: 64(k) x 32(r) tiles through shared memory; writes hi|lo rows as half2 pairs (a warp stores
// 128 contiguous bytes per row and term), zero padding up to Kp (Kp is a multiple of 64)
__global__ void __launch_bounds__(256)
split_transpose_kernel(const float* __restrict__ src, int ld, int K, int rows, int Kp,
                       const unsigned* __restrict__ amax_bits, __half* __restrict__ cat,
                       float* __restrict__ inv_scale) {
  __shared__ float tile[64][33];
  const int r0 = blockIdx.y * 32, k0 = blockIdx.x * 64;     // K (up to millions of nodes) on grid.x
  for (int i = threadIdx.y; i < 64; i += 8) {
    const int k = k0 + i, r = r0 + threadIdx.x;
    tile[i][threadIdx.x] = (k < K && r < rows) ? __ldg(src + (size_t)k * ld + r) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i;
    if (r >= rows) continue;
    const float s = split_scale(__uint_as_float(__ldg(amax_bits + r)));
    const float x0 = tile[2 * threadIdx.x][i] * s, x1 = tile[2 * threadIdx.x + 1][i] * s;
    const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
    __half* row = cat + (size_t)r * 2 * Kp + k0 + 2 * threadIdx.x;
    *reinterpret_cast<__half2*>(row) = __halves2half2(h0, h1);
    *reinterpret_cast<__half2*>(row + Kp) =
        __halves2half2(__float2half_rn(x0 - __half2float(h0)), __float2half_rn(x1 - __half2float(h1)));
    if (blockIdx.x == 0 && threadIdx.x == 0) inv_scale[r] = 1.f / s;
  }
}

// ------------------------------------------------------------------------------------------
// main kernel
// ------------------------------------------------------------------------------------------
struct GemmArgs {
  float* C;
  int ldc, M, N;
  const float* inv_sa;   // [M]
  const float* inv_sb;   // [n_tiles * BN] (zero beyond N)
  const float* bias;     // [N] or null
  const float* addend;   // [M, ldc] or null
  const float* mask;     // [M, ldc] or null
  int relu, accumulate;
  float* partial;        // split-K: [splits][M][N] raw products (scales applied), else null
  TileMap tmap;
  int Kp, n_terms;
};

__device__ __forceinline__ float finish(float v, const GemmArgs& g, int n, size_t off) {
  if (g.bias) v += __ldg(g.bias + n);
  if (g.addend) v += __ldg(g.addend + off);
  if (g.relu) v = fmaxf(v, 0.f);
  if (g.mask) v *= __ldg(g.mask + off);
  if (g.accumulate) v += g.C[off];
  return v;
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  const Pipe P = pipe_setup(smem_raw, &tm_a, &tm_b, EPI_WARPS);
  float* sb_s = reinterpret_cast<float*>(P.scratch);      // [2][BN] column scales
  float* epi_t = sb_s + 2 * 2 * BN;                       // [8][32 x 20] epilogue transpose blocks
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = g.tmap.total();

  if (warp == 0) {
    pipe_producer_wide(P, &tm_a, &tm_b, g.tmap, g.Kp, g.n_terms);
  } else if (warp == 1) {
    pipe_mma_wide(P, g.tmap, g.n_terms);
  } else {
    // TMEM lanes are reachable by warp id % 4: warps 2..5 take columns [0, 128) of their quadrant, 6..9 [128, 256)
    const int quad = warp & 3, etid = threadIdx.x - 64, half = (warp - 2) >> 2;
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      int m0, n0, split, ks0, nks;
      g.tmap.decode(tile, m0, n0, split, ks0, nks);
      const int buf = it & 1;
      float* sb = sb_s + buf * BN;
      sb[etid] = __ldg(g.inv_sb + n0 + etid);
      const int m = m0 + quad * 32 + lane;
      const float sa = m < g.M ? __ldg(g.inv_sa + m) : 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");            // the 8 epilogue warps: column scales staged
      // while the accumulator is still being produced: pull this warp's lines of addend / mask / C towards L2
      if (!g.partial && m < g.M && (g.addend || g.mask || g.accumulate)) {
        const size_t off0 = (size_t)m * g.ldc + n0 + half * (BN / 2);
#pragma unroll
        for (int q = 0; q < BN / 2; q += 32) {
          if (n0 + half * (BN / 2) + q < g.N) {
            if (g.addend) asm volatile("prefetch.global.L2 [%0];" ::"l"(g.addend + off0 + q));
            if (g.mask) asm volatile("prefetch.global.L2 [%0];" ::"l"(g.mask + off0 + q));
            if (g.accumulate) asm volatile("prefetch.global.L2 [%0];" ::"l"(g.C + off0 + q));
          }
        }
      }
      const uint32_t taddr = epi_acquire(P, it);
      float* out = g.partial ? g.partial + ((size_t)split * g.M + m) * g.N : g.C + (size_t)m * g.ldc;
      const int ld_out = g.partial ? g.N : g.ldc;
      const bool vec = (ld_out & 3) == 0 && (g.N & 3) == 0 &&
                       (reinterpret_cast<uintptr_t>(g.partial ? g.partial : g.C) & 15) == 0 &&
                       (!g.addend || (reinterpret_cast<uintptr_t>(g.addend) & 15) == 0) &&
                       (!g.mask || (reinterpret_cast<uintptr_t>(g.mask) & 15) == 0) &&
                       (!g.bias || (reinterpret_cast<uintptr_t>(g.bias) & 15) == 0);
      float* T = epi_t + (warp - 2) * EPI_T;
#pragma unroll 1
      for (int c = half * (BN / 64); c < (half + 1) * (BN / 64); ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c * 32, v);
        tmem_ld_wait();
        const int nb = n0 + c * 32;
        if (vec && nb < g.N) {
          // The accumulator arrives one ROW per lane.  Written out that way a warp store is 32 strided
          // 16-byte pieces; instead the warp parks 32 x 16 blocks in shared memory and walks them with
          // 4 lanes per row (4 columns each): C / addend / mask are touched in whole 64-byte runs.
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 s4 = *reinterpret_cast<const float4*>(sb + c * 32 + 16 * hh + j);
              *reinterpret_cast<float4*>(T + lane * EPI_PITCH + j) = make_float4(
                  __uint_as_float(v[16 * hh + j]) * sa * s4.x, __uint_as_float(v[16 * hh + j + 1]) * sa * s4.y,
                  __uint_as_float(v[16 * hh + j + 2]) * sa * s4.z, __uint_as_float(v[16 * hh + j + 3]) * sa * s4.w);
            }
            __syncwarp();
            const int n = nb + 16 * hh + 4 * (lane & 3);
            if (n < g.N) {
              float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
              if (g.bias && !g.partial) b4 = __ldg(reinterpret_cast<const float4*>(g.bias + n));
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int rl = 8 * i + (lane >> 2), mm = m0 + quad * 32 + rl;
                if (mm < g.M) {
                  float4 r = *reinterpret_cast<const float4*>(T + rl * EPI_PITCH + 4 * (lane & 3));
                  if (g.partial) {
                    *reinterpret_cast<float4*>(g.partial + ((size_t)split * g.M + mm) * g.N + n) = r;
                  } else {
                    const size_t off = (size_t)mm * g.ldc + n;
                    r.x += b4.x; r.y += b4.y; r.z += b4.z; r.w += b4.w;
                    if (g.addend) { const float4 a = __ldg(reinterpret_cast<const float4*>(g.addend + off)); r.x += a.x; r.y += a.y; r.z += a.z; r.w += a.w; }
                    if (g.relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
                    if (g.mask) { const float4 k = __ldg(reinterpret_cast<const float4*>(g.mask + off)); r.x *= k.x; r.y *= k.y; r.z *= k.z; r.w *= k.w; }
                    if (g.accumulate) { const float4 o = *reinterpret_cast<const float4*>(g.C + off); r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w; }
                    *reinterpret_cast<float4*>(g.C + off) = r;
                  }
                }
              }
            }
            __syncwarp();                                   // the block is consumed before the next one is parked
          }
        } else if (m < g.M && nb < g.N) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const int n = nb + j;
            const float4 s4 = *reinterpret_cast<const float4*>(sb + c * 32 + j);
            float r0 = __uint_as_float(v[j]) * sa * s4.x, r1 = __uint_as_float(v[j + 1]) * sa * s4.y;
            float r2 = __uint_as_float(v[j + 2]) * sa * s4.z, r3 = __uint_as_float(v[j + 3]) * sa * s4.w;
            if (g.partial) {
              if (vec && n + 3 < g.N) {
                *reinterpret_cast<float4*>(out + n) = make_float4(r0, r1, r2, r3);
              } else {
                if (n < g.N) out[n] = r0;
                if (n + 1 < g.N) out[n + 1] = r1;
                if (n + 2 < g.N) out[n + 2] = r2;
                if (n + 3 < g.N) out[n + 3] = r3;
              }
            } else {
              const size_t off = (size_t)m * g.ldc + n;
              if (vec && n + 3 < g.N) {
                if (g.bias) { r0 += __ldg(g.bias + n); r1 += __ldg(g.bias + n + 1); r2 += __ldg(g.bias + n + 2); r3 += __ldg(g.bias + n + 3); }
                if (g.addend) { const float4 a = __ldg(reinterpret_cast<const float4*>(g.addend + off)); r0 += a.x; r1 += a.y; r2 += a.z; r3 += a.w; }
                if (g.relu) { r0 = fmaxf(r0, 0.f); r1 = fmaxf(r1, 0.f); r2 = fmaxf(r2, 0.f); r3 = fmaxf(r3, 0.f); }
                if (g.mask) { const float4 k = __ldg(reinterpret_cast<const float4*>(g.mask + off)); r0 *= k.x; r1 *= k.y; r2 *= k.z; r3 *= k.w; }
                if (g.accumulate) { const float4 o = *reinterpret_cast<const float4*>(g.C + off); r0 += o.x; r1 += o.y; r2 += o.z; r3 += o.w; }
                *reinterpret_cast<float4*>(g.C + off) = make_float4(r0, r1, r2, r3);
              } else {
                if (n < g.N) g.C[off] = finish(r0, g, n, off);
                if (n + 1 < g.N) g.C[off + 1] = finish(r1, g, n + 1, off + 1);
                if (n + 2 < g.N) g.C[off + 2] = finish(r2, g, n + 2, off + 2);
                if (n + 3 < g.N) g.C[off + 3] = finish(r3, g, n + 3, off + 3);
              }
            }
          }
        }
      }
      epi_release(P, it);
    }
  }
  pipe_teardown(P);
}

// split-K: C = epilogue(sum_s partial[s]) in a fixed order
__global__ void __launch_bounds__(256)
splitk_finish_kernel(GemmArgs g, int splits) {
  const long long total = (long long)g.M * g.N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / g.N), n = (int)(i % g.N);
    float v = 0.f;
    for (int s = 0; s < splits; ++s) v += g.partial[(size_t)s * total + i];
    const size_t off = (size_t)m * g.ldc + n;
    g.C[off] = finish(v, g, n, off);
  }
}

struct Layout {
  size_t acat, bcat, inv_sa, inv_sb, amax, partial, total;
  int Kp, m_tiles, n_tiles, splits, k_per_split;
};

Layout layout(int M, int N, int K) {
  Layout L;
  L.Kp = kg_div_up(K, BK) * BK;
  L.m_tiles = kg_div_up(M, BM);
  L.n_tiles = kg_div_up(N, BN);
  const int k_steps = L.Kp / BK, tiles = L.m_tiles * L.n_tiles, sms = kg_sm_count();
  int splits = 1;
  if (tiles * 2 <= sms && k_steps >= 16) {
    splits = sms / tiles;
    if (splits > k_steps / 8) splits = k_steps / 8;
    if (splits < 1) splits = 1;
  }
  L.k_per_split = kg_div_up(k_steps, splits);
  L.splits = kg_div_up(k_steps, L.k_per_split);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += kg_align_up(bytes, 1024); return o; };
  L.acat = take((size_t)M * 2 * L.Kp * sizeof(__half));
  L.bcat = take((size_t)N * 2 * L.Kp * sizeof(__half));
  L.inv_sa = take((size_t)M * sizeof(float));
  L.inv_sb = take((size_t)L.n_tiles * BN * sizeof(float));
  L.amax = take((size_t)(M > N ? M : N) * sizeof(unsigned));
  L.partial = take(L.splits > 1 ? (size_t)L.splits * M * N * sizeof(float) : 0);
  L.total = off;
  return L;
}

int convert_operand(const float* src, int ld, bool k_contig, int rows, int K, int Kp, __half* cat,
                    float* inv_scale, unsigned* amax, cudaStream_t st) {
  if (k_contig) {
    split_rows_kernel<<<kg_div_up((long long)rows * 32, 256), 256, 0, st>>>(src, ld, rows, K, Kp, cat, inv_scale);
    KG_LAUNCH_OK();
  } else {
    KG_CUDA(cudaMemsetAsync(amax, 0, sizeof(unsigned) * rows, st));
    const int k_chunk = 512;
    colmax_kernel<<<dim3(kg_div_up(rows, 32), kg_div_up(K, k_chunk)), dim3(32, 8), 0, st>>>(src, ld, K, rows, k_chunk, amax);
    KG_LAUNCH_OK();
    split_transpose_kernel<<<dim3(Kp / 64, kg_div_up(rows, 32)), dim3(32, 8), 0, st>>>(src, ld, K, rows, Kp, amax, cat, inv_scale);
    KG_LAUNCH_OK();
  }
  return KG_OK;
}

}  // namespace

size_t kg_gemm_tc_workspace_bytes(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 1024;
  return layout(M, N, K).total + 1024;
}

bool kg_gemm_tc_eligible(int M, int N, int K) {
  return K >= 32 && M >= 64 && N >= 64 && (long long)M * N * K >= (1LL << 22);
}

int kg_gemm_tc_run(const float* A, int lda, int trans_a, const float* B, int ldb, int trans_b, float* C,
                   int ldc, int M, int N, int K, const float* bias, const float* addend, int relu,
                   const float* mask, int accumulate, void* workspace, size_t workspace_bytes,
                   cudaStream_t st) {
  const Layout L = layout(M, N, K);
  uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023;
  if (!workspace || base + L.total > reinterpret_cast<uintptr_t>(workspace) + workspace_bytes)
    return kg_fail(KG_ERR_WORKSPACE, "gemm: workspace too small (%zu needed)", L.total + 1024);
  char* ws = reinterpret_cast<char*>(base);
  __half* acat = reinterpret_cast<__half*>(ws + L.acat);
  __half* bcat = reinterpret_cast<__half*>(ws + L.bcat);
  float* inv_sa = reinterpret_cast<float*>(ws + L.inv_sa);
  float* inv_sb = reinterpret_cast<float*>(ws + L.inv_sb);
  unsigned* amax = reinterpret_cast<unsigned*>(ws + L.amax);
  float* partial = L.splits > 1 ? reinterpret_cast<float*>(ws + L.partial) : nullptr;

  KG_CUDA(cudaMemsetAsync(inv_sb, 0, sizeof(float) * L.n_tiles * BN, st));
  int rc = convert_operand(A, lda, !trans_a, M, K, L.Kp, acat, inv_sa, amax, st);
  if (rc != KG_OK) return rc;
  rc = convert_operand(B, ldb, trans_b != 0, N, K, L.Kp, bcat, inv_sb, amax, st);
  if (rc != KG_OK) return rc;

  CUtensorMap tm_a, tm_b;
  const uint64_t row_bytes = (uint64_t)2 * L.Kp * sizeof(__half);
  rc = make_tensor_map_2d_b16(&tm_a, acat, M, 2 * L.Kp, row_bytes, BM);
  if (rc != KG_OK) return rc;
  rc = make_tensor_map_2d_b16(&tm_b, bcat, N, 2 * L.Kp, row_bytes, BN);
  if (rc != KG_OK) return rc;

  if (kg_attr_needed(0))
    KG_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  GemmArgs g;
  g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.inv_sa = inv_sa; g.inv_sb = inv_sb;
  g.bias = bias; g.addend = addend; g.mask = mask; g.relu = relu; g.accumulate = accumulate;
  g.partial = partial;
  g.tmap = TileMap{L.m_tiles, L.n_tiles, L.splits, L.Kp / BK, L.k_per_split};
  g.Kp = L.Kp;
  g.n_terms = tc05::tc_terms();
  const int total = g.tmap.total();
  const int grid = total < kg_sm_count() ? total : kg_sm_count();
  gemm_tc_kernel<<<grid, GEMM_THREADS, SMEM_BYTES, st>>>(tm_a, tm_b, g);
  KG_LAUNCH_OK();
  if (partial) {
    const long long n = (long long)M * N;
    int blocks = kg_div_up(n, 256);
    if (blocks > 8 * kg_sm_count()) blocks = 8 * kg_sm_count();
    splitk_finish_kernel<<<blocks, 256, 0, st>>>(g, L.splits);
    KG_LAUNCH_OK();
  }
  return KG_OK;
}
