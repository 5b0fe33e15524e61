// Dense fp32-accurate GEMM on the 5th-generation tensor cores (the large products of the path):
// the RelGraphConv self-loop  x @ loop_weight  (reference kgvae/model.py:55,58 via DGL) and its
// two backward products, and the MaskedLinear stacks of the IAF flow (kgvae/flow_network.py:15,
// 53-63; 90 products per forward at n_flows = 3) with theirs.
//
// Both operands are "prepared" once (two-term fp16 split under one power-of-two scale per matrix, see
// below) and multiplied by the shared TMA/tcgen05 pipeline, which reads a prepared matrix K-major or
// MN-major as the product needs - no transposing conversion exists.  A prepared operand can be handed in
// by the caller (kg_gemm_prepare / kg_gemm_f32_prepared) and reused: a layer splits x, W and its upstream
// gradient once for  x W,  g W^T  and  x^T g.  The epilogue undoes the two scales and applies bias /
// addend / ReLU / dropout mask / accumulate straight from tensor memory.
// Long-K, small-output products (the weight gradients: K = number of nodes) are split along K
// into a partials buffer and finished by a deterministic reduction kernel.
#include "split_pipe.cuh"

using namespace splitpipe;

namespace {

constexpr int EPI_WARPS = 8;                                    // two per 32-row quadrant: columns [0,128) and [128,256)
constexpr int GEMM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int EPI_PITCH = 20;                                   // padded row of a 32 x 16 transpose block (16-byte aligned)
constexpr int EPI_T = 32 * EPI_PITCH;
constexpr int SMEM_BYTES = PIPE_SMEM + 1024 + EPI_WARPS * EPI_T * 4;   // + transpose blocks

// ------------------------------------------------------------------------------------------
// prepared operands
// ------------------------------------------------------------------------------------------
// A prepared operand is the two-term fp16 split of a row-major fp32 matrix src[rows, cols] under ONE
// power-of-two scale for the whole matrix:  1024-byte header {amax bits, 1 / scale}, then rows of
// 2 * Cp halves  hi | lo  (Cp = cols rounded up to 64, zero padded).  Because the scale does not depend
// on the row, the same array serves a product that contracts over the COLUMNS of src (rows are read
// K-major, as before) and one that contracts over its ROWS (the tile is read MN-major: tcgen05 shared
// memory descriptors take either).  So x, W and the upstream gradient g of a layer are each split once
// and feed  x W,  g W^T  and  x^T g  without any transposing pass.
// Precision: an element keeps 22 significant bits while |x| >= 2^-17 max|src|; below that its absolute
// error is 2^-40 max|src| (fp16 subnormal spacing of the lo term) - far inside fp32 rounding of any sum
// the element takes part in.
struct PrepHeader {
  unsigned amax_bits;
  float inv_scale;
};
constexpr size_t PREP_HEADER_BYTES = 1024;

inline int prep_cp(int cols) { return kg_div_up(cols, BK) * BK; }
inline size_t prep_bytes(int rows, int cols) {
  return PREP_HEADER_BYTES + kg_align_up((size_t)rows * 2 * prep_cp(cols) * sizeof(__half), 1024);
}

// pass 1: bit pattern of max |src| (a non-negative float orders like an unsigned); warp per row, rows grid-strided
__global__ void __launch_bounds__(256)
amax_kernel(const float* __restrict__ src, int ld, int rows, int cols, PrepHeader* __restrict__ hdr) {
  __shared__ float red[8];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const bool vec4 = (ld & 3) == 0 && (cols & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  float m = 0.f;
  for (int r = blockIdx.x * 8 + wib; r < rows; r += gridDim.x * 8) {
    const float* x = src + (size_t)r * ld;
    if (vec4) {
      for (int k = 4 * lane; k < cols; k += 128) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(x + k));
        m = fmaxf(m, fmaxf(fmaxf(fabsf(t.x), fabsf(t.y)), fmaxf(fabsf(t.z), fabsf(t.w))));
      }
    } else {
      for (int k = lane; k < cols; k += 32) m = fmaxf(m, fabsf(__ldg(x + k)));
    }
  }
  m = warp_max(m);
  if (lane == 0) red[wib] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    atomicMax(&hdr->amax_bits, __float_as_uint(m));
  }
}

__device__ __forceinline__ uint2 pack_half4(float a, float b, float c, float d) {
  const __half2 p = __floats2half2_rn(a, b), q = __floats2half2_rn(c, d);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&p), *reinterpret_cast<const uint32_t*>(&q));
}

// pass 2: warp per row, four columns per lane: 128-bit loads, 64-bit stores of the hi and of the lo term
__global__ void __launch_bounds__(256)
split_kernel(const float* __restrict__ src, int ld, int rows, int cols, int Cp, PrepHeader* __restrict__ hdr,
             __half* __restrict__ cat) {
  const float s = split_scale(__uint_as_float(hdr->amax_bits));
  if (blockIdx.x == 0 && threadIdx.x == 0) hdr->inv_scale = 1.f / s;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const bool vec4 = (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  for (int r = blockIdx.x * 8 + wib; r < rows; r += gridDim.x * 8) {
    const float* x = src + (size_t)r * ld;
    __half* row = cat + (size_t)r * 2 * Cp;
    for (int k = 4 * lane; k < Cp; k += 128) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (vec4 && k + 3 < cols) {
        t = __ldg(reinterpret_cast<const float4*>(x + k));
      } else {
        if (k < cols) t.x = __ldg(x + k);
        if (k + 1 < cols) t.y = __ldg(x + k + 1);
        if (k + 2 < cols) t.z = __ldg(x + k + 2);
        if (k + 3 < cols) t.w = __ldg(x + k + 3);
      }
      t.x *= s; t.y *= s; t.z *= s; t.w *= s;
      const __half h0 = __float2half_rn(t.x), h1 = __float2half_rn(t.y), h2 = __float2half_rn(t.z), h3 = __float2half_rn(t.w);
      const __half2 p = __halves2half2(h0, h1), q = __halves2half2(h2, h3);
      *reinterpret_cast<uint2*>(row + k) = make_uint2(*reinterpret_cast<const uint32_t*>(&p), *reinterpret_cast<const uint32_t*>(&q));
      *reinterpret_cast<uint2*>(row + Cp + k) = pack_half4(t.x - __half2float(h0), t.y - __half2float(h1),
                                                           t.z - __half2float(h2), t.w - __half2float(h3));
    }
  }
}

int prepare_operand(const float* src, int ld, int rows, int cols, void* prep, cudaStream_t st) {
  PrepHeader* hdr = reinterpret_cast<PrepHeader*>(prep);
  __half* cat = reinterpret_cast<__half*>(reinterpret_cast<char*>(prep) + PREP_HEADER_BYTES);
  KG_CUDA(cudaMemsetAsync(hdr, 0, sizeof(PrepHeader), st));
  int blocks = kg_div_up(rows, 8);
  const int cap = 8 * kg_sm_count();
  if (blocks > cap) blocks = cap;
  amax_kernel<<<blocks, 256, 0, st>>>(src, ld, rows, cols, hdr);
  KG_LAUNCH_OK();
  split_kernel<<<blocks, 256, 0, st>>>(src, ld, rows, cols, prep_cp(cols), hdr, cat);
  KG_LAUNCH_OK();
  return KG_OK;
}

// ------------------------------------------------------------------------------------------
// main kernel
// ------------------------------------------------------------------------------------------
struct GemmArgs {
  float* C;
  int ldc, M, N;
  const float* inv_sa;   // 1 / scale of the prepared A (one float, device memory)
  const float* inv_sb;   // 1 / scale of the prepared B
  const float* bias;     // [N] or null
  const float* addend;   // [M, ldc] or null
  const float* mask;     // [M, ldc] or null
  int relu, accumulate;
  float* partial;        // split-K: [splits][M][N] raw products (scales applied), else null
  TileMap tmap;
  WideOperands op;
  int n_terms;
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
  float* epi_t = reinterpret_cast<float*>(P.scratch);     // [8][32 x 20] epilogue transpose blocks
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = g.tmap.total();

  if (warp == 0) {
    pipe_producer_wide(P, &tm_a, &tm_b, g.tmap, g.op, g.n_terms);
  } else if (warp == 1) {
    pipe_mma_wide(P, g.tmap, g.op, g.n_terms);
  } else {
    // TMEM lanes are reachable by warp id % 4: warps 2..5 take columns [0, 128) of their quadrant, 6..9 [128, 256)
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const float sc = __ldg(g.inv_sa) * __ldg(g.inv_sb);       // both powers of two: exact
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      int m0, n0, split, ks0, nks;
      g.tmap.decode(tile, m0, n0, split, ks0, nks);
      const int m = m0 + quad * 32 + lane;
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
              *reinterpret_cast<float4*>(T + lane * EPI_PITCH + j) = make_float4(
                  __uint_as_float(v[16 * hh + j]) * sc, __uint_as_float(v[16 * hh + j + 1]) * sc,
                  __uint_as_float(v[16 * hh + j + 2]) * sc, __uint_as_float(v[16 * hh + j + 3]) * sc);
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
            float r0 = __uint_as_float(v[j]) * sc, r1 = __uint_as_float(v[j + 1]) * sc;
            float r2 = __uint_as_float(v[j + 2]) * sc, r3 = __uint_as_float(v[j + 3]) * sc;
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
  size_t a_prep, b_prep, partial, total;
  int m_tiles, n_tiles, splits, k_steps, k_per_split;
};

// with_prep: the workspace also holds the prepared forms of A (stored a_rows x a_cols) and B
Layout layout(int M, int N, int K, bool with_prep, int trans_a, int trans_b) {
  Layout L;
  L.m_tiles = kg_div_up(M, BM);
  L.n_tiles = kg_div_up(N, BN);
  L.k_steps = kg_div_up(K, BK);
  const int tiles = L.m_tiles * L.n_tiles, sms = kg_sm_count();
  int splits = 1;
  if (tiles * 2 <= sms && L.k_steps >= 16) {
    splits = sms / tiles;
    if (splits > L.k_steps / 8) splits = L.k_steps / 8;
    if (splits < 1) splits = 1;
  }
  L.k_per_split = kg_div_up(L.k_steps, splits);
  L.splits = kg_div_up(L.k_steps, L.k_per_split);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += kg_align_up(bytes, 1024); return o; };
  L.a_prep = take(with_prep ? (trans_a ? prep_bytes(K, M) : prep_bytes(M, K)) : 0);
  L.b_prep = take(with_prep ? (trans_b ? prep_bytes(N, K) : prep_bytes(K, N)) : 0);
  L.partial = take(L.splits > 1 ? (size_t)L.splits * M * N * sizeof(float) : 0);
  L.total = off;
  return L;
}

int run_prepared(const void* prep_a, int trans_a, const void* prep_b, int trans_b, float* C, int ldc, int M, int N,
                 int K, const float* bias, const float* addend, int relu, const float* mask, int accumulate,
                 const Layout& L, float* partial, cudaStream_t st) {
  // stored shapes: A is [M, K] (read K-major) or, transposed, [K, M] (read MN-major); B is [N, K] when trans_b
  // (K-major), else [K, N] (MN-major)
  const int a_mn = trans_a ? 1 : 0, b_mn = trans_b ? 0 : 1;
  const int a_rows = trans_a ? K : M, a_cols = trans_a ? M : K;
  const int b_rows = trans_b ? N : K, b_cols = trans_b ? K : N;
  const int a_cp = prep_cp(a_cols), b_cp = prep_cp(b_cols);
  const char* pa = reinterpret_cast<const char*>(prep_a);
  const char* pb = reinterpret_cast<const char*>(prep_b);

  CUtensorMap tm_a, tm_b;
  int rc = make_tensor_map_2d_b16(&tm_a, pa + PREP_HEADER_BYTES, a_rows, 2 * a_cp, (uint64_t)2 * a_cp * sizeof(__half),
                                  a_mn ? 64 : BM);
  if (rc != KG_OK) return rc;
  rc = make_tensor_map_2d_b16(&tm_b, pb + PREP_HEADER_BYTES, b_rows, 2 * b_cp, (uint64_t)2 * b_cp * sizeof(__half),
                              b_mn ? 64 : BN);
  if (rc != KG_OK) return rc;

  if (kg_attr_needed(0))
    KG_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  GemmArgs g;
  g.C = C; g.ldc = ldc; g.M = M; g.N = N;
  g.inv_sa = &reinterpret_cast<const PrepHeader*>(pa)->inv_scale;
  g.inv_sb = &reinterpret_cast<const PrepHeader*>(pb)->inv_scale;
  g.bias = bias; g.addend = addend; g.mask = mask; g.relu = relu; g.accumulate = accumulate;
  g.partial = partial;
  g.tmap = TileMap{L.m_tiles, L.n_tiles, L.splits, L.k_steps, L.k_per_split};
  g.op = WideOperands{a_mn, b_mn, a_cp, b_cp};
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

char* aligned_ws(void* workspace, size_t workspace_bytes, size_t need) {
  uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023;
  if (!workspace || base + need > reinterpret_cast<uintptr_t>(workspace) + workspace_bytes) return nullptr;
  return reinterpret_cast<char*>(base);
}

}  // namespace

size_t kg_gemm_tc_workspace_bytes(int M, int N, int K, int trans_a, int trans_b) {
  if (M <= 0 || N <= 0 || K <= 0) return 1024;
  return layout(M, N, K, true, trans_a, trans_b).total + 1024;
}

bool kg_gemm_tc_eligible(int M, int N, int K) {
  return K >= 32 && M >= 64 && N >= 64 && (long long)M * N * K >= (1LL << 22);
}

int kg_gemm_tc_run(const float* A, int lda, int trans_a, const float* B, int ldb, int trans_b, float* C,
                   int ldc, int M, int N, int K, const float* bias, const float* addend, int relu,
                   const float* mask, int accumulate, void* workspace, size_t workspace_bytes,
                   cudaStream_t st) {
  const Layout L = layout(M, N, K, true, trans_a, trans_b);
  char* ws = aligned_ws(workspace, workspace_bytes, L.total);
  if (!ws) return kg_fail(KG_ERR_WORKSPACE, "gemm: workspace too small (%zu needed)", L.total + 1024);
  int rc = prepare_operand(A, lda, trans_a ? K : M, trans_a ? M : K, ws + L.a_prep, st);
  if (rc != KG_OK) return rc;
  rc = prepare_operand(B, ldb, trans_b ? N : K, trans_b ? K : N, ws + L.b_prep, st);
  if (rc != KG_OK) return rc;
  float* partial = L.splits > 1 ? reinterpret_cast<float*>(ws + L.partial) : nullptr;
  return run_prepared(ws + L.a_prep, trans_a, ws + L.b_prep, trans_b, C, ldc, M, N, K, bias, addend, relu, mask,
                      accumulate, L, partial, st);
}

// ------------------------------------------------------------------------------------------
// prepared-operand entry points
// ------------------------------------------------------------------------------------------
extern "C" size_t kg_gemm_prep_bytes(int rows, int cols) {
  if (rows <= 0 || cols <= 0) return PREP_HEADER_BYTES;
  return prep_bytes(rows, cols);
}

extern "C" int kg_gemm_prepare(const float* src, int ld, int rows, int cols, void* prep, size_t prep_size,
                               void* stream) {
  KG_REQUIRE(src && prep && rows > 0 && cols > 0 && ld >= cols, "gemm prepare: bad arguments");
  KG_REQUIRE((reinterpret_cast<uintptr_t>(prep) & 15) == 0, "gemm prepare: buffer must be 16-byte aligned");
  if (prep_size < prep_bytes(rows, cols))
    return kg_fail(KG_ERR_WORKSPACE, "gemm prepare: buffer too small (%zu needed)", prep_bytes(rows, cols));
  return prepare_operand(src, ld, rows, cols, prep, kg_stream(stream));
}

extern "C" size_t kg_gemm_f32_prepared_workspace_bytes(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 1024;
  return layout(M, N, K, false, 0, 0).total + 1024;
}

extern "C" int kg_gemm_f32_prepared(const void* prep_a, int trans_a, const void* prep_b, int trans_b, float* C,
                                    int ldc, int M, int N, int K, const float* bias, const float* addend, int relu,
                                    const float* mask, int accumulate, void* workspace, size_t workspace_bytes,
                                    void* stream) {
  KG_REQUIRE(prep_a && prep_b && C, "gemm (prepared): null operand");
  KG_REQUIRE(kg_gemm_tc_eligible(M, N, K), "gemm (prepared): the product is below the tensor-core size (use kg_gemm_f32)");
  KG_REQUIRE(((reinterpret_cast<uintptr_t>(prep_a) | reinterpret_cast<uintptr_t>(prep_b)) & 15) == 0,
             "gemm (prepared): operands must be 16-byte aligned");
  const Layout L = layout(M, N, K, false, 0, 0);
  float* partial = nullptr;
  if (L.splits > 1) {
    char* ws = aligned_ws(workspace, workspace_bytes, L.total);
    if (!ws) return kg_fail(KG_ERR_WORKSPACE, "gemm (prepared): workspace too small (%zu needed)", L.total + 1024);
    partial = reinterpret_cast<float*>(ws + L.partial);
  }
  return run_prepared(prep_a, trans_a, prep_b, trans_b, C, ldc, M, N, K, bias, addend, relu, mask, accumulate, L,
                      partial, kg_stream(stream));
}
