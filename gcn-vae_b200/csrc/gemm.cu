// Dense fp32 GEMM with fused epilogue.
// Serves: the RelGraphConv self-loop  x @ loop_weight  (reference kgvae/model.py:55,58 via DGL
// matmul_maybe_select) fused with  + agg + h_bias -> activation -> dropout,  its backward
// (dW_loop = x^T g, dx += g W_loop^T), and MaskedLinear (kgvae/flow_network.py:15) fwd/bwd.
// Two kernels behind one entry point: products large enough to fill tensor-core tiles go to the
// tcgen05 split-fp16 kernel (gemm_tc.cu, fp32-accurate); small ones (and the K = 0 "epilogue
// only" form) use the fp32 FMA tile kernel below.  Both meet the fp32 parity bar (north_star:
// 1e-4 relative vs the reference).
#include "gemm_tile.cuh"

size_t kg_gemm_tc_workspace_bytes(int M, int N, int K, int trans_a, int trans_b);
bool kg_gemm_tc_eligible(int M, int N, int K);
int kg_gemm_tc_run(const float* A, int lda, int trans_a, const float* B, int ldb, int trans_b, float* C,
                   int ldc, int M, int N, int K, const float* bias, const float* addend, int relu,
                   const float* mask, int accumulate, void* workspace, size_t workspace_bytes,
                   cudaStream_t st);

using namespace kg_gemm;

struct Epilogue {
  float* C;
  int ldc, M, N;
  const float* bias;
  const float* addend;
  const float* mask;
  int relu, accumulate, atomic;
};

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(THREADS, 2)
gemm_kernel(TileLoader<A_KC> la, TileLoader<B_KC> lb, int K, int k_chunk, Epilogue ep) {
  __shared__ Smem sm;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * k_chunk;
  const int k_end = min(K, k_begin + k_chunk);
  float acc[8][8];
  mainloop<A_KC, B_KC>(la, lb, m0, n0, k_begin, k_end, sm, acc);

  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + tile_row(ty, i);
    if (m >= ep.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + tile_col(tx, j);
      if (n >= ep.N) continue;
      const size_t off = (size_t)m * ep.ldc + n;
      float v = acc[i][j];
      if (ep.bias) v += __ldg(ep.bias + n);
      if (ep.addend) v += __ldg(ep.addend + off);
      if (ep.relu) v = fmaxf(v, 0.f);
      if (ep.mask) v *= __ldg(ep.mask + off);
      if (ep.atomic) atomicAdd(ep.C + off, v);
      else if (ep.accumulate) ep.C[off] += v;
      else ep.C[off] = v;
    }
  }
}

// K = 0: the epilogue alone, one thread per element (narrow outputs would waste a 128 x 128 tile)
__global__ void __launch_bounds__(256)
epilogue_only_kernel(Epilogue ep) {
  const long long total = (long long)ep.M * ep.N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / ep.N), n = (int)(i - (long long)m * ep.N);
    const size_t off = (size_t)m * ep.ldc + n;
    float v = ep.bias ? __ldg(ep.bias + n) : 0.f;
    if (ep.addend) v += ep.addend == ep.C ? ep.C[off] : __ldg(ep.addend + off);   // in place: no read-only path
    if (ep.relu) v = fmaxf(v, 0.f);
    if (ep.mask) v *= __ldg(ep.mask + off);
    if (ep.accumulate) v += ep.C[off];
    ep.C[off] = v;
  }
}

// the same for contiguous rows (ldc == N), N % 4 == 0, 16-byte aligned pointers: 128-bit accesses
__global__ void __launch_bounds__(256)
epilogue_only_vec4_kernel(Epilogue ep) {
  const long long total4 = (long long)ep.M * ep.N / 4;
  const int n4 = ep.N >> 2;
  float4* C4 = reinterpret_cast<float4*>(ep.C);
  const bool inplace = ep.addend == ep.C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ep.bias) v = __ldg(reinterpret_cast<const float4*>(ep.bias) + (int)(i % n4));
    if (ep.addend) {
      const float4 a = inplace ? C4[i] : __ldg(reinterpret_cast<const float4*>(ep.addend) + i);
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
    if (ep.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    if (ep.mask) {
      const float4 k = __ldg(reinterpret_cast<const float4*>(ep.mask) + i);
      v.x *= k.x; v.y *= k.y; v.z *= k.z; v.w *= k.w;
    }
    if (ep.accumulate) { const float4 o = C4[i]; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
    C4[i] = v;
  }
}

// M <= 8 (the first pass of every MADE call multiplies ONE row by the masked weights, flow_network.py:91; its
// input-gradient product has the same shape): a 128 x 128 tile kernel would run one nearly empty CTA row through
// K sequential steps.  B k-contiguous: a warp per output column, lanes striding k (coalesced rows of B);
// otherwise a lane per output column (coalesced across n) with the CTA's eight warps splitting K.
constexpr int kSmallM = 8;

__device__ __forceinline__ void small_m_store(const Epilogue& ep, int m, int n, float v) {
  const size_t off = (size_t)m * ep.ldc + n;
  if (ep.bias) v += __ldg(ep.bias + n);
  if (ep.addend) v += __ldg(ep.addend + off);
  if (ep.relu) v = fmaxf(v, 0.f);
  if (ep.mask) v *= __ldg(ep.mask + off);
  if (ep.accumulate) v += ep.C[off];
  ep.C[off] = v;
}

template <bool B_KC>
__global__ void __launch_bounds__(256)
gemm_small_m_kernel(const float* __restrict__ A, int lda, int trans_a, const float* __restrict__ B, int ldb, int K,
                    Epilogue ep) {
  __shared__ float red[8][kSmallM][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, M = ep.M, N = ep.N;
  float acc[kSmallM];
#pragma unroll
  for (int m = 0; m < kSmallM; ++m) acc[m] = 0.f;
  auto a_at = [&](int m, int k) { return trans_a ? __ldg(A + (size_t)k * lda + m) : __ldg(A + (size_t)m * lda + k); };
  if (B_KC) {
    const int n = blockIdx.x * 8 + warp;
    if (n >= N) return;
    const float* b = B + (size_t)n * ldb;
    for (int k = lane; k < K; k += 32) {
      const float bv = __ldg(b + k);
#pragma unroll
      for (int m = 0; m < kSmallM; ++m)
        if (m < M) acc[m] = fmaf(a_at(m, k), bv, acc[m]);
    }
#pragma unroll
    for (int m = 0; m < kSmallM; ++m) {
      const float v = kg_warp_sum(acc[m]);
      if (m < M && lane == 0) small_m_store(ep, m, n, v);
    }
  } else {
    const int n = blockIdx.x * 32 + lane;
    const int kc = (K + 7) / 8, k0 = warp * kc, k1 = min(K, k0 + kc);
    if (n < N)
      for (int k = k0; k < k1; ++k) {
        const float bv = __ldg(B + (size_t)k * ldb + n);
#pragma unroll
        for (int m = 0; m < kSmallM; ++m)
          if (m < M) acc[m] = fmaf(a_at(m, k), bv, acc[m]);
      }
#pragma unroll
    for (int m = 0; m < kSmallM; ++m) red[warp][m][lane] = acc[m];
    __syncthreads();
    if (warp == 0 && n < N) {
#pragma unroll
      for (int m = 0; m < kSmallM; ++m)
        if (m < M) {
          float v = 0.f;
#pragma unroll
          for (int w8 = 0; w8 < 8; ++w8) v += red[w8][m][lane];
          small_m_store(ep, m, n, v);
        }
    }
  }
}

template <bool A_KC, bool B_KC>
static int launch(const float* A, int lda, const float* B, int ldb, int M, int N, int K, int splits,
                  int k_chunk, Epilogue ep, cudaStream_t st) {
  TileLoader<A_KC> la;
  la.ptr = A; la.ld = lda; la.rows = M; la.K = K;
  la.vec = ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && (lda % 4 == 0);
  TileLoader<B_KC> lb;
  lb.ptr = B; lb.ld = ldb; lb.rows = N; lb.K = K;
  lb.vec = ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && (ldb % 4 == 0);
  dim3 grid(kg_div_up(N, BN), kg_div_up(M, BM), splits);
  gemm_kernel<A_KC, B_KC><<<grid, THREADS, 0, st>>>(la, lb, K, k_chunk, ep);
  KG_LAUNCH_OK();
  return KG_OK;
}

extern "C" size_t kg_gemm_f32_workspace_bytes(int M, int N, int K) {
  // sized for the larger of the two storage orders of each operand (the padded widths differ)
  if (!kg_gemm_tc_eligible(M, N, K)) return 0;
  size_t best = 0;
  for (int t = 0; t < 4; ++t) {
    const size_t b = kg_gemm_tc_workspace_bytes(M, N, K, t & 1, t >> 1);
    if (b > best) best = b;
  }
  return best;
}

extern "C" int kg_gemm_f32_uses_tensor_cores(int M, int N, int K) { return kg_gemm_tc_eligible(M, N, K) ? 1 : 0; }

extern "C" int kg_gemm_f32(const float* A, int lda, int trans_a, const float* B, int ldb, int trans_b,
                           float* C, int ldc, int M, int N, int K, const float* bias,
                           const float* addend, int relu, const float* mask, int accumulate,
                           void* workspace, size_t workspace_bytes, void* stream) {
  KG_REQUIRE(M >= 0 && N >= 0 && K >= 0, "gemm: negative size");
  KG_REQUIRE(C && (K == 0 || (A && B)), "gemm: null operand");
  if (M == 0 || N == 0) return KG_OK;
  cudaStream_t st = kg_stream(stream);
  if (K == 0) {
    Epilogue ep{C, ldc, M, N, bias, addend, mask, relu, accumulate, 0};
    const long long cap = 16LL * kg_sm_count();
    const uintptr_t bits = reinterpret_cast<uintptr_t>(C) | reinterpret_cast<uintptr_t>(bias) |
                           reinterpret_cast<uintptr_t>(addend) | reinterpret_cast<uintptr_t>(mask);
    if (ldc == N && N % 4 == 0 && (bits & 15) == 0) {
      const long long blocks = ((long long)M * N / 4 + 255) / 256;
      epilogue_only_vec4_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(ep);
    } else {
      const long long blocks = ((long long)M * N + 255) / 256;
      epilogue_only_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(ep);
    }
    KG_LAUNCH_OK();
    return KG_OK;
  }
  if (kg_gemm_tc_eligible(M, N, K))
    return kg_gemm_tc_run(A, lda, trans_a, B, ldb, trans_b, C, ldc, M, N, K, bias, addend, relu, mask,
                          accumulate, workspace, workspace_bytes, st);

  if (M <= kSmallM && K >= 32) {
    Epilogue ep{C, ldc, M, N, bias, addend, mask, relu, accumulate, 0};
    if (trans_b) gemm_small_m_kernel<true><<<kg_div_up(N, 8), 256, 0, st>>>(A, lda, trans_a, B, ldb, K, ep);
    else gemm_small_m_kernel<false><<<kg_div_up(N, 32), 256, 0, st>>>(A, lda, trans_a, B, ldb, K, ep);
    KG_LAUNCH_OK();
    return KG_OK;
  }

  // split-K only for plain / masked products whose output tiling cannot fill the machine
  int splits = 1;
  const long long tiles = (long long)kg_div_up(M, BM) * kg_div_up(N, BN);
  const int sms = kg_sm_count();
  if (!bias && !addend && !relu && tiles < sms && K >= 8 * BK) {
    splits = (int)((2LL * sms + tiles - 1) / tiles);
    int max_splits = K / (4 * BK);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  int k_chunk = kg_div_up(kg_div_up(K, splits), BK) * BK;
  if (k_chunk < BK) k_chunk = BK;
  splits = K > 0 ? kg_div_up(K, k_chunk) : 1;

  Epilogue ep{C, ldc, M, N, bias, addend, mask, relu, accumulate, splits > 1 ? 1 : 0};
  if (splits > 1 && !accumulate) {
    KG_CUDA(cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * N, M, st));
  }
  // A is k-contiguous unless transposed; B is k-contiguous only when given as [N, K]
  if (!trans_a && trans_b) return launch<true, true>(A, lda, B, ldb, M, N, K, splits, k_chunk, ep, st);
  if (!trans_a && !trans_b) return launch<true, false>(A, lda, B, ldb, M, N, K, splits, k_chunk, ep, st);
  if (trans_a && trans_b) return launch<false, true>(A, lda, B, ldb, M, N, K, splits, k_chunk, ep, st);
  return launch<false, false>(A, lda, B, ldb, M, N, K, splits, k_chunk, ep, st);
}
