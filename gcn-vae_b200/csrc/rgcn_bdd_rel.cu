// a3: RelGraphConv(regularizer="bdd") message passing in RELATION-MAJOR order.
//
// Per edge the layer needs the whole block-diagonal weight of the edge's relation: B*si*so
// floats (10 KB for 5x5 blocks, 20 KB for 5x10) against a 2 KB source row.  Walking edges in
// destination order (rgcn_bdd.cu, first version) re-reads those weights from L2 for every
// edge - 5-10x the feature bytes.  Here edges are walked in (relation, destination) order -
// the etype-major record list kg_graph_index already builds - so a CTA keeps W_r in shared
// memory for a whole run of edges and the per-edge traffic is the source row in and the message
// out:
//
//   forward   out[dst] += norm * blockdiag(W_r) x[src]          message accumulated with
//   dX        dx[src]  += norm * blockdiag(W_r)^T dagg[dst]     128-bit vector reductions (RED.v4)
//   dW        dW_r     += norm * x[src] (x) dagg[dst] blockwise registers, one flush per run
//
// dX and dW share one kernel (both need the gathered dagg row).  The reductions into out/dx hit
// rows that are L2-resident at knowledge-graph sizes; summation order across CTAs is not fixed,
// so results are reproducible to fp32 rounding, not bitwise.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kChunk = 128;   // consecutive relation-sorted edges per CTA
constexpr int kGroup = 4;     // edges staged in shared memory at a time

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// cooperative asynchronous copy of `n` floats (n % 4 == 0, 16-byte aligned rows) global -> shared
__device__ __forceinline__ void stage_row_async(float* dst, const float* __restrict__ src, int n) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
  for (int i = threadIdx.x; i < n / 4; i += kThreads)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 16 * i), "l"(src + 4 * i) : "memory");
}
__device__ __forceinline__ void async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// next run of <= kGroup consecutive edges of one relation starting at e (uniform across the CTA)
__device__ __forceinline__ int group_len(const int4* __restrict__ pack, int e, int e1) {
  if (e >= e1) return 0;
  const int r = __ldg(&pack[e].z);
  int g = 1;
  while (g < kGroup && e + g < e1 && __ldg(&pack[e + g].z) == r) ++g;
  return g;
}

// message of one staged neighbour row: 4 consecutive output columns j0..j0+3
//   msg[j] = sum_{i<FI} xs[(j / FO) * FI + i] * W_s[i * width + j]
template <int FI, int FO>
__device__ __forceinline__ float4 block_message(const float* xs, const float* W_s, int width, int j0) {
  const int b0 = j0 / FO, b3 = (j0 + 3) / FO;          // the 4 columns touch at most 2 blocks
  const bool s1 = (j0 + 1) / FO == b0, s2 = (j0 + 2) / FO == b0;
  float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < FI; ++i) {
    const float xa = xs[b0 * FI + i], xb = xs[b3 * FI + i];
    const float4 w = *reinterpret_cast<const float4*>(W_s + i * width + j0);
    m.x = fmaf(xa, w.x, m.x);
    m.y = fmaf(s1 ? xa : xb, w.y, m.y);
    m.z = fmaf(s2 ? xa : xb, w.z, m.z);
    m.w = fmaf(xb, w.w, m.w);
  }
  return m;
}

// out[tgt] += norm * blockdiag(W_etype) feat[nbr]; W layout [R][FI][B*FO]; out zero-filled by caller
template <int FI, int FO>
__global__ void __launch_bounds__(kThreads)
bdd_rel_scatter_kernel(const float* __restrict__ feat, const int4* __restrict__ pack, int E,
                       const float* __restrict__ wl, int B, int swap, float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  const int width = B * FO, in_w = B * FI;
  float* W_s = sm;                        // [FI][width]
  float* X_s = sm + FI * width;           // [2][kGroup][in_w]   (double-buffered cp.async staging)
  const int tpe = width / 4;              // threads per edge
  const int slots = kThreads / tpe;       // edges processed concurrently
  const int slot = threadIdx.x / tpe, j0 = (threadIdx.x % tpe) * 4;
  const int e0 = blockIdx.x * kChunk, e1 = min(E, e0 + kChunk);

  auto prefetch = [&](int e, int g, int buf) {
    for (int q = 0; q < g; ++q) {
      const int4 p = __ldg(pack + e + q);                 // {src, dst, etype, norm}
      stage_row_async(X_s + (buf * kGroup + q) * in_w, feat + (size_t)(swap ? p.y : p.x) * in_w, in_w);
    }
    async_commit();
  };

  int cur = -1, buf = 0;
  int e = e0, g = group_len(pack, e, e1);
  prefetch(e, g, 0);
  while (g > 0) {
    const int en = e + g, gn = group_len(pack, en, e1);
    prefetch(en, gn, buf ^ 1);                            // empty commit group when gn == 0
    const int r = __ldg(&pack[e].z);
    if (r != cur) {                                       // W_s is idle here: last reads were before the
      const float4* w4 = reinterpret_cast<const float4*>(wl + (size_t)r * FI * width);   // trailing barrier
      float4* d4 = reinterpret_cast<float4*>(W_s);
      for (int i = threadIdx.x; i < FI * width / 4; i += kThreads) d4[i] = __ldg(w4 + i);
      cur = r;
    }
    async_wait<1>();
    __syncthreads();
    if (slot < slots) {
      for (int q = slot; q < g; q += slots) {
        const int4 p = __ldg(pack + e + q);
        const float nv = __int_as_float(p.w);
        const float4 m = block_message<FI, FO>(X_s + (buf * kGroup + q) * in_w, W_s, width, j0);
        red_add_v4(out + (size_t)(swap ? p.x : p.y) * width + j0, nv * m.x, nv * m.y, nv * m.z, nv * m.w);
      }
    }
    __syncthreads();                                      // buffer and W_s free again
    e = en;
    g = gn;
    buf ^= 1;
  }
}

// smallest divisor of SO that leaves at most 32 accumulators (SI * SO / OS) per thread
__host__ __device__ constexpr int col_splits(int si, int so) {
  for (int os = 1; os <= so; ++os)
    if (so % os == 0 && si * so / os <= 32) return os;
  return so;
}

// fused backward: dx[src] += norm * blockdiag(W_r)^T dagg[dst]   and
//                 dW[r][b][i][o] += norm * x[src][b*SI+i] * dagg[dst][b*SO+o]
// w_bwd layout [R][SO][B*SI]; dx, dW zero-filled by the caller; dx may be null.
// Weight gradient: a thread owns one block b (or 1/OS of its output columns) and keeps the
// SI x SO/OS outer-product accumulators in registers: SI + SO/OS shared-memory reads per
// SI*SO/OS FMAs, stride-5 addresses (conflict-free).
template <int SI, int SO>
__global__ void __launch_bounds__(kThreads)
bdd_rel_backward_kernel(const float* __restrict__ x, const float* __restrict__ dagg,
                        const int4* __restrict__ pack, int E, const float* __restrict__ w_bwd, int B,
                        float* __restrict__ dx, float* __restrict__ dW) {
  extern __shared__ __align__(16) float sm[];
  constexpr int OS = col_splits(SI, SO);   // output-column splits per block: <= 32 accumulators per thread
  constexpr int SOS = SO / OS;
  const int in_w = B * SI, out_w = B * SO, KW = B * SI * SO;
  float* W_s = sm;                        // [SO][in_w]
  float* X_s = W_s + SO * in_w;           // [2][kGroup][in_w]    (double-buffered cp.async staging)
  float* D_s = X_s + 2 * kGroup * in_w;   // [2][kGroup][out_w]
  float acc[SI][SOS];
#pragma unroll
  for (int i = 0; i < SI; ++i)
#pragma unroll
    for (int o = 0; o < SOS; ++o) acc[i][o] = 0.f;
  // weight-gradient role
  const int tpw = B * OS, wslots = kThreads / tpw;
  const int wslot = threadIdx.x / tpw, wb = (threadIdx.x % tpw) / OS, oh = (threadIdx.x % tpw) % OS;
  // input-gradient role
  const int tpe = in_w / 4, slots = kThreads / tpe;
  const int slot = threadIdx.x / tpe, j0 = (threadIdx.x % tpe) * 4;
  const int e0 = blockIdx.x * kChunk, e1 = min(E, e0 + kChunk);
  int cur = -1;

  auto flush = [&](int r) {
    if (wslot < wslots) {
      float* dst = dW + (size_t)r * KW + wb * SI * SO + oh * SOS;
#pragma unroll
      for (int i = 0; i < SI; ++i)
#pragma unroll
        for (int o = 0; o < SOS; ++o) {
          atomicAdd(dst + i * SO + o, acc[i][o]);
          acc[i][o] = 0.f;
        }
    }
  };

  auto prefetch = [&](int e, int g, int buf) {
    for (int q = 0; q < g; ++q) {
      const int4 p = __ldg(pack + e + q);
      stage_row_async(X_s + (buf * kGroup + q) * in_w, x + (size_t)p.x * in_w, in_w);
      stage_row_async(D_s + (buf * kGroup + q) * out_w, dagg + (size_t)p.y * out_w, out_w);
    }
    async_commit();
  };

  int buf = 0;
  int e = e0, g = group_len(pack, e, e1);
  prefetch(e, g, 0);
  while (g > 0) {
    const int en = e + g, gn = group_len(pack, en, e1);
    prefetch(en, gn, buf ^ 1);
    const int r = __ldg(&pack[e].z);
    if (r != cur) {
      if (cur >= 0) flush(cur);
      if (dx) {
        const float4* w4 = reinterpret_cast<const float4*>(w_bwd + (size_t)r * SO * in_w);
        float4* d4 = reinterpret_cast<float4*>(W_s);
        for (int i = threadIdx.x; i < SO * in_w / 4; i += kThreads) d4[i] = __ldg(w4 + i);
      }
      cur = r;
    }
    async_wait<1>();
    __syncthreads();
    const float* Xb = X_s + buf * kGroup * in_w;
    const float* Db = D_s + buf * kGroup * out_w;
    if (wslot < wslots) {
      for (int t = wslot; t < g; t += wslots) {
        const float nv = __int_as_float(__ldg(&pack[e + t].w));
        const float* xs = Xb + t * in_w + wb * SI;
        const float* ds = Db + t * out_w + wb * SO + oh * SOS;
        float xv[SI], dv[SOS];
#pragma unroll
        for (int i = 0; i < SI; ++i) xv[i] = nv * xs[i];
#pragma unroll
        for (int o = 0; o < SOS; ++o) dv[o] = ds[o];
#pragma unroll
        for (int i = 0; i < SI; ++i)
#pragma unroll
          for (int o = 0; o < SOS; ++o) acc[i][o] = fmaf(xv[i], dv[o], acc[i][o]);
      }
    }
    // input gradient: same shape as the forward message with the transposed blocks
    if (dx && slot < slots) {
      for (int q = slot; q < g; q += slots) {
        const int4 p = __ldg(pack + e + q);
        const float nv = __int_as_float(p.w);
        const float4 m = block_message<SO, SI>(Db + q * out_w, W_s, in_w, j0);
        red_add_v4(dx + (size_t)p.x * in_w + j0, nv * m.x, nv * m.y, nv * m.z, nv * m.w);
      }
    }
    __syncthreads();
    e = en;
    g = gn;
    buf ^= 1;
  }
  if (cur >= 0) flush(cur);
}

// ------------------------------------------------------------------------------------------
// any block shape: run-time FI/FO, scalar reductions, dW accumulated in shared memory
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_row_any(float* dst, const float* __restrict__ src, int n) {
  for (int i = threadIdx.x; i < n; i += kThreads) dst[i] = __ldg(src + i);
}

__global__ void __launch_bounds__(kThreads)
bdd_rel_scatter_generic(const float* __restrict__ feat, const int4* __restrict__ pack, int E,
                        const float* __restrict__ wl, int B, int FI, int FO, int swap, float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  const int width = B * FO, in_w = B * FI;
  float* W_s = sm;
  float* X_s = sm + FI * width;
  const int e0 = blockIdx.x * kChunk, e1 = min(E, e0 + kChunk);
  int cur = -1;
  for (int e = e0; e < e1; ++e) {
    const int4 p = __ldg(pack + e);
    __syncthreads();
    if (p.z != cur) {
      stage_row_any(W_s, wl + (size_t)p.z * FI * width, FI * width);
      cur = p.z;
    }
    stage_row_any(X_s, feat + (size_t)(swap ? p.y : p.x) * in_w, in_w);
    __syncthreads();
    const float nv = __int_as_float(p.w);
    float* orow = out + (size_t)(swap ? p.x : p.y) * width;
    for (int j = threadIdx.x; j < width; j += kThreads) {
      const float* xs = X_s + (j / FO) * FI;
      float m = 0.f;
      for (int i = 0; i < FI; ++i) m = fmaf(xs[i], W_s[i * width + j], m);
      atomicAdd(orow + j, nv * m);
    }
  }
}

__global__ void __launch_bounds__(kThreads)
bdd_rel_backward_generic(const float* __restrict__ x, const float* __restrict__ dagg,
                         const int4* __restrict__ pack, int E, const float* __restrict__ w_bwd, int B, int SI,
                         int SO, float* __restrict__ dx, float* __restrict__ dW) {
  extern __shared__ __align__(16) float sm[];
  const int in_w = B * SI, out_w = B * SO, KW = B * SI * SO;
  float* W_s = sm;                 // [SO][in_w]
  float* A_s = W_s + SO * in_w;    // [KW] weight-gradient accumulators of the current relation
  float* X_s = A_s + KW;           // [in_w]
  float* D_s = X_s + in_w;         // [out_w]
  const int e0 = blockIdx.x * kChunk, e1 = min(E, e0 + kChunk);
  int cur = -1;
  for (int e = e0; e < e1; ++e) {
    const int4 p = __ldg(pack + e);
    __syncthreads();
    if (p.z != cur) {
      for (int k = threadIdx.x; k < KW; k += kThreads) {
        if (cur >= 0) atomicAdd(dW + (size_t)cur * KW + k, A_s[k]);
        A_s[k] = 0.f;
      }
      stage_row_any(W_s, w_bwd + (size_t)p.z * SO * in_w, SO * in_w);
      cur = p.z;
    }
    stage_row_any(X_s, x + (size_t)p.x * in_w, in_w);
    stage_row_any(D_s, dagg + (size_t)p.y * out_w, out_w);
    __syncthreads();
    const float nv = __int_as_float(p.w);
    for (int k = threadIdx.x; k < KW; k += kThreads) {
      const int b = k / (SI * SO), rem = k - b * (SI * SO);
      A_s[k] = fmaf(nv * X_s[b * SI + rem / SO], D_s[b * SO + rem % SO], A_s[k]);
    }
    if (dx) {
      float* drow = dx + (size_t)p.x * in_w;
      for (int j = threadIdx.x; j < in_w; j += kThreads) {
        const float* ds = D_s + (j / SI) * SO;
        float m = 0.f;
        for (int o = 0; o < SO; ++o) m = fmaf(ds[o], W_s[o * in_w + j], m);
        atomicAdd(drow + j, nv * m);
      }
    }
  }
  __syncthreads();
  if (cur >= 0)
    for (int k = threadIdx.x; k < KW; k += kThreads) atomicAdd(dW + (size_t)cur * KW + k, A_s[k]);
}

template <int FI, int FO>
int launch_scatter(const float* feat, const void* pack, int E, const float* wl, int B, int swap, float* out,
                   cudaStream_t st) {
  const int width = B * FO, in_w = B * FI;
  const size_t smem = sizeof(float) * ((size_t)FI * width + (size_t)2 * kGroup * in_w);
  auto kern = bdd_rel_scatter_kernel<FI, FO>;
  KG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<kg_div_up(E, kChunk), kThreads, smem, st>>>(feat, reinterpret_cast<const int4*>(pack), E, wl, B, swap, out);
  KG_LAUNCH_OK();
  return KG_OK;
}

template <int SI, int SO>
int launch_backward(const float* x, const float* dagg, const void* pack, int E, const float* w_bwd, int B,
                    float* dx, float* dW, cudaStream_t st) {
  const int in_w = B * SI, out_w = B * SO;
  const size_t smem = sizeof(float) * ((size_t)SO * in_w + (size_t)2 * kGroup * (in_w + out_w));
  auto kern = bdd_rel_backward_kernel<SI, SO>;
  KG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<kg_div_up(E, kChunk), kThreads, smem, st>>>(x, dagg, reinterpret_cast<const int4*>(pack), E, w_bwd, B, dx, dW);
  KG_LAUNCH_OK();
  return KG_OK;
}

bool fast_shape(int B, int si, int so) {
  const bool known = (si == 5 && so == 5) || (si == 5 && so == 10) || (si == 10 && so == 10) ||
                     (si == 4 && so == 4) || (si == 8 && so == 8);
  // vector width 4 on both feature widths, one thread per 4 columns, one thread per (block, column split)
  return known && (B * si) % 4 == 0 && (B * so) % 4 == 0 && (B * so) / 4 <= kThreads && (B * si) / 4 <= kThreads &&
         B * col_splits(si, so) <= kThreads;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

#define KG_BDD_DISPATCH(FN, SI_, SO_, ...) \
  if (si == SI_ && so == SO_) return FN<SI_, SO_>(__VA_ARGS__)

// agg[dst] += norm * blockdiag(W[etype]) x[src] over relation-sorted edges; agg zero-filled by the caller
extern "C" int kg_bdd_rel_fwd(const float* x, const void* rel_pack, int n_edges, const float* w_fwd,
                              int num_bases, int si, int so, float* agg, void* stream) {
  KG_REQUIRE(n_edges >= 0 && num_bases > 0 && si > 0 && so > 0, "bdd rel fwd: bad sizes");
  if (n_edges == 0) return KG_OK;
  cudaStream_t st = kg_stream(stream);
  if (fast_shape(num_bases, si, so) && aligned16(x) && aligned16(w_fwd) && aligned16(agg)) {
    KG_BDD_DISPATCH(launch_scatter, 5, 5, x, rel_pack, n_edges, w_fwd, num_bases, 0, agg, st);
    KG_BDD_DISPATCH(launch_scatter, 5, 10, x, rel_pack, n_edges, w_fwd, num_bases, 0, agg, st);
    KG_BDD_DISPATCH(launch_scatter, 10, 10, x, rel_pack, n_edges, w_fwd, num_bases, 0, agg, st);
    KG_BDD_DISPATCH(launch_scatter, 4, 4, x, rel_pack, n_edges, w_fwd, num_bases, 0, agg, st);
    KG_BDD_DISPATCH(launch_scatter, 8, 8, x, rel_pack, n_edges, w_fwd, num_bases, 0, agg, st);
  }
  const size_t smem = sizeof(float) * ((size_t)si * num_bases * so + (size_t)num_bases * si);
  KG_REQUIRE(smem <= 200 * 1024, "bdd rel fwd: block weights of one relation exceed shared memory");
  KG_CUDA(cudaFuncSetAttribute(bdd_rel_scatter_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  bdd_rel_scatter_generic<<<kg_div_up(n_edges, kChunk), kThreads, smem, st>>>(
      x, reinterpret_cast<const int4*>(rel_pack), n_edges, w_fwd, num_bases, si, so, 0, agg);
  KG_LAUNCH_OK();
  return KG_OK;
}

// dx (zero-filled, may be NULL) and dweight (zero-filled) of the same layer
extern "C" int kg_bdd_rel_bwd(const float* x, const float* dagg, const void* rel_pack, int n_edges,
                              const float* w_bwd, int num_bases, int si, int so, float* dx, float* dweight,
                              void* stream) {
  KG_REQUIRE(n_edges >= 0 && num_bases > 0 && si > 0 && so > 0, "bdd rel bwd: bad sizes");
  if (n_edges == 0) return KG_OK;
  cudaStream_t st = kg_stream(stream);
  if (fast_shape(num_bases, si, so) && aligned16(x) && aligned16(dagg) && aligned16(w_bwd) && aligned16(dx)) {
    KG_BDD_DISPATCH(launch_backward, 5, 5, x, dagg, rel_pack, n_edges, w_bwd, num_bases, dx, dweight, st);
    KG_BDD_DISPATCH(launch_backward, 5, 10, x, dagg, rel_pack, n_edges, w_bwd, num_bases, dx, dweight, st);
    KG_BDD_DISPATCH(launch_backward, 10, 10, x, dagg, rel_pack, n_edges, w_bwd, num_bases, dx, dweight, st);
    KG_BDD_DISPATCH(launch_backward, 4, 4, x, dagg, rel_pack, n_edges, w_bwd, num_bases, dx, dweight, st);
    KG_BDD_DISPATCH(launch_backward, 8, 8, x, dagg, rel_pack, n_edges, w_bwd, num_bases, dx, dweight, st);
  }
  const int in_w = num_bases * si, out_w = num_bases * so;
  const size_t smem = sizeof(float) * ((size_t)so * in_w + (size_t)num_bases * si * so + in_w + out_w);
  KG_REQUIRE(smem <= 200 * 1024, "bdd rel bwd: block weights of one relation exceed shared memory");
  KG_CUDA(cudaFuncSetAttribute(bdd_rel_backward_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  bdd_rel_backward_generic<<<kg_div_up(n_edges, kChunk), kThreads, smem, st>>>(
      x, dagg, reinterpret_cast<const int4*>(rel_pack), n_edges, w_bwd, num_bases, si, so, dx, dweight);
  KG_LAUNCH_OK();
  return KG_OK;
}
