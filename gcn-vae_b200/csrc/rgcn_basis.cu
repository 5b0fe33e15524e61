// a4: RelGraphConv(regularizer="basis") message passing - the entity-classification layers
// (reference kgvae/entity_classify.py:30-43, DGL basis_message_func).
//
//   W_r = sum_b w_comp[r, b] V_b            (V [NB, in, out]; W_r = V_r when NB == R)
//   dense features   msg_e = norm_e * x[src_e] @ W_{r_e}
//   integer node ids msg_e = norm_e * W_{r_e}[id_src_e, :]     (embedding-style lookup, in = num_nodes)
//
// The reference materialises W as [R, in, out] and, for integer features, indexes it as a
// [R*in, out] table - 8.9 GB at the AM shape (R = 133, in = 1.67 M nodes, out = 10), which is why it
// runs that dataset on the CPU.  Here the id path never forms W: every edge combines the NB basis
// rows V[b, id, :] with its relation's coefficients on the fly.  The dense path keeps a column
// tile of W_r in shared memory while a run of same-relation edges streams through (edges in
// (etype, dst) order, as in rgcn_bdd_rel.cu); W itself ([R, in, out], small whenever features are
// dense) is composed by the caller with the GEMM entry point.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kChunk = 128;     // relation-sorted edges per CTA
constexpr int kMaxOut = 32;     // id path: output columns kept in registers per lane

// ------------------------------------------------------------------------------------------
// integer-id features, forward: one warp per destination row (dst-CSR), lane b over the bases
//   out[v, :] += sum_{e -> v} norm_e * sum_b coef[r_e, b] * V[b, id_e, :]
// coef == nullptr means NB == R and W_r = V_r (one basis row per edge).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
basis_id_fwd_kernel(const float* __restrict__ V, const float* __restrict__ coef, const int* __restrict__ ids,
                    const int* __restrict__ row_ptr, const int4* __restrict__ fwd_pack, int n_dst, int n_in,
                    int NB, int out_f, float* __restrict__ out) {
  const int v = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (v >= n_dst) return;
  float acc[kMaxOut];
#pragma unroll
  for (int o = 0; o < kMaxOut; ++o) acc[o] = 0.f;
  const int e1 = __ldg(row_ptr + v + 1);
  for (int e = __ldg(row_ptr + v); e < e1; ++e) {
    const int4 p = __ldg(fwd_pack + e);              // {src, etype, norm, dst}
    const float nv = __int_as_float(p.z);
    const int id = ids ? __ldg(ids + p.x) : p.x;
    if (coef) {
      for (int b = lane; b < NB; b += 32) {
        const float c = nv * __ldg(coef + (size_t)p.y * NB + b);
        const float* row = V + ((size_t)b * n_in + id) * out_f;
#pragma unroll
        for (int o = 0; o < kMaxOut; ++o)
          if (o < out_f) acc[o] = fmaf(c, __ldg(row + o), acc[o]);
      }
    } else if (lane == 0) {
      const float* row = V + ((size_t)p.y * n_in + id) * out_f;
#pragma unroll
      for (int o = 0; o < kMaxOut; ++o)
        if (o < out_f) acc[o] = fmaf(nv, __ldg(row + o), acc[o]);
    }
  }
#pragma unroll
  for (int o = 0; o < kMaxOut; ++o) {
    if (o < out_f) {
      const float s = kg_warp_sum(acc[o]);
      if (lane == 0) out[(size_t)v * out_f + o] += s;   // row owned by this warp: out may hold the self-loop rows
    }
  }
}

// integer-id features, backward over relation-sorted edges: warp per chunk, lane b over the bases
//   dV[b, id, :]   += coef[r, b] * norm * g[dst, :]
//   dcoef[r, b]    += norm * <V[b, id, :], g[dst, :]>
__global__ void __launch_bounds__(kThreads)
basis_id_bwd_kernel(const float* __restrict__ V, const float* __restrict__ coef, const int* __restrict__ ids,
                    const float* __restrict__ g, const int4* __restrict__ rel_pack, int E, int n_in, int NB,
                    int out_f, float* __restrict__ dV, float* __restrict__ dcoef) {
  const int warp = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  const int e0 = warp * 32, e1 = min(E, e0 + 32);
  if (e0 >= E) return;
  for (int b = lane; b < (coef ? NB : 1); b += 32) {
    float dc = 0.f;
    int cur = -1;
    for (int e = e0; e < e1; ++e) {
      const int4 p = __ldg(rel_pack + e);            // {src, dst, etype, norm}
      const float nv = __int_as_float(p.w);
      const int id = ids ? __ldg(ids + p.x) : p.x;
      if (p.z != cur) {
        if (cur >= 0 && coef) atomicAdd(dcoef + (size_t)cur * NB + b, dc);
        dc = 0.f;
        cur = p.z;
      }
      const int bb = coef ? b : p.z;
      const float c = coef ? nv * __ldg(coef + (size_t)p.z * NB + b) : nv;
      const float* row = V + ((size_t)bb * n_in + id) * out_f;
      float* drow = dV + ((size_t)bb * n_in + id) * out_f;
      const float* gr = g + (size_t)p.y * out_f;
      for (int o = 0; o < out_f; ++o) {
        const float gv = __ldg(gr + o);
        dc = fmaf(nv * __ldg(row + o), gv, dc);
        atomicAdd(drow + o, c * gv);
      }
    }
    if (cur >= 0 && coef) atomicAdd(dcoef + (size_t)cur * NB + b, dc);
  }
}

// ------------------------------------------------------------------------------------------
// dense features over relation-sorted edges; CTA = edge chunk x column tile [c0, c0 + cw) of W_r
//   forward : out[dst, c] += norm * sum_i x[src, i] W_r[i, c]
//   backward: dx[src, i]  += norm * sum_c W_r[i, c] g[dst, c];   dW_r[i, c] += norm * x[src, i] g[dst, c]
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
basis_dense_fwd_kernel(const float* __restrict__ x, const int4* __restrict__ pack, int E,
                       const float* __restrict__ W, int in_f, int out_f, int cw, float* __restrict__ out) {
  extern __shared__ float sm[];
  float* W_s = sm;               // [in_f][cw]
  float* X_s = sm + in_f * cw;   // [in_f]
  const int c0 = blockIdx.y * cw, cn = min(cw, out_f - c0);
  const int e0 = blockIdx.x * kChunk, e1 = min(E, e0 + kChunk);
  int cur = -1;
  for (int e = e0; e < e1; ++e) {
    const int4 p = __ldg(pack + e);
    __syncthreads();
    if (p.z != cur) {
      for (int t = threadIdx.x; t < in_f * cn; t += kThreads)
        W_s[(t / cn) * cw + t % cn] = __ldg(W + ((size_t)p.z * in_f + t / cn) * out_f + c0 + t % cn);
      cur = p.z;
    }
    for (int i = threadIdx.x; i < in_f; i += kThreads) X_s[i] = __ldg(x + (size_t)p.x * in_f + i);
    __syncthreads();
    const float nv = __int_as_float(p.w);
    for (int c = threadIdx.x; c < cn; c += kThreads) {
      float m = 0.f;
      for (int i = 0; i < in_f; ++i) m = fmaf(X_s[i], W_s[i * cw + c], m);
      atomicAdd(out + (size_t)p.y * out_f + c0 + c, nv * m);
    }
  }
}

__global__ void __launch_bounds__(kThreads)
basis_dense_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g, const int4* __restrict__ pack,
                       int E, const float* __restrict__ W, int in_f, int out_f, int cw,
                       float* __restrict__ dx, float* __restrict__ dW) {
  extern __shared__ float sm[];
  float* W_s = sm;                    // [in_f][cw]
  float* A_s = W_s + in_f * cw;       // [in_f][cw] dW accumulators of the current relation
  float* X_s = A_s + in_f * cw;       // [in_f]
  float* G_s = X_s + in_f;            // [cw]
  const int c0 = blockIdx.y * cw, cn = min(cw, out_f - c0);
  const int e0 = blockIdx.x * kChunk, e1 = min(E, e0 + kChunk);
  int cur = -1;
  for (int e = e0; e < e1; ++e) {
    const int4 p = __ldg(pack + e);
    __syncthreads();
    if (p.z != cur) {
      for (int t = threadIdx.x; t < in_f * cn; t += kThreads) {
        const int i = t / cn, c = t % cn;
        if (cur >= 0) atomicAdd(dW + ((size_t)cur * in_f + i) * out_f + c0 + c, A_s[i * cw + c]);
        A_s[i * cw + c] = 0.f;
        W_s[i * cw + c] = __ldg(W + ((size_t)p.z * in_f + i) * out_f + c0 + c);
      }
      cur = p.z;
    }
    for (int i = threadIdx.x; i < in_f; i += kThreads) X_s[i] = __ldg(x + (size_t)p.x * in_f + i);
    for (int c = threadIdx.x; c < cn; c += kThreads) G_s[c] = __ldg(g + (size_t)p.y * out_f + c0 + c);
    __syncthreads();
    const float nv = __int_as_float(p.w);
    for (int t = threadIdx.x; t < in_f * cn; t += kThreads) {
      const int i = t / cn, c = t % cn;
      A_s[i * cw + c] = fmaf(nv * X_s[i], G_s[c], A_s[i * cw + c]);
    }
    if (dx) {
      for (int i = threadIdx.x; i < in_f; i += kThreads) {
        float m = 0.f;
        for (int c = 0; c < cn; ++c) m = fmaf(W_s[i * cw + c], G_s[c], m);
        atomicAdd(dx + (size_t)p.x * in_f + i, nv * m);
      }
    }
  }
  __syncthreads();
  if (cur >= 0)
    for (int t = threadIdx.x; t < in_f * cn; t += kThreads)
      atomicAdd(dW + ((size_t)cur * in_f + t / cn) * out_f + c0 + t % cn, A_s[(t / cn) * cw + t % cn]);
}

int col_tile(int in_f, int out_f, int arrays) {
  // widest column tile whose `arrays` [in_f][cw] shared arrays stay under 96 KB
  long long cw = (96LL * 1024 / 4) / ((long long)arrays * in_f);
  if (cw > out_f) cw = out_f;
  if (cw > 256) cw = 256;
  return (int)(cw < 1 ? 0 : cw);
}

}  // namespace

// out must hold the self-loop contribution (or zeros) on entry; messages are added on top.
// ids: optional [n_src] int32 map from source row to row of V (the 1-D feature tensor); NULL = identity.
extern "C" int kg_basis_id_fwd(const float* V, const float* coef, const int32_t* ids, const int32_t* row_ptr,
                               const void* fwd_pack, int n_dst, int n_in, int num_bases, int out_feat, float* out,
                               void* stream) {
  KG_REQUIRE(n_dst >= 0 && n_in > 0 && num_bases > 0 && out_feat > 0, "basis id fwd: bad sizes");
  KG_REQUIRE(out_feat <= kMaxOut, "basis id fwd: out_feat > 32 with integer features is not supported");
  if (n_dst == 0) return KG_OK;
  basis_id_fwd_kernel<<<kg_div_up((long long)n_dst * 32, kThreads), kThreads, 0, kg_stream(stream)>>>(
      V, coef, ids, row_ptr, reinterpret_cast<const int4*>(fwd_pack), n_dst, n_in, num_bases, out_feat, out);
  KG_LAUNCH_OK();
  return KG_OK;
}

// dV [NB, n_in, out] and dcoef [R, NB] (NULL when coef is NULL) zero-filled by the caller
extern "C" int kg_basis_id_bwd(const float* V, const float* coef, const int32_t* ids, const float* g,
                               const void* rel_pack, int n_edges, int n_in, int num_bases, int out_feat,
                               float* dV, float* dcoef, void* stream) {
  KG_REQUIRE(n_edges >= 0 && n_in > 0 && num_bases > 0 && out_feat > 0, "basis id bwd: bad sizes");
  KG_REQUIRE((coef == nullptr) == (dcoef == nullptr), "basis id bwd: coef and dcoef go together");
  if (n_edges == 0) return KG_OK;
  const int warps = kg_div_up(n_edges, 32);
  basis_id_bwd_kernel<<<kg_div_up((long long)warps * 32, kThreads), kThreads, 0, kg_stream(stream)>>>(
      V, coef, ids, g, reinterpret_cast<const int4*>(rel_pack), n_edges, n_in, num_bases, out_feat, dV, dcoef);
  KG_LAUNCH_OK();
  return KG_OK;
}

// W [R, in, out] composed by the caller; out zero-filled (or holding the self-loop term)
extern "C" int kg_basis_dense_fwd(const float* x, const void* rel_pack, int n_edges, const float* W, int in_feat,
                                  int out_feat, float* out, void* stream) {
  KG_REQUIRE(n_edges >= 0 && in_feat > 0 && out_feat > 0, "basis dense fwd: bad sizes");
  if (n_edges == 0) return KG_OK;
  const int cw = col_tile(in_feat, out_feat, 1);
  KG_REQUIRE(cw > 0, "basis dense fwd: in_feat too large for one shared-memory column");
  const size_t smem = sizeof(float) * ((size_t)in_feat * cw + in_feat);
  KG_CUDA(cudaFuncSetAttribute(basis_dense_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(kg_div_up(n_edges, kChunk), kg_div_up(out_feat, cw));
  basis_dense_fwd_kernel<<<grid, kThreads, smem, kg_stream(stream)>>>(
      x, reinterpret_cast<const int4*>(rel_pack), n_edges, W, in_feat, out_feat, cw, out);
  KG_LAUNCH_OK();
  return KG_OK;
}

// dx [n_src, in] (may be NULL) and dW [R, in, out] zero-filled by the caller
extern "C" int kg_basis_dense_bwd(const float* x, const float* g, const void* rel_pack, int n_edges,
                                  const float* W, int in_feat, int out_feat, float* dx, float* dW, void* stream) {
  KG_REQUIRE(n_edges >= 0 && in_feat > 0 && out_feat > 0, "basis dense bwd: bad sizes");
  if (n_edges == 0) return KG_OK;
  const int cw = col_tile(in_feat, out_feat, 2);
  KG_REQUIRE(cw > 0, "basis dense bwd: in_feat too large for one shared-memory column");
  const size_t smem = sizeof(float) * ((size_t)2 * in_feat * cw + in_feat + cw);
  KG_CUDA(cudaFuncSetAttribute(basis_dense_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(kg_div_up(n_edges, kChunk), kg_div_up(out_feat, cw));
  basis_dense_bwd_kernel<<<grid, kThreads, smem, kg_stream(stream)>>>(
      x, g, reinterpret_cast<const int4*>(rel_pack), n_edges, W, in_feat, out_feat, cw, dx, dW);
  KG_LAUNCH_OK();
  return KG_OK;
}
