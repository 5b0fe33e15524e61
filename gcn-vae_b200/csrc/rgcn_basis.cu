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
// integer-id features with ids == arange (the reference's `feats = torch.arange(num_nodes)`,
// kgvae/entity_classify.py:63): SOURCE-TILED kernels over the src-major edge list (col_ptr +
// records {dst, etype, norm, edge}).  A CTA takes NT consecutive source nodes, so the basis rows it
// needs - V[b, n0 .. n0+NT, :] for every b - are NB contiguous runs: the 2.67 GB table (AM shape)
// is read ONCE, coalesced, instead of one 40-byte row per (edge, basis), and its gradient is
// WRITTEN once with plain stores (each (b, node) row has exactly one owner thread): no atomics on
// dV, no zero-fill.  The rows reduced into (out[dst], 67 MB at the AM shape) are L2-resident.
// Nodes with more than kHeavy out-edges are finished by the whole CTA (hubs of real RDF graphs).
// ------------------------------------------------------------------------------------------
constexpr int kHeavy = 512;

// forward: thread (n, o) keeps the column V[0..NB, n, o] in registers; per out-edge NB FMAs against
// the relation's coefficient row (shared memory) and one atomic add into out[dst, o]
template <int NBR>
__global__ void __launch_bounds__(256, 2)
basis_id_src_fwd_kernel(const float* __restrict__ V, const float* __restrict__ coef, const int* __restrict__ col_ptr,
                        const int4* __restrict__ pack, int n_src, int R, int NB, int out_f, int NT,
                        float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  float* coef_s = sm;                                  // [R][NBR] (zero padded)
  float* Vh_s = sm + (size_t)R * NBR;                  // [NBR] column of a heavy node, per o: [out_f][NBR]
  for (int i = threadIdx.x; i < R * NBR; i += blockDim.x) {
    const int r = i / NBR, b = i - r * NBR;
    coef_s[i] = b < NB ? __ldg(coef + (size_t)r * NB + b) : 0.f;
  }
  __syncthreads();
  const int per_tile = NT * out_f;
  for (int n0 = blockIdx.x * NT; n0 < n_src; n0 += gridDim.x * NT) {
    const int nt = min(NT, n_src - n0);
    const int t = threadIdx.x;
    const bool mine = t < nt * out_f;
    const int n = n0 + (mine ? t / out_f : 0), o = mine ? t % out_f : 0;
    float v[NBR];
#pragma unroll
    for (int b = 0; b < NBR; ++b) v[b] = (mine && b < NB) ? __ldg(V + ((size_t)b * n_src + n) * out_f + o) : 0.f;
    int e0 = 0, e1 = 0;
    if (mine) { e0 = __ldg(col_ptr + n); e1 = __ldg(col_ptr + n + 1); }
    const bool heavy = e1 - e0 > kHeavy;
    if (!heavy) {
      for (int e = e0; e < e1; ++e) {
        const int4 p = __ldg(pack + e);                // {dst, etype, norm, edge}
        const float4* c4 = reinterpret_cast<const float4*>(coef_s + (size_t)p.y * NBR);
        float m = 0.f;
#pragma unroll
        for (int b = 0; b < NBR; b += 4) {
          const float4 c = c4[b / 4];
          m = fmaf(c.x, v[b], m); m = fmaf(c.y, v[b + 1], m); m = fmaf(c.z, v[b + 2], m); m = fmaf(c.w, v[b + 3], m);
        }
        atomicAdd(out + (size_t)p.x * out_f + o, __int_as_float(p.z) * m);
      }
    }
    // heavy nodes of this tile: the whole CTA shares the edges of one node at a time
    if (__syncthreads_or(heavy)) {
      for (int hn = 0; hn < nt; ++hn) {
        const int h0 = __ldg(col_ptr + n0 + hn), h1 = __ldg(col_ptr + n0 + hn + 1);
        if (h1 - h0 <= kHeavy) continue;               // uniform across the CTA
        __syncthreads();
        if (mine && t / out_f == hn)
#pragma unroll
          for (int b = 0; b < NBR; ++b) Vh_s[o * NBR + b] = v[b];
        __syncthreads();
        const int groups = blockDim.x / out_f, grp = t / out_f, oo = t % out_f;
        if (grp < groups) {
          const float4* v4 = reinterpret_cast<const float4*>(Vh_s + oo * NBR);
          for (int e = h0 + grp; e < h1; e += groups) {
            const int4 p = __ldg(pack + e);
            const float4* c4 = reinterpret_cast<const float4*>(coef_s + (size_t)p.y * NBR);
            float m = 0.f;
#pragma unroll
            for (int b = 0; b < NBR; b += 4) {
              const float4 c = c4[b / 4], w = v4[b / 4];
              m = fmaf(c.x, w.x, m); m = fmaf(c.y, w.y, m); m = fmaf(c.z, w.z, m); m = fmaf(c.w, w.w, m);
            }
            atomicAdd(out + (size_t)p.x * out_f + oo, __int_as_float(p.z) * m);
          }
        }
      }
      __syncthreads();
    }
    (void)per_tile;
  }
}

// backward: thread (n, b) keeps V[b, n, :] and the gradient row dV[b, n, :] in registers
//   dV[b, n, :]  = sum_{e: src = n} coef[r_e, b] * norm_e * g[dst_e, :]          (plain store, once)
//   dcoef[r, b] += sum_e norm_e * <V[b, n, :], g[dst_e, :]>   (shared-memory table, flushed per CTA)
// Global traffic is whole contiguous runs: the V tile ([b][nt*out_f] runs, odd pitch so that lanes
// walking b hit distinct banks) and the dV tile go through shared memory.  The tile's out-edges are
// one contiguous range of the src-major list: they are staged in chunks - record + norm-scaled
// g[dst] row, every load of a chunk in flight at once - and the (n, b) threads then walk their
// node's share of the chunk from shared memory, so the dependent latencies (col_ptr -> record ->
// g row) are paid once per chunk, not once per edge; hub nodes are simply many chunks.
constexpr int kEdgeChunk = 128;

template <int OF>
__global__ void __launch_bounds__(640, 2)
basis_id_src_bwd_kernel(const float* __restrict__ V, const float* __restrict__ coef, const float* __restrict__ g,
                        const int* __restrict__ col_ptr, const int4* __restrict__ pack, int n_src, int R, int NB,
                        int out_f, int NT, float* __restrict__ dV, float* __restrict__ dcoef) {
  extern __shared__ __align__(16) float sm[];
  const int pitch = (NT * out_f) | 1;
  float* G_s = sm;                                     // [kEdgeChunk][OF] norm * g[dst] (16-byte aligned rows)
  int4* E_s = reinterpret_cast<int4*>(G_s + kEdgeChunk * OF);   // [kEdgeChunk] records
  float* coef_s = reinterpret_cast<float*>(E_s + kEdgeChunk);   // [R][NB]
  float* dc_s = coef_s + (size_t)R * NB;               // [R][NB]
  float* T_base = dc_s + (size_t)R * NB;               // [2][NB][pitch]  V tile (double-buffered), then dV tile
  for (int i = threadIdx.x; i < R * NB; i += blockDim.x) {
    coef_s[i] = __ldg(coef + i);
    dc_s[i] = 0.f;
  }
  const int t = threadIdx.x;
  const int nl = t / NB, b = t - nl * NB;              // thread (node slot, basis)
  // The V tile of the NEXT node tile is fetched with cp.async while this tile's edges are walked: the launch
  // ran at 19 % of HBM peak because a CTA loaded, computed and stored strictly one after the other.
  auto fetch_tile = [&](int n0_, float* buf) {
    if (n0_ < n_src) {
      const int run_ = min(NT, n_src - n0_) * out_f;
      for (int i = t; i < NB * run_; i += blockDim.x) {
        const int bb = i / run_, j = i - bb * run_;
        const unsigned dst_s = static_cast<unsigned>(__cvta_generic_to_shared(buf + bb * pitch + j));
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_s), "l"(V + ((size_t)bb * n_src + n0_) * out_f + j)
                     : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  fetch_tile(blockIdx.x * NT, T_base);
  int it = 0;
  for (int n0 = blockIdx.x * NT; n0 < n_src; n0 += gridDim.x * NT, ++it) {
    const int nt = min(NT, n_src - n0), run = nt * out_f;
    const bool mine = nl < nt;
    float* T_s = T_base + (size_t)(it & 1) * NB * pitch;
    __syncthreads();                                   // previous tile written out (its buffer is refilled now); tables ready
    fetch_tile(n0 + gridDim.x * NT, T_base + (size_t)((it & 1) ^ 1) * NB * pitch);
    asm volatile("cp.async.wait_group 1;" ::: "memory");   // this tile has landed (the next one may still be in flight)
    const int e_lo = __ldg(col_ptr + n0), e_hi = __ldg(col_ptr + n0 + nt);
    int e0 = 0, e1 = 0;
    if (mine) { e0 = __ldg(col_ptr + n0 + nl); e1 = __ldg(col_ptr + n0 + nl + 1); }
    __syncthreads();
    float v[OF], acc[OF];
#pragma unroll
    for (int o = 0; o < OF; ++o) {
      v[o] = (mine && o < out_f) ? T_s[b * pitch + nl * out_f + o] : 0.f;
      acc[o] = 0.f;
    }
    for (int c0 = e_lo; c0 < e_hi; c0 += kEdgeChunk) { // uniform across the CTA
      const int cn = min(kEdgeChunk, e_hi - c0);
      __syncthreads();                                 // previous chunk consumed
      for (int j = t; j < cn; j += blockDim.x) {
        const int4 p = __ldg(pack + c0 + j);           // {dst, etype, norm, edge}
        E_s[j] = p;
        const float nv = __int_as_float(p.z);
        const float* gr = g + (size_t)p.x * out_f;
#pragma unroll
        for (int o = 0; o < OF; ++o) G_s[j * OF + o] = o < out_f ? nv * __ldg(gr + o) : 0.f;
      }
      __syncthreads();
      const int a0 = max(e0, c0) - c0, a1 = min(e1, c0 + cn) - c0;
      for (int e = a0; e < a1; ++e) {
        const int r = E_s[e].y;
        const float c = coef_s[r * NB + b];
        float d = 0.f;
#pragma unroll
        for (int o = 0; o < OF; o += 4) {
          const float4 gv = *reinterpret_cast<const float4*>(G_s + e * OF + o);
          d = fmaf(v[o], gv.x, d); d = fmaf(v[o + 1], gv.y, d); d = fmaf(v[o + 2], gv.z, d); d = fmaf(v[o + 3], gv.w, d);
          acc[o] = fmaf(c, gv.x, acc[o]); acc[o + 1] = fmaf(c, gv.y, acc[o + 1]);
          acc[o + 2] = fmaf(c, gv.z, acc[o + 2]); acc[o + 3] = fmaf(c, gv.w, acc[o + 3]);
        }
        atomicAdd(dc_s + r * NB + b, d);
      }
    }
    __syncthreads();                                   // every thread holds its V row: the tile becomes dV
    if (mine)
#pragma unroll
      for (int o = 0; o < OF; ++o)
        if (o < out_f) T_s[b * pitch + nl * out_f + o] = acc[o];
    __syncthreads();
    for (int i = t; i < NB * run; i += blockDim.x) {
      const int bb = i / run, j = i - bb * run;
      dV[((size_t)bb * n_src + n0) * out_f + j] = T_s[bb * pitch + j];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < R * NB; i += blockDim.x)
    if (dc_s[i] != 0.f) atomicAdd(dcoef + i, dc_s[i]);
}

// node-tile size and CTA size of the source-tiled id kernels; 0 when the shape does not fit
struct SrcPlan {
  int nt, threads;
  size_t smem_fwd, smem_bwd;
  int nbr;
};
SrcPlan src_plan(int R, int NB, int out_f) {
  SrcPlan p{0, 0, 0, 0, 0};
  if (NB > 64 || out_f > 16 || NB < 2) return p;
  p.nbr = NB <= 16 ? 16 : NB <= 32 ? 32 : NB <= 48 ? 48 : 64;
  p.nt = 640 / NB;                                     // backward: NT * NB threads
  if (p.nt > 32) p.nt = 32;
  if (p.nt < 1) return SrcPlan{0, 0, 0, 0, 0};
  p.threads = (p.nt * NB + 31) / 32 * 32;
  p.smem_fwd = sizeof(float) * ((size_t)R * p.nbr + (size_t)out_f * p.nbr);
  p.smem_bwd = sizeof(float) * ((size_t)kEdgeChunk * (16 + 4) + (size_t)2 * R * NB + (size_t)2 * NB * ((p.nt * out_f) | 1));
  if (p.smem_fwd > 100 * 1024 || p.smem_bwd > 112 * 1024) return SrcPlan{0, 0, 0, 0, 0};   // two CTAs per SM (227 KB)
  return p;
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


// ------------------------------------------------------------------------------------------
// dense features, SMALL layers (in_feat, out_feat <= 16: the entity-classification hidden / output
// layers, e.g. 10 -> 11 at the AM shape): every W_r ([R, in, out], 58 KB at AM) sits in shared
// memory, ONE THREAD PER EDGE does the whole in x out product from registers - no per-edge CTA
// barrier, lanes of a warp read the same W_r (relation-sorted records: a broadcast).
//   forward  out[dst, o] += sum_i (norm x[src, i]) W_r[i, o]                (W transposed + padded)
//   backward dx[src, i]  += norm sum_o W_r[i, o] g[dst, o]
//            dW_r[i, o]  += sum_e norm x[src, i] g[dst, o]: warp-reduced over the 32 edges of a
//            batch, accumulated per warp in shared memory, flushed when the relation changes
// ------------------------------------------------------------------------------------------
constexpr int kSmall = 16;          // register rows: in_feat, out_feat <= 16
constexpr int kSmallRange = 2048;   // consecutive edges per warp visit (backward)

__device__ __forceinline__ void load_row16(float (&v)[kSmall], const float* __restrict__ row, int n, float s) {
#pragma unroll
  for (int i = 0; i < kSmall; ++i) v[i] = i < n ? s * __ldg(row + i) : 0.f;
}

__global__ void __launch_bounds__(kThreads)
basis_small_fwd_kernel(const float* __restrict__ x, const int4* __restrict__ pack, int E, const float* __restrict__ W,
                       int R, int in_f, int out_f, float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];          // Wt_s[R][out_f][ips]
  const int ips = (in_f + 3) & ~3;
  for (int idx = threadIdx.x; idx < R * out_f * ips; idx += blockDim.x) {
    const int i = idx % ips, ro = idx / ips, o = ro % out_f, r = ro / out_f;
    sm[idx] = i < in_f ? __ldg(W + ((size_t)r * in_f + i) * out_f + o) : 0.f;
  }
  __syncthreads();
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (long long)gridDim.x * blockDim.x) {
    const int4 p = __ldg(pack + e);                    // {src, dst, etype, norm}
    float xv[kSmall];
    load_row16(xv, x + (size_t)p.x * in_f, in_f, __int_as_float(p.w));
    const float* wr = sm + (size_t)p.z * out_f * ips;
    float* orow = out + (size_t)p.y * out_f;
    for (int o = 0; o < out_f; ++o) {
      const float4* w4 = reinterpret_cast<const float4*>(wr + o * ips);
      float m = 0.f;
#pragma unroll
      for (int q = 0; q < kSmall / 4; ++q)
        if (4 * q < ips) {
          const float4 w = w4[q];
          m = fmaf(xv[4 * q], w.x, m); m = fmaf(xv[4 * q + 1], w.y, m);
          m = fmaf(xv[4 * q + 2], w.z, m); m = fmaf(xv[4 * q + 3], w.w, m);
        }
      atomicAdd(orow + o, m);
    }
  }
}

__global__ void __launch_bounds__(kThreads)
basis_small_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g, const int4* __restrict__ pack, int E,
                       const float* __restrict__ W, int R, int in_f, int out_f, float* __restrict__ dx,
                       float* __restrict__ dW) {
  extern __shared__ __align__(16) float sm[];          // W_s[R][in_f][ops], then dW_w[warps][in_f * out_f]
  const int ops = (out_f + 3) & ~3, io = in_f * out_f;
  float* acc_all = sm + (size_t)R * in_f * ops;
  for (int idx = threadIdx.x; idx < R * in_f * ops; idx += blockDim.x) {
    const int o = idx % ops, ri = idx / ops;
    sm[idx] = o < out_f ? __ldg(W + (size_t)ri * out_f + o) : 0.f;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  float* acc = acc_all + warp * io;
  for (int idx = lane; idx < io; idx += 32) acc[idx] = 0.f;
  __syncthreads();
  int cur = -1;
  auto flush = [&]() {
    if (cur >= 0) {
      __syncwarp();
      for (int idx = lane; idx < io; idx += 32) {
        const float v = acc[idx];
        if (v != 0.f) atomicAdd(dW + (size_t)cur * io + idx, v);
        acc[idx] = 0.f;
      }
      __syncwarp();
    }
  };
  const long long n_ranges = ((long long)E + kSmallRange - 1) / kSmallRange;
  for (long long rg = (long long)blockIdx.x * warps + warp; rg < n_ranges; rg += (long long)gridDim.x * warps) {
    const long long r0 = rg * kSmallRange, r1 = r0 + kSmallRange < E ? r0 + kSmallRange : E;
    for (long long b0 = r0; b0 < r1; b0 += 32) {
      const long long e = b0 + lane;
      const bool valid = e < r1;
      const int4 p = valid ? __ldg(pack + e) : make_int4(0, 0, -1, 0);
      const float nv = __int_as_float(p.w);
      float xv[kSmall], gv[kSmall];
      load_row16(xv, x + (size_t)p.x * in_f, valid ? in_f : 0, nv);
      load_row16(gv, g + (size_t)p.y * out_f, valid ? out_f : 0, 1.f);
      if (dx != nullptr && valid) {
        const float* wr = sm + (size_t)p.z * in_f * ops;
        float* drow = dx + (size_t)p.x * in_f;
        for (int i = 0; i < in_f; ++i) {
          const float4* w4 = reinterpret_cast<const float4*>(wr + i * ops);
          float m = 0.f;
#pragma unroll
          for (int q = 0; q < kSmall / 4; ++q)
            if (4 * q < ops) {
              const float4 w = w4[q];
              m = fmaf(gv[4 * q], w.x, m); m = fmaf(gv[4 * q + 1], w.y, m);
              m = fmaf(gv[4 * q + 2], w.z, m); m = fmaf(gv[4 * q + 3], w.w, m);
            }
          atomicAdd(drow + i, nv * m);
        }
      }
      // weight gradient: one pass per distinct relation in the batch (sorted records: usually one)
      unsigned pending = __ballot_sync(0xffffffffu, valid);
      while (pending) {
        const int r = __shfl_sync(0xffffffffu, p.z, __ffs(pending) - 1);
        if (r != cur) {
          flush();
          cur = r;
        }
        const bool mine = valid && p.z == r;
#pragma unroll
        for (int i = 0; i < kSmall; ++i)
          if (i < in_f) {
            const float xi = mine ? xv[i] : 0.f;
#pragma unroll
            for (int o = 0; o < kSmall; ++o)
              if (o < out_f) {
                const float t = kg_warp_sum(xi * gv[o]);
                if (lane == 0) acc[i * out_f + o] += t;
              }
          }
        pending &= ~__ballot_sync(0xffffffffu, mine);
      }
    }
  }
  flush();
}

// shared memory of the small-layer kernels; 0 when the layer does not qualify
size_t small_smem(int R, int in_f, int out_f, bool backward) {
  if (in_f > kSmall || out_f > kSmall) return 0;
  const int ips = (in_f + 3) & ~3, ops = (out_f + 3) & ~3;
  const size_t bytes = backward ? sizeof(float) * ((size_t)R * in_f * ops + (size_t)(kThreads / 32) * in_f * out_f)
                                : sizeof(float) * (size_t)R * out_f * ips;
  return bytes <= 160 * 1024 ? bytes : 0;
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


// 1 when kg_basis_id_src_fwd / kg_basis_id_src_bwd cover this shape (identity ids, composed basis)
extern "C" int kg_basis_id_src_eligible(int num_rels, int num_bases, int out_feat) {
  return src_plan(num_rels, num_bases, out_feat).nt > 0 ? 1 : 0;
}

// Source-tiled forward for ids == arange(n_src) and coef != NULL: out (holding the self-loop term or
// zeros) += messages; col_ptr / bwd_pack: the src-major list of kg_graph_index.
extern "C" int kg_basis_id_src_fwd(const float* V, const float* coef, const int32_t* col_ptr, const void* bwd_pack,
                                   int n_src, int num_rels, int num_bases, int out_feat, float* out, void* stream) {
  KG_REQUIRE(V && coef && col_ptr && bwd_pack && out, "basis id src fwd: null argument");
  const SrcPlan pl = src_plan(num_rels, num_bases, out_feat);
  KG_REQUIRE(pl.nt > 0, "basis id src fwd: shape not covered (num_bases <= 64, out_feat <= 16)");
  if (n_src == 0) return KG_OK;
  // forward tile: as many nodes as give <= 256 (node, column) threads (two CTAs share an SM)
  int nt = 256 / out_feat;
  if (nt > 32) nt = 32;
  const int threads = (nt * out_feat + 31) / 32 * 32;
  const int tiles = kg_div_up(n_src, nt);
  const int grid = tiles < 8 * kg_sm_count() ? tiles : 8 * kg_sm_count();
  cudaStream_t st = kg_stream(stream);
  const int4* pack = reinterpret_cast<const int4*>(bwd_pack);
#define KG_SRC_FWD(NBR_)                                                                                          \
  if (pl.nbr == NBR_) {                                                                                           \
    KG_CUDA(cudaFuncSetAttribute(basis_id_src_fwd_kernel<NBR_>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                 (int)pl.smem_fwd));                                                              \
    basis_id_src_fwd_kernel<NBR_><<<grid, threads, pl.smem_fwd, st>>>(V, coef, col_ptr, pack, n_src, num_rels,    \
                                                                      num_bases, out_feat, nt, out);              \
  }
  KG_SRC_FWD(16) KG_SRC_FWD(32) KG_SRC_FWD(48) KG_SRC_FWD(64)
#undef KG_SRC_FWD
  KG_LAUNCH_OK();
  return KG_OK;
}

// Source-tiled backward: dV [NB, n_src, out] is WRITTEN (every row, no zero-fill needed); dcoef
// [R, NB] is accumulated into (zero-filled by the caller).
extern "C" int kg_basis_id_src_bwd(const float* V, const float* coef, const float* g, const int32_t* col_ptr,
                                   const void* bwd_pack, int n_src, int num_rels, int num_bases, int out_feat,
                                   float* dV, float* dcoef, void* stream) {
  KG_REQUIRE(V && coef && g && col_ptr && bwd_pack && dV && dcoef, "basis id src bwd: null argument");
  const SrcPlan pl = src_plan(num_rels, num_bases, out_feat);
  KG_REQUIRE(pl.nt > 0, "basis id src bwd: shape not covered (num_bases <= 64, out_feat <= 16)");
  if (n_src == 0) return KG_OK;
  const int tiles = kg_div_up(n_src, pl.nt);
  const int grid = tiles < 6 * kg_sm_count() ? tiles : 6 * kg_sm_count();
  cudaStream_t st = kg_stream(stream);
  const int4* pack = reinterpret_cast<const int4*>(bwd_pack);
  if (out_feat <= 12) {
    KG_CUDA(cudaFuncSetAttribute(basis_id_src_bwd_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bwd));
    basis_id_src_bwd_kernel<12><<<grid, pl.threads, pl.smem_bwd, st>>>(V, coef, g, col_ptr, pack, n_src, num_rels,
                                                                       num_bases, out_feat, pl.nt, dV, dcoef);
  } else {
    KG_CUDA(cudaFuncSetAttribute(basis_id_src_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bwd));
    basis_id_src_bwd_kernel<16><<<grid, pl.threads, pl.smem_bwd, st>>>(V, coef, g, col_ptr, pack, n_src, num_rels,
                                                                       num_bases, out_feat, pl.nt, dV, dcoef);
  }
  KG_LAUNCH_OK();
  return KG_OK;
}

// W [R, in, out] composed by the caller; out zero-filled (or holding the self-loop term)
extern "C" int kg_basis_dense_fwd(const float* x, const void* rel_pack, int n_edges, const float* W, int num_rels,
                                  int in_feat, int out_feat, float* out, void* stream) {
  KG_REQUIRE(n_edges >= 0 && in_feat > 0 && out_feat > 0 && num_rels > 0, "basis dense fwd: bad sizes");
  if (n_edges == 0) return KG_OK;
  if (const size_t smem = small_smem(num_rels, in_feat, out_feat, false)) {
    KG_CUDA(cudaFuncSetAttribute(basis_small_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int want = kg_div_up(n_edges, kThreads), cap = 4 * kg_sm_count();
    basis_small_fwd_kernel<<<want < cap ? want : cap, kThreads, smem, kg_stream(stream)>>>(
        x, reinterpret_cast<const int4*>(rel_pack), n_edges, W, num_rels, in_feat, out_feat, out);
    KG_LAUNCH_OK();
    return KG_OK;
  }
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
                                  const float* W, int num_rels, int in_feat, int out_feat, float* dx, float* dW,
                                  void* stream) {
  KG_REQUIRE(n_edges >= 0 && in_feat > 0 && out_feat > 0 && num_rels > 0, "basis dense bwd: bad sizes");
  if (n_edges == 0) return KG_OK;
  if (const size_t smem = small_smem(num_rels, in_feat, out_feat, true)) {
    KG_CUDA(cudaFuncSetAttribute(basis_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int want = kg_div_up(kg_div_up(n_edges, kSmallRange), kThreads / 32), cap = 2 * kg_sm_count();
    basis_small_bwd_kernel<<<want < cap ? want : cap, kThreads, smem, kg_stream(stream)>>>(
        x, g, reinterpret_cast<const int4*>(rel_pack), n_edges, W, num_rels, in_feat, out_feat, dx, dW);
    KG_LAUNCH_OK();
    return KG_OK;
  }
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
