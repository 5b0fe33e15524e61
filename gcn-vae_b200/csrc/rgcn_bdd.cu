// a3: RelGraphConv(regularizer="bdd") message passing, forward and backward, plus the small
// row-wise kernels around it (embedding lookup a2, activation/dropout backward, column sums).
//
// Replaces DGL 0.4.x bdd_message_func + update_all(fn.sum) behind RelGraphConv.forward
// (constructed at reference kgvae/model.py:54-59, called at :110-111).  The reference
// materialises a per-edge weight tensor [E, B*si*so] (10-20 KB per edge), runs E*B tiny bmm's,
// multiplies by the norm and scatter-adds.  Here nothing per-edge is materialised:
//
//   forward   one thread per (destination row, output column): walks the row's incoming edges in
//             the reference's edge order (dst-CSR), 128-bit packed edge record, the si inputs of
//             its block from the source row, the matching weight column from a layout in which
//             consecutive output columns are contiguous (coalesced across the warp).  Pure
//             gather + register accumulate: no atomics, deterministic.
//   dX        the same kernel over the src-CSR with the transposed layout.
//   dW        edges grouped by relation; one thread per weight element accumulates a run of
//             edges of the same relation in a register and flushes once per run.
#include "common.cuh"

static constexpr int kThreads = 256;

// ------------------------------------------------------------------------------------------
// weight layouts
// ------------------------------------------------------------------------------------------
__global__ void bdd_layouts_kernel(const float* __restrict__ w, int R, int B, int si, int so,
                                   float* __restrict__ w_fwd, float* __restrict__ w_bwd) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_rel = (long long)B * si * so;
  if (idx >= (long long)R * per_rel) return;
  int r = (int)(idx / per_rel);
  int rem = (int)(idx % per_rel);
  int b = rem / (si * so), i = (rem / so) % si, o = rem % so;
  float v = w[idx];
  w_fwd[((size_t)r * si + i) * (B * so) + b * so + o] = v;
  w_bwd[((size_t)r * so + o) * (B * si) + b * si + i] = v;
}

extern "C" int kg_bdd_weight_layouts(const float* weight, int num_etypes, int num_bases, int si, int so,
                                     float* w_fwd, float* w_bwd, void* stream) {
  KG_REQUIRE(num_etypes > 0 && num_bases > 0 && si > 0 && so > 0, "bdd layouts: bad sizes");
  long long n = (long long)num_etypes * num_bases * si * so;
  bdd_layouts_kernel<<<kg_div_up(n, kThreads), kThreads, 0, kg_stream(stream)>>>(
      weight, num_etypes, num_bases, si, so, w_fwd, w_bwd);
  KG_LAUNCH_OK();
  return KG_OK;
}

// ------------------------------------------------------------------------------------------
// gather-aggregate (forward over dst-CSR, dX over src-CSR)
//   out[row][b*FO + o] = sum_{e in row} norm_e * sum_{i<FI} feat[nbr_e][b*FI + i] * wl[etype_e][i][b*FO + o]
// ------------------------------------------------------------------------------------------
template <int FI_T>
__global__ void __launch_bounds__(kThreads)
bdd_gather_kernel(const float* __restrict__ feat, const int* __restrict__ ptr,
                  const int4* __restrict__ pack, const float* __restrict__ wl, int n_rows, int B,
                  int FI_rt, int FO, float* __restrict__ out) {
  const int FI = FI_T > 0 ? FI_T : FI_rt;
  const int width = B * FO, in_width = B * FI;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n_rows * width) return;
  const int row = (int)(t / width), j = (int)(t % width);
  const int xoff = (j / FO) * FI;
  const int e_end = __ldg(ptr + row + 1);
  float acc = 0.f;
#pragma unroll 2
  for (int e = __ldg(ptr + row); e < e_end; ++e) {
    const int4 p = __ldg(pack + e);                 // {neighbour, etype, norm bits, -}
    const float* xr = feat + (size_t)p.x * in_width + xoff;
    const float* wr = wl + (size_t)p.y * FI * width + j;
    float m = 0.f;
#pragma unroll
    for (int i = 0; i < FI; ++i) m = fmaf(__ldg(xr + i), __ldg(wr + (size_t)i * width), m);
    acc = fmaf(__int_as_float(p.z), m, acc);
  }
  out[t] = acc;
}

static int launch_gather(const float* feat, const int* ptr, const void* pack, const float* wl,
                         int n_rows, int B, int FI, int FO, float* out, cudaStream_t st) {
  long long total = (long long)n_rows * B * FO;
  if (total == 0) return KG_OK;
  int grid = kg_div_up(total, kThreads);
  const int4* pk = reinterpret_cast<const int4*>(pack);
  switch (FI) {
    case 5: bdd_gather_kernel<5><<<grid, kThreads, 0, st>>>(feat, ptr, pk, wl, n_rows, B, FI, FO, out); break;
    case 10: bdd_gather_kernel<10><<<grid, kThreads, 0, st>>>(feat, ptr, pk, wl, n_rows, B, FI, FO, out); break;
    default: bdd_gather_kernel<0><<<grid, kThreads, 0, st>>>(feat, ptr, pk, wl, n_rows, B, FI, FO, out); break;
  }
  KG_LAUNCH_OK();
  return KG_OK;
}

extern "C" int kg_bdd_aggregate_fwd(const float* x, const int32_t* row_ptr, const void* fwd_pack,
                                    const float* w_fwd, int n_dst, int num_bases, int si, int so,
                                    float* agg, void* stream) {
  KG_REQUIRE(n_dst >= 0 && num_bases > 0 && si > 0 && so > 0, "bdd fwd: bad sizes");
  return launch_gather(x, row_ptr, fwd_pack, w_fwd, n_dst, num_bases, si, so, agg, kg_stream(stream));
}

extern "C" int kg_bdd_aggregate_bwd_dx(const float* dagg, const int32_t* col_ptr, const void* bwd_pack,
                                       const float* w_bwd, int n_src, int num_bases, int si, int so,
                                       float* dx, void* stream) {
  KG_REQUIRE(n_src >= 0 && num_bases > 0 && si > 0 && so > 0, "bdd dx: bad sizes");
  return launch_gather(dagg, col_ptr, bwd_pack, w_bwd, n_src, num_bases, so, si, dx, kg_stream(stream));
}

// ------------------------------------------------------------------------------------------
// dW over relation-grouped edges
// ------------------------------------------------------------------------------------------
static constexpr int kDwChunk = 128;

__global__ void __launch_bounds__(kThreads)
bdd_dw_kernel(const float* __restrict__ x, const float* __restrict__ dagg,
              const int4* __restrict__ rel_pack, int E, int B, int si, int so,
              float* __restrict__ dW) {
  const int KW = B * si * so;
  const int kidx = blockIdx.y * blockDim.x + threadIdx.x;
  const bool active = kidx < KW;
  const int b = kidx / (si * so), rem = kidx % (si * so);
  const int xin = b * si + rem / so, yin = b * so + rem % so;
  const int in_w = B * si, out_w = B * so;
  const int e0 = blockIdx.x * kDwChunk, e1 = min(E, e0 + kDwChunk);
  float acc = 0.f;
  int cur = -1;
  for (int e = e0; e < e1; ++e) {
    const int4 p = __ldg(rel_pack + e);             // {src, dst, etype, norm bits}
    if (p.z != cur) {
      if (cur >= 0 && active) atomicAdd(dW + (size_t)cur * KW + kidx, acc);
      acc = 0.f;
      cur = p.z;
    }
    if (active)
      acc = fmaf(__int_as_float(p.w) * __ldg(x + (size_t)p.x * in_w + xin),
                 __ldg(dagg + (size_t)p.y * out_w + yin), acc);
  }
  if (cur >= 0 && active) atomicAdd(dW + (size_t)cur * KW + kidx, acc);
}

extern "C" int kg_bdd_aggregate_bwd_dw(const float* x, const float* dagg, const void* rel_pack,
                                       int n_edges, int num_bases, int si, int so, float* dweight,
                                       void* stream) {
  KG_REQUIRE(n_edges >= 0 && num_bases > 0 && si > 0 && so > 0, "bdd dw: bad sizes");
  if (n_edges == 0) return KG_OK;
  dim3 grid(kg_div_up(n_edges, kDwChunk), kg_div_up((long long)num_bases * si * so, kThreads));
  bdd_dw_kernel<<<grid, kThreads, 0, kg_stream(stream)>>>(
      x, dagg, reinterpret_cast<const int4*>(rel_pack), n_edges, num_bases, si, so, dweight);
  KG_LAUNCH_OK();
  return KG_OK;
}

// ------------------------------------------------------------------------------------------
// a2: embedding rows (reference kgvae/model.py:185-191)
// ------------------------------------------------------------------------------------------
__global__ void embedding_fwd_kernel(const float* __restrict__ table, const int* __restrict__ ids,
                                     int n, int dim, float* __restrict__ out) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * dim) return;
  int i = (int)(t / dim), d = (int)(t % dim);
  out[t] = __ldg(table + (size_t)__ldg(ids + i) * dim + d);
}

__global__ void embedding_bwd_kernel(const float* __restrict__ g, const int* __restrict__ ids, int n,
                                     int dim, float* __restrict__ grad_table) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * dim) return;
  int i = (int)(t / dim), d = (int)(t % dim);
  atomicAdd(grad_table + (size_t)__ldg(ids + i) * dim + d, g[t]);
}

extern "C" int kg_embedding_fwd(const float* table, const int32_t* ids, int n, int dim, float* out,
                                void* stream) {
  KG_REQUIRE(n >= 0 && dim > 0, "embedding fwd: bad sizes");
  if (n == 0) return KG_OK;
  embedding_fwd_kernel<<<kg_div_up((long long)n * dim, kThreads), kThreads, 0, kg_stream(stream)>>>(
      table, ids, n, dim, out);
  KG_LAUNCH_OK();
  return KG_OK;
}

extern "C" int kg_embedding_bwd(const float* grad_out, const int32_t* ids, int n, int dim,
                                float* grad_table, void* stream) {
  KG_REQUIRE(n >= 0 && dim > 0, "embedding bwd: bad sizes");
  if (n == 0) return KG_OK;
  embedding_bwd_kernel<<<kg_div_up((long long)n * dim, kThreads), kThreads, 0, kg_stream(stream)>>>(
      grad_out, ids, n, dim, grad_table);
  KG_LAUNCH_OK();
  return KG_OK;
}

// ------------------------------------------------------------------------------------------
// backward of  out = dropout(act(pre))   (tail of RelGraphConv.forward)
// ------------------------------------------------------------------------------------------
__global__ void act_dropout_bwd_kernel(const float* __restrict__ go, const float* __restrict__ out,
                                       const float* __restrict__ mask, int act, long long n,
                                       float* __restrict__ gp) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  float g = go[t];
  if (mask) g *= mask[t];
  // relu: with a positive keep-scale, out > 0 exactly when pre > 0 and the unit was kept
  if (act == 1 && !(out[t] > 0.f)) g = 0.f;
  gp[t] = g;
}

extern "C" int kg_act_dropout_bwd(const float* grad_out, const float* out, const float* drop_mask,
                                  int act, long long numel, float* grad_pre, void* stream) {
  KG_REQUIRE(numel >= 0 && (act == 0 || act == 1), "act_dropout_bwd: bad arguments");
  if (numel == 0) return KG_OK;
  act_dropout_bwd_kernel<<<kg_div_up(numel, kThreads), kThreads, 0, kg_stream(stream)>>>(
      grad_out, out, drop_mask, act, numel, grad_pre);
  KG_LAUNCH_OK();
  return KG_OK;
}

// ------------------------------------------------------------------------------------------
// column sums, two deterministic stages
// ------------------------------------------------------------------------------------------
static int colsum_chunks(int rows) {
  int c = kg_div_up(rows, 256);
  return c < 1 ? 1 : (c > 128 ? 128 : c);
}

__global__ void colsum_stage1(const float* __restrict__ x, int rows, int cols, int chunks,
                              float* __restrict__ partial) {
  __shared__ float red[8][33];
  const int col = blockIdx.x * 32 + threadIdx.x;
  const int per = (rows + chunks - 1) / chunks;
  const int r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  float s = 0.f;
  if (col < cols)
    for (int r = r0 + threadIdx.y; r < r1; r += 8) s += x[(size_t)r * cols + col];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && col < cols) {
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) tot += red[k][threadIdx.x];
    partial[(size_t)blockIdx.y * cols + col] = tot;
  }
}

__global__ void colsum_stage2(const float* __restrict__ partial, int cols, int chunks,
                              float* __restrict__ out) {
  int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= cols) return;
  float s = 0.f;
  for (int c = 0; c < chunks; ++c) s += partial[(size_t)c * cols + col];
  out[col] = s;
}

extern "C" size_t kg_colsum_workspace_bytes(int rows, int cols) {
  return kg_align_up((size_t)colsum_chunks(rows) * (cols > 0 ? cols : 1) * sizeof(float));
}

extern "C" int kg_colsum(const float* x, int rows, int cols, float* out, void* workspace,
                         size_t workspace_bytes, void* stream) {
  KG_REQUIRE(rows >= 0 && cols > 0, "colsum: bad sizes");
  cudaStream_t st = kg_stream(stream);
  if (rows == 0) {
    KG_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * cols, st));
    return KG_OK;
  }
  const int chunks = colsum_chunks(rows);
  if (workspace_bytes < (size_t)chunks * cols * sizeof(float))
    return kg_fail(KG_ERR_WORKSPACE, "colsum: workspace too small");
  float* partial = reinterpret_cast<float*>(workspace);
  colsum_stage1<<<dim3(kg_div_up(cols, 32), chunks), dim3(32, 8), 0, st>>>(x, rows, cols, chunks, partial);
  KG_LAUNCH_OK();
  colsum_stage2<<<kg_div_up(cols, 128), 128, 0, st>>>(partial, cols, chunks, out);
  KG_LAUNCH_OK();
  return KG_OK;
}
