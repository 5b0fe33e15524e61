// Small row-wise kernels around the RelGraphConv("bdd") layer (a3; the message passing itself is
// in rgcn_bdd_rel.cu): the two derived weight layouts, the embedding lookup (a2), the
// activation/dropout backward and deterministic column sums.
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

// the same, plus the bias gradient in the same pass: colsum[c] += sum_r grad_pre[r, c]  (colsum zero-filled
// here; a thread owns 4 consecutive columns, a warp 512 contiguous bytes of a row, a block 64 rows)
__global__ void __launch_bounds__(256)
act_dropout_bwd_colsum_kernel(const float* __restrict__ go, const float* __restrict__ out,
                              const float* __restrict__ mask, int act, int rows, int cols,
                              float* __restrict__ gp, float* __restrict__ colsum) {
  __shared__ float4 red[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + lane) * 4;
  const int r0 = blockIdx.y * 64, r1 = min(rows, r0 + 64);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < cols)
    for (int r = r0 + warp; r < r1; r += 8) {
      const size_t off = (size_t)r * cols + c;
      float4 g = *reinterpret_cast<const float4*>(go + off);
      if (mask) {
        const float4 m = *reinterpret_cast<const float4*>(mask + off);
        g.x *= m.x; g.y *= m.y; g.z *= m.z; g.w *= m.w;
      }
      if (act == 1) {
        const float4 o = *reinterpret_cast<const float4*>(out + off);
        if (!(o.x > 0.f)) g.x = 0.f;
        if (!(o.y > 0.f)) g.y = 0.f;
        if (!(o.z > 0.f)) g.z = 0.f;
        if (!(o.w > 0.f)) g.w = 0.f;
      }
      *reinterpret_cast<float4*>(gp + off) = g;
      acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
    }
  red[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && c < cols) {
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      const float4 t = red[w][lane];
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    atomicAdd(colsum + c, acc.x);
    atomicAdd(colsum + c + 1, acc.y);
    atomicAdd(colsum + c + 2, acc.z);
    atomicAdd(colsum + c + 3, acc.w);
  }
}

// grad_pre and the column sums of grad_pre (the bias gradient) in one pass; cols % 4 == 0 and 16-byte
// aligned tensors take the fused kernel, anything else the two separate ones
extern "C" int kg_act_dropout_bwd_colsum(const float* grad_out, const float* out, const float* drop_mask, int act,
                                         int rows, int cols, float* grad_pre, float* colsum, void* workspace,
                                         size_t workspace_bytes, void* stream) {
  KG_REQUIRE(rows >= 0 && cols > 0 && (act == 0 || act == 1), "act_dropout_bwd_colsum: bad arguments");
  cudaStream_t st = kg_stream(stream);
  const uintptr_t bits = reinterpret_cast<uintptr_t>(grad_out) | reinterpret_cast<uintptr_t>(out) |
                         reinterpret_cast<uintptr_t>(drop_mask) | reinterpret_cast<uintptr_t>(grad_pre);
  if (rows > 0 && cols % 4 == 0 && (bits & 15) == 0) {
    KG_CUDA(cudaMemsetAsync(colsum, 0, sizeof(float) * cols, st));
    act_dropout_bwd_colsum_kernel<<<dim3(kg_div_up(cols, 128), kg_div_up(rows, 64)), 256, 0, st>>>(
        grad_out, out, drop_mask, act, rows, cols, grad_pre, colsum);
    KG_LAUNCH_OK();
    return KG_OK;
  }
  int rc = kg_act_dropout_bwd(grad_out, out, drop_mask, act, (long long)rows * cols, grad_pre, stream);
  if (rc != KG_OK) return rc;
  return kg_colsum(grad_pre, rows, cols, colsum, workspace, workspace_bytes, stream);
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

// narrow matrices (cols <= 32, e.g. the 10 / 11 columns of the entity-classification layers):
// consecutive threads walk the matrix as one flat array (coalesced), thread t owns column t % cols
__global__ void __launch_bounds__(256)
colsum_narrow(const float* __restrict__ x, int rows, int cols, float* __restrict__ out) {
  __shared__ float red[256];
  const int tpr = 256 / cols;                            // rows per block pass
  const int rig = threadIdx.x / cols, col = threadIdx.x - rig * cols;
  float s = 0.f;
  if (rig < tpr)
    for (long long r = (long long)blockIdx.x * tpr + rig; r < rows; r += (long long)gridDim.x * tpr)
      s += x[(size_t)r * cols + col];
  red[threadIdx.x] = rig < tpr ? s : 0.f;
  __syncthreads();
  if (threadIdx.x < cols) {
    float tot = 0.f;
    for (int g = 0; g < tpr; ++g) tot += red[g * cols + threadIdx.x];
    atomicAdd(out + threadIdx.x, tot);
  }
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
  if (cols <= 32 && rows >= 4096) {
    KG_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * cols, st));
    const int want = kg_div_up(rows, (256 / cols) * 16), cap = 8 * kg_sm_count();
    colsum_narrow<<<want < cap ? want : cap, 256, 0, st>>>(x, rows, cols, out);
    KG_LAUNCH_OK();
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
