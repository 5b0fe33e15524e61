// a9: DistMult decoder, BCE-with-logits and the L2 regulariser, forward and backward.
// Replaces LinkPredict.calc_score / get_loss / regularization_loss (reference
// kgvae/link_predict.py:57-78).  The reference gathers three [S, h] tensors, materialises two
// [S, h] products and, in backward, scatter-adds with index_put(accumulate) into z.
// Here: the score is one warp-level pass over three gathered rows; the backward into z is a
// gather-reduce over an entity-major index of the triplets (no atomics, deterministic), the
// backward into w_relation runs over relation-grouped triplets.
#include <cub/cub.cuh>

#include "common.cuh"

static constexpr int kThreads = 256;

// ------------------------------------------------------------------------------------------
// score_i = sum_d z[s,d] w[r,d] z[o,d] + shift          (link_predict.py:57-63, :75-76)
// ------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(kThreads)
distmult_score_kernel(const float* __restrict__ z, const float* __restrict__ w,
                      const int* __restrict__ trip, int S, int h, const float* __restrict__ shift,
                      float* __restrict__ score) {
  const int t = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= S) return;
  const int s = __ldg(trip + 3 * (size_t)t), r = __ldg(trip + 3 * (size_t)t + 1),
            o = __ldg(trip + 3 * (size_t)t + 2);
  float acc = 0.f;
  if (VEC == 4) {
    const float4* zs = reinterpret_cast<const float4*>(z + (size_t)s * h);
    const float4* wr = reinterpret_cast<const float4*>(w + (size_t)r * h);
    const float4* zo = reinterpret_cast<const float4*>(z + (size_t)o * h);
    for (int c = lane; c < h / 4; c += 32) {
      const float4 a = __ldg(zs + c), b = __ldg(wr + c), d = __ldg(zo + c);
      acc = fmaf(a.x * b.x, d.x, acc);
      acc = fmaf(a.y * b.y, d.y, acc);
      acc = fmaf(a.z * b.z, d.z, acc);
      acc = fmaf(a.w * b.w, d.w, acc);
    }
  } else {
    for (int c = lane; c < h; c += 32)
      acc = fmaf(__ldg(z + (size_t)s * h + c) * __ldg(w + (size_t)r * h + c), __ldg(z + (size_t)o * h + c), acc);
  }
  acc = kg_warp_sum(acc);
  if (lane == 0) score[t] = shift ? acc + __ldg(shift) : acc;
}

extern "C" int kg_distmult_score(const float* z, const float* w, const int32_t* triplets,
                                 int n_triplets, int h, const float* shift, float* score, void* stream) {
  KG_REQUIRE(n_triplets >= 0 && h > 0, "distmult score: bad sizes");
  if (n_triplets == 0) return KG_OK;
  const int grid = kg_div_up((long long)n_triplets * 32, kThreads);
  const bool vec = (h % 4 == 0) && ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(w)) & 15) == 0;
  if (vec) distmult_score_kernel<4><<<grid, kThreads, 0, kg_stream(stream)>>>(z, w, triplets, n_triplets, h, shift, score);
  else distmult_score_kernel<1><<<grid, kThreads, 0, kg_stream(stream)>>>(z, w, triplets, n_triplets, h, shift, score);
  KG_LAUNCH_OK();
  return KG_OK;
}

// ------------------------------------------------------------------------------------------
// deterministic two-stage reductions
// ------------------------------------------------------------------------------------------
static constexpr int kMaxPartials = 1024;

static int reduce_blocks(long long n) {
  long long b = (n + 4 * kThreads - 1) / (4 * kThreads);
  return (int)(b < 1 ? 1 : (b > kMaxPartials ? kMaxPartials : b));
}

__device__ __forceinline__ float block_sum(float v) {
  __shared__ float red[kThreads / 32];
  v = kg_warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float tot = 0.f;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) tot += red[i];
  }
  return tot;   // valid on thread 0
}

// mode 0: sum x; 1: sum x^2; 2: BCE-with-logits(x, y) and dscore = (sigmoid(x) - y) * inv_n
__global__ void __launch_bounds__(kThreads)
reduce_stage1(const float* __restrict__ x, const float* __restrict__ y, long long n, int mode,
              float inv_n, float* __restrict__ dscore, float* __restrict__ partial) {
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    if (mode == 0) s += v;
    else if (mode == 1) s = fmaf(v, v, s);
    else {
      const float lab = y[i];
      s += fmaxf(v, 0.f) - v * lab + log1pf(expf(-fabsf(v)));   // F.binary_cross_entropy_with_logits
      if (dscore) dscore[i] = (kg_sigmoid(v) - lab) * inv_n;
    }
  }
  const float tot = block_sum(s);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

__global__ void reduce_stage2(const float* __restrict__ partial, int count, float scale,
                              float* __restrict__ out) {
  // one warp, fixed order: lane-strided partial sums then a shuffle tree
  float s = 0.f;
  for (int i = threadIdx.x; i < count; i += 32) s += partial[i];
  s = kg_warp_sum(s);
  if (threadIdx.x == 0) out[0] = s * scale;
}

extern "C" size_t kg_reduce_workspace_bytes(long long n) {
  (void)n;
  return kg_align_up(sizeof(float) * kMaxPartials);
}

static int run_reduce(const float* x, const float* y, long long n, int mode, float scale, float inv_n,
                      float* dscore, float* out, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (workspace_bytes < sizeof(float) * kMaxPartials)
    return kg_fail(KG_ERR_WORKSPACE, "reduce: workspace too small");
  if (n == 0) {
    KG_CUDA(cudaMemsetAsync(out, 0, sizeof(float), st));
    return KG_OK;
  }
  const int blocks = reduce_blocks(n);
  float* partial = reinterpret_cast<float*>(workspace);
  reduce_stage1<<<blocks, kThreads, 0, st>>>(x, y, n, mode, inv_n, dscore, partial);
  KG_LAUNCH_OK();
  reduce_stage2<<<1, 32, 0, st>>>(partial, blocks, scale, out);
  KG_LAUNCH_OK();
  return KG_OK;
}

extern "C" int kg_bce_logits_fwd(const float* score, const float* labels, int n, float* loss_out,
                                 float* dscore, void* workspace, size_t workspace_bytes, void* stream) {
  KG_REQUIRE(n >= 0, "bce: bad size");
  const float inv = n > 0 ? 1.0f / (float)n : 0.f;
  return run_reduce(score, labels, n, 2, inv, inv, dscore, loss_out, workspace, workspace_bytes, kg_stream(stream));
}

extern "C" int kg_sum_squares(const float* x, long long n, float* out, void* workspace,
                              size_t workspace_bytes, void* stream) {
  KG_REQUIRE(n >= 0, "sum_squares: bad size");
  return run_reduce(x, nullptr, n, 1, 1.f, 0.f, nullptr, out, workspace, workspace_bytes, kg_stream(stream));
}

extern "C" int kg_sum(const float* x, long long n, float* out, void* workspace, size_t workspace_bytes,
                      void* stream) {
  KG_REQUIRE(n >= 0, "sum: bad size");
  return run_reduce(x, nullptr, n, 0, 1.f, 0.f, nullptr, out, workspace, workspace_bytes, kg_stream(stream));
}

// ------------------------------------------------------------------------------------------
// triplet index for the backward pass
// ------------------------------------------------------------------------------------------
__global__ void triplet_keys(const int* __restrict__ trip, int S, int* ent_key, int* ent_val,
                             int* rel_key, int* rel_val, int* ent_cnt, int* rel_cnt) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= S) return;
  const int s = trip[3 * (size_t)t], r = trip[3 * (size_t)t + 1], o = trip[3 * (size_t)t + 2];
  ent_key[t] = s;      ent_val[t] = t;         // subject side
  ent_key[S + t] = o;  ent_val[S + t] = S + t; // object side
  rel_key[t] = r;      rel_val[t] = t;
  atomicAdd(ent_cnt + s, 1);
  atomicAdd(ent_cnt + o, 1);
  atomicAdd(rel_cnt + r, 1);
}

__global__ void fill_ent_pack(const int* __restrict__ trip, const int* __restrict__ sorted_val, int S,
                              int4* __restrict__ pack) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= 2 * S) return;
  const int v = sorted_val[k];
  const int t = v < S ? v : v - S;
  const int other = v < S ? trip[3 * (size_t)t + 2] : trip[3 * (size_t)t];
  pack[k] = make_int4(other, trip[3 * (size_t)t + 1], t, 0);
}

static int bits_for(int n) {
  int b = 1;
  while (b < 31 && (1ll << b) < (long long)n) ++b;
  return b;
}

static size_t triplet_cub_bytes(int S) {
  size_t a = 0, c = 0;
  cub::DeviceRadixSort::SortPairs((void*)nullptr, a, (int*)nullptr, (int*)nullptr, (int*)nullptr,
                                  (int*)nullptr, 2 * S);
  cub::DeviceScan::ExclusiveSum((void*)nullptr, c, (int*)nullptr, (int*)nullptr, (1 << 24) + 1);
  return kg_align_up((a > c ? a : c) + 256);
}

extern "C" size_t kg_triplet_index_workspace_bytes(int n_triplets) {
  int S = n_triplets > 0 ? n_triplets : 1;
  return 4 * kg_align_up(((size_t)2 * S + 1) * 4) + 2 * kg_align_up(((size_t)S + 1) * 4) + triplet_cub_bytes(S) + 1024;
}

extern "C" int kg_triplet_index(const int32_t* triplets, int n_triplets, int n_nodes, int n_rels,
                                int32_t* ent_ptr, void* ent_pack, int32_t* rel_ptr, int32_t* rel_perm,
                                void* workspace, size_t workspace_bytes, void* stream) {
  KG_REQUIRE(n_triplets >= 0 && n_nodes > 0 && n_rels > 0, "triplet index: bad sizes");
  KG_REQUIRE(n_nodes < (1 << 24), "triplet index: n_nodes < 2^24");
  cudaStream_t st = kg_stream(stream);
  const int S = n_triplets;
  KG_CUDA(cudaMemsetAsync(ent_ptr, 0, sizeof(int) * (n_nodes + 1), st));
  KG_CUDA(cudaMemsetAsync(rel_ptr, 0, sizeof(int) * (n_rels + 1), st));
  KgArena ws(workspace, workspace_bytes);
  int* ek_in = ws.take<int>(2 * (size_t)S + 1);
  int* ek_out = ws.take<int>(2 * (size_t)S + 1);
  int* ev_in = ws.take<int>(2 * (size_t)S + 1);
  int* ev_out = ws.take<int>(2 * (size_t)S + 1);
  int* rk_in = ws.take<int>((size_t)S + 1);
  int* rv_in = ws.take<int>((size_t)S + 1);
  size_t temp_bytes = triplet_cub_bytes(S > 0 ? S : 1);
  void* temp = ws.take<char>(temp_bytes);
  if (!ek_in || !ek_out || !ev_in || !ev_out || !rk_in || !rv_in || !temp)
    return kg_fail(KG_ERR_WORKSPACE, "triplet index: workspace too small");
  if (S > 0) {
    triplet_keys<<<kg_div_up(S, kThreads), kThreads, 0, st>>>(triplets, S, ek_in, ev_in, rk_in, rv_in, ent_ptr, rel_ptr);
    KG_LAUNCH_OK();
  }
  size_t tb = temp_bytes;
  KG_CUDA(cub::DeviceScan::ExclusiveSum(temp, tb, ent_ptr, ent_ptr, n_nodes + 1, st));
  tb = temp_bytes;
  KG_CUDA(cub::DeviceScan::ExclusiveSum(temp, tb, rel_ptr, rel_ptr, n_rels + 1, st));
  if (S == 0) return KG_OK;
  tb = temp_bytes;
  KG_CUDA(cub::DeviceRadixSort::SortPairs(temp, tb, ek_in, ek_out, ev_in, ev_out, 2 * S, 0, bits_for(n_nodes), st));
  fill_ent_pack<<<kg_div_up(2LL * S, kThreads), kThreads, 0, st>>>(triplets, ev_out, S, reinterpret_cast<int4*>(ent_pack));
  KG_LAUNCH_OK();
  tb = temp_bytes;
  // relation grouping: reuse ek_out as the (unused) sorted-key output
  KG_CUDA(cub::DeviceRadixSort::SortPairs(temp, tb, rk_in, ek_out, rv_in, rel_perm, S, 0, bits_for(n_rels), st));
  return KG_OK;
}

// ------------------------------------------------------------------------------------------
// dz[v, :] = sum over the entity-major index of gscore[t] * w[r, :] * z[other, :]
// one thread per (entity, VEC consecutive columns); no atomics
// ------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(kThreads)
distmult_dz_kernel(const float* __restrict__ z, const float* __restrict__ w,
                   const float* __restrict__ gscore, const int* __restrict__ ent_ptr,
                   const int4* __restrict__ ent_pack, int n_nodes, int h, float* __restrict__ dz) {
  const int cols = h / VEC;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n_nodes * cols) return;
  const int v = (int)(t / cols), c = (int)(t % cols);
  const int e_end = __ldg(ent_ptr + v + 1);
  float acc[VEC];
#pragma unroll
  for (int q = 0; q < VEC; ++q) acc[q] = 0.f;
#pragma unroll 2
  for (int e = __ldg(ent_ptr + v); e < e_end; ++e) {
    const int4 p = __ldg(ent_pack + e);          // {other, rel, triplet, -}
    const float g = __ldg(gscore + p.z);
    if (VEC == 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(w + (size_t)p.y * h) + c);
      const float4 b = __ldg(reinterpret_cast<const float4*>(z + (size_t)p.x * h) + c);
      acc[0] = fmaf(g * a.x, b.x, acc[0]);
      acc[1 % VEC] = fmaf(g * a.y, b.y, acc[1 % VEC]);
      acc[2 % VEC] = fmaf(g * a.z, b.z, acc[2 % VEC]);
      acc[3 % VEC] = fmaf(g * a.w, b.w, acc[3 % VEC]);
    } else {
      acc[0] = fmaf(g * __ldg(w + (size_t)p.y * h + c), __ldg(z + (size_t)p.x * h + c), acc[0]);
    }
  }
  if (VEC == 4) {
    reinterpret_cast<float4*>(dz + (size_t)v * h)[c] = make_float4(acc[0], acc[1 % VEC], acc[2 % VEC], acc[3 % VEC]);
  } else {
    dz[(size_t)v * h + c] = acc[0];
  }
}

extern "C" int kg_distmult_bwd_dz(const float* z, const float* w, const float* gscore,
                                  const int32_t* ent_ptr, const void* ent_pack, int n_nodes, int h,
                                  float* dz, void* stream) {
  KG_REQUIRE(n_nodes >= 0 && h > 0, "distmult dz: bad sizes");
  if (n_nodes == 0) return KG_OK;
  const bool vec = (h % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(dz)) & 15) == 0;
  const int4* pk = reinterpret_cast<const int4*>(ent_pack);
  if (vec) {
    distmult_dz_kernel<4><<<kg_div_up((long long)n_nodes * (h / 4), kThreads), kThreads, 0, kg_stream(stream)>>>(
        z, w, gscore, ent_ptr, pk, n_nodes, h, dz);
  } else {
    distmult_dz_kernel<1><<<kg_div_up((long long)n_nodes * h, kThreads), kThreads, 0, kg_stream(stream)>>>(
        z, w, gscore, ent_ptr, pk, n_nodes, h, dz);
  }
  KG_LAUNCH_OK();
  return KG_OK;
}

// ------------------------------------------------------------------------------------------
// dw[r, :] += sum_{t in rel group r} gscore[t] * z[s_t, :] * z[o_t, :]
// CTA = chunk of relation-grouped triplets x 128 columns; a run of equal relations is
// accumulated in a register and flushed once
// ------------------------------------------------------------------------------------------
static constexpr int kDwChunk = 256;

__global__ void __launch_bounds__(128)
distmult_dw_kernel(const float* __restrict__ z, const float* __restrict__ gscore,
                   const int* __restrict__ trip, const int* __restrict__ rel_perm, int S, int h,
                   float* __restrict__ dw) {
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  const bool active = c < h;
  const int k0 = blockIdx.x * kDwChunk, k1 = min(S, k0 + kDwChunk);
  float acc = 0.f;
  int cur = -1;
  for (int k = k0; k < k1; ++k) {
    const int t = __ldg(rel_perm + k);
    const int s = __ldg(trip + 3 * (size_t)t), r = __ldg(trip + 3 * (size_t)t + 1),
              o = __ldg(trip + 3 * (size_t)t + 2);
    if (r != cur) {
      if (cur >= 0 && active) atomicAdd(dw + (size_t)cur * h + c, acc);
      acc = 0.f;
      cur = r;
    }
    if (active)
      acc = fmaf(__ldg(gscore + t) * __ldg(z + (size_t)s * h + c), __ldg(z + (size_t)o * h + c), acc);
  }
  if (cur >= 0 && active) atomicAdd(dw + (size_t)cur * h + c, acc);
}

extern "C" int kg_distmult_bwd_dw(const float* z, const float* gscore, const int32_t* triplets,
                                  const int32_t* rel_perm, int n_triplets, int h, float* dw, void* stream) {
  KG_REQUIRE(n_triplets >= 0 && h > 0, "distmult dw: bad sizes");
  if (n_triplets == 0) return KG_OK;
  dim3 grid(kg_div_up(n_triplets, kDwChunk), kg_div_up(h, 128));
  distmult_dw_kernel<<<grid, 128, 0, kg_stream(stream)>>>(z, gscore, triplets, rel_perm, n_triplets, h, dw);
  KG_LAUNCH_OK();
  return KG_OK;
}
