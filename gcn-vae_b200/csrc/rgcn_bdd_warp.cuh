// a3: RelGraphConv(regularizer="bdd") message passing, WARP-AUTONOMOUS block-owner kernels (the
// reference model's 5x5 and 5x10 blocks; DGL RelGraphConv as constructed at kgvae/model.py:54-59).
//
// Same register tiling as rgcn_bdd_own.cuh - a lane owns TB whole diagonal blocks of W_r (100
// floats) for a run of edges - but every WARP runs its own pipeline: it gathers only the slice of
// the row its lanes consume (one bulk copy per edge into its own mbarrier ring), computes, and
// issues its own bulk reduction.  ncu on the CTA-synchronised version showed the warps of a slot
// waiting for each other at the named barrier (1.97 stall cycles per issue in the fused backward,
// 0.51 in the forward) with 3 warps per scheduler; here there is no barrier in the edge loop at all
// (only __syncwarp), the two halves of an edge drift apart freely, and idle lanes shadow a valid
// owner instead of branching around the math.
//
//   forward   warp = (edge share, half of the blocks): x[src] slice -> out[dst] slice
//   backward  input-gradient warps: dagg[dst] slice -> dx[src] slice (bulk reduce-add)
//             weight-gradient warps: x[src] slice + dagg[dst] slice -> dW_r in registers
//             Two variants: fully independent warps (every matrix L2-resident: FB15k-237 / WN18
//             shapes; fastest there, 0.39 vs 0.49 ms) and PAIRED warps that share one ring so that
//             dagg[dst] crosses HBM once per edge (streaming regime, wikikg2 shape: the independent
//             warps drift apart and fetched it twice - 148 GB instead of 103 GB of DRAM traffic).
#pragma once
#include "rgcn_bdd_own.cuh"

namespace bddwarp {

using namespace bddown;

constexpr int kWarps = 4;

// 16-byte aligned superset [start, start + bytes/4) of the floats [first, first + n) of a row
struct Span {
  int start;        // float offset of the copy inside the row (multiple of 4)
  uint32_t bytes;   // copy size (multiple of 16)
  int skip;         // floats between `start` and `first`
};
__host__ __device__ __forceinline__ Span make_span(int first, int n) {
  const int lo = first & ~3, hi = (first + n + 3) & ~3;
  return Span{lo, (uint32_t)(hi - lo) * 4u, first - lo};
}
__host__ __device__ __forceinline__ int round4(int v) { return (v + 3) & ~3; }

// a lane's WN weights from global memory: 128-bit loads when WN * 4 is a multiple of 16 (every lane's slice is then
// 16-byte aligned), 64-bit loads otherwise (5x5 blocks, two per lane: 50 floats)
template <int WN>
__device__ __forceinline__ void ldg_weights(float (&w)[WN], const float* __restrict__ p) {
  if (WN % 4 == 0) {
    const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int i = 0; i < WN / 4; ++i) {
      const float4 t = __ldg(q + i);
      w[4 * i] = t.x; w[4 * i + 1] = t.y; w[4 * i + 2] = t.z; w[4 * i + 3] = t.w;
    }
  } else {
    const float2* q = reinterpret_cast<const float2*>(p);
#pragma unroll
    for (int i = 0; i < WN / 2; ++i) {
      const float2 t = __ldg(q + i);
      w[2 * i] = t.x; w[2 * i + 1] = t.y;
    }
  }
}

// dst[0..WN) += r[0..WN) with vector reductions, then r = 0 (owners only reduce; everyone clears)
template <int WN>
__device__ __forceinline__ void red_flush(float* dst, float (&r)[WN], bool owner) {
  if (WN % 4 == 0) {
#pragma unroll
    for (int i = 0; i < WN; i += 4) {
      if (owner) red_add_v4(dst + i, r[i], r[i + 1], r[i + 2], r[i + 3]);
      r[i] = r[i + 1] = r[i + 2] = r[i + 3] = 0.f;
    }
  } else {
#pragma unroll
    for (int i = 0; i < WN; i += 2) {
      if (owner) red_add_v2(dst + i, r[i], r[i + 1]);
      r[i] = r[i + 1] = 0.f;
    }
  }
}

template <int N>
__device__ __forceinline__ void park_raw(float* mine, const float (&v)[N]) {
  if (N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N; i += 4) *reinterpret_cast<float4*>(mine + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < N; i += 2) *reinterpret_cast<float2*>(mine + i) = make_float2(v[i], v[i + 1]);
  }
}

// ------------------------------------------------------------------------------------------
// forward: out[dst] += norm * blockdiag(W_etype) feat[src]
// ------------------------------------------------------------------------------------------
template <int FI, int FO, int TB, int WPS>
struct FwdLayout {
  static constexpr int XN = TB * FI, CN = TB * FO;
  int lpw, stage_f;
  __host__ __device__ FwdLayout(int B) {
    lpw = lanes_per_warp(B / TB, WPS, CN);
    stage_f = round4(lpw * XN) + 4;
  }
  __host__ __device__ int per_warp(int depth, int nt) const { return depth * stage_f + nt * 32 * CN; }
  __host__ size_t smem(int depth, int nt) const {
    return sizeof(float) * ((size_t)4 * kChunk + (size_t)kWarps * per_warp(depth, nt)) + sizeof(uint64_t) * kWarps * depth;
  }
};

// NT = parked-output buffers per warp = bulk reductions a warp keeps in flight.  ncu on the 5x10 layer showed the
// L1 -> XBAR path 80 % busy at 5.3 TB/s with two buffers: with 12 warps per SM that is 48 KB of reductions in
// flight, i.e. the launch was bound by the LATENCY of a bulk reduction, not by the path; four buffers (and a
// shallower gather ring to stay at three CTAs per SM) double the bytes in flight.
template <int FI, int FO, int TB, int WPS, int DEPTH, int NT, int MINB>
__global__ void __launch_bounds__(kCta, MINB)
fwd_kernel(RowSource feat, const int4* __restrict__ pack, int E, const float* __restrict__ weight, int B, int ldw,
           int ldo, int hints, float* __restrict__ out) {
  // B = blocks handled by this launch; ldw / ldo = row strides (floats) of weight / out.  A launch over a COLUMN
  // CHUNK of the layer (blocks [b0, b0 + B) of B_total) gets weight + b0 * FI * FO, out + b0 * FO, the chunk's own
  // compact x matrix, ldw = B_total * FI * FO and ldo = B_total * FO: block-diagonal weights make chunks independent.
  constexpr int XN = TB * FI, CN = TB * FO, WN = TB * FI * FO, SLOTS = kWarps / WPS;
  extern __shared__ __align__(16) float sm[];
  const FwdLayout<FI, FO, TB, WPS> L(B);
  const int width = ldo, in_w = B * FI, per = B / TB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = warp / WPS, wsl = warp % WPS;
  int4* P_s = reinterpret_cast<int4*>(sm);                     // [kChunk] {src, dst, etype, norm}
  float* ring = sm + 4 * kChunk + warp * L.per_warp(DEPTH, NT);   // [DEPTH][stage_f] this warp's slices
  float* tbuf = ring + DEPTH * L.stage_f;                         // [NT][32 * CN] parked outputs
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + 4 * kChunk + kWarps * L.per_warp(DEPTH, NT)) + warp * DEPTH;
  const int e0 = blockIdx.x * kChunk, n = min(E - e0, kChunk);
  for (int i = threadIdx.x; i < n; i += kCta) P_s[i] = __ldg(pack + e0 + i);
  if (lane == 0) {
    for (int i = 0; i < DEPTH; ++i) mbar_init(full + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();                                             // the only CTA-wide barrier
  const int share = (n + SLOTS - 1) / SLOTS;
  const int k_lo = slot * share, n_my = max(0, min(n - k_lo, share));
  const int g_lo = wsl * L.lpw, cnt = max(0, min(L.lpw, per - g_lo));
  if (n_my == 0 || cnt == 0) return;
  const int4* rec = P_s + k_lo;
  const Span xs = make_span(g_lo * XN, cnt * XN);
  const uint64_t pol = l2_policy((hints & 1) && feat.parts == nullptr), pol_red = l2_policy_last(hints & 4);
  auto gather = [&](int k, int st) {                           // lane 0
    mbar_expect_tx(full + st, xs.bytes);
    bulk_g2s(ring + st * L.stage_f, feat.row(rec[k].x, in_w) + xs.start, xs.bytes, full + st, pol);
  };
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < DEPTH; ++k)
      if (k < n_my) gather(k, k);
  }
  const int gl = min(lane, cnt - 1);          // idle lanes shadow the last owner: parked, never sent
  const float* xg = ring + xs.skip + gl * XN;
  const float* w_g = weight + (size_t)(g_lo + gl) * WN;
  float* out_w = out + g_lo * CN;
  const uint32_t warp_bytes = (uint32_t)cnt * CN * 4;
  float w[WN];
  int cur = -1;
  _Pragma("unroll 1") for (int k = 0; k < n_my; ++k) {
    const int st = k % DEPTH;
    mbar_wait(full + st, (k / DEPTH) & 1);                     // my slice of row k has landed
    const int4 p = rec[k];
    if (p.z != cur) {                                          // relation run starts (warp-uniform)
      cur = p.z;
      ldg_weights<WN>(w, row_at(w_g, cur, ldw));
    }
    float xv[XN], m[CN];
    lds_vec<XN>(xv, xg + st * L.stage_f);
    const float nv = __int_as_float(p.w);
#pragma unroll
    for (int i = 0; i < XN; ++i) xv[i] *= nv;
#pragma unroll
    for (int tb = 0; tb < TB; ++tb)
#pragma unroll
      for (int o = 0; o < FO; ++o) {
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < FI; ++i) a = fmaf(xv[tb * FI + i], w[(tb * FI + i) * FO + o], a);
        m[tb * FO + o] = a;
      }
    float* tb_cur = tbuf + (k % NT) * 32 * CN;
    bulk_wait_read<NT - 1>();                 // (lane 0 owns the groups) the reduction of edge k-NT has read this buffer
    __syncwarp();
    park_raw<CN>(tb_cur + lane * CN, m);
    fence_async_smem();
    __syncwarp();                             // outputs parked; every lane has consumed ring stage st
    if (lane == 0) {
      bulk_red_add(const_cast<float*>(row_at(out_w, p.y, width)), tb_cur, warp_bytes, pol_red);
      bulk_commit();
      if (k + DEPTH < n_my) gather(k + DEPTH, st);
    }
  }
  if (lane == 0) bulk_wait_all();             // shared memory must outlive the reductions reading it
}

// ------------------------------------------------------------------------------------------
// fused backward: dx[src] += norm * blockdiag(W_r)^T dagg[dst]
//                 dW[r][b][i][o] += norm * x[src][b*SI+i] * dagg[dst][b*SO+o]
// slot = WPR input-gradient warps then WPR weight-gradient warps
// ------------------------------------------------------------------------------------------
template <int SI, int SO, int TB, int WPR>
struct BwdLayout {
  static constexpr int XN = TB * SI, DN = TB * SO, WPS = 2 * WPR, SLOTS = kWarps / WPS;
  int lpw_x, lpw_w, stage_x, stage_w, xoff_w;
  __host__ __device__ BwdLayout(int B) {
    lpw_x = lanes_per_warp(B / TB, WPR, XN);
    lpw_w = lanes_per_warp(B / TB, WPR, 4);
    stage_x = round4(lpw_x * DN) + 4;                          // dagg slice
    xoff_w = round4(lpw_w * DN) + 4;                           // dagg slice, then x slice
    stage_w = xoff_w + round4(lpw_w * XN) + 4;
  }
  __host__ __device__ int x_warp(int depth, int nt) const { return depth * stage_x + nt * 32 * XN; }
  __host__ __device__ int w_warp(int depth) const { return depth * stage_w; }
  __host__ __device__ int total_f(int depth, int nt) const { return SLOTS * WPR * (x_warp(depth, nt) + w_warp(depth)); }
  __host__ size_t smem(int depth, int nt) const {
    return sizeof(float) * ((size_t)4 * kChunk + (size_t)total_f(depth, nt)) + sizeof(uint64_t) * kWarps * depth;
  }
};

template <int SI, int SO, int TB, int WPR, int DEPTH, int NT>
__global__ void __launch_bounds__(kCta)
bwd_kernel(RowSource x, const float* __restrict__ dagg, const int4* __restrict__ pack, int E,
           const float* __restrict__ weight, int B, int ldw, int ldd, int hints, float* __restrict__ dx,
           float* __restrict__ dW) {
  constexpr int XN = TB * SI, DN = TB * SO, WN = TB * SI * SO, WPS = 2 * WPR, SLOTS = kWarps / WPS;
  extern __shared__ __align__(16) float sm[];
  const BwdLayout<SI, SO, TB, WPR> L(B);
  const int in_w = B * SI, out_w = ldd, per = B / TB;          // out_w: row stride of dagg (see fwd_kernel on column chunks)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = warp / WPS, wsl = warp % WPS;
  const bool xrole = wsl < WPR;               // warp-uniform: input-gradient warps come first
  const int wr_i = xrole ? wsl : wsl - WPR;   // warp index inside its role
  const int nx = slot * WPR + (xrole ? wsl : WPR), nw = slot * WPR + (xrole ? 0 : wsl - WPR);   // warps of each role before me
  int4* P_s = reinterpret_cast<int4*>(sm);
  float* ring = sm + 4 * kChunk + nx * L.x_warp(DEPTH, NT) + nw * L.w_warp(DEPTH);
  const int stage_f = xrole ? L.stage_x : L.stage_w;
  float* tbuf = ring + DEPTH * stage_f;       // input-gradient warps only: [NT][32 * XN]
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + 4 * kChunk + L.total_f(DEPTH, NT)) + warp * DEPTH;
  const int e0 = blockIdx.x * kChunk, n = min(E - e0, kChunk);
  for (int i = threadIdx.x; i < n; i += kCta) P_s[i] = __ldg(pack + e0 + i);
  if (lane == 0) {
    for (int i = 0; i < DEPTH; ++i) mbar_init(full + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();                            // the only CTA-wide barrier
  const int share = (n + SLOTS - 1) / SLOTS;
  const int k_lo = slot * share, n_my = max(0, min(n - k_lo, share));
  const int lpw = xrole ? L.lpw_x : L.lpw_w;
  const int g_lo = wr_i * lpw, cnt = max(0, min(lpw, per - g_lo));
  if (n_my == 0 || cnt == 0 || (xrole && dx == nullptr)) return;
  const int4* rec = P_s + k_lo;
  const Span ds = make_span(g_lo * DN, cnt * DN), xsp = make_span(g_lo * XN, cnt * XN);
  const uint64_t pol_x = x.parts != nullptr ? l2_policy(false) : (hints & 4) ? l2_policy_last(true) : l2_policy(hints & 1);
  const uint64_t pol_d = l2_policy(hints & 2);
  const uint64_t pol_red = l2_policy_last(hints & 4);
  auto gather = [&](int k, int st) {          // lane 0
    const int4 p = rec[k];
    mbar_expect_tx(full + st, ds.bytes + (xrole ? 0u : xsp.bytes));
    bulk_g2s(ring + st * stage_f, row_at(dagg, p.y, out_w) + ds.start, ds.bytes, full + st, pol_d);
    if (!xrole) bulk_g2s(ring + st * stage_f + L.xoff_w, x.row(p.x, in_w) + xsp.start, xsp.bytes, full + st, pol_x);
  };
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < DEPTH; ++k)
      if (k < n_my) gather(k, k);
  }
  const int gl = min(lane, cnt - 1);          // idle lanes shadow the last owner (never stored)
  const bool owner = lane < cnt;
  const float* dg = ring + ds.skip + gl * DN;
  const float* xg = ring + L.xoff_w + xsp.skip + gl * XN;
  const size_t KW = (size_t)ldw;
  const float* w_g = weight + (size_t)(g_lo + gl) * WN;
  float* dW_g = dW + (size_t)(g_lo + gl) * WN;
  float* dx_w = dx + g_lo * XN;
  const uint32_t warp_bytes = (uint32_t)cnt * XN * 4;
  float r[WN];                                // dX role: my blocks of W_r; dW role: their gradient
#pragma unroll
  for (int i = 0; i < WN; ++i) r[i] = 0.f;
  int cur = -1;

  auto flush = [&](int rel) {                 // dW role, owners only
    red_flush<WN>(dW_g + (size_t)(unsigned)rel * KW, r, owner);
  };

  if (xrole) {                                // warp-uniform: one loop per role keeps the live ranges apart
    _Pragma("unroll 1") for (int k = 0; k < n_my; ++k) {
      const int st = k % DEPTH;
      mbar_wait(full + st, (k / DEPTH) & 1);
      const int4 p = rec[k];
      const float nv = __int_as_float(p.w);
      if (p.z != cur) {
        cur = p.z;
        ldg_weights<WN>(r, w_g + (size_t)(unsigned)cur * KW);
      }
      float dv[DN], m[XN];
      lds_vec<DN>(dv, dg + st * stage_f);
#pragma unroll
      for (int tb = 0; tb < TB; ++tb)
#pragma unroll
        for (int i = 0; i < SI; ++i) {
          float a = 0.f;
#pragma unroll
          for (int o = 0; o < SO; ++o) a = fmaf(dv[tb * SO + o], r[(tb * SI + i) * SO + o], a);
          m[tb * SI + i] = nv * a;
        }
      float* tb_cur = tbuf + (k % NT) * 32 * XN;
      bulk_wait_read<NT - 1>();
      __syncwarp();
      park_raw<XN>(tb_cur + lane * XN, m);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        bulk_red_add(const_cast<float*>(row_at(dx_w, p.x, in_w)), tb_cur, warp_bytes, pol_red);
        bulk_commit();
        if (k + DEPTH < n_my) gather(k + DEPTH, st);
      }
    }
    if (lane == 0) bulk_wait_all();
  } else {
    int k = 0;
    _Pragma("unroll 1") while (k < n_my) {    // one relation run at a time: the accumulators start from zero
      const int rel = rec[k].z;
      _Pragma("unroll 1") do {
        const int st = k % DEPTH;
        mbar_wait(full + st, (k / DEPTH) & 1);
        const float nv = __int_as_float(rec[k].w);
        float dv[DN], xv[XN];
        lds_vec<DN>(dv, dg + st * stage_f);
        lds_vec<XN>(xv, xg + st * stage_f);
#pragma unroll
        for (int tb = 0; tb < TB; ++tb)
#pragma unroll
          for (int i = 0; i < SI; ++i) {
            const float xs = nv * xv[tb * SI + i];
#pragma unroll
            for (int o = 0; o < SO; ++o)
              r[(tb * SI + i) * SO + o] = fmaf(xs, dv[tb * SO + o], r[(tb * SI + i) * SO + o]);
          }
        __syncwarp();                         // every lane has consumed ring stage st
        if (lane == 0 && k + DEPTH < n_my) gather(k + DEPTH, st);
        ++k;
      } while (k < n_my && rec[k].z == rel);
      flush(rel);
    }
  }
}

// ------------------------------------------------------------------------------------------
// fused backward, PAIRED variant for the streaming regime (dagg larger than L2)
// ------------------------------------------------------------------------------------------
// The input-gradient warp i and the weight-gradient warp i of a slot own the SAME blocks, hence the
// same slice of dagg[dst]: they are PARTNERS on one ring (stage = dagg slice + x slice, gathered
// once by the pair's issuer).  `full[stage]` is waited on by both; `empty[stage]` collects one
// arrival per partner before the issuer refills the stage, so the partners stay within DEPTH edges
// of each other and the dagg row crosses HBM / L2 once per edge, not once per role.
template <int SI, int SO, int TB, int WPR>
struct PairedLayout {
  static constexpr int XN = TB * SI, DN = TB * SO, WPS = 2 * WPR, SLOTS = kWarps / WPS, PAIRS = kWarps / 2;
  int lpw, xoff, stage_f;
  __host__ __device__ PairedLayout(int B) {
    lpw = lanes_per_warp(B / TB, WPR, XN);                     // whole 16-byte pieces of dx per warp
    xoff = round4(lpw * DN) + 4;                               // dagg slice, then x slice
    stage_f = xoff + round4(lpw * XN) + 4;
  }
  __host__ __device__ int pair_f(int depth) const { return depth * stage_f + 2 * 32 * XN; }   // ring + parked dx
  __host__ size_t smem(int depth) const {
    return sizeof(float) * ((size_t)4 * kChunk + (size_t)PAIRS * pair_f(depth)) + sizeof(uint64_t) * 2 * PAIRS * depth;
  }
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int SI, int SO, int TB, int WPR, int DEPTH>
__global__ void __launch_bounds__(kCta)
bwd_paired_kernel(RowSource x, const float* __restrict__ dagg, const int4* __restrict__ pack, int E,
           const float* __restrict__ weight, int B, int ldw, int ldd, int hints, float* __restrict__ dx,
           float* __restrict__ dW) {
  constexpr int XN = TB * SI, DN = TB * SO, WN = TB * SI * SO, WPS = 2 * WPR, SLOTS = kWarps / WPS, PAIRS = kWarps / 2;
  extern __shared__ __align__(16) float sm[];
  const PairedLayout<SI, SO, TB, WPR> L(B);
  const int in_w = B * SI, out_w = ldd, per = B / TB;          // out_w: row stride of dagg (see fwd_kernel on column chunks)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = warp / WPS, wsl = warp % WPS;
  const bool xrole = wsl < WPR;               // warp-uniform: input-gradient warps come first
  const int wr_i = xrole ? wsl : wsl - WPR;   // warp index inside its role = pair index inside the slot
  const int pair = slot * WPR + wr_i;
  const bool have_dx = dx != nullptr;
  const bool issuer = have_dx ? xrole : !xrole;               // who gathers for the pair
  int4* P_s = reinterpret_cast<int4*>(sm);
  float* ring = sm + 4 * kChunk + pair * L.pair_f(DEPTH);
  float* tbuf = ring + DEPTH * L.stage_f;     // [2][32 * XN] parked input gradients
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + 4 * kChunk + PAIRS * L.pair_f(DEPTH)) + pair * 2 * DEPTH;
  uint64_t* empty = full + DEPTH;
  const int e0 = blockIdx.x * kChunk, n = min(E - e0, kChunk);
  for (int i = threadIdx.x; i < n; i += kCta) P_s[i] = __ldg(pack + e0 + i);
  if (issuer && lane == 0) {
    for (int i = 0; i < DEPTH; ++i) {
      mbar_init(full + i, 1);
      mbar_init(empty + i, have_dx ? 2 : 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();                            // the only CTA-wide barrier
  const int share = (n + SLOTS - 1) / SLOTS;
  const int k_lo = slot * share, n_my = max(0, min(n - k_lo, share));
  const int g_lo = wr_i * L.lpw, cnt = max(0, min(L.lpw, per - g_lo));
  if (n_my == 0 || cnt == 0 || (xrole && !have_dx)) return;
  const int4* rec = P_s + k_lo;
  const Span ds = make_span(g_lo * DN, cnt * DN), xsp = make_span(g_lo * XN, cnt * XN);
  const uint64_t pol_x = x.parts != nullptr ? l2_policy(false) : (hints & 4) ? l2_policy_last(true) : l2_policy(hints & 1);
  const uint64_t pol_d = l2_policy(hints & 2);
  const uint64_t pol_red = l2_policy_last(hints & 4);
  auto gather = [&](int k, int st) {          // issuer lane 0: both slices of the pair's stage
    const int4 p = rec[k];
    mbar_expect_tx(full + st, ds.bytes + xsp.bytes);
    bulk_g2s(ring + st * L.stage_f, row_at(dagg, p.y, out_w) + ds.start, ds.bytes, full + st, pol_d);
    bulk_g2s(ring + st * L.stage_f + L.xoff, x.row(p.x, in_w) + xsp.start, xsp.bytes, full + st, pol_x);
  };
  // after this warp has consumed stage st for edge k: tell the pair, and (issuer) refill it for edge k + DEPTH
  auto release = [&](int k, int st) {         // lane 0
    mbar_arrive(empty + st);
    if (issuer && k + DEPTH < n_my) {
      mbar_wait(empty + st, (k / DEPTH) & 1); // the partner is done with it as well
      gather(k + DEPTH, st);
    }
  };
  if (issuer && lane == 0) {
#pragma unroll
    for (int k = 0; k < DEPTH; ++k)
      if (k < n_my) gather(k, k);
  }
  const int gl = min(lane, cnt - 1);          // idle lanes shadow the last owner (never stored)
  const bool owner = lane < cnt;
  const float* dg = ring + ds.skip + gl * DN;
  const float* xg = ring + L.xoff + xsp.skip + gl * XN;
  const size_t KW = (size_t)ldw;
  const float* w_g = weight + (size_t)(g_lo + gl) * WN;
  float* dW_g = dW + (size_t)(g_lo + gl) * WN;
  float* dx_w = dx + g_lo * XN;
  const uint32_t warp_bytes = (uint32_t)cnt * XN * 4;
  float r[WN];                                // dX role: my blocks of W_r; dW role: their gradient
#pragma unroll
  for (int i = 0; i < WN; ++i) r[i] = 0.f;

  if (xrole) {                                // warp-uniform: one loop per role keeps the live ranges apart
    int cur = -1;
    _Pragma("unroll 1") for (int k = 0; k < n_my; ++k) {
      const int st = k % DEPTH;
      mbar_wait(full + st, (k / DEPTH) & 1);
      const int4 p = rec[k];
      const float nv = __int_as_float(p.w);
      if (p.z != cur) {
        cur = p.z;
        ldg_weights<WN>(r, w_g + (size_t)(unsigned)cur * KW);
      }
      float dv[DN], m[XN];
      lds_vec<DN>(dv, dg + st * L.stage_f);
#pragma unroll
      for (int tb = 0; tb < TB; ++tb)
#pragma unroll
        for (int i = 0; i < SI; ++i) {
          float a = 0.f;
#pragma unroll
          for (int o = 0; o < SO; ++o) a = fmaf(dv[tb * SO + o], r[(tb * SI + i) * SO + o], a);
          m[tb * SI + i] = nv * a;
        }
      float* tb_cur = tbuf + (k & 1) * 32 * XN;
      bulk_wait_read<1>();
      __syncwarp();
      park_raw<XN>(tb_cur + lane * XN, m);
      fence_async_smem();
      __syncwarp();                           // gradients parked; every lane has consumed ring stage st
      if (lane == 0) {
        bulk_red_add(const_cast<float*>(row_at(dx_w, p.x, in_w)), tb_cur, warp_bytes, pol_red);
        bulk_commit();
        release(k, st);
      }
    }
    if (lane == 0) bulk_wait_all();
  } else {
    auto flush = [&](int rel) {               // owners only
      red_flush<WN>(dW_g + (size_t)(unsigned)rel * KW, r, owner);
    };
    int k = 0;
    _Pragma("unroll 1") while (k < n_my) {    // one relation run at a time: the accumulators start from zero
      const int rel = rec[k].z;
      _Pragma("unroll 1") do {
        const int st = k % DEPTH;
        mbar_wait(full + st, (k / DEPTH) & 1);
        const float nv = __int_as_float(rec[k].w);
        float dv[DN], xv[XN];
        lds_vec<DN>(dv, dg + st * L.stage_f);
        lds_vec<XN>(xv, xg + st * L.stage_f);
#pragma unroll
        for (int tb = 0; tb < TB; ++tb)
#pragma unroll
          for (int i = 0; i < SI; ++i) {
            const float xs = nv * xv[tb * SI + i];
#pragma unroll
            for (int o = 0; o < SO; ++o)
              r[(tb * SI + i) * SO + o] = fmaf(xs, dv[tb * SO + o], r[(tb * SI + i) * SO + o]);
          }
        __syncwarp();                         // every lane has consumed ring stage st
        if (lane == 0) release(k, st);
        ++k;
      } while (k < n_my && rec[k].z == rel);
      flush(rel);
    }
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
template <int FI, int FO, int TB, int WPS, int DEPTH, int NT = 2, int MINB = 1>
int launch_fwd(RowSource feat, const void* pack, int E, const float* weight, int B, int hints, float* out,
               cudaStream_t st, int ldw = 0, int ldo = 0) {
  if (ldw == 0) ldw = B * FI * FO;
  if (ldo == 0) ldo = B * FO;
  const size_t smem = FwdLayout<FI, FO, TB, WPS>(B).smem(DEPTH, NT);
  auto kern = fwd_kernel<FI, FO, TB, WPS, DEPTH, NT, MINB>;
  KG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<kg_div_up(E, kChunk), kCta, smem, st>>>(feat, reinterpret_cast<const int4*>(pack), E, weight, B, ldw, ldo, hints, out);
  KG_LAUNCH_OK();
  return KG_OK;
}

template <int SI, int SO, int TB, int WPR, int DEPTH, int NT = 2>
int launch_bwd(RowSource x, const float* dagg, const void* pack, int E, const float* weight, int B, int hints,
               float* dx, float* dW, cudaStream_t st, int ldw = 0, int ldd = 0) {
  if (ldw == 0) ldw = B * SI * SO;
  if (ldd == 0) ldd = B * SO;
  const size_t smem = BwdLayout<SI, SO, TB, WPR>(B).smem(DEPTH, NT);
  auto kern = bwd_kernel<SI, SO, TB, WPR, DEPTH, NT>;
  KG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<kg_div_up(E, kChunk), kCta, smem, st>>>(x, dagg, reinterpret_cast<const int4*>(pack), E, weight, B, ldw, ldd,
                                                 hints, dx, dW);
  KG_LAUNCH_OK();
  return KG_OK;
}

template <int SI, int SO, int TB, int WPR, int DEPTH>
int launch_bwd_paired(RowSource x, const float* dagg, const void* pack, int E, const float* weight, int B, int hints,
                      float* dx, float* dW, cudaStream_t st, int ldw = 0, int ldd = 0) {
  if (ldw == 0) ldw = B * SI * SO;
  if (ldd == 0) ldd = B * SO;
  const size_t smem = PairedLayout<SI, SO, TB, WPR>(B).smem(DEPTH);
  auto kern = bwd_paired_kernel<SI, SO, TB, WPR, DEPTH>;
  KG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<kg_div_up(E, kChunk), kCta, smem, st>>>(x, dagg, reinterpret_cast<const int4*>(pack), E, weight, B, ldw, ldd,
                                                 hints, dx, dW);
  KG_LAUNCH_OK();
  return KG_OK;
}

}  // namespace bddwarp
