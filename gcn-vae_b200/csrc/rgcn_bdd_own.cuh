// a3: RelGraphConv(regularizer="bdd") message passing, BLOCK-OWNER kernels (the reference model's
// 5x5 and 5x10 blocks; DGL RelGraphConv as constructed at kgvae/model.py:54-59).
//
// A thread owns TB whole diagonal blocks of the relation's weight - TB*si*so = 100 floats, read
// straight from the DGL-layout weight row [B][si][so] (they are contiguous) - and keeps them in
// REGISTERS while the relation lasts.  Per edge it reads its TB*si inputs with 64/128-bit shared
// memory loads, does TB*si*so FMAs and emits TB*so outputs with vector reductions: no per-edge
// weight traffic, no selects, 2/3 of the issued instructions are FMAs.  With B = 100 an edge needs
// 25 (5x5, TB = 4) or 50 (5x10, TB = 2) lanes: ONE or TWO warps per edge, so a 128-thread CTA runs
// 4 (2) independent slots with __syncwarp / a 64-thread named barrier as the only synchronisation.
//
// Data movement is bulk-asynchronous (TMA engine, no per-lane address arithmetic):
//   gather   one cp.async.bulk global -> shared per row (2 KB) into a 4-deep ring, completion on an
//            mbarrier per stage; L2 evict-first hint when the gathered matrix streams from HBM
//   scatter  the warp parks its scaled outputs in shared memory and ONE lane issues
//            cp.reduce.async.bulk.add.f32 shared -> global for the warp's contiguous 1-2 KB of the
//            destination row: the atomic adds happen in L2 on whole lines, the SM issues no REDs.
//
// Backward is fused: the warps of a slot split into an input-gradient role (same shape as the
// forward, transposed contraction) and a weight-gradient role (TB*si*so outer-product accumulators
// in the same registers, flushed with vector reductions when the relation changes); both read the
// gathered dagg[dst] row, the weight-gradient role also x[src].
#pragma once
#include "common.cuh"

namespace bddown {

constexpr int kCta = 128;     // threads per CTA (4 warps)
constexpr int kChunk = 256;   // consecutive relation-sorted edges per CTA
constexpr int kDepth = 4;     // gathered rows in flight per slot

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}
__device__ __forceinline__ void red_add_v2(float* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy(bool evict_first) {
  uint64_t pol;
  if (evict_first) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_last(bool evict_last) {
  uint64_t pol;
  if (evict_last) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// bulk copy global -> shared (bytes % 16 == 0, both 16-byte aligned), completes on `bar`
__device__ __forceinline__ void bulk_g2s(float* dst, const float* __restrict__ src, uint32_t bytes, uint64_t* bar,
                                         uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}
// bulk reduction shared -> global: dst[0..bytes/4) += src[0..bytes/4) (fp32), part of the thread's bulk group
__device__ __forceinline__ void bulk_red_add(float* dst, const float* src, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.L2::cache_hint.add.f32 [%0], [%1], %2, %3;"
               ::"l"(dst), "r"(smem_u32(src)), "r"(bytes), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// order this thread's generic-proxy shared-memory writes before later async-proxy (bulk) reads
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// synchronise the WARPS warps of slot `slot` (immediate barrier ids: the CTA reserves 1 + slots)
template <int WARPS>
__device__ __forceinline__ void slot_sync(int slot) {
  if (WARPS == 1) {
    __syncwarp();
  } else if (WARPS == 2) {
    if (slot == 0) asm volatile("bar.sync 1, 64;" ::: "memory");
    else asm volatile("bar.sync 2, 64;" ::: "memory");
  } else {
    asm volatile("bar.sync 1, 128;" ::: "memory");
  }
}

__device__ __forceinline__ const float* row_at(const float* base, int row, int width) {
  return base + (size_t)(unsigned)row * (unsigned)width;
}

// Where gathered rows live: one matrix, or (destination-partitioned training) one row block per
// rank, each in the owner's HBM and mapped into this process (CUDA IPC): row r is row r % part_rows
// of parts[r / part_rows].  A peer row is fetched by the same bulk copy, over NVLink - the
// "all-gather" of layer inputs is fused into the gather stage of the message-passing kernel and
// moves only the rows this rank's edges reference.
struct RowSource {
  const float* base;
  const float* const* parts;
  int part_rows;
  __device__ __forceinline__ const float* row(int r, int width) const {
    if (parts == nullptr) return row_at(base, r, width);
    const int owner = r / part_rows;
    return row_at(parts[owner], r - owner * part_rows, width);
  }
};

// N contiguous floats from shared memory with the widest loads the alignment of N allows
template <int N>
__device__ __forceinline__ void lds_vec(float (&v)[N], const float* p) {
  if (N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N; i += 4) {
      const float4 t = *reinterpret_cast<const float4*>(p + i);
      v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; i += 2) {
      const float2 t = *reinterpret_cast<const float2*>(p + i);
      v[i] = t.x; v[i + 1] = t.y;
    }
  }
}

// park s * v[0..N) (this lane's N consecutive output columns) in the warp's staging buffer
template <int N>
__device__ __forceinline__ void park(float* mine, const float (&v)[N], float s) {
  if (N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N; i += 4)
      *reinterpret_cast<float4*>(mine + i) = make_float4(s * v[i], s * v[i + 1], s * v[i + 2], s * v[i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < N; i += 2) *reinterpret_cast<float2*>(mine + i) = make_float2(s * v[i], s * v[i + 1]);
  }
}

// owner lanes per warp when `per` groups of N columns are spread over `warps` warps; even when N * 4
// is not a multiple of 16 so that every warp's share of a row is a whole number of 16-byte pieces
__device__ __host__ __forceinline__ int lanes_per_warp(int per, int warps, int n_cols) {
  int lpw = (per + warps - 1) / warps;
  if ((n_cols * 4) % 16 != 0 && (lpw & 1)) ++lpw;
  return lpw;
}

// group (owned block set) of this lane, or -1
__device__ __forceinline__ int group_of(int warp_in_role, int lane, int per, int lpw) {
  const int g = warp_in_role * lpw + lane;
  return (lane < lpw && g < per) ? g : -1;
}

// ------------------------------------------------------------------------------------------
// forward: out[dst] += norm * blockdiag(W_etype) feat[src]
// weight [R][B][FI][FO] (DGL layout); out zero-filled by the caller
// ------------------------------------------------------------------------------------------
template <int FI, int FO, int TB, int WPS>
__global__ void __launch_bounds__(kCta)
fwd_kernel(RowSource feat, const int4* __restrict__ pack, int E,
           const float* __restrict__ weight, int B, int hints, float* __restrict__ out) {
  constexpr int XN = TB * FI, CN = TB * FO, WN = TB * FI * FO, SLOTS = 4 / WPS;
  extern __shared__ __align__(16) float sm[];
  const int width = B * FO, in_w = B * FI;
  int4* P_s = reinterpret_cast<int4*>(sm);    // [kChunk] {src, dst, etype, norm}
  float* X_s = sm + 4 * kChunk;               // [SLOTS][kDepth][in_w] gathered rows
  float* T_s = X_s + SLOTS * kDepth * in_w;   // [4 warps][2][32 * CN] parked outputs
  uint64_t* bars = reinterpret_cast<uint64_t*>(T_s + 4 * 2 * 32 * CN);   // [SLOTS][kDepth]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = warp / WPS, wsl = warp % WPS;
  const int e0 = blockIdx.x * kChunk, n = min(E - e0, kChunk);
  for (int i = threadIdx.x; i < n; i += kCta) P_s[i] = __ldg(pack + e0 + i);
  if (threadIdx.x == 0) {
    for (int i = 0; i < SLOTS * kDepth; ++i) mbar_init(bars + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // a slot takes a contiguous share of the chunk: the relation (and so the registers) changes rarely
  const int share = (n + SLOTS - 1) / SLOTS;
  const int k_lo = slot * share, n_my = max(0, min(n - k_lo, share));
  if (n_my == 0) return;
  const int4* rec = P_s + k_lo;
  float* ring = X_s + slot * kDepth * in_w;
  uint64_t* full = bars + slot * kDepth;
  // peer rows bypass the local L2 anyway; a cache hint on them only slows the copies down (measured 4x)
  const uint64_t pol = l2_policy((hints & 1) && feat.parts == nullptr), pol_red = l2_policy_last(hints & 4);
  const uint32_t row_bytes = (uint32_t)in_w * 4;
  const bool leader = wsl == 0 && lane == 0;  // issues the slot's gathers

  if (leader) {
#pragma unroll
    for (int k = 0; k < kDepth; ++k)
      if (k < n_my) {
        mbar_expect_tx(full + k, row_bytes);
        bulk_g2s(ring + k * in_w, feat.row(rec[k].x, in_w), row_bytes, full + k, pol);
      }
  }

  const int lpw = lanes_per_warp(B / TB, WPS, CN);
  const int g = group_of(wsl, lane, B / TB, lpw);
  const int g_lo = wsl * lpw;                                        // first group of this warp
  const uint32_t warp_bytes = (uint32_t)max(0, min(lpw, B / TB - g_lo)) * CN * 4;   // this warp's share of a row
  float* tbuf = T_s + warp * 2 * 32 * CN;
  float* out_w = out + g_lo * CN;
  const float* xg = ring + max(g, 0) * XN;
  const float* w_g = weight + max(g, 0) * WN;
  float w[WN];
  int cur = -1;
  for (int k = 0; k < n_my; k += kDepth) {
#pragma unroll
    for (int u = 0; u < kDepth; ++u) {
      if (k + u < n_my) {                     // uniform across the slot
        mbar_wait(full + u, (k / kDepth) & 1);              // row k+u has landed
        const int4 p = rec[k + u];
        float m[CN];
        if (g >= 0) {
          if (p.z != cur) {                   // relation run starts: my TB blocks of W_r
            cur = p.z;
            const float4* wr = reinterpret_cast<const float4*>(row_at(w_g, cur, B * FI * FO));
#pragma unroll
            for (int i = 0; i < WN / 4; ++i) {
              const float4 t = __ldg(wr + i);
              w[4 * i] = t.x; w[4 * i + 1] = t.y; w[4 * i + 2] = t.z; w[4 * i + 3] = t.w;
            }
          }
          float xv[XN];
          lds_vec<XN>(xv, xg + u * in_w);
#pragma unroll
          for (int tb = 0; tb < TB; ++tb)
#pragma unroll
            for (int o = 0; o < FO; ++o) {
              float a = 0.f;
#pragma unroll
              for (int i = 0; i < FI; ++i) a = fmaf(xv[tb * FI + i], w[(tb * FI + i) * FO + o], a);
              m[tb * FO + o] = a;
            }
        }
        float* tb_cur = tbuf + (u & 1) * 32 * CN;
        if (lane == 0) bulk_wait_read<1>();   // the reduction that read this buffer two edges ago is done with it
        __syncwarp();
        if (g >= 0) {
          park<CN>(tb_cur + lane * CN, m, __int_as_float(p.w));
          fence_async_smem();
        }
        slot_sync<WPS>(slot);                 // outputs parked; everybody has consumed ring stage u
        if (lane == 0) {
          if (warp_bytes) bulk_red_add(const_cast<float*>(row_at(out_w, p.y, width)), tb_cur, warp_bytes, pol_red);
          bulk_commit();
        }
        if (leader && k + u + kDepth < n_my) {              // refill the stage just consumed
          mbar_expect_tx(full + u, row_bytes);
          bulk_g2s(ring + u * in_w, feat.row(rec[k + u + kDepth].x, in_w), row_bytes, full + u, pol);
        }
      }
    }
  }
  if (lane == 0) bulk_wait_all();             // shared memory must outlive the reductions reading it
}

// ------------------------------------------------------------------------------------------
// fused backward: dx[src] += norm * blockdiag(W_r)^T dagg[dst]
//                 dW[r][b][i][o] += norm * x[src][b*SI+i] * dagg[dst][b*SO+o]
// weight, dW [R][B][SI][SO]; dx, dW zero-filled by the caller; dx may be null
// slot = WPR input-gradient warps then WPR weight-gradient warps
// ------------------------------------------------------------------------------------------
template <int SI, int SO, int TB, int WPR>
__global__ void __launch_bounds__(kCta)
bwd_kernel(RowSource x, const float* __restrict__ dagg, const int4* __restrict__ pack, int E,
           const float* __restrict__ weight, int B, int hints, float* __restrict__ dx, float* __restrict__ dW) {
  constexpr int XN = TB * SI, DN = TB * SO, WN = TB * SI * SO, WPS = 2 * WPR, SLOTS = 4 / WPS;
  extern __shared__ __align__(16) float sm[];
  const int in_w = B * SI, out_w = B * SO, row_w = in_w + out_w;
  int4* P_s = reinterpret_cast<int4*>(sm);    // [kChunk]
  float* R_s = sm + 4 * kChunk;               // [SLOTS][kDepth][in_w + out_w]: x row then dagg row
  float* T_s = R_s + SLOTS * kDepth * row_w;  // [4 warps][2][32 * XN] parked input gradients
  uint64_t* bars = reinterpret_cast<uint64_t*>(T_s + 4 * 2 * 32 * XN);   // [SLOTS][kDepth]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = warp / WPS, wsl = warp % WPS;
  const int e0 = blockIdx.x * kChunk, n = min(E - e0, kChunk);
  for (int i = threadIdx.x; i < n; i += kCta) P_s[i] = __ldg(pack + e0 + i);
  if (threadIdx.x == 0) {
    for (int i = 0; i < SLOTS * kDepth; ++i) mbar_init(bars + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int share = (n + SLOTS - 1) / SLOTS;
  const int k_lo = slot * share, n_my = max(0, min(n - k_lo, share));
  if (n_my == 0) return;
  const int4* rec = P_s + k_lo;
  float* ring = R_s + slot * kDepth * row_w;
  uint64_t* full = bars + slot * kDepth;
  const uint64_t pol_x = x.parts != nullptr ? l2_policy(false) : (hints & 4) ? l2_policy_last(true) : l2_policy(hints & 1);
  const uint64_t pol_d = l2_policy(hints & 2);
  const uint64_t pol_red = l2_policy_last(hints & 4);
  const uint32_t x_bytes = (uint32_t)in_w * 4, d_bytes = (uint32_t)out_w * 4;
  const bool leader = wsl == 0 && lane == 0;

  auto gather = [&](int k, int stage) {       // leader only
    const int4 p = rec[k];
    mbar_expect_tx(full + stage, x_bytes + d_bytes);
    bulk_g2s(ring + stage * row_w, x.row(p.x, in_w), x_bytes, full + stage, pol_x);
    bulk_g2s(ring + stage * row_w + in_w, row_at(dagg, p.y, out_w), d_bytes, full + stage, pol_d);
  };
  if (leader) {
#pragma unroll
    for (int k = 0; k < kDepth; ++k)
      if (k < n_my) gather(k, k);
  }

  const bool xrole = wsl < WPR;               // warp-uniform: input-gradient warps come first
  const int lpw = lanes_per_warp(B / TB, WPR, xrole ? XN : 4);
  const int wr_i = xrole ? wsl : wsl - WPR;   // warp index inside its role
  const int g = (xrole && dx == nullptr) ? -1 : group_of(wr_i, lane, B / TB, lpw);
  const int gg = max(g, 0);
  const int g_lo = wr_i * lpw;
  const uint32_t warp_bytes = (uint32_t)max(0, min(lpw, B / TB - g_lo)) * XN * 4;   // dx share of an x-role warp
  float* tbuf = T_s + warp * 2 * 32 * XN;
  float* dx_w = dx + g_lo * XN;
  const float* xg = ring + gg * XN;
  const float* dg = ring + in_w + gg * DN;
  const size_t KW = (size_t)B * SI * SO;
  float r[WN];                                // dX role: my blocks of W_r; dW role: their gradient
#pragma unroll
  for (int i = 0; i < WN; ++i) r[i] = 0.f;
  int cur = -1;

  auto flush = [&](int rel) {                 // dW role
    float* dst = dW + (size_t)(unsigned)rel * KW + gg * WN;
#pragma unroll
    for (int i = 0; i < WN; i += 4) {
      red_add_v4(dst + i, r[i], r[i + 1], r[i + 2], r[i + 3]);
      r[i] = r[i + 1] = r[i + 2] = r[i + 3] = 0.f;
    }
  };

  for (int k = 0; k < n_my; k += kDepth) {
#pragma unroll
    for (int u = 0; u < kDepth; ++u) {
      if (k + u < n_my) {
        mbar_wait(full + u, (k / kDepth) & 1);
        const int4 p = rec[k + u];
        const float nv = __int_as_float(p.w);
        float* tb_cur = tbuf + (u & 1) * 32 * XN;
        if (xrole) {                          // warp-uniform
          float m[XN];
          if (g >= 0) {
            if (p.z != cur) {
              cur = p.z;
              const float4* wr = reinterpret_cast<const float4*>(weight + (size_t)(unsigned)cur * KW + gg * WN);
#pragma unroll
              for (int i = 0; i < WN / 4; ++i) {
                const float4 t = __ldg(wr + i);
                r[4 * i] = t.x; r[4 * i + 1] = t.y; r[4 * i + 2] = t.z; r[4 * i + 3] = t.w;
              }
            }
            float dv[DN];
            lds_vec<DN>(dv, dg + u * row_w);
#pragma unroll
            for (int tb = 0; tb < TB; ++tb)
#pragma unroll
              for (int i = 0; i < SI; ++i) {
                float a = 0.f;
#pragma unroll
                for (int o = 0; o < SO; ++o) a = fmaf(dv[tb * SO + o], r[(tb * SI + i) * SO + o], a);
                m[tb * SI + i] = a;
              }
          }
          if (lane == 0) bulk_wait_read<1>();
          __syncwarp();
          if (g >= 0) {
            park<XN>(tb_cur + lane * XN, m, nv);
            fence_async_smem();
          }
        } else if (g >= 0) {
          if (p.z != cur) {
            if (cur >= 0) flush(cur);
            cur = p.z;
          }
          float dv[DN], xv[XN];
          lds_vec<DN>(dv, dg + u * row_w);
          lds_vec<XN>(xv, xg + u * row_w);
#pragma unroll
          for (int tb = 0; tb < TB; ++tb)
#pragma unroll
            for (int i = 0; i < SI; ++i) {
              const float xs = nv * xv[tb * SI + i];
#pragma unroll
              for (int o = 0; o < SO; ++o)
                r[(tb * SI + i) * SO + o] = fmaf(xs, dv[tb * SO + o], r[(tb * SI + i) * SO + o]);
            }
        }
        slot_sync<WPS>(slot);                 // gradients parked; everybody has consumed ring stage u
        if (xrole && lane == 0) {
          if (dx != nullptr && warp_bytes)
            bulk_red_add(const_cast<float*>(row_at(dx_w, p.x, in_w)), tb_cur, warp_bytes, pol_red);
          bulk_commit();
        }
        if (leader && k + u + kDepth < n_my) gather(k + u + kDepth, u);
      }
    }
  }
  if (g >= 0 && !xrole && cur >= 0) flush(cur);
  if (xrole && lane == 0) bulk_wait_all();
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
template <int FI, int FO, int TB, int WPS>
int launch_fwd(RowSource feat, const void* pack, int E, const float* weight, int B, int hints, float* out,
               cudaStream_t st) {
  constexpr int SLOTS = 4 / WPS;
  const size_t smem = sizeof(float) * ((size_t)4 * kChunk + (size_t)SLOTS * kDepth * B * FI + 4 * 2 * 32 * TB * FO) +
                      sizeof(uint64_t) * SLOTS * kDepth;
  auto kern = fwd_kernel<FI, FO, TB, WPS>;
  KG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<kg_div_up(E, kChunk), kCta, smem, st>>>(feat, reinterpret_cast<const int4*>(pack), E, weight, B, hints, out);
  KG_LAUNCH_OK();
  return KG_OK;
}

template <int SI, int SO, int TB, int WPR>
int launch_bwd(RowSource x, const float* dagg, const void* pack, int E, const float* weight, int B, int hints,
               float* dx, float* dW, cudaStream_t st) {
  constexpr int SLOTS = 4 / (2 * WPR);
  const size_t smem = sizeof(float) * ((size_t)4 * kChunk + (size_t)SLOTS * kDepth * B * (SI + SO) + 4 * 2 * 32 * TB * SI) +
                      sizeof(uint64_t) * SLOTS * kDepth;
  auto kern = bwd_kernel<SI, SO, TB, WPR>;
  KG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<kg_div_up(E, kChunk), kCta, smem, st>>>(x, dagg, reinterpret_cast<const int4*>(pack), E, weight, B, hints,
                                                 dx, dW);
  KG_LAUNCH_OK();
  return KG_OK;
}

// shapes the block-owner kernels cover: 5x5 blocks (4 per thread) and 5x10 blocks (2 per thread)
// with at most 32 / 64 owner lanes per edge
inline bool eligible(int B, int si, int so) {
  if (si == 5 && so == 5) return B % 4 == 0 && B / 4 <= 32;
  if (si == 5 && so == 10) return B % 2 == 0 && B / 2 <= 64;
  return false;
}

}  // namespace bddown
