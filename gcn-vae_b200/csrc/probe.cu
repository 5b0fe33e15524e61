// Measurement aid (not on the product path): what the L2 of this GPU sustains for the access pattern of
// the fused DistMult pass (decoder.cu) - whole fp32 rows of an L2-resident matrix read at random, whole
// rows of a second L2-resident matrix reduced into at random.  bench.py divides the DistMult launch's
// L2 bytes per second by the figure this probe reaches on the same GPU in the same process: the
// roofline denominator for a kernel whose working set (z and dz, 29 MB each at the FB15k-237 shape)
// never leaves L2, where the HBM figure of MEASURED_PEAKS.json says nothing.
//
// One warp per op stream, no dependence between ops, UNROLL ops in flight per warp, rows picked by a
// multiplicative hash: the only limits left are the SM<->L2 paths and the L2 slices themselves.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ void probe_red_v4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

__device__ __forceinline__ unsigned probe_row(unsigned op, unsigned salt, unsigned rows) {
  unsigned x = (op + salt) * 2654435761u;
  x ^= x >> 15;
  x *= 2246822519u;
  x ^= x >> 13;
  return x % rows;
}

// mode bit 0: read a random row of src;  bit 1: reduce a row into a random row of dst
template <int NV, int UNROLL>
__global__ void __launch_bounds__(kThreads)
l2_probe_kernel(const float* __restrict__ src, float* __restrict__ dst, unsigned rows, int row_floats, long long n_ops,
                int mode, float* __restrict__ sink) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31, nvec = row_floats >> 2;
  float4 keep = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long op0 = warp * UNROLL; op0 < n_ops; op0 += n_warps * UNROLL) {
    float4 v[UNROLL][NV];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const unsigned r = probe_row((unsigned)(op0 + u), 0x9e3779b9u, rows);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        v[u][i] = make_float4(1e-30f, 1e-30f, 1e-30f, 1e-30f);
        if ((mode & 1) && c < nvec) v[u][i] = __ldg(reinterpret_cast<const float4*>(src + (size_t)r * row_floats) + c);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const unsigned r = probe_row((unsigned)(op0 + u), 0x85ebca6bu, rows);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        if (mode & 2) {
          if (c < nvec) probe_red_v4(dst + (size_t)r * row_floats + 4 * c, v[u][i]);
        } else {
          keep.x += v[u][i].x; keep.y += v[u][i].y; keep.z += v[u][i].z; keep.w += v[u][i].w;
        }
      }
    }
  }
  if (keep.x + keep.y + keep.z + keep.w == 12345.678f) sink[0] = keep.x;   // keeps the loads alive
}

}  // namespace

// Runs n_ops independent (read a random row of src [rows, row_floats]) / (reduce into a random row of
// dst [rows, row_floats]) operations; the caller times the call with CUDA events.  Bytes moved through L2:
// n_ops * 4 * row_floats per enabled direction.  row_floats % 4 == 0, row_floats <= 1024, 16-byte aligned.
extern "C" int kg_probe_l2(const float* src, float* dst, int rows, int row_floats, long long n_ops, int mode,
                           float* sink, void* stream) {
  KG_REQUIRE(rows > 0 && row_floats > 0 && row_floats % 4 == 0 && row_floats <= 1024, "l2 probe: bad row shape");
  KG_REQUIRE(n_ops >= 0 && (mode & 3) != 0 && sink, "l2 probe: bad arguments");
  KG_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0, "l2 probe: unaligned");
  if (n_ops == 0) return KG_OK;
  const int grid = kg_sm_count() * 8;       // 2048 threads per SM: full occupancy at <= 32 registers... (64 regs: 4 CTAs)
  cudaStream_t st = kg_stream(stream);
  if (row_floats <= 512)
    l2_probe_kernel<4, 2><<<grid, kThreads, 0, st>>>(src, dst, (unsigned)rows, row_floats, n_ops, mode, sink);
  else
    l2_probe_kernel<8, 1><<<grid, kThreads, 0, st>>>(src, dst, (unsigned)rows, row_floats, n_ops, mode, sink);
  KG_LAUNCH_OK();
  return KG_OK;
}
