// Shared helpers for the kgvae_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/kgvae_b200.h"

int kg_fail(int code, const char* fmt, ...);

#define KG_REQUIRE(cond, ...)                                   \
  do {                                                          \
    if (!(cond)) return kg_fail(KG_ERR_INVALID, __VA_ARGS__);   \
  } while (0)

#define KG_CUDA(expr)                                                                        \
  do {                                                                                       \
    cudaError_t e_ = (expr);                                                                 \
    if (e_ != cudaSuccess)                                                                   \
      return kg_fail(KG_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, \
                     __LINE__);                                                              \
  } while (0)

#define KG_LAUNCH_OK() KG_CUDA(cudaGetLastError())

static inline cudaStream_t kg_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int kg_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

static inline size_t kg_align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-provided workspace.
struct KgArena {
  char* base;
  size_t size, used;
  KgArena(void* p, size_t n) : base(reinterpret_cast<char*>(p)), size(n), used(0) {}
  template <typename T>
  T* take(size_t count) {
    size_t bytes = kg_align_up(count * sizeof(T));
    if (used + bytes > size) return nullptr;
    T* out = reinterpret_cast<T*>(base + used);
    used += bytes;
    return out;
  }
};

__device__ __forceinline__ float kg_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ int kg_warp_sum_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// torch.nn.functional.softplus (beta=1, threshold=20)
__device__ __forceinline__ float kg_softplus(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float kg_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

// Number of SMs of the current device (cached per device; immutable attribute).
int kg_sm_count();
// true the first time it is called for (slot, current device): guards per-device cudaFuncSetAttribute calls
bool kg_attr_needed(int slot);
