// Microbenchmark: issue rate of 3-register FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
// Decides whether the block-owner message-passing kernels should use packed FMAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma_probe tools/ffma_probe.cu && /tmp/ffma_probe
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm volatile(
      "{ .reg .b64 ra, rb, rd;\n mov.b64 ra, {%2,%3};\n mov.b64 rb, {%4,%5};\n mov.b64 rd, {%0,%1};\n"
      " fma.rn.f32x2 rd, ra, rb, rd;\n mov.b64 {%0,%1}, rd; }"
      : "+f"(d0), "+f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, const float* in, int iters) {
  float a[16], w[16], x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    a[i] = in[threadIdx.x + i];
    w[i] = in[threadIdx.x + 16 + i];
    x[i] = in[threadIdx.x + 32 + i];
  }
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(x[i], w[i], a[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 16; i += 2) ffma2(a[i], a[i + 1], x[i], x[i + 1], w[i], w[i + 1]);
    }
  }
  float t = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) t += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
int main() {
  float *o, *in;
  cudaMalloc(&o, 148 * 8 * 256 * 4);
  cudaMalloc(&in, 4096);
  cudaMemset(in, 0, 4096);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000;
  const int ctas[3] = {1, 2, 8};
  for (int c = 0; c < 3; ++c)
    for (int mode = 0; mode < 2; ++mode)
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148 * ctas[c], 256>>>(o, in, iters);
        else k<1><<<148 * ctas[c], 256>>>(o, in, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl = 2.0 * 16 * iters * 148.0 * ctas[c] * 256;
        if (rep) printf("%s ctas/SM=%d: %.3f ms  %.1f TFLOP/s fp32\n", mode ? "FFMA2" : "FFMA ", ctas[c], ms, fl / ms * 1e-9);
      }
  return 0;
}
