// Micro-benchmark: throughput of three ways to round fp32 -> tf32 precision (used by the 3xTF32 operand split).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cvt_tf32 cvt_tf32.cu && ./cvt_tf32
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float r_cvt(float v) { uint32_t r; asm volatile("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v)); return __uint_as_float(r); }
__device__ __forceinline__ float r_int(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ float r_velt(float v) { float g = v * 8193.0f; float d = v - g; return g + d; }

template <int MODE>
__global__ void k(float* out, int iters) {
  float a[8];
  for (int j = 0; j < 8; ++j) a[j] = (float)(threadIdx.x + j) * 1.0001f + 0.37f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float h = MODE == 0 ? r_cvt(a[j]) : (MODE == 1 ? r_int(a[j]) : r_velt(a[j]));
      a[j] = a[j] - h * 0.999f + 1.0f;     // dependent chain per j, 8 independent chains
    }
  }
  float s = 0; for (int j = 0; j < 8; ++j) s += a[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 3; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 8, 256>>>(out, iters);
      if (mode == 1) k<1><<<148 * 8, 256>>>(out, iters);
      if (mode == 2) k<2><<<148 * 8, 256>>>(out, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double roundings = 148.0 * 8 * 256 * 8 * iters;
      if (rep) printf("mode %d (%s): %.3f ms, %.1f G roundings/s, %.2f roundings/clk/SM @1.9GHz\n", mode,
                      mode == 0 ? "cvt.rna.tf32" : (mode == 1 ? "int add+and" : "veltkamp 3 flops"), ms, roundings / ms / 1e6,
                      roundings / (ms * 1e-3) / 148 / 1.9e9);
    }
  }
  return 0;
}
