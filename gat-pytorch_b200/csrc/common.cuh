// Shared helpers for libgat_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/gat_b200.h"

namespace gat {

void set_error(const char* fmt, ...);
void count_launch();   // bumps the counter behind gat_launch_count()

#define GAT_CHECK_ARG(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      gat::set_error(__VA_ARGS__);          \
      return GAT_EINVAL;                    \
    }                                       \
  } while (0)

#define GAT_CUDA(expr)                                                                  \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      gat::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                                   \
    }                                                                                   \
  } while (0)

#define GAT_LAUNCH_CHECK()                                                              \
  do {                                                                                  \
    gat::count_launch();                                                                \
    cudaError_t _e = cudaGetLastError();                                                \
    if (_e != cudaSuccess) {                                                            \
      gat::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                                   \
    }                                                                                   \
  } while (0)

constexpr float kLeakySlope = 0.01f;  // nn.LeakyReLU() default, gat_layer.py:87
constexpr float kSoftmaxEps = 1e-8f;  // gat_layer.py:109
constexpr int kNumSMs = 148;          // B200

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: the "already opted in" caches of the launchers are
// therefore indexed by the current device (a process that drives cuda:1 after cuda:0 must opt in again).  A racing
// second thread at worst repeats the (idempotent) attribute call.
constexpr int kMaxDevices = 64;
static inline int cur_device() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0) d = 0;
  return d % kMaxDevices;
}

// Philox-4x32-10 keyed on (seed), counter = (edge id, head/4 block, offset): one call yields the
// masks of 4 consecutive heads of one edge.  Stateless, so backward regenerates the same mask.
__device__ __forceinline__ uint4 philox4x32(uint64_t seed, uint64_t offset, uint32_t edge, uint32_t blk) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = edge, c1 = blk, c2 = (uint32_t)offset, c3 = (uint32_t)(offset >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// Dropout keep-scale for (edge, head): 0 or 1/(1-p).  nn.Dropout semantics, gat_layer.py:113-115.
__device__ __forceinline__ float dropout_scale(uint64_t seed, uint64_t offset, uint32_t edge, int head, float p) {
  uint4 r = philox4x32(seed, offset, edge, (uint32_t)(head >> 2));
  uint32_t v = (head & 3) == 0 ? r.x : (head & 3) == 1 ? r.y : (head & 3) == 2 ? r.z : r.w;
  float u = (float)(v >> 8) * (1.0f / 16777216.0f);  // [0,1)
  return u < p ? 0.0f : 1.0f / (1.0f - p);
}

// Output glue: what GATModel.forward does between a layer and the next one -- skip add, ELU, the next layer's
// input dropout (GATModel.py:130, :135-149) -- folded into the kernel that produces the layer's output:
//     y = keep(row, col) * E(out + skip),   E = ELU or identity,   keep = 0 or 1/(1-p) from Philox keyed on (row, col/4)
// so the mask is never stored; the backward regenerates it.
__device__ __forceinline__ float4 glue_keep4(uint64_t seed, int64_t row, int chunk, float p) {
  const uint4 r = philox4x32(seed, 0x676c7565ull, (uint32_t)row, (uint32_t)chunk);
  const float keep = 1.0f / (1.0f - p);
  return make_float4(((float)(r.x >> 8) * (1.0f / 16777216.0f)) < p ? 0.f : keep, ((float)(r.y >> 8) * (1.0f / 16777216.0f)) < p ? 0.f : keep,
                     ((float)(r.z >> 8) * (1.0f / 16777216.0f)) < p ? 0.f : keep, ((float)(r.w >> 8) * (1.0f / 16777216.0f)) < p ? 0.f : keep);
}
__device__ __forceinline__ float glue_keep1(uint64_t seed, int64_t row, int col, float p) {
  const float4 k = glue_keep4(seed, row, col >> 2, p);
  return (col & 3) == 0 ? k.x : (col & 3) == 1 ? k.y : (col & 3) == 2 ? k.z : k.w;
}
__device__ __forceinline__ float glue_elu(float v) { return v > 0.f ? v : expm1f(v); }
// adjoint of one element: dL/d(out+skip) from dL/dy and the stored y (keep = this element's keep scale, 1 without dropout)
__device__ __forceinline__ float glue_adjoint1(float g, float y, int act, float keep, float one_minus_p) {
  if (keep == 0.f) return 0.f;
  const float hval = y * one_minus_p;                       // E(out + skip)
  const float d = (act && hval <= 0.f) ? hval + 1.0f : 1.0f;   // ELU' = 1 (h > 0) or h + 1
  return g * keep * d;
}

// Order-preserving float <-> uint mapping so that a float max can use atomicMax (deterministic).
__device__ __forceinline__ unsigned int float_to_ordered(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

}  // namespace gat
