// Shared pieces of the edge kernels (Kernel 3 / Kernel 4).
//
// Work decomposition: a GROUP of G lanes (G = 1..32, power of two) owns one CSR row.  A row of the
// padded feature matrices is dp = NH*Fp floats = dp/4 float4 "chunks"; lane g of the group owns chunks
// g, g+G, g+2G, ... (SLOTS of them), so one group-wide load instruction reads a contiguous G*16 bytes
// of the row.  Because Fp % 4 == 0 every chunk belongs to exactly one head.
//
// Per batch of G edges, lane t first does the per-(edge, head) scalar work for edge t of the batch
// (logit, exp, alpha, dropout...) and publishes (source id, per-head weights) in shared memory; then the
// whole group walks the batch, gathering Wh[src] rows with all loads of UNROLL edges issued before
// the first use.
#pragma once
#include "common.cuh"

namespace gat {

constexpr int kMaxHeads = 8;
constexpr int kEdgeThreads = 256;

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

template <int G>
__device__ __forceinline__ unsigned group_mask(int lane) {
  if (G == 32) return 0xffffffffu;
  return ((1u << G) - 1u) << (lane & ~(G - 1));
}

template <int G>
__device__ __forceinline__ float group_sum(float v, unsigned mask) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}

// Per-head dropout keep-scales of one edge (Philox keyed on the edge's position in the rewritten list).
__device__ __forceinline__ void dropout_scales(uint64_t seed, uint64_t offset, uint32_t edge, int nh, float p,
                                               float (&m)[kMaxHeads]) {
  const float keep = 1.0f / (1.0f - p);
#pragma unroll
  for (int b = 0; b < kMaxHeads / 4; ++b) {
    if (b * 4 < nh) {
      uint4 r = philox4x32(seed, offset, edge, (uint32_t)b);
      m[b * 4 + 0] = ((float)(r.x >> 8) * (1.0f / 16777216.0f)) < p ? 0.f : keep;
      m[b * 4 + 1] = ((float)(r.y >> 8) * (1.0f / 16777216.0f)) < p ? 0.f : keep;
      m[b * 4 + 2] = ((float)(r.z >> 8) * (1.0f / 16777216.0f)) < p ? 0.f : keep;
      m[b * 4 + 3] = ((float)(r.w >> 8) * (1.0f / 16777216.0f)) < p ? 0.f : keep;
    }
  }
}

// Un-normalised attention of one (edge, head): exp(LeakyReLU_0.01(l - M)), gat_layer.py:85-96, with the
// reference's operation order (subtract, activate, exp) so that equal inputs give equal bits up to expf.
__device__ __forceinline__ float attn_exp(float logit, float gmax) {
  float t = logit - gmax;
  float u = t >= 0.f ? t : t * kLeakySlope;
  return expf(u);
}

__device__ __forceinline__ float atomic_max_float(float* addr, float value) {
  return (value >= 0.f) ? __int_as_float(atomicMax((int*)addr, __float_as_int(value)))
                        : __uint_as_float(atomicMin((unsigned int*)addr, __float_as_uint(value)));
}

// Host-side choice of (G, SLOTS) for a padded row of `chunks` float4s.
struct GroupShape { int g, slots; };
static inline GroupShape pick_group(int chunks) {
  GroupShape s;
  if (chunks <= 1) { s.g = 1; s.slots = 1; }
  else if (chunks <= 2) { s.g = 2; s.slots = 1; }
  else if (chunks <= 4) { s.g = 4; s.slots = 1; }
  else if (chunks <= 8) { s.g = 8; s.slots = 1; }
  else if (chunks <= 16) { s.g = 16; s.slots = 1; }
  else {
    s.g = 32;
    int need = (chunks + 31) / 32;
    const int allowed[] = {1, 2, 3, 4, 6, 8};
    s.slots = -1;
    for (int a : allowed) if (a >= need) { s.slots = a; break; }
  }
  return s;
}

#define GAT_DISPATCH_GROUP(shape, LAUNCH)                                         \
  do {                                                                            \
    if ((shape).g == 1) { LAUNCH(1, 1); }                                         \
    else if ((shape).g == 2) { LAUNCH(2, 1); }                                    \
    else if ((shape).g == 4) { LAUNCH(4, 1); }                                    \
    else if ((shape).g == 8) { LAUNCH(8, 1); }                                    \
    else if ((shape).g == 16) { LAUNCH(16, 1); }                                  \
    else if ((shape).slots == 1) { LAUNCH(32, 1); }                               \
    else if ((shape).slots == 2) { LAUNCH(32, 2); }                               \
    else if ((shape).slots == 3) { LAUNCH(32, 3); }                               \
    else if ((shape).slots == 4) { LAUNCH(32, 4); }                               \
    else if ((shape).slots == 6) { LAUNCH(32, 6); }                               \
    else { LAUNCH(32, 8); }                                                       \
  } while (0)

}  // namespace gat
