// Kernel 3a (global max of the logits) and Kernel 3 (fused edge forward), plus the head merge.
// Replaces gat_layer.py:70-132: no (E', NH, F) tensor is ever materialised.
#include "edge_common.cuh"

namespace gat {

// ------------------------------------------------------------------------------------------
// Kernel 3a: M = max_{e,h} (s_src[src_e,h] + s_tgt[dst_e,h])      (gat_layer.py:85)
// Persistent grid, one warp per destination row handed out long rows first (same scheduler as Kernel 3);
// s_src (n*NH floats) is L2-resident, so DRAM traffic is ~ col only.  max() is order independent, so the
// float atomic max keeps the result deterministic.
// ------------------------------------------------------------------------------------------
struct EdgeMaxParams {
  const int32_t* rowptr; const int32_t* col; RowSched sched;
  const float* s_src; const float* s_tgt; int nh; float* gmax;
};

__device__ __forceinline__ float edge_max_span(const EdgeMaxParams& P, const int64_t row, const int first, const int step,
                                               float m) {
  const int nh = P.nh;
  float st[kMaxHeads];
#pragma unroll
  for (int h = 0; h < kMaxHeads; ++h) st[h] = h < nh ? __ldg(P.s_tgt + row * nh + h) : 0.f;
  const int start = __ldg(P.rowptr + row), end = __ldg(P.rowptr + row + 1);
  for (int e = start + first; e < end; e += step) {
    const float* ss = P.s_src + (int64_t)__ldg(P.col + e) * nh;
    if (nh == 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(ss));
      m = fmaxf(m, fmaxf(fmaxf(v.x + st[0], v.y + st[1]), fmaxf(v.z + st[2], v.w + st[3])));
    } else {
#pragma unroll
      for (int h = 0; h < kMaxHeads; ++h)
        if (h < nh) m = fmaxf(m, __ldg(ss + h) + st[h]);
    }
  }
  return m;
}

__global__ void __launch_bounds__(256)
edge_max_kernel(const EdgeMaxParams P) {
  __shared__ float warp_max[8];
  __shared__ int sh_ctl;
  const int tid = threadIdx.x, lane = tid & 31;
  float m = -INFINITY;
  // long rows: the whole CTA strides over the row (max is exact and order independent)
  for (;;) {
    const int64_t row = grab_long_row(P.sched, P.rowptr, &sh_ctl);
    if (row < 0) break;
    m = edge_max_span(P, row, tid, 256, m);
  }
  int64_t base;
  while (grab_rows<32>(P.sched, lane, base)) {
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
      const int64_t row = sched_row<32>(P.sched, base, k, lane);
      if (row < 0 || taken_by_cta_phase(P.sched, P.rowptr, row)) continue;
      m = edge_max_span(P, row, lane, 32, m);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) warp_max[tid >> 5] = m;
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, warp_max[w]);
    if (m > -INFINITY) atomic_max_float(P.gmax, m);
  }
}

// ------------------------------------------------------------------------------------------
// Kernel 3: fused forward over destination rows (persistent grid, rows handed out per warp).
// ------------------------------------------------------------------------------------------
struct EdgeFwdParams {
  const int32_t* rowptr; const int32_t* col; const int32_t* eid; RowSched sched;
  const float* wh; int nh; int dp; int chunks; int chunks_per_head;
  const float* s_src; const float* s_tgt; const float* gmax;
  int const_attention; float dropout_p; uint64_t seed; uint64_t offset;
  float* out; float* alpha_out; float* z_out;
  int out_act;   // 1: the stored row is ELU(out) -- the F.elu that follows a hidden layer (GATModel.py:148-149), fused
  int32_t* tie_dst; int32_t* tie_src; unsigned long long* tie_total;
  // rest of the output glue (common.cuh): skip rows added before the activation, dropout of the stored row
  const float* skip; int64_t ld_skip; float out_drop_p; uint64_t out_drop_seed;
};

__device__ __forceinline__ float4 elu_f4(float4 v) {
  return make_float4(v.x > 0.f ? v.x : expm1f(v.x), v.y > 0.f ? v.y : expm1f(v.y), v.z > 0.f ? v.z : expm1f(v.z),
                     v.w > 0.f ? v.w : expm1f(v.w));
}

// what is stored for chunk c of row `row`: keep * E(t + skip)
__device__ __forceinline__ float4 out_glue(const EdgeFwdParams& P, const int64_t row, const int c, float4 t) {
  if (P.skip) {
    const float4 k = ldg4(P.skip + row * P.ld_skip + c * 4);
    t.x += k.x; t.y += k.y; t.z += k.z; t.w += k.w;
  }
  if (P.out_act) t = elu_f4(t);
  if (P.out_drop_p > 0.f) {
    const float4 k = glue_keep4(P.out_drop_seed, row, c, P.out_drop_p);
    t.x *= k.x; t.y *= k.y; t.z *= k.z; t.w *= k.w;
  }
  return t;
}

template <int NHT>
__device__ __forceinline__ void edge_probs(const EdgeFwdParams& P, int e, bool valid, const float (&st)[NHT],
                                           float gmax, int& src, float (&p)[NHT], unsigned& tiemask) {
  tiemask = 0;
  src = 0;
#pragma unroll
  for (int h = 0; h < NHT; ++h) p[h] = 0.f;
  if (!valid) return;
  src = __ldg(P.col + e);
  if (P.const_attention) {
#pragma unroll
    for (int h = 0; h < NHT; ++h) p[h] = h < P.nh ? 1.f : 0.f;   // exp(0), gat_layer.py:89-96
    return;
  }
  const float* ss = P.s_src + (int64_t)src * P.nh;
#pragma unroll
  for (int h = 0; h < NHT; ++h) {
    if (h < P.nh) {
      float l = __ldg(ss + h) + st[h];
      p[h] = attn_exp(l, gmax);
      if (l == gmax) tiemask |= 1u << h;
    }
  }
}

// One 4-element chunk of a gathered feature row.  BF16 = true (the opt-in "bf16 variant" of SURVEY.md 8-d): the gathered
// matrix is a bfloat16 copy of Wh (half the bytes per edge); the chunk is 8 bytes and is widened to fp32 in registers, all
// accumulation stays fp32.  `base` points at the matrix, `off` is the chunk's element offset inside the row.
template <bool BF16>
__device__ __forceinline__ float4 gather_chunk(const float* base, const int64_t row, const int dp, const int off) {
  if (BF16) {
    const uint16_t* p = reinterpret_cast<const uint16_t*>(base) + row * dp + off;
    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(p));
    return make_float4(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u),
                       __uint_as_float(raw.y << 16), __uint_as_float(raw.y & 0xffff0000u));
  }
  return ldg4(base + row * dp + off);
}

// COOP = false: the group owns the whole row.  COOP = true (long rows): the CTA's NG = 256/G groups take the row's
// batches round-robin; Z and the output row are combined over the groups through `coop` (NG*dp floats of dynamic
// shared memory) in group order.  Called by ALL threads of the CTA in that case.
template <int G, int SLOTS, int NHT, bool COOP, bool BF16 = false>
__device__ __forceinline__ void edge_fwd_row(const EdgeFwdParams& P, const int64_t row, const int start, const int end, const int tid, const int gl,
                                             const int gbase, const unsigned gmask, const float gmax,
                                             int* sh_src, float* sh_w, float* coop) {
  constexpr int U = SLOTS >= 4 ? 2 : (SLOTS >= 2 ? 4 : 8);
  constexpr int NG = kEdgeThreads / G;
  const int grp = tid / G;
  const int first = COOP ? grp * G : 0, step = COOP ? NG * G : G;
  const int nh = P.nh;
  int head[SLOTS], coff[SLOTS];
  bool ok[SLOTS];
  float4 acc[SLOTS];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    int c = s * G + gl;
    ok[s] = c < P.chunks;
    head[s] = ok[s] ? c / P.chunks_per_head : 0;
    coff[s] = ok[s] ? c * 4 : 0;      // float offset of the lane's chunk in a gathered row
    acc[s] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float st[NHT];
#pragma unroll
  for (int h = 0; h < NHT; ++h) st[h] = (!P.const_attention && h < nh) ? __ldg(P.s_tgt + row * nh + h) : 0.f;

  // ---- phase A: softmax denominators Z[h] = sum_e p[e,h]  (gat_layer.py:99-103)
  float p[NHT], z[NHT];
  int my_src = 0;
  unsigned tiemask = 0;
#pragma unroll
  for (int h = 0; h < NHT; ++h) z[h] = 0.f;
  for (int base = start + first; base < end; base += step) {
    edge_probs<NHT>(P, base + gl, base + gl < end, st, gmax, my_src, p, tiemask);
#pragma unroll
    for (int h = 0; h < NHT; ++h) z[h] += p[h];
  }
#pragma unroll
  for (int h = 0; h < NHT; ++h)
    if (h < nh) z[h] = group_sum<G>(z[h], gmask);
  if (COOP) {
    if (gl == 0) {
#pragma unroll
      for (int h = 0; h < NHT; ++h) coop[grp * NHT + h] = z[h];
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < NHT; ++h) {
      float t = 0.f;
      for (int j = 0; j < NG; ++j) t += coop[j * NHT + h];
      z[h] = t;
    }
    __syncthreads();   // coop is reused for the output row below
  }
  if (P.z_out && (COOP ? tid == 0 : gl == 0)) {
#pragma unroll
    for (int h = 0; h < NHT; ++h)
      if (h < nh) P.z_out[row * nh + h] = z[h];
  }
  const bool single = !COOP && (end - start) <= G;   // p[] of the only batch is still in registers

  // ---- phase B: alpha, dropout, weighted gather-accumulate  (gat_layer.py:106-127)
  for (int base = start + first; base < end; base += step) {
    const int e = base + gl;
    const bool valid = e < end;
    if (!single) edge_probs<NHT>(P, e, valid, st, gmax, my_src, p, tiemask);
    if (valid) {
      float w[NHT];
#pragma unroll
      for (int h = 0; h < NHT; ++h) w[h] = h < nh ? p[h] / (z[h] + kSoftmaxEps) : 0.f;
      int edge_id = 0;
      if (P.alpha_out || P.dropout_p > 0.f) edge_id = __ldg(P.eid + e);
      if (P.alpha_out) {
        float* ao = P.alpha_out + (int64_t)edge_id * nh;
#pragma unroll
        for (int h = 0; h < NHT; ++h)
          if (h < nh) ao[h] = w[h];
      }
      if (P.tie_total && tiemask) {
#pragma unroll
        for (int h = 0; h < NHT; ++h) {
          if (tiemask & (1u << h)) {
            atomicAdd(P.tie_dst + row * nh + h, 1);
            atomicAdd(P.tie_src + (int64_t)my_src * nh + h, 1);
            atomicAdd(P.tie_total, 1ull);
          }
        }
      }
      if (P.dropout_p > 0.f) {
        float m[NHT];
        dropout_scales<NHT>(P.seed, P.offset, (uint32_t)edge_id, nh, P.dropout_p, m);
#pragma unroll
        for (int h = 0; h < NHT; ++h)
          if (h < nh) w[h] *= m[h];
      }
      sh_src[tid] = my_src;
#pragma unroll
      for (int h = 0; h < NHT; ++h) sh_w[tid * NHT + h] = w[h];
    }
    __syncwarp(gmask);
    const int cnt = min(G, end - base);
    // A full group of U edges is one branch-free block (all loads, then all FMAs); a slot that does not exist (row
    // narrower than SLOTS*G chunks) reads chunk 0 of the row instead of being predicated per lane -- what it accumulates
    // is never stored -- so narrow rows (192 floats on a 256-float lane grid) run the same code as full ones.
#pragma unroll 1
    for (int t = 0; t < cnt; t += U) {
      float4 v[U][SLOTS];
      const int* sp = sh_src + gbase + t;
      const float* wp[SLOTS];
#pragma unroll
      for (int s = 0; s < SLOTS; ++s) wp[s] = sh_w + (gbase + t) * NHT + head[s];
      if (t + U <= cnt) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t srow = sp[u];
#pragma unroll
          for (int s = 0; s < SLOTS; ++s) v[u][s] = gather_chunk<BF16>(P.wh, srow, P.dp, coff[s]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int s = 0; s < SLOTS; ++s) {
            const float w = wp[s][u * NHT];
            acc[s].x = fmaf(w, v[u][s].x, acc[s].x);
            acc[s].y = fmaf(w, v[u][s].y, acc[s].y);
            acc[s].z = fmaf(w, v[u][s].z, acc[s].z);
            acc[s].w = fmaf(w, v[u][s].w, acc[s].w);
          }
        }
      } else {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const bool on = t + u < cnt;
          const int64_t srow = on ? sp[u] : 0;
#pragma unroll
          for (int s = 0; s < SLOTS; ++s)
            v[u][s] = on ? gather_chunk<BF16>(P.wh, srow, P.dp, coff[s]) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (t + u < cnt) {
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
              const float w = wp[s][u * NHT];
              acc[s].x = fmaf(w, v[u][s].x, acc[s].x);
              acc[s].y = fmaf(w, v[u][s].y, acc[s].y);
              acc[s].z = fmaf(w, v[u][s].z, acc[s].z);
              acc[s].w = fmaf(w, v[u][s].w, acc[s].w);
            }
          }
        }
      }
    }
    __syncwarp(gmask);
  }
  if (COOP) {
#pragma unroll
    for (int s = 0; s < SLOTS; ++s)
      if (ok[s]) *reinterpret_cast<float4*>(coop + grp * P.dp + (s * G + gl) * 4) = acc[s];
    __syncthreads();
    for (int c = tid; c < P.chunks; c += kEdgeThreads) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j = 0; j < NG; ++j) {
        const float4 v = *reinterpret_cast<const float4*>(coop + j * P.dp + c * 4);
        t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
      }
      *reinterpret_cast<float4*>(P.out + row * P.dp + c * 4) = out_glue(P, row, c, t);
    }
    // the next grab_long_row() starts with a __syncthreads(), which also protects `coop`
  } else {
#pragma unroll
    for (int s = 0; s < SLOTS; ++s)
      if (ok[s]) *reinterpret_cast<float4*>(P.out + row * P.dp + (s * G + gl) * 4) = out_glue(P, row, s * G + gl, acc[s]);
  }
}

// COOP = false: every warp takes short rows, group per row.  COOP = true: every CTA takes long rows, CTA per row (its
// own launch, so that neither path pays for the other's registers).
template <int G, int SLOTS, int NHT, bool COOP, bool BF16 = false>
__global__ void __launch_bounds__(kEdgeThreads, (SLOTS <= 2 ? 3 : (SLOTS <= 4 ? 2 : 1)))
edge_fwd_kernel(const EdgeFwdParams P) {
  extern __shared__ __align__(16) float coop[];   // COOP: (256/G) * dp floats, cross-group reduction
  __shared__ int sh_src[kEdgeThreads];
  __shared__ float sh_w[kEdgeThreads * NHT];
  const int tid = threadIdx.x, lane = tid & 31, gl = tid & (G - 1), gbase = tid - gl;
  const unsigned gmask = group_mask<G>(lane);
  const float gmax = P.const_attention ? 0.f : __ldg(P.gmax);
  if (COOP) {
    __shared__ int sh_ctl;
    pdl_release_dependents();   // the short-row launch that follows may fill SMs as this grid drains
    for (;;) {
      const int64_t row = grab_long_row(P.sched, P.rowptr, &sh_ctl);
      if (row < 0) break;
      edge_fwd_row<G, SLOTS, NHT, true, BF16>(P, row, __ldg(P.rowptr + row), __ldg(P.rowptr + row + 1), tid, gl, gbase, gmask, gmax, sh_src, sh_w, coop);
    }
  } else {
    int64_t base;
    while (grab_rows<G>(P.sched, lane, base)) {
      int pr, ps, pe;
      prefetch_rows<G>(P.sched, P.rowptr, base, lane, pr, ps, pe);
#pragma unroll 1
      for (int k = 0; k < kGrabIters<G>; ++k) {
        int64_t row;
        int start, end;
        if (prefetched_row<G>(P.sched, k, lane, pr, ps, pe, row, start, end))
          edge_fwd_row<G, SLOTS, NHT, false, BF16>(P, row, start, end, tid, gl, gbase, gmask, gmax, sh_src, sh_w, nullptr);
      }
    }
    pdl_wait_for_primary();     // no-op unless launched behind the cooperative kernel
  }
}

static size_t coop_smem_bytes(int g, int dp) { return (size_t)(kEdgeThreads / g) * dp * sizeof(float); }

// ------------------------------------------------------------------------------------------
// Head merge: padded (n, NH, Fp) -> (n, NH*F) or head mean (n, F)   (gat_layer.py:129-132)
// ------------------------------------------------------------------------------------------
// 2-D indexing (threadIdx.x walks the columns of a row, blockIdx.x / threadIdx.y the rows): no 64-bit division per element, which
// made these streaming kernels instruction bound (0.56 / 0.46 ms on the products output layer for 0.36 / 0.14 ms of traffic).
constexpr int kMergeRows = 8;      // rows per CTA (blockDim = 32 x 8)

__global__ void __launch_bounds__(256)
head_merge_fwd_kernel(const float* __restrict__ o, int64_t n, int nh, int f, int fp, int concat, float* __restrict__ out) {
  const int width = concat ? nh * f : f;
  for (int64_t i = (int64_t)blockIdx.x * kMergeRows + threadIdx.y; i < n; i += (int64_t)gridDim.x * kMergeRows) {
    const float* r = o + i * (int64_t)nh * fp;
    float* w = out + i * (int64_t)width;
    if (concat) {
      for (int h = 0; h < nh; ++h)
        for (int j = threadIdx.x; j < f; j += 32) w[h * f + j] = r[h * fp + j];
    } else {
      for (int c = threadIdx.x; c < f; c += 32) {
        float s = 0.f;
        for (int h = 0; h < nh; ++h) s += r[h * fp + c];
        w[c] = s / (float)nh;   // torch.mean(dim=1): sum then divide
      }
    }
  }
}

// Head merge with the output glue applied to the MERGED row (layers whose padded rows are not the caller's rows: head-mean layers,
// F % 4 != 0): out = keep * E(merge(o) + skip).
__global__ void head_merge_fwd_glue_kernel(const float* __restrict__ o, int64_t n, int nh, int f, int fp, int concat,
                                           const float* __restrict__ skip, int64_t ld_skip, int act, float drop_p, uint64_t drop_seed,
                                           float* __restrict__ out) {
  const int width = concat ? nh * f : f;
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= n * width) return;
  int64_t i = idx / width;
  int c = (int)(idx % width);
  const float* r = o + i * (int64_t)nh * fp;
  float v;
  if (concat) {
    v = r[(c / f) * fp + (c % f)];
  } else {
    float s = 0.f;
    for (int h = 0; h < nh; ++h) s += r[h * fp + c];
    v = s / (float)nh;
  }
  if (skip) v += skip[i * ld_skip + c];
  if (act) v = glue_elu(v);
  if (drop_p > 0.f) v *= glue_keep1(drop_seed, i, c, drop_p);
  out[idx] = v;
}

// dL/d(out + skip) from dL/dy and the stored y = keep * E(out + skip), element-wise over an (n, width) matrix
__global__ void out_glue_adjoint_kernel(const float* __restrict__ g, const float* __restrict__ y, int64_t n, int width, int act,
                                        float drop_p, uint64_t drop_seed, float* __restrict__ out) {
  const float omp = 1.0f - drop_p;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n * width; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / width;
    const int c = (int)(idx - i * width);
    const float keep = drop_p > 0.f ? glue_keep1(drop_seed, i, c, drop_p) : 1.0f;
    out[idx] = glue_adjoint1(g[idx], y[idx], act, keep, omp);
  }
}

// Head-mean layers: the adjoint of mean(dim=1) hands every head the same vector g/NH, so it is stored ONCE as a padded
// (n, fp) row that the backward kernels share across heads (go_shared) -- a quarter of the gather traffic at NH = 4.
__global__ void __launch_bounds__(256)
head_mean_bwd_shared_kernel(const float* __restrict__ g, int64_t n, int nh, int f, int fp, float* __restrict__ go) {
  for (int64_t i = (int64_t)blockIdx.x * kMergeRows + threadIdx.y; i < n; i += (int64_t)gridDim.x * kMergeRows) {
    const float* r = g + i * (int64_t)f;
    float* w = go + i * (int64_t)fp;
    for (int j = threadIdx.x; j < fp; j += 32) w[j] = j < f ? r[j] / (float)nh : 0.f;
  }
}

__global__ void head_merge_bwd_kernel(const float* __restrict__ g, int64_t n, int nh, int f, int fp, int concat,
                                      float* __restrict__ go) {
  const int dp = nh * fp;
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= n * dp) return;
  int64_t i = idx / dp;
  int c = (int)(idx % dp), h = c / fp, j = c % fp;
  float v = 0.f;
  if (j < f) v = concat ? g[i * (int64_t)nh * f + h * f + j] : g[i * (int64_t)f + j] / (float)nh;
  go[idx] = v;
}

}  // namespace gat

extern "C" int gat_edge_max(const int32_t* rowptr, const int32_t* col, const int32_t* row_order, int64_t n_long, int64_t n,
                            const float* s_src, const float* s_tgt, int nh, float* gmax,
                            void* workspace, size_t workspace_bytes, gat_stream_t stream) {
  using namespace gat;
  GAT_CHECK_ARG(nh >= 1 && nh <= kMaxHeads, "gat_edge_max: num_heads %d not in [1, %d]", nh, kMaxHeads);
  GAT_CHECK_ARG(workspace != nullptr && workspace_bytes >= 256, "gat_edge_max: workspace too small");
  if (n == 0) return GAT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // the row counters (warp-level, CTA-level) live at words 16/17 of the shared 256-byte workspace (gat_edge_fwd uses 0/1)
  unsigned int* counter = (unsigned int*)workspace + 16;
  GAT_CUDA(cudaMemsetAsync(counter, 0, 2 * sizeof(unsigned int), st));
  EdgeMaxParams P;
  (void)n_long;   // one kernel serves both phases here (no register pressure to protect)
  P.rowptr = rowptr; P.col = col; P.sched.order = row_order; P.sched.counter = counter; P.sched.cta_counter = counter + 1; P.sched.n = n;
  P.s_src = s_src; P.s_tgt = s_tgt; P.nh = nh; P.gmax = gmax;
  edge_max_kernel<<<persistent_grid(edge_max_kernel, 256, 0, (n + 7) / 8), 256, 0, st>>>(P);
  GAT_LAUNCH_CHECK();
  return GAT_OK;
}

extern "C" size_t gat_edge_fwd_workspace_bytes(void) { return 256; }

static int edge_fwd_impl(bool gather_bf16, const int32_t* rowptr, const int32_t* col, const int32_t* eid, const int32_t* row_order, int64_t n_long,
                         int64_t n, const float* wh, int nh, int fp, const float* s_src, const float* s_tgt,
                         const float* gmax, int const_attention, float dropout_p, uint64_t seed, uint64_t offset,
                         float* out, int out_act, float* alpha_out, float* z_out,
                         int32_t* tie_dst, int32_t* tie_src, unsigned long long* tie_total,
                         void* workspace, size_t workspace_bytes, gat_stream_t stream,
                         const float* skip = nullptr, int64_t ld_skip = 0, float out_drop_p = 0.f, uint64_t out_drop_seed = 0) {
  using namespace gat;
  GAT_CHECK_ARG(nh >= 1 && nh <= kMaxHeads, "gat_edge_fwd: num_heads %d not in [1, %d]", nh, kMaxHeads);
  GAT_CHECK_ARG(out_drop_p >= 0.f && out_drop_p < 1.f, "gat_edge_fwd: output dropout %f not in [0, 1)", out_drop_p);
  GAT_CHECK_ARG(skip == nullptr || (ld_skip >= (int64_t)nh * fp && ld_skip % 4 == 0 && ((uintptr_t)skip & 15) == 0),
                "gat_edge_fwd: skip rows must be 16-byte aligned with a stride of at least nh*fp floats");
  GAT_CHECK_ARG(workspace != nullptr && workspace_bytes >= gat_edge_fwd_workspace_bytes(), "gat_edge_fwd: workspace too small");
  GAT_CHECK_ARG(fp > 0 && fp % 4 == 0, "gat_edge_fwd: padded head width %d must be a positive multiple of 4", fp);
  GAT_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "gat_edge_fwd: dropout %f not in [0, 1)", dropout_p);
  GAT_CHECK_ARG(const_attention || (s_src && s_tgt && gmax), "gat_edge_fwd: score buffers missing");
  GAT_CHECK_ARG((tie_total == nullptr) == (tie_dst == nullptr) && (tie_total == nullptr) == (tie_src == nullptr),
                "gat_edge_fwd: tie buffers must be given together");
  if (n == 0) return GAT_OK;
  EdgeFwdParams P;
  P.rowptr = rowptr; P.col = col; P.eid = eid; P.wh = wh; P.nh = nh; P.dp = nh * fp;
  P.sched.order = row_order; P.sched.counter = (unsigned int*)workspace; P.sched.cta_counter = (unsigned int*)workspace + 1; P.sched.n = n;
  P.chunks = nh * fp / 4; P.chunks_per_head = fp / 4;
  P.s_src = s_src; P.s_tgt = s_tgt; P.gmax = gmax; P.const_attention = const_attention;
  P.dropout_p = dropout_p; P.seed = seed; P.offset = offset;
  P.out = out; P.alpha_out = alpha_out; P.z_out = z_out; P.out_act = out_act ? 1 : 0;
  P.skip = skip; P.ld_skip = ld_skip; P.out_drop_p = out_drop_p; P.out_drop_seed = out_drop_seed;
  P.tie_dst = const_attention ? nullptr : tie_dst; P.tie_src = const_attention ? nullptr : tie_src;
  P.tie_total = const_attention ? nullptr : tie_total;
  GroupShape shape = pick_group(P.chunks, n);
  if (shape.slots < 0) {
    set_error("gat_edge_fwd: row width %d floats exceeds the supported 1024", P.dp);
    return GAT_EUNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  GAT_CUDA(cudaMemsetAsync(workspace, 0, 2 * sizeof(unsigned int), st));
  // long rows first, CTA per row (skipped when the caller knows there are none; n_long < 0 = unknown); the short-row
  // launch overlaps its tail
  const bool coop_launch = row_order != nullptr && n_long != 0;
  if (gather_bf16) {   // bf16 variant: instantiated (and tested) for the shapes it is meant for: G = 32, 2 chunks per lane, NH <= 4
    if (!(shape.g == 32 && shape.slots == 2 && nh <= 4)) {
      set_error("gat_edge_fwd_bf16: supported for NH <= 4 and padded rows of 132..256 floats (got NH = %d, %d floats)", nh, P.dp);
      return GAT_EUNSUPPORTED;
    }
#define LAUNCH_BF16(S_)                                                                                               \
    do {                                                                                                              \
      if (coop_launch) {                                                                                              \
        GAT_CUDA(launch_kernel(edge_fwd_kernel<32, S_, 4, true, true>,                                                \
                               persistent_grid(edge_fwd_kernel<32, S_, 4, true, true>, kEdgeThreads, coop_smem_bytes(32, P.dp), \
                                               n_long < 0 ? n : n_long),                                              \
                               kEdgeThreads, coop_smem_bytes(32, P.dp), st, P, false));                               \
        GAT_LAUNCH_CHECK();                                                                                           \
      }                                                                                                               \
      GAT_CUDA(launch_kernel(edge_fwd_kernel<32, S_, 4, false, true>,                                                 \
                             persistent_grid(edge_fwd_kernel<32, S_, 4, false, true>, kEdgeThreads, 0, (n + 7) / 8),  \
                             kEdgeThreads, 0, st, P, coop_launch));                                                   \
      GAT_LAUNCH_CHECK();                                                                                             \
    } while (0)
    LAUNCH_BF16(2);
#undef LAUNCH_BF16
    return GAT_OK;
  }
  if (coop_launch) {
    const int64_t ctas = n_long < 0 ? n : n_long;
#define LAUNCH(G_, S_, N_)                                                                                            \
  GAT_CUDA(launch_kernel(edge_fwd_kernel<G_, S_, N_, true>,                                                           \
                         persistent_grid(edge_fwd_kernel<G_, S_, N_, true>, kEdgeThreads, coop_smem_bytes(G_, P.dp), ctas), \
                         kEdgeThreads, coop_smem_bytes(G_, P.dp), st, P, false))
    GAT_DISPATCH_GROUP(shape, nh, LAUNCH);
#undef LAUNCH
    GAT_LAUNCH_CHECK();
  }
#define LAUNCH(G_, S_, N_)                                                                                            \
  GAT_CUDA(launch_kernel(edge_fwd_kernel<G_, S_, N_, false>,                                                          \
                         persistent_grid(edge_fwd_kernel<G_, S_, N_, false>, kEdgeThreads, 0,                         \
                                         (n + (kEdgeThreads / G_) - 1) / (kEdgeThreads / G_)),                        \
                         kEdgeThreads, 0, st, P, coop_launch))
  GAT_DISPATCH_GROUP(shape, nh, LAUNCH);
#undef LAUNCH
  GAT_LAUNCH_CHECK();
  return GAT_OK;
}

extern "C" int gat_edge_fwd(const int32_t* rowptr, const int32_t* col, const int32_t* eid, const int32_t* row_order, int64_t n_long,
                            int64_t n, const float* wh, int nh, int fp, const float* s_src, const float* s_tgt,
                            const float* gmax, int const_attention, float dropout_p, uint64_t seed, uint64_t offset,
                            float* out, int out_act, float* alpha_out, float* z_out,
                            int32_t* tie_dst, int32_t* tie_src, unsigned long long* tie_total,
                            void* workspace, size_t workspace_bytes, gat_stream_t stream) {
  return edge_fwd_impl(false, rowptr, col, eid, row_order, n_long, n, wh, nh, fp, s_src, s_tgt, gmax, const_attention, dropout_p, seed,
                       offset, out, out_act, alpha_out, z_out, tie_dst, tie_src, tie_total, workspace, workspace_bytes, stream);
}

// gat_edge_fwd with the whole output glue (include/gat_b200.h): out = keep * E(out + skip)
extern "C" int gat_edge_fwd_glue(const int32_t* rowptr, const int32_t* col, const int32_t* eid, const int32_t* row_order, int64_t n_long,
                                 int64_t n, const float* wh, int nh, int fp, const float* s_src, const float* s_tgt,
                                 const float* gmax, int const_attention, float dropout_p, uint64_t seed, uint64_t offset,
                                 float* out, int out_act, const float* skip, int64_t ld_skip, float out_drop_p, uint64_t out_drop_seed,
                                 float* alpha_out, float* z_out,
                                 int32_t* tie_dst, int32_t* tie_src, unsigned long long* tie_total,
                                 void* workspace, size_t workspace_bytes, gat_stream_t stream) {
  return edge_fwd_impl(false, rowptr, col, eid, row_order, n_long, n, wh, nh, fp, s_src, s_tgt, gmax, const_attention, dropout_p, seed,
                       offset, out, out_act, alpha_out, z_out, tie_dst, tie_src, tie_total, workspace, workspace_bytes, stream,
                       skip, ld_skip, out_drop_p, out_drop_seed);
}

// bf16 variant: `wh_bf16` is a bfloat16 copy of Wh (gat_f32_to_bf16), same (n, nh*fp) row-major layout; everything else as above.
extern "C" int gat_edge_fwd_bf16(const int32_t* rowptr, const int32_t* col, const int32_t* eid, const int32_t* row_order, int64_t n_long,
                                 int64_t n, const void* wh_bf16, int nh, int fp, const float* s_src, const float* s_tgt,
                                 const float* gmax, int const_attention, float dropout_p, uint64_t seed, uint64_t offset,
                                 float* out, int out_act, float* alpha_out, float* z_out,
                                 int32_t* tie_dst, int32_t* tie_src, unsigned long long* tie_total,
                                 void* workspace, size_t workspace_bytes, gat_stream_t stream) {
  return edge_fwd_impl(true, rowptr, col, eid, row_order, n_long, n, (const float*)wh_bf16, nh, fp, s_src, s_tgt, gmax, const_attention,
                       dropout_p, seed, offset, out, out_act, alpha_out, z_out, tie_dst, tie_src, tie_total, workspace, workspace_bytes, stream);
}

extern "C" int gat_head_merge_fwd(const float* o_padded, int64_t n, int nh, int f, int fp, int concat, float* out,
                                  gat_stream_t stream) {
  using namespace gat;
  GAT_CHECK_ARG(nh >= 1 && f >= 1 && fp >= f, "gat_head_merge_fwd: bad shape");
  int64_t total = n * (concat ? nh * f : f);
  if (total == 0) return GAT_OK;
  const int64_t want_m = (n + kMergeRows - 1) / kMergeRows;
  head_merge_fwd_kernel<<<(unsigned)(want_m < kNumSMs * 16 ? want_m : kNumSMs * 16), dim3(32, kMergeRows), 0, (cudaStream_t)stream>>>(
      o_padded, n, nh, f, fp, concat, out);
  GAT_LAUNCH_CHECK();
  return GAT_OK;
}

extern "C" int gat_head_merge_fwd_glue(const float* o_padded, int64_t n, int nh, int f, int fp, int concat,
                                       const float* skip, int64_t ld_skip, int act, float drop_p, uint64_t drop_seed, float* out,
                                       gat_stream_t stream) {
  using namespace gat;
  GAT_CHECK_ARG(nh >= 1 && f >= 1 && fp >= f, "gat_head_merge_fwd_glue: bad shape");
  GAT_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "gat_head_merge_fwd_glue: dropout %f not in [0, 1)", drop_p);
  const int64_t width = concat ? (int64_t)nh * f : f;
  GAT_CHECK_ARG(skip == nullptr || ld_skip >= width, "gat_head_merge_fwd_glue: skip stride too small");
  const int64_t total = n * width;
  if (total == 0) return GAT_OK;
  head_merge_fwd_glue_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(o_padded, n, nh, f, fp, concat, skip, ld_skip,
                                                                                                 act ? 1 : 0, drop_p, drop_seed, out);
  GAT_LAUNCH_CHECK();
  return GAT_OK;
}

extern "C" int gat_out_glue_adjoint(const float* grad_y, const float* y, int64_t n, int width, int act, float drop_p, uint64_t drop_seed,
                                    float* grad_pre, gat_stream_t stream) {
  using namespace gat;
  GAT_CHECK_ARG(grad_y && y && grad_pre && width >= 1 && drop_p >= 0.f && drop_p < 1.f, "gat_out_glue_adjoint: bad arguments");
  const int64_t total = n * width;
  if (total == 0) return GAT_OK;
  int64_t blocks = (total + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  out_glue_adjoint_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(grad_y, y, n, width, act ? 1 : 0, drop_p, drop_seed, grad_pre);
  GAT_LAUNCH_CHECK();
  return GAT_OK;
}

extern "C" int gat_head_merge_bwd(const float* grad_out, int64_t n, int nh, int f, int fp, int concat, float* go_padded,
                                  gat_stream_t stream) {
  using namespace gat;
  GAT_CHECK_ARG(nh >= 1 && f >= 1 && fp >= f, "gat_head_merge_bwd: bad shape");
  int64_t total = n * (int64_t)nh * fp;
  if (total == 0) return GAT_OK;
  head_merge_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(grad_out, n, nh, f, fp, concat, go_padded);
  GAT_LAUNCH_CHECK();
  return GAT_OK;
}

extern "C" int gat_head_mean_bwd_shared(const float* grad_out, int64_t n, int nh, int f, int fp, float* go_shared,
                                        gat_stream_t stream) {
  using namespace gat;
  GAT_CHECK_ARG(nh >= 1 && f >= 1 && fp >= f && fp % 4 == 0, "gat_head_mean_bwd_shared: bad shape");
  int64_t total = n * (int64_t)fp;
  if (total == 0) return GAT_OK;
  const int64_t want_m = (n + kMergeRows - 1) / kMergeRows;
  head_mean_bwd_shared_kernel<<<(unsigned)(want_m < kNumSMs * 16 ? want_m : kNumSMs * 16), dim3(32, kMergeRows), 0, (cudaStream_t)stream>>>(
      grad_out, n, nh, f, fp, go_shared);
  GAT_LAUNCH_CHECK();
  return GAT_OK;
}
