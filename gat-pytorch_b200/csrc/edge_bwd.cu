// Kernel 4: atomic-free deterministic backward of the edge stage (autograd of gat_layer.py:70-132;
// formulas in SURVEY.md section 9.2).  Two passes:
//   dst pass over the target-sorted CSR  -> per-edge records {g, m*alpha}, ds_tgt, partial sums of g
//   src pass over the source-sorted CSR  -> d_wh (total), ds_src, arg-max correction
// Every sum is a fixed-order register / shuffle / shared-memory reduction; no floating-point atomics.
#include "edge_common.cuh"
#include <cstddef>

namespace gat {

struct BwdHeader {          // first 256 bytes of the backward workspace
  double gamma;             // sum over all (e,h) of g
  float corr;               // gamma / |T|   (0 when the arg-max set is empty)
  int n_partials;           // number of per-block partials written by the dst pass
  unsigned int pad0[12];
  unsigned int counter_dst; // row scheduler of the dst pass   (byte offset 64)
  unsigned int pad1[15];
  unsigned int counter_src; // row scheduler of the src pass   (byte offset 128)
};
static_assert(offsetof(BwdHeader, counter_dst) == 64 && offsetof(BwdHeader, counter_src) == 128, "header layout");
constexpr size_t kBwdHeaderBytes = 256;

struct EdgeBwdDstParams {
  const int32_t* rowptr; const int32_t* col; const int32_t* eid; RowSched sched;
  const float* wh; int nh; int dp; int chunks; int chunks_per_head;
  const float* s_src; const float* s_tgt; const float* gmax; const float* z;
  int const_attention; float dropout_p; uint64_t seed; uint64_t offset;
  const float* go; const float* grad_alpha;
  float* rec; float* ds_tgt;
};

template <int G, int SLOTS>
struct DstShape {
  static constexpr int TB = (G < 8) ? G : (SLOTS >= 6 ? 4 : 8);     // edges per transpose-reduce sub-batch
  static constexpr int U = (SLOTS >= 4) ? 2 : (TB < 4 ? TB : 4);    // edges in flight
};

template <int G, int SLOTS, int NHT>
__device__ __forceinline__ void edge_bwd_dst_row(const EdgeBwdDstParams& P, const int64_t row, const int tid, const int gl,
                                                 const int gbase, const unsigned gmask, const float gmax,
                                                 int* sh_src, float* sh_da, float* part) {
  constexpr int TB = DstShape<G, SLOTS>::TB, U = DstShape<G, SLOTS>::U;
  const int nh = P.nh;
  const int pstride = P.chunks + 1;
  bool ok[SLOTS];
  float4 go[SLOTS];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    int c = s * G + gl;
    ok[s] = c < P.chunks;
    go[s] = ok[s] ? ldg4(P.go + row * P.dp + c * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int start = __ldg(P.rowptr + row), end = __ldg(P.rowptr + row + 1);
  float st[NHT], z[NHT];
#pragma unroll
  for (int h = 0; h < NHT; ++h) {
    st[h] = (!P.const_attention && h < nh) ? __ldg(P.s_tgt + row * nh + h) : 0.f;
    z[h] = h < nh ? __ldg(P.z + row * nh + h) : 0.f;
  }
  const bool single = (end - start) <= G;
  float alpha[NHT], dal[NHT], msk[NHT], ssum[NHT];
#pragma unroll
  for (int h = 0; h < NHT; ++h) { alpha[h] = 0.f; dal[h] = 0.f; msk[h] = 1.f; ssum[h] = 0.f; }

  // ---- pass 1: d_alpha[e,h] = m * <go[i,h,:], wh[src,h,:]> + grad_alpha;  S[h] = sum alpha*d_alpha
  for (int base = start; base < end; base += G) {
    const int e = base + gl;
    const bool valid = e < end;
    int my_src = 0;
    float ga[NHT];
#pragma unroll
    for (int h = 0; h < NHT; ++h) { alpha[h] = 0.f; msk[h] = 1.f; ga[h] = 0.f; }
    if (valid) {
      my_src = __ldg(P.col + e);
      if (P.const_attention) {
#pragma unroll
        for (int h = 0; h < NHT; ++h) alpha[h] = h < nh ? 1.f / (z[h] + kSoftmaxEps) : 0.f;
      } else {
        const float* ss = P.s_src + (int64_t)my_src * nh;
#pragma unroll
        for (int h = 0; h < NHT; ++h)
          if (h < nh) alpha[h] = attn_exp(__ldg(ss + h) + st[h], gmax) / (z[h] + kSoftmaxEps);
      }
      if (P.dropout_p > 0.f || P.grad_alpha) {
        const int edge_id = __ldg(P.eid + e);
        if (P.dropout_p > 0.f) dropout_scales<NHT>(P.seed, P.offset, (uint32_t)edge_id, nh, P.dropout_p, msk);
        if (P.grad_alpha) {
#pragma unroll
          for (int h = 0; h < NHT; ++h)
            if (h < nh) ga[h] = __ldg(P.grad_alpha + (int64_t)edge_id * nh + h);
        }
      }
      sh_src[tid] = my_src;
    }
    __syncwarp(gmask);
    const int cnt = min(G, end - base);
    for (int t0 = 0; t0 < cnt; t0 += TB) {
      const int tcnt = min(TB, cnt - t0);
#pragma unroll
      for (int tt = 0; tt < TB; tt += U) {
        float4 v[U][SLOTS];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const bool on = tt + u < tcnt;
          const int sidx = on ? sh_src[gbase + t0 + tt + u] : 0;
          const float* rowp = P.wh + (int64_t)sidx * P.dp + gl * 4;
#pragma unroll
          for (int s = 0; s < SLOTS; ++s)
            v[u][s] = (on && ok[s]) ? ldg4(rowp + s * G * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (tt + u < tcnt) {
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
              if (ok[s]) {
                float d = go[s].x * v[u][s].x;
                d = fmaf(go[s].y, v[u][s].y, d);
                d = fmaf(go[s].z, v[u][s].z, d);
                d = fmaf(go[s].w, v[u][s].w, d);
                part[(tt + u) * pstride + s * G + gl] = d;
              }
            }
          }
        }
      }
      __syncwarp(gmask);
      // transpose-reduce: (edge, head) pair q sums the chunks of that head
      for (int q = gl; q < tcnt * nh; q += G) {
        const int t = q / nh, h = q - t * nh;
        const float* pp = part + t * pstride + h * P.chunks_per_head;
        float d = 0.f;
        for (int c = 0; c < P.chunks_per_head; ++c) d += pp[c];
        sh_da[(gbase + t0 + t) * NHT + h] = d;
      }
      __syncwarp(gmask);
    }
    if (valid) {
#pragma unroll
      for (int h = 0; h < NHT; ++h) {
        if (h < nh) {
          dal[h] = fmaf(msk[h], sh_da[tid * NHT + h], ga[h]);
          ssum[h] = fmaf(alpha[h], dal[h], ssum[h]);
        }
      }
      if (!single) {   // stage {d_alpha, alpha} in the record slot; finalised in pass 2
        float* r = P.rec + (int64_t)e * 2 * nh;
#pragma unroll
        for (int h = 0; h < NHT; ++h)
          if (h < nh) { r[h] = dal[h]; r[nh + h] = alpha[h]; }
      }
    }
    __syncwarp(gmask);
  }
#pragma unroll
  for (int h = 0; h < NHT; ++h)
    if (h < nh) ssum[h] = group_sum<G>(ssum[h], gmask);

  // ---- pass 2: g = slope * alpha * (d_alpha - S);  record {g, m*alpha};  ds_tgt = sum g
  float gsum[NHT];
#pragma unroll
  for (int h = 0; h < NHT; ++h) gsum[h] = 0.f;
  for (int base = start; base < end; base += G) {
    const int e = base + gl;
    if (e < end) {
      float* r = P.rec + (int64_t)e * 2 * nh;
      if (!single) {
#pragma unroll
        for (int h = 0; h < NHT; ++h)
          if (h < nh) { dal[h] = r[h]; alpha[h] = r[nh + h]; }
        if (P.dropout_p > 0.f) dropout_scales<NHT>(P.seed, P.offset, (uint32_t)__ldg(P.eid + e), nh, P.dropout_p, msk);
      }
#pragma unroll
      for (int h = 0; h < NHT; ++h) {
        if (h < nh) {
          // LeakyReLU'(l - M) = 0.01 everywhere: l - M <= 0, and torch uses the slope at exactly 0
          const float g = P.const_attention ? 0.f : kLeakySlope * alpha[h] * (dal[h] - ssum[h]);
          r[h] = g;
          r[nh + h] = msk[h] * alpha[h];
          gsum[h] += g;
        }
      }
    }
  }
  if (!P.const_attention) {
#pragma unroll
    for (int h = 0; h < NHT; ++h)
      if (h < nh) gsum[h] = group_sum<G>(gsum[h], gmask);
    if (gl == 0) {
#pragma unroll
      for (int h = 0; h < NHT; ++h)
        if (h < nh) P.ds_tgt[row * nh + h] = gsum[h];
    }
  }
}

template <int G, int SLOTS, int NHT>
__global__ void __launch_bounds__(kEdgeThreads, (SLOTS <= 2 ? 3 : (SLOTS <= 4 ? 2 : 1)))
edge_bwd_dst_kernel(const EdgeBwdDstParams P) {
  constexpr int TB = DstShape<G, SLOTS>::TB;
  extern __shared__ float dyn_smem[];
  __shared__ int sh_src[kEdgeThreads];
  __shared__ float sh_da[kEdgeThreads * NHT];
  const int tid = threadIdx.x, lane = tid & 31, gl = tid & (G - 1), gbase = tid - gl;
  const unsigned gmask = group_mask<G>(lane);
  float* part = dyn_smem + (size_t)(tid / G) * TB * (P.chunks + 1);   // [TB][chunks+1] of my group
  const float gmax = P.const_attention ? 0.f : __ldg(P.gmax);
  int64_t base;
  while (grab_rows<G>(P.sched, lane, base)) {
#pragma unroll 1
    for (int k = 0; k < (G == 32 ? 4 : 2); ++k) {
      const int64_t row = sched_row<G>(P.sched, base, k, lane);
      if (row >= 0) edge_bwd_dst_row<G, SLOTS, NHT>(P, row, tid, gl, gbase, gmask, gmax, sh_src, sh_da, part);
    }
  }
}

// Gamma = sum over all (row, head) of ds_tgt (each entry is already the fixed-order sum of g over a row), reduced
// in two fixed-order stages so the result does not depend on how the persistent dst pass was scheduled.
constexpr int kGammaBlocks = 592;   // 4 per SM
__global__ void __launch_bounds__(256)
gamma_partial_kernel(const float* __restrict__ ds_tgt, int64_t count, BwdHeader* header, double* __restrict__ partials) {
  __shared__ double sh[256];
  const int64_t per = (count + kGammaBlocks - 1) / kGammaBlocks;
  const int64_t lo = (int64_t)blockIdx.x * per, hi = min(count, lo + per);
  double t = 0.0;
  for (int64_t i = lo + threadIdx.x; i < hi; i += 256) t += (double)ds_tgt[i];
  sh[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = sh[0];
    if (blockIdx.x == 0) header->n_partials = kGammaBlocks;
  }
}

// Gamma = sum of the dst-pass partials (fixed order); corr = Gamma / |T|.
__global__ void __launch_bounds__(1024)
gamma_finalize_kernel(BwdHeader* header, const double* __restrict__ partials, const unsigned long long* __restrict__ tie_total) {
  __shared__ double sh[1024];
  const int n = header->n_partials;
  double t = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) t += partials[i];
  sh[threadIdx.x] = t;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    header->gamma = sh[0];
    unsigned long long ties = tie_total ? *tie_total : 0ull;
    header->corr = ties ? (float)(sh[0] / (double)ties) : 0.f;
  }
}

struct EdgeBwdSrcParams {
  const int32_t* rowptr_t; const int32_t* col_t; const int32_t* pos_t; RowSched sched;
  int nh; int dp; int chunks; int chunks_per_head;
  const float* rec; const float* go; const float* a_src; const float* a_tgt; int const_attention;
  const int32_t* tie_dst; const int32_t* tie_src; const BwdHeader* header; const float* corr_override;
  int64_t tgt_lo; int64_t tgt_hi;   // rows that this call owns as TARGETS (ds_tgt / tie_dst are indexed row - tgt_lo)
  float* ds_src; float* ds_tgt; float* d_wh;
};

template <int G, int SLOTS, int NHT>
__device__ __forceinline__ void edge_bwd_src_row(const EdgeBwdSrcParams& P, const int64_t row, const int tid, const int gl,
                                                 const int gbase, const unsigned gmask, const float corr,
                                                 int* sh_dst, float* sh_w) {
  constexpr int U = SLOTS >= 4 ? 2 : (SLOTS >= 2 ? 4 : 8);
  const int nh = P.nh;
  int head[SLOTS];
  bool ok[SLOTS];
  float4 acc[SLOTS];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    int c = s * G + gl;
    ok[s] = c < P.chunks;
    head[s] = ok[s] ? c / P.chunks_per_head : 0;
    acc[s] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int start = __ldg(P.rowptr_t + row), end = __ldg(P.rowptr_t + row + 1);
  float gsum[NHT];
#pragma unroll
  for (int h = 0; h < NHT; ++h) gsum[h] = 0.f;

  for (int base = start; base < end; base += G) {
    const int e = base + gl;
    if (e < end) {
      sh_dst[tid] = __ldg(P.col_t + e);
      const float* r = P.rec + (int64_t)__ldg(P.pos_t + e) * 2 * nh;
#pragma unroll
      for (int h = 0; h < NHT; ++h) {
        float g = 0.f, w = 0.f;
        if (h < nh) { g = __ldg(r + h); w = __ldg(r + nh + h); }
        gsum[h] += g;
        sh_w[tid * NHT + h] = w;
      }
    }
    __syncwarp(gmask);
    const int cnt = min(G, end - base);
    for (int t = 0; t < cnt; t += U) {
      float4 v[U][SLOTS];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool on = t + u < cnt;
        const int didx = on ? sh_dst[gbase + t + u] : 0;
        const float* rowp = P.go + (int64_t)didx * P.dp + gl * 4;
#pragma unroll
        for (int s = 0; s < SLOTS; ++s)
          v[u][s] = (on && ok[s]) ? ldg4(rowp + s * G * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (t + u < cnt) {
#pragma unroll
          for (int s = 0; s < SLOTS; ++s) {
            const float w = sh_w[(gbase + t + u) * NHT + head[s]];
            acc[s].x = fmaf(w, v[u][s].x, acc[s].x);
            acc[s].y = fmaf(w, v[u][s].y, acc[s].y);
            acc[s].z = fmaf(w, v[u][s].z, acc[s].z);
            acc[s].w = fmaf(w, v[u][s].w, acc[s].w);
          }
        }
      }
    }
    __syncwarp(gmask);
  }

  if (!P.const_attention) {
    // ds_src = sum g - |T_src|*Gamma/|T|;  ds_tgt -= |T_dst|*Gamma/|T|   (gradient through max(), section 9.2)
    const bool own_tgt = row >= P.tgt_lo && row < P.tgt_hi;   // single GPU: always; partitioned: owner rank only
    const int64_t trow = row - P.tgt_lo;
    float dss[NHT], dst_[NHT];
#pragma unroll
    for (int h = 0; h < NHT; ++h) {
      dss[h] = 0.f; dst_[h] = 0.f;
      if (h < nh) {
        float g = group_sum<G>(gsum[h], gmask);
        int ts = P.tie_src ? __ldg(P.tie_src + row * nh + h) : 0;
        dss[h] = ts ? g - (float)ts * corr : g;
        if (own_tgt) {
          int td = P.tie_dst ? __ldg(P.tie_dst + trow * nh + h) : 0;
          float t = P.ds_tgt[trow * nh + h];
          dst_[h] = td ? t - (float)td * corr : t;
        }
      }
    }
    __syncwarp(gmask);   // every lane has read ds_tgt before lane 0 overwrites it
    if (gl == 0) {
#pragma unroll
      for (int h = 0; h < NHT; ++h)
        if (h < nh) {
          P.ds_src[row * nh + h] = dss[h];
          if (own_tgt) P.ds_tgt[trow * nh + h] = dst_[h];
        }
    }
    // d_wh_total = d_wh + ds_src * A_src + ds_tgt * A_tgt
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      if (ok[s]) {
        const int c4 = (s * G + gl) * 4;
#pragma unroll
        for (int h = 0; h < NHT; ++h) {
          if (h < nh) {
            const float4 as = ldg4(P.a_src + (int64_t)h * P.dp + c4);
            const float4 at = ldg4(P.a_tgt + (int64_t)h * P.dp + c4);
            acc[s].x = fmaf(dss[h], as.x, fmaf(dst_[h], at.x, acc[s].x));
            acc[s].y = fmaf(dss[h], as.y, fmaf(dst_[h], at.y, acc[s].y));
            acc[s].z = fmaf(dss[h], as.z, fmaf(dst_[h], at.z, acc[s].z));
            acc[s].w = fmaf(dss[h], as.w, fmaf(dst_[h], at.w, acc[s].w));
          }
        }
      }
    }
  }
#pragma unroll
  for (int s = 0; s < SLOTS; ++s)
    if (ok[s]) *reinterpret_cast<float4*>(P.d_wh + row * P.dp + (s * G + gl) * 4) = acc[s];
}

template <int G, int SLOTS, int NHT>
__global__ void __launch_bounds__(kEdgeThreads, (SLOTS <= 2 ? 3 : (SLOTS <= 4 ? 2 : 1)))
edge_bwd_src_kernel(const EdgeBwdSrcParams P) {
  __shared__ int sh_dst[kEdgeThreads];
  __shared__ float sh_w[kEdgeThreads * NHT];
  const int tid = threadIdx.x, lane = tid & 31, gl = tid & (G - 1), gbase = tid - gl;
  const unsigned gmask = group_mask<G>(lane);
  float corr = 0.f;
  if (!P.const_attention) corr = P.corr_override ? __ldg(P.corr_override) : P.header->corr;
  int64_t base;
  while (grab_rows<G>(P.sched, lane, base)) {
#pragma unroll 1
    for (int k = 0; k < (G == 32 ? 4 : 2); ++k) {
      const int64_t row = sched_row<G>(P.sched, base, k, lane);
      if (row >= 0) edge_bwd_src_row<G, SLOTS, NHT>(P, row, tid, gl, gbase, gmask, corr, sh_dst, sh_w);
    }
  }
}

template <int G, int SLOTS>
static size_t dst_dyn_smem(int chunks) {
  return (size_t)(kEdgeThreads / G) * DstShape<G, SLOTS>::TB * (chunks + 1) * sizeof(float);
}

}  // namespace gat

extern "C" size_t gat_edge_bwd_workspace_bytes(int64_t n, int64_t n_edges, int nh) {
  (void)n; (void)n_edges; (void)nh;
  return gat::kBwdHeaderBytes + (size_t)(gat::kGammaBlocks + 1) * sizeof(double);
}

extern "C" int gat_edge_bwd_dst(const int32_t* rowptr, const int32_t* col, const int32_t* eid, const int32_t* row_order, int64_t n,
                                const float* wh, int nh, int fp, const float* s_src, const float* s_tgt,
                                const float* gmax, const float* z, int const_attention,
                                float dropout_p, uint64_t seed, uint64_t offset,
                                const float* go_padded, const float* grad_alpha,
                                float* rec, float* ds_tgt, void* workspace, size_t workspace_bytes,
                                gat_stream_t stream) {
  using namespace gat;
  GAT_CHECK_ARG(nh >= 1 && nh <= kMaxHeads, "gat_edge_bwd_dst: num_heads %d not in [1, %d]", nh, kMaxHeads);
  GAT_CHECK_ARG(fp > 0 && fp % 4 == 0, "gat_edge_bwd_dst: padded head width %d must be a positive multiple of 4", fp);
  GAT_CHECK_ARG(const_attention || (s_src && s_tgt && gmax), "gat_edge_bwd_dst: score buffers missing");
  if (workspace == nullptr || workspace_bytes < gat_edge_bwd_workspace_bytes(n, 0, nh)) {
    set_error("gat_edge_bwd_dst: workspace too small");
    return GAT_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  GAT_CUDA(cudaMemsetAsync(workspace, 0, kBwdHeaderBytes, st));
  if (n == 0) return GAT_OK;
  EdgeBwdDstParams P;
  P.rowptr = rowptr; P.col = col; P.eid = eid; P.wh = wh; P.nh = nh; P.dp = nh * fp;
  P.sched.order = row_order; P.sched.counter = &((BwdHeader*)workspace)->counter_dst; P.sched.n = n;
  P.chunks = nh * fp / 4; P.chunks_per_head = fp / 4;
  P.s_src = s_src; P.s_tgt = s_tgt; P.gmax = gmax; P.z = z; P.const_attention = const_attention;
  P.dropout_p = dropout_p; P.seed = seed; P.offset = offset; P.go = go_padded; P.grad_alpha = grad_alpha;
  P.rec = rec; P.ds_tgt = ds_tgt;
  GroupShape shape = pick_group(P.chunks);
  if (shape.slots < 0) {
    set_error("gat_edge_bwd_dst: row width %d floats exceeds the supported 1024", P.dp);
    return GAT_EUNSUPPORTED;
  }
#define LAUNCH(G_, S_, N_)                                                                               \
  edge_bwd_dst_kernel<G_, S_, N_><<<persistent_grid(edge_bwd_dst_kernel<G_, S_, N_>, kEdgeThreads,               \
                                                    dst_dyn_smem<G_, S_>(P.chunks),                              \
                                                    (n + (kEdgeThreads / G_) - 1) / (kEdgeThreads / G_)),        \
                                    kEdgeThreads, dst_dyn_smem<G_, S_>(P.chunks), st>>>(P)
  GAT_DISPATCH_GROUP(shape, nh, LAUNCH);
#undef LAUNCH
  GAT_LAUNCH_CHECK();
  if (!const_attention) {
    gamma_partial_kernel<<<kGammaBlocks, 256, 0, st>>>(ds_tgt, n * nh, (BwdHeader*)workspace,
                                                       (double*)((char*)workspace + kBwdHeaderBytes));
    GAT_LAUNCH_CHECK();
  }
  return GAT_OK;
}

extern "C" int gat_edge_bwd_gamma(void* workspace, size_t workspace_bytes, double* gamma_out, gat_stream_t stream) {
  using namespace gat;
  GAT_CHECK_ARG(workspace && gamma_out && workspace_bytes >= kBwdHeaderBytes, "gat_edge_bwd_gamma: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  BwdHeader* header = (BwdHeader*)workspace;
  gamma_finalize_kernel<<<1, 1024, 0, st>>>(header, (const double*)((char*)workspace + kBwdHeaderBytes), nullptr);
  GAT_LAUNCH_CHECK();
  GAT_CUDA(cudaMemcpyAsync(gamma_out, &header->gamma, sizeof(double), cudaMemcpyDeviceToDevice, st));
  return GAT_OK;
}

extern "C" int gat_edge_bwd_src(const int32_t* rowptr_t, const int32_t* col_t, const int32_t* pos_t, const int32_t* row_order_t, int64_t n,
                                int nh, int fp, const float* rec, const float* go_padded,
                                const float* a_src, const float* a_tgt, int const_attention,
                                const int32_t* tie_dst, const int32_t* tie_src, const unsigned long long* tie_total,
                                const float* corr_override, int64_t tgt_lo, int64_t tgt_hi,
                                float* ds_src, float* ds_tgt, float* d_wh,
                                void* workspace, size_t workspace_bytes, gat_stream_t stream) {
  using namespace gat;
  GAT_CHECK_ARG(nh >= 1 && nh <= kMaxHeads, "gat_edge_bwd_src: num_heads %d not in [1, %d]", nh, kMaxHeads);
  GAT_CHECK_ARG(fp > 0 && fp % 4 == 0, "gat_edge_bwd_src: padded head width %d must be a positive multiple of 4", fp);
  GAT_CHECK_ARG(const_attention || (a_src && a_tgt && ds_src && ds_tgt), "gat_edge_bwd_src: attention buffers missing");
  if (workspace == nullptr || workspace_bytes < gat_edge_bwd_workspace_bytes(n, 0, nh)) {
    set_error("gat_edge_bwd_src: workspace too small");
    return GAT_EWORKSPACE;
  }
  if (n == 0) return GAT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  BwdHeader* header = (BwdHeader*)workspace;
  if (!const_attention && corr_override == nullptr) {
    gamma_finalize_kernel<<<1, 1024, 0, st>>>(header, (const double*)((char*)workspace + kBwdHeaderBytes), tie_total);
    GAT_LAUNCH_CHECK();
  }
  EdgeBwdSrcParams P;
  GAT_CUDA(cudaMemsetAsync(&header->counter_src, 0, sizeof(unsigned int), st));
  P.rowptr_t = rowptr_t; P.col_t = col_t; P.pos_t = pos_t; P.nh = nh; P.dp = nh * fp;
  P.sched.order = row_order_t; P.sched.counter = &header->counter_src; P.sched.n = n;
  P.chunks = nh * fp / 4; P.chunks_per_head = fp / 4;
  P.rec = rec; P.go = go_padded; P.a_src = a_src; P.a_tgt = a_tgt; P.const_attention = const_attention;
  P.tie_dst = tie_dst; P.tie_src = tie_src; P.header = header; P.corr_override = corr_override;
  P.tgt_lo = tgt_lo; P.tgt_hi = tgt_hi;
  P.ds_src = ds_src; P.ds_tgt = ds_tgt; P.d_wh = d_wh;
  GroupShape shape = pick_group(P.chunks);
  if (shape.slots < 0) {
    set_error("gat_edge_bwd_src: row width %d floats exceeds the supported 1024", P.dp);
    return GAT_EUNSUPPORTED;
  }
#define LAUNCH(G_, S_, N_)                                                                      \
  edge_bwd_src_kernel<G_, S_, N_><<<persistent_grid(edge_bwd_src_kernel<G_, S_, N_>, kEdgeThreads, 0,   \
                                                    (n + (kEdgeThreads / G_) - 1) / (kEdgeThreads / G_)), kEdgeThreads, 0, st>>>(P)
  GAT_DISPATCH_GROUP(shape, nh, LAUNCH);
#undef LAUNCH
  GAT_LAUNCH_CHECK();
  return GAT_OK;
}
