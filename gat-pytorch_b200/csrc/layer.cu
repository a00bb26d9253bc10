// One GATLayer forward / backward as ONE C-ABI call each (gat_layer_fwd / gat_layer_bwd): the whole kernel sequence of
// gat_layer.py:42-140 and of its autograd is issued from here, out of two caller-provided arenas, with the reference's own
// parameter layouts (W (NH*F, F_in), a (NH, NH*2F)) on both sides.
//
// Why: on the small named graphs (Cora, Pubmed, PPI, PATTERN) a layer is 8-10 kernels of 3-10 us each, and the Python
// host path between them (one ctypes call, a few torch.empty / view / pad / contiguous ops and their autograd per kernel)
// took longer than the kernels: 1.5 ms per Cora step at 38 % kernel time.  From C a launch costs ~2 us, so the GPU stays
// fed.  The per-kernel entry points remain the boundary for callers that want them (the partitioned layer, bench.py's
// per-kernel timing); this file only composes them.
#include "common.cuh"
#include <math.h>

namespace gat {

namespace {

struct Bump {           // sizes a plan (base == nullptr) or carves it out of an arena; 256-byte granules
  char* base;
  size_t off;
  template <typename T>
  T* take(size_t count) {
    const size_t bytes = align_up(count * sizeof(T), 256);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += bytes;
    return p;
  }
};

struct FwdPlan {        // everything the backward needs again lives here (the "saved tensors" of the layer)
  float *w_p, *w_pT, *a_src_p, *a_tgt_p, *wh, *s_src, *s_tgt, *gmax, *z, *out_p;
  float* norm_t;        // (n, nh) per-target sums of the fused attention-norm regulariser (desc->norm_out given)
  void* nws; size_t nws_bytes;
  int32_t* ties;        // [tie_total (2 x int32 = one uint64) | tie_dst (n*nh) | tie_src (n*nh)]
  void* fws; size_t fws_bytes;
  void* gws; size_t gws_bytes;
  bool out_is_user;     // out_p aliases the caller's `out` (concat layer, F % 4 == 0): no merge kernel, nothing stored twice
};

bool needs_merge(const gat_layer_desc* d) { return d->fp != d->f || !d->concat; }
// output glue (include/gat_b200.h): applied in the edge kernel's epilogue when the padded rows are the caller's rows, else
// to the merged rows by the head-merge kernel
bool has_glue(const gat_layer_desc* d) { return d->out_act || d->skip != nullptr || d->out_drop_p > 0.f; }
int64_t out_width(const gat_layer_desc* d) { return d->concat ? (int64_t)d->nh * d->f : d->f; }

void plan_fwd(const gat_layer_desc* d, Bump& b, FwdPlan& p) {
  const int64_t n = d->n, dp = (int64_t)d->nh * d->fp;
  p.w_p = d->fp != d->f ? b.take<float>((size_t)dp * d->f_in) : nullptr;
  p.w_pT = b.take<float>((size_t)dp * d->f_in);
  p.a_src_p = p.a_tgt_p = p.s_src = p.s_tgt = p.gmax = nullptr;
  p.ties = nullptr;
  if (!d->const_attention) {
    p.a_src_p = b.take<float>((size_t)d->nh * dp);
    p.a_tgt_p = b.take<float>((size_t)d->nh * dp);
    p.s_src = b.take<float>((size_t)n * d->nh);
    p.s_tgt = b.take<float>((size_t)n * d->nh);
    p.gmax = b.take<float>(64);
    p.ties = b.take<int32_t>((size_t)2 * n * d->nh + 2);
  }
  p.wh = b.take<float>((size_t)n * dp);
  p.z = b.take<float>((size_t)n * d->nh);
  p.out_is_user = !needs_merge(d);
  p.out_p = p.out_is_user ? nullptr : b.take<float>((size_t)n * dp);
  // always planned (n*nh floats), so that the arena layout does not depend on whether the norm was asked for
  p.norm_t = b.take<float>((size_t)n * d->nh);
  p.nws_bytes = gat_attention_norm_workspace_bytes();
  p.nws = b.take<char>(p.nws_bytes);
  // forward-only scratch at the tail
  p.fws_bytes = gat_edge_fwd_workspace_bytes();
  p.fws = b.take<char>(p.fws_bytes);
  p.gws_bytes = gat_gemm_workspace_bytes(0, 1, n, dp, d->f_in, d->gemm_algo);
  p.gws = p.gws_bytes ? b.take<char>(p.gws_bytes) : nullptr;
}

struct BwdPlan {
  float *go_p, *d_wh, *ds_src, *ds_tgt, *s_sum, *go_pre, *tpack, *rec, *gw_p, *ga_src_p, *ga_tgt_p;
  float* grad_pre;      // dL/d(out + skip) in the caller's layout (glue applied to merged rows, or the three-pass backward)
  void* ws; size_t ws_bytes;
  void* gws; size_t gws_bytes;
  void* sws; size_t sws_bytes;
  bool go_is_user;      // the padded upstream gradient is grad_out itself
  int go_shared;
};

void plan_bwd(const gat_layer_desc* d, bool has_grad_alpha, bool want_gx, bool want_gw, bool want_ga, Bump& b, BwdPlan& p) {
  const int64_t n = d->n, dp = (int64_t)d->nh * d->fp;
  const bool fused = !d->const_attention && !has_grad_alpha;
  p.go_shared = (!d->concat && d->nh > 1) ? 1 : 0;
  const bool glue = has_glue(d), glue_in_edge = glue && !needs_merge(d);
  // the adjoint kernel runs wherever rowdot cannot apply it: merged rows, and the three-pass backward
  p.grad_pre = (glue && !(glue_in_edge && fused)) ? b.take<float>((size_t)n * out_width(d)) : nullptr;
  p.go_is_user = !p.go_shared && !needs_merge(d);
  p.go_p = p.go_is_user ? nullptr : b.take<float>((size_t)n * (p.go_shared ? d->fp : dp));
  p.d_wh = b.take<float>((size_t)n * dp);
  p.ds_src = p.ds_tgt = p.s_sum = p.go_pre = p.tpack = p.rec = nullptr;
  if (!d->const_attention) {
    p.ds_src = b.take<float>((size_t)n * d->nh);
    p.ds_tgt = b.take<float>((size_t)n * d->nh);
    p.s_sum = b.take<float>((size_t)n * d->nh);
    if (fused) {
      if (glue_in_edge) p.go_pre = b.take<float>((size_t)n * dp);
      p.tpack = b.take<float>((size_t)n * gat_tgt_pack_stride(d->nh));
    } else {
      p.rec = b.take<float>((size_t)d->n_edges * 2 * d->nh);
    }
  }
  p.ws_bytes = gat_edge_bwd_workspace_bytes(n, d->n_edges, d->nh);
  p.ws = b.take<char>(p.ws_bytes);
  p.gw_p = (want_gw && d->fp != d->f) ? b.take<float>((size_t)dp * d->f_in) : nullptr;
  p.ga_src_p = p.ga_tgt_p = nullptr;
  p.sws = nullptr; p.sws_bytes = 0;
  if (want_ga && !d->const_attention) {
    p.ga_src_p = b.take<float>((size_t)d->nh * dp);
    p.ga_tgt_p = b.take<float>((size_t)d->nh * dp);
    p.sws_bytes = gat_scores_bwd_workspace_bytes((int)dp, d->nh);
    p.sws = b.take<char>(p.sws_bytes);
  }
  size_t g1 = want_gx ? gat_gemm_workspace_bytes(0, 1, n, d->f_in, dp, d->gemm_algo) : 0;
  size_t g2 = want_gw ? gat_gemm_workspace_bytes(1, 0, dp, d->f_in, n, d->gemm_algo) : 0;
  p.gws_bytes = g1 > g2 ? g1 : g2;
  p.gws = p.gws_bytes ? b.take<char>(p.gws_bytes) : nullptr;
}

int check_desc(const gat_layer_desc* d, const char* who, bool need_params = true) {
  if (d == nullptr) { set_error("%s: null descriptor", who); return GAT_EINVAL; }
  if (d->n < 0 || d->n_edges < 0 || d->f_in < 1 || d->nh < 1 || d->f < 1 || d->fp < d->f || d->fp % 4 != 0 || d->fp - d->f > 3) {
    set_error("%s: bad layer shape (n=%lld f_in=%lld nh=%d f=%d fp=%d)", who, (long long)d->n, (long long)d->f_in, d->nh, d->f, d->fp);
    return GAT_EINVAL;
  }
  if (need_params && (d->W == nullptr || (!d->const_attention && d->a == nullptr))) { set_error("%s: parameter pointer missing", who); return GAT_EINVAL; }
  if (!(d->out_drop_p >= 0.f && d->out_drop_p < 1.f)) { set_error("%s: output dropout %f not in [0, 1)", who, d->out_drop_p); return GAT_EINVAL; }
  if (d->skip != nullptr && d->ld_skip < out_width(d)) { set_error("%s: skip stride %lld below the output width", who, (long long)d->ld_skip); return GAT_EINVAL; }
  return GAT_OK;
}

// gmax = -inf, tie counters = 0: one launch instead of a fill and a memset
__global__ void layer_init_kernel(float* gmax, int32_t* ties, int64_t n_ties) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i == 0) *gmax = -INFINITY;
  for (int64_t j = i; j < n_ties; j += (int64_t)gridDim.x * blockDim.x) ties[j] = 0;
}

unsigned grid_for(int64_t elems) {
  int64_t b = (elems + 255) / 256;
  if (b < 1) b = 1;
  if (b > kNumSMs * 8) b = kNumSMs * 8;
  return (unsigned)b;
}

}  // namespace
}  // namespace gat

#define GAT_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != GAT_OK) return _rc; \
  } while (0)

extern "C" size_t gat_layer_fwd_arena_bytes(const gat_layer_desc* d) {
  using namespace gat;
  if (check_desc(d, "gat_layer_fwd_arena_bytes", false) != GAT_OK) return 0;
  Bump b{nullptr, 0};
  FwdPlan p;
  plan_fwd(d, b, p);
  return b.off + 256;
}

extern "C" size_t gat_layer_bwd_scratch_bytes(const gat_layer_desc* d, int has_grad_alpha, int want_gx, int want_gw, int want_ga) {
  using namespace gat;
  if (check_desc(d, "gat_layer_bwd_scratch_bytes", false) != GAT_OK) return 0;
  Bump b{nullptr, 0};
  BwdPlan p;
  plan_bwd(d, has_grad_alpha != 0, want_gx != 0, want_gw != 0, want_ga != 0, b, p);
  return b.off + 256;
}

extern "C" int gat_layer_fwd(const gat_layer_desc* d, const float* x, int64_t ldx, void* arena, size_t arena_bytes,
                             float* out, float* alpha, int want_ties, gat_stream_t stream) {
  using namespace gat;
  GAT_TRY(check_desc(d, "gat_layer_fwd"));
  GAT_CHECK_ARG(x != nullptr && out != nullptr && arena != nullptr && ldx >= d->f_in, "gat_layer_fwd: null buffer");
  GAT_CHECK_ARG(((uintptr_t)arena & 255) == 0, "gat_layer_fwd: arena must be 256-byte aligned");
  Bump b{reinterpret_cast<char*>(arena), 0};
  FwdPlan p;
  plan_fwd(d, b, p);
  if (b.off > arena_bytes) { set_error("gat_layer_fwd: arena too small (%zu < %zu)", arena_bytes, b.off); return GAT_EWORKSPACE; }
  if (d->n == 0) return GAT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = d->n;
  const int dp = d->nh * d->fp;
  float* out_p = p.out_is_user ? out : p.out_p;
  // parameters: reference layouts -> operand layouts (padded-head W, its transpose for dX, the two halves of a)
  GAT_TRY(gat_pack_params(d->W, d->const_attention ? nullptr : d->a, d->nh, d->f, d->fp, d->f_in, p.w_p, p.w_pT, p.a_src_p, p.a_tgt_p, stream));
  const float* w_p = p.w_p ? p.w_p : d->W;
  GAT_TRY(gat_project_fwd(x, n, d->f_in, ldx, d->x_act, w_p, d->f_in, dp, p.a_src_p, p.a_tgt_p, d->nh, p.wh, p.s_src, p.s_tgt,
                          d->gemm_algo, p.gws, p.gws_bytes, stream));
  int32_t *tie_dst = nullptr, *tie_src = nullptr;
  unsigned long long* tie_total = nullptr;
  if (!d->const_attention) {
    const int64_t n_ties = want_ties ? 2 * n * d->nh + 2 : 0;
    layer_init_kernel<<<grid_for(n_ties > 0 ? n_ties : 1), 256, 0, st>>>(p.gmax, p.ties, n_ties);
    GAT_LAUNCH_CHECK();
    if (want_ties) {
      tie_total = reinterpret_cast<unsigned long long*>(p.ties);
      tie_dst = p.ties + 2;
      tie_src = p.ties + 2 + n * d->nh;
    }
    GAT_TRY(gat_edge_max(d->rowptr, d->col, d->order, d->n_long, n, p.s_src, p.s_tgt, d->nh, p.gmax, p.fws, p.fws_bytes, stream));
  }
  const bool glue = has_glue(d);
  if (glue && p.out_is_user) {
    // a skip matrix whose rows cannot be read as float4 (odd stride / alignment) is not expected from torch; reject it loudly
    GAT_TRY(gat_edge_fwd_glue(d->rowptr, d->col, d->eid, d->order, d->n_long, n, p.wh, d->nh, d->fp, p.s_src, p.s_tgt, p.gmax,
                              d->const_attention, d->p_drop, d->seed, 0, out_p, d->out_act, d->skip, d->ld_skip, d->out_drop_p,
                              d->out_drop_seed, alpha, p.z, tie_dst, tie_src, tie_total, p.fws, p.fws_bytes, stream));
  } else {
    GAT_TRY(gat_edge_fwd(d->rowptr, d->col, d->eid, d->order, d->n_long, n, p.wh, d->nh, d->fp, p.s_src, p.s_tgt, p.gmax,
                         d->const_attention, d->p_drop, d->seed, 0, out_p, 0, alpha, p.z, tie_dst, tie_src, tie_total,
                         p.fws, p.fws_bytes, stream));
  }
  if (d->norm_out != nullptr)
    GAT_TRY(gat_attention_norm_scores(d->rowptr, d->col, n, d->n_edges, p.s_src, p.s_tgt, p.gmax, p.z, d->nh, d->const_attention, p.norm_t,
                                      d->norm_out, p.nws, p.nws_bytes, stream));
  if (!p.out_is_user) {
    if (glue) GAT_TRY(gat_head_merge_fwd_glue(out_p, n, d->nh, d->f, d->fp, d->concat, d->skip, d->ld_skip, d->out_act, d->out_drop_p,
                                              d->out_drop_seed, out, stream));
    else GAT_TRY(gat_head_merge_fwd(out_p, n, d->nh, d->f, d->fp, d->concat, out, stream));
  }
  return GAT_OK;
}

extern "C" int gat_layer_bwd(const gat_layer_desc* d, const float* x, int64_t ldx, const void* arena, const float* out,
                             const float* grad_out, const float* grad_alpha, void* scratch, size_t scratch_bytes,
                             float* gx, float* gW, float* ga, gat_stream_t stream) {
  using namespace gat;
  GAT_TRY(check_desc(d, "gat_layer_bwd"));
  GAT_CHECK_ARG(x != nullptr && arena != nullptr && grad_out != nullptr && scratch != nullptr, "gat_layer_bwd: null buffer");
  GAT_CHECK_ARG(((uintptr_t)arena & 255) == 0 && ((uintptr_t)scratch & 255) == 0, "gat_layer_bwd: arenas must be 256-byte aligned");
  if (d->n == 0) return GAT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  Bump fb{const_cast<char*>(reinterpret_cast<const char*>(arena)), 0};
  FwdPlan f;
  plan_fwd(d, fb, f);
  const bool want_gx = gx != nullptr, want_gw = gW != nullptr, want_ga = ga != nullptr && !d->const_attention;
  Bump bb{reinterpret_cast<char*>(scratch), 0};
  BwdPlan p;
  plan_bwd(d, grad_alpha != nullptr, want_gx, want_gw, want_ga, bb, p);
  if (bb.off > scratch_bytes) { set_error("gat_layer_bwd: scratch too small (%zu < %zu)", scratch_bytes, bb.off); return GAT_EWORKSPACE; }
  const int64_t n = d->n;
  const int dp = d->nh * d->fp;
  const bool fused = !d->const_attention && grad_alpha == nullptr;
  const float* out_p = f.out_is_user ? out : f.out_p;
  GAT_CHECK_ARG(out_p != nullptr, "gat_layer_bwd: the forward's output is needed (out)");
  const int32_t *tie_dst = nullptr, *tie_src = nullptr;
  const unsigned long long* tie_total = nullptr;
  if (!d->const_attention) {
    tie_total = reinterpret_cast<const unsigned long long*>(f.ties);
    tie_dst = f.ties + 2;
    tie_src = f.ties + 2 + n * d->nh;
  }
  // output glue: dL/d(out + skip) from dL/dy and the stored y = `out` -- inside rowdot when the glue ran in the edge kernel and
  // the one-pass backward follows, by the element-wise adjoint kernel otherwise; either way it is also dL/dskip
  const bool glue = has_glue(d), glue_in_edge = glue && f.out_is_user;
  GAT_CHECK_ARG(!glue || out != nullptr, "gat_layer_bwd: the forward's output y is needed for the output glue (out)");
  const float* go_user = grad_out;
  if (glue && !(glue_in_edge && fused)) {
    float* gp = d->grad_skip ? d->grad_skip : p.grad_pre;
    GAT_TRY(gat_out_glue_adjoint(grad_out, out, n, (int)out_width(d), d->out_act, d->out_drop_p, d->out_drop_seed, gp, stream));
    go_user = gp;
  }
  // upstream gradient -> padded-head layout (shared single row for a head mean)
  const float* go = go_user;
  if (p.go_shared) {
    GAT_TRY(gat_head_mean_bwd_shared(go_user, n, d->nh, d->f, d->fp, p.go_p, stream));
    go = p.go_p;
  } else if (needs_merge(d)) {
    GAT_TRY(gat_head_merge_bwd(go_user, n, d->nh, d->f, d->fp, d->concat, p.go_p, stream));
    go = p.go_p;
  }
  GAT_CHECK_ARG(d->grad_norm == nullptr || fused || d->const_attention,
                "gat_layer_bwd: the fused attention norm and an upstream dL/dalpha cannot be combined (use one or the other)");
  if (fused) {
    const float* norm_coef = d->grad_norm;          // NULL: no regulariser on this layer
    const float norm_scale = d->n_edges > 0 ? 1.0f / (float)d->n_edges : 0.f;
    if (glue_in_edge || norm_coef) {
      float* go_pre = glue_in_edge ? (d->grad_skip ? d->grad_skip : p.go_pre) : nullptr;
      GAT_TRY(gat_edge_bwd_rowdot_glue(go, p.go_shared, out_p, glue_in_edge ? d->out_act : 0, glue_in_edge ? d->skip : nullptr,
                                       glue_in_edge ? d->ld_skip : 0, glue_in_edge ? d->out_drop_p : 0.f, d->out_drop_seed, go_pre,
                                       d->rowptr, norm_coef ? f.norm_t : nullptr, norm_coef, norm_scale, f.z, n, d->nh,
                                       d->fp, p.s_sum, p.ds_tgt, f.s_tgt, p.tpack, p.ws, p.ws_bytes, stream));
      if (glue_in_edge) go = go_pre;
    } else {
      GAT_TRY(gat_edge_bwd_rowdot(go, p.go_shared, out_p, 0, nullptr, f.z, n, d->nh, d->fp, p.s_sum, p.ds_tgt, f.s_tgt, p.tpack,
                                  p.ws, p.ws_bytes, stream));
    }
    if (norm_coef) {
      GAT_TRY(gat_edge_bwd_fused_norm(d->rowptr_t, d->col_t, d->pos_t, d->order_t, d->n_long_t, d->eid, n, f.wh, d->nh, d->fp, f.s_src, f.s_tgt,
                                      f.gmax, f.z, d->p_drop, d->seed, 0, go, p.go_shared, p.s_sum, p.tpack, f.a_src_p, f.a_tgt_p,
                                      tie_dst, tie_src, tie_total, norm_coef, norm_scale, p.ds_src, p.ds_tgt, p.d_wh, p.ws, p.ws_bytes, stream));
    } else {
      GAT_TRY(gat_edge_bwd_fused(d->rowptr_t, d->col_t, d->pos_t, d->order_t, d->n_long_t, d->eid, n, f.wh, d->nh, d->fp, f.s_src, f.s_tgt,
                                 f.gmax, f.z, d->p_drop, d->seed, 0, go, p.go_shared, p.s_sum, p.tpack, f.a_src_p, f.a_tgt_p,
                                 tie_dst, tie_src, tie_total, nullptr, 0, n, p.ds_src, p.ds_tgt, p.d_wh, nullptr, 0, 0, 0,
                                 p.ws, p.ws_bytes, stream));
    }
  } else {
    GAT_TRY(gat_edge_bwd_main(d->rowptr_t, d->col_t, d->pos_t, d->order_t, d->n_long_t, d->eid, n, f.wh, d->nh, d->fp, f.s_src, f.s_tgt,
                              f.gmax, f.z, d->const_attention, d->p_drop, d->seed, 0, go, p.go_shared, grad_alpha, p.rec, p.d_wh,
                              p.ws, p.ws_bytes, stream));
    if (!d->const_attention) {
      GAT_TRY(gat_edge_bwd_rowsum(d->rowptr, d->tpos, d->order, d->n_long, n, d->nh, p.rec, f.z, p.s_sum, p.ds_tgt, p.ws, p.ws_bytes, stream));
      GAT_TRY(gat_edge_bwd_finish(d->rowptr_t, d->col_t, d->order_t, d->n_long_t, n, d->nh, d->fp, p.rec, p.s_sum, f.a_src_p, f.a_tgt_p,
                                  tie_dst, tie_src, tie_total, nullptr, 0, n, p.ds_src, p.ds_tgt, p.d_wh, p.ws, p.ws_bytes, stream));
    }
  }
  if (want_gx)   // dX = dWh W (K-major x K-major through the packed transpose); with a fused input activation * ELU'(x)
    GAT_TRY(gat_gemm_ex(0, 1, n, d->f_in, dp, p.d_wh, dp, f.w_pT, dp, gx, d->f_in, 0, 0, d->x_act ? x : nullptr, d->x_act ? ldx : 0,
                        d->gemm_algo, p.gws, p.gws_bytes, stream));
  if (want_gw) {
    float* gw_p = p.gw_p ? p.gw_p : gW;
    GAT_TRY(gat_gemm_ex(1, 0, dp, d->f_in, n, p.d_wh, dp, x, ldx, gw_p, d->f_in, 0, d->x_act, nullptr, 0, d->gemm_algo, p.gws, p.gws_bytes, stream));
  }
  if (want_ga)
    GAT_TRY(gat_scores_bwd(f.wh, n, dp, d->nh, p.ds_src, p.ds_tgt, p.ga_src_p, p.ga_tgt_p, p.sws, p.sws_bytes, stream));
  if ((want_gw && p.gw_p) || want_ga)
    GAT_TRY(gat_unpack_param_grads(p.gw_p, p.ga_src_p, p.ga_tgt_p, d->nh, d->f, d->fp, d->f_in, (want_gw && p.gw_p) ? gW : nullptr,
                                   want_ga ? ga : nullptr, stream));
  return GAT_OK;
}
