// Library-wide entry points: version, last-error string, GEMM dispatch.
#include "common.cuh"
#include <string.h>
#include <atomic>

namespace gat {

static thread_local char g_last_error[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

int gemm_simt(int ta, int tb, int64_t m, int64_t n, int64_t k, const float* a, int64_t lda, const float* b,
              int64_t ldb, float* c, int64_t ldc, void* workspace, size_t workspace_bytes, cudaStream_t st,
              int act_a, int act_b, const float* mul_src, int64_t mul_ld);
size_t simt_workspace_bytes(int64_t m, int64_t n, int64_t k);

int gemm_tc(int ta, int tb, int64_t m, int64_t n, int64_t k, const float* a, int64_t lda, const float* b,
            int64_t ldb, float* c, int64_t ldc, void* workspace, size_t workspace_bytes, cudaStream_t st,
            int act_a, int act_b, const float* mul_src, int64_t mul_ld);
size_t tc_workspace_bytes(int ta, int tb, int64_t m, int64_t n, int64_t k);
bool tc_supported(int ta, int tb, int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb, int64_t ldc);
bool tc_project_supported(int64_t n_rows, int64_t dp, int64_t k, int64_t ldx, int64_t ldw);
int gemm_tc_project(int64_t n_rows, int64_t dp, int64_t k, const float* x, int64_t ldx, const float* w, int64_t ldw,
                    float* wh, const float* a_src, const float* a_tgt, int nh, float* s_src, float* s_tgt, cudaStream_t st,
                    float* const* wh_dests, int n_dests, int64_t row_offset, int x_act);

}  // namespace gat

extern "C" int gat_version(void) { return 100; }

extern "C" const char* gat_last_error(void) { return gat::g_last_error; }

extern "C" unsigned long long gat_launch_count(void) { return gat::g_launches.load(std::memory_order_relaxed); }

extern "C" int gat_gemm_tc_supported(int ta, int tb, int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb,
                                     int64_t ldc) {
  return gat::tc_supported(ta, tb, m, n, k, lda, ldb, ldc) ? 1 : 0;
}

extern "C" size_t gat_gemm_workspace_bytes(int ta, int tb, int64_t m, int64_t n, int64_t k, int algo) {
  size_t simt = gat::simt_workspace_bytes(m, n, k);
  if (algo == 1) return simt;
  size_t tc = gat::tc_workspace_bytes(ta, tb, m, n, k);
  return simt > tc ? simt : tc;
}

extern "C" int gat_gemm(int ta, int tb, int64_t m, int64_t n, int64_t k, const float* a, int64_t lda,
                        const float* b, int64_t ldb, float* c, int64_t ldc, int algo, void* workspace,
                        size_t workspace_bytes, gat_stream_t stream) {
  return gat_gemm_ex(ta, tb, m, n, k, a, lda, b, ldb, c, ldc, 0, 0, nullptr, 0, algo, workspace, workspace_bytes, stream);
}

extern "C" int gat_gemm_ex(int ta, int tb, int64_t m, int64_t n, int64_t k, const float* a, int64_t lda,
                           const float* b, int64_t ldb, float* c, int64_t ldc,
                           int act_a, int act_b, const float* mul_elu_grad_src, int64_t mul_ld,
                           int algo, void* workspace, size_t workspace_bytes, gat_stream_t stream) {
  using namespace gat;
  GAT_CHECK_ARG(m >= 0 && n >= 0 && k >= 0, "gat_gemm: negative dimension");
  GAT_CHECK_ARG(algo >= 0 && algo <= 2, "gat_gemm: unknown algo %d", algo);
  GAT_CHECK_ARG(mul_elu_grad_src == nullptr || (ta == 0 && mul_ld >= n), "gat_gemm: the ELU' output multiplier needs ta = 0 and mul_ld >= n");
  cudaStream_t st = (cudaStream_t)stream;
  bool tc_ok = tc_supported(ta, tb, m, n, k, lda, ldb, ldc);
  if (algo == 2 && !tc_ok) {
    set_error("gat_gemm: tcgen05 path does not support this shape/alignment (m=%lld n=%lld k=%lld)", (long long)m,
              (long long)n, (long long)k);
    return GAT_EUNSUPPORTED;
  }
  // auto: tensor cores once the problem is big enough to fill the machine; tiny problems stay on the FFMA path
  const bool big = (double)m * (double)n * (double)k >= 1.6e7 && (ta ? k >= 4096 : m >= 512);
  if (algo == 2 || (algo == 0 && tc_ok && big))
    return gemm_tc(ta, tb, m, n, k, a, lda, b, ldb, c, ldc, workspace, workspace_bytes, st, act_a, act_b, mul_elu_grad_src, mul_ld);
  return gemm_simt(ta, tb, m, n, k, a, lda, b, ldb, c, ldc, workspace, workspace_bytes, st, act_a, act_b, mul_elu_grad_src, mul_ld);
}

// Kernel 2 as the north star names it: the projection GEMM that also emits the per-node score terms.
extern "C" int gat_project_fwd(const float* x, int64_t n, int64_t f_in, int64_t ldx, int x_act, const float* w, int64_t ldw, int dp,
                               const float* a_src, const float* a_tgt, int nh, float* wh, float* s_src, float* s_tgt,
                               int algo, void* workspace, size_t workspace_bytes, gat_stream_t stream) {
  using namespace gat;
  GAT_CHECK_ARG(n >= 0 && f_in >= 1 && dp >= 4 && dp % 4 == 0, "gat_project_fwd: bad shape");
  GAT_CHECK_ARG(algo >= 0 && algo <= 2, "gat_project_fwd: unknown algo %d", algo);
  GAT_CHECK_ARG((a_src == nullptr) == (a_tgt == nullptr), "gat_project_fwd: attention halves must be given together");
  if (n == 0) return GAT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool want_scores = a_src != nullptr;
  const bool big = (double)n * dp * (double)f_in >= 1.6e7 && n >= 512;
  const bool fused_ok = want_scores && tc_project_supported(n, dp, f_in, ldx, ldw) &&
                        ((uintptr_t)x | (uintptr_t)w | (uintptr_t)wh | (uintptr_t)a_src | (uintptr_t)a_tgt) % 16 == 0;
  if (fused_ok && (algo == 2 || (algo == 0 && big)))
    return gemm_tc_project(n, dp, f_in, x, ldx, w, ldw, wh, a_src, a_tgt, nh, s_src, s_tgt, st, nullptr, 0, 0, x_act);
  int rc = gat_gemm_ex(0, 1, n, dp, f_in, x, ldx, w, ldw, wh, dp, x_act, 0, nullptr, 0,
                       algo == 2 && !tc_supported(0, 1, n, dp, f_in, ldx, ldw, dp) ? 0 : algo, workspace, workspace_bytes, stream);
  if (rc != GAT_OK || !want_scores) return rc;
  return gat_scores_fwd(wh, n, dp, a_src, a_tgt, nh, s_src, s_tgt, stream);
}

// Fused projection -> all-gather over NVLink peer memory (the partitioned layer's one exchange step).
// Every output tile of wh = x W^T is staged in shared memory once and written by TMA to row `row_offset + i` of each
// of the n_dests gathered buffers -- this rank's own and the peers' (pointers mapped through CUDA peer / symmetric
// memory) -- so the transfer overlaps the GEMM tile by tile and no separate collective moves the features.  The
// score terms are computed from the same accumulator tile and stay local (n rows).
extern "C" int gat_project_fwd_allgather(const float* x, int64_t n, int64_t f_in, int64_t ldx, int x_act, const float* w, int64_t ldw, int dp,
                                         const float* a_src, const float* a_tgt, int nh,
                                         float* const* h_wh_dests, int n_dests, int64_t row_offset,
                                         float* s_src, float* s_tgt, gat_stream_t stream) {
  using namespace gat;
  GAT_CHECK_ARG(n >= 0 && f_in >= 1 && dp >= 4 && dp % 4 == 0 && row_offset >= 0, "gat_project_fwd_allgather: bad shape");
  GAT_CHECK_ARG(h_wh_dests != nullptr && n_dests >= 1 && n_dests <= 8, "gat_project_fwd_allgather: 1..8 destinations");
  GAT_CHECK_ARG(a_src != nullptr && a_tgt != nullptr && s_src != nullptr && s_tgt != nullptr, "gat_project_fwd_allgather: score buffers missing");
  if (n == 0) return GAT_OK;
  if (!tc_project_supported(n, dp, f_in, ldx, ldw)) {
    set_error("gat_project_fwd_allgather: needs the tcgen05 path (dp <= 256, 16-byte aligned leading dimensions)");
    return GAT_EUNSUPPORTED;
  }
  return gemm_tc_project(n, dp, f_in, x, ldx, w, ldw, nullptr, a_src, a_tgt, nh, s_src, s_tgt, (cudaStream_t)stream,
                         h_wh_dests, n_dests, row_offset, x_act);
}
