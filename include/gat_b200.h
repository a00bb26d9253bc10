/*
 * gat_b200.h -- C ABI of libgat_b200.so: the B200 (sm_100a) kernels behind the drop-in GATLayer.
 *
 * This is the boundary a maintainer of loodvn/gat-pytorch binds to replace the hot path of
 * `models/gat_layer.py:42-140` + `models/utils.py:6-72` (the reference has no FFI of its own:
 * it is pure Python on ATen, so each entry point below cites the reference lines it replaces).
 *
 * Conventions
 *  - Every pointer is a DEVICE pointer unless the name starts with `h_`.  The library never
 *    allocates, frees or keeps device memory: the caller (torch's caching allocator) owns all
 *    buffers, including workspaces sized by the *_workspace_bytes queries.
 *  - Every call only enqueues work on `stream` (a cudaStream_t); nothing synchronises.
 *  - Return value: 0 = ok; otherwise a negative GAT_E* code or a positive cudaError_t.
 *    gat_last_error() returns a thread-local human-readable message for the last failure.
 *  - Node features are row-major float32.  Internal feature buffers use the "padded head"
 *    layout (n, NH, Fp) with Fp = roundup(F, 4) and zero pad lanes, so that a 16-byte vector
 *    never straddles two heads; `dp` below is NH*Fp.
 *  - Graph arrays are int32 (N < 2^31, E' < 2^31); the user-facing edge_index stays int64.
 */
#ifndef GAT_B200_H_
#define GAT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GAT_OK 0
#define GAT_EINVAL (-1)      /* bad argument */
#define GAT_EWORKSPACE (-2)  /* workspace too small */
#define GAT_EUNSUPPORTED (-3)

/* Rows (destination rows of the CSR, source rows of the transposed CSR) with more than this many edges are "long":
 * the persistent edge kernels process them cooperatively, one CTA per row, instead of one warp per row.  A
 * `row_order` permutation handed to an edge kernel must list every long row before every short row
 * (gat_csr_build emits exactly that). */
#define GAT_LONG_ROW_EDGES 256

typedef void* gat_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define GAT_API __attribute__((visibility("default")))
#else
#define GAT_API
#endif

GAT_API int gat_version(void);
GAT_API const char* gat_last_error(void);
/* Number of kernels this library has launched so far in this process (bench.py's gpu_launches). */
GAT_API unsigned long long gat_launch_count(void);

/* ---------------------------------------------------------------------------------------
 * Kernel 1 -- edge_index -> rewritten edge list + destination-sorted CSR + transposed CSR.
 * Replaces utils.py:47-72 (add_remaining_self_loops / maybe_num_nodes) and the implicit
 * edge grouping of every scatter_add_/index_select in gat_layer.py:99-127.
 * ------------------------------------------------------------------------------------- */

/* Pass 1: d_stats[0] = max(edge_index)+1 (utils.py:72), d_stats[1] = #edges with src != dst
 * (utils.py:61), d_stats[2] = min(edge_index, 0) (negative ids are rejected by the caller).  `edge_index` is (2, E) with row stride `row_stride` elements, int64 when
 * index_is_int64 else int32.  The caller reads d_stats back (the one host sync per new graph,
 * replacing the `int(index.max())` sync at utils.py:72). */
GAT_API int gat_edges_scan(const void* edge_index, int64_t n_edges, int64_t row_stride, int index_is_int64,
                   int64_t* d_stats, gat_stream_t stream);

GAT_API size_t gat_csr_workspace_bytes(int64_t n_edges_in, int64_t n_edges_out, int64_t n_nodes);

/* Pass 2.  If add_self_loops: drops every (i,i), keeps the other edges in input order, appends
 * (k,k) for k < n_idx (utils.py:61-65); n_edges_out must equal d_stats[1] + n_idx.  Otherwise
 * the list is used as is (n_edges_out == n_edges_in).
 *  ei_out   (2, n_edges_out) int64, the rewritten list returned to the caller (gat_layer.py:137);
 *  rowptr   (n_nodes+1) / col / eid (n_edges_out): CSR by target, stable in edge order;
 *           rowptr diffs are the reference's degree counts (GATModel.py:196-201);
 *  rowptr_t (n_nodes+1) / col_t / pos_t: CSR by source; col_t = target ids, pos_t = slot of that
 *           edge in the target-sorted CSR; tpos (n_edges_out) = its inverse (CSR^T slot of each CSR slot);
 *  row_order / row_order_t (n_nodes) or NULL: scheduling permutations for the persistent edge kernels: rows with
 *           more than GAT_LONG_ROW_EDGES edges first (in row order), then the rest in row order;
 *  n_long   device int64[2] or NULL: number of long rows in row_order / row_order_t.  The edge kernels take it as a
 *           HOST value `n_long` (>= 0 exact, < 0 unknown): it only sizes / skips the cooperative launch. */
GAT_API int gat_csr_build(const void* edge_index, int64_t n_edges_in, int64_t row_stride, int index_is_int64,
                  int add_self_loops, int64_t n_idx, int64_t n_edges_out, int64_t n_nodes,
                  int64_t* ei_out, int32_t* rowptr, int32_t* col, int32_t* eid,
                  int32_t* rowptr_t, int32_t* col_t, int32_t* pos_t, int32_t* tpos,
                  int32_t* row_order, int32_t* row_order_t, int64_t* n_long,
                  void* workspace, size_t workspace_bytes, gat_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Kernel 2 -- dense projections.  Replaces gat_layer.py:64-65 (W) and :76-82 (a) and their
 * autograd (SURVEY.md section 9.2 "dense tail").
 * ------------------------------------------------------------------------------------- */

/* C[M,N] = op(A)[M,K] * op(B)[K,N], row-major float32.  A is stored (M,K) (ta=0) or (K,M) (ta=1);
 * B is stored (K,N) (tb=0) or (N,K) (tb=1).  algo: 0 = auto, 1 = fp32 FFMA tiles,
 * 2 = tcgen05 3xTF32 (TMA-fed; needs the alignment gat_gemm_tc_supported reports).
 * Deterministic: split-K partials are reduced in a fixed order inside `workspace`. */
GAT_API size_t gat_gemm_workspace_bytes(int ta, int tb, int64_t m, int64_t n, int64_t k, int algo);
GAT_API int gat_gemm_tc_supported(int ta, int tb, int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb, int64_t ldc);
GAT_API int gat_gemm(int ta, int tb, int64_t m, int64_t n, int64_t k,
             const float* a, int64_t lda, const float* b, int64_t ldb, float* c, int64_t ldc,
             int algo, void* workspace, size_t workspace_bytes, gat_stream_t stream);
/* gat_gemm with the reference's inter-layer glue fused in (GATModel.py:148-149 applies F.elu between layers; SURVEY.md
 * 8-f1).  act_a / act_b != 0: the A / B operand is ELU(stored values), applied to the tile on its way to the tensor
 * cores, so the activated tensor is never written.  mul_elu_grad_src != NULL (ta = 0 only): C[i,j] *= ELU'(src[i*mul_ld+j])
 * -- the adjoint of that activation, applied to dX in the epilogue. */
GAT_API int gat_gemm_ex(int ta, int tb, int64_t m, int64_t n, int64_t k,
                const float* a, int64_t lda, const float* b, int64_t ldb, float* c, int64_t ldc,
                int act_a, int act_b, const float* mul_elu_grad_src, int64_t mul_ld,
                int algo, void* workspace, size_t workspace_bytes, gat_stream_t stream);

/* The forward projection as one entry point: wh (n, dp) = x (n, f_in) * w (dp, f_in)^T  [gat_layer.py:64-65] and, when
 * a_src/a_tgt (nh, dp) are given, s_src/s_tgt (n, nh) = wh * a_src^T / wh * a_tgt^T  [gat_layer.py:76-82].  When the
 * tcgen05 path applies and dp <= 256 the score terms are computed in the GEMM epilogue from the accumulator tile
 * (fp64 accumulation over the fp32-rounded wh, same arithmetic as gat_scores_fwd); otherwise gat_gemm + gat_scores_fwd.
 * x_act != 0: the layer input is ELU(x) (fused glue, see gat_gemm_ex).
 * workspace: gat_gemm_workspace_bytes(0, 1, n, dp, f_in, algo). */
GAT_API int gat_project_fwd(const float* x, int64_t n, int64_t f_in, int64_t ldx, int x_act, const float* w, int64_t ldw, int dp,
                            const float* a_src, const float* a_tgt, int nh, float* wh, float* s_src, float* s_tgt,
                            int algo, void* workspace, size_t workspace_bytes, gat_stream_t stream);

/* Fused projection -> all-gather for the destination-range partitioned layer (one process per GPU).  Computes this
 * rank's slab wh[row_offset .. row_offset+n) = x W^T exactly like gat_project_fwd (tcgen05 path only: dp <= 256,
 * 16-byte aligned leading dimensions) and writes every output tile, from shared memory by TMA, into row
 * row_offset + i of EACH of the n_dests (1..8) gathered (>= row_offset + n, dp) buffers in h_wh_dests -- a HOST array of
 * device pointers: this rank's own buffer and the peers' buffers mapped into this process (CUDA peer / symmetric
 * memory over NVLink).  No separate collective moves the features; the caller only needs a cross-rank barrier (any
 * small collective) before reading its gathered buffer.  s_src / s_tgt (n, nh) are this rank's rows only. */
GAT_API int gat_project_fwd_allgather(const float* x, int64_t n, int64_t f_in, int64_t ldx, int x_act, const float* w, int64_t ldw, int dp,
                                      const float* a_src, const float* a_tgt, int nh,
                                      float* const* h_wh_dests, int n_dests, int64_t row_offset,
                                      float* s_src, float* s_tgt, gat_stream_t stream);

/* s_src[i,h] = <wh[i,:], a_src[h,:]>, s_tgt[i,h] = <wh[i,:], a_tgt[h,:]>  (fp64 accumulate).
 * The decomposition of gat_layer.py:76-82: logit[e,h] = s_src[src_e,h] + s_tgt[dst_e,h]. */
GAT_API int gat_scores_fwd(const float* wh, int64_t n, int dp, const float* a_src, const float* a_tgt, int nh,
                   float* s_src, float* s_tgt, gat_stream_t stream);

/* Adjoint of gat_scores_fwd w.r.t. the attention matrix: da_src = ds_src^T wh, da_tgt = ds_tgt^T wh, (nh, dp)
 * each, one streaming pass over wh with a fixed-order two-stage reduction (deterministic). */
GAT_API size_t gat_scores_bwd_workspace_bytes(int dp, int nh);
GAT_API int gat_scores_bwd(const float* wh, int64_t n, int dp, int nh, const float* ds_src, const float* ds_tgt,
                           float* da_src, float* da_tgt, void* workspace, size_t workspace_bytes, gat_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Kernel 3 -- fused edge forward.  Replaces gat_layer.py:70-132.
 * ------------------------------------------------------------------------------------- */

/* Kernel 3a: *gmax = max over all (e,h) of s_src[col[e],h] + s_tgt[dst(e),h]  (gat_layer.py:85).
 * gmax must hold -inf on entry (the kernel combines with an order-independent atomic max).
 * workspace: the same gat_edge_fwd_workspace_bytes() buffer that is later handed to gat_edge_fwd. */
GAT_API int gat_edge_max(const int32_t* rowptr, const int32_t* col, const int32_t* row_order, int64_t n_long, int64_t n,
                 const float* s_src, const float* s_tgt, int nh, float* gmax,
                 void* workspace, size_t workspace_bytes, gat_stream_t stream);

/* Kernel 3: per destination row, p = exp(0.01*(l-M)) (gat_layer.py:85-96), Z = sum p (:99-103),
 * alpha = p/(Z+1e-8) (:106-109), Philox dropout on alpha (:113-115), out = sum alpha*wh[src]
 * (:119-127), written in padded-head concat layout (n, dp).
 *  out_act     1: store ELU(out) -- the F.elu GATModel.forward applies to a hidden layer's output (GATModel.py:148-149) fused
 *               into this kernel's epilogue (opt-in); 0: the reference layer's pre-activation output;
 *  alpha_out   (n_edges, nh) in REWRITTEN EDGE ORDER (row eid[j]), pre-dropout, or NULL;
 *  z_out       (n, nh) saved for backward, or NULL;
 *  tie_dst/tie_src (n, nh) int32 + tie_total (1) uint64: arg-max-set bookkeeping for the
 *               gradient through max(), zero-initialised by the caller, or NULL;
 *  const_attention: logits are 0, gmax/s_src/s_tgt ignored (gat_layer.py:89-92);
 *  dropout_p > 0 enables the mask with (seed, offset) keyed on (edge id, head);
 *  row_order: scheduling permutation from gat_csr_build (long rows first: they take the cooperative CTA-per-row path,
 *               whose summation order differs from the warp-per-row path in the last bits) or NULL (natural order,
 *               every row warp-per-row);
 *  workspace: gat_edge_fwd_workspace_bytes() bytes (the persistent grid's row counters). */
GAT_API size_t gat_edge_fwd_workspace_bytes(void);
GAT_API int gat_edge_fwd(const int32_t* rowptr, const int32_t* col, const int32_t* eid, const int32_t* row_order, int64_t n_long,
                 int64_t n, const float* wh, int nh, int fp, const float* s_src, const float* s_tgt,
                 const float* gmax, int const_attention, float dropout_p, uint64_t seed, uint64_t offset,
                 float* out, int out_act, float* alpha_out, float* z_out,
                 int32_t* tie_dst, int32_t* tie_src, unsigned long long* tie_total,
                 void* workspace, size_t workspace_bytes, gat_stream_t stream);

/* OUTPUT GLUE (opt-in): everything GATModel.forward does between this layer and the next one --
 *     x_next = dropout_p( ELU( layer(x) + skip ) )          GATModel.py:130 (next layer's input dropout), :135-145 (skip add), :148-149
 * -- folded into the kernel that writes the layer's output, so no (n, D) pass of its own exists in either direction:
 *     y[i, c] = keep(i, c) * E(out[i, c] + skip[i, c]),   E = ELU (out_act) or identity, keep = 0 or 1/(1 - out_drop_p) from Philox keyed
 *     on (out_drop_seed, row i, column c / 4) -- the mask is never stored, the backward regenerates it.
 * gat_edge_fwd_glue = gat_edge_fwd with the whole glue in its epilogue, for layers whose padded rows ARE the caller's rows
 * (concat, F % 4 == 0); `skip` (n, ld_skip >= nh*fp) is indexed like `out`, 16-byte aligned rows, or NULL. */
GAT_API int gat_edge_fwd_glue(const int32_t* rowptr, const int32_t* col, const int32_t* eid, const int32_t* row_order, int64_t n_long,
                 int64_t n, const float* wh, int nh, int fp, const float* s_src, const float* s_tgt,
                 const float* gmax, int const_attention, float dropout_p, uint64_t seed, uint64_t offset,
                 float* out, int out_act, const float* skip, int64_t ld_skip, float out_drop_p, uint64_t out_drop_seed,
                 float* alpha_out, float* z_out,
                 int32_t* tie_dst, int32_t* tie_src, unsigned long long* tie_total,
                 void* workspace, size_t workspace_bytes, gat_stream_t stream);

/* Head merge (gat_layer.py:129-132): padded (n, nh, fp) -> (n, nh*f) concat or (n, f) head mean. */
GAT_API int gat_head_merge_fwd(const float* o_padded, int64_t n, int nh, int f, int fp, int concat,
                       float* out, gat_stream_t stream);
/* and its adjoint: grad (n, nh*f | f) -> padded (n, nh, fp) with zero pad lanes. */
GAT_API int gat_head_merge_bwd(const float* grad_out, int64_t n, int nh, int f, int fp, int concat,
                       float* go_padded, gat_stream_t stream);

/* Adjoint of the head mean in SHARED form: every head receives the same vector grad_out/nh, so it is stored once as
 * go_shared (n, fp) with zero pad lanes; gat_edge_bwd_main / gat_edge_bwd_rowdot read it with go_shared = 1 (a
 * 1/nh-th of the per-edge gather traffic of the expanded (n, nh, fp) form). */
GAT_API int gat_head_mean_bwd_shared(const float* grad_out, int64_t n, int nh, int f, int fp, float* go_shared,
                                     gat_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Kernel 4 -- atomic-free deterministic backward with ONE feature-row gather per edge.
 * Replaces autograd of gat_layer.py:70-132 (formulas: SURVEY.md section 9.2).
 * d_alpha[e,h] = m*<dOut[dst,h,:], Wh[src,h,:]> + dL/dalpha is formed in the source-major pass, where dOut[dst] is
 * gathered anyway and Wh[src] is the pass's own row, so no second gather of Wh[src] is needed.
 * ------------------------------------------------------------------------------------- */

GAT_API size_t gat_edge_bwd_workspace_bytes(int64_t n, int64_t n_edges, int nh);

/* Pass 1 (transposed CSR, rows = SOURCE nodes 0..n_rows-1): recomputes alpha from (s_src[src], s_tgt[dst], z[dst], gmax),
 * gathers go_padded[dst] once per edge, writes d_wh[src] = sum_e m*alpha*go[dst] (value path only) and the per-edge
 * record rec[j] = {d_alpha[0..nh), alpha[0..nh)} in CSR^T slot order.  s_tgt, z, go_padded are indexed by TARGET id
 * (a partitioned caller passes pointers shifted by its first owned row).  eid maps CSR slots to positions in the
 * rewritten edge list (dropout key / row of grad_alpha).  grad_alpha is (n_edges, nh) or NULL.
 * go_shared = 0: go_padded is (n, nh, fp); go_shared = 1 (head-mean layers): go_padded is (n, fp), the same row for
 * every head (gat_head_mean_bwd_shared). */
GAT_API int gat_edge_bwd_main(const int32_t* rowptr_t, const int32_t* col_t, const int32_t* pos_t, const int32_t* row_order_t,
                              int64_t n_long, const int32_t* eid, int64_t n_rows, const float* wh, int nh, int fp,
                              const float* s_src, const float* s_tgt, const float* gmax, const float* z,
                              int const_attention, float dropout_p, uint64_t seed, uint64_t offset,
                              const float* go_padded, int go_shared, const float* grad_alpha, float* rec, float* d_wh,
                              void* workspace, size_t workspace_bytes, gat_stream_t stream);

/* The backward in TWO passes, for the common case without an upstream dL/dalpha (nothing consumed the returned
 * attention): run gat_edge_bwd_rowdot FIRST -- S = <dOut, out>, ds_tgt and the Gamma partials need no per-edge data --
 * then this source-major pass, which does everything gat_edge_bwd_main + gat_edge_bwd_finish do, without records:
 * alpha recomputed, dOut[dst] gathered once, g = 0.01*alpha*(d_alpha - S[dst]) summed into ds_src, the arg-max
 * correction Gamma/|T| applied to ds_src / ds_tgt through the tie counts, and
 * d_wh[src] = sum_e m*alpha*dOut[dst] + ds_src*A_src + ds_tgt*A_tgt written once.  Arguments as in gat_edge_bwd_main /
 * gat_edge_bwd_finish (s_sum indexed by TARGET id; corr_override NULL on one GPU: Gamma is then finalised here from the
 * partials gat_edge_bwd_rowdot left in `workspace`, which must be the same buffer). */
GAT_API int gat_edge_bwd_fused(const int32_t* rowptr_t, const int32_t* col_t, const int32_t* pos_t, const int32_t* row_order_t,
                               int64_t n_long, const int32_t* eid, int64_t n_rows, const float* wh, int nh, int fp,
                               const float* s_src, const float* s_tgt, const float* gmax, const float* z,
                               float dropout_p, uint64_t seed, uint64_t offset,
                               const float* go_padded, int go_shared, const float* s_sum, const float* tgt_pack,
                               const float* a_src, const float* a_tgt,
                               const int32_t* tie_dst, const int32_t* tie_src, const unsigned long long* tie_total,
                               const float* corr_override, int64_t tgt_lo, int64_t tgt_hi,
                               float* ds_src, float* ds_tgt, float* d_wh,
                               float* const* h_push_dst, int n_push, int my_rank, int64_t rows_per_rank,
                               void* workspace, size_t workspace_bytes, gat_stream_t stream);
/* tgt_pack (optional): the per-target records written by gat_edge_bwd_rowdot; when given, s_tgt / z / s_sum are not read.
 * PUSH mode of gat_edge_bwd_fused (partitioned graphs: the reduce-scatter of dWh fused into the pass).  With n_push = P > 0,
 * h_push_dst is a HOST array of the P ranks' receive buffers, each (P, rows_per_rank, dp) floats and mapped into this
 * process (peer / symmetric memory); the finished dWh row of source `row` is stored into slab `my_rank`, row
 * row - owner*rows_per_rank of its owner's buffer (owner = row / rows_per_rank) instead of d_wh.  After a cross-rank barrier
 * the owner adds its P slabs in rank order with gat_slab_sum.  n_push = 0: plain d_wh (n_rows, dp). */
GAT_API int gat_slab_sum(const float* recv, int n_slabs, int64_t slab_rows, int dp, float* out, gat_stream_t stream);

/* Pass 2 (CSR by target, rows = owned TARGET nodes): s_sum[d,h] = sum_e alpha*d_alpha over the in-edges of d (records
 * gathered through tpos = CSR^T slot of each CSR slot); ds_tgt[d,h] = sum_e g = 0.01*s_sum*eps/(z+eps) (before the arg-max
 * correction); then reduces Gamma = sum ds_tgt in two fixed-order stages into the workspace. */
GAT_API int gat_edge_bwd_rowsum(const int32_t* rowptr, const int32_t* tpos, const int32_t* row_order, int64_t n_long, int64_t n_rows, int nh,
                                const float* rec, const float* z, float* s_sum, float* ds_tgt,
                                void* workspace, size_t workspace_bytes, gat_stream_t stream);

/* Pass 2 when there is no upstream dL/dalpha: s_sum[d,h] = <go_padded[d,h,:], out_padded[d,h,:]> (the forward output in
 * padded-head layout) -- identical to the record sum because out = sum_e m*alpha*Wh[src]; no per-edge gather at all.
 * Same outputs and Gamma reduction as gat_edge_bwd_rowsum.
 * out_is_act = 1 (the forward ran with out_act): out_padded holds h = ELU(out) and go_padded is dL/dh; the pass recovers
 * out = h > 0 ? h : log1p(h) and ELU'(out) = h > 0 ? 1 : h + 1, writes dL/dout = go*ELU' to go_out (n_rows, nh*fp) -- the buffer
 * the source-major pass must then gather -- and uses it in S: the ELU backward costs no pass of its own.
 * tgt_pack (optional, with s_tgt): the pass also writes one record per target, {s_tgt[d,:] | z[d,:] | s_sum[d,:]} in three
 * groups of (nh <= 4 ? 4 : 8) floats at a stride of gat_tgt_pack_stride(nh) floats (16-byte aligned), which gat_edge_bwd_fused
 * then gathers with ONE memory transaction per edge instead of three. */
GAT_API int gat_tgt_pack_stride(int nh);
GAT_API int gat_edge_bwd_rowdot(const float* go_padded, int go_shared, const float* out_padded, int out_is_act, float* go_out,
                                const float* z, int64_t n_rows, int nh, int fp,
                                float* s_sum, float* ds_tgt, const float* s_tgt, float* tgt_pack,
                                void* workspace, size_t workspace_bytes, gat_stream_t stream);

/* gat_edge_bwd_rowdot with the opt-in extras:
 *  - behind gat_edge_fwd_glue (out_is_act / skip / drop_p): y_padded holds y = keep * E(out + skip) and go_padded is dL/dy; the pass
 *    regenerates the mask, recovers out = E^-1(y * (1 - drop_p)) - skip, writes dL/d(out + skip) = go * keep * E' to go_out
 *    (n_rows, nh*fp) -- which is ALSO dL/dskip, and the buffer the source-major pass gathers -- and uses both in S;
 *  - fused attention-norm regulariser (norm_t / norm_coef / rowptr given, tgt_pack required): the loss carries
 *    c * sum |alpha*deg - 1| of this layer with c = *norm_coef * norm_scale (device scalar = upstream gradient of the layer's
 *    norm, host scale = 1/E'); then S[d,h] += c * deg(d) * norm_t[d,h] (norm_t from gat_attention_norm_scores) and deg(d) is
 *    stored in the target's record for gat_edge_bwd_fused_norm.  rowptr is the CSR by target of the rows handled here. */
GAT_API int gat_edge_bwd_rowdot_glue(const float* go_padded, int go_shared, const float* y_padded, int out_is_act, const float* skip,
                                     int64_t ld_skip, float drop_p, uint64_t drop_seed, float* go_out,
                                     const int32_t* rowptr, const float* norm_t, const float* norm_coef, float norm_scale,
                                     const float* z, int64_t n_rows, int nh, int fp,
                                     float* s_sum, float* ds_tgt, const float* s_tgt, float* tgt_pack,
                                     void* workspace, size_t workspace_bytes, gat_stream_t stream);

/* gat_edge_bwd_fused (one GPU) for a loss that also carries c * sum |alpha*deg - 1| of this layer: dL/dalpha[e,h] =
 * c * deg(d) * sign(alpha*deg(d) - 1) is formed inside the pass from the recomputed alpha and the deg(d) of the target's record,
 * so the attention-regularised training step (planetoid_gat.py:19-27, ppi_gat.py:22-33) keeps the ONE-pass backward and writes
 * no per-edge record.  The records must come from gat_edge_bwd_rowdot_glue called with the same coefficient. */
GAT_API int gat_edge_bwd_fused_norm(const int32_t* rowptr_t, const int32_t* col_t, const int32_t* pos_t, const int32_t* row_order_t,
                                    int64_t n_long, const int32_t* eid, int64_t n_rows, const float* wh, int nh, int fp,
                                    const float* s_src, const float* s_tgt, const float* gmax, const float* z,
                                    float dropout_p, uint64_t seed, uint64_t offset,
                                    const float* go_padded, int go_shared, const float* s_sum, const float* tgt_pack,
                                    const float* a_src, const float* a_tgt,
                                    const int32_t* tie_dst, const int32_t* tie_src, const unsigned long long* tie_total,
                                    const float* norm_coef, float norm_scale,
                                    float* ds_src, float* ds_tgt, float* d_wh,
                                    void* workspace, size_t workspace_bytes, gat_stream_t stream);

/* The output glue for layers whose padded rows are NOT the caller's rows (head-mean layers, F % 4 != 0): applied to the
 * merged row, out[i, c] = keep * E(merge(o_padded)[i, c] + skip[i, c]), skip (n, ld_skip >= width) or NULL. */
GAT_API int gat_head_merge_fwd_glue(const float* o_padded, int64_t n, int nh, int f, int fp, int concat,
                                    const float* skip, int64_t ld_skip, int act, float drop_p, uint64_t drop_seed, float* out,
                                    gat_stream_t stream);

/* Adjoint of the output glue, element-wise over an (n, width) matrix: grad_pre = dL/d(out + skip) (= dL/dskip) from grad_y = dL/dy
 * and the stored y.  Used where the glue was applied to merged rows, and on the three-pass backward (upstream dL/dalpha). */
GAT_API int gat_out_glue_adjoint(const float* grad_y, const float* y, int64_t n, int width, int act, float drop_p, uint64_t drop_seed,
                                 float* grad_pre, gat_stream_t stream);

/* Partitioned graphs only: *gamma_out = this rank's Gamma.  The caller all-reduces Gamma and tie_total over ranks and hands
 * Gamma/|T| to gat_edge_bwd_finish as `corr_override` (a device scalar).  On one GPU pass corr_override = NULL. */
GAT_API int gat_edge_bwd_gamma(void* workspace, size_t workspace_bytes, double* gamma_out, gat_stream_t stream);

/* Pass 3 (transposed CSR, rows = SOURCE nodes): g = 0.01*alpha*(d_alpha - s_sum[dst]); ds_src = sum g; applies the arg-max
 * correction Gamma/|T| to ds_src and ds_tgt via the tie counts; d_wh[src] += ds_src*A_src + ds_tgt*A_tgt, so that d_wh is the
 * total gradient of Wh.  [tgt_lo, tgt_hi) is the range of nodes this call also owns as targets: ds_tgt and tie_dst have
 * tgt_hi - tgt_lo rows and only those rows receive the ds_tgt*A_tgt term; s_sum is indexed by TARGET id. */
GAT_API int gat_edge_bwd_finish(const int32_t* rowptr_t, const int32_t* col_t, const int32_t* row_order_t, int64_t n_long, int64_t n_rows,
                                int nh, int fp, const float* rec, const float* s_sum, const float* a_src, const float* a_tgt,
                                const int32_t* tie_dst, const int32_t* tie_src, const unsigned long long* tie_total,
                                const float* corr_override, int64_t tgt_lo, int64_t tgt_hi,
                                float* ds_src, float* ds_tgt, float* d_wh,
                                void* workspace, size_t workspace_bytes, gat_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Caller-side glue: the attention-norm regulariser of GATModel.calc_attention_norm
 * (GATModel.py:189-234) for ONE layer:  *norm_out = sum_{e,h} |alpha[e,h]*deg(dst_e) - 1| / n_edges, with deg(dst) the
 * rowptr difference of gat_csr_build (== the scatter_add/index_select degrees of GATModel.py:196-201).  edge_dst is row 1 of
 * the rewritten edge list (n_edges entries, int64 or int32); alpha is (n_edges, nh) in that edge order.  The backward
 * writes grad_alpha = *upstream * sign(alpha*deg - 1) * deg / n_edges.  Deterministic (fixed-order fp64 reduction).
 * ------------------------------------------------------------------------------------- */
GAT_API size_t gat_attention_norm_workspace_bytes(void);
GAT_API int gat_attention_norm_fwd(const void* edge_dst, int index_is_int64, const int32_t* rowptr, const float* alpha,
                                   int64_t n_edges, int nh, float* norm_out, void* workspace, size_t workspace_bytes,
                                   gat_stream_t stream);
GAT_API int gat_attention_norm_bwd(const void* edge_dst, int index_is_int64, const int32_t* rowptr, const float* alpha,
                                   int64_t n_edges, int nh, const float* upstream, float* grad_alpha, gat_stream_t stream);
/* The same regulariser WITHOUT the (n_edges, nh) attention tensor: alpha is recomputed per CSR slot from the layer's score
 * terms (s_src, s_tgt, gmax, z as gat_project_fwd / gat_edge_max / gat_edge_fwd left them; pre-dropout, like the returned
 * attention).  *norm_out = sum |alpha*deg - 1| / n_edges; tsum[d,h] = sum_{e into d} alpha*sign(alpha*deg(d) - 1) is what the
 * backward needs (gat_edge_bwd_rowdot_glue / gat_edge_bwd_fused_norm).  workspace: gat_attention_norm_workspace_bytes(). */
GAT_API int gat_attention_norm_scores(const int32_t* rowptr, const int32_t* col, int64_t n, int64_t n_edges, const float* s_src,
                                      const float* s_tgt, const float* gmax, const float* z, int nh, int const_attention,
                                      float* tsum, float* norm_out, void* workspace, size_t workspace_bytes, gat_stream_t stream);

/* Micro-averaged F1 of a multilabel prediction, as the PPI task model logs it every step (ppi_gat.py:38, :48, :56:
 * sklearn.metrics.f1_score(y_pred = out > 0, y_true = y, average="micro") on host copies of both (n, classes) matrices -- 29 of
 * every 36 ms of a PPI-shaped training step).  counts[0..2] = TP, FP, FN over `count` elements (zeroed by the call; integer sums,
 * exact); F1 = 2 TP / (2 TP + FP + FN), 0 when the denominator is 0 (sklearn's zero_division default). */
GAT_API int gat_micro_f1_counts(const float* logits, const float* y_true, int64_t count, unsigned long long* counts, gat_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Visualisation feed.  Replaces the per-node `target_nodes == node_id` masks of
 * visualisation/entropy_histograms.py:103-115 and visualisation/weight_histograms.py:74-87 by one pass over the CSR of
 * gat_csr_build (a node's incoming edges are one segment, in the reference's edge order).  alpha is (n_edges, nh) in
 * edge-list order (what GATLayer returns); eid maps a CSR slot to its edge.
 *   gat_attention_entropy:        entropy[i,h] = scipy.stats.entropy(alpha[dst==i, h], base=2)   (n, nh)
 *                                 uniform[i]   = log2(in-degree of i)                             (n)
 *   gat_attention_degree_scaled:  scaled[j,h]  = alpha[eid[j],h] * in-degree(row of slot j)      (n_edges, nh), CSR order
 *                                 (== the concatenation over node_id of weight_histograms.py:81 before its `< 5` filter)
 * ------------------------------------------------------------------------------------- */
GAT_API int gat_attention_entropy(const int32_t* rowptr, const int32_t* eid, int64_t n, const float* alpha, int nh,
                                  float* entropy, float* uniform, gat_stream_t stream);
GAT_API int gat_attention_degree_scaled(const int32_t* rowptr, const int32_t* eid, int64_t n, const float* alpha, int nh,
                                        float* scaled, gat_stream_t stream);

/*   gat_attention_neighbourhood:  the star-plot feed of visualisation/neighbourhood_attention_weights.py:45-58.  For request i
 *                                 (target node nodes[i]) writes, at out_off[i] (a prefix sum of the requested in-degrees),
 *                                 out_src = its neighbours' ids in edge-list order and
 *                                 out_w   = alpha[.., head] over them / its maximum * 60 / neighbourhood size. */
GAT_API int gat_attention_neighbourhood(const int32_t* rowptr, const int32_t* col, const int32_t* eid, const float* alpha, int nh,
                                        int head, const int64_t* nodes, int64_t n_nodes_req, const int64_t* out_off,
                                        int64_t* out_src, float* out_w, gat_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * bf16 variant (BASELINE.json north_star "bf16 variant stated separately").  Opt-in: the matrices the edge
 * kernels GATHER per edge -- Wh in the forward, the upstream gradient dL/dout in the fused backward -- are bfloat16 copies
 * (half the bytes per edge); the row a pass owns, every accumulation and every output stay fp32.  Same arguments as the
 * fp32 entry points except the gathered matrix.  Supported for NH <= 4 and padded rows of 132..256 floats;
 * the backward additionally needs an unshared gradient (concat layers).  Parity bar of this variant: 2e-2 tensor-relative.
 * gat_edge_bf16_native(nh, fp, go_shared) tells whether such a kernel exists for a shape.  Everywhere else (narrow or very wide
 * rows, more than 4 heads, shared head-mean gradients, the three-pass backward) the variant keeps its NUMERICS -- the gathered
 * matrix is rounded to bfloat16 by gat_f32_round_bf16, stored as fp32 -- and runs the fp32 kernels: same results as a bf16
 * gather, without the halved bytes (those shapes are the L2-resident small graphs, where the bytes are not the bound).
 * ------------------------------------------------------------------------------------- */
GAT_API int gat_edge_bf16_native(int nh, int fp, int go_shared);
GAT_API int gat_f32_round_bf16(const float* src, float* dst, int64_t count, gat_stream_t stream);
GAT_API int gat_f32_to_bf16(const float* src, void* dst, int64_t count, gat_stream_t stream);
GAT_API int gat_edge_fwd_bf16(const int32_t* rowptr, const int32_t* col, const int32_t* eid, const int32_t* row_order, int64_t n_long,
                 int64_t n, const void* wh_bf16, int nh, int fp, const float* s_src, const float* s_tgt,
                 const float* gmax, int const_attention, float dropout_p, uint64_t seed, uint64_t offset,
                 float* out, int out_act, float* alpha_out, float* z_out,
                 int32_t* tie_dst, int32_t* tie_src, unsigned long long* tie_total,
                 void* workspace, size_t workspace_bytes, gat_stream_t stream);
GAT_API int gat_edge_bwd_fused_bf16(const int32_t* rowptr_t, const int32_t* col_t, const int32_t* pos_t, const int32_t* row_order_t,
                               int64_t n_long, const int32_t* eid, int64_t n_rows, const float* wh, int nh, int fp,
                               const float* s_src, const float* s_tgt, const float* gmax, const float* z,
                               float dropout_p, uint64_t seed, uint64_t offset,
                               const void* go_bf16, int go_shared, const float* s_sum, const float* tgt_pack,
                               const float* a_src, const float* a_tgt,
                               const int32_t* tie_dst, const int32_t* tie_src, const unsigned long long* tie_total,
                               const float* corr_override, int64_t tgt_lo, int64_t tgt_hi,
                               float* ds_src, float* ds_tgt, float* d_wh,
                               float* const* h_push_dst, int n_push, int my_rank, int64_t rows_per_rank,
                               void* workspace, size_t workspace_bytes, gat_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Parameter packing: the reference's parameter layouts <-> the kernels' operand layouts, ONE launch per direction
 * (replaces the view / slice / pad / cat / transpose chain and its autograd, a dozen tiny kernels per layer and step).
 *  W (NH*F, F_in)  [gat_layer.py:27]  -> W_p (NH*Fp, F_in) padded-head rows (NULL: not needed, i.e. Fp == F, use W itself)
 *                                      -> W_pT (F_in, NH*Fp), the K-major operand of dX = dWh W (NULL: not needed)
 *  a (NH, NH*2F)   [gat_layer.py:31]  -> a_src_p, a_tgt_p (NH, NH*Fp): a.view(NH, NH, 2F)[:, :, :F] and [:, :, F:]
 *                                         (gat_layer.py:76-82); a == NULL for const_attention layers
 * gat_unpack_param_grads is the adjoint: gW (NH*F, F_in) from gW_p (NULL when Fp == F: gW_p already is gW) and
 * ga (NH, NH*2F) from the two halves.
 * ------------------------------------------------------------------------------------- */
GAT_API int gat_pack_params(const float* W, const float* a, int nh, int f, int fp, int64_t f_in,
                            float* W_p, float* W_pT, float* a_src_p, float* a_tgt_p, gat_stream_t stream);
GAT_API int gat_unpack_param_grads(const float* gW_p, const float* ga_src_p, const float* ga_tgt_p, int nh, int f, int fp,
                                   int64_t f_in, float* gW, float* ga, gat_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * One layer per call.  gat_layer_fwd is GATLayer.forward (gat_layer.py:42-140) and gat_layer_bwd its autograd backward,
 * each issuing the whole kernel sequence above from ONE host call, with the parameters in the REFERENCE's layouts
 * (W (NH*F, F_in) gat_layer.py:27, a (NH, NH*2F) gat_layer.py:31) on both sides.  On the small named graphs a layer is
 * 8-10 kernels of a few microseconds; issuing them from one call keeps the GPU fed (csrc/layer.cu).
 *   desc       graph structure (gat_csr_build outputs) + layer configuration + parameter pointers;
 *   arena      gat_layer_fwd_arena_bytes(desc) bytes, 256-byte aligned: packed operands, Wh, score terms, Z, tie counts and
 *              (when a head merge / un-padding is needed) the padded output -- everything the backward reads again.  The
 *              caller keeps it alive until gat_layer_bwd and must not reuse it for another forward in between;
 *   out        (n, NH*F) concat or (n, F) head mean; alpha (n_edges, NH) in the rewritten edge order, or NULL;
 *   want_ties  0 when no gradient will be asked for (inference): the arg-max bookkeeping is skipped;
 *   scratch    gat_layer_bwd_scratch_bytes(...) bytes, 256-byte aligned, dead after the call;
 *   grad_alpha (n_edges, NH) or NULL (then the backward is rowdot + ONE fused source-major pass);
 *   gx (n, F_in) / gW (NH*F, F_in) / ga (NH, NH*2F): outputs, each may be NULL when not needed.
 * Output glue (see gat_edge_fwd_glue): out_act, skip (n, ld_skip; same columns as `out`), out_drop_p / out_drop_seed make
 * gat_layer_fwd return y = keep * E(out + skip); gat_layer_bwd then takes grad_out = dL/dy, `out` = y, and writes dL/dskip
 * (n, width of out; contiguous) to grad_skip when it is not NULL.
 * Fused attention-norm regulariser: norm_out makes gat_layer_fwd also return this layer's
 * sum |alpha*deg - 1| / E' (GATModel.py:196-225, one layer) without materialising alpha; grad_norm hands its upstream gradient
 * to gat_layer_bwd, which stays rowdot + ONE source-major pass.
 * ------------------------------------------------------------------------------------- */
typedef struct gat_layer_desc {
  const int32_t *rowptr, *col, *eid, *order;            /* CSR by target + scheduling permutation */
  const int32_t *rowptr_t, *col_t, *pos_t, *order_t;    /* CSR by source */
  const int32_t *tpos;
  int64_t n_long, n_long_t;
  int64_t n, n_edges;                                   /* nodes, rewritten edges */
  int64_t f_in;
  int32_t nh, f, fp;                                    /* heads, out_features, roundup(out_features, 4) */
  int32_t concat, const_attention, x_act, out_act, gemm_algo;
  float p_drop;                                         /* 0 outside training */
  uint64_t seed;
  const float *W, *a;                                   /* reference layouts; a = NULL for const_attention */
  const float* skip;                                    /* output glue: rows added before the activation, or NULL */
  int64_t ld_skip;
  float out_drop_p;                                     /* output glue: dropout of the stored output (the next layer's input dropout) */
  uint64_t out_drop_seed;
  float* grad_skip;                                     /* gat_layer_bwd: dL/dskip output, or NULL */
  float* norm_out;                                      /* gat_layer_fwd: device scalar receiving sum |alpha*deg - 1| / E' of this layer
                                                           (gat_attention_norm_scores), or NULL */
  const float* grad_norm;                               /* gat_layer_bwd: device scalar dL/dnorm, or NULL; needs grad_alpha == NULL */
} gat_layer_desc;

GAT_API size_t gat_layer_fwd_arena_bytes(const gat_layer_desc* desc);
GAT_API size_t gat_layer_bwd_scratch_bytes(const gat_layer_desc* desc, int has_grad_alpha, int want_gx, int want_gw, int want_ga);
GAT_API int gat_layer_fwd(const gat_layer_desc* desc, const float* x, int64_t ldx, void* arena, size_t arena_bytes,
                          float* out, float* alpha, int want_ties, gat_stream_t stream);
GAT_API int gat_layer_bwd(const gat_layer_desc* desc, const float* x, int64_t ldx, const void* arena, const float* out,
                          const float* grad_out, const float* grad_alpha, void* scratch, size_t scratch_bytes,
                          float* gx, float* gW, float* ga, gat_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GAT_B200_H_ */
