// Kernel 4: atomic-free deterministic backward of the edge stage (autograd of gat_layer.py:70-132; formulas in
// SURVEY.md section 9.2) with ONE feature-row gather per edge.
//
// The obvious split (a dst pass that gathers Wh[src] for d_alpha, then a src pass that gathers dOut[dst] for dWh) moves
// two rows per edge.  But d_alpha[e,h] = <dOut[dst_e,h,:], Wh[src_e,h,:]> can be formed where dOut[dst_e] is gathered
// anyway -- in the SOURCE-major pass, whose own row Wh[src] sits in registers.  So:
//
//   pass 1  gat_edge_bwd_main    (CSR^T, heavy)  per source row s: recompute alpha from (s_src[s], s_tgt[d], Z[d], M),
//                                                gather dOut[d] once, dWh[s] += m*alpha*dOut[d],
//                                                d_alpha = m*<dOut[d],Wh[s]> + dL/dalpha, record {d_alpha, alpha} per edge
//                                                (coalesced, CSR^T order)
//   pass 2  gat_edge_bwd_rowsum  (CSR,  light)   per target row d: S = sum_e alpha*d_alpha (32-byte record gathers);
//                                                ds_tgt = sum_e g = 0.01*S*eps/(Z+eps)   [closed form: sum_e alpha = Z/(Z+eps)]
//           Gamma = sum ds_tgt (two fixed-order stages)          -> gradient through the global max()
//   pass 3  gat_edge_bwd_finish  (CSR^T, light)  per source row s: g = 0.01*alpha*(d_alpha - S[d]), ds_src = sum g,
//                                                arg-max corrections, dWh[s] += ds_src*A_src + ds_tgt*A_tgt
//
// Every sum is a fixed-order register / shuffle / shared-memory reduction; the only atomics are the integer row
// counters of the persistent schedulers, which never influence a result.
#include "edge_common.cuh"
#include <cstddef>
#include <cstdlib>

namespace gat {

struct BwdHeader {          // first 256 bytes of the backward workspace
  double gamma;             // sum over all (e,h) of g
  float corr;               // gamma / |T|   (0 when the arg-max set is empty)
  int n_partials;           // number of partials written by gamma_partial_kernel
  unsigned int pad0[12];
  unsigned int counter_a;   // row scheduler of pass 1 / pass 3 (byte offset 64): warp-level counter,
  unsigned int counter_a_cta;   //                                                  CTA-level long-row counter
  unsigned int pad1[14];
  unsigned int counter_b;   // row scheduler of pass 2          (byte offset 128)
  unsigned int counter_b_cta;
};
static_assert(offsetof(BwdHeader, counter_a) == 64 && offsetof(BwdHeader, counter_b) == 128, "header layout");
constexpr size_t kBwdHeaderBytes = 256;
constexpr int kGammaBlocks = 592;   // 4 per SM

// ------------------------------------------------------------------------------------------------------------
// pass 1
// ------------------------------------------------------------------------------------------------------------
struct BwdMainParams {
  const int32_t* rowptr_t; const int32_t* col_t; const int32_t* pos_t; const int32_t* eid; RowSched sched;
  const float* wh; int nh; int dp; int chunks; int chunks_per_head;
  const float* s_src; const float* s_tgt; const float* gmax; const float* z;   // s_tgt / z indexed by TARGET id
  int const_attention; float dropout_p; uint64_t seed; uint64_t offset;
  const float* go; const float* grad_alpha;                                    // go indexed by TARGET id
  int go_ld; int go_shared;   // go row stride (floats); go_shared: one (Fp)-wide row serves every head (head-mean layers)
  float* rec; float* d_wh;
  // FUSED mode (no upstream dL/dalpha): S = <dOut, out> and Gamma are known BEFORE this pass (gat_edge_bwd_rowdot runs
  // first), so g, ds_src, the arg-max corrections and dWh += ds_src*A_src + ds_tgt*A_tgt are all done here; no records.
  const float* s_sum;                      // indexed by TARGET id
  // per-target record {s_tgt[NHT] | Z[NHT] | S[NHT] | pad} written by gat_edge_bwd_rowdot (stride 4*NHT floats = one 64 B
  // DRAM atom for NH <= 4): the three per-edge gathers of a target's scalars become ONE.  ncu: the fused pass moved
  // 84 GB for 74.5 GB of algorithmic bytes on the products graph because each of the three 16-byte gathers (three
  // 39 MB arrays, not L2 resident under the streaming traffic) cost its own 64-byte DRAM atom.
  const float* tpack;
  const float* a_src; const float* a_tgt;
  const int32_t* tie_dst; const int32_t* tie_src; const BwdHeader* header; const float* corr_override;
  int64_t tgt_lo; int64_t tgt_hi;
  float* ds_src; float* ds_tgt;
  // PUSH mode (partitioned graphs, fused reduce-scatter): the finished dWh row of source `row` is stored straight into
  // its OWNER's receive buffer -- slab `my_rank`, row `row - owner*rows_per_rank` -- through the peer-mapped pointers
  // push_dst[owner] (this rank's own buffer included), so the exchange rides NVLink while the pass is still running.
  float* push_dst[8]; int push; int my_rank; int64_t rows_per_rank;
  int stage_offset_floats;   // GS: where the staging area starts inside the dynamic shared memory
  // PUSH with G == 32: a finished row is staged in shared memory (two dp-float buffers per warp) and leaves as ONE bulk
  // copy (cp.async.bulk, the TMA engine).  Per-thread 16-byte stores to peer memory reached ~280 GB/s over NVLink (the
  // kernel "finished" and the barrier behind it then waited 2.4 ms per layer for the posted writes to drain at 4 GPUs);
  // the projection's TMA stores reach 700 GB/s on the same links.
  int rowbuf_offset_floats;  // < 0: per-thread stores
  // PUSH: every rank walks the source rows in the same order, so without a rotation all P ranks push into the SAME owner's
  // slab at the same time (its NVLink ingress serialises them: measured as a 2.4 ms wait per layer at 4 GPUs behind a
  // 4.4 ms pass); rank r starts at the slab of owner r+1 instead.
  int64_t sched_rot;
  int go_bf16;               // host-side switch: launch the BF16 instantiation (go is a bfloat16 matrix)
  // attention-norm regulariser fused: the loss carries c * |alpha*deg - 1|_1 with c = *norm_coef * norm_scale
  // (upstream gradient of the layer's norm, device scalar, times 1/E'); dL/dalpha[e,h] = c*deg(d)*sign(alpha*deg(d) - 1) is
  // formed here from the recomputed alpha and deg(d) -- float 3*NHT of the target's record -- and S[d] already includes its share
  const float* norm_coef; float norm_scale;
};

__device__ __forceinline__ float* dwh_row_ptr(const BwdMainParams& P, const int64_t row) {
  if (!P.push) return P.d_wh + row * P.dp;
  const int64_t owner = row / P.rows_per_rank;
  return P.push_dst[owner] + ((int64_t)P.my_rank * P.rows_per_rank + (row - owner * P.rows_per_rank)) * P.dp;
}

template <int G, int SLOTS>
struct MainShape {
  static constexpr int TB = (G < 8) ? G : (SLOTS >= 6 ? 4 : 8);     // edges per transpose-reduce sub-batch
  static constexpr int U = (SLOTS >= 4) ? 2 : (TB < 4 ? TB : 4);    // edges in flight
  static constexpr int PSTRIDE = SLOTS * G + 1;                      // floats per row of the transpose tile (odd: conflict free)
};

// COOP (long source rows): the CTA's 256/G groups take the row's batches round-robin (each writes the records of its
// own edges) and the dWh row is combined over the groups in group order through `coop`, which ALIASES the groups'
// `part` tiles -- hence the CTA barrier before it is written.  Called by all threads of the CTA in that case.
// What a lane owns of a padded row: fixed for the whole kernel, so it is computed once per thread, not once per row.
template <int SLOTS>
struct LaneShape {
  int head[SLOTS];   // head of the lane's chunk in slot s
  int goff[SLOTS];   // float offset of that chunk inside a gathered dOut row
  bool ok[SLOTS];    // chunk exists (row narrower than SLOTS*G chunks)
};

template <int G, int SLOTS>
__device__ __forceinline__ LaneShape<SLOTS> make_lane_shape(const BwdMainParams& P, const int gl) {
  LaneShape<SLOTS> L;
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int c = s * G + gl;
    L.ok[s] = c < P.chunks;
    L.head[s] = L.ok[s] ? c / P.chunks_per_head : 0;
    // the head-mean layer's upstream gradient is the same (Fp)-wide vector for every head (gat_layer.py:132), so it is
    // stored once and every head reads the same chunk
    // a slot that does not exist (row narrower than SLOTS*G chunks) points at chunk 0: the gather loop treats every slot
    // alike (no per-lane predicates, no divergence); what such a slot accumulates is never stored or read
    L.goff[s] = L.ok[s] ? (P.go_shared ? c - L.head[s] * P.chunks_per_head : c) * 4 : 0;
  }
  return L;
}

// FULL: the row has exactly SLOTS*G chunks (e.g. products' 256 floats), so every slot of every lane exists and the
// per-slot predicates fold away at compile time.
// GS ("gathered rows staged"): head-mean layers hand every head the SAME (Fp)-wide upstream-gradient row, so the row an edge
// gathers is narrow (products' last layer: 192 B) and four edges in flight per warp are only 768 B -- the pass was
// latency-bound at a third of the DRAM rate (ncu: 33 % DRAM, 10.3 ms for 26 GB).  With GS the warp copies the rows of 16
// edges at a time into shared memory with cp.async (no registers held: 3 KB in flight per warp) and the gather loop reads
// its chunks from there.
constexpr int kStageEdges = 16;
constexpr int kStageMaxChunks = 16;     // widest staged row: 16 float4 = 64 floats per head
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void bulk_store(void* gmem_dst, const void* smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
               "r"((unsigned)__cvta_generic_to_shared(smem_src)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// at most one bulk store of this thread may still be READING its shared-memory source (the other buffer)
__device__ __forceinline__ void bulk_wait_read_keep1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// BF16 (opt-in bf16 variant): the gathered upstream-gradient matrix is a bfloat16 copy (half the bytes per
// edge); a chunk is 8 bytes, widened to fp32 in registers; the row's own Wh, every accumulation and every output stay fp32.
template <bool BF16>
__device__ __forceinline__ float4 gathered_chunk(const float* rowp, const int off) {
  if (BF16) {
    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(rowp) + off));
    return make_float4(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u),
                       __uint_as_float(raw.y << 16), __uint_as_float(raw.y & 0xffff0000u));
  }
  return ldg4(rowp + off);
}

template <int G, int SLOTS, int NHT, bool COOP, bool FUSED, bool FULL, bool GS = false, bool BF16 = false>
__device__ __forceinline__ void bwd_main_row(const BwdMainParams& P, const LaneShape<SLOTS>& L, const int64_t row, const int start,
                                             const int end, const int tid, const int gl, const int gbase, const unsigned gmask, const float gmax, const float corr,
                                             int* sh_dst, const float** sh_gp, float* sh_w, float* sh_da, float* sh_s, float* part, float* coop,
                                             float4* stage, int* sh_flip) {
  constexpr int TB = MainShape<G, SLOTS>::TB, U = MainShape<G, SLOTS>::U;
  constexpr int PSTRIDE = MainShape<G, SLOTS>::PSTRIDE;   // compile-time stride of the transpose tile rows
  constexpr int NG = kEdgeThreads / G;
  const int grp = tid / G;
  const int first = COOP ? grp * G : 0, step = COOP ? NG * G : G;
  const int nh = P.nh;
  float4 whr[SLOTS], acc[SLOTS];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    acc[s] = make_float4(0.f, 0.f, 0.f, 0.f);
    whr[s] = ((FULL || L.ok[s]) && (FUSED || !P.const_attention)) ? ldg4(P.wh + row * P.dp + (s * G + gl) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float ss[NHT], gsum[NHT];
#pragma unroll
  for (int h = 0; h < NHT; ++h) { ss[h] = ((FUSED || !P.const_attention) && h < nh) ? __ldg(P.s_src + row * nh + h) : 0.f; gsum[h] = 0.f; }
  // FUSED, group per row: lane h < nh prefetches the row's tie counts / ds_tgt now, so the epilogue needs no extra round trip
  const bool own_tgt = FUSED && row >= P.tgt_lo && row < P.tgt_hi;
  const int64_t trow = row - P.tgt_lo;
  int p_ts = 0, p_td = 0;
  float p_t = 0.f;
  if (FUSED && !COOP && gl < nh) {
    if (P.tie_src) p_ts = __ldg(P.tie_src + row * nh + gl);
    if (own_tgt) {
      if (P.tie_dst) p_td = __ldg(P.tie_dst + trow * nh + gl);
      p_t = P.ds_tgt[trow * nh + gl];
    }
  }
  const float* wbase[SLOTS];   // this lane's weight of edge k in slot s: wbase[s][k * NHT]
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) wbase[s] = sh_w + gbase * NHT + L.head[s];

  for (int base = start + first; base < end; base += step) {
    const int e = base + gl;
    const bool valid = e < end;
    float alpha[NHT], msk[NHT], ga[NHT];
#pragma unroll
    for (int h = 0; h < NHT; ++h) { alpha[h] = 0.f; msk[h] = 1.f; ga[h] = 0.f; }
    if (valid) {
      const int d = __ldg(P.col_t + e);
      const float* zp = P.z + (int64_t)d * nh;
      if (!FUSED && P.const_attention) {
#pragma unroll
        for (int h = 0; h < NHT; ++h) alpha[h] = h < nh ? 1.f / (__ldg(zp + h) + kSoftmaxEps) : 0.f;
      } else {
        const float* tp = P.s_tgt + (int64_t)d * nh;
        float sv[NHT];
        if (FUSED && P.tpack != nullptr) {   // one 16-byte-aligned record per target: s_tgt | Z | S
          const float* pk = P.tpack + (int64_t)d * (4 * NHT);
          float tv[NHT], zv[NHT];
#pragma unroll
          for (int q = 0; q < NHT / 4; ++q) {
            const float4 t4 = ldg4(pk + 4 * q), z4 = ldg4(pk + NHT + 4 * q), s4 = ldg4(pk + 2 * NHT + 4 * q);
            tv[4 * q] = t4.x; tv[4 * q + 1] = t4.y; tv[4 * q + 2] = t4.z; tv[4 * q + 3] = t4.w;
            zv[4 * q] = z4.x; zv[4 * q + 1] = z4.y; zv[4 * q + 2] = z4.z; zv[4 * q + 3] = z4.w;
            sv[4 * q] = s4.x; sv[4 * q + 1] = s4.y; sv[4 * q + 2] = s4.z; sv[4 * q + 3] = s4.w;
          }
#pragma unroll
          for (int h = 0; h < NHT; ++h)
            if (h < nh) alpha[h] = attn_exp(ss[h] + tv[h], gmax) / (zv[h] + kSoftmaxEps);
          if (P.norm_coef != nullptr) {   // g = 0.01*alpha*(m*<dOut,Wh> + dL/dalpha - S): fold dL/dalpha into the S term
            const float cn = __ldg(P.norm_coef) * P.norm_scale, deg = __ldg(pk + 3 * NHT);
#pragma unroll
            for (int h = 0; h < NHT; ++h) {
              const float u = alpha[h] * deg - 1.0f;
              sv[h] -= u > 0.f ? cn * deg : (u < 0.f ? -cn * deg : 0.f);
            }
          }
        } else {
          if (FUSED) {   // S[dst] travels with s_tgt[dst] / Z[dst] (same round trip) and waits in shared memory
            const float* sp = P.s_sum + (int64_t)d * nh;
#pragma unroll
            for (int h = 0; h < NHT; ++h) sv[h] = h < nh ? __ldg(sp + h) : 0.f;
          }
#pragma unroll
          for (int h = 0; h < NHT; ++h)
            if (h < nh) alpha[h] = attn_exp(ss[h] + __ldg(tp + h), gmax) / (__ldg(zp + h) + kSoftmaxEps);
        }
        if (FUSED) {   // alpha * S[dst]: all the epilogue of this batch needs besides m*alpha (sh_w), so that no per-edge
                       // register array stays live across the gather loop (the loop needs them for its loads in flight)
#pragma unroll
          for (int h = 0; h < NHT; ++h) sh_s[tid * NHT + h] = alpha[h] * sv[h];
        }
      }
      if (P.dropout_p > 0.f || P.grad_alpha) {
        const int edge_id = __ldg(P.eid + __ldg(P.pos_t + e));
        if (P.dropout_p > 0.f) dropout_scales<NHT>(P.seed, P.offset, (uint32_t)edge_id, nh, P.dropout_p, msk);
        if (P.grad_alpha) {
#pragma unroll
          for (int h = 0; h < NHT; ++h)
            if (h < nh) ga[h] = __ldg(P.grad_alpha + (int64_t)edge_id * nh + h);
        }
      }
      sh_dst[tid] = d;
      // the 64-bit row address is formed once, by the lane that owns the edge (BF16: rows of 2-byte elements)
      sh_gp[tid] = BF16 ? reinterpret_cast<const float*>(reinterpret_cast<const uint16_t*>(P.go) + (int64_t)d * P.go_ld)
                        : P.go + (int64_t)d * P.go_ld;
#pragma unroll
      for (int h = 0; h < NHT; ++h) sh_w[tid * NHT + h] = msk[h] * alpha[h];
    }
    __syncwarp(gmask);
    const int cnt = min(G, end - base);
    for (int t0 = 0; t0 < cnt; t0 += TB) {
      const int tcnt = min(TB, cnt - t0);
      if (GS && (t0 % kStageEdges) == 0) {   // stage the gathered rows of edges t0 .. t0+15 of this batch (G == 32)
        const int gch = P.chunks_per_head;
        const int total = min(kStageEdges, cnt - t0) * gch;
        const int q32 = 32 / gch, r32 = 32 - q32 * gch;
        int k = gl / gch, j = gl - k * gch;
        __syncwarp(gmask);                   // the previous 16 edges have been consumed by every lane
        for (int i = gl; i < total; i += 32) {
          cp_async16(stage + i, sh_gp[gbase + t0 + k] + j * 4);
          k += q32; j += r32;
          if (j >= gch) { j -= gch; ++k; }
        }
        cp_async_wait_all();
        __syncwarp(gmask);
      }
      // groups of U edges: all loads of a group are issued before its first use; only the groups that exist are executed.
      // A FULL group (the common case) is one branch-free block whose shared-memory addresses are a base formed once per
      // group plus compile-time offsets; with per-edge `exists` branches inside it the compiler re-derived every address
      // from %tid for every edge (ncu source page: ~22 of the ~55 instructions per edge).
#pragma unroll 1
      for (int tt = 0; tt < tcnt; tt += U) {
        const int k0 = t0 + tt;
        const float* const* gp = sh_gp + gbase + k0;       // row address of edge k0 + u: gp[u]
        const float* wp[SLOTS];                            // weight of edge k0 + u in slot s: wp[s][u * NHT]
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) wp[s] = wbase[s] + k0 * NHT;
        float* prow = part + tt * PSTRIDE + gl;
        float4 v[U][SLOTS];
        if (tt + U <= tcnt) {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (GS) {
              const float4* srow = stage + ((k0 + u) % kStageEdges) * P.chunks_per_head;
#pragma unroll
              for (int s = 0; s < SLOTS; ++s)
                v[u][s] = srow[L.goff[s] >> 2];
            } else {
              const float* rowp = gp[u];
#pragma unroll
              for (int s = 0; s < SLOTS; ++s)
                v[u][s] = gathered_chunk<BF16>(rowp, L.goff[s]);
            }
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
              if (FUSED || !P.const_attention) {
                float dd = whr[s].x * v[u][s].x;
                dd = fmaf(whr[s].y, v[u][s].y, dd);
                dd = fmaf(whr[s].z, v[u][s].z, dd);
                dd = fmaf(whr[s].w, v[u][s].w, dd);
                prow[u * PSTRIDE + s * G] = dd;
              }
            }
          }
        } else {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const bool on = tt + u < tcnt;
            if (GS) {
              const float4* srow = stage + ((k0 + (on ? u : 0)) % kStageEdges) * P.chunks_per_head;
#pragma unroll
              for (int s = 0; s < SLOTS; ++s)
                v[u][s] = on ? srow[L.goff[s] >> 2] : make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
              const float* rowp = gp[on ? u : 0];
#pragma unroll
              for (int s = 0; s < SLOTS; ++s)
                v[u][s] = on ? gathered_chunk<BF16>(rowp, L.goff[s]) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (tt + u < tcnt) {
#pragma unroll
              for (int s = 0; s < SLOTS; ++s) {
                const float w = wp[s][u * NHT];
                acc[s].x = fmaf(w, v[u][s].x, acc[s].x);
                acc[s].y = fmaf(w, v[u][s].y, acc[s].y);
                acc[s].z = fmaf(w, v[u][s].z, acc[s].z);
                acc[s].w = fmaf(w, v[u][s].w, acc[s].w);
                if (FUSED || !P.const_attention) {
                  float dd = whr[s].x * v[u][s].x;
                  dd = fmaf(whr[s].y, v[u][s].y, dd);
                  dd = fmaf(whr[s].z, v[u][s].z, dd);
                  dd = fmaf(whr[s].w, v[u][s].w, dd);
                  prow[u * PSTRIDE + s * G] = dd;
                }
              }
            }
          }
        }
      }
      if ((FUSED || !P.const_attention)) {
        __syncwarp(gmask);
        // transpose-reduce: (edge, head) pair q sums the chunks of that head (four independent partial sums)
        for (int q = gl; q < tcnt * nh; q += G) {
          int t, h;
          if (nh == NHT) { t = q / NHT; h = q - t * NHT; } else { t = q / nh; h = q - t * nh; }
          const float* pp = part + t * PSTRIDE + h * P.chunks_per_head;
          float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
          int c = 0;
          for (; c + 4 <= P.chunks_per_head; c += 4) { d0 += pp[c]; d1 += pp[c + 1]; d2 += pp[c + 2]; d3 += pp[c + 3]; }
          for (; c < P.chunks_per_head; ++c) d0 += pp[c];
          sh_da[(gbase + t0 + t) * NHT + h] = (d0 + d1) + (d2 + d3);
        }
        __syncwarp(gmask);
      }
    }
    if (valid && (FUSED || !P.const_attention)) {
      if (FUSED) {   // g = 0.01*alpha*(d_alpha - S[dst]) = 0.01*((m*alpha)*<dOut,Wh> - alpha*S[dst]), summed per source row  
#pragma unroll
        for (int h = 0; h < NHT; ++h)
          if (h < nh) gsum[h] = fmaf(kLeakySlope, fmaf(sh_w[tid * NHT + h], sh_da[tid * NHT + h], -sh_s[tid * NHT + h]), gsum[h]);
      } else {
        float* r = P.rec + (int64_t)e * 2 * nh;
#pragma unroll
        for (int h = 0; h < NHT; ++h)
          if (h < nh) { r[h] = fmaf(msk[h], sh_da[tid * NHT + h], ga[h]); r[nh + h] = alpha[h]; }
      }
    }
    __syncwarp(gmask);
  }
  float* const drow = dwh_row_ptr(P, row);
  // FUSED epilogue inputs: ds_src = sum g - |T_src|*Gamma/|T|, ds_tgt -= |T_dst|*Gamma/|T| (gradient through max())
  float dss[NHT], dst_[NHT];
#pragma unroll
  for (int h = 0; h < NHT; ++h) { dss[h] = 0.f; dst_[h] = 0.f; }
  if (FUSED && !COOP && (FUSED || !P.const_attention)) {
#pragma unroll
    for (int h = 0; h < NHT; ++h) {
      if (h < nh) {
        const float g = group_sum<G>(gsum[h], gmask);
        const int src_lane = (tid & 31 & ~(G - 1)) + h;     // lane h of this group holds head h's prefetched values
        const int ts = __shfl_sync(gmask, p_ts, src_lane);
        const int td = __shfl_sync(gmask, p_td, src_lane);
        const float t = __shfl_sync(gmask, p_t, src_lane);
        dss[h] = ts ? g - (float)ts * corr : g;
        if (own_tgt) dst_[h] = td ? t - (float)td * corr : t;
      }
    }
    if (gl == 0) {
#pragma unroll
      for (int h = 0; h < NHT; ++h)
        if (h < nh) {
          P.ds_src[row * nh + h] = dss[h];
          if (own_tgt) P.ds_tgt[trow * nh + h] = dst_[h];
        }
    }
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      if ((FULL || L.ok[s])) {
        const int c = s * G + gl;
#pragma unroll
        for (int h = 0; h < NHT; ++h) {
          if (h < nh) {
            const float4 as = ldg4(P.a_src + (int64_t)h * P.dp + c * 4);
            const float4 at = ldg4(P.a_tgt + (int64_t)h * P.dp + c * 4);
            acc[s].x = fmaf(dss[h], as.x, fmaf(dst_[h], at.x, acc[s].x));
            acc[s].y = fmaf(dss[h], as.y, fmaf(dst_[h], at.y, acc[s].y));
            acc[s].z = fmaf(dss[h], as.z, fmaf(dst_[h], at.z, acc[s].z));
            acc[s].w = fmaf(dss[h], as.w, fmaf(dst_[h], at.w, acc[s].w));
          }
        }
      }
    }
  }
  if (COOP) {
    __syncthreads();   // every group is done with its `part` tile, which `coop` aliases
#pragma unroll
    for (int s = 0; s < SLOTS; ++s)
      if ((FULL || L.ok[s])) *reinterpret_cast<float4*>(coop + grp * P.dp + (s * G + gl) * 4) = acc[s];
    float* coop_g = coop + NG * P.dp;   // [NG][NHT] per-group sums of g
    if (FUSED && (FUSED || !P.const_attention)) {
#pragma unroll
      for (int h = 0; h < NHT; ++h) gsum[h] = group_sum<G>(gsum[h], gmask);
      if (gl == 0) {
#pragma unroll
        for (int h = 0; h < NHT; ++h) coop_g[grp * NHT + h] = gsum[h];
      }
    }
    __syncthreads();
    if (FUSED && (FUSED || !P.const_attention)) {
#pragma unroll
      for (int h = 0; h < NHT; ++h) {
        if (h < nh) {
          float g = 0.f;
          for (int j = 0; j < NG; ++j) g += coop_g[j * NHT + h];
          const int ts = P.tie_src ? __ldg(P.tie_src + row * nh + h) : 0;
          dss[h] = ts ? g - (float)ts * corr : g;
          if (own_tgt) {
            const int td = P.tie_dst ? __ldg(P.tie_dst + trow * nh + h) : 0;
            const float t = P.ds_tgt[trow * nh + h];
            dst_[h] = td ? t - (float)td * corr : t;
          }
        }
      }
      __syncthreads();   // every thread has read ds_tgt before thread 0 overwrites it
      if (tid == 0) {
#pragma unroll
        for (int h = 0; h < NHT; ++h)
          if (h < nh) {
            P.ds_src[row * nh + h] = dss[h];
            if (own_tgt) P.ds_tgt[trow * nh + h] = dst_[h];
          }
      }
    }
    for (int c = tid; c < P.chunks; c += kEdgeThreads) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j = 0; j < NG; ++j) {
        const float4 v = *reinterpret_cast<const float4*>(coop + j * P.dp + c * 4);
        t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
      }
      if (FUSED && (FUSED || !P.const_attention)) {
#pragma unroll
        for (int h = 0; h < NHT; ++h) {
          if (h < nh) {
            const float4 as = ldg4(P.a_src + (int64_t)h * P.dp + c * 4);
            const float4 at = ldg4(P.a_tgt + (int64_t)h * P.dp + c * 4);
            t.x = fmaf(dss[h], as.x, fmaf(dst_[h], at.x, t.x));
            t.y = fmaf(dss[h], as.y, fmaf(dst_[h], at.y, t.y));
            t.z = fmaf(dss[h], as.z, fmaf(dst_[h], at.z, t.z));
            t.w = fmaf(dss[h], as.w, fmaf(dst_[h], at.w, t.w));
          }
        }
      }
      *reinterpret_cast<float4*>(drow + c * 4) = t;
    }
    // the next grab_long_row() starts with a __syncthreads()
  } else if (G == 32 && P.rowbuf_offset_floats >= 0) {
    // bulk-copy push: lane 0 owns the bulk-async group of this warp; which of the warp's two row buffers is next lives in
    // shared memory (nothing about the push stays in registers across the row)
    extern __shared__ __align__(16) float dyn_smem_rb[];
    const int flip = sh_flip[tid >> 5];
    float* rb = dyn_smem_rb + P.rowbuf_offset_floats + ((tid >> 5) * 2 + flip) * P.dp;
    if (gl == 0) bulk_wait_read_keep1();          // the copy issued from this buffer two rows ago has read it
    __syncwarp(gmask);
#pragma unroll
    for (int s = 0; s < SLOTS; ++s)
      if ((FULL || L.ok[s])) *reinterpret_cast<float4*>(rb + (s * G + gl) * 4) = acc[s];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the bulk copy
    __syncwarp(gmask);
    if (gl == 0) { bulk_store(drow, rb, (unsigned)(P.dp * sizeof(float))); sh_flip[tid >> 5] = flip ^ 1; }
  } else {
#pragma unroll
    for (int s = 0; s < SLOTS; ++s)
      if ((FULL || L.ok[s])) *reinterpret_cast<float4*>(drow + (s * G + gl) * 4) = acc[s];
  }
}

template <int G, int SLOTS, int NHT, bool COOP, bool FUSED, bool FULL, bool GS = false, bool BF16 = false>
__global__ void __launch_bounds__(kEdgeThreads, (SLOTS <= 2 ? 3 : (SLOTS <= 4 ? 2 : 1)))
edge_bwd_main_kernel(const BwdMainParams P) {
  constexpr int TB = MainShape<G, SLOTS>::TB;
  extern __shared__ __align__(16) float dyn_smem[];
  __shared__ int sh_dst[kEdgeThreads];
  __shared__ float sh_w[kEdgeThreads * NHT];
  __shared__ float sh_da[kEdgeThreads * NHT];
  __shared__ const float* sh_gp[kEdgeThreads];
  __shared__ float sh_s[FUSED ? kEdgeThreads * NHT : 1];
  const int tid = threadIdx.x, lane = tid & 31, gl = tid & (G - 1), gbase = tid - gl;
  const unsigned gmask = group_mask<G>(lane);
  const LaneShape<SLOTS> L = make_lane_shape<G, SLOTS>(P, gl);
  float* part = dyn_smem + (size_t)(tid / G) * TB * MainShape<G, SLOTS>::PSTRIDE;   // [TB][PSTRIDE] of my group
  // GS: per-group staging area behind the part / coop region (main_dyn_smem(), 16-byte aligned)
  float4* stage = GS ? reinterpret_cast<float4*>(dyn_smem + P.stage_offset_floats) + (size_t)(tid / G) * kStageEdges * kStageMaxChunks : nullptr;
  __shared__ int sh_flip[kEdgeThreads / 32];
  if (lane == 0) sh_flip[tid >> 5] = 0;
  __syncwarp();
  const float gmax = P.const_attention ? 0.f : __ldg(P.gmax);
  const float corr = !FUSED ? 0.f : (P.corr_override ? __ldg(P.corr_override) : P.header->corr);
  if (COOP) {   // long source rows, CTA per row (its own launch)
    __shared__ int sh_ctl;
    pdl_release_dependents();   // the short-row launch that follows may fill SMs as this grid drains
    for (;;) {
      const int64_t row = grab_long_row(P.sched, P.rowptr_t, &sh_ctl);
      if (row < 0) break;
      bwd_main_row<G, SLOTS, NHT, true, FUSED, FULL, GS, BF16>(P, L, row, __ldg(P.rowptr_t + row), __ldg(P.rowptr_t + row + 1), tid, gl, gbase, gmask, gmax, corr, sh_dst, sh_gp, sh_w, sh_da, sh_s, part, dyn_smem, stage, sh_flip);
    }
  } else {
    int64_t base;
    while (grab_rows<G>(P.sched, lane, base)) {
      int pr, ps, pe;
      prefetch_rows<G>(P.sched, P.rowptr_t, base, lane, pr, ps, pe, P.sched_rot);
#pragma unroll 1
      for (int k = 0; k < kGrabIters<G>; ++k) {
        int64_t row;
        int start, end;
        if (prefetched_row<G>(P.sched, k, lane, pr, ps, pe, row, start, end))
          bwd_main_row<G, SLOTS, NHT, false, FUSED, FULL, GS, BF16>(P, L, row, start, end, tid, gl, gbase, gmask, gmax, corr, sh_dst, sh_gp, sh_w, sh_da,
                                                              sh_s, part, nullptr, stage, sh_flip);
      }
    }
    if (G == 32 && P.rowbuf_offset_floats >= 0 && lane == 0) bulk_wait_all();   // every pushed row has left before the grid may complete
    pdl_wait_for_primary();     // no-op unless launched behind the cooperative kernel
  }
}

// per-group transpose tiles; the cooperative path's (256/G) x dp reduction buffer aliases them
template <int G, int SLOTS>
static size_t main_dyn_smem(int chunks) {
  const size_t part = (size_t)(kEdgeThreads / G) * MainShape<G, SLOTS>::TB * MainShape<G, SLOTS>::PSTRIDE * sizeof(float);
  const size_t coop = (size_t)(kEdgeThreads / G) * (chunks * 4 + 8) * sizeof(float);   // + [NG][NHT] sums of g (FUSED)
  return part > coop ? part : coop;
}

}  // namespace gat
#include "edge_bwd_hm.cuh"   // head-mean layers with four heads: lane = (edge slot, head, quarter)
namespace gat {

// ------------------------------------------------------------------------------------------------------------
// pass 2: per target row, S[h] = sum_e alpha*d_alpha; ds_tgt[h] = 0.01 * S * eps / (Z + eps)
// ------------------------------------------------------------------------------------------------------------
struct BwdRowsumParams {
  const int32_t* rowptr; const int32_t* tpos; RowSched sched; int nh;
  const float* rec; const float* z; float* s_sum; float* ds_tgt;
};

__device__ __forceinline__ void rowsum_span(const BwdRowsumParams& P, const int start, const int end, const int first,
                                            const int step, float (&s)[kMaxHeads]) {
  const int nh = P.nh;
#pragma unroll
  for (int h = 0; h < kMaxHeads; ++h) s[h] = 0.f;
  for (int j = start + first; j < end; j += step) {
    const float* r = P.rec + (int64_t)__ldg(P.tpos + j) * 2 * nh;
    if (nh == 4) {
      const float4 da = __ldg(reinterpret_cast<const float4*>(r)), al = __ldg(reinterpret_cast<const float4*>(r) + 1);
      s[0] = fmaf(al.x, da.x, s[0]); s[1] = fmaf(al.y, da.y, s[1]); s[2] = fmaf(al.z, da.z, s[2]); s[3] = fmaf(al.w, da.w, s[3]);
    } else {
#pragma unroll
      for (int h = 0; h < kMaxHeads; ++h)
        if (h < nh) s[h] = fmaf(__ldg(r + nh + h), __ldg(r + h), s[h]);
    }
  }
#pragma unroll
  for (int h = 0; h < kMaxHeads; ++h) {
    if (h < nh) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s[h] += __shfl_xor_sync(0xffffffffu, s[h], o);
    }
  }
}

__device__ __forceinline__ void rowsum_store(const BwdRowsumParams& P, const int64_t row, const float (&s)[kMaxHeads]) {
#pragma unroll
  for (int h = 0; h < kMaxHeads; ++h) {
    if (h < P.nh) {
      const float zz = __ldg(P.z + row * P.nh + h);
      P.s_sum[row * P.nh + h] = s[h];
      // sum_e g = 0.01*(S - S*sum_e alpha) with sum_e alpha = Z/(Z+eps): exact, and free of the cancellation a
      // direct fp32 sum of g would suffer
      P.ds_tgt[row * P.nh + h] = kLeakySlope * s[h] * (kSoftmaxEps / (zz + kSoftmaxEps));
    }
  }
}

__global__ void __launch_bounds__(256)
edge_bwd_rowsum_kernel(const BwdRowsumParams P) {
  __shared__ float sh_part[8][kMaxHeads];
  __shared__ int sh_ctl;
  const int tid = threadIdx.x, lane = tid & 31;
  float s[kMaxHeads];
  // long rows: the 8 warps stride over the row, warp partials are added in warp order
  for (;;) {
    const int64_t row = grab_long_row(P.sched, P.rowptr, &sh_ctl);
    if (row < 0) break;
    rowsum_span(P, __ldg(P.rowptr + row), __ldg(P.rowptr + row + 1), tid, 256, s);
    if (lane == 0) {
#pragma unroll
      for (int h = 0; h < kMaxHeads; ++h) sh_part[tid >> 5][h] = s[h];
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
      for (int h = 0; h < kMaxHeads; ++h) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += sh_part[w][h];
        s[h] = t;
      }
      rowsum_store(P, row, s);
    }
  }
  int64_t base;
  while (grab_rows<32>(P.sched, lane, base)) {
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
      const int64_t row = sched_row<32>(P.sched, base, k, lane);
      if (row < 0 || taken_by_cta_phase(P.sched, P.rowptr, row)) continue;
      rowsum_span(P, __ldg(P.rowptr + row), __ldg(P.rowptr + row + 1), lane, 32, s);
      if (lane == 0) rowsum_store(P, row, s);
    }
  }
}

// pass 2, common case (no upstream dL/dalpha): sum_e alpha*d_alpha = sum_e m*alpha*<dOut[d,h,:], Wh[src,h,:]>
//   = <dOut[d,h,:], sum_e m*alpha*Wh[src,h,:]> = <dOut[d,h,:], out[d,h,:]>  -- a per-node dot product of the upstream gradient
// with the forward output, so no per-edge record has to be gathered at all.
struct BwdRowdotParams {
  const float* go; const float* out; const float* z; int64_t n; int nh; int dp; int chunks; int chunks_per_head;
  int go_ld; int go_shared;
  // out_is_act: `out` holds ELU(out) (gat_edge_fwd out_act) and `go` is the gradient w.r.t. that activated output; the pass
  // then recovers out = h > 0 ? h : log1p(h) and ELU'(out) = h > 0 ? 1 : h + 1, forms the gradient w.r.t. the pre-activation
  // output go*ELU', writes it to go_out (what the source-major pass gathers) and uses it in S -- the ELU backward fused.
  int out_is_act; float* go_out;
  // the rest of the output glue (common.cuh): `out` holds y = keep * E(out + skip); `glue` is set when any part is active
  int glue; const float* skip; int64_t ld_skip; float drop_p; uint64_t drop_seed;
  // fused attention-norm regulariser: S[d,h] += c * deg(d) * norm_t[d,h]; deg(d) is also packed into the target's record
  const int32_t* rowptr; const float* norm_t; const float* norm_coef; float norm_scale;
  float* s_sum; float* ds_tgt;
  const float* s_tgt; float* tpack;   // optional: write the per-target record {s_tgt | Z | S} for gat_edge_bwd_fused
};

// out = log1p(h) for h in (-1, 0], cheaply: the series where 1 + h would round away h's low bits, the fast log elsewhere
// (|error| <= ~1e-7 absolute; the term it enters, go*ELU'*out, is scaled by ELU' = h + 1 wherever out is large).
__device__ __forceinline__ float log1p_neg_fast(float h) {
  return h > -1e-3f ? h * fmaf(h, fmaf(h, 0.33333334f, -0.5f), 1.0f) : __logf(1.0f + h);
}

// FULLGLUE: skip rows and / or output dropout are part of the glue (Philox per chunk: its registers stay out of the plain
// instantiation, which the headline's hidden layers run with the ELU adjoint only).
template <int NHT, bool FULLGLUE>
__global__ void __launch_bounds__(256, (NHT == 4 && !FULLGLUE) ? 3 : 2)
edge_bwd_rowdot_kernel(const BwdRowdotParams P) {
  constexpr int R = 4;   // rows per warp iteration: 2*R*chunks/32 independent 16-byte loads in flight per lane
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * 8;
  const int nh = P.nh;
  // head / in-head chunk of this lane's chunk c = lane + 32k, advanced without a division per chunk (the two runtime
  // divisions per chunk made this streaming pass instruction-bound: 52 % issue slots busy at 23-54 % of the DRAM rate)
  const int cph = P.chunks_per_head;
  const int hh0 = lane / cph, gc0 = lane - hh0 * cph, q32 = 32 / cph, r32 = 32 - q32 * cph;
  for (int64_t row0 = warp * R; row0 < P.n; row0 += nwarps * R) {
    float s[R][NHT];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int h = 0; h < NHT; ++h) s[r][h] = 0.f;
    // this lane's (row, head) pair of the epilogue below: its scalars are requested now, used after the streaming loop
    float pre_z = 1.f, pre_t = 0.f, pre_nt = 0.f, pre_deg = 0.f, pre_cdeg = 0.f;
    {
      constexpr int V0 = R * NHT;
      const int v0 = V0 == 32 ? (((lane >> 4) & 1) << 4 | ((lane >> 3) & 1) << 3 | ((lane >> 2) & 1) << 2 | ((lane >> 1) & 1) << 1 | (lane & 1))
                              : (((lane >> 4) & 1) << 3 | ((lane >> 3) & 1) << 2 | ((lane >> 2) & 1) << 1 | ((lane >> 1) & 1));
      const int pr = v0 / NHT, ph = v0 - pr * NHT;
      const int64_t prow = row0 + pr;
      if (prow < P.n && ph < nh) {
        pre_z = __ldg(P.z + prow * nh + ph);
        if (P.tpack) pre_t = __ldg(P.s_tgt + prow * nh + ph);
        if (P.rowptr) {
          pre_deg = (float)(__ldg(P.rowptr + prow + 1) - __ldg(P.rowptr + prow));
          if (P.norm_coef) { pre_cdeg = __ldg(P.norm_coef) * P.norm_scale * pre_deg; pre_nt = __ldg(P.norm_t + prow * nh + ph); }
        }
      }
    }
    int hh = hh0 - q32, gc = gc0 - r32;
    for (int c = lane; c < P.chunks; c += 32) {
      hh += q32; gc += r32;
      if (gc >= cph) { gc -= cph; ++hh; }
      float4 g[R], o[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const bool on = row0 + r < P.n;
        g[r] = on ? ldg4(P.go + (row0 + r) * P.go_ld + (P.go_shared ? gc : c) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        o[r] = on ? ldg4(P.out + (row0 + r) * P.dp + c * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (!FULLGLUE && P.out_is_act) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          float* gv = reinterpret_cast<float*>(&g[r]);
          float* ov = reinterpret_cast<float*>(&o[r]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float h = ov[j];
            const bool neg = h <= 0.f;
            gv[j] = neg ? gv[j] * (h + 1.0f) : gv[j];      // dL/dout = dL/dh * ELU'(out)
            ov[j] = neg ? log1p_neg_fast(h) : h;           // out recovered from h = ELU(out)
          }
          if (row0 + r < P.n) *reinterpret_cast<float4*>(P.go_out + (row0 + r) * P.dp + c * 4) = g[r];
        }
      }
      if (FULLGLUE) {
        const float omp = 1.0f - P.drop_p;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          float* gv = reinterpret_cast<float*>(&g[r]);
          float* ov = reinterpret_cast<float*>(&o[r]);
          const bool on = row0 + r < P.n;
          float4 k4 = make_float4(1.f, 1.f, 1.f, 1.f), s4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (P.drop_p > 0.f) k4 = glue_keep4(P.drop_seed, row0 + r, c, P.drop_p);
          if (P.skip && on) s4 = ldg4(P.skip + (row0 + r) * P.ld_skip + c * 4);
          const float* kv = reinterpret_cast<const float*>(&k4);
          const float* sv = reinterpret_cast<const float*>(&s4);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float h = ov[j] * omp;                   // E(out + skip) where the element was kept
            const bool neg = P.out_is_act && h <= 0.f;
            gv[j] = glue_adjoint1(gv[j], ov[j], P.out_is_act, kv[j], omp);      // dL/d(out + skip)
            // out recovered from the stored value; a dropped element's gradient is 0, so its (unknown) out never matters
            ov[j] = kv[j] == 0.f ? 0.f : (neg ? log1p_neg_fast(h) : h) - sv[j];
          }
          if (on) *reinterpret_cast<float4*>(P.go_out + (row0 + r) * P.dp + c * 4) = g[r];
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float d = fmaf(g[r].x, o[r].x, fmaf(g[r].y, o[r].y, fmaf(g[r].z, o[r].z, g[r].w * o[r].w)));
#pragma unroll
        for (int h = 0; h < NHT; ++h) s[r][h] += (h == hh) ? d : 0.f;
      }
    }
    // ---- the R*NHT per-(row, head) sums of the 32 lanes are TRANSPOSE-REDUCED: at butterfly distance o a lane keeps the half of
    // its values whose index has the matching bit and sends the other half, so R*NHT - 1 (+1) shuffles reduce all values at once
    // (16 for 16 values; the plain butterfly takes 80) and lane L ends up with the COMPLETE sum of ONE (row, head) pair -- whose
    // epilogue it then runs itself, in parallel with the others: the scalars it needs were requested before the streaming loop,
    // its stores are contiguous across lanes.  (Was: 80 shuffles, then lane 0 alone loaded, computed and stored all 16 pairs
    // while the warp waited; the pass ran at 0.6 of the DRAM rate with half its issue slots idle.)
    constexpr int V = R * NHT;                     // 16 or 32 values, index v = r*NHT + h
    float val[V];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int h = 0; h < NHT; ++h) val[r * NHT + h] = s[r][h];
    int width = V;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      if (width > 1) {
        const bool up = (lane & o) != 0;           // keep the upper half of the indices
        const int half = width / 2;
#pragma unroll
        for (int i = 0; i < V / 2; ++i) {
          if (i < half) {
            const float send = up ? val[i] : val[i + half];
            const float keep = up ? val[i + half] : val[i];
            val[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
          }
        }
        width = half;
      } else {
        val[0] += __shfl_xor_sync(0xffffffffu, val[0], o);     // V = 16: the last stage sums two lanes holding the same index
      }
    }
    // lane -> value index: the stages at o = 16, 8, 4, 2 (, 1) consumed index bits from the top
    const int vidx = V == 32 ? (((lane >> 4) & 1) << 4 | ((lane >> 3) & 1) << 3 | ((lane >> 2) & 1) << 2 | ((lane >> 1) & 1) << 1 | (lane & 1))
                             : (((lane >> 4) & 1) << 3 | ((lane >> 3) & 1) << 2 | ((lane >> 2) & 1) << 1 | ((lane >> 1) & 1));
    const int er = vidx / NHT, eh = vidx - er * NHT;
    const int64_t erow = row0 + er;
    if ((V == 32 || (lane & 1) == 0) && erow < P.n && eh < nh) {
      float sv = val[0];
      if (P.norm_t) sv = fmaf(pre_cdeg, pre_nt, sv);
      P.s_sum[erow * nh + eh] = sv;
      P.ds_tgt[erow * nh + eh] = kLeakySlope * sv * (kSoftmaxEps / (pre_z + kSoftmaxEps));
      if (P.tpack) {
        float* pk = P.tpack + erow * (4 * NHT);
        pk[eh] = pre_t; pk[NHT + eh] = pre_z; pk[2 * NHT + eh] = sv;
        if (P.rowptr && eh == 0) pk[3 * NHT] = pre_deg;
      }
    } else if ((V == 32 || (lane & 1) == 0) && erow < P.n && P.tpack) {   // head slots beyond nh: defined (zero) record fields
      float* pk = P.tpack + erow * (4 * NHT);
      pk[eh] = 0.f; pk[NHT + eh] = 1.f; pk[2 * NHT + eh] = 0.f;
    }
  }
}

// Gamma = sum over all (row, head) of ds_tgt, reduced in two fixed-order stages (independent of the dynamic schedule).
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

// ------------------------------------------------------------------------------------------------------------
// pass 3
// ------------------------------------------------------------------------------------------------------------
struct BwdFinishParams {
  const int32_t* rowptr_t; const int32_t* col_t; RowSched sched;
  int nh; int dp; int chunks;
  const float* rec; const float* s_sum;              // s_sum indexed by TARGET id
  const float* a_src; const float* a_tgt;
  const int32_t* tie_dst; const int32_t* tie_src; const BwdHeader* header; const float* corr_override;
  int64_t tgt_lo; int64_t tgt_hi;   // rows this call owns as TARGETS (ds_tgt / tie_dst are indexed row - tgt_lo)
  float* ds_src; float* ds_tgt; float* d_wh;
};

// COOP (long source rows): all 256 threads stride over the row's records; the per-head sums are combined warp by warp
// in a fixed order through `sh_part`, then group 0 alone runs the row epilogue.  Called by all threads in that case.
template <int G, int SLOTS, int NHT, bool COOP>
__device__ __forceinline__ void bwd_finish_row(const BwdFinishParams& P, const int64_t row, const int gl,
                                               const unsigned gmask, const float corr, float (*sh_part)[kMaxHeads]) {
  const int nh = P.nh;
  const int tid = threadIdx.x;
  const int start = __ldg(P.rowptr_t + row), end = __ldg(P.rowptr_t + row + 1);
  const bool epilogue = !COOP || tid < G;
  float4 own[SLOTS];   // the row to update: issued first so its latency overlaps the edge loop
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int c = s * G + gl;
    own[s] = (epilogue && c < P.chunks) ? *reinterpret_cast<const float4*>(P.d_wh + row * P.dp + c * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float gsum[NHT];
#pragma unroll
  for (int h = 0; h < NHT; ++h) gsum[h] = 0.f;
  for (int e = start + (COOP ? tid : gl); e < end; e += (COOP ? kEdgeThreads : G)) {
    const float* r = P.rec + (int64_t)e * 2 * nh;
    const float* sp = P.s_sum + (int64_t)__ldg(P.col_t + e) * nh;
#pragma unroll
    for (int h = 0; h < NHT; ++h)
      if (h < nh) gsum[h] = fmaf(kLeakySlope * __ldg(r + nh + h), __ldg(r + h) - __ldg(sp + h), gsum[h]);   // g = 0.01*alpha*(d_alpha - S)
  }
  if (COOP) {
#pragma unroll
    for (int h = 0; h < NHT; ++h) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) gsum[h] += __shfl_xor_sync(0xffffffffu, gsum[h], o);
    }
    if ((tid & 31) == 0) {
#pragma unroll
      for (int h = 0; h < NHT; ++h) sh_part[tid >> 5][h] = gsum[h];
    }
    __syncthreads();
    if (!epilogue) return;
#pragma unroll
    for (int h = 0; h < NHT; ++h) {
      float t = 0.f;
      for (int w = 0; w < kEdgeThreads / 32; ++w) t += sh_part[w][h];
      gsum[h] = t;
    }
  }
  // ds_src = sum g - |T_src|*Gamma/|T|;  ds_tgt -= |T_dst|*Gamma/|T|   (gradient through max(), section 9.2)
  const bool own_tgt = row >= P.tgt_lo && row < P.tgt_hi;   // single GPU: always; partitioned: owner rank only
  const int64_t trow = row - P.tgt_lo;
  float dss[NHT], dst_[NHT];
#pragma unroll
  for (int h = 0; h < NHT; ++h) {
    dss[h] = 0.f; dst_[h] = 0.f;
    if (h < nh) {
      float g = COOP ? gsum[h] : group_sum<G>(gsum[h], gmask);
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
  // d_wh_total = d_wh + ds_src * A_src + ds_tgt * A_tgt   (in place on the row this group owns)
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int c = s * G + gl;
    if (c < P.chunks) {
      float4* dp4 = reinterpret_cast<float4*>(P.d_wh + row * P.dp + c * 4);
      float4 a = own[s];
#pragma unroll
      for (int h = 0; h < NHT; ++h) {
        if (h < nh) {
          const float4 as = ldg4(P.a_src + (int64_t)h * P.dp + c * 4);
          const float4 at = ldg4(P.a_tgt + (int64_t)h * P.dp + c * 4);
          a.x = fmaf(dss[h], as.x, fmaf(dst_[h], at.x, a.x));
          a.y = fmaf(dss[h], as.y, fmaf(dst_[h], at.y, a.y));
          a.z = fmaf(dss[h], as.z, fmaf(dst_[h], at.z, a.z));
          a.w = fmaf(dss[h], as.w, fmaf(dst_[h], at.w, a.w));
        }
      }
      *dp4 = a;
    }
  }
}

template <int G, int SLOTS, int NHT>
__global__ void __launch_bounds__(kEdgeThreads)
edge_bwd_finish_kernel(const BwdFinishParams P) {
  __shared__ float sh_part[kEdgeThreads / 32][kMaxHeads];
  __shared__ int sh_ctl;
  const int tid = threadIdx.x, lane = tid & 31, gl = tid & (G - 1);
  const unsigned gmask = group_mask<G>(lane);
  const float corr = P.corr_override ? __ldg(P.corr_override) : P.header->corr;
  for (;;) {
    const int64_t row = grab_long_row(P.sched, P.rowptr_t, &sh_ctl);
    if (row < 0) break;
    bwd_finish_row<G, SLOTS, NHT, true>(P, row, gl, gmask, corr, sh_part);
  }
  int64_t base;
  while (grab_rows<G>(P.sched, lane, base)) {
#pragma unroll 1
    for (int k = 0; k < kGrabIters<G>; ++k) {
      const int64_t row = sched_row<G>(P.sched, base, k, lane);
      if (row >= 0 && !taken_by_cta_phase(P.sched, P.rowptr_t, row))
        bwd_finish_row<G, SLOTS, NHT, false>(P, row, gl, gmask, corr, nullptr);
    }
  }
}

// Owner-side half of the fused reduce-scatter: out[r] = sum over slabs q (ranks, fixed order) of recv[q][r].
__global__ void __launch_bounds__(256)
slab_sum_kernel(const float* __restrict__ recv, int n_slabs, int64_t slab_floats, int64_t count4, float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 t = __ldg(reinterpret_cast<const float4*>(recv) + i);
    for (int q = 1; q < n_slabs; ++q) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(recv + q * slab_floats) + i);
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    reinterpret_cast<float4*>(out)[i] = t;
  }
}

static int check_common(const char* who, int nh, int fp, void* workspace, size_t workspace_bytes) {
  if (nh < 1 || nh > kMaxHeads) { set_error("%s: num_heads %d not in [1, %d]", who, nh, kMaxHeads); return GAT_EINVAL; }
  if (fp <= 0 || fp % 4) { set_error("%s: padded head width %d must be a positive multiple of 4", who, fp); return GAT_EINVAL; }
  if (workspace == nullptr || workspace_bytes < gat_edge_bwd_workspace_bytes(0, 0, nh)) {
    set_error("%s: workspace too small", who);
    return GAT_EWORKSPACE;
  }
  return GAT_OK;
}

}  // namespace gat

extern "C" size_t gat_edge_bwd_workspace_bytes(int64_t n, int64_t n_edges, int nh) {
  (void)n; (void)n_edges; (void)nh;
  return gat::kBwdHeaderBytes + (size_t)(gat::kGammaBlocks + 1) * sizeof(double);
}

namespace gat {

template <bool FUSED>
static int launch_bwd_main(const BwdMainParams& P, const int32_t* row_order_t, int64_t n_long, int64_t n_rows, cudaStream_t st) {
  const int nh = P.nh;
  GroupShape shape = pick_group(P.chunks, n_rows);
  if (shape.slots < 0) {
    set_error("gat_edge_bwd_main: row width %d floats exceeds the supported 1024", P.dp);
    return GAT_EUNSUPPORTED;
  }
  const bool coop_launch = row_order_t != nullptr && n_long != 0;
  const bool full = FUSED && P.chunks == shape.g * shape.slots && !P.go_shared;   // predicate-free instantiation
  BwdMainParams Q = P;
  // bulk-copy push (see BwdMainParams::rowbuf_offset_floats): two row buffers per warp behind everything else
  const size_t push_bytes = (P.push && shape.g == 32) ? (size_t)(kEdgeThreads / 32) * 2 * P.dp * sizeof(float) : 0;
  Q.rowbuf_offset_floats = -1;
  Q.sched_rot = P.push ? (((int64_t)P.my_rank + 1) * P.rows_per_rank) % (n_rows > 0 ? n_rows : 1) : 0;
  if (P.go_bf16) {   // bf16 variant: fused path, wide unshared rows only (the shapes it is meant for)
    if (!(FUSED && !P.go_shared && shape.g == 32 && shape.slots == 2 && nh <= 4)) {
      set_error("gat_edge_bwd_fused_bf16: supported for NH <= 4, unshared gradient rows of 132..256 floats (got NH = %d, %d floats)", nh, P.dp);
      return GAT_EUNSUPPORTED;
    }
#define LAUNCH_BF16(S_)                                                                                                \
    do {                                                                                                               \
      const size_t base_ = (main_dyn_smem<32, S_>(P.chunks) + 15) / 16 * 16;                                           \
      const size_t smem_ = base_ + push_bytes;                                                                         \
      if (push_bytes) Q.rowbuf_offset_floats = (int)(base_ / sizeof(float));                                           \
      static size_t optin_bf_dev_[kMaxDevices] = {0};                                                                  \
      size_t& optin_bf_ = optin_bf_dev_[cur_device()];                                                                 \
      if (smem_ > optin_bf_) {                                                                                         \
        optin_bf_ = smem_;                                                                                             \
        GAT_CUDA(cudaFuncSetAttribute(edge_bwd_main_kernel<32, S_, 4, true, FUSED, false, false, true>,                \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)optin_bf_));                   \
        GAT_CUDA(cudaFuncSetAttribute(edge_bwd_main_kernel<32, S_, 4, false, FUSED, false, false, true>,               \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)optin_bf_));                   \
      }                                                                                                                \
      if (coop_launch) {                                                                                               \
        GAT_CUDA(launch_kernel(edge_bwd_main_kernel<32, S_, 4, true, FUSED, false, false, true>,                       \
                               persistent_grid(edge_bwd_main_kernel<32, S_, 4, true, FUSED, false, false, true>, kEdgeThreads, smem_, \
                                               n_long < 0 ? n_rows : n_long),                                          \
                               kEdgeThreads, smem_, st, Q, false));                                                    \
        GAT_LAUNCH_CHECK();                                                                                            \
      }                                                                                                                \
      GAT_CUDA(launch_kernel(edge_bwd_main_kernel<32, S_, 4, false, FUSED, false, false, true>,                        \
                             persistent_grid(edge_bwd_main_kernel<32, S_, 4, false, FUSED, false, false, true>, kEdgeThreads, smem_, \
                                             (n_rows + 7) / 8),                                                        \
                             kEdgeThreads, smem_, st, Q, coop_launch));                                                \
      GAT_LAUNCH_CHECK();                                                                                              \
    } while (0)
    LAUNCH_BF16(2);
#undef LAUNCH_BF16
    return GAT_OK;
  }
  // GS: narrow shared upstream-gradient rows are staged through shared memory (see bwd_main_row)
  if (FUSED && P.go_shared && shape.g == 32 && shape.slots <= 2 && nh <= 4 && P.chunks_per_head <= kStageMaxChunks) {
#define LAUNCH_GS(S_)                                                                                                  \
    do {                                                                                                               \
      const size_t base_ = (main_dyn_smem<32, S_>(P.chunks) + 15) / 16 * 16;                                           \
      const size_t stage_ = (size_t)(kEdgeThreads / 32) * kStageEdges * kStageMaxChunks * sizeof(float4);              \
      const size_t smem_ = base_ + stage_ + push_bytes;                                                                \
      Q.stage_offset_floats = (int)(base_ / sizeof(float));                                                            \
      if (push_bytes) Q.rowbuf_offset_floats = (int)((base_ + stage_) / sizeof(float));                                \
      static bool optin_gs_dev_[kMaxDevices] = {false};                                                                \
      bool& optin_gs_ = optin_gs_dev_[cur_device()];                                                                   \
      if (!optin_gs_) {                                                                                                \
        GAT_CUDA(cudaFuncSetAttribute(edge_bwd_main_kernel<32, S_, 4, true, FUSED, false, true>,                       \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(96 * 1024)));                 \
        GAT_CUDA(cudaFuncSetAttribute(edge_bwd_main_kernel<32, S_, 4, false, FUSED, false, true>,                      \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(96 * 1024)));                 \
        optin_gs_ = true;                                                                                              \
      }                                                                                                                \
      if (coop_launch) {                                                                                               \
        GAT_CUDA(launch_kernel(edge_bwd_main_kernel<32, S_, 4, true, FUSED, false, true>,                              \
                               persistent_grid(edge_bwd_main_kernel<32, S_, 4, true, FUSED, false, true>, kEdgeThreads, smem_, \
                                               n_long < 0 ? n_rows : n_long),                                          \
                               kEdgeThreads, smem_, st, Q, false));                                                    \
        GAT_LAUNCH_CHECK();                                                                                            \
      }                                                                                                                \
      if (hm_cpl == 0) {                                                                                               \
        GAT_CUDA(launch_kernel(edge_bwd_main_kernel<32, S_, 4, false, FUSED, false, true>,                             \
                               persistent_grid(edge_bwd_main_kernel<32, S_, 4, false, FUSED, false, true>, kEdgeThreads, smem_, \
                                               (n_rows + 7) / 8),                                                      \
                               kEdgeThreads, smem_, st, Q, coop_launch));                                              \
        GAT_LAUNCH_CHECK();                                                                                            \
      }                                                                                                                \
    } while (0)
    // short rows of a four-head head-mean layer: the (edge slot, head, quarter) kernel of edge_bwd_hm.cuh (the long rows keep
    // the cooperative launch above)
    int hm_cpl = (FUSED && nh == 4 && P.tpack != nullptr && getenv("GAT_BWD_HM_OFF") == nullptr) ? (P.chunks_per_head + 3) / 4 : 0;
    if (hm_cpl < 2 || hm_cpl > 4) hm_cpl = 0;     // 5..16 chunks per head (narrower rows: the general kernel's lane grid fits them)
    if (shape.slots == 1) LAUNCH_GS(1); else LAUNCH_GS(2);
#undef LAUNCH_GS
#define LAUNCH_HM(C_, X_)                                                                                              \
    do {                                                                                                               \
      const size_t hm_smem_ = HmShape<C_>::kSmem + push_bytes;                                                         \
      BwdMainParams H = Q;                                                                                             \
      H.rowbuf_offset_floats = push_bytes ? (int)(HmShape<C_>::kSmem / sizeof(float)) : -1;                            \
      static size_t optin_hm_dev_[kMaxDevices] = {0};                                                                  \
      size_t& optin_hm_ = optin_hm_dev_[cur_device()];                                                                 \
      if (hm_smem_ > optin_hm_) {                                                                                      \
        optin_hm_ = hm_smem_;                                                                                          \
        GAT_CUDA(cudaFuncSetAttribute(edge_bwd_hm4_kernel<C_, X_>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                      (int)optin_hm_));                                                                \
      }                                                                                                                \
      GAT_CUDA(launch_kernel(edge_bwd_hm4_kernel<C_, X_>,                                                              \
                             persistent_grid(edge_bwd_hm4_kernel<C_, X_>, kEdgeThreads, hm_smem_, (n_rows + 7) / 8),   \
                             kEdgeThreads, hm_smem_, st, H, coop_launch));                                             \
      GAT_LAUNCH_CHECK();                                                                                              \
    } while (0)
    if (hm_cpl == 3 && P.chunks_per_head == 12) LAUNCH_HM(3, true);
    else if (hm_cpl == 2) LAUNCH_HM(2, false); else if (hm_cpl == 3) LAUNCH_HM(3, false); else if (hm_cpl == 4) LAUNCH_HM(4, false);
#undef LAUNCH_HM
    return GAT_OK;
  }
#define LAUNCH_BOTH(G_, S_, N_, FULL_)                                                                                 \
  do {                                                                                                                 \
    /* static + dynamic shared memory can exceed the 48 KB default (wide rows, 8 heads): opt in once per size */       \
    static size_t optin_dev_[kMaxDevices] = {0};                                                                       \
    size_t& optin_ = optin_dev_[cur_device()];                                                                         \
    const size_t base_ = (main_dyn_smem<G_, S_>(P.chunks) + 15) / 16 * 16;                                             \
    const size_t smem_ = base_ + push_bytes;                                                                           \
    if (push_bytes) Q.rowbuf_offset_floats = (int)(base_ / sizeof(float));                                             \
    if (smem_ > optin_) {                                                                                              \
      optin_ = smem_;                                                                                                  \
      GAT_CUDA(cudaFuncSetAttribute(edge_bwd_main_kernel<G_, S_, N_, true, FUSED, FULL_>,                              \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)optin_));                        \
      GAT_CUDA(cudaFuncSetAttribute(edge_bwd_main_kernel<G_, S_, N_, false, FUSED, FULL_>,                             \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)optin_));                        \
    }                                                                                                                  \
    if (coop_launch) {   /* long source rows first, CTA per row; the short-row launch overlaps its tail */             \
      GAT_CUDA(launch_kernel(edge_bwd_main_kernel<G_, S_, N_, true, FUSED, FULL_>,                                     \
                             persistent_grid(edge_bwd_main_kernel<G_, S_, N_, true, FUSED, FULL_>, kEdgeThreads,       \
                                             smem_, n_long < 0 ? n_rows : n_long),                                     \
                             kEdgeThreads, smem_, st, Q, false));                                                      \
      GAT_LAUNCH_CHECK();                                                                                              \
    }                                                                                                                  \
    GAT_CUDA(launch_kernel(edge_bwd_main_kernel<G_, S_, N_, false, FUSED, FULL_>,                                      \
                           persistent_grid(edge_bwd_main_kernel<G_, S_, N_, false, FUSED, FULL_>, kEdgeThreads,        \
                                           smem_,                                                                      \
                                           (n_rows + (kEdgeThreads / G_) - 1) / (kEdgeThreads / G_)),                  \
                           kEdgeThreads, smem_, st, Q, coop_launch));                                                  \
    GAT_LAUNCH_CHECK();                                                                                                \
  } while (0)
#define LAUNCH(G_, S_, N_)                                                                                             \
  do {                                                                                                                 \
    if (FUSED && full) LAUNCH_BOTH(G_, S_, N_, FUSED);                                                                 \
    else LAUNCH_BOTH(G_, S_, N_, false);                                                                               \
  } while (0)
  GAT_DISPATCH_GROUP(shape, nh, LAUNCH);
#undef LAUNCH
#undef LAUNCH_BOTH
  return GAT_OK;
}

}  // namespace gat

extern "C" int gat_edge_bwd_main(const int32_t* rowptr_t, const int32_t* col_t, const int32_t* pos_t, const int32_t* row_order_t,
                                 int64_t n_long, const int32_t* eid, int64_t n_rows, const float* wh, int nh, int fp,
                                 const float* s_src, const float* s_tgt, const float* gmax, const float* z,
                                 int const_attention, float dropout_p, uint64_t seed, uint64_t offset,
                                 const float* go_padded, int go_shared, const float* grad_alpha, float* rec, float* d_wh,
                                 void* workspace, size_t workspace_bytes, gat_stream_t stream) {
  using namespace gat;
  int rc = check_common("gat_edge_bwd_main", nh, fp, workspace, workspace_bytes);
  if (rc) return rc;
  GAT_CHECK_ARG(const_attention || (s_src && s_tgt && gmax && rec), "gat_edge_bwd_main: score buffers missing");
  cudaStream_t st = (cudaStream_t)stream;
  GAT_CUDA(cudaMemsetAsync(workspace, 0, kBwdHeaderBytes, st));
  if (n_rows == 0) return GAT_OK;
  BwdMainParams P = {};
  P.rowptr_t = rowptr_t; P.col_t = col_t; P.pos_t = pos_t; P.eid = eid;
  P.sched.order = row_order_t; P.sched.counter = &((BwdHeader*)workspace)->counter_a;
  P.sched.cta_counter = &((BwdHeader*)workspace)->counter_a_cta; P.sched.n = n_rows;
  P.wh = wh; P.nh = nh; P.dp = nh * fp; P.chunks = nh * fp / 4; P.chunks_per_head = fp / 4;
  P.s_src = s_src; P.s_tgt = s_tgt; P.gmax = gmax; P.z = z; P.const_attention = const_attention;
  P.dropout_p = dropout_p; P.seed = seed; P.offset = offset; P.go = go_padded; P.grad_alpha = grad_alpha;
  P.go_shared = go_shared ? 1 : 0; P.go_ld = go_shared ? fp : nh * fp;
  P.rec = rec; P.d_wh = d_wh;
  return launch_bwd_main<false>(P, row_order_t, n_long, n_rows, st);
}

static int edge_bwd_fused_impl(bool go_bf16, const int32_t* rowptr_t, const int32_t* col_t, const int32_t* pos_t, const int32_t* row_order_t,
                                  int64_t n_long, const int32_t* eid, int64_t n_rows, const float* wh, int nh, int fp,
                                  const float* s_src, const float* s_tgt, const float* gmax, const float* z,
                                  float dropout_p, uint64_t seed, uint64_t offset,
                                  const float* go_padded, int go_shared, const float* s_sum, const float* tgt_pack,
                                  const float* a_src, const float* a_tgt,
                                  const int32_t* tie_dst, const int32_t* tie_src, const unsigned long long* tie_total,
                                  const float* corr_override, int64_t tgt_lo, int64_t tgt_hi,
                                  float* ds_src, float* ds_tgt, float* d_wh,
                                  float* const* h_push_dst, int n_push, int my_rank, int64_t rows_per_rank,
                                  void* workspace, size_t workspace_bytes, gat_stream_t stream,
                                  const float* norm_coef = nullptr, float norm_scale = 0.f) {
  using namespace gat;
  int rc = check_common("gat_edge_bwd_fused", nh, fp, workspace, workspace_bytes);
  GAT_CHECK_ARG(norm_coef == nullptr || (tgt_pack != nullptr && !go_bf16), "gat_edge_bwd_fused_norm: needs the per-target records (tgt_pack)");
  if (rc) return rc;
  GAT_CHECK_ARG(s_src && gmax && (tgt_pack || (s_tgt && z && s_sum)) && a_src && a_tgt && ds_src && ds_tgt && (d_wh || n_push > 0),
                "gat_edge_bwd_fused: buffers missing");
  GAT_CHECK_ARG(n_push == 0 || (h_push_dst && n_push >= 1 && n_push <= 8 && my_rank >= 0 && my_rank < n_push && rows_per_rank >= 1 &&
                                rows_per_rank * n_push >= n_rows),
                "gat_edge_bwd_fused: bad push configuration");
  if (n_rows == 0) return GAT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  BwdHeader* header = (BwdHeader*)workspace;
  if (corr_override == nullptr) {   // Gamma partials were left in the workspace by gat_edge_bwd_rowdot
    gamma_finalize_kernel<<<1, 1024, 0, st>>>(header, (const double*)((char*)workspace + kBwdHeaderBytes), tie_total);
    GAT_LAUNCH_CHECK();
  }
  GAT_CUDA(cudaMemsetAsync(&header->counter_a, 0, 2 * sizeof(unsigned int), st));
  BwdMainParams P = {};
  P.rowptr_t = rowptr_t; P.col_t = col_t; P.pos_t = pos_t; P.eid = eid;
  P.sched.order = row_order_t; P.sched.counter = &header->counter_a; P.sched.cta_counter = &header->counter_a_cta; P.sched.n = n_rows;
  P.wh = wh; P.nh = nh; P.dp = nh * fp; P.chunks = nh * fp / 4; P.chunks_per_head = fp / 4;
  P.s_src = s_src; P.s_tgt = s_tgt; P.gmax = gmax; P.z = z; P.const_attention = 0;
  P.dropout_p = dropout_p; P.seed = seed; P.offset = offset; P.go = go_padded; P.grad_alpha = nullptr;
  P.go_shared = go_shared ? 1 : 0; P.go_ld = go_shared ? fp : nh * fp;
  P.rec = nullptr; P.d_wh = d_wh;
  P.s_sum = s_sum; P.tpack = tgt_pack; P.a_src = a_src; P.a_tgt = a_tgt; P.tie_dst = tie_dst; P.tie_src = tie_src; P.header = header;
  P.corr_override = corr_override; P.tgt_lo = tgt_lo; P.tgt_hi = tgt_hi; P.ds_src = ds_src; P.ds_tgt = ds_tgt;
  P.push = n_push > 0; P.my_rank = my_rank; P.rows_per_rank = rows_per_rank > 0 ? rows_per_rank : 1;
  P.go_bf16 = go_bf16 ? 1 : 0;
  P.norm_coef = norm_coef; P.norm_scale = norm_scale;
  for (int q = 0; q < n_push; ++q) P.push_dst[q] = h_push_dst[q];
  return launch_bwd_main<true>(P, row_order_t, n_long, n_rows, st);
}

extern "C" int gat_edge_bwd_fused(const int32_t* rowptr_t, const int32_t* col_t, const int32_t* pos_t, const int32_t* row_order_t,
                                  int64_t n_long, const int32_t* eid, int64_t n_rows, const float* wh, int nh, int fp,
                                  const float* s_src, const float* s_tgt, const float* gmax, const float* z,
                                  float dropout_p, uint64_t seed, uint64_t offset,
                                  const float* go_padded, int go_shared, const float* s_sum, const float* tgt_pack,
                                  const float* a_src, const float* a_tgt,
                                  const int32_t* tie_dst, const int32_t* tie_src, const unsigned long long* tie_total,
                                  const float* corr_override, int64_t tgt_lo, int64_t tgt_hi,
                                  float* ds_src, float* ds_tgt, float* d_wh,
                                  float* const* h_push_dst, int n_push, int my_rank, int64_t rows_per_rank,
                                  void* workspace, size_t workspace_bytes, gat_stream_t stream) {
  return edge_bwd_fused_impl(false, rowptr_t, col_t, pos_t, row_order_t, n_long, eid, n_rows, wh, nh, fp, s_src, s_tgt, gmax, z, dropout_p, seed, offset,
                             go_padded, go_shared, s_sum, tgt_pack, a_src, a_tgt, tie_dst, tie_src, tie_total, corr_override, tgt_lo, tgt_hi,
                             ds_src, ds_tgt, d_wh, h_push_dst, n_push, my_rank, rows_per_rank, workspace, workspace_bytes, stream);
}

// gat_edge_bwd_fused for a loss that also carries c * |alpha*deg - 1|_1 of this layer (c = *norm_coef * norm_scale): see
// BwdMainParams::norm_coef; the records must come from gat_edge_bwd_rowdot_glue called with the same coefficient.
extern "C" int gat_edge_bwd_fused_norm(const int32_t* rowptr_t, const int32_t* col_t, const int32_t* pos_t, const int32_t* row_order_t,
                                  int64_t n_long, const int32_t* eid, int64_t n_rows, const float* wh, int nh, int fp,
                                  const float* s_src, const float* s_tgt, const float* gmax, const float* z,
                                  float dropout_p, uint64_t seed, uint64_t offset,
                                  const float* go_padded, int go_shared, const float* s_sum, const float* tgt_pack,
                                  const float* a_src, const float* a_tgt,
                                  const int32_t* tie_dst, const int32_t* tie_src, const unsigned long long* tie_total,
                                  const float* norm_coef, float norm_scale,
                                  float* ds_src, float* ds_tgt, float* d_wh,
                                  void* workspace, size_t workspace_bytes, gat_stream_t stream) {
  return edge_bwd_fused_impl(false, rowptr_t, col_t, pos_t, row_order_t, n_long, eid, n_rows, wh, nh, fp, s_src, s_tgt, gmax, z, dropout_p, seed, offset,
                             go_padded, go_shared, s_sum, tgt_pack, a_src, a_tgt, tie_dst, tie_src, tie_total, nullptr, 0, n_rows,
                             ds_src, ds_tgt, d_wh, nullptr, 0, 0, 0, workspace, workspace_bytes, stream, norm_coef, norm_scale);
}

// bf16 variant: `go_bf16` is a bfloat16 copy (gat_f32_to_bf16) of the (n_targets, nh*fp) upstream gradient the pass gathers.
extern "C" int gat_edge_bwd_fused_bf16(const int32_t* rowptr_t, const int32_t* col_t, const int32_t* pos_t, const int32_t* row_order_t,
                                  int64_t n_long, const int32_t* eid, int64_t n_rows, const float* wh, int nh, int fp,
                                  const float* s_src, const float* s_tgt, const float* gmax, const float* z,
                                  float dropout_p, uint64_t seed, uint64_t offset,
                                  const void* go_bf16, int go_shared, const float* s_sum, const float* tgt_pack,
                                  const float* a_src, const float* a_tgt,
                                  const int32_t* tie_dst, const int32_t* tie_src, const unsigned long long* tie_total,
                                  const float* corr_override, int64_t tgt_lo, int64_t tgt_hi,
                                  float* ds_src, float* ds_tgt, float* d_wh,
                                  float* const* h_push_dst, int n_push, int my_rank, int64_t rows_per_rank,
                                  void* workspace, size_t workspace_bytes, gat_stream_t stream) {
  return edge_bwd_fused_impl(true, rowptr_t, col_t, pos_t, row_order_t, n_long, eid, n_rows, wh, nh, fp, s_src, s_tgt, gmax, z, dropout_p, seed, offset,
                             (const float*)go_bf16, go_shared, s_sum, tgt_pack, a_src, a_tgt, tie_dst, tie_src, tie_total, corr_override, tgt_lo, tgt_hi,
                             ds_src, ds_tgt, d_wh, h_push_dst, n_push, my_rank, rows_per_rank, workspace, workspace_bytes, stream);
}

extern "C" int gat_edge_bwd_rowsum(const int32_t* rowptr, const int32_t* tpos, const int32_t* row_order, int64_t n_long, int64_t n_rows, int nh,
                                   const float* rec, const float* z, float* s_sum, float* ds_tgt,
                                   void* workspace, size_t workspace_bytes, gat_stream_t stream) {
  using namespace gat;
  (void)n_long;
  int rc = check_common("gat_edge_bwd_rowsum", nh, 4, workspace, workspace_bytes);
  if (rc) return rc;
  if (n_rows == 0) return GAT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  BwdHeader* header = (BwdHeader*)workspace;
  GAT_CUDA(cudaMemsetAsync(&header->counter_b, 0, 2 * sizeof(unsigned int), st));
  BwdRowsumParams P;
  P.rowptr = rowptr; P.tpos = tpos; P.sched.order = row_order; P.sched.counter = &header->counter_b;
  P.sched.cta_counter = &header->counter_b_cta; P.sched.n = n_rows;
  P.nh = nh; P.rec = rec; P.z = z; P.s_sum = s_sum; P.ds_tgt = ds_tgt;
  edge_bwd_rowsum_kernel<<<persistent_grid(edge_bwd_rowsum_kernel, 256, 0, (n_rows + 7) / 8), 256, 0, st>>>(P);
  GAT_LAUNCH_CHECK();
  gamma_partial_kernel<<<kGammaBlocks, 256, 0, st>>>(ds_tgt, n_rows * nh, header, (double*)((char*)workspace + kBwdHeaderBytes));
  GAT_LAUNCH_CHECK();
  return GAT_OK;
}

extern "C" int gat_tgt_pack_stride(int nh) { return nh <= 4 ? 16 : 32; }

static int edge_bwd_rowdot_impl(const float* go_padded, int go_shared, const float* out_padded, int out_is_act, float* go_out,
                                   const float* z, int64_t n_rows, int nh, int fp,
                                   float* s_sum, float* ds_tgt, const float* s_tgt, float* tgt_pack,
                                   void* workspace, size_t workspace_bytes, gat_stream_t stream,
                                   const float* skip, int64_t ld_skip, float drop_p, uint64_t drop_seed,
                                   const int32_t* rowptr = nullptr, const float* norm_t = nullptr, const float* norm_coef = nullptr,
                                   float norm_scale = 0.f) {
  using namespace gat;
  int rc = check_common("gat_edge_bwd_rowdot", nh, fp, workspace, workspace_bytes);
  GAT_CHECK_ARG((norm_t == nullptr) == (norm_coef == nullptr) && (norm_t == nullptr || (rowptr != nullptr && tgt_pack != nullptr)),
                "gat_edge_bwd_rowdot: the fused attention norm needs rowptr, norm_t, norm_coef and tgt_pack together");
  if (rc) return rc;
  const bool glue = out_is_act || skip != nullptr || drop_p > 0.f;
  GAT_CHECK_ARG(!glue || (go_out != nullptr && !go_shared), "gat_edge_bwd_rowdot: a fused output glue needs go_out and an unshared gradient");
  GAT_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "gat_edge_bwd_rowdot: output dropout %f not in [0, 1)", drop_p);
  GAT_CHECK_ARG(skip == nullptr || (ld_skip >= (int64_t)nh * fp && ld_skip % 4 == 0 && ((uintptr_t)skip & 15) == 0),
                "gat_edge_bwd_rowdot: skip rows must be 16-byte aligned with a stride of at least nh*fp floats");
  GAT_CHECK_ARG(tgt_pack == nullptr || (s_tgt != nullptr && ((uintptr_t)tgt_pack & 15) == 0), "gat_edge_bwd_rowdot: tgt_pack needs s_tgt and 16-byte alignment");
  if (n_rows == 0) return GAT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  BwdRowdotParams P;
  P.go = go_padded; P.out = out_padded; P.z = z; P.n = n_rows; P.nh = nh; P.dp = nh * fp; P.chunks = nh * fp / 4;
  P.chunks_per_head = fp / 4; P.s_sum = s_sum; P.ds_tgt = ds_tgt;
  P.go_shared = go_shared ? 1 : 0; P.go_ld = go_shared ? fp : nh * fp;
  P.out_is_act = out_is_act ? 1 : 0; P.go_out = go_out;
  P.glue = glue ? 1 : 0; P.skip = skip; P.ld_skip = ld_skip; P.drop_p = drop_p; P.drop_seed = drop_seed;
  P.rowptr = rowptr; P.norm_t = norm_t; P.norm_coef = norm_coef; P.norm_scale = norm_scale;
  P.s_tgt = s_tgt; P.tpack = tgt_pack;
  int64_t want = (n_rows + 31) / 32;
  const unsigned grid = (unsigned)(want < kNumSMs * 8 ? (want < 1 ? 1 : want) : kNumSMs * 8);
  const bool full = skip != nullptr || drop_p > 0.f;
  if (nh <= 4) { if (full) edge_bwd_rowdot_kernel<4, true><<<grid, 256, 0, st>>>(P); else edge_bwd_rowdot_kernel<4, false><<<grid, 256, 0, st>>>(P); }
  else { if (full) edge_bwd_rowdot_kernel<8, true><<<grid, 256, 0, st>>>(P); else edge_bwd_rowdot_kernel<8, false><<<grid, 256, 0, st>>>(P); }
  GAT_LAUNCH_CHECK();
  gamma_partial_kernel<<<kGammaBlocks, 256, 0, st>>>(ds_tgt, n_rows * nh, (BwdHeader*)workspace, (double*)((char*)workspace + kBwdHeaderBytes));
  GAT_LAUNCH_CHECK();
  return GAT_OK;
}

extern "C" int gat_edge_bwd_rowdot(const float* go_padded, int go_shared, const float* out_padded, int out_is_act, float* go_out,
                                   const float* z, int64_t n_rows, int nh, int fp,
                                   float* s_sum, float* ds_tgt, const float* s_tgt, float* tgt_pack,
                                   void* workspace, size_t workspace_bytes, gat_stream_t stream) {
  return edge_bwd_rowdot_impl(go_padded, go_shared, out_padded, out_is_act, go_out, z, n_rows, nh, fp, s_sum, ds_tgt, s_tgt, tgt_pack,
                              workspace, workspace_bytes, stream, nullptr, 0, 0.f, 0);
}

// gat_edge_bwd_rowdot for a forward that ran with the whole output glue (gat_edge_fwd_glue): out_padded holds y = keep * E(out + skip)
extern "C" int gat_edge_bwd_rowdot_glue(const float* go_padded, int go_shared, const float* y_padded, int out_is_act, const float* skip,
                                        int64_t ld_skip, float drop_p, uint64_t drop_seed, float* go_out,
                                        const int32_t* rowptr, const float* norm_t, const float* norm_coef, float norm_scale,
                                        const float* z, int64_t n_rows, int nh, int fp,
                                        float* s_sum, float* ds_tgt, const float* s_tgt, float* tgt_pack,
                                        void* workspace, size_t workspace_bytes, gat_stream_t stream) {
  return edge_bwd_rowdot_impl(go_padded, go_shared, y_padded, out_is_act, go_out, z, n_rows, nh, fp, s_sum, ds_tgt, s_tgt, tgt_pack,
                              workspace, workspace_bytes, stream, skip, ld_skip, drop_p, drop_seed, rowptr, norm_t, norm_coef, norm_scale);
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

extern "C" int gat_edge_bwd_finish(const int32_t* rowptr_t, const int32_t* col_t, const int32_t* row_order_t, int64_t n_long, int64_t n_rows,
                                   int nh, int fp, const float* rec, const float* s_sum, const float* a_src, const float* a_tgt,
                                   const int32_t* tie_dst, const int32_t* tie_src, const unsigned long long* tie_total,
                                   const float* corr_override, int64_t tgt_lo, int64_t tgt_hi,
                                   float* ds_src, float* ds_tgt, float* d_wh,
                                   void* workspace, size_t workspace_bytes, gat_stream_t stream) {
  using namespace gat;
  (void)n_long;
  int rc = check_common("gat_edge_bwd_finish", nh, fp, workspace, workspace_bytes);
  if (rc) return rc;
  GAT_CHECK_ARG(rec && s_sum && a_src && a_tgt && ds_src && ds_tgt && d_wh, "gat_edge_bwd_finish: buffers missing");
  if (n_rows == 0) return GAT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  BwdHeader* header = (BwdHeader*)workspace;
  if (corr_override == nullptr) {
    gamma_finalize_kernel<<<1, 1024, 0, st>>>(header, (const double*)((char*)workspace + kBwdHeaderBytes), tie_total);
    GAT_LAUNCH_CHECK();
  }
  GAT_CUDA(cudaMemsetAsync(&header->counter_a, 0, 2 * sizeof(unsigned int), st));
  BwdFinishParams P;
  P.rowptr_t = rowptr_t; P.col_t = col_t; P.sched.order = row_order_t; P.sched.counter = &header->counter_a;
  P.sched.cta_counter = &header->counter_a_cta; P.sched.n = n_rows;
  P.nh = nh; P.dp = nh * fp; P.chunks = nh * fp / 4;
  P.rec = rec; P.s_sum = s_sum; P.a_src = a_src; P.a_tgt = a_tgt;
  P.tie_dst = tie_dst; P.tie_src = tie_src; P.header = header; P.corr_override = corr_override;
  P.tgt_lo = tgt_lo; P.tgt_hi = tgt_hi; P.ds_src = ds_src; P.ds_tgt = ds_tgt; P.d_wh = d_wh;
  GroupShape shape = pick_group(P.chunks, n_rows);
  if (shape.slots < 0) {
    set_error("gat_edge_bwd_finish: row width %d floats exceeds the supported 1024", P.dp);
    return GAT_EUNSUPPORTED;
  }
#define LAUNCH(G_, S_, N_)                                                                                \
  edge_bwd_finish_kernel<G_, S_, N_><<<persistent_grid(edge_bwd_finish_kernel<G_, S_, N_>, kEdgeThreads, 0,       \
                                                       (n_rows + (kEdgeThreads / G_) - 1) / (kEdgeThreads / G_)), \
                                       kEdgeThreads, 0, st>>>(P)
  GAT_DISPATCH_GROUP(shape, nh, LAUNCH);
#undef LAUNCH
  GAT_LAUNCH_CHECK();
  return GAT_OK;
}

extern "C" int gat_slab_sum(const float* recv, int n_slabs, int64_t slab_rows, int dp, float* out, gat_stream_t stream) {
  using namespace gat;
  GAT_CHECK_ARG(recv && out && n_slabs >= 1 && slab_rows >= 0 && dp > 0 && dp % 4 == 0, "gat_slab_sum: bad arguments");
  const int64_t count4 = slab_rows * dp / 4;
  if (count4 == 0) return GAT_OK;
  const int64_t want = (count4 + 255) / 256;
  slab_sum_kernel<<<(unsigned)(want < kNumSMs * 16 ? want : kNumSMs * 16), 256, 0, (cudaStream_t)stream>>>(
      recv, n_slabs, slab_rows * dp, count4, out);
  GAT_LAUNCH_CHECK();
  return GAT_OK;
}
