// Kernel 4, head-mean layers with four heads (products' output layer: NH = 4, F = 47 -> padded 48 floats per head): the fused
// source-major backward pass with a lane mapping built for the SHARED upstream-gradient row.
//
// A head-mean layer hands every head the same (Fp)-wide gradient row dOut[d]/NH (gat_layer.py:132), so an edge gathers 192 B
// while the source's own row Wh[s] is four times as wide.  The general kernel (edge_bwd_main_kernel, lane = chunk of the WIDE
// row) then spends ~92 warp instructions per edge on 48 chunks over 32 lanes x 2 slots plus a shared-memory transpose per
// batch, and ran at a third of the DRAM rate (ncu r01c: 33 % DRAM, 69 % issue slots).  Here
//
//     lane = (edge slot j in {0,1}) x (head h in 0..3) x (quarter q in 0..3)
//
// so a warp step handles TWO edges with all 32 lanes busy; a lane owns the in-head chunks c = q, q+4, q+8, ... (CPL of them) of
// head h of Wh[s] (registers, loaded once per source row) and of the accumulated dWh[s]:
//     acc[h][c] += w[e,h] * go[d_e][c]                 CPL float4 FMAs
//     <go[d_e], Wh[s,h,:]> = sum over the 4 quarter lanes (two shuffles) of  sum_c whr[c].go[c]
//     g[e,h] = 0.01 * (w*<.,.> - alpha*S[d])           summed per lane, combined per row with one shuffle
// The gathered rows of 16 edges at a time are staged in shared memory with cp.async (no registers held), double buffered, and
// issued as soon as the target ids are known, so the gather overlaps the per-edge softmax recomputation (which needs the
// target's {s_tgt | Z | S} record, a second DRAM round trip).  Row epilogue: the j = 0 half adds ds_src*A_src, the j = 1 half
// ds_tgt*A_tgt, then the halves are summed with one shuffle per register -- every sum in a fixed order, no atomics.
// Long rows (> kLongRow edges) keep the cooperative CTA-per-row launch of the general kernel.
#pragma once

namespace gat {

constexpr int kHmStage = 16;         // edges per staging buffer
constexpr int kHmWarps = kEdgeThreads / 32;

template <int CPL>
struct HmShape {
  static constexpr int SROW = 4 * CPL;                                  // float4 per staged row
  static constexpr int kStageF4 = 2 * kHmStage * SROW;                  // per warp, both buffers
  // per warp: stage | w[32] (float4) | aS[32] (float4) | d[32] (int)
  static constexpr size_t kWarpBytes = (size_t)kStageF4 * 16 + 32 * 16 + 32 * 16 + 32 * 4;
  static constexpr size_t kSmem = kWarpBytes * kHmWarps;
};

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// EXACT: chunks_per_head == 4*CPL (products: 12), so every lane's CPL chunks exist and the per-chunk predicates fold away.
template <int CPL, bool EXACT>
__global__ void __launch_bounds__(kEdgeThreads, 3)
edge_bwd_hm4_kernel(const BwdMainParams P) {
  constexpr int SROW = HmShape<CPL>::SROW;
  extern __shared__ __align__(16) unsigned char hm_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned char* wbase = hm_smem + (size_t)warp * HmShape<CPL>::kWarpBytes;
  float4* stage = reinterpret_cast<float4*>(wbase);                                   // [2][kHmStage][SROW]
  float4* sh_w = stage + HmShape<CPL>::kStageF4;                                      // [32] m*alpha per head
  float4* sh_s = sh_w + 32;                                                           // [32] alpha*S[dst] per head
  int* sh_d = reinterpret_cast<int*>(sh_s + 32);                                      // [32] target ids
  const int j = lane >> 4, h = (lane >> 2) & 3, q = lane & 3;
  const int cph = EXACT ? 4 * CPL : P.chunks_per_head;
  const float gmax = __ldg(P.gmax);
  const float corr = P.corr_override ? __ldg(P.corr_override) : P.header->corr;
  bool ok[CPL];
  int coff[CPL];            // float offset of this lane's chunk i inside a full (NH*Fp) row
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    const int c = q + 4 * i;
    ok[i] = EXACT || c < cph;
    coff[i] = (h * cph + (ok[i] ? c : 0)) * 4;
  }
  // staging: lane -> (edge k, chunk jj) advanced without a division per chunk
  const int q32 = 32 / cph, r32 = 32 - q32 * cph;
  const int k0s = lane / cph, j0s = lane - k0s * cph;

  int flip = 0;             // which of the warp's two push buffers is next
  int64_t base_rows;
  while (grab_rows<32>(P.sched, lane, base_rows)) {
    int pr, ps, pe;
    prefetch_rows<32>(P.sched, P.rowptr_t, base_rows, lane, pr, ps, pe, P.sched_rot);
#pragma unroll 1
    for (int kk = 0; kk < kGrabIters<32>; ++kk) {
      int64_t row;
      int start, end;
      if (!prefetched_row<32>(P.sched, kk, lane, pr, ps, pe, row, start, end)) continue;
      // ---- row prologue ----
      float4 whr[CPL], acc[CPL];
#pragma unroll
      for (int i = 0; i < CPL; ++i) {
        acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        whr[i] = ok[i] ? ldg4(P.wh + row * P.dp + coff[i]) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const float4 ss = ldg4(P.s_src + row * 4);
      const bool own_tgt = row >= P.tgt_lo && row < P.tgt_hi;
      const int64_t trow = row - P.tgt_lo;
      int p_ts = 0, p_td = 0;
      float p_t = 0.f;
      if (lane < 4) {
        if (P.tie_src) p_ts = __ldg(P.tie_src + row * 4 + lane);
        if (own_tgt) {
          if (P.tie_dst) p_td = __ldg(P.tie_dst + trow * 4 + lane);
          p_t = P.ds_tgt[trow * 4 + lane];
        }
      }
      float gsum = 0.f;
      for (int base = start; base < end; base += 32) {
        const int cnt = min(32, end - base);
        const int e = base + lane;
        const bool valid = lane < cnt;
        // target id; lanes past the end repeat the last valid edge (weight 0), so the step loop needs no predicates and reads
        // nothing an existing edge does not read
        int d = __ldg(P.col_t + (valid ? e : base + cnt - 1));
        __syncwarp();                         // the previous batch has been consumed by every lane
        sh_d[lane] = d;
        __syncwarp();
        // ---- issue the gathers of both halves (edges 0..15 -> buffer 0, 16..31 -> buffer 1) ----
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int ne = min(kHmStage, ((cnt + 1) & ~1) - hf * kHmStage);   // edges of this half, rounded up to a pair
          if (ne > 0) {
            const int total = ne * cph;
            int k = k0s, jj = j0s;
            for (int i = lane; i < total; i += 32) {
              cp_async16(stage + (hf * kHmStage + k) * SROW + jj, P.go + (int64_t)sh_d[hf * kHmStage + k] * P.go_ld + jj * 4);
              k += q32; jj += r32;
              if (jj >= cph) { jj -= cph; ++k; }
            }
          }
          cp_async_commit();
        }
        // ---- per-edge scalars (lane = edge): alpha from the target's record, dropout mask ----
        {
          const float* pk = P.tpack + (int64_t)d * 16;
          const float4 t4 = ldg4(pk), z4 = ldg4(pk + 4), s4 = ldg4(pk + 8);
          float4 al, w;
          al.x = attn_exp(ss.x + t4.x, gmax) / (z4.x + kSoftmaxEps);
          al.y = attn_exp(ss.y + t4.y, gmax) / (z4.y + kSoftmaxEps);
          al.z = attn_exp(ss.z + t4.z, gmax) / (z4.z + kSoftmaxEps);
          al.w = attn_exp(ss.w + t4.w, gmax) / (z4.w + kSoftmaxEps);
          w = al;
          float4 s4e = s4;
          if (P.norm_coef != nullptr) {   // fused attention-norm regulariser: dL/dalpha folded into the S term (BwdMainParams::norm_coef)
            const float cn = __ldg(P.norm_coef) * P.norm_scale, deg = __ldg(pk + 12);
            const float ux = al.x * deg - 1.0f, uy = al.y * deg - 1.0f, uz = al.z * deg - 1.0f, uw = al.w * deg - 1.0f;
            s4e.x -= ux > 0.f ? cn * deg : (ux < 0.f ? -cn * deg : 0.f);
            s4e.y -= uy > 0.f ? cn * deg : (uy < 0.f ? -cn * deg : 0.f);
            s4e.z -= uz > 0.f ? cn * deg : (uz < 0.f ? -cn * deg : 0.f);
            s4e.w -= uw > 0.f ? cn * deg : (uw < 0.f ? -cn * deg : 0.f);
          }
          if (P.dropout_p > 0.f && valid) {
            float msk[4];
            const int edge_id = __ldg(P.eid + __ldg(P.pos_t + e));
            dropout_scales<4>(P.seed, P.offset, (uint32_t)edge_id, 4, P.dropout_p, msk);
            w.x *= msk[0]; w.y *= msk[1]; w.z *= msk[2]; w.w *= msk[3];
          }
          if (!valid) { w = make_float4(0.f, 0.f, 0.f, 0.f); al = w; }
          sh_w[lane] = w;
          sh_s[lane] = make_float4(al.x * s4e.x, al.y * s4e.y, al.z * s4e.z, al.w * s4e.w);
        }
        // ---- the two halves ----
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          if (hf == 0) cp_async_wait_group<1>(); else cp_async_wait_group<0>();
          __syncwarp();                       // staged rows and the per-edge scalars of every lane are visible
          const int ne = min(kHmStage, ((cnt + 1) & ~1) - hf * kHmStage);
          const float4* srow = stage + (hf * kHmStage + j) * SROW + q;
          const float* wp = reinterpret_cast<const float*>(sh_w + hf * kHmStage + j) + h;
          const float* sp = reinterpret_cast<const float*>(sh_s + hf * kHmStage + j) + h;
#pragma unroll 2
          for (int t = 0; t < ne; t += 2) {
            float4 v[CPL];
#pragma unroll
            for (int i = 0; i < CPL; ++i) v[i] = ok[i] ? srow[t * SROW + 4 * i] : make_float4(0.f, 0.f, 0.f, 0.f);
            const float w = wp[t * 4], as = sp[t * 4];
            float dd = 0.f;
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
              acc[i].x = fmaf(w, v[i].x, acc[i].x);
              acc[i].y = fmaf(w, v[i].y, acc[i].y);
              acc[i].z = fmaf(w, v[i].z, acc[i].z);
              acc[i].w = fmaf(w, v[i].w, acc[i].w);
              dd = fmaf(whr[i].x, v[i].x, dd);
              dd = fmaf(whr[i].y, v[i].y, dd);
              dd = fmaf(whr[i].z, v[i].z, dd);
              dd = fmaf(whr[i].w, v[i].w, dd);
            }
            dd += __shfl_xor_sync(0xffffffffu, dd, 1);
            dd += __shfl_xor_sync(0xffffffffu, dd, 2);
            // g = 0.01*alpha*(d_alpha - S[dst]) = 0.01*((m*alpha)*<dOut,Wh> - alpha*S[dst])   
            gsum = fmaf(kLeakySlope, fmaf(w, dd, -as), gsum);
          }
        }
      }
      // ---- row epilogue ----
      gsum += __shfl_xor_sync(0xffffffffu, gsum, 16);       // the two edge slots; lanes of head h now hold sum_e g[e,h]
      float coef[4];
#pragma unroll
      for (int hh = 0; hh < 4; ++hh) {
        const float g = __shfl_sync(0xffffffffu, gsum, hh * 4);
        const int ts = __shfl_sync(0xffffffffu, p_ts, hh);
        const int td = __shfl_sync(0xffffffffu, p_td, hh);
        const float t = __shfl_sync(0xffffffffu, p_t, hh);
        // ds_src = sum g - |T_src|*Gamma/|T|, ds_tgt -= |T_dst|*Gamma/|T| (gradient through max())
        const float dss = ts ? g - (float)ts * corr : g;
        const float dst_ = own_tgt ? (td ? t - (float)td * corr : t) : 0.f;
        if (lane == hh) {
          P.ds_src[row * 4 + hh] = dss;
          if (own_tgt) P.ds_tgt[trow * 4 + hh] = dst_;
        }
        coef[hh] = j ? dst_ : dss;
      }
      // dWh += ds_src*A_src (edge-slot half 0) + ds_tgt*A_tgt (half 1), then the halves are summed
      const float* amat = j ? P.a_tgt : P.a_src;
#pragma unroll
      for (int i = 0; i < CPL; ++i) {
        if (ok[i]) {
#pragma unroll
          for (int hh = 0; hh < 4; ++hh) {
            const float4 a4 = ldg4(amat + (int64_t)hh * P.dp + coff[i]);
            acc[i].x = fmaf(coef[hh], a4.x, acc[i].x);
            acc[i].y = fmaf(coef[hh], a4.y, acc[i].y);
            acc[i].z = fmaf(coef[hh], a4.z, acc[i].z);
            acc[i].w = fmaf(coef[hh], a4.w, acc[i].w);
          }
        }
      }
      float* const drow = dwh_row_ptr(P, row);
      // PUSH (partitioned graphs): the finished row is staged in one of the warp's two row buffers and leaves as ONE bulk copy
      // into its owner's receive slab over NVLink (see BwdMainParams::rowbuf_offset_floats); otherwise straight to d_wh
      float* dst = drow;
      if (P.rowbuf_offset_floats >= 0) {
        dst = reinterpret_cast<float*>(hm_smem) + P.rowbuf_offset_floats + (warp * 2 + flip) * P.dp;
        if (lane == 0) bulk_wait_read_keep1();          // the copy issued from this buffer two rows ago has read it
        __syncwarp();
      }
#pragma unroll
      for (int i = 0; i < CPL; ++i) {
        acc[i].x += __shfl_xor_sync(0xffffffffu, acc[i].x, 16);
        acc[i].y += __shfl_xor_sync(0xffffffffu, acc[i].y, 16);
        acc[i].z += __shfl_xor_sync(0xffffffffu, acc[i].z, 16);
        acc[i].w += __shfl_xor_sync(0xffffffffu, acc[i].w, 16);
        // chunk i leaves from half (i & 1): both halves hold the same sum (x + y is commutative bit for bit)
        if (ok[i] && j == (i & 1)) *reinterpret_cast<float4*>(dst + coff[i]) = acc[i];
      }
      if (P.rowbuf_offset_floats >= 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the bulk copy
        __syncwarp();
        if (lane == 0) bulk_store(drow, dst, (unsigned)(P.dp * sizeof(float)));
        flip ^= 1;
      }
    }
  }
  if (P.rowbuf_offset_floats >= 0 && lane == 0) bulk_wait_all();   // every pushed row has left before the grid may complete
  pdl_wait_for_primary();     // no-op unless launched behind the cooperative long-row kernel
}

}  // namespace gat
