// Kernel 2: tcgen05 / TMA projection GEMM with fp32-grade accuracy ("3xTF32").
//
//   NT:  C[M,N] = A[M,K] * B[N,K]^T    A, B row-major fp32, both K-major   (gat_gemm ta=0, tb=1: Wh = x W^T, dX = dWh W)
//   TN:  C[M,N] = A[K,M]^T * B[K,N]    A, B row-major fp32, both MN-major  (gat_gemm ta=1, tb=0: dW = dWh^T x), split-K
//
// tcgen05 has no fp32 MMA kind, and plain TF32 (~5e-4) cannot meet the 1e-5 parity bar, so
// every operand tile gets a companion lo = v - tf32(v) tile IN SHARED MEMORY, and three MMAs accumulate
// hi*lo + lo*hi + hi*hi into fp32 TMEM accumulators.  The global operands are read exactly once, as
// fp32, by TMA (no pre-split pass, no extra HBM traffic).
//
// CTA = one 128 x BN output tile, 192 threads, warp-specialised:
//   warp 0      TMA producer: cp.async.bulk.tensor (SWIZZLE_128B boxes of 32 fp32 = 128 B per row) -> smem stage
//   warp 1      TMEM allocator + MMA issuer: one elected lane issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8)
//   warps 2..5  splitters (hi/lo in place, position preserving, so the swizzle is irrelevant), then the epilogue:
//               tcgen05.ld 32x32b from TMEM -> registers -> 32-byte row segments to global
// mbarrier pipeline per stage: full (TMA -> splitters), ready (splitters -> MMA), empty (tcgen05.commit -> TMA).
// Every wait is bounded and traps instead of hanging.
#include "tc_common.cuh"

namespace gat {

namespace tc {

// TN products (dW, K = number of nodes): the split-K plan caps one accumulator at 2048 rows and sums the partials in fp64, so
// the truncation split of the NT path (hi = the raw word, only lo is written) is accurate enough here too (measured in
// tests/test_gpu_parity.py::test_gemm_tcgen05_3xtf32) and saves one of the three shared-memory writes per element.
constexpr bool kRoundSplitTN = false;

template <int BN, bool MN>
struct Smem {
  static constexpr int kABytes = BM * BK * 4;     // 8 KB
  static constexpr int kBBytes = BN * BK * 4;
  static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;   // raw/hi + lo for both operands
  static constexpr int kCtasPerSm = (!MN && BN <= 128) ? 2 : 1;
  static constexpr int kStages = MN ? (BN >= 256 ? 4 : (BN >= 128 ? 6 : 8)) : (BN >= 256 ? 4 : (BN >= 128 ? 3 : 4));
  static constexpr int kTotal = kStages * kStageBytes + 1024 /*alignment*/ + 256 /*barriers*/;
  static_assert(kTotal * kCtasPerSm <= 227 * 1024, "shared memory budget");
};

// MN = false: NT product, one CTA per output tile, whole K.   MN = true: TN product, blockIdx.z = K split, the CTA
// writes its partial tile to C + blockIdx.z * split_stride (reduced in a fixed order by splitk_reduce_kernel).
template <int BN, bool MN>
__global__ void __launch_bounds__(kThreads, (Smem<BN, MN>::kCtasPerSm))
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CStoreMaps cmaps, float* __restrict__ C, int64_t ldc, int64_t M, int64_t N, int64_t K, int kb_per_split, int64_t split_stride,
               uint32_t mn_lbo, uint32_t mn_sbo,
               // fused score epilogue (NT only, one N tile): s_src = C A_src^T, s_tgt = C A_tgt^T in fp64, or nullptr
               const float* __restrict__ a_src, const float* __restrict__ a_tgt, int nh,
               float* __restrict__ s_src, float* __restrict__ s_tgt,
               // fused glue: ELU on the A / B operand tiles; output multiplied by ELU'(mul_src[row, col]) (NT only)
               int act_a, int act_b, const float* __restrict__ mul_src, int64_t mul_ld) {
  using S = Smem<BN, MN>;
  constexpr int kStages = S::kStages;
  // two accumulators: [0, BN) leading term hi*hi, [BN, 2BN) cross terms hi*lo + lo*hi.  The tensor core rounds its
  // fp32 accumulator toward zero at every step, a one-sided error proportional to the accumulator's magnitude; keeping
  // the (2^-12 smaller) cross terms apart cuts the number of roundings the large accumulator sees by 3x.
  constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;   // BN is a power of two here
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B atoms need 1024 B alignment
  uint64_t* bars = (uint64_t*)(smem + kStages * S::kStageBytes);
  uint64_t* full = bars;                    // TMA landed
  uint64_t* ready = bars + kStages;         // hi/lo split done
  uint64_t* empty = bars + 2 * kStages;     // MMAs that read the stage have completed
  uint64_t* accum = bars + 3 * kStages;     // all MMAs of the tile have completed
  uint32_t* tmem_ptr = (uint32_t*)(bars + 3 * kStages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // NT: 1-D grid, the N tiles of one M tile are adjacent (they share the A tile: the second read hits L2)
  const int n_tiles = (int)((N + BN - 1) / BN);
  const int64_t m0 = (MN ? (int64_t)blockIdx.x : (int64_t)(blockIdx.x / n_tiles)) * BM;
  const int64_t n0 = (MN ? (int64_t)blockIdx.y : (int64_t)(blockIdx.x % n_tiles)) * BN;
  const int total_kb = (int)((K + BK - 1) / BK);
  const int kb_begin = MN ? (int)blockIdx.z * kb_per_split : 0;
  const int num_kb = MN ? max(0, min(kb_per_split, total_kb - kb_begin)) : total_kb;
  C += (int64_t)blockIdx.z * split_stride;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&ready[s], kSplitThreads); mbar_init(&empty[s], 1); }
    mbar_init(accum, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        if (kb >= kStages) mbar_wait(&empty[s], ((kb / kStages) - 1) & 1);
        uint8_t* st = smem + s * S::kStageBytes;
        mbar_expect_tx(&full[s], S::kABytes + S::kBBytes);
        if (!MN) {
          tma_load_2d(st, &map_a, &full[s], kb * BK, (int)m0);
          tma_load_2d(st + 2 * S::kABytes, &map_b, &full[s], kb * BK, (int)n0);
        } else {
          const int krow = (kb_begin + kb) * BK;
#pragma unroll
          for (int j = 0; j < BM / 32; ++j) tma_load_2d(st + j * (BK * 128), &map_a, &full[s], (int)m0 + 32 * j, krow);
#pragma unroll
          for (int j = 0; j < BN / 32; ++j)
            tma_load_2d(st + 2 * S::kABytes + j * (BK * 128), &map_b, &full[s], (int)n0 + 32 * j, krow);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(BM, BN, MN);
      const uint32_t tmem_main = tmem_base, tmem_cross = tmem_base + BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        mbar_wait(&ready[s], (kb / kStages) & 1);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(smem + s * S::kStageBytes), a_lo = a_hi + S::kABytes;
        const uint32_t b_hi = a_hi + 2 * S::kABytes, b_lo = b_hi + S::kBBytes;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          const uint32_t first = (kb | k) != 0;
          if (!MN) {
            const uint32_t off = k * UMMA_K * 4;   // bytes along the 64-byte swizzled row
            umma_tf32(tmem_cross, make_desc_k_sw64(a_hi + off), make_desc_k_sw64(b_lo + off), idesc, first);
            umma_tf32(tmem_cross, make_desc_k_sw64(a_lo + off), make_desc_k_sw64(b_hi + off), idesc, 1);
            umma_tf32(tmem_main, make_desc_k_sw64(a_hi + off), make_desc_k_sw64(b_hi + off), idesc, first);
          } else {
            const uint32_t off = k * 1024;         // one 8-row swizzle atom per UMMA_K
            umma_tf32(tmem_cross, make_desc_mn_sw128(a_hi + off, mn_lbo, mn_sbo), make_desc_mn_sw128(b_lo + off, mn_lbo, mn_sbo), idesc, first);
            umma_tf32(tmem_cross, make_desc_mn_sw128(a_lo + off, mn_lbo, mn_sbo), make_desc_mn_sw128(b_hi + off, mn_lbo, mn_sbo), idesc, 1);
            umma_tf32(tmem_main, make_desc_mn_sw128(a_hi + off, mn_lbo, mn_sbo), make_desc_mn_sw128(b_hi + off, mn_lbo, mn_sbo), idesc, first);
          }
        }
        umma_commit(&empty[s]);                  // implies tcgen05.fence::before_thread_sync
      }
      if (num_kb > 0) umma_commit(accum);
    }
  } else {
    // ===== splitters (warps 2..5), then epilogue =====
    const int t = threadIdx.x - 64;              // 0..127
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % kStages;
      mbar_wait(&full[s], (kb / kStages) & 1);
      float4* a_hi = (float4*)(smem + s * S::kStageBytes);
      float4* a_lo = (float4*)(smem + s * S::kStageBytes + S::kABytes);
      float4* b_hi = (float4*)(smem + s * S::kStageBytes + 2 * S::kABytes);
      float4* b_lo = (float4*)(smem + s * S::kStageBytes + 2 * S::kABytes + S::kBBytes);
#pragma unroll 4
      for (int i = t; i < S::kABytes / 16; i += kSplitThreads) {
        float4 v = a_hi[i];
        if (MN && kRoundSplitTN) {   // round-to-nearest hi, written back (4x smaller representation error, one more shared-memory write per element)
          if (act_a) v = elu4(v);
          const float4 h = make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
          a_hi[i] = h;
          a_lo[i] = make_float4(tf32_round(v.x - h.x), tf32_round(v.y - h.y), tf32_round(v.z - h.z), tf32_round(v.w - h.w));
        } else {
          if (act_a) { v = elu4(v); a_hi[i] = v; }
          a_lo[i] = lo4(v);
        }
      }
#pragma unroll 4
      for (int i = t; i < S::kBBytes / 16; i += kSplitThreads) {
        float4 v = b_hi[i];
        if (MN && kRoundSplitTN) {
          if (act_b) v = elu4(v);
          const float4 h = make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
          b_hi[i] = h;
          b_lo[i] = make_float4(tf32_round(v.x - h.x), tf32_round(v.y - h.y), tf32_round(v.z - h.z), tf32_round(v.w - h.w));
        } else {
          if (act_b) { v = elu4(v); b_hi[i] = v; }
          b_lo[i] = lo4(v);
        }
      }
      fence_proxy_async();                       // generic-proxy writes -> visible to the tensor core (async proxy)
      mbar_arrive(&ready[s]);
    }
    // epilogue: warp w may touch TMEM lanes 32*(w%4) .. +31
    if (num_kb > 0) mbar_wait(accum, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int64_t row = m0 + q * 32 + lane;
    float* crow = C + row * ldc + n0;
    // Fused score terms: the attention halves (2*nh x N floats) are staged in the now idle pipeline memory; every
    // epilogue thread owns one full row of Wh (this CTA covers all N columns), so s = Wh A^T needs no reduction.
    // Fused score terms s = Wh A^T in fp64 (gat_layer.py:76-82).  Vector FP64 runs at ~1/16 of the FP32 rate on this
    // part (measured: 2.3 T DFMA/s), so the contraction goes through the FP64 tensor cores: mma.sync m8n8k4.f64,
    // M = 8 rows, N = 8 score columns, K = 4.  The attention halves are converted to fp64 once per CTA into the now
    // idle pipeline memory; each warp transposes its 32 rows x 8 columns of Wh through a private shared-memory
    // tile into the A-fragment layout.
    const bool fuse = (!MN) && a_src != nullptr;
    // NT: the tile is staged in the (now idle) pipeline memory as BN/32 boxes of 128 rows x 128 B in the SWIZZLE_128B
    // pattern the C tensor maps expect (16-byte chunk j of row r sits at chunk j ^ (r % 8)), then one thread writes
    // it with TMA to every destination.  TN (split-K partials) keeps plain 32-byte row-segment stores.
    constexpr int kBoxes = BN / 32;
    constexpr int kBoxBytes = BM * 128;
    uint8_t* stage_c = smem;
    constexpr int ASTR = BN + 4;                     // row stride (doubles) of the staged attention matrix
    constexpr int TSTR = 36;                         // row stride (doubles) of the per-warp transpose tile
    double* a_s = reinterpret_cast<double*>(smem + (MN ? 0 : kBoxes * kBoxBytes));   // [16][ASTR], rows >= 2*nh are zero
    double* tile = a_s + 16 * ASTR + (warp - 2) * (8 * TSTR);
    static_assert(MN || kBoxes * kBoxBytes + (16 * ASTR + 4 * 8 * TSTR) * 8 <= S::kStages * S::kStageBytes,
                  "the epilogue's staging buffers alias the pipeline memory and must fit inside it");
    const int nj = 2 * nh, nbk = nj > 8 ? 2 : 1;
    const int rl = q * 32 + lane;                    // row inside the tile
    double sacc[4][2][2];
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int nb = 0; nb < 2; ++nb) sacc[g][nb][0] = sacc[g][nb][1] = 0.0;
    if (fuse) {
      for (int i = t; i < 16 * BN; i += kSplitThreads) {
        const int j = i / BN, col = i - j * BN;
        float v = 0.f;
        if (j < nj && n0 + col < N) v = (j < nh) ? __ldg(a_src + (int64_t)j * N + n0 + col) : __ldg(a_tgt + (int64_t)(j - nh) * N + n0 + col);
        a_s[j * ASTR + col] = (double)v;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kSplitThreads) : "memory");
    }
#pragma unroll 1
    for (int c = 0; c < BN; c += 8) {
      uint32_t r[8], x[8];
      float v[8];
      if (num_kb > 0) {
        tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
        tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(BN + c), x);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]) + __uint_as_float(x[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
      }
      if (fuse) {
#pragma unroll
        for (int k = 0; k < 8; ++k) tile[k * TSTR + lane] = (double)v[k];      // T[col][row], conflict free
        __syncwarp();
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          double bf[2];
#pragma unroll
          for (int nb = 0; nb < 2; ++nb)
            bf[nb] = nb < nbk ? a_s[(nb * 8 + (lane >> 2)) * ASTR + c + ks * 4 + (lane & 3)] : 0.0;   // B[k][n] = A[j=n][col=k]
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const double af = tile[(ks * 4 + (lane & 3)) * TSTR + g * 8 + (lane >> 2)];              // A[row][k]
#pragma unroll
            for (int nb = 0; nb < 2; ++nb) {
              if (nb < nbk)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                             : "+d"(sacc[g][nb][0]), "+d"(sacc[g][nb][1]) : "d"(af), "d"(bf[nb]));
            }
          }
        }
        __syncwarp();
      }
      if (!MN && mul_src != nullptr && row < M) {
        const float* mp = mul_src + row * mul_ld + n0 + c;
        if (n0 + c + 8 <= N && (mul_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(mul_src) & 15) == 0) {
          const float4 m0v = __ldg(reinterpret_cast<const float4*>(mp)), m1v = __ldg(reinterpret_cast<const float4*>(mp) + 1);
          v[0] *= elu_grad1(m0v.x); v[1] *= elu_grad1(m0v.y); v[2] *= elu_grad1(m0v.z); v[3] *= elu_grad1(m0v.w);
          v[4] *= elu_grad1(m1v.x); v[5] *= elu_grad1(m1v.y); v[6] *= elu_grad1(m1v.z); v[7] *= elu_grad1(m1v.w);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (n0 + c + j < N) v[j] *= elu_grad1(__ldg(mp + j));
        }
      }
      if (!MN) {
        uint8_t* bx = stage_c + (c >> 5) * kBoxBytes + rl * 128;
        const int j0 = (c & 31) >> 2, sw = rl & 7;
        *reinterpret_cast<float4*>(bx + (((j0) ^ sw) << 4)) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(bx + (((j0 + 1) ^ sw) << 4)) = make_float4(v[4], v[5], v[6], v[7]);
      } else if (row < M) {
        if (n0 + c + 8 <= N) {
          *reinterpret_cast<float4*>(crow + c) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(crow + c + 4) = make_float4(v[4], v[5], v[6], v[7]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (n0 + c + j < N) crow[c + j] = v[j];
        }
      }
    }
    if (!MN) {
      fence_proxy_async();                           // generic-proxy writes of the staged tile -> visible to TMA
      asm volatile("bar.sync 1, %0;" ::"n"(kSplitThreads) : "memory");
      if (t == 0) {
        for (int d = 0; d < cmaps.count; ++d) {
#pragma unroll 1
          for (int b = 0; b < kBoxes; ++b)
            if (n0 + 32 * b < N) tma_store_2d(&cmaps.maps[d], stage_c + b * kBoxBytes, (int)n0 + 32 * b, cmaps.row_offset + (int)m0);
        }
        tma_store_commit_and_wait();
      }
    }
    if (fuse) {
      // C fragment: lane holds rows g*8 + lane/4, score columns nb*8 + 2*(lane%4) + {0,1}
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int64_t srow = m0 + q * 32 + g * 8 + (lane >> 2);
        if (srow < M) {
#pragma unroll
          for (int nb = 0; nb < 2; ++nb) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int j = nb * 8 + 2 * (lane & 3) + e;
              if (nb < nbk && j < nj) {
                float* sp = j < nh ? s_src + srow * nh + j : s_tgt + srow * nh + (j - nh);
                // two N tiles: each adds its half of the dot product to the zero-initialised score (at most two addends per
                // element, and a + b == b + a, so the result does not depend on which CTA arrives first)
                if (n_tiles > 1) atomicAdd(sp, (float)sacc[g][nb][e]); else *sp = (float)sacc[g][nb][e];
              }
            }
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
  }
}

struct TnPlan { int splits; int kb_per_split; };
static TnPlan tn_plan(int64_t m, int64_t n, int64_t k, int bn) {
  const int64_t tiles = ((m + BM - 1) / BM) * ((n + bn - 1) / bn);
  const int total_kb = (int)((k + BK - 1) / BK);
  int64_t want = (kNumSMs + tiles - 1) / tiles;            // at least one wave of CTAs
  if (want > total_kb) want = total_kb;
  if (want < 1) want = 1;
  TnPlan p;
  p.kb_per_split = (int)((total_kb + want - 1) / want);
  // The tensor core rounds its fp32 accumulator toward zero at every MMA, so the error of one accumulator grows
  // linearly with the number of k-steps (measured: 7e-5 after 4136 steps, ~1e-6 after 128).  Cap the steps per
  // split; the partials are then summed in fp64 by splitk_reduce_kernel.
  constexpr int kMaxKbPerSplit = 2048 / BK;
  if (p.kb_per_split > kMaxKbPerSplit) p.kb_per_split = kMaxKbPerSplit;
  p.splits = (total_kb + p.kb_per_split - 1) / p.kb_per_split;
  if (p.splits < 1) p.splits = 1;
  return p;
}

void splitk_reduce_launch(const float* partial, int splits, int64_t m, int64_t n, float* c, int64_t ldc, cudaStream_t st);

struct ScoreFuse { const float* a_src; const float* a_tgt; int nh; float* s_src; float* s_tgt; };

// Extra destinations of the output (fused projection -> all-gather): `count` base pointers of (row_offset + m, n)
// matrices with leading dimension ldc; the tile rows are written at row_offset.  count == 0: only `c`.
struct Dests { float* const* ptrs; int count; int64_t row_offset; };
struct Glue { int act_a; int act_b; const float* mul_src; int64_t mul_ld; };

template <int BN, bool MN>
static int launch(int64_t m, int64_t n, int64_t k, const float* a, int64_t lda, const float* b, int64_t ldb, float* c,
                  int64_t ldc, void* workspace, size_t workspace_bytes, cudaStream_t st, ScoreFuse f = ScoreFuse{nullptr, nullptr, 0, nullptr, nullptr},
                  Dests dests = Dests{nullptr, 0, 0}, Glue g = Glue{0, 0, nullptr, 0}) {
  CUtensorMap map_a, map_b;
  CStoreMaps cm;          // only .count/.row_offset and the first `count` maps are meaningful; copied by value at launch
  cm.count = 0; cm.row_offset = 0;
  int rc;
  if (!MN) {
    rc = make_map(&map_a, a, m, k, lda, BM, kMapK64);
    if (!rc) rc = make_map(&map_b, b, n, k, ldb, BN, kMapK64);
    if (dests.count > kMaxDests || dests.row_offset + m >= ((int64_t)1 << 31)) { set_error("gat_gemm: too many destinations"); return GAT_EINVAL; }
    if (dests.count == 0) {
      if (!rc) rc = make_map(&cm.maps[0], c, m, n, ldc, BM, kMapC128);
      cm.count = 1;
    } else {
      for (int d = 0; d < dests.count && !rc; ++d) rc = make_map(&cm.maps[d], dests.ptrs[d], dests.row_offset + m, n, ldc, BM, kMapC128);
      cm.count = dests.count; cm.row_offset = (int)dests.row_offset;
    }
  } else {   // stored (K, M) and (K, N): boxes of 32 MN-elements x BK k-rows
    rc = make_map(&map_a, a, k, m, lda, BK, kMapMN);
    if (!rc) rc = make_map(&map_b, b, k, n, ldb, BK, kMapMN);
  }
  if (rc) return rc;
  static bool attr_set_dev[kMaxDevices] = {false};     // per device (and per template instantiation)
  bool& attr_set = attr_set_dev[cur_device()];
  if (!attr_set) {
    GAT_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<BN, MN>::kTotal));
    attr_set = true;
  }
  if (!MN) {
    const int64_t tiles = ((m + BM - 1) / BM) * ((n + BN - 1) / BN);
    if (tiles >= ((int64_t)1 << 31)) { set_error("gat_gemm: too many tiles"); return GAT_EINVAL; }
    if (f.a_src != nullptr && n > BN) {   // partial dot products of the N tiles are added into the scores
      if (n > 2 * BN) { set_error("gat_project_fwd: fused scores support at most two N tiles"); return GAT_EUNSUPPORTED; }
      GAT_CUDA(cudaMemsetAsync(f.s_src, 0, (size_t)m * f.nh * sizeof(float), st));
      GAT_CUDA(cudaMemsetAsync(f.s_tgt, 0, (size_t)m * f.nh * sizeof(float), st));
    }
    dim3 grid((unsigned)tiles, 1, 1);
    gemm_tc_kernel<BN, MN><<<grid, kThreads, Smem<BN, MN>::kTotal, st>>>(map_a, map_b, cm, c, ldc, m, n, k, 0, 0, 0, 0, f.a_src, f.a_tgt, f.nh, f.s_src, f.s_tgt,
                                                                     g.act_a, g.act_b, g.mul_src, g.mul_ld);
    GAT_LAUNCH_CHECK();
    return GAT_OK;
  }
  TnPlan p = tn_plan(m, n, k, BN);
  const uint32_t mn_lbo = BK * 128, mn_sbo = 512;   // measured on B200: the swapped assignment gives wrong products
  dim3 grid((unsigned)((m + BM - 1) / BM), (unsigned)((n + BN - 1) / BN), (unsigned)p.splits);
  if (p.splits == 1) {
    gemm_tc_kernel<BN, MN><<<grid, kThreads, Smem<BN, MN>::kTotal, st>>>(map_a, map_b, cm, c, ldc, m, n, k, p.kb_per_split, 0, mn_lbo, mn_sbo, nullptr, nullptr, 0, nullptr, nullptr,
                                                                     g.act_a, g.act_b, nullptr, 0);
    GAT_LAUNCH_CHECK();
    return GAT_OK;
  }
  const size_t need = (size_t)p.splits * m * n * sizeof(float);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("gat_gemm: workspace too small (%zu < %zu)", workspace_bytes, need);
    return GAT_EWORKSPACE;
  }
  gemm_tc_kernel<BN, MN><<<grid, kThreads, Smem<BN, MN>::kTotal, st>>>(map_a, map_b, cm, (float*)workspace, n, m, n, k, p.kb_per_split, m * n, mn_lbo, mn_sbo, nullptr, nullptr, 0, nullptr, nullptr,
                                                                   g.act_a, g.act_b, nullptr, 0);
  GAT_LAUNCH_CHECK();
  splitk_reduce_launch((const float*)workspace, p.splits, m, n, c, ldc, st);
  GAT_LAUNCH_CHECK();
  return GAT_OK;
}

static int bn_for(int64_t n) { return n > 128 ? 256 : (n > 64 ? 128 : 64); }      // TN
static int bn_for_nt(int64_t n) { return n > 64 ? 128 : 64; }                      // NT: two CTAs per SM

}  // namespace tc

bool pair_supported(int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb, int64_t ldc);
int gemm_pair(int64_t m, int64_t n, int64_t k, const float* a, int64_t lda, const float* b, int64_t ldb, float* c, int64_t ldc,
              const float* a_src, const float* a_tgt, int nh, float* s_src, float* s_tgt,
              float* const* dests, int n_dests, int64_t row_offset,
              int act_a, const float* mul_src, int64_t mul_ld, cudaStream_t st);

bool pair_tn_supported(int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb, int64_t ldc);
int gemm_pair_tn(int64_t m, int64_t n, int64_t k, const float* a, int64_t lda, const float* b, int64_t ldb, float* c, int64_t ldc,
                 int act_b, cudaStream_t st);

bool tc_supported(int ta, int tb, int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb, int64_t ldc) {
  const bool nt = (ta == 0 && tb == 1), tn = (ta == 1 && tb == 0);
  if (!nt && !tn) return false;
  if (lda % 4 || ldb % 4 || ldc % 4) return false;             // TMA global strides / float4 stores need 16-byte multiples
  if (m < 1 || n < 8 || k < 1) return false;
  if (m >= ((int64_t)1 << 31) || n >= ((int64_t)1 << 31) || k >= ((int64_t)1 << 31)) return false;
  return true;
}

size_t tc_workspace_bytes(int ta, int tb, int64_t m, int64_t n, int64_t k) {
  if (!(ta == 1 && tb == 0) || m < 1 || n < 1 || k < 1) return 0;
  tc::TnPlan p = tc::tn_plan(m, n, k, tc::bn_for(n));
  return p.splits > 1 ? (size_t)p.splits * m * n * sizeof(float) : 0;
}

// Forward projection with the score epilogue fused: wh = x W^T, s_src = wh A_src^T, s_tgt = wh A_tgt^T.  Needs the whole
// row of wh in one CTA (dp <= 256) and dense attention halves (leading dimension dp).
bool tc_project_supported(int64_t n_rows, int64_t dp, int64_t k, int64_t ldx, int64_t ldw) {
  return dp <= 256 && dp % 4 == 0 && tc_supported(0, 1, n_rows, dp, k, ldx, ldw, dp);
}

int gemm_tc_project(int64_t n_rows, int64_t dp, int64_t k, const float* x, int64_t ldx, const float* w, int64_t ldw,
                    float* wh, const float* a_src, const float* a_tgt, int nh, float* s_src, float* s_tgt, cudaStream_t st,
                    float* const* wh_dests, int n_dests, int64_t row_offset, int x_act) {
  uintptr_t bits = (uintptr_t)x | (uintptr_t)w | (uintptr_t)a_src | (uintptr_t)a_tgt | (n_dests ? 0 : (uintptr_t)wh);
  for (int d = 0; d < n_dests; ++d) bits |= (uintptr_t)wh_dests[d];
  if (!tc_project_supported(n_rows, dp, k, ldx, ldw) || bits % 16) {
    set_error("gat_project_fwd: fused tcgen05 path unsupported for this shape/alignment");
    return GAT_EUNSUPPORTED;
  }
  if (pair_supported(n_rows, dp, k, ldx, ldw, dp) && nh <= 8)   // large M, short K: persistent CTA-pair kernel, scores as extra columns
    return gemm_pair(n_rows, dp, k, x, ldx, w, ldw, wh, dp, a_src, a_tgt, nh, s_src, s_tgt, wh_dests, n_dests, row_offset, x_act, nullptr, 0, st);
  tc::ScoreFuse f{a_src, a_tgt, nh, s_src, s_tgt};
  tc::Dests dd{wh_dests, n_dests, row_offset};
  tc::Glue g{x_act, 0, nullptr, 0};
  const int bn = tc::bn_for_nt(dp);
  if (bn == 128) return tc::launch<128, false>(n_rows, dp, k, x, ldx, w, ldw, wh, dp, nullptr, 0, st, f, dd, g);
  return tc::launch<64, false>(n_rows, dp, k, x, ldx, w, ldw, wh, dp, nullptr, 0, st, f, dd, g);
}

int gemm_tc(int ta, int tb, int64_t m, int64_t n, int64_t k, const float* a, int64_t lda, const float* b, int64_t ldb,
            float* c, int64_t ldc, void* workspace, size_t workspace_bytes, cudaStream_t st,
            int act_a, int act_b, const float* mul_src, int64_t mul_ld) {
  if (!tc_supported(ta, tb, m, n, k, lda, ldb, ldc) || ((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) % 16) {
    set_error("gat_gemm: tcgen05 path needs (ta,tb) = (0,1) or (1,0) and 16-byte aligned pointers / leading dimensions");
    return GAT_EUNSUPPORTED;
  }
  if (ta == 0 && act_b == 0 && pair_supported(m, n, k, lda, ldb, ldc))
    return gemm_pair(m, n, k, a, lda, b, ldb, c, ldc, nullptr, nullptr, 0, nullptr, nullptr, nullptr, 0, 0, act_a, mul_src, mul_ld, st);
  const int bn = ta == 0 ? tc::bn_for_nt(n) : tc::bn_for(n);
  const tc::ScoreFuse nf{nullptr, nullptr, 0, nullptr, nullptr};
  const tc::Dests nd{nullptr, 0, 0};
  const tc::Glue g{act_a, act_b, mul_src, mul_ld};
  if (ta == 0) {
    if (bn == 128) return tc::launch<128, false>(m, n, k, a, lda, b, ldb, c, ldc, workspace, workspace_bytes, st, nf, nd, g);
    return tc::launch<64, false>(m, n, k, a, lda, b, ldb, c, ldc, workspace, workspace_bytes, st, nf, nd, g);
  }
  if (mul_src != nullptr) { set_error("gat_gemm: the ELU' output multiplier is only implemented for ta = 0"); return GAT_EUNSUPPORTED; }
  if (act_a == 0 && pair_tn_supported(m, n, k, lda, ldb, ldc))     // dW of a large graph: persistent CTA-pair kernel
    return gemm_pair_tn(m, n, k, a, lda, b, ldb, c, ldc, act_b, st);
  if (bn == 256) return tc::launch<256, true>(m, n, k, a, lda, b, ldb, c, ldc, workspace, workspace_bytes, st, nf, nd, g);
  if (bn == 128) return tc::launch<128, true>(m, n, k, a, lda, b, ldb, c, ldc, workspace, workspace_bytes, st, nf, nd, g);
  return tc::launch<64, true>(m, n, k, a, lda, b, ldb, c, ldc, workspace, workspace_bytes, st, nf, nd, g);
}

}  // namespace gat
