// Kernel 2, large-M form: persistent CTA-pair 3xTF32 GEMM for the NT products with a short contraction (K <= 256):
//   Wh = x W^T (+ the per-node score terms), dX = dWh W.        gat_layer.py:64-65, :76-82
//
// The one-CTA-per-tile kernel in gemm_tc.cu is bound by shared-memory bandwidth: every 16-wide k-block it fills both
// operand tiles by TMA, reads and re-writes them for the hi/lo split, and the tensor core reads each of them three
// times.  This kernel removes most of that traffic:
//   * cta_group::2 -- a cluster of two CTAs (one per SM of a TPC) computes a 256 x 128 tile; each CTA keeps only HALF of
//     the B operand (64..72 rows of W), and the tensor cores exchange the halves themselves;
//   * B (W, hi and lo parts, all of K) is RESIDENT in shared memory for the whole kernel: split once per launch by a
//     small prep kernel, loaded once per CTA -- no per-tile B traffic at all;
//   * A never sits in shared memory as an MMA operand: TMA lands the raw fp32 rows in a small staging ring, the four
//     splitter warps read each row once and write hi (the raw value: the tensor core reads the top 19 bits) and
//     lo = round_tf32(v - trunc_tf32(v)) straight into TENSOR MEMORY (tcgen05.st), and the MMAs take A from TMEM
//     (tcgen05.mma [d], [a], b_desc).  Per k-block the shared-memory traffic drops from 96 KB to 28 KB per CTA and the main
//     loop becomes MMA bound (3 x 2 MMAs of 64 clk per 16-wide k-block);
//   * persistent: 74 clusters walk the row tiles; accumulators (main: hi*hi, cross: hi*lo + lo*hi, see gemm_tc.cu) are
//     drained into registers in ~300 clk (tcgen05.ld runs at ~1 KB/clk/SM, tools/ubench/tmem_bw.cu) and released before
//     the epilogue stages / stores the tile, so the next tile's MMAs overlap the whole output path;
//   * the score terms s_src = Wh A_src^T, s_tgt = Wh A_tgt^T = x (A W)^T are 16 MORE COLUMNS OF THE SAME GEMM: the prep
//     kernel forms A W in fp64 and appends it to the first N tile's B operand (UMMA N = 144), which replaces the FP64
//     tensor-core epilogue (shared-memory transposes + DMMA, +1.2 ms per layer) of gemm_tc.cu.
//
// Roles (320 threads): warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer (leader CTA only issues), warps 2..5 splitters
// (thread = row), warps 6..9 epilogue (thread = row).  TMEM (512 columns): [0,144) main, [144,288) cross, 7 A stages of 32
// columns (16 hi + 16 lo).  Every wait is bounded (trap instead of hang).
#include "tc_common.cuh"
#include <stdlib.h>

namespace gat {
namespace tcp {

using namespace tc;

constexpr int PBM = 128;               // rows per CTA (UMMA M = 256 over the pair)
constexpr int PBN = 128;               // output columns per N tile
constexpr int NSC = 16;                // score columns appended to N tile 0
constexpr int NBROWS = PBN + NSC;      // rows of the split B operand per N tile
constexpr int BHALF = NBROWS / 2;      // rows resident per CTA (72; an N = 128 tile uses the first 64)
constexpr int KMAX = 256;
constexpr int KBMAX = KMAX / BK;       // 16 k-blocks
constexpr int kBTileBytes = BHALF * BK * 4;        // 4608: one k-block of the resident half, SWIZZLE_64B
constexpr int kBBytes = KBMAX * kBTileBytes;       // 73728 per hi / lo
constexpr int AK = 32;                 // fp32 elements of A per pipeline stage = two 16-wide k-blocks of B (one 128-byte swizzle row)
constexpr int kRawStages = 3;
constexpr int kRawBytes = PBM * AK * 4;            // 16384
constexpr int kAStages = 3;
constexpr int kAStageCols = 2 * AK;                // 32 hi + 32 lo columns
constexpr int kStageCBytes = PBM * 128;            // one 32-column box of the output tile
constexpr int kSmemTotal = 2 * kBBytes + kRawStages * kRawBytes + 2 * kStageCBytes + 1024 + 256;
static_assert(kSmemTotal <= 227 * 1024, "shared memory budget");
constexpr int kAccMain = 0, kAccCross = NBROWS, kACol0 = 2 * NBROWS;   // TMEM columns
static_assert(kACol0 + kAStages * kAStageCols <= 512, "TMEM budget");
constexpr int kPairThreads = 320;

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster.  Default (.release.cta) semantics on
// purpose: .release.cluster compiles to MEMBAR.ALL.GPU + ERRBAR before every arrival (one per k-block and warp), and what is
// handed over here is tensor-memory state, ordered by tcgen05.fence::before_thread_sync / after_thread_sync.
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
               "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (spins > (1u << 26)) __trap();
  }
}
// D[tmem] (+)= A[tmem] * B[smem]^T over the CTA pair.  Executed by ALL lanes of the leader CTA's MMA warp with warp-uniform
// operands; `elected` (one lane, chosen once by elect.sync) predicates the instruction itself.  Keeping the control flow
// around it warp-uniform matters: inside an `if (lane == 0)` region ptxas re-derives every uniform-register operand with
// ELECT + R2UR.BROADCAST (about 16 instructions and 70 clk per MMA), which made the issuing thread, not the tensor core, the
// limiter of the main loop (52 % tensor-pipe activity in ncu).
__device__ __forceinline__ void umma_tf32_ts2(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, uint32_t accumulate, uint32_t elected) {
  asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
               "@q tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(tmem_d), "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(elected) : "memory");
}
// completion of all MMAs issued so far -> one arrival on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit2(uint64_t* bar, uint32_t elected) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t"
               "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3), "r"(elected) : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t e;
  asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(e));
  return e;
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]),
                 "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
               ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]),
                 "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]),
                 "f"(v[16]), "f"(v[17]), "f"(v[18]), "f"(v[19]), "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]),
                 "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]), "f"(v[28]), "f"(v[29]), "f"(v[30]), "f"(v[31]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]),
                 "=f"(r[16]), "=f"(r[17]), "=f"(r[18]), "=f"(r[19]), "=f"(r[20]), "=f"(r[21]), "=f"(r[22]), "=f"(r[23]), "=f"(r[24]), "=f"(r[25]), "=f"(r[26]), "=f"(r[27]), "=f"(r[28]), "=f"(r[29]), "=f"(r[30]), "=f"(r[31])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15])
               : "r"(taddr) : "memory");
}
// explicit shared-space accesses (a generic LD/ST through a computed pointer takes the slower generic path and, being
// asynchronous, may still be in flight when a later mbarrier arrival hands the buffer back to TMA)
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void tma_store_wait_read(int pending) {
  if (pending == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}

// Instruction descriptor for the pair: D = f32, A = B = tf32, both K-major, M = 256, N = n.
__host__ __device__ constexpr uint32_t make_idesc_pair(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

// Split operand B' per N tile t (rows t*144 ..): 128 rows of W (zero beyond N), then 16 score rows (tile 0: row j < nh is
// A_src[j,:] W, row 8 + j is A_tgt[j,:] W, formed in fp64; zero otherwise).  hi = round_tf32(v),
// lo = round_tf32(v - hi); out[0 .. R*Kp) = hi, out[R*Kp .. 2*R*Kp) = lo with R = n_tiles*144 rows of Kp floats.
__global__ void pair_prep_b_kernel(const float* __restrict__ w, int64_t ldw, int n, int k, int kp, int n_tiles,
                                   const float* __restrict__ a_src, const float* __restrict__ a_tgt, int nh,
                                   float* __restrict__ out) {
  const int R = n_tiles * NBROWS;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < R * kp; i += gridDim.x * blockDim.x) {
    const int row = i / kp, kk = i - row * kp;
    const int t = row / NBROWS, rr = row - t * NBROWS;
    float v = 0.f;
    if (kk < k) {
      if (rr < PBN) {
        const int col = t * PBN + rr;
        if (col < n) v = __ldg(w + (int64_t)col * ldw + kk);
      } else if (t == 0 && a_src != nullptr && ((rr - PBN) & 7) < nh) {
        const int j = rr - PBN;                   // score column: j < 8 source term of head j, j >= 8 target term of head j - 8
        const float* av = j < 8 ? a_src + (int64_t)j * n : a_tgt + (int64_t)(j - 8) * n;
        double acc = 0.0;
        for (int d = 0; d < n; ++d) acc += (double)__ldg(av + d) * (double)__ldg(w + (int64_t)d * ldw + kk);
        v = (float)acc;
      }
    }
    const float hi = tf32_round(v);
    out[i] = hi;
    out[(int64_t)R * kp + i] = tf32_round(v - hi);
  }
}

struct PairArgs {
  int64_t M;            // rows of A / C
  int N;                // output columns (<= 256)
  int num_kb;           // ceil(K / 32) (<= 8) A stages per tile; B holds 2 * num_kb 16-wide k-blocks
  int n_tiles;          // ceil(N / 128)
  int m_pairs;          // ceil(M / 256)
  int clusters0;        // clusters walking the row tiles of N tile 0; the remaining gridDim/2 - clusters0 walk N tile 1
  int act_a;            // ELU on the A operand
  int nh;               // heads of the fused score terms (0: none)
  float* s_src;
  float* s_tgt;
  const float* mul_src; // output multiplied by ELU'(mul_src[row, col]) or nullptr
  int64_t mul_ld;
  int debug;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CStoreMaps cmaps, const PairArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* b_hi = smem;
  uint8_t* b_lo = smem + kBBytes;
  uint8_t* raw = smem + 2 * kBBytes;
  uint8_t* stage_c = raw + kRawStages * kRawBytes;
  uint64_t* bars = (uint64_t*)(stage_c + 2 * kStageCBytes);
  uint64_t* raw_full = bars;                       // TMA landed a raw A k-block            (1 + tx)
  uint64_t* raw_empty = raw_full + kRawStages;     // the 4 splitter warps have read it      (4)
  uint64_t* a_ready = raw_empty + kRawStages;      // hi/lo of a k-block are in TMEM of BOTH CTAs (8 warps; leader's copy is used)
  uint64_t* a_free = a_ready + kAStages;           // the MMAs that read the TMEM stage have completed (commit, both CTAs)
  uint64_t* acc_full = a_free + kAStages;          // all MMAs of the tile have completed    (commit, both CTAs)
  uint64_t* acc_empty = acc_full + 1;              // both CTAs' epilogue warps drained the accumulators (8 warps; leader's copy)
  uint64_t* b_full = acc_empty + 1;                // the resident B half has landed         (1 + tx)
  uint32_t* tmem_ptr = (uint32_t*)(b_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  // Each cluster owns one N tile (its B half stays resident) and walks the row tiles with a stride.  The two groups of
  // clusters are sized in proportion to their MMA work per tile (144 : 128 columns when the score columns ride on tile 0),
  // so both groups advance through the rows at the same pace and the second read of every A tile hits L2 -- with equal
  // groups the slower one fell ~15 row waves behind by the end and DRAM saw A twice (ncu: 5.0 GB read for 2.5 GB).
  const int n_clusters = (int)(gridDim.x >> 1);
  const int nt = cluster_id < p.clusters0 ? 0 : 1;            // this cluster's N tile
  const int first_pair = nt == 0 ? cluster_id : cluster_id - p.clusters0;
  const int pair_stride = nt == 0 ? p.clusters0 : n_clusters - p.clusters0;
  const int n0 = nt * PBN;
  const bool scores = (nt == 0) && p.nh > 0;
  const int umma_n = scores ? NBROWS : PBN;
  const int num_kb = p.num_kb;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kRawStages; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], 4); }
    for (int s = 0; s < kAStages; ++s) { mbar_init(&a_ready[s], 8); mbar_init(&a_free[s], 1); }
    mbar_init(acc_full, 1); mbar_init(acc_empty, 8); mbar_init(b_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_ptr)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===== TMA producer: the resident B half once, then the raw A k-blocks of every tile =====
    if (lane == 0) {
      const int brow = nt * NBROWS + (int)rank * (umma_n / 2);
      const int lo_row0 = p.n_tiles * NBROWS;
      mbar_expect_tx(b_full, 2 * (2 * num_kb) * kBTileBytes);
      for (int kb = 0; kb < 2 * num_kb; ++kb) {
        tma_load_2d(b_hi + kb * kBTileBytes, &map_b, b_full, kb * BK, brow);
        tma_load_2d(b_lo + kb * kBTileBytes, &map_b, b_full, kb * BK, lo_row0 + brow);
      }
      // The staging ring holds two loads in flight (32 KB per SM), far less than HBM latency x the bandwidth this CTA
      // consumes; so every tile's rows are pulled into L2 one tile (~3 us) ahead and the ring only has to cover L2 latency.
      auto prefetch_tile = [&](int pr_) {
        if (pr_ < p.m_pairs)
          for (int kb = 0; kb < num_kb; ++kb)
            asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(&map_a), "r"(kb * AK), "r"(pr_ * 256 + (int)rank * PBM) : "memory");
      };
      prefetch_tile(first_pair);
      uint32_t it = 0;
      for (int pr = first_pair; pr < p.m_pairs; pr += pair_stride) {
        const int m0 = pr * 256 + (int)rank * PBM;
        prefetch_tile(pr + pair_stride);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % kRawStages;
          if (it >= kRawStages) mbar_wait(&raw_empty[s], ((it / kRawStages) - 1) & 1);
          mbar_expect_tx(&raw_full[s], kRawBytes);
          tma_load_2d(raw + s * kRawBytes, &map_a, &raw_full[s], kb * AK, m0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA): the whole warp walks the loop, one elected lane issues =====
    if (rank == 0) {
      const uint32_t elected = elect_one();
      const uint32_t idesc = make_idesc_pair(umma_n);
      const uint32_t bh = smem_u32(b_hi), bl = smem_u32(b_lo);
      uint32_t it = 0, t = 0;
      for (int pr = first_pair; pr < p.m_pairs; pr += pair_stride, ++t) {
        if (t > 0) mbar_wait_cluster(acc_empty, (t - 1) & 1);          // previous tile drained by both epilogues
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t a = it % kAStages;
          mbar_wait_cluster(&a_ready[a], (it / kAStages) & 1);
          tc_fence_after();
          const uint32_t acol = tmem_base + kACol0 + a * kAStageCols;
#pragma unroll
          for (int k = 0; k < AK / UMMA_K; ++k) {
            const uint32_t first = (kb | k) != 0;
            const uint32_t a_hi = acol + k * UMMA_K, a_lo = a_hi + AK;
            const uint32_t boff = (2 * kb + (k >> 1)) * kBTileBytes + (k & 1) * UMMA_K * 4;   // 16-wide k-block of B, 8-wide step inside it
            const uint64_t d_hi = make_desc_k_sw64(bh + boff);
            const uint64_t d_lo = make_desc_k_sw64(bl + boff);
            umma_tf32_ts2(tmem_base + kAccCross, a_hi, d_lo, idesc, first, elected);
            umma_tf32_ts2(tmem_base + kAccCross, a_lo, d_hi, idesc, 1, elected);
            umma_tf32_ts2(tmem_base + kAccMain, a_hi, d_hi, idesc, first, elected);
          }
          umma_commit2(&a_free[a], elected);
        }
        umma_commit2(acc_full, elected);
      }
    }
    __syncwarp();
  } else if (warp < 6) {
    // ===== splitters: raw fp32 row (shared memory) -> hi / lo in tensor memory =====
    const int q = warp & 3;                           // TMEM lane quarter this warp may touch
    const int r = q * 32 + lane;                      // row of the CTA's 128-row tile
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + kACol0;
    const int sw = r & 7;                             // SWIZZLE_128B: 16-byte chunk j of row r sits at chunk j ^ (r % 8)
    const uint32_t raw_u32 = smem_u32(raw) + r * 128;
    mbar_wait(b_full, 0);                             // the first a_ready arrival also tells the leader that B is resident
    const int my_tiles = first_pair < p.m_pairs ? (p.m_pairs - first_pair + pair_stride - 1) / pair_stride : 0;
    const uint32_t total = (uint32_t)my_tiles * (uint32_t)num_kb;
    // One k-block per iteration; the next k-block's shared-memory loads are issued before this one is processed, so the
    // chain wait -> load -> split -> tcgen05.st -> wait::st -> arrive of one k-block overlaps the next one's load.
    float cur[AK], nxt[AK];
    if (total > 0) {
      mbar_wait(&raw_full[0], 0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 v = lds128(raw_u32 + ((j ^ sw) << 4));
        cur[4 * j] = v.x; cur[4 * j + 1] = v.y; cur[4 * j + 2] = v.z; cur[4 * j + 3] = v.w;
      }
    }
    for (uint32_t it = 0; it < total; ++it) {
      const uint32_t s = it % kRawStages, a = it % kAStages;
      if (it + 1 < total) {
        const uint32_t s1 = (it + 1) % kRawStages;
        mbar_wait(&raw_full[s1], ((it + 1) / kRawStages) & 1);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 v = lds128(raw_u32 + s1 * kRawBytes + ((j ^ sw) << 4));
          nxt[4 * j] = v.x; nxt[4 * j + 1] = v.y; nxt[4 * j + 2] = v.z; nxt[4 * j + 3] = v.w;
        }
      }
      if (p.act_a) {
#pragma unroll
        for (int j = 0; j < AK; ++j) cur[j] = elu1(cur[j]);
      }
      float lo[AK];
#pragma unroll
      for (int j = 0; j < AK; ++j) lo[j] = lo1(cur[j]);
      // the loads of stage s must have LANDED (not merely been issued) before it is handed back to TMA: make the arrival
      // depend on one value of each 16-byte load
      asm volatile("" ::"f"(lo[0]), "f"(lo[4]), "f"(lo[8]), "f"(lo[12]), "f"(lo[16]), "f"(lo[20]), "f"(lo[24]), "f"(lo[28]) : "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&raw_empty[s]);      // the raw stage can be refilled
      if (it >= kAStages) { mbar_wait(&a_free[a], ((it / kAStages) - 1) & 1); tc_fence_after(); }
      tmem_st32(trow + a * kAStageCols, cur);
      tmem_st32(trow + a * kAStageCols + AK, lo);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&a_ready[a], 0);
#pragma unroll
      for (int j = 0; j < AK; ++j) cur[j] = nxt[j];
    }
  } else {
    // ===== epilogue: TMEM -> registers (accumulators released at once) -> swizzled staging -> TMA store =====
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int et = threadIdx.x - 6 * 32;              // 0..127
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t t = 0, boxes = 0;
    // The ELU' multiplier rows come from HBM (latency > 1 us): pull the NEXT tile's rows into L2 one tile ahead.
    auto prefetch_mul = [&](int pr_) {
      if (p.mul_src != nullptr && pr_ < p.m_pairs) {
        const int64_t g = (int64_t)pr_ * 256 + (int)rank * PBM + r;
        if (g < p.M) {
          const float* mp = p.mul_src + g * p.mul_ld + n0;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (n0 + c * 32 < p.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(mp + c * 32));
        }
      }
    };
    prefetch_mul(first_pair);
    for (int pr = first_pair; pr < p.m_pairs; pr += pair_stride, ++t) {
      const int m0 = pr * 256 + (int)rank * PBM;
      const int64_t grow = (int64_t)m0 + r;
      prefetch_mul(pr + pair_stride);
      mbar_wait(acc_full, t & 1);
      tc_fence_after();
      if (scores) {                                   // the 16 score columns first: their registers are dead before the tile is loaded
        float sc[16], y[16];
        tmem_ld16(trow + kAccMain + PBN, sc);
        tmem_ld16(trow + kAccCross + PBN, y);
        tmem_ld_wait();
        if (grow < p.M) {
          float* ss = p.s_src + grow * p.nh;
          float* st = p.s_tgt + grow * p.nh;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (j < p.nh) { ss[j] = sc[j] + y[j]; st[j] = sc[8 + j] + y[8 + j]; }
          }
        }
      }
      float v[4][32];
      {
        float x[32];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tmem_ld32(trow + kAccMain + c * 32, v[c]);
          tmem_ld32(trow + kAccCross + c * 32, x);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[c][j] += x[j];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (!(p.debug & 2)) { if (lane == 0) mbar_arrive_cluster(acc_empty, 0); }          // the next tile's MMAs may overwrite the accumulators
      if (p.debug & 1) { for (int z = 0; z < 20; ++z) __nanosleep(1000); }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (n0 + c * 32 < p.N) {                                  // uniform over the CTA
          if (p.mul_src != nullptr && grow < p.M) {
            const float* mp = p.mul_src + grow * p.mul_ld + n0 + c * 32;
            if (n0 + c * 32 + 32 <= p.N && (p.mul_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(p.mul_src) & 15) == 0) {
              float4 mv[8];                                    // all eight loads in flight before the first use
#pragma unroll
              for (int j = 0; j < 8; ++j) mv[j] = __ldg(reinterpret_cast<const float4*>(mp) + j);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                v[c][4 * j] *= elu_grad1(mv[j].x); v[c][4 * j + 1] *= elu_grad1(mv[j].y);
                v[c][4 * j + 2] *= elu_grad1(mv[j].z); v[c][4 * j + 3] *= elu_grad1(mv[j].w);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (n0 + c * 32 + j < p.N) v[c][j] *= elu_grad1(__ldg(mp + j));
            }
          }
          uint8_t* box = stage_c + (boxes & 1) * kStageCBytes;
          const uint32_t bx = smem_u32(box) + r * 128;
          if (et == 0) tma_store_wait_read(1);                   // the store that last read this buffer (two boxes ago) is done
          asm volatile("bar.sync 1, 128;" ::: "memory");
          const int swc = r & 7;                                  // SWIZZLE_128B: chunk j of row r at j ^ (r % 8)
#pragma unroll
          for (int j = 0; j < 8; ++j)
            sts128(bx + ((j ^ swc) << 4), v[c][4 * j], v[c][4 * j + 1], v[c][4 * j + 2], v[c][4 * j + 3]);
          fence_proxy_async();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (et == 0) {
            for (int d = 0; d < cmaps.count; ++d) tma_store_2d(&cmaps.maps[d], box, n0 + c * 32, cmaps.row_offset + m0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          ++boxes;
        }
      }
      if (p.debug & 2) { __syncwarp(); if (lane == 0) mbar_arrive_cluster(acc_empty, 0); }
    }
    if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores complete before the CTA exits
  }
  tc_fence_before();
  cluster_sync_all();                                  // both CTAs are done with TMEM, barriers and each other's memory
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
  }
}

// ======================================================================================================================
// TN form for the weight gradient: C[M <= 256, N <= 256] = A[K, M]^T B[K, N], K = number of nodes (dW = dWh^T x, the adjoint of
// gat_layer.py:64).  Same machinery as above with the roles of the operands adapted to the long contraction:
//   * every cluster owns one 128-column half of C and one contiguous range of node rows (split-K); per 32-node stage each
//     CTA loads a [32 x 128] slice of A (its half of the M = 256 rows of C) and a [32 x 64] slice of B (its half of the
//     cluster's 128 C columns) by TMA;
//   * A goes through tensor memory: thread m reads COLUMN m of the raw tile (conflict-free LDS.32), i.e. the transposition the
//     MN-major operand would need happens on the way into TMEM (lane = m, column = node), hi = raw word, lo = round_tf32(v - trunc);
//   * B stays in shared memory in the MN-major SWIZZLE_128B_BASE32B layout the tensor core reads (see gemm_tc.cu); the
//     splitters write only its lo companion tile (position preserving);
//   * the tensor core rounds its fp32 accumulator toward zero at every MMA, so one accumulator sees at most 64 stages (2048
//     node rows): then both CTAs drain it and add it IN FP64 into the cluster's own slot of a (splits, 256, 256) fp64 buffer
//     (read-modify-write by one owner: deterministic); a small kernel sums the slots.
// The one-CTA-per-tile TN kernel it replaces on the products shapes ran at 2.1 ms (shared-memory bound: both operands split in
// shared memory and read three times); this one moves 88 KB of shared memory per stage and CTA against 768 clk of MMA.
constexpr int TK = 32;                       // node rows per stage
constexpr int kTnStages = 6;
constexpr int kTnABytes = TK * 128 * 4;      // raw A slice [32][128], unswizzled (read by columns)
constexpr int kTnBBytes = TK * 64 * 4;       // B slice: two boxes of [32 k][32 n] in the MN-major atom layout
constexpr int kTnStageBytes = kTnABytes + 2 * kTnBBytes;     // + lo companion of B
constexpr int kTnSmemTotal = kTnStages * kTnStageBytes + 1024 + 256;
static_assert(kTnSmemTotal <= 227 * 1024, "shared memory budget");
constexpr int kTnAStages = 3, kTnAccCross = 128, kTnACol0 = 256;   // TMEM: [0,128) main, [128,256) cross, 3 x 64 A columns
constexpr int kTnFlushStages = 64;

struct PairTnArgs {
  int64_t K;             // nodes
  int M, N;              // C is M x N
  int n_halves;          // ceil(N / 128)
  int n_splits;          // clusters per half
  int stages_per_split;  // 32-node stages per split (multiple of 1)
  int act_b;             // ELU on the B operand (fused input activation of the layer)
  double* slots;         // (n_splits, 256, 256) fp64
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
gemm_pair_tn_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const PairTnArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + kTnStages * kTnStageBytes);
  uint64_t* full = bars;                         // TMA landed stage s                          (1 + tx)
  uint64_t* stage_free = full + kTnStages;       // the MMAs that read B of stage s completed   (commit, both CTAs)
  uint64_t* a_ready = stage_free + kTnStages;    // A hi/lo in TMEM + B lo written, both CTAs   (8 warps, leader's copy)
  uint64_t* a_free = a_ready + kTnAStages;       // the MMAs that read TMEM stage a completed   (commit, both CTAs)
  uint64_t* acc_full = a_free + kTnAStages;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_ptr = (uint32_t*)(acc_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int nhalf = cluster_id % p.n_halves, split = cluster_id / p.n_halves;
  const int64_t total_stages = (p.K + TK - 1) / TK;
  const int64_t st0 = (int64_t)split * p.stages_per_split;
  const int my_stages = (int)max((int64_t)0, min((int64_t)p.stages_per_split, total_stages - st0));

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTnStages; ++s) { mbar_init(&full[s], 1); mbar_init(&stage_free[s], 1); }
    for (int s = 0; s < kTnAStages; ++s) { mbar_init(&a_ready[s], 8); mbar_init(&a_free[s], 1); }
    mbar_init(acc_full, 1); mbar_init(acc_empty, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_ptr)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      const int m_col = (int)rank * 128;                           // this CTA's rows of C = columns of A
      const int n_col = nhalf * 128 + (int)rank * 64;              // this CTA's half of the cluster's C columns = columns of B
      for (int it = 0; it < my_stages; ++it) {
        const int s = it % kTnStages;
        if (it >= kTnStages) mbar_wait(&stage_free[s], ((it / kTnStages) - 1) & 1);
        uint8_t* st = smem + s * kTnStageBytes;
        const int krow = (int)((st0 + it) * TK);
        mbar_expect_tx(&full[s], kTnABytes + kTnBBytes);
        tma_load_2d(st, &map_a, &full[s], m_col, krow);
        tma_load_2d(st + kTnABytes, &map_b, &full[s], n_col, krow);
        tma_load_2d(st + kTnABytes + TK * 128, &map_b, &full[s], n_col + 32, krow);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA), warp-uniform, one elected lane issues =====
    if (rank == 0) {
      const uint32_t elected = elect_one();
      // D = f32, A = B = tf32, A K-major (TMEM), B MN-major, M = 256, N = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      uint32_t flush = 0;
      for (int it = 0; it < my_stages; ++it) {
        const uint32_t s = it % kTnStages, a = it % kTnAStages;
        const int in_flush = it % kTnFlushStages;
        if (in_flush == 0 && it > 0) { mbar_wait_cluster(acc_empty, (flush - 1) & 1); tc_fence_after(); }
        mbar_wait_cluster(&a_ready[a], (it / kTnAStages) & 1);
        tc_fence_after();
        const uint32_t acol = tmem_base + kTnACol0 + a * 64;
        const uint32_t b_hi = smem_u32(smem + s * kTnStageBytes + kTnABytes), b_lo = b_hi + kTnBBytes;
#pragma unroll
        for (int k = 0; k < TK / UMMA_K; ++k) {
          const uint32_t first = (in_flush | k) != 0;
          const uint32_t a_hi = acol + k * UMMA_K, a_lo = a_hi + TK;
          const uint64_t d_hi = make_desc_mn_sw128(b_hi + k * 1024, TK * 128, 512);     // 8 node rows = two 4-row atoms = 1 KB
          const uint64_t d_lo = make_desc_mn_sw128(b_lo + k * 1024, TK * 128, 512);
          umma_tf32_ts2(tmem_base + kTnAccCross, a_hi, d_lo, idesc, first, elected);
          umma_tf32_ts2(tmem_base + kTnAccCross, a_lo, d_hi, idesc, 1, elected);
          umma_tf32_ts2(tmem_base, a_hi, d_hi, idesc, first, elected);
        }
        umma_commit2(&a_free[a], elected);
        umma_commit2(&stage_free[s], elected);
        if (in_flush == kTnFlushStages - 1 || it == my_stages - 1) { umma_commit2(acc_full, elected); ++flush; }
      }
    }
    __syncwarp();
  } else if (warp < 6) {
    // ===== splitters: column m of the raw A slice -> hi / lo in tensor memory; lo companion of the B slice =====
    // (measured: giving the B half of this work to the epilogue warps, with its own barrier ring, was SLOWER -- 2.13 vs 1.77 ms)
    const int q = warp & 3;
    const int r = q * 32 + lane;                      // row of C inside the CTA's half = column of the A slice
    const int t = (warp - 2) * 32 + lane;             // 0..127, for the B slice
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + kTnACol0;
    for (int it = 0; it < my_stages; ++it) {
      const uint32_t s = it % kTnStages, a = it % kTnAStages;
      mbar_wait(&full[s], (it / kTnStages) & 1);
      const uint32_t st = smem_u32(smem + s * kTnStageBytes);
      float hi[TK], lo[TK];
#pragma unroll
      for (int k = 0; k < TK; ++k) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(hi[k]) : "r"(st + k * 512 + r * 4));
#pragma unroll
      for (int k = 0; k < TK; ++k) lo[k] = lo1(hi[k]);
      // B: 512 16-byte chunks per CTA and stage, 4 per thread; hi stays where TMA put it unless the ELU is applied
      const uint32_t bh = st + kTnABytes, bl = bh + kTnBBytes;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t off = (uint32_t)(t + i * 128) * 16;
        float4 v = lds128(bh + off);
        if (p.act_b) { v = elu4(v); sts128(bh + off, v.x, v.y, v.z, v.w); }
        const float4 l = lo4(v);
        sts128(bl + off, l.x, l.y, l.z, l.w);
      }
      fence_proxy_async();                            // generic-proxy writes of B lo -> visible to the tensor core
      if (it >= kTnAStages) { mbar_wait(&a_free[a], ((it / kTnAStages) - 1) & 1); tc_fence_after(); }
      tmem_st32(trow + a * 64, hi);
      tmem_st32(trow + a * 64 + TK, lo);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&a_ready[a], 0);
    }
  } else {
    // ===== epilogue: every 64 stages drain main + cross and add them in fp64 into this cluster's slot =====
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const int m = (int)rank * 128 + r;
    double* slot = p.slots + ((int64_t)split * 256 + m) * 256 + nhalf * 128;
    const int n_flush = (my_stages + kTnFlushStages - 1) / kTnFlushStages;
    for (int f = 0; f < n_flush; ++f) {
      mbar_wait(acc_full, f & 1);
      tc_fence_after();
      float v[4][32];
      {
        float x[32];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tmem_ld32(trow + c * 32, v[c]);
          tmem_ld32(trow + kTnAccCross + c * 32, x);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[c][j] += x[j];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acc_empty, 0);
      if (m < p.M) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            double2* dp = reinterpret_cast<double2*>(slot + c * 32 + j);
            double2 acc = f == 0 ? make_double2(0.0, 0.0) : *dp;
            acc.x += (double)v[c][j]; acc.y += (double)v[c][j + 1];
            *dp = acc;
          }
        }
      }
    }
    if (n_flush == 0 && m < p.M) {                    // a cluster without rows still owns a slot: zero it
      for (int j = 0; j < 128; ++j) slot[j] = 0.0;
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
  }
}

__global__ void pair_tn_reduce_kernel(const double* __restrict__ slots, int n_splits, int M, int N, float* __restrict__ C, int64_t ldc) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * N) return;
  const int m = idx / N, n = idx - m * N;
  double s = 0.0;
  for (int z = 0; z < n_splits; ++z) s += slots[((int64_t)z * 256 + m) * 256 + n];
  C[(int64_t)m * ldc + n] = (float)s;
}

static int max_clusters() {
  static int cached[kMaxDevices] = {0};
  int& c = cached[cur_device()];
  if (c == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * kNumSMs, 1, 1);
    cfg.blockDim = dim3(kPairThreads, 1, 1);
    cfg.dynamicSmemBytes = kSmemTotal;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gemm_pair_kernel, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); n = kNumSMs / 2; }
    if (n > kNumSMs / 2) n = kNumSMs / 2;
    c = n;
  }
  return c;
}

}  // namespace tcp

// Shapes the pair kernel takes: NT, K <= 256, N <= 256, enough rows to fill the machine.  GAT_GEMM_PAIR=0 disables it
// (A/B timing against the one-CTA-per-tile kernel).
bool pair_supported(int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb, int64_t ldc) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("GAT_GEMM_PAIR"); enabled = (e && e[0] == '0') ? 0 : 1; }
  if (!enabled) return false;
  if (k < 1 || k > tcp::KMAX || n < 8 || n > 2 * tcp::PBN || m < 16384 || m >= ((int64_t)1 << 31) - 512) return false;
  if (lda % 4 || ldb % 4 || ldc % 4) return false;
  return true;
}

// C[M,N] = act(A)[M,K] * B[N,K]^T (* ELU'(mul_src)), optionally the score terms s = C A^T as extra columns; the output tile
// goes to `c` or, for the fused all-gather, to row_offset of every destination.
int gemm_pair(int64_t m, int64_t n, int64_t k, const float* a, int64_t lda, const float* b, int64_t ldb, float* c, int64_t ldc,
              const float* a_src, const float* a_tgt, int nh, float* s_src, float* s_tgt,
              float* const* dests, int n_dests, int64_t row_offset,
              int act_a, const float* mul_src, int64_t mul_ld, cudaStream_t st) {
  using namespace tcp;
  if (a_src != nullptr && (nh < 1 || 2 * nh > NSC)) { set_error("gemm_pair: fused scores support at most 8 heads"); return GAT_EUNSUPPORTED; }
  const int kp = (int)((k + AK - 1) / AK) * AK;      // B is padded with zero k-blocks to whole A stages
  const int n_tiles = (int)((n + PBN - 1) / PBN);
  const int R = n_tiles * NBROWS;
  // Scratch for the split operand comes from the device's stream-ordered pool.  Its default release threshold is 0: every
  // synchronisation would hand the memory back to the driver and the next call would map it again (device-wide stalls in
  // the middle of a step), so the pool is told once per device to keep what it has.
  static bool pool_set_dev[kMaxDevices] = {false};
  if (!pool_set_dev[cur_device()]) {
    cudaMemPool_t pool;
    int dev_id = 0;
    if (cudaGetDevice(&dev_id) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev_id) == cudaSuccess) {
      uint64_t keep = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
    pool_set_dev[cur_device()] = true;
  }
  float* bsplit = nullptr;
  GAT_CUDA(cudaMallocAsync((void**)&bsplit, (size_t)2 * R * kp * sizeof(float), st));
  {
    const int total = R * kp, threads = 256;
    pair_prep_b_kernel<<<(total + threads - 1) / threads, threads, 0, st>>>(b, ldb, (int)n, (int)k, kp, n_tiles, a_src, a_tgt, nh, bsplit);
    GAT_LAUNCH_CHECK();
  }
  CUtensorMap map_a, map_b;
  CStoreMaps cm;
  cm.count = 0; cm.row_offset = 0;
  int rc = make_map(&map_a, a, m, k, lda, PBM, kMapC128);   // 32-wide (128-byte) SWIZZLE_128B boxes of 128 rows
  if (!rc) rc = make_map(&map_b, bsplit, 2 * R, kp, kp, BHALF, kMapK64);
  if (n_dests > kMaxDests || row_offset + m >= ((int64_t)1 << 31) - 512) { set_error("gemm_pair: too many destinations / rows"); rc = GAT_EINVAL; }
  if (!rc) {
    if (n_dests == 0) { rc = make_map(&cm.maps[0], c, m, n, ldc, PBM, kMapC128); cm.count = 1; }
    else {
      for (int d = 0; d < n_dests && !rc; ++d) rc = make_map(&cm.maps[d], dests[d], row_offset + m, n, ldc, PBM, kMapC128);
      cm.count = n_dests; cm.row_offset = (int)row_offset;
    }
  }
  if (rc) { cudaFreeAsync(bsplit, st); return rc; }
  static bool attr_set_dev[kMaxDevices] = {false};
  bool& attr_set = attr_set_dev[cur_device()];
  if (!attr_set) {
    GAT_CUDA(cudaFuncSetAttribute(gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    attr_set = true;
  }
  PairArgs p;
  p.M = m; p.N = (int)n; p.num_kb = kp / AK; p.n_tiles = n_tiles; p.m_pairs = (int)((m + 255) / 256);
  int clusters = max_clusters();
  if (clusters > n_tiles * p.m_pairs) clusters = n_tiles * p.m_pairs;
  if (n_tiles == 1) p.clusters0 = clusters;
  else {
    if (clusters < 2) clusters = 2;
    const int w0 = a_src != nullptr ? NBROWS : PBN;     // MMA columns per tile of the two groups
    p.clusters0 = (clusters * w0 + (w0 + PBN) / 2) / (w0 + PBN);
    if (p.clusters0 < 1) p.clusters0 = 1;
    if (p.clusters0 > clusters - 1) p.clusters0 = clusters - 1;
  }
  p.act_a = act_a; p.nh = a_src != nullptr ? nh : 0; p.s_src = s_src; p.s_tgt = s_tgt; p.mul_src = mul_src; p.mul_ld = mul_ld;
  { const char* e = getenv("GAT_PAIR_DEBUG"); p.debug = e ? atoi(e) : 0; }
  const unsigned grid = (unsigned)(2 * clusters);
  gemm_pair_kernel<<<grid, kPairThreads, kSmemTotal, st>>>(map_a, map_b, cm, p);
  GAT_LAUNCH_CHECK();
  GAT_CUDA(cudaFreeAsync(bsplit, st));
  return GAT_OK;
}

// dW-shaped products: both operands row-major with the contraction over rows.
bool pair_tn_supported(int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb, int64_t ldc) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("GAT_GEMM_PAIR"); enabled = (e && e[0] == '0') ? 0 : 1; }
  if (!enabled) return false;
  if (m < 8 || m > 256 || n < 8 || n > 256 || k < 65536 || k >= ((int64_t)1 << 31) - 64) return false;
  if (lda % 4 || ldb % 4 || ldc < n) return false;
  return true;
}

static int make_map_plain(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols, bool atom32) {
  tc::EncodeTiledFn fn = tc::encode_fn();
  if (!fn) { set_error("gat_gemm: cuTensorMapEncodeTiled is unavailable"); return GAT_EUNSUPPORTED; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("gat_gemm: cuTensorMapEncodeTiled failed (%d)", (int)r); return GAT_EINVAL; }
  return GAT_OK;
}

int gemm_pair_tn(int64_t m, int64_t n, int64_t k, const float* a, int64_t lda, const float* b, int64_t ldb, float* c, int64_t ldc,
                 int act_b, cudaStream_t st) {
  using namespace tcp;
  CUtensorMap map_a, map_b;
  int rc = make_map_plain(&map_a, a, k, m, lda, TK, 128, false);
  if (!rc) rc = make_map_plain(&map_b, b, k, n, ldb, TK, 32, true);
  if (rc) return rc;
  PairTnArgs p;
  p.K = k; p.M = (int)m; p.N = (int)n; p.n_halves = n > 128 ? 2 : 1; p.act_b = act_b;
  const int64_t total_stages = (k + TK - 1) / TK;
  int clusters = max_clusters();
  int splits = clusters / p.n_halves;
  if (splits < 1) splits = 1;
  if (splits > total_stages) splits = (int)total_stages;
  p.stages_per_split = (int)((total_stages + splits - 1) / splits);
  p.n_splits = (int)((total_stages + p.stages_per_split - 1) / p.stages_per_split);
  static bool pool_set_dev[kMaxDevices] = {false};
  if (!pool_set_dev[cur_device()]) {
    cudaMemPool_t pool;
    int dev_id = 0;
    if (cudaGetDevice(&dev_id) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev_id) == cudaSuccess) {
      uint64_t keep = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
    pool_set_dev[cur_device()] = true;
  }
  double* slots = nullptr;
  GAT_CUDA(cudaMallocAsync((void**)&slots, (size_t)p.n_splits * 256 * 256 * sizeof(double), st));
  p.slots = slots;
  static bool attr_set_dev[kMaxDevices] = {false};
  bool& attr_set = attr_set_dev[cur_device()];
  if (!attr_set) {
    GAT_CUDA(cudaFuncSetAttribute(gemm_pair_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTnSmemTotal));
    attr_set = true;
  }
  gemm_pair_tn_kernel<<<(unsigned)(2 * p.n_splits * p.n_halves), kPairThreads, kTnSmemTotal, st>>>(map_a, map_b, p);
  GAT_LAUNCH_CHECK();
  pair_tn_reduce_kernel<<<(unsigned)((m * n + 255) / 256), 256, 0, st>>>(slots, p.n_splits, (int)m, (int)n, c, ldc);
  GAT_LAUNCH_CHECK();
  GAT_CUDA(cudaFreeAsync(slots, st));
  return GAT_OK;
}

}  // namespace gat
