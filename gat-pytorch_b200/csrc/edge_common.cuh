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
template <int NHT>
__device__ __forceinline__ void dropout_scales(uint64_t seed, uint64_t offset, uint32_t edge, int nh, float p,
                                               float (&m)[NHT]) {
  static_assert(NHT % 4 == 0, "head bound must be a multiple of 4");
  const float keep = 1.0f / (1.0f - p);
#pragma unroll
  for (int b = 0; b < NHT / 4; ++b) {
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
  value += 0.0f;   // canonicalise -0.0 to +0.0: as an int, -0.0 is INT_MIN and would lose against every stored pattern
  return (value >= 0.f) ? __int_as_float(atomicMax((int*)addr, __float_as_int(value)))
                        : __uint_as_float(atomicMin((unsigned int*)addr, __float_as_uint(value)));
}

// Persistent-grid row scheduler.  Rows are handed out to WARPS from a global counter in the order given by
// `order` (long rows first, built by gat_csr_build), so a hub row starts at t=0 and short rows fill in
// behind it; which warp computes a row never changes the row's result, so the output stays deterministic.
//
// LONG rows (more than kLongRow edges) are not handed to a single warp: one hub row of 16k edges would keep one warp
// busy for milliseconds after every other row has finished (measured on the products-shaped graph: the four longest
// rows, 63k edges, landed on one warp and set a ~8 ms floor under the forward kernel no matter how many GPUs shared the
// rest).  Instead every CTA first takes long rows from a second counter and processes each one COOPERATIVELY: its
// 256/G groups split the row's batches round-robin and their partial sums are combined through shared memory in a
// fixed order.  The path a row takes depends only on its own length, so results stay deterministic and independent of
// the schedule and of the partition.  Contract: when `order` is given, every long row precedes every short row in it
// (gat_csr_build emits exactly that); with order == nullptr there is no cooperative phase.
struct RowSched {
  const int32_t* order;      // permutation of [0, n), long rows first, or nullptr for natural order
  unsigned int* counter;     // warp-level row counter, zeroed before the launch
  int64_t n;
  unsigned int* cta_counter; // CTA-level long-row counter (the word after `counter`), zeroed before the launch
};

constexpr int kLongRow = GAT_LONG_ROW_EDGES;

// CTA-uniform: next long row for this CTA, or -1 when the long rows are exhausted.  Every thread must call it.
__device__ __forceinline__ int64_t grab_long_row(const RowSched& S, const int32_t* rowptr, int* sh_ctl) {
  __syncthreads();           // the previous row's shared-memory traffic (and reads of *sh_ctl) are done
  if (threadIdx.x == 0) {
    int r = -1;
    if (S.order) {
      const unsigned int c = atomicAdd(S.cta_counter, 1u);
      if ((int64_t)c < S.n) {
        const int row = __ldg(S.order + c);
        if (__ldg(rowptr + row + 1) - __ldg(rowptr + row) > kLongRow) r = row;
      }
    }
    *sh_ctl = r;
  }
  __syncthreads();
  return (int64_t)*sh_ctl;
}

// Warp phase: rows the cooperative phase has taken are skipped.
__device__ __forceinline__ bool taken_by_cta_phase(const RowSched& S, const int32_t* rowptr, int64_t row) {
  return S.order != nullptr && (__ldg(rowptr + row + 1) - __ldg(rowptr + row)) > kLongRow;
}

// A warp takes kGrabIters<G> rounds of 32/G rows per grab (never more than 32 rows: one prefetched entry per lane).
template <int G> constexpr int kGrabIters = G == 32 ? 4 : (G == 1 ? 1 : 2);

template <int G>
__device__ __forceinline__ bool grab_rows(const RowSched& S, int lane, int64_t& base) {
  constexpr int kRowsPerGrab = (32 / G) * kGrabIters<G>;
  __syncwarp();
  unsigned int b = 0;
  if (lane == 0) b = atomicAdd(S.counter, (unsigned int)kRowsPerGrab);
  b = __shfl_sync(0xffffffffu, b, 0);
  base = (int64_t)b;
  return base < S.n;
}

template <int G>
__device__ __forceinline__ int64_t sched_row(const RowSched& S, int64_t base, int k, int lane) {
  const int64_t idx = base + (int64_t)k * (32 / G) + lane / G;
  if (idx >= S.n) return -1;
  return S.order ? (int64_t)__ldg(S.order + idx) : idx;
}

// Row metadata of a whole grab, fetched by the first lanes in ONE round trip (order[idx] -> rowptr[row], rowptr[row+1]) instead
// of two dependent loads per row on the critical path of every row; iteration k reads its entry with shuffles.
// `rot` rotates the schedule (position idx -> idx + rot mod n): ranks of a partitioned run start their walk over the source
// rows at different owners' slabs, so that at any time they push to DIFFERENT peers (see BwdMainParams::sched_rot).
template <int G>
__device__ __forceinline__ void prefetch_rows(const RowSched& S, const int32_t* rowptr, const int64_t base, const int lane,
                                              int& pr, int& ps, int& pe, const int64_t rot = 0) {
  constexpr int kRowsPerGrab = (32 / G) * kGrabIters<G>;
  static_assert(kRowsPerGrab <= 32, "one prefetched row per lane");
  pr = -1; ps = 0; pe = 0;
  int64_t idx = base + lane;
  if (lane < kRowsPerGrab && idx < S.n) {
    idx += rot;
    if (idx >= S.n) idx -= S.n;
    pr = S.order ? __ldg(S.order + idx) : (int)idx;
    ps = __ldg(rowptr + pr);
    pe = __ldg(rowptr + pr + 1);
  }
}

// Entry of iteration k for this lane's group; returns false when there is no row (or the cooperative phase owns it).
template <int G>
__device__ __forceinline__ bool prefetched_row(const RowSched& S, const int k, const int lane, const int pr, const int ps, const int pe,
                                               int64_t& row, int& start, int& end) {
  const int j = k * (32 / G) + lane / G;
  row = (int64_t)__shfl_sync(0xffffffffu, pr, j);
  start = __shfl_sync(0xffffffffu, ps, j);
  end = __shfl_sync(0xffffffffu, pe, j);
  return row >= 0 && !(S.order != nullptr && end - start > kLongRow);
}

// Host-side choice of (G, SLOTS) for a padded row of `chunks` float4s.
// Narrow rows on SMALL graphs: the narrowest group that covers the row leaves most of the machine idle (PATTERN's 1x1 output
// layer: 15 341 rows of one chunk = 480 warps for 3 552 warp slots) and walks a hub row edge by edge (Cora's 169-edge row on a
// 2-lane group: 60 us of a 500 us step).  While the whole launch still fits one wave of resident warps the group is widened:
// a batch is then G edges whose softmax terms are computed by G lanes in parallel; the surplus lanes of the gather loop re-read
// chunk 0 (never stored).  Which G a row gets depends on (row width, rows in the launch) only, so results stay deterministic.
struct GroupShape { int g, slots; };
static inline GroupShape pick_group(int chunks, int64_t n_rows = -1) {
  GroupShape s;
  if (chunks <= 16) {
    s.slots = 1;
    s.g = chunks <= 1 ? 1 : (chunks <= 2 ? 2 : (chunks <= 4 ? 4 : (chunks <= 8 ? 8 : 16)));
    constexpr int64_t kWave = (int64_t)kNumSMs * 24;                 // resident warps at 3 CTAs of 8 warps per SM
    while (n_rows >= 0 && s.g < 32 && n_rows * (2 * s.g) <= kWave * 32) s.g *= 2;
  }
  else {
    s.g = 32;
    int need = (chunks + 31) / 32;
    const int allowed[] = {1, 2, 3, 4, 6, 8};
    s.slots = -1;
    for (int a : allowed) if (a >= need) { s.slots = a; break; }
  }
  return s;
}

#define GAT_DISPATCH_GROUP_NHT(shape, NHT_, LAUNCH)                               \
  do {                                                                            \
    if ((shape).g == 1) { LAUNCH(1, 1, NHT_); }                                   \
    else if ((shape).g == 2) { LAUNCH(2, 1, NHT_); }                              \
    else if ((shape).g == 4) { LAUNCH(4, 1, NHT_); }                              \
    else if ((shape).g == 8) { LAUNCH(8, 1, NHT_); }                              \
    else if ((shape).g == 16) { LAUNCH(16, 1, NHT_); }                            \
    else if ((shape).slots == 1) { LAUNCH(32, 1, NHT_); }                         \
    else if ((shape).slots == 2) { LAUNCH(32, 2, NHT_); }                         \
    else if ((shape).slots == 3) { LAUNCH(32, 3, NHT_); }                         \
    else if ((shape).slots == 4) { LAUNCH(32, 4, NHT_); }                         \
    else if ((shape).slots == 6) { LAUNCH(32, 6, NHT_); }                         \
    else { LAUNCH(32, 8, NHT_); }                                                 \
  } while (0)

// NHT = compile-time bound on the head count (4 or 8): halves the per-head register arrays for NH <= 4.
#define GAT_DISPATCH_GROUP(shape, nh, LAUNCH)                                     \
  do {                                                                            \
    if ((nh) <= 4) GAT_DISPATCH_GROUP_NHT(shape, 4, LAUNCH);                      \
    else GAT_DISPATCH_GROUP_NHT(shape, 8, LAUNCH);                                \
  } while (0)

// Programmatic dependent launch, used to OVERLAP two independent launches of one entry point (the cooperative long-row
// kernel and the short-row kernel touch disjoint rows): the first kernel releases its dependents at once
// (pdl_release_dependents), the second is launched with programmatic stream serialisation, so its CTAs become resident
// as soon as CTAs of the first retire, and each of its threads ends with pdl_wait_for_primary(), so the second grid --
// and with it the stream -- completes only after the first has fully completed and flushed.
__device__ __forceinline__ void pdl_release_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait_for_primary() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename Params>
static inline cudaError_t launch_kernel(void (*kernel)(const Params), unsigned grid, int threads, size_t dyn_smem, cudaStream_t st,
                                        const Params& params, bool overlap_with_previous) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3((unsigned)threads, 1, 1);
  cfg.dynamicSmemBytes = dyn_smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = overlap_with_previous ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, params);
}

// Persistent grid: enough CTAs to fill every SM at the kernel's occupancy, never more than the work.
template <typename K>
static inline unsigned persistent_grid(K kernel, int threads, size_t dyn_smem, int64_t work_blocks) {
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, dyn_smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  int64_t g = (int64_t)per_sm * kNumSMs;
  if (g > work_blocks) g = work_blocks;
  return (unsigned)(g < 1 ? 1 : g);
}

}  // namespace gat
