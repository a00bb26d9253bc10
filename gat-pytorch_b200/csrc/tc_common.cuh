// Shared tcgen05 / TMA / mbarrier helpers of the tensor-core GEMMs (gemm_tc.cu, gemm_pair.cu).  sm_100a only.
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace gat {
namespace tc {

constexpr int BM = 128;          // UMMA M (cta_group::1)
constexpr int BK = 16;           // fp32 elements per k-block = one 64-byte swizzle row (NT) / 16 k-rows of a 128-byte atom column (TN)
constexpr int CBOX = 32;         // columns per TMA box of the staged output tile (one 128-byte swizzle row)
constexpr int UMMA_K = 8;        // tf32
constexpr int kThreads = 192;
constexpr int kSplitThreads = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (spins > (1u << 26)) __trap();   // a broken pipeline must fault, not hang the GPU
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// TMA store of one swizzled shared-memory box; rows / columns beyond the tensor map's extent are clipped by the hardware.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit_and_wait() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_64B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4,
// LBO unused (1), SBO = 512 B (8 rows x 64 B) >> 4, version 1 (Blackwell), layout type 4 (SWIZZLE_64B).  A k-block is
// 16 fp32 = 64 B per row (not 128): a stage is half as large, so the pipeline is twice as deep in the same memory --
// load, split and MMA are three phases and need more than two stages to overlap.
__device__ __forceinline__ uint64_t make_desc_k_sw64(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
// MN-major tf32: the only layout the tensor core accepts is SWIZZLE_128B_BASE32B (layout type 1; TMA mode
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): an atom is 32 MN-elements (128 B) x 4 K-rows (512 B, 32-byte chunks XORed with
// the row index); atoms tile along MN with stride LBO and along K with stride SBO.  A TMA box of {32 MN-elements, BK
// k-rows} lands as BK/4 atoms stacked along K (SBO = 512 B); consecutive boxes along MN are BK*128 B apart (LBO).
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32, majorness bits 15/16, N>>3, M>>4.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int m, int n, bool mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((mn_major ? 1u : 0u) << 15) | ((mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Inter-layer glue fused into the GEMM (GATModel.py:148-149 applies F.elu between layers):
//   ELU(x)  on an operand tile while it is split in shared memory (the activated tensor is never written to HBM);
//   ELU'(x) as a multiplier of the output tile (the adjoint, for dX).
// (fast exp: the 4 splitter warps touch every operand element, so the activation must cost a handful of instructions;
// |error| <= 2e-7 absolute, far inside the 1e-5 parity bar)
__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : __expf(x) - 1.0f; }
__device__ __forceinline__ float elu_grad1(float x) { return x > 0.f ? 1.f : __expf(x); }
__device__ __forceinline__ float4 elu4(float4 v) { return make_float4(elu1(v.x), elu1(v.y), elu1(v.z), elu1(v.w)); }

// Round to tf32 precision (11 significant bits).  sm_100a has no hardware cvt.rna.tf32.f32: ptxas expands it into
// VIADD + LOP3 + FSETP + SEL on the half-rate integer pipe, which made the 4 splitter warps -- not the tensor core -- the
// limiter of the main loop (1400 clk per 16-wide k-block against 786 clk of MMA).  Veltkamp's splitting does the same
// rounding (to nearest, ties to even) in three full-rate FP32 operations and propagates NaN; |v| > 4e34 overflows to
// NaN, which no feature matrix reaches.  __fmul_rn / __fadd_rn keep the compiler from contracting the sequence into FMAs.
__device__ __forceinline__ float tf32_round(float v) {
  const float g = __fmul_rn(v, 8193.0f);      // 2^13 + 1
  return __fadd_rn(g, __fsub_rn(v, g));
}

// NT tiles are at most 128 columns wide and sized so that TWO CTAs are resident per SM (<= 113 KB of shared memory and
// 256 TMEM columns each): a non-persistent CTA spends ~10 us per tile outside its main loop (launch, TMEM allocation,
// pipeline fill, TMEM drain, output store) -- measured as the K-independent part of the tile time -- and with a second
// CTA on the SM that time is covered by the other CTA's main loop.  TN (split-K, long K, tiny epilogue) keeps one CTA
// per SM with 256-wide tiles and a deeper pipeline.
// The "hi" operand is the fp32 value itself, left where TMA put it: the tensor core reads only the top 19 bits of an
// fp32 word for kind::tf32, i.e. it sees hi = trunc_tf32(v).  The splitters therefore write only
// lo = round_tf32(v - trunc_tf32(v)) (the subtraction is exact; lo is rounded because the tensor core would otherwise
// truncate its low bits too, a one-sided error).  Dropped term lo_a*lo_b <= 2^-20 |a b|.  One shared-memory write per
// element instead of two: the main loop is bound by shared-memory traffic (TMA fill + split + MMA operand reads).
// Used by the NT products (K <= a few thousand); the TN product (K = number of nodes) keeps the round-to-nearest split.
__device__ __forceinline__ float lo1(float v) { return tf32_round(v - __uint_as_float(__float_as_uint(v) & 0xffffe000u)); }
__device__ __forceinline__ float4 lo4(float4 v) { return make_float4(lo1(v.x), lo1(v.y), lo1(v.z), lo1(v.w)); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D fp32 row-major (rows x cols, leading dimension ld elements); box = box_rows x 32 columns, SWIZZLE_128B,
// out-of-bounds elements read as zero (so M, N and K tails need no special casing).
enum MapKind { kMapK64, kMapC128, kMapMN };   // K-major operand (64-byte rows), output tile (128-byte rows), MN-major operand
static inline int make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, MapKind kind) {
  const int box_cols = kind == kMapK64 ? BK : 32;
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("gat_gemm: cuTensorMapEncodeTiled is unavailable"); return GAT_EUNSUPPORTED; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  kind == kMapMN ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : (kind == kMapK64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("gat_gemm: cuTensorMapEncodeTiled failed (%d)", (int)r); return GAT_EINVAL; }
  return GAT_OK;
}


// Destinations of an NT output tile.  The tile is staged in shared memory and written by TMA to `count` row-major
// matrices with the same geometry: the caller's C and, for the fused projection -> all-gather, the same slab of every
// peer GPU's gathered buffer (NVLink peer memory: the stores leave over the switch while the next CTAs compute).  The
// tile's rows land at row_offset + m0; each map's row extent is row_offset + M, so tail rows are clipped.
constexpr int kMaxDests = 8;
struct CStoreMaps {
  CUtensorMap maps[kMaxDests];
  int count;
  int row_offset;
};

}  // namespace tc
}  // namespace gat
