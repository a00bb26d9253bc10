// tcgen05 / TMA 3xTF32 GEMM (algo 2).  Placeholder until the kernel lands: reports "unsupported" so
// that algo 0 (auto) routes everything to the exact fp32 FFMA path.
#include "common.cuh"

namespace gat {

bool tc_supported(int, int, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t) { return false; }
size_t tc_workspace_bytes(int, int, int64_t, int64_t, int64_t) { return 0; }
int gemm_tc(int, int, int64_t, int64_t, int64_t, const float*, int64_t, const float*, int64_t, float*, int64_t,
            void*, size_t, cudaStream_t) {
  set_error("gat_gemm: tcgen05 path not built");
  return GAT_EUNSUPPORTED;
}

}  // namespace gat
