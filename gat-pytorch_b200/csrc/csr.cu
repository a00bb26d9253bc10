// Kernel 1: edge_index -> rewritten edge list, destination-sorted CSR and transposed CSR.
//
// Integer-only.  Reproduces utils.py:47-72 exactly: N_idx = max+1, every (i,i) dropped, remaining
// edges in input order, loops 0..N_idx-1 appended; then groups edges by target (stable, so the
// order inside a row is the reference's edge order, which is also the order in which the CPU
// scatter_add_ of utils.py:20 visits them) and by source.
//
// The two stable key sorts use CUB's LSD radix sort (header-only, ships with the toolkit) over
// ceil(log2 N) key bits; everything else is hand-written.  The structure is cached per graph by the
// host layer, so this is off the per-step path.
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace gat {

template <typename T>
__global__ void edges_scan_kernel(const T* __restrict__ src, const T* __restrict__ dst, int64_t n_edges,
                                  unsigned long long* __restrict__ stats /* max+1, keep, min(neg flag) */) {
  long long vmax = -1, vmin = 0;
  unsigned long long keep = 0;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_edges; e += (int64_t)gridDim.x * blockDim.x) {
    long long s = (long long)src[e], d = (long long)dst[e];
    vmax = max(vmax, max(s, d));
    vmin = min(vmin, min(s, d));
    keep += (s != d);
  }
  for (int o = 16; o > 0; o >>= 1) {
    vmax = max(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    vmin = min(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    keep += __shfl_xor_sync(0xffffffffu, keep, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax((long long*)&stats[0], vmax + 1);
    atomicAdd(&stats[1], keep);
    atomicMin((long long*)&stats[2], vmin);
  }
}

template <typename T>
__global__ void keep_flags_kernel(const T* __restrict__ src, const T* __restrict__ dst, int64_t n_edges,
                                  int* __restrict__ flags) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e < n_edges) flags[e] = src[e] != dst[e];
}

// Writes the rewritten list: kept edges at their scanned position, then the appended loops.
template <typename T>
__global__ void rewrite_kernel(const T* __restrict__ src, const T* __restrict__ dst, int64_t n_edges,
                               const int* __restrict__ pos, int add_self_loops, int64_t n_keep, int64_t n_idx,
                               int32_t* __restrict__ src32, int32_t* __restrict__ dst32,
                               int64_t* __restrict__ ei_out, int64_t n_edges_out) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t < n_edges) {
    long long s = (long long)src[t], d = (long long)dst[t];
    if (!add_self_loops || s != d) {
      int64_t p = add_self_loops ? (int64_t)pos[t] : t;
      src32[p] = (int32_t)s;
      dst32[p] = (int32_t)d;
      if (ei_out) { ei_out[p] = s; ei_out[n_edges_out + p] = d; }
    }
  }
  if (add_self_loops && t < n_idx) {
    int64_t p = n_keep + t;
    src32[p] = (int32_t)t;
    dst32[p] = (int32_t)t;
    if (ei_out) { ei_out[p] = t; ei_out[n_edges_out + p] = t; }
  }
}

__global__ void iota_kernel(int32_t* __restrict__ v, int64_t n) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t < n) v[t] = (int32_t)t;
}

// rowptr[r] = first slot j with sorted_key[j] >= r, for r in [0, n_nodes].
__global__ void rowptr_kernel(const int32_t* __restrict__ sorted_key, int64_t n_edges, int64_t n_nodes,
                              int32_t* __restrict__ rowptr) {
  int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j > n_edges) return;
  int64_t prev = (j == 0) ? -1 : sorted_key[j - 1];
  int64_t cur = (j == n_edges) ? n_nodes : sorted_key[j];
  for (int64_t r = prev + 1; r <= cur; ++r) rowptr[r] = (int32_t)j;
}

// CSR by target: col[j] = src[eid[j]]; slot_of_edge[eid[j]] = j.
__global__ void gather_csr_kernel(const int32_t* __restrict__ eid, const int32_t* __restrict__ src32, int64_t n_edges,
                                  int32_t* __restrict__ col, int32_t* __restrict__ slot_of_edge) {
  int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j < n_edges) {
    int32_t e = eid[j];
    col[j] = src32[e];
    slot_of_edge[e] = (int32_t)j;
  }
}

// CSR by source: col_t[j] = dst[perm_t[j]]; pos_t[j] = slot_of_edge[perm_t[j]].
__global__ void gather_csrt_kernel(const int32_t* __restrict__ perm_t, const int32_t* __restrict__ dst32,
                                   const int32_t* __restrict__ slot_of_edge, int64_t n_edges,
                                   int32_t* __restrict__ col_t, int32_t* __restrict__ pos_t, int32_t* __restrict__ tpos) {
  int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j < n_edges) {
    int32_t e = perm_t[j];
    col_t[j] = dst32[e];
    int32_t slot = slot_of_edge[e];
    pos_t[j] = slot;
    if (tpos) tpos[slot] = (int32_t)j;
  }
}

// Scheduling order for the persistent edge kernels: rows longer than `thresh` first, LONGEST first among them (the
// cooperative CTA-per-row phase then starts its biggest jobs at t = 0), then the rest in row order.  One stable
// descending radix sort on key = (degree > thresh ? degree : 0).
__global__ void order_key_kernel(const int32_t* __restrict__ rowptr, int64_t n, int thresh, int32_t* __restrict__ keys,
                                 int32_t* __restrict__ rows, int64_t* __restrict__ n_long_out) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int is_long = 0;
  if (r < n) {
    const int deg = rowptr[r + 1] - rowptr[r];
    is_long = deg > thresh;
    keys[r] = is_long ? deg : 0;
    rows[r] = (int32_t)r;
  }
  const unsigned ballot = __ballot_sync(0xffffffffu, is_long);
  if ((threadIdx.x & 31) == 0 && ballot && n_long_out) atomicAdd((unsigned long long*)n_long_out, (unsigned long long)__popc(ballot));
}

constexpr int kLongRowThreshold = GAT_LONG_ROW_EDGES;

static int key_bits(int64_t n_nodes) {
  int b = 1;
  while (((int64_t)1 << b) < n_nodes) ++b;
  return b;
}

struct CsrWorkspace {
  size_t off_src32, off_dst32, off_keys_out, off_iota, off_perm, off_slot, off_flags, off_pos, off_rows, off_cub, cub_bytes, total;
};
static inline int64_t max64(int64_t a, int64_t b) { return a > b ? a : b; }

static int plan_workspace(int64_t e_in, int64_t e_out, int64_t n_nodes, CsrWorkspace* w) {
  size_t sort_bytes = 0, scan_bytes = 0;
  cudaError_t e1 = cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const int32_t*)nullptr, (int32_t*)nullptr,
                                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)e_out, 0, key_bits(n_nodes));
  cudaError_t e2 = cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const int*)nullptr, (int*)nullptr, (int)max64(e_in, n_nodes));
  size_t order_bytes = 0;
  cudaError_t e3 = cub::DeviceRadixSort::SortPairsDescending(nullptr, order_bytes, (const int32_t*)nullptr, (int32_t*)nullptr,
                                                            (const int32_t*)nullptr, (int32_t*)nullptr, (int)max64(n_nodes, 1), 0, 31);
  if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) return (int)(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3));
  if (order_bytes > sort_bytes) sort_bytes = order_bytes;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 256); return r; };
  size_t eb = (size_t)(e_out > 0 ? e_out : 1) * sizeof(int32_t);
  size_t ib = (size_t)(max64(max64(e_in, n_nodes), 1)) * sizeof(int);   // flags/pos also serve the row-order scan
  w->off_src32 = take(eb); w->off_dst32 = take(eb); w->off_keys_out = take(eb); w->off_iota = take(eb);
  w->off_perm = take(eb); w->off_slot = take(eb); w->off_flags = take(ib); w->off_pos = take(ib); w->off_rows = take(ib);
  w->cub_bytes = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
  w->off_cub = take(w->cub_bytes + 256);
  w->total = o;
  return 0;
}

template <typename T>
static int csr_build_impl(const T* src, const T* dst, int64_t e_in, int add_self_loops, int64_t n_idx, int64_t e_out,
                          int64_t n_nodes, int64_t* ei_out, int32_t* rowptr, int32_t* col, int32_t* eid,
                          int32_t* rowptr_t, int32_t* col_t, int32_t* pos_t, int32_t* tpos, int32_t* row_order, int32_t* row_order_t,
                          int64_t* n_long, char* ws, const CsrWorkspace& w, cudaStream_t st) {
  int32_t* src32 = (int32_t*)(ws + w.off_src32);
  int32_t* dst32 = (int32_t*)(ws + w.off_dst32);
  int32_t* keys_out = (int32_t*)(ws + w.off_keys_out);
  int32_t* iota = (int32_t*)(ws + w.off_iota);
  int32_t* perm = (int32_t*)(ws + w.off_perm);
  int32_t* slot = (int32_t*)(ws + w.off_slot);
  int* flags = (int*)(ws + w.off_flags);
  int* pos = (int*)(ws + w.off_pos);
  int32_t* rows = (int32_t*)(ws + w.off_rows);
  void* cub_tmp = (void*)(ws + w.off_cub);
  size_t cub_bytes = w.cub_bytes;
  const int T256 = 256;
  auto blocks = [&](int64_t n) { return (unsigned)((n + T256 - 1) / T256 > 0 ? (n + T256 - 1) / T256 : 1); };
  int64_t n_keep = e_out - (add_self_loops ? n_idx : 0);
  if (add_self_loops && e_in > 0) {
    keep_flags_kernel<T><<<blocks(e_in), T256, 0, st>>>(src, dst, e_in, flags);
    GAT_LAUNCH_CHECK();
    GAT_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, flags, pos, (int)e_in, st));
  }
  int64_t span = e_in > n_idx ? e_in : n_idx;
  if (span > 0) {
    rewrite_kernel<T><<<blocks(span), T256, 0, st>>>(src, dst, e_in, pos, add_self_loops, n_keep,
                                                     add_self_loops ? n_idx : 0, src32, dst32, ei_out, e_out);
    GAT_LAUNCH_CHECK();
  }
  int bits = key_bits(n_nodes);
  if (e_out > 0) {
    iota_kernel<<<blocks(e_out), T256, 0, st>>>(iota, e_out);
    GAT_LAUNCH_CHECK();
    // by target
    GAT_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, dst32, keys_out, iota, eid, (int)e_out, 0, bits, st));
  }
  rowptr_kernel<<<blocks(e_out + 1), T256, 0, st>>>(keys_out, e_out, n_nodes, rowptr);
  GAT_LAUNCH_CHECK();
  if (e_out > 0) {
    gather_csr_kernel<<<blocks(e_out), T256, 0, st>>>(eid, src32, e_out, col, slot);
    GAT_LAUNCH_CHECK();
    // by source
    GAT_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, src32, keys_out, iota, perm, (int)e_out, 0, bits, st));
  }
  rowptr_kernel<<<blocks(e_out + 1), T256, 0, st>>>(keys_out, e_out, n_nodes, rowptr_t);
  GAT_LAUNCH_CHECK();
  if (e_out > 0) {
    gather_csrt_kernel<<<blocks(e_out), T256, 0, st>>>(perm, dst32, slot, e_out, col_t, pos_t, tpos);
    GAT_LAUNCH_CHECK();
  }
  if (n_nodes > 0) {
    const int32_t* rps[2] = {rowptr, rowptr_t};
    int32_t* outs[2] = {row_order, row_order_t};
    for (int i = 0; i < 2; ++i) {
      if (!outs[i]) continue;
      order_key_kernel<<<blocks(n_nodes), T256, 0, st>>>(rps[i], n_nodes, kLongRowThreshold, flags, rows, n_long ? n_long + i : nullptr);
      GAT_LAUNCH_CHECK();
      GAT_CUDA(cub::DeviceRadixSort::SortPairsDescending(cub_tmp, cub_bytes, (const int32_t*)flags, (int32_t*)pos, (const int32_t*)rows,
                                                         outs[i], (int)n_nodes, 0, 31, st));
    }
  }
  return GAT_OK;
}

}  // namespace gat

extern "C" int gat_edges_scan(const void* edge_index, int64_t n_edges, int64_t row_stride, int index_is_int64,
                              int64_t* d_stats, gat_stream_t stream) {
  using namespace gat;
  GAT_CHECK_ARG(n_edges >= 0 && d_stats != nullptr, "gat_edges_scan: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const long long init[3] = {0, 0, 0};
  GAT_CUDA(cudaMemcpyAsync(d_stats, init, sizeof(init), cudaMemcpyHostToDevice, st));
  if (n_edges == 0) return GAT_OK;
  int blocks = (int)((n_edges + 1023) / 1024);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  if (index_is_int64) {
    const int64_t* p = (const int64_t*)edge_index;
    edges_scan_kernel<int64_t><<<blocks, 256, 0, st>>>(p, p + row_stride, n_edges, (unsigned long long*)d_stats);
  } else {
    const int32_t* p = (const int32_t*)edge_index;
    edges_scan_kernel<int32_t><<<blocks, 256, 0, st>>>(p, p + row_stride, n_edges, (unsigned long long*)d_stats);
  }
  GAT_LAUNCH_CHECK();
  return GAT_OK;
}

extern "C" size_t gat_csr_workspace_bytes(int64_t n_edges_in, int64_t n_edges_out, int64_t n_nodes) {
  gat::CsrWorkspace w;
  if (gat::plan_workspace(n_edges_in, n_edges_out, n_nodes, &w) != 0) return 0;
  return w.total;
}

extern "C" int gat_csr_build(const void* edge_index, int64_t n_edges_in, int64_t row_stride, int index_is_int64,
                             int add_self_loops, int64_t n_idx, int64_t n_edges_out, int64_t n_nodes,
                             int64_t* ei_out, int32_t* rowptr, int32_t* col, int32_t* eid,
                             int32_t* rowptr_t, int32_t* col_t, int32_t* pos_t, int32_t* tpos,
                             int32_t* row_order, int32_t* row_order_t, int64_t* n_long,
                             void* workspace, size_t workspace_bytes, gat_stream_t stream) {
  using namespace gat;
  GAT_CHECK_ARG(n_edges_in >= 0 && n_edges_out >= 0 && n_nodes >= 0, "gat_csr_build: negative size");
  GAT_CHECK_ARG(n_nodes < (int64_t)2147483647 && n_edges_out < (int64_t)2147483647 && n_edges_in < (int64_t)2147483647,
                "gat_csr_build: graph exceeds int32 indexing");
  GAT_CHECK_ARG(n_idx <= n_nodes, "gat_csr_build: edge_index refers to node %lld but x has %lld rows",
                (long long)n_idx - 1, (long long)n_nodes);
  GAT_CHECK_ARG(add_self_loops ? n_edges_out >= n_idx && n_edges_out - n_idx <= n_edges_in : n_edges_out == n_edges_in,
                "gat_csr_build: inconsistent n_edges_out");
  CsrWorkspace w;
  int rc = plan_workspace(n_edges_in, n_edges_out, n_nodes, &w);
  if (rc != 0) { set_error("gat_csr_build: workspace planning failed (%d)", rc); return rc; }
  if (workspace_bytes < w.total || workspace == nullptr) {
    set_error("gat_csr_build: workspace too small (%zu < %zu)", workspace_bytes, w.total);
    return GAT_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (n_long) GAT_CUDA(cudaMemsetAsync(n_long, 0, 2 * sizeof(int64_t), st));
  if (index_is_int64) {
    const int64_t* p = (const int64_t*)edge_index;
    return csr_build_impl<int64_t>(p, p + row_stride, n_edges_in, add_self_loops, n_idx, n_edges_out, n_nodes, ei_out,
                                   rowptr, col, eid, rowptr_t, col_t, pos_t, tpos, row_order, row_order_t, n_long, (char*)workspace, w, st);
  }
  const int32_t* p = (const int32_t*)edge_index;
  return csr_build_impl<int32_t>(p, p + row_stride, n_edges_in, add_self_loops, n_idx, n_edges_out, n_nodes, ei_out,
                                 rowptr, col, eid, rowptr_t, col_t, pos_t, tpos, row_order, row_order_t, n_long, (char*)workspace, w, st);
}
