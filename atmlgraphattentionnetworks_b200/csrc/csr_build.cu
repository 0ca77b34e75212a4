// K0 — graph ingestion: COO int64 edge_index (+ appended self loops, GAT.py:38 -> [PyG] add_self_loops: no dedup)
// -> destination-sorted CSR and source-sorted CSC, int32.
//
// Canonical order = STABLE sort by destination of [edge_index ; (n,n) n<N]  (so inside a row the original edge
// order is kept and the appended self loop is last); CSC = STABLE sort of the CSR-ordered entries by source.
// Both sorts are LSD radix sorts (cub::DeviceRadixSort, stable) over ceil(log2 N) key bits; everything else is
// one coalesced streaming pass.  Every node owns a self loop, so every CSR row and CSC column is non-empty and
// rowptr / colptr fall out of a boundary scan of the sorted keys — no histogram, no atomics, bit-exact.
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>

namespace b200gat {

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int key_bits(int64_t n) {
  int b = 1;
  while ((int64_t(1) << b) < n) ++b;
  return b;
}

// keys[p] = destination, iota[p] = p for the E' = E + N entries of [edges ; loops]; counts out-of-range indices.
__global__ void csr_fill_keys(const int64_t* __restrict__ ei, int64_t E, int64_t N, int32_t* __restrict__ keys,
                              int32_t* __restrict__ iota, int32_t* __restrict__ status) {
  int64_t EP = E + N;
  int bad = 0, span = 0;
  for (int64_t p = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; p < EP; p += int64_t(gridDim.x) * blockDim.x) {
    int64_t d, s;
    if (p < E) { s = ei[p]; d = ei[E + p]; } else { s = d = p - E; }
    if (d < 0 || d >= N || s < 0 || s >= N) { ++bad; d = d < 0 ? 0 : (d >= N ? N - 1 : d); }
    else { const int64_t w = s > d ? s - d : d - s; span = w > span ? static_cast<int>(w) : span; }
    keys[p] = static_cast<int32_t>(d);
    iota[p] = static_cast<int32_t>(p);
  }
  if (bad) atomicAdd(status, bad);
  // status[1] = max |source - destination|: how far apart (in node rows) an edge's gather can be from the row that
  // is being processed — block-diagonal graph batches have a small span (their gathered operand stays L2-resident)
  span = __reduce_max_sync(0xffffffffu, span);
  if ((threadIdx.x & 31) == 0 && span) atomicMax(status + 1, span);
}

// After the destination sort: col[pos] = source of entry eid[pos]; rowptr from key boundaries.
__global__ void csr_finish_rows(const int64_t* __restrict__ ei, int64_t E, int64_t N,
                                const int32_t* __restrict__ row_of, const int32_t* __restrict__ eid,
                                int32_t* __restrict__ col, int32_t* __restrict__ rowptr) {
  int64_t EP = E + N;
  for (int64_t p = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; p < EP; p += int64_t(gridDim.x) * blockDim.x) {
    int32_t e = eid[p];
    int64_t s = e < E ? ei[e] : int64_t(e) - E;
    s = s < 0 ? 0 : (s >= N ? N - 1 : s);
    col[p] = static_cast<int32_t>(s);
    int32_t r = row_of[p];
    if (p == 0 || row_of[p - 1] != r) rowptr[r] = static_cast<int32_t>(p);
    if (p == EP - 1) rowptr[N] = static_cast<int32_t>(EP);
  }
}

// After the source sort: crow / ceid through the CSR position, colptr from key boundaries.
__global__ void csr_finish_cols(int64_t EP, int64_t N, const int32_t* __restrict__ src_sorted,
                                const int32_t* __restrict__ cpos, const int32_t* __restrict__ row_of,
                                const int32_t* __restrict__ eid, int32_t* __restrict__ crow,
                                int32_t* __restrict__ ceid, int32_t* __restrict__ colptr) {
  for (int64_t p = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; p < EP; p += int64_t(gridDim.x) * blockDim.x) {
    int32_t q = cpos[p];
    crow[p] = row_of[q];
    ceid[p] = eid[q];
    int32_t s = src_sorted[p];
    if (p == 0 || src_sorted[p - 1] != s) colptr[s] = static_cast<int32_t>(p);
    if (p == EP - 1) colptr[N] = static_cast<int32_t>(EP);
  }
}

// rows whose degree exceeds B200GAT_HUB_DEGREE (b200gat_graph.hub_rows / hub_cols)
__global__ void hub_list_kernel(const int32_t* __restrict__ ptr, int64_t n, int32_t* __restrict__ list, int64_t cap,
                                int32_t* __restrict__ count, int32_t* __restrict__ ends) {
  int maxdeg = 0;
  for (int64_t r = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; r < n; r += int64_t(gridDim.x) * blockDim.x) {
    const int32_t b = ptr[r], e = ptr[r + 1];
    const bool hub = e - b > B200GAT_HUB_DEGREE;
    ends[r] = hub ? b : e;
    maxdeg = e - b > maxdeg ? e - b : maxdeg;
    if (hub) {
      const int pos = atomicAdd(count, 1);
      if (pos < cap) list[pos] = static_cast<int32_t>(r);
    }
  }
  maxdeg = __reduce_max_sync(0xffffffffu, maxdeg);
  if ((threadIdx.x & 31) == 0) atomicMax(count + 1, maxdeg);
}

struct CsrWorkspace {
  size_t off_a, off_iota, off_row, off_pos, off_cub, cub_bytes, total;
};

static int plan_workspace(int64_t N, int64_t E, CsrWorkspace* w, bool query_cub) {
  int64_t EP = E + N;
  size_t arr = align_up(size_t(EP > 0 ? EP : 1) * sizeof(int32_t), 256);
  w->off_a = 0;
  w->off_iota = arr;
  w->off_row = 2 * arr;
  w->off_pos = 3 * arr;
  w->off_cub = 4 * arr;
  size_t cub_bytes = 0;
  if (query_cub && EP > 0) {
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const int32_t*)nullptr, (int32_t*)nullptr,
                                                    (const int32_t*)nullptr, (int32_t*)nullptr,
                                                    static_cast<int64_t>(EP), 0, key_bits(N));
    if (e != cudaSuccess) {
      set_error("csr workspace query: %s", cudaGetErrorString(e));
      return static_cast<int>(e);
    }
  }
  w->cub_bytes = align_up(cub_bytes + 256, 256);
  w->total = w->off_cub + w->cub_bytes;
  return 0;
}

}  // namespace b200gat

using namespace b200gat;

extern "C" size_t b200gat_csr_workspace_bytes(int64_t num_nodes, int64_t num_input_edges) {
  if (num_nodes < 0 || num_input_edges < 0) return 0;
  CsrWorkspace w;
  if (plan_workspace(num_nodes, num_input_edges, &w, true) != 0) return 0;
  return w.total;
}

extern "C" int b200gat_csr_build(const int64_t* edge_index, int64_t E, int64_t N, int32_t* rowptr, int32_t* col,
                                 int32_t* eid, int32_t* colptr, int32_t* crow, int32_t* ceid, int32_t* status,
                                 void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200GAT_REQUIRE(N >= 0 && E >= 0, B200GAT_E_SHAPE, "csr_build: negative size");
  B200GAT_REQUIRE(E + N < (int64_t(1) << 31), B200GAT_E_SHAPE, "csr_build: E + N must be < 2^31");
  B200GAT_REQUIRE(rowptr && colptr && status, B200GAT_E_NULL, "csr_build: NULL output");
  B200GAT_REQUIRE(E == 0 || edge_index, B200GAT_E_NULL, "csr_build: NULL edge_index");
  cudaError_t ce = cudaMemsetAsync(status, 0, 2 * sizeof(int32_t), stream);
  if (ce != cudaSuccess) return fail(static_cast<int>(ce), "csr_build: memset: %s", cudaGetErrorString(ce));
  int64_t EP = E + N;
  if (N == 0) {
    ce = cudaMemsetAsync(rowptr, 0, sizeof(int32_t), stream);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(colptr, 0, sizeof(int32_t), stream);
    if (ce != cudaSuccess) return fail(static_cast<int>(ce), "csr_build: memset: %s", cudaGetErrorString(ce));
    return 0;
  }
  B200GAT_REQUIRE(col && eid && crow && ceid && workspace, B200GAT_E_NULL, "csr_build: NULL array");
  CsrWorkspace w;
  int rc = plan_workspace(N, E, &w, true);
  if (rc) return rc;
  B200GAT_REQUIRE(workspace_bytes >= w.total, B200GAT_E_WORKSPACE, "csr_build: workspace %zu < %zu bytes",
                  workspace_bytes, w.total);
  B200GAT_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, B200GAT_E_ALIGN,
                  "csr_build: workspace must be 256-byte aligned");
  char* base = static_cast<char*>(workspace);
  int32_t* buf_a = reinterpret_cast<int32_t*>(base + w.off_a);      // dst keys, later: sorted source keys
  int32_t* iota = reinterpret_cast<int32_t*>(base + w.off_iota);
  int32_t* row_of = reinterpret_cast<int32_t*>(base + w.off_row);   // destination of each CSR entry
  int32_t* cpos = reinterpret_cast<int32_t*>(base + w.off_pos);     // CSR position of each CSC entry
  void* cub_tmp = base + w.off_cub;
  size_t cub_bytes = w.cub_bytes;
  const int bits = key_bits(N);
  const int threads = 256;
  const int blocks = static_cast<int>(ceil_div(EP, threads) < 148 * 16 ? ceil_div(EP, threads) : 148 * 16);

  csr_fill_keys<<<blocks, threads, 0, stream>>>(edge_index, E, N, buf_a, iota, status);
  if ((rc = check_launch("csr_fill_keys"))) return rc;
  ce = cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, (const int32_t*)buf_a, row_of, (const int32_t*)iota, eid,
                                       EP, 0, bits, stream);
  if (ce != cudaSuccess) return fail(static_cast<int>(ce), "csr_build: sort by destination: %s", cudaGetErrorString(ce));
  csr_finish_rows<<<blocks, threads, 0, stream>>>(edge_index, E, N, row_of, eid, col, rowptr);
  if ((rc = check_launch("csr_finish_rows"))) return rc;
  ce = cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, (const int32_t*)col, buf_a, (const int32_t*)iota, cpos,
                                       EP, 0, bits, stream);
  if (ce != cudaSuccess) return fail(static_cast<int>(ce), "csr_build: sort by source: %s", cudaGetErrorString(ce));
  csr_finish_cols<<<blocks, threads, 0, stream>>>(EP, N, buf_a, cpos, row_of, eid, crow, ceid, colptr);
  return check_launch("csr_finish_cols");
}

extern "C" int b200gat_hub_rows(const int32_t* ptr, int64_t num_rows, int32_t* list, int64_t cap, int32_t* count,
                                int32_t* ends, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  B200GAT_REQUIRE(num_rows >= 0 && cap >= 0, B200GAT_E_SHAPE, "hub_rows: negative size");
  B200GAT_REQUIRE(count && (num_rows == 0 || (ptr && ends)) && (cap == 0 || list), B200GAT_E_NULL, "hub_rows: NULL pointer");
  cudaError_t ce = cudaMemsetAsync(count, 0, 2 * sizeof(int32_t), stream);
  if (ce != cudaSuccess) return fail(static_cast<int>(ce), "hub_rows: memset: %s", cudaGetErrorString(ce));
  if (num_rows == 0) return 0;
  const int64_t want = ceil_div(num_rows, 256);
  hub_list_kernel<<<static_cast<int>(want < 148 * 16 ? want : 148 * 16), 256, 0, stream>>>(ptr, num_rows, list, cap, count, ends);
  return check_launch("hub_list_kernel");
}
