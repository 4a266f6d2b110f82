// kNN similarity-graph construction on device (data_loader.py:278-344, utils.py:11-27):
//   1. per-row top-k of a float64 similarity block under the tie rule (value desc, column asc)
//      -- replaces np.argpartition(-S, k)[:, :k] (data_loader.py:293);
//   2. neighbour lists -> A + A^T (entries 1/2), + I, D^-1 A in float64, cast to fp32, emitted as a
//      canonical (row, col)-sorted COO plus CSR indptr -- replaces the scipy.sparse pipeline at
//      data_loader.py:294-308 / utils.py:11-27. Integer work is exact; the normalised values match
//      numpy bit for bit (IEEE double reciprocal and product, then one rounding to fp32).
#include <math_constants.h>

#include "common.cuh"
#include "primitives.cuh"

namespace dg {

// ------------------------------------------------------------------------------------------------
// top-k: one warp per row, warp-distributed sorted list of up to 64 (value, column) entries
// (position p lives in lane p%32, slot p/32). Candidates are tested 32 at a time against the
// current k-th best; the few that pass are inserted one by one with a warp-parallel shift.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool better(double v, int c, double w, int wc) { return v > w || (v == w && c < wc); }

template <int SLOTS>
__global__ void __launch_bounds__(256) topk_rows_kernel(const double* __restrict__ sim, int64_t n_rows, int64_t n_cols,
                                                        int64_t ld, int k, int* __restrict__ out_idx) {
  const int lane = lane_id();
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_rows) return;
  const double* srow = sim + row * ld;
  double lv[SLOTS];
  int lc[SLOTS];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) { lv[s] = -CUDART_INF; lc[s] = 0x7fffffff; }
  const int kth_lane = (k - 1) & 31, kth_slot = (k - 1) >> 5;
  double thr_v = -CUDART_INF;
  int thr_c = 0x7fffffff;

  for (int64_t base = 0; base < n_cols; base += 32 * 4) {
    double cv[4];
    int cc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {            // 4 independent coalesced loads in flight per lane
      const int64_t c = base + u * 32 + lane;
      cc[u] = c < n_cols ? static_cast<int>(c) : 0x7fffffff;
      cv[u] = c < n_cols ? srow[c] : -CUDART_INF;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      unsigned pass = __ballot_sync(kFull, cc[u] != 0x7fffffff && better(cv[u], cc[u], thr_v, thr_c));
      while (pass) {
        const int srcl = __ffs(pass) - 1;
        pass &= pass - 1;
        const double v = __shfl_sync(kFull, cv[u], srcl);
        const int c = __shfl_sync(kFull, cc[u], srcl);
        if (!better(v, c, thr_v, thr_c)) continue;   // threshold may have risen since the ballot
        // new[p] = old[p] while old[p] is better than the candidate; the candidate takes the first
        // position that is not; everything after shifts down by one (last entry falls off).
#pragma unroll
        for (int s = SLOTS - 1; s >= 0; --s) {
          double pv = __shfl_up_sync(kFull, lv[s], 1);
          int pc = __shfl_up_sync(kFull, lc[s], 1);
          // lane 0 continues from lane 31 of the previous slot (s is a compile-time constant here, so
          // every lane executes the same shuffles); position 0 has an always-better sentinel before it
          double wv = CUDART_INF;
          int wc = -1;
          if (s > 0) { wv = __shfl_sync(kFull, lv[s > 0 ? s - 1 : 0], 31); wc = __shfl_sync(kFull, lc[s > 0 ? s - 1 : 0], 31); }
          if (lane == 0) { pv = wv; pc = wc; }
          const bool mine_better = better(lv[s], lc[s], v, c);
          const bool prev_better = better(pv, pc, v, c);
          if (!mine_better) {
            if (prev_better) { lv[s] = v; lc[s] = c; }
            else { lv[s] = pv; lc[s] = pc; }
          }
        }
        thr_v = __shfl_sync(kFull, (SLOTS > 1 && kth_slot) ? lv[SLOTS - 1] : lv[0], kth_lane);
        thr_c = __shfl_sync(kFull, (SLOTS > 1 && kth_slot) ? lc[SLOTS - 1] : lc[0], kth_lane);
      }
    }
  }
  // emit the k winners in ascending column order (rank = number of winners with a smaller column)
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int p = s * 32 + lane;
    int rank = 0;
#pragma unroll
    for (int s2 = 0; s2 < SLOTS; ++s2)
      for (int l2 = 0; l2 < 32; ++l2) {
        const int oc = __shfl_sync(kFull, lc[s2], l2);
        if (s2 * 32 + l2 < k && oc < lc[s]) ++rank;
      }
    if (p < k) out_idx[row * k + rank] = lc[s];
  }
}

// ------------------------------------------------------------------------------------------------
// neighbour lists -> symmetric normalised adjacency
// ------------------------------------------------------------------------------------------------
__global__ void knn_emit_keys(const int* __restrict__ nbr, int64_t n, int k, uint64_t* __restrict__ keys) {
  const int64_t nk = n * k;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < nk + n; t += stride) {
    if (t < nk) {
      const uint64_t i = static_cast<uint64_t>(t / k), j = static_cast<uint64_t>(nbr[t]);
      keys[2 * t] = i * n + j;          // A
      keys[2 * t + 1] = j * n + i;      // A^T
    } else {
      const uint64_t i = static_cast<uint64_t>(t - nk);
      keys[2 * nk + i] = i * n + i;     // + I
    }
  }
}

__global__ void knn_head_flags(const uint64_t* __restrict__ keys, int64_t m, int* __restrict__ flags) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t s = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; s < m; s += stride)
    flags[s] = (s == 0 || keys[s] != keys[s - 1]) ? 1 : 0;
}

__global__ void knn_emit_entries(const uint64_t* __restrict__ keys, const int* __restrict__ pos, int64_t m, int64_t n,
                                 int* __restrict__ row, int* __restrict__ col, int* __restrict__ count,
                                 int* __restrict__ rowsum, int* __restrict__ rowcnt) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t s = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; s < m; s += stride) {
    const uint64_t key = keys[s];
    if (s != 0 && keys[s - 1] == key) continue;
    int c = 1;
    while (s + c < m && keys[s + c] == key) ++c;      // multiplicity <= 3 (A, A^T, I)
    const int u = pos[s];
    const int r = static_cast<int>(key / static_cast<uint64_t>(n));
    row[u] = r;
    col[u] = static_cast<int>(key % static_cast<uint64_t>(n));
    count[u] = c;
    atomicAdd(&rowsum[r], c);     // integers: exact and order-independent
    atomicAdd(&rowcnt[r], 1);
  }
}

__global__ void knn_normalise(const int* __restrict__ row, const int* __restrict__ count, const int* __restrict__ rowsum,
                              const int* __restrict__ nnz, float* __restrict__ val, int* __restrict__ nnz_out) {
  const int total = *nnz;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t u = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; u < total; u += stride) {
    // utils.py:13-17: r_inv = rowsum ** -1 (float64), r_mat_inv.dot(mx) -> product in float64;
    // utils.py:22: astype(float32)
    const double r_inv = __drcp_rn(static_cast<double>(rowsum[row[u]]));
    val[u] = __double2float_rn(__dmul_rn(r_inv, static_cast<double>(count[u])));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *nnz_out = total;
}

static unsigned knn_grid(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = static_cast<int64_t>(kNumSM) * 16;
  if (b > cap) b = cap;
  return static_cast<unsigned>(b < 1 ? 1 : b);
}

static int bit_length64(uint64_t x) { int b = 0; while (x) { ++b; x >>= 1; } return b; }

size_t knn_graph_workspace_bytes(int64_t n, int k) {
  const size_t m = static_cast<size_t>(2 * n * k + n);
  size_t b = 0;
  b = ws_add(b, m * sizeof(uint64_t));
  b = ws_add(b, m * sizeof(uint64_t));
  b = ws_add(b, (m + 1) * sizeof(int));     // flags / positions
  b = ws_add(b, m * sizeof(int));           // multiplicities
  b = ws_add(b, static_cast<size_t>(n + 1) * sizeof(int) * 2);   // rowsum | rowcnt
  b = ws_add(b, sort_workspace_bytes(static_cast<int64_t>(m)));
  b = ws_add(b, scan_workspace_bytes(static_cast<int64_t>(m)));
  return b;
}

}  // namespace dg

extern "C" {

int dg_topk_rows_f64(const double* sim, int64_t n_rows, int64_t n_cols, int64_t ld, int k, int32_t* out_idx,
                     dg_stream_t stream) {
  using namespace dg;
  DG_REQUIRE(n_rows >= 0 && n_cols >= 0 && ld >= n_cols, "bad shape");
  DG_REQUIRE(k >= 1 && k <= 64 && k <= n_cols, "k must be in [1, min(64, n_cols)]");
  DG_REQUIRE(n_cols <= 0x7fffffffLL, "n_cols out of int32 range");
  if (n_rows == 0) return DG_OK;
  const unsigned blocks = static_cast<unsigned>((n_rows + 7) / 8);
  if (k <= 32) topk_rows_kernel<1><<<blocks, 256, 0, as_stream(stream)>>>(sim, n_rows, n_cols, ld, k, out_idx);
  else topk_rows_kernel<2><<<blocks, 256, 0, as_stream(stream)>>>(sim, n_rows, n_cols, ld, k, out_idx);
  DG_CHECK_LAUNCH("topk_rows");
  return DG_OK;
}

size_t dg_knn_graph_workspace_bytes(int64_t n, int k) { return dg::knn_graph_workspace_bytes(n, k); }

int dg_knn_graph_from_neighbors(const int32_t* nbr, int64_t n, int k, int32_t* indptr, int32_t* row, int32_t* col,
                                float* val, int32_t* nnz_out, void* workspace, size_t workspace_bytes,
                                dg_stream_t stream) {
  using namespace dg;
  DG_REQUIRE(n >= 1 && k >= 1, "n and k must be positive");
  DG_REQUIRE(n <= 0x7fffffffLL && (2 * n * k + n) <= 0x7fffffffLL, "graph too large for int32 positions");
  cudaStream_t st = as_stream(stream);
  const int64_t m = 2 * n * k + n;
  Workspace w(workspace, workspace_bytes);
  uint64_t* keys_a = w.take<uint64_t>(m);
  uint64_t* keys_b = w.take<uint64_t>(m);
  int* pos = w.take<int>(m + 1);
  int* count = w.take<int>(m);
  int* rowsum = w.take<int>(2 * (n + 1));
  const size_t sort_bytes = sort_workspace_bytes(m);
  char* sort_ws = w.take<char>(sort_bytes);
  const size_t scan_bytes = scan_workspace_bytes(m);
  char* scan_ws = w.take<char>(scan_bytes);
  if (!keys_a || !keys_b || !pos || !count || !rowsum || !sort_ws || !scan_ws) {
    set_error("knn_graph: workspace too small");
    return DG_ERR_WORKSPACE_TOO_SMALL;
  }
  int* rowcnt = rowsum + (n + 1);
  DG_CHECK_CUDA(cudaMemsetAsync(rowsum, 0, static_cast<size_t>(2 * (n + 1)) * sizeof(int), st));
  knn_emit_keys<<<knn_grid(n * k + n), 256, 0, st>>>(nbr, n, k, keys_a);
  DG_CHECK_LAUNCH("knn_emit_keys");
  const int bits = bit_length64(static_cast<uint64_t>(n) * static_cast<uint64_t>(n) - 1);
  DG_PROPAGATE(sort_pairs_u64(keys_a, nullptr, keys_b, nullptr, m, bits < 1 ? 1 : bits, sort_ws, sort_bytes, st));
  knn_head_flags<<<knn_grid(m), 256, 0, st>>>(keys_b, m, pos);
  DG_CHECK_LAUNCH("knn_head_flags");
  DG_PROPAGATE(exclusive_scan_i32(pos, pos, m, scan_ws, scan_bytes, st));
  knn_emit_entries<<<knn_grid(m), 256, 0, st>>>(keys_b, pos, m, n, row, col, count, rowsum, rowcnt);
  DG_CHECK_LAUNCH("knn_emit_entries");
  DG_PROPAGATE(exclusive_scan_i32(rowcnt, indptr, n, scan_ws, scan_bytes, st));
  knn_normalise<<<knn_grid(m), 256, 0, st>>>(row, count, rowsum, pos + m, val, nnz_out);
  DG_CHECK_LAUNCH("knn_normalise");
  return DG_OK;
}
}
