// On-device graph construction: COO -> canonical CSR (stable radix sort on (row, col)), degree
// normalisers, and the per-iteration edge-dropout rebuild as an order-preserving CSR compaction.
//
// Reference behaviour replaced: DGL's lazy COO->CSC / COO->CSR inside update_all (layers.py:229-232),
// in_degrees / out_degrees + _calc_norm (data_loader.py:454-488), the heterograph rebuild after
// th.randperm edge dropout (augmentation.py:35-65) and sparse-COO dropout (augmentation.py:107-124).
#include "common.cuh"
#include "primitives.cuh"

namespace dg {

static int bit_length(uint64_t x) {
  int b = 0;
  while (x) { ++b; x >>= 1; }
  return b;
}

__global__ void csr_make_keys(const int* __restrict__ row, const int* __restrict__ col, int64_t n, int col_bits,
                              uint64_t* __restrict__ keys, int* __restrict__ ids, int* __restrict__ row_hist) {
  int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < n; e += stride) {
    int r = row[e];
    keys[e] = (static_cast<uint64_t>(static_cast<uint32_t>(r)) << col_bits) | static_cast<uint32_t>(col[e]);
    ids[e] = static_cast<int>(e);
    atomicAdd(&row_hist[r], 1);   // integer counts: order-independent result
  }
}

__global__ void csr_extract_cols(const uint64_t* __restrict__ keys, int64_t n, int col_bits, int* __restrict__ indices) {
  const uint64_t mask = (col_bits >= 64) ? ~0ull : ((1ull << col_bits) - 1);
  int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < n; e += stride)
    indices[e] = static_cast<int>(keys[e] & mask);
}

static unsigned grid_for(int64_t n, int threads, int max_waves = 8) {
  int64_t blocks = (n + threads - 1) / threads;
  int64_t cap = static_cast<int64_t>(kNumSM) * max_waves * (2048 / threads);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<unsigned>(blocks);
}

size_t csr_build_workspace_bytes(int64_t n_edges, int64_t n_rows) {
  size_t e = static_cast<size_t>(n_edges > 0 ? n_edges : 1);
  size_t b = 0;
  b = ws_add(b, e * sizeof(uint64_t));       // keys a
  b = ws_add(b, e * sizeof(uint64_t));       // keys b
  b = ws_add(b, e * sizeof(int));            // ids a
  b = ws_add(b, static_cast<size_t>(n_rows + 1) * sizeof(int));   // row histogram
  b = ws_add(b, sort_workspace_bytes(n_edges));
  b = ws_add(b, scan_workspace_bytes(n_rows));
  return b;
}

int csr_build(const int32_t* row, const int32_t* col, int64_t n_edges, int64_t n_rows, int64_t n_cols,
              int32_t* indptr, int32_t* indices, int32_t* eid, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (n_edges < 0 || n_rows < 0 || n_cols < 0 || n_edges > 0x7fffffffLL || n_rows > 0x7fffffffLL || n_cols > 0x7fffffffLL) {
    set_error("csr_build: sizes out of int32 range");
    return DG_ERR_INVALID_ARGUMENT;
  }
  size_t e = static_cast<size_t>(n_edges > 0 ? n_edges : 1);
  Workspace w(ws, ws_bytes);
  uint64_t* keys_a = w.take<uint64_t>(e);
  uint64_t* keys_b = w.take<uint64_t>(e);
  int* ids_a = w.take<int>(e);
  int* hist = w.take<int>(n_rows + 1);
  size_t sort_bytes = sort_workspace_bytes(n_edges);
  char* sort_ws = w.take<char>(sort_bytes);
  size_t scan_bytes = scan_workspace_bytes(n_rows);
  char* scan_ws = w.take<char>(scan_bytes);
  if (!keys_a || !keys_b || !ids_a || !hist || !sort_ws || !scan_ws) {
    set_error("csr_build: workspace too small (%zu bytes given)", ws_bytes);
    return DG_ERR_WORKSPACE_TOO_SMALL;
  }
  DG_CHECK_CUDA(cudaMemsetAsync(hist, 0, static_cast<size_t>(n_rows + 1) * sizeof(int), st));
  int col_bits = bit_length(n_cols > 0 ? static_cast<uint64_t>(n_cols - 1) : 0);
  int row_bits = bit_length(n_rows > 0 ? static_cast<uint64_t>(n_rows - 1) : 0);
  if (col_bits < 1) col_bits = 1;
  if (n_edges > 0) {
    csr_make_keys<<<grid_for(n_edges, 256), 256, 0, st>>>(row, col, n_edges, col_bits, keys_a, ids_a, hist);
    DG_CHECK_LAUNCH("csr_make_keys");
    DG_PROPAGATE(sort_pairs_u64(keys_a, ids_a, keys_b, eid, n_edges, col_bits + row_bits, sort_ws, sort_bytes, st));
    csr_extract_cols<<<grid_for(n_edges, 256), 256, 0, st>>>(keys_b, n_edges, col_bits, indices);
    DG_CHECK_LAUNCH("csr_extract_cols");
  }
  DG_PROPAGATE(exclusive_scan_i32(hist, indptr, n_rows, scan_ws, scan_bytes, st));
  return DG_OK;
}

// ------------------------------------------------------------------------------------------------
__global__ void degree_norm_kernel(const int* __restrict__ indptr, int64_t n, float* __restrict__ norm) {
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int deg = indptr[i + 1] - indptr[i];
  // data_loader.py:454-457: float32 degree, 0 -> inf, 1/sqrt -- IEEE sqrt and divide, no rsqrt approx
  norm[i] = deg == 0 ? 0.0f : __fdiv_rn(1.0f, __fsqrt_rn(static_cast<float>(deg)));
}

// ------------------------------------------------------------------------------------------------
__global__ void keep_flags_kernel(const int64_t* __restrict__ perm, int64_t num_keep, int64_t offset,
                                  uint8_t* __restrict__ flags) {
  int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < num_keep; i += stride)
    flags[offset + perm[i]] = 1;
}

// one warp per row: number of kept slots
__global__ void __launch_bounds__(256) compact_count(const int* __restrict__ indptr, const int* __restrict__ eid,
                                                     const uint8_t* __restrict__ keep, int64_t n_rows,
                                                     int* __restrict__ counts) {
  const int lane = lane_id();
  int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = warp; r < n_rows; r += n_warps) {
    int beg = indptr[r], end = indptr[r + 1];
    int c = 0;
    for (int s = beg + lane; s < end; s += 32) c += keep[eid[s]] ? 1 : 0;
#pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(kFull, c, o);
    if (lane == 0) counts[r] = c;
  }
}

template <bool kHasVals>
__global__ void __launch_bounds__(256) compact_write(const int* __restrict__ indptr, const int* __restrict__ indices,
                                                     const int* __restrict__ eid, const float* __restrict__ vals,
                                                     const uint8_t* __restrict__ keep, int64_t n_rows,
                                                     const int* __restrict__ out_indptr, int* __restrict__ out_indices,
                                                     int* __restrict__ out_eid, float* __restrict__ out_vals) {
  const int lane = lane_id();
  int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = warp; r < n_rows; r += n_warps) {
    int beg = indptr[r], end = indptr[r + 1];
    int run = out_indptr[r];
    for (int base = beg; base < end; base += 32) {
      int s = base + lane;
      int e = 0;
      bool k = false;
      if (s < end) { e = eid[s]; k = keep[e] != 0; }
      unsigned m = __ballot_sync(kFull, k);
      if (k) {
        int pos = run + __popc(m & ((1u << lane) - 1));
        out_indices[pos] = indices[s];
        out_eid[pos] = e;
        if (kHasVals) out_vals[pos] = vals[s];
      }
      run += __popc(m);
    }
  }
}

size_t csr_compact_workspace_bytes(int64_t n_rows) {
  size_t b = ws_add(0, static_cast<size_t>(n_rows + 1) * sizeof(int));
  return ws_add(b, scan_workspace_bytes(n_rows));
}

int csr_compact(const int32_t* indptr, const int32_t* indices, const int32_t* eid, const float* vals, int64_t n_rows,
                const uint8_t* keep, int32_t* out_indptr, int32_t* out_indices, int32_t* out_eid, float* out_vals,
                void* ws, size_t ws_bytes, cudaStream_t st) {
  if (n_rows < 0) { set_error("csr_compact: n_rows < 0"); return DG_ERR_INVALID_ARGUMENT; }
  if ((vals == nullptr) != (out_vals == nullptr)) { set_error("csr_compact: vals/out_vals mismatch"); return DG_ERR_INVALID_ARGUMENT; }
  Workspace w(ws, ws_bytes);
  int* counts = w.take<int>(n_rows + 1);
  size_t scan_bytes = scan_workspace_bytes(n_rows);
  char* scan_ws = w.take<char>(scan_bytes);
  if (!counts || !scan_ws) { set_error("csr_compact: workspace too small"); return DG_ERR_WORKSPACE_TOO_SMALL; }
  if (n_rows > 0) {
    unsigned grid = grid_for(n_rows * 32, 256);
    compact_count<<<grid, 256, 0, st>>>(indptr, eid, keep, n_rows, counts);
    DG_CHECK_LAUNCH("compact_count");
    DG_PROPAGATE(exclusive_scan_i32(counts, out_indptr, n_rows, scan_ws, scan_bytes, st));
    if (vals)
      compact_write<true><<<grid, 256, 0, st>>>(indptr, indices, eid, vals, keep, n_rows, out_indptr, out_indices, out_eid, out_vals);
    else
      compact_write<false><<<grid, 256, 0, st>>>(indptr, indices, eid, nullptr, keep, n_rows, out_indptr, out_indices, out_eid, nullptr);
    DG_CHECK_LAUNCH("compact_write");
  } else {
    DG_PROPAGATE(exclusive_scan_i32(counts, out_indptr, 0, scan_ws, scan_bytes, st));
  }
  return DG_OK;
}

__global__ void __launch_bounds__(256) expand_rows_kernel(const int* __restrict__ indptr, int64_t n_rows, int* __restrict__ row) {
  const int lane = lane_id();
  int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = warp; r < n_rows; r += n_warps) {
    int beg = indptr[r], end = indptr[r + 1];
    for (int s = beg + lane; s < end; s += 32) row[s] = static_cast<int>(r);
  }
}

}  // namespace dg

extern "C" {
size_t dg_csr_build_workspace_bytes(int64_t n_edges, int64_t n_rows) { return dg::csr_build_workspace_bytes(n_edges, n_rows); }
int dg_csr_build(const int32_t* row, const int32_t* col, int64_t n_edges, int64_t n_rows, int64_t n_cols,
                 int32_t* indptr, int32_t* indices, int32_t* eid, void* workspace, size_t workspace_bytes,
                 dg_stream_t stream) {
  return dg::csr_build(row, col, n_edges, n_rows, n_cols, indptr, indices, eid, workspace, workspace_bytes,
                       dg::as_stream(stream));
}
int dg_degree_norm(const int32_t* indptr, int64_t n_rows, float* norm, dg_stream_t stream) {
  DG_REQUIRE(n_rows >= 0, "n_rows < 0");
  if (n_rows == 0) return DG_OK;
  dg::degree_norm_kernel<<<static_cast<unsigned>((n_rows + 255) / 256), 256, 0, dg::as_stream(stream)>>>(indptr, n_rows, norm);
  DG_CHECK_LAUNCH("degree_norm");
  return DG_OK;
}
int dg_keep_flags_from_perm(const int64_t* perm, int64_t num_keep, int64_t offset, uint8_t* flags, dg_stream_t stream) {
  DG_REQUIRE(num_keep >= 0 && offset >= 0, "negative size");
  if (num_keep == 0) return DG_OK;
  dg::keep_flags_kernel<<<dg::grid_for(num_keep, 256), 256, 0, dg::as_stream(stream)>>>(perm, num_keep, offset, flags);
  DG_CHECK_LAUNCH("keep_flags");
  return DG_OK;
}
size_t dg_csr_compact_workspace_bytes(int64_t n_rows) { return dg::csr_compact_workspace_bytes(n_rows); }
int dg_csr_compact(const int32_t* indptr, const int32_t* indices, const int32_t* eid, const float* vals,
                   int64_t n_rows, const uint8_t* keep_by_eid, int32_t* out_indptr, int32_t* out_indices,
                   int32_t* out_eid, float* out_vals, void* workspace, size_t workspace_bytes, dg_stream_t stream) {
  return dg::csr_compact(indptr, indices, eid, vals, n_rows, keep_by_eid, out_indptr, out_indices, out_eid, out_vals,
                         workspace, workspace_bytes, dg::as_stream(stream));
}
int dg_csr_expand_rows(const int32_t* indptr, int64_t n_rows, int32_t* row, dg_stream_t stream) {
  DG_REQUIRE(n_rows >= 0, "n_rows < 0");
  if (n_rows == 0) return DG_OK;
  dg::expand_rows_kernel<<<dg::grid_for(n_rows * 32, 256), 256, 0, dg::as_stream(stream)>>>(indptr, n_rows, row);
  DG_CHECK_LAUNCH("expand_rows");
  return DG_OK;
}
}
