// Index primitives used by the graph builders: exclusive scan (int32) and a stable LSD radix sort
// of (uint64 key, int32 value) pairs. Both are deterministic; the sort is stable so that equal
// (row, col) keys keep edge-id order.
#include <stdlib.h>

#include "common.cuh"
#include "primitives.cuh"

namespace dg {

// ------------------------------------------------------------------------------------------------
// exclusive scan: tile sums -> spine scan (one CTA) -> per-tile scan with carried offset
// ------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ int warp_inclusive_scan(int v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(kFull, v, o);
    if (lane_id() >= o) v += t;
  }
  return v;
}

// Block-wide exclusive scan of one value per thread (blockDim.x <= 1024); returns the exclusive
// prefix and writes the block total to *total (valid for all threads after return).
__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
  __shared__ int warp_sums[32];
  __shared__ int block_total;
  int inc = warp_inclusive_scan(v);
  int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if (lane_id() == 31) warp_sums[w] = inc;
  __syncthreads();
  if (w == 0) {
    int s = lane_id() < nw ? warp_sums[lane_id()] : 0;
    int si = warp_inclusive_scan(s);
    warp_sums[lane_id()] = si - s;
    if (lane_id() == 31) block_total = si;
  }
  __syncthreads();
  int r = inc - v + warp_sums[w];
  *total = block_total;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums(const int* in, int64_t n,
                                                               int* __restrict__ tile_sums) {
  int64_t base = static_cast<int64_t>(blockIdx.x) * kScanTile;
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    int64_t idx = base + static_cast<int64_t>(i) * kScanThreads + threadIdx.x;   // coalesced
    if (idx < n) s += in[idx];
  }
  int total;
  block_exclusive_scan(s, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) scan_spine(int* __restrict__ tile_sums, int64_t n_tiles) {
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < n_tiles; base += blockDim.x) {
    int64_t idx = base + threadIdx.x;
    int v = idx < n_tiles ? tile_sums[idx] : 0;
    int total;
    int ex = block_exclusive_scan(v, &total);
    int carry = carry_s;
    if (idx < n_tiles) tile_sums[idx] = ex + carry;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + total;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kScanThreads) scan_tiles(const int* in, int* out,   // may alias (in-place scan)
                                                           int64_t n, const int* __restrict__ tile_offsets) {
  // Each thread owns kScanItems CONSECUTIVE items so the prefix is a plain running sum.
  int64_t base = static_cast<int64_t>(blockIdx.x) * kScanTile + static_cast<int64_t>(threadIdx.x) * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    int64_t idx = base + i;
    v[i] = idx < n ? in[idx] : 0;
    s += v[i];
  }
  int total;
  int ex = block_exclusive_scan(s, &total) + (tile_offsets ? tile_offsets[blockIdx.x] : 0);
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    int64_t idx = base + i;
    if (idx < n) out[idx] = ex;
    ex += v[i];
  }
  // the thread that owns the last element also writes the grand total to out[n]
  if (base < n && base + kScanItems >= n) out[n] = ex;
}

__global__ void scan_empty(int* out) { out[0] = 0; }

size_t scan_workspace_bytes(int64_t n) {
  int64_t tiles = (n + kScanTile - 1) / kScanTile;
  return ws_add(0, static_cast<size_t>(tiles > 0 ? tiles : 1) * sizeof(int));
}

int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (n < 0) { set_error("exclusive_scan: n < 0"); return DG_ERR_INVALID_ARGUMENT; }
  if (n == 0) {
    scan_empty<<<1, 1, 0, st>>>(out);
    DG_CHECK_LAUNCH("scan_empty");
    return DG_OK;
  }
  int64_t tiles = (n + kScanTile - 1) / kScanTile;
  static const bool small_on = [] { const char* v = getenv("DG_SMALL_PRIMS"); return v == nullptr || atoi(v) != 0; }();   // A/B switch
  if (tiles == 1 && small_on) {                                  // one tile (<= 4 096 items: the row counts of the real-dataset graphs): its
    scan_tiles<<<1, kScanThreads, 0, st>>>(in, out, n, nullptr);      // offset is zero, one launch instead of three
    DG_CHECK_LAUNCH("scan_tiles");
    return DG_OK;
  }
  Workspace w(ws, ws_bytes);
  int* tile_sums = w.take<int>(tiles);
  if (!tile_sums) { set_error("exclusive_scan: workspace too small"); return DG_ERR_WORKSPACE_TOO_SMALL; }
  scan_tile_sums<<<static_cast<unsigned>(tiles), kScanThreads, 0, st>>>(in, n, tile_sums);
  DG_CHECK_LAUNCH("scan_tile_sums");
  scan_spine<<<1, 1024, 0, st>>>(tile_sums, tiles);
  DG_CHECK_LAUNCH("scan_spine");
  scan_tiles<<<static_cast<unsigned>(tiles), kScanThreads, 0, st>>>(in, out, n, tile_sums);
  DG_CHECK_LAUNCH("scan_tiles");
  return DG_OK;
}

// ------------------------------------------------------------------------------------------------
// stable LSD radix sort, 8 bits per pass
//   tile = 8 warps x 256 keys; warp w owns a CONTIGUOUS run of 256 keys and walks it 32 at a time,
//   so "rank among equal digits" = (keys of that digit in earlier tiles) + (earlier warps of this
//   tile) + (earlier rounds of this warp) + (lower lanes of this round): input order is preserved.
// ------------------------------------------------------------------------------------------------
constexpr int kSortWarps = 8;
constexpr int kSortRounds = 8;
constexpr int kSortTile = kSortWarps * 32 * kSortRounds;   // 2048
constexpr int kRadix = 256;

__device__ __forceinline__ void sort_warp_histogram(const uint64_t* __restrict__ keys, int64_t n, int64_t warp_base,
                                                    int shift, int* __restrict__ hist /* [256] for this warp */) {
  const int lane = lane_id();
  for (int i = lane; i < kRadix; i += 32) hist[i] = 0;
  __syncwarp();
#pragma unroll 1
  for (int r = 0; r < kSortRounds; ++r) {
    int64_t idx = warp_base + r * 32 + lane;
    bool valid = idx < n;
    unsigned digit = valid ? static_cast<unsigned>((keys[idx] >> shift) & 0xff) : 0x100u + lane;  // unique if invalid
    unsigned peers = __match_any_sync(kFull, digit);
    if (valid && (__ffs(peers) - 1) == lane) hist[digit] += __popc(peers);
    __syncwarp();
  }
}

__global__ void __launch_bounds__(kSortWarps * 32) sort_count(const uint64_t* __restrict__ keys, int64_t n, int shift,
                                                              int* __restrict__ counts, int64_t n_tiles) {
  __shared__ int hist[kSortWarps][kRadix];
  const int w = threadIdx.x >> 5;
  int64_t warp_base = static_cast<int64_t>(blockIdx.x) * kSortTile + w * (32 * kSortRounds);
  sort_warp_histogram(keys, n, warp_base, shift, hist[w]);
  __syncthreads();
  for (int d = threadIdx.x; d < kRadix; d += blockDim.x) {
    int s = 0;
#pragma unroll
    for (int ww = 0; ww < kSortWarps; ++ww) s += hist[ww][d];
    counts[static_cast<int64_t>(d) * n_tiles + blockIdx.x] = s;   // digit-major so one scan orders it
  }
}

template <bool kHasVals>
__global__ void __launch_bounds__(kSortWarps * 32) sort_scatter(const uint64_t* __restrict__ keys_in,
                                                                const int* __restrict__ vals_in,
                                                                uint64_t* __restrict__ keys_out,
                                                                int* __restrict__ vals_out, int64_t n, int shift,
                                                                const int* __restrict__ offsets, int64_t n_tiles) {
  __shared__ int hist[kSortWarps][kRadix];
  const int w = threadIdx.x >> 5, lane = lane_id();
  int64_t warp_base = static_cast<int64_t>(blockIdx.x) * kSortTile + w * (32 * kSortRounds);
  sort_warp_histogram(keys_in, n, warp_base, shift, hist[w]);
  __syncthreads();
  // hist[w][d] <- global start of (digit d, this tile, warp w)
  for (int d = threadIdx.x; d < kRadix; d += blockDim.x) {
    int run = offsets[static_cast<int64_t>(d) * n_tiles + blockIdx.x];
#pragma unroll
    for (int ww = 0; ww < kSortWarps; ++ww) {
      int c = hist[ww][d];
      hist[ww][d] = run;
      run += c;
    }
  }
  __syncthreads();
  int* cursor = hist[w];
#pragma unroll 1
  for (int r = 0; r < kSortRounds; ++r) {
    int64_t idx = warp_base + r * 32 + lane;
    bool valid = idx < n;
    uint64_t key = valid ? keys_in[idx] : 0;
    unsigned digit = valid ? static_cast<unsigned>((key >> shift) & 0xff) : 0x100u + lane;
    unsigned peers = __match_any_sync(kFull, digit);
    int rank = __popc(peers & ((1u << lane) - 1));
    int pos = 0;
    if (valid) pos = cursor[digit] + rank;
    __syncwarp();
    if (valid && (__ffs(peers) - 1) == lane) cursor[digit] += __popc(peers);
    __syncwarp();
    if (valid) {
      keys_out[pos] = key;
      if (kHasVals) vals_out[pos] = vals_in[idx];
    }
  }
}

size_t sort_workspace_bytes(int64_t n) {
  int64_t tiles = (n + kSortTile - 1) / kSortTile;
  if (tiles < 1) tiles = 1;
  size_t b = ws_add(0, static_cast<size_t>(tiles) * kRadix * sizeof(int) + sizeof(int));   // counts (+1 total)
  b = ws_add(b, scan_workspace_bytes(tiles * kRadix));
  return b;
}

int sort_pairs_u64(uint64_t* keys_in, int32_t* vals_in, uint64_t* keys_out, int32_t* vals_out, int64_t n,
                   int key_bits, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (n < 0 || key_bits < 0 || key_bits > 64) { set_error("sort_pairs: bad arguments"); return DG_ERR_INVALID_ARGUMENT; }
  if ((vals_in == nullptr) != (vals_out == nullptr)) { set_error("sort_pairs: vals_in/vals_out mismatch"); return DG_ERR_INVALID_ARGUMENT; }
  if (n == 0) return DG_OK;
  int passes = (key_bits + 7) / 8;
  if (passes < 1) passes = 1;
  int64_t tiles = (n + kSortTile - 1) / kSortTile;
  Workspace w(ws, ws_bytes);
  int* counts = w.take<int>(tiles * kRadix + 1);
  size_t scan_bytes = scan_workspace_bytes(tiles * kRadix);
  char* scan_ws = w.take<char>(scan_bytes);
  if (!counts || !scan_ws) { set_error("sort_pairs: workspace too small"); return DG_ERR_WORKSPACE_TOO_SMALL; }
  uint64_t* kin = keys_in; uint64_t* kout = keys_out;
  int32_t* vin = vals_in; int32_t* vout = vals_out;
  for (int p = 0; p < passes; ++p) {
    int shift = p * 8;
    sort_count<<<static_cast<unsigned>(tiles), kSortWarps * 32, 0, st>>>(kin, n, shift, counts, tiles);
    DG_CHECK_LAUNCH("sort_count");
    DG_PROPAGATE(exclusive_scan_i32(counts, counts, tiles * kRadix, scan_ws, scan_bytes, st));
    if (vin)
      sort_scatter<true><<<static_cast<unsigned>(tiles), kSortWarps * 32, 0, st>>>(kin, vin, kout, vout, n, shift, counts, tiles);
    else
      sort_scatter<false><<<static_cast<unsigned>(tiles), kSortWarps * 32, 0, st>>>(kin, nullptr, kout, nullptr, n, shift, counts, tiles);
    DG_CHECK_LAUNCH("sort_scatter");
    uint64_t* tk = kin; kin = kout; kout = tk;
    int32_t* tv = vin; vin = vout; vout = tv;
  }
  // after the loop the sorted data is in `kin`; make sure it ends in keys_out
  if (kin != keys_out) {
    DG_CHECK_CUDA(cudaMemcpyAsync(keys_out, kin, static_cast<size_t>(n) * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
    if (vals_out) DG_CHECK_CUDA(cudaMemcpyAsync(vals_out, vin, static_cast<size_t>(n) * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  }
  return DG_OK;
}

}  // namespace dg

extern "C" {
size_t dg_scan_workspace_bytes(int64_t n) { return dg::scan_workspace_bytes(n); }
int dg_exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, void* workspace, size_t workspace_bytes,
                          dg_stream_t stream) {
  return dg::exclusive_scan_i32(in, out, n, workspace, workspace_bytes, dg::as_stream(stream));
}
size_t dg_sort_workspace_bytes(int64_t n) { return dg::sort_workspace_bytes(n); }
int dg_sort_pairs_u64(uint64_t* keys_in, int32_t* vals_in, uint64_t* keys_out, int32_t* vals_out, int64_t n,
                      int key_bits, void* workspace, size_t workspace_bytes, dg_stream_t stream) {
  return dg::sort_pairs_u64(keys_in, vals_in, keys_out, vals_out, n, key_bits, workspace, workspace_bytes,
                            dg::as_stream(stream));
}
}
