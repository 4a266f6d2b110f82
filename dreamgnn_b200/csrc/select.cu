// Uniform random k-subset of n edges without a sort: "keep the edges listed first in a random permutation"
// (augmentation.py:48-52, :113-118) only needs the SET of the k smallest of n i.i.d. random keys, i.e. a selection.
//
//   key_i = (r_i << b) | i     r_i = 63 random bits drawn by the caller (torch's generator), b = bits of n - 1,
//                              so all keys are distinct and the k-th smallest is a well-defined threshold T
//   8 MSB-first radix passes:  256-bin histogram of the current digit over the keys that still match the prefix of T
//                              (shared-memory counts, integer adds: order-independent), then one CTA picks the digit
//                              in which the running count crosses k
//   flags[i] = key_i <= T      exactly k ones, written sequentially (no scatter)
//
// 9 reads of the 8 n bytes of r (HBM streaming) instead of torch.randperm's seven 32 n byte radix-sort passes. Used for
// the edge dropout inside captured CUDA graphs (bench.py, train --cuda_graph), where the generator's offsets differ
// from an eager run anyway; the eager path keeps th.randperm so that its kept sets are the reference's for a given
// generator state.
#include <stdlib.h>

#include "common.cuh"

namespace dg {
namespace {

struct SelectState {
  unsigned long long prefix;     // digits of T decided so far (in their final bit positions)
  unsigned long long k_rem;      // rank of T among the keys matching the prefix (1-based)
  unsigned int hist[256];
};

__device__ __forceinline__ unsigned long long make_key(long long r, int64_t i, int idx_bits) {
  return (static_cast<unsigned long long>(r) << idx_bits) | static_cast<unsigned long long>(i);
}

__global__ void select_init_kernel(SelectState* st, unsigned long long k) {
  const int t = threadIdx.x;
  st->hist[t] = 0;
  if (t == 0) { st->prefix = 0; st->k_rem = k; }
}

__global__ void __launch_bounds__(256)
select_hist_kernel(const long long* __restrict__ rnd, int64_t n, int idx_bits, int pass, SelectState* __restrict__ st) {
  __shared__ unsigned int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int shift = 56 - 8 * pass;
  const unsigned long long prefix = st->prefix;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const unsigned long long key = make_key(rnd[i], i, idx_bits);
    // pass 0 looks at every key; later passes only at those whose higher digits equal the prefix
    if (pass == 0 || (key >> (shift + 8)) == (prefix >> (shift + 8)))
      atomicAdd(&h[(key >> shift) & 255ull], 1u);
  }
  __syncthreads();
  const unsigned int c = h[threadIdx.x];
  if (c) atomicAdd(&st->hist[threadIdx.x], c);
}

__global__ void __launch_bounds__(256) select_pick_kernel(SelectState* st, int pass) {
  __shared__ unsigned long long wsum[8];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const unsigned long long mine_cnt = st->hist[t];
  unsigned long long incl = mine_cnt;                          // inclusive scan of the 256 bins: warp shuffles + 8 warp sums
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long up = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  for (int w = 0; w < warp; ++w) incl += wsum[w];
  const unsigned long long excl = incl - mine_cnt;
  const unsigned long long k = st->k_rem;
  __syncthreads();                                            // everyone has read k_rem before the owner rewrites it
  if (excl < k && k <= incl) {                                // exactly one digit satisfies this (k <= total)
    st->prefix |= static_cast<unsigned long long>(t) << (56 - 8 * pass);
    st->k_rem = k - excl;
  }
  st->hist[t] = 0;                                            // ready for the next pass
}

__global__ void select_flags_kernel(const long long* __restrict__ rnd, int64_t n, int idx_bits,
                                    const SelectState* __restrict__ st, uint8_t* __restrict__ flags) {
  const unsigned long long T = st->prefix;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    flags[i] = make_key(rnd[i], i, idx_bits) <= T ? 1 : 0;
}

// The whole selection for a small relation (n <= kSelectSmall keys: the kNN graphs and the positive-label relations of the
// real-dataset shapes, a few thousand edges) in ONE single-CTA launch instead of 18: the same eight histogram / pick passes
// over the same keys (read from L1 / L2 each pass: <= 128 KB), the state in shared memory, then the flags.
constexpr int kSelectSmall = 16384;
constexpr int kSelectSmallThreads = 1024;
__global__ void __launch_bounds__(kSelectSmallThreads)
select_small_kernel(const long long* __restrict__ rnd, int n, int idx_bits, unsigned long long k, uint8_t* __restrict__ flags) {
  __shared__ unsigned int h[256];
  __shared__ unsigned long long wsum[8];
  __shared__ unsigned long long s_prefix, s_krem;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t == 0) { s_prefix = 0; s_krem = k; }
  for (int pass = 0; pass < 8; ++pass) {
    if (t < 256) h[t] = 0;
    __syncthreads();
    const int shift = 56 - 8 * pass;
    const unsigned long long prefix = s_prefix;
    for (int i = t; i < n; i += kSelectSmallThreads) {
      const unsigned long long key = make_key(rnd[i], i, idx_bits);
      if (pass == 0 || (key >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&h[(key >> shift) & 255ull], 1u);
    }
    __syncthreads();
    // pick: inclusive scan of the 256 bins by the first 8 warps, exactly as select_pick_kernel
    unsigned long long mine_cnt = 0, incl = 0;
    if (t < 256) {
      mine_cnt = h[t];
      incl = mine_cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long up = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += up;
      }
      if (lane == 31) wsum[warp] = incl;
    }
    __syncthreads();
    const unsigned long long kr = s_krem;
    if (t < 256) for (int w = 0; w < warp; ++w) incl += wsum[w];
    __syncthreads();                                          // everyone has read k_rem / wsum before the owner rewrites
    if (t < 256) {
      const unsigned long long excl = incl - mine_cnt;
      if (excl < kr && kr <= incl) {
        s_prefix = prefix | (static_cast<unsigned long long>(t) << shift);
        s_krem = kr - excl;
      }
    }
    __syncthreads();
  }
  const unsigned long long T = s_prefix;
  for (int i = t; i < n; i += kSelectSmallThreads) flags[i] = make_key(rnd[i], i, idx_bits) <= T ? 1 : 0;
}

}  // namespace
}  // namespace dg

extern "C" {

size_t dg_random_subset_workspace_bytes(void) { return 256 + sizeof(dg::SelectState); }

int dg_random_subset_flags(const int64_t* rnd, int64_t n, int64_t k, uint8_t* flags, void* workspace, size_t workspace_bytes,
                           dg_stream_t stream) {
  using namespace dg;
  DG_REQUIRE(n >= 0 && k >= 0 && k <= n, "need 0 <= k <= n");
  if (n == 0) return DG_OK;
  DG_REQUIRE(rnd != nullptr && flags != nullptr, "null pointer");
  cudaStream_t s = as_stream(stream);
  if (k == 0 || k == n) {
    DG_CHECK_CUDA(cudaMemsetAsync(flags, k == n ? 1 : 0, static_cast<size_t>(n), s));
    return DG_OK;
  }
  Workspace ws(workspace, workspace_bytes);
  SelectState* st = ws.take<SelectState>(1);
  if (st == nullptr) {
    set_error("dg_random_subset_flags: workspace too small");
    return DG_ERR_WORKSPACE_TOO_SMALL;
  }
  int idx_bits = 0;
  while (idx_bits < 40 && (1ll << idx_bits) < n) ++idx_bits;           // bits of n - 1 (n >= 2 here)
  static const bool small_on = [] { const char* v = getenv("DG_SMALL_PRIMS"); return v == nullptr || atoi(v) != 0; }();   // A/B switch
  if (n <= kSelectSmall && small_on) {
    select_small_kernel<<<1, kSelectSmallThreads, 0, s>>>(reinterpret_cast<const long long*>(rnd), static_cast<int>(n), idx_bits,
                                                          static_cast<unsigned long long>(k), flags);
    DG_CHECK_LAUNCH("select_small");
    return DG_OK;
  }
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = static_cast<int64_t>(kNumSM) * 16;
  if (blocks > cap) blocks = cap;
  select_init_kernel<<<1, 256, 0, s>>>(st, static_cast<unsigned long long>(k));
  DG_CHECK_LAUNCH("select_init");
  for (int pass = 0; pass < 8; ++pass) {
    select_hist_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(reinterpret_cast<const long long*>(rnd), n, idx_bits, pass, st);
    DG_CHECK_LAUNCH("select_hist");
    select_pick_kernel<<<1, 256, 0, s>>>(st, pass);
    DG_CHECK_LAUNCH("select_pick");
  }
  select_flags_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(reinterpret_cast<const long long*>(rnd), n, idx_bits, st, flags);
  DG_CHECK_LAUNCH("select_flags");
  return DG_OK;
}

}  // extern "C"
