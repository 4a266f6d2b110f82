// fp32 GEMM for the small dense layers of the real-dataset shapes (nn.Linear ufc / ifc / fusion / lin1 and the 128 -> 128
// relation projections of GCMC layers 1-2 at N ~ 600-800 nodes: 10-100 MFLOP each; layers.py:139-142, 220-221, 281-282,
// 366-369).
//
// At these sizes a product is bound by latency, not by any pipe: the tcgen05 kernel of gemm_tc.cu has a ~20 us floor
// (tensor-map fetch, TMEM allocation, a 2-3 stage TMA ring that 8 k-blocks never fill), and the library runs each one as
// a SIMT kernel PLUS a split-K reduction launch PLUS a bias epilogue launch, often behind a layout copy. This kernel is
// one launch per product whatever the operand layout:
//   * C[b] = op(A[b]) . op(B[b])^T (+ bias), op(A) [M, K], op(B) [N, K]; either operand may be stored transposed
//     ([K, M] / [K, N]) and is read as stored (an operand contiguous along k lands in a row-major shared-memory tile, a
//     transposed one in a k-major tile; the thread -> output mapping of each side follows its tile so that every
//     shared-memory read is a conflict-free LDS.64 / LDS.128);
//   * 32 x 32 output tiles (a 763 x 128 output is 96 CTAs, not 24: what bounds a product this small is how long ONE CTA
//     works), 128 threads, 2 x 4 outputs per thread, 32-deep k-blocks through a 4-stage ring of 16-byte cp.async copies
//     (4-byte copies when a leading dimension or extent is not a multiple of 4: no alignment or padding rules);
//   * `reduce_batch`: C = sum_b op(A[b]) . op(B[b])^T with the batch walked inside the k-loop (dx = sum_r dy[r] . W_r^T of
//     the relation projections): no partial tiles at all;
//   * split-K over CTAs only for very long K with few tiles, the partial tiles then added IN THE KERNEL by whichever CTA
//     of a tile finishes last -- in split order, so the result does not depend on which one that is (deterministic, no
//     atomics on data, no second launch);
//   * plain fp32 FMA in ascending k: no TF32 split to compensate.
#include "common.cuh"

namespace dg {
namespace {

constexpr int kTgBM = 32, kTgBN = 32, kTgBK = 32, kTgThreads = 128, kTgStages = 4;
constexpr int kTgLd = kTgBK + 4;                       // row stride of a row-major tile: LDS.128 along k over 8 rows hits 8 bank groups
constexpr int kTgTile = kTgBM * kTgLd;                 // floats per operand tile (the k-major form needs 32 x 32 of them)
constexpr int kTgMaxSplits = 32;

struct SmallGemmParams {
  const float* A; const float* B; const float* bias; float* C; float* partial; int* tickets;
  int64_t lda, ldb, ldc, stride_a, stride_b, stride_c;
  int M, N, K, batch, splits, k_per_split, reduce_batch, vec_a, vec_b;
};

// asynchronous copy global -> shared; `valid == false` writes zeros without reading
__device__ __forceinline__ void tg_cp16(float* smem, const float* gmem, bool valid) {
  const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem), "r"(n) : "memory");
}
__device__ __forceinline__ void tg_cp4(float* smem, const float* gmem, bool valid) {
  const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  const int n = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(gmem), "r"(n) : "memory");
}

// One operand tile (32 rows of op(X) x 32 k) -> shared memory. kT: X is stored [K, rows] and lands k-major ([k][row], row
// stride 32); else stored [rows, K] and lands row-major ([row][k], row stride 36). Each thread moves two groups of four
// elements that are contiguous in global AND shared memory.
template <bool kT>
__device__ __forceinline__ void tg_issue(float* tile, const float* __restrict__ X, int64_t ld, int row0, int rows, int kb, int k_end,
                                         bool vec) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int v = threadIdx.x + h * kTgThreads;        // 0 .. 255
    const int outer = v >> 3, inner = (v & 7) * 4;
    const int r = kT ? inner : outer, k = kT ? outer : inner;           // first element of the group
    float* dst = tile + (kT ? k * kTgBM + r : r * kTgLd + k);
    const float* src = kT ? X + static_cast<int64_t>(kb + k) * ld + row0 + r : X + static_cast<int64_t>(row0 + r) * ld + kb + k;
    if (vec) {
      const bool ok = (row0 + r < rows) && (kb + k < k_end);            // extents are multiples of 4: never partial
      tg_cp16(dst, ok ? src : X, ok);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool ok = kT ? (row0 + r + e < rows && kb + k < k_end) : (row0 + r < rows && kb + k + e < k_end);
        tg_cp4(dst + e, ok ? src + e : X, ok);
      }
    }
  }
}

template <bool kTA, bool kTB>
__global__ void __launch_bounds__(kTgThreads)
small_gemm_kernel(const SmallGemmParams p) {
  __shared__ __align__(16) float As[kTgStages][kTgTile];
  __shared__ __align__(16) float Bs[kTgStages][kTgTile];
  __shared__ int is_last;
  const int t = threadIdx.x, tx = t & 7, ty = t >> 3;
  const int m_tile = blockIdx.y * kTgBM, n_tile = blockIdx.x * kTgBN;
  const int z = blockIdx.z;
  const int b = p.reduce_batch ? 0 : z / p.splits, s = p.reduce_batch ? 0 : z - b * p.splits;
  const int k0 = s * p.k_per_split, k1 = min(p.K, k0 + p.k_per_split);
  const int nkb = (k1 - k0 + kTgBK - 1) / kTgBK;       // k-blocks per batch entry
  const int n_blocks = nkb * (p.reduce_batch ? p.batch : 1);
  auto issue = [&](int stage, int q) {                 // block q of this CTA's sequence (batch-major when reducing over it)
    const int bq = p.reduce_batch ? q / nkb : b, kb = k0 + (q - (p.reduce_batch ? bq * nkb : 0)) * kTgBK;
    tg_issue<kTA>(As[stage], p.A + static_cast<int64_t>(bq) * p.stride_a, p.lda, m_tile, p.M, kb, k1, p.vec_a != 0);
    tg_issue<kTB>(Bs[stage], p.B + static_cast<int64_t>(bq) * p.stride_b, p.ldb, n_tile, p.N, kb, k1, p.vec_b != 0);
  };
  float acc[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // kTgStages - 1 blocks in flight ahead of the one being multiplied; one commit group per block (empty past the end, so
  // that the wait count stays a compile-time constant)
#pragma unroll
  for (int st = 0; st < kTgStages - 1; ++st) {
    if (st < n_blocks) issue(st, st);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  int stage = 0;
  for (int q = 0; q < n_blocks; ++q) {
    asm volatile("cp.async.wait_group %0;" ::"n"(kTgStages - 2) : "memory");
    __syncthreads();                                   // block q has landed for every thread; the stage refilled below was
                                                       // read by everyone in the previous trip
    if (q + kTgStages - 1 < n_blocks) issue((stage + kTgStages - 1) % kTgStages, q + kTgStages - 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const float* at = As[stage];
    const float* bt = Bs[stage];
#pragma unroll
    for (int k = 0; k < kTgBK; k += 4) {
      float a[2][4], bb[4][4];
      if (kTA) {                                       // k-major: rows 2 ty, 2 ty + 1
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float2 v = *reinterpret_cast<const float2*>(at + (k + kk) * kTgBM + ty * 2);
          a[0][kk] = v.x; a[1][kk] = v.y;
        }
      } else {                                         // row-major: rows ty, ty + 16
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float4 v = *reinterpret_cast<const float4*>(at + (ty + 16 * i) * kTgLd + k);
          a[i][0] = v.x; a[i][1] = v.y; a[i][2] = v.z; a[i][3] = v.w;
        }
      }
      if (kTB) {                                       // k-major: columns 4 tx .. 4 tx + 3
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float4 v = *reinterpret_cast<const float4*>(bt + (k + kk) * kTgBN + tx * 4);
          bb[0][kk] = v.x; bb[1][kk] = v.y; bb[2][kk] = v.z; bb[3][kk] = v.w;
        }
      } else {                                         // row-major: columns tx, tx + 8, tx + 16, tx + 24
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 v = *reinterpret_cast<const float4*>(bt + (tx + 8 * j) * kTgLd + k);
          bb[j][0] = v.x; bb[j][1] = v.y; bb[j][2] = v.z; bb[j][3] = v.w;
        }
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i][kk], bb[j][kk], acc[i][j]);
    }
    stage = (stage + 1) % kTgStages;
  }
  int row[2], col[4];
#pragma unroll
  for (int i = 0; i < 2; ++i) row[i] = m_tile + (kTA ? ty * 2 + i : ty + 16 * i);
#pragma unroll
  for (int j = 0; j < 4; ++j) col[j] = n_tile + (kTB ? tx * 4 + j : tx + 8 * j);
  if (p.splits == 1) {                                  // the tile is complete: bias and out
    float* C = p.C + static_cast<int64_t>(b) * p.stride_c;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (row[i] < p.M && col[j] < p.N) C[static_cast<int64_t>(row[i]) * p.ldc + col[j]] = acc[i][j] + (p.bias ? __ldg(p.bias + col[j]) : 0.f);
    return;
  }
  // split-K: partial tile -> workspace [z][M][N]; the CTA that arrives last at this (output, tile) adds the parts in order
  const int64_t mn = static_cast<int64_t>(p.M) * p.N;
  float* mine = p.partial + static_cast<int64_t>(z) * mn;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (row[i] < p.M && col[j] < p.N) mine[static_cast<int64_t>(row[i]) * p.N + col[j]] = acc[i][j];
  __threadfence();
  __syncthreads();
  int* ticket = p.tickets + (static_cast<int64_t>(b) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  if (t == 0) {
    const int got = atomicAdd(ticket, 1);
    is_last = (got == p.splits - 1);
    if (is_last) *ticket = 0;                           // zero again for the next call / graph replay
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  float* C = p.C + static_cast<int64_t>(b) * p.stride_c;
  float sum[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) sum[i][j] = 0.f;
  for (int q = 0; q < p.splits; ++q) {                  // eight independent loads per part
    const float* part = p.partial + static_cast<int64_t>(b * p.splits + q) * mn;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (row[i] < p.M && col[j] < p.N) sum[i][j] += __ldcg(part + static_cast<int64_t>(row[i]) * p.N + col[j]);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (row[i] < p.M && col[j] < p.N) C[static_cast<int64_t>(row[i]) * p.ldc + col[j]] = sum[i][j] + (p.bias ? __ldg(p.bias + col[j]) : 0.f);
}

int small_gemm_splits(int64_t M, int64_t N, int64_t K, int64_t batch, int reduce_batch) {
  if (reduce_batch && batch > 1) return 1;              // the batch is walked inside the k-loop
  const int64_t tiles = ((M + kTgBM - 1) / kTgBM) * ((N + kTgBN - 1) / kTgBN) * batch;
  int64_t s = 1;
  if (tiles < kNumSM / 2 && K >= 2048) {                // a 32-deep k-block is ~0.1 us: only very long K is worth parts
    s = (kNumSM + tiles - 1) / tiles;
    if (s > K / 512) s = K / 512;
    if (s > kTgMaxSplits) s = kTgMaxSplits;
    if (s < 1) s = 1;
  }
  return static_cast<int>(s);
}

inline bool tg_vec_ok(const float* p, int64_t ld, int64_t stride, int64_t inner_extent) {
  return reinterpret_cast<uintptr_t>(p) % 16 == 0 && ld % 4 == 0 && stride % 4 == 0 && inner_extent % 4 == 0;
}

}  // namespace
}  // namespace dg

extern "C" {

size_t dg_small_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int64_t batch) {
  const size_t splits = static_cast<size_t>(dg::small_gemm_splits(M, N, K, batch, 0));
  return dg::ws_add(0, (splits > 1 ? splits * static_cast<size_t>(batch) : 0) * static_cast<size_t>(M) * static_cast<size_t>(N) * sizeof(float));
}

int64_t dg_small_gemm_tickets(int64_t M, int64_t N, int64_t K, int64_t batch) {
  if (dg::small_gemm_splits(M, N, K, batch, 0) == 1) return 0;
  return ((M + dg::kTgBM - 1) / dg::kTgBM) * ((N + dg::kTgBN - 1) / dg::kTgBN) * batch;
}

int dg_small_gemm_f32(const float* A, int64_t lda, int64_t stride_a, int trans_a, const float* B, int64_t ldb, int64_t stride_b,
                      int trans_b, const float* bias, float* C, int64_t ldc, int64_t stride_c, int64_t M, int64_t N, int64_t K,
                      int64_t batch, int reduce_batch, void* workspace, size_t workspace_bytes, int32_t* tickets,
                      dg_stream_t stream) {
  using namespace dg;
  DG_REQUIRE(A && B && C, "null pointer");
  DG_REQUIRE(M > 0 && N > 0 && K > 0 && batch > 0 && batch < 65536, "M, N, K, batch must be positive");
  DG_REQUIRE(M < (1 << 30) && N < (1 << 30) && K < (1 << 30), "shape too large");
  DG_REQUIRE(lda >= (trans_a ? M : K) && ldb >= (trans_b ? N : K) && ldc >= N, "leading dimension too small");
  SmallGemmParams p;
  p.A = A; p.B = B; p.bias = bias; p.C = C;
  p.lda = lda; p.ldb = ldb; p.ldc = ldc; p.stride_a = stride_a; p.stride_b = stride_b;
  p.M = static_cast<int>(M); p.N = static_cast<int>(N); p.K = static_cast<int>(K); p.batch = static_cast<int>(batch);
  p.reduce_batch = (reduce_batch && batch > 1) ? 1 : 0;
  p.stride_c = p.reduce_batch ? 0 : stride_c;
  // 16-byte copies need every group of four elements aligned and whole: base, leading dimension, batch stride and the
  // extent along the contiguous direction (K, or the row count of a transposed operand) multiples of 4
  p.vec_a = tg_vec_ok(A, lda, stride_a, trans_a ? M : K) ? 1 : 0;
  p.vec_b = tg_vec_ok(B, ldb, stride_b, trans_b ? N : K) ? 1 : 0;
  p.splits = small_gemm_splits(M, N, K, batch, p.reduce_batch);
  const int kb = static_cast<int>((K + kTgBK - 1) / kTgBK);
  p.k_per_split = ((kb + p.splits - 1) / p.splits) * kTgBK;
  p.splits = static_cast<int>((K + p.k_per_split - 1) / p.k_per_split);           // no empty split
  p.partial = nullptr;
  p.tickets = tickets;
  if (p.splits > 1) {
    DG_REQUIRE(tickets != nullptr, "split-K needs zeroed tickets (dg_small_gemm_tickets ints)");
    Workspace ws(workspace, workspace_bytes);
    p.partial = ws.take<float>(static_cast<size_t>(p.batch) * p.splits * static_cast<size_t>(M) * static_cast<size_t>(N));
    if (!p.partial) {
      set_error("dg_small_gemm_f32: workspace too small");
      return DG_ERR_WORKSPACE_TOO_SMALL;
    }
  }
  const dim3 grid(static_cast<unsigned>((N + kTgBN - 1) / kTgBN), static_cast<unsigned>((M + kTgBM - 1) / kTgBM),
                  static_cast<unsigned>(p.reduce_batch ? 1 : p.batch * p.splits));
  DG_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "grid too large");
  cudaStream_t st = as_stream(stream);
  if (trans_a && trans_b) small_gemm_kernel<true, true><<<grid, kTgThreads, 0, st>>>(p);
  else if (trans_a) small_gemm_kernel<true, false><<<grid, kTgThreads, 0, st>>>(p);
  else if (trans_b) small_gemm_kernel<false, true><<<grid, kTgThreads, 0, st>>>(p);
  else small_gemm_kernel<false, false><<<grid, kTgThreads, 0, st>>>(p);
  DG_CHECK_LAUNCH("small_gemm");
  return DG_OK;
}

}  // extern "C"
