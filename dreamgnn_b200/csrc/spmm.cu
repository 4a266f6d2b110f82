// CSR SpMM for the GCMC and FGCN aggregations, forward and backward (the backward is the same kernel
// on the transposed CSR with the two scale vectors swapped -- no atomics, bit-reproducible).
//
//   out[i,:] = epi( dst_scale[i] * sum_{s in row i} vals[s] * src_scale[idx[s]] * x[idx[s], :] )
//
// replaces DGL update_all(copy_u,sum) with `feat * dropout(cj)` and `rst * ci` fused
// (layers.py:224-234) and torch.spmm(adj, support) + bias (layers.py:312-314).
//
// Mapping: one warp per (row, column slab). A slab is 32 lanes x NCHUNK 16-byte vectors, so a lane
// issues NCHUNK independent 128-bit loads per gathered row and a warp reads whole 512 B segments
// of it. The warp first pulls 32 column indices (+ weights) with one coalesced load, then walks
// them kUnroll at a time: all loads of the group are issued before any FMA so kUnroll*NCHUNK
// 128-bit requests per lane are in flight. HBM/L2-bound integer-indexed gather: no tensor cores.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"

namespace dg {

struct LoadF32 {
  static constexpr int kVec = 4;
  using Elem = float;
  using Raw = float4;
  __device__ static __forceinline__ Raw load(const Elem* p) { return ldg_f4_stream(reinterpret_cast<const float4*>(p)); }
  __device__ static __forceinline__ Raw load_l1(const Elem* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
  __device__ static __forceinline__ void unpack(const Raw& t, float (&v)[kVec]) {
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
};

struct LoadBF16 {
  static constexpr int kVec = 8;
  using Elem = __nv_bfloat16;
  using Raw = uint4;
  __device__ static __forceinline__ Raw load(const Elem* p) { return ldg_u4_stream(reinterpret_cast<const uint4*>(p)); }
  __device__ static __forceinline__ Raw load_l1(const Elem* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ static __forceinline__ void unpack(const Raw& t, float (&v)[kVec]) {
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {   // bf16 -> fp32 is a 16-bit shift
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};

// L2 prefetch of the 128-byte lines [first, last] of one gathered row, spread over `lanes` lanes (sub = lane's slot)
// (kL1: into L1 instead -- the demand loads of that instance allocate in / hit L1, so the L2 -> SM latency of the next
// group is covered too, again without holding registers)
template <bool kL1 = false>
__device__ __forceinline__ void prefetch_row_l2(const void* row, unsigned bytes, int sub, int lanes) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(row);
  const uintptr_t first = a >> 7, last = (a + bytes - 1) >> 7;
  for (uintptr_t ln = first + sub; ln <= last; ln += lanes) {
    if (kL1) asm volatile("prefetch.global.L1 [%0];" ::"l"(ln << 7) : "memory");
    else asm volatile("prefetch.global.L2 [%0];" ::"l"(ln << 7) : "memory");
  }
}

// Prefetch distance in groups of kUnroll rows (0 = off). DG_SPMM_PREFETCH overrides the default for experiments.
static int spmm_prefetch_distance() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DG_SPMM_PREFETCH");
    v = e ? atoi(e) : 1;
    if (v < 0) v = 0;
  }
  return v;
}

// kPf: 0 = no prefetch, 1 = L2 prefetch (streaming demand loads), 2 = L1 prefetch (demand loads through L1)
static int spmm_prefetch_l1() {          // DG_SPMM_PF_L1=<distance in groups>: L1-prefetch instance (0 / unset = off)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DG_SPMM_PF_L1");
    v = e ? atoi(e) : 0;
    if (v < 0) v = 0;
  }
  return v;
}

template <typename L, int NCHUNK, int kUnroll, bool kWeighted, int kMinBlocks, int kPf>
__global__ void __launch_bounds__(256, kMinBlocks)
spmm_csr_kernel(const int* __restrict__ indptr, const int* __restrict__ indices, const float* __restrict__ vals,
                const float* __restrict__ src_scale, const float* __restrict__ dst_scale,
                const float* __restrict__ bias, const typename L::Elem* __restrict__ x, int64_t ldx,
                float* __restrict__ out, int64_t ldo, int64_t n_rows, int d, int n_slabs, int flags, int pf_dist) {
  constexpr int V = L::kVec;
  constexpr int kSlabCols = 32 * V * NCHUNK;
  const int lane = lane_id();
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t row = warp / n_slabs;
  if (row >= n_rows) return;
  const int slab = static_cast<int>(warp - row * n_slabs);
  const int col0 = slab * kSlabCols + lane * V;     // chunk c covers columns col0 + c*32*V .. +V
  // L2 prefetch of the rows this warp will gather `pf_dist` groups from now (prefetch.global.L2 on each 128-byte line
  // of the row, 32 / kUnroll lanes per row). Holds no registers and no shared memory, so the bytes in flight towards
  // DRAM are no longer bounded by the register file; the demand loads that follow find their lines in L2.
  const int slab_c0 = slab * kSlabCols;
  const unsigned pf_bytes = static_cast<unsigned>((min(d, slab_c0 + kSlabCols) - slab_c0) * static_cast<int>(sizeof(typename L::Elem)));
  const int pf_lo = pf_dist * kUnroll;               // prefetch window [t + pf_lo, t + pf_lo + kUnroll) in batch positions
  constexpr int kPfLanes = 32 / kUnroll;             // lanes sharing one row of the window
  const int pf_r = lane / kPfLanes, pf_sub = lane % kPfLanes;

  float acc[NCHUNK][V];
  bool live[NCHUNK];
#pragma unroll
  for (int c = 0; c < NCHUNK; ++c) {
    live[c] = col0 + c * 32 * V < d;
#pragma unroll
    for (int v = 0; v < V; ++v) acc[c][v] = 0.f;
  }

  const int beg = indptr[row], end = indptr[row + 1];
  // software prefetch of the next 32 (column, weight) pairs: the index load, and the dependent gather
  // of src_scale[column], are issued one batch ahead so neither latency sits in front of the row loads
  int j_nx = 0;
  float w_nx = 0.f, sc_nx = 1.f;
  if (beg + lane < end) {
    j_nx = ldg_i32_stream(indices + beg + lane);
    if (kWeighted) {
      w_nx = vals ? vals[beg + lane] : 1.f;
      if (src_scale) sc_nx = src_scale[j_nx];
    }
  }
  for (int base = beg; base < end; base += 32) {
    const int j = j_nx;
    const float w = kWeighted ? w_nx * sc_nx : 0.f;
    const int sn = base + 32 + lane;
    const bool has_next = sn < end;
    j_nx = 0;
    if (has_next) {
      j_nx = ldg_i32_stream(indices + sn);
      if (kWeighted) w_nx = vals ? vals[sn] : 1.f;
    } else if (kWeighted) {
      w_nx = 0.f;
    }
    const int cnt = min(32, end - base);
    int t = 0;
    if (kPf && base == beg) {
      // start of the row: the groups 1 .. pf_dist-1 of the first batch, which no earlier window covered
      for (int q0 = kUnroll; q0 < pf_lo; q0 += kUnroll) {
        const int q = q0 + pf_r;
        const int jj = __shfl_sync(kFull, j, q & 31);
        if (q < cnt) prefetch_row_l2<kPf == 2>(x + static_cast<int64_t>(jj) * ldx + slab_c0, pf_bytes, pf_sub, kPfLanes);
      }
    }
    for (; t + kUnroll <= cnt; t += kUnroll) {
      typename L::Raw buf[kUnroll][NCHUNK];   // raw 128-bit vectors; unpacked only at FMA time
      float wt[kUnroll];
      if (kPf) {
        const int q = t + pf_lo + pf_r;                 // batch position of the row this lane helps to prefetch
        if (t + pf_lo + kUnroll <= 32) {                // window inside this batch (warp-uniform branch)
          const int jj = __shfl_sync(kFull, j, q & 31);
          if (q < cnt) prefetch_row_l2<kPf == 2>(x + static_cast<int64_t>(jj) * ldx + slab_c0, pf_bytes, pf_sub, kPfLanes);
        } else if (t + pf_lo >= 32) {                   // window inside the next batch: its indices landed long ago
          const int jj = __shfl_sync(kFull, j_nx, (q - 32) & 31);
          if (base + q < end) prefetch_row_l2<kPf == 2>(x + static_cast<int64_t>(jj) * ldx + slab_c0, pf_bytes, pf_sub, kPfLanes);
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int jj = __shfl_sync(kFull, j, t + u);
        const typename L::Elem* xr = x + static_cast<int64_t>(jj) * ldx + col0;
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c)
          if (live[c]) buf[u][c] = (kPf == 2) ? L::load_l1(xr + c * 32 * V) : L::load(xr + c * 32 * V);
      }
      if (kWeighted) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) wt[u] = __shfl_sync(kFull, w, t + u);
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u)
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c)
          if (live[c]) {
            float b[V];
            L::unpack(buf[u][c], b);
#pragma unroll
            for (int v = 0; v < V; ++v) acc[c][v] = kWeighted ? fmaf(wt[u], b[v], acc[c][v]) : acc[c][v] + b[v];
          }
    }
    for (; t < cnt; ++t) {
      const int jj = __shfl_sync(kFull, j, t);
      const float wt = kWeighted ? __shfl_sync(kFull, w, t) : 1.f;
      const typename L::Elem* xr = x + static_cast<int64_t>(jj) * ldx + col0;
#pragma unroll
      for (int c = 0; c < NCHUNK; ++c)
        if (live[c]) {
          float b[V];
          L::unpack(L::load(xr + c * 32 * V), b);
#pragma unroll
          for (int v = 0; v < V; ++v) acc[c][v] = kWeighted ? fmaf(wt, b[v], acc[c][v]) : acc[c][v] + b[v];
        }
    }
    if (kWeighted) sc_nx = (src_scale && has_next) ? src_scale[j_nx] : 1.f;   // next batch's scale, index has landed by now
  }

  const float ds = dst_scale ? dst_scale[row] : 1.f;
  float* orow = out + row * ldo;
#pragma unroll
  for (int c = 0; c < NCHUNK; ++c) {
    if (!live[c]) continue;
    const int col = col0 + c * 32 * V;
#pragma unroll
    for (int q = 0; q < V / 4; ++q) {
      float4 r = make_float4(acc[c][4 * q] * ds, acc[c][4 * q + 1] * ds, acc[c][4 * q + 2] * ds, acc[c][4 * q + 3] * ds);
      if (bias) {
        const float4 b = *reinterpret_cast<const float4*>(bias + col + 4 * q);
        r.x += b.x; r.y += b.y; r.z += b.z; r.w += b.w;
      }
      float4* op = reinterpret_cast<float4*>(orow + col + 4 * q);
      if (flags & DG_SPMM_ACCUMULATE) {
        const float4 o = *op;
        r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w;
      }
      if (flags & DG_SPMM_RELU) {
        r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f);
      }
      *op = r;
    }
  }
}

// ---- staged variant: gathered rows travel L2 -> shared memory by cp.async, not through registers ------------------------
// The warp-per-row kernel above holds its in-flight loads in registers: 16 warps x 4 rows x 1376 B = 88 KB per SM at
// d = 344, and measures ~10 TB/s of gather where the same access shape with twice the bytes in flight reaches 18-19 TB/s
// (profiles/l2_peak.json): under load the L2 round trip is ~1.5 us, so throughput is (bytes in flight) / latency. Here
// every lane copies ITS OWN 16-byte pieces of the next kRing - 1 rows into a per-warp ring with cp.async (LDGSTS: no
// destination register) and reads them back after cp.async.wait_group -- shared memory as an extension of the register
// file for loads in flight. A lane only ever reads bytes it copied itself, so no barrier of any kind is needed; one
// commit group per row keeps the wait count a compile-time constant. fp32 features only.
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NCHUNK, int kRing, bool kWeighted, int kMinBlocks>
__global__ void __launch_bounds__(256, kMinBlocks)
spmm_csr_stage_kernel(const int* __restrict__ indptr, const int* __restrict__ indices, const float* __restrict__ vals,
                      const float* __restrict__ src_scale, const float* __restrict__ dst_scale,
                      const float* __restrict__ bias, const float* __restrict__ x, int64_t ldx, float* __restrict__ out,
                      int64_t ldo, int64_t n_rows, int d, int n_slabs, int flags) {
  static_assert((kRing & (kRing - 1)) == 0 && kRing <= 32, "ring size: power of two, at most one index batch");
  constexpr int kSlabCols = 128 * NCHUNK;
  extern __shared__ float4 ring_all[];                  // [8 warps][kRing][NCHUNK][32 lanes]
  const int lane = lane_id(), wib = threadIdx.x >> 5;
  const int64_t warp = static_cast<int64_t>(blockIdx.x) * 8 + wib;
  const int64_t row = warp / n_slabs;
  if (row >= n_rows) return;
  const int slab = static_cast<int>(warp - row * n_slabs);
  const int col0 = slab * kSlabCols + lane * 4;
  float4* ring = ring_all + wib * (kRing * NCHUNK * 32);
  const uint32_t ring_s = static_cast<uint32_t>(__cvta_generic_to_shared(ring)) + lane * 16;
  float acc[NCHUNK][4];
  bool live[NCHUNK];
#pragma unroll
  for (int c = 0; c < NCHUNK; ++c) {
    live[c] = col0 + c * 128 < d;
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[c][v] = 0.f;
  }
  const int beg = indptr[row], n = indptr[row + 1] - beg;
  auto load_batch = [&](int pos, int& jj, float& ww) {   // column (+ weight) of edge `pos` of the row, 0 past the end
    jj = 0;
    ww = 0.f;
    if (pos < n) {
      jj = ldg_i32_stream(indices + beg + pos);
      if (kWeighted) ww = (vals ? vals[beg + pos] : 1.f) * (src_scale ? src_scale[jj] : 1.f);
    }
  };
  auto issue = [&](int jj, int slot) {
    const float* xr = x + static_cast<int64_t>(jj) * ldx + col0;
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c)
      if (live[c]) cp_async16(ring_s + (slot * NCHUNK + c) * 512, xr + c * 128);
  };
  int j, j_nx;
  float w, w_nx;
  load_batch(lane, j, w);
  load_batch(32 + lane, j_nx, w_nx);
#pragma unroll
  for (int p = 0; p < kRing - 1; ++p) {                  // prologue: rows 0 .. kRing-2 on their way
    const int jj = __shfl_sync(kFull, j, p);
    if (p < n) issue(jj, p);
    cp_async_commit();
  }
  for (int base = 0; base < n; base += 32) {
    const int cnt = min(32, n - base);
    for (int t = 0; t < cnt; ++t) {
      const int pa = t + kRing - 1;                       // the row that enters the ring now (warp-uniform control flow)
      const int jj = __shfl_sync(kFull, pa < 32 ? j : j_nx, pa & 31);
      if (base + pa < n) issue(jj, (base + pa) & (kRing - 1));
      cp_async_commit();
      cp_async_wait<kRing - 1>();                         // row base + t has landed (this lane's pieces of it)
      const float4* rs = ring + ((base + t) & (kRing - 1)) * NCHUNK * 32 + lane;
      const float wt = kWeighted ? __shfl_sync(kFull, w, t) : 1.f;
#pragma unroll
      for (int c = 0; c < NCHUNK; ++c)
        if (live[c]) {
          const float4 b = rs[c * 32];
          if (kWeighted) {
            acc[c][0] = fmaf(wt, b.x, acc[c][0]); acc[c][1] = fmaf(wt, b.y, acc[c][1]);
            acc[c][2] = fmaf(wt, b.z, acc[c][2]); acc[c][3] = fmaf(wt, b.w, acc[c][3]);
          } else {
            acc[c][0] += b.x; acc[c][1] += b.y; acc[c][2] += b.z; acc[c][3] += b.w;
          }
        }
    }
    j = j_nx;
    w = w_nx;
    load_batch(base + 64 + lane, j_nx, w_nx);
  }
  cp_async_wait<0>();
  const float ds = dst_scale ? dst_scale[row] : 1.f;
  float* orow = out + row * ldo;
#pragma unroll
  for (int c = 0; c < NCHUNK; ++c) {
    if (!live[c]) continue;
    const int col = col0 + c * 128;
    float4 r = make_float4(acc[c][0] * ds, acc[c][1] * ds, acc[c][2] * ds, acc[c][3] * ds);
    if (bias) {
      const float4 b = *reinterpret_cast<const float4*>(bias + col);
      r.x += b.x; r.y += b.y; r.z += b.z; r.w += b.w;
    }
    float4* op = reinterpret_cast<float4*>(orow + col);
    if (flags & DG_SPMM_ACCUMULATE) {
      const float4 o = *op;
      r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w;
    }
    if (flags & DG_SPMM_RELU) {
      r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f);
    }
    *op = r;
  }
}

// DG_SPMM_STAGE: bit 0 = staged instance for wide rows (d > 256; default on: measured -28 % on the d = 344 launches and
// -20 % on d = 768 inside the syn20m step), bit 1 = also for narrow rows (d <= 128; measured +18 %: off).
static int spmm_stage_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DG_SPMM_STAGE");
    v = e ? atoi(e) : 1;
  }
  return v;
}
// DG_SPMM_RING=4|8|16: ring depth (x 4 / 2 / 1 resident CTAs: ~190 KB of shared memory per SM either way). Measured inside
// the syn20m step, d = 344 class / d = 768 class per step: 6.95 / 3.33 ms (4), 7.27 / 3.48 ms (8), 10.94 / 4.77 ms (16) --
// more warps beat a deeper ring, so 4 x 4 is the default.
static int spmm_stage_ring() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DG_SPMM_RING");
    v = e ? atoi(e) : 4;
  }
  return v;
}

template <int NCHUNK, int kRing, int kMinBlocks>
static int launch_spmm_stage(const int* indptr, const int* indices, const float* vals, const float* src_scale,
                             const float* dst_scale, const float* bias, const float* x, int64_t ldx, float* out, int64_t ldo,
                             int64_t n_rows, int d, int flags, cudaStream_t st) {
  constexpr int kSlabCols = 128 * NCHUNK;
  constexpr size_t kSmem = static_cast<size_t>(8) * kRing * NCHUNK * 512;
  const int n_slabs = (d + kSlabCols - 1) / kSlabCols;
  const int64_t blocks = (n_rows * n_slabs + 7) / 8;
  if (blocks > 0x7fffffffLL) { set_error("spmm: grid too large"); return DG_ERR_INVALID_ARGUMENT; }
  static bool attr_set = false;
  if (!attr_set) {
    DG_CHECK_CUDA(cudaFuncSetAttribute(spmm_csr_stage_kernel<NCHUNK, kRing, true, kMinBlocks>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmem)));
    DG_CHECK_CUDA(cudaFuncSetAttribute(spmm_csr_stage_kernel<NCHUNK, kRing, false, kMinBlocks>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmem)));
    attr_set = true;
  }
  if (vals || src_scale)
    spmm_csr_stage_kernel<NCHUNK, kRing, true, kMinBlocks><<<static_cast<unsigned>(blocks), 256, kSmem, st>>>(
        indptr, indices, vals, src_scale, dst_scale, bias, x, ldx, out, ldo, n_rows, d, n_slabs, flags);
  else
    spmm_csr_stage_kernel<NCHUNK, kRing, false, kMinBlocks><<<static_cast<unsigned>(blocks), 256, kSmem, st>>>(
        indptr, indices, vals, src_scale, dst_scale, bias, x, ldx, out, ldo, n_rows, d, n_slabs, flags);
  DG_CHECK_LAUNCH("spmm_csr_stage");
  return DG_OK;
}

// ---- row-split variant for graphs with FEW, LONG rows (the real datasets: ~700 rows of ~600 edges) ---------------
// One CTA per (row, slab): its 8 warps take contiguous eighths of the row's 32-edge batches, accumulate exactly like
// the warp-per-row kernel, and warp 0 adds the eight partial rows in warp order (fixed order: bit-reproducible) before
// the same epilogue. 8x the parallelism when warp-per-row would leave most of the 148 SMs idle.
template <typename L, int NCHUNK, int kUnroll, bool kWeighted>
__global__ void __launch_bounds__(256)
spmm_csr_rowsplit_kernel(const int* __restrict__ indptr, const int* __restrict__ indices, const float* __restrict__ vals,
                         const float* __restrict__ src_scale, const float* __restrict__ dst_scale,
                         const float* __restrict__ bias, const typename L::Elem* __restrict__ x, int64_t ldx,
                         float* __restrict__ out, int64_t ldo, int64_t n_rows, int d, int n_slabs, int flags) {
  constexpr int V = L::kVec;
  constexpr int kSlabCols = 32 * V * NCHUNK;
  constexpr int kWarps = 8;
  __shared__ float red[kWarps - 1][NCHUNK][V][32];
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int64_t row = blockIdx.x / n_slabs;
  const int slab = static_cast<int>(blockIdx.x - row * n_slabs);
  const int col0 = slab * kSlabCols + lane * V;
  float acc[NCHUNK][V];
  bool live[NCHUNK];
#pragma unroll
  for (int c = 0; c < NCHUNK; ++c) {
    live[c] = col0 + c * 32 * V < d;
#pragma unroll
    for (int v = 0; v < V; ++v) acc[c][v] = 0.f;
  }
  const int rbeg = indptr[row], rend = indptr[row + 1];
  const int nb = (rend - rbeg + 31) >> 5;                       // 32-edge batches of the row
  const int beg = rbeg + ((nb * warp) / kWarps) * 32;
  const int end = min(rend, rbeg + ((nb * (warp + 1)) / kWarps) * 32);
  for (int base = beg; base < end; base += 32) {
    const int cnt = min(32, end - base);
    int j = 0;
    float w = 0.f;
    if (lane < cnt) {
      j = ldg_i32_stream(indices + base + lane);
      if (kWeighted) w = (vals ? vals[base + lane] : 1.f) * (src_scale ? src_scale[j] : 1.f);
    }
    int t = 0;
    for (; t + kUnroll <= cnt; t += kUnroll) {
      typename L::Raw buf[kUnroll][NCHUNK];
      float wt[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int jj = __shfl_sync(kFull, j, t + u);
        const typename L::Elem* xr = x + static_cast<int64_t>(jj) * ldx + col0;
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c)
          if (live[c]) buf[u][c] = L::load(xr + c * 32 * V);
      }
      if (kWeighted) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) wt[u] = __shfl_sync(kFull, w, t + u);
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u)
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c)
          if (live[c]) {
            float b[V];
            L::unpack(buf[u][c], b);
#pragma unroll
            for (int v = 0; v < V; ++v) acc[c][v] = kWeighted ? fmaf(wt[u], b[v], acc[c][v]) : acc[c][v] + b[v];
          }
    }
    for (; t < cnt; ++t) {
      const int jj = __shfl_sync(kFull, j, t);
      const float wt = kWeighted ? __shfl_sync(kFull, w, t) : 1.f;
      const typename L::Elem* xr = x + static_cast<int64_t>(jj) * ldx + col0;
#pragma unroll
      for (int c = 0; c < NCHUNK; ++c)
        if (live[c]) {
          float b[V];
          L::unpack(L::load(xr + c * 32 * V), b);
#pragma unroll
          for (int v = 0; v < V; ++v) acc[c][v] = kWeighted ? fmaf(wt, b[v], acc[c][v]) : acc[c][v] + b[v];
        }
    }
  }
  if (warp > 0) {
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c)
#pragma unroll
      for (int v = 0; v < V; ++v) red[warp - 1][c][v][lane] = acc[c][v];
  }
  __syncthreads();
  if (warp != 0) return;
#pragma unroll
  for (int wq = 0; wq < kWarps - 1; ++wq)                       // fixed order: warp 0 + warp 1 + ... + warp 7
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c)
#pragma unroll
      for (int v = 0; v < V; ++v) acc[c][v] += red[wq][c][v][lane];
  const float ds = dst_scale ? dst_scale[row] : 1.f;
  float* orow = out + row * ldo;
#pragma unroll
  for (int c = 0; c < NCHUNK; ++c) {
    if (!live[c]) continue;
    const int col = col0 + c * 32 * V;
#pragma unroll
    for (int q = 0; q < V / 4; ++q) {
      float4 r = make_float4(acc[c][4 * q] * ds, acc[c][4 * q + 1] * ds, acc[c][4 * q + 2] * ds, acc[c][4 * q + 3] * ds);
      if (bias) {
        const float4 b = *reinterpret_cast<const float4*>(bias + col + 4 * q);
        r.x += b.x; r.y += b.y; r.z += b.z; r.w += b.w;
      }
      float4* op = reinterpret_cast<float4*>(orow + col + 4 * q);
      if (flags & DG_SPMM_ACCUMULATE) {
        const float4 o = *op;
        r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w;
      }
      if (flags & DG_SPMM_RELU) {
        r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f);
      }
      *op = r;
    }
  }
}

template <typename L, int NCHUNK, int kUnroll, int kMinBlocks>
static int launch_spmm(const int* indptr, const int* indices, const float* vals, const float* src_scale,
                       const float* dst_scale, const float* bias, const typename L::Elem* x, int64_t ldx, float* out,
                       int64_t ldo, int64_t n_rows, int d, int flags, cudaStream_t st) {
  constexpr int kSlabCols = 32 * L::kVec * NCHUNK;
  const int n_slabs = (d + kSlabCols - 1) / kSlabCols;
  const int64_t warps = n_rows * n_slabs;
  if (flags & DG_SPMM_ROWSPLIT) {
    if (warps > 0x7fffffffLL) { set_error("spmm: grid too large"); return DG_ERR_INVALID_ARGUMENT; }
    if (vals || src_scale)
      spmm_csr_rowsplit_kernel<L, NCHUNK, kUnroll, true><<<static_cast<unsigned>(warps), 256, 0, st>>>(
          indptr, indices, vals, src_scale, dst_scale, bias, x, ldx, out, ldo, n_rows, d, n_slabs, flags);
    else
      spmm_csr_rowsplit_kernel<L, NCHUNK, kUnroll, false><<<static_cast<unsigned>(warps), 256, 0, st>>>(
          indptr, indices, vals, src_scale, dst_scale, bias, x, ldx, out, ldo, n_rows, d, n_slabs, flags);
    DG_CHECK_LAUNCH("spmm_csr_rowsplit");
    return DG_OK;
  }
  const int64_t blocks = (warps + 7) / 8;
  if (blocks > 0x7fffffffLL) { set_error("spmm: grid too large"); return DG_ERR_INVALID_ARGUMENT; }
  int pf = (flags & DG_SPMM_PREFETCH) ? spmm_prefetch_distance() : 0;
  const int l1 = (NCHUNK >= 3 && pf > 0) ? spmm_prefetch_l1() : 0;     // wide rows of out-of-L2 operands: prefetch into L1
  if (l1 > 0) pf = l1;
  if (pf * kUnroll > 32 - kUnroll) pf = (32 - kUnroll) / kUnroll;      // the window stays within one batch + the next
#define DG_SPMM_LAUNCH(W, P)                                                                                   \
  spmm_csr_kernel<L, NCHUNK, kUnroll, W, kMinBlocks, P><<<static_cast<unsigned>(blocks), 256, 0, st>>>(          \
      indptr, indices, vals, src_scale, dst_scale, bias, x, ldx, out, ldo, n_rows, d, n_slabs, flags, pf)
  const bool weighted = vals || src_scale;
  if (pf > 0 && l1 > 0) {
    if (weighted) DG_SPMM_LAUNCH(true, 2); else DG_SPMM_LAUNCH(false, 2);
  } else if (pf > 0) {
    if (weighted) DG_SPMM_LAUNCH(true, 1); else DG_SPMM_LAUNCH(false, 1);
  } else {
    if (weighted) DG_SPMM_LAUNCH(true, 0); else DG_SPMM_LAUNCH(false, 0);
  }
#undef DG_SPMM_LAUNCH
  DG_CHECK_LAUNCH("spmm_csr");
  return DG_OK;
}

// Tuning knob for experiments only (scripts/spmm_bench.py): DG_SPMM_VARIANT=<n> picks another
// (unroll, resident blocks) point; unset = the measured best.
static int spmm_variant() {
  const char* e = getenv("DG_SPMM_VARIANT");
  return e ? atoi(e) : 0;
}

#define DG_SPMM_ARGS indptr, indices, vals, src_scale, dst_scale, bias, x, ldx, out, ldo, n_rows, dd, flags, st

template <typename L>
static int spmm_dispatch(const int* indptr, const int* indices, const float* vals, const float* src_scale,
                         const float* dst_scale, const float* bias, const typename L::Elem* x, int64_t ldx, float* out,
                         int64_t ldo, int64_t n_rows, int64_t d, int flags, cudaStream_t st) {
  if (n_rows < 0 || d <= 0 || d > (1 << 20)) { set_error("spmm: bad n_rows / d"); return DG_ERR_INVALID_ARGUMENT; }
  if (d % L::kVec || ldx % L::kVec || ldo % 4 || ldx < d || ldo < d) {
    set_error("spmm: d=%lld ldx=%lld ldo=%lld must be multiples of the vector width and ld >= d",
              (long long)d, (long long)ldx, (long long)ldo);
    return DG_ERR_INVALID_ARGUMENT;
  }
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(bias)) & 15) {
    set_error("spmm: x / out / bias must be 16-byte aligned");
    return DG_ERR_INVALID_ARGUMENT;
  }
  if (n_rows == 0) return DG_OK;
  if (sizeof(typename L::Elem) == 4 && !(flags & DG_SPMM_ROWSPLIT)) {
    const int mode = spmm_stage_mode();          // bit 0: wide rows (d > 256), bit 1: narrow rows (d <= 128)
    const float* xf = reinterpret_cast<const float*>(x);
    if ((mode & 1) && d > 256) {
#define DG_STAGE_ARGS indptr, indices, vals, src_scale, dst_scale, bias, xf, ldx, out, ldo, n_rows, static_cast<int>(d), flags, st
      switch (spmm_stage_ring()) {
        case 8: return launch_spmm_stage<3, 8, 2>(DG_STAGE_ARGS);
        case 16: return launch_spmm_stage<3, 16, 1>(DG_STAGE_ARGS);
        default: return launch_spmm_stage<3, 4, 4>(DG_STAGE_ARGS);
      }
#undef DG_STAGE_ARGS
    }
    if ((mode & 2) && d <= 128)
      return launch_spmm_stage<1, 16, 3>(indptr, indices, vals, src_scale, dst_scale, bias, xf, ldx, out, ldo, n_rows,
                                         static_cast<int>(d), flags, st);
  }
  // pick the slab shape with the least lane waste: d <= 32V -> 1 chunk, <= 64V -> 2, else 3-chunk slabs
  const int per = 32 * L::kVec;
  const int dd = static_cast<int>(d);
  const int var = spmm_variant();
  if (dd <= per) {
    switch (var) {
      case 1: return launch_spmm<L, 1, 4, 6>(DG_SPMM_ARGS);
      case 2: return launch_spmm<L, 1, 8, 4>(DG_SPMM_ARGS);
      case 3: return launch_spmm<L, 1, 16, 2>(DG_SPMM_ARGS);
      default: return launch_spmm<L, 1, 8, 3>(DG_SPMM_ARGS);
    }
  }
  if (dd <= 2 * per || (dd > 3 * per && dd % (3 * per) && dd % (2 * per) == 0)) {
    switch (var) {
      case 1: return launch_spmm<L, 2, 2, 5>(DG_SPMM_ARGS);
      case 2: return launch_spmm<L, 2, 4, 4>(DG_SPMM_ARGS);
      default: return launch_spmm<L, 2, 4, 3>(DG_SPMM_ARGS);
    }
  }
  switch (var) {
    case 1: return launch_spmm<L, 3, 2, 4>(DG_SPMM_ARGS);
    case 2: return launch_spmm<L, 3, 3, 3>(DG_SPMM_ARGS);
    case 3: return launch_spmm<L, 3, 4, 3>(DG_SPMM_ARGS);
    case 4: return launch_spmm<L, 3, 1, 6>(DG_SPMM_ARGS);
    default: return launch_spmm<L, 3, 4, 2>(DG_SPMM_ARGS);
  }
}

}  // namespace dg

extern "C" {
int dg_spmm_csr_f32(const int32_t* indptr, const int32_t* indices, const float* vals, const float* src_scale,
                    const float* dst_scale, const float* bias, const float* x, int64_t ldx, float* out, int64_t ldo,
                    int64_t n_rows, int64_t d, int flags, dg_stream_t stream) {
  return dg::spmm_dispatch<dg::LoadF32>(indptr, indices, vals, src_scale, dst_scale, bias, x, ldx, out, ldo, n_rows, d,
                                        flags, dg::as_stream(stream));
}
int dg_spmm_csr_bf16(const int32_t* indptr, const int32_t* indices, const float* vals, const float* src_scale,
                     const float* dst_scale, const float* bias, const void* x_bf16, int64_t ldx, float* out,
                     int64_t ldo, int64_t n_rows, int64_t d, int flags, dg_stream_t stream) {
  return dg::spmm_dispatch<dg::LoadBF16>(indptr, indices, vals, src_scale, dst_scale, bias,
                                         static_cast<const __nv_bfloat16*>(x_bf16), ldx, out, ldo, n_rows, d, flags,
                                         dg::as_stream(stream));
}
}
