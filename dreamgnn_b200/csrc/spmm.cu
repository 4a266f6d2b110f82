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
  __device__ static __forceinline__ void unpack(const Raw& t, float (&v)[kVec]) {
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
};

struct LoadBF16 {
  static constexpr int kVec = 8;
  using Elem = __nv_bfloat16;
  using Raw = uint4;
  __device__ static __forceinline__ Raw load(const Elem* p) { return ldg_u4_stream(reinterpret_cast<const uint4*>(p)); }
  __device__ static __forceinline__ void unpack(const Raw& t, float (&v)[kVec]) {
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {   // bf16 -> fp32 is a 16-bit shift
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};

template <typename L, int NCHUNK, int kUnroll, bool kWeighted, int kMinBlocks>
__global__ void __launch_bounds__(256, kMinBlocks)
spmm_csr_kernel(const int* __restrict__ indptr, const int* __restrict__ indices, const float* __restrict__ vals,
                const float* __restrict__ src_scale, const float* __restrict__ dst_scale,
                const float* __restrict__ bias, const typename L::Elem* __restrict__ x, int64_t ldx,
                float* __restrict__ out, int64_t ldo, int64_t n_rows, int d, int n_slabs, int flags) {
  constexpr int V = L::kVec;
  constexpr int kSlabCols = 32 * V * NCHUNK;
  const int lane = lane_id();
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t row = warp / n_slabs;
  if (row >= n_rows) return;
  const int slab = static_cast<int>(warp - row * n_slabs);
  const int col0 = slab * kSlabCols + lane * V;     // chunk c covers columns col0 + c*32*V .. +V

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
    for (; t + kUnroll <= cnt; t += kUnroll) {
      typename L::Raw buf[kUnroll][NCHUNK];   // raw 128-bit vectors; unpacked only at FMA time
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

template <typename L, int NCHUNK, int kUnroll, int kMinBlocks>
static int launch_spmm(const int* indptr, const int* indices, const float* vals, const float* src_scale,
                       const float* dst_scale, const float* bias, const typename L::Elem* x, int64_t ldx, float* out,
                       int64_t ldo, int64_t n_rows, int d, int flags, cudaStream_t st) {
  constexpr int kSlabCols = 32 * L::kVec * NCHUNK;
  const int n_slabs = (d + kSlabCols - 1) / kSlabCols;
  const int64_t warps = n_rows * n_slabs;
  const int64_t blocks = (warps + 7) / 8;
  if (blocks > 0x7fffffffLL) { set_error("spmm: grid too large"); return DG_ERR_INVALID_ARGUMENT; }
  if (vals || src_scale)
    spmm_csr_kernel<L, NCHUNK, kUnroll, true, kMinBlocks><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
        indptr, indices, vals, src_scale, dst_scale, bias, x, ldx, out, ldo, n_rows, d, n_slabs, flags);
  else
    spmm_csr_kernel<L, NCHUNK, kUnroll, false, kMinBlocks><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
        indptr, indices, vals, src_scale, dst_scale, bias, x, ldx, out, ldo, n_rows, d, n_slabs, flags);
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
