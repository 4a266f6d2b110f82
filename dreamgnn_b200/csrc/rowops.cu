// Row-streaming helpers around the message-passing kernels (all HBM-bound, one pass over the matrix):
//   * colsum:            deterministic column sums (bias gradients of GraphConvolution / nn.Linear, the column means
//                        of common_loss), optionally fused with the ReLU-backward mask of the aggregation epilogue
//   * center_normalize:  rows of common_loss (utils.py:87-95): centre by the column mean, L2-normalise, widen to fp64
//   * its backward
// Layout: warp per row-slice, lane = one float4 of a 128-column chunk, so every warp load is one 512-byte line group.
#include "common.cuh"

namespace dg {
namespace {

constexpr int kColChunk = 128;    // columns per CTA column-chunk: 32 lanes x float4
constexpr int kColsumWarps = 8;
constexpr int64_t kColsumOneLaunchRows = 16384;   // below this the finishing launch costs more than it computes

__host__ __device__ inline int colsum_slabs(int64_t n_rows, int64_t d) {
  const int chunks = static_cast<int>((d + kColChunk - 1) / kColChunk);
  int64_t s = (2 * kNumSM + chunks - 1) / chunks;                       // ~2 CTAs per SM over all chunks
  const int64_t max_s = (n_rows + kColsumWarps - 1) / kColsumWarps;      // at least one row per warp
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  return static_cast<int>(s);
}

// partial[slab][col] = sum over the slab's rows (warps interleaved, then summed in warp order) of y[row][col],
// y = x or x * (gate > 0); y is optionally written back (the masked gradient the transposed SpMM gathers next).
template <bool kGate, bool kWrite>
__global__ void __launch_bounds__(kColsumWarps * 32)
colsum_partial_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ gate, int64_t ldg,
                      float* __restrict__ y, int64_t ldy, int64_t n_rows, int d, float* __restrict__ partial,
                      int* __restrict__ tickets, float* __restrict__ out) {
  __shared__ float4 red[kColsumWarps][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = blockIdx.x * kColChunk + lane * 4;
  const bool active = col < d;
  const int slabs = gridDim.y;
  const int64_t per = (n_rows + slabs - 1) / slabs;
  const int64_t r0 = per * blockIdx.y;
  const int64_t r1 = (r0 + per < n_rows) ? r0 + per : n_rows;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (active) {
    int64_t r = r0 + warp;
    // four independent rows in flight per lane
    for (; r + 3 * kColsumWarps < r1; r += 4 * kColsumWarps) {
      float4 v[4], g[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = ldg_f4_stream(reinterpret_cast<const float4*>(x + (r + u * kColsumWarps) * ldx + col));
      if (kGate) {
#pragma unroll
        for (int u = 0; u < 4; ++u) g[u] = ldg_f4_stream(reinterpret_cast<const float4*>(gate + (r + u * kColsumWarps) * ldg + col));
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          v[u].x = g[u].x > 0.f ? v[u].x : 0.f; v[u].y = g[u].y > 0.f ? v[u].y : 0.f;
          v[u].z = g[u].z > 0.f ? v[u].z : 0.f; v[u].w = g[u].w > 0.f ? v[u].w : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (kWrite) *reinterpret_cast<float4*>(y + (r + u * kColsumWarps) * ldy + col) = v[u];
        acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
      }
    }
    for (; r < r1; r += kColsumWarps) {
      float4 v = ldg_f4_stream(reinterpret_cast<const float4*>(x + r * ldx + col));
      if (kGate) {
        const float4 g = ldg_f4_stream(reinterpret_cast<const float4*>(gate + r * ldg + col));
        v.x = g.x > 0.f ? v.x : 0.f; v.y = g.y > 0.f ? v.y : 0.f; v.z = g.z > 0.f ? v.z : 0.f; v.w = g.w > 0.f ? v.w : 0.f;
      }
      if (kWrite) *reinterpret_cast<float4*>(y + r * ldy + col) = v;
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  red[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && active) {
    float4 s = red[0][lane];
#pragma unroll
    for (int w = 1; w < kColsumWarps; ++w) {
      const float4 t = red[w][lane];
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    *reinterpret_cast<float4*>(partial + static_cast<int64_t>(blockIdx.y) * d + col) = s;
  }
  if (tickets == nullptr) return;                    // two-launch form: colsum_finish_kernel adds the slabs
  // One-launch form: the CTA that finishes LAST among this column chunk's slabs adds the slab partials -- in slab order,
  // so the result does not depend on which CTA that is. (Matrices of a few hundred rows: the second launch costs more
  // than both kernels' work, on a dependency chain that has ~40 of these per training iteration.)
  __shared__ int is_last;
  __shared__ double fin[2][kColChunk];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int t = atomicAdd(tickets + blockIdx.x, 1);
    is_last = (t == static_cast<int>(gridDim.y) - 1);
    if (is_last) tickets[blockIdx.x] = 0;            // zero again for the next call / graph replay
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const int c = threadIdx.x & (kColChunk - 1), half = threadIdx.x >> 7;       // 256 threads: 128 columns x 2 slab ranges
  const int ccol = blockIdx.x * kColChunk + c;
  const int per_half = (slabs + 1) / 2;
  const int s0 = half * per_half, s1 = min(slabs, s0 + per_half);
  double a[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (ccol < d) {
    int i = s0;
    for (; i + 7 < s1; i += 8) {                      // eight independent L2 loads in flight
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldcg(partial + static_cast<int64_t>(i + u) * d + ccol);
#pragma unroll
      for (int u = 0; u < 8; ++u) a[u] += v[u];
    }
    for (; i < s1; ++i) a[0] += __ldcg(partial + static_cast<int64_t>(i) * d + ccol);
  }
  fin[half][c] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  __syncthreads();
  if (half == 0 && ccol < d) out[ccol] = static_cast<float>(fin[0][c] + fin[1][c]);
}

// out[col] = sum of the slab partials in a fixed order: 32 warps each sum a contiguous range of slabs for 32 columns
// (every load independent, double accumulators), then the 32 range sums are added in range order.
constexpr int kFinishGroups = 32;
__global__ void __launch_bounds__(32 * kFinishGroups)
colsum_finish_kernel(const float* __restrict__ partial, int slabs, int d, float* __restrict__ out) {
  __shared__ double red[kFinishGroups][33];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + lane;
  const int per = (slabs + kFinishGroups - 1) / kFinishGroups;
  const int s0 = grp * per, s1 = min(slabs, s0 + per);
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  if (col < d) {
    int i = s0;
    for (; i + 3 < s1; i += 4) {
      const float v0 = partial[static_cast<int64_t>(i) * d + col], v1 = partial[static_cast<int64_t>(i + 1) * d + col];
      const float v2 = partial[static_cast<int64_t>(i + 2) * d + col], v3 = partial[static_cast<int64_t>(i + 3) * d + col];
      a0 += v0; a1 += v1; a2 += v2; a3 += v3;
    }
    for (; i < s1; ++i) a0 += partial[static_cast<int64_t>(i) * d + col];
  }
  red[grp][lane] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (grp == 0 && col < d) {
    double s = red[0][lane];
#pragma unroll
    for (int g = 1; g < kFinishGroups; ++g) s += red[g][lane];
    out[col] = static_cast<float>(s);
  }
}

// z[row,:] = (x[row,:] - colsum/n) * inv, inv = 1 / max(||x[row,:] - colsum/n||_2, eps); warp per row
__global__ void __launch_bounds__(256)
center_normalize_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ colsum, int64_t n_rows, int d,
                        double eps, double* __restrict__ z, int64_t ldz, double* __restrict__ inv_norm) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const float inv_n = 1.0f / static_cast<float>(n_rows);
  const float* xr = x + row * ldx;
  float ss = 0.f;
  for (int c = lane * 4; c < d; c += kColChunk) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    const float4 m = *reinterpret_cast<const float4*>(colsum + c);
    const float a = v.x - m.x * inv_n, b = v.y - m.y * inv_n, e = v.z - m.z * inv_n, f = v.w - m.w * inv_n;
    ss += a * a + b * b + e * e + f * f;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(kFull, ss, o);
  const double nrm = sqrt(static_cast<double>(ss));
  const double inv = 1.0 / (nrm > eps ? nrm : eps);
  double* zr = z + row * ldz;
  for (int c = lane * 4; c < d; c += kColChunk) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    const float4 m = *reinterpret_cast<const float4*>(colsum + c);
    double2 lo, hi;
    lo.x = static_cast<double>(v.x - m.x * inv_n) * inv; lo.y = static_cast<double>(v.y - m.y * inv_n) * inv;
    hi.x = static_cast<double>(v.z - m.z * inv_n) * inv; hi.y = static_cast<double>(v.w - m.w * inv_n) * inv;
    *reinterpret_cast<double2*>(zr + c) = lo;
    *reinterpret_cast<double2*>(zr + c + 2) = hi;
  }
  if (lane == 0) inv_norm[row] = inv;
}

// gradient w.r.t. the centred row c (before the mean is subtracted again): dc = (gz - z (z . gz)) * inv, fp32 out
__global__ void __launch_bounds__(256)
center_normalize_bwd_kernel(const double* __restrict__ gz, int64_t ldgz, const double* __restrict__ z, int64_t ldz,
                            const double* __restrict__ inv_norm, int64_t n_rows, int d, float* __restrict__ dc, int64_t lddc,
                            const float* __restrict__ gout, double scale) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const double* gr = gz + row * ldgz;
  const double* zr = z + row * ldz;
  double dot = 0.0;
  for (int c = lane * 4; c < d; c += kColChunk) {
    const double2 g0 = *reinterpret_cast<const double2*>(gr + c), g1 = *reinterpret_cast<const double2*>(gr + c + 2);
    const double2 z0 = *reinterpret_cast<const double2*>(zr + c), z1 = *reinterpret_cast<const double2*>(zr + c + 2);
    dot += g0.x * z0.x + g0.y * z0.y + g1.x * z1.x + g1.y * z1.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(kFull, dot, o);
  // the upstream scalar gradient and the loss's constant factor ride along (the product is linear in gz)
  const double inv = inv_norm[row] * (gout ? static_cast<double>(__ldg(gout)) * scale : scale);
  float* dr = dc + row * lddc;
  for (int c = lane * 4; c < d; c += kColChunk) {
    const double2 g0 = *reinterpret_cast<const double2*>(gr + c), g1 = *reinterpret_cast<const double2*>(gr + c + 2);
    const double2 z0 = *reinterpret_cast<const double2*>(zr + c), z1 = *reinterpret_cast<const double2*>(zr + c + 2);
    float4 o;
    o.x = static_cast<float>((g0.x - z0.x * dot) * inv); o.y = static_cast<float>((g0.y - z0.y * dot) * inv);
    o.z = static_cast<float>((g1.x - z1.x * dot) * inv); o.w = static_cast<float>((g1.y - z1.y * dot) * inv);
    *reinterpret_cast<float4*>(dr + c) = o;
  }
}

inline bool aligned16(const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; }

}  // namespace
}  // namespace dg

extern "C" {

size_t dg_colsum_workspace_bytes(int64_t n_rows, int64_t d) {
  if (n_rows <= 0 || d <= 0) return 256;
  return dg::ws_add(0, static_cast<size_t>(dg::colsum_slabs(n_rows, d)) * static_cast<size_t>(d) * sizeof(float));
}

int dg_colsum_f32(const float* x, int64_t ldx, const float* gate, int64_t ldg, float* y, int64_t ldy, int64_t n_rows,
                  int64_t d, float* out, void* workspace, size_t workspace_bytes, int32_t* tickets, dg_stream_t stream) {
  using namespace dg;
  DG_REQUIRE(n_rows >= 0 && d > 0 && d <= (1 << 20), "bad shape");
  DG_REQUIRE(x != nullptr && out != nullptr, "null pointer");
  DG_REQUIRE(d % 4 == 0 && ldx % 4 == 0 && ldx >= d && aligned16(x) && aligned16(out), "x rows must be 16-byte aligned, d % 4 == 0");
  DG_REQUIRE(gate == nullptr || (ldg % 4 == 0 && ldg >= d && aligned16(gate)), "gate rows must be 16-byte aligned");
  DG_REQUIRE(y == nullptr || (ldy % 4 == 0 && ldy >= d && aligned16(y)), "y rows must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  if (n_rows == 0) {
    DG_CHECK_CUDA(cudaMemsetAsync(out, 0, static_cast<size_t>(d) * sizeof(float), st));
    return DG_OK;
  }
  int slabs = colsum_slabs(n_rows, d);
  // one launch (last CTA adds the slabs) when the caller lends zeroed tickets and the matrix is small enough for the
  // launch, not the bandwidth, to be what costs; then with at least four rows per warp, so that few slabs are left to add
  int* tk = (tickets != nullptr && n_rows <= kColsumOneLaunchRows) ? tickets : nullptr;
  if (tk != nullptr) {
    const int64_t cap = (n_rows + 4 * kColsumWarps - 1) / (4 * kColsumWarps);
    if (slabs > cap) slabs = static_cast<int>(cap < 1 ? 1 : cap);
  }
  Workspace ws(workspace, workspace_bytes);
  float* partial = ws.take<float>(static_cast<size_t>(slabs) * d);
  if (partial == nullptr) {
    set_error("dg_colsum_f32: workspace too small");
    return DG_ERR_WORKSPACE_TOO_SMALL;
  }
  const dim3 grid(static_cast<unsigned>((d + kColChunk - 1) / kColChunk), static_cast<unsigned>(slabs));
  const int di = static_cast<int>(d);
  if (gate != nullptr && y != nullptr)
    colsum_partial_kernel<true, true><<<grid, kColsumWarps * 32, 0, st>>>(x, ldx, gate, ldg, y, ldy, n_rows, di, partial, tk, out);
  else if (gate != nullptr)
    colsum_partial_kernel<true, false><<<grid, kColsumWarps * 32, 0, st>>>(x, ldx, gate, ldg, nullptr, 0, n_rows, di, partial, tk, out);
  else if (y != nullptr)
    colsum_partial_kernel<false, true><<<grid, kColsumWarps * 32, 0, st>>>(x, ldx, nullptr, 0, y, ldy, n_rows, di, partial, tk, out);
  else
    colsum_partial_kernel<false, false><<<grid, kColsumWarps * 32, 0, st>>>(x, ldx, nullptr, 0, nullptr, 0, n_rows, di, partial, tk, out);
  DG_CHECK_LAUNCH("colsum_partial");
  if (tk != nullptr) return DG_OK;
  colsum_finish_kernel<<<static_cast<unsigned>((d + 31) / 32), 32 * kFinishGroups, 0, st>>>(partial, slabs, di, out);
  DG_CHECK_LAUNCH("colsum_finish");
  return DG_OK;
}

int dg_center_normalize_f64(const float* x, int64_t ldx, const float* colsum, int64_t n_rows, int64_t d, double eps,
                            double* z, int64_t ldz, double* inv_norm, dg_stream_t stream) {
  using namespace dg;
  DG_REQUIRE(n_rows >= 0 && d > 0, "bad shape");
  if (n_rows == 0) return DG_OK;
  DG_REQUIRE(x && colsum && z && inv_norm, "null pointer");
  DG_REQUIRE(d % 4 == 0 && ldx % 4 == 0 && ldx >= d && aligned16(x) && aligned16(colsum), "x rows must be 16-byte aligned, d % 4 == 0");
  DG_REQUIRE(ldz % 2 == 0 && ldz >= d && aligned16(z), "z rows must be 16-byte aligned");
  const unsigned blocks = static_cast<unsigned>((n_rows + 7) / 8);
  center_normalize_kernel<<<blocks, 256, 0, as_stream(stream)>>>(x, ldx, colsum, n_rows, static_cast<int>(d), eps, z, ldz, inv_norm);
  DG_CHECK_LAUNCH("center_normalize");
  return DG_OK;
}

int dg_center_normalize_bwd_f64(const double* gz, int64_t ldgz, const double* z, int64_t ldz, const double* inv_norm,
                                int64_t n_rows, int64_t d, float* dc, int64_t lddc, const float* gout, double scale,
                                dg_stream_t stream) {
  using namespace dg;
  DG_REQUIRE(n_rows >= 0 && d > 0, "bad shape");
  if (n_rows == 0) return DG_OK;
  DG_REQUIRE(gz && z && inv_norm && dc, "null pointer");
  DG_REQUIRE(d % 4 == 0 && ldgz % 2 == 0 && ldz % 2 == 0 && ldgz >= d && ldz >= d && aligned16(gz) && aligned16(z),
             "gz / z rows must be 16-byte aligned, d % 4 == 0");
  DG_REQUIRE(lddc % 4 == 0 && lddc >= d && aligned16(dc), "dc rows must be 16-byte aligned");
  const unsigned blocks = static_cast<unsigned>((n_rows + 7) / 8);
  center_normalize_bwd_kernel<<<blocks, 256, 0, as_stream(stream)>>>(gz, ldgz, z, ldz, inv_norm, n_rows, static_cast<int>(d), dc, lddc,
                                                                      gout, scale);
  DG_CHECK_LAUNCH("center_normalize_bwd");
  return DG_OK;
}

}  // extern "C"
