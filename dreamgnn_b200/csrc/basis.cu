// Basis decomposition of the GCMC relation weights (layers.py:120-121):  W_r = sum_b att[r][b] * basis[b].
//
// The reference writes it as matmul(att [R, B], basis.view(B, in * msg)). With R = B = 2 that GEMM is degenerate both
// ways: the library runs the forward as a GEMV-like kernel and the backward's d_att = dW . basis^T as a 2 x 2 output over
// K = in * msg = 262 k with a 147-way split-K (56 us + 10 us reduction at the very end of every backward, nothing left
// to overlap). Here it is what it is -- an elementwise weighted sum, and four dot products:
//
//   forward   W[r][i][0..Dp) = sum_b att[r][b] * basis[b][i][0..D), zero in the padded columns [D, Dp): the message width
//             padding (341 -> 344) the aggregation kernels need is written here instead of by a separate pad copy
//   backward  dbasis[b] = sum_r att[r][b] * dW[r]   (same elementwise pass, reading the padded gradient as stored)
//             datt[r][b] = <dW[r], basis[b]>        per-thread fp32 over a few elements, then float64 by warp / CTA / grid
//                                                   in a fixed order (deterministic)
#include "common.cuh"

namespace dg {
namespace {

constexpr int kBasisMax = 4;            // R, B <= 4
constexpr int kBasisThreads = 256;
constexpr int kBasisMaxCtas = 2 * kNumSM;

struct BasisAtt { float a[kBasisMax][kBasisMax]; };

__global__ void __launch_bounds__(kBasisThreads)
basis_combine_fwd_kernel(const float* __restrict__ att, const float* __restrict__ basis, int R, int B, int64_t rows, int D, int Dp,
                         float* __restrict__ W) {
  __shared__ BasisAtt s;
  if (threadIdx.x < kBasisMax * kBasisMax) {
    const int r = threadIdx.x / kBasisMax, b = threadIdx.x % kBasisMax;
    s.a[r][b] = (r < R && b < B) ? att[r * B + b] : 0.f;
  }
  __syncthreads();
  const int64_t n_out = rows * Dp, n_in = rows * D;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kBasisThreads;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * kBasisThreads + threadIdx.x; idx < n_out; idx += stride) {
    const int64_t i = idx / Dp;
    const int j = static_cast<int>(idx - i * Dp);
    float v[kBasisMax];
#pragma unroll
    for (int b = 0; b < kBasisMax; ++b) v[b] = (b < B && j < D) ? __ldg(basis + b * n_in + i * D + j) : 0.f;
#pragma unroll
    for (int r = 0; r < kBasisMax; ++r) {
      if (r < R) {
        float w = 0.f;
#pragma unroll
        for (int b = 0; b < kBasisMax; ++b) w = fmaf(s.a[r][b], v[b], w);
        W[r * n_out + idx] = w;
      }
    }
  }
}

__global__ void __launch_bounds__(kBasisThreads)
basis_combine_bwd_kernel(const float* __restrict__ att, const float* __restrict__ basis, const float* __restrict__ dW, int R, int B,
                         int64_t rows, int D, int Dp, float* __restrict__ dbasis, double* __restrict__ partial) {
  __shared__ BasisAtt s;
  __shared__ double red[kBasisThreads / 32][kBasisMax * kBasisMax];
  if (threadIdx.x < kBasisMax * kBasisMax) {
    const int r = threadIdx.x / kBasisMax, b = threadIdx.x % kBasisMax;
    s.a[r][b] = (r < R && b < B) ? att[r * B + b] : 0.f;
  }
  __syncthreads();
  const int64_t n_out = rows * Dp, n_in = rows * D;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kBasisThreads;
  float acc[kBasisMax][kBasisMax];
#pragma unroll
  for (int r = 0; r < kBasisMax; ++r)
#pragma unroll
    for (int b = 0; b < kBasisMax; ++b) acc[r][b] = 0.f;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * kBasisThreads + threadIdx.x; e < n_in; e += stride) {
    const int64_t i = e / D;
    const int j = static_cast<int>(e - i * D);
    float g[kBasisMax], v[kBasisMax];
#pragma unroll
    for (int r = 0; r < kBasisMax; ++r) g[r] = r < R ? __ldg(dW + r * n_out + i * Dp + j) : 0.f;
#pragma unroll
    for (int b = 0; b < kBasisMax; ++b) v[b] = b < B ? __ldg(basis + b * n_in + e) : 0.f;
#pragma unroll
    for (int b = 0; b < kBasisMax; ++b) {
      if (b < B) {
        float d = 0.f;
#pragma unroll
        for (int r = 0; r < kBasisMax; ++r) d = fmaf(s.a[r][b], g[r], d);
        dbasis[b * n_in + e] = d;
      }
    }
#pragma unroll
    for (int r = 0; r < kBasisMax; ++r)
#pragma unroll
      for (int b = 0; b < kBasisMax; ++b) acc[r][b] = fmaf(g[r], v[b], acc[r][b]);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < kBasisMax; ++r)
#pragma unroll
    for (int b = 0; b < kBasisMax; ++b) {
      double x = static_cast<double>(acc[r][b]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
      if (lane == 0) red[warp][r * kBasisMax + b] = x;
    }
  __syncthreads();
  if (threadIdx.x < kBasisMax * kBasisMax) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kBasisThreads / 32; ++w) t += red[w][threadIdx.x];
    partial[static_cast<int64_t>(blockIdx.x) * (kBasisMax * kBasisMax) + threadIdx.x] = t;
  }
}

__global__ void basis_att_finish_kernel(const double* __restrict__ partial, int n_ctas, int R, int B, float* __restrict__ datt) {
  const int r = threadIdx.x / kBasisMax, b = threadIdx.x % kBasisMax;
  if (threadIdx.x < kBasisMax * kBasisMax && r < R && b < B) {
    double t = 0.0;
    for (int c = 0; c < n_ctas; ++c) t += partial[static_cast<int64_t>(c) * (kBasisMax * kBasisMax) + threadIdx.x];
    datt[r * B + b] = static_cast<float>(t);
  }
}

int basis_grid(int64_t n) {
  int64_t g = (n + kBasisThreads * 4 - 1) / (kBasisThreads * 4);
  if (g > kBasisMaxCtas) g = kBasisMaxCtas;
  return static_cast<int>(g < 1 ? 1 : g);
}

}  // namespace
}  // namespace dg

extern "C" {

int dg_basis_combine_fwd_f32(const float* att, const float* basis, int n_rel, int n_basis, int64_t rows, int64_t d, int64_t d_pad,
                             float* w, dg_stream_t stream) {
  using namespace dg;
  DG_REQUIRE(att && basis && w, "null pointer");
  DG_REQUIRE(n_rel >= 1 && n_rel <= kBasisMax && n_basis >= 1 && n_basis <= kBasisMax, "at most 4 relations / bases");
  DG_REQUIRE(rows > 0 && d > 0 && d_pad >= d && d_pad < (1 << 30), "bad shape");
  basis_combine_fwd_kernel<<<basis_grid(rows * d_pad), kBasisThreads, 0, as_stream(stream)>>>(
      att, basis, n_rel, n_basis, rows, static_cast<int>(d), static_cast<int>(d_pad), w);
  DG_CHECK_LAUNCH("basis_combine_fwd");
  return DG_OK;
}

size_t dg_basis_combine_bwd_workspace_bytes(int64_t rows, int64_t d) {
  return dg::ws_add(0, static_cast<size_t>(dg::basis_grid(rows * d)) * dg::kBasisMax * dg::kBasisMax * sizeof(double));
}

int dg_basis_combine_bwd_f32(const float* att, const float* basis, const float* dw, int n_rel, int n_basis, int64_t rows,
                             int64_t d, int64_t d_pad, float* dbasis, float* datt, void* workspace, size_t workspace_bytes,
                             dg_stream_t stream) {
  using namespace dg;
  DG_REQUIRE(att && basis && dw && dbasis && datt, "null pointer");
  DG_REQUIRE(n_rel >= 1 && n_rel <= kBasisMax && n_basis >= 1 && n_basis <= kBasisMax, "at most 4 relations / bases");
  DG_REQUIRE(rows > 0 && d > 0 && d_pad >= d && d_pad < (1 << 30), "bad shape");
  const int grid = basis_grid(rows * d);
  Workspace ws(workspace, workspace_bytes);
  double* partial = ws.take<double>(static_cast<size_t>(grid) * kBasisMax * kBasisMax);
  if (!partial) {
    set_error("dg_basis_combine_bwd_f32: workspace too small");
    return DG_ERR_WORKSPACE_TOO_SMALL;
  }
  cudaStream_t st = as_stream(stream);
  basis_combine_bwd_kernel<<<grid, kBasisThreads, 0, st>>>(att, basis, dw, n_rel, n_basis, rows, static_cast<int>(d),
                                                           static_cast<int>(d_pad), dbasis, partial);
  DG_CHECK_LAUNCH("basis_combine_bwd");
  basis_att_finish_kernel<<<1, 32, 0, st>>>(partial, grid, n_rel, n_basis, datt);
  DG_CHECK_LAUNCH("basis_att_finish");
  return DG_OK;
}

}  // extern "C"
