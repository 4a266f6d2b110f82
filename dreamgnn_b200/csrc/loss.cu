// The two scalar losses of an iteration (train.py:286-294) as explicit kernels:
//
//   * binary cross entropy with logits over the scored pairs (nn.BCEWithLogitsLoss / LabelSmoothingBCELoss,
//     train.py:15-23, 291): mean loss in two launches (per-CTA float64 partials, then one CTA adds them in order), its
//     gradient in one. ATen: log_sigmoid, fill, add, mul, add, mean + sigmoid, add, mul, mul = 10 launches.
//   * the tail of common_loss in Gram form (utils.py:87-95, ops.GramCommonLoss): from G = Z^T Z, the matrix G * S
//     (S = +1 on the diagonal blocks, -1 off them) and loss = sum(G * G * S) / n^2 in ONE single-CTA launch. ATen: clone,
//     two negations, multiply, a single-CTA float64 reduction (20 us), two divisions, a cast = 8 launches.
//
// All sums run in a fixed order (deterministic), in float64.
#include <math.h>

#include "common.cuh"

namespace dg {
namespace {

constexpr int kLossThreads = 256;
constexpr int kBceMaxCtas = 4 * kNumSM;

__device__ __forceinline__ double loss_block_sum(double v, double* red /* [kLossThreads / 32] */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < kLossThreads / 32; ++w) s += red[w];
  }
  return s;                                            // valid in thread 0
}

__device__ __forceinline__ float smooth_target(float t, float smoothing) {
  return smoothing > 0.f ? t * (1.f - smoothing) + 0.5f * smoothing : t;
}

// partial[cta] = sum over the CTA's elements (grid-stride, fixed) of  max(x, 0) - x t + log1p(exp(-|x|))
__global__ void __launch_bounds__(kLossThreads)
bce_partial_kernel(const float* __restrict__ x, const float* __restrict__ target, int64_t n, float smoothing,
                   double* __restrict__ partial) {
  __shared__ double red[kLossThreads / 32];
  double acc = 0.0;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kLossThreads;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kLossThreads + threadIdx.x; i < n; i += stride) {
    const float xv = __ldg(x + i), t = smooth_target(__ldg(target + i), smoothing);
    acc += static_cast<double>(fmaxf(xv, 0.f) - xv * t + log1pf(expf(-fabsf(xv))));
  }
  const double s = loss_block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void __launch_bounds__(kLossThreads)
bce_finish_kernel(const double* __restrict__ partial, int n_partial, int64_t n, float* __restrict__ loss) {
  __shared__ double red[kLossThreads / 32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n_partial; i += kLossThreads) acc += partial[i];
  const double s = loss_block_sum(acc, red);
  if (threadIdx.x == 0) *loss = static_cast<float>(s / static_cast<double>(n));
}

// dx = gout * (sigmoid(x) - t) / n
__global__ void __launch_bounds__(kLossThreads)
bce_grad_kernel(const float* __restrict__ x, const float* __restrict__ target, int64_t n, float smoothing,
                const float* __restrict__ gout, float* __restrict__ dx) {
  const float scale = __ldg(gout) / static_cast<float>(n);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kLossThreads;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kLossThreads + threadIdx.x; i < n; i += stride) {
    const float xv = __ldg(x + i), t = smooth_target(__ldg(target + i), smoothing);
    const float sg = 1.f / (1.f + expf(-xv));
    dx[i] = (sg - t) * scale;
  }
}

int bce_grid(int64_t n) {
  int64_t g = (n + kLossThreads * 4 - 1) / (kLossThreads * 4);
  if (g > kBceMaxCtas) g = kBceMaxCtas;
  return static_cast<int>(g < 1 ? 1 : g);
}

// gs = G * S, loss = sum(G * gs) / n^2 for G [2d, 2d] float64 (row-major, leading dimension ldg). One CTA (the sum must
// end in one place and the matrix is 512 KB at d = 128): 16 independent loads in flight per thread -- with one element
// per trip the kernel is a chain of 64 dependent L2 round trips (37 us measured; this: ~6).
constexpr int kGramThreads = 1024;
constexpr int kGramUnroll = 16;
__global__ void __launch_bounds__(kGramThreads)
gram_loss_kernel(const double* __restrict__ G, int64_t ldg, int d, double n_rows, double* __restrict__ gs, float* __restrict__ loss) {
  __shared__ double red[kGramThreads / 32];
  const int w = 2 * d, n = w * w;
  double acc = 0.0;
  for (int base = 0; base < n; base += kGramThreads * kGramUnroll) {
    double g[kGramUnroll];
#pragma unroll
    for (int u = 0; u < kGramUnroll; ++u) {
      const int i = base + u * kGramThreads + threadIdx.x;
      const int r = i / w, c = i - r * w;
      g[u] = i < n ? G[static_cast<int64_t>(r) * ldg + c] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < kGramUnroll; ++u) {
      const int i = base + u * kGramThreads + threadIdx.x;
      if (i < n) {
        const int r = i / w, c = i - r * w;
        const double s = ((r < d) == (c < d)) ? g[u] : -g[u];
        gs[i] = s;                                     // gs is dense [2d, 2d]
        acc += g[u] * s;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int k = 0; k < kGramThreads / 32; ++k) s += red[k];
    *loss = static_cast<float>(s / n_rows / n_rows);
  }
}

}  // namespace
}  // namespace dg

extern "C" {

size_t dg_bce_logits_workspace_bytes(int64_t n) { return dg::ws_add(0, static_cast<size_t>(dg::bce_grid(n)) * sizeof(double)); }

int dg_bce_logits_fwd_f32(const float* logits, const float* target, int64_t n, float smoothing, float* loss, void* workspace,
                          size_t workspace_bytes, dg_stream_t stream) {
  using namespace dg;
  DG_REQUIRE(n > 0 && logits && target && loss, "bad arguments");
  DG_REQUIRE(smoothing >= 0.f && smoothing <= 1.f, "smoothing outside [0, 1]");
  Workspace ws(workspace, workspace_bytes);
  const int grid = bce_grid(n);
  double* partial = ws.take<double>(static_cast<size_t>(grid));
  if (!partial) {
    set_error("dg_bce_logits_fwd_f32: workspace too small");
    return DG_ERR_WORKSPACE_TOO_SMALL;
  }
  cudaStream_t st = as_stream(stream);
  bce_partial_kernel<<<grid, kLossThreads, 0, st>>>(logits, target, n, smoothing, partial);
  DG_CHECK_LAUNCH("bce_partial");
  bce_finish_kernel<<<1, kLossThreads, 0, st>>>(partial, grid, n, loss);
  DG_CHECK_LAUNCH("bce_finish");
  return DG_OK;
}

int dg_bce_logits_bwd_f32(const float* logits, const float* target, int64_t n, float smoothing, const float* gout,
                          float* dlogits, dg_stream_t stream) {
  using namespace dg;
  DG_REQUIRE(n > 0 && logits && target && gout && dlogits, "bad arguments");
  bce_grad_kernel<<<bce_grid(n), kLossThreads, 0, as_stream(stream)>>>(logits, target, n, smoothing, gout, dlogits);
  DG_CHECK_LAUNCH("bce_grad");
  return DG_OK;
}

int dg_gram_common_loss_f64(const double* G, int64_t ldg, int64_t d, double n_rows, double* gs, float* loss, dg_stream_t stream) {
  using namespace dg;
  DG_REQUIRE(G && gs && loss, "null pointer");
  DG_REQUIRE(d > 0 && d <= 4096 && ldg >= 2 * d && n_rows > 0, "bad shape");
  gram_loss_kernel<<<1, kGramThreads, 0, as_stream(stream)>>>(G, ldg, static_cast<int>(d), n_rows, gs, loss);
  DG_CHECK_LAUNCH("gram_loss");
  return DG_OK;
}

}  // extern "C"
