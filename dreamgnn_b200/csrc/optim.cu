// clip_grad_norm_ + Adam over the whole parameter list in two launches (train.py:297-300).
//
// torch runs this tail of the iteration as ~17 launches (per-tensor norms, their norm, the clip coefficient, the scaled
// gradients, and the foreach Adam update as eight multi-tensor passes): ~250 us at the end of every iteration with nothing
// to overlap -- 10 % of a step at the real-dataset shapes. Here:
//
//   adam_gradnorm_kernel   one CTA per 4 096-element chunk of one tensor: sum of squares of the gradient, threads in a
//                          fixed interleave, warps summed in warp order, float64 partial per chunk (no atomics)
//   adam_update_kernel     every CTA adds the partials in the same fixed order (they are L2-resident), derives the clip
//                          coefficient min(1, max_norm / (norm + 1e-6)) and applies the Adam update to its chunk
//
// The tensors (parameter, gradient, two moments, length) travel BY VALUE in the kernel arguments, up to kAdamBatch per
// launch: no device-side table to keep in sync with gradients that autograd re-allocates every iteration, and a captured
// CUDA graph bakes them in. Element order inside a chunk is t, t + 256, ...: plain coalesced 4-byte accesses, so the
// tensors need no alignment beyond their element size.
#include <math.h>

#include "common.cuh"

namespace dg {
namespace {

constexpr int kAdamBatch = DG_ADAM_MAX_TENSORS_PER_LAUNCH;   // 48
constexpr int kAdamThreads = 256;
constexpr int kAdamPerThread = 16;
constexpr int kAdamChunk = kAdamThreads * kAdamPerThread;    // 4 096 elements

struct AdamBatch {
  dg_adam_tensor_t t[kAdamBatch];
  int chunk_start[kAdamBatch + 1];   // first chunk (CTA) of each tensor inside this launch
  int n_tensors;
};

__device__ __forceinline__ int adam_find_tensor(const AdamBatch& b, int chunk) {
  int lo = 0, hi = b.n_tensors - 1;          // last tensor whose first chunk is <= chunk
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (b.chunk_start[mid] <= chunk) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// fixed-order CTA sum: lanes by shuffle tree, warps in warp order; every thread returns the total
__device__ __forceinline__ double adam_block_sum(double v, double* red /* [kAdamThreads / 32 + 1] */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();                                     // `red` may still be read from a previous call
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kAdamThreads / 32; ++w) s += red[w];
    red[kAdamThreads / 32] = s;
  }
  __syncthreads();
  return red[kAdamThreads / 32];
}

__global__ void __launch_bounds__(kAdamThreads)
adam_gradnorm_kernel(const __grid_constant__ AdamBatch b, double* __restrict__ partials, float* step, int bump_step) {
  __shared__ double red[kAdamThreads / 32 + 1];
  const int ti = adam_find_tensor(b, blockIdx.x);
  const dg_adam_tensor_t& T = b.t[ti];
  const int64_t off = static_cast<int64_t>(blockIdx.x - b.chunk_start[ti]) * kAdamChunk;
  const float* g = static_cast<const float*>(T.grad) + off;
  const int64_t left = T.numel - off;
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < kAdamPerThread; ++k) {
    const int i = threadIdx.x + k * kAdamThreads;
    if (i < left) { const float x = g[i]; acc = fmaf(x, x, acc); }
  }
  const double s = adam_block_sum(static_cast<double>(acc), red);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = s;
    if (bump_step && blockIdx.x == 0) *step += 1.0f;   // nothing in this kernel reads it; the update kernel does
  }
}

__global__ void __launch_bounds__(kAdamThreads)
adam_update_kernel(const __grid_constant__ AdamBatch b, const double* __restrict__ partials, int n_partials,
                   const float* __restrict__ step, const float* __restrict__ lr_dev, double lr, double beta1, double beta2,
                   float eps, float weight_decay, float max_norm, float* __restrict__ norm_out) {
  __shared__ double red[kAdamThreads / 32 + 1];
  __shared__ float cfg[3];                   // clip coefficient, step size, sqrt(bias correction 2)
  double part = 0.0;
  for (int i = threadIdx.x; i < n_partials; i += kAdamThreads) part += partials[i];
  const double total = adam_block_sum(part, red);        // same order in every CTA: one coefficient for all
  if (threadIdx.x == 0) {
    const double norm = sqrt(total);
    float coef = 1.f;
    if (max_norm > 0.f) {
      const double c = static_cast<double>(max_norm) / (norm + 1e-6);   // nn.utils.clip_grad_norm_
      coef = c < 1.0 ? static_cast<float>(c) : 1.f;
    }
    const double t = static_cast<double>(*step);
    const double bc1 = 1.0 - pow(beta1, t), bc2 = 1.0 - pow(beta2, t);
    const double rate = lr_dev ? static_cast<double>(*lr_dev) : lr;
    cfg[0] = coef;
    cfg[1] = static_cast<float>(rate / bc1);
    cfg[2] = static_cast<float>(sqrt(bc2));
    if (norm_out && blockIdx.x == 0) *norm_out = static_cast<float>(norm);
  }
  __syncthreads();
  const float coef = cfg[0], step_size = cfg[1], sqrt_bc2 = cfg[2];
  const int ti = adam_find_tensor(b, blockIdx.x);
  const dg_adam_tensor_t& T = b.t[ti];
  const int64_t off = static_cast<int64_t>(blockIdx.x - b.chunk_start[ti]) * kAdamChunk;
  float* p = static_cast<float*>(T.param) + off;
  float* g = static_cast<float*>(T.grad) + off;
  float* m = static_cast<float*>(T.exp_avg) + off;
  float* v = static_cast<float*>(T.exp_avg_sq) + off;
  const int64_t left = T.numel - off;
  // torch evaluates 1 - beta in Python floats (double) and rounds once: 1 - 0.999 -> 0.001f, not 1.f - 0.999f
  const float omb1 = static_cast<float>(1.0 - beta1), omb2 = static_cast<float>(1.0 - beta2), b2 = static_cast<float>(beta2);
#pragma unroll 4
  for (int k = 0; k < kAdamPerThread; ++k) {
    const int i = threadIdx.x + k * kAdamThreads;
    if (i < left) {
      const float pv = p[i];
      const float gc = g[i] * coef;                     // the clipped gradient stays in .grad, as after clip_grad_norm_
      const float gw = weight_decay != 0.f ? fmaf(weight_decay, pv, gc) : gc;
      const float mv = fmaf(gw - m[i], omb1, m[i]);     // lerp(exp_avg, grad, 1 - beta1)
      const float vv = fmaf(omb2 * gw, gw, v[i] * b2);
      const float denom = sqrtf(vv) / sqrt_bc2 + eps;
      g[i] = gc;
      m[i] = mv;
      v[i] = vv;
      p[i] = pv - step_size * (mv / denom);
    }
  }
}

}  // namespace
}  // namespace dg

extern "C" {

size_t dg_adam_workspace_bytes(const dg_adam_tensor_t* tensors, int n_tensors) {
  size_t chunks = 0;
  for (int i = 0; i < n_tensors; ++i)
    if (tensors && tensors[i].numel > 0) chunks += static_cast<size_t>((tensors[i].numel + dg::kAdamChunk - 1) / dg::kAdamChunk);
  return dg::ws_add(0, (chunks ? chunks : 1) * sizeof(double));
}

int dg_adam_clip_step_f32(const dg_adam_tensor_t* tensors, int n_tensors, float* step, const float* lr_dev, double lr,
                          double beta1, double beta2, double eps, double weight_decay, double max_norm, float* norm_out,
                          void* workspace, size_t workspace_bytes, dg_stream_t stream) {
  using namespace dg;
  DG_REQUIRE(n_tensors >= 0 && (n_tensors == 0 || tensors != nullptr), "bad tensor list");
  DG_REQUIRE(step != nullptr, "null step counter");
  DG_REQUIRE(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0, "bad hyper-parameters");
  int64_t total_chunks = 0;
  for (int i = 0; i < n_tensors; ++i) {
    const dg_adam_tensor_t& T = tensors[i];
    DG_REQUIRE(T.numel >= 0, "negative tensor length");
    if (T.numel == 0) continue;
    DG_REQUIRE(T.param && T.grad && T.exp_avg && T.exp_avg_sq, "null tensor pointer");
    total_chunks += (T.numel + kAdamChunk - 1) / kAdamChunk;
  }
  DG_REQUIRE(total_chunks < (1 << 30), "too many chunks");
  Workspace ws(workspace, workspace_bytes);
  double* partials = ws.take<double>(static_cast<size_t>(total_chunks ? total_chunks : 1));
  if (!partials) {
    set_error("dg_adam_clip_step_f32: workspace too small");
    return DG_ERR_WORKSPACE_TOO_SMALL;
  }
  cudaStream_t st = as_stream(stream);
  // pass 0: sum of squares per chunk (the first launch also advances the step counter); pass 1: clip + update.
  // One launch per batch of kAdamBatch tensors in each pass.
  for (int pass = 0; pass < 2; ++pass) {
    int i = 0;
    int64_t chunk_base = 0;
    bool first = true;
    while (i < n_tensors) {
      AdamBatch b;
      b.n_tensors = 0;
      int chunks = 0;
      for (; i < n_tensors && b.n_tensors < kAdamBatch; ++i) {
        if (tensors[i].numel == 0) continue;
        b.t[b.n_tensors] = tensors[i];
        b.chunk_start[b.n_tensors] = chunks;
        chunks += static_cast<int>((tensors[i].numel + kAdamChunk - 1) / kAdamChunk);
        ++b.n_tensors;
      }
      b.chunk_start[b.n_tensors] = chunks;
      if (chunks == 0) break;                          // only empty tensors were left
      if (pass == 0) {
        adam_gradnorm_kernel<<<chunks, kAdamThreads, 0, st>>>(b, partials + chunk_base, step, first ? 1 : 0);
        DG_CHECK_LAUNCH("adam_gradnorm");
      } else {
        adam_update_kernel<<<chunks, kAdamThreads, 0, st>>>(b, partials, static_cast<int>(total_chunks), step, lr_dev, lr, beta1,
                                                            beta2, static_cast<float>(eps), static_cast<float>(weight_decay),
                                                            static_cast<float>(max_norm), first ? norm_out : nullptr);
        DG_CHECK_LAUNCH("adam_update");
      }
      first = false;
      chunk_base += chunks;
    }
  }
  return DG_OK;
}

}  // extern "C"
