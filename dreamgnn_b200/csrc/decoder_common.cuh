// Pieces shared by the SIMT decoder kernels (decoder.cu) and the tensor-core ones (decoder_tc.cu).
#pragma once
#include <math.h>

#include "common.cuh"

namespace dg {

constexpr int H1 = DG_DEC_H1;   // 128
constexpr int H2 = DG_DEC_H2;   // 64
constexpr int kDecThreads = 256;

struct DropCfg {
  uint32_t thresh;   // round(p * 65536) on 16-bit uniforms; 0 disables
  float scale;       // 1 / (actual keep probability)
  uint64_t seed;
  const uint64_t* seed_dev;   // when non-null the seed is read from device memory (CUDA-graph replays)
};

inline DropCfg make_drop(float p, uint64_t seed, const uint64_t* seed_dev) {
  DropCfg c;
  if (p <= 0.f) { c.thresh = 0; c.scale = 1.f; }
  else {
    long t = lround(static_cast<double>(p) * 65536.0);
    c.thresh = static_cast<uint32_t>(t < 1 ? 1 : (t > 65535 ? 65535 : t));
    c.scale = static_cast<float>(65536.0 / (65536.0 - c.thresh));     // unbiased for the quantised keep rate
  }
  c.seed = seed;
  c.seed_dev = seed_dev;
  return c;
}

// per-CTA partial block of the backward: dW2 [H2][H1] | db2 [H2] | dw3 [H2] | db3 (first slot of the last H2 block)
constexpr int kPartial = H2 * H1 + 3 * H2;

// tensor-core (tcgen05) variants, decoder_tc.cu; same contracts as the SIMT kernels
int launch_decoder_fwd_tc(const int* src, const int* dst, const int* perm, int64_t n_pairs, const float* pd, const float* ps, const float* w2,
                          const float* b2, const float* w3, const float* b3, DropCfg drop, float* out, float* z2_save,
                          cudaStream_t st);
int launch_decoder_bwd_tc(const int* src, const int* dst, const int* perm, int64_t n_pairs, const float* pd, const float* ps, const float* w2,
                          const float* w3, DropCfg drop, const float* z2, const float* dout, float* dz1, float* partials,
                          int* n_ctas, const int* pair_slot, float* slot_rows, cudaStream_t st);

}  // namespace dg
