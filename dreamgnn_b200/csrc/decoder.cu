// Fused MLP decoder over scored (drug, disease) pairs -- replaces DGL apply_edges(udf_u_mul_e) +
// lin1/lin2/lin3 with ReLU and dropout (layers.py:360-379).
//
// lin1 is split exactly (lin1(cat(a,b)) = a W1a^T + b W1b^T + b1), so the [E,256] concat (479 MB at
// lrssl, 20 GB at the 20M-pair shape) is never built: the kernel gathers the two 128-wide node
// projections per pair (SDDMM-style), adds them, and runs the 128->64->1 tail on the tile in shared
// memory. Backward regenerates z1 (same counter-based dropout mask), reads the saved z2, and writes
// dz1 per pair; the scatter into node gradients is done by the caller as two deterministic segment
// sums over the decoder graph's CSR/CSC (spmm.cu) -- no atomics anywhere.
//
// fp32 SIMT FMA (the 1e-5 parity path): E x 8192 MAC forward, E x 16384 MAC backward.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "decoder_common.cuh"

namespace dg {

// Gather one tile of z1 = drop(relu(pd[src] + ps[dst])) into shared memory, layout Z[p][H1].
// One warp per pair at a time, a lane owns 4 consecutive hidden units (one float4 per operand row).
template <int TP>
__device__ __forceinline__ void gather_z1_tile(const int* __restrict__ src, const int* __restrict__ dst,
                                               int64_t tile_base, int64_t n_pairs, const float* __restrict__ pd,
                                               const float* __restrict__ ps, const DropCfg drop, float* Z) {
  const int lane = lane_id(), w = threadIdx.x >> 5;
  constexpr int kWarps = kDecThreads / 32;
  constexpr int kPer = TP / kWarps;          // pairs per warp
  constexpr int U = 4;                       // pairs in flight per warp
  static_assert(kPer % U == 0, "tile/warp mismatch");
#pragma unroll 1
  for (int i0 = 0; i0 < kPer; i0 += U) {
    float4 a[U], b[U];
    int64_t e[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = w * kPer + i0 + u;
      e[u] = tile_base + p;
      if (e[u] < n_pairs) {
        const int s = src[e[u]], d = dst[e[u]];
        a[u] = __ldg(reinterpret_cast<const float4*>(pd + static_cast<int64_t>(s) * H1) + lane);
        b[u] = __ldg(reinterpret_cast<const float4*>(ps + static_cast<int64_t>(d) * H1) + lane);
      } else {
        a[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        b[u] = a[u];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = w * kPer + i0 + u;
      float z[4] = {a[u].x + b[u].x, a[u].y + b[u].y, a[u].z + b[u].z, a[u].w + b[u].w};
      DropBits bits;
      if (drop.thresh) bits = dropout_bits(drop.seed, static_cast<uint32_t>(e[u]), lane);      // units 4*lane .. +3
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float v = fmaxf(z[q], 0.f);
        if (drop.thresh) v = dropout_keep16(bits, q, drop.thresh) ? v * drop.scale : 0.f;
        z[q] = v;
      }
      *reinterpret_cast<float4*>(Z + p * H1 + lane * 4) = make_float4(z[0], z[1], z[2], z[3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// forward: tile of 128 pairs; thread (tp = t/16, tj = t%16) owns pairs tp*8..+8 and units tj*4..+4
// ------------------------------------------------------------------------------------------------
constexpr int kFwdTile = 128;
constexpr size_t kFwdSmem = (static_cast<size_t>(H1) * H2 + static_cast<size_t>(kFwdTile) * H1) * sizeof(float);

__global__ void __launch_bounds__(kDecThreads, 2)
decoder_fwd_kernel(const int* __restrict__ src, const int* __restrict__ dst, const int* __restrict__ perm, int64_t n_pairs,
                   const float* __restrict__ pd, const float* __restrict__ ps, const float* __restrict__ w2,
                   const float* __restrict__ b2, const float* __restrict__ w3, const float* __restrict__ b3,
                   DropCfg drop, float* __restrict__ out, float* __restrict__ z2_save) {
  if (drop.seed_dev) drop.seed = *drop.seed_dev;
  extern __shared__ __align__(16) float smem[];
  float* W2T = smem;               // [H1][H2]  (k-major so a thread reads its 4 units as one float4)
  float* Z = smem + H1 * H2;       // [kFwdTile][H1]
  const int t = threadIdx.x;
  for (int i = t; i < H1 * H2; i += kDecThreads) {
    const int j = i / H1, k = i - j * H1;       // w2 is [H2][H1] row-major (nn.Linear weight)
    W2T[k * H2 + j] = w2[i];
  }
  const int tj = t & 15, tp = t >> 4;
  const int j0 = tj * 4, p0 = tp * 8;
  const float4 bias2 = *reinterpret_cast<const float4*>(b2 + j0);
  const float4 wv3 = *reinterpret_cast<const float4*>(w3 + j0);
  const float bias3 = b3[0];
  const int64_t n_tiles = (n_pairs + kFwdTile - 1) / kFwdTile;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t base = tile * kFwdTile;
    __syncthreads();   // W2T ready (first trip) / previous tile's Z fully consumed
    gather_z1_tile<kFwdTile>(src, dst, base, n_pairs, pd, ps, drop, Z);
    __syncthreads();
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 2
    for (int k = 0; k < H1; k += 4) {
      float4 wk[4];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) wk[kk] = *reinterpret_cast<const float4*>(W2T + (k + kk) * H2 + j0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 z = *reinterpret_cast<const float4*>(Z + (p0 + i) * H1 + k);
        acc[i][0] = fmaf(z.x, wk[0].x, acc[i][0]); acc[i][1] = fmaf(z.x, wk[0].y, acc[i][1]);
        acc[i][2] = fmaf(z.x, wk[0].z, acc[i][2]); acc[i][3] = fmaf(z.x, wk[0].w, acc[i][3]);
        acc[i][0] = fmaf(z.y, wk[1].x, acc[i][0]); acc[i][1] = fmaf(z.y, wk[1].y, acc[i][1]);
        acc[i][2] = fmaf(z.y, wk[1].z, acc[i][2]); acc[i][3] = fmaf(z.y, wk[1].w, acc[i][3]);
        acc[i][0] = fmaf(z.z, wk[2].x, acc[i][0]); acc[i][1] = fmaf(z.z, wk[2].y, acc[i][1]);
        acc[i][2] = fmaf(z.z, wk[2].z, acc[i][2]); acc[i][3] = fmaf(z.z, wk[2].w, acc[i][3]);
        acc[i][0] = fmaf(z.w, wk[3].x, acc[i][0]); acc[i][1] = fmaf(z.w, wk[3].y, acc[i][1]);
        acc[i][2] = fmaf(z.w, wk[3].z, acc[i][2]); acc[i][3] = fmaf(z.w, wk[3].w, acc[i][3]);
      }
    }
    // epilogue: z2 = drop(relu(acc + b2)); out = w3 . z2 + b3, reduced over the 16 tj lanes
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t e = base + p0 + i;
      float z2[4] = {acc[i][0] + bias2.x, acc[i][1] + bias2.y, acc[i][2] + bias2.z, acc[i][3] + bias2.w};
      DropBits bits;
      if (drop.thresh) bits = dropout_bits(drop.seed, static_cast<uint32_t>(e), H1 / 4 + tj);   // units j0 .. j0+3
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float v = fmaxf(z2[q], 0.f);
        if (drop.thresh) v = dropout_keep16(bits, q, drop.thresh) ? v * drop.scale : 0.f;
        z2[q] = v;
      }
      if (z2_save && e < n_pairs) *reinterpret_cast<float4*>(z2_save + e * H2 + j0) = make_float4(z2[0], z2[1], z2[2], z2[3]);
      float part = z2[0] * wv3.x + z2[1] * wv3.y + z2[2] * wv3.z + z2[3] * wv3.w;
#pragma unroll
      for (int o = 8; o; o >>= 1) part += __shfl_xor_sync(kFull, part, o);   // fixed tree: deterministic
      if (tj == 0 && e < n_pairs) out[perm ? perm[e] : e] = part + bias3;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward: tile of 64 pairs.
//   dz2[p][j] = dout[p] * w3[j] * scale * [z2[p][j] > 0]
//   dz1[p][k] = (sum_j dz2[p][j] W2[j][k]) * scale * [z1[p][k] > 0]          -> global, per pair
//   dW2[j][k] += sum_p dz2[p][j] z1[p][k]   (register accumulators across the CTA's tiles)
//   db2[j] += sum_p dz2[p][j];  dw3[j] += sum_p dout[p] z2[p][j];  db3 += sum_p dout[p]
// Per-CTA partials go to the workspace and are summed in CTA order by decoder_reduce_partials.
// ------------------------------------------------------------------------------------------------
constexpr int kBwdTile = 64;
constexpr size_t kBwdSmem = (static_cast<size_t>(H2) * H1 + static_cast<size_t>(kBwdTile) * H1 +
                             static_cast<size_t>(kBwdTile) * H2 + kBwdTile) * sizeof(float);

__global__ void __launch_bounds__(kDecThreads, 2)
decoder_bwd_kernel(const int* __restrict__ src, const int* __restrict__ dst, const int* __restrict__ perm, int64_t n_pairs,
                   const float* __restrict__ pd, const float* __restrict__ ps, const float* __restrict__ w2,
                   const float* __restrict__ w3, DropCfg drop, const float* __restrict__ z2,
                   const float* __restrict__ dout, float* __restrict__ dz1, float* __restrict__ partials) {
  if (drop.seed_dev) drop.seed = *drop.seed_dev;
  extern __shared__ __align__(16) float smem[];
  float* W2S = smem;                          // [H2][H1] as stored
  float* Z = W2S + H2 * H1;                   // [kBwdTile][H1]
  float* DZ2 = Z + kBwdTile * H1;             // [kBwdTile][H2]
  float* DO = DZ2 + kBwdTile * H2;            // [kBwdTile]
  const int t = threadIdx.x;
  for (int i = t; i < H2 * H1 / 4; i += kDecThreads)
    reinterpret_cast<float4*>(W2S)[i] = __ldg(reinterpret_cast<const float4*>(w2) + i);

  // persistent accumulators
  const int gj = (t >> 4) * 4;                // dW2 rows gj..gj+3
  const int gk = (t & 15) * 4;                // dW2 cols gk..gk+3 and 64+gk..
  float dW[4][8];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) dW[a][b] = 0.f;
  // dz2 phase: thread (pz = t/16 -> pairs pz*4..+4, jz = t%16 -> units jz*4..+4)
  const int pz = (t >> 4) * 4, jz = (t & 15) * 4;
  const float4 wv3 = *reinterpret_cast<const float4*>(w3 + jz);
  float db2_p[4] = {0.f, 0.f, 0.f, 0.f}, dw3_p[4] = {0.f, 0.f, 0.f, 0.f}, db3_p = 0.f;

  const int64_t n_tiles = (n_pairs + kBwdTile - 1) / kBwdTile;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t base = tile * kBwdTile;
    __syncthreads();
    gather_z1_tile<kBwdTile>(src, dst, base, n_pairs, pd, ps, drop, Z);
    if (t < kBwdTile) DO[t] = (base + t < n_pairs) ? dout[perm ? perm[base + t] : base + t] : 0.f;
    __syncthreads();
    // ---- dz2 tile + db2 / dw3 / db3 partials -------------------------------------------------
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p = pz + i;
      const int64_t e = base + p;
      const float g = DO[p];
      float4 zz = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e < n_pairs) zz = __ldg(reinterpret_cast<const float4*>(z2 + e * H2 + jz));
      const float gs = g * drop.scale;
      float4 d;
      d.x = zz.x > 0.f ? gs * wv3.x : 0.f; d.y = zz.y > 0.f ? gs * wv3.y : 0.f;
      d.z = zz.z > 0.f ? gs * wv3.z : 0.f; d.w = zz.w > 0.f ? gs * wv3.w : 0.f;
      *reinterpret_cast<float4*>(DZ2 + p * H2 + jz) = d;
      db2_p[0] += d.x; db2_p[1] += d.y; db2_p[2] += d.z; db2_p[3] += d.w;
      dw3_p[0] = fmaf(g, zz.x, dw3_p[0]); dw3_p[1] = fmaf(g, zz.y, dw3_p[1]);
      dw3_p[2] = fmaf(g, zz.z, dw3_p[2]); dw3_p[3] = fmaf(g, zz.w, dw3_p[3]);
      if (jz == 0) db3_p += g;
    }
    __syncthreads();
    // ---- dz1 = (DZ2 @ W2) masked: thread owns pairs pz..pz+3 and cols gk..+3, 64+gk..+3 -------
    {
      float acc[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[i][b] = 0.f;
#pragma unroll 2
      for (int j = 0; j < H2; j += 4) {
        float4 wa[4], wb[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          wa[jj] = *reinterpret_cast<const float4*>(W2S + (j + jj) * H1 + gk);
          wb[jj] = *reinterpret_cast<const float4*>(W2S + (j + jj) * H1 + 64 + gk);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 dz = *reinterpret_cast<const float4*>(DZ2 + (pz + i) * H2 + j);
          const float dv[4] = {dz.x, dz.y, dz.z, dz.w};
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            acc[i][0] = fmaf(dv[jj], wa[jj].x, acc[i][0]); acc[i][1] = fmaf(dv[jj], wa[jj].y, acc[i][1]);
            acc[i][2] = fmaf(dv[jj], wa[jj].z, acc[i][2]); acc[i][3] = fmaf(dv[jj], wa[jj].w, acc[i][3]);
            acc[i][4] = fmaf(dv[jj], wb[jj].x, acc[i][4]); acc[i][5] = fmaf(dv[jj], wb[jj].y, acc[i][5]);
            acc[i][6] = fmaf(dv[jj], wb[jj].z, acc[i][6]); acc[i][7] = fmaf(dv[jj], wb[jj].w, acc[i][7]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int p = pz + i;
        const int64_t e = base + p;
        if (e >= n_pairs) continue;
        const float4 za = *reinterpret_cast<const float4*>(Z + p * H1 + gk);
        const float4 zb = *reinterpret_cast<const float4*>(Z + p * H1 + 64 + gk);
        const float s = drop.scale;
        float4 ra = make_float4(za.x > 0.f ? acc[i][0] * s : 0.f, za.y > 0.f ? acc[i][1] * s : 0.f,
                                za.z > 0.f ? acc[i][2] * s : 0.f, za.w > 0.f ? acc[i][3] * s : 0.f);
        float4 rb = make_float4(zb.x > 0.f ? acc[i][4] * s : 0.f, zb.y > 0.f ? acc[i][5] * s : 0.f,
                                zb.z > 0.f ? acc[i][6] * s : 0.f, zb.w > 0.f ? acc[i][7] * s : 0.f);
        *reinterpret_cast<float4*>(dz1 + e * H1 + gk) = ra;
        *reinterpret_cast<float4*>(dz1 + e * H1 + 64 + gk) = rb;
      }
    }
    // ---- dW2 += DZ2^T Z: thread owns rows gj..gj+3, cols gk..+3 and 64+gk..+3 -------------------
#pragma unroll 4
    for (int p = 0; p < kBwdTile; ++p) {
      const float4 dz = *reinterpret_cast<const float4*>(DZ2 + p * H2 + gj);
      const float4 za = *reinterpret_cast<const float4*>(Z + p * H1 + gk);
      const float4 zb = *reinterpret_cast<const float4*>(Z + p * H1 + 64 + gk);
      const float dv[4] = {dz.x, dz.y, dz.z, dz.w};
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        dW[a][0] = fmaf(dv[a], za.x, dW[a][0]); dW[a][1] = fmaf(dv[a], za.y, dW[a][1]);
        dW[a][2] = fmaf(dv[a], za.z, dW[a][2]); dW[a][3] = fmaf(dv[a], za.w, dW[a][3]);
        dW[a][4] = fmaf(dv[a], zb.x, dW[a][4]); dW[a][5] = fmaf(dv[a], zb.y, dW[a][5]);
        dW[a][6] = fmaf(dv[a], zb.z, dW[a][6]); dW[a][7] = fmaf(dv[a], zb.w, dW[a][7]);
      }
    }
  }
  // ---- per-CTA partials ------------------------------------------------------------------------
  float* my = partials + static_cast<size_t>(blockIdx.x) * kPartial;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    *reinterpret_cast<float4*>(my + (gj + a) * H1 + gk) = make_float4(dW[a][0], dW[a][1], dW[a][2], dW[a][3]);
    *reinterpret_cast<float4*>(my + (gj + a) * H1 + 64 + gk) = make_float4(dW[a][4], dW[a][5], dW[a][6], dW[a][7]);
  }
  // db2 / dw3 / db3: sum the 16 pair-groups (t/16) for each unit group (t%16) in a fixed order
  __syncthreads();
  float* red = Z;   // reuse: [16 groups][16 jz][9]
  float* mine = red + ((t >> 4) * 16 + (t & 15)) * 9;
#pragma unroll
  for (int q = 0; q < 4; ++q) { mine[q] = db2_p[q]; mine[4 + q] = dw3_p[q]; }
  mine[8] = db3_p;
  __syncthreads();
  if (t < 16) {
    float s[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int g = 0; g < 16; ++g)
#pragma unroll
      for (int q = 0; q < 9; ++q) s[q] += red[(g * 16 + t) * 9 + q];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      my[H2 * H1 + t * 4 + q] = s[q];                 // db2
      my[H2 * H1 + H2 + t * 4 + q] = s[4 + q];        // dw3
    }
    if (t == 0) my[H2 * H1 + 2 * H2] = s[8];          // db3
  }
}

__global__ void decoder_reduce_partials(const float* __restrict__ partials, int n_ctas, float* __restrict__ dw2,
                                        float* __restrict__ db2, float* __restrict__ dw3, float* __restrict__ db3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H2 * H1 + 2 * H2 + 1) return;
  float s = 0.f;
  for (int c = 0; c < n_ctas; ++c) s += partials[static_cast<size_t>(c) * kPartial + i];   // CTA order: deterministic
  if (i < H2 * H1) dw2[i] = s;
  else if (i < H2 * H1 + H2) db2[i - H2 * H1] = s;
  else if (i < H2 * H1 + 2 * H2) dw3[i - H2 * H1 - H2] = s;
  else db3[0] = s;
}

// DG_DECODER=simt selects the fp32 FMA kernels below (A/B runs); the default is the tcgen05 path (decoder_tc.cu)
static bool use_simt_decoder() {
  const char* v = getenv("DG_DECODER");
  return v && strcmp(v, "simt") == 0;
}

static int decoder_grid(int64_t n_tiles) {
  int64_t g = static_cast<int64_t>(kNumSM) * 2;     // 2 resident CTAs per SM
  if (n_tiles < g) g = n_tiles;
  return static_cast<int>(g < 1 ? 1 : g);
}

}  // namespace dg

extern "C" {

int dg_decoder_fwd_f32(const int32_t* src, const int32_t* dst, const int32_t* perm, int64_t n_pairs, const float* pd, const float* ps,
                       const float* w2, const float* b2, const float* w3, const float* b3, float dropout_p,
                       uint64_t seed, const uint64_t* seed_dev, float* out, float* z2_save, dg_stream_t stream) {
  using namespace dg;
  DG_REQUIRE(n_pairs >= 0, "n_pairs < 0");
  DG_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "dropout_p must be in [0,1)");
  if (n_pairs == 0) return DG_OK;
  if (!use_simt_decoder())
    return launch_decoder_fwd_tc(src, dst, perm, n_pairs, pd, ps, w2, b2, w3, b3, make_drop(dropout_p, seed, seed_dev), out, z2_save,
                                 as_stream(stream));
  static bool attr_set = false;
  if (!attr_set) {
    DG_CHECK_CUDA(cudaFuncSetAttribute(decoder_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kFwdSmem)));
    attr_set = true;
  }
  const int64_t n_tiles = (n_pairs + kFwdTile - 1) / kFwdTile;
  decoder_fwd_kernel<<<decoder_grid(n_tiles), kDecThreads, kFwdSmem, as_stream(stream)>>>(
      src, dst, perm, n_pairs, pd, ps, w2, b2, w3, b3, make_drop(dropout_p, seed, seed_dev), out, z2_save);
  DG_CHECK_LAUNCH("decoder_fwd");
  return DG_OK;
}

size_t dg_decoder_bwd_workspace_bytes(int64_t n_pairs) {
  (void)n_pairs;
  return dg::ws_add(0, static_cast<size_t>(dg::kNumSM) * 2 * dg::kPartial * sizeof(float));
}

int dg_decoder_bwd_f32(const int32_t* src, const int32_t* dst, const int32_t* perm, int64_t n_pairs, const float* pd, const float* ps,
                       const float* w2, const float* w3, float dropout_p, uint64_t seed, const uint64_t* seed_dev, const float* z2,
                       const float* dout, float* dz1, float* dw2, float* db2, float* dw3, float* db3,
                       const int32_t* pair_slot, float* slot_rows, void* workspace, size_t workspace_bytes, dg_stream_t stream) {
  using namespace dg;
  DG_REQUIRE(n_pairs >= 0, "n_pairs < 0");
  DG_REQUIRE((pair_slot == nullptr) == (slot_rows == nullptr), "pair_slot and slot_rows go together");
  DG_REQUIRE(pair_slot == nullptr || !use_simt_decoder(), "the fused source-node segment sum needs the tensor-core kernel");
  DG_REQUIRE((reinterpret_cast<uintptr_t>(pair_slot) & 15) == 0, "pair_slot must be 16-byte aligned");
  DG_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "dropout_p must be in [0,1)");
  Workspace w(workspace, workspace_bytes);
  float* partials = w.take<float>(static_cast<size_t>(kNumSM) * 2 * kPartial);
  if (!partials) { set_error("decoder_bwd: workspace too small"); return DG_ERR_WORKSPACE_TOO_SMALL; }
  int grid = 0;
  if (!use_simt_decoder()) {
    DG_PROPAGATE(launch_decoder_bwd_tc(src, dst, perm, n_pairs, pd, ps, w2, w3, make_drop(dropout_p, seed, seed_dev), z2, dout, dz1,
                                       partials, &grid, pair_slot, slot_rows, as_stream(stream)));
  } else {
    static bool attr_set = false;
    if (!attr_set) {
      DG_CHECK_CUDA(cudaFuncSetAttribute(decoder_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kBwdSmem)));
      attr_set = true;
    }
    const int64_t n_tiles = (n_pairs + kBwdTile - 1) / kBwdTile;
    grid = decoder_grid(n_tiles);
    decoder_bwd_kernel<<<grid, kDecThreads, kBwdSmem, as_stream(stream)>>>(
        src, dst, perm, n_pairs, pd, ps, w2, w3, make_drop(dropout_p, seed, seed_dev), z2, dout, dz1, partials);
    DG_CHECK_LAUNCH("decoder_bwd");
  }
  constexpr int kOut = H2 * H1 + 2 * H2 + 1;
  decoder_reduce_partials<<<(kOut + 255) / 256, 256, 0, as_stream(stream)>>>(partials, grid, dw2, db2, dw3, db3);
  DG_CHECK_LAUNCH("decoder_reduce_partials");
  return DG_OK;
}
}
