// MLP decoder over scored pairs on the 5th-generation tensor cores (layers.py:341-379).
//
// Same contract as the SIMT kernels in decoder.cu (same dropout counters, same saved z2, same per-CTA
// partial layout), but the per-tile GEMMs run as tcgen05.mma.kind::tf32 with the 3xTF32 error compensation
// of gemm_tc.cu (D = A_lo B_hi + A_hi B_lo + A_hi B_hi; here hi = rna_tf32(x), lo = rna_tf32(x - hi) since the
// gathering threads write both halves anyway), so the fp32 1e-5 parity bar still holds while the FMA pipe is taken
// off the critical path:
//
//   forward   z2 tile [128 pairs x 64]  = z1 [128 x 128] . W2^T            (A, B K-major)
//   backward  dz1^T tile [128 x 64 pairs] = W2^T [128 x 64] . dz2^T          (A = W2^T, B = dz2, both K-major; the
//                                                                            transposed product fills all 128 rows)
//             dW2^T [128 x 64]         += z1^T . dz2  over the tile's pairs (A and B MN-major: pairs are K)
//
// The operand tiles are written by the gathering threads straight into the swizzled layouts the UMMA
// shared-memory descriptors expect (no TMA: the rows are an indexed gather): K-major operands in the usual
// 128-byte swizzle, MN-major ones in the 128-byte swizzle with a 32-byte base, the only MN-major layout fp32 /
// tf32 operands have -- dz2 is the one tile read both ways and is written twice. The accumulators live in
// TMEM, and dW2 stays in TMEM across tiles (flushed to the CTA's fp32 partial every 16 tiles = K 1024, because
// the tensor-core accumulator truncates when it aligns addends).
#include <stdlib.h>

#include "common.cuh"
#include "decoder_common.cuh"
#include "tc_common.cuh"

namespace dg {

namespace {

constexpr int kTcThreads = 512;           // 16 warps, one CTA per SM

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(static_cast<uint32_t>(kCols)));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
template <int kCols>
__device__ __forceinline__ void tmem_free(uint32_t base) {     // one full warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(static_cast<uint32_t>(kCols)));
}

// error-compensated product of one K = 8 slice: small terms first, the dominant one last
__device__ __forceinline__ void mma3(uint32_t d, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo, uint32_t idesc,
                                     uint32_t accumulate) {
  umma_tf32(d, a_lo, b_hi, idesc, accumulate);
  umma_tf32(d, a_hi, b_lo, idesc, 1u);
  umma_tf32(d, a_hi, b_hi, idesc, 1u);
}

// W2 [H2][H1] (nn.Linear weight, row-major) -> shared image: 4 blocks of 32 hidden-1 units, each [64 rows j][128 B]
// swizzled; hi = rna_tf32(x), lo = rna_tf32(x - hi). Read K-major (K = unit) by the forward GEMM and MN-major
// (N = unit, K = j; 32-byte-base swizzle) by the backward dz1 GEMM.
constexpr int kWBlk = H2 * 128;                      // 8 KB
template <bool kMnMajor>
__device__ __forceinline__ void stage_w2(const float* __restrict__ w2, uint8_t* W_hi, uint8_t* W_lo) {
  for (int i = threadIdx.x; i < H2 * H1 / 4; i += kTcThreads) {
    const int j = i >> 5, c4 = i & 31;
    const float4 x = __ldg(reinterpret_cast<const float4*>(w2) + i);
    const uint32_t off = static_cast<uint32_t>(c4 >> 3) * kWBlk + (kMnMajor ? mn32_off(j, c4 & 7) : sw128_off(j, c4 & 7));
    float4 hi, lo;
    split_tf32(x, hi, lo);
    *reinterpret_cast<float4*>(W_hi + off) = hi;
    *reinterpret_cast<float4*>(W_lo + off) = lo;
  }
}

// Gather of z1 = drop(relu(pd[src] + ps[dst])) for one tile, software-pipelined through registers: a warp owns
// U consecutive pairs of the tile, a lane 4 consecutive hidden-1 units of each (one float4 per operand row).
// `issue_idx` loads a tile's endpoints, `issue_rows` -- one tile later -- starts its row loads (2 * U float4 in flight
// per lane, 512 threads: 64 KB per SM at U = 4), `commit` -- another tile later, after the previous tile's MMAs have
// retired -- finishes z1 and writes the
// hi / lo halves into the swizzled operand tiles: 4 blocks of 32 units, each [kTile pairs][128 B], K-major for
// the forward GEMM, MN-major (pairs are K) for the backward dW2 GEMM.
template <int U>
struct GatherRegs {
  float4 a[U], b[U];
  int s, d;                // lane u < U: endpoints of pair u of the tile after the one in a / b (the index loads run two
                           // tiles ahead, so the row loads never wait on them); -1 past the end
  uint32_t keep;           // dropout keep decisions of the tile in a / b: bit 4u + q = unit 4*lane + q of pair u survives
};

// The keep decisions depend on (seed, pair, unit) only, not on loaded data: they are hashed for the NEXT tile right before
// the wait for the current tile's MMAs -- ~36 integer instructions per pair and lane that would otherwise sit on the
// critical path in front of the operand commit run in the shadow of the tensor core (all 16 warps idle-spin there).
template <int U>
__device__ __forceinline__ uint32_t gather_keep_bits(int64_t base, const DropCfg& drop) {
  static_assert(U * 4 <= 32, "one mask word per tile and lane");
  uint32_t m = 0;
  if (drop.thresh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const DropBits bits = dropout_bits(drop.seed, static_cast<uint32_t>(base + warp * U + u), lane);      // units 4*lane .. +3
#pragma unroll
      for (int q = 0; q < 4; ++q) m |= static_cast<uint32_t>(dropout_keep16(bits, q, drop.thresh)) << (u * 4 + q);
    }
  }
  return m;
}

template <int U>
__device__ __forceinline__ void gather_issue_idx(const int* __restrict__ src, const int* __restrict__ dst, int64_t base,
                                                 int64_t n_pairs, GatherRegs<U>& r) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t e = base + warp * U + lane;
  r.s = -1;
  r.d = -1;
  if (lane < U && e < n_pairs) { r.s = __ldg(src + e); r.d = __ldg(dst + e); }
}

template <int U>
__device__ __forceinline__ void gather_issue_rows(const float* __restrict__ pd, const float* __restrict__ ps, GatherRegs<U>& r) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int su = __shfl_sync(kFull, r.s, u), du = __shfl_sync(kFull, r.d, u);
    if (su >= 0) {
      r.a[u] = __ldg(reinterpret_cast<const float4*>(pd + static_cast<int64_t>(su) * H1) + lane);
      r.b[u] = __ldg(reinterpret_cast<const float4*>(ps + static_cast<int64_t>(du) * H1) + lane);
    } else {
      r.a[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      r.b[u] = r.a[u];
    }
  }
}

template <int kTile, int U, bool kMnMajor>
__device__ __forceinline__ void gather_commit(const GatherRegs<U>& r, int64_t base, const DropCfg& drop, uint8_t* Z_hi,
                                              uint8_t* Z_lo) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr uint32_t kBlk = kTile * 128;
  const uint32_t blk_off = static_cast<uint32_t>(lane >> 3) * kBlk;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int p = warp * U + u;
    float z[4] = {r.a[u].x + r.b[u].x, r.a[u].y + r.b[u].y, r.a[u].z + r.b[u].z, r.a[u].w + r.b[u].w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float v = fmaxf(z[q], 0.f);
      if (drop.thresh) v = ((r.keep >> (u * 4 + q)) & 1u) ? v * drop.scale : 0.f;
      z[q] = v;
    }
    const uint32_t off = blk_off + (kMnMajor ? mn32_off(p, lane & 7) : sw128_off(p, lane & 7));
    float4 hi, lo;
    split_tf32(make_float4(z[0], z[1], z[2], z[3]), hi, lo);
    *reinterpret_cast<float4*>(Z_hi + off) = hi;
    *reinterpret_cast<float4*>(Z_lo + off) = lo;
  }
}

// ------------------------------------------------------------------------------------------------
// forward: kFT-pair tiles (64: accumulator rows 64..127 are padding; 128: full tile, 8 pairs per warp in flight)
// ------------------------------------------------------------------------------------------------
// The tensor-core accumulator truncates when it aligns addends, so a long accumulation chain picks up a bias of
// ~0.5 ulp per MMA (measured: mean |err| 3e-7 on z2 with all 48 MMAs of a tile in one accumulator, 6x the fp32
// FMA kernel -- enough to flip relu masks of pre-activations next to zero and with them whole gradient terms).
// The hi*hi products of each 32-unit block therefore get their own accumulator (4 MMAs each), the two small
// cross terms share a fifth, and the epilogue adds the five in fp32 (round to nearest).
constexpr int kFwdTmemCols = 512;                    // 5 accumulators x 64 columns (power-of-two allocation)
constexpr uint32_t kIdescFwd = tf32_idesc(128, H2, 0, 0);
template <int kFT>
constexpr size_t fwd_tc_smem() {
  return 1024 /*alignment slack*/ + 8 * (kFT * 128) + 8 * kWBlk + (4 * kFT + 2 * H2) * sizeof(float) + 64;
}

template <int kFT>
__global__ void __launch_bounds__(kTcThreads, 1)
decoder_fwd_tc_kernel(const int* __restrict__ src, const int* __restrict__ dst, const int* __restrict__ perm, int64_t n_pairs,
                      const float* __restrict__ pd, const float* __restrict__ ps, const float* __restrict__ w2,
                      const float* __restrict__ b2, const float* __restrict__ w3, const float* __restrict__ b3,
                      DropCfg drop, float* __restrict__ out, float* __restrict__ z2_save) {
  constexpr int kFBlk = kFT * 128;                   // [kFT pairs][32 units]
  constexpr int U = kFT / (kTcThreads / 32);         // pairs per warp
  if (drop.seed_dev) drop.seed = *drop.seed_dev;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the swizzle atoms, kept as an offset from the __shared__ array so that the compiler
  // still knows the address space (LDS / STS instead of generic LD / ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* A_hi = smem;
  uint8_t* A_lo = A_hi + 4 * kFBlk;                  // (kFT = 64: the M=128 descriptors over-read into what follows --
  uint8_t* W_hi = A_lo + 4 * kFBlk;                  //  finite values whose accumulator rows 64..127 are never read)
  uint8_t* W_lo = W_hi + 4 * kWBlk;
  float* part = reinterpret_cast<float*>(W_lo + 4 * kWBlk);          // [4 column groups][kFT pairs]
  float* b2s = part + 4 * kFT;                       // b2 | w3
  uint64_t* bar = reinterpret_cast<uint64_t*>(b2s + 2 * H2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;

  stage_w2<false>(w2, W_hi, W_lo);
  if (t < H2) { b2s[t] = __ldg(b2 + t); b2s[H2 + t] = __ldg(w3 + t); }
  const uint32_t bar_a = smem_u32(bar);
  if (t == 0) {
    mbar_init(bar_a, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc<kFwdTmemCols>(tmem_slot);
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // epilogue role: TMEM lanes 32*(warp % 4) .. +31 (pairs), columns 16*(warp / 4) .. +15 (hidden-2 units)
  const int q = warp & 3, g = warp >> 2;
  const int j0 = g * 16;
  const float bias3 = __ldg(b3);

  uint32_t phase = 0;
  const int64_t n_tiles = (n_pairs + kFT - 1) / kFT;
  GatherRegs<U> regs;
  int64_t tile = blockIdx.x;
  gather_issue_idx<U>(src, dst, tile * kFT, n_pairs, regs);               // (past-the-end tiles load nothing)
  gather_issue_rows<U>(pd, ps, regs);
  gather_issue_idx<U>(src, dst, (tile + gridDim.x) * kFT, n_pairs, regs);
  regs.keep = gather_keep_bits<U>(tile * kFT, drop);
  for (; tile < n_tiles; tile += gridDim.x) {
    const int64_t base = tile * kFT;
    gather_commit<kFT, U, false>(regs, base, drop, A_hi, A_lo);
    proxy_fence();
    __syncthreads();
    if (t == 0) {
      tc_fence_after();
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {
        const uint64_t a_hi = smem_desc_sw128(smem_u32(A_hi + kb * kFBlk)), a_lo = smem_desc_sw128(smem_u32(A_lo + kb * kFBlk));
        const uint64_t b_hi = smem_desc_sw128(smem_u32(W_hi + kb * kWBlk)), b_lo = smem_desc_sw128(smem_u32(W_lo + kb * kWBlk));
#pragma unroll
        for (int k = 0; k < 4; ++k) {                    // UMMA_K = 8 tf32 = 32 bytes -> +2 in the address field
          umma_tf32(tmem_base, a_lo + 2 * k, b_hi + 2 * k, kIdescFwd, (kb | k) ? 1u : 0u);
          umma_tf32(tmem_base, a_hi + 2 * k, b_lo + 2 * k, kIdescFwd, 1u);
          umma_tf32(tmem_base + (kb + 1) * H2, a_hi + 2 * k, b_hi + 2 * k, kIdescFwd, k ? 1u : 0u);
        }
      }
      umma_commit(bar_a);
    }
    // next tile's rows (and the endpoints of the one after): in flight while the tensor core and the epilogue work
    gather_issue_rows<U>(pd, ps, regs);
    gather_issue_idx<U>(src, dst, (tile + 2 * static_cast<int64_t>(gridDim.x)) * kFT, n_pairs, regs);
    // dropout decisions in the shadow of the MMAs: the next tile's hidden-1 units and this tile's hidden-2 units (epilogue)
    regs.keep = gather_keep_bits<U>((tile + gridDim.x) * kFT, drop);
    uint32_t keep2 = 0;                                   // bit 4c + r = unit j0 + 4c + r of pair base + 32q + lane survives
    if (drop.thresh && q * 32 < kFT) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const DropBits bits = dropout_bits(drop.seed, static_cast<uint32_t>(base + q * 32 + lane), H1 / 4 + (j0 >> 2) + c);
#pragma unroll
        for (int r = 0; r < 4; ++r) keep2 |= static_cast<uint32_t>(dropout_keep16(bits, r, drop.thresh)) << (c * 4 + r);
      }
    }
    mbar_wait(bar_a, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- epilogue: z2 = drop(relu(acc + b2)); partial of w3 . z2 over this thread's 16 columns ----
    if (q * 32 < kFT) {
      uint32_t v[16];
      float acc[16];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + j0;
      tmem_ld16(taddr + 1 * H2, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(v[j]);
      tmem_ld16(taddr + 2 * H2, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] += __uint_as_float(v[j]);
      float acc2[16];
      tmem_ld16(taddr + 3 * H2, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) acc2[j] = __uint_as_float(v[j]);
      tmem_ld16(taddr + 4 * H2, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] += acc2[j] + __uint_as_float(v[j]);
      tmem_ld16(taddr, v);                               // the two small cross terms
      tmem_ld_wait();
      const int p = q * 32 + lane;
      const int64_t e = base + p;
      float partial = 0.f;
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {                   // 8 units = one full 32-byte sector per store
        float z2[8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = c2 * 2 + h;
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            float x = fmaxf((acc[c * 4 + r] + __uint_as_float(v[c * 4 + r])) + b2s[j0 + c * 4 + r], 0.f);
            if (drop.thresh) x = ((keep2 >> (c * 4 + r)) & 1u) ? x * drop.scale : 0.f;
            z2[h * 4 + r] = x;
          }
        }
        // 256-bit streaming store: a 16-byte store leaves half a sector, and a partially written sector evicted
        // (evict-first) before its other half arrives costs a DRAM read-modify-write (measured: 5.2 GB of reads)
        if (z2_save && e < n_pairs) st_global_cs_v8(z2_save + e * H2 + j0 + c2 * 8, z2);
#pragma unroll
        for (int r = 0; r < 8; ++r) partial += z2[r] * b2s[H2 + j0 + c2 * 8 + r];
      }
      part[g * kFT + p] = partial;
    }
    tc_fence_before();
    __syncthreads();
    if (t < kFT && base + t < n_pairs)                  // fixed order: deterministic
      out[perm ? perm[base + t] : base + t] = ((part[t] + part[kFT + t]) + (part[2 * kFT + t] + part[3 * kFT + t])) + bias3;
    // `part` is next written two barriers from here; A and the accumulators are free (MMAs retired, TMEM read)
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_free<kFwdTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// backward: 64-pair tiles
// ------------------------------------------------------------------------------------------------
constexpr int kBT = 64;
constexpr int kBBlk = kBT * 128;                     // 8 KB: [64 pairs][32 columns]
constexpr int kBwdTmemCols = 128;                    // D1 (dz1^T: units x pairs) columns 0..63, D2 (dW2^T: units x j) 64..127
constexpr int kWtBlk = H1 * 128;                     // 16 KB: W2^T block [128 units][32 j]
constexpr int kFlushTiles = 4;                       // K = 256 pairs per TMEM accumulation of dW2 (truncation bias, see above)
constexpr size_t kBwdTcSmem = 1024 + 8 * kBBlk /*z1 hi,lo*/ + 8 * kBBlk /*dz2 hi,lo in both layouts*/ + 4 * kWtBlk + 64;
constexpr uint32_t kIdescDz1 = tf32_idesc(128, kBT, 0, 0);    // A = W2^T, B = dz2, both K-major (K = j)
constexpr uint32_t kIdescDw2 = tf32_idesc(128, H2, 1, 1);     // A = z1 MN-major, B = dz2 MN-major
constexpr int kPairGroups = kTcThreads / 16;         // 32 pair groups x 16 column groups in the dz2 phase
constexpr int kPP = kBT / kPairGroups;               // 2 pairs per thread
constexpr int kBU = kBT / (kTcThreads / 32);         // 4 pairs per warp in the gather

__global__ void __launch_bounds__(kTcThreads, 1)
decoder_bwd_tc_kernel(const int* __restrict__ src, const int* __restrict__ dst, const int* __restrict__ perm, int64_t n_pairs,
                      const float* __restrict__ pd, const float* __restrict__ ps, const float* __restrict__ w2,
                      const float* __restrict__ w3, DropCfg drop, const float* __restrict__ z2,
                      const float* __restrict__ dout, float* __restrict__ dz1, float* __restrict__ partials, int flush_tiles,
                      const int* __restrict__ pair_slot, float* __restrict__ slot_rows) {
  if (drop.seed_dev) drop.seed = *drop.seed_dev;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the swizzle atoms, kept as an offset from the __shared__ array so that the compiler
  // still knows the address space (LDS / STS instead of generic LD / ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* Z_hi = smem;                              // 4 blocks [64 pairs][32 units], MN-major (pairs are K)
  uint8_t* Z_lo = Z_hi + 4 * kBBlk;
  uint8_t* DZ_hi = Z_lo + 4 * kBBlk;                 // 2 blocks [64 pairs][32 hidden-2 units], K-major (B of the dz1 GEMM)
  uint8_t* DZ_lo = DZ_hi + 2 * kBBlk;
  uint8_t* DM_hi = DZ_lo + 2 * kBBlk;                // the same dz2 tile MN-major (B of the dW2 GEMM)
  uint8_t* DM_lo = DM_hi + 2 * kBBlk;
  uint8_t* W_hi = DM_lo + 2 * kBBlk;                 // W2^T image: 2 blocks [128 units][32 j], K-major (A of the dz1 GEMM)
  uint8_t* W_lo = W_hi + 2 * kWtBlk;
  uint64_t* bars = reinterpret_cast<uint64_t*>(W_lo + 2 * kWtBlk);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;

  for (int i = t; i < H2 * H1 / 4; i += kTcThreads) {      // w2 [j][unit] row-major -> W2^T rows = units, K = j contiguous
    const int j = i >> 5, u0 = (i & 31) * 4;
    const float4 x = __ldg(reinterpret_cast<const float4*>(w2) + i);
    const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const uint32_t off = static_cast<uint32_t>(j >> 5) * kWtBlk + sw128_off(u0 + r, (j & 31) >> 2) + (j & 3) * 4;
      float hi, lo;
      split_tf32(xv[r], hi, lo);
      *reinterpret_cast<float*>(W_hi + off) = hi;
      *reinterpret_cast<float*>(W_lo + off) = lo;
    }
  }
  const uint32_t bar1 = smem_u32(bars), bar2 = smem_u32(bars + 1);
  if (t == 0) {
    mbar_init(bar1, 1);
    mbar_init(bar2, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc<kBwdTmemCols>(tmem_slot);
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_d1 = tmem_base, tmem_d2 = tmem_base + kBT;

  // dz2 phase role: pairs pz .. pz+kPP-1, hidden-2 units jz .. jz+3
  const int pz = (t >> 4) * kPP, jz = (t & 15) * 4;
  const float4 wv3 = *reinterpret_cast<const float4*>(w3 + jz);
  float db2_p[4] = {0.f, 0.f, 0.f, 0.f}, dw3_p[4] = {0.f, 0.f, 0.f, 0.f}, db3_p = 0.f;
  // TMEM roles: lanes 32*(warp % 4) .. +31, column group warp / 4
  const int q = warp & 3, g = warp >> 2;
  float* my = partials + static_cast<size_t>(blockIdx.x) * kPartial;
  bool flushed_once = false;
  int since_flush = 0;

  // dW2^T accumulator [128 units (lanes)][64 j (columns)] -> this CTA's partial dW2 [j][unit] (L2-resident)
  auto flush_dw2 = [&]() {
    uint32_t v[16];
    tmem_ld16(tmem_d2 + (static_cast<uint32_t>(q * 32) << 16) + g * 16, v);
    tmem_ld_wait();
    const int unit = q * 32 + lane;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float* dstp = my + (g * 16 + j) * H1 + unit;
      *dstp = flushed_once ? *dstp + __uint_as_float(v[j]) : __uint_as_float(v[j]);
    }
  };
  // z2 / dout of this thread's pairs in the dz2 phase
  float4 zz[kPP];
  float go[kPP];
  auto issue_z2 = [&](int64_t base) {
#pragma unroll
    for (int i = 0; i < kPP; ++i) {
      const int64_t e = base + pz + i;
      zz[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      go[i] = 0.f;
      if (e < n_pairs) {
        zz[i] = __ldcs(reinterpret_cast<const float4*>(z2 + e * H2 + jz));
        go[i] = __ldg(dout + (perm ? __ldg(perm + e) : e));
      }
    }
  };

  uint32_t it = 0;
  const int64_t n_tiles = (n_pairs + kBT - 1) / kBT;
  GatherRegs<kBU> regs;
  int64_t tile = blockIdx.x;
  issue_z2(tile * kBT);                                                    // (past-the-end tiles load nothing)
  gather_issue_idx<kBU>(src, dst, tile * kBT, n_pairs, regs);
  gather_issue_rows<kBU>(pd, ps, regs);
  gather_issue_idx<kBU>(src, dst, (tile + gridDim.x) * kBT, n_pairs, regs);
  regs.keep = gather_keep_bits<kBU>(tile * kBT, drop);
  for (; tile < n_tiles; tile += gridDim.x, ++it) {
    const int64_t base = tile * kBT;
    gather_commit<kBT, kBU, true>(regs, base, drop, Z_hi, Z_lo);
    // ---- dz2 tile (K-major copy for the dz1 GEMM, MN-major copy for the dW2 GEMM) + db2 / dw3 / db3 partials ----
#pragma unroll
    for (int i = 0; i < kPP; ++i) {
      const float gs = go[i] * drop.scale;
      float4 d;
      d.x = zz[i].x > 0.f ? gs * wv3.x : 0.f; d.y = zz[i].y > 0.f ? gs * wv3.y : 0.f;
      d.z = zz[i].z > 0.f ? gs * wv3.z : 0.f; d.w = zz[i].w > 0.f ? gs * wv3.w : 0.f;
      const uint32_t blk = static_cast<uint32_t>(jz >> 5) * kBBlk;
      const uint32_t off_k = blk + sw128_off(pz + i, (jz & 31) >> 2), off_m = blk + mn32_off(pz + i, (jz & 31) >> 2);
      float4 hi, lo;
      split_tf32(d, hi, lo);
      *reinterpret_cast<float4*>(DZ_hi + off_k) = hi;
      *reinterpret_cast<float4*>(DZ_lo + off_k) = lo;
      *reinterpret_cast<float4*>(DM_hi + off_m) = hi;
      *reinterpret_cast<float4*>(DM_lo + off_m) = lo;
      db2_p[0] += d.x; db2_p[1] += d.y; db2_p[2] += d.z; db2_p[3] += d.w;
      dw3_p[0] = fmaf(go[i], zz[i].x, dw3_p[0]); dw3_p[1] = fmaf(go[i], zz[i].y, dw3_p[1]);
      dw3_p[2] = fmaf(go[i], zz[i].z, dw3_p[2]); dw3_p[3] = fmaf(go[i], zz[i].w, dw3_p[3]);
      if (jz == 0) db3_p += go[i];
    }
    proxy_fence();
    __syncthreads();
    if (t == 0) {
      tc_fence_after();
      // dz1^T [units x pairs] = W2^T [units x j] . dz2^T [j x pairs]: K = 64 j in 8 slices of 32 bytes along the rows
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        const uint32_t a_off = (s >> 2) * kWtBlk + (s & 3) * 32, b_off = (s >> 2) * kBBlk + (s & 3) * 32;
        mma3(tmem_d1, smem_desc_sw128(smem_u32(W_hi + a_off)), smem_desc_sw128(smem_u32(W_lo + a_off)),
             smem_desc_sw128(smem_u32(DZ_hi + b_off)), smem_desc_sw128(smem_u32(DZ_lo + b_off)), kIdescDz1, s ? 1u : 0u);
      }
      umma_commit(bar1);
      // dW2^T [units x j] += z1^T [units x pairs] . dz2 [pairs x j]: K = 64 pairs in 8 slices
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        const uint32_t off = s * 1024;
        mma3(tmem_d2, smem_desc_mn32(smem_u32(Z_hi + off), kBBlk, 512), smem_desc_mn32(smem_u32(Z_lo + off), kBBlk, 512),
             smem_desc_mn32(smem_u32(DM_hi + off), kBBlk, 512), smem_desc_mn32(smem_u32(DM_lo + off), kBBlk, 512),
             kIdescDw2, (s || since_flush) ? 1u : 0u);
      }
      umma_commit(bar2);
    }
    // next tile's rows / z2 / dout (and the endpoints of the tile after): in flight while the tensor core and the
    // epilogue work on this one
    issue_z2((tile + gridDim.x) * kBT);
    gather_issue_rows<kBU>(pd, ps, regs);
    gather_issue_idx<kBU>(src, dst, (tile + 2 * static_cast<int64_t>(gridDim.x)) * kBT, n_pairs, regs);
    regs.keep = gather_keep_bits<kBU>((tile + gridDim.x) * kBT, drop);     // next tile's masks, in the shadow of the MMAs
    // ---- dz1 epilogue: this thread holds unit 32q + lane of pairs 16g .. 16g+15; a warp stores 128 contiguous
    // bytes of one pair's row per instruction ----
    mbar_wait(bar1, it & 1);
    tc_fence_after();
    {
      uint32_t v[16];
      tmem_ld16(tmem_d1 + (static_cast<uint32_t>(q * 32) << 16) + g * 16, v);
      tmem_ld_wait();
      const float sc = drop.scale;
      const uint8_t* zrow = Z_hi + q * kBBlk + (lane & 7) * 4;             // unit 32q + lane: 32-byte chunk lane >> 3
      // Segment sums by source node, fused: `pair_slot[e]` numbers the runs of equal source inside aligned 16-pair groups
      // (the pairs are walked in an order that is sorted by source inside each label class, so a run is usually the whole
      // group). This thread owns unit 32q + lane of exactly one group, so it sums its 16 values run by run in pair order
      // and writes one row element per run into slot_rows [n_slots, 128]: no other thread touches that (slot, unit), no
      // atomics, fixed order. The host adds the few slots of each source node in slot order (a 1.5 M-row segment sum
      // instead of re-reading the 10 GB dz1).
      int sl[16];
      if (pair_slot) {
        const int64_t e0 = base + g * 16;
        if (e0 + 16 <= n_pairs) {
          const int4* ps4 = reinterpret_cast<const int4*>(pair_slot + e0);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int4 t4 = __ldg(ps4 + i);
            sl[4 * i] = t4.x; sl[4 * i + 1] = t4.y; sl[4 * i + 2] = t4.z; sl[4 * i + 3] = t4.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) sl[j] = (e0 + j < n_pairs) ? __ldg(pair_slot + e0 + j) : -1;
        }
      }
      float run = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int p = g * 16 + j;
        const int64_t e = base + p;
        const float z = *reinterpret_cast<const float*>(zrow + p * 128 + (((lane >> 3) ^ (p & 3)) << 5));
        const float val = z > 0.f ? __uint_as_float(v[j]) * sc : 0.f;
        if (e < n_pairs) __stcs(dz1 + e * H1 + q * 32 + lane, val);
        run += val;                                                        // pairs past the end contribute exact zeros
      }
      if (pair_slot) {
        float* srow = slot_rows + q * 32 + lane;
        if (sl[0] == sl[15]) {                                             // the whole group is one run (~9 groups in 10)
          if (sl[0] >= 0) srow[static_cast<int64_t>(sl[0]) * H1] = run;
        } else {                                                           // run boundaries inside the group: walk it again
          int cur = -1;
          run = 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int p = g * 16 + j;
            const float z = *reinterpret_cast<const float*>(zrow + p * 128 + (((lane >> 3) ^ (p & 3)) << 5));
            const float val = z > 0.f ? __uint_as_float(v[j]) * sc : 0.f;
            if (sl[j] != cur) {                                            // warp-uniform: every lane sees the same pairs
              if (cur >= 0) srow[static_cast<int64_t>(cur) * H1] = run;
              cur = sl[j];
              run = 0.f;
            }
            run += val;
          }
          if (cur >= 0) srow[static_cast<int64_t>(cur) * H1] = run;
        }
      }
    }
    mbar_wait(bar2, it & 1);                             // z1 / dz2 tiles free again
    tc_fence_after();
    if (++since_flush == flush_tiles) {
      flush_dw2();
      flushed_once = true;
      since_flush = 0;
    }
    tc_fence_before();
    __syncthreads();                                     // every warp is done with Z_hi / TMEM before the next tile
  }
  if (since_flush) { flush_dw2(); flushed_once = true; }
  if (!flushed_once) {                                   // CTA without tiles (n_pairs == 0)
    for (int i = t; i < H2 * H1; i += kTcThreads) my[i] = 0.f;
  }
  tc_fence_before();
  // ---- db2 / dw3 / db3: sum the pair groups for each column group in a fixed order ----
  __syncthreads();
  float* red = reinterpret_cast<float*>(Z_hi);           // [kPairGroups][16][9]
  float* mine = red + ((t >> 4) * 16 + (t & 15)) * 9;
#pragma unroll
  for (int r = 0; r < 4; ++r) { mine[r] = db2_p[r]; mine[4 + r] = dw3_p[r]; }
  mine[8] = db3_p;
  __syncthreads();
  if (t < 16) {
    float s[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int grp = 0; grp < kPairGroups; ++grp)
#pragma unroll
      for (int r = 0; r < 9; ++r) s[r] += red[(grp * 16 + t) * 9 + r];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      my[H2 * H1 + t * 4 + r] = s[r];                    // db2
      my[H2 * H1 + H2 + t * 4 + r] = s[4 + r];           // dw3
    }
    if (t == 0) my[H2 * H1 + 2 * H2] = s[8];             // db3
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_free<kBwdTmemCols>(tmem_base);
  }
}

int tc_grid(int64_t n_tiles) {
  int64_t g = kNumSM;                                    // one resident CTA per SM
  if (n_tiles < g) g = n_tiles;
  return static_cast<int>(g < 1 ? 1 : g);
}

}  // namespace

template <int kFT>
static int launch_fwd(const int* src, const int* dst, const int* perm, int64_t n_pairs, const float* pd, const float* ps, const float* w2,
                      const float* b2, const float* w3, const float* b3, DropCfg drop, float* out, float* z2_save, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    DG_CHECK_CUDA(cudaFuncSetAttribute(decoder_fwd_tc_kernel<kFT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(fwd_tc_smem<kFT>())));
    attr_set = true;
  }
  const int64_t n_tiles = (n_pairs + kFT - 1) / kFT;
  decoder_fwd_tc_kernel<kFT><<<tc_grid(n_tiles), kTcThreads, fwd_tc_smem<kFT>(), st>>>(src, dst, perm, n_pairs, pd, ps, w2, b2, w3, b3,
                                                                                      drop, out, z2_save);
  DG_CHECK_LAUNCH("decoder_fwd_tc");
  return DG_OK;
}

int launch_decoder_fwd_tc(const int* src, const int* dst, const int* perm, int64_t n_pairs, const float* pd, const float* ps, const float* w2,
                          const float* b2, const float* w3, const float* b3, DropCfg drop, float* out, float* z2_save,
                          cudaStream_t st) {
  const char* v = getenv("DG_DEC_FT");               // tuning switch: forward tile of 64 or 128 pairs
  if (v && atoi(v) == 64) return launch_fwd<64>(src, dst, perm, n_pairs, pd, ps, w2, b2, w3, b3, drop, out, z2_save, st);
  return launch_fwd<128>(src, dst, perm, n_pairs, pd, ps, w2, b2, w3, b3, drop, out, z2_save, st);
}

int launch_decoder_bwd_tc(const int* src, const int* dst, const int* perm, int64_t n_pairs, const float* pd, const float* ps, const float* w2,
                          const float* w3, DropCfg drop, const float* z2, const float* dout, float* dz1, float* partials,
                          int* n_ctas, const int* pair_slot, float* slot_rows, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    DG_CHECK_CUDA(cudaFuncSetAttribute(decoder_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kBwdTcSmem)));
    attr_set = true;
  }
  const int64_t n_tiles = (n_pairs + kBT - 1) / kBT;
  const int grid = tc_grid(n_tiles);
  const char* ft = getenv("DG_DEC_FLUSH");           // tuning switch
  decoder_bwd_tc_kernel<<<grid, kTcThreads, kBwdTcSmem, st>>>(src, dst, perm, n_pairs, pd, ps, w2, w3, drop, z2, dout, dz1, partials,
                                                              ft ? atoi(ft) : kFlushTiles, pair_slot, slot_rows);
  DG_CHECK_LAUNCH("decoder_bwd_tc");
  *n_ctas = grid;
  return DG_OK;
}

}  // namespace dg
