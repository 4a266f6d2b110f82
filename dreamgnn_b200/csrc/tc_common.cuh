// tcgen05 / TMEM / mbarrier / TMA PTX wrappers shared by the tensor-core kernels (gemm_tc.cu, decoder_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dg {

constexpr uint32_t kSpinLimit = 1u << 22;                // watchdog: trap instead of hanging the GPU

// ---- PTX wrappers ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > kSpinLimit) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// same, with an L2 eviction-priority hint (createpolicy result) for operands that other CTAs re-read
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_load_3d_hint(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format): start>>4 | LBO(ignored)=1 |
// SBO = 1024 B between 8-row groups | version 1 | layout SWIZZLE_128B (= 2 in bits 61..63)
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr) {
  return static_cast<uint64_t>((addr & 0x3ffff) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major, SWIZZLE_128B descriptor: the operand is stored [k][mn] with 32 fp32 (128 B) of MN contiguous per
// row and 8 K-rows per 1024-byte swizzle atom; LBO = bytes between 32-element blocks along MN, SBO = bytes
// between 8-row groups along K (canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units).
__device__ __forceinline__ uint64_t smem_desc_sw128_mn(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return static_cast<uint64_t>((addr & 0x3ffff) >> 4) | (static_cast<uint64_t>(lbo_bytes >> 4) << 16) |
         (static_cast<uint64_t>(sbo_bytes >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor for kind::tf32 with an fp32 accumulator: D = f32 (bits 4-5 = 1), A/B = tf32 (2 at
// bits 7-9 / 10-12), a_major / b_major at bits 15 / 16 (0 = K-major, 1 = MN-major), N >> 3 at bits 17-22,
// M >> 4 at bits 24-28
__host__ __device__ constexpr uint32_t tf32_idesc(int m, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}
// byte offset of 16-byte chunk `chunk` (0..7) of row `row` inside a 128-byte-swizzled block whose base is
// 1024-byte aligned (rows are 128 B apart)
__device__ __forceinline__ uint32_t sw128_off(int row, int chunk) {
  return static_cast<uint32_t>(row) * 128u + (static_cast<uint32_t>(chunk ^ (row & 7)) << 4);
}
__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// MN-major fp32 / tf32 operands only exist in the 128-byte swizzle with a 32-byte base (layout type 1):
// rows of 32 MN-contiguous elements (128 B), one row per K index, 4 K-rows per 512-byte swizzle atom, the 32-byte
// chunk index of a row XOR-ed with (k & 3). LBO = bytes between 32-element blocks along MN, SBO = bytes between
// 4-row groups along K (512 for densely packed rows). One kind::tf32 MMA (K = 8) reads 8 consecutive rows.
__device__ __forceinline__ uint64_t smem_desc_mn32(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return static_cast<uint64_t>((addr & 0x3ffff) >> 4) | (static_cast<uint64_t>(lbo_bytes >> 4) << 16) |
         (static_cast<uint64_t>(sbo_bytes >> 4) << 32) | (1ull << 46) | (1ull << 61);
}
// byte offset of 16-byte chunk `chunk16` (0..7) of K-row `k` inside such a block (base 1024-byte aligned)
__device__ __forceinline__ uint32_t mn32_off(int k, int chunk16) {
  return static_cast<uint32_t>(k) * 128u + (static_cast<uint32_t>((chunk16 >> 1) ^ (k & 3)) << 5) +
         (static_cast<uint32_t>(chunk16 & 1) << 4);
}
// round-to-nearest tf32 split: hi = rna_tf32(x), lo = rna_tf32(x - hi); both are exact tf32 values, so the
// tensor core's operand truncation is a no-op and the dropped terms are ~2^-23 relative
// (integer form of cvt.rna.tf32.f32 -- add half an ulp of the 10-bit mantissa to the magnitude, clear the low 13
// bits; the cvt instruction runs on the quarter-rate conversion pipe, these two run on the integer ALUs)
__device__ __forceinline__ float rna_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = rna_tf32(x);
  lo = rna_tf32(x - hi);
}
__device__ __forceinline__ void split_tf32(const float4& x, float4& hi, float4& lo) {
  split_tf32(x.x, hi.x, lo.x); split_tf32(x.y, hi.y, lo.y); split_tf32(x.z, hi.z, lo.z); split_tf32(x.w, hi.w, lo.w);
}

}  // namespace dg
