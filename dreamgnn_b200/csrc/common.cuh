// Shared helpers for libdreamgnn.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dreamgnn.h"

namespace dg {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

void set_error(const char* fmt, ...);
void count_launch();

inline cudaStream_t as_stream(dg_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Every launcher ends with this: counts the launch and turns a launch error into a return code.
#define DG_CHECK_LAUNCH(name)                                                   \
  do {                                                                          \
    ::dg::count_launch();                                                       \
    cudaError_t e__ = cudaGetLastError();                                       \
    if (e__ != cudaSuccess) {                                                   \
      ::dg::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return static_cast<int>(e__);                                             \
    }                                                                           \
  } while (0)

#define DG_CHECK_CUDA(call)                                                         \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      ::dg::set_error("%s failed: %s", #call, cudaGetErrorString(e__));             \
      return static_cast<int>(e__);                                                 \
    }                                                                               \
  } while (0)

#define DG_REQUIRE(cond, msg)                                   \
  do {                                                          \
    if (!(cond)) {                                              \
      ::dg::set_error("%s: %s", __func__, msg);                 \
      return DG_ERR_INVALID_ARGUMENT;                           \
    }                                                           \
  } while (0)

#define DG_PROPAGATE(call)      \
  do {                          \
    int rc__ = (call);          \
    if (rc__ != 0) return rc__; \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-provided workspace.
struct Workspace {
  char* base;
  size_t size, off;
  Workspace(void* p, size_t n) : base(static_cast<char*>(p)), size(n), off(0) {}
  template <typename T>
  T* take(size_t count) {
    size_t o = align_up(off, 256);
    size_t end = o + count * sizeof(T);
    if (end > size || base == nullptr) return nullptr;
    off = end;
    return reinterpret_cast<T*>(base + o);
  }
};
inline size_t ws_add(size_t acc, size_t bytes) { return align_up(acc, 256) + bytes; }

// 148 SMs on B200; grids for grid-stride kernels are sized as a multiple of this.
constexpr int kNumSM = 148;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// Streaming 128-bit loads/stores that do not allocate in L1 (gathered rows are not re-used by the
// same SM; index arrays are read once).
__device__ __forceinline__ float4 ldg_f4_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ldg_u4_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ int ldg_i32_stream(const int* p) {
  int r;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}

// 256-bit streaming store (sm_100: STG.256): one full 32-byte sector per thread
__device__ __forceinline__ void st_global_cs_v8(float* p, const float (&v)[8]) {
  asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}

// Counter-based dropout generator shared by forward and backward kernels: the keep decisions of the four
// units (4g .. 4g+3) of pair / row `a` under `seed` are a pure function, so the backward regenerates the
// forward's mask. Two rounds of a 32-bit integer hash give four 16-bit uniforms per call (one call per
// float4), i.e. ~4 integer instructions per element instead of a 64-bit mix per element.
__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du;
  x ^= x >> 15; x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
struct DropBits { uint32_t lo, hi; };
__device__ __forceinline__ DropBits dropout_bits(uint64_t seed, uint32_t a, uint32_t g) {
  DropBits b;
  b.lo = hash32((a * 0x9e3779b1u) ^ (g * 0x85ebca77u) ^ static_cast<uint32_t>(seed));
  b.hi = hash32(b.lo ^ static_cast<uint32_t>(seed >> 32) ^ 0xc2b2ae3du);
  return b;
}
// q-th (0..3) 16-bit uniform of the group; keep iff it is >= thresh16 = round(p * 65536)
__device__ __forceinline__ bool dropout_keep16(const DropBits& b, int q, uint32_t thresh16) {
  const uint32_t w = (q & 2) ? b.hi : b.lo;
  return ((q & 1) ? (w >> 16) : (w & 0xffffu)) >= thresh16;
}

}  // namespace dg
