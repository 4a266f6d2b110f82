// Many small device-to-device copies in one launch: the ~60 tensors of a freshly drawn augmentation (CSRs of both
// orientations, kNN adjacencies, normalisers, noised features) written into their persistent staging buffers
// (graphed.StagedAugmentation.refresh). As separate copies they are ~60 nodes on the augmentation branch of every captured
// iteration; here the (dst, src, bytes) triples travel BY VALUE in the kernel arguments, one CTA per 16 KB chunk of one
// tensor. (Own kernel: plain loads and stores.)
#include "common.cuh"

namespace dg {
namespace {

constexpr int kCopyBatch = DG_COPY_MAX_PER_LAUNCH;   // 96
constexpr int kCopyThreads = 256;
constexpr int64_t kCopyChunk = 16384;                // bytes per CTA

struct CopyBatch {
  dg_copy_t item[kCopyBatch];
  int chunk_start[kCopyBatch + 1];
  int n;
};

__global__ void __launch_bounds__(kCopyThreads) multi_copy_kernel(const __grid_constant__ CopyBatch b) {
  int lo = 0, hi = b.n - 1;                           // last item whose first chunk is <= this CTA
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (b.chunk_start[mid] <= static_cast<int>(blockIdx.x)) lo = mid; else hi = mid - 1;
  }
  const dg_copy_t& it = b.item[lo];
  const int64_t off = static_cast<int64_t>(blockIdx.x - b.chunk_start[lo]) * kCopyChunk;
  const int64_t len = (it.bytes - off < kCopyChunk) ? it.bytes - off : kCopyChunk;
  const char* src = static_cast<const char*>(it.src) + off;
  char* dst = static_cast<char*>(it.dst) + off;
  if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
    const int64_t n16 = len >> 4;
    for (int64_t i = threadIdx.x; i < n16; i += kCopyThreads)
      reinterpret_cast<uint4*>(dst)[i] = ldg_u4_stream(reinterpret_cast<const uint4*>(src) + i);
    for (int64_t i = (n16 << 4) + threadIdx.x; i < len; i += kCopyThreads) dst[i] = src[i];
  } else {
    for (int64_t i = threadIdx.x; i < len; i += kCopyThreads) dst[i] = src[i];
  }
}

}  // namespace
}  // namespace dg

extern "C" {

int dg_multi_copy(const dg_copy_t* items, int n, dg_stream_t stream) {
  using namespace dg;
  DG_REQUIRE(n >= 0 && (n == 0 || items != nullptr), "bad item list");
  cudaStream_t st = as_stream(stream);
  int i = 0;
  while (i < n) {
    CopyBatch b;
    b.n = 0;
    int chunks = 0;
    for (; i < n && b.n < kCopyBatch; ++i) {
      DG_REQUIRE(items[i].bytes >= 0, "negative length");
      if (items[i].bytes == 0) continue;
      DG_REQUIRE(items[i].dst != nullptr && items[i].src != nullptr, "null pointer");
      const int64_t c = (items[i].bytes + kCopyChunk - 1) / kCopyChunk;
      if (chunks + c > (1 << 30)) break;
      b.item[b.n] = items[i];
      b.chunk_start[b.n] = chunks;
      chunks += static_cast<int>(c);
      ++b.n;
    }
    b.chunk_start[b.n] = chunks;
    if (chunks == 0) { if (i >= n) break; continue; }
    multi_copy_kernel<<<chunks, kCopyThreads, 0, st>>>(b);
    DG_CHECK_LAUNCH("multi_copy");
  }
  return DG_OK;
}

}  // extern "C"
