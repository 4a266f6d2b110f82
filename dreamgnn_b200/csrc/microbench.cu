// Read-bandwidth microbenchmark behind profiles/l2_peak.json (scripts/l2_peak.py): the denominators of the SpMM
// roofline. The SpMM kernels are row gathers served partly by L2, so next to the HBM copy peak of MEASURED_PEAKS.json
// the repo measures, with the SAME load instruction and access shape the SpMM uses (one warp per row, 128-bit
// ld.global.nc.L1::no_allocate per lane, several rows in flight per warp):
//   * sequential rows over a buffer that fits L2      -> L2 -> SM streaming bandwidth
//   * pseudo-random rows over a buffer that fits L2   -> L2 -> SM gather bandwidth at a given row width
//   * the same over a buffer much larger than L2      -> HBM gather bandwidth at that row width
#include "common.cuh"

namespace dg {
namespace {

constexpr int kBenchThreads = 256;

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du;
  x ^= x >> 15; x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}

template <int kInFlight>                       // rows in flight per warp (the SpMM instances keep 4-8)
__global__ void __launch_bounds__(kBenchThreads)
bench_read_kernel(const float4* __restrict__ buf, int64_t n_rows, int row_f4, int64_t rows_per_warp, int random,
                  float* __restrict__ sink) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * kBenchThreads + threadIdx.x) >> 5;
  const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * kBenchThreads) >> 5;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  // row ids cost a handful of integer instructions (multiply-shift range reduction / add-and-wrap): the loop must stay
  // load-bound even for 512-byte rows
  const uint32_t nr = static_cast<uint32_t>(n_rows);
  const uint32_t salt = mix32(static_cast<uint32_t>(warp) * 0x9e3779b1u + 0x7f4a7c15u);
  uint32_t seq = static_cast<uint32_t>(warp % n_rows);
  const uint32_t step = static_cast<uint32_t>(n_warps % n_rows);
  for (int64_t i = 0; i < rows_per_warp; i += kInFlight) {
    const float4* rp[kInFlight];
#pragma unroll
    for (int u = 0; u < kInFlight; ++u) {
      uint32_t r;
      if (random) {
        r = __umulhi(mix32(salt ^ static_cast<uint32_t>(i + u)), nr);
      } else {
        r = seq;
        seq += step;
        if (seq >= nr) seq -= nr;
      }
      rp[u] = buf + static_cast<int64_t>(r) * row_f4;
    }
    for (int c = lane; c < row_f4; c += 32) {
      float4 v[kInFlight];
#pragma unroll
      for (int u = 0; u < kInFlight; ++u) v[u] = ldg_f4_stream(rp[u] + c);
#pragma unroll
      for (int u = 0; u < kInFlight; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
  }
  if (acc.x + acc.y + acc.z + acc.w == 1234.5678f) sink[0] = acc.x;       // keeps the loads alive
}

}  // namespace
}  // namespace dg

extern "C" int dg_bench_read_rows(const float* buf, int64_t n_rows, int64_t row_floats, int64_t rows_per_warp, int random,
                                  int ctas_per_sm, int rows_in_flight, float* sink, dg_stream_t stream) {
  using namespace dg;
  DG_REQUIRE(buf != nullptr && sink != nullptr, "null pointer");
  DG_REQUIRE(n_rows > 0 && n_rows < (1ll << 31) && row_floats > 0 && row_floats % 4 == 0 && rows_per_warp > 0, "bad shape");
  DG_REQUIRE(rows_in_flight == 2 || rows_in_flight == 4 || rows_in_flight == 8 || rows_in_flight == 16, "rows_in_flight: 2, 4, 8 or 16");
  DG_REQUIRE(rows_per_warp % rows_in_flight == 0, "rows_per_warp must be a multiple of rows_in_flight");
  DG_REQUIRE(ctas_per_sm >= 1 && ctas_per_sm <= 8, "ctas_per_sm out of range");
  const float4* b4 = reinterpret_cast<const float4*>(buf);
  const int f4 = static_cast<int>(row_floats / 4);
  const unsigned grid = kNumSM * ctas_per_sm;
  cudaStream_t st = as_stream(stream);
  switch (rows_in_flight) {
    case 2: bench_read_kernel<2><<<grid, kBenchThreads, 0, st>>>(b4, n_rows, f4, rows_per_warp, random, sink); break;
    case 4: bench_read_kernel<4><<<grid, kBenchThreads, 0, st>>>(b4, n_rows, f4, rows_per_warp, random, sink); break;
    case 8: bench_read_kernel<8><<<grid, kBenchThreads, 0, st>>>(b4, n_rows, f4, rows_per_warp, random, sink); break;
    default: bench_read_kernel<16><<<grid, kBenchThreads, 0, st>>>(b4, n_rows, f4, rows_per_warp, random, sink); break;
  }
  DG_CHECK_LAUNCH("bench_read_rows");
  return DG_OK;
}
