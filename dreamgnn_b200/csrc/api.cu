// Library bookkeeping: ABI version, thread-local error string, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace dg {
namespace {
thread_local char g_error[512] = "";
std::atomic<unsigned long long> g_launches{0};
}  // namespace

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace dg

extern "C" {
int dg_abi_version(void) { return DG_ABI_VERSION; }
const char* dg_last_error(void) { return dg::g_error; }
unsigned long long dg_launch_count(void) { return dg::g_launches.load(std::memory_order_relaxed); }
void dg_reset_launch_count(void) { dg::g_launches.store(0, std::memory_order_relaxed); }
}
