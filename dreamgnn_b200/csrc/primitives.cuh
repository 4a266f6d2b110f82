// Internal (non-ABI) declarations of the index primitives so other translation units can chain them.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace dg {
size_t scan_workspace_bytes(int64_t n);
int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, void* ws, size_t ws_bytes, cudaStream_t st);
size_t sort_workspace_bytes(int64_t n);
int sort_pairs_u64(uint64_t* keys_in, int32_t* vals_in, uint64_t* keys_out, int32_t* vals_out, int64_t n,
                   int key_bits, void* ws, size_t ws_bytes, cudaStream_t st);
}  // namespace dg
