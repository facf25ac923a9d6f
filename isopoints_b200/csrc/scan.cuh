#pragma once
#include "common.cuh"
namespace isob200 {
size_t scan_ws_bytes(int n, int rows);
// out[row][i] = sum_{j<i} in[row][j]; rows are `*_stride` elements apart. `out` may alias `in`.
int exclusive_scan_i32(const int* in, int* out, int n, int rows, long long in_stride,
                       long long out_stride, void* ws, size_t ws_bytes, cudaStream_t stream);
}  // namespace isob200
