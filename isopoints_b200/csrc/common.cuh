// isob200 -- shared device/host helpers for the sm_100a kernels.
// All kernels in this library are stream-ordered, never allocate, and report
// errors through an int status + isob200_last_error().
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ISOB200_OK 0
#define ISOB200_ERR_INVALID 1   // bad argument (shape, K out of range, null pointer)
#define ISOB200_ERR_CUDA 2      // CUDA runtime error (launch failure, ...)
#define ISOB200_ERR_WORKSPACE 3 // workspace too small

namespace isob200 {

void set_error(const char* fmt, ...);
void count_launch();  // bumps the library-wide kernel launch counter (isob200_launch_count)

constexpr int kNumSMs = 148;  // B200

static inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// Grid size for grid-stride kernels: enough CTAs to fill all 148 SMs, a multiple of the SM count.
static inline int grid_for(long long work_items, int block, int ctas_per_sm) {
  long long need = (work_items + block - 1) / block;
  long long cap = (long long)kNumSMs * ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

}  // namespace isob200

#define ISO_CHECK_ARG(cond, ...)                   \
  do {                                             \
    if (!(cond)) {                                 \
      isob200::set_error(__VA_ARGS__);             \
      return ISOB200_ERR_INVALID;                  \
    }                                              \
  } while (0)

#define ISO_CHECK_LAUNCH(name)                                                        \
  do {                                                                                \
    cudaError_t e__ = cudaGetLastError();                                             \
    if (e__ != cudaSuccess) {                                                         \
      isob200::set_error("%s: CUDA error: %s", name, cudaGetErrorString(e__));        \
      return ISOB200_ERR_CUDA;                                                        \
    }                                                                                 \
    isob200::count_launch();                                                          \
  } while (0)

#define ISO_CUDA(call)                                                                \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      isob200::set_error("%s: CUDA error: %s", #call, cudaGetErrorString(e__));       \
      return ISOB200_ERR_CUDA;                                                        \
    }                                                                                 \
  } while (0)

// ---- FRNN grid parameter layout (same slots as the reference: grid.h:5-24) ----
#define ISO_G3_MIN_X 0
#define ISO_G3_MIN_Y 1
#define ISO_G3_MIN_Z 2
#define ISO_G3_DELTA 3
#define ISO_G3_RES_X 4
#define ISO_G3_RES_Y 5
#define ISO_G3_RES_Z 6
#define ISO_G3_TOTAL 7
#define ISO_G3_SIZE 8
#define ISO_G3_MAX_RES 128
#define ISO_G2_MIN_X 0
#define ISO_G2_MIN_Y 1
#define ISO_G2_DELTA 2
#define ISO_G2_RES_X 3
#define ISO_G2_RES_Y 4
#define ISO_G2_TOTAL 5
#define ISO_G2_SIZE 6
#define ISO_G2_MAX_RES 1024
