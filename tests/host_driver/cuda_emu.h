// TEST INFRASTRUCTURE ONLY: the handful of CUDA language pieces the simple grid-stride kernels of this library use, for compiling a
// kernel header (e.g. csrc/jvp_kernels.cuh) as plain C++ and running it thread by thread on the host.  No shared memory, no warp
// intrinsics, no atomics: kernels that need those are not emulated.
#pragma once
#include <math.h>
#include <stdint.h>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline

struct emu_dim3 { unsigned x, y, z; };
static emu_dim3 blockIdx, blockDim, gridDim, threadIdx;

static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }

// run `kernel(args...)` for every thread of a 1-D grid, serially
#define EMU_LAUNCH(grid, block, kernel, ...)                       \
  do {                                                            \
    gridDim = {(unsigned)(grid), 1, 1};                           \
    blockDim = {(unsigned)(block), 1, 1};                         \
    for (unsigned _b = 0; _b < (unsigned)(grid); ++_b)            \
      for (unsigned _t = 0; _t < (unsigned)(block); ++_t) {       \
        blockIdx = {_b, 0, 0};                                    \
        threadIdx = {_t, 0, 0};                                   \
        kernel(__VA_ARGS__);                                      \
      }                                                           \
  } while (0)
