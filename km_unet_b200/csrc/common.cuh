// common.cuh -- shared host/device helpers of libkmunet (error channel, launch accounting, small math).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kmunet.h"

namespace kmu {

void set_error(const char* fmt, ...);
int finish_launch(const char* what);  // counts the launch; returns KMU_OK or KMU_ERR_LAUNCH with message set
void count_launches(int n);
bool deterministic();          // kmu_set_deterministic / KMU_DETERMINISTIC=1: bit-reproducible variants where a faster atomic one exists
void set_deterministic(int on);

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

#define KMU_REQUIRE(cond, code, ...)   \
  do {                                 \
    if (!(cond)) {                     \
      ::kmu::set_error(__VA_ARGS__);   \
      return (code);                   \
    }                                  \
  } while (0)

#define KMU_LAUNCH_CHECK(what)                       \
  do {                                               \
    int _st = ::kmu::finish_launch(what);            \
    if (_st != KMU_OK) return _st;                   \
  } while (0)

// precise variants (fp32 family); the tensor-core producers use their own fast-math versions
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float siluf_(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float silu_gradf_(float x) {
  float s = 1.0f / (1.0f + expf(-x));
  return s * (1.0f + x * (1.0f - s));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace kmu
