// common.cuh - small device/host helpers shared by the kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define TPP_CUDA_CHECK(expr)                                                            \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      fprintf(stderr, "tpp-xsmm-cuda: %s failed at %s:%d: %s\n", #expr, __FILE__,       \
              __LINE__, cudaGetErrorString(_e));                                        \
      exit(-1);                                                                         \
    }                                                                                   \
  } while (0)

namespace tpp {

constexpr int64_t kF32 = 1;
constexpr int64_t kBF16 = 2;

__device__ __forceinline__ float bf16_bits_to_f32(uint16_t h) {
  return __uint_as_float(static_cast<uint32_t>(h) << 16);
}

// round-to-nearest-even f32 -> bf16 bits (same rule as the oracle's
// xo_f32_to_bf16 / mlir Float16bits.h for every finite value)
__device__ __forceinline__ uint16_t f32_to_bf16_bits(float f) {
  return __bfloat16_as_ushort(__float2bfloat16_rn(f));
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  return static_cast<uint32_t>(f32_to_bf16_bits(lo)) |
         (static_cast<uint32_t>(f32_to_bf16_bits(hi)) << 16);
}

// relu as a select: negative, -0 and NaN all become +0 (matches the oracle)
__device__ __forceinline__ float relu_f32(float x) { return x > 0.0f ? x : 0.0f; }

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

} // namespace tpp
