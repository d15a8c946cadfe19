// dtype helpers: 16-byte vector load/store of fp32 or bf16 rows as float lanes.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "rt_internal.h"

namespace rt {

using bf16 = __nv_bfloat16;

template <typename T>
struct VecT;
template <>
struct VecT<float> {
  static constexpr int N = 4;
};
template <>
struct VecT<bf16> {
  static constexpr int N = 8;
};

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) {
  return v;
}
template <>
__device__ __forceinline__ bf16 from_f<bf16>(float v) {
  return __float2bfloat16_rn(v);
}

// load VecT<T>::N consecutive elements (16-byte aligned) as floats
__device__ __forceinline__ void ldvec(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void ldvec(const bf16* p, float (&v)[8]) {
  uint4 t = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ void stvec(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void stvec(bf16* p, const float (&v)[8]) {
  uint4 t;
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  t.x = w[0]; t.y = w[1]; t.z = w[2]; t.w = w[3];
  *reinterpret_cast<uint4*>(p) = t;
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

// exact-ish gelu(tanh) used by the fp32 path; the bf16 tensor-core epilogue uses tanh.approx
__device__ __forceinline__ float gelu_tanh_ref(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float inner = k0 * (x + k1 * x * x * x);
  return 0.5f * x * (1.0f + tanhf(inner));
}

#define RT_DISPATCH_DTYPE(dt, T, ...)                            \
  do {                                                           \
    if ((dt) == RT_BF16) {                                       \
      using T = ::rt::bf16;                                      \
      __VA_ARGS__;                                               \
    } else if ((dt) == RT_F32) {                                 \
      using T = float;                                           \
      __VA_ARGS__;                                               \
    } else {                                                     \
      throw ::rt::Error(RT_ERR_INVALID, "unknown dtype");        \
    }                                                            \
  } while (0)

inline void count_launch() { ++g_launch_count; }
#define RT_POST_LAUNCH()                 \
  do {                                   \
    ::rt::count_launch();                \
    RT_CHECK_CUDA(cudaGetLastError());   \
  } while (0)

}  // namespace rt
