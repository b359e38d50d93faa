// Shared host/device helpers for libtinyrec (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>

#include "../../include/tinyrec.h"

namespace tnr {

// ---- error plumbing (thread-local message, C-ABI returns int) -------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define TNR_CHECK_CUDA(expr)                                   \
  do {                                                         \
    cudaError_t _e = (expr);                                   \
    if (_e != cudaSuccess) return ::tnr::cuda_fail(_e, #expr); \
  } while (0)

#define TNR_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      ::tnr::set_error(__VA_ARGS__);  \
      return 1;                       \
    }                                 \
  } while (0)

#define TNR_LAUNCH_CHECK() TNR_CHECK_CUDA(cudaGetLastError())

int num_sms();           // cached SM count of the current device (148 on B200)

// ---- device helpers ---------------------------------------------------------
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

__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float bf16_to_f(__nv_bfloat16 v) { return __bfloat162float(v); }

// erf-GELU exactly as torch F.gelu (transformers ACT2FN['gelu'])
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// 16-byte vector of 8 bf16
struct __align__(16) bf16x8 { uint32_t u[4]; };

__device__ __forceinline__ void unpack8(const bf16x8& v, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = bf16_lo(v.u[i]); f[2 * i + 1] = bf16_hi(v.u[i]); }
}
__device__ __forceinline__ bf16x8 pack8(const float* f) {
  bf16x8 v;
#pragma unroll
  for (int i = 0; i < 4; ++i) v.u[i] = pack_bf16(f[2 * i], f[2 * i + 1]);
  return v;
}

}  // namespace tnr
