// Shared host/device helpers for libtinyrec (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <atomic>

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

// Dynamic shared-memory opt-in of a kernel, done once PER DEVICE (cudaFuncSetAttribute applies to the current
// device's copy of the function) and safe under concurrent host threads: a process that launches on a second
// device, or from the loader thread, gets the attribute set there too.  `bytes` may grow between calls.
constexpr int MAX_DEVICES = 64;
#define TNR_SET_SMEM(kern, bytes)                                                                         \
  do {                                                                                                    \
    static std::atomic<int> tnr_smem_done_[::tnr::MAX_DEVICES];                                           \
    int tnr_dev_ = -1;                                                                                    \
    TNR_CHECK_CUDA(cudaGetDevice(&tnr_dev_));                                                             \
    TNR_REQUIRE(tnr_dev_ >= 0 && tnr_dev_ < ::tnr::MAX_DEVICES, "device index %d out of range", tnr_dev_); \
    if (tnr_smem_done_[tnr_dev_].load(std::memory_order_acquire) < (int)(bytes)) {                         \
      TNR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
      tnr_smem_done_[tnr_dev_].store((int)(bytes), std::memory_order_release);                            \
    }                                                                                                     \
  } while (0)

int num_sms();           // cached SM count of the current device (148 on B200)
int sm_reserve();        // SMs the persistent GEMM grids leave free on the current device (tnr_set_sm_reserve)

// ---- device helpers ---------------------------------------------------------
// threadIdx.x / 32 read through a shuffle: the compiler then treats the warp index -- and every row / item index derived
// from it -- as warp-uniform, and stops wrapping the warp collectives behind `if (row < rows)` in divergence handling
// (BRA.DIV + WARPSYNC per shuffle: a quarter of the LayerNorm forward's instructions)
__device__ __forceinline__ int uniform_warp_id() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
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

// erf-GELU (torch F.gelu, transformers ACT2FN['gelu']) for the GEMM epilogues, which are
// instruction-issue bound at K = 768 (about 24 issue slots per output element; erff() alone costs
// more).  Phi(z) = 0.5 (1 + erf(z / sqrt 2)) is evaluated as sigmoid(2 q(z)) with the odd quintic
// q(z) = z (a + b z^2 + c z^4) fitted (minimax) to atanh(erf(z / sqrt 2)):
//   max |z Phi_approx - gelu_erf(z)| = 2.6e-5 over all z (fp32), below bf16 output resolution for
//   |gelu| > 0.013 and 20x closer to the erf form than the classic tanh GELU (4.7e-4);
//   7 instructions: 3 FMA, 2 FMUL/FADD, MUFU.EX2, MUFU.RCP.  The coefficients below are
//   -2 log2(e) * (a, b, c).  The derivative uses the same Phi plus the exact Gaussian density.
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_erf(float z) {
  const float zc = fminf(fmaxf(z, -8.0f), 8.0f);      // the fit holds on |z| <= 8, where Phi is 0 / 1 to 1e-12
  const float u = zc * zc;
  const float w = fmaf(fmaf(0.001014264184050262f, u, -0.10677573084831238f), u, -2.301121234893799f);
  return __fdividef(z, 1.0f + ex2_approx(w * zc));    // z * sigmoid(2 q)
}
// d/dz of the same approximant: s + z s (1 - s) 2 q'(z), s = sigmoid(2 q); max |.. - gelu_erf'| = 1.1e-4
__device__ __forceinline__ float gelu_erf_grad(float z) {
  const float zc = fminf(fmaxf(z, -8.0f), 8.0f);
  const float u = zc * zc;
  const float w = fmaf(fmaf(0.001014264184050262f, u, -0.10677573084831238f), u, -2.301121234893799f);
  const float e = ex2_approx(w * zc);
  const float s = __fdividef(1.0f, 1.0f + e);
  const float qp2 = fmaf(fmaf(-0.003515171750f, u, 0.2220338916f), u, 1.595015762f);   // 2 q'(z)
  return fmaf(s * s * e * qp2, zc, s);
}

// ---- packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2: two fp32 lanes per issue slot) ----------------
// The K = 768 GEMM epilogues have ~45 issue slots per output element and thread; the scalar GELU alone took
// ~10 of the ~24 they used, which made the FFN1 / dGELU GEMMs epilogue-bound (1050 vs 1250 TFLOP/s).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// sigmoid(2 q(z)) for a pair: s = 1 / (1 + 2^(w(u) z)), u = min(z^2, 64).  Clamping u (not z) keeps the
// fitted range of the polynomial and still saturates: |z| > 8 gives 2^(-4.98 |z|) -> 0 or inf -> s = 1 or 0.
__device__ __forceinline__ void gelu_sig2(f32x2 z, float& s0, float& s1) {
  float u0, u1;
  upk2(mul2(z, z), u0, u1);
  const f32x2 u = pk2(fminf(u0, 64.0f), fminf(u1, 64.0f));
  const f32x2 w = fma2(fma2(pk2(0.001014264184050262f, 0.001014264184050262f), u,
                            pk2(-0.10677573084831238f, -0.10677573084831238f)), u,
                       pk2(-2.301121234893799f, -2.301121234893799f));
  float a0, a1;
  upk2(mul2(w, z), a0, a1);
  float d0, d1;
  upk2(add2(pk2(ex2_approx(a0), ex2_approx(a1)), pk2(1.0f, 1.0f)), d0, d1);
  s0 = rcp_approx(d0);
  s1 = rcp_approx(d1);
}
// same function as gelu_erf() / gelu_erf_grad(), two elements per call
__device__ __forceinline__ f32x2 gelu_erf2(f32x2 z) {
  float s0, s1;
  gelu_sig2(z, s0, s1);
  return mul2(z, pk2(s0, s1));
}
__device__ __forceinline__ f32x2 gelu_erf_grad2(f32x2 z) {
  float s0, s1;
  gelu_sig2(z, s0, s1);
  const f32x2 s = pk2(s0, s1);
  float u0, u1;
  upk2(mul2(z, z), u0, u1);
  const f32x2 u = pk2(fminf(u0, 64.0f), fminf(u1, 64.0f));
  const f32x2 qp2 = fma2(fma2(pk2(-0.003515171750f, -0.003515171750f), u, pk2(0.2220338916f, 0.2220338916f)), u,
                         pk2(1.595015762f, 1.595015762f));                       // 2 q'(z)
  const f32x2 ss = fma2(mul2(s, pk2(-1.0f, -1.0f)), s, s);                       // s (1 - s) = s - s^2 (no inf * 0)
  return fma2(mul2(ss, qp2), z, s);
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


// ---- counter-based dropout (Philox4x32-7) -----------------------------------------
// One Philox call yields 8 x 16-bit lanes: element e of group g is KEPT iff lane_e >= thr16,
// thr16 = round(p * 65536).  ctr = (g_lo, g_hi, site, 0), key = (seed_lo, seed_hi).
// The numpy restatement used by the tests is oracle/dropout.py.
__device__ __forceinline__ uint4 philox4x32_7(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
  }
  return c;
}

struct DropCfg {            // device-side view of tnr_dropout
  uint64_t seed; uint32_t site; uint32_t thr16; float scale;   // thr16 == 0: disabled
};

__device__ __forceinline__ DropCfg load_drop(const tnr_dropout& d) {
  DropCfg c;
  c.site = d.site;
  if (d.seed == nullptr || !(d.p > 0.f)) { c.seed = 0; c.thr16 = 0; c.scale = 1.f; return c; }
  c.seed = *d.seed;
  c.thr16 = min((uint32_t)(d.p * 65536.0f + 0.5f), 65535u);
  c.scale = 1.0f / (1.0f - d.p);
  return c;
}

// bit e of the result = keep flag of element e (0..7) of group g
__device__ __forceinline__ uint32_t dropout_keep8(const DropCfg& c, uint64_t g) {
  const uint4 r = philox4x32_7(make_uint4((uint32_t)g, (uint32_t)(g >> 32), c.site, 0u),
                               make_uint2((uint32_t)c.seed, (uint32_t)(c.seed >> 32)));
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m |= ((w[i] & 0xffffu) >= c.thr16 ? 1u : 0u) << (2 * i);
    m |= ((w[i] >> 16) >= c.thr16 ? 1u : 0u) << (2 * i + 1);
  }
  return m;
}

// the same keep flags as dropout_keep8(), delivered as the four packed multiplier pairs (scale or 0) an epilogue applies
// with one FMUL2 / FFMA2 per pair: the high lane of a word compares in place against thr16 << 16 (its low bits
// cannot change the outcome), the low lane after one shift -- 2.5 instructions per element instead of 3.5.
// thr16 <= 65535 (load_drop).
__device__ __forceinline__ void dropout_mul8(const DropCfg& c, uint64_t g, f32x2* m) {
  const uint4 r = philox4x32_7(make_uint4((uint32_t)g, (uint32_t)(g >> 32), c.site, 0u),
                               make_uint2((uint32_t)c.seed, (uint32_t)(c.seed >> 32)));
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  const uint32_t thr_hi = c.thr16 << 16;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    m[i] = pk2((w[i] << 16) >= thr_hi ? c.scale : 0.f, w[i] >= thr_hi ? c.scale : 0.f);
}

inline tnr_dropout drop_or_none(const tnr_dropout* d) {
  tnr_dropout z; z.seed = nullptr; z.site = 0; z.p = 0.f;
  return d ? *d : z;
}

}  // namespace tnr
