// Additive-attention pooling over words (news encoder head), forward and backward.
// Reference: Tiny-NewsRec/model_bert.py:15-34 (AttentionPooling), used UNMASKED on word
// vectors at model_bert.py:133.  The fc1 + tanh contraction runs on the tensor cores
// (tnr_gemm_bf16 with the TANH epilogue); these kernels do the HBM-bound rest in one pass
// over the last hidden state:  alpha = exp(e.w2 + b2) [* mask];  a = alpha / (sum + 1e-8);
// out = sum_s a_s x_s.
#include "common.cuh"

namespace tnr {

constexpr int POOL_THREADS = 256;
constexpr int POOL_SMAX = 512;

// x [n, S, C] bf16; e [n*S, ldq] bf16 (first Q columns valid); out bf16 [n, C]; a_out fp32 [n, S]
__global__ void __launch_bounds__(POOL_THREADS)
attnpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ e, int ldq, int Q,
                    const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ mask,
                    __nv_bfloat16* __restrict__ out, float* __restrict__ a_out, int S, int C) {
  __shared__ float s_a[POOL_SMAX];
  __shared__ float s_sum;
  const int n = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float bias2 = b2[0];
  for (int s = warp; s < S; s += POOL_THREADS / 32) {
    const __nv_bfloat16* er = e + ((size_t)n * S + s) * ldq;
    float acc = 0.f;
    for (int q = lane * 8; q < Q; q += 256) {        // Q % 8 == 0
      float f[8];
      unpack8(*reinterpret_cast<const bf16x8*>(er + q), f);
      const float4 wa = *reinterpret_cast<const float4*>(w2 + q), wb = *reinterpret_cast<const float4*>(w2 + q + 4);
      acc += f[0] * wa.x + f[1] * wa.y + f[2] * wa.z + f[3] * wa.w + f[4] * wb.x + f[5] * wb.y + f[6] * wb.z + f[7] * wb.w;
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      float al = __expf(acc + bias2);
      if (mask != nullptr) al *= mask[(size_t)n * S + s];
      s_a[s] = al;
    }
  }
  __syncthreads();
  if (warp == 0) {
    float t = 0.f;
    for (int s = lane; s < S; s += 32) t += s_a[s];
    t = warp_sum(t);
    if (lane == 0) s_sum = 1.0f / (t + 1e-8f);
  }
  __syncthreads();
  const float inv = s_sum;
  for (int s = threadIdx.x; s < S; s += POOL_THREADS) {
    const float a = s_a[s] * inv;
    a_out[(size_t)n * S + s] = a;
  }
  for (int col = threadIdx.x * 8; col < C; col += POOL_THREADS * 8) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    const __nv_bfloat16* xp = x + (size_t)n * S * C + col;
    for (int s = 0; s < S; ++s) {
      float f[8];
      unpack8(*reinterpret_cast<const bf16x8*>(xp + (size_t)s * C), f);
      const float a = s_a[s] * inv;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(a, f[i], acc[i]);
    }
    *reinterpret_cast<bf16x8*>(out + (size_t)n * C + col) = pack8(acc);
  }
}

// backward:  dout fp32 [n, C]  ->  dx_direct bf16 [n,S,C] (= a_s * dout), du bf16 [n*S, ldq]
// (gradient at the fc1 pre-activation), dw2 [Q] / db2 [1] (fp32 atomics).
__global__ void __launch_bounds__(POOL_THREADS)
attnpool_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ e, int ldq, int Q,
                    const float* __restrict__ w2, const float* __restrict__ a_in, const float* __restrict__ dout,
                    __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ du, float* __restrict__ dw2,
                    float* __restrict__ db2, int S, int C) {
  __shared__ float s_a[POOL_SMAX];
  __shared__ float s_dz[POOL_SMAX];
  __shared__ float s_dot;
  const int n = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int s = threadIdx.x; s < S; s += POOL_THREADS) s_a[s] = a_in[(size_t)n * S + s];
  __syncthreads();
  // da_s = dout . x_s ; dx_direct = a_s * dout
  for (int s = warp; s < S; s += POOL_THREADS / 32) {
    const float a = s_a[s];
    float acc = 0.f;
    for (int col = lane * 8; col < C; col += 256) {
      float f[8], o[8];
      unpack8(*reinterpret_cast<const bf16x8*>(x + ((size_t)n * S + s) * C + col), f);
      const float4 d0 = *reinterpret_cast<const float4*>(dout + (size_t)n * C + col);
      const float4 d1 = *reinterpret_cast<const float4*>(dout + (size_t)n * C + col + 4);
      const float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) { acc = fmaf(d[i], f[i], acc); o[i] = a * d[i]; }
      *reinterpret_cast<bf16x8*>(dx + ((size_t)n * S + s) * C + col) = pack8(o);
    }
    acc = warp_sum(acc);
    if (lane == 0) s_dz[s] = acc;          // da_s for now
  }
  __syncthreads();
  if (warp == 0) {
    float t = 0.f;
    for (int s = lane; s < S; s += 32) t += s_a[s] * s_dz[s];
    t = warp_sum(t);
    if (lane == 0) s_dot = t;
  }
  __syncthreads();
  const float dot = s_dot;
  for (int s = threadIdx.x; s < S; s += POOL_THREADS) s_dz[s] = s_a[s] * (s_dz[s] - dot);   // dz_s
  __syncthreads();
  // du_sq = dz_s * w2_q * (1 - e_sq^2);  dw2_q += sum_s dz_s e_sq;  db2 += sum_s dz_s
  for (int q = threadIdx.x; q < Q; q += POOL_THREADS) {
    const float w = w2[q];
    float gw = 0.f;
    for (int s = 0; s < S; ++s) {
      const size_t off = ((size_t)n * S + s) * ldq + q;
      const float ev = bf16_to_f(e[off]);
      const float dz = s_dz[s];
      gw = fmaf(dz, ev, gw);
      du[off] = __float2bfloat16_rn(dz * w * (1.0f - ev * ev));
    }
    atomicAdd(dw2 + q, gw);
  }
  if (warp == 0) {
    float t = 0.f;
    for (int s = lane; s < S; s += 32) t += s_dz[s];
    t = warp_sum(t);
    if (lane == 0) atomicAdd(db2, t);
  }
}

}  // namespace tnr

using namespace tnr;

extern "C" __attribute__((visibility("default"))) int tnr_attnpool_fwd(const void* x_bf16, const void* e_bf16, int ldq, int Q, const float* w2, const float* b2,
                                const float* mask, void* out_bf16, float* a_out, int n, int S, int C, void* stream) {
  TNR_REQUIRE(S >= 1 && S <= POOL_SMAX, "tnr_attnpool_fwd: S=%d out of range (1..%d)", S, POOL_SMAX);
  TNR_REQUIRE(C % 8 == 0 && Q % 8 == 0 && ldq % 8 == 0, "tnr_attnpool_fwd: C, Q, ldq must be multiples of 8");
  if (n == 0) return 0;
  attnpool_fwd_kernel<<<n, POOL_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x_bf16), reinterpret_cast<const __nv_bfloat16*>(e_bf16), ldq, Q, w2, b2, mask,
      reinterpret_cast<__nv_bfloat16*>(out_bf16), a_out, S, C);
  TNR_LAUNCH_CHECK();
  return 0;
}

extern "C" __attribute__((visibility("default"))) int tnr_attnpool_bwd(const void* x_bf16, const void* e_bf16, int ldq, int Q, const float* w2, const float* a_in,
                                const float* dout, void* dx_bf16, void* du_bf16, float* dw2, float* db2, int n, int S,
                                int C, void* stream) {
  TNR_REQUIRE(S >= 1 && S <= POOL_SMAX, "tnr_attnpool_bwd: S=%d out of range (1..%d)", S, POOL_SMAX);
  TNR_REQUIRE(C % 8 == 0 && Q % 8 == 0 && ldq % 8 == 0, "tnr_attnpool_bwd: C, Q, ldq must be multiples of 8");
  if (n == 0) return 0;
  attnpool_bwd_kernel<<<n, POOL_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x_bf16), reinterpret_cast<const __nv_bfloat16*>(e_bf16), ldq, Q, w2, a_in,
      dout, reinterpret_cast<__nv_bfloat16*>(dx_bf16), reinterpret_cast<__nv_bfloat16*>(du_bf16), dw2, db2, S, C);
  TNR_LAUNCH_CHECK();
  return 0;
}
