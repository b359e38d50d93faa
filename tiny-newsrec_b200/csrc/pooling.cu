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
// Block per news, ONE pass: warp w takes rows w, w + 8, ... two at a time -- the e row (score) and the x
// row (value) of both are requested before either is used -- and accumulates alpha_s * x_s unnormalised in
// registers; the 8 partial rows and the alpha sum are combined through shared memory at the end.
// (The first version pooled with C/8 = 96 of 256 threads walking the S rows serially: one 16-byte load in
// flight per thread, 36 % of HBM peak.)
template <int VPL>
__global__ void __launch_bounds__(POOL_THREADS)
attnpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ e, int ldq, int Q,
                    const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ mask,
                    __nv_bfloat16* __restrict__ out, float* __restrict__ a_out, int S) {
  constexpr int C = VPL * 256;
  constexpr int NW = POOL_THREADS / 32;
  __shared__ float s_a[POOL_SMAX];
  __shared__ __align__(16) float s_red[NW][C];
  __shared__ float s_part[NW];
  const int n = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float bias2 = b2[0];
  const int q0 = lane * 8;
  float wv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) wv[i] = q0 + i < Q ? w2[q0 + i] : 0.f;      // Q % 8 == 0: all-or-nothing per lane
  float acc[VPL * 8];
#pragma unroll
  for (int i = 0; i < VPL * 8; ++i) acc[i] = 0.f;
  float asum = 0.f;
  const __nv_bfloat16* xn = x + (size_t)n * S * C;
  const __nv_bfloat16* en = e + (size_t)n * S * ldq;
  for (int s0 = warp; s0 < S; s0 += 2 * NW) {
    const int s1 = s0 + NW;
    const bool two = s1 < S;
    bf16x8 er[2], xr[2][VPL];
    er[0] = er[1] = bf16x8{{0u, 0u, 0u, 0u}};
    if (q0 < Q) {
      er[0] = *reinterpret_cast<const bf16x8*>(en + (size_t)s0 * ldq + q0);
      if (two) er[1] = *reinterpret_cast<const bf16x8*>(en + (size_t)s1 * ldq + q0);
    }
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      xr[0][v] = *reinterpret_cast<const bf16x8*>(xn + (size_t)s0 * C + (v * 32 + lane) * 8);
      if (two) xr[1][v] = *reinterpret_cast<const bf16x8*>(xn + (size_t)s1 * C + (v * 32 + lane) * 8);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int s = k ? s1 : s0;
      if (k && !two) break;
      float f[8];
      unpack8(er[k], f);
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) d = fmaf(f[i], wv[i], d);
      for (int q = q0 + 256; q < Q; q += 256) {                          // Q > 256 (not the reference's 200)
        unpack8(*reinterpret_cast<const bf16x8*>(en + (size_t)s * ldq + q), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) d = fmaf(f[i], w2[q + i], d);
      }
      d = warp_sum(d);
      float al = __expf(d + bias2);
      if (mask != nullptr) al *= mask[(size_t)n * S + s];
      if (lane == 0) s_a[s] = al;
      asum += al;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        unpack8(xr[k][v], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[v * 8 + i] = fmaf(al, f[i], acc[v * 8 + i]);
      }
    }
  }
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    float4* dst = reinterpret_cast<float4*>(&s_red[warp][(v * 32 + lane) * 8]);
    dst[0] = make_float4(acc[v * 8 + 0], acc[v * 8 + 1], acc[v * 8 + 2], acc[v * 8 + 3]);
    dst[1] = make_float4(acc[v * 8 + 4], acc[v * 8 + 5], acc[v * 8 + 6], acc[v * 8 + 7]);
  }
  if (lane == 0) s_part[warp] = asum;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < NW; ++w) tot += s_part[w];
  const float inv = 1.0f / (tot + 1e-8f);
  for (int s = threadIdx.x; s < S; s += POOL_THREADS) a_out[(size_t)n * S + s] = s_a[s] * inv;
  for (int c = threadIdx.x * 4; c < C; c += POOL_THREADS * 4) {
    float4 t = *reinterpret_cast<const float4*>(&s_red[0][c]);
#pragma unroll
    for (int w = 1; w < NW; ++w) {
      const float4 u = *reinterpret_cast<const float4*>(&s_red[w][c]);
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    uint2 o;
    o.x = pack_bf16(t.x * inv, t.y * inv);
    o.y = pack_bf16(t.z * inv, t.w * inv);
    *reinterpret_cast<uint2*>(out + (size_t)n * C + c) = o;
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
  TNR_REQUIRE(C % 256 == 0 && C <= 1024 && Q % 8 == 0 && ldq % 8 == 0,
              "tnr_attnpool_fwd: C must be 256/512/768/1024 and Q, ldq multiples of 8 (C=%d Q=%d ldq=%d)", C, Q, ldq);
  if (n == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define POOL_FWD(VPL_)                                                                                         \
  attnpool_fwd_kernel<VPL_><<<n, POOL_THREADS, 0, st>>>(                                                      \
      reinterpret_cast<const __nv_bfloat16*>(x_bf16), reinterpret_cast<const __nv_bfloat16*>(e_bf16), ldq, Q, w2, b2, mask, \
      reinterpret_cast<__nv_bfloat16*>(out_bf16), a_out, S)
  switch (C / 256) {
    case 1: POOL_FWD(1); break;
    case 2: POOL_FWD(2); break;
    case 3: POOL_FWD(3); break;
    default: POOL_FWD(4); break;
  }
#undef POOL_FWD
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
