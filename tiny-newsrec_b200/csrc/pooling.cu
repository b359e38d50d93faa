// Additive-attention pooling over words (news encoder head), forward and backward.
// Reference: Tiny-NewsRec/model_bert.py:15-34 (AttentionPooling), used UNMASKED on word
// vectors at model_bert.py:133.  The fc1 + tanh contraction runs on the tensor cores
// (tnr_gemm_bf16 with the TANH epilogue); these kernels do the HBM-bound rest in one pass
// over the last hidden state:  alpha = exp(e.w2 + b2) [* mask];  a = alpha / (sum + 1e-8);
// out = sum_s a_s x_s.
#include "common.cuh"

namespace tnr {

constexpr int POOL_SMAX = 512;
constexpr int PW_WARPS = 4;           // warps per block; ONE NEWS PER WARP
constexpr int PW_THREADS = PW_WARPS * 32;
constexpr int PW_RING = 6;            // rows in flight per warp (cp.async ring in shared memory)

__device__ __forceinline__ void pw_cp16(void* smem_dst, const void* src) {
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void pw_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void pw_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// x [n, S, C] bf16; e [n*S, ldq] bf16 (first Q columns valid); out bf16 [n, C]; a_out fp32 [n, S]
// v3: a warp owns a news item and streams its S rows (value row 2 C bytes + score row 2 Q bytes) through a
// private PW_RING-deep cp.async ring: 12 KB in flight per warp, ~12 warps per SM, no block-level barrier and no
// cross-warp combine (v2: block per news, 8 warps x 2 rows in flight, partial rows combined through shared
// memory behind two __syncthreads -- 50 % of HBM peak; the short-lived blocks spent as long starting and
// draining as streaming).  Every lane reads back exactly the 16-byte pieces it copied itself, so the ring
// needs no synchronisation beyond cp.async.wait_group.
template <int VPL>
__global__ void __launch_bounds__(PW_THREADS)
attnpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ e, int ldq, int Q,
                    const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ mask,
                    __nv_bfloat16* __restrict__ out, float* __restrict__ a_out, int n_news, int S) {
  constexpr int C = VPL * 256;
  constexpr int ROW16 = (VPL + 1) * 32;                 // 16-byte pieces per ring row: VPL x 32 of x, 32 of e
  extern __shared__ __align__(16) unsigned char pw_sm[];
  const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
  const int n = blockIdx.x * PW_WARPS + warp;
  if (n >= n_news) return;                               // no block-level barrier below
  bf16x8* ring = reinterpret_cast<bf16x8*>(pw_sm) + (size_t)warp * PW_RING * ROW16;
  float* s_a = reinterpret_cast<float*>(pw_sm + (size_t)PW_WARPS * PW_RING * ROW16 * 16) + warp * POOL_SMAX;
  const float bias2 = b2[0];
  const int q0 = lane * 8;
  const bool has_e = q0 < Q;                             // Q % 8 == 0: all-or-nothing per lane
  float wv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) wv[i] = has_e ? w2[q0 + i] : 0.f;
  const __nv_bfloat16* xn = x + (size_t)n * S * C;
  const __nv_bfloat16* en = e + (size_t)n * S * ldq;
  auto issue = [&](int s, int slot) {
    if (s < S) {
      bf16x8* dst = ring + slot * ROW16;
#pragma unroll
      for (int v = 0; v < VPL; ++v) pw_cp16(dst + v * 32 + lane, xn + (size_t)s * C + (v * 32 + lane) * 8);
      if (has_e) pw_cp16(dst + VPL * 32 + lane, en + (size_t)s * ldq + q0);
    }
    pw_commit();                                         // always: keeps the group count in step with s
  };
#pragma unroll
  for (int s = 0; s < PW_RING; ++s) issue(s, s);
  float acc[VPL * 8];
#pragma unroll
  for (int i = 0; i < VPL * 8; ++i) acc[i] = 0.f;
  float asum = 0.f;
  int slot = 0;
  for (int s = 0; s < S; ++s) {
    pw_wait<PW_RING - 1>();
    const bf16x8* row = ring + slot * ROW16;
    float f[8];
    float d = 0.f;
    if (has_e) {
      unpack8(row[VPL * 32 + lane], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) d = fmaf(f[i], wv[i], d);
    }
    for (int q = q0 + 256; q < Q; q += 256) {            // Q > 256 (not the reference's 200): straight from global
      unpack8(*reinterpret_cast<const bf16x8*>(en + (size_t)s * ldq + q), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) d = fmaf(f[i], w2[q + i], d);
    }
    bf16x8 xr[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) xr[v] = row[v * 32 + lane];
    issue(s + PW_RING, slot);                            // the slot's pieces are in registers: refill it
    slot = slot + 1 == PW_RING ? 0 : slot + 1;
    d = warp_sum(d);
    float al = __expf(d + bias2);
    if (mask != nullptr) al *= mask[(size_t)n * S + s];
    if (lane == 0) s_a[s] = al;
    asum += al;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      unpack8(xr[v], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[v * 8 + i] = fmaf(al, f[i], acc[v * 8 + i]);
    }
  }
  pw_wait<0>();
  const float inv = 1.0f / (asum + 1e-8f);
  __syncwarp();
  for (int s = lane; s < S; s += 32) a_out[(size_t)n * S + s] = s_a[s] * inv;
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = acc[v * 8 + i] * inv;
    *reinterpret_cast<bf16x8*>(out + (size_t)n * C + (v * 32 + lane) * 8) = pack8(o);
  }
}

// backward:  dout fp32 [n, C]  ->  dx_direct bf16 [n,S,C] (= a_s * dout), du bf16 [n*S, ldq]
// (gradient at the fc1 pre-activation), dw2 [Q] / db2 [1] (fp32 atomics, one per block and column).
// Same shape as the forward: a warp per news, pass A streams the x rows (da_s = dout . x_s, writes a_s dout),
// pass B streams the e rows (du, dw2); the dw2 / db2 partials of the block's warps meet in shared memory.
template <int VPL>
__global__ void __launch_bounds__(PW_THREADS)
attnpool_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ e, int ldq, int Q,
                    const float* __restrict__ w2, const float* __restrict__ a_in, const float* __restrict__ dout,
                    __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ du, float* __restrict__ dw2,
                    float* __restrict__ db2, int n_news, int S) {
  constexpr int C = VPL * 256;
  constexpr int ROW16 = VPL * 32;
  extern __shared__ __align__(16) unsigned char pw_sm[];
  __shared__ float s_gw[PW_WARPS][256];
  __shared__ float s_gb[PW_WARPS];
  const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
  const int n = blockIdx.x * PW_WARPS + warp;
  const bool active = n < n_news;
  bf16x8* ring = reinterpret_cast<bf16x8*>(pw_sm) + (size_t)warp * PW_RING * ROW16;
  float* s_a = reinterpret_cast<float*>(pw_sm + (size_t)PW_WARPS * PW_RING * ROW16 * 16) + warp * 2 * POOL_SMAX;
  float* s_dz = s_a + POOL_SMAX;
  const int q0 = lane * 8;
  const bool has_e = q0 < Q;
  float gw[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) gw[i] = 0.f;
  float gb = 0.f;
  if (active) {
    const __nv_bfloat16* xn = x + (size_t)n * S * C;
    const __nv_bfloat16* en = e + (size_t)n * S * ldq;
    __nv_bfloat16* dxn = dx + (size_t)n * S * C;
    __nv_bfloat16* dun = du + (size_t)n * S * ldq;
    for (int s = lane; s < S; s += 32) s_a[s] = a_in[(size_t)n * S + s];
    float dr[VPL * 8];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const float4 d0 = *reinterpret_cast<const float4*>(dout + (size_t)n * C + (v * 32 + lane) * 8);
      const float4 d1 = *reinterpret_cast<const float4*>(dout + (size_t)n * C + (v * 32 + lane) * 8 + 4);
      dr[v * 8 + 0] = d0.x; dr[v * 8 + 1] = d0.y; dr[v * 8 + 2] = d0.z; dr[v * 8 + 3] = d0.w;
      dr[v * 8 + 4] = d1.x; dr[v * 8 + 5] = d1.y; dr[v * 8 + 6] = d1.z; dr[v * 8 + 7] = d1.w;
    }
    __syncwarp();
    // ---- pass A: x rows
    auto issue_x = [&](int s, int slot) {
      if (s < S) {
#pragma unroll
        for (int v = 0; v < VPL; ++v) pw_cp16(ring + slot * ROW16 + v * 32 + lane, xn + (size_t)s * C + (v * 32 + lane) * 8);
      }
      pw_commit();
    };
#pragma unroll
    for (int s = 0; s < PW_RING; ++s) issue_x(s, s);
    int slot = 0;
    for (int s = 0; s < S; ++s) {
      pw_wait<PW_RING - 1>();
      bf16x8 xr[VPL];
#pragma unroll
      for (int v = 0; v < VPL; ++v) xr[v] = ring[slot * ROW16 + v * 32 + lane];
      issue_x(s + PW_RING, slot);
      slot = slot + 1 == PW_RING ? 0 : slot + 1;
      const float a = s_a[s];
      float da = 0.f;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        float f[8], o[8];
        unpack8(xr[v], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) { da = fmaf(dr[v * 8 + i], f[i], da); o[i] = a * dr[v * 8 + i]; }
        *reinterpret_cast<bf16x8*>(dxn + (size_t)s * C + (v * 32 + lane) * 8) = pack8(o);
      }
      da = warp_sum(da);
      if (lane == 0) s_dz[s] = da;                     // da_s for now
    }
    pw_wait<0>();
    __syncwarp();
    float dot = 0.f;
    for (int s = lane; s < S; s += 32) dot = fmaf(s_a[s], s_dz[s], dot);
    dot = warp_sum(dot);
    for (int s = lane; s < S; s += 32) {
      const float dz = s_a[s] * (s_dz[s] - dot);
      s_dz[s] = dz;
      gb += dz;
    }
    __syncwarp();
    // ---- pass B: e rows (one 16-byte piece per lane and row; the ring slots are reused)
    float wv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) wv[i] = has_e ? w2[q0 + i] : 0.f;
    auto issue_e = [&](int s, int sl) {
      if (s < S && has_e) pw_cp16(ring + sl * ROW16 + lane, en + (size_t)s * ldq + q0);
      pw_commit();
    };
#pragma unroll
    for (int s = 0; s < PW_RING; ++s) issue_e(s, s);
    slot = 0;
    for (int s = 0; s < S; ++s) {
      pw_wait<PW_RING - 1>();
      const float dz = s_dz[s];
      if (has_e) {
        float f[8], o[8];
        unpack8(ring[slot * ROW16 + lane], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          gw[i] = fmaf(dz, f[i], gw[i]);
          o[i] = dz * wv[i] * (1.0f - f[i] * f[i]);
        }
        *reinterpret_cast<bf16x8*>(dun + (size_t)s * ldq + q0) = pack8(o);
      }
      issue_e(s + PW_RING, slot);
      slot = slot + 1 == PW_RING ? 0 : slot + 1;
      for (int q = q0 + 256; q < Q; q += 256) {          // Q > 256: straight from global
        float f[8], o[8];
        unpack8(*reinterpret_cast<const bf16x8*>(en + (size_t)s * ldq + q), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          atomicAdd(dw2 + q + i, dz * f[i]);
          o[i] = dz * w2[q + i] * (1.0f - f[i] * f[i]);
        }
        *reinterpret_cast<bf16x8*>(dun + (size_t)s * ldq + q) = pack8(o);
      }
    }
    pw_wait<0>();
    gb = warp_sum(gb);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) s_gw[warp][q0 + i] = gw[i];
  if (lane == 0) s_gb[warp] = gb;
  __syncthreads();
  for (int q = threadIdx.x; q < min(Q, 256); q += PW_THREADS) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < PW_WARPS; ++w) t += s_gw[w][q];
    atomicAdd(dw2 + q, t);
  }
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < PW_WARPS; ++w) t += s_gb[w];
    atomicAdd(db2, t);
  }
}

}  // namespace tnr

using namespace tnr;

template <int VPL>
static int pool_fwd_launch(const void* x, const void* e, int ldq, int Q, const float* w2, const float* b2, const float* mask,
                           void* out, float* a_out, int n, int S, cudaStream_t st) {
  const int smem = PW_WARPS * (PW_RING * (VPL + 1) * 512 + POOL_SMAX * 4);
  TNR_SET_SMEM(attnpool_fwd_kernel<VPL>, smem);
  attnpool_fwd_kernel<VPL><<<(n + PW_WARPS - 1) / PW_WARPS, PW_THREADS, smem, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(e), ldq, Q, w2, b2, mask,
      reinterpret_cast<__nv_bfloat16*>(out), a_out, n, S);
  TNR_LAUNCH_CHECK();
  return 0;
}

template <int VPL>
static int pool_bwd_launch(const void* x, const void* e, int ldq, int Q, const float* w2, const float* a_in, const float* dout,
                           void* dx, void* du, float* dw2, float* db2, int n, int S, cudaStream_t st) {
  const int smem = PW_WARPS * (PW_RING * VPL * 512 + 2 * POOL_SMAX * 4);
  TNR_SET_SMEM(attnpool_bwd_kernel<VPL>, smem);
  attnpool_bwd_kernel<VPL><<<(n + PW_WARPS - 1) / PW_WARPS, PW_THREADS, smem, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(e), ldq, Q, w2, a_in, dout,
      reinterpret_cast<__nv_bfloat16*>(dx), reinterpret_cast<__nv_bfloat16*>(du), dw2, db2, n, S);
  TNR_LAUNCH_CHECK();
  return 0;
}

extern "C" __attribute__((visibility("default"))) int tnr_attnpool_fwd(const void* x_bf16, const void* e_bf16, int ldq, int Q, const float* w2, const float* b2,
                                const float* mask, void* out_bf16, float* a_out, int n, int S, int C, void* stream) {
  TNR_REQUIRE(S >= 1 && S <= POOL_SMAX, "tnr_attnpool_fwd: S=%d out of range (1..%d)", S, POOL_SMAX);
  TNR_REQUIRE(C % 256 == 0 && C <= 1024 && Q % 8 == 0 && ldq % 8 == 0,
              "tnr_attnpool_fwd: C must be 256/512/768/1024 and Q, ldq multiples of 8 (C=%d Q=%d ldq=%d)", C, Q, ldq);
  TNR_REQUIRE(((uintptr_t)x_bf16 | (uintptr_t)e_bf16 | (uintptr_t)out_bf16) % 16 == 0, "tnr_attnpool_fwd: 16-byte alignment required");
  if (n == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (C / 256) {
    case 1: return pool_fwd_launch<1>(x_bf16, e_bf16, ldq, Q, w2, b2, mask, out_bf16, a_out, n, S, st);
    case 2: return pool_fwd_launch<2>(x_bf16, e_bf16, ldq, Q, w2, b2, mask, out_bf16, a_out, n, S, st);
    case 3: return pool_fwd_launch<3>(x_bf16, e_bf16, ldq, Q, w2, b2, mask, out_bf16, a_out, n, S, st);
    default: return pool_fwd_launch<4>(x_bf16, e_bf16, ldq, Q, w2, b2, mask, out_bf16, a_out, n, S, st);
  }
}

extern "C" __attribute__((visibility("default"))) int tnr_attnpool_bwd(const void* x_bf16, const void* e_bf16, int ldq, int Q, const float* w2, const float* a_in,
                                const float* dout, void* dx_bf16, void* du_bf16, float* dw2, float* db2, int n, int S,
                                int C, void* stream) {
  TNR_REQUIRE(S >= 1 && S <= POOL_SMAX, "tnr_attnpool_bwd: S=%d out of range (1..%d)", S, POOL_SMAX);
  TNR_REQUIRE(C % 256 == 0 && C <= 1024 && Q % 8 == 0 && ldq % 8 == 0,
              "tnr_attnpool_bwd: C must be 256/512/768/1024 and Q, ldq multiples of 8 (C=%d Q=%d ldq=%d)", C, Q, ldq);
  TNR_REQUIRE(((uintptr_t)x_bf16 | (uintptr_t)e_bf16 | (uintptr_t)dx_bf16 | (uintptr_t)du_bf16 | (uintptr_t)dout) % 16 == 0,
              "tnr_attnpool_bwd: 16-byte alignment required");
  if (n == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (C / 256) {
    case 1: return pool_bwd_launch<1>(x_bf16, e_bf16, ldq, Q, w2, a_in, dout, dx_bf16, du_bf16, dw2, db2, n, S, st);
    case 2: return pool_bwd_launch<2>(x_bf16, e_bf16, ldq, Q, w2, a_in, dout, dx_bf16, du_bf16, dw2, db2, n, S, st);
    case 3: return pool_bwd_launch<3>(x_bf16, e_bf16, ldq, Q, w2, a_in, dout, dx_bf16, du_bf16, dw2, db2, n, S, st);
    default: return pool_bwd_launch<4>(x_bf16, e_bf16, ldq, Q, w2, a_in, dout, dx_bf16, du_bf16, dw2, db2, n, S, st);
  }
}
