// fp32 "head" GEMM-shaped work on the warp-level tensor path (mma.sync m16n8k8, TF32 inputs
// rounded with cvt.rna, fp32 accumulate): the additive-attention user encoders (student + M
// teachers in ONE launch), the student user-encoder backward, and the two small batched fp32
// GEMMs of the per-teacher projection (transform_matrix forward / weight gradient).
// These are a few MFLOP..GFLOP per step on [B*H, 256] fp32 matrices: far below a tcgen05 tile
// economy, but they must not serialise on 32 blocks either -- every kernel here spreads over
// >= 148 CTAs or finishes in a few microseconds.
//
// Reference: Tiny-NewsRec/model_bert.py:155-176 (UserEncoder, NAML branches), :15-34
// (AttentionPooling), :277-278,283 (transform_matrix), and their autograd backward.
#include "common.cuh"

namespace tnr {

__device__ __forceinline__ uint32_t f2tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* src, bool valid) {
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ----------------------------------------------------------------------------------
// batched fp32 GEMMs, 64 x 64 tiles, 4 warps (2 x 2), K chunks of 32 double-buffered with cp.async
//   NT: C[b][M,N]   = A[b][M,K] . B[b][N,K]^T + bias[b][N]
//   TN: C[b][N1,N2] += A[b][R,N1]^T . B[b][R,N2];  cbias[b][N1] += colsum(A[b])  (split over R, fp32 atomics)
// ----------------------------------------------------------------------------------
constexpr int SG_T = 64, SG_K = 32, SG_THREADS = 128;

__global__ void __launch_bounds__(SG_THREADS)
sgemm_nt_kernel(const float* __restrict__ A, const float* __restrict__ Bm, const float* __restrict__ bias,
                float* __restrict__ C, int M, int N, int K, long long sA, long long sB, long long sbias, long long sC) {
  __shared__ __align__(16) float As[2][SG_T][SG_K + 4];
  __shared__ __align__(16) float Bs[2][SG_T][SG_K + 4];
  A += blockIdx.z * sA; Bm += blockIdx.z * sB; C += blockIdx.z * sC;
  if (bias) bias += blockIdx.z * sbias;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wm = warp >> 1, wn = warp & 1;
  const int m0 = blockIdx.y * SG_T, n0 = blockIdx.x * SG_T;
  float acc[2][4][4] = {};
  const int nk = (K + SG_K - 1) / SG_K;
  auto load = [&](int st, int kc) {
    const int k0 = kc * SG_K;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * SG_THREADS, r = idx >> 3, c4 = (idx & 7) * 4;
      const bool kv = k0 + c4 + 4 <= K;
      const bool va = kv && (m0 + r < M), vb = kv && (n0 + r < N);
      cp_async16_zfill(&As[st][r][c4], va ? A + (size_t)(m0 + r) * K + k0 + c4 : A, va);
      cp_async16_zfill(&Bs[st][r][c4], vb ? Bm + (size_t)(n0 + r) * K + k0 + c4 : Bm, vb);
    }
    cp_async_commit();
  };
  load(0, 0);
  for (int kc = 0; kc < nk; ++kc) {
    const int st = kc & 1;
    if (kc + 1 < nk) { load(st ^ 1, kc + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < SG_K / 8; ++ks) {
      uint32_t b[4][2];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        b[nt][0] = f2tf32(Bs[st][wn * 32 + nt * 8 + g][ks * 8 + t]);
        b[nt][1] = f2tf32(Bs[st][wn * 32 + nt * 8 + g][ks * 8 + t + 4]);
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        uint32_t a[4];
        const int r = wm * 32 + mt * 16 + g;
        a[0] = f2tf32(As[st][r][ks * 8 + t]);
        a[1] = f2tf32(As[st][r + 8][ks * 8 + t]);
        a[2] = f2tf32(As[st][r][ks * 8 + t + 4]);
        a[3] = f2tf32(As[st][r + 8][ks * 8 + t + 4]);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_tf32(acc[mt][nt], a, b[nt][0], b[nt][1]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int hi = 0; hi < 2; ++hi) {
        const int m = m0 + wm * 32 + mt * 16 + g + hi * 8;
        if (m >= M) continue;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int n = n0 + wn * 32 + nt * 8 + 2 * t + e;
          if (n < N) C[(size_t)m * N + n] = acc[mt][nt][hi * 2 + e] + (bias ? bias[n] : 0.f);
        }
      }
}

__global__ void __launch_bounds__(SG_THREADS)
sgemm_tn_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ C, float* __restrict__ cbias,
                int R, int N1, int N2, int splits, int rows_per_split, long long sA, long long sB, long long sC,
                long long sbias) {
  __shared__ __align__(16) float As[2][SG_K][SG_T + 8];
  __shared__ __align__(16) float Bs[2][SG_K][SG_T + 8];
  const int batch = blockIdx.z / splits, split = blockIdx.z - batch * splits;
  A += batch * sA; Bm += batch * sB; C += batch * sC;
  if (cbias) cbias += batch * sbias;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wm = warp >> 1, wn = warp & 1;
  const int i0 = blockIdx.y * SG_T, j0 = blockIdx.x * SG_T;
  const int r_begin = split * rows_per_split;
  const int r_end = min(R, r_begin + rows_per_split);
  float acc[2][4][4] = {};
  float bsum = 0.f;
  const int nk = (r_end - r_begin + SG_K - 1) / SG_K;
  auto load = [&](int st, int kc) {
    const int r0 = r_begin + kc * SG_K;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * SG_THREADS, r = idx >> 4, c4 = (idx & 15) * 4;
      const bool rv = r0 + r < r_end;
      const bool va = rv && (i0 + c4 + 4 <= N1), vb = rv && (j0 + c4 + 4 <= N2);
      cp_async16_zfill(&As[st][r][c4], va ? A + (size_t)(r0 + r) * N1 + i0 + c4 : A, va);
      cp_async16_zfill(&Bs[st][r][c4], vb ? Bm + (size_t)(r0 + r) * N2 + j0 + c4 : Bm, vb);
    }
    cp_async_commit();
  };
  if (nk > 0) load(0, 0);
  for (int kc = 0; kc < nk; ++kc) {
    const int st = kc & 1;
    if (kc + 1 < nk) { load(st ^ 1, kc + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    if (cbias != nullptr && blockIdx.x == 0 && tid < SG_T) {
#pragma unroll 8
      for (int r = 0; r < SG_K; ++r) bsum += As[st][r][tid];
    }
#pragma unroll
    for (int ks = 0; ks < SG_K / 8; ++ks) {
      uint32_t b[4][2];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        b[nt][0] = f2tf32(Bs[st][ks * 8 + t][wn * 32 + nt * 8 + g]);
        b[nt][1] = f2tf32(Bs[st][ks * 8 + t + 4][wn * 32 + nt * 8 + g]);
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        uint32_t a[4];
        const int m = wm * 32 + mt * 16 + g;
        a[0] = f2tf32(As[st][ks * 8 + t][m]);
        a[1] = f2tf32(As[st][ks * 8 + t][m + 8]);
        a[2] = f2tf32(As[st][ks * 8 + t + 4][m]);
        a[3] = f2tf32(As[st][ks * 8 + t + 4][m + 8]);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_tf32(acc[mt][nt], a, b[nt][0], b[nt][1]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int hi = 0; hi < 2; ++hi) {
        const int m = i0 + wm * 32 + mt * 16 + g + hi * 8;
        if (m >= N1) continue;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int n = j0 + wn * 32 + nt * 8 + 2 * t + e;
          if (n < N2) atomicAdd(C + (size_t)m * N2 + n, acc[mt][nt][hi * 2 + e]);
        }
      }
  if (cbias != nullptr && blockIdx.x == 0 && tid < SG_T && i0 + tid < N1) atomicAdd(cbias + i0 + tid, bsum);
}

// ----------------------------------------------------------------------------------
// user encoders, forward.  grid (B, n_enc): block (b, i) handles impression b of encoder i.
//   blend (use_mask == 0): v = vec*m + pad_doc*(1-m); alpha unmasked
//   mask  (use_mask == 1): v = vec;  alpha *= m
//   e = tanh(v W1^T + b1) [H, Q] (TF32 mma, W1 fragments straight from L2); alpha = exp(e.w2 + b2)
//   a = alpha / (sum + 1e-8); user = sum_h a_h v_h   (fp32)
// ----------------------------------------------------------------------------------
constexpr int UE_THREADS = 256;
constexpr int UE_HMAX = 64;
constexpr int UE_MAX_ENC = 9;

struct UeFwdParams {
  tnr_user_encoder_io enc[UE_MAX_ENC];
  const float* mask;
  int use_mask, H, D, Q;
};

__global__ void __launch_bounds__(UE_THREADS)
user_encoder_fwd_kernel(const __grid_constant__ UeFwdParams p) {
  extern __shared__ __align__(16) float sm[];
  const int H = p.H, D = p.D, Q = p.Q, DS = D + 4;
  float* sv = sm;                              // [64][D+4]
  float* slog = sv + UE_HMAX * DS;             // [64]
  float* sz = slog + UE_HMAX;                  // [64]
  __shared__ float s_inv;
  const tnr_user_encoder_io io = p.enc[blockIdx.y];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  for (int i = tid; i < UE_HMAX * (D / 4); i += UE_THREADS) {
    const int h = i / (D / 4), d = (i - h * (D / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (h < H) {
      v = *reinterpret_cast<const float4*>(io.vecs + ((size_t)b * H + h) * D + d);
      if (!p.use_mask) {
        const float m = p.mask[(size_t)b * H + h];
        const float4 pd = *reinterpret_cast<const float4*>(io.pad_doc + d);
        v.x = v.x * m + pd.x * (1.0f - m); v.y = v.y * m + pd.y * (1.0f - m);
        v.z = v.z * m + pd.z * (1.0f - m); v.w = v.w * m + pd.w * (1.0f - m);
      }
    }
    *reinterpret_cast<float4*>(sv + h * DS + d) = v;
  }
  if (tid < UE_HMAX) slog[tid] = 0.f;
  __syncthreads();
  const int MT = (H + 15) >> 4;
  float acc[4][4][4];
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[mt][i][c] = 0.f;
  // warp w owns n-tiles w, w+8, w+16, w+24 (Q <= 256)
  const float* wrow[4];
  bool wv[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = (warp + 8 * i) * 8 + g;
    wv[i] = q < Q;
    wrow[i] = io.W1 + (size_t)(wv[i] ? q : 0) * D + t;
  }
  float bn[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) { bn[i][0] = wv[i] ? __ldg(wrow[i]) : 0.f; bn[i][1] = wv[i] ? __ldg(wrow[i] + 4) : 0.f; }
  for (int ks = 0; ks < D / 8; ++ks) {
    uint32_t bf[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) { bf[i][0] = f2tf32(bn[i][0]); bf[i][1] = f2tf32(bn[i][1]); }
    if (ks + 1 < D / 8) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        bn[i][0] = wv[i] ? __ldg(wrow[i] + (ks + 1) * 8) : 0.f;
        bn[i][1] = wv[i] ? __ldg(wrow[i] + (ks + 1) * 8 + 4) : 0.f;
      }
    }
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      if (mt < MT) {
        uint32_t a[4];
        const float* r0 = sv + (mt * 16 + g) * DS + ks * 8 + t;
        a[0] = f2tf32(r0[0]); a[1] = f2tf32(r0[8 * DS]); a[2] = f2tf32(r0[4]); a[3] = f2tf32(r0[8 * DS + 4]);
#pragma unroll
        for (int i = 0; i < 4; ++i) mma_tf32(acc[mt][i], a, bf[i][0], bf[i][1]);
      }
    }
  }
  // epilogue: e = tanh(acc + b1); partial logits
  float lp[4][2];
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) { lp[mt][0] = 0.f; lp[mt][1] = 0.f; }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int q = (warp + 8 * i) * 8 + 2 * t + e;
      if (q >= Q) continue;
      const float bq = io.b1[q], wq = io.w2[q];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        if (mt >= MT) continue;
#pragma unroll
        for (int hi = 0; hi < 2; ++hi) {
          const int h = mt * 16 + g + hi * 8;
          const float ev = tanhf(acc[mt][i][hi * 2 + e] + bq);
          lp[mt][hi] = fmaf(ev, wq, lp[mt][hi]);
          if (io.e_out != nullptr && h < H) io.e_out[((size_t)b * H + h) * Q + q] = ev;
        }
      }
    }
  }
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
      float v = lp[mt][hi];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      if (t == 0 && mt < MT) atomicAdd(&slog[mt * 16 + g + hi * 8], v);
    }
  __syncthreads();
  if (tid < UE_HMAX) {
    float al = 0.f;
    if (tid < H) {
      al = __expf(slog[tid] + io.b2[0]);
      if (p.use_mask) al *= p.mask[(size_t)b * H + tid];
    }
    sz[tid] = al;
  }
  __syncthreads();
  if (warp == 0) {
    float s = sz[lane] + sz[lane + 32];
    s = warp_sum(s);
    if (lane == 0) s_inv = 1.0f / (s + 1e-8f);
  }
  __syncthreads();
  const float inv = s_inv;
  if (tid < H) io.a_out[(size_t)b * H + tid] = sz[tid] * inv;
  for (int d = tid; d < D; d += UE_THREADS) {
    float u = 0.f;
    for (int h = 0; h < H; ++h) u = fmaf(sz[h] * inv, sv[h * DS + d], u);
    io.user[(size_t)b * D + d] = u;
  }
}

// ----------------------------------------------------------------------------------
// user encoder backward (student).  d_user [B, D] -> d_vecs (+=) [B*H, D]; dpad / db1 / dw2 / db2
// via fp32 atomics; dU (grad at the fc1 pre-activation) and the blended inputs are written to
// scratch so dW1 += dU^T V runs as one TN GEMM over all impressions afterwards.
// ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(UE_THREADS)
user_encoder_bwd_kernel(const float* __restrict__ vecs, const float* __restrict__ mask, const float* __restrict__ pad_doc,
                        const float* __restrict__ W1, const float* __restrict__ w2, int use_mask,
                        const float* __restrict__ a_in, const float* __restrict__ e_in, const float* __restrict__ d_user,
                        float* __restrict__ d_vecs, float* __restrict__ dpad, float* __restrict__ db1,
                        float* __restrict__ dw2, float* __restrict__ db2, float* __restrict__ dU, float* __restrict__ Vb,
                        int H, int D, int Q) {
  extern __shared__ __align__(16) float sm[];
  const int DS = D + 4, Qp = (Q + 7) & ~7, QS = Qp + 4;
  float* sv = sm;                         // [64][D+4]   blended inputs
  float* sdu = sv + UE_HMAX * DS;         // [64][Qp+4]  grad at fc1 pre-activation (rows >= H, cols >= Q zero)
  float* sa = sdu + UE_HMAX * QS;         // [64]
  float* sdz = sa + UE_HMAX;              // [64]
  float* smk = sdz + UE_HMAX;             // [64]
  float* sdusr = smk + UE_HMAX;           // [D]
  __shared__ float s_dot;
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  for (int i = tid; i < UE_HMAX * (D / 4); i += UE_THREADS) {
    const int h = i / (D / 4), d = (i - h * (D / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (h < H) {
      v = *reinterpret_cast<const float4*>(vecs + ((size_t)b * H + h) * D + d);
      if (!use_mask) {
        const float m = mask[(size_t)b * H + h];
        const float4 pd = *reinterpret_cast<const float4*>(pad_doc + d);
        v.x = v.x * m + pd.x * (1.0f - m); v.y = v.y * m + pd.y * (1.0f - m);
        v.z = v.z * m + pd.z * (1.0f - m); v.w = v.w * m + pd.w * (1.0f - m);
      }
      *reinterpret_cast<float4*>(Vb + ((size_t)b * H + h) * D + d) = v;
    }
    *reinterpret_cast<float4*>(sv + h * DS + d) = v;
  }
  for (int i = tid; i < UE_HMAX * QS; i += UE_THREADS) sdu[i] = 0.f;
  if (tid < UE_HMAX) {
    sa[tid] = tid < H ? a_in[(size_t)b * H + tid] : 0.f;
    smk[tid] = tid < H ? mask[(size_t)b * H + tid] : 0.f;
  }
  for (int d = tid; d < D; d += UE_THREADS) sdusr[d] = d_user[(size_t)b * D + d];
  __syncthreads();
  for (int h = warp; h < H; h += UE_THREADS / 32) {       // da_h = d_user . v_h
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s = fmaf(sdusr[d], sv[h * DS + d], s);
    s = warp_sum(s);
    if (lane == 0) sdz[h] = s;
  }
  __syncthreads();
  if (warp == 0) {
    float s = 0.f;
    for (int h = lane; h < H; h += 32) s += sa[h] * sdz[h];
    s = warp_sum(s);
    if (lane == 0) s_dot = s;
  }
  __syncthreads();
  const float dot = s_dot;
  if (tid < H) sdz[tid] = sa[tid] * (sdz[tid] - dot);
  __syncthreads();
  for (int q = tid; q < Q; q += UE_THREADS) {               // du, dw2, db1
    const float w = w2[q];
    float gw2 = 0.f, gb1 = 0.f;
    for (int h = 0; h < H; ++h) {
      const float ev = e_in[((size_t)b * H + h) * Q + q];
      const float dz = sdz[h];
      gw2 = fmaf(dz, ev, gw2);
      const float du = dz * w * (1.0f - ev * ev);
      sdu[h * QS + q] = du;
      dU[((size_t)b * H + h) * Q + q] = du;
      gb1 += du;
    }
    atomicAdd(dw2 + q, gw2);
    atomicAdd(db1 + q, gb1);
  }
  if (warp == 0) {
    float s = 0.f;
    for (int h = lane; h < H; h += 32) s += sdz[h];
    s = warp_sum(s);
    if (lane == 0) atomicAdd(db2, s);
  }
  __syncthreads();
  // dv[h][d] = a_h dusr_d + sum_q du[h][q] W1[q][d]   (M = 64 rows, N = D, K = Qp), 4 n-tiles per warp per pass
  const int MT = (H + 15) >> 4;
  for (int pass = 0; pass * 32 < D / 8; ++pass) {
    float acc[4][4][4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[mt][i][c] = 0.f;
    int ncol[4];
    bool nv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { ncol[i] = (pass * 32 + warp + 8 * i) * 8; nv[i] = ncol[i] < D; }
    for (int ks = 0; ks < Qp / 8; ++ks) {
      uint32_t bf[4][2];
      const int q0 = ks * 8 + t;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        bf[i][0] = f2tf32((nv[i] && q0 < Q) ? __ldg(W1 + (size_t)q0 * D + ncol[i] + g) : 0.f);
        bf[i][1] = f2tf32((nv[i] && q0 + 4 < Q) ? __ldg(W1 + (size_t)(q0 + 4) * D + ncol[i] + g) : 0.f);
      }
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        if (mt < MT) {
          uint32_t a[4];
          const float* r0 = sdu + (mt * 16 + g) * QS + ks * 8 + t;
          a[0] = f2tf32(r0[0]); a[1] = f2tf32(r0[8 * QS]); a[2] = f2tf32(r0[4]); a[3] = f2tf32(r0[8 * QS + 4]);
#pragma unroll
          for (int i = 0; i < 4; ++i) mma_tf32(acc[mt][i], a, bf[i][0], bf[i][1]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (!nv[i]) continue;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int d = ncol[i] + 2 * t + e;
        float gpad = 0.f;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
          if (mt >= MT) continue;
#pragma unroll
          for (int hi = 0; hi < 2; ++hi) {
            const int h = mt * 16 + g + hi * 8;
            if (h >= H) continue;
            const float tv = fmaf(sa[h], sdusr[d], acc[mt][i][hi * 2 + e]);
            float* dst = d_vecs + ((size_t)b * H + h) * D + d;
            if (use_mask) {
              *dst += tv;
            } else {
              const float m = smk[h];
              *dst += tv * m;
              gpad = fmaf(tv, 1.0f - m, gpad);
            }
          }
        }
        if (!use_mask) {
          gpad += __shfl_xor_sync(0xffffffffu, gpad, 4);
          gpad += __shfl_xor_sync(0xffffffffu, gpad, 8);
          gpad += __shfl_xor_sync(0xffffffffu, gpad, 16);
          if (g == 0) atomicAdd(dpad + d, gpad);
        }
      }
    }
  }
}

}  // namespace tnr

using namespace tnr;

#define TNR_API extern "C" __attribute__((visibility("default")))

static int launch_sgemm_tn(const float* A, const float* Bm, float* C, float* cbias, int R, int N1, int N2, int batch,
                           long long sA, long long sB, long long sC, long long sbias, cudaStream_t st) {
  TNR_REQUIRE(N1 % 4 == 0 && N2 % 4 == 0, "tnr_sgemm_tn_acc: N1=%d and N2=%d must be multiples of 4", N1, N2);
  TNR_REQUIRE(((uintptr_t)A % 16 == 0) && ((uintptr_t)Bm % 16 == 0) && sA % 4 == 0 && sB % 4 == 0,
              "tnr_sgemm_tn_acc: operands must be 16-byte aligned");
  if (N1 == 0 || N2 == 0 || batch == 0 || R == 0) return 0;
  const int tiles = ((N1 + SG_T - 1) / SG_T) * ((N2 + SG_T - 1) / SG_T) * batch;
  int splits = (2 * num_sms() + tiles - 1) / tiles;
  const int max_splits = (R + 4 * SG_K - 1) / (4 * SG_K);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int rows_per_split = (R + splits - 1) / splits;
  rows_per_split = (rows_per_split + SG_K - 1) / SG_K * SG_K;
  splits = (R + rows_per_split - 1) / rows_per_split;
  dim3 grid((N2 + SG_T - 1) / SG_T, (N1 + SG_T - 1) / SG_T, batch * splits);
  sgemm_tn_kernel<<<grid, SG_THREADS, 0, st>>>(A, Bm, C, cbias, R, N1, N2, splits, rows_per_split, sA, sB, sC, sbias);
  TNR_LAUNCH_CHECK();
  return 0;
}

TNR_API int tnr_sgemm_nt(const float* A, const float* Bm, const float* bias, float* C, int M, int N, int K, int batch,
                         long long sA, long long sB, long long sbias, long long sC, void* stream) {
  TNR_REQUIRE(K % 4 == 0, "tnr_sgemm_nt: K=%d must be a multiple of 4", K);
  TNR_REQUIRE(((uintptr_t)A % 16 == 0) && ((uintptr_t)Bm % 16 == 0) && sA % 4 == 0 && sB % 4 == 0,
              "tnr_sgemm_nt: operands must be 16-byte aligned");
  if (M == 0 || N == 0 || batch == 0) return 0;
  dim3 grid((N + SG_T - 1) / SG_T, (M + SG_T - 1) / SG_T, batch);
  sgemm_nt_kernel<<<grid, SG_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(A, Bm, bias, C, M, N, K, sA, sB, sbias, sC);
  TNR_LAUNCH_CHECK();
  return 0;
}

TNR_API int tnr_sgemm_tn_acc(const float* A, const float* Bm, float* C, float* cbias, int R, int N1, int N2, int batch,
                             long long sA, long long sB, long long sC, long long sbias, void* stream) {
  return launch_sgemm_tn(A, Bm, C, cbias, R, N1, N2, batch, sA, sB, sC, sbias, reinterpret_cast<cudaStream_t>(stream));
}

static int ue_check(const char* who, int H, int D, int Q) {
  TNR_REQUIRE(H >= 1 && H <= UE_HMAX, "%s: history length %d not supported (1..%d)", who, H, UE_HMAX);
  TNR_REQUIRE(D % 8 == 0 && D >= 8 && D <= 768, "%s: D=%d must be a multiple of 8 in 8..768", who, D);
  TNR_REQUIRE(Q >= 1 && Q <= 256, "%s: query dim %d not supported (1..256)", who, Q);
  return 0;
}

TNR_API int tnr_user_encoder_fwd_multi(const tnr_user_encoder_io* enc, int n_enc, const float* mask, int use_mask, int B,
                                       int H, int D, int Q, void* stream) {
  if (ue_check("tnr_user_encoder_fwd", H, D, Q)) return 1;
  TNR_REQUIRE(n_enc >= 1 && n_enc <= UE_MAX_ENC, "tnr_user_encoder_fwd_multi: n_enc=%d out of range (1..%d)", n_enc, UE_MAX_ENC);
  if (B == 0) return 0;
  UeFwdParams p;
  for (int i = 0; i < n_enc; ++i) p.enc[i] = enc[i];
  for (int i = n_enc; i < UE_MAX_ENC; ++i) p.enc[i] = enc[0];
  p.mask = mask; p.use_mask = use_mask; p.H = H; p.D = D; p.Q = Q;
  const int smem = (UE_HMAX * (D + 4) + 2 * UE_HMAX) * 4;
  static int attr_smem = 0;
  if (smem > attr_smem) {
    TNR_CHECK_CUDA(cudaFuncSetAttribute(user_encoder_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  user_encoder_fwd_kernel<<<dim3(B, n_enc), UE_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  TNR_LAUNCH_CHECK();
  return 0;
}

TNR_API int tnr_user_encoder_fwd(const float* vecs, const float* mask, const float* pad_doc, const float* W1,
                                 const float* b1, const float* w2, const float* b2, int use_mask, float* user,
                                 float* a_out, float* e_out, int B, int H, int D, int Q, void* stream) {
  tnr_user_encoder_io io;
  io.vecs = vecs; io.pad_doc = pad_doc; io.W1 = W1; io.b1 = b1; io.w2 = w2; io.b2 = b2;
  io.user = user; io.a_out = a_out; io.e_out = e_out;
  return tnr_user_encoder_fwd_multi(&io, 1, mask, use_mask, B, H, D, Q, stream);
}

TNR_API int tnr_user_encoder_bwd(const float* vecs, const float* mask, const float* pad_doc, const float* W1,
                                 const float* w2, int use_mask, const float* a_in, const float* e_in,
                                 const float* d_user, float* d_vecs, float* dpad, float* dW1, float* db1, float* dw2,
                                 float* db2, float* scratch, int B, int H, int D, int Q, void* stream) {
  if (ue_check("tnr_user_encoder_bwd", H, D, Q)) return 1;
  TNR_REQUIRE(Q % 4 == 0, "tnr_user_encoder_bwd: query dim %d must be a multiple of 4", Q);
  TNR_REQUIRE(D <= 512, "tnr_user_encoder_bwd: D=%d too large for shared memory (<= 512)", D);
  TNR_REQUIRE(scratch != nullptr && (uintptr_t)scratch % 16 == 0, "tnr_user_encoder_bwd: scratch [B*H*(Q+D)] fp32 required");
  if (B == 0) return 0;
  const int Qp = (Q + 7) & ~7;
  const int smem = (UE_HMAX * (D + 4) + UE_HMAX * (Qp + 4) + 3 * UE_HMAX + D) * 4;
  static int attr_smem = 0;
  if (smem > attr_smem) {
    TNR_CHECK_CUDA(cudaFuncSetAttribute(user_encoder_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* dU = scratch;
  float* Vb = scratch + (size_t)B * H * Q;
  user_encoder_bwd_kernel<<<B, UE_THREADS, smem, st>>>(vecs, mask, pad_doc, W1, w2, use_mask, a_in, e_in, d_user, d_vecs,
                                                      dpad, db1, dw2, db2, dU, Vb, H, D, Q);
  TNR_LAUNCH_CHECK();
  return launch_sgemm_tn(dU, Vb, dW1, nullptr, B * H, Q, D, 1, 0, 0, 0, 0, st);
}
