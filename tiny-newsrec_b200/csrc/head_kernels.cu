// fp32 "head" of the model: additive-attention user encoder (forward / backward), click
// scoring + cross-entropy + multi-teacher KD loss with its gradients, and the small fp32
// GEMMs of the per-teacher projection (transform_matrix) forward / weight gradient.
// All of this is latency / HBM bound (a few MB per step) -> warp-shuffle reductions,
// shared-memory staging, one block per impression.
//
// Reference: Tiny-NewsRec/model_bert.py:155-176 (UserEncoder, NAML branches), :15-34
// (AttentionPooling), :204 (score), :208-219 (kd_ce_loss), :222-244 (hid_mse_loss),
// :262-306 (Model.forward).
#include "common.cuh"

namespace tnr {

constexpr int UE_THREADS = 256;

__device__ __forceinline__ float block_sum_256(float v, float* red /*[8]*/) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < UE_THREADS / 32; ++i) t += red[i];
  return t;
}

// ----------------------------------------------------------------------------------
// user encoder forward.  vecs: row (b, h) at vecs + (b*H + h) * D.
//   blend (user_log_mask == 0): v = vec*m + pad_doc*(1-m); alpha unmasked
//   mask  (user_log_mask == 1): v = vec;  alpha *= m
// outputs user [B, D]; a [B, H] (normalised weights); e [B, H, Q] (tanh activations)
// smem: v [H][D] + e [H][Qp] + z[H]
// ----------------------------------------------------------------------------------
template <int HMAX>
__global__ void __launch_bounds__(UE_THREADS)
user_encoder_fwd_kernel(const float* __restrict__ vecs, const float* __restrict__ mask, const float* __restrict__ pad_doc,
                        const float* __restrict__ W1, const float* __restrict__ b1, const float* __restrict__ w2,
                        const float* __restrict__ b2, int use_mask, float* __restrict__ user, float* __restrict__ a_out,
                        float* __restrict__ e_out, int H, int D, int Q) {
  extern __shared__ __align__(16) float sm[];
  float* sv = sm;                       // [H][D]
  float* se = sv + (size_t)H * D;       // [H][Q]
  float* sz = se + (size_t)H * Q;       // [H]
  __shared__ float s_inv;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < H * D; i += UE_THREADS) {
    const int h = i / D, d = i - h * D;
    float v = vecs[((size_t)b * H + h) * D + d];
    if (!use_mask) {
      const float m = mask[(size_t)b * H + h];
      v = v * m + pad_doc[d] * (1.0f - m);
    }
    sv[i] = v;
  }
  __syncthreads();
  // e[h][q] = tanh(b1[q] + v[h] . W1[q])  -- thread q keeps acc[h] in registers
  for (int q = tid; q < Q; q += UE_THREADS) {
    float acc[HMAX];
#pragma unroll
    for (int h = 0; h < HMAX; ++h) acc[h] = 0.f;
    const float* wrow = W1 + (size_t)q * D;
    for (int d = 0; d < D; d += 4) {
      const float4 w = *reinterpret_cast<const float4*>(wrow + d);
#pragma unroll
      for (int h = 0; h < HMAX; ++h) {
        if (h < H) {
          const float4 x = *reinterpret_cast<const float4*>(sv + (size_t)h * D + d);
          acc[h] = fmaf(x.x, w.x, fmaf(x.y, w.y, fmaf(x.z, w.z, fmaf(x.w, w.w, acc[h]))));
        }
      }
    }
    const float bq = b1[q];
#pragma unroll
    for (int h = 0; h < HMAX; ++h)
      if (h < H) se[(size_t)h * Q + q] = tanhf(acc[h] + bq);
  }
  __syncthreads();
  for (int h = warp; h < H; h += UE_THREADS / 32) {
    float t = 0.f;
    for (int q = lane; q < Q; q += 32) t = fmaf(se[(size_t)h * Q + q], w2[q], t);
    t = warp_sum(t);
    if (lane == 0) {
      float al = __expf(t + b2[0]);
      if (use_mask) al *= mask[(size_t)b * H + h];
      sz[h] = al;
    }
  }
  __syncthreads();
  if (warp == 0) {
    float t = 0.f;
    for (int h = lane; h < H; h += 32) t += sz[h];
    t = warp_sum(t);
    if (lane == 0) s_inv = 1.0f / (t + 1e-8f);
  }
  __syncthreads();
  const float inv = s_inv;
  for (int h = tid; h < H; h += UE_THREADS) a_out[(size_t)b * H + h] = sz[h] * inv;
  for (int d = tid; d < D; d += UE_THREADS) {
    float acc = 0.f;
    for (int h = 0; h < H; ++h) acc = fmaf(sz[h] * inv, sv[(size_t)h * D + d], acc);
    user[(size_t)b * D + d] = acc;
  }
  if (e_out != nullptr)
    for (int i = tid; i < H * Q; i += UE_THREADS) e_out[(size_t)b * H * Q + i] = se[i];
}

// ----------------------------------------------------------------------------------
// user encoder backward (student).  d_user [B, D] -> d_vecs (+=) [B*H, D] and parameter
// gradients (fp32 atomics into dpad [D], dW1 [Q, D], db1 [Q], dw2 [Q], db2 [1]).
// ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(UE_THREADS)
user_encoder_bwd_kernel(const float* __restrict__ vecs, const float* __restrict__ mask, const float* __restrict__ pad_doc,
                        const float* __restrict__ W1, const float* __restrict__ w2, int use_mask,
                        const float* __restrict__ a_in, const float* __restrict__ e_in, const float* __restrict__ d_user,
                        float* __restrict__ d_vecs, float* __restrict__ dpad, float* __restrict__ dW1,
                        float* __restrict__ db1, float* __restrict__ dw2, float* __restrict__ db2, int H, int D, int Q) {
  extern __shared__ __align__(16) float sm[];
  float* sv = sm;                        // [H][D] blended inputs
  float* sdu = sv + (size_t)H * D;       // [H][Q]  grad at fc1 pre-activation
  float* sa = sdu + (size_t)H * Q;       // [H]
  float* sdz = sa + H;                   // [H]
  float* sdusr = sdz + H;                // [D]
  __shared__ float s_dot;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < H * D; i += UE_THREADS) {
    const int h = i / D, d = i - h * D;
    float v = vecs[((size_t)b * H + h) * D + d];
    if (!use_mask) {
      const float m = mask[(size_t)b * H + h];
      v = v * m + pad_doc[d] * (1.0f - m);
    }
    sv[i] = v;
  }
  for (int h = tid; h < H; h += UE_THREADS) sa[h] = a_in[(size_t)b * H + h];
  for (int d = tid; d < D; d += UE_THREADS) sdusr[d] = d_user[(size_t)b * D + d];
  __syncthreads();
  for (int h = warp; h < H; h += UE_THREADS / 32) {       // da_h = d_user . v_h
    float t = 0.f;
    for (int d = lane; d < D; d += 32) t = fmaf(sdusr[d], sv[(size_t)h * D + d], t);
    t = warp_sum(t);
    if (lane == 0) sdz[h] = t;
  }
  __syncthreads();
  if (warp == 0) {
    float t = 0.f;
    for (int h = lane; h < H; h += 32) t += sa[h] * sdz[h];
    t = warp_sum(t);
    if (lane == 0) s_dot = t;
  }
  __syncthreads();
  const float dot = s_dot;
  for (int h = tid; h < H; h += UE_THREADS) sdz[h] = sa[h] * (sdz[h] - dot);
  __syncthreads();
  // du[h][q] = dz_h * w2_q * (1 - e^2); dw2_q, db1_q
  for (int q = tid; q < Q; q += UE_THREADS) {
    const float w = w2[q];
    float gw2 = 0.f, gb1 = 0.f;
    for (int h = 0; h < H; ++h) {
      const float ev = e_in[((size_t)b * H + h) * Q + q];
      const float dz = sdz[h];
      gw2 = fmaf(dz, ev, gw2);
      const float du = dz * w * (1.0f - ev * ev);
      sdu[(size_t)h * Q + q] = du;
      gb1 += du;
    }
    atomicAdd(dw2 + q, gw2);
    atomicAdd(db1 + q, gb1);
  }
  if (warp == 0) {
    float t = 0.f;
    for (int h = lane; h < H; h += 32) t += sdz[h];
    t = warp_sum(t);
    if (lane == 0) atomicAdd(db2, t);
  }
  __syncthreads();
  // thread d: dW1[q][d] += sum_h du[h][q] v[h][d];   dv[h][d] = a_h dusr_d + sum_q du[h][q] W1[q][d]
  for (int d = tid; d < D; d += UE_THREADS) {
    for (int q = 0; q < Q; ++q) {
      float t = 0.f;
      for (int h = 0; h < H; ++h) t = fmaf(sdu[(size_t)h * Q + q], sv[(size_t)h * D + d], t);
      atomicAdd(dW1 + (size_t)q * D + d, t);
    }
    float gpad = 0.f;
    for (int h = 0; h < H; ++h) {
      float t = sa[h] * sdusr[d];
      for (int q = 0; q < Q; ++q) t = fmaf(sdu[(size_t)h * Q + q], W1[(size_t)q * D + d], t);
      float* dst = d_vecs + ((size_t)b * H + h) * D + d;
      if (use_mask) {
        *dst += t;
      } else {
        const float m = mask[(size_t)b * H + h];
        *dst += t * m;
        gpad = fmaf(t, 1.0f - m, gpad);
      }
    }
    if (!use_mask) atomicAdd(dpad + d, gpad);
  }
}

// ----------------------------------------------------------------------------------
// click scoring + CE + multi-teacher KD loss, forward and gradients in one pass.
// Row layout of every [R, D] news matrix (student news vecs, T_ext, TP_ext, G_ext, d_news):
//   rows [0, B*H)           history block   (b, h) -> b*H + h
//   rows [B*H, B*(H+K))     candidate block (b, k) -> B*H + b*K + k
//   rows [B*(H+K), +B)      user rows (teacher matrices only)
// T_ext / TP_ext / G_ext are [M, R_ext, D] with R_ext = B*(H+K) + B.
// losses[0..2] += {distill, emb, target} (already divided by B); M may be 0 (PLM-NR: CE only).
// ----------------------------------------------------------------------------------
constexpr int KD_MAXM = 8;
constexpr int KD_MAXK = 32;

__global__ void __launch_bounds__(UE_THREADS)
kd_loss_kernel(const float* __restrict__ s_news, const float* __restrict__ s_user, const int64_t* __restrict__ label,
               const float* __restrict__ T_ext, const float* __restrict__ TP_ext, int M, int B, int H, int K, int D,
               float temperature, float coef, int want_grad, float* __restrict__ score_out, float* __restrict__ losses,
               float* __restrict__ d_news, float* __restrict__ d_user, float* __restrict__ G_ext) {
  __shared__ float s_sc[KD_MAXK];                 // student scores
  __shared__ float s_tsc[KD_MAXM][KD_MAXK];       // teacher scores
  __shared__ float s_w[KD_MAXM];                  // teacher weights
  __shared__ float s_ds[KD_MAXK];                 // d loss / d student score
  __shared__ float s_mse[KD_MAXM];                // NE_i + UE_i
  __shared__ float red[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int R = B * (H + K);
  const size_t Rext = (size_t)R + B;
  const float* cand = s_news + ((size_t)B * H + (size_t)b * K) * D;
  const float* hist = s_news + (size_t)b * H * D;
  const float* usr = s_user + (size_t)b * D;
  const int lab = (int)label[b];
  const float invB = 1.0f / (float)B;

  // scores: (1 + M) * K dot products of length D, one warp each
  for (int item = warp; item < (1 + M) * K; item += UE_THREADS / 32) {
    const int i = item / K, k = item - i * K;
    const float *c, *u;
    if (i == 0) { c = cand + (size_t)k * D; u = usr; }
    else {
      const float* Ti = T_ext + (size_t)(i - 1) * Rext * D;
      c = Ti + ((size_t)B * H + (size_t)b * K + k) * D;
      u = Ti + ((size_t)R + b) * D;
    }
    float t = 0.f;
    for (int d = lane; d < D; d += 32) t = fmaf(c[d], u[d], t);
    t = warp_sum(t);
    if (lane == 0) { if (i == 0) s_sc[k] = t; else s_tsc[i - 1][k] = t; }
  }
  // embedding MSEs per teacher: NE_i (mean over (H+K)*D) + UE_i (mean over D)
  for (int i = 0; i < M; ++i) {
    const float* TPi = TP_ext + (size_t)i * Rext * D;
    float acc = 0.f;
    for (int idx = tid; idx < (H + K) * D; idx += UE_THREADS) {
      const int r = idx / D, d = idx - r * D;
      const size_t row = r < H ? (size_t)b * H + r : (size_t)B * H + (size_t)b * K + (r - H);
      const float diff = s_news[row * D + d] - TPi[row * D + d];
      acc = fmaf(diff, diff, acc);
    }
    float ne = block_sum_256(acc, red) / (float)((H + K) * D);
    float acc2 = 0.f;
    for (int d = tid; d < D; d += UE_THREADS) {
      const float diff = usr[d] - TPi[((size_t)R + b) * D + d];
      acc2 = fmaf(diff, diff, acc2);
    }
    float ue = block_sum_256(acc2, red) / (float)D;
    if (tid == 0) s_mse[i] = ne + ue;
  }
  __syncthreads();
  if (tid == 0) {
    // student CE
    float mx = -INFINITY;
    for (int k = 0; k < K; ++k) mx = fmaxf(mx, s_sc[k]);
    float se = 0.f;
    for (int k = 0; k < K; ++k) se += expf(s_sc[k] - mx);
    const float lse = mx + logf(se);
    const float target = lse - s_sc[lab];
    float distill = 0.f, emb = 0.f;
    float pT[KD_MAXK];
    for (int k = 0; k < K; ++k) pT[k] = 0.f;
    if (M > 0) {
      float tl[KD_MAXM];
      float mn = INFINITY;
      for (int i = 0; i < M; ++i) {
        float m2 = -INFINITY;
        for (int k = 0; k < K; ++k) m2 = fmaxf(m2, s_tsc[i][k]);
        float s2 = 0.f;
        for (int k = 0; k < K; ++k) s2 += expf(s_tsc[i][k] - m2);
        tl[i] = m2 + logf(s2) - s_tsc[i][lab];
        mn = fminf(mn, tl[i]);
      }
      float ws = 0.f;
      for (int i = 0; i < M; ++i) { s_w[i] = expf(-(tl[i] - mn)); ws += s_w[i]; }
      for (int i = 0; i < M; ++i) { s_w[i] /= ws; emb += s_w[i] * s_mse[i]; }
      // weighted teacher score, soft-label CE at temperature tau
      float tmx = -INFINITY, smx = -INFINITY;
      float tsc[KD_MAXK];
      for (int k = 0; k < K; ++k) {
        float t = 0.f;
        for (int i = 0; i < M; ++i) t = fmaf(s_tsc[i][k], s_w[i], t);
        tsc[k] = t / temperature;
        tmx = fmaxf(tmx, tsc[k]);
        smx = fmaxf(smx, s_sc[k] / temperature);
      }
      float tsum = 0.f, ssum = 0.f;
      for (int k = 0; k < K; ++k) { pT[k] = expf(tsc[k] - tmx); tsum += pT[k]; ssum += expf(s_sc[k] / temperature - smx); }
      const float slse = smx + logf(ssum);
      for (int k = 0; k < K; ++k) {
        pT[k] /= tsum;
        distill -= pT[k] * (s_sc[k] / temperature - slse);
        // d distill / d s_k = (softmax(s/tau)_k - pT_k) / tau
        s_ds[k] = (expf(s_sc[k] / temperature - slse) - pT[k]) / temperature;
      }
    } else {
      for (int k = 0; k < K; ++k) s_ds[k] = 0.f;
    }
    for (int k = 0; k < K; ++k)
      s_ds[k] = (s_ds[k] + coef * (expf(s_sc[k] - lse) - (k == lab ? 1.0f : 0.0f))) * invB;
    atomicAdd(losses + 0, distill * invB);
    atomicAdd(losses + 1, emb * invB);
    atomicAdd(losses + 2, target * invB);
    atomicAdd(losses + 3, (distill + coef * target + emb) * invB);     // model_bert.py:305
  }
  __syncthreads();
  for (int k = tid; k < K; k += UE_THREADS) score_out[(size_t)b * K + k] = s_sc[k];
  if (!want_grad) return;
  const float cne = 2.0f / (float)((H + K) * D) * invB;
  const float cue = 2.0f / (float)D * invB;
  // d_news rows of this impression (assigned, not accumulated) and per-teacher residual grads
  for (int idx = tid; idx < (H + K) * D; idx += UE_THREADS) {
    const int r = idx / D, d = idx - r * D;
    const size_t row = r < H ? (size_t)b * H + r : (size_t)B * H + (size_t)b * K + (r - H);
    const float sv = s_news[row * D + d];
    float g = 0.f;
    for (int i = 0; i < M; ++i) {
      const float gi = s_w[i] * cne * (sv - TP_ext[((size_t)i * Rext + row) * D + d]);
      g += gi;
      G_ext[((size_t)i * Rext + row) * D + d] = -gi;
    }
    if (r >= H) g = fmaf(s_ds[r - H], usr[d], g);
    d_news[row * D + d] = g;
  }
  for (int d = tid; d < D; d += UE_THREADS) {
    float g = 0.f;
    for (int k = 0; k < K; ++k) g = fmaf(s_ds[k], cand[(size_t)k * D + d], g);
    for (int i = 0; i < M; ++i) {
      const float gi = s_w[i] * cue * (usr[d] - TP_ext[((size_t)i * Rext + R + b) * D + d]);
      g += gi;
      G_ext[((size_t)i * Rext + R + b) * D + d] = -gi;
    }
    d_user[(size_t)b * D + d] = g;
  }
}

// ----------------------------------------------------------------------------------
// small fp32 GEMMs (SIMT, 64x64 tiles, 4x4 per thread), batched over blockIdx.z
//   NT: C[M,N] = A[M,K] . B[N,K]^T + bias[N]
//   TN: C[N1,N2] += A[R,N1]^T . B[R,N2];  cbias[N1] += colsum(A)
// ----------------------------------------------------------------------------------
constexpr int SG_T = 64, SG_K = 16;

__global__ void __launch_bounds__(256)
sgemm_nt_kernel(const float* __restrict__ A, const float* __restrict__ Bm, const float* __restrict__ bias,
                float* __restrict__ C, int M, int N, int K, long long sA, long long sB, long long sbias, long long sC) {
  __shared__ float As[SG_K][SG_T + 4];
  __shared__ float Bs[SG_K][SG_T + 4];
  A += blockIdx.z * sA; Bm += blockIdx.z * sB; C += blockIdx.z * sC;
  if (bias) bias += blockIdx.z * sbias;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * SG_T, n0 = blockIdx.x * SG_T;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += SG_K) {
    for (int i = tid; i < SG_T * SG_K; i += 256) {
      const int r = i / SG_K, c = i - r * SG_K;
      As[c][r] = (m0 + r < M && k0 + c < K) ? A[(size_t)(m0 + r) * K + k0 + c] : 0.f;
      Bs[c][r] = (n0 + r < N && k0 + c < K) ? Bm[(size_t)(n0 + r) * K + k0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SG_K; ++k) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; bb[i] = Bs[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) C[(size_t)m * N + n] = acc[i][j] + (bias ? bias[n] : 0.f);
    }
  }
}

__global__ void __launch_bounds__(256)
sgemm_tn_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ C, float* __restrict__ cbias,
                int R, int N1, int N2, long long sA, long long sB, long long sC, long long sbias) {
  __shared__ float As[SG_K][SG_T + 4];
  __shared__ float Bs[SG_K][SG_T + 4];
  A += blockIdx.z * sA; Bm += blockIdx.z * sB; C += blockIdx.z * sC;
  if (cbias) cbias += blockIdx.z * sbias;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int i0 = blockIdx.y * SG_T, j0 = blockIdx.x * SG_T;
  float acc[4][4] = {};
  float bsum[4] = {};
  for (int r0 = 0; r0 < R; r0 += SG_K) {
    for (int i = tid; i < SG_T * SG_K; i += 256) {
      const int r = i / SG_T, c = i - r * SG_T;
      As[r][c] = (r0 + r < R && i0 + c < N1) ? A[(size_t)(r0 + r) * N1 + i0 + c] : 0.f;
      Bs[r][c] = (r0 + r < R && j0 + c < N2) ? Bm[(size_t)(r0 + r) * N2 + j0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SG_K; ++k) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; bb[i] = Bs[k][tx * 4 + i]; bsum[i] += a[i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = i0 + ty * 4 + i;
    if (m >= N1) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = j0 + tx * 4 + j;
      if (n < N2) C[(size_t)m * N2 + n] += acc[i][j];
    }
    if (cbias != nullptr && blockIdx.x == 0 && tx == 0) cbias[m] += bsum[i];
  }
}

}  // namespace tnr

using namespace tnr;

#define TNR_API extern "C" __attribute__((visibility("default")))

static int ue_smem_fwd(int H, int D, int Q) { return (H * D + H * Q + H) * 4; }
static int ue_smem_bwd(int H, int D, int Q) { return (H * D + H * Q + 2 * H + D) * 4; }

TNR_API int tnr_user_encoder_fwd(const float* vecs, const float* mask, const float* pad_doc, const float* W1,
                                 const float* b1, const float* w2, const float* b2, int use_mask, float* user,
                                 float* a_out, float* e_out, int B, int H, int D, int Q, void* stream) {
  TNR_REQUIRE(H >= 1 && H <= 64, "tnr_user_encoder_fwd: history length %d not supported (1..64)", H);
  TNR_REQUIRE(D % 4 == 0, "tnr_user_encoder_fwd: D must be a multiple of 4");
  if (B == 0) return 0;
  const int smem = ue_smem_fwd(H, D, Q);
  TNR_REQUIRE(smem <= 200 * 1024, "tnr_user_encoder_fwd: H*D too large for shared memory");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (H <= 16) {
    TNR_CHECK_CUDA(cudaFuncSetAttribute(user_encoder_fwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    user_encoder_fwd_kernel<16><<<B, UE_THREADS, smem, st>>>(vecs, mask, pad_doc, W1, b1, w2, b2, use_mask, user, a_out, e_out, H, D, Q);
  } else {
    TNR_CHECK_CUDA(cudaFuncSetAttribute(user_encoder_fwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    user_encoder_fwd_kernel<64><<<B, UE_THREADS, smem, st>>>(vecs, mask, pad_doc, W1, b1, w2, b2, use_mask, user, a_out, e_out, H, D, Q);
  }
  TNR_LAUNCH_CHECK();
  return 0;
}

TNR_API int tnr_user_encoder_bwd(const float* vecs, const float* mask, const float* pad_doc, const float* W1,
                                 const float* w2, int use_mask, const float* a_in, const float* e_in,
                                 const float* d_user, float* d_vecs, float* dpad, float* dW1, float* db1, float* dw2,
                                 float* db2, int B, int H, int D, int Q, void* stream) {
  if (B == 0) return 0;
  const int smem = ue_smem_bwd(H, D, Q);
  TNR_REQUIRE(smem <= 200 * 1024, "tnr_user_encoder_bwd: H*D too large for shared memory");
  TNR_CHECK_CUDA(cudaFuncSetAttribute(user_encoder_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  user_encoder_bwd_kernel<<<B, UE_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      vecs, mask, pad_doc, W1, w2, use_mask, a_in, e_in, d_user, d_vecs, dpad, dW1, db1, dw2, db2, H, D, Q);
  TNR_LAUNCH_CHECK();
  return 0;
}

TNR_API int tnr_kd_loss_fwdbwd(const float* s_news, const float* s_user, const int64_t* label, const float* T_ext,
                               const float* TP_ext, int M, int B, int H, int K, int D, float temperature, float coef,
                               int want_grad, float* score_out, float* losses, float* d_news, float* d_user,
                               float* G_ext, void* stream) {
  TNR_REQUIRE(M >= 0 && M <= KD_MAXM, "tnr_kd_loss_fwdbwd: num_teachers %d out of range (0..%d)", M, KD_MAXM);
  TNR_REQUIRE(K >= 1 && K <= KD_MAXK, "tnr_kd_loss_fwdbwd: candidates per impression %d out of range (1..%d)", K, KD_MAXK);
  if (B == 0) return 0;
  kd_loss_kernel<<<B, UE_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      s_news, s_user, label, T_ext, TP_ext, M, B, H, K, D, temperature, coef, want_grad, score_out, losses, d_news, d_user,
      G_ext);
  TNR_LAUNCH_CHECK();
  return 0;
}

TNR_API int tnr_sgemm_nt(const float* A, const float* Bm, const float* bias, float* C, int M, int N, int K, int batch,
                         long long sA, long long sB, long long sbias, long long sC, void* stream) {
  if (M == 0 || N == 0 || batch == 0) return 0;
  dim3 grid((N + SG_T - 1) / SG_T, (M + SG_T - 1) / SG_T, batch);
  sgemm_nt_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(A, Bm, bias, C, M, N, K, sA, sB, sbias, sC);
  TNR_LAUNCH_CHECK();
  return 0;
}

TNR_API int tnr_sgemm_tn_acc(const float* A, const float* Bm, float* C, float* cbias, int R, int N1, int N2, int batch,
                             long long sA, long long sB, long long sC, long long sbias, void* stream) {
  if (N1 == 0 || N2 == 0 || batch == 0) return 0;
  dim3 grid((N2 + SG_T - 1) / SG_T, (N1 + SG_T - 1) / SG_T, batch);
  sgemm_tn_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(A, Bm, C, cbias, R, N1, N2, sA, sB, sC, sbias);
  TNR_LAUNCH_CHECK();
  return 0;
}
