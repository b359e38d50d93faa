// fp32 "head" of the model: click scoring + cross-entropy + multi-teacher KD loss with its
// gradients (the user encoders and the transform_matrix GEMMs live in head_mma.cu).
// Latency / HBM bound (a few MB per step) -> warp-shuffle reductions, one block per impression.
//
// Reference: Tiny-NewsRec/model_bert.py:155-176 (UserEncoder, NAML branches), :15-34
// (AttentionPooling), :204 (score), :208-219 (kd_ce_loss), :222-244 (hid_mse_loss),
// :262-306 (Model.forward).
#include "common.cuh"

namespace tnr {

// 1024 threads per impression: the block count is the batch size (32 at the demo shape, on 148 SMs), so the only
// parallelism to add is inside the block -- with 256 threads the launch took 90 us for 7.6 MB.
constexpr int UE_THREADS = 1024;

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
  __shared__ float s_part[KD_MAXM][UE_THREADS / 32];   // per-warp partial MSEs
  __shared__ int s_last;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int warp = uniform_warp_id(), lane = tid & 31;
  const int R = B * (H + K);
  const size_t Rext = (size_t)R + B;
  const float* cand = s_news + ((size_t)B * H + (size_t)b * K) * D;
  const float* usr = s_user + (size_t)b * D;
  // a label outside 0..K-1 (torch's cross_entropy device-asserts on it) must not index shared memory out of bounds:
  // it is read as 0 and the impression's target loss -- hence the batch losses -- become NaN
  const long long lab_raw = label[b];
  const bool lab_bad = lab_raw < 0 || lab_raw >= (long long)K;
  const int lab = lab_bad ? 0 : (int)lab_raw;
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
  // embedding MSEs per teacher: NE_i (mean over (H+K)*D) + UE_i (mean over D).  One sweep over the impression's rows
  // with all M teachers' loads in flight and ONE block reduction of the 2M partial sums (a sweep and two block
  // reductions per teacher cost 4M barriers and M dependent rounds of global loads)
  {
    float ne[KD_MAXM], ue[KD_MAXM];
#pragma unroll
    for (int i = 0; i < KD_MAXM; ++i) { ne[i] = 0.f; ue[i] = 0.f; }
    for (int idx = tid; idx < (H + K) * D; idx += UE_THREADS) {
      const int r = idx / D, d = idx - r * D;
      const size_t row = r < H ? (size_t)b * H + r : (size_t)B * H + (size_t)b * K + (r - H);
      const float sv = s_news[row * D + d];
#pragma unroll
      for (int i = 0; i < KD_MAXM; ++i)
        if (i < M) { const float diff = sv - TP_ext[((size_t)i * Rext + row) * D + d]; ne[i] = fmaf(diff, diff, ne[i]); }
    }
    for (int d = tid; d < D; d += UE_THREADS) {
      const float uv = usr[d];
#pragma unroll
      for (int i = 0; i < KD_MAXM; ++i)
        if (i < M) { const float diff = uv - TP_ext[((size_t)i * Rext + R + b) * D + d]; ue[i] = fmaf(diff, diff, ue[i]); }
    }
#pragma unroll
    for (int i = 0; i < KD_MAXM; ++i)
      if (i < M) {
        ne[i] = warp_sum(ne[i]);
        ue[i] = warp_sum(ue[i]);
        if (lane == 0) { s_part[i][warp] = ne[i] / (float)((H + K) * D) + ue[i] / (float)D; }
      }
    __syncthreads();
    if (tid < M) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < UE_THREADS / 32; ++w) t += s_part[tid][w];      // fixed order: deterministic
      s_mse[tid] = t;
    }
  }
  __syncthreads();
  if (tid == 0) {
    // student CE
    float mx = -INFINITY;
    for (int k = 0; k < K; ++k) mx = fmaxf(mx, s_sc[k]);
    float se = 0.f;
    for (int k = 0; k < K; ++k) se += expf(s_sc[k] - mx);
    const float lse = mx + logf(se);
    const float target = lab_bad ? __int_as_float(0x7fc00000) : lse - s_sc[lab];
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
    // deterministic batch means: per-impression terms go to losses[4 + 4b ..]; the block that takes the last
    // ticket adds them up in a fixed order (below) -- fp32 atomics would make the loss depend on block
    // scheduling in its last bits.
    float* part = losses + 4 + 4 * (size_t)b;
    part[0] = distill; part[1] = emb; part[2] = target; part[3] = distill + coef * target + emb;   // model_bert.py:305
    __threadfence();
    unsigned int* ticket = reinterpret_cast<unsigned int*>(losses + 4 + 4 * (size_t)B);
    s_last = (atomicAdd(ticket, 1u) == (unsigned int)B - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_last && warp == 0) {
    // lane i sums impressions i, i + 32, ... in order, then a fixed butterfly over the 32 lanes
    __threadfence();
    float acc4[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = lane; i < B; i += 32) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(losses + 4) + i);
      acc4[0] += v.x; acc4[1] += v.y; acc4[2] += v.z; acc4[3] += v.w;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) acc4[c] = warp_sum(acc4[c]);
    if (lane == 0) {
      for (int c = 0; c < 4; ++c) losses[c] = acc4[c] * invB;
      *reinterpret_cast<unsigned int*>(losses + 4 + 4 * (size_t)B) = 0u;     // ready for the next launch
    }
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

}  // namespace tnr

using namespace tnr;

#define TNR_API extern "C" __attribute__((visibility("default")))

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

