// Fused short-sequence self-attention with relative-position bias (L <= 32, head dim 64).
//   P = softmax(Q K^T / sqrt(dh) + (1 - mask) * -10000 + relpos[h]);  ctx = P V
// Reference: Tiny-NewsRec/tnlrv3/modeling.py:205-231 (multi_head_attention), mask :446-454,
// rel-pos bias :458-463 (batch-invariant [A, L, L] table).
// One warp per (news, head): K/V (and Q/dO in backward) tiles staged in shared memory as
// bf16, lane i owns query row i, scores / probabilities live in registers.  The backward
// recomputes P from Q,K (nothing but QKV is saved by the forward).
#include "common.cuh"

namespace tnr {

constexpr int DH = 64;
constexpr int LMAX = 32;
constexpr int ATT_WARPS = 4;

__device__ __forceinline__ void load_tile_bf16(__nv_bfloat16* s, const __nv_bfloat16* g, int L, int ld, int lane) {
  // L rows x 64 bf16 (128 B per row): 8 x 16 B chunks per row
  for (int idx = lane; idx < L * 8; idx += 32) {
    const int r = idx >> 3, c = idx & 7;
    *reinterpret_cast<bf16x8*>(s + r * DH + c * 8) = *reinterpret_cast<const bf16x8*>(g + (size_t)r * ld + c * 8);
  }
}

__device__ __forceinline__ float dot_row(const float* a, const __nv_bfloat16* srow) {
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < DH / 8; ++c) {
    float f[8];
    unpack8(*reinterpret_cast<const bf16x8*>(srow + c * 8), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc = fmaf(a[c * 8 + i], f[i], acc);
  }
  return acc;
}

__device__ __forceinline__ void axpy_row(float* acc, float a, const __nv_bfloat16* srow) {
#pragma unroll
  for (int c = 0; c < DH / 8; ++c) {
    float f[8];
    unpack8(*reinterpret_cast<const bf16x8*>(srow + c * 8), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[c * 8 + i] = fmaf(a, f[i], acc[c * 8 + i]);
  }
}

__device__ __forceinline__ void load_row_regs(float* r, const __nv_bfloat16* g) {
#pragma unroll
  for (int c = 0; c < DH / 8; ++c) unpack8(*reinterpret_cast<const bf16x8*>(g + c * 8), r + c * 8);
}

__device__ __forceinline__ void store_row_regs(__nv_bfloat16* g, const float* r) {
#pragma unroll
  for (int c = 0; c < DH / 8; ++c) *reinterpret_cast<bf16x8*>(g + c * 8) = pack8(r + c * 8);
}

// scores -> normalised probabilities of row `lane` (valid for lane < L); p[j] for j >= L is 0
__device__ __forceinline__ void softmax_row(float (&p)[LMAX], const float* q, const __nv_bfloat16* sK,
                                            const float* __restrict__ relrow, float my_mask_add, int L, int lane) {
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < LMAX; ++j) {
    const float madd = __shfl_sync(0xffffffffu, my_mask_add, j);
    if (j < L) {
      float s = dot_row(q, sK + j * DH) + madd;
      if (lane < L) s += relrow[j];
      p[j] = s;
      mx = fmaxf(mx, s);
    } else {
      p[j] = 0.f;
    }
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < LMAX; ++j)
    if (j < L) { p[j] = __expf(p[j] - mx); sum += p[j]; }
  const float inv = 1.0f / sum;
#pragma unroll
  for (int j = 0; j < LMAX; ++j) p[j] *= inv;
}

__global__ void __launch_bounds__(ATT_WARPS * 32)
attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, const int64_t* __restrict__ mask, int mask_ld,
                const float* __restrict__ relpos, __nv_bfloat16* __restrict__ ctx, int n_news, int L, int A, int E) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem) + (size_t)warp * 2 * LMAX * DH;
  __nv_bfloat16* sV = sK + LMAX * DH;
  const long long item = (long long)blockIdx.x * ATT_WARPS + warp;
  if (item >= (long long)n_news * A) return;
  const int n = (int)(item / A), h = (int)(item % A);
  const int ld = 3 * E;
  const __nv_bfloat16* base = qkv + (size_t)n * L * ld + h * DH;
  load_tile_bf16(sK, base + E, L, ld, lane);
  load_tile_bf16(sV, base + 2 * E, L, ld, lane);
  float q[DH];
  float my_mask_add = 0.f;
  if (lane < L) {
    load_row_regs(q, base + (size_t)lane * ld);
#pragma unroll
    for (int i = 0; i < DH; ++i) q[i] *= 0.125f;          // 1/sqrt(64), exact
    my_mask_add = (1.0f - (float)mask[(size_t)n * mask_ld + lane]) * -10000.0f;
  } else {
#pragma unroll
    for (int i = 0; i < DH; ++i) q[i] = 0.f;
  }
  __syncwarp();
  float p[LMAX];
  softmax_row(p, q, sK, relpos + ((size_t)h * L + (lane < L ? lane : 0)) * L, my_mask_add, L, lane);
  float acc[DH];
#pragma unroll
  for (int i = 0; i < DH; ++i) acc[i] = 0.f;
#pragma unroll
  for (int j = 0; j < LMAX; ++j)
    if (j < L) axpy_row(acc, p[j], sV + j * DH);
  if (lane < L) store_row_regs(ctx + ((size_t)n * L + lane) * E + h * DH, acc);
}

// backward: dQKV from dCtx (recompute P).  smem per warp: Q,K,V,dO tiles (bf16) + P, dS (fp32, padded)
constexpr int ATT_BWD_SMEM_PER_WARP = 4 * LMAX * DH * 2 + 2 * LMAX * (LMAX + 1) * 4;

__global__ void __launch_bounds__(ATT_WARPS * 32)
attn_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const int64_t* __restrict__ mask, int mask_ld,
                const float* __restrict__ relpos, const __nv_bfloat16* __restrict__ dctx,
                __nv_bfloat16* __restrict__ dqkv, int n_news, int L, int A, int E) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* wbase = smem + (size_t)warp * ATT_BWD_SMEM_PER_WARP;
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(wbase);
  __nv_bfloat16* sK = sQ + LMAX * DH;
  __nv_bfloat16* sV = sK + LMAX * DH;
  __nv_bfloat16* sO = sV + LMAX * DH;
  float* sP = reinterpret_cast<float*>(sO + LMAX * DH);
  float* sS = sP + LMAX * (LMAX + 1);
  const long long item = (long long)blockIdx.x * ATT_WARPS + warp;
  if (item >= (long long)n_news * A) return;
  const int n = (int)(item / A), h = (int)(item % A);
  const int ld = 3 * E;
  const __nv_bfloat16* base = qkv + (size_t)n * L * ld + h * DH;
  load_tile_bf16(sQ, base, L, ld, lane);
  load_tile_bf16(sK, base + E, L, ld, lane);
  load_tile_bf16(sV, base + 2 * E, L, ld, lane);
  load_tile_bf16(sO, dctx + (size_t)n * L * E + h * DH, L, E, lane);
  float my_mask_add = 0.f;
  if (lane < L) my_mask_add = (1.0f - (float)mask[(size_t)n * mask_ld + lane]) * -10000.0f;
  __syncwarp();
  const int li = lane < L ? lane : 0;
  float p[LMAX];
  {
    float q[DH];
    load_row_regs(q, sQ + li * DH);
#pragma unroll
    for (int i = 0; i < DH; ++i) q[i] *= 0.125f;
    softmax_row(p, q, sK, relpos + ((size_t)h * L + li) * L, my_mask_add, L, lane);
  }
  float ds[LMAX];
  {
    float dO[DH];
    load_row_regs(dO, sO + li * DH);
    float delta = 0.f;
#pragma unroll
    for (int j = 0; j < LMAX; ++j) {
      ds[j] = (j < L) ? dot_row(dO, sV + j * DH) : 0.f;      // dP_ij
      delta = fmaf(p[j], ds[j], delta);
    }
#pragma unroll
    for (int j = 0; j < LMAX; ++j) ds[j] = p[j] * (ds[j] - delta) * 0.125f;   // dS_ij / sqrt(dh)
  }
  if (lane < L) {
#pragma unroll
    for (int j = 0; j < LMAX; ++j) { sP[lane * (LMAX + 1) + j] = p[j]; sS[lane * (LMAX + 1) + j] = ds[j]; }
  }
  __nv_bfloat16* dbase = dqkv + (size_t)n * L * ld + h * DH;
  {
    float acc[DH];
#pragma unroll
    for (int i = 0; i < DH; ++i) acc[i] = 0.f;
#pragma unroll
    for (int j = 0; j < LMAX; ++j)
      if (j < L) axpy_row(acc, ds[j], sK + j * DH);           // dQ_i = sum_j dS_ij K_j
    if (lane < L) store_row_regs(dbase + (size_t)lane * ld, acc);
  }
  __syncwarp();
  // phase B: lane j owns key/value row j
  {
    float acc[DH];
#pragma unroll
    for (int i = 0; i < DH; ++i) acc[i] = 0.f;
    for (int i = 0; i < L; ++i) axpy_row(acc, sP[i * (LMAX + 1) + li], sO + i * DH);   // dV_j = sum_i P_ij dO_i
    if (lane < L) store_row_regs(dbase + 2 * E + (size_t)lane * ld, acc);
#pragma unroll
    for (int i = 0; i < DH; ++i) acc[i] = 0.f;
    for (int i = 0; i < L; ++i) axpy_row(acc, sS[i * (LMAX + 1) + li], sQ + i * DH);   // dK_j = sum_i dS_ij Q_i
    if (lane < L) store_row_regs(dbase + E + (size_t)lane * ld, acc);
  }
}

}  // namespace tnr

using namespace tnr;

static int check_attn(const char* who, int L, int A, int E) {
  TNR_REQUIRE(L >= 1 && L <= LMAX, "%s: L=%d not supported by the short-sequence kernel (1..%d)", who, L, LMAX);
  TNR_REQUIRE(A * DH == E, "%s: needs head dim 64 (A=%d, E=%d)", who, A, E);
  return 0;
}

extern "C" __attribute__((visibility("default"))) int tnr_attn_relpos_fwd(const void* qkv_bf16, const int64_t* mask, int mask_ld, const float* relpos,
                                   void* ctx_bf16, int n_news, int L, int A, int E, void* stream) {
  if (check_attn("tnr_attn_relpos_fwd", L, A, E)) return 1;
  if (n_news == 0) return 0;
  const long long items = (long long)n_news * A;
  const int grid = (int)((items + ATT_WARPS - 1) / ATT_WARPS);
  const int smem = ATT_WARPS * 2 * LMAX * DH * 2;
  attn_fwd_kernel<<<grid, ATT_WARPS * 32, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv_bf16), mask, mask_ld, relpos, reinterpret_cast<__nv_bfloat16*>(ctx_bf16),
      n_news, L, A, E);
  TNR_LAUNCH_CHECK();
  return 0;
}

extern "C" __attribute__((visibility("default"))) int tnr_attn_relpos_bwd(const void* qkv_bf16, const int64_t* mask, int mask_ld, const float* relpos,
                                   const void* dctx_bf16, void* dqkv_bf16, int n_news, int L, int A, int E, void* stream) {
  if (check_attn("tnr_attn_relpos_bwd", L, A, E)) return 1;
  if (n_news == 0) return 0;
  const long long items = (long long)n_news * A;
  const int grid = (int)((items + ATT_WARPS - 1) / ATT_WARPS);
  const int smem = ATT_WARPS * ATT_BWD_SMEM_PER_WARP;
  static bool attr_done = false;
  if (!attr_done) {
    TNR_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done = true;
  }
  attn_bwd_kernel<<<grid, ATT_WARPS * 32, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv_bf16), mask, mask_ld, relpos,
      reinterpret_cast<const __nv_bfloat16*>(dctx_bf16), reinterpret_cast<__nv_bfloat16*>(dqkv_bf16), n_news, L, A, E);
  TNR_LAUNCH_CHECK();
  return 0;
}
