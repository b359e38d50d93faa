// Fused short-sequence self-attention with relative-position bias (L <= 32, head dim 64).
//   P = softmax(Q K^T / sqrt(dh) + (1 - mask) * -10000 + relbias[h][j - i]);  ctx = dropout(P) V
// Reference: Tiny-NewsRec/tnlrv3/modeling.py:205-231 (multi_head_attention), mask :446-454,
// rel-pos bias :458-463.  position_ids are always arange(L) (:162-163), so the reference's per-forward
// [n, A, L, L] bias is batch-invariant and Toeplitz: a [A, 2L-1] vector indexed by (j - i) + L - 1, which
// each warp keeps in shared memory.  L > 32 is handled by attention_long.cu (forward and backward).
//
// The op is HBM-bound (arithmetic intensity 4*L*E / (8*E) = 15 FLOP/B at L = 30): one warp per
// (news, head) stages its 32x64 Q/K/V (and dO) tiles with cp.async into padded shared memory,
// runs the two small contractions on the warp-level tensor path (ldmatrix + mma.sync m16n8k16,
// bf16 in / fp32 accumulate; the 30x30x64 tiles are far below a tcgen05 128-row atom), keeps
// scores / probabilities in registers and writes the result through shared memory as full
// 128-byte rows.  The backward recomputes P from Q,K (nothing but QKV is saved), regenerates
// the dropout mask from the Philox counter and keeps its tiles unpadded and XOR-swizzled (12 warps per SM).
#include "common.cuh"
#include "mma_sync.cuh"

namespace tnr {

constexpr int DH = 64;
constexpr int LMAX = 32;
constexpr int ATT_WARPS = 4;
constexpr int TS = 72;            // smem row stride (bf16) of a 32 x 64 tile: 144 B, ldmatrix conflict-free
constexpr int TILE_BYTES = LMAX * TS * 2;
constexpr int ATT_FWD_SMEM_PER_WARP = 3 * TILE_BYTES + LMAX * 4 + 2 * LMAX * 4;        // tiles, key mask, rel-pos vector
// backward: Q, K, V, dO as UNPADDED 32 x 128-byte tiles whose 16-byte chunks are XOR-swizzled with the row (ldmatrix,
// cp.async, staging and row stores all conflict-free without the 16 pad bytes per row); the dropout(P) and dS tiles
// (32 x 64 bytes each, swizzled the same way) ALIAS the V tile, which is dead once dP = dO V^T is in registers.
// 16.4 KB per warp: six warps per block, two blocks = 12 warps per SM (padded tiles + a separate dS tile: 21.4 KB, 10).
constexpr int ATT_BWD_WARPS = 6;
constexpr int SW_TILE_BYTES = LMAX * 128;
constexpr int ATT_BWD_SMEM_PER_WARP = 4 * SW_TILE_BYTES + LMAX * 4 + 2 * LMAX * 4;

int attn_long_bwd_launch(const void* qkv_bf16, const int64_t* mask, int mask_ld, const float* relbias, const void* dctx_bf16,
                         void* dqkv_bf16, float* dbias, float* ws, int n_news, int L, int A, int E, const tnr_dropout* drop,
                         cudaStream_t st);
int attn_long_fwd_launch(const void* qkv_bf16, const int64_t* mask, int mask_ld, const float* relbias, void* ctx_bf16,
                         int n_news, int L, int A, int E, const tnr_dropout* drop, cudaStream_t st);

// L rows x 64 bf16 from global (row stride ld) -> padded smem tile; rows >= L are zero-filled.  A lane keeps its 16-byte
// column c = lane % 8 and walks rows r0, r0 + 4, ...: one 64-bit multiply-add per copy for the global address and an
// immediate for the shared one (indexing by idx = lane + 32 it cost 9 - 13 instructions per 16-byte copy: a quarter of
// the backward kernel's instructions were tile addressing)
__device__ __forceinline__ void load_tile_async(__nv_bfloat16* s, const __nv_bfloat16* g, int L, int ld, int lane) {
  const int r0 = lane >> 3, c = lane & 7;
  uint8_t* sp = reinterpret_cast<uint8_t*>(s + r0 * TS + c * 8);
  const uint32_t sa = smem_addr(sp);
  const uint8_t* gp = reinterpret_cast<const uint8_t*>(g + (size_t)r0 * ld + c * 8);
  const size_t gs = (size_t)ld * 8;              // four rows further, in bytes
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    if (r0 + 4 * it < L) cp_async16(sa + it * (4 * TS * 2), gp + it * gs);
    else *reinterpret_cast<uint4*>(sp + it * (4 * TS * 2)) = make_uint4(0, 0, 0, 0);
  }
}

// padded smem tile -> global rows (full 128-byte rows, 16 B per lane)
__device__ __forceinline__ void store_tile(__nv_bfloat16* g, const __nv_bfloat16* s, int L, int ld, int lane) {
  const int r0 = lane >> 3, c = lane & 7;
  const uint8_t* sp = reinterpret_cast<const uint8_t*>(s + r0 * TS + c * 8);
  uint8_t* gp = reinterpret_cast<uint8_t*>(g + (size_t)r0 * ld + c * 8);
  const size_t gs = (size_t)ld * 8;
#pragma unroll
  for (int it = 0; it < 8; ++it)
    if (r0 + 4 * it < L) *reinterpret_cast<uint4*>(gp + it * gs) = *reinterpret_cast<const uint4*>(sp + it * (4 * TS * 2));
}

// C fragments (2 m-tiles x 8 n-tiles of a 32 x 64 fp32 result) -> bf16 smem tile
__device__ __forceinline__ void stage_c64(__nv_bfloat16* s, const float (&o)[2][8][4], int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      *reinterpret_cast<uint32_t*>(s + (mt * 16 + g) * TS + nt * 8 + 2 * t) = pack_bf16(o[mt][nt][0], o[mt][nt][1]);
      *reinterpret_cast<uint32_t*>(s + (mt * 16 + g + 8) * TS + nt * 8 + 2 * t) = pack_bf16(o[mt][nt][2], o[mt][nt][3]);
    }
}

// acc[2][4] (32 x 32) = X[32 x 64] . Y[32 x 64]^T, both row-major padded smem tiles
__device__ __forceinline__ void mma_xyT(float (&acc)[2][4][4], const __nv_bfloat16* sX, const __nv_bfloat16* sY, int lane) {
  uint32_t yb[4][2][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int half = 0; half < 2; ++half)
      ldsm_x4(smem_addr(sY + (nt * 8 + (lane & 7)) * TS + half * 32 + (lane >> 3) * 8), yb[nt][half]);
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t a[4];
      ldsm_x4(smem_addr(sX + (mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * TS + ks * 16 + (lane >> 4) * 8), a);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) mma_bf16(acc[mt][nt], a, yb[nt][ks >> 1][(ks & 1) * 2], yb[nt][ks >> 1][(ks & 1) * 2 + 1]);
    }
}

// out[2][8] (32 x 64) += P[32 x 32] . Y[32 x 64]; P given as A fragments pa[mt][ks], Y row-major smem tile
__device__ __forceinline__ void mma_pY(float (&out)[2][8][4], const uint32_t (&pa)[2][2][4], const __nv_bfloat16* sY, int lane) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    uint32_t b[4];
    ldsm_x4_t(smem_addr(sY + lane * TS + nt * 8), b);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      mma_bf16(out[mt][nt], pa[mt][0], b[0], b[1]);
      mma_bf16(out[mt][nt], pa[mt][1], b[2], b[3]);
    }
  }
}

// C-layout fp32 32x32 -> bf16 A fragments
__device__ __forceinline__ void c_to_a(uint32_t (&pa)[2][2][4], const float (&p)[2][4][4]) {
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      pa[mt][ks][0] = pack_bf16(p[mt][2 * ks][0], p[mt][2 * ks][1]);
      pa[mt][ks][1] = pack_bf16(p[mt][2 * ks][2], p[mt][2 * ks][3]);
      pa[mt][ks][2] = pack_bf16(p[mt][2 * ks + 1][0], p[mt][2 * ks + 1][1]);
      pa[mt][ks][3] = pack_bf16(p[mt][2 * ks + 1][2], p[mt][2 * ks + 1][3]);
    }
}

// per-head rel-pos bias vector [2L-1] (index (j - i) + L - 1) -> this warp's shared-memory copy
__device__ __forceinline__ void load_relbias(float* srel, const float* __restrict__ relbias, int h, int L, int lane) {
  const float* src = relbias + (size_t)h * (2 * L - 1);
  srel[lane] = lane < 2 * L - 1 ? __ldg(src + lane) : 0.f;                 // all 2 * LMAX entries defined: columns >= L read
  srel[lane + 32] = lane + 32 < 2 * L - 1 ? __ldg(src + lane + 32) : 0.f;   // them (and are masked with -inf)
}

// raw QK^T accumulators -> normalised probabilities (C layout; columns >= L become 0)
__device__ __forceinline__ void softmax_frag(float (&s)[2][4][4], const float* smadd, const float* srel, int L, int lane) {
  const int g = lane >> 2, t = lane & 3;
  // the key-side addend of this lane's eight columns (row independent): mask term, -inf past L
  float mb[4][2];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j = nt * 8 + 2 * t + e;
      mb[nt][e] = j < L ? smadd[j] : -INFINITY;
    }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
      const int i = mt * 16 + g + hi * 8;
      const float* relrow = srel + (L - 1 - (i < L ? i : L - 1)) + 2 * t;      // relrow[nt*8 + e] = bias of key j for query i
      float mx = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float v = fmaf(s[mt][nt][hi * 2 + e], 0.125f, mb[nt][e] + relrow[nt * 8 + e]);
          s[mt][nt][hi * 2 + e] = v;
          mx = fmaxf(mx, v);
        }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      // exp(v - mx) = 2^(v log2e - mx log2e): one FFMA + one MUFU (column 0 is always live, so mx is finite)
      const float nm = -mx * 1.4426950408889634f;
      float sum = 0.f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float p = ex2_approx(fmaf(s[mt][nt][hi * 2 + e], 1.4426950408889634f, nm));
          s[mt][nt][hi * 2 + e] = p;
          sum += p;
        }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      const float inv = __fdividef(1.0f, sum);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) s[mt][nt][hi * 2 + e] *= inv;
    }
}

// Dropout keep bits of the probabilities this thread holds for row (mt, hi): bit (nt*2 + e).
// Element (item, i, j) lives in Philox group (item*32 + i)*4 + ((j >> 1) & 3), lane (j >> 3)*2 + (j & 1),
// i.e. linear dropout index ((item*32 + i)*4 + ((j>>1)&3))*8 + (j>>3)*2 + (j&1)  (oracle/dropout.py).
__device__ __forceinline__ uint32_t attn_keep8(const DropCfg& dc, long long item, int i, int t) {
  return dropout_keep8(dc, ((uint64_t)item * 32 + (uint64_t)i) * 4 + (uint64_t)t);
}

__global__ void __launch_bounds__(ATT_WARPS * 32)
attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, const int64_t* __restrict__ mask, int mask_ld,
                const float* __restrict__ relbias, __nv_bfloat16* __restrict__ ctx, int n_news, int L, int A, int E,
                const tnr_dropout drop) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
  const long long item = (long long)blockIdx.x * ATT_WARPS + warp;
  if (item >= (long long)n_news * A) return;
  uint8_t* wbase = smem + (size_t)warp * ATT_FWD_SMEM_PER_WARP;
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(wbase);
  __nv_bfloat16* sK = sQ + LMAX * TS;
  __nv_bfloat16* sV = sK + LMAX * TS;
  float* smadd = reinterpret_cast<float*>(sV + LMAX * TS);
  float* srel = smadd + LMAX;
  const int n = (int)(item / A), h = (int)(item % A);
  const int ld = 3 * E;
  const __nv_bfloat16* base = qkv + (size_t)n * L * ld + h * DH;
  load_tile_async(sQ, base, L, ld, lane);
  load_tile_async(sK, base + E, L, ld, lane);
  load_tile_async(sV, base + 2 * E, L, ld, lane);
  smadd[lane] = lane < L ? (1.0f - (float)mask[(size_t)n * mask_ld + lane]) * -10000.0f : 0.f;
  load_relbias(srel, relbias, h, L, lane);
  const DropCfg dc = load_drop(drop);
  cp_async_wait_all();
  __syncwarp();

  float s[2][4][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int c = 0; c < 4; ++c) s[mt][nt][c] = 0.f;
  mma_xyT(s, sQ, sK, lane);
  softmax_frag(s, smadd, srel, L, lane);
  if (dc.thr16 != 0) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int hi = 0; hi < 2; ++hi) {
        const uint32_t keep = attn_keep8(dc, item, mt * 16 + g + hi * 8, t);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e)
            s[mt][nt][hi * 2 + e] = ((keep >> (nt * 2 + e)) & 1u) ? s[mt][nt][hi * 2 + e] * dc.scale : 0.f;
      }
  }
  uint32_t pa[2][2][4];
  c_to_a(pa, s);
  float o[2][8][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int c = 0; c < 4; ++c) o[mt][nt][c] = 0.f;
  mma_pY(o, pa, sV, lane);
  __syncwarp();
  stage_c64(sQ, o, lane);
  __syncwarp();
  store_tile(ctx + (size_t)n * L * E + h * DH, sQ, L, E, lane);
}

// ---- swizzled tiles of the backward kernel ------------------------------------------------------------------
// 32 x 64 bf16 tile, 128-byte rows: element (r, c) at r*128 + ((c/8 ^ (r & 7)) * 16) + (c % 8) * 2 bytes
__device__ __forceinline__ uint32_t sw_off(int r, int chunk) { return (uint32_t)(r * 128 + ((chunk ^ (r & 7)) << 4)); }
// 32 x 32 bf16 tile, 64-byte rows (two rows per 128-byte line): chunk (0..3) XOR ((r >> 1) & 3)
__device__ __forceinline__ uint32_t pt_off(int r, int chunk) { return (uint32_t)(r * 64 + ((chunk ^ ((r >> 1) & 3)) << 4)); }

// a lane keeps its 16-byte column c and walks rows r0 + 4 it (r0 = lane / 8 < 4): r & 7 alternates between r0 and r0 + 4,
// so the swizzled chunk alternates between x0 and x0 ^ 64 bytes and every shared address is base + immediate
__device__ __forceinline__ void load_tile_async_sw(uint8_t* s, const __nv_bfloat16* g, int L, int ld, int lane) {
  const int r0 = lane >> 3, c = lane & 7;
  const uint32_t x0 = (uint32_t)((c ^ r0) << 4);
  uint8_t* sp0 = s + r0 * 128 + x0;
  uint8_t* sp1 = s + r0 * 128 + (x0 ^ 64u);
  const uint32_t sa0 = smem_addr(sp0), sa1 = smem_addr(sp1);
  const uint8_t* gp = reinterpret_cast<const uint8_t*>(g + (size_t)r0 * ld + c * 8);
  const size_t gs = (size_t)ld * 8;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    if (r0 + 4 * it < L) cp_async16(((it & 1) ? sa1 : sa0) + it * 512, gp + it * gs);
    else *reinterpret_cast<uint4*>(((it & 1) ? sp1 : sp0) + it * 512) = make_uint4(0, 0, 0, 0);
  }
}
__device__ __forceinline__ void store_tile_sw(__nv_bfloat16* g, const uint8_t* s, int L, int ld, int lane) {
  const int r0 = lane >> 3, c = lane & 7;
  const uint32_t x0 = (uint32_t)((c ^ r0) << 4);
  const uint8_t* sp0 = s + r0 * 128 + x0;
  const uint8_t* sp1 = s + r0 * 128 + (x0 ^ 64u);
  uint8_t* gp = reinterpret_cast<uint8_t*>(g + (size_t)r0 * ld + c * 8);
  const size_t gs = (size_t)ld * 8;
#pragma unroll
  for (int it = 0; it < 8; ++it)
    if (r0 + 4 * it < L)
      *reinterpret_cast<uint4*>(gp + it * gs) = *reinterpret_cast<const uint4*>(((it & 1) ? sp1 : sp0) + it * 512);
}
__device__ __forceinline__ void stage_c64_sw(uint8_t* s, const float (&o)[2][8][4], int lane) {
  const int m = lane >> 3;                       // matrix of the x4 store this lane addresses: rows (m & 1) * 8.., chunk +(m >> 1)
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      const uint32_t r[4] = {pack_bf16(o[mt][2 * np][0], o[mt][2 * np][1]), pack_bf16(o[mt][2 * np][2], o[mt][2 * np][3]),
                             pack_bf16(o[mt][2 * np + 1][0], o[mt][2 * np + 1][1]),
                             pack_bf16(o[mt][2 * np + 1][2], o[mt][2 * np + 1][3])};
      stsm_x4(smem_addr(s + sw_off(mt * 16 + (m & 1) * 8 + (lane & 7), 2 * np + (m >> 1))), r);
    }
}
// A-fragment-shaped registers of a 32 x 32 bf16 tile (c_to_a) -> swizzled 64-byte-row tile
__device__ __forceinline__ void stage_p32_sw(uint8_t* s, const uint32_t (&pa)[2][2][4], int lane) {
  const int m = lane >> 3;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
      stsm_x4(smem_addr(s + pt_off(mt * 16 + (m & 1) * 8 + (lane & 7), 2 * ks + (m >> 1))), pa[mt][ks]);
}
// column sums of a staged 32 x 64 tile on the tensor path: ones(16 x 32) . tile; every accumulator row holds the sums,
// lane (g, t) adds those of columns g*8 + 2t, +1 (rows >= L of the staged gradients are exact zeros)
__device__ __forceinline__ void tile_colsum_mma(float* __restrict__ out, const uint8_t* s, int lane) {
  const uint32_t ones[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u};
  const int g = lane >> 2, t = lane & 3;
  float v0 = 0.f, v1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    uint32_t b[4];
    ldsm_x4_t(smem_addr(s + sw_off(lane, nt)), b);
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    mma_bf16(c, ones, b[0], b[1]);
    mma_bf16(c, ones, b[2], b[3]);
    if (nt == g) { v0 = c[0]; v1 = c[1]; }
  }
  atomicAdd(reinterpret_cast<float2*>(out + g * 8 + 2 * t), make_float2(v0, v1));      // one 8-byte RED (sm_90+)
}
// acc (32 x 32) = X . Y^T, both 32 x 64 swizzled tiles
__device__ __forceinline__ void mma_xyT_sw(float (&acc)[2][4][4], const uint8_t* sX, const uint8_t* sY, int lane) {
  uint32_t yb[4][2][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int half = 0; half < 2; ++half)
      ldsm_x4(smem_addr(sY + sw_off(nt * 8 + (lane & 7), half * 4 + (lane >> 3))), yb[nt][half]);
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t a[4];
      ldsm_x4(smem_addr(sX + sw_off(mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, ks * 2 + (lane >> 4))), a);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) mma_bf16(acc[mt][nt], a, yb[nt][ks >> 1][(ks & 1) * 2], yb[nt][ks >> 1][(ks & 1) * 2 + 1]);
    }
}
// out (32 x 64) += P (32 x 32, A fragments) . Y (32 x 64 swizzled tile)
__device__ __forceinline__ void mma_pY_sw(float (&out)[2][8][4], const uint32_t (&pa)[2][2][4], const uint8_t* sY, int lane) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    uint32_t b[4];
    ldsm_x4_t(smem_addr(sY + sw_off(lane, nt)), b);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      mma_bf16(out[mt][nt], pa[mt][0], b[0], b[1]);
      mma_bf16(out[mt][nt], pa[mt][1], b[2], b[3]);
    }
  }
}
// A fragments of X^T, X a 32 x 32 swizzled (64-byte rows) tile
__device__ __forceinline__ void load_xT_frags_sw(uint32_t (&pa)[2][2][4], const uint8_t* sX, int lane) {
  const int mi = lane >> 3;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
      ldsm_x4_t(smem_addr(sX + pt_off(ks * 16 + (lane & 7) + (mi >> 1) * 8, mt * 2 + (mi & 1))), pa[mt][ks]);
}

__global__ void __launch_bounds__(ATT_BWD_WARPS * 32, 2)
attn_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const int64_t* __restrict__ mask, int mask_ld,
                const float* __restrict__ relbias, const __nv_bfloat16* __restrict__ dctx,
                __nv_bfloat16* __restrict__ dqkv, float* __restrict__ dbias, int n_news, int L, int A, int E,
                const tnr_dropout drop) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
  const long long item = (long long)blockIdx.x * ATT_BWD_WARPS + warp;
  if (item >= (long long)n_news * A) return;
  uint8_t* sQ = smem + (size_t)warp * ATT_BWD_SMEM_PER_WARP;
  uint8_t* sK = sQ + SW_TILE_BYTES;
  uint8_t* sV = sK + SW_TILE_BYTES;
  uint8_t* sO = sV + SW_TILE_BYTES;
  uint8_t* sP = sV;                               // dropout(P) and dS alias V: written only after the dP product has consumed it
  uint8_t* sS = sV + LMAX * 64;
  float* smadd = reinterpret_cast<float*>(sO + SW_TILE_BYTES);
  float* srel = smadd + LMAX;
  const int n = (int)(item / A), h = (int)(item % A);
  const int ld = 3 * E;
  const int g = lane >> 2, t = lane & 3;
  const __nv_bfloat16* base = qkv + (size_t)n * L * ld + h * DH;
  load_tile_async_sw(sQ, base, L, ld, lane);
  load_tile_async_sw(sK, base + E, L, ld, lane);
  load_tile_async_sw(sV, base + 2 * E, L, ld, lane);
  load_tile_async_sw(sO, dctx + (size_t)n * L * E + h * DH, L, E, lane);
  smadd[lane] = lane < L ? (1.0f - (float)mask[(size_t)n * mask_ld + lane]) * -10000.0f : 0.f;
  load_relbias(srel, relbias, h, L, lane);
  const DropCfg dc = load_drop(drop);
  cp_async_wait_all();
  __syncwarp();

  float p[2][4][4], dp[2][4][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int c = 0; c < 4; ++c) { p[mt][nt][c] = 0.f; dp[mt][nt][c] = 0.f; }
  mma_xyT_sw(p, sQ, sK, lane);
  softmax_frag(p, smadd, srel, L, lane);
  mma_xyT_sw(dp, sO, sV, lane);                 // dP = dO V^T
  __syncwarp();                                 // every lane is done reading V before dropout(P) / dS go into its tile
  // dS = P o (dP_eff - delta) / 8, P_drop = P o keep * scale; both to smem (bf16) for the transposed products
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
      const int i = mt * 16 + g + hi * 8;
      uint32_t keep = 0xffu;
      if (dc.thr16 != 0) keep = attn_keep8(dc, item, i, t);
      float delta = 0.f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float k = ((keep >> (nt * 2 + e)) & 1u) ? dc.scale : 0.f;
          const float de = dp[mt][nt][hi * 2 + e] * k;          // gradient w.r.t. the pre-dropout probability
          dp[mt][nt][hi * 2 + e] = de;
          delta = fmaf(p[mt][nt][hi * 2 + e], de, delta);
        }
      delta += __shfl_xor_sync(0xffffffffu, delta, 1);
      delta += __shfl_xor_sync(0xffffffffu, delta, 2);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        float pd[2], ds[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float pv = p[mt][nt][hi * 2 + e];
          const float k = ((keep >> (nt * 2 + e)) & 1u) ? dc.scale : 0.f;
          pd[e] = pv * k;
          ds[e] = pv * (dp[mt][nt][hi * 2 + e] - delta) * 0.125f;
          dp[mt][nt][hi * 2 + e] = ds[e];
          p[mt][nt][hi * 2 + e] = pd[e];
        }
      }
    }
  __nv_bfloat16* dbase = dqkv + (size_t)n * L * ld + h * DH;
  float acc[2][8][4];
  uint32_t pa[2][2][4];
  c_to_a(pa, p);
  stage_p32_sw(sP, pa, lane);                     // dropout(P), bf16
  // dQ = dS K
  c_to_a(pa, dp);
  stage_p32_sw(sS, pa, lane);                     // dS, bf16 (the same registers are the A fragments of dS K)
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[mt][nt][c] = 0.f;
  mma_pY_sw(acc, pa, sK, lane);
  __syncwarp();                                   // all lanes done reading sK; sP / sS visible
  stage_c64_sw(sK, acc, lane);
  __syncwarp();
  store_tile_sw(dbase, sK, L, ld, lane);
  if (dbias != nullptr) tile_colsum_mma(dbias + h * DH, sK, lane);
  // dV = P_drop^T dO, staged over dO once every lane has its fragments
  load_xT_frags_sw(pa, sP, lane);
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[mt][nt][c] = 0.f;
  mma_pY_sw(acc, pa, sO, lane);
  __syncwarp();
  stage_c64_sw(sO, acc, lane);
  __syncwarp();
  store_tile_sw(dbase + 2 * E, sO, L, ld, lane);
  if (dbias != nullptr) tile_colsum_mma(dbias + 2 * E + h * DH, sO, lane);
  // dK = dS^T Q
  load_xT_frags_sw(pa, sS, lane);
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[mt][nt][c] = 0.f;
  mma_pY_sw(acc, pa, sQ, lane);
  __syncwarp();
  stage_c64_sw(sQ, acc, lane);
  __syncwarp();
  store_tile_sw(dbase + E, sQ, L, ld, lane);
  if (dbias != nullptr) tile_colsum_mma(dbias + E + h * DH, sQ, lane);
}

}  // namespace tnr

using namespace tnr;

static int check_attn(const char* who, int L, int A, int E, int lmax) {
  TNR_REQUIRE(L >= 1 && L <= lmax, "%s: L=%d not supported (1..%d)", who, L, lmax);
  TNR_REQUIRE(A * DH == E, "%s: needs head dim 64 (A=%d, E=%d)", who, A, E);
  return 0;
}

extern "C" __attribute__((visibility("default"))) int tnr_attn_relpos_fwd(const void* qkv_bf16, const int64_t* mask, int mask_ld, const float* relbias,
                                   void* ctx_bf16, int n_news, int L, int A, int E, const tnr_dropout* drop, void* stream) {
  if (check_attn("tnr_attn_relpos_fwd", L, A, E, 512)) return 1;
  if (n_news == 0) return 0;
  if (L > LMAX)
    return attn_long_fwd_launch(qkv_bf16, mask, mask_ld, relbias, ctx_bf16, n_news, L, A, E, drop,
                                reinterpret_cast<cudaStream_t>(stream));
  const long long items = (long long)n_news * A;
  const int grid = (int)((items + ATT_WARPS - 1) / ATT_WARPS);
  const int smem = ATT_WARPS * ATT_FWD_SMEM_PER_WARP;
  TNR_SET_SMEM(attn_fwd_kernel, smem);
  attn_fwd_kernel<<<grid, ATT_WARPS * 32, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv_bf16), mask, mask_ld, relbias, reinterpret_cast<__nv_bfloat16*>(ctx_bf16),
      n_news, L, A, E, drop_or_none(drop));
  TNR_LAUNCH_CHECK();
  return 0;
}

extern "C" __attribute__((visibility("default"))) int tnr_attn_relpos_bwd(const void* qkv_bf16, const int64_t* mask, int mask_ld, const float* relbias,
                                   const void* dctx_bf16, void* dqkv_bf16, float* dbias_qkv, float* workspace, int n_news, int L,
                                   int A, int E, const tnr_dropout* drop, void* stream) {
  if (check_attn("tnr_attn_relpos_bwd", L, A, E, 512)) return 1;
  if (n_news == 0) return 0;
  if (L > LMAX)
    return attn_long_bwd_launch(qkv_bf16, mask, mask_ld, relbias, dctx_bf16, dqkv_bf16, dbias_qkv, workspace, n_news, L, A, E,
                                drop, reinterpret_cast<cudaStream_t>(stream));
  const long long items = (long long)n_news * A;
  const int grid = (int)((items + ATT_BWD_WARPS - 1) / ATT_BWD_WARPS);
  const int smem = ATT_BWD_WARPS * ATT_BWD_SMEM_PER_WARP;
  TNR_SET_SMEM(attn_bwd_kernel, smem);
  attn_bwd_kernel<<<grid, ATT_BWD_WARPS * 32, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv_bf16), mask, mask_ld, relbias,
      reinterpret_cast<const __nv_bfloat16*>(dctx_bf16), reinterpret_cast<__nv_bfloat16*>(dqkv_bf16), dbias_qkv, n_news, L,
      A, E, drop_or_none(drop));
  TNR_LAUNCH_CHECK();
  return 0;
}
