// HBM-bound row kernels of the encoder: embedding gather + LayerNorm, LayerNorm forward /
// backward, column sums (bias gradients).  One warp per row, 16-byte vector loads, the row
// lives in registers (E <= 1024), warp-shuffle reductions.
#include "common.cuh"
#include <cstdlib>

namespace tnr {

constexpr int ROWS_PER_BLOCK = 8;   // 8 warps

template <int VPL>   // 8-element vectors per lane: E = VPL * 256
__device__ __forceinline__ void ln_row(float (&x)[VPL * 8], const float* __restrict__ gamma,
                                       const float* __restrict__ beta, float eps, int lane,
                                       __nv_bfloat16* __restrict__ out, const DropCfg& dc, uint64_t row) {
  constexpr int E = VPL * 256;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL * 8; ++i) s += x[i];
  const float mean = warp_sum(s) * (1.0f / E);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL * 8; ++i) { const float d = x[i] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / E) + eps);
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int col = (v * 32 + lane) * 8;
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + col);
    const float4 g1 = *reinterpret_cast<const float4*>(gamma + col + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(beta + col);
    const float4 b1 = *reinterpret_cast<const float4*>(beta + col + 4);
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = (x[v * 8 + i] - mean) * rstd * g[i] + b[i];
    if (dc.thr16 != 0) {
      const uint32_t keep = dropout_keep8(dc, (row * (uint64_t)E + (uint64_t)col) >> 3);
#pragma unroll
      for (int i = 0; i < 8; ++i) y[i] = ((keep >> i) & 1u) ? y[i] * dc.scale : 0.f;
    }
    *reinterpret_cast<bf16x8*>(out + col) = pack8(y);
  }
}

// out[t, :] = dropout(LN(word[id[t]] + pos[t % L] + type[0]))         (tnlrv3/modeling.py:168-177)
// ids: int64, row n at ids + n * ids_ld, L tokens per row.  word table bf16 or fp32.
// Persistent warps, each bound to ONE position l: gamma, beta and (pos[l] + type0) stay in registers
// and only the gathered word row and the output row move per token (the first version re-read the
// four fp32 parameter rows per token: 13.5 KB of L1 traffic for 3 KB of HBM traffic, 37 % of HBM peak).
// Warp w of the grid owns position w % L and news rows w / L, w / L + groups, ...; two rows in flight.
template <int VPL, bool WORD_BF16>
__device__ __forceinline__ void embed_load_row(bf16x8 (&raw)[VPL * (WORD_BF16 ? 1 : 2)], const void* __restrict__ word,
                                               long long id, int lane) {
  constexpr int E = VPL * 256;
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int col = (v * 32 + lane) * 8;
    if (WORD_BF16) {
      raw[v] = *reinterpret_cast<const bf16x8*>(reinterpret_cast<const __nv_bfloat16*>(word) + (size_t)id * E + col);
    } else {
      const float* wp = reinterpret_cast<const float*>(word) + (size_t)id * E + col;
      raw[2 * v] = *reinterpret_cast<const bf16x8*>(wp);          // 16 raw bytes = 4 floats
      raw[2 * v + 1] = *reinterpret_cast<const bf16x8*>(wp + 4);
    }
  }
}

constexpr int EMB_WARPS = 12;      // 12 warps x <=168 registers: one block per SM, 24 rows in flight
template <int VPL, bool WORD_BF16>
__global__ void __launch_bounds__(EMB_WARPS * 32)
embed_ln_kernel(const int64_t* __restrict__ ids, int ids_ld, int L, int n_rows, int groups, int vocab,
                const void* __restrict__ word, const float* __restrict__ pos, const float* __restrict__ type0,
                const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                __nv_bfloat16* __restrict__ out, const tnr_dropout drop) {
  constexpr int E = VPL * 256;
  constexpr int NRAW = VPL * (WORD_BF16 ? 1 : 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * EMB_WARPS + warp;
  const int l = w % L, grp = w / L;
  if (grp >= groups) return;
  const DropCfg dc = load_drop(drop);
  float g[VPL * 8], bt[VPL * 8], pt[VPL * 8];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int col = (v * 32 + lane) * 8;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 gg = *reinterpret_cast<const float4*>(gamma + col + 4 * h);
      const float4 bb = *reinterpret_cast<const float4*>(beta + col + 4 * h);
      const float4 pp = *reinterpret_cast<const float4*>(pos + (size_t)l * E + col + 4 * h);
      const float4 tt = *reinterpret_cast<const float4*>(type0 + col + 4 * h);
      g[v * 8 + 4 * h + 0] = gg.x; g[v * 8 + 4 * h + 1] = gg.y; g[v * 8 + 4 * h + 2] = gg.z; g[v * 8 + 4 * h + 3] = gg.w;
      bt[v * 8 + 4 * h + 0] = bb.x; bt[v * 8 + 4 * h + 1] = bb.y; bt[v * 8 + 4 * h + 2] = bb.z; bt[v * 8 + 4 * h + 3] = bb.w;
      // batch-invariant part folded once: w + (p + t) vs the reference's (w + p) + t differ by an fp32
      // ulp at most, far below the bf16 output resolution
      pt[v * 8 + 4 * h + 0] = pp.x + tt.x; pt[v * 8 + 4 * h + 1] = pp.y + tt.y;
      pt[v * 8 + 4 * h + 2] = pp.z + tt.z; pt[v * 8 + 4 * h + 3] = pp.w + tt.w;
    }
  }
  for (int n0 = grp; n0 < n_rows; n0 += 2 * groups) {
    const int n1 = n0 + groups;
    long long id0 = ids[(size_t)n0 * ids_ld + l];
    long long id1 = n1 < n_rows ? ids[(size_t)n1 * ids_ld + l] : 0;
    id0 = id0 < 0 ? 0 : (id0 >= vocab ? vocab - 1 : id0);     // defensive clamp (torch would raise)
    id1 = id1 < 0 ? 0 : (id1 >= vocab ? vocab - 1 : id1);
    bf16x8 raw0[NRAW], raw1[NRAW];
    embed_load_row<VPL, WORD_BF16>(raw0, word, id0, lane);
    if (n1 < n_rows) embed_load_row<VPL, WORD_BF16>(raw1, word, id1, lane);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int n = k ? n1 : n0;
      if (n >= n_rows) break;
      float x[VPL * 8];
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        if (WORD_BF16) {
          unpack8(k ? raw1[v] : raw0[v], &x[v * 8]);
        } else {
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < 4; ++i) x[v * 8 + 4 * h + i] = __uint_as_float((k ? raw1 : raw0)[2 * v + h].u[i]);
        }
      }
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < VPL * 8; ++i) { x[i] += pt[i]; s += x[i]; }
      const float mean = warp_sum(s) * (1.0f / E);
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < VPL * 8; ++i) { const float d = x[i] - mean; q += d * d; }
      const float rstd = rsqrtf(warp_sum(q) * (1.0f / E) + eps);
      const size_t t = (size_t)n * L + l;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int col = (v * 32 + lane) * 8;
        float y[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = (x[v * 8 + i] - mean) * rstd * g[v * 8 + i] + bt[v * 8 + i];
        if (dc.thr16 != 0) {
          const uint32_t keep = dropout_keep8(dc, ((uint64_t)t * (uint64_t)E + (uint64_t)col) >> 3);
#pragma unroll
          for (int i = 0; i < 8; ++i) y[i] = ((keep >> i) & 1u) ? y[i] * dc.scale : 0.f;
        }
        *reinterpret_cast<bf16x8*>(out + t * E + col) = pack8(y);
      }
    }
  }
}

// y = LN(x) for a bf16 "pre-LN" buffer (dense + bias + residual written by the GEMM epilogue).
// Persistent warps: gamma / beta stay in registers (re-loading them per row made the kernel
// L1-bound: ncu l1tex 85 %, dram 36 %), two rows in flight per warp.
template <int VPL>
__global__ void __launch_bounds__(ROWS_PER_BLOCK * 32)
layernorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, int rows, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ y, int reverse) {
  constexpr int E = VPL * 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float g[VPL * 8], bt[VPL * 8];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int col = (v * 32 + lane) * 8;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 gg = *reinterpret_cast<const float4*>(gamma + col + 4 * h);
      const float4 bb = *reinterpret_cast<const float4*>(beta + col + 4 * h);
      g[v * 8 + 4 * h + 0] = gg.x; g[v * 8 + 4 * h + 1] = gg.y; g[v * 8 + 4 * h + 2] = gg.z; g[v * 8 + 4 * h + 3] = gg.w;
      bt[v * 8 + 4 * h + 0] = bb.x; bt[v * 8 + 4 * h + 1] = bb.y; bt[v * 8 + 4 * h + 2] = bb.z; bt[v * 8 + 4 * h + 3] = bb.w;
    }
  }
  const int stride = gridDim.x * ROWS_PER_BLOCK;
  for (int r0 = blockIdx.x * ROWS_PER_BLOCK + warp; r0 < rows; r0 += 2 * stride) {
    const int r1 = r0 + stride;
    // reverse: sweep the rows from the END -- the rows the producing GEMM wrote last are the ones still in L2
    const int a0 = reverse ? rows - 1 - r0 : r0, a1 = reverse ? rows - 1 - r1 : r1;
    bf16x8 raw0[VPL], raw1[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) raw0[v] = *reinterpret_cast<const bf16x8*>(x + (size_t)a0 * E + (v * 32 + lane) * 8);
    if (r1 < rows) {
#pragma unroll
      for (int v = 0; v < VPL; ++v) raw1[v] = *reinterpret_cast<const bf16x8*>(x + (size_t)a1 * E + (v * 32 + lane) * 8);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if ((k ? r1 : r0) >= rows) break;
      const int r = k ? a1 : a0;
      float xv[VPL * 8];
#pragma unroll
      for (int v = 0; v < VPL; ++v) unpack8(k ? raw1[v] : raw0[v], &xv[v * 8]);
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < VPL * 8; ++i) s += xv[i];
      const float mean = warp_sum(s) * (1.0f / E);
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < VPL * 8; ++i) { const float d = xv[i] - mean; q += d * d; }
      const float rstd = rsqrtf(warp_sum(q) * (1.0f / E) + eps);
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = (xv[v * 8 + i] - mean) * rstd * g[v * 8 + i] + bt[v * 8 + i];
        *reinterpret_cast<bf16x8*>(y + (size_t)r * E + (v * 32 + lane) * 8) = pack8(o);
      }
    }
  }
}

// LayerNorm backward.  xhat recomputed from the saved pre-LN row.
//   dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma
//   dgamma += sum_rows dy * xhat ; dbeta += sum_rows dy ; dsum += sum_rows dx_drop (the bias gradient of
//   the dense layer in front of this LayerNorm -- saves a separate column-sum pass over dx)
// Persistent grid, one 12-warp block per SM.  gamma and the three per-column accumulators (dgamma, dbeta,
// dsum) live in registers; each warp streams its rows through a private 3-stage shared-memory ring filled by
// cp.async, so two further rows are in flight while the current one is reduced.  History: rows held in
// registers (171 registers, 29 % of HBM peak) -> double-buffered stage with gamma re-read from L1 and dbeta
// accumulated in shared memory (~20 KB of LSU traffic per 6 KB row pair, 55 %) -> this version.
__device__ __forceinline__ void ln_cp_async16(void* smem_dst, const void* src) {
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

constexpr int LNB_WARPS = 12;      // 384 threads, <= 168 registers: one block per SM
constexpr int LNB_STAGES = 3;

template <int VPL>
__global__ void __launch_bounds__(LNB_WARPS * 32, 1)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x, int rows,
                     const float* __restrict__ gamma, float eps, __nv_bfloat16* __restrict__ dx,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, __nv_bfloat16* __restrict__ dx_drop,
                     float* __restrict__ dsum, const tnr_dropout drop) {
  constexpr int E = VPL * 256;
  extern __shared__ __align__(16) uint8_t ln_smem[];
  const DropCfg dc = load_drop(drop);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* s_acc = reinterpret_cast<float*>(ln_smem);                       // [3][E] block totals of dgamma, dbeta, dsum
  // per warp: LNB_STAGES x {x row, dy row} bf16
  __nv_bfloat16* stage = reinterpret_cast<__nv_bfloat16*>(ln_smem + 3 * E * 4) + (size_t)warp * LNB_STAGES * 2 * E;
  for (int i = threadIdx.x; i < 3 * E; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  float gm[VPL * 8], acc_dg[VPL * 8], acc_db[VPL * 8], acc_ds[VPL * 8];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int col = (v * 32 + lane) * 8;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + col));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + col + 4));
    gm[v * 8 + 0] = g0.x; gm[v * 8 + 1] = g0.y; gm[v * 8 + 2] = g0.z; gm[v * 8 + 3] = g0.w;
    gm[v * 8 + 4] = g1.x; gm[v * 8 + 5] = g1.y; gm[v * 8 + 6] = g1.z; gm[v * 8 + 7] = g1.w;
  }
#pragma unroll
  for (int i = 0; i < VPL * 8; ++i) { acc_dg[i] = 0.f; acc_db[i] = 0.f; acc_ds[i] = 0.f; }
  const int stride = gridDim.x * LNB_WARPS;
  auto prefetch = [&](int st, int r) {
    if (r < rows) {
      __nv_bfloat16* sx = stage + st * 2 * E;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int col = (v * 32 + lane) * 8;
        ln_cp_async16(sx + col, x + (size_t)r * E + col);
        ln_cp_async16(sx + E + col, dy + (size_t)r * E + col);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");       // one group per slot, empty past the end
  };
  int r = blockIdx.x * LNB_WARPS + warp;
  prefetch(0, r);
  prefetch(1, r + stride);
  int st = 0;
  for (; r < rows; r += stride) {
    int st2 = st + 2; if (st2 >= LNB_STAGES) st2 -= LNB_STAGES;
    prefetch(st2, r + 2 * stride);
    asm volatile("cp.async.wait_group 2;" ::: "memory");
    __syncwarp();
    bf16x8 rx[VPL], rd[VPL];
    {
      const __nv_bfloat16* sx = stage + st * 2 * E;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        rx[v] = *reinterpret_cast<const bf16x8*>(sx + (v * 32 + lane) * 8);
        rd[v] = *reinterpret_cast<const bf16x8*>(sx + E + (v * 32 + lane) * 8);
      }
    }
    __syncwarp();                      // stage consumed: a later prefetch may overwrite it
    if (++st == LNB_STAGES) st = 0;
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      float t[8];
      unpack8(rx[v], t);
#pragma unroll
      for (int i = 0; i < 8; ++i) s += t[i];
    }
    const float mean = warp_sum(s) * (1.0f / E);
    float q = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      float t[8];
      unpack8(rx[v], t);
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float d = t[i] - mean; q += d * d; }
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / E) + eps);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      float t[8], d[8];
      unpack8(rx[v], t);
      unpack8(rd[v], d);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float xh = (t[i] - mean) * rstd;
        acc_dg[v * 8 + i] = fmaf(d[i], xh, acc_dg[v * 8 + i]);
        acc_db[v * 8 + i] += d[i];
        const float gg = d[i] * gm[v * 8 + i];
        sg += gg;
        sgx = fmaf(gg, xh, sgx);
      }
    }
    sg = warp_sum(sg) * (1.0f / E);
    sgx = warp_sum(sgx) * (1.0f / E);
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      float t[8], d[8], o[8];
      const int col = (v * 32 + lane) * 8;
      unpack8(rx[v], t);
      unpack8(rd[v], d);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = rstd * (d[i] * gm[v * 8 + i] - sg - (t[i] - mean) * rstd * sgx);
      *reinterpret_cast<bf16x8*>(dx + (size_t)r * E + col) = pack8(o);
      if (dc.thr16 != 0) {
        const uint32_t keep = dropout_keep8(dc, ((uint64_t)r * (uint64_t)E + (uint64_t)col) >> 3);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = ((keep >> i) & 1u) ? o[i] * dc.scale : 0.f;
      }
      if (dx_drop != nullptr) *reinterpret_cast<bf16x8*>(dx_drop + (size_t)r * E + col) = pack8(o);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc_ds[v * 8 + i] += o[i];
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
  for (int v = 0; v < VPL; ++v)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int col = (v * 32 + lane) * 8 + i;
      atomicAdd(&s_acc[col], acc_dg[v * 8 + i]);
      atomicAdd(&s_acc[E + col], acc_db[v * 8 + i]);
      if (dsum != nullptr) atomicAdd(&s_acc[2 * E + col], acc_ds[v * 8 + i]);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < E; i += blockDim.x) {
    atomicAdd(dgamma + i, s_acc[i]);
    atomicAdd(dbeta + i, s_acc[E + i]);
    if (dsum != nullptr) atomicAdd(dsum + i, s_acc[2 * E + i]);
  }
}

// out[c] += sum_r x[r, c]   (bias gradients).  bf16 input, fp32 atomics; cols % 8 == 0.
__global__ void __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16* __restrict__ x, int rows, int cols, int ld, float* __restrict__ out) {
  // block handles a 64-column strip (8 lanes x 8 cols) and a row slab; 32 row-lanes per block
  const int cgroup = threadIdx.x & 7;            // 8 column vectors of 8
  const int rlane = threadIdx.x >> 3;            // 32 row lanes
  const int col = (blockIdx.x * 8 + cgroup) * 8;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (col < cols) {
    for (int r = blockIdx.y * 32 + rlane; r < rows; r += gridDim.y * 32) {
      float v[8];
      unpack8(*reinterpret_cast<const bf16x8*>(x + (size_t)r * ld + col), v);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += v[i];
    }
  }
  __shared__ float red[32][65];
#pragma unroll
  for (int i = 0; i < 8; ++i) red[rlane][cgroup * 8 + i] = acc[i];
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 32; ++r) s += red[r][threadIdx.x];
    const int c = blockIdx.x * 64 + threadIdx.x;
    if (c < cols) atomicAdd(out + c, s);
  }
}

}  // namespace tnr

using namespace tnr;

#define DISPATCH_VPL(E, CALL)                                   \
  switch ((E) / 256) {                                          \
    case 1: { constexpr int VPL = 1; CALL; break; }             \
    case 2: { constexpr int VPL = 2; CALL; break; }             \
    case 3: { constexpr int VPL = 3; CALL; break; }             \
    case 4: { constexpr int VPL = 4; CALL; break; }             \
    default: set_error("hidden size %d not supported (need 256/512/768/1024)", (E)); return 1; \
  }

extern "C" __attribute__((visibility("default"))) int tnr_embed_ln_fwd(const int64_t* ids, int ids_ld, int n_rows, int L, int vocab, const void* word,
                                int word_dtype, const float* pos, const float* type0, const float* gamma,
                                const float* beta, float eps, int E, void* out_bf16, const tnr_dropout* drop, void* stream) {
  TNR_REQUIRE(E % 256 == 0, "tnr_embed_ln_fwd: E=%d must be a multiple of 256", E);
  TNR_REQUIRE(n_rows >= 0 && L > 0, "tnr_embed_ln_fwd: bad shape");
  if (n_rows == 0) return 0;
  // warps = groups * L: every warp owns one position; one block per SM
  const long long want_warps = (long long)num_sms() * EMB_WARPS;
  int groups = (int)(want_warps / L);
  if (groups < 1) groups = 1;
  if (groups > n_rows) groups = n_rows;
  const int grid = (int)(((long long)groups * L + EMB_WARPS - 1) / EMB_WARPS);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  if (word_dtype == TNR_BF16) {
    DISPATCH_VPL(E, (embed_ln_kernel<VPL, true><<<grid, EMB_WARPS * 32, 0, st>>>(
                        ids, ids_ld, L, n_rows, groups, vocab, word, pos, type0, gamma, beta, eps, out, drop_or_none(drop))));
  } else {
    DISPATCH_VPL(E, (embed_ln_kernel<VPL, false><<<grid, EMB_WARPS * 32, 0, st>>>(
                        ids, ids_ld, L, n_rows, groups, vocab, word, pos, type0, gamma, beta, eps, out, drop_or_none(drop))));
  }
  TNR_LAUNCH_CHECK();
  return 0;
}

extern "C" __attribute__((visibility("default"))) int tnr_layernorm_fwd(const void* x_bf16, int rows, int E, const float* gamma, const float* beta, float eps,
                                 void* y_bf16, void* stream) {
  if (rows == 0) return 0;
  int grid = (rows + 2 * ROWS_PER_BLOCK - 1) / (2 * ROWS_PER_BLOCK);
  const int cap_f = num_sms() * 2;
  if (grid > cap_f) grid = cap_f;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int reverse = 0;   // sweeping from the end (to catch the producing GEMM's tail in L2) measured no gain: 0.270 -> 0.265 ms / 8 launches
  DISPATCH_VPL(E, (layernorm_fwd_kernel<VPL><<<grid, ROWS_PER_BLOCK * 32, 0, st>>>(
                      reinterpret_cast<const __nv_bfloat16*>(x_bf16), rows, gamma, beta, eps,
                      reinterpret_cast<__nv_bfloat16*>(y_bf16), reverse)));
  TNR_LAUNCH_CHECK();
  return 0;
}

extern "C" __attribute__((visibility("default"))) int tnr_layernorm_bwd(const void* dy_bf16, const void* x_bf16, int rows, int E, const float* gamma, float eps,
                                 void* dx_bf16, float* dgamma, float* dbeta, void* dx_drop_bf16, float* dsum,
                                 const tnr_dropout* drop, void* stream) {
  if (rows == 0) return 0;
  int grid = (rows + LNB_WARPS - 1) / LNB_WARPS;
  const int cap = num_sms();
  if (grid > cap) grid = cap;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int smem = 3 * E * 4 + LNB_WARPS * LNB_STAGES * 2 * E * 2;
  DISPATCH_VPL(E, (cudaFuncSetAttribute(layernorm_bwd_kernel<VPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)));
  DISPATCH_VPL(E, (layernorm_bwd_kernel<VPL><<<grid, LNB_WARPS * 32, smem, st>>>(
                      reinterpret_cast<const __nv_bfloat16*>(dy_bf16), reinterpret_cast<const __nv_bfloat16*>(x_bf16),
                      rows, gamma, eps, reinterpret_cast<__nv_bfloat16*>(dx_bf16), dgamma, dbeta,
                      reinterpret_cast<__nv_bfloat16*>(dx_drop_bf16), dsum, drop_or_none(drop))));
  TNR_LAUNCH_CHECK();
  return 0;
}

extern "C" __attribute__((visibility("default"))) int tnr_colsum_bf16(const void* x_bf16, int rows, int cols, int ld, float* out, void* stream) {
  TNR_REQUIRE(cols % 8 == 0 && ld % 8 == 0, "tnr_colsum_bf16: cols/ld must be multiples of 8");
  if (rows == 0) return 0;
  dim3 grid((cols + 63) / 64, 1);
  int gy = (num_sms() * 8) / (int)grid.x;
  if (gy < 1) gy = 1;
  const int max_gy = (rows + 31) / 32;
  grid.y = gy < max_gy ? gy : max_gy;
  colsum_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x_bf16), rows, cols, ld, out);
  TNR_LAUNCH_CHECK();
  return 0;
}
