// HBM-bound row kernels of the encoder: embedding gather + LayerNorm, LayerNorm forward /
// backward, column sums (bias gradients).  One warp per row, 16-byte vector loads, the row
// lives in registers (E <= 1024), warp-shuffle reductions.
#include "common.cuh"
#include <cstdlib>

namespace tnr {

constexpr int ROWS_PER_BLOCK = 8;   // 8 warps
constexpr int LNF_ROWS = 3;         // rows in flight per warp of the LayerNorm forward: 1 -> 69.6 %, 2 -> 78.9 %, 3 -> 80.7 % of HBM peak (128 registers)

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
  const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
  const int w = blockIdx.x * EMB_WARPS + warp;
  const int l = w % L, grp = w / L;
  if (grp >= groups) return;
  const DropCfg dc = load_drop(drop);
  // packed fp32 pairs (FFMA2 / FADD2): pair i of vector v = columns (v*32+lane)*8 + 2i, +1
  f32x2 g[VPL * 4], bt[VPL * 4], pt[VPL * 4];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int col = (v * 32 + lane) * 8;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 gg = *reinterpret_cast<const float4*>(gamma + col + 4 * h);
      const float4 bb = *reinterpret_cast<const float4*>(beta + col + 4 * h);
      const float4 pp = *reinterpret_cast<const float4*>(pos + (size_t)l * E + col + 4 * h);
      const float4 tt = *reinterpret_cast<const float4*>(type0 + col + 4 * h);
      g[v * 4 + 2 * h] = pk2(gg.x, gg.y); g[v * 4 + 2 * h + 1] = pk2(gg.z, gg.w);
      bt[v * 4 + 2 * h] = pk2(bb.x, bb.y); bt[v * 4 + 2 * h + 1] = pk2(bb.z, bb.w);
      // batch-invariant part folded once: w + (p + t) vs the reference's (w + p) + t differ by an fp32
      // ulp at most, far below the bf16 output resolution
      pt[v * 4 + 2 * h] = pk2(pp.x + tt.x, pp.y + tt.y); pt[v * 4 + 2 * h + 1] = pk2(pp.z + tt.z, pp.w + tt.w);
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
      f32x2 x[VPL * 4];
      f32x2 s2 = pk2(0.f, 0.f);
#pragma unroll
      for (int v = 0; v < VPL; ++v)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          f32x2 w2;
          if (WORD_BF16) {
            const uint32_t u = (k ? raw1[v] : raw0[v]).u[i];
            w2 = pk2(bf16_lo(u), bf16_hi(u));
          } else {
            const bf16x8& rr = (k ? raw1 : raw0)[2 * v + (i >> 1)];
            w2 = pk2(__uint_as_float(rr.u[(i & 1) * 2]), __uint_as_float(rr.u[(i & 1) * 2 + 1]));
          }
          x[v * 4 + i] = add2(w2, pt[v * 4 + i]);
          s2 = add2(s2, x[v * 4 + i]);
        }
      float s0, s1;
      upk2(s2, s0, s1);
      const float mean = warp_sum(s0 + s1) * (1.0f / E);
      const f32x2 nmean2 = pk2(-mean, -mean);
      f32x2 q2 = pk2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < VPL * 4; ++i) { const f32x2 d = add2(x[i], nmean2); q2 = fma2(d, d, q2); }
      upk2(q2, s0, s1);
      const float rstd = rsqrtf(warp_sum(s0 + s1) * (1.0f / E) + eps);
      const f32x2 rstd2 = pk2(rstd, rstd), nmr2 = pk2(-mean * rstd, -mean * rstd);
      const size_t t = (size_t)n * L + l;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int col = (v * 32 + lane) * 8;
        f32x2 yv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) yv[i] = fma2(fma2(x[v * 4 + i], rstd2, nmr2), g[v * 4 + i], bt[v * 4 + i]);
        if (dc.thr16 != 0) {
          f32x2 m[4];
          dropout_mul8(dc, ((uint64_t)t * (uint64_t)E + (uint64_t)col) >> 3, m);
#pragma unroll
          for (int i = 0; i < 4; ++i) yv[i] = mul2(yv[i], m[i]);
        }
        bf16x8 o;
#pragma unroll
        for (int i = 0; i < 4; ++i) { float lo, hi; upk2(yv[i], lo, hi); o.u[i] = pack_bf16(lo, hi); }
        *reinterpret_cast<bf16x8*>(out + t * E + col) = o;
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
  constexpr int NR = LNF_ROWS;                    // rows in flight per warp (registers)
  const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
  // packed fp32 pairs (FFMA2 / FADD2): pair i of vector v = columns (v*32+lane)*8 + 2i, +1
  f32x2 g[VPL * 4], bt[VPL * 4];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int col = (v * 32 + lane) * 8;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 gg = *reinterpret_cast<const float4*>(gamma + col + 4 * h);
      const float4 bb = *reinterpret_cast<const float4*>(beta + col + 4 * h);
      g[v * 4 + 2 * h] = pk2(gg.x, gg.y); g[v * 4 + 2 * h + 1] = pk2(gg.z, gg.w);
      bt[v * 4 + 2 * h] = pk2(bb.x, bb.y); bt[v * 4 + 2 * h + 1] = pk2(bb.z, bb.w);
    }
  }
  const int stride = gridDim.x * ROWS_PER_BLOCK;
  for (int r0 = blockIdx.x * ROWS_PER_BLOCK + warp; r0 < rows; r0 += NR * stride) {
    // reverse: sweep the rows from the END -- the rows the producing GEMM wrote last are the ones still in L2
    bf16x8 raw[NR][VPL];
#pragma unroll
    for (int k = 0; k < NR; ++k) {
      const int rk = r0 + k * stride;
      if (rk < rows) {
        const int a = reverse ? rows - 1 - rk : rk;
#pragma unroll
        for (int v = 0; v < VPL; ++v) raw[k][v] = *reinterpret_cast<const bf16x8*>(x + (size_t)a * E + (v * 32 + lane) * 8);
      }
    }
#pragma unroll
    for (int k = 0; k < NR; ++k) {
      const int rk = r0 + k * stride;
      if (rk >= rows) break;
      const int r = reverse ? rows - 1 - rk : rk;
      f32x2 xv[VPL * 4];
      f32x2 s2 = pk2(0.f, 0.f);
#pragma unroll
      for (int v = 0; v < VPL; ++v)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          xv[v * 4 + i] = pk2(bf16_lo(raw[k][v].u[i]), bf16_hi(raw[k][v].u[i]));
          s2 = add2(s2, xv[v * 4 + i]);
        }
      float s0, s1;
      upk2(s2, s0, s1);
      const float mean = warp_sum(s0 + s1) * (1.0f / E);
      const f32x2 nmean2 = pk2(-mean, -mean);
      f32x2 q2 = pk2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < VPL * 4; ++i) { const f32x2 d = add2(xv[i], nmean2); q2 = fma2(d, d, q2); }
      upk2(q2, s0, s1);
      const float rstd = rsqrtf(warp_sum(s0 + s1) * (1.0f / E) + eps);
      const f32x2 rstd2 = pk2(rstd, rstd), nmr2 = pk2(-mean * rstd, -mean * rstd);
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        bf16x8 o;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float lo, hi;
          upk2(fma2(fma2(xv[v * 4 + i], rstd2, nmr2), g[v * 4 + i], bt[v * 4 + i]), lo, hi);
          o.u[i] = pack_bf16(lo, hi);
        }
        *reinterpret_cast<bf16x8*>(y + (size_t)r * E + (v * 32 + lane) * 8) = o;
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
constexpr int LNB_STAGES = 3;       // ring depth 2 / 3 / 4 measure the same (the kernel is not latency-bound); 3 also sizes the tail's slots

template <int VPL>
__global__ void __launch_bounds__(LNB_WARPS * 32, 1)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x, int rows,
                     const float* __restrict__ gamma, float eps, __nv_bfloat16* __restrict__ dx,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, __nv_bfloat16* __restrict__ dx_drop,
                     float* __restrict__ dsum, const tnr_dropout drop) {
  constexpr int E = VPL * 256;
  extern __shared__ __align__(16) uint8_t ln_smem[];
  const DropCfg dc = load_drop(drop);
  const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
  // per warp: LNB_STAGES x {x row, dy row} bf16
  __nv_bfloat16* stage = reinterpret_cast<__nv_bfloat16*>(ln_smem) + (size_t)warp * LNB_STAGES * 2 * E;
  // packed fp32 pairs (FFMA2 / FADD2 / FMUL2, common.cuh): element pair i of vector v is columns (v*32+lane)*8 + 2i, +1
  f32x2 gm[VPL * 4], acc_dg[VPL * 4], acc_db[VPL * 4], acc_ds[VPL * 4];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int col = (v * 32 + lane) * 8;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + col));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + col + 4));
    gm[v * 4 + 0] = pk2(g0.x, g0.y); gm[v * 4 + 1] = pk2(g0.z, g0.w);
    gm[v * 4 + 2] = pk2(g1.x, g1.y); gm[v * 4 + 3] = pk2(g1.z, g1.w);
  }
#pragma unroll
  for (int i = 0; i < VPL * 4; ++i) { acc_dg[i] = pk2(0.f, 0.f); acc_db[i] = pk2(0.f, 0.f); acc_ds[i] = pk2(0.f, 0.f); }
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
#pragma unroll
  for (int i = 0; i < LNB_STAGES - 1; ++i) prefetch(i, r + i * stride);
  int st = 0;
  for (; r < rows; r += stride) {
    int st2 = st + LNB_STAGES - 1; if (st2 >= LNB_STAGES) st2 -= LNB_STAGES;
    prefetch(st2, r + (LNB_STAGES - 1) * stride);
    asm volatile("cp.async.wait_group %0;" ::"n"(LNB_STAGES - 1) : "memory");
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
    // mean and variance from one sweep: sum and sum of squares of the row SHIFTED by its first element (the shift removes
    // the cancellation a large mean would cause; bf16 inputs, fp32 sums, E <= 1024), the two warp reductions interleaved
    const float shift = __shfl_sync(0xffffffffu, bf16_lo(rx[0].u[0]), 0);
    const f32x2 nshift2 = pk2(-shift, -shift);
    f32x2 s2 = pk2(0.f, 0.f), q2 = pk2(0.f, 0.f);
#pragma unroll
    for (int v = 0; v < VPL; ++v)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const f32x2 d = add2(pk2(bf16_lo(rx[v].u[i]), bf16_hi(rx[v].u[i])), nshift2);
        s2 = add2(s2, d);
        q2 = fma2(d, d, q2);
      }
    float s, q;
    { float a0, a1; upk2(s2, a0, a1); s = a0 + a1; upk2(q2, a0, a1); q = a0 + a1; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    const float ms = s * (1.0f / E);
    const float mean = shift + ms;
    const float rstd = rsqrtf(fmaxf(q * (1.0f / E) - ms * ms, 0.f) + eps);
    // xhat = x * rstd - mean * rstd (one FFMA2 per pair)
    const f32x2 rstd2 = pk2(rstd, rstd), nmr2 = pk2(-mean * rstd, -mean * rstd);
    f32x2 sg2 = pk2(0.f, 0.f), sgx2 = pk2(0.f, 0.f);
#pragma unroll
    for (int v = 0; v < VPL; ++v)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const f32x2 xh = fma2(pk2(bf16_lo(rx[v].u[i]), bf16_hi(rx[v].u[i])), rstd2, nmr2);
        const f32x2 d = pk2(bf16_lo(rd[v].u[i]), bf16_hi(rd[v].u[i]));
        acc_dg[v * 4 + i] = fma2(d, xh, acc_dg[v * 4 + i]);
        acc_db[v * 4 + i] = add2(acc_db[v * 4 + i], d);
        const f32x2 gg = mul2(d, gm[v * 4 + i]);
        sg2 = add2(sg2, gg);
        sgx2 = fma2(gg, xh, sgx2);
      }
    float sg, sgx;
    { float a0, a1; upk2(sg2, a0, a1); sg = a0 + a1; upk2(sgx2, a0, a1); sgx = a0 + a1; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sg += __shfl_xor_sync(0xffffffffu, sg, o);
      sgx += __shfl_xor_sync(0xffffffffu, sgx, o);
    }
    sg *= (1.0f / E);
    sgx *= (1.0f / E);
    // dx = rstd * g - rstd * mean(g) - xhat * rstd * mean(g xhat): two FFMA2 per pair on top of g and xhat
    const f32x2 c0 = pk2(-rstd * sg, -rstd * sg), c1 = pk2(-rstd * sgx, -rstd * sgx);
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int col = (v * 32 + lane) * 8;
      f32x2 o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const f32x2 xh = fma2(pk2(bf16_lo(rx[v].u[i]), bf16_hi(rx[v].u[i])), rstd2, nmr2);
        const f32x2 gg = mul2(pk2(bf16_lo(rd[v].u[i]), bf16_hi(rd[v].u[i])), gm[v * 4 + i]);
        o[i] = fma2(xh, c1, fma2(gg, rstd2, c0));
      }
      bf16x8 ob;
#pragma unroll
      for (int i = 0; i < 4; ++i) { float lo, hi; upk2(o[i], lo, hi); ob.u[i] = pack_bf16(lo, hi); }
      *reinterpret_cast<bf16x8*>(dx + (size_t)r * E + col) = ob;
      if (dc.thr16 != 0) {
        f32x2 m[4];
        dropout_mul8(dc, ((uint64_t)r * (uint64_t)E + (uint64_t)col) >> 3, m);
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = mul2(o[i], m[i]);
#pragma unroll
        for (int i = 0; i < 4; ++i) { float lo, hi; upk2(o[i], lo, hi); ob.u[i] = pack_bf16(lo, hi); }
      }
      if (dx_drop != nullptr) *reinterpret_cast<bf16x8*>(dx_drop + (size_t)r * E + col) = ob;
#pragma unroll
      for (int i = 0; i < 4; ++i) acc_ds[v * 4 + i] = add2(acc_ds[v * 4 + i], o[i]);
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();
  // block totals without shared-memory float atomics (CAS loops: 10 instructions each, 72 per lane, contended by 12
  // warps): every warp parks its three accumulator rows in its own -- now idle -- ring (3 stages x 2 rows x 2 B = 3 x 4 B
  // per column), then a thread adds the 12 partials of its columns in warp order and issues one global RED per total
  static_assert(LNB_STAGES >= 3, "the ring doubles as the per-warp accumulator slot");
  {
    float* mine = reinterpret_cast<float*>(stage);
#pragma unroll
    for (int v = 0; v < VPL; ++v)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int col = (v * 32 + lane) * 8 + 2 * i;
        float lo, hi;
        upk2(acc_dg[v * 4 + i], lo, hi);
        *reinterpret_cast<float2*>(mine + col) = make_float2(lo, hi);
        upk2(acc_db[v * 4 + i], lo, hi);
        *reinterpret_cast<float2*>(mine + E + col) = make_float2(lo, hi);
        upk2(acc_ds[v * 4 + i], lo, hi);
        *reinterpret_cast<float2*>(mine + 2 * E + col) = make_float2(lo, hi);
      }
  }
  __syncthreads();
  const float* part0 = reinterpret_cast<const float*>(ln_smem);
  constexpr int PART_STRIDE = LNB_STAGES * 2 * E * 2 / 4;          // floats between two warps' slots
  for (int i = threadIdx.x; i < 3 * E; i += blockDim.x) {
    if (i >= 2 * E && dsum == nullptr) break;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < LNB_WARPS; ++w) t += part0[w * PART_STRIDE + i];
    float* dst = i < E ? dgamma + i : (i < 2 * E ? dbeta + (i - E) : dsum + (i - 2 * E));
    atomicAdd(dst, t);
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
  const int smem = LNB_WARPS * LNB_STAGES * 2 * E * 2;
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
