// bf16 GEMM on 5th-gen tensor cores (tcgen05.mma, fp32 accumulators in TMEM), operands
// staged by TMA into 128B-swizzled shared memory, warp-specialised persistent kernel:
//   warp 0      TMA producer   } the whole warp walks the loop and waits on the mbarriers converged, one elect.sync
//   warp 1      MMA issuer     } lane issues (operands in uniform registers); tcgen05.mma cta_group::1, 128 x BN x 16
//   warp 2      TMEM allocator / deallocator
//   warps 4-19  epilogue: tcgen05.ld -> bias / GELU / tanh / dGELU / dropout / residual
//               bf16 outputs: each warp stages 32 x 64 boxes in 128B-swizzled shared memory and
//               writes them with TMA stores; the residual (or the dGELU pre-activation) box is
//               prefetched by TMA at tile start.  (Per-thread-row 16 B global accesses made the
//               epilogue L1-tag bound: 32 lines per request, ncu l1tex 68 % / tensor 18 %.)
//               fp32 / split-K outputs: direct stores / fp32 atomics.
// Two TMEM accumulator stages (2 x BN columns) let the epilogue of tile i overlap the
// main loop of tile i+1.  Split-K work items accumulate with fp32 atomics (wgrad).
//
// CTA2 variant (large-M forward / dgrad GEMMs): the two CTAs of a cluster (one TPC) work on one
// 256 x BN tile with tcgen05.mma.cta_group::2.  Each CTA TMA-loads its own 128 rows of A and HALF of the
// B tile (BN/2 rows) -- 32 KB instead of 48 KB of L2->SM operand traffic per CTA and k-block -- and both
// loads signal the LEADER's "full" barrier; the leader's elected thread issues the 256-row MMAs and
// multicasts the "stage free" / "accumulator ready" commits to both CTAs; each CTA's epilogue warps
// drain their own 128 TMEM lanes and release the accumulator on the leader's barrier.
//
// Replaces every nn.Linear on the path (reference call sites listed in include/tinyrec.h).
#include <cuda.h>
#include <cstdlib>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace tnr {
namespace gemm {

constexpr int BM = 128;
constexpr int BK = 64;          // 64 bf16 = 128 B = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 16;     // 4 per SM sub-partition: 4 TMEM lane quarters x 4 column slices of the tile
constexpr int NUM_THREADS = 128 + NUM_EPI_WARPS * 32;

struct Params {
  int M, N, K;
  int m_tiles, n_tiles, k_blocks, k_per_split, splits;
  void* C; int ldc; int c_f32;
  const float* bias;
  const __nv_bfloat16* residual; int ldr;
  int act;
  __nv_bfloat16* aux; int ldaux;
  int atomic;
  tnr_dropout drop;
  int epi_tma;        // bf16 C through smem + TMA store
  int in_mode;        // 0 none, 1 residual box prefetched by TMA, 2 dGELU pre-activation box
  int aux_out;        // GELU pre-activation written through TMA
  float* colsum;      // column sums of the bf16 output (bias gradient), fp32 atomics; NULL = off
};

// Shared-memory matrix descriptor (sm_100 format, version 1, SWIZZLE_128B).
//   K-major : rows of 128 B; 8-row groups 1024 B apart (SBO); LBO unused.
//   MN-major: atoms of (64 MN x 8 K) = 1024 B; next 8 K-rows +1024 B (SBO);
//             next 64 MN elements = next TMA box, +8192 B (LBO).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, bool mn_major) {
  const uint64_t sbo = 1024 >> 4;
  const uint64_t lbo = mn_major ? (uint64_t)((BK * 128) >> 4) : 0;
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= lbo << 16;
  d |= sbo << 32;
  d |= (uint64_t)1 << 46;     // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;     // SWIZZLE_128B
  return d;
}

template <int BN, bool A_MN, bool B_MN, bool CTA2 = false>
__host__ __device__ constexpr uint32_t make_idesc() {
  return (1u << 4)                      // D format f32
         | (1u << 7)                    // A bf16
         | (1u << 10)                   // B bf16
         | ((A_MN ? 1u : 0u) << 15)     // A major
         | ((B_MN ? 1u : 0u) << 16)     // B major
         | ((uint32_t)(BN >> 3) << 17)  // N
         | ((uint32_t)((CTA2 ? 2 * BM : BM) >> 4) << 24); // M (256 across the CTA pair)
}

template <int BN, bool EPI_TMA, bool CTA2 = false, bool LEAN = false>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_ROWS = CTA2 ? BN / 2 : BN;    // B rows (of N) staged by one CTA
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // CTA pair, BN = 256, bf16 output: 5 stages; the two-output GELU + GELU' epilogue (LEAN) runs 3.5 % faster with 4 (same-box
  // A/B in profiles/r02_gemm_issue_loop.txt); every other epilogue is 1-2 % slower
  static constexpr int STAGES = CTA2 ? ((BN == 256) ? (EPI_TMA ? (LEAN ? 4 : 5) : 6) : (EPI_TMA ? 6 : 8))
                                     : ((BN == 256) ? (EPI_TMA ? 3 : 4) : (EPI_TMA ? 4 : 6));
  static constexpr int TMEM_COLS = 2 * BN;             // 2 accumulator stages (power of 2 >= 32)
  static constexpr int EPI_BOX_BYTES = 32 * 128;       // 32 rows x 64 bf16, SWIZZLE_128B
  static constexpr int EPI_BYTES = EPI_TMA ? NUM_EPI_WARPS * EPI_BOX_BYTES : 0;      // one staging box per epilogue warp
  static constexpr int NUM_BARS = 2 * STAGES + 4 + NUM_EPI_WARPS;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align slack*/ + 8 * NUM_BARS + 16;
};

// ---------------------------------------------------------------- epilogue math
template <int ACT>
__device__ __forceinline__ void epilogue_chunk(const Params& p, const DropCfg& dc, const uint32_t* acc, int row, int col0) {
  // 32 consecutive columns [col0, col0+32) of one output row
  const bool row_ok = row < p.M;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int col = col0 + g * 8;
    if (!row_ok || col >= p.N) continue;      // N % 8 == 0 is required by the host wrapper
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(acc[g * 8 + i]);
    if (p.bias != nullptr) {
      const float4 b0 = *reinterpret_cast<const float4*>(p.bias + col);
      const float4 b1 = *reinterpret_cast<const float4*>(p.bias + col + 4);
      v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
      v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    }
    if (ACT == TNR_ACT_GELU) {
      if (p.aux != nullptr)
        *reinterpret_cast<bf16x8*>(p.aux + (size_t)row * p.ldaux + col) = pack8(v);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = gelu_erf(v[i]);
    } else if (ACT == TNR_ACT_TANH) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = tanhf(v[i]);
    } else if (ACT == TNR_ACT_DGELU) {
      float z[8];
      unpack8(*reinterpret_cast<const bf16x8*>(p.aux + (size_t)row * p.ldaux + col), z);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] *= gelu_erf_grad(z[i]);
    } else if (ACT == TNR_ACT_GELU_DAUX) {
      if (p.aux != nullptr) {
        float d[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = gelu_erf_grad(v[i]);
        *reinterpret_cast<bf16x8*>(p.aux + (size_t)row * p.ldaux + col) = pack8(d);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = gelu_erf(v[i]);
    } else if (ACT == TNR_ACT_MULAUX) {
      float z[8];
      unpack8(*reinterpret_cast<const bf16x8*>(p.aux + (size_t)row * p.ldaux + col), z);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] *= z[i];
    }
    if (dc.thr16 != 0) {       // dropout on (dense + bias), before the residual (BertSelfOutput / BertOutput)
      const uint32_t keep = dropout_keep8(dc, ((uint64_t)row * (uint64_t)p.N + (uint64_t)col) >> 3);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = ((keep >> i) & 1u) ? v[i] * dc.scale : 0.f;
    }
    if (p.residual != nullptr) {
      float r[8];
      unpack8(*reinterpret_cast<const bf16x8*>(p.residual + (size_t)row * p.ldr + col), r);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += r[i];
    }
    if (p.c_f32) {
      float* c = reinterpret_cast<float*>(p.C) + (size_t)row * p.ldc + col;
      if (p.atomic) {
        atomicAdd(reinterpret_cast<float4*>(c), make_float4(v[0], v[1], v[2], v[3]));
        atomicAdd(reinterpret_cast<float4*>(c + 4), make_float4(v[4], v[5], v[6], v[7]));
      } else {
        *reinterpret_cast<float4*>(c) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(c + 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
    } else {
      __nv_bfloat16* c = reinterpret_cast<__nv_bfloat16*>(p.C) + (size_t)row * p.ldc + col;
      *reinterpret_cast<bf16x8*>(c) = pack8(v);
    }
  }
}

// bf16 output path: 8 consecutive columns of one row, staged in 128B-swizzled shared memory
// (box row = lane, 16-byte chunk index c) for the TMA store; `in` holds the residual / dGELU box.
// z = acc + bias as bf16 into the staging box (GELU pre-activation kept for the backward)
__device__ __forceinline__ void epilogue_preact8(const Params& p, const uint32_t* acc, int col, uint8_t* box, uint32_t swz) {
  f32x2 v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = pk2(__uint_as_float(acc[2 * i]), __uint_as_float(acc[2 * i + 1]));
  if (p.bias != nullptr) {
    const float* bp = p.bias + min(col, p.N - 8);
    const float4 b0 = *reinterpret_cast<const float4*>(bp);
    const float4 b1 = *reinterpret_cast<const float4*>(bp + 4);
    v[0] = add2(v[0], pk2(b0.x, b0.y)); v[1] = add2(v[1], pk2(b0.z, b0.w));
    v[2] = add2(v[2], pk2(b1.x, b1.y)); v[3] = add2(v[3], pk2(b1.z, b1.w));
  }
  bf16x8 o;
#pragma unroll
  for (int i = 0; i < 4; ++i) { float lo, hi; upk2(v[i], lo, hi); o.u[i] = pack_bf16(lo, hi); }
  *reinterpret_cast<bf16x8*>(box + swz) = o;
}

// GELU_DAUX, first pass: z = acc + bias; gelu'(z) as bf16 into the staging box (the backward only ever needs the
// derivative: storing it instead of z turns the dGELU dgrad epilogue into one multiply), gelu(z) as bf16 back into
// acc[0..3] for the second pass.  Both come from ONE sigmoid evaluation.
__device__ __forceinline__ void epilogue_gelu_both8(const Params& p, uint32_t* acc, int col, uint8_t* box, uint32_t swz) {
  f32x2 v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = pk2(__uint_as_float(acc[2 * i]), __uint_as_float(acc[2 * i + 1]));
  if (p.bias != nullptr) {
    const float* bp = p.bias + min(col, p.N - 8);
    const float4 b0 = *reinterpret_cast<const float4*>(bp);
    const float4 b1 = *reinterpret_cast<const float4*>(bp + 4);
    v[0] = add2(v[0], pk2(b0.x, b0.y)); v[1] = add2(v[1], pk2(b0.z, b0.w));
    v[2] = add2(v[2], pk2(b1.x, b1.y)); v[3] = add2(v[3], pk2(b1.z, b1.w));
  }
  bf16x8 d;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float s0, s1;
    gelu_sig2(v[i], s0, s1);
    const f32x2 sg = pk2(s0, s1);
    float u0, u1;
    upk2(mul2(v[i], v[i]), u0, u1);
    const f32x2 u = pk2(fminf(u0, 64.0f), fminf(u1, 64.0f));
    const f32x2 qp2 = fma2(fma2(pk2(-0.003515171750f, -0.003515171750f), u, pk2(0.2220338916f, 0.2220338916f)), u,
                           pk2(1.595015762f, 1.595015762f));
    const f32x2 ss = fma2(mul2(sg, pk2(-1.0f, -1.0f)), sg, sg);
    float lo, hi;
    upk2(fma2(mul2(ss, qp2), v[i], sg), lo, hi);            // gelu'(z), same form as gelu_erf_grad2()
    d.u[i] = pack_bf16(lo, hi);
    upk2(mul2(v[i], sg), lo, hi);                          // gelu(z)
    acc[i] = pack_bf16(lo, hi);
  }
  *reinterpret_cast<bf16x8*>(box + swz) = d;
}

template <int ACT>
__device__ __forceinline__ void epilogue_staged8(const Params& p, const DropCfg& dc, const uint32_t* acc, uint64_t drop_group,
                                                 int col, uint8_t* out_box, uint8_t* in_box, uint32_t swz) {
  // four packed fp32 pairs (FFMA2 path, common.cuh)
  f32x2 v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = pk2(__uint_as_float(acc[2 * i]), __uint_as_float(acc[2 * i + 1]));
  if (p.bias != nullptr) {
    // N % 8 == 0 (host check): a live chunk is entirely inside the row; dead columns re-read the last chunk
    const float* bp = p.bias + min(col, p.N - 8);
    const float4 b0 = *reinterpret_cast<const float4*>(bp);
    const float4 b1 = *reinterpret_cast<const float4*>(bp + 4);
    v[0] = add2(v[0], pk2(b0.x, b0.y)); v[1] = add2(v[1], pk2(b0.z, b0.w));
    v[2] = add2(v[2], pk2(b1.x, b1.y)); v[3] = add2(v[3], pk2(b1.z, b1.w));
  }
  if (ACT == TNR_ACT_GELU) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = gelu_erf2(v[i]);
  } else if (ACT == TNR_ACT_TANH) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { float lo, hi; upk2(v[i], lo, hi); v[i] = pk2(tanhf(lo), tanhf(hi)); }
  } else if (ACT == TNR_ACT_DGELU) {
    const bf16x8 zr = *reinterpret_cast<const bf16x8*>(in_box + swz);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = mul2(v[i], gelu_erf_grad2(pk2(bf16_lo(zr.u[i]), bf16_hi(zr.u[i]))));
  } else if (ACT == TNR_ACT_MULAUX) {
    const bf16x8 zr = *reinterpret_cast<const bf16x8*>(in_box + swz);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = mul2(v[i], pk2(bf16_lo(zr.u[i]), bf16_hi(zr.u[i])));
  } else if (ACT == TNR_ACT_GELU_DAUX) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = gelu_erf2(v[i]);       // only reached without an aux output
  }
  const bool has_res = ACT != TNR_ACT_DGELU && ACT != TNR_ACT_MULAUX && p.in_mode == 1;
  if (dc.thr16 != 0) {
    f32x2 m[4];
    dropout_mul8(dc, drop_group, m);
    if (has_res) {                                            // residual + dropout(v): one FFMA2 per pair
      const bf16x8 rr = *reinterpret_cast<const bf16x8*>(in_box + swz);
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = fma2(v[i], m[i], pk2(bf16_lo(rr.u[i]), bf16_hi(rr.u[i])));
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = mul2(v[i], m[i]);
    }
  } else if (has_res) {
    const bf16x8 rr = *reinterpret_cast<const bf16x8*>(in_box + swz);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = add2(v[i], pk2(bf16_lo(rr.u[i]), bf16_hi(rr.u[i])));
  }
  bf16x8 o;
#pragma unroll
  for (int i = 0; i < 4; ++i) { float lo, hi; upk2(v[i], lo, hi); o.u[i] = pack_bf16(lo, hi); }
  *reinterpret_cast<bf16x8*>(out_box + swz) = o;
}

#ifdef GEMM_TIMING
__device__ long long g_gemm_dbg[8192];       // debug build: clock64 stamps of CTA 0 (tools/gemm_timing.py)
#define GEMM_STAMP(slot) do { if (blockIdx.x == 0 && (slot) < 8192) g_gemm_dbg[(slot)] = clock64(); } while (0)
#else
#define GEMM_STAMP(slot) do { } while (0)
#endif

// ---------------------------------------------------------------- kernel
template <int BN, bool A_MN, bool B_MN, int ACT, bool EPI_TMA, bool CTA2>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_in,
            const __grid_constant__ CUtensorMap tmap_aux, const Params p) {
  using C = Cfg<BN, EPI_TMA, CTA2, CTA2 && ACT == TNR_ACT_GELU_DAUX>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // [pipeline stages][epilogue staging boxes][barriers]
  const uint32_t epi_base = smem_base + C::STAGES * C::STAGE_BYTES;
  const uint32_t bar_base = epi_base + C::EPI_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + 2 + s); };
  auto ld_bar = [&](int w) { return bar_base + 8u * (2 * C::STAGES + 4 + w); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + C::STAGES * C::STAGE_BYTES + C::EPI_BYTES + 8 * C::NUM_BARS);

  const int warp = uniform_warp_id();
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = CTA2 ? cluster_ctarank() : 0u;     // 0 = leader of the CTA pair

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (EPI_TMA) tma_prefetch_desc(&tmap_c);
    if (p.in_mode) tma_prefetch_desc(&tmap_in);
    if (p.aux_out) tma_prefetch_desc(&tmap_aux);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), (CTA2 ? 2 : 1) * NUM_EPI_WARPS); }
    for (int w = 0; w < NUM_EPI_WARPS; ++w) mbar_init(ld_bar(w), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if (CTA2) {   // pair allocation: the same warp of both CTAs, same columns in both SMs
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "n"(C::TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "n"(C::TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (CTA2) cluster_sync_all(); else __syncthreads();     // the peer's barriers must be initialised before remote arrives
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work items: (n_blk fastest, m_unit, k split); an m_unit is one 128-row tile, or a 256-row pair tile (CTA2)
  const int m_units = CTA2 ? (p.m_tiles + 1) / 2 : p.m_tiles;
  const int total_tiles = m_units * p.n_tiles * p.splits;
  const int t_first = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int t_stride = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0) {
    // ===================== TMA producer =====================
    {                                             // whole warp walks the loop, one elected lane issues (see the MMA warp)
      int stage = 0; uint32_t phase = 0;
      for (int t = t_first; t < total_tiles; t += t_stride) {
        const int n_blk = t % p.n_tiles;
        const int m_unit = (t / p.n_tiles) % m_units;
        const int m_blk = CTA2 ? 2 * m_unit + (int)cta_rank : m_unit;
        const int ks = t / (p.n_tiles * m_units);
        const int kb0 = ks * p.k_per_split;
        const int kb1 = min(p.k_blocks, kb0 + p.k_per_split);
        const int n0 = n_blk * BN + (CTA2 ? (int)cta_rank * (BN / 2) : 0);   // this CTA's rows of the B tile
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
          const uint32_t sb = sa + C::A_BYTES;
          if (!elect_one()) {
          } else if (CTA2) {
            // both CTAs' bytes land on the leader's barrier; only the leader arrives on it
            if (cta_rank == 0) mbar_expect_tx(full_bar(stage), 2 * C::STAGE_BYTES);
            const uint32_t fb = mapa_cluster(full_bar(stage), 0);
            if (!A_MN) {
              tma_load_2d_cta2(sa, &tmap_a, fb, kb * BK, m_blk * BM);
            } else {
#pragma unroll
              for (int i = 0; i < BM / 64; ++i)
                tma_load_2d_cta2(sa + i * (BK * 128), &tmap_a, fb, m_blk * BM + i * 64, kb * BK);
            }
            if (!B_MN) {
              tma_load_2d_cta2(sb, &tmap_b, fb, kb * BK, n0);
            } else {
#pragma unroll
              for (int i = 0; i < C::B_ROWS / 64; ++i)
                tma_load_2d_cta2(sb + i * (BK * 128), &tmap_b, fb, n0 + i * 64, kb * BK);
            }
          } else {
            mbar_expect_tx(full_bar(stage), C::STAGE_BYTES);
            if (!A_MN) {
              tma_load_2d(sa, &tmap_a, full_bar(stage), kb * BK, m_blk * BM);
            } else {
#pragma unroll
              for (int i = 0; i < BM / 64; ++i)
                tma_load_2d(sa + i * (BK * 128), &tmap_a, full_bar(stage), m_blk * BM + i * 64, kb * BK);
            }
            if (!B_MN) {
              tma_load_2d(sb, &tmap_b, full_bar(stage), kb * BK, n0);
            } else {
#pragma unroll
              for (int i = 0; i < BN / 64; ++i)
                tma_load_2d(sb + i * (BK * 128), &tmap_b, full_bar(stage), n0 + i * 64, kb * BK);
            }
          }
          __syncwarp();
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp walks the loop (barrier waits included) and one elected lane issues: with warp-uniform
    // control flow the descriptors, the TMEM address and the barrier addresses live in uniform registers, so a
    // tcgen05.mma costs one UTCHMMA instead of an ELECT + R2UR broadcast chain per operand (measured with
    // -DGEMM_TIMING: ~75 cycles per issue before, the issue thread -- not the tensor pipe -- paced a k-block).
    if (cta_rank == 0) {                          // CTA2: the leader issues for the pair
      constexpr uint32_t idesc = make_idesc<BN, A_MN, B_MN, CTA2>();
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int t = t_first; t < total_tiles; t += t_stride) {
        const int ks = t / (p.n_tiles * m_units);
        const int kb0 = ks * p.k_per_split;
        const int kb1 = min(p.k_blocks, kb0 + p.k_per_split);
#ifdef GEMM_TIMING
        const int tl_dbg = (t - t_first) / t_stride;
#endif
        if (lane == 0) GEMM_STAMP(tl_dbg * 32 + 0);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        if (lane == 0) GEMM_STAMP(tl_dbg * 32 + 1);
        const uint32_t tmem_d = tmem_u + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (lane == 0 && kb - kb0 < 14) GEMM_STAMP(tl_dbg * 32 + 2 + 2 * (kb - kb0));
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
          const uint32_t sb = sa + C::A_BYTES;
          const uint64_t adesc = make_desc(sa, A_MN);
          const uint64_t bdesc = make_desc(sb, B_MN);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // advance inside the swizzle atom: K-major +32 B per UMMA_K, MN-major +16 rows * 128 B
              const uint64_t aoff = (uint64_t)((A_MN ? k * UMMA_K * 128 : k * UMMA_K * 2) >> 4);
              const uint64_t boff = (uint64_t)((B_MN ? k * UMMA_K * 128 : k * UMMA_K * 2) >> 4);
              if (CTA2) tc_mma_bf16_cta2(tmem_d, adesc + aoff, bdesc + boff, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              else tc_mma_bf16(tmem_d, adesc + aoff, bdesc + boff, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
            // frees the smem stage (in both CTAs of a pair) when these MMAs retire
            if (CTA2) tc_commit_mc2(empty_bar(stage), 3); else tc_commit(empty_bar(stage));
          }
          __syncwarp();
          if (lane == 0 && kb - kb0 < 14) GEMM_STAMP(tl_dbg * 32 + 3 + 2 * (kb - kb0));
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
        // accumulator ready for the epilogue warps (of both CTAs)
        if (elect_one()) {
          if (CTA2) tc_commit_mc2(tfull_bar(acc), 3); else tc_commit(tfull_bar(acc));
        }
        __syncwarp();
        acc ^= 1; if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    // 16 warps: TMEM lane quarter (warp & 3) x 64-column slice of the tile.  Each warp drains its 32 x 64
    // accumulator block into registers as soon as the tile is complete and hands the TMEM stage back right
    // away; the math, the shared-memory staging and the TMA store then overlap the next tile's MMAs, with
    // four warps per scheduler to hide the tcgen05.ld / MUFU / TMA latencies (8 warps were latency-bound:
    // issue slots 40 % busy, K = 768 GEMMs with GELU / dropout epilogues at 700-1050 TFLOP/s).
    const int e = warp - 4;
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int slice = e >> 2;                     // 64-column slice of the tile
    int acc = 0; uint32_t acc_phase = 0;
    const DropCfg dc = load_drop(p.drop);
    // accumulator release goes to the leader's barrier (the MMA issuer waits there for both CTAs)
    auto release_acc = [&](int a) {
      if (CTA2) mbar_arrive_cluster(mapa_cluster(tempty_bar(a), 0));
      else mbar_arrive_relaxed(tempty_bar(a));
    };
    const bool slice_live = slice * 64 < BN;       // BN = 128: half of the warps only keep the barrier counts
    if (EPI_TMA) {
      uint8_t* box = smem_gen + C::STAGES * C::STAGE_BYTES + e * C::EPI_BOX_BYTES;
      const uint32_t box_u32 = epi_base + e * C::EPI_BOX_BYTES;
      const uint32_t swz_row = (uint32_t)lane * 128u;
      uint32_t ld_phase = 0;
      for (int t = t_first; t < total_tiles; t += t_stride) {
        const int n_blk = t % p.n_tiles;
        const int m_unit = (t / p.n_tiles) % m_units;
        const int m_blk = CTA2 ? 2 * m_unit + (int)cta_rank : m_unit;
        const int row0 = m_blk * BM + quarter * 32;
        const int col0 = n_blk * BN + slice * 64;
        const bool live = slice_live && row0 < p.M && col0 < p.N;      // warp-uniform
        if (p.in_mode && live) {
          // the box is about to be overwritten by the prefetch: its previous store must have been read
          if (lane == 0) {
            bulk_wait_read<0>();
            mbar_expect_tx(ld_bar(e), C::EPI_BOX_BYTES);
            tma_load_2d(box_u32, &tmap_in, ld_bar(e), col0, row0);
          }
          __syncwarp();
        }
#ifdef GEMM_TIMING
        const int tle_dbg = (t - t_first) / t_stride;
        if (e == 0 && lane == 0) GEMM_STAMP(4096 + tle_dbg * 8 + 0);
#endif
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
#ifdef GEMM_TIMING
        if (e == 0 && lane == 0) GEMM_STAMP(4096 + tle_dbg * 8 + 1);
#endif
        uint32_t r[64];
        if (slice_live) {
          const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + slice * 64);
          tmem_ld32(taddr, r);
          tmem_ld32(taddr + 32, r + 32);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release_acc(acc);             // accumulator block is in registers
#ifdef GEMM_TIMING
        if (e == 0 && lane == 0) GEMM_STAMP(4096 + tle_dbg * 8 + 2);
#endif
        acc ^= 1; if (acc == 0) acc_phase ^= 1u;
        if (!live) continue;
        if (p.in_mode) {
          mbar_wait(ld_bar(e), ld_phase);
          ld_phase ^= 1u;
        } else {
          if (lane == 0) bulk_wait_read<0>();
          __syncwarp();
        }
        // dropout groups of this thread's row: 8 consecutive columns each, N % 8 == 0 (host check)
        const uint64_t drop_g0 = ((uint64_t)(row0 + lane) * (uint64_t)p.N + (uint64_t)col0) >> 3;
        if ((ACT == TNR_ACT_GELU || ACT == TNR_ACT_GELU_DAUX) && p.aux_out) {
          // pre-activation z = acc + bias (GELU) or gelu'(z) (GELU_DAUX) goes out first through the same box
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint32_t swz = swz_row + (uint32_t)((c ^ (lane & 7)) << 4);
            if (ACT == TNR_ACT_GELU_DAUX) epilogue_gelu_both8(p, r + c * 8, col0 + c * 8, box, swz);
            else epilogue_preact8(p, r + c * 8, col0 + c * 8, box, swz);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmap_aux, box_u32, col0, row0);
            bulk_commit();
            bulk_wait_read<0>();
          }
          __syncwarp();
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t swz = swz_row + (uint32_t)((c ^ (lane & 7)) << 4);
          if (ACT == TNR_ACT_GELU_DAUX && p.aux_out) {          // gelu(z) was packed into r[c*8 .. +3] by the first pass
            bf16x8 o;
#pragma unroll
            for (int i = 0; i < 4; ++i) o.u[i] = r[c * 8 + i];
            *reinterpret_cast<bf16x8*>(box + swz) = o;
          } else {
            epilogue_staged8<ACT>(p, dc, r + c * 8, drop_g0 + (uint64_t)c, col0 + c * 8, box, box, swz);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmap_c, box_u32, col0, row0);
          bulk_commit();
        }
#ifdef GEMM_TIMING
        if (e == 0 && lane == 0) GEMM_STAMP(4096 + tle_dbg * 8 + 3);
#endif
        if (p.colsum != nullptr) {
          // columns 2*lane, 2*lane+1 of the staged 32 x 64 box (128B-swizzled rows); rows past M are skipped
          const int rows_ok = min(32, p.M - row0);
          const uint32_t cw = (uint32_t)(lane & 3) * 4u;
          float a0 = 0.f, a1 = 0.f;
          for (int r = 0; r < rows_ok; ++r) {
            const uint32_t u = *reinterpret_cast<const uint32_t*>(box + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + cw);
            a0 += bf16_lo(u);
            a1 += bf16_hi(u);
          }
          const int c = col0 + 2 * lane;
          if (c < p.N) { atomicAdd(p.colsum + c, a0); atomicAdd(p.colsum + c + 1, a1); }
        }
      }
      if (lane == 0) bulk_wait_all();
    } else {
      for (int t = t_first; t < total_tiles; t += t_stride) {
        const int n_blk = t % p.n_tiles;
        const int m_unit = (t / p.n_tiles) % m_units;
        const int m_blk = CTA2 ? 2 * m_unit + (int)cta_rank : m_unit;
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        const int row = m_blk * BM + quarter * 32 + lane;
        uint32_t r[64];
        if (slice_live) {
          const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + slice * 64);
          tmem_ld32(taddr, r);
          tmem_ld32(taddr + 32, r + 32);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release_acc(acc);
        acc ^= 1; if (acc == 0) acc_phase ^= 1u;
        if (!slice_live) continue;
        epilogue_chunk<ACT>(p, dc, r, row, n_blk * BN + slice * 64);
        epilogue_chunk<ACT>(p, dc, r + 32, row, n_blk * BN + slice * 64 + 32);
      }
    }
  }

  tc_fence_before();
  if (CTA2) cluster_sync_all(); else __syncthreads();   // pair: the leader's MMAs also write the peer's TMEM
  if (warp == 2) {
    if (CTA2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D bf16 tensor map: dim0 (contiguous) = inner, dim1 = outer with row stride ld elements.
static int make_map(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                    uint32_t box_outer) {
  EncodeTiledFn enc = get_encode();
  TNR_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled driver entry point not available");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TNR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (CUresult %d) inner=%llu outer=%llu ld=%llu", (int)r,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld);
  return 0;
}

template <int BN, bool A_MN, bool B_MN, int ACT, bool EPI_TMA, bool CTA2 = false>
static int launch(const CUtensorMap* maps, const Params& p, int grid, cudaStream_t st) {
  using C = Cfg<BN, EPI_TMA, CTA2, CTA2 && ACT == TNR_ACT_GELU_DAUX>;
  auto kern = gemm_kernel<BN, A_MN, B_MN, ACT, EPI_TMA, CTA2>;
  TNR_SET_SMEM(kern, C::SMEM_BYTES);
  if (CTA2) {
    // `grid` counts CTA pairs: launch them as clusters of 2 (the two SMs of one TPC)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    TNR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, maps[0], maps[1], maps[2], maps[3], maps[4], p));
    return 0;
  }
  kern<<<grid, NUM_THREADS, C::SMEM_BYTES, st>>>(maps[0], maps[1], maps[2], maps[3], maps[4], p);
  TNR_LAUNCH_CHECK();
  return 0;
}

// CTA-pair kernels: bf16 outputs, BN = 256, A K-major (forward and dgrad GEMMs of the encoder)
template <bool B_MN>
static int dispatch_act_cta2(const CUtensorMap* maps, const Params& p, int grid, cudaStream_t st) {
  switch (p.act) {
    case TNR_ACT_NONE: return launch<256, false, B_MN, TNR_ACT_NONE, true, true>(maps, p, grid, st);
    case TNR_ACT_GELU: return launch<256, false, B_MN, TNR_ACT_GELU, true, true>(maps, p, grid, st);
    case TNR_ACT_TANH: return launch<256, false, B_MN, TNR_ACT_TANH, true, true>(maps, p, grid, st);
    case TNR_ACT_DGELU: return launch<256, false, B_MN, TNR_ACT_DGELU, true, true>(maps, p, grid, st);
    case TNR_ACT_GELU_DAUX: return launch<256, false, B_MN, TNR_ACT_GELU_DAUX, true, true>(maps, p, grid, st);
    case TNR_ACT_MULAUX: return launch<256, false, B_MN, TNR_ACT_MULAUX, true, true>(maps, p, grid, st);
  }
  set_error("tnr_gemm_bf16: unknown act %d", p.act);
  return 1;
}

template <int BN, bool A_MN, bool B_MN>
static int dispatch_act(const CUtensorMap* maps, const Params& p, int grid, cudaStream_t st) {
  switch (p.act) {
    case TNR_ACT_NONE:
      if (p.epi_tma) return launch<BN, A_MN, B_MN, TNR_ACT_NONE, true>(maps, p, grid, st);
      return launch<BN, A_MN, B_MN, TNR_ACT_NONE, false>(maps, p, grid, st);
    case TNR_ACT_GELU: return launch<BN, A_MN, B_MN, TNR_ACT_GELU, true>(maps, p, grid, st);
    case TNR_ACT_TANH: return launch<BN, A_MN, B_MN, TNR_ACT_TANH, true>(maps, p, grid, st);
    case TNR_ACT_DGELU: return launch<BN, A_MN, B_MN, TNR_ACT_DGELU, true>(maps, p, grid, st);
    case TNR_ACT_GELU_DAUX: return launch<BN, A_MN, B_MN, TNR_ACT_GELU_DAUX, true>(maps, p, grid, st);
    case TNR_ACT_MULAUX: return launch<BN, A_MN, B_MN, TNR_ACT_MULAUX, true>(maps, p, grid, st);
  }
  set_error("tnr_gemm_bf16: unknown act %d", p.act);
  return 1;
}

template <int BN>
static int dispatch_major(bool a_mn, bool b_mn, const CUtensorMap* maps, const Params& p, int grid, cudaStream_t st) {
  if (!a_mn && !b_mn) return dispatch_act<BN, false, false>(maps, p, grid, st);
  if (!a_mn && b_mn) return dispatch_act<BN, false, true>(maps, p, grid, st);
  if (a_mn && b_mn) {
    // wgrad only ever uses the plain epilogue
    TNR_REQUIRE(p.act == TNR_ACT_NONE, "tnr_gemm_bf16: A MN-major supports act=NONE only");
    TNR_REQUIRE(!p.epi_tma, "tnr_gemm_bf16: A MN-major (wgrad) writes fp32 only");
    return launch<BN, true, true, TNR_ACT_NONE, false>(maps, p, grid, st);
  }
  set_error("tnr_gemm_bf16: A MN-major with B K-major is not instantiated");
  return 1;
}

}  // namespace gemm
}  // namespace tnr

#ifdef GEMM_TIMING
extern "C" __attribute__((visibility("default"))) int tnr_debug_gemm_stamps(long long* host_out, int n) {
  return (int)cudaMemcpyFromSymbol(host_out, tnr::gemm::g_gemm_dbg, sizeof(long long) * (size_t)n);
}
#endif

extern "C" __attribute__((visibility("default"))) int tnr_gemm_bf16(const tnr_gemm_args* a, void* stream) {
  using namespace tnr;
  using namespace tnr::gemm;
  TNR_REQUIRE(a != nullptr, "tnr_gemm_bf16: null args");
  TNR_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "tnr_gemm_bf16: empty problem M=%d N=%d K=%d", a->M, a->N, a->K);
  TNR_REQUIRE(a->N % 8 == 0, "tnr_gemm_bf16: N=%d must be a multiple of 8", a->N);
  TNR_REQUIRE(a->lda % 8 == 0 && a->ldb % 8 == 0 && a->ldc % 8 == 0, "tnr_gemm_bf16: leading dims must be multiples of 8");
  TNR_REQUIRE(((uintptr_t)a->A % 16 == 0) && ((uintptr_t)a->B % 16 == 0) && ((uintptr_t)a->C % 16 == 0),
              "tnr_gemm_bf16: operands must be 16-byte aligned");
  TNR_REQUIRE(a->residual == nullptr || (a->ldr % 8 == 0 && (uintptr_t)a->residual % 16 == 0),
              "tnr_gemm_bf16: residual alignment");
  TNR_REQUIRE(a->aux == nullptr || (a->ldaux % 8 == 0 && (uintptr_t)a->aux % 16 == 0), "tnr_gemm_bf16: aux alignment");
  TNR_REQUIRE((a->act != TNR_ACT_DGELU && a->act != TNR_ACT_MULAUX) || a->aux != nullptr,
              "tnr_gemm_bf16: DGELU / MULAUX need aux (pre-activation / multiplier)");
  TNR_REQUIRE(a->bias == nullptr || (uintptr_t)a->bias % 16 == 0, "tnr_gemm_bf16: bias alignment");
  const bool atomic = a->split_k > 1 || a->accumulate != 0;
  TNR_REQUIRE(!atomic || a->c_dtype == TNR_F32, "tnr_gemm_bf16: split_k/accumulate require fp32 C");
  TNR_REQUIRE(!atomic || (a->act == TNR_ACT_NONE && a->residual == nullptr),
              "tnr_gemm_bf16: split_k/accumulate support the plain epilogue only");
  const bool a_mn = a->a_mn_major != 0, b_mn = a->b_mn_major != 0;
  if (a_mn) TNR_REQUIRE(a->M % 8 == 0, "tnr_gemm_bf16: MN-major A needs M %% 8 == 0");

  const int BN = (a->N > 128) ? 256 : 128;
  const bool atomic_out = a->split_k > 1 || a->accumulate != 0;
  // CTA-pair (cta_group::2) path: large-M GEMMs with a bf16 TMA-stored output.  TNR_GEMM_CTA2=0 disables it.
  static const int cta2_env = [] { const char* e = getenv("TNR_GEMM_CTA2"); return e ? atoi(e) : 1; }();
  //   forward / dgrad: bf16 output, A K-major, M >= 4096;  wgrad: both operands MN-major, fp32 (atomic) output.
  const bool cta2_fwd = BN == 256 && !a_mn && !atomic_out && a->c_dtype == TNR_BF16 && a->M >= 4096;
  const bool cta2_wgrad = BN == 256 && a_mn && b_mn && a->c_dtype == TNR_F32 && a->act == TNR_ACT_NONE && a->M >= 512 &&
                          a->residual == nullptr;
  const bool cta2 = cta2_env != 0 && (cta2_fwd || (cta2_wgrad && cta2_env != 2));
  Params p;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.m_tiles = (a->M + BM - 1) / BM;
  p.n_tiles = (a->N + BN - 1) / BN;
  p.k_blocks = (a->K + BK - 1) / BK;
  int splits = a->split_k > 1 ? a->split_k : 1;
  if (splits > p.k_blocks) splits = p.k_blocks;
  p.k_per_split = (p.k_blocks + splits - 1) / splits;
  p.splits = (p.k_blocks + p.k_per_split - 1) / p.k_per_split;
  p.C = a->C; p.ldc = a->ldc; p.c_f32 = (a->c_dtype == TNR_F32);
  p.bias = a->bias;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(a->residual); p.ldr = a->ldr;
  p.act = a->act;
  p.aux = reinterpret_cast<__nv_bfloat16*>(a->aux); p.ldaux = a->ldaux;
  p.atomic = atomic ? 1 : 0;
  p.drop = drop_or_none(a->drop);
  p.colsum = a->colsum;
  TNR_REQUIRE(a->colsum == nullptr || (a->c_dtype == TNR_BF16 && !atomic),
              "tnr_gemm_bf16: colsum needs a bf16 (TMA-staged) output");
  TNR_REQUIRE(p.drop.seed == nullptr || !(p.drop.p > 0.f) || (!atomic && a->act == TNR_ACT_NONE),
              "tnr_gemm_bf16: dropout is supported with the plain (bias + residual) epilogue only");
  TNR_REQUIRE(a->c_dtype != TNR_BF16 || atomic || a->N % 8 == 0, "tnr_gemm_bf16: a bf16 output needs N %% 8 == 0");

  CUtensorMap maps[5];
  CUtensorMap &ta = maps[0], &tb = maps[1];
  if (!a_mn) { if (make_map(&ta, a->A, a->K, a->M, a->lda, BK, BM)) return 1; }
  else       { if (make_map(&ta, a->A, a->M, a->K, a->lda, 64, BK)) return 1; }
  if (!b_mn) { if (make_map(&tb, a->B, a->K, a->N, a->ldb, BK, cta2 ? BN / 2 : BN)) return 1; }
  else       { if (make_map(&tb, a->B, a->N, a->K, a->ldb, 64, BK)) return 1; }
  // bf16 outputs go through shared memory + TMA (32-row x 64-column boxes, 128B swizzle)
  p.epi_tma = (!p.c_f32 && !atomic) ? 1 : 0;
  TNR_REQUIRE(p.epi_tma || a->act == TNR_ACT_NONE, "tnr_gemm_bf16: activation epilogues need a bf16 output");
  p.in_mode = 0; p.aux_out = 0;
  maps[2] = ta; maps[3] = ta; maps[4] = ta;        // placeholders when unused
  if (p.epi_tma) {
    TNR_REQUIRE(!((a->act == TNR_ACT_DGELU || a->act == TNR_ACT_MULAUX) && a->residual != nullptr),
                "tnr_gemm_bf16: DGELU / MULAUX with a residual is not supported for bf16 outputs");
    if (make_map(&maps[2], a->C, a->N, a->M, a->ldc, 64, 32)) return 1;
    if (a->act == TNR_ACT_DGELU || a->act == TNR_ACT_MULAUX) {
      p.in_mode = 2;
      if (make_map(&maps[3], a->aux, a->N, a->M, a->ldaux, 64, 32)) return 1;
    } else if (a->residual != nullptr) {
      p.in_mode = 1;
      if (make_map(&maps[3], a->residual, a->N, a->M, a->ldr, 64, 32)) return 1;
    }
    if ((a->act == TNR_ACT_GELU || a->act == TNR_ACT_GELU_DAUX) && a->aux != nullptr) {
      p.aux_out = 1;
      if (make_map(&maps[4], a->aux, a->N, a->M, a->ldaux, 64, 32)) return 1;
    }
  }

  int sms = num_sms();
  TNR_REQUIRE(sms > 0, "tnr_gemm_bf16: no CUDA device");
  sms -= sm_reserve();          // SMs left to a concurrently running collective (tnr_set_sm_reserve)
  TNR_REQUIRE(sms >= 2, "tnr_gemm_bf16: SM reserve leaves no SMs");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (cta2) {
    const int pair_tiles = ((p.m_tiles + 1) / 2) * p.n_tiles * p.splits;
    const int pairs = pair_tiles < sms / 2 ? pair_tiles : sms / 2;
    if (a_mn) return launch<256, true, true, TNR_ACT_NONE, false, true>(maps, p, pairs, st);
    return b_mn ? dispatch_act_cta2<true>(maps, p, pairs, st) : dispatch_act_cta2<false>(maps, p, pairs, st);
  }
  const int tiles = p.m_tiles * p.n_tiles * p.splits;
  const int grid = tiles < sms ? tiles : sms;
  if (BN == 256) return dispatch_major<256>(a_mn, b_mn, maps, p, grid, st);
  return dispatch_major<128>(a_mn, b_mn, maps, p, grid, st);
}
