// Gradient all-reduce over NVLink 5 / NVSwitch PEER MEMORY, one kernel per bucket (replaces the Horovod all-reduce of
// Tiny-NewsRec/run.py:144-149 for the data-parallel train step; NCCL stays the fallback when the symmetric mapping is
// not available).
//
// Every rank holds the flat fp32 gradient buffer at the same offset of a symmetric allocation whose peer copies are
// mapped into this process (ptrs[r] = rank r's buffer).  Two-shot all-reduce in ONE launch of a few CTAs:
//   barrier A   block b of every rank tells block b of every other rank "my bucket is final" (one 4-byte flag store per
//               peer into the peer's flag page, st.release.sys) and waits for the world's flags in its own page
//   reduce      rank r owns slice r of the bucket: it reads that slice from every rank IN RANK ORDER (ld over NVLink),
//               adds in fp32 and stores the sum into every rank's buffer -- each element is summed by exactly one rank,
//               so all replicas receive bit-identical gradients (what keeps data-parallel replicas identical)
//   barrier B   "my slice is written everywhere"; when the kernel ends, the local bucket holds the sum.
// A bucket of the train step is 2.4 - 31 MB: NCCL's all-reduce costs ~130 us of launch / protocol latency whatever the
// size (measured at N = 2, profiles/), which is what the trailing buckets of a step expose; here the floor is two flag
// round trips over NVLink (~2 x 3 us).  Flags carry a per-block sequence number kept in the flag page itself, so the
// kernel can be replayed inside a CUDA graph without host-side state.
#include "common.cuh"

namespace tnr {

constexpr int AR_THREADS = 512;
constexpr int AR_MAX_WORLD = 16;
constexpr int AR_MAX_CTAS = 32;
// flag page layout (uint32 words): [phase (2)][block (AR_MAX_CTAS)][source rank (AR_MAX_WORLD)] flags, then the
// per-block sequence counters
constexpr int AR_FLAG_WORDS = 2 * AR_MAX_CTAS * AR_MAX_WORLD + AR_MAX_CTAS;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer(const float4* p) {       // peer data: never from this SM's L1
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void ar_barrier(uint32_t* const* flags, int rank, int world, int phase, uint32_t seq) {
  __syncthreads();                                   // this block's work before the barrier is done
  const int slot = (phase * AR_MAX_CTAS + (int)blockIdx.x) * AR_MAX_WORLD;
  if ((int)threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(flags[threadIdx.x] + slot + rank, seq);                 // to peer threadIdx.x: "rank `rank`, block b: here"
    const uint32_t* mine = flags[rank] + slot + threadIdx.x;               // from peer threadIdx.x
    unsigned long long spins = 0;
    while ((int)(ld_acquire_sys(mine) - seq) < 0)
      if (++spins > (1ull << 31)) __trap();          // a rank that never arrives: fail loudly instead of hanging the box
  }
  __syncthreads();
}

// WORLD > 0: compile-time world size (the rank loop unrolls: all U x WORLD loads of a thread are in flight together --
// with a runtime loop each rank's loads waited for the previous rank's adds, ~3 us of NVLink latency per rank and pass);
// WORLD == 0: any world size up to AR_MAX_WORLD.
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float4* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float4* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// NVLS form of the same two-shot exchange: the buffer is also mapped at a MULTICAST address; one multimem.ld_reduce
// returns the sum of an element over all ranks, added inside the NVSwitch, and one multimem.st writes the result to
// every rank -- world x fewer requests per thread than the peer-load loop below, which at 8 GPUs could not keep
// enough bytes in flight from 4 CTAs (31 MB in 1.1 ms).  A slice is still reduced by exactly one rank: replicas get
// bit-identical sums.
__global__ void __launch_bounds__(AR_THREADS)
allreduce_nvls_kernel(float4* __restrict__ mc, uint32_t* const* __restrict__ flags, int rank, int world, long long off4,
                      long long n4) {
  __shared__ uint32_t* s_flag[AR_MAX_WORLD];
  __shared__ uint32_t s_seq;
  if ((int)threadIdx.x < world) s_flag[threadIdx.x] = flags[threadIdx.x];
  if (threadIdx.x == 0) {
    uint32_t* cnt = flags[rank] + 2 * AR_MAX_CTAS * AR_MAX_WORLD + blockIdx.x;
    s_seq = *cnt + 1u;
    *cnt = s_seq;
  }
  __syncthreads();
  const uint32_t seq = s_seq;
  ar_barrier(s_flag, rank, world, 0, seq);
  const long long per = (n4 + world - 1) / world;
  const long long lo = (long long)rank * per, hi = (lo + per < n4) ? lo + per : n4;
  const long long stride = (long long)gridDim.x * AR_THREADS;
  float4* base = mc + off4;
  constexpr int U = 8;
  for (long long i0 = lo + (long long)blockIdx.x * AR_THREADS + threadIdx.x; i0 < hi; i0 += U * stride) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < hi) v[u] = multimem_ld_reduce_add(base + i);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < hi) multimem_st(base + i, v[u]);
    }
  }
  ar_barrier(s_flag, rank, world, 1, seq);
}

template <int WORLD>
__global__ void __launch_bounds__(AR_THREADS)
allreduce_p2p_kernel(float* const* __restrict__ ptrs, uint32_t* const* __restrict__ flags, int rank, int world_rt,
                     long long off4, long long n4) {
  const int world = WORLD > 0 ? WORLD : world_rt;
  __shared__ float4* s_ptr[AR_MAX_WORLD];
  __shared__ uint32_t* s_flag[AR_MAX_WORLD];
  __shared__ uint32_t s_seq;
  if ((int)threadIdx.x < world) {
    s_ptr[threadIdx.x] = reinterpret_cast<float4*>(ptrs[threadIdx.x]) + off4;
    s_flag[threadIdx.x] = flags[threadIdx.x];
  }
  if (threadIdx.x == 0) {
    uint32_t* cnt = flags[rank] + 2 * AR_MAX_CTAS * AR_MAX_WORLD + blockIdx.x;   // this block's own counter: nobody else touches it
    s_seq = *cnt + 1u;
    *cnt = s_seq;
  }
  __syncthreads();
  const uint32_t seq = s_seq;
  ar_barrier(s_flag, rank, world, 0, seq);
  // slice `rank` of the bucket, in float4 units, spread over the blocks
  const long long per = (n4 + world - 1) / world;
  const long long lo = (long long)rank * per, hi = (lo + per < n4) ? lo + per : n4;
  const long long stride = (long long)gridDim.x * AR_THREADS;
  constexpr int U = WORLD > 0 ? (WORLD <= 2 ? 8 : (WORLD <= 4 ? 4 : 2)) : 2;      // U x WORLD 16-byte loads in flight per thread
  float4* pr[WORLD > 0 ? WORLD : 1];
  if (WORLD > 0) {
#pragma unroll
    for (int r = 0; r < WORLD; ++r) pr[r] = s_ptr[r];
  }
  for (long long i0 = lo + (long long)blockIdx.x * AR_THREADS + threadIdx.x; i0 < hi; i0 += U * stride) {
    float4 acc[U];
    if (WORLD > 0) {
      float4 v[WORLD > 0 ? WORLD : 1][U];
#pragma unroll
      for (int r = 0; r < WORLD; ++r)
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const long long i = i0 + u * stride;
          v[r][u] = i < hi ? ld_peer(pr[r] + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        acc[u] = v[0][u];
#pragma unroll
        for (int r = 1; r < WORLD; ++r) {               // rank order: the same sum on every replica
          acc[u].x += v[r][u].x; acc[u].y += v[r][u].y; acc[u].z += v[r][u].z; acc[u].w += v[r][u].w;
        }
      }
#pragma unroll
      for (int r = 0; r < WORLD; ++r)
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const long long i = i0 + u * stride;
          if (i < hi) pr[r][i] = acc[u];
        }
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < world; ++r) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const long long i = i0 + u * stride;
          v[u] = i < hi ? ld_peer(s_ptr[r] + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) { acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w; }
      }
      for (int r = 0; r < world; ++r) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const long long i = i0 + u * stride;
          if (i < hi) s_ptr[r][i] = acc[u];
        }
      }
    }
  }
  ar_barrier(s_flag, rank, world, 1, seq);
}

}  // namespace tnr

using namespace tnr;
#define TNR_API extern "C" __attribute__((visibility("default")))

TNR_API long long tnr_allreduce_p2p_flag_words(void) { return AR_FLAG_WORDS; }

TNR_API int tnr_allreduce_p2p(void* const* ptrs_dev, void* multicast_ptr, void* const* flags_dev, int rank, int world,
                              long long off, long long n, int n_ctas, void* stream) {
  TNR_REQUIRE(ptrs_dev != nullptr && flags_dev != nullptr, "tnr_allreduce_p2p: null pointer tables");
  TNR_REQUIRE(world >= 1 && world <= AR_MAX_WORLD && rank >= 0 && rank < world, "tnr_allreduce_p2p: rank %d / world %d", rank, world);
  TNR_REQUIRE(off % 4 == 0 && n % 4 == 0 && off >= 0 && n >= 0, "tnr_allreduce_p2p: offset and count must be multiples of 4 floats");
  TNR_REQUIRE(n_ctas >= 1 && n_ctas <= AR_MAX_CTAS, "tnr_allreduce_p2p: 1 <= n_ctas <= %d", AR_MAX_CTAS);
  if (n == 0 || world == 1) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* const* pp = reinterpret_cast<float* const*>(ptrs_dev);
  uint32_t* const* ff = reinterpret_cast<uint32_t* const*>(flags_dev);
  if (multicast_ptr != nullptr) {
    TNR_REQUIRE((uintptr_t)multicast_ptr % 16 == 0, "tnr_allreduce_p2p: multicast pointer must be 16-byte aligned");
    allreduce_nvls_kernel<<<n_ctas, AR_THREADS, 0, st>>>(reinterpret_cast<float4*>(multicast_ptr), ff, rank, world, off / 4, n / 4);
    TNR_LAUNCH_CHECK();
    return 0;
  }
  switch (world) {
    case 2: allreduce_p2p_kernel<2><<<n_ctas, AR_THREADS, 0, st>>>(pp, ff, rank, world, off / 4, n / 4); break;
    case 4: allreduce_p2p_kernel<4><<<n_ctas, AR_THREADS, 0, st>>>(pp, ff, rank, world, off / 4, n / 4); break;
    case 8: allreduce_p2p_kernel<8><<<n_ctas, AR_THREADS, 0, st>>>(pp, ff, rank, world, off / 4, n / 4); break;
    default: allreduce_p2p_kernel<0><<<n_ctas, AR_THREADS, 0, st>>>(pp, ff, rank, world, off / 4, n / 4); break;
  }
  TNR_LAUNCH_CHECK();
  return 0;
}
