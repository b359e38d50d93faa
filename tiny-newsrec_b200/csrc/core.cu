// Error plumbing, device check.
#include "common.cuh"
#include <cstring>
#include <thread>
#include <vector>

namespace tnr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return 2;
}

static std::atomic<int> g_sm_reserve[MAX_DEVICES];

int sm_reserve() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) return 0;
  return g_sm_reserve[dev].load(std::memory_order_relaxed);
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace tnr

extern "C" __attribute__((visibility("default"))) const char* tnr_last_error(void) { return tnr::g_err; }
extern "C" __attribute__((visibility("default"))) int tnr_abi_version(void) { return TNR_ABI_VERSION; }

extern "C" __attribute__((visibility("default"))) int tnr_device_check(int* sms) {
  int dev = 0, major = 0, minor = 0;
  TNR_CHECK_CUDA(cudaGetDevice(&dev));
  TNR_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  TNR_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  TNR_REQUIRE(major == 10, "libtinyrec is built for sm_100a only; device %d is sm_%d%d (no fallback path)", dev, major,
              minor);
  if (sms) *sms = tnr::num_sms();
  return 0;
}

extern "C" __attribute__((visibility("default"))) int tnr_set_sm_reserve(int n_sms) {
  int dev = 0;
  TNR_CHECK_CUDA(cudaGetDevice(&dev));
  TNR_REQUIRE(dev >= 0 && dev < tnr::MAX_DEVICES, "device index %d out of range", dev);
  TNR_REQUIRE(n_sms >= 0 && n_sms <= 64 && n_sms % 2 == 0, "tnr_set_sm_reserve: 0 <= n_sms <= 64, even (CTA pairs), got %d", n_sms);
  tnr::g_sm_reserve[dev].store(n_sms, std::memory_order_relaxed);
  return 0;
}

// Host-side staging copy for the loaders: dst (pinned) <- src (pageable numpy memory), `bytes` split over up to
// `n_threads` host threads.  One core moves ~10 GB/s; the eval driver stages ~220 MB of index arrays for the
// 376k-impression dev set, and Python threads cannot do this in parallel without fighting over the GIL (ctypes drops
// the GIL for the duration of this call).  Replaces the reference's loader thread (dataloader.py:303-314).
extern "C" __attribute__((visibility("default"))) int tnr_host_copy_mt(void* dst, const void* src, long long bytes, int n_threads) {
  TNR_REQUIRE(bytes >= 0 && (bytes == 0 || (dst != nullptr && src != nullptr)), "tnr_host_copy_mt: null buffer");
  if (bytes == 0) return 0;
  const long long min_chunk = 1 << 20;
  long long nt = n_threads < 1 ? 1 : n_threads;
  if (nt > 16) nt = 16;
  if (nt > (bytes + min_chunk - 1) / min_chunk) nt = (bytes + min_chunk - 1) / min_chunk;
  if (nt <= 1) { memcpy(dst, src, (size_t)bytes); return 0; }
  const long long per = ((bytes + nt - 1) / nt + 63) / 64 * 64;
  std::vector<std::thread> th;
  th.reserve((size_t)nt - 1);
  for (long long t = 1; t < nt; ++t) {
    const long long lo = t * per, hi = lo + per < bytes ? lo + per : bytes;
    if (lo >= hi) break;
    th.emplace_back([=] { memcpy((char*)dst + lo, (const char*)src + lo, (size_t)(hi - lo)); });
  }
  memcpy(dst, src, (size_t)(per < bytes ? per : bytes));
  for (auto& x : th) x.join();
  return 0;
}
