// Host-side shared declarations of libgspb200: context, error plumbing, device buffers.
#pragma once
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/gsp_b200.h"
#include "gsp_rt.h"

namespace gsp {

extern long long g_launches;  // kernels launched by this library (gsp_kernel_launches)

// Optional per-kernel-class timing with CUDA events on the launching stream (gsp_profile_enable).
// Off by default: the timed paths carry no events besides the per-call pair.
struct Prof {
  struct Pending { std::string name; cudaEvent_t e0, e1; int dev; cudaStream_t st; };
  bool on = false;
  std::vector<Pending> pending;
  std::vector<std::pair<int, cudaEvent_t>> ref;  // timeline origin per device (mark_reference)
  // GSP_PROF_TIMELINE: idle all `devs`, then record one origin event per device back to back (aligned to ~10 us across devices)
  void mark_reference(const std::vector<int>& devs);
  std::vector<std::pair<std::string, std::pair<double, long long>>> acc;  // name -> (ms, launches)
  void add(const std::string& n, double ms);
  void flush();
};
extern Prof g_prof;
struct ProfScope {
  cudaStream_t st;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  const char* name;
  ProfScope(const char* n, cudaStream_t s) : st(s), name(n) {
    if (g_prof.on) {
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      cudaEventRecord(e0, st);
    }
  }
  ~ProfScope() {
    if (e0) {
      cudaEventRecord(e1, st);
      int dev = 0;
      cudaGetDevice(&dev);
      g_prof.pending.push_back({name, e0, e1, dev, st});
    }
  }
};

struct DevCtx {
  int dev = 0;
  cudaStream_t stream = nullptr;   // compute (highest priority: carries the critical path)
  cudaStream_t aux = nullptr;      // medium priority: panel solves and look-ahead updates next to the panel chain
  cudaStream_t h2d = nullptr;      // copy-in
  cudaStream_t d2h = nullptr;      // copy-out
  static constexpr int kSide = 8;
  cudaStream_t side[kSide] = {};   // low-priority streams for look-ahead work (one per recursion depth)
  int sms = 148;
  size_t smem_optin = 227 * 1024;
};

}  // namespace gsp

struct gsp_ctx {
  std::vector<gsp::DevCtx> devs;
  bool peer_ok = true;  // every device of the context can map every other one's memory (needed by the distributed factorization)
  std::string err;
  std::mutex mu;
  double last_sample_ms = 0.0;
};

namespace gsp {

int set_err(gsp_ctx* ctx, int code, const std::string& msg);
// " [device d: x of y MiB free]" when the failure is an allocation (which device ran out, and by how much)
inline std::string mem_note(cudaError_t e) {
  if (e != cudaErrorMemoryAllocation) return "";
  int dev = 0;
  size_t fr = 0, tot = 0;
  cudaGetLastError();
  if (cudaGetDevice(&dev) != cudaSuccess || cudaMemGetInfo(&fr, &tot) != cudaSuccess) return "";
  return " [device " + std::to_string(dev) + ": " + std::to_string(fr >> 20) + " of " + std::to_string(tot >> 20) + " MiB free]";
}

#define GSP_CUDA_OK(ctx, expr)                                                                          \
  do {                                                                                                  \
    cudaError_t e_ = (expr);                                                                            \
    if (e_ != cudaSuccess)                                                                              \
      return ::gsp::set_err((ctx), GSP_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_) + ::gsp::mem_note(e_)); \
  } while (0)

#define GSP_TRY(expr)            \
  do {                           \
    int rc_ = (expr);            \
    if (rc_ != GSP_OK) return rc_; \
  } while (0)

// RAII device buffer (freed on the device it was allocated on).  Memory comes from the device's default stream-ordered pool, whose
// release threshold gsp_ctx_create raises (GSP_MEMPOOL_KEEP, default "everything"): a freed 2-9 GB covariance matrix stays mapped
// and the next plan's allocation costs microseconds instead of the 2-35 ms a cudaMalloc of that size was measured at.  release()
// keeps cudaFree's semantics (the device is idle before the block can be handed out again): plans use non-blocking streams.
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int dev = 0;
  cudaStream_t owner = nullptr;  // stream every use of the block is ordered on (or joined into) - see release()
  bool pooled = true;            // false: the block came from cudaMalloc (fallback, see alloc)
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  // `st`: the stream that all work on the block is ordered on; the block is then freed stream-ordered (cudaFreeAsync on st), with
  // no synchronisation at all.  Without it (nullptr) release() falls back to a device-wide synchronisation before the free.
  cudaError_t alloc(int device, size_t n, cudaStream_t st = nullptr) {
    release();
    dev = device;
    owner = st;
    pooled = true;
    cudaSetDevice(dev);
    cudaError_t e = cudaMallocAsync(&p, n ? n : 16, (cudaStream_t)0);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)0);
#ifndef GSP_EMU
    if (e == cudaErrorMemoryAllocation) {
      // Seen on a 2-GPU box with > 170 GB free: the pool refused a 2 GB block right after peer access had been granted on a pool
      // that held ~20 GB of cached blocks.  Hand the cache back and retry; then fall back to a plain cudaMalloc (peer access to
      // it comes from cudaDeviceEnablePeerAccess).
      cudaGetLastError();
      cudaMemPool_t pool = nullptr;
      if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess && pool) {
        cudaDeviceSynchronize();
        cudaMemPoolTrimTo(pool, 0);
      }
      p = nullptr;
      e = cudaMallocAsync(&p, n ? n : 16, (cudaStream_t)0);
      if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)0);
      if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        p = nullptr;
        pooled = false;
        e = cudaMalloc(&p, n ? n : 16);
      }
    }
#endif
    if (e == cudaSuccess) bytes = n;
    else p = nullptr;
    return e;
  }
  void release() {
    if (p) {
      cudaSetDevice(dev);
      if (!pooled) {
        cudaFree(p);  // synchronises the device
      } else if (owner) {
        cudaFreeAsync(p, owner);
      } else {
        cudaDeviceSynchronize();
        cudaFreeAsync(p, (cudaStream_t)0);
      }
      p = nullptr;
      bytes = 0;
    }
  }
  template <class T> T* as() const { return (T*)p; }
};

inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

}  // namespace gsp
