// TEST-ONLY: a tiny single-process CUDA execution-model emulator (fibers), used to run the
// product's kernels and host orchestration on the CPU at toy sizes inside `pytest -m "not gpu"`.
// It exists because this container has no GPU: it lets index math, barrier structure and the
// host-side pass orchestration be checked (optionally under ASan) before GPU time is spent.
//
// It is NOT a fallback: the product library (libgspb200.so) never includes this header.  Only
// tests/emu/Makefile builds libgspb200_emu.so from the same .cu sources with -DGSP_EMU, and only
// tests/ load it.  See DESIGN.md "Emulated build".
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

// ---------------------------------------------------------------- vector types
struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct double2 { double x, y; };
struct int2 { int x, y; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct int4 { int x, y, z, w; };
struct alignas(16) longlong2 { long long x, y; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }

// ---------------------------------------------------------------- runtime API subset
typedef int cudaError_t;
typedef struct emuStream_* cudaStream_t;
// ---- stream-dependency recorder (GSP_DEPCHECK=1): every launch becomes a node ordered by its stream, by the events its stream
// waited on and by host synchronisations; launches annotate the matrix regions they read / write; dep_check() reports every pair of
// launches that touch overlapping regions (at least one writing) WITHOUT a happens-before path - a missing cudaStreamWaitEvent that
// the sequential emulator would otherwise execute correctly by accident.
namespace emu {
void dep_enable(bool on);
bool dep_enabled();
void* dep_new_stream();
void dep_access(const void* base, long long r0, long long r1, long long c0, long long c1, bool write);
void dep_launch(void* stream, const char* name);
void dep_event_record(void* ev, void* stream);
void dep_stream_wait(void* stream, void* ev);
void dep_host_sync(void* stream);  // nullptr: every stream
long long dep_check(int verbose);
}  // namespace emu
typedef struct emuEvent_ { double t; }* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
static inline const char* cudaGetErrorString(cudaError_t e) { return e ? "emu error" : "no error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { ::emu::dep_host_sync(nullptr); return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocAsync(void** p, size_t n, cudaStream_t) { return cudaMalloc(p, n); }
static inline cudaError_t cudaMemGetInfo(size_t* fr, size_t* tot) { *fr = *tot = 0; return cudaSuccess; }
static inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { return cudaFree(p); }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t st = 0) { ::emu::dep_launch((void*)st, "memcpy"); std::memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t = 0) {
  for (size_t r = 0; r < h; ++r) std::memmove((char*)d + r * dp, (const char*)s + r * sp, w);
  return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t st = 0) { ::emu::dep_launch((void*)st, "memset"); std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)::emu::dep_new_stream(); return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = (cudaStream_t)::emu::dep_new_stream(); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = (cudaStream_t)::emu::dep_new_stream(); return cudaSuccess; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = -5; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t s) { ::emu::dep_host_sync(s ? (void*)s : (void*)1); return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned = 0) { ::emu::dep_stream_wait((void*)s, (void*)e); return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emuEvent_{0}; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = 0) { ::emu::dep_event_record((void*)e, (void*)s); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
template <class F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 2; return cudaSuccess; }
struct cudaDeviceProp { int multiProcessorCount; size_t sharedMemPerBlockOptin; char name[64]; int major, minor; };
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
  p->multiProcessorCount = 4; p->sharedMemPerBlockOptin = 227 * 1024; std::strcpy(p->name, "gsp-emu"); p->major = 10; p->minor = 0;
  return cudaSuccess;
}
enum { cudaDevAttrMultiProcessorCount = 16 };
static inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 4; return cudaSuccess; }
static inline cudaError_t cudaDeviceCanAccessPeer(int* can, int, int) { *can = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaMemcpyPeerAsync(void* d, int, const void* s, int, size_t n, cudaStream_t st = 0) { ::emu::dep_launch((void*)st, "memcpy_peer"); std::memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaHostRegister(void*, size_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaHostUnregister(void*) { return cudaSuccess; }

// ---------------------------------------------------------------- execution model
namespace emu {

struct Fiber {
  ucontext_t ctx;
  char* stack = nullptr;
  bool done = false;
  // wait state
  int wait_kind = 0;   // 0 runnable, 1 block barrier, 2 named barrier, 3 warp rendezvous, 4 spin (always runnable)
  int wait_id = 0;
  int block = 0;  // index into State::blocks
  uint3 tid;
};

struct BlockState {
  uint3 bid;
  unsigned char* dyn_smem = nullptr;
  int live = 0;
  int bar_waiting = 0;
  int named_count[16] = {0};
  int named_gen[16] = {0};
};

struct State {
  std::vector<Fiber> fibers;
  ucontext_t sched;
  int cur = -1;
  int nthreads = 0;
  int tpb = 0;     // threads per block
  int live = 0;
  std::function<void()> body;
  std::vector<BlockState> blocks;
  // warp rendezvous
  std::vector<int> warp_arrived, warp_gen;
  std::vector<uint64_t> warp_buf;   // 32 slots per warp
  // set whenever a wait condition is satisfied (warp rendezvous completes, mbarrier phase flips, a flag is released): kernels
  // whose only yield points are spin-type waits (mbarrier pipelines) would otherwise look deadlocked to the scheduler
  bool event = false;
};

extern State* g;
extern uint3 g_threadIdx, g_blockIdx;
extern dim3 g_blockDim, g_gridDim;

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
void launch_coop(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);  // all blocks co-resident
inline unsigned char* dyn_smem() { return g->blocks[g->fibers[g->cur].block].dyn_smem; }
void yield_to_sched();
inline void note_event();
void syncthreads();
void named_bar_sync(int id, int count);
void named_bar_arrive(int id, int count);
// warp collective: every live lane of the warp deposits `v`, gets the whole table back
void warp_exchange(uint64_t v, uint64_t out[32]);
void warp_sync();
// cooperative spin: call inside a polling loop so other fibers can make progress
inline void spin_yield() { g->fibers[g->cur].wait_kind = 4; yield_to_sched(); g->fibers[g->cur].wait_kind = 0; }
inline void note_event() { g->event = true; }
inline int lane() { return (int)(g_threadIdx.x + g_blockDim.x * (g_threadIdx.y + g_blockDim.y * g_threadIdx.z)) & 31; }

}  // namespace emu

#define threadIdx (::emu::g_threadIdx)
#define blockIdx (::emu::g_blockIdx)
#define blockDim (::emu::g_blockDim)
#define gridDim (::emu::g_gridDim)

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __shared__ static
#define __launch_bounds__(...)
#define __constant__ static

static inline void __syncthreads() { emu::syncthreads(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_sync(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

template <class T> static inline T __ldg(const T* p) { return *p; }

template <class T> static inline T emu_shfl_idx(T v, int src) {
  static_assert(sizeof(T) <= 8, "shfl payload");
  uint64_t bits = 0, tab[32];
  std::memcpy(&bits, &v, sizeof(T));
  emu::warp_exchange(bits, tab);
  T out;
  std::memcpy(&out, &tab[src & 31], sizeof(T));
  return out;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) { return emu_shfl_idx(v, src); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return emu_shfl_idx(v, emu::lane() ^ m); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d, int = 32) {
  int s = emu::lane() + d;
  return emu_shfl_idx(v, s > 31 ? emu::lane() : s);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  uint64_t tab[32];
  emu::warp_exchange(pred ? 1 : 0, tab);
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= (tab[i] ? 1u : 0u) << i;
  return r;
}

static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, !pred) == 0u; }
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { auto o = *p; *p = o + v; return o; }
template <class T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }
template <class T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }

static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline double sinpi(double x) { return std::sin(M_PI * x); }
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
static inline void sincospi(double x, double* s, double* c) { *s = std::sin(M_PI * x); *c = std::cos(M_PI * x); }
static inline double __dsqrt_rn(double x) { return std::sqrt(x); }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
using std::pow; using std::sinh; using std::cosh; using std::sin; using std::fma; using std::sqrt; using std::exp; using std::fabs; using std::log; using std::floor; using std::fmin; using std::fmax;

namespace emu {
// arguments are evaluated eagerly (like a real launch) and copied into the per-thread closure
template <class K, class... Args>
inline void launch_k(dim3 grid, dim3 block, size_t smem, K kernel, Args... args) {
  launch(grid, block, smem, [=]() { kernel(args...); });
}
}  // namespace emu
template <class K, class... Args>
inline void emu_launch_coop_k(dim3 grid, dim3 block, size_t smem, K kernel, Args... args) {
  ::emu::launch_coop(grid, block, smem, [=]() { kernel(args...); });
}
#define GSP_LAUNCH(kernel, grid, block, smem, stream, ...) \
  (::emu::dep_launch((void*)(stream), #kernel), ::emu::launch_k((grid), (block), (smem), (kernel), __VA_ARGS__))
// persistent kernels whose CTAs wait on each other: the whole grid must be co-resident
#define GSP_LAUNCH_COOP(kernel, grid, block, smem, stream, ...) \
  (::emu::dep_launch((void*)(stream), #kernel), ::emu_launch_coop_k((grid), (block), (smem), (kernel), __VA_ARGS__))
