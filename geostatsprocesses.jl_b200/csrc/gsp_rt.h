// Thin device-runtime layer: everything that is inline PTX (DMMA, bulk async copies, mbarriers)
// lives here behind small functions.  Product builds compile this with nvcc for sm_100a.
// With -DGSP_EMU (tests/emu only) the same names map onto a CPU fiber emulator so that kernel
// index math can be checked without a GPU; that build is never shipped or loaded by the package.
#pragma once
#include <stdint.h>

#ifdef GSP_EMU
#include "gsp_emu.h"
#define GSP_DYN_SMEM(name) unsigned char* name = ::emu::dyn_smem()
#define GSP_HD
#else
#include <cuda_runtime.h>
#define GSP_DYN_SMEM(name)                                   \
  extern __shared__ __align__(128) unsigned char name##_raw_[]; \
  unsigned char* name = name##_raw_
#define GSP_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define GSP_LAUNCH_COOP(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define GSP_HD __host__ __device__
#endif

// Region annotations for the emulator's stream-dependency checker (tests/emu: GSP_DEPCHECK=1); nothing in the product build.
// Units are the callers' choice per base pointer (the factorization uses 128-block rows / columns of a device's matrix).
#ifdef GSP_EMU
#define GSP_DEP_ACCESS(base, r0, r1, c0, c1, w) ::emu::dep_access((const void*)(base), (long long)(r0), (long long)(r1), (long long)(c0), (long long)(c1), (w))
#else
#define GSP_DEP_ACCESS(base, r0, r1, c0, c1, w) ((void)0)
#endif

#define GSP_DEV __device__ __forceinline__
#define GSP_DEV_NOINLINE static __device__ __noinline__  // big scalar routines (bessel_k): one copy per kernel, not one per call site

namespace gsp {

struct cplx {
  double re, im;
};
GSP_DEV cplx cmul(cplx a, cplx b) { return cplx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
GSP_DEV cplx cadd(cplx a, cplx b) { return cplx{a.re + b.re, a.im + b.im}; }
GSP_DEV cplx csub(cplx a, cplx b) { return cplx{a.re - b.re, a.im - b.im}; }
GSP_DEV cplx cconj(cplx a) { return cplx{a.re, -a.im}; }

// ------------------------------------------------------------------ DMMA 8x8x4 (FP64 tensor core)
// D(8x8) += A(8x4, row) * B(4x8, col).  Lane l holds A[l/4][l%4], B[l%4][l/4], C[l/4][2*(l%4)+{0,1}].
GSP_DEV void dmma884(double& c0, double& c1, double a, double b) {
#ifdef GSP_EMU
  uint64_t ta[32], tb[32], ua, ub;
  memcpy(&ua, &a, 8);
  memcpy(&ub, &b, 8);
  emu::warp_exchange(ua, ta);
  emu::warp_exchange(ub, tb);
  int l = emu::lane(), r = l / 4, c = 2 * (l % 4);
  for (int k = 0; k < 4; ++k) {
    double av, b0, b1;
    memcpy(&av, &ta[r * 4 + k], 8);
    memcpy(&b0, &tb[c * 4 + k], 8);
    memcpy(&b1, &tb[(c + 1) * 4 + k], 8);
    c0 = fma(av, b0, c0);
    c1 = fma(av, b1, c1);
  }
#else
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
#endif
}

// ------------------------------------------------------------------ mbarrier + bulk async copy (TMA engine)
#ifdef GSP_EMU
struct mbar_t {
  int expected;  // arrivals per phase
  int pending;   // arrivals still missing in the current phase
  int tx;        // outstanding transaction bytes
  int phase;     // completed-phase parity
};
GSP_DEV void mbar_init(mbar_t* b, int count) { b->expected = count; b->pending = count; b->tx = 0; b->phase = 0; }
GSP_DEV void mbar_check_(mbar_t* b) {
  if (b->pending == 0 && b->tx == 0) { b->phase ^= 1; b->pending = b->expected; emu::note_event(); }
}
GSP_DEV void mbar_arrive(mbar_t* b) { b->pending--; mbar_check_(b); }
GSP_DEV void mbar_arrive_expect_tx(mbar_t* b, uint32_t bytes) { b->tx += (int)bytes; b->pending--; mbar_check_(b); }
GSP_DEV bool mbar_try_wait(mbar_t* b, uint32_t parity) { return (uint32_t)b->phase != parity; }
GSP_DEV void mbar_wait(mbar_t* b, uint32_t parity) {
  while (!mbar_try_wait(b, parity)) emu::spin_yield();
}
GSP_DEV void fence_mbar_init() {}
GSP_DEV void fence_proxy_async() {}
GSP_DEV void fence_proxy_async_all() {}
// dst: shared, src: global, bytes % 16 == 0, both 16-byte aligned
GSP_DEV void bulk_g2s(void* dst, const void* src, uint32_t bytes, mbar_t* b) {
  if ((((uintptr_t)dst) & 15) || (((uintptr_t)src) & 15) || (bytes & 15)) {
    fprintf(stderr, "gsp_emu: misaligned bulk copy dst=%p src=%p bytes=%u\n", dst, src, bytes);
    abort();
  }
  memcpy(dst, src, bytes);
  b->tx -= (int)bytes;
  mbar_check_(b);
}
#else
typedef uint64_t mbar_t;
GSP_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
GSP_DEV void mbar_init(mbar_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
GSP_DEV void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
GSP_DEV void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// all state spaces: global data written through the generic proxy (plain stores, possibly by another CTA) before TMA reads it
GSP_DEV void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
GSP_DEV void mbar_arrive(mbar_t* b) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(smem_u32(b)) : "memory");
}
GSP_DEV void mbar_arrive_expect_tx(mbar_t* b, uint32_t bytes) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(smem_u32(b)), "r"(bytes)
               : "memory");
}
GSP_DEV bool mbar_try_wait(mbar_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
      : "=r"(ok)
      : "r"(smem_u32(b)), "r"(parity)
      : "memory");
  return ok != 0;
}
GSP_DEV void mbar_wait(mbar_t* b, uint32_t parity) {
  while (!mbar_try_wait(b, parity)) {
  }
}
GSP_DEV void bulk_g2s(void* dst, const void* src, uint32_t bytes, mbar_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(b))
               : "memory");
}
#endif

// ------------------------------------------------------------------ TMA tensor copies (3-D tiled, no swizzle)
#ifdef GSP_EMU
struct TensorMap {
  const unsigned char* base;
  unsigned long long dims[3];     // extents in elements, dim 0 fastest
  unsigned long long strides[3];  // byte strides (strides[0] = element size)
  unsigned box[3];
  unsigned esize;
};
#define GSP_GRID_CONSTANT
// dst (shared, dense box, dim 0 fastest) <- tile at element coordinates (c0, c1, c2); out-of-bounds elements are zero
GSP_DEV void tma_load_3d(void* dst, const TensorMap* tm, int c0, int c1, int c2, mbar_t* b) {
  unsigned char* d = (unsigned char*)dst;
  size_t bytes = 0;
  for (unsigned k = 0; k < tm->box[2]; ++k)
    for (unsigned j = 0; j < tm->box[1]; ++j)
      for (unsigned i = 0; i < tm->box[0]; ++i) {
        const long long x = c0 + (long long)i, y = c1 + (long long)j, z = c2 + (long long)k;
        const bool in = x >= 0 && y >= 0 && z >= 0 && (unsigned long long)x < tm->dims[0] && (unsigned long long)y < tm->dims[1] &&
                        (unsigned long long)z < tm->dims[2];
        if (in)
          memcpy(d + bytes, tm->base + x * tm->strides[0] + y * tm->strides[1] + z * tm->strides[2], tm->esize);
        else
          memset(d + bytes, 0, tm->esize);
        bytes += tm->esize;
      }
  b->tx -= (int)bytes;
  mbar_check_(b);
}
#else
}  // namespace gsp
#include <cuda.h>
namespace gsp {
typedef CUtensorMap TensorMap;
#define GSP_GRID_CONSTANT __grid_constant__
GSP_DEV void tma_load_3d(void* dst, const TensorMap* tm, int c0, int c1, int c2, mbar_t* b) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(tm), "r"(smem_u32(b)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
#endif
// host: FP64 tensor `base` with extents dims[3] (elements), byte strides of dims 1 and 2, tile box[3].  Returns 0 on success.
int make_tensor_map_f64(TensorMap* out, const void* base, const unsigned long long dims[3], unsigned long long stride1_bytes,
                        unsigned long long stride2_bytes, const unsigned box[3]);

// named barrier over `count` threads of the CTA (a subset of warps, e.g. the consumer warps of a producer/consumer kernel)
GSP_DEV void named_bar_sync(int id, int count) {
#ifdef GSP_EMU
  emu::named_bar_sync(id, count);
#else
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
#endif
}

// inter-CTA flags in global memory (persistent kernels): release after the CTA's stores, acquire before dependent loads
GSP_DEV void spin_pause() {
#ifdef GSP_EMU
  emu::spin_yield();
#else
  __nanosleep(64);
#endif
}
GSP_DEV int ld_acquire_gpu(const int* p) {
#ifdef GSP_EMU
  return *p;
#else
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
#endif
}
GSP_DEV void red_release_gpu_add(int* p, int v) {
#ifdef GSP_EMU
  *p += v;
  emu::note_event();
#else
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}

// streaming (read-once / write-once) 16-byte global accesses
GSP_DEV double2 ld_stream2(const double* p) {
#ifdef GSP_EMU
  return double2{p[0], p[1]};
#else
  double2 v;
  asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
#endif
}
GSP_DEV void st_stream2(double* p, double2 v) {
#ifdef GSP_EMU
  p[0] = v.x;
  p[1] = v.y;
#else
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
#endif
}

}  // namespace gsp
