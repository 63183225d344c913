// Context management, argument marshalling and the small standalone entry points of the C ABI.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>

#include "chol.h"
#include "cov.cuh"

namespace gsp {

long long g_launches = 0;
Prof g_prof;

void Prof::add(const std::string& n, double ms) {
  for (auto& a : acc)
    if (a.first == n) {
      a.second.first += ms;
      a.second.second += 1;
      return;
    }
  acc.push_back({n, {ms, 1}});
}

static bool prof_timeline() {
  static int timeline = -1;
  if (timeline < 0) {
    const char* env = getenv("GSP_PROF_TIMELINE");
    timeline = (env && env[0] == '1') ? 1 : 0;
  }
  return timeline != 0;
}

void Prof::mark_reference(const std::vector<int>& devs) {
  if (!on || !prof_timeline()) return;
  flush();
  for (auto& r : ref) cudaEventDestroy(r.second);
  ref.clear();
  for (int d : devs) {
    cudaSetDevice(d);
    cudaDeviceSynchronize();
  }
  for (int d : devs) {
    cudaSetDevice(d);
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, (cudaStream_t)0);
    ref.push_back({d, e});
  }
  for (int d : devs) {
    cudaSetDevice(d);
    cudaDeviceSynchronize();
  }
}

void Prof::flush() {
  // GSP_PROF_TIMELINE=1 (development): print "TL dev stream name start_ms duration_ms" of every launch; start is relative to the
  // device's origin event (mark_reference) or, without one, to the first pending launch of that device
  const bool timeline = prof_timeline();
  for (auto& p : pending) {
    cudaSetDevice(p.dev);
    cudaEventSynchronize(p.e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, p.e0, p.e1);
    if (timeline) {
      cudaEvent_t origin = nullptr;
      for (auto& r : ref)
        if (r.first == p.dev) origin = r.second;
      if (!origin)
        for (auto& q : pending)
          if (q.dev == p.dev) {
            origin = q.e0;
            break;
          }
      float t0 = 0.f;
      cudaEventElapsedTime(&t0, origin, p.e0);
      fprintf(stderr, "TL %d %p %s %.4f %.4f\n", p.dev, (void*)p.st, p.name.c_str(), t0, ms);
    }
    add(p.name, ms);
  }
  for (auto& p : pending) {
    cudaEventDestroy(p.e0);
    cudaEventDestroy(p.e1);
  }
  pending.clear();
}

int make_tensor_map_f64(TensorMap* out, const void* base, const unsigned long long dims[3], unsigned long long stride1_bytes,
                        unsigned long long stride2_bytes, const unsigned box[3]) {
#ifdef GSP_EMU
  out->base = (const unsigned char*)base;
  out->esize = 8;
  for (int i = 0; i < 3; ++i) {
    out->dims[i] = dims[i];
    out->box[i] = box[i];
  }
  out->strides[0] = 8;
  out->strides[1] = stride1_bytes;
  out->strides[2] = stride2_bytes;
  return 0;
#else
  // resolved at run time so that the library carries no link-time dependency on libcuda
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess || !sym) return -1;
    fn = (encode_fn)sym;
  }
  cuuint64_t gdim[3] = {dims[0], dims[1], dims[2]};
  cuuint64_t gstr[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t bx[3] = {box[0], box[1], box[2]};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
#endif
}

int set_err(gsp_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}

int make_cov_dev(gsp_ctx* ctx, const gsp_cov_model* cov, int dim, int argpos, CovDev* out) {
  if (!cov || !cov->structs) return set_err(ctx, -argpos, "covariance model is NULL");
  if (cov->nstruct < 1 || cov->nstruct > GSP_MAX_STRUCTS) return set_err(ctx, -argpos, "nstruct must be in 1..GSP_MAX_STRUCTS");
  std::memset(out, 0, sizeof(*out));
  out->nstruct = cov->nstruct;
  out->dim = dim;
  for (int s = 0; s < cov->nstruct; ++s) {
    const gsp_structure& st = cov->structs[s];
    if (st.kind < GSP_NUGGET || st.kind > GSP_MATERN) return set_err(ctx, -argpos, "unknown structure kind");
    if (st.kind == GSP_MATERN) {
      if (!(st.param > 0.0) || !(st.param <= 100.0)) return set_err(ctx, -argpos, "Matern order must be in (0, 100]");
      out->param[s] = st.param;
      out->aux[s] = std::exp2(1.0 - st.param) / std::tgamma(st.param);  // 2^(1-nu) / Gamma(nu)
    }
    if (!(st.sill >= 0.0)) return set_err(ctx, -argpos, "structure sill must be >= 0");
    out->kind[s] = st.kind;
    out->sill[s] = st.sill;
    for (int i = 0; i < 9; ++i) {
      if (!std::isfinite(st.A[i])) return set_err(ctx, -argpos, "structure metric is not finite");
      out->A[s][i] = st.A[i];
    }
  }
  return GSP_OK;
}

double cov_sill(const CovDev& m) {
  double s = 0.0;
  for (int i = 0; i < m.nstruct; ++i) s += m.sill[i];
  return s;
}

// NOTE: for kind 0 the caller must replace out->coords by a device copy before launching kernels.
int make_dom_dev(gsp_ctx* ctx, const gsp_domain* dom, int argpos, DomDev* out) {
  if (!dom) return set_err(ctx, -argpos, "domain is NULL");
  if (dom->dim < 1 || dom->dim > 3) return set_err(ctx, -argpos, "domain dim must be 1..3");
  std::memset(out, 0, sizeof(*out));
  out->kind = dom->kind;
  out->dim = dom->dim;
  if (dom->kind == 1) {
    long long n = 1;
    for (int a = 0; a < 3; ++a) {
      out->dims[a] = a < dom->dim ? dom->dims[a] : 1;
      out->origin[a] = a < dom->dim ? dom->origin[a] : 0.0;
      out->spacing[a] = a < dom->dim ? dom->spacing[a] : 1.0;
      if (out->dims[a] < 1) return set_err(ctx, -argpos, "grid extents must be >= 1");
      if (a < dom->dim && !(out->spacing[a] > 0.0)) return set_err(ctx, -argpos, "grid spacing must be > 0");
      n *= out->dims[a];
    }
    if (dom->nelems != 0 && dom->nelems != n) return set_err(ctx, -argpos, "nelems does not match prod(dims)");
    out->nelems = n;
  } else if (dom->kind == 0) {
    if (dom->nelems < 1 || !dom->coords) return set_err(ctx, -argpos, "point domain needs coords and nelems >= 1");
    out->nelems = dom->nelems;
    out->coords = dom->coords;
    for (int a = 0; a < 3; ++a) out->dims[a] = 1;
  } else {
    return set_err(ctx, -argpos, "domain kind must be 0 (points) or 1 (grid)");
  }
  return GSP_OK;
}

}  // namespace gsp

using namespace gsp;

extern "C" const char* gsp_version(void) {
#ifdef GSP_EMU
  return "gsp_b200 0.1.0 (CPU EMULATION BUILD - tests only)";
#else
  return "gsp_b200 0.1.0 (sm_100a)";
#endif
}

extern "C" int gsp_ctx_create(int32_t ndev, const int32_t* devs, gsp_ctx** out) {
  if (!out) return -3;
  *out = nullptr;
  if (ndev < 1 || ndev > 64) return -1;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count < 1) return GSP_E_CUDA;
  gsp_ctx* ctx = new gsp_ctx;
  for (int i = 0; i < ndev; ++i) {
    DevCtx dc;
    dc.dev = devs ? devs[i] : i;
    if (dc.dev < 0 || dc.dev >= count) {
      delete ctx;
      return -2;
    }
    ctx->devs.push_back(dc);
  }
  for (auto& dc : ctx->devs) {
    cudaError_t e = cudaSetDevice(dc.dev);
    int prio_lo = 0, prio_hi = 0;
#ifndef GSP_EMU
    if (e == cudaSuccess) {
      // keep freed device memory in the default pool (DevBuf); GSP_MEMPOOL_KEEP=<bytes> lowers the threshold, 0 = return it at once
      cudaMemPool_t pool = nullptr;
      if (cudaDeviceGetDefaultMemPool(&pool, dc.dev) == cudaSuccess && pool) {
        unsigned long long keep = ~0ull;
        const char* env = getenv("GSP_MEMPOOL_KEEP");
        if (env && env[0]) keep = strtoull(env, nullptr, 10);
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      cudaGetLastError();
    }
#endif
    if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&dc.stream, cudaStreamNonBlocking, prio_hi);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&dc.aux, cudaStreamNonBlocking, (prio_hi + prio_lo) / 2);  // between the chain and the wide updates
    for (int k = 0; k < DevCtx::kSide && e == cudaSuccess; ++k) e = cudaStreamCreateWithPriority(&dc.side[k], cudaStreamNonBlocking, prio_lo);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&dc.h2d, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&dc.d2h, cudaStreamNonBlocking);
    cudaDeviceProp prop;
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, dc.dev);
    if (e != cudaSuccess) {
      delete ctx;
      return GSP_E_CUDA;
    }
    dc.sms = prop.multiProcessorCount;
    dc.smem_optin = prop.sharedMemPerBlockOptin;
  }
  // peer access between all device pairs (panel broadcasts of the multi-GPU factorization, copies of L / d2)
  for (auto& a : ctx->devs)
    for (auto& b : ctx->devs) {
      if (a.dev == b.dev) continue;
      int can = 0;
      const bool asked = cudaDeviceCanAccessPeer(&can, a.dev, b.dev) == cudaSuccess;
#ifndef GSP_EMU
      if (!asked || !can) ctx->peer_ok = false;  // no NVLink / PCIe peer path: the factorization then runs on device 0 and L is copied
#endif
      if (asked && can) {
        cudaSetDevice(a.dev);
        cudaDeviceEnablePeerAccess(b.dev, 0);
        cudaGetLastError();  // "already enabled" is fine
#ifndef GSP_EMU
        // stream-ordered pool memory is not covered by cudaDeviceEnablePeerAccess: device a may map what b's pool hands out
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, b.dev) == cudaSuccess && pool) {
          // blocks an earlier context of this process left cached in the pool are released first: granting peer access on a pool
          // that held ~20 GB of cached blocks was followed by "out of memory" on the next 2 GB request (170 GB free)
          cudaSetDevice(b.dev);
          cudaDeviceSynchronize();
          cudaMemPoolTrimTo(pool, 0);
          cudaSetDevice(a.dev);
          cudaMemAccessDesc desc{};
          desc.location.type = cudaMemLocationTypeDevice;
          desc.location.id = a.dev;
          desc.flags = cudaMemAccessFlagsProtReadWrite;
          cudaMemPoolSetAccess(pool, &desc, 1);
        }
        cudaGetLastError();
#endif
      }
    }
  *out = ctx;
  return GSP_OK;
}

extern "C" int gsp_ctx_destroy(gsp_ctx* ctx) {
  if (!ctx) return GSP_OK;
  for (auto& dc : ctx->devs) {
    cudaSetDevice(dc.dev);
    if (dc.stream) cudaStreamDestroy(dc.stream);
    for (int k = 0; k < DevCtx::kSide; ++k)
      if (dc.side[k]) cudaStreamDestroy(dc.side[k]);
    if (dc.h2d) cudaStreamDestroy(dc.h2d);
    if (dc.d2h) cudaStreamDestroy(dc.d2h);
#ifndef GSP_EMU
    cudaMemPool_t pool = nullptr;  // hand the cached blocks back to the driver
    if (cudaDeviceGetDefaultMemPool(&pool, dc.dev) == cudaSuccess && pool) {
      cudaDeviceSynchronize();
      cudaMemPoolTrimTo(pool, 0);
    }
    cudaGetLastError();
#endif
  }
  delete ctx;
  return GSP_OK;
}

extern "C" const char* gsp_last_error(gsp_ctx* ctx) { return ctx ? ctx->err.c_str() : "context is NULL"; }
extern "C" int gsp_ctx_ndev(gsp_ctx* ctx) { return ctx ? (int)ctx->devs.size() : 0; }
extern "C" int64_t gsp_kernel_launches(void) { return g_launches; }
extern "C" double gsp_last_sample_ms(gsp_ctx* ctx) { return ctx ? ctx->last_sample_ms : 0.0; }

extern "C" int gsp_profile_enable(gsp_ctx* ctx, int32_t on) {
  if (!ctx) return -1;
  std::lock_guard<std::mutex> lk(ctx->mu);
  g_prof.flush();
  g_prof.on = on != 0;
  if (on) g_prof.acc.clear();
  return GSP_OK;
}

extern "C" int64_t gsp_profile_read(gsp_ctx* ctx, char* buf, int64_t buflen) {
  if (!ctx || !buf || buflen < 2) return -1;
  std::lock_guard<std::mutex> lk(ctx->mu);
  g_prof.flush();
  std::string js = "{";
  bool first = true;
  for (auto& a : g_prof.acc) {
    char tmp[256];
    snprintf(tmp, sizeof(tmp), "%s\"%s\": {\"ms\": %.6f, \"launches\": %lld}", first ? "" : ", ", a.first.c_str(), a.second.first,
             a.second.second);
    js += tmp;
    first = false;
  }
  js += "}";
  if ((int64_t)js.size() + 1 > buflen) return -(int64_t)js.size() - 1;
  std::memcpy(buf, js.c_str(), js.size() + 1);
  return (int64_t)js.size();
}

extern "C" int gsp_host_alloc(void** out, int64_t bytes) {
  if (!out || bytes < 0) return -1;
  return cudaMallocHost(out, (size_t)bytes) == cudaSuccess ? GSP_OK : GSP_E_NOMEM;
}
extern "C" int gsp_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? GSP_OK : GSP_E_CUDA; }

extern "C" int gsp_pairwise(gsp_ctx* ctx, const gsp_cov_model* cov, int32_t dim, int64_t n1, const double* X1, int64_t n2,
                            const double* X2, double* out) {
  if (!ctx) return -1;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (dim < 1 || dim > 3) return set_err(ctx, -3, "dim must be 1..3");
  if (n1 < 1 || !X1) return set_err(ctx, -4, "X1 / n1 invalid");
  if (!out) return set_err(ctx, -8, "out is NULL");
  const bool sym = (X2 == nullptr);
  if (sym) n2 = n1;
  if (n2 < 1) return set_err(ctx, -6, "n2 invalid");
  CovDev cd;
  GSP_TRY(make_cov_dev(ctx, cov, dim, 2, &cd));
  DevCtx& dc = ctx->devs[0];
  cudaSetDevice(dc.dev);
  DevBuf d1, d2, dout;
  GSP_CUDA_OK(ctx, d1.alloc(dc.dev, (size_t)n1 * dim * sizeof(double)));
  GSP_CUDA_OK(ctx, cudaMemcpyAsync(d1.p, X1, (size_t)n1 * dim * sizeof(double), cudaMemcpyHostToDevice, dc.stream));
  if (!sym) {
    GSP_CUDA_OK(ctx, d2.alloc(dc.dev, (size_t)n2 * dim * sizeof(double)));
    GSP_CUDA_OK(ctx, cudaMemcpyAsync(d2.p, X2, (size_t)n2 * dim * sizeof(double), cudaMemcpyHostToDevice, dc.stream));
  }
  GSP_CUDA_OK(ctx, dout.alloc(dc.dev, (size_t)n1 * n2 * sizeof(double)));
  DomDev r{}, c{};
  r.kind = 0; r.dim = dim; r.nelems = n1; r.coords = d1.as<double>();
  c.kind = 0; c.dim = dim; c.nelems = n2; c.coords = sym ? d1.as<double>() : d2.as<double>();
  launch_assemble(dc.stream, cd, r, c, nullptr, nullptr, n1, n2, dout.as<double>(), n1, false);
  GSP_CUDA_OK(ctx, cudaGetLastError());
  GSP_CUDA_OK(ctx, cudaMemcpyAsync(out, dout.p, (size_t)n1 * n2 * sizeof(double), cudaMemcpyDeviceToHost, dc.stream));
  GSP_CUDA_OK(ctx, cudaStreamSynchronize(dc.stream));
  return GSP_OK;
}

extern "C" int gsp_potrf(gsp_ctx* ctx, int64_t n, double* A) {
  if (!ctx) return -1;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (n < 1) return set_err(ctx, -2, "n must be >= 1");
  if (!A) return set_err(ctx, -3, "A is NULL");
  DevCtx& dc = ctx->devs[0];
  cudaSetDevice(dc.dev);
  const long long np = round_up(n, 128);
  const int nb = (int)(np / 128);
  DevBuf dA, dinv, dinfo;
  GSP_CUDA_OK(ctx, dA.alloc(dc.dev, (size_t)np * np * sizeof(double)));
  GSP_CUDA_OK(ctx, dinv.alloc(dc.dev, (size_t)nb * 128 * 128 * sizeof(double)));
  GSP_CUDA_OK(ctx, dinfo.alloc(dc.dev, sizeof(int)));
  // identity padding, then the user matrix in the leading n x n corner
  std::vector<double> pad((size_t)np * np, 0.0);
  for (long long j = 0; j < np; ++j) {
    if (j < n) std::memcpy(&pad[(size_t)j * np], A + (size_t)j * n, (size_t)n * sizeof(double));
    else pad[(size_t)j * np + j] = 1.0;
  }
  GSP_CUDA_OK(ctx, cudaMemcpyAsync(dA.p, pad.data(), pad.size() * sizeof(double), cudaMemcpyHostToDevice, dc.stream));
  const char* algo = getenv("GSP_CHOL_ALGO");
  if (algo && algo[0] == 'p') {  // the panel algorithm of the distributed factorization, on this one device
    DevBuf rows, flags;
    GSP_CUDA_OK(ctx, rows.alloc(dc.dev, (size_t)nb * sizeof(int)));
    GSP_CUDA_OK(ctx, flags.alloc(dc.dev, (size_t)chol_dist_flag_ints() * sizeof(int)));
    const int pb_env = getenv("GSP_CHOL_PB") ? atoi(getenv("GSP_CHOL_PB")) : 0;
    std::vector<DistDev> dv{DistDev{dc.dev, dc.stream, dc.aux, dc.side[0], dA.as<double>(), dinv.as<double>(), dinfo.as<int>(), rows.as<int>(), flags.as<int>()}};
    GSP_CUDA_OK(ctx, chol_factor_dist(dv, np, nb, std::min(nb, pb_env > 0 ? std::min(pb_env, 8) : 4)));
  } else {
    GSP_CUDA_OK(ctx, chol_factor(dc.stream, dc.side, DevCtx::kSide, dA.as<double>(), np, nb, dinv.as<double>(), dinfo.as<int>()));
  }
  int info = 0;
  GSP_CUDA_OK(ctx, cudaMemcpyAsync(&info, dinfo.p, sizeof(int), cudaMemcpyDeviceToHost, dc.stream));
  GSP_CUDA_OK(ctx, cudaMemcpyAsync(pad.data(), dA.p, pad.size() * sizeof(double), cudaMemcpyDeviceToHost, dc.stream));
  GSP_CUDA_OK(ctx, cudaStreamSynchronize(dc.stream));
  if (info > 0 && info <= n) return info;
  for (long long j = 0; j < n; ++j)
    for (long long i = 0; i < n; ++i) A[(size_t)j * n + i] = (i >= j) ? pad[(size_t)j * np + i] : 0.0;
  return GSP_OK;
}
