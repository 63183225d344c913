// Ensemble statistics on the device - src/ensembles.jl:42-52 and the `ereduce` loop :76-85.
//
// The reference reduces over realizations with scalar generator loops on host tables (O(n R) fetches per statistic).
// Here the realizations never leave HBM: Z is n x R (realization index slowest), sharded over the devices of the
// context, and every statistic is one streaming pass (mean / var / cdf / ccdf: 8 n R bytes read, HBM-bound) or one
// shared-memory sort per node tile (quantile).  Only the n-vector of results crosses PCIe.
#include <cmath>
#include <memory>

#include "ensemble.h"

namespace gsp {

// ------------------------------------------------------------------ kernels
// Partial moments of every node over the nr local realizations.  Shifted sums (shift = the node's first value) keep
// sum((x-K)^2) - sum(x-K)^2 / nr well conditioned; partials of different devices are merged with Chan's update.
__global__ void __launch_bounds__(256) ens_moments_kernel(const double* __restrict__ Z, long long n, long long nr, double* __restrict__ mean,
                                                          double* __restrict__ m2) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double* p = Z + i;
    const double K = p[0];
    double s1a = 0.0, s2a = 0.0, s1b = 0.0, s2b = 0.0;
    long long r = 1;
    for (; r + 3 < nr; r += 4) {  // four independent loads in flight per thread
      const double x0 = p[r * n], x1 = p[(r + 1) * n], x2 = p[(r + 2) * n], x3 = p[(r + 3) * n];
      const double d0 = x0 - K, d1 = x1 - K, d2 = x2 - K, d3 = x3 - K;
      s1a += d0 + d2;
      s1b += d1 + d3;
      s2a += d0 * d0 + d2 * d2;
      s2b += d1 * d1 + d3 * d3;
    }
    for (; r < nr; ++r) {
      const double d = p[r * n] - K;
      s1a += d;
      s2a += d * d;
    }
    const double s1 = s1a + s1b, s2 = s2a + s2b;
    mean[i] = K + s1 / (double)nr;
    m2[i] = s2 - s1 * s1 / (double)nr;
  }
}

// (mean, m2) of na values  <-  merged with (mean_b, m2_b) of nb values
__global__ void __launch_bounds__(256) ens_merge_kernel(double* __restrict__ mean, double* __restrict__ m2, const double* __restrict__ mean_b,
                                                        const double* __restrict__ m2_b, long long n, double na, double nb) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const double nt = na + nb;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double d = mean_b[i] - mean[i];
    mean[i] += d * (nb / nt);
    m2[i] += m2_b[i] + d * d * (na * nb / nt);
  }
}

// v = (a + b * v) / c: a true division like the reference's count / length and sum / (n - 1), so that integer counts give
// bit-identical frequencies
__global__ void __launch_bounds__(256) ens_scale_kernel(double* __restrict__ v, long long n, double a, double b, double c) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
#ifdef GSP_EMU
    v[i] = (a + b * v[i]) / c;
#else
    v[i] = __ddiv_rn(__fma_rn(b, v[i], a), c);  // explicit: nvcc otherwise rewrites the loop as fma(v, b/c, a/c) (seen in SASS)
#endif
  }
}

__global__ void __launch_bounds__(256) ens_add_kernel(double* __restrict__ v, const double* __restrict__ w, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) v[i] += w[i];
}

// cnt[i] = #{r : Z[i, r] <= x}   (count(<=(x), vals), ensembles.jl:46)
__global__ void __launch_bounds__(256) ens_count_kernel(const double* __restrict__ Z, long long n, long long nr, double x, double* __restrict__ cnt) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double* p = Z + i;
    int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    long long r = 0;
    for (; r + 3 < nr; r += 4) {  // four independent loads in flight per thread, compared afterwards
      const double x0 = p[r * n], x1 = p[(r + 1) * n], x2 = p[(r + 2) * n], x3 = p[(r + 3) * n];
      c0 += x0 <= x;
      c1 += x1 <= x;
      c2 += x2 <= x;
      c3 += x3 <= x;
    }
    for (; r < nr; ++r) c0 += p[r * n] <= x;
    cnt[i] = (double)(c0 + c1 + c2 + c3);
  }
}

// quantile(vals, p) per node, Julia's default definition (Statistics.quantile, alpha = beta = 1 = R type 7), ensembles.jl:50.
// One CTA sorts a tile of T nodes x Rp values (Rp = R rounded up to a power of two, padded with +inf) in shared memory with a
// bitonic network; element (k, node) lives at s[k * T + node], so threads with consecutive ids touch consecutive nodes:
// conflict-free shared accesses and contiguous T*8-byte global runs.
__global__ void __launch_bounds__(256) ens_quantile_kernel(const double* __restrict__ Z, long long ldz, long long nloc, long long R, int Rp, int T,
                                                           int np, const double* __restrict__ ps, double* __restrict__ out, long long ldo) {
  GSP_DYN_SMEM(smem);
  double* s = reinterpret_cast<double*>(smem);
  const int tid = threadIdx.x, nth = blockDim.x;
  const int lt = 31 - __clz(T);
  const long long ntiles = (nloc + T - 1) / T;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long i0 = tile * T;
    const int nt = (int)((nloc - i0 < T) ? nloc - i0 : T);
    for (int e = tid; e < T * Rp; e += nth) {
      const int node = e & (T - 1), k = e >> lt;  // T and Rp are powers of two: no integer divisions in the hot loops
      s[e] = (k < R && node < nt) ? Z[(i0 + node) + (long long)k * ldz] : INFINITY;
    }
    __syncthreads();
    const int half = T * (Rp / 2);
    for (int size = 2; size <= Rp; size <<= 1) {
      for (int stride = size >> 1, ls = 31 - __clz(size >> 1); stride > 0; stride >>= 1, --ls) {
        for (int q = tid; q < half; q += nth) {
          const int node = q & (T - 1), h = q >> lt;
          const int lo = ((h >> ls) << (ls + 1)) | (h & (stride - 1)), hi = lo + stride;
          const bool asc = (lo & size) == 0;
          const double a = s[lo * T + node], b = s[hi * T + node];
          if ((a > b) == asc) {
            s[lo * T + node] = b;
            s[hi * T + node] = a;
          }
        }
        __syncthreads();
      }
    }
    for (int e = tid; e < nt * np; e += nth) {
      const int node = e % nt, k = e / nt;
      const double p = ps[k];
      // Statistics.quantile: m = alpha + p (1 - alpha - beta) = 1 - p;  aleph = n p + m;  j = clamp(trunc(aleph), 1, n-1)
      const double m = 1.0 + p * (1.0 - 1.0 - 1.0);
      const double aleph = (double)R * p + m;
      long long j = (long long)aleph;
      if (j > R - 1) j = R - 1;
      if (j < 1) j = 1;
      double g = aleph - (double)j;
      g = g < 0.0 ? 0.0 : (g > 1.0 ? 1.0 : g);
      double a, b;
      if (R == 1) {
        a = b = s[node];
      } else {
        a = s[(j - 1) * T + node];
        b = s[j * T + node];
      }
      out[(i0 + node) + (long long)k * ldo] = a + g * (b - a);
    }
    __syncthreads();
  }
}

namespace {

inline unsigned grid_1d(long long n, int sms) {
  long long b = (n + 255) / 256;
  if (b > (long long)sms * 16) b = (long long)sms * 16;
  if (b < 1) b = 1;
  return (unsigned)b;
}

int sync_all(gsp_ensemble* e) {
  for (auto& d : e->dev) {
    cudaSetDevice(d->dc->dev);
    GSP_CUDA_OK(e->ctx, cudaStreamSynchronize(d->dc->stream));
  }
  return GSP_OK;
}

// per-device (mean, m2) of the local realizations, merged on device 0; on return mean0/m20 (device 0) hold the ensemble's
int moments_on_device0(gsp_ensemble* e, DevBuf* mean0, DevBuf* m20) {
  gsp_ctx* ctx = e->ctx;
  const long long n = e->n;
  std::vector<std::unique_ptr<DevBuf>> pm, p2;
  for (auto& d : e->dev) {
    pm.emplace_back(new DevBuf);
    p2.emplace_back(new DevBuf);
    if (d->nr == 0) continue;
    cudaSetDevice(d->dc->dev);
    DevBuf* m = d.get() == e->dev[0].get() ? mean0 : pm.back().get();
    DevBuf* q = d.get() == e->dev[0].get() ? m20 : p2.back().get();
    GSP_CUDA_OK(ctx, m->alloc(d->dc->dev, (size_t)n * sizeof(double)));
    GSP_CUDA_OK(ctx, q->alloc(d->dc->dev, (size_t)n * sizeof(double)));
    ProfScope prof_("ens_moments", d->dc->stream);
    GSP_LAUNCH(ens_moments_kernel, dim3(grid_1d(n, d->dc->sms)), dim3(256), 0, d->dc->stream, (const double*)d->Z.as<double>(), n, d->nr,
               m->as<double>(), q->as<double>());
    g_launches++;
    GSP_CUDA_OK(ctx, cudaGetLastError());
  }
  GSP_TRY(sync_all(e));
  EnsDev* d0 = e->dev[0].get();
  cudaSetDevice(d0->dc->dev);
  DevBuf tm, t2;
  double na = (double)d0->nr;
  for (size_t i = 1; i < e->dev.size(); ++i) {
    EnsDev* d = e->dev[i].get();
    if (d->nr == 0) continue;
    if (!tm.p) {
      GSP_CUDA_OK(ctx, tm.alloc(d0->dc->dev, (size_t)n * sizeof(double)));
      GSP_CUDA_OK(ctx, t2.alloc(d0->dc->dev, (size_t)n * sizeof(double)));
    }
    GSP_CUDA_OK(ctx, cudaMemcpyPeerAsync(tm.p, d0->dc->dev, pm[i]->p, d->dc->dev, (size_t)n * sizeof(double), d0->dc->stream));
    GSP_CUDA_OK(ctx, cudaMemcpyPeerAsync(t2.p, d0->dc->dev, p2[i]->p, d->dc->dev, (size_t)n * sizeof(double), d0->dc->stream));
    GSP_LAUNCH(ens_merge_kernel, dim3(grid_1d(n, d0->dc->sms)), dim3(256), 0, d0->dc->stream, mean0->as<double>(), m20->as<double>(),
               (const double*)tm.as<double>(), (const double*)t2.as<double>(), n, na, (double)d->nr);
    g_launches++;
    GSP_CUDA_OK(ctx, cudaGetLastError());
    na += (double)d->nr;
  }
  GSP_CUDA_OK(ctx, cudaStreamSynchronize(d0->dc->stream));
  return GSP_OK;
}

int check_nonempty(gsp_ensemble* e, const void* out, int argpos) {
  if (!out) return set_err(e->ctx, -argpos, "output pointer is NULL");
  if (e->R < 1 || e->dev.empty() || e->dev[0]->nr < 1) return set_err(e->ctx, GSP_E_STATE, "the ensemble holds no realizations");
  return GSP_OK;
}

}  // namespace
}  // namespace gsp

using namespace gsp;

extern "C" int gsp_ensemble_create(gsp_ctx* ctx, int64_t n, int64_t R, gsp_ensemble** out) {
  if (!ctx) return -1;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!out) return set_err(ctx, -4, "out is NULL");
  *out = nullptr;
  if (n < 1) return set_err(ctx, -2, "n < 1");
  if (R < 1) return set_err(ctx, -3, "R < 1");
  std::unique_ptr<gsp_ensemble> e(new gsp_ensemble);
  e->ctx = ctx;
  e->n = n;
  e->R = R;
  const int ndev = (int)ctx->devs.size();
  long long r0 = 0;
  for (int i = 0; i < ndev; ++i) {  // same contiguous split as gsp_*_sample (field.jl:103-121 shards over workers)
    std::unique_ptr<EnsDev> d(new EnsDev);
    d->dc = &ctx->devs[i];
    d->r0 = r0;
    d->nr = R / ndev + (i < R % ndev ? 1 : 0);
    r0 += d->nr;
    if (d->nr > 0) {
      cudaError_t er = d->Z.alloc(d->dc->dev, (size_t)n * d->nr * sizeof(double));
      if (er != cudaSuccess) return set_err(ctx, GSP_E_NOMEM, std::string("ensemble: ") + cudaGetErrorString(er));
    }
    e->dev.push_back(std::move(d));
  }
  *out = e.release();
  return GSP_OK;
}

extern "C" int gsp_ensemble_destroy(gsp_ensemble* e) {
  if (!e) return GSP_OK;
  for (auto& d : e->dev) {
    cudaSetDevice(d->dc->dev);
    cudaStreamSynchronize(d->dc->stream);
  }
  delete e;
  return GSP_OK;
}

extern "C" int gsp_ensemble_sizes(gsp_ensemble* e, int64_t sizes[2]) {
  if (!e) return -1;
  if (!sizes) return set_err(e->ctx, -2, "sizes is NULL");
  sizes[0] = e->n;
  sizes[1] = e->R;
  return GSP_OK;
}

namespace {
// realizations [r0, r0 + nr) <-> host buffer (n x nr), walking over the owning devices
int transfer(gsp_ensemble* e, int64_t r0, int64_t nr, double* host, bool to_device) {
  gsp_ctx* ctx = e->ctx;
  if (r0 < 0 || nr < 0 || r0 + nr > e->R) return set_err(ctx, -2, "realization range out of bounds");
  if (!host) return set_err(ctx, -4, "host buffer is NULL");
  for (auto& d : e->dev) {
    const long long a = std::max<long long>(r0, d->r0), b = std::min<long long>(r0 + nr, d->r0 + d->nr);
    if (a >= b) continue;
    cudaSetDevice(d->dc->dev);
    double* dp = d->Z.as<double>() + (a - d->r0) * e->n;
    double* hp = host + (a - r0) * e->n;
    const size_t bytes = (size_t)(b - a) * e->n * sizeof(double);
    if (to_device)
      GSP_CUDA_OK(ctx, cudaMemcpyAsync(dp, hp, bytes, cudaMemcpyHostToDevice, d->dc->stream));
    else
      GSP_CUDA_OK(ctx, cudaMemcpyAsync(hp, dp, bytes, cudaMemcpyDeviceToHost, d->dc->stream));
  }
  return sync_all(e);
}
}  // namespace

extern "C" int gsp_ensemble_put(gsp_ensemble* e, int64_t r0, int64_t nr, const double* Z) {
  if (!e) return -1;
  std::lock_guard<std::mutex> lk(e->mu);
  return transfer(e, r0, nr, const_cast<double*>(Z), true);
}

extern "C" int gsp_ensemble_fetch(gsp_ensemble* e, int64_t r0, int64_t nr, double* Z) {
  if (!e) return -1;
  std::lock_guard<std::mutex> lk(e->mu);
  return transfer(e, r0, nr, Z, false);
}

extern "C" int gsp_ensemble_moments(gsp_ensemble* e, double* mean, double* m2) {
  if (!e) return -1;
  std::lock_guard<std::mutex> lk(e->mu);
  gsp_ctx* ctx = e->ctx;
  GSP_TRY(check_nonempty(e, mean ? (void*)mean : (void*)m2, 2));
  DevBuf dm, d2;
  GSP_TRY(moments_on_device0(e, &dm, &d2));
  EnsDev* d0 = e->dev[0].get();
  cudaSetDevice(d0->dc->dev);
  if (mean) GSP_CUDA_OK(ctx, cudaMemcpyAsync(mean, dm.p, (size_t)e->n * sizeof(double), cudaMemcpyDeviceToHost, d0->dc->stream));
  if (m2) GSP_CUDA_OK(ctx, cudaMemcpyAsync(m2, d2.p, (size_t)e->n * sizeof(double), cudaMemcpyDeviceToHost, d0->dc->stream));
  GSP_CUDA_OK(ctx, cudaStreamSynchronize(d0->dc->stream));
  return GSP_OK;
}

extern "C" int gsp_ensemble_mean(gsp_ensemble* e, double* out) { return gsp_ensemble_moments(e, out, nullptr); }

extern "C" int gsp_ensemble_var(gsp_ensemble* e, double* out) {
  if (!e) return -1;
  std::lock_guard<std::mutex> lk(e->mu);
  gsp_ctx* ctx = e->ctx;
  GSP_TRY(check_nonempty(e, out, 2));
  DevBuf dm, d2;
  GSP_TRY(moments_on_device0(e, &dm, &d2));
  EnsDev* d0 = e->dev[0].get();
  cudaSetDevice(d0->dc->dev);
  // Statistics.var: corrected, m2 / (R - 1); R == 1 gives 0/0 = NaN like the reference
  GSP_LAUNCH(ens_scale_kernel, dim3(grid_1d(e->n, d0->dc->sms)), dim3(256), 0, d0->dc->stream, d2.as<double>(), e->n, 0.0, 1.0,
             (double)e->R - 1.0);
  g_launches++;
  GSP_CUDA_OK(ctx, cudaGetLastError());
  GSP_CUDA_OK(ctx, cudaMemcpyAsync(out, d2.p, (size_t)e->n * sizeof(double), cudaMemcpyDeviceToHost, d0->dc->stream));
  GSP_CUDA_OK(ctx, cudaStreamSynchronize(d0->dc->stream));
  if (e->R == 1)
    for (long long i = 0; i < e->n; ++i) out[i] = NAN;
  return GSP_OK;
}

namespace {
int cdf_impl(gsp_ensemble* e, double x, double* out, bool complement) {
  gsp_ctx* ctx = e->ctx;
  GSP_TRY(check_nonempty(e, out, 3));
  const long long n = e->n;
  std::vector<std::unique_ptr<DevBuf>> cnt;
  for (auto& d : e->dev) {
    cnt.emplace_back(new DevBuf);
    if (d->nr == 0) continue;
    cudaSetDevice(d->dc->dev);
    GSP_CUDA_OK(ctx, cnt.back()->alloc(d->dc->dev, (size_t)n * sizeof(double)));
    ProfScope prof_("ens_count", d->dc->stream);
    GSP_LAUNCH(ens_count_kernel, dim3(grid_1d(n, d->dc->sms)), dim3(256), 0, d->dc->stream, (const double*)d->Z.as<double>(), n, d->nr, x,
               cnt.back()->as<double>());
    g_launches++;
    GSP_CUDA_OK(ctx, cudaGetLastError());
  }
  GSP_TRY(sync_all(e));
  EnsDev* d0 = e->dev[0].get();
  cudaSetDevice(d0->dc->dev);
  DevBuf tmp;
  for (size_t i = 1; i < e->dev.size(); ++i) {
    EnsDev* d = e->dev[i].get();
    if (d->nr == 0) continue;
    if (!tmp.p) GSP_CUDA_OK(ctx, tmp.alloc(d0->dc->dev, (size_t)n * sizeof(double)));
    GSP_CUDA_OK(ctx, cudaMemcpyPeerAsync(tmp.p, d0->dc->dev, cnt[i]->p, d->dc->dev, (size_t)n * sizeof(double), d0->dc->stream));
    GSP_LAUNCH(ens_add_kernel, dim3(grid_1d(n, d0->dc->sms)), dim3(256), 0, d0->dc->stream, cnt[0]->as<double>(), (const double*)tmp.as<double>(), n);
    g_launches++;
    GSP_CUDA_OK(ctx, cudaGetLastError());
  }
  // cdf = count(<= x) / R;  ccdf = count(> x) / R = (R - count(<= x)) / R   (ensembles.jl:46-48)
  const double R = (double)e->R;
  GSP_LAUNCH(ens_scale_kernel, dim3(grid_1d(n, d0->dc->sms)), dim3(256), 0, d0->dc->stream, cnt[0]->as<double>(), n, complement ? R : 0.0,
             complement ? -1.0 : 1.0, R);
  g_launches++;
  GSP_CUDA_OK(ctx, cudaGetLastError());
  GSP_CUDA_OK(ctx, cudaMemcpyAsync(out, cnt[0]->p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, d0->dc->stream));
  GSP_CUDA_OK(ctx, cudaStreamSynchronize(d0->dc->stream));
  return GSP_OK;
}
}  // namespace

extern "C" int gsp_ensemble_cdf(gsp_ensemble* e, double x, double* out) {
  if (!e) return -1;
  std::lock_guard<std::mutex> lk(e->mu);
  return cdf_impl(e, x, out, false);
}

extern "C" int gsp_ensemble_ccdf(gsp_ensemble* e, double x, double* out) {
  if (!e) return -1;
  std::lock_guard<std::mutex> lk(e->mu);
  return cdf_impl(e, x, out, true);
}

extern "C" int gsp_ensemble_quantile(gsp_ensemble* e, int64_t np, const double* ps, double* out) {
  if (!e) return -1;
  std::lock_guard<std::mutex> lk(e->mu);
  gsp_ctx* ctx = e->ctx;
  GSP_TRY(check_nonempty(e, out, 4));
  if (np < 1 || !ps) return set_err(ctx, -3, "ps is NULL or np < 1");
  for (long long k = 0; k < np; ++k)
    if (!(ps[k] >= 0.0 && ps[k] <= 1.0)) return set_err(ctx, -3, "quantile probabilities must be in [0, 1] (Statistics.quantile throws ArgumentError)");
  const long long n = e->n, R = e->R;
  int Rp = 2;
  while (Rp < R) Rp <<= 1;
  const size_t budget = 160 * 1024;
  if ((size_t)Rp * sizeof(double) > budget) return set_err(ctx, GSP_E_UNSUPPORTED, "quantile: more than 20480 realizations per node do not fit one shared-memory tile");
  int T = 32;  // nodes per tile: a power of two (shifts instead of divisions in the kernel), 256-byte global runs at 32
  while (T > 1 && (size_t)T * Rp * sizeof(double) > budget) T >>= 1;
  const size_t smem = (size_t)T * Rp * sizeof(double);
  const int ndev = (int)e->dev.size();
  // node ranges: device i sorts the nodes [b[i], b[i+1]) of ALL realizations; with more than one device the other devices'
  // columns of that range are pulled over NVLink first (2-D peer copies), in node chunks that bound the extra memory
  std::vector<long long> b(ndev + 1, 0);
  for (int i = 0; i < ndev; ++i) b[i + 1] = b[i] + n / ndev + (i < n % ndev ? 1 : 0);
  std::vector<std::unique_ptr<DevBuf>> dps(ndev), gat(ndev), res(ndev);
  const long long chunk_cap = std::max<long long>(1, (1ll << 31) / (R * (long long)sizeof(double)));  // <= 2 GB of gathered values per device
  int rc = GSP_OK;
  for (int i = 0; i < ndev && rc == GSP_OK; ++i) {
    EnsDev* d = e->dev[i].get();
    const long long nloc = b[i + 1] - b[i];
    dps[i].reset(new DevBuf);
    gat[i].reset(new DevBuf);
    res[i].reset(new DevBuf);
    if (nloc == 0) continue;
    cudaSetDevice(d->dc->dev);
    GSP_CUDA_OK(ctx, dps[i]->alloc(d->dc->dev, (size_t)np * sizeof(double)));
    GSP_CUDA_OK(ctx, cudaMemcpyAsync(dps[i]->p, ps, (size_t)np * sizeof(double), cudaMemcpyHostToDevice, d->dc->stream));
    GSP_CUDA_OK(ctx, res[i]->alloc(d->dc->dev, (size_t)nloc * np * sizeof(double)));
    auto kfn = ens_quantile_kernel;
    GSP_CUDA_OK(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (ndev == 1) {
      const long long ntiles = (nloc + T - 1) / T;
      long long grid = std::min<long long>(ntiles, (long long)d->dc->sms * 8);
      ProfScope prof_("ens_quantile", d->dc->stream);
      GSP_LAUNCH(kfn, dim3((unsigned)grid), dim3(256), smem, d->dc->stream, (const double*)d->Z.as<double>(), n, nloc, R, Rp, T, (int)np,
                 (const double*)dps[i]->as<double>(), res[i]->as<double>(), nloc);
      g_launches++;
      GSP_CUDA_OK(ctx, cudaGetLastError());
    } else {
      const long long cw = std::min(nloc, chunk_cap);
      GSP_CUDA_OK(ctx, gat[i]->alloc(d->dc->dev, (size_t)cw * R * sizeof(double)));
      for (long long c0 = 0; c0 < nloc; c0 += cw) {
        const long long w = std::min(cw, nloc - c0);
        for (int j = 0; j < ndev; ++j) {
          EnsDev* s = e->dev[j].get();
          if (s->nr == 0) continue;
          GSP_CUDA_OK(ctx, cudaMemcpy2DAsync(gat[i]->as<double>() + s->r0 * w, (size_t)w * sizeof(double), s->Z.as<double>() + b[i] + c0,
                                             (size_t)n * sizeof(double), (size_t)w * sizeof(double), (size_t)s->nr, cudaMemcpyDefault, d->dc->stream));
        }
        const long long ntiles = (w + T - 1) / T;
        long long grid = std::min<long long>(ntiles, (long long)d->dc->sms * 8);
        ProfScope prof_("ens_quantile", d->dc->stream);
        GSP_LAUNCH(kfn, dim3((unsigned)grid), dim3(256), smem, d->dc->stream, (const double*)gat[i]->as<double>(), w, w, R, Rp, T, (int)np,
                   (const double*)dps[i]->as<double>(), res[i]->as<double>() + c0, nloc);
        g_launches++;
        GSP_CUDA_OK(ctx, cudaGetLastError());
      }
    }
    for (long long k = 0; k < np; ++k)
      GSP_CUDA_OK(ctx, cudaMemcpyAsync(out + b[i] + k * n, res[i]->as<double>() + k * nloc, (size_t)nloc * sizeof(double), cudaMemcpyDeviceToHost,
                                       d->dc->stream));
  }
  GSP_TRY(sync_all(e));
  return rc;
}
