// FFTSIM on the device: spectrum build (fftsim.jl:77-91) and the per-realization
// pipeline (fftsim.jl:124-135) as real-to-complex / complex-to-real passes over a half spectrum.
//
// Layout in HBM: real fields [z][y][x] (x fastest = Julia column-major), half spectrum
// H[z][y][kx] with kx = 0..nx/2 (16-byte complex), F stored as a REAL half spectrum (4N bytes).
// Per realization:  x-pass R2C  ->  (y-pass fwd)  ->  last-axis pass: fwd, P = s*F*W/|W|, inverse
//                   ->  (y-pass inverse)  ->  x-pass C2R with the "+mu" epilogue.
// sigma^2 = var(Z, mean=0) (fftsim.jl:131) does not depend on the noise: |P_k| = F_k, so by Parseval
// sigma^2 = sum(F^2) / (N (N-1)); it is computed once per plan and folded into the scalar s.
#include <cmath>
#include <cstdlib>
#include <memory>

#include "cov.cuh"
#include "chol.h"
#include "fft_kernels.cuh"
#include "fft_pow2.cuh"
#ifdef GSP_EXPERIMENTAL
#include "fft_plane.cuh"
#endif
#include "ensemble.h"
#include "krige.cuh"
#include "rng.cuh"

namespace gsp {

// Variants that were measured SLOWER than the default schedule on the B200 (fused x+y plane kernels through L2, L2 slab schedules,
// 16-kx bundles) are compiled only with -DGSP_EXPERIMENTAL (the test-only emulator build defines it, so their
// index math stays covered); the product library neither contains their kernels nor reads their switches.
#ifdef GSP_EXPERIMENTAL
static inline const char* exp_env(const char* name) { return getenv(name); }
#else
static inline const char* exp_env(const char*) { return nullptr; }
#endif

// ------------------------------------------------------------------ kernels
// x-axis forward pass: real rows -> half spectrum rows.  One CTA = B consecutive rows.
// packed != 0 (even nx): rows are read as nx/2 complex numbers, transformed with a half-length FFT
// and untangled; packed == 0 (odd nx): plain complex transform of the zero-extended row.
__global__ void __launch_bounds__(256) xpass_fwd_kernel(LinePlan lp, int nx, int hx, long long nrows, int B, int packed,
                                                        const double* __restrict__ in, cplx* __restrict__ H) {
  GSP_DYN_SMEM(smem);
  const int L = packed ? nx / 2 + 1 : nx;
  cplx* X = reinterpret_cast<cplx*>(smem);
  cplx* Y = X + (size_t)L * B;
  const long long row0 = (long long)blockIdx.x * B;
  const int nb = (int)((nrows - row0 < B) ? nrows - row0 : B);
  const int n = lp.n;
  for (int idx = threadIdx.x; idx < B * n; idx += blockDim.x) {
    const int b = idx / n, j = idx - b * n;
    cplx v{0.0, 0.0};
    if (b < nb) {
      const double* p = in + (row0 + b) * nx;
      if (packed) {
        const double2 t = *reinterpret_cast<const double2*>(p + 2 * j);
        v = cplx{t.x, t.y};
      } else {
        v = cplx{p[j], 0.0};
      }
    }
    X[j * B + b] = v;
  }
  __syncthreads();
  const cplx* Z = fft_bundle<false>(lp, X, Y, B);
  for (int idx = threadIdx.x; idx < nb * hx; idx += blockDim.x) {
    const int b = idx / hx, k = idx - b * hx;
    cplx o;
    if (packed) {
      const int h = n;
      const cplx zk = Z[(k % h) * B + b];
      const cplx zc = cconj(Z[((h - k) % h) * B + b]);
      const cplx e = cplx{0.5 * (zk.re + zc.re), 0.5 * (zk.im + zc.im)};
      const cplx d = csub(zk, zc);
      const cplx od = cplx{0.5 * d.im, -0.5 * d.re};  // (-i/2) * d
      o = cadd(e, cmul(lp.tw[k], od));                // tw has stride 1 over the length-nx table here
    } else {
      o = Z[k * B + b];
    }
    H[(row0 + b) * hx + k] = o;
  }
}

// x-axis inverse pass: half spectrum rows -> real rows, out = scale * (unnormalised inverse DFT) + mu
__global__ void __launch_bounds__(256) xpass_inv_kernel(LinePlan lp, int nx, int hx, long long nrows, int B, int packed,
                                                        const cplx* __restrict__ H, double* __restrict__ out, double scale,
                                                        double mu) {
  GSP_DYN_SMEM(smem);
  const int L = packed ? nx / 2 + 1 : nx;
  cplx* X = reinterpret_cast<cplx*>(smem);
  cplx* Y = X + (size_t)L * B;
  const long long row0 = (long long)blockIdx.x * B;
  const int nb = (int)((nrows - row0 < B) ? nrows - row0 : B);
  const int n = lp.n;
  for (int idx = threadIdx.x; idx < B * hx; idx += blockDim.x) {
    const int b = idx / hx, k = idx - b * hx;
    X[k * B + b] = (b < nb) ? H[(row0 + b) * hx + k] : cplx{0.0, 0.0};
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < B * n; idx += blockDim.x) {
    const int b = idx / n, k = idx - b * n;
    cplx v;
    if (packed) {
      const int h = n;
      const cplx xk = X[k * B + b];
      const cplx xc = cconj(X[(h - k) * B + b]);
      const cplx s = cadd(xk, xc);
      const cplx d = csub(xk, xc);
      const cplx t = cmul(cconj(lp.tw[k]), d);
      v = cplx{s.re - t.im, s.im + t.re};  // s + i*t
    } else {
      v = (k < hx) ? X[k * B + b] : cconj(X[(nx - k) * B + b]);
    }
    Y[k * B + b] = v;
  }
  __syncthreads();
  const cplx* z = fft_bundle<true>(lp, Y, X, B);
  for (int idx = threadIdx.x; idx < nb * n; idx += blockDim.x) {
    const int b = idx / n, j = idx - b * n;
    const cplx v = z[j * B + b];
    double* p = out + (row0 + b) * nx;
    if (packed)
      *reinterpret_cast<double2*>(p + 2 * j) = make_double2(v.re * scale + mu, v.im * scale + mu);
    else
      p[j] = v.re * scale + mu;
  }
}

enum { PASS_FWD = 1, PASS_MUL = 2, PASS_INV = 4 };

// strided-axis pass (y or z) over the half spectrum, in place.  blockIdx.x = bundle of B adjacent kx,
// blockIdx.y = index along the remaining axis.  With PASS_MUL the spectrum is replaced by
// P = s * F * W/|W| (angle(0) = 0 => P = s*F) between the forward and inverse transforms (fftsim.jl:125).
__global__ void __launch_bounds__(256) strided_pass_kernel(LinePlan lp, cplx* __restrict__ H, long long es, int hx, int B,
                                                           long long other_stride, int flags, const double* __restrict__ Fh,
                                                           long long esF, long long other_strideF, double s) {
  GSP_DYN_SMEM(smem);
  const int n = lp.n;
  cplx* X = reinterpret_cast<cplx*>(smem);
  cplx* Y = X + (size_t)n * B;
  const int b0 = blockIdx.x * B;
  const int nb = (hx - b0 < B) ? hx - b0 : B;
  const long long base = (long long)blockIdx.y * other_stride + b0;
  const long long baseF = (long long)blockIdx.y * other_strideF + b0;
  for (int idx = threadIdx.x; idx < n * B; idx += blockDim.x) {
    const int j = idx / B, b = idx - j * B;
    X[idx] = (b < nb) ? H[base + (long long)j * es + b] : cplx{0.0, 0.0};
  }
  __syncthreads();
  cplx* cur = X;
  cplx* oth = Y;
  if (flags & PASS_FWD) {
    cur = fft_bundle<false>(lp, X, Y, B);
    oth = (cur == X) ? Y : X;
  }
  if (flags & PASS_MUL) {
    for (int idx = threadIdx.x; idx < n * B; idx += blockDim.x) {
      const int j = idx / B, b = idx - j * B;
      if (b < nb) {
        const double f = s * Fh[baseF + (long long)j * esF + b];
        const cplx w = cur[idx];
        const double m2 = w.re * w.re + w.im * w.im;
        if (m2 > 0.0) {
          const double g = f * rsqrt(m2);
          cur[idx] = cplx{g * w.re, g * w.im};
        } else {
          cur[idx] = cplx{f, 0.0};
        }
      }
    }
    __syncthreads();
  }
  if (flags & PASS_INV) cur = fft_bundle<true>(lp, cur, oth, B);
  for (int idx = threadIdx.x; idx < n * B; idx += blockDim.x) {
    const int j = idx / B, b = idx - j * B;
    if (b < nb) H[base + (long long)j * es + b] = cur[idx];
  }
}

// 1-D grids only: P = s * F * W/|W| elementwise on the half spectrum
// (`total` = nh * batch elements; F repeats with period nh)
__global__ void __launch_bounds__(256) spectral_mul_kernel(cplx* __restrict__ H, const double* __restrict__ Fh, long long nh,
                                                           long long total, double s) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const double f = s * Fh[i % nh];
    const cplx w = H[i];
    const double m2 = w.re * w.re + w.im * w.im;
    H[i] = (m2 > 0.0) ? cplx{f * rsqrt(m2) * w.re, f * rsqrt(m2) * w.im} : cplx{f, 0.0};
  }
}

// F = sqrt(|H|), F[0] = 0 (fftsim.jl:90-91); partial[blockIdx.x] = sum of w_k F_k^2 over the block's
// elements, w_k = 1 on the self-conjugate planes kx = 0 and (nx even) kx = nx/2, else 2.
__global__ void __launch_bounds__(256) spectrum_finalize_kernel(const cplx* __restrict__ H, double* __restrict__ Fh, long long nh,
                                                                int hx, int hxF, int nx, double* __restrict__ partial) {
  __shared__ double red[256];
  double acc = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nh; i += stride) {
    const cplx h = H[i];
    double f2 = sqrt(h.re * h.re + h.im * h.im);  // F^2 = |fft(C)|
    if (i == 0) f2 = 0.0;
    const int kx = (int)(i % hx);
    Fh[(i / hx) * hxF + kx] = sqrt(f2);
    const double w = (kx == 0 || ((nx & 1) == 0 && kx == nx / 2)) ? 1.0 : 2.0;
    acc += w * f2;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

__global__ void __launch_bounds__(256) sum_partials_kernel(const double* __restrict__ partial, int n, double* __restrict__ out) {
  __shared__ double red[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) acc += partial[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = red[0];
}

// out[q + r*n_inds] = src[(inds[q]-1) + r*N]   (Z[parentindices(sdom)], fftsim.jl:135; inds 1-based)
__global__ void __launch_bounds__(256) gather_kernel(const double* __restrict__ src, long long N, const long long* __restrict__ inds,
                                                     long long n_inds, long long R, double* __restrict__ out) {
  const long long total = n_inds * R;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const long long r = t / n_inds, q = t - r * n_inds;
    out[t] = src[(inds[q] - 1) + r * N];
  }
}

// ------------------------------------------------------------------ host side
namespace {

bool choose_radices(int n, LinePlan* lp) {
  lp->n = n;
  lp->nst = 0;
  int m = n;
  const int cand[] = {16, 8, 4, 2, 3, 5, 7, 11, 13};
  for (int c : cand) {
    while (m % c == 0 && m > 1) {
      if (lp->nst >= FFT_MAX_STAGES) return false;
      lp->radix[lp->nst++] = c;
      m /= c;
    }
  }
  // what is left is a product of primes > 13: one generic stage per prime (fft_stage_generic)
  for (int q = 17; m > 1; q += 2) {
    if ((long long)q * q > m) q = m;
    while (m % q == 0) {
      if (lp->nst >= FFT_MAX_STAGES) return false;
      lp->radix[lp->nst++] = q;
      m /= q;
    }
  }
  return m == 1;
}

// per-stage tables of fft_pow2.cuh: stage s >= 1 of the length-n transform, w[r][k] = exp(-2*pi*i*r*k/(Ns*R)), k fastest
std::vector<cplx> make_stage_twiddles(int n, bool inv) {
  std::vector<cplx> t((size_t)p2_stw_size(n, inv) + 1);
  const long double two_pi = 6.283185307179586476925286766559005768L;
  for (int s = 1; s < p2_stages(n); ++s) {
    const int R = p2_radix(n, inv, s), Ns = p2_ns(n, inv, s), off = p2_stw_offset(n, inv, s);
    for (int r = 0; r < R; ++r)
      for (int k = 0; k < Ns; ++k) {
        const long double a = two_pi * (long double)(r * k) / (long double)(Ns * R);
        t[(size_t)off + (size_t)r * Ns + k] = cplx{(double)cosl(a), (double)-sinl(a)};
      }
  }
  return t;
}

std::vector<cplx> make_twiddles(int n) {
  std::vector<cplx> t((size_t)n);
  const long double two_pi = 6.283185307179586476925286766559005768L;
  for (int i = 0; i < n; ++i) {
    // reduce to the first octant for accuracy
    long double a = two_pi * (long double)i / (long double)n;
    t[i] = cplx{(double)cosl(a), (double)-sinl(a)};
  }
  return t;
}

struct AxisPlan {
  int len = 1;       // extent
  LinePlan lp{};     // device-visible plan
  int B = 1;         // bundle width
  size_t smem = 0;
  int packed = 0;    // x axis only
  bool fast = false; // power-of-two register-resident kernels (fft_pow2.cuh)
  TensorMap tmH, tmF; // strided fast passes: TMA tiles of the work spectrum / of F along this axis
  int bundle = 0;     // kx per work item of the fast strided passes along this axis (= inner extent of the TMA box / 2)
};

// One realization in flight: its own stream and half-spectrum work buffer (3-D grids run several lanes concurrently so that
// the small slab kernels of different realizations fill each other's launch gaps and tails).  Lane 0 = the compute stream + d->H.
struct Lane {
  cudaStream_t st = nullptr;
  bool own_stream = false;
  DevBuf Hbuf;
  cplx* H = nullptr;
  TensorMap tmH[3];
  cudaEvent_t done = nullptr;
  DevBuf syncbuf;  // fused x+y kernels: [claim counter, nz plane counters] forward, then the same for the inverse
};

// sub-range of a strided pass: bundles [bx0, bx0 + nb) (nb == 0: all) for the indices [o0, o1) of the remaining axis (o1 == 0: all)
struct Sub {
  int bx0 = 0, nb = 0;
  long long o0 = 0, o1 = 0;
};

// conditioning by Kriging of residuals (krige.cuh): per-node weight table of the second Kriging + the conditional mean
struct CondDev {
  DevBuf zbar;    // n
  DevBuf lam;     // weight table, tile-major: [n / 32][kk][32]
  DevBuf nbr;     // same layout, int32: index into the residual table
  DevBuf knodes;  // nk: 0-based positions of the data nodes within sdom
  DevBuf res;     // nk x chunk residuals
  long long res_cap = 0;
};

struct FftDev {
  DevCtx* dc = nullptr;
  CondDev cond;
  std::vector<std::unique_ptr<Lane>> lanes;
  cudaEvent_t ev_fork = nullptr;
  int slab_mode = 0;     // 0: full-grid passes; 1: z-plane slabs for the x/y pairs; 2: kx-bundle groups for y fwd / z / y inv
  int slab_planes = 32;  // mode 1
  int slab_bundles = 4;  // mode 2
  DevBuf tw[3];
  DevBuf stw_fwd, stw_inv;  // x axis: per-stage twiddle tables of the length-nx/2 transform (forward / inverse radix order)
  DevBuf stw_ax_fwd[3], stw_ax_inv[3];  // strided axes: the same for the full-length transforms
  DevBuf Fh;       // real half spectrum
  DevBuf Fperm;    // 3-D fast grids: F in the item order of the last axis' fused pass (p2_permute_F_kernel)
  DevBuf H;        // complex half-spectrum work buffer
  DevBuf win[2], zout[2];  // staging for host-pointer sampling
  DevBuf inds;
  long long inds_cap = 0;
  bool fused_xy = false;   // 3-D grids whose x and y extents are covered by p2_plane_kernel
  AxisPlan ax[3];
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

}  // namespace
}  // namespace gsp

using namespace gsp;

struct gsp_fft_plan {
  gsp_ctx* ctx = nullptr;
  int ndim = 1;
  long long dims[3] = {1, 1, 1};
  long long N = 1, nh = 1, nhF = 1;
  int hx = 1, hxF = 2;  // F rows are padded to an even length: 16-byte aligned rows for the bulk copies
  double sumF2 = 0.0;  // full-spectrum sum of F^2
  long long rb = 1;    // realizations per launch (1-D / 2-D grids are batched so that small grids still fill the GPU)
  gsp::CovDev cov;     // the model and the parent grid (kept for gsp_fft_plan_condition)
  gsp::DomDev dom;
  // conditional simulation (gsp_fft_plan_condition): z = zbar + (zu - zbaru), fftsim.jl:140-153
  bool cond = false;
  double cond_mu = 0.0;
  long long cond_n = 0, cond_nk = 0, cond_ninds = 0;
  unsigned long long cond_inds_hash = 0;  // FNV-1a of the view's parent indices given to gsp_fft_plan_condition
  int cond_kk = 0;
  std::vector<std::unique_ptr<FftDev>> dev;
  std::mutex mu;
};

namespace gsp {
namespace {

const size_t kMaxSmem = 200 * 1024;
#ifndef GSP_STRIDED_STAGES
#define GSP_STRIDED_STAGES 1  // measured on B200 at 256^3: z pass 93 -> 76 us, y passes unchanged
#endif
#ifndef GSP_FFT_FUSE_DEFAULT
#define GSP_FFT_FUSE_DEFAULT 0
#endif
bool g_force_generic = false;  // GSP_FFT_GENERIC=1: use the mixed-radix kernels for every extent (A/B checks)

// Per-kernel launch setup, done once per device: opt in to the dynamic shared memory and ask the occupancy API how many CTAs
// are resident per SM (at most 4 are used).  Both calls cost microseconds of host time, which the slab schedules (dozens of
// short launches per realization) cannot afford per launch.  SLOT is a distinct static per kernel instantiation.
#ifndef GSP_MAX_PER_SM
#define GSP_MAX_PER_SM 4
#endif
struct KernelSetup {
  int per_sm[32] = {};
  cudaError_t err = cudaSuccess;
  template <class K>
  int get(K kfn, int threads, size_t smem) {
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 31;
    if (per_sm[dev] == 0) {
      err = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (err != cudaSuccess) return 0;
      int n = 1;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kfn, threads, smem) != cudaSuccess || n < 1) n = 1;
      per_sm[dev] = n > GSP_MAX_PER_SM ? GSP_MAX_PER_SM : n;
    }
    return per_sm[dev];
  }
};

// persistent grid: as many CTAs as are resident per SM, never more than there are items
inline unsigned persistent_grid(int per_sm, int sms, long long items) {
  long long g = (long long)per_sm * sms;
  if (g > items) g = items;
  if (g < 1) g = 1;
  return (unsigned)g;
}

template <int HN, bool RNG>
cudaError_t launch_p2_xfwd_t(cudaStream_t st, int sms, const double* in, cplx* H, const cplx* tw, const cplx* stw, long long nrows, const XRng& rng) {
  using C = XCfg<HN, false>;
  auto kfn = p2_xfwd_kernel<HN, RNG>;
  static KernelSetup ks;
  const int per_sm = ks.get(kfn, C::THREADS, C::SMEM);
  if (per_sm == 0) return ks.err;
  const long long ngroups = (nrows + C::ROWS - 1) / C::ROWS;
  ProfScope prof_(RNG ? "fft_xpass_fwd_rng" : "fft_xpass_fwd", st);
  GSP_LAUNCH(kfn, dim3(persistent_grid(per_sm, sms, ngroups)), dim3(C::THREADS), C::SMEM, st, in, H, tw, stw, nrows, rng);
  g_launches++;
  return cudaGetLastError();
}

// in == nullptr: the noise comes from the counter RNG inside the kernel (rng describes which realization / rows)
template <int HN>
cudaError_t launch_p2_xfwd(cudaStream_t st, int sms, const double* in, cplx* H, const cplx* tw, const cplx* stw, long long nrows, const XRng& rng) {
  return in ? launch_p2_xfwd_t<HN, false>(st, sms, in, H, tw, stw, nrows, rng) : launch_p2_xfwd_t<HN, true>(st, sms, in, H, tw, stw, nrows, rng);
}

template <int HN>
cudaError_t launch_p2_xinv(cudaStream_t st, int sms, const cplx* H, double* out, const cplx* tw, const cplx* stw, long long nrows, double scale,
                           double mu) {
  using C = XCfg<HN, true>;
  auto kfn = p2_xinv_kernel<HN>;
  static KernelSetup ks;
  const int per_sm = ks.get(kfn, C::THREADS, C::SMEM);
  if (per_sm == 0) return ks.err;
  const long long ngroups = (nrows + C::ROWS - 1) / C::ROWS;
  ProfScope prof_("fft_xpass_inv", st);
  GSP_LAUNCH(kfn, dim3(persistent_grid(per_sm, sms, ngroups)), dim3(C::THREADS), C::SMEM, st, H, out, tw, stw, nrows, scale, mu);
  g_launches++;
  return cudaGetLastError();
}

template <int N, int FLAGS, int B>
cudaError_t launch_p2_strided_f(cudaStream_t st, int sms, const TensorMap& tmH, int axis, cplx* H, const cplx* twp, const cplx* tw, const cplx* twi,
                                long long es, int hx,
                                long long nother, long long other_stride, const double* Fh, long long esF, long long other_strideF, double s,
                                const Sub& sub, const double* Fperm = nullptr) {
  constexpr int STAGES = GSP_STRIDED_STAGES;
  using C = StridedCfg<N, B, FLAGS, STAGES>;
  auto kfn = p2_strided_kernel<N, B, FLAGS, STAGES>;
  static KernelSetup ks;
  const int per_sm = ks.get(kfn, C::THREADS, C::SMEM);
  if (per_sm == 0) return ks.err;
  const int nball = (hx + B - 1) / B;
  int bx0 = sub.bx0, nbundles = sub.nb > 0 ? sub.nb : nball;
  if (bx0 + nbundles > nball) nbundles = nball - bx0;
  const long long o0 = sub.o0, o1 = sub.o1 > 0 ? sub.o1 : nother;
  if (nbundles <= 0 || o1 <= o0) return cudaSuccess;
  const long long ubeg = o0 * nbundles, uend = o1 * nbundles;
  ProfScope prof_((FLAGS & P2_MUL) ? "fft_strided_fwd_mul_inv" : ((FLAGS & P2_FWD) ? "fft_strided_fwd" : "fft_strided_inv"), st);
  GSP_LAUNCH(kfn, dim3(persistent_grid(per_sm, sms, uend - ubeg)), dim3(C::THREADS), C::SMEM, st, tmH, axis, H, twp, tw, twi, es, hx, nbundles, uend,
             other_stride, Fh, esF, other_strideF, s, bx0, ubeg, (FLAGS & P2_MUL) ? Fperm : (const double*)nullptr);
  g_launches++;
  return cudaGetLastError();
}

// wide bundles (16 kx = 256-byte runs) exist for the middle axis of 3-D grids with extents <= 256 (opt-in, GSP_FFT_WIDE=1): the first
// measurements on the B200 at 256^3 favoured them for the y passes (57.5 -> 53.4 us, 51.1 -> 50.4 us; the fused pass of the last axis
// lost, 77.4 -> 81.1 us, and never used them), a repeated same-box A/B of the final tree did not (y inverse 52 -> 56 us)
constexpr int P2_WIDE = 16;
GSP_HD constexpr bool p2_wide_ok(int N) { return N >= 64 && N <= 256 && p2_bundle(N) < P2_WIDE; }

template <int N, int B>
cudaError_t launch_p2_strided_b(cudaStream_t st, int sms, int flags, const TensorMap& tmH, int axis, cplx* H, const cplx* twp, const cplx* tw,
                                const cplx* twi, long long es,
                                int hx, long long nother, long long other_stride, const double* Fh, long long esF, long long other_strideF,
                                double s, const Sub& sub, const double* Fperm) {
  if (flags == P2_FWD) return launch_p2_strided_f<N, P2_FWD, B>(st, sms, tmH, axis, H, twp, tw, twi, es, hx, nother, other_stride, Fh, esF, other_strideF, s, sub);
  if (flags == P2_INV) return launch_p2_strided_f<N, P2_INV, B>(st, sms, tmH, axis, H, twp, tw, twi, es, hx, nother, other_stride, Fh, esF, other_strideF, s, sub);
  return launch_p2_strided_f<N, P2_FWD | P2_MUL | P2_INV, B>(st, sms, tmH, axis, H, twp, tw, twi, es, hx, nother, other_stride, Fh, esF, other_strideF, s, sub,
                                                             Fperm);
}

template <int N>
cudaError_t launch_p2_strided(cudaStream_t st, int sms, int flags, int bundle, const TensorMap& tmH, int axis, cplx* H, const cplx* twp, const cplx* tw,
                              const cplx* twi, long long es,
                              int hx, long long nother, long long other_stride, const double* Fh, long long esF, long long other_strideF,
                              double s, const Sub& sub, const double* Fperm) {
  if constexpr (p2_wide_ok(N)) {
    if (bundle == P2_WIDE)
      return launch_p2_strided_b<N, P2_WIDE>(st, sms, flags, tmH, axis, H, twp, tw, twi, es, hx, nother, other_stride, Fh, esF, other_strideF, s, sub,
                                             nullptr);
  }
  if (bundle != p2_bundle(N)) return cudaErrorInvalidValue;
  return launch_p2_strided_b<N, p2_bundle(N)>(st, sms, flags, tmH, axis, H, twp, tw, twi, es, hx, nother, other_stride, Fh, esF, other_strideF, s, sub,
                                              Fperm);
}

// item-major copy of F for the fused pass of the last axis of a 3-D grid (bundles of p2_bundle(N) kx)
template <int N>
cudaError_t launch_p2_permute(cudaStream_t st, const double* Fh, long long esF, long long other_strideF, int hx, long long nother, double* Fperm) {
  constexpr int B = p2_bundle(N);
  using C = StridedCfg<N, B, P2_FWD | P2_MUL | P2_INV, 1>;
  const long long items = nother * ((hx + B - 1) / B);
  GSP_LAUNCH((p2_permute_F_kernel<N, B>), dim3((unsigned)items), dim3(C::THREADS), 0, st, Fh, esF, other_strideF, hx, Fperm);
  g_launches++;
  return cudaGetLastError();
}

#ifdef GSP_EXPERIMENTAL
template <int HN, int NY, bool INV, bool RNG>
cudaError_t launch_plane(cudaStream_t st, int sms, const TensorMap& tmHy, const double* in, double* out, cplx* H, const cplx* twx,
                         const cplx* stwx, const cplx* stwy, int nz, int* sync, double scale, double mu, const XRng& rng) {
  using C = PlaneCfg<HN, NY, INV>;
  if constexpr (!C::OK) {
    return cudaErrorInvalidValue;
  } else {
    auto kfn = p2_plane_kernel<HN, NY, INV, RNG>;
    static KernelSetup ks;
    const int per_sm = ks.get(kfn, PLANE_THREADS, C::SMEM);
    if (per_sm == 0) return ks.err;
    long long grid = (long long)per_sm * sms;
    if (grid > (long long)nz * C::PB) grid = (long long)nz * C::PB;
    // every CTA holds one claimed item: lag the dependent kind beyond the window of items in flight
    const int lag = (int)((grid + C::PB - 1) / C::PB) + 2;
    ProfScope prof_(INV ? "fft_plane_yx_inv" : (RNG ? "fft_plane_xy_fwd_rng" : "fft_plane_xy_fwd"), st);
    // (claims are dynamic, so no co-residency is needed on the GPU; the emulator runs the CTAs concurrently to exercise the waits)
    GSP_LAUNCH_COOP(kfn, dim3((unsigned)grid), dim3(PLANE_THREADS), C::SMEM, st, tmHy, in, out, H, twx, stwx, stwy, nz, lag, sync, scale, mu,
                    rng);
    g_launches++;
    return cudaGetLastError();
  }
}

// (HN, NY) combinations compiled for the fused x+y kernels; everything else runs the separate passes
inline bool plane_supported(int hn, int ny) { return (hn == 64 || hn == 128 || hn == 256) && (ny == 128 || ny == 256); }

template <bool INV, bool RNG>
cudaError_t launch_plane_dispatch(int hn, int ny, cudaStream_t st, int sms, const TensorMap& tmHy, const double* in, double* out, cplx* H,
                                  const cplx* twx, const cplx* stwx, const cplx* stwy, int nz, int* sync, double scale, double mu,
                                  const XRng& rng) {
#define GSP_PL(HN_, NY_) \
  if (hn == HN_ && ny == NY_) return launch_plane<HN_, NY_, INV, RNG>(st, sms, tmHy, in, out, H, twx, stwx, stwy, nz, sync, scale, mu, rng);
  GSP_PL(64, 128) GSP_PL(64, 256) GSP_PL(128, 128) GSP_PL(128, 256) GSP_PL(256, 128) GSP_PL(256, 256)
#undef GSP_PL
  return cudaErrorInvalidValue;
}

#else
inline bool plane_supported(int, int) { return false; }
#endif

#define GSP_P2_SWITCH(n, CALL)        \
  switch (n) {                        \
    case 16: return CALL(16);         \
    case 32: return CALL(32);         \
    case 64: return CALL(64);         \
    case 128: return CALL(128);       \
    case 256: return CALL(256);       \
    case 512: return CALL(512);       \
    case 1024: return CALL(1024);     \
    case 2048: return CALL(2048);     \
    case 4096: return CALL(4096);     \
    default: return cudaErrorInvalidValue; \
  }

int setup_axes(gsp_ctx* ctx, gsp_fft_plan* p, FftDev* d) {
  const int nx = (int)p->dims[0];
  // x axis
  {
    AxisPlan& a = d->ax[0];
    a.len = nx;
    a.packed = (nx % 2 == 0) ? 1 : 0;
    const int n = a.packed ? nx / 2 : nx;
    if (!choose_radices(n, &a.lp)) return set_err(ctx, GSP_E_UNSUPPORTED, "x extent needs more than 16 FFT stages");
    std::vector<cplx> tw = make_twiddles(nx);
    GSP_CUDA_OK(ctx, d->tw[0].alloc(d->dc->dev, tw.size() * sizeof(cplx)));
    GSP_CUDA_OK(ctx, cudaMemcpyAsync(d->tw[0].p, tw.data(), tw.size() * sizeof(cplx), cudaMemcpyHostToDevice, d->dc->stream));
    GSP_CUDA_OK(ctx, cudaStreamSynchronize(d->dc->stream));
    a.lp.tw = d->tw[0].as<cplx>();
    a.lp.tw_stride = a.packed ? 2 : 1;
    const size_t L = a.packed ? (size_t)nx / 2 + 1 : (size_t)nx;
    const long long nrows = p->dims[1] * p->dims[2];
    int B = (int)(32 * 1024 / (L * sizeof(cplx)));
    if (B < 1) B = 1;
    if (B > 15) B = 15;
    if (B > nrows) B = (int)nrows;
    if (B % 2 == 0) B -= 1;  // odd: conflict-free global<->shared transposition
    if (B < 1) B = 1;
    a.B = B;
    a.smem = 2 * L * B * sizeof(cplx);
    a.fast = a.packed && p2_supported(nx / 2) && !g_force_generic;
    if (a.fast) {
      for (int inv = 0; inv < 2; ++inv) {
        std::vector<cplx> stw = make_stage_twiddles(nx / 2, inv != 0);
        DevBuf& b = inv ? d->stw_inv : d->stw_fwd;
        GSP_CUDA_OK(ctx, b.alloc(d->dc->dev, stw.size() * sizeof(cplx)));
        GSP_CUDA_OK(ctx, cudaMemcpyAsync(b.p, stw.data(), stw.size() * sizeof(cplx), cudaMemcpyHostToDevice, d->dc->stream));
        GSP_CUDA_OK(ctx, cudaStreamSynchronize(d->dc->stream));
      }
    }
    if (!a.fast && a.smem > kMaxSmem) return set_err(ctx, GSP_E_UNSUPPORTED, "x extent too large for one shared-memory line");
  }
  for (int axis = 1; axis < p->ndim; ++axis) {
    AxisPlan& a = d->ax[axis];
    const int n = (int)p->dims[axis];
    a.len = n;
    if (!choose_radices(n, &a.lp)) return set_err(ctx, GSP_E_UNSUPPORTED, "grid extent needs more than 16 FFT stages");
    std::vector<cplx> tw = make_twiddles(n);
    GSP_CUDA_OK(ctx, d->tw[axis].alloc(d->dc->dev, tw.size() * sizeof(cplx)));
    GSP_CUDA_OK(ctx, cudaMemcpyAsync(d->tw[axis].p, tw.data(), tw.size() * sizeof(cplx), cudaMemcpyHostToDevice, d->dc->stream));
    GSP_CUDA_OK(ctx, cudaStreamSynchronize(d->dc->stream));
    a.lp.tw = d->tw[axis].as<cplx>();
    a.lp.tw_stride = 1;
    int B = 8;
    while (B > 1 && 2 * (size_t)n * B * sizeof(cplx) > 96 * 1024) B /= 2;
    if (B > p->hx) B = p->hx;
    a.B = B;
    a.smem = 2 * (size_t)n * B * sizeof(cplx);
    a.fast = p2_supported(n) && !g_force_generic;
    if (a.fast) {
      for (int inv = 0; inv < 2; ++inv) {
        std::vector<cplx> stw = make_stage_twiddles(n, inv != 0);
        DevBuf& b = inv ? d->stw_ax_inv[axis] : d->stw_ax_fwd[axis];
        GSP_CUDA_OK(ctx, b.alloc(d->dc->dev, stw.size() * sizeof(cplx)));
        GSP_CUDA_OK(ctx, cudaMemcpyAsync(b.p, stw.data(), stw.size() * sizeof(cplx), cudaMemcpyHostToDevice, d->dc->stream));
        GSP_CUDA_OK(ctx, cudaStreamSynchronize(d->dc->stream));
      }
    }
    if (!a.fast && a.smem > kMaxSmem) return set_err(ctx, GSP_E_UNSUPPORTED, "grid extent too large for one shared-memory line");
  }
  return GSP_OK;
}

// x passes over the rows [row0, row0 + nrows) of the lane's work spectrum (`in` / `out` point at the first of those rows).
// 1-D / 2-D grids: a batch of realizations, stored back to back, is simply more rows of ONE launch.
// in == nullptr (fast x axis only): noise of realization rng.first_real (+ row / rows_per_real) generated inside the kernel
cudaError_t run_xfwd(FftDev* d, gsp_fft_plan* p, const Lane& L, const double* in, long long row0, long long nrows, XRng rng = XRng{}) {
  const AxisPlan& a = d->ax[0];
  cplx* H = L.H + row0 * p->hx;
  rng.row_base = row0;
  rng.rows_per_real = p->dims[1] * p->dims[2];
  if (!in && !a.fast) return cudaErrorInvalidValue;
  if (a.fast) {
#define GSP_CALL(HN) launch_p2_xfwd<HN>(L.st, d->dc->sms, in, H, a.lp.tw, d->stw_fwd.as<cplx>(), nrows, rng)
    GSP_P2_SWITCH((int)p->dims[0] / 2, GSP_CALL)
#undef GSP_CALL
  }
  auto kfn = xpass_fwd_kernel;
  cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a.smem);
  if (e != cudaSuccess) return e;
  ProfScope prof_("fft_xpass_fwd", L.st);
  GSP_LAUNCH(kfn, dim3((unsigned)((nrows + a.B - 1) / a.B)), dim3(256), a.smem, L.st, a.lp, (int)p->dims[0], p->hx, nrows,
             a.B, a.packed, in, H);
  g_launches++;
  return cudaGetLastError();
}

cudaError_t run_xinv(FftDev* d, gsp_fft_plan* p, const Lane& L, double* out, long long row0, long long nrows, double scale, double mu) {
  const AxisPlan& a = d->ax[0];
  const cplx* H = L.H + row0 * p->hx;
  if (a.fast) {
#define GSP_CALL(HN) launch_p2_xinv<HN>(L.st, d->dc->sms, H, out, a.lp.tw, d->stw_inv.as<cplx>(), nrows, scale, mu)
    GSP_P2_SWITCH((int)p->dims[0] / 2, GSP_CALL)
#undef GSP_CALL
  }
  auto kfn = xpass_inv_kernel;
  cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a.smem);
  if (e != cudaSuccess) return e;
  ProfScope prof_("fft_xpass_inv", L.st);
  GSP_LAUNCH(kfn, dim3((unsigned)((nrows + a.B - 1) / a.B)), dim3(256), a.smem, L.st, a.lp, (int)p->dims[0], p->hx, nrows,
             a.B, a.packed, H, out, scale, mu);
  g_launches++;
  return cudaGetLastError();
}

// `sub` (fast kernels only) restricts the pass to a slab; the generic kernels always run the whole grid
cudaError_t run_strided(FftDev* d, gsp_fft_plan* p, const Lane& L, int axis, int flags, const double* Fh, double s, long long batch = 1,
                        const Sub& sub = Sub()) {
  const AxisPlan& a = d->ax[axis];
  cplx* H = L.H;
  const long long hx = p->hx;
  long long es, other_stride, nother, esF, other_strideF;
  const long long hxF = p->hxF;
  if (axis == 1) {
    es = hx;
    other_stride = hx * p->dims[1];
    nother = p->dims[2];
    esF = hxF;
    other_strideF = hxF * p->dims[1];
    if (p->ndim == 2) {  // the "other" index runs over the realizations of a batch; every realization sees the same F
      nother = batch;
      other_strideF = 0;
    }
  } else {
    es = hx * p->dims[1];
    other_stride = hx;
    nother = p->dims[1];
    esF = hxF * p->dims[1];
    other_strideF = hxF;
  }
  if (a.fast) {
#define GSP_CALL(NN) \
  launch_p2_strided<NN>(L.st, d->dc->sms, flags, a.bundle, L.tmH[axis], axis, H, a.lp.tw, d->stw_ax_fwd[axis].as<cplx>(), d->stw_ax_inv[axis].as<cplx>(), es, \
                        (int)hx, \
                        nother, other_stride, Fh, esF, other_strideF, s, sub, \
                        (axis == 2 && sub.nb == 0 && a.bundle == p2_bundle(a.len)) ? d->Fperm.as<double>() : (const double*)nullptr)
    GSP_P2_SWITCH(a.len, GSP_CALL)
#undef GSP_CALL
  }
  auto kfn = strided_pass_kernel;
  cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a.smem);
  if (e != cudaSuccess) return e;
  dim3 grid((unsigned)((hx + a.B - 1) / a.B), (unsigned)nother);
  ProfScope prof_((flags & PASS_MUL) ? "fft_strided_fwd_mul_inv" : ((flags & PASS_FWD) ? "fft_strided_fwd" : "fft_strided_inv"),
                  L.st);
  GSP_LAUNCH(kfn, grid, dim3(256), a.smem, L.st, a.lp, H, es, (int)hx, a.B, other_stride, flags, Fh, esF, other_strideF, s);
  g_launches++;
  return cudaGetLastError();
}

// fused x+y forward of one realization on a lane; in == nullptr: noise from the counter RNG.  Zeroes the lane's claim / plane
// counters first (`both`: also those of the inverse kernel that follows in `realization`).
cudaError_t run_plane_fwd(FftDev* d, gsp_fft_plan* p, const Lane& L, const double* in, const XRng& rng, bool both) {
  const int nz = (int)p->dims[2];
  cudaError_t e = cudaMemsetAsync(L.syncbuf.p, 0, (size_t)(both ? 2 : 1) * (nz + 1) * sizeof(int), L.st);
  if (e != cudaSuccess) return e;
  const int hn = (int)p->dims[0] / 2, ny = (int)p->dims[1];
  const cplx *twx = d->ax[0].lp.tw, *stwx = d->stw_fwd.as<cplx>(), *stwy = d->stw_ax_fwd[1].as<cplx>();
#ifdef GSP_EXPERIMENTAL
  // (the RNG = true instantiation - noise drawn inside the x items - hung on the B200 in round 1 and is not instantiated: the fused
  // mode takes its noise from rng_fill_kernel's scratch array)
  if (!in) return cudaErrorInvalidValue;
  return launch_plane_dispatch<false, false>(hn, ny, L.st, d->dc->sms, L.tmH[1], in, nullptr, L.H, twx, stwx, stwy, nz, L.syncbuf.as<int>(), 0.0, 0.0, rng);
#else
  (void)hn; (void)ny; (void)twx; (void)stwx; (void)stwy; (void)in; (void)rng;
  return cudaErrorInvalidValue;
#endif
}
cudaError_t run_plane_inv(FftDev* d, gsp_fft_plan* p, const Lane& L, double* out, double scale, double mu) {
#ifdef GSP_EXPERIMENTAL
  const int nz = (int)p->dims[2];
  return launch_plane_dispatch<true, false>((int)p->dims[0] / 2, (int)p->dims[1], L.st, d->dc->sms, L.tmH[1], nullptr, out, L.H, d->ax[0].lp.tw,
                                            d->stw_inv.as<cplx>(), d->stw_ax_inv[1].as<cplx>(), nz, L.syncbuf.as<int>() + nz + 1, scale, mu, XRng{});
#else
  (void)d; (void)p; (void)L; (void)out; (void)scale; (void)mu;
  return cudaErrorInvalidValue;
#endif
}

// forward transform of a real field into d->H (all axes)
cudaError_t forward_all(FftDev* d, gsp_fft_plan* p, const double* in) {
  const Lane& L = *d->lanes[0];
  if (d->fused_xy) {
    cudaError_t e = run_plane_fwd(d, p, L, in, XRng{}, false);
    if (e == cudaSuccess) e = run_strided(d, p, L, 2, PASS_FWD, nullptr, 0.0);
    return e;
  }
  cudaError_t e = run_xfwd(d, p, L, in, 0, p->dims[1] * p->dims[2]);
  for (int axis = 1; axis < p->ndim && e == cudaSuccess; ++axis) e = run_strided(d, p, L, axis, PASS_FWD, nullptr, 0.0);
  return e;
}

// one realization on one lane: real noise (device) -> real field (device), full grid
// w == nullptr: the forward x pass draws realization rng.first_real from the counter RNG itself (fast x axis, no fused planes)
cudaError_t realization(FftDev* d, gsp_fft_plan* p, const Lane& L, const double* w, double* out, double s, double scale_out, double mu,
                        XRng rng = XRng{}) {
  const double* Fh = d->Fh.as<double>();
  const long long nrows = p->dims[1] * p->dims[2];
  if (d->fused_xy) {
    // 3 kernels per realization: (x+y forward) -> (z forward, spectral multiply, z inverse) -> (y+x inverse)
    XRng rr = rng;
    rr.row_base = 0;
    rr.rows_per_real = nrows;
    cudaError_t e = run_plane_fwd(d, p, L, w, rr, true);
    if (e == cudaSuccess) e = run_strided(d, p, L, 2, PASS_FWD | PASS_MUL | PASS_INV, Fh, s);
    if (e == cudaSuccess) e = run_plane_inv(d, p, L, out, scale_out, mu);
    return e;
  }
  cudaError_t e = cudaSuccess;
  if (p->ndim == 3 && d->slab_mode == 1) {
    // z-plane slabs: the x pass of a slab leaves its half-spectrum rows in L2, the y pass of the same slab picks them up
    // there and overwrites them in place -> the x/y intermediate never makes a round trip through HBM (52 N instead of 84 N bytes)
    const long long nz = p->dims[2], ny = p->dims[1], zs = d->slab_planes;
    for (long long z0 = 0; z0 < nz && e == cudaSuccess; z0 += zs) {
      const long long z1 = z0 + zs < nz ? z0 + zs : nz;
      Sub sub;
      sub.o0 = z0;
      sub.o1 = z1;
      e = run_xfwd(d, p, L, w ? w + z0 * ny * p->dims[0] : nullptr, z0 * ny, (z1 - z0) * ny, rng);
      if (e == cudaSuccess) e = run_strided(d, p, L, 1, PASS_FWD, nullptr, 0.0, 1, sub);
    }
    if (e == cudaSuccess) e = run_strided(d, p, L, 2, PASS_FWD | PASS_MUL | PASS_INV, Fh, s);
    for (long long z0 = 0; z0 < nz && e == cudaSuccess; z0 += zs) {
      const long long z1 = z0 + zs < nz ? z0 + zs : nz;
      Sub sub;
      sub.o0 = z0;
      sub.o1 = z1;
      e = run_strided(d, p, L, 1, PASS_INV, nullptr, 0.0, 1, sub);
      if (e == cudaSuccess) e = run_xinv(d, p, L, out + z0 * ny * p->dims[0], z0 * ny, (z1 - z0) * ny, scale_out, mu);
    }
    return e;
  }
  e = run_xfwd(d, p, L, w, 0, nrows, rng);
  if (e != cudaSuccess) return e;
  const int last = p->ndim - 1;
  if (last == 0) {
    long long blocks = (p->nh + 255) / 256;
    if (blocks > (long long)d->dc->sms * 8) blocks = (long long)d->dc->sms * 8;
    GSP_LAUNCH(spectral_mul_kernel, dim3((unsigned)blocks), dim3(256), 0, L.st, L.H, Fh, p->nh, p->nh, s);
    g_launches++;
    e = cudaGetLastError();
  } else if (p->ndim == 3 && d->slab_mode == 2) {
    // kx-bundle groups: y forward, z (forward, multiply, inverse) and y inverse of one group of bundles back to back;
    // the group's slab (group x ny x nz) stays in L2 between the three kernels
    const int B = d->ax[1].bundle;
    const int nball = (p->hx + B - 1) / B;
    for (int bx0 = 0; bx0 < nball && e == cudaSuccess; bx0 += d->slab_bundles) {
      Sub sub;
      sub.bx0 = bx0;
      sub.nb = d->slab_bundles;
      e = run_strided(d, p, L, 1, PASS_FWD, nullptr, 0.0, 1, sub);
      if (e == cudaSuccess) e = run_strided(d, p, L, 2, PASS_FWD | PASS_MUL | PASS_INV, Fh, s, 1, sub);
      if (e == cudaSuccess) e = run_strided(d, p, L, 1, PASS_INV, nullptr, 0.0, 1, sub);
    }
  } else {
    for (int axis = 1; axis < last && e == cudaSuccess; ++axis) e = run_strided(d, p, L, axis, PASS_FWD, nullptr, 0.0);
    if (e == cudaSuccess) e = run_strided(d, p, L, last, PASS_FWD | PASS_MUL | PASS_INV, Fh, s);
    for (int axis = last - 1; axis >= 1 && e == cudaSuccess; --axis) e = run_strided(d, p, L, axis, PASS_INV, nullptr, 0.0);
  }
  if (e != cudaSuccess) return e;
  return run_xinv(d, p, L, out, 0, nrows, scale_out, mu);
}

// `nb` realizations at once (1-D / 2-D grids: one launch per pass for the whole batch; 3-D: one realization per lane, the lanes
// run concurrently and join the compute stream at the end)
cudaError_t realization_batch(FftDev* d, gsp_fft_plan* p, const double* w, double* out, long long nb, double s, double scale_out, double mu,
                              XRng rng = XRng{}) {
  const Lane& L0 = *d->lanes[0];
  if (p->ndim == 3 || nb == 1) {
    const long long nl = (long long)d->lanes.size() < nb ? (long long)d->lanes.size() : nb;
    cudaError_t e = cudaSuccess;
    if (nl > 1) {
      e = cudaEventRecord(d->ev_fork, L0.st);
      for (long long l = 1; l < nl && e == cudaSuccess; ++l) e = cudaStreamWaitEvent(d->lanes[l]->st, d->ev_fork, 0);
    }
    for (long long r = 0; r < nb && e == cudaSuccess; ++r) {
      XRng rr = rng;
      rr.first_real = rng.first_real + r;
      e = realization(d, p, *d->lanes[r % nl], w ? w + r * p->N : nullptr, out + r * p->N, s, scale_out, mu, rr);
    }
    for (long long l = 1; l < nl; ++l) {  // always join, also after an error: nothing may outlive the call on a side stream
      cudaError_t e2 = cudaEventRecord(d->lanes[l]->done, d->lanes[l]->st);
      if (e2 == cudaSuccess) e2 = cudaStreamWaitEvent(L0.st, d->lanes[l]->done, 0);
      if (e == cudaSuccess) e = e2;
    }
    return e;
  }
  const double* Fh = d->Fh.as<double>();
  const long long nrows = p->dims[1] * p->dims[2] * nb;
  cudaError_t e = run_xfwd(d, p, L0, w, 0, nrows, rng);
  if (e != cudaSuccess) return e;
  if (p->ndim == 1) {
    const long long total = p->nh * nb;
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)d->dc->sms * 8) blocks = (long long)d->dc->sms * 8;
    GSP_LAUNCH(spectral_mul_kernel, dim3((unsigned)blocks), dim3(256), 0, L0.st, L0.H, Fh, p->nh, total, s);
    g_launches++;
    e = cudaGetLastError();
  } else {
    e = run_strided(d, p, L0, 1, PASS_FWD | PASS_MUL | PASS_INV, Fh, s, nb);
  }
  if (e != cudaSuccess) return e;
  return run_xinv(d, p, L0, out, 0, nrows, scale_out, mu);
}

int build_device(gsp_ctx* ctx, gsp_fft_plan* p, FftDev* d, const CovDev& cov, const DomDev& dom, long long eref) {
  cudaSetDevice(d->dc->dev);
  GSP_TRY(setup_axes(ctx, p, d));
  GSP_CUDA_OK(ctx, d->Fh.alloc(d->dc->dev, (size_t)p->nhF * sizeof(double)));
  GSP_CUDA_OK(ctx, cudaMemsetAsync(d->Fh.p, 0, (size_t)p->nhF * sizeof(double), d->dc->stream));
  GSP_CUDA_OK(ctx, d->H.alloc(d->dc->dev, (size_t)p->nh * (p->ndim == 3 ? 1 : p->rb) * sizeof(cplx)));
  GSP_CUDA_OK(ctx, cudaEventCreate(&d->ev0));
  GSP_CUDA_OK(ctx, cudaEventCreate(&d->ev1));
  GSP_CUDA_OK(ctx, cudaEventCreateWithFlags(&d->ev_fork, cudaEventDisableTiming));
  // Lanes and slabs (3-D grids on the register-resident kernels only).  Defaults measured on B200 at 256^3; GSP_FFT_LANES,
  // GSP_FFT_SLAB (0 | 1 | 2), GSP_FFT_SLAB_PLANES and GSP_FFT_SLAB_BUNDLES override them for A/B runs.
  int nlanes = 1;
  const bool all_fast = p->ndim == 3 && d->ax[0].fast && d->ax[1].fast && d->ax[2].fast;
  if (all_fast) {
    auto env_int = [](const char* name, int dflt) {
      const char* v = exp_env(name);
      return v && v[0] ? atoi(v) : dflt;
    };
    nlanes = (int)p->rb;  // 3-D: realizations per chunk = concurrent lanes (gsp_fft_plan_create)
    d->slab_mode = env_int("GSP_FFT_SLAB", 0);
    if (d->slab_mode < 0 || d->slab_mode > 2) d->slab_mode = 0;
    d->slab_planes = env_int("GSP_FFT_SLAB_PLANES", 32);
    if (d->slab_planes < 1) d->slab_planes = 1;
    d->slab_bundles = env_int("GSP_FFT_SLAB_BUNDLES", 4);
    if (d->slab_bundles < 1) d->slab_bundles = 1;
  }
  for (int l = 0; l < nlanes; ++l) {
    std::unique_ptr<Lane> L(new Lane);
    if (l == 0) {
      L->st = d->dc->stream;
      L->H = d->H.as<cplx>();
    } else {
      GSP_CUDA_OK(ctx, cudaStreamCreateWithFlags(&L->st, cudaStreamNonBlocking));
      L->own_stream = true;
      GSP_CUDA_OK(ctx, L->Hbuf.alloc(d->dc->dev, (size_t)p->nh * sizeof(cplx)));
      L->H = L->Hbuf.as<cplx>();
      GSP_CUDA_OK(ctx, cudaEventCreateWithFlags(&L->done, cudaEventDisableTiming));
    }
    d->lanes.push_back(std::move(L));
  }
  for (int axis = 1; axis < p->ndim; ++axis) {
    AxisPlan& a = d->ax[axis];
    if (!a.fast) continue;
    // GSP_FFT_WIDE=1 turns the 16-kx bundles on (off by default: a repeated same-box A/B of the final tree showed 280 vs 274 us per
    // realization on one lane and no difference with 4 lanes); the fused plane kernels and the bundle-group slabs are built for 8
    const char* wenv = exp_env("GSP_FFT_WIDE");
    const char* fenv = exp_env("GSP_FFT_FUSE");
    const bool fuse_req = fenv && fenv[0] ? fenv[0] == '1' : GSP_FFT_FUSE_DEFAULT != 0;
    const bool wide = p->ndim == 3 && axis == 1 && p2_wide_ok(a.len) && (wenv && wenv[0] == '1') && !fuse_req && d->slab_mode != 2;
    a.bundle = wide ? P2_WIDE : p2_bundle(a.len);
    const unsigned B = (unsigned)a.bundle;
    const unsigned long long dH[3] = {2ull * p->hx, (unsigned long long)p->dims[1],
                                      (unsigned long long)(p->ndim == 2 ? p->rb : p->dims[2])};  // 2-D: 3rd extent = batch
    const unsigned long long dF[3] = {(unsigned long long)p->hxF, (unsigned long long)p->dims[1], (unsigned long long)p->dims[2]};
    const unsigned lbox = a.len < 256 ? (unsigned)a.len : 256u;
    const unsigned boxH[3] = {2 * B, axis == 1 ? lbox : 1u, axis == 2 ? lbox : 1u};
    const unsigned boxF[3] = {B, boxH[1], boxH[2]};
    int r1 = 0;
    for (auto& L : d->lanes) {
      const int r = make_tensor_map_f64(&L->tmH[axis], L->H, dH, 16ull * p->hx, 16ull * p->hx * p->dims[1], boxH);
      if (r != 0) r1 = r;
    }
    a.tmH = d->lanes[0]->tmH[axis];
    int r2 = make_tensor_map_f64(&a.tmF, d->Fh.p, dF, 8ull * p->hxF, 8ull * p->hxF * p->dims[1], boxF);
    if (r1 != 0 || r2 != 0) return set_err(ctx, GSP_E_CUDA, "cuTensorMapEncodeTiled failed for an FFT pass (code " + std::to_string(r1 ? r1 : r2) + ")");
  }
  {
    // Fused x+y kernels (fft_plane.cuh): GSP_FFT_FUSE=1 turns them on, =0 off (A/B runs).
    const char* env = exp_env("GSP_FFT_FUSE");
    const bool allow = env && env[0] ? env[0] == '1' : GSP_FFT_FUSE_DEFAULT != 0;
    d->fused_xy = allow && all_fast && plane_supported((int)p->dims[0] / 2, (int)p->dims[1]) && d->slab_mode == 0;
    if (d->fused_xy) {
      for (auto& L : d->lanes) GSP_CUDA_OK(ctx, L->syncbuf.alloc(d->dc->dev, (size_t)(2 * p->dims[2] + 2) * sizeof(int)));
    }
  }
  DevBuf C, partial, total;
  GSP_CUDA_OK(ctx, C.alloc(d->dc->dev, (size_t)p->N * sizeof(double)));
  const int nblocks = d->dc->sms * 4;
  GSP_CUDA_OK(ctx, partial.alloc(d->dc->dev, (size_t)nblocks * sizeof(double)));
  GSP_CUDA_OK(ctx, total.alloc(d->dc->dev, sizeof(double)));
  launch_cov_to_center(d->dc->stream, d->dc->sms, cov, dom, eref, C.as<double>());
  GSP_CUDA_OK(ctx, cudaGetLastError());
  GSP_CUDA_OK(ctx, forward_all(d, p, C.as<double>()));
  GSP_LAUNCH(spectrum_finalize_kernel, dim3((unsigned)nblocks), dim3(256), 0, d->dc->stream, d->H.as<cplx>(), d->Fh.as<double>(),
             p->nh, p->hx, p->hxF, (int)p->dims[0], partial.as<double>());
  g_launches++;
  GSP_LAUNCH(sum_partials_kernel, dim3(1), dim3(256), 0, d->dc->stream, partial.as<double>(), nblocks, total.as<double>());
  g_launches++;
  GSP_CUDA_OK(ctx, cudaGetLastError());
  double s2 = 0.0;
  GSP_CUDA_OK(ctx, cudaMemcpyAsync(&s2, total.p, sizeof(double), cudaMemcpyDeviceToHost, d->dc->stream));
  GSP_CUDA_OK(ctx, cudaStreamSynchronize(d->dc->stream));
  p->sumF2 = s2;
  {
    // item-major copy of F for the last axis' fused pass (3-D, register-resident kernels); GSP_FFT_FPERM=0 keeps the strided reads
    const char* env = getenv("GSP_FFT_FPERM");
    const AxisPlan& a = d->ax[2];
    if (all_fast && !(env && env[0] == '0') && a.bundle == p2_bundle(a.len)) {
      const long long nball = (p->hx + a.bundle - 1) / a.bundle;
      const size_t count = (size_t)p->dims[1] * (size_t)nball * (size_t)a.len * (size_t)a.bundle;
      GSP_CUDA_OK(ctx, d->Fperm.alloc(d->dc->dev, count * sizeof(double)));
      auto build = [&]() -> cudaError_t {
#define GSP_CALL(NN) \
  launch_p2_permute<NN>(d->dc->stream, d->Fh.as<double>(), (long long)p->hxF * p->dims[1], (long long)p->hxF, p->hx, p->dims[1], d->Fperm.as<double>())
        GSP_P2_SWITCH(a.len, GSP_CALL)
#undef GSP_CALL
      };
      GSP_CUDA_OK(ctx, build());
      GSP_CUDA_OK(ctx, cudaStreamSynchronize(d->dc->stream));
    }
  }
  return GSP_OK;
}

// scalar folded into the spectrum multiply: sqrt(sill / sigma^2) / N  (fftsim.jl:128-132)
double fold_scale(const gsp_fft_plan* p, double sill) {
  const double N = (double)p->N;
  const double sigma2 = p->sumF2 / (N * (N - 1.0));
  return std::sqrt(sill / sigma2) / N;
}

}  // namespace
}  // namespace gsp

extern "C" int gsp_fft_plan_create(gsp_ctx* ctx, const gsp_cov_model* cov, const gsp_domain* grid, gsp_fft_plan** out) {
  if (!ctx) return -1;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!out) return set_err(ctx, -4, "out is NULL");
  *out = nullptr;
  if (!grid || grid->kind != 1) return set_err(ctx, -3, "FFTSIM needs a CartesianGrid domain (kind 1)");
  {
    const char* env = getenv("GSP_FFT_GENERIC");
    g_force_generic = env && env[0] == '1';
  }
  DomDev dom;
  GSP_TRY(make_dom_dev(ctx, grid, 3, &dom));
  CovDev cd;
  GSP_TRY(make_cov_dev(ctx, cov, grid->dim, 2, &cd));
  std::unique_ptr<gsp_fft_plan> p(new gsp_fft_plan);
  p->ctx = ctx;
  p->cov = cd;
  p->dom = dom;
  p->ndim = grid->dim;
  p->N = 1;
  for (int a = 0; a < 3; ++a) {
    p->dims[a] = a < grid->dim ? grid->dims[a] : 1;
    p->N *= p->dims[a];
  }
  if (p->N < 2) return set_err(ctx, -3, "grid must have at least 2 elements");
  long long eref = 0, stride = 1;
  for (int a = 0; a < p->ndim; ++a) {
    const long long c = p->dims[a] / 2;  // CartesianIndex(dims .÷ 2), 1-based (fftsim.jl:80)
    if (c < 1) return set_err(ctx, -3, "grid extent < 2: dims .÷ 2 is not a valid index in the reference");
    eref += (c - 1) * stride;
    stride *= p->dims[a];
  }
  p->hx = (int)(p->dims[0] / 2 + 1);
  p->hxF = (p->hx + 1) & ~1;
  p->nh = (long long)p->hx * p->dims[1] * p->dims[2];
  p->nhF = (long long)p->hxF * p->dims[1] * p->dims[2];
  if (p->ndim < 3) {
    static long long batch_mb = -1;  // GSP_FFT_BATCH_MB: cap of a batch's work spectrum
    if (batch_mb < 0) {
      const char* env = getenv("GSP_FFT_BATCH_MB");
      batch_mb = env ? atoll(env) : 1024;  // measured on B200 (1024^2 x 64): 24 MB 1.65 ms, 256 MB 1.09 ms, 1 GB 1.02 ms
      if (batch_mb < 1) batch_mb = 1;
    }
    p->rb = (batch_mb << 20) / (p->nh * (long long)sizeof(cplx));
    if (p->rb > 64) p->rb = 64;
    if (p->rb < 1) p->rb = 1;
  }
  if (p->ndim == 3) {
    // 3-D: a chunk = one realization per lane (concurrent streams with their own work spectrum); only the register-resident
    // power-of-two kernels are worth it, everything else keeps one lane
    bool fast = !g_force_generic && p->dims[0] % 2 == 0 && p2_supported((int)p->dims[0] / 2);
    for (int a = 1; a < 3; ++a) fast = fast && p2_supported((int)p->dims[a]);
    const char* env = getenv("GSP_FFT_LANES");
    long long lanes = env && env[0] ? atoll(env) : 4;  // measured on B200 at 256^3: 1 lane 281 us, 2 lanes 258 us, 4 lanes 252 us per realization
    if (lanes < 1) lanes = 1;
    if (lanes > 8) lanes = 8;
    p->rb = fast ? lanes : 1;
  }
  for (auto& dc : ctx->devs) {
    std::unique_ptr<FftDev> d(new FftDev);
    d->dc = &dc;
    GSP_TRY(build_device(ctx, p.get(), d.get(), cd, dom, eref));
    p->dev.push_back(std::move(d));
  }
  *out = p.release();
  return GSP_OK;
}

extern "C" int gsp_fft_plan_destroy(gsp_fft_plan* p) {
  if (!p) return GSP_OK;
  for (auto& d : p->dev) {
    cudaSetDevice(d->dc->dev);
    cudaStreamSynchronize(d->dc->stream);
    if (d->ev0) cudaEventDestroy(d->ev0);
    if (d->ev1) cudaEventDestroy(d->ev1);
    if (d->ev_fork) cudaEventDestroy(d->ev_fork);
    for (auto& L : d->lanes) {
      if (L->own_stream) {
        cudaStreamSynchronize(L->st);
        cudaStreamDestroy(L->st);
      }
      if (L->done) cudaEventDestroy(L->done);
    }
  }
  delete p;
  return GSP_OK;
}

extern "C" int gsp_fft_plan_get(gsp_fft_plan* p, double* F) {
  if (!p) return -1;
  gsp_ctx* ctx = p->ctx;
  std::lock_guard<std::mutex> lk(p->mu);
  if (!F) return set_err(ctx, -2, "F is NULL");
  FftDev* d = p->dev[0].get();
  cudaSetDevice(d->dc->dev);
  std::vector<double> Fh((size_t)p->nhF);
  GSP_CUDA_OK(ctx, cudaMemcpyAsync(Fh.data(), d->Fh.p, Fh.size() * sizeof(double), cudaMemcpyDeviceToHost, d->dc->stream));
  GSP_CUDA_OK(ctx, cudaStreamSynchronize(d->dc->stream));
  const long long nx = p->dims[0], ny = p->dims[1], nz = p->dims[2], hx = p->hx, hxF = p->hxF;
  for (long long z = 0; z < nz; ++z)
    for (long long y = 0; y < ny; ++y)
      for (long long x = 0; x < nx; ++x) {
        double v;
        if (x < hx)
          v = Fh[(size_t)(x + hxF * (y + ny * z))];
        else
          v = Fh[(size_t)((nx - x) + hxF * (((ny - y) % ny) + ny * ((nz - z) % nz)))];
        F[x + nx * (y + ny * z)] = v;
      }
  return GSP_OK;
}

namespace gsp {
namespace {

// z = zbar + (zu - zbaru) for the nb <= 32 unconditional realizations at Z (n x nb), in place (fftsim.jl:140-153)
int apply_conditioning(gsp_fft_plan* p, FftDev* d, double* Z, long long nb, double mu) {
  gsp_ctx* ctx = p->ctx;
  CondDev& c = d->cond;
  const long long n = p->cond_n, nk = p->cond_nk;
  cudaStream_t st = d->dc->stream;
  if (c.res_cap < nk * KRIGE_RB) {
    GSP_CUDA_OK(ctx, c.res.alloc(d->dc->dev, (size_t)(nk * KRIGE_RB) * sizeof(double)));
    c.res_cap = nk * KRIGE_RB;
  }
  long long blocks = (nk * KRIGE_RB + 255) / 256;
  if (blocks > (long long)d->dc->sms * 8) blocks = (long long)d->dc->sms * 8;
  {
    ProfScope prof_("krige_residual", st);
    GSP_LAUNCH(krige_residual_kernel, dim3((unsigned)blocks), dim3(256), 0, st, (const double*)Z, n, (const long long*)c.knodes.as<long long>(), nk,
               nb, mu, c.res.as<double>());
    g_launches++;
  }
  blocks = (n + 31) / 32;
  if (blocks > (long long)d->dc->sms * 8) blocks = (long long)d->dc->sms * 8;
  ProfScope prof_("krige_apply", st);
  GSP_LAUNCH(krige_apply_kernel, dim3((unsigned)blocks), dim3(256), 0, st, Z, n, (int)nb, (const double*)c.zbar.as<double>(),
             (const double*)c.lam.as<double>(), (const int*)c.nbr.as<int>(), p->cond_kk, (const double*)c.res.as<double>(), mu);
  g_launches++;
  GSP_CUDA_OK(ctx, cudaGetLastError());
  return GSP_OK;
}

unsigned long long hash_inds(const int64_t* inds, long long n) {
  unsigned long long h = 1469598103934665603ull;
  for (long long q = 0; q < n; ++q) {
    h ^= (unsigned long long)inds[q];
    h *= 1099511628211ull;
  }
  return h;
}

// conditional plans were built for ONE simulation domain (the view) and ONE mean.  `inds_host`: the caller's index list when it is
// a host array (the _dev entry point only checks the length): another view of the same length must not silently get the zbar and
// the weight tables of the domain the plan was conditioned on.
int check_cond_args(gsp_fft_plan* p, double mu, long long n_inds, const int64_t* inds_host = nullptr) {
  if (!p->cond) return GSP_OK;
  if (n_inds != p->cond_ninds) return set_err(p->ctx, -8, "conditional plan: n_inds / inds differ from those given to gsp_fft_plan_condition");
  if (inds_host && n_inds > 0 && hash_inds(inds_host, n_inds) != p->cond_inds_hash)
    return set_err(p->ctx, -9, "conditional plan: inds differ from the view given to gsp_fft_plan_condition");
  if (mu != p->cond_mu) return set_err(p->ctx, -7, "conditional plan: mu differs from the mean given to gsp_fft_plan_condition");
  return GSP_OK;
}

// GSP_FFT_FUSED_RNG=0: draw the noise into a scratch array first (rng_fill_kernel) instead of inside the forward x pass (A/B runs)
bool fused_rng_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* env = getenv("GSP_FFT_FUSED_RNG");
    on = (env && env[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

// sample R realizations on one device; w/out/inds are DEVICE pointers (w may be NULL => RNG into scratch)
int sample_on_device(gsp_fft_plan* p, FftDev* d, long long R, const double* w, unsigned long long seed, long long first_real,
                     double sill, double mu, long long n_inds, const long long* inds_dev, double* out, DevBuf* scratch_w,
                     DevBuf* scratch_z) {
  gsp_ctx* ctx = p->ctx;
  const double s = fold_scale(p, sill);
  long long rb = p->rb;  // realizations per batch (1-D / 2-D grids: one launch per pass for the batch; 3-D: concurrent lanes)
  // 3-D grids writing whole fields from a noise array (or the fused RNG) need no per-batch scratch: ALL R realizations go to the
  // lanes in one round-robin, each lane streams through its realizations and the lanes join the compute stream once at the end,
  // instead of a fork / join every `rb` realizations (GSP_FFT_STREAM=0: the batched fork / join, for A/B runs)
  static const bool stream_all = !(getenv("GSP_FFT_STREAM") && getenv("GSP_FFT_STREAM")[0] == '0');
  if (stream_all && p->ndim == 3 && n_inds == 0 && (w || (d->ax[0].fast && !d->fused_xy && fused_rng_enabled()))) rb = std::max<long long>(R, 1);
  for (long long r = 0; r < R; r += rb) {
    const long long nb = (R - r < rb) ? R - r : rb;
    const double* wr;
    XRng rng{};
    if (w) {
      wr = w + r * p->N;
    } else if (d->ax[0].fast && !d->fused_xy && fused_rng_enabled()) {
      // no noise array at all: the forward x pass generates realization first_real + r (+ batch offset) in its registers
      wr = nullptr;
      rng.seed = seed;
      rng.first_real = first_real + r;
    } else {
      if (!scratch_w->p) GSP_CUDA_OK(ctx, scratch_w->alloc(d->dc->dev, (size_t)p->N * rb * sizeof(double)));
      GSP_CUDA_OK(ctx, launch_rng_fill(d->dc->stream, d->dc->sms, scratch_w->as<double>(), p->N, p->N, nb, seed, 0,
                                       (unsigned long long)(first_real + r), false));
      wr = scratch_w->as<double>();
    }
    if (n_inds > 0) {
      if (!scratch_z->p) GSP_CUDA_OK(ctx, scratch_z->alloc(d->dc->dev, (size_t)p->N * rb * sizeof(double)));
      GSP_CUDA_OK(ctx, realization_batch(d, p, wr, scratch_z->as<double>(), nb, s, 1.0, mu, rng));
      long long blocks = (n_inds * nb + 255) / 256;
      if (blocks > (long long)d->dc->sms * 8) blocks = (long long)d->dc->sms * 8;
      GSP_LAUNCH(gather_kernel, dim3((unsigned)blocks), dim3(256), 0, d->dc->stream, scratch_z->as<double>(), p->N, inds_dev, n_inds,
                 nb, out + r * n_inds);
      g_launches++;
      GSP_CUDA_OK(ctx, cudaGetLastError());
    } else {
      GSP_CUDA_OK(ctx, realization_batch(d, p, wr, out + r * p->N, nb, s, 1.0, mu, rng));
    }
  }
  // conditioning runs over larger chunks than the simulation: a node's weight row (12 kk bytes) is read once per chunk
  if (p->cond) {
    const long long nout = n_inds > 0 ? n_inds : p->N, cb = KRIGE_RB;
    for (long long r = 0; r < R; r += cb) GSP_TRY(apply_conditioning(p, d, out + r * nout, (R - r < cb) ? R - r : cb, mu));
  }
  return GSP_OK;
}

}  // namespace
}  // namespace gsp

extern "C" int gsp_fft_sample_dev(gsp_fft_plan* p, int64_t R, const double* w, uint64_t seed, int64_t first_real, double sill,
                                  double mu, int64_t n_inds, const int64_t* inds_dev, double* out) {
  if (!p) return -1;
  gsp_ctx* ctx = p->ctx;
  std::lock_guard<std::mutex> lk(p->mu);
  if (R < 0) return set_err(ctx, -2, "R < 0");
  if (!(sill > 0.0)) return set_err(ctx, -6, "sill must be positive");
  if (!out) return set_err(ctx, -10, "out is NULL");
  if (n_inds > 0 && !inds_dev) return set_err(ctx, -9, "inds is NULL");
  GSP_TRY(check_cond_args(p, mu, n_inds));
  FftDev* d = p->dev[0].get();
  cudaSetDevice(d->dc->dev);
  DevBuf sw, sz;
  GSP_CUDA_OK(ctx, cudaEventRecord(d->ev0, d->dc->stream));
  GSP_TRY(sample_on_device(p, d, R, w, seed, first_real, sill, mu, n_inds, (const long long*)inds_dev, out, &sw, &sz));
  GSP_CUDA_OK(ctx, cudaEventRecord(d->ev1, d->dc->stream));
  GSP_CUDA_OK(ctx, cudaStreamSynchronize(d->dc->stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, d->ev0, d->ev1);
  ctx->last_sample_ms = ms;
  return GSP_OK;
}

namespace gsp {
namespace {

// Host-facing sampling core.  Realizations are sharded contiguously over the devices of the context; the result goes either
// to the caller's host buffer `out` (3-stream H2D / compute / D2H pipeline per device) or stays on the devices in `ens`
// (same sharding rule as gsp_ensemble_create), in which case nothing but the noise (if injected) crosses PCIe.
int fft_sample_impl(gsp_fft_plan* p, int64_t R, const double* w, uint64_t seed, int64_t first_real, double sill, double mu, int64_t n_inds,
                    const int64_t* inds, double* out, gsp_ensemble* ens) {
  gsp_ctx* ctx = p->ctx;
  if (R < 0) return set_err(ctx, -2, "R < 0");
  if (!(sill > 0.0)) return set_err(ctx, -6, "sill must be positive");
  if (!out && !ens) return set_err(ctx, -10, "out is NULL");
  if (n_inds > 0 && !inds) return set_err(ctx, -9, "inds is NULL");
  GSP_TRY(check_cond_args(p, mu, n_inds, inds));
  if (n_inds > 0)
    for (long long q = 0; q < n_inds; ++q)
      if (inds[q] < 1 || inds[q] > p->N) return set_err(ctx, -9, "inds out of range (1-based parent indices)");
  const long long nout = n_inds > 0 ? n_inds : p->N;
  const int ndev = (int)p->dev.size();
  if (ens) {
    if (ens->ctx != ctx || (int)ens->dev.size() != ndev) return set_err(ctx, -11, "the ensemble belongs to a different context");
    if (ens->n != nout || ens->R != R) return set_err(ctx, -11, "ensemble shape does not match (n = n_inds or prod(dims), R realizations)");
  }
  // host pipeline granularity: 3-D grids move one realization (8 N bytes each way) per step - PCIe is the bound there and a
  // finer grain overlaps better; 1-D / 2-D grids move a batch
  const long long hrb = p->ndim == 3 ? 1 : p->rb;
  // contiguous shards of realizations per device (the reference shards over worker processes, field.jl:103-121)
  std::vector<long long> r0(ndev + 1, 0);
  for (int i = 0; i < ndev; ++i) r0[i + 1] = r0[i] + (R / ndev) + (i < R % ndev ? 1 : 0);
  // upload indices
  for (int i = 0; i < ndev; ++i) {
    FftDev* d = p->dev[i].get();
    cudaSetDevice(d->dc->dev);
    if (n_inds > 0) {
      if (d->inds_cap < n_inds) {
        GSP_CUDA_OK(ctx, d->inds.alloc(d->dc->dev, (size_t)n_inds * sizeof(long long)));
        d->inds_cap = n_inds;
      }
      GSP_CUDA_OK(ctx, cudaMemcpyAsync(d->inds.p, inds, (size_t)n_inds * sizeof(long long), cudaMemcpyHostToDevice, d->dc->stream));
    }
    for (int k = 0; k < 2; ++k) {
      if (w && !d->win[k].p) GSP_CUDA_OK(ctx, d->win[k].alloc(d->dc->dev, (size_t)p->N * hrb * sizeof(double)));
      if (!ens && (!d->zout[k].p || d->zout[k].bytes < (size_t)nout * hrb * sizeof(double)))
        GSP_CUDA_OK(ctx, d->zout[k].alloc(d->dc->dev, (size_t)nout * hrb * sizeof(double)));
    }
  }
  // software pipeline per device: H2D(r+1) | compute(r) | D2H(r-1) on three streams, double-buffered
  std::vector<std::vector<cudaEvent_t>> ev(ndev);
  int rc = GSP_OK;
  std::vector<DevBuf> sw(ndev), sz(ndev);
  long long maxshard = 0;
  for (int i = 0; i < ndev; ++i) maxshard = std::max(maxshard, r0[i + 1] - r0[i]);
  // events: per device, per buffer: in_ready, compute_done, out_drained
  struct Ev { cudaEvent_t in_ready[2], done[2], drained[2], in_free[2]; };
  std::vector<Ev> evs(ndev);
  for (int i = 0; i < ndev; ++i) {
    cudaSetDevice(p->dev[i]->dc->dev);
    for (int k = 0; k < 2; ++k) {
      cudaEventCreateWithFlags(&evs[i].in_ready[k], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&evs[i].done[k], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&evs[i].drained[k], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&evs[i].in_free[k], cudaEventDisableTiming);
    }
  }
  {
    FftDev* d0 = p->dev[0].get();
    cudaSetDevice(d0->dc->dev);
    cudaEventRecord(d0->ev0, d0->dc->stream);
  }
  // one pipeline step = one chunk of up to hrb realizations (1 for 3-D grids, a batch for 1-D / 2-D grids); a resident
  // ensemble with on-device noise has nothing to pipeline: the whole shard is one step (all lanes busy)
  const long long rb = (ens && !w) ? std::max<long long>(maxshard, 1) : hrb;
  const long long nsteps = (maxshard + rb - 1) / rb;
  for (long long step = 0; step < nsteps && rc == GSP_OK; ++step) {
    for (int i = 0; i < ndev && rc == GSP_OK; ++i) {
      const long long nloc = r0[i + 1] - r0[i];
      if (step * rb >= nloc) continue;
      FftDev* d = p->dev[i].get();
      cudaSetDevice(d->dc->dev);
      const int k = (int)(step & 1);
      const long long r = r0[i] + step * rb;
      const long long nbk = (nloc - step * rb < rb) ? nloc - step * rb : rb;
      const double* wdev = nullptr;
      if (w) {
        if (step >= 2) cudaStreamWaitEvent(d->dc->h2d, evs[i].in_free[k], 0);
        cudaError_t e = cudaMemcpyAsync(d->win[k].p, w + r * p->N, (size_t)p->N * nbk * sizeof(double), cudaMemcpyHostToDevice, d->dc->h2d);
        if (e != cudaSuccess) { rc = set_err(ctx, GSP_E_CUDA, cudaGetErrorString(e)); break; }
        cudaEventRecord(evs[i].in_ready[k], d->dc->h2d);
        cudaStreamWaitEvent(d->dc->stream, evs[i].in_ready[k], 0);
        wdev = d->win[k].as<double>();
      }
      if (ens) {
        rc = sample_on_device(p, d, nbk, wdev, seed, first_real + r, sill, mu, n_inds, d->inds.as<long long>(),
                              ens->dev[i]->Z.as<double>() + step * rb * nout, &sw[i], &sz[i]);
        if (rc != GSP_OK) break;
        if (w) cudaEventRecord(evs[i].in_free[k], d->dc->stream);
        continue;
      }
      if (step >= 2) cudaStreamWaitEvent(d->dc->stream, evs[i].drained[k], 0);
      rc = sample_on_device(p, d, nbk, wdev, seed, first_real + r, sill, mu, n_inds, d->inds.as<long long>(), d->zout[k].as<double>(),
                            &sw[i], &sz[i]);
      if (rc != GSP_OK) break;
      cudaEventRecord(evs[i].done[k], d->dc->stream);
      if (w) cudaEventRecord(evs[i].in_free[k], d->dc->stream);
      cudaStreamWaitEvent(d->dc->d2h, evs[i].done[k], 0);
      cudaError_t e = cudaMemcpyAsync(out + r * nout, d->zout[k].p, (size_t)nout * nbk * sizeof(double), cudaMemcpyDeviceToHost, d->dc->d2h);
      if (e != cudaSuccess) { rc = set_err(ctx, GSP_E_CUDA, cudaGetErrorString(e)); break; }
      cudaEventRecord(evs[i].drained[k], d->dc->d2h);
    }
  }
  {
    FftDev* d0 = p->dev[0].get();
    cudaSetDevice(d0->dc->dev);
    cudaEventRecord(d0->ev1, d0->dc->stream);
  }
  for (int i = 0; i < ndev; ++i) {
    FftDev* d = p->dev[i].get();
    cudaSetDevice(d->dc->dev);
    cudaError_t e1 = cudaStreamSynchronize(d->dc->h2d);
    cudaError_t e2 = cudaStreamSynchronize(d->dc->stream);
    cudaError_t e3 = cudaStreamSynchronize(d->dc->d2h);
    if (rc == GSP_OK && (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess))
      rc = set_err(ctx, GSP_E_CUDA, std::string("fft_sample: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3)));
    for (int k = 0; k < 2; ++k) {
      cudaEventDestroy(evs[i].in_ready[k]);
      cudaEventDestroy(evs[i].done[k]);
      cudaEventDestroy(evs[i].drained[k]);
      cudaEventDestroy(evs[i].in_free[k]);
    }
  }
  if (rc == GSP_OK) {
    FftDev* d0 = p->dev[0].get();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, d0->ev0, d0->ev1);
    ctx->last_sample_ms = ms;
  }
  return rc;
}

}  // namespace
}  // namespace gsp

extern "C" int gsp_fft_sample(gsp_fft_plan* p, int64_t R, const double* w, uint64_t seed, int64_t first_real, double sill,
                              double mu, int64_t n_inds, const int64_t* inds, double* out) {
  if (!p) return -1;
  std::lock_guard<std::mutex> lk(p->mu);
  if (!out) return set_err(p->ctx, -10, "out is NULL");
  return fft_sample_impl(p, R, w, seed, first_real, sill, mu, n_inds, inds, out, nullptr);
}

extern "C" int gsp_fft_sample_ensemble(gsp_fft_plan* p, gsp_ensemble* ens, const double* w, uint64_t seed, int64_t first_real, double sill,
                                       double mu, int64_t n_inds, const int64_t* inds) {
  if (!p) return -1;
  std::lock_guard<std::mutex> lk(p->mu);
  if (!ens) return set_err(p->ctx, -2, "ensemble is NULL");
  std::lock_guard<std::mutex> lk2(ens->mu);
  return fft_sample_impl(p, ens->R, w, seed, first_real, sill, mu, n_inds, inds, nullptr, ens);
}

// ------------------------------------------------------------------ conditional FFTSIM (fftsim.jl:94-101,140-153)
extern "C" int gsp_fft_plan_condition(gsp_fft_plan* p, double mu, int32_t minneighbors, int32_t maxneighbors, int64_t nd,
                                      const double* dcoords, const double* dvals, int64_t nk, const int64_t* knodes, int64_t n_inds,
                                      const int64_t* inds) {
  if (!p) return -1;
  gsp_ctx* ctx = p->ctx;
  std::lock_guard<std::mutex> lk(p->mu);
  if (!std::isfinite(mu)) return set_err(ctx, -2, "mu is not finite");
  if (nd < 1 || !dcoords || !dvals) return set_err(ctx, -5, "conditioning data: nd >= 1, dcoords and dvals are required");
  if (nk < 1 || !knodes) return set_err(ctx, -8, "data nodes: nk >= 1 and knodes are required");
  if (n_inds < 0 || (n_inds > 0 && !inds)) return set_err(ctx, -10, "inds is NULL");
  const long long n = n_inds > 0 ? n_inds : p->N;
  for (long long q = 0; q < n_inds; ++q)
    if (inds[q] < 1 || inds[q] > p->N) return set_err(ctx, -10, "inds out of range (1-based parent indices)");
  for (long long j = 0; j < nk; ++j)
    if (knodes[j] < 1 || knodes[j] > n || (j > 0 && knodes[j] <= knodes[j - 1]))
      return set_err(ctx, -8, "knodes must be strictly ascending 1-based positions within the simulation domain (findall(mask))");
  // GeoStatsModels.fitpredict fixes the limits the same way: maxneighbors outside [1, nobs] -> nobs, minneighbors outside [1, max] -> 1
  long long kmax_d = maxneighbors, kmax_k = maxneighbors;
  if (kmax_d > nd || kmax_d < 1) kmax_d = nd;
  if (kmax_k > nk || kmax_k < 1) kmax_k = nk;
  (void)minneighbors;  // a k-nearest search always returns min(k, nobs) >= 1 neighbours: the minimum never binds
  if (kmax_d > KRIGE_MAXK || kmax_k > KRIGE_MAXK)
    return set_err(ctx, GSP_E_UNSUPPORTED, "maxneighbors (after clamping to the number of data) must be <= 32");
  for (long long j = 0; j < nd; ++j)
    if (!std::isfinite(dvals[j])) return set_err(ctx, -6, "dvals must be finite (drop missing rows before the call)");
  // centroids of the data nodes: the samples of the second Kriging (view(sdom, dinds), fftsim.jl:143)
  std::vector<double> kc((size_t)nk * p->ndim);
  std::vector<long long> k0((size_t)nk);
  for (long long j = 0; j < nk; ++j) {
    const long long pos = knodes[j] - 1;
    long long e = n_inds > 0 ? inds[pos] - 1 : pos;
    k0[(size_t)j] = pos;
    for (int a = 0; a < p->ndim; ++a) {
      const long long ia = e % p->dom.dims[a];
      e /= p->dom.dims[a];
      kc[(size_t)j * p->ndim + a] = p->dom.origin[a] + ((double)ia + 0.5) * p->dom.spacing[a];
    }
  }
  p->cond = false;
  for (auto& dptr : p->dev) {
    FftDev* d = dptr.get();
    CondDev& c = d->cond;
    cudaSetDevice(d->dc->dev);
    cudaStream_t st = d->dc->stream;
    DevBuf dx, dv, kx, di, info;
    GSP_CUDA_OK(ctx, dx.alloc(d->dc->dev, (size_t)nd * p->ndim * sizeof(double)));
    GSP_CUDA_OK(ctx, dv.alloc(d->dc->dev, (size_t)nd * sizeof(double)));
    GSP_CUDA_OK(ctx, kx.alloc(d->dc->dev, kc.size() * sizeof(double)));
    GSP_CUDA_OK(ctx, info.alloc(d->dc->dev, sizeof(int)));
    GSP_CUDA_OK(ctx, cudaMemcpyAsync(dx.p, dcoords, (size_t)nd * p->ndim * sizeof(double), cudaMemcpyHostToDevice, st));
    GSP_CUDA_OK(ctx, cudaMemcpyAsync(dv.p, dvals, (size_t)nd * sizeof(double), cudaMemcpyHostToDevice, st));
    GSP_CUDA_OK(ctx, cudaMemcpyAsync(kx.p, kc.data(), kc.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    GSP_CUDA_OK(ctx, cudaMemsetAsync(info.p, 0, sizeof(int), st));
    if (n_inds > 0) {
      GSP_CUDA_OK(ctx, di.alloc(d->dc->dev, (size_t)n_inds * sizeof(long long)));
      GSP_CUDA_OK(ctx, cudaMemcpyAsync(di.p, inds, (size_t)n_inds * sizeof(long long), cudaMemcpyHostToDevice, st));
    }
    GSP_CUDA_OK(ctx, c.zbar.alloc(d->dc->dev, (size_t)n * sizeof(double)));
    const size_t nslots = (size_t)((n + 31) / 32) * 32 * (size_t)kmax_k;  // tile-major weight table, whole tiles
    GSP_CUDA_OK(ctx, c.lam.alloc(d->dc->dev, nslots * sizeof(double)));
    GSP_CUDA_OK(ctx, c.nbr.alloc(d->dc->dev, nslots * sizeof(int)));
    GSP_CUDA_OK(ctx, c.knodes.alloc(d->dc->dev, (size_t)nk * sizeof(long long)));
    GSP_CUDA_OK(ctx, cudaMemcpyAsync(c.knodes.p, k0.data(), (size_t)nk * sizeof(long long), cudaMemcpyHostToDevice, st));
    const unsigned blocks = (unsigned)((n + KW_THREADS - 1) / KW_THREADS);
    const long long* dinds = n_inds > 0 ? di.as<long long>() : nullptr;
    // sample-to-sample covariances of a Kriging, assembled once (up to 4,096 samples = 128 MB; beyond that the kernel evaluates them)
    DevBuf ktab;
    auto sample_table = [&](const double* coords, long long ns) -> const double* {
      const char* tenv = getenv("GSP_KRIGE_TABLE");  // =0: evaluate the covariances inside the kernel (the path of > 4,096 samples)
      if (ns > 4096 || (tenv && tenv[0] == '0')) return nullptr;
      if (ktab.bytes < (size_t)ns * ns * sizeof(double) && ktab.alloc(d->dc->dev, (size_t)ns * ns * sizeof(double)) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
      }
      DomDev sd{};
      sd.kind = 0; sd.dim = p->ndim; sd.nelems = ns; sd.coords = coords;
      launch_assemble(st, p->cov, sd, sd, nullptr, nullptr, ns, ns, ktab.as<double>(), ns, false);
      return ktab.as<double>();
    };
    {
      const double* kt = sample_table(dx.as<double>(), nd);
      ProfScope prof_("krige_weights", st);
      // first Kriging: the data where they are -> zbar (fftsim.jl:99); nothing else of it is needed later
      GSP_LAUNCH(krige_weights_kernel, dim3(blocks), dim3(KW_THREADS), 0, st, p->cov, p->dom, dinds, n, (int)kmax_d, (long long)nd,
                 (const double*)dx.as<double>(), (const double*)dv.as<double>(), mu, c.zbar.as<double>(), (double*)nullptr, (int*)nullptr,
                 info.as<int>(), kt);
      g_launches++;
    }
    {
      const double* kt = sample_table(kx.as<double>(), nk);
      ProfScope prof_("krige_weights", st);
      // second Kriging: samples at the centroids of the data nodes -> weight table, applied to every realization
      GSP_LAUNCH(krige_weights_kernel, dim3(blocks), dim3(KW_THREADS), 0, st, p->cov, p->dom, dinds, n, (int)kmax_k, (long long)nk,
                 (const double*)kx.as<double>(), (const double*)nullptr, mu, (double*)nullptr, c.lam.as<double>(), c.nbr.as<int>(), info.as<int>(),
                 kt);
      g_launches++;
    }
    GSP_CUDA_OK(ctx, cudaGetLastError());
    int h_info = 0;
    GSP_CUDA_OK(ctx, cudaMemcpyAsync(&h_info, info.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    GSP_CUDA_OK(ctx, cudaStreamSynchronize(st));
    if (h_info != 0)
      return set_err(ctx, GSP_E_STATE, "Kriging matrix of element " + std::to_string(h_info) +
                                           " is not positive definite (coincident data locations?): cholesky would throw PosDefException");
  }
  p->cond = true;
  p->cond_mu = mu;
  p->cond_n = n;
  p->cond_nk = nk;
  p->cond_ninds = n_inds;
  p->cond_inds_hash = n_inds > 0 ? hash_inds(inds, n_inds) : 0;
  p->cond_kk = (int)kmax_k;
  return GSP_OK;
}

extern "C" int gsp_fft_plan_condmean(gsp_fft_plan* p, double* zbar) {
  if (!p) return -1;
  gsp_ctx* ctx = p->ctx;
  std::lock_guard<std::mutex> lk(p->mu);
  if (!zbar) return set_err(ctx, -2, "zbar is NULL");
  if (!p->cond) return set_err(ctx, GSP_E_STATE, "the plan is unconditional (call gsp_fft_plan_condition first)");
  FftDev* d = p->dev[0].get();
  cudaSetDevice(d->dc->dev);
  GSP_CUDA_OK(ctx, cudaMemcpyAsync(zbar, d->cond.zbar.p, (size_t)p->cond_n * sizeof(double), cudaMemcpyDeviceToHost, d->dc->stream));
  GSP_CUDA_OK(ctx, cudaStreamSynchronize(d->dc->stream));
  return GSP_OK;
}

