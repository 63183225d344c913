// LUSIM on the device: plan = joint covariance assembly over [data; simulation]
// nodes + blocked Cholesky + conditional mean d2 (lusim.jl:38-110); sample = L22 * W for all
// realizations with the fused epilogue d2 + ., scatter to sinds, data rows, +mu (lusim.jl:112-175).
//
// ONE Cholesky of the joint matrix ordered [dinds; sinds] replaces the reference's block algebra
// (lusim.jl:95-103): with K = [C11 C12; C21 C22] = L L',  L = [L11 0; A21 L22] where
// A21 = (L11 \ C12)' and L22 = chol(C22 - A21 B12), and d2 = A21 (L11 \ z1).
//
// Layout in HBM: joint matrix (Np x Np doubles, column-major, only lower tiles touched) with
// the data block padded to a multiple of 128 by identity rows ("dummy data", z = 0) so that L22
// starts on a tile boundary, and the simulation block padded likewise at the end.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <memory>

#include "chol.h"
#include "ensemble.h"
#include "rng.cuh"

namespace gsp {

// Wp = rho * W1p + sqrt(1 - rho^2) * Wp   (lusim.jl:164)
__global__ void __launch_bounds__(256) premix_kernel(double* __restrict__ Wp, const double* __restrict__ W1p, long long n, double rho,
                                                     double c2) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) Wp[i] = rho * W1p[i] + c2 * Wp[i];
}

// Z[dinds[j] + r*ldz] = z1[j]   (lusim.jl:168; data honoured exactly)
__global__ void __launch_bounds__(256) scatter_data_kernel(double* __restrict__ Z, long long ldz, const long long* __restrict__ dinds,
                                                           const double* __restrict__ z1, long long nd, long long R) {
  const long long total = nd * R;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const long long r = t / nd, j = t - r * nd;
    Z[dinds[j] + r * ldz] = z1[j];
  }
}

// realization-major staging of the noise: dst[r + i*ldd] = (i < rows && r < cols) ? src[i + r*lds] : 0
// for i < rows_pad, r < cols_pad (32x32 shared-memory transpose; both sides coalesced)
__global__ void __launch_bounds__(256) pad_transpose_kernel(double* __restrict__ dst, long long ldd, long long rows_pad, long long cols_pad,
                                                            const double* __restrict__ src, long long lds, long long rows, long long cols) {
  __shared__ double tile[32][33];
  const long long i0 = (long long)blockIdx.x * 32, r0 = (long long)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int rr = ty; rr < 32; rr += 8) {
    const long long i = i0 + tx, r = r0 + rr;
    tile[rr][tx] = (i < rows && r < cols) ? src[i + r * lds] : 0.0;
  }
  __syncthreads();
  for (int ii = ty; ii < 32; ii += 8) {
    const long long i = i0 + ii, r = r0 + tx;
    if (i < rows_pad && r < cols_pad) dst[r + i * ldd] = tile[tx][ii];
  }
}

namespace {

struct LuDev {
  DevCtx* dc = nullptr;
  std::shared_ptr<DevBuf> A;     // Np x Np joint matrix -> L   (shared with the plans made by gsp_lu_plan_create_like)
  std::shared_ptr<DevBuf> invD;  // inverses of the diagonal 128-blocks
  DevBuf rows;    // block rows this device owns (distributed factorization)
  DevBuf flags;   // inter-CTA flags of the fused diagonal-square kernel
  DevBuf d2;      // Ns_pad
  DevBuf sinds;   // Ns (0-based)
  DevBuf dinds;   // Nd (0-based)
  DevBuf z1;      // Nd
  DevBuf info;
  // sampling scratch (per chunk)
  DevBuf Wp, W1p, Zc, Wraw, W1raw;
  DevBuf Zc2, Wraw2, W1raw2;   // second slot of the host-pointer pipeline (H2D / compute / D2H of consecutive chunks overlap)
  long long chunk_cols = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
};

inline unsigned grid_for(long long n, int sms) {
  long long b = (n + 255) / 256;
  if (b > (long long)sms * 16) b = (long long)sms * 16;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace
}  // namespace gsp

using namespace gsp;

struct gsp_lu_plan {
  gsp_ctx* ctx = nullptr;
  long long N = 0, Nd = 0, Ns = 0, Ndp = 0, Nsp = 0, Np = 0;
  double mu = 0.0;
  double t_assemble_ms = 0.0, t_factor_ms = 0.0, t_solve_ms = 0.0;  // device times of the plan stages (CUDA events)
  std::vector<std::unique_ptr<LuDev>> dev;
  std::vector<long long> h_sinds, h_dind0;  // 0-based simulation / data nodes (kept for gsp_lu_plan_create_like)
  std::mutex mu_lock;
};

namespace gsp {
namespace {

int ensure_chunk(gsp_ctx* ctx, gsp_lu_plan* p, LuDev* d, long long cols, bool need_w1, bool two_slots = false) {
  const long long cpad = round_up(cols, 128);
  if (d->chunk_cols < cpad) {
    GSP_CUDA_OK(ctx, d->Wp.alloc(d->dc->dev, (size_t)p->Nsp * cpad * sizeof(double)));
    GSP_CUDA_OK(ctx, d->Zc.alloc(d->dc->dev, (size_t)p->N * cpad * sizeof(double)));
    GSP_CUDA_OK(ctx, d->Wraw.alloc(d->dc->dev, (size_t)p->Ns * cpad * sizeof(double)));
    d->W1p.release();
    d->W1raw.release();
    d->Zc2.release();
    d->Wraw2.release();
    d->W1raw2.release();
    d->chunk_cols = cpad;
  }
  if (need_w1 && !d->W1p.p) {
    GSP_CUDA_OK(ctx, d->W1p.alloc(d->dc->dev, (size_t)p->Nsp * d->chunk_cols * sizeof(double)));
    GSP_CUDA_OK(ctx, d->W1raw.alloc(d->dc->dev, (size_t)p->Ns * d->chunk_cols * sizeof(double)));
  }
  if (two_slots) {
    if (!d->Zc2.p) GSP_CUDA_OK(ctx, d->Zc2.alloc(d->dc->dev, (size_t)p->N * d->chunk_cols * sizeof(double)));
    if (!d->Wraw2.p) GSP_CUDA_OK(ctx, d->Wraw2.alloc(d->dc->dev, (size_t)p->Ns * d->chunk_cols * sizeof(double)));
    if (need_w1 && !d->W1raw2.p) GSP_CUDA_OK(ctx, d->W1raw2.alloc(d->dc->dev, (size_t)p->Ns * d->chunk_cols * sizeof(double)));
    for (int k = 0; k < 2; ++k) {
      if (!d->ev_in[k]) GSP_CUDA_OK(ctx, cudaEventCreateWithFlags(&d->ev_in[k], cudaEventDisableTiming));
      if (!d->ev_comp[k]) GSP_CUDA_OK(ctx, cudaEventCreateWithFlags(&d->ev_comp[k], cudaEventDisableTiming));
      if (!d->ev_out[k]) GSP_CUDA_OK(ctx, cudaEventCreateWithFlags(&d->ev_out[k], cudaEventDisableTiming));
    }
  }
  return GSP_OK;
}

// Build the padded, realization-major noise Wt (cpad x Nsp, ld = cpad) for `cols` realizations starting at `real0`.
// src: device pointer (ld = lds) or NULL => Philox stream `stream`.
int stage_noise(gsp_ctx* ctx, gsp_lu_plan* p, LuDev* d, double* Wp, const double* src, long long lds, long long cols, long long cpad,
                unsigned long long seed, unsigned stream, long long real0) {
  cudaStream_t st = d->dc->stream;
  if (src) {
    GSP_DEP_ACCESS(src, 0, 1, 0, 1, false);
    GSP_DEP_ACCESS(Wp, 0, 1, 0, 1, true);
    ProfScope prof_("pad_transpose", st);
    GSP_LAUNCH(pad_transpose_kernel, dim3((unsigned)(p->Nsp / 32), (unsigned)(cpad / 32)), dim3(256), 0, st, Wp, cpad, p->Nsp, cpad, src, lds, p->Ns,
               cols);
    g_launches++;
  } else {
    GSP_CUDA_OK(ctx, cudaMemsetAsync(Wp, 0, (size_t)p->Nsp * cpad * sizeof(double), st));
    GSP_CUDA_OK(ctx, launch_rng_fill(st, d->dc->sms, Wp, p->Ns, cpad, cols, seed, stream, (unsigned long long)real0, true, true));
  }
  GSP_CUDA_OK(ctx, cudaGetLastError());
  return GSP_OK;
}

// all-device-pointer core: `cols` realizations into Z (ldz), on device d
int sample_core(gsp_ctx* ctx, gsp_lu_plan* p, LuDev* d, long long cols, const double* W, long long ldw, unsigned long long seed,
                int stream, long long real0, double rho, const double* W1, double* Z, long long ldz) {
  const bool mix = !std::isnan(rho);
  const long long cpad = round_up(cols, 128);
  cudaStream_t st = d->dc->stream;
  double* Wp = d->Wp.as<double>();
  GSP_TRY(stage_noise(ctx, p, d, Wp, W, ldw, cols, cpad, seed, (unsigned)stream, real0));
  if (mix) {
    double* W1p = d->W1p.as<double>();
    GSP_TRY(stage_noise(ctx, p, d, W1p, W1, ldw, cols, cpad, seed, 0u, real0));
    const long long n = p->Nsp * cpad;
    GSP_DEP_ACCESS(W1p, 0, 1, 0, 1, false);
    GSP_DEP_ACCESS(Wp, 0, 1, 0, 1, true);
    GSP_LAUNCH(premix_kernel, dim3(grid_for(n, d->dc->sms)), dim3(256), 0, st, Wp, (const double*)W1p, n, rho, std::sqrt(1.0 - rho * rho));
    g_launches++;
    GSP_CUDA_OK(ctx, cudaGetLastError());
  }
  const double* L22 = d->A->as<double>() + p->Ndp * (p->Np + 1);
  const double addmu = (p->Nd == 0) ? p->mu : 0.0;  // lusim.jl:172
  GSP_DEP_ACCESS(Wp, 0, 1, 0, 1, false);
  GSP_DEP_ACCESS(Z, 0, 1, 0, 1, true);
  GSP_CUDA_OK(ctx, sample_gemm(st, L22, p->Np, (int)(p->Nsp / 128), Wp, cpad, (int)(cpad / 128), Z, ldz, d->d2.as<double>(),
                               d->sinds.as<long long>(), addmu, p->Ns, cols));
  if (p->Nd > 0) {
    GSP_DEP_ACCESS(Z, 0, 1, 0, 1, true);
    GSP_LAUNCH(scatter_data_kernel, dim3(grid_for(p->Nd * cols, d->dc->sms)), dim3(256), 0, st, Z, ldz, (const long long*)d->dinds.as<long long>(),
               (const double*)d->z1.as<double>(), p->Nd, cols);
    g_launches++;
    GSP_CUDA_OK(ctx, cudaGetLastError());
  }
  return GSP_OK;
}

}  // namespace
}  // namespace gsp

namespace gsp {
namespace {

// RAII for the timing events of gsp_lu_plan_create (every return path destroys them)
struct EventSet {
  std::vector<cudaEvent_t> ev;
  ~EventSet() {
    for (cudaEvent_t e : ev)
      if (e) cudaEventDestroy(e);
  }
  cudaError_t make(int n) {
    ev.assign((size_t)n, nullptr);
    for (auto& e : ev) {
      cudaError_t rc = cudaEventCreate(&e);
      if (rc != cudaSuccess) return rc;
    }
    return cudaSuccess;
  }
};

void destroy_plan_events(gsp_lu_plan* p) {
  for (auto& d : p->dev) {
    cudaSetDevice(d->dc->dev);
    cudaStreamSynchronize(d->dc->stream);
    for (cudaEvent_t* e : {&d->ev0, &d->ev1, &d->ev_in[0], &d->ev_in[1], &d->ev_comp[0], &d->ev_comp[1], &d->ev_out[0], &d->ev_out[1]})
      if (*e) {
        cudaEventDestroy(*e);
        *e = nullptr;
      }
  }
}

// Which factorization runs (GSP_CHOL_ALGO): "panel" = chol_factor_dist (row-panel ownership over the devices of the context, also on
// one device), "recursive" = chol_factor on device 0 (+ copy of L to the other devices).  Default: panel whenever the matrix has
// at least GSP_CHOL_DIST_MIN_BLOCKS 128-blocks (below that the panel chain dominates and the recursion on one device is as fast).
// d2 = A21 * (L11 \ z1) on device 0 (lusim.jl:102); zero when unconditional (lusim.jl:91).  Enqueued on device 0's stream.
int compute_d2(gsp_ctx* ctx, gsp_lu_plan* p, const double* z1) {
  LuDev* d0 = p->dev[0].get();
  DevCtx& dc = *d0->dc;
  cudaSetDevice(dc.dev);
  cudaStream_t st = dc.stream;
  DevBuf y;
  GSP_CUDA_OK(ctx, cudaMemsetAsync(d0->d2.p, 0, (size_t)p->Nsp * sizeof(double), st));
  if (p->Nd > 0) {
    GSP_CUDA_OK(ctx, y.alloc(dc.dev, (size_t)p->Ndp * sizeof(double), st));
    GSP_CUDA_OK(ctx, cudaMemsetAsync(y.p, 0, (size_t)p->Ndp * sizeof(double), st));
    GSP_CUDA_OK(ctx, cudaMemcpyAsync(y.p, z1, (size_t)p->Nd * sizeof(double), cudaMemcpyHostToDevice, st));
    GSP_CUDA_OK(ctx, chol_forward_solve(st, d0->A->as<double>(), p->Np, d0->invD->as<double>(), (int)(p->Ndp / 128), y.as<double>()));
    GSP_CUDA_OK(ctx, chol_gemv_rows(st, d0->A->as<double>(), p->Np, p->Ndp, p->Nsp, (int)p->Ndp, y.as<double>(), d0->d2.as<double>()));
  }
  return GSP_OK;  // y is freed stream-ordered on st
}

struct CholChoice {
  bool dist;
  int PB;
};
CholChoice choose_chol(int ndev, int nb) {
  static const char* algo = getenv("GSP_CHOL_ALGO");
  static const int min_blocks = getenv("GSP_CHOL_DIST_MIN_BLOCKS") ? atoi(getenv("GSP_CHOL_DIST_MIN_BLOCKS")) : 24;
  static const int min_blocks_1 = getenv("GSP_CHOL_PANEL_MIN_BLOCKS") ? atoi(getenv("GSP_CHOL_PANEL_MIN_BLOCKS")) : 48;
  static const int pb_env = getenv("GSP_CHOL_PB") ? atoi(getenv("GSP_CHOL_PB")) : 0;
  CholChoice c{};
  if (algo && algo[0] == 'r') c.dist = false;
  else if (algo && algo[0] == 'p') c.dist = true;
  else c.dist = ndev > 1 ? nb >= min_blocks : nb >= min_blocks_1;  // one device: measured C3 51.9 vs 59.4 ms, 32k nodes 359 vs 367 ms
  // panel width in 128-blocks, measured on B200: 4 everywhere (C3 51.8 vs 52.4 ms on one GPU, C5 63 vs 64 ms and C3 17.5 vs 17.7 ms
  // on eight) except large matrices on ONE device, where 8 wins (32k nodes: 359 vs 371 ms)
  c.PB = pb_env > 0 ? std::min(pb_env, 8) : ((ndev == 1 && nb >= 192) ? 8 : 4);
  if (c.PB > nb) c.PB = nb;
  return c;
}

}  // namespace
}  // namespace gsp

extern "C" int gsp_lu_plan_create(gsp_ctx* ctx, const gsp_cov_model* cov, const gsp_domain* dom, int64_t nd, const int64_t* dinds,
                                  const double* z1, double mu, gsp_lu_plan** out) {
  if (!ctx) return -1;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!out) return set_err(ctx, -8, "out is NULL");
  *out = nullptr;
  DomDev dd;
  GSP_TRY(make_dom_dev(ctx, dom, 3, &dd));
  CovDev cd;
  GSP_TRY(make_cov_dev(ctx, cov, dom->dim, 2, &cd));
  const long long N = dd.nelems;
  if (nd < 0 || nd >= N) return set_err(ctx, -4, "nd must satisfy 0 <= nd < nelems");
  if (nd > 0 && (!dinds || !z1)) return set_err(ctx, nd > 0 && !dinds ? -5 : -6, "dinds / z1 is NULL");
  for (long long j = 0; j < nd; ++j) {
    if (dinds[j] < 1 || dinds[j] > N) return set_err(ctx, -5, "dinds out of range (1-based)");
    if (j > 0 && dinds[j] <= dinds[j - 1]) return set_err(ctx, -5, "dinds must be strictly ascending (findall(mask))");
  }
  // the plan owns CUDA events: on every error return below they are destroyed with it
  struct PlanGuard {
    gsp_lu_plan* p;
    ~PlanGuard() {
      if (p) {
        destroy_plan_events(p);
        delete p;
      }
    }
  } guard{new gsp_lu_plan};
  gsp_lu_plan* p = guard.p;
  p->ctx = ctx;
  p->N = N;
  p->Nd = nd;
  p->Ns = N - nd;
  p->Ndp = round_up(nd, 128);
  p->Nsp = round_up(p->Ns, 128);
  p->Np = p->Ndp + p->Nsp;
  p->mu = mu;

  // host index maps (0-based): perm = [dinds; pad; sinds; pad]
  std::vector<long long> perm((size_t)p->Np, -1), sinds((size_t)p->Ns), dind0((size_t)nd);
  {
    long long j = 0, s = 0;
    for (long long e = 0; e < N; ++e) {
      if (j < nd && dinds[j] - 1 == e) {
        dind0[(size_t)j] = e;
        perm[(size_t)j] = e;
        ++j;
      } else {
        sinds[(size_t)s] = e;
        perm[(size_t)(p->Ndp + s)] = e;
        ++s;
      }
    }
  }

  p->h_sinds = sinds;
  p->h_dind0 = dind0;
  const int ndev = (int)ctx->devs.size();
  const int nb = (int)(p->Np / 128);
  CholChoice cc = choose_chol(ndev, nb);
  if (ndev > 1 && !ctx->peer_ok) cc.dist = false;  // peer stores need peer access; without it: factor on device 0, copy L (cudaMemcpyPeer)
  const int nbuild = cc.dist ? ndev : 1;   // devices that take part in the factorization
#ifdef GSP_EMU
  {  // test-only: record the stream / event graph of this plan build and check it (chol_factor_dist)
    const char* env = getenv("GSP_DEPCHECK");
    emu::dep_enable(env && env[0] == '1');
  }
#endif

  EventSet tev;
  cudaSetDevice(ctx->devs[0].dev);
  GSP_CUDA_OK(ctx, tev.make(4));
  GSP_CUDA_OK(ctx, cudaEventRecord(tev.ev[0], ctx->devs[0].stream));

  // ---- per-device buffers; a1: joint covariance (lusim.jl:88,95,96) - with the distributed factorization every device assembles
  // only the block rows it owns (lower part), otherwise device 0 assembles the lower tiles of the whole matrix
  for (int i = 0; i < ndev; ++i) {
    std::unique_ptr<LuDev> dptr(new LuDev);
    LuDev* d = dptr.get();
    d->dc = &ctx->devs[i];
    p->dev.push_back(std::move(dptr));
    DevCtx& dc = *d->dc;
    cudaSetDevice(dc.dev);
    cudaStream_t st = dc.stream;
    GSP_CUDA_OK(ctx, d->d2.alloc(dc.dev, (size_t)p->Nsp * sizeof(double), st));
    GSP_CUDA_OK(ctx, d->sinds.alloc(dc.dev, (size_t)p->Ns * sizeof(long long), st));
    GSP_CUDA_OK(ctx, d->dinds.alloc(dc.dev, (size_t)std::max<long long>(nd, 1) * sizeof(long long), st));
    GSP_CUDA_OK(ctx, d->z1.alloc(dc.dev, (size_t)std::max<long long>(nd, 1) * sizeof(double), st));
    GSP_CUDA_OK(ctx, d->info.alloc(dc.dev, sizeof(int), st));
    GSP_CUDA_OK(ctx, cudaEventCreate(&d->ev0));
    GSP_CUDA_OK(ctx, cudaEventCreate(&d->ev1));
    GSP_CUDA_OK(ctx, cudaMemcpyAsync(d->sinds.p, sinds.data(), sinds.size() * sizeof(long long), cudaMemcpyHostToDevice, st));
    if (nd > 0) {
      GSP_CUDA_OK(ctx, cudaMemcpyAsync(d->dinds.p, dind0.data(), dind0.size() * sizeof(long long), cudaMemcpyHostToDevice, st));
      GSP_CUDA_OK(ctx, cudaMemcpyAsync(d->z1.p, z1, (size_t)nd * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    d->A = std::make_shared<DevBuf>();
    d->invD = std::make_shared<DevBuf>();
    GSP_CUDA_OK(ctx, d->A->alloc(dc.dev, (size_t)p->Np * p->Np * sizeof(double), st));
    GSP_CUDA_OK(ctx, d->invD->alloc(dc.dev, (size_t)nb * 128 * 128 * sizeof(double), st));
    if (i >= nbuild) continue;
    GSP_CUDA_OK(ctx, d->rows.alloc(dc.dev, (size_t)nb * sizeof(int), st));
    GSP_CUDA_OK(ctx, d->flags.alloc(dc.dev, (size_t)chol_dist_flag_ints() * sizeof(int), st));
    DevBuf dperm, dcoords;
    GSP_CUDA_OK(ctx, dperm.alloc(dc.dev, (size_t)p->Np * sizeof(long long), st));
    GSP_CUDA_OK(ctx, cudaMemcpyAsync(dperm.p, perm.data(), perm.size() * sizeof(long long), cudaMemcpyHostToDevice, st));
    DomDev ddl = dd;
    if (dd.kind == 0) {
      GSP_CUDA_OK(ctx, dcoords.alloc(dc.dev, (size_t)N * dd.dim * sizeof(double), st));
      GSP_CUDA_OK(ctx, cudaMemcpyAsync(dcoords.p, dom->coords, (size_t)N * dd.dim * sizeof(double), cudaMemcpyHostToDevice, st));
      ddl.coords = dcoords.as<double>();
    }
    if (!cc.dist) {
      GSP_DEP_ACCESS(d->A->as<double>(), 0, nb, 0, nb, true);
      launch_assemble(st, cd, ddl, ddl, dperm.as<long long>(), dperm.as<long long>(), p->Np, p->Np, d->A->as<double>(), p->Np, true);
    } else {
      std::vector<int> own;
      chol_dist_owned_rows(nb, cc.PB, nbuild, i, &own);
      for (size_t a = 0; a < own.size();) {  // one launch per run of consecutive own block rows: rows [r0, r1) x columns [0, r1)
        size_t b = a + 1;
        while (b < own.size() && own[b] == own[b - 1] + 1) ++b;
        const long long r0 = (long long)own[a] * 128, r1 = (long long)(own[b - 1] + 1) * 128;
        GSP_DEP_ACCESS(d->A->as<double>(), own[a], own[b - 1] + 1, 0, own[b - 1] + 1, true);
        launch_assemble(st, cd, ddl, ddl, dperm.as<long long>() + r0, dperm.as<long long>(), r1 - r0, r1, d->A->as<double>() + r0, p->Np, true, r0);
        a = b;
      }
    }
    GSP_CUDA_OK(ctx, cudaGetLastError());
    // dperm / dcoords are freed stream-ordered on st when they go out of scope; the pageable host sources were consumed by the copies
  }
  LuDev* d0 = p->dev[0].get();
  DevCtx& dc = *d0->dc;
  cudaSetDevice(dc.dev);
  cudaStream_t st = dc.stream;
  GSP_CUDA_OK(ctx, cudaEventRecord(tev.ev[1], st));
  // a2/a3: one joint Cholesky (lusim.jl:92 or 98-103)
  if (cc.dist) {
    std::vector<DistDev> dv;
    for (int i = 0; i < nbuild; ++i) {
      LuDev* d = p->dev[i].get();
      dv.push_back(DistDev{d->dc->dev, d->dc->stream, d->dc->aux, d->dc->side[0], d->A->as<double>(), d->invD->as<double>(), d->info.as<int>(),
                           d->rows.as<int>(), d->flags.as<int>()});
    }
    GSP_CUDA_OK(ctx, chol_factor_dist(dv, p->Np, nb, cc.PB));
    cudaSetDevice(dc.dev);
  } else {
    GSP_CUDA_OK(ctx, chol_factor(st, dc.side, DevCtx::kSide, d0->A->as<double>(), p->Np, nb, d0->invD->as<double>(), d0->info.as<int>()));
  }
  GSP_CUDA_OK(ctx, cudaEventRecord(tev.ev[2], st));
#ifdef GSP_EMU
  // test-only (GSP_DEPCHECK=1): every pair of launches of the assembly + factorization that touch the same blocks, one of them writing,
  // must be ordered by streams / events - for the panel algorithm on G devices and for the recursive one with its look-ahead streams
  if (emu::dep_enabled()) {
    const long long bad = emu::dep_check(1);
    emu::dep_enable(false);
    if (bad != 0) return set_err(ctx, GSP_E_STATE, "stream-dependency check failed: " + std::to_string(bad) + " unordered conflicting launch pairs");
  }
#endif
  GSP_TRY(compute_d2(ctx, p, z1));
  GSP_CUDA_OK(ctx, cudaEventRecord(tev.ev[3], st));
  int info = 0;
  for (int i = 0; i < nbuild; ++i) {  // the first non-positive pivot may have been met on any panel owner
    int inf_i = 0;
    cudaSetDevice(p->dev[i]->dc->dev);
    GSP_CUDA_OK(ctx, cudaMemcpyAsync(&inf_i, p->dev[i]->info.p, sizeof(int), cudaMemcpyDeviceToHost, p->dev[i]->dc->stream));
    GSP_CUDA_OK(ctx, cudaStreamSynchronize(p->dev[i]->dc->stream));
    if (inf_i > 0 && (info == 0 || inf_i < info)) info = inf_i;
  }
  cudaSetDevice(dc.dev);
  GSP_CUDA_OK(ctx, cudaStreamSynchronize(st));
  {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, tev.ev[0], tev.ev[1]); p->t_assemble_ms = ms;
    cudaEventElapsedTime(&ms, tev.ev[1], tev.ev[2]); p->t_factor_ms = ms;
    cudaEventElapsedTime(&ms, tev.ev[2], tev.ev[3]); p->t_solve_ms = ms;
  }
  if (info > 0) {
    // map the padded position back to the reference ordering [dinds; sinds] (1-based)
    long long pos = info - 1;
    long long ref = pos < p->Ndp ? pos + 1 : (pos - p->Ndp) + nd + 1;
    set_err(ctx, (int)ref, "matrix is not positive definite (PosDefException)");
    return (int)ref;
  }
  // the other devices receive d2 (and L, unless the distributed factorization already left it everywhere)
  for (int i = 1; i < ndev; ++i) {
    LuDev* e = p->dev[i].get();
    cudaSetDevice(e->dc->dev);
    if (!cc.dist) GSP_CUDA_OK(ctx, cudaMemcpyPeerAsync(e->A->p, e->dc->dev, d0->A->p, d0->dc->dev, d0->A->bytes, e->dc->stream));
    GSP_CUDA_OK(ctx, cudaMemcpyPeerAsync(e->d2.p, e->dc->dev, d0->d2.p, d0->dc->dev, d0->d2.bytes, e->dc->stream));
    GSP_CUDA_OK(ctx, cudaStreamSynchronize(e->dc->stream));
  }
  guard.p = nullptr;
  *out = p;
  return GSP_OK;
}

extern "C" int gsp_lu_plan_create_like(gsp_lu_plan* base, int64_t nd, const int64_t* dinds, const double* z1, double mu, gsp_lu_plan** out) {
  if (!base) return -1;
  gsp_ctx* ctx = base->ctx;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!out) return set_err(ctx, -6, "out is NULL");
  *out = nullptr;
  if (nd != base->Nd) return set_err(ctx, -2, "nd differs from the base plan: the factor cannot be shared");
  if (nd > 0 && (!dinds || !z1)) return set_err(ctx, !dinds ? -3 : -4, "dinds / z1 is NULL");
  for (long long j = 0; j < nd; ++j)
    if (dinds[j] - 1 != base->h_dind0[(size_t)j]) return set_err(ctx, -3, "dinds differ from the base plan: the factor cannot be shared");
  struct PlanGuard {
    gsp_lu_plan* p;
    ~PlanGuard() {
      if (p) {
        destroy_plan_events(p);
        delete p;
      }
    }
  } guard{new gsp_lu_plan};
  gsp_lu_plan* p = guard.p;
  p->ctx = ctx;
  p->N = base->N; p->Nd = base->Nd; p->Ns = base->Ns; p->Ndp = base->Ndp; p->Nsp = base->Nsp; p->Np = base->Np;
  p->mu = mu;
  p->h_sinds = base->h_sinds;
  p->h_dind0 = base->h_dind0;
  EventSet tev;
  cudaSetDevice(ctx->devs[0].dev);
  GSP_CUDA_OK(ctx, tev.make(2));
  for (size_t i = 0; i < base->dev.size(); ++i) {
    std::unique_ptr<LuDev> dptr(new LuDev);
    LuDev* d = dptr.get();
    const LuDev* b = base->dev[i].get();
    d->dc = b->dc;
    p->dev.push_back(std::move(dptr));
    DevCtx& dc = *d->dc;
    cudaSetDevice(dc.dev);
    cudaStream_t st = dc.stream;
    d->A = b->A;        // the factor and the block inverses are shared (reference counted)
    d->invD = b->invD;
    GSP_CUDA_OK(ctx, d->d2.alloc(dc.dev, (size_t)p->Nsp * sizeof(double), st));
    GSP_CUDA_OK(ctx, d->sinds.alloc(dc.dev, (size_t)p->Ns * sizeof(long long), st));
    GSP_CUDA_OK(ctx, d->dinds.alloc(dc.dev, (size_t)std::max<long long>(nd, 1) * sizeof(long long), st));
    GSP_CUDA_OK(ctx, d->z1.alloc(dc.dev, (size_t)std::max<long long>(nd, 1) * sizeof(double), st));
    GSP_CUDA_OK(ctx, d->info.alloc(dc.dev, sizeof(int), st));
    GSP_CUDA_OK(ctx, cudaEventCreate(&d->ev0));
    GSP_CUDA_OK(ctx, cudaEventCreate(&d->ev1));
    GSP_CUDA_OK(ctx, cudaMemcpyAsync(d->sinds.p, p->h_sinds.data(), p->h_sinds.size() * sizeof(long long), cudaMemcpyHostToDevice, st));
    if (nd > 0) {
      GSP_CUDA_OK(ctx, cudaMemcpyAsync(d->dinds.p, p->h_dind0.data(), p->h_dind0.size() * sizeof(long long), cudaMemcpyHostToDevice, st));
      GSP_CUDA_OK(ctx, cudaMemcpyAsync(d->z1.p, z1, (size_t)nd * sizeof(double), cudaMemcpyHostToDevice, st));
    }
  }
  LuDev* d0 = p->dev[0].get();
  cudaSetDevice(d0->dc->dev);
  GSP_CUDA_OK(ctx, cudaEventRecord(tev.ev[0], d0->dc->stream));
  GSP_TRY(compute_d2(ctx, p, z1));
  GSP_CUDA_OK(ctx, cudaEventRecord(tev.ev[1], d0->dc->stream));
  GSP_CUDA_OK(ctx, cudaStreamSynchronize(d0->dc->stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, tev.ev[0], tev.ev[1]);
  p->t_solve_ms = ms;  // assembly and factorization were paid by the base plan: their times stay 0 here
  for (size_t i = 1; i < p->dev.size(); ++i) {
    LuDev* e = p->dev[i].get();
    cudaSetDevice(e->dc->dev);
    GSP_CUDA_OK(ctx, cudaMemcpyPeerAsync(e->d2.p, e->dc->dev, d0->d2.p, d0->dc->dev, d0->d2.bytes, e->dc->stream));
    GSP_CUDA_OK(ctx, cudaStreamSynchronize(e->dc->stream));
  }
  for (auto& d : p->dev) {  // the host staging of sinds / dinds / z1 must be consumed before the caller's arrays go away
    cudaSetDevice(d->dc->dev);
    GSP_CUDA_OK(ctx, cudaStreamSynchronize(d->dc->stream));
  }
  guard.p = nullptr;
  *out = p;
  return GSP_OK;
}

extern "C" int gsp_lu_plan_destroy(gsp_lu_plan* p) {
  if (!p) return GSP_OK;
  destroy_plan_events(p);
  delete p;
  return GSP_OK;
}

extern "C" int gsp_lu_plan_sizes(gsp_lu_plan* p, int64_t sizes[3]) {
  if (!p || !sizes) return -1;
  sizes[0] = p->N;
  sizes[1] = p->Nd;
  sizes[2] = p->Ns;
  return GSP_OK;
}

extern "C" int gsp_lu_plan_times(gsp_lu_plan* p, double ms[3]) {
  if (!p || !ms) return -1;
  ms[0] = p->t_assemble_ms;
  ms[1] = p->t_factor_ms;
  ms[2] = p->t_solve_ms;
  return GSP_OK;
}

extern "C" int gsp_lu_plan_get(gsp_lu_plan* p, double* d2, double* L22) {
  if (!p) return -1;
  gsp_ctx* ctx = p->ctx;
  std::lock_guard<std::mutex> lk(p->mu_lock);
  LuDev* d = p->dev[0].get();
  cudaSetDevice(d->dc->dev);
  cudaStream_t st = d->dc->stream;
  if (d2) GSP_CUDA_OK(ctx, cudaMemcpyAsync(d2, d->d2.p, (size_t)p->Ns * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (L22) {
    const double* src = d->A->as<double>() + p->Ndp * (p->Np + 1);
    GSP_CUDA_OK(ctx, cudaMemcpy2DAsync(L22, (size_t)p->Ns * sizeof(double), src, (size_t)p->Np * sizeof(double), (size_t)p->Ns * sizeof(double),
                                       (size_t)p->Ns, cudaMemcpyDeviceToHost, st));
  }
  GSP_CUDA_OK(ctx, cudaStreamSynchronize(st));
  if (L22)  // strict upper triangle of off-diagonal tiles was never written: report zeros like `.L`
    for (long long j = 0; j < p->Ns; ++j)
      for (long long i = 0; i < j; ++i) L22[(size_t)j * p->Ns + i] = 0.0;
  return GSP_OK;
}

extern "C" int gsp_lu_sample_dev(gsp_lu_plan* p, int64_t R, const double* W, int64_t ldw, uint64_t seed, int32_t stream,
                                 int64_t first_real, double rho, const double* W1, double* Z, int64_t ldz) {
  if (!p) return -1;
  gsp_ctx* ctx = p->ctx;
  std::lock_guard<std::mutex> lk(p->mu_lock);
  if (R < 0) return set_err(ctx, -2, "R < 0");
  if (W && ldw < p->Ns) return set_err(ctx, -4, "ldw < Ns");
  if (!Z || ldz < p->N) return set_err(ctx, -10, "Z is NULL or ldz < N");
  const bool mix = !std::isnan(rho);
  if (mix && !(rho >= -1.0 && rho <= 1.0)) return set_err(ctx, -8, "rho must be in [-1, 1] (or NaN for the first variable)");
  if (mix && W && !W1) return set_err(ctx, -9, "W1 is required when W is given and rho is set");
  if (!W && W1) return set_err(ctx, -9, "W1 without W: both noises are injected or both come from the device RNG");
  LuDev* d = p->dev[0].get();
  cudaSetDevice(d->dc->dev);
  const long long chunk = std::min<long long>(std::max<long long>(R, 1), 1024);
  GSP_TRY(ensure_chunk(ctx, p, d, chunk, mix));
  GSP_CUDA_OK(ctx, cudaEventRecord(d->ev0, d->dc->stream));
  for (long long c0 = 0; c0 < R; c0 += chunk) {
    const long long cols = std::min(chunk, R - c0);
    GSP_TRY(sample_core(ctx, p, d, cols, W ? W + c0 * ldw : nullptr, ldw, seed, stream, first_real + c0, rho,
                        (mix && W1) ? W1 + c0 * ldw : nullptr, Z + c0 * ldz, ldz));
  }
  GSP_CUDA_OK(ctx, cudaEventRecord(d->ev1, d->dc->stream));
  GSP_CUDA_OK(ctx, cudaStreamSynchronize(d->dc->stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, d->ev0, d->ev1);
  ctx->last_sample_ms = ms;
  return GSP_OK;
}

namespace gsp {
namespace {

// host-facing sampling core: fields go to the caller's host buffer Z, or stay on the devices in `ens` (same sharding rule)
int lu_sample_impl(gsp_lu_plan* p, int64_t R, const double* W, uint64_t seed, int32_t stream, int64_t first_real, double rho,
                   const double* W1, double* Z, gsp_ensemble* ens) {
  gsp_ctx* ctx = p->ctx;
  if (R < 0) return set_err(ctx, -2, "R < 0");
  if (!Z && !ens) return set_err(ctx, -9, "Z is NULL");
  if (ens) {
    if (ens->ctx != ctx || ens->dev.size() != p->dev.size()) return set_err(ctx, -9, "the ensemble belongs to a different context");
    if (ens->n != p->N || ens->R != R) return set_err(ctx, -9, "ensemble shape does not match (n = N nodes, R realizations)");
  }
  const bool mix = !std::isnan(rho);
  if (mix && !(rho >= -1.0 && rho <= 1.0)) return set_err(ctx, -7, "rho must be in [-1, 1] (or NaN for the first variable)");
  if (mix && W && !W1) return set_err(ctx, -8, "W1 is required when W is given and rho is set");
  if (!W && W1) return set_err(ctx, -8, "W1 without W: both noises are injected or both come from the device RNG");
  const int ndev = (int)p->dev.size();
  std::vector<long long> r0(ndev + 1, 0);
  for (int i = 0; i < ndev; ++i) r0[i + 1] = r0[i] + (R / ndev) + (i < R % ndev ? 1 : 0);
  long long maxshard = 0;
  for (int i = 0; i < ndev; ++i) maxshard = std::max(maxshard, r0[i + 1] - r0[i]);
  const long long chunk = std::min<long long>(std::max<long long>(maxshard, 1), 512);
  // host-pointer calls with more than one chunk per device run a two-slot pipeline: the H2D copy of chunk c+1 (copy-in stream) and the
  // D2H copy of chunk c-1 (copy-out stream) overlap the GEMM of chunk c (compute stream).  GSP_LU_PIPELINE=0: everything in stream order.
  static int pipe_on = -1;
  if (pipe_on < 0) {
    const char* env = getenv("GSP_LU_PIPELINE");
    pipe_on = (env && env[0] == '0') ? 0 : 1;
  }
  const bool piped = pipe_on && !ens && maxshard > chunk;
#ifdef GSP_EMU
  {  // test-only: record the H2D / compute / D2H stream graph of this call and check it at the end
    const char* env = getenv("GSP_DEPCHECK");
    emu::dep_enable(env && env[0] == '1');
  }
#endif
  for (int i = 0; i < ndev; ++i) {
    cudaSetDevice(p->dev[i]->dc->dev);
    GSP_TRY(ensure_chunk(ctx, p, p->dev[i].get(), chunk, mix, piped));
  }
  {
    LuDev* d0 = p->dev[0].get();
    cudaSetDevice(d0->dc->dev);
    cudaEventRecord(d0->ev0, d0->dc->stream);
  }
  int rc = GSP_OK;
  auto cuda_ok = [&](cudaError_t e) {
    if (e != cudaSuccess && rc == GSP_OK) rc = set_err(ctx, GSP_E_CUDA, cudaGetErrorString(e));
    return e == cudaSuccess;
  };
  // chunks are issued round-robin over the devices; each device runs H2D -> compute -> D2H in stream order (or pipelined, see above).
  // The D2H copies are issued in a second sweep so that a (host-blocking) copy into pageable memory overlaps the other devices' compute.
  long long ci = 0;
  for (long long c0 = 0; c0 < maxshard && rc == GSP_OK; c0 += chunk, ++ci) {
    const int slot = piped ? (int)(ci & 1) : 0;
    for (int i = 0; i < ndev && rc == GSP_OK; ++i) {
      const long long nloc = r0[i + 1] - r0[i];
      if (c0 >= nloc) continue;
      LuDev* d = p->dev[i].get();
      cudaSetDevice(d->dc->dev);
      cudaStream_t st = d->dc->stream;
      cudaStream_t sin = piped ? d->dc->h2d : st;
      const long long cols = std::min(chunk, nloc - c0);
      const long long ra = r0[i] + c0;  // absolute first realization of this chunk
      const double* Wd = nullptr;
      const double* W1d = nullptr;
      DevBuf& wraw = slot ? d->Wraw2 : d->Wraw;
      DevBuf& w1raw = slot ? d->W1raw2 : d->W1raw;
      if (W) {
        // the slot's noise buffer is free once the compute of the chunk that used it two steps ago has run
        if (piped && ci >= 2 && !cuda_ok(cudaStreamWaitEvent(sin, d->ev_comp[slot], 0))) break;
        GSP_DEP_ACCESS(wraw.p, 0, 1, 0, 1, true);
        if (!cuda_ok(cudaMemcpyAsync(wraw.p, W + ra * p->Ns, (size_t)p->Ns * cols * sizeof(double), cudaMemcpyHostToDevice, sin))) break;
        Wd = wraw.as<double>();
        if (mix) {
          GSP_DEP_ACCESS(w1raw.p, 0, 1, 0, 1, true);
          if (!cuda_ok(cudaMemcpyAsync(w1raw.p, W1 + ra * p->Ns, (size_t)p->Ns * cols * sizeof(double), cudaMemcpyHostToDevice, sin))) break;
          W1d = w1raw.as<double>();
        }
        if (piped) {
          if (!cuda_ok(cudaEventRecord(d->ev_in[slot], sin))) break;
          if (!cuda_ok(cudaStreamWaitEvent(st, d->ev_in[slot], 0))) break;
        }
      }
      // ... and its field buffer once that chunk's copy-out has finished
      if (piped && ci >= 2 && !cuda_ok(cudaStreamWaitEvent(st, d->ev_out[slot], 0))) break;
      double* Zt = ens ? ens->dev[i]->Z.as<double>() + c0 * p->N : (slot ? d->Zc2.as<double>() : d->Zc.as<double>());
      rc = sample_core(ctx, p, d, cols, Wd, p->Ns, seed, stream, first_real + ra, rho, W1d, Zt, p->N);
      if (rc == GSP_OK && piped) cuda_ok(cudaEventRecord(d->ev_comp[slot], st));
    }
    for (int i = 0; i < ndev && rc == GSP_OK && !ens; ++i) {
      const long long nloc = r0[i + 1] - r0[i];
      if (c0 >= nloc) continue;
      LuDev* d = p->dev[i].get();
      cudaSetDevice(d->dc->dev);
      const long long cols = std::min(chunk, nloc - c0);
      const long long ra = r0[i] + c0;
      cudaStream_t sout = piped ? d->dc->d2h : d->dc->stream;
      if (piped && !cuda_ok(cudaStreamWaitEvent(sout, d->ev_comp[slot], 0))) break;
      const double* Zs = slot ? d->Zc2.as<double>() : d->Zc.as<double>();
      GSP_DEP_ACCESS(Zs, 0, 1, 0, 1, false);
      if (!cuda_ok(cudaMemcpyAsync(Z + ra * p->N, Zs, (size_t)p->N * cols * sizeof(double), cudaMemcpyDeviceToHost, sout))) break;
      if (piped) cuda_ok(cudaEventRecord(d->ev_out[slot], sout));
    }
  }
  if (piped) {
    // the compute streams (timed, and synchronised below) wait for the copy streams
    for (int i = 0; i < ndev; ++i) {
      LuDev* d = p->dev[i].get();
      cudaSetDevice(d->dc->dev);
      for (int k = 0; k < 2; ++k) {
        cudaStreamWaitEvent(d->dc->stream, d->ev_out[k], 0);   // never-recorded events are complete: no-op
        cudaStreamWaitEvent(d->dc->stream, d->ev_in[k], 0);
      }
    }
  }
  {
    LuDev* d0 = p->dev[0].get();
    cudaSetDevice(d0->dc->dev);
    cudaEventRecord(d0->ev1, d0->dc->stream);
  }
  for (int i = 0; i < ndev; ++i) {
    cudaSetDevice(p->dev[i]->dc->dev);
    cudaError_t e = cudaStreamSynchronize(p->dev[i]->dc->stream);
    if (rc == GSP_OK && e != cudaSuccess) rc = set_err(ctx, GSP_E_CUDA, std::string("lu_sample: ") + cudaGetErrorString(e));
  }
  if (rc == GSP_OK) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, p->dev[0]->ev0, p->dev[0]->ev1);
    ctx->last_sample_ms = ms;
  }
#ifdef GSP_EMU
  if (emu::dep_enabled()) {
    const long long bad = emu::dep_check(1);
    emu::dep_enable(false);
    if (bad != 0 && rc == GSP_OK)
      rc = set_err(ctx, GSP_E_STATE, "stream-dependency check failed: " + std::to_string(bad) + " unordered conflicting launch pairs");
  }
#endif
  return rc;
}

}  // namespace
}  // namespace gsp

extern "C" int gsp_lu_sample(gsp_lu_plan* p, int64_t R, const double* W, uint64_t seed, int32_t stream, int64_t first_real,
                             double rho, const double* W1, double* Z) {
  if (!p) return -1;
  std::lock_guard<std::mutex> lk(p->mu_lock);
  if (!Z) return set_err(p->ctx, -9, "Z is NULL");
  return lu_sample_impl(p, R, W, seed, stream, first_real, rho, W1, Z, nullptr);
}

extern "C" int gsp_lu_sample_ensemble(gsp_lu_plan* p, gsp_ensemble* ens, const double* W, uint64_t seed, int32_t stream,
                                      int64_t first_real, double rho, const double* W1) {
  if (!p) return -1;
  std::lock_guard<std::mutex> lk(p->mu_lock);
  if (!ens) return set_err(p->ctx, -2, "ensemble is NULL");
  std::lock_guard<std::mutex> lk2(ens->mu);
  return lu_sample_impl(p, ens->R, W, seed, stream, first_real, rho, W1, nullptr, ens);
}
