// Conditioning of FFTSIM realizations by simple Kriging of residuals:
//   preprocess  fftsim.jl:94-101   zbar  = fitpredict(Kriging(f, mu), data, sdom; minneighbors, maxneighbors, distance)
//   randsingle  fftsim.jl:140-153  zbaru = fitpredict(Kriging(f, mu), georef(zu[dinds], view(sdom, dinds)), sdom; ...)
//                                  z     = zbar + (zu - zbaru)
// GeoStatsModels' neighbourhood path (dependency, not vendored; restated from its published algorithm): for every
// element of sdom take the `maxneighbors` nearest samples (KNearestSearch, Euclidean, sorted by distance), fit simple
// Kriging to them - covariance matrix C of the samples, factorised with Cholesky - and predict at the element's centroid:
// zhat = mu + sum_a lambda_a (z_a - mu), lambda = C^-1 c0.  The weights depend on geometry only, so the second Kriging,
// which the reference re-fits for every realization, is reduced here to a per-node weight table built ONCE per plan.
#pragma once
#include "cov.cuh"

namespace gsp {

constexpr int KRIGE_MAXK = 32;  // neighbours per node (the reference's default maxneighbors is 26)

// One thread per element of sdom.  Samples: coordinates sx (dim x ns, column-major) and, when zbar != nullptr, values sv.
//   zbar != nullptr : zbar[i] = mu + sum lambda_a (sv[nbr_a] - mu)            (fftsim.jl:99)
//   lam  != nullptr : slot (i, a) = lambda_a and the sample index (a < kk; unused slots: lambda 0, index 0), stored TILE-major
//                     at (i / 32) * kk * 32 + a * 32 + i % 32: the weight rows of 32 consecutive nodes are one contiguous block
// Ties in distance keep the sample with the lower index (stable insertion while scanning in index order).
// info: 1-based index of the first element whose Kriging matrix is not positive definite (0 = ok).
__global__ void __launch_bounds__(128) krige_weights_kernel(CovDev cov, DomDev dom, const long long* __restrict__ inds, long long n, int kk,
                                                            long long ns, const double* __restrict__ sx, const double* __restrict__ sv, double mu,
                                                            double* __restrict__ zbar, double* __restrict__ lam, int* __restrict__ nbr,
                                                            int* __restrict__ info) {
  __shared__ double tile[3 * 128];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  double tx = 0.0, ty = 0.0, tz = 0.0;
  if (live) centroid(dom, inds ? inds[i] - 1 : i, tx, ty, tz);
  double bd[KRIGE_MAXK];
  int bi[KRIGE_MAXK];
  int cnt = 0;
  const int dim = dom.dim;
  for (long long s0 = 0; s0 < ns; s0 += 128) {
    const int m = (int)((ns - s0 < 128) ? ns - s0 : 128);
    __syncthreads();
    if ((int)threadIdx.x < m) {
      const double* p = sx + (s0 + threadIdx.x) * dim;
      tile[threadIdx.x] = p[0];
      tile[128 + threadIdx.x] = dim > 1 ? p[1] : 0.0;
      tile[256 + threadIdx.x] = dim > 2 ? p[2] : 0.0;
    }
    __syncthreads();
    if (!live) continue;
    for (int j = 0; j < m; ++j) {
      const double dx = tile[j] - tx, dy = tile[128 + j] - ty, dz = tile[256 + j] - tz;
      const double d2 = dx * dx + dy * dy + dz * dz;
      if (cnt == kk && !(d2 < bd[kk - 1])) continue;
      int pos = cnt < kk ? cnt : kk - 1;  // slot that is freed (or appended)
      while (pos > 0 && d2 < bd[pos - 1]) {
        bd[pos] = bd[pos - 1];
        bi[pos] = bi[pos - 1];
        --pos;
      }
      bd[pos] = d2;
      bi[pos] = (int)(s0 + j);
      if (cnt < kk) ++cnt;
    }
  }
  if (!live) return;
  // simple-Kriging system of the cnt neighbours: packed lower C, right-hand side c0 = cov(sample, target)
  double C[KRIGE_MAXK * (KRIGE_MAXK + 1) / 2];
  double x[KRIGE_MAXK];
  for (int a = 0; a < cnt; ++a) {
    const double* pa = sx + (long long)bi[a] * dim;
    const double ax = pa[0], ay = dim > 1 ? pa[1] : 0.0, az = dim > 2 ? pa[2] : 0.0;
    for (int b = 0; b <= a; ++b) {
      const double* pb = sx + (long long)bi[b] * dim;
      C[a * (a + 1) / 2 + b] = cov_eval(cov, ax - pb[0], dim > 1 ? ay - pb[1] : 0.0, dim > 2 ? az - pb[2] : 0.0);
    }
    x[a] = cov_eval(cov, ax - tx, ay - ty, az - tz);
  }
  // Cholesky (row by row), then L y = c0 and L' lambda = y
  bool bad = false;
  for (int a = 0; a < cnt; ++a) {
    for (int b = 0; b <= a; ++b) {
      double s = C[a * (a + 1) / 2 + b];
      for (int k = 0; k < b; ++k) s -= C[a * (a + 1) / 2 + k] * C[b * (b + 1) / 2 + k];
      if (b < a) {
        C[a * (a + 1) / 2 + b] = s / C[b * (b + 1) / 2 + b];
      } else {
        if (!(s > 0.0)) bad = true;
        C[a * (a + 1) / 2 + a] = sqrt(s);
      }
    }
  }
  if (bad) {
    atomicCAS(info, 0, (int)(i < 2147483646LL ? i + 1 : 2147483647LL));
    return;
  }
  for (int a = 0; a < cnt; ++a) {
    double s = x[a];
    for (int k = 0; k < a; ++k) s -= C[a * (a + 1) / 2 + k] * x[k];
    x[a] = s / C[a * (a + 1) / 2 + a];
  }
  for (int a = cnt - 1; a >= 0; --a) {
    double s = x[a];
    for (int k = a + 1; k < cnt; ++k) s -= C[k * (k + 1) / 2 + a] * x[k];
    x[a] = s / C[a * (a + 1) / 2 + a];
  }
  if (zbar) {
    double acc = 0.0;
    for (int a = 0; a < cnt; ++a) acc += x[a] * (sv[bi[a]] - mu);
    zbar[i] = mu + acc;
  }
  if (lam) {
    const long long base = (i >> 5) * kk * 32 + (i & 31);
    for (int a = 0; a < kk; ++a) {
      lam[base + a * 32] = a < cnt ? x[a] : 0.0;
      nbr[base + a * 32] = a < cnt ? bi[a] : 0;
    }
  }
}

constexpr int KRIGE_RB = 32;  // realizations per conditioning chunk = lanes of a warp

// res[j * 32 + r] = zu[knode_j, r] - mu (r < nb, zero beyond): the table georef(zu[dinds], view(sdom, dinds)) of fftsim.jl:142-144,
// realization index fastest so that the 32 lanes of a warp (one realization each) read one sample's residuals as one 256-byte run
__global__ void __launch_bounds__(256) krige_residual_kernel(const double* __restrict__ Z, long long n, const long long* __restrict__ knodes,
                                                             long long nk, long long nb, double mu, double* __restrict__ res) {
  const long long total = nk * KRIGE_RB;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const long long j = t / KRIGE_RB, r = t - j * KRIGE_RB;
    res[t] = r < nb ? Z[knodes[j] + r * n] - mu : 0.0;
  }
}

// z = zbar + (zu - zbaru), zbaru = mu + sum_a lambda_a res_a (fftsim.jl:148-152), in place for nb <= 32 realizations.
// A CTA walks over tiles of 32 nodes x 32 realizations: the tile of Z and the nodes' weight rows are staged in shared memory with
// coalesced 256-byte runs; in the compute phase a LANE IS A REALIZATION, so the weight / neighbour index of a node is a
// shared-memory broadcast and the residual gather res[nbr * 32 + lane] is one contiguous run per neighbour (a thread-per-node
// version spent its time replaying 32-way divergent gathers: 504 us per realization at 256^3 against ~90 us of traffic).
__global__ void __launch_bounds__(256) krige_apply_kernel(double* __restrict__ Z, long long n, int nb, const double* __restrict__ zbar,
                                                          const double* __restrict__ lam, const int* __restrict__ nbr, int kk,
                                                          const double* __restrict__ res, double mu) {
  alignas(16) __shared__ double lamS[KRIGE_MAXK][36];
  alignas(16) __shared__ int nbrS[KRIGE_MAXK][36];
  __shared__ double Zs[KRIGE_RB][33];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long ntiles = (n + 31) / 32;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long i0 = tile * 32;
    // the tile's weight rows: one contiguous block of kk * 32 entries (tile-major layout written by krige_weights_kernel);
    // slots of nodes >= n in the last tile were never written: masked here
    for (int idx = tid; idx < kk * 32; idx += 256) {
      const int a = idx >> 5, c = idx & 31;
      const bool in = i0 + c < n;
      lamS[a][c] = in ? lam[tile * kk * 32 + idx] : 0.0;
      nbrS[a][c] = in ? nbr[tile * kk * 32 + idx] * KRIGE_RB : 0;
    }
    for (int idx = tid; idx < nb * 32; idx += 256) {
      const int r = idx >> 5, c = idx & 31;
      Zs[r][c] = (i0 + c < n) ? Z[i0 + c + (long long)r * n] : 0.0;
    }
    __syncthreads();
    // warp w owns the four ADJACENT nodes c = 4w .. 4w + 3 of the tile: their weights / neighbour offsets are fetched with
    // 16-byte broadcast loads (rows are padded to 36 entries: 16-byte aligned), four independent accumulation chains per lane
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    const int c0 = 4 * warp;
    const double* resl = res + lane;
    for (int a = 0; a < kk; ++a) {
      const double2 l01 = *reinterpret_cast<const double2*>(&lamS[a][c0]);
      const double2 l23 = *reinterpret_cast<const double2*>(&lamS[a][c0 + 2]);
      const int4 o = *reinterpret_cast<const int4*>(&nbrS[a][c0]);
      acc0 += l01.x * resl[o.x];
      acc1 += l01.y * resl[o.y];
      acc2 += l23.x * resl[o.z];
      acc3 += l23.y * resl[o.w];
    }
    if (lane < nb) {
      if (i0 + c0 < n) Zs[lane][c0] = zbar[i0 + c0] + (Zs[lane][c0] - (mu + acc0));
      if (i0 + c0 + 1 < n) Zs[lane][c0 + 1] = zbar[i0 + c0 + 1] + (Zs[lane][c0 + 1] - (mu + acc1));
      if (i0 + c0 + 2 < n) Zs[lane][c0 + 2] = zbar[i0 + c0 + 2] + (Zs[lane][c0 + 2] - (mu + acc2));
      if (i0 + c0 + 3 < n) Zs[lane][c0 + 3] = zbar[i0 + c0 + 3] + (Zs[lane][c0 + 3] - (mu + acc3));
    }
    __syncthreads();
    for (int idx = tid; idx < nb * 32; idx += 256) {
      const int r = idx >> 5, c = idx & 31;
      if (i0 + c < n) Z[i0 + c + (long long)r * n] = Zs[r][c];
    }
    __syncthreads();
  }
}

}  // namespace gsp
