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

// Samples: coordinates sx (dim x ns, column-major) and, when zbar != nullptr, values sv.
//   zbar != nullptr : zbar[i] = mu + sum lambda_a (sv[nbr_a] - mu)            (fftsim.jl:99)
//   lam  != nullptr : slot (i, a) = lambda_a and the sample index (a < kk; unused slots: lambda 0, index 0), stored TILE-major
//                     at (i / 32) * kk * 32 + a * 32 + i % 32: the weight rows of 32 consecutive nodes are one contiguous block
// Ties in distance keep the sample with the lower index (stable insertion while scanning in index order).
// info: 1-based index of the first element whose Kriging matrix is not positive definite (0 = ok).
//
// Two phases per warp of 32 consecutive elements:
//  A. one thread per element: brute-force scan of the samples (staged in shared memory) for its kk nearest;
//  B. the WARP solves the elements' Kriging systems one after the other, lane a = neighbour a: the neighbour list is sorted by sample
//     index across the lanes (bitonic network; the order of the neighbours only permutes the system), the covariance matrix lives in
//     shared memory one row per lane (ld 33: conflict-free rows AND columns), Cholesky / forward / backward substitution are short
//     ROLLED loops over it with one broadcast per step.  Consecutive elements mostly have the SAME neighbour set (an order-26 Voronoi
//     cell of 1,000 data in 256^3 is ~7 elements across): then the factor is simply reused and only the right-hand side and the two
//     substitutions remain.
// History (256^3, 1,000 data, per Kriging): everything per thread in 4.9 KB of local memory each: 276 ms; rows in registers with fully
// unrolled shuffle recursions: 655 ms - 8k warp instructions per element of straight-line code, stalled on instruction fetch
// (ncu: no_instruction 12.6 per issue); this version: see profiles/r01_s3_notes.md.
constexpr int KW_THREADS = 64;
constexpr int KW_LD = KRIGE_MAXK + 1;
constexpr int KW_PRUNE = KRIGE_MAXK * KW_LD;  // samples whose centre distances fit the (still unused) matrix storage of a warp

__global__ void __launch_bounds__(KW_THREADS) krige_weights_kernel(CovDev cov, DomDev dom, const long long* __restrict__ inds, long long n, int kk,
                                                                   long long ns, const double* __restrict__ sx, const double* __restrict__ sv,
                                                                   double mu, double* __restrict__ zbar, double* __restrict__ lam,
                                                                   int* __restrict__ nbr, int* __restrict__ info,
                                                                   const double* __restrict__ ktab) {
  // ktab (optional): ns x ns covariances between the samples, assembled once per Kriging (8 MB for 1,000 samples: L2-resident) - a
  // matrix entry is then one load instead of a covariance evaluation
  __shared__ double tile[3 * KW_THREADS];
  __shared__ int nbrS[KW_THREADS / 32][32][KW_LD];
  __shared__ double Ls[KW_THREADS / 32][KRIGE_MAXK][KW_LD];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  double tx = 0.0, ty = 0.0, tz = 0.0;
  if (live) centroid(dom, inds ? inds[i] - 1 : i, tx, ty, tz);
  const int dim = dom.dim;
  // ---- A: the kk nearest samples of this thread's element.  The list is kept UNSORTED (phase B orders it by sample index anyway):
  // a candidate replaces the current worst entry - largest distance, ties: highest sample index, which is what a stable sorted
  // insertion would push out - and the new worst is found by one pass over the list (independent local loads; the sorted insertion
  // of the first version was a chain of dependent local-memory loads and took most of the kernel's time).  The worst distance
  // lives in a register, so a sample that does not enter costs no local-memory access.
  // The warp first PRUNES the samples together (ns <= KW_PRUNE): with c, rho the centre and half-diagonal of the box around its 32
  // elements and U >= the kk-th smallest distance from c (bisection on the cached distances), every element's kk nearest samples lie
  // within U + 2 rho of c (triangle inequality), typically ~120 of 1,000.  The candidates keep their index order (ballot compaction),
  // so the tie rule is unchanged.
  {
    double bd[KRIGE_MAXK];
    int bi[KRIGE_MAXK];
    int cntA = 0, wpos = 0;
    double worst = -1.0;
    auto offer = [&](double d2, int sidx) {
      if (cntA < kk) {
        bd[cntA] = d2;
        bi[cntA] = sidx;
        if (d2 >= worst) {  // later index wins the tie for "worst"
          worst = d2;
          wpos = cntA;
        }
        ++cntA;
        return;
      }
      if (!(d2 < worst)) return;
      bd[wpos] = d2;
      bi[wpos] = sidx;
      worst = -1.0;
      int wi = -1;
      for (int a = 0; a < kk; ++a) {
        const double da = bd[a];
        const int ia = bi[a];
        if (da > worst || (da == worst && ia > wi)) {
          worst = da;
          wi = ia;
          wpos = a;
        }
      }
    };
    if (ns <= KW_PRUNE) {
      double* dc = &Ls[warp][0][0];          // phase B's matrix is not in use yet: squared distances from the box centre
      int* cand = &nbrS[warp][0][0];         // ... and its neighbour table: the candidate list
      // box of the warp's elements (lanes past the end copy lane 0)
      const unsigned livemask = __ballot_sync(0xffffffffu, live);
      int ncand = 0;
      if (livemask != 0u) {
        const int src = __ffs((int)livemask) - 1;
        const double fx = __shfl_sync(0xffffffffu, tx, src), fy = __shfl_sync(0xffffffffu, ty, src), fz = __shfl_sync(0xffffffffu, tz, src);
        double lx = live ? tx : fx, ly = live ? ty : fy, lz = live ? tz : fz;
        double hx2 = lx, hy2 = ly, hz2 = lz;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          lx = fmin(lx, __shfl_xor_sync(0xffffffffu, lx, o)); hx2 = fmax(hx2, __shfl_xor_sync(0xffffffffu, hx2, o));
          ly = fmin(ly, __shfl_xor_sync(0xffffffffu, ly, o)); hy2 = fmax(hy2, __shfl_xor_sync(0xffffffffu, hy2, o));
          lz = fmin(lz, __shfl_xor_sync(0xffffffffu, lz, o)); hz2 = fmax(hz2, __shfl_xor_sync(0xffffffffu, hz2, o));
        }
        const double cx = 0.5 * (lx + hx2), cy = 0.5 * (ly + hy2), cz = 0.5 * (lz + hz2);
        const double rho = 0.5 * sqrt((hx2 - lx) * (hx2 - lx) + (hy2 - ly) * (hy2 - ly) + (hz2 - lz) * (hz2 - lz));
        double dmax = 0.0;
        for (int sidx = lane; sidx < (int)ns; sidx += 32) {
          const double* p = sx + (long long)sidx * dim;
          const double dx = p[0] - cx, dy = (dim > 1 ? p[1] : 0.0) - cy, dz = (dim > 2 ? p[2] : 0.0) - cz;
          const double d2 = dx * dx + dy * dy + dz * dz;
          dc[sidx] = d2;
          dmax = fmax(dmax, d2);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
        __syncwarp();
        // U^2: smallest bisection point with at least kk samples inside (any upper bound of the kk-th distance is valid)
        double lo = 0.0, hi = dmax;
        const int need = (int)(ns < kk ? ns : kk);
        for (int itb = 0; itb < 14; ++itb) {
          const double mid = 0.5 * (lo + hi);
          int c = 0;
          for (int sidx = lane; sidx < (int)ns; sidx += 32) c += dc[sidx] <= mid ? 1 : 0;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
          if (c >= need) hi = mid;
          else lo = mid;
        }
        const double rr = sqrt(hi) + 2.0 * rho;
        const double T = rr * rr * (1.0 + 1e-9);  // slack for the rounding of the three distances in the triangle inequality
        // candidates in index order
        for (int s0 = 0; s0 < (int)ns; s0 += 32) {
          const int sidx = s0 + lane;
          const bool in = sidx < (int)ns && dc[sidx] <= T;
          const unsigned mk = __ballot_sync(0xffffffffu, in);
          __syncwarp();
          if (in) cand[ncand + __popc(mk & ((1u << lane) - 1u))] = sidx;
          ncand += __popc(mk);
        }
        __syncwarp();
        if (live) {
          for (int t = 0; t < ncand; ++t) {
            const int sidx = cand[t];
            const double* p = sx + (long long)sidx * dim;
            const double dx = p[0] - tx, dy = (dim > 1 ? p[1] : 0.0) - ty, dz = (dim > 2 ? p[2] : 0.0) - tz;
            offer(dx * dx + dy * dy + dz * dz, sidx);
          }
        }
        __syncwarp();   // the candidate list is dead: its storage becomes the neighbour table
      }
    } else {
      for (long long s0 = 0; s0 < ns; s0 += KW_THREADS) {
        const int m = (int)((ns - s0 < KW_THREADS) ? ns - s0 : KW_THREADS);
        __syncthreads();
        if ((int)threadIdx.x < m) {
          const double* p = sx + (s0 + threadIdx.x) * dim;
          tile[threadIdx.x] = p[0];
          tile[KW_THREADS + threadIdx.x] = dim > 1 ? p[1] : 0.0;
          tile[2 * KW_THREADS + threadIdx.x] = dim > 2 ? p[2] : 0.0;
        }
        __syncthreads();
        if (!live) continue;
        for (int j = 0; j < m; ++j) {
          const double dx = tile[j] - tx, dy = tile[KW_THREADS + j] - ty, dz = tile[2 * KW_THREADS + j] - tz;
          offer(dx * dx + dy * dy + dz * dz, (int)(s0 + j));
        }
      }
    }
    for (int a = 0; a < KRIGE_MAXK; ++a) nbrS[warp][lane][a] = (live && a < cntA) ? bi[a] : 0x7fffffff;
  }
  __syncwarp();
  // ---- B: the warp's 32 elements one after the other, lane = neighbour
  const int cnt = (int)(ns < kk ? ns : kk);      // every element finds the same number of neighbours
  const long long wbase = i - lane;              // first element of this warp
  double (*L)[KW_LD] = Ls[warp];
  double rd_own = 1.0;                           // 1 / L[lane][lane]
  double ax = 0.0, ay = 0.0, az = 0.0;
  int prev = -1;
  bool have = false;
  const bool act = lane < cnt;
  for (int nn = 0; nn < 32; ++nn) {
    const long long inode = wbase + nn;
    if (inode >= n) break;                       // uniform: elements of a warp are consecutive
    // neighbour list sorted by sample index (sentinels last)
    int my = nbrS[warp][nn][lane];
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
      for (int j = k >> 1; j > 0; j >>= 1) {
        const int other = __shfl_xor_sync(0xffffffffu, my, j);
        const bool up = (lane & k) == 0, lower = (lane & j) == 0;
        my = (lower == up) ? (my < other ? my : other) : (my > other ? my : other);
      }
    }
    const bool same = have && (__ballot_sync(0xffffffffu, my != prev) == 0u);
    if (!same) {
      if (act) {
        const double* pa = sx + (long long)my * dim;
        ax = pa[0];
        ay = dim > 1 ? pa[1] : 0.0;
        az = dim > 2 ? pa[2] : 0.0;
      }
      // row `lane` of the covariance matrix of the neighbours (lower part)
      if (ktab) {
        const double* krow = ktab + (long long)(act ? my : 0) * ns;
        for (int b = 0; b < cnt; ++b) {
          const int mb = __shfl_sync(0xffffffffu, my, b);
          if (act && b <= lane) L[lane][b] = krow[mb];
        }
      } else {
        for (int b = 0; b < cnt; ++b) {
          const double bx = __shfl_sync(0xffffffffu, ax, b), by = __shfl_sync(0xffffffffu, ay, b), bz = __shfl_sync(0xffffffffu, az, b);
          if (act && b <= lane) L[lane][b] = cov_eval(cov, ax - bx, ay - by, az - bz);
        }
      }
      __syncwarp();
      // left-looking Cholesky, one column per step: lane r forms L[r][J] from its own row and row J (a broadcast) - two loads and
      // one FMA per term, the sum in registers (the right-looking form read-modify-wrote the whole trailing block in shared memory)
      bool bad = false;
      for (int J = 0; J < cnt; ++J) {
        const double* rl = L[lane];
        const double* rj = L[J];
        double s0 = 0.0, s1 = 0.0;
        int k = 0;
        for (; k + 1 < J; k += 2) {
          s0 += rl[k] * rj[k];
          s1 += rl[k + 1] * rj[k + 1];
        }
        if (k < J) s0 += rl[k] * rj[k];
        const double v = rl[J] - (s0 + s1);
        const double piv = __shfl_sync(0xffffffffu, v, J);
        if (!(piv > 0.0)) bad = true;
        const double rinv = rsqrt(piv);
        if (lane == J) {
          L[J][J] = piv * rinv;
          rd_own = rinv;
        } else if (lane > J && act) {
          L[lane][J] = v * rinv;
        }
        __syncwarp();
      }
      if (bad && lane == 0) atomicCAS(info, 0, (int)(inode < 2147483646LL ? inode + 1 : 2147483647LL));
      prev = my;
      have = true;
    }
    // right-hand side: covariance between neighbour `lane` and the element, then L y = c0 and L' lambda = y, both column-oriented:
    // one broadcast of the finished component per step, every other lane updates its own
    const double ex = __shfl_sync(0xffffffffu, tx, nn), ey = __shfl_sync(0xffffffffu, ty, nn), ez = __shfl_sync(0xffffffffu, tz, nn);
    double x = act ? cov_eval(cov, ax - ex, ay - ey, az - ez) : 0.0;
    for (int J = 0; J < cnt; ++J) {
      const double yj = __shfl_sync(0xffffffffu, x * rd_own, J);
      if (lane == J) x = yj;
      else if (lane > J && act) x -= L[lane][J] * yj;
    }
    for (int J = cnt - 1; J >= 0; --J) {
      const double lj = __shfl_sync(0xffffffffu, x * rd_own, J);
      if (lane == J) x = lj;
      else if (lane < J) x -= L[J][lane] * lj;
    }
    if (zbar) {
      double t = act ? x * (sv[my] - mu) : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0) zbar[inode] = mu + t;
    }
    if (lam && lane < kk) {
      const long long base = (inode >> 5) * kk * 32 + (inode & 31);
      lam[base + lane * 32] = act ? x : 0.0;
      nbr[base + lane * 32] = act ? my : 0;
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
      // the weight table lists every element's neighbours in sample order (krige_weights_kernel sorts them), so adjacent elements
      // mostly name the same sample at the same rank: one gather then serves several of the four (uniform branches: o is a broadcast)
      const double r0 = resl[o.x];
      const double r1 = (o.y == o.x) ? r0 : resl[o.y];
      const double r2 = (o.z == o.y) ? r1 : resl[o.z];
      const double r3 = (o.w == o.z) ? r2 : resl[o.w];
      acc0 += l01.x * r0;
      acc1 += l01.y * r1;
      acc2 += l23.x * r2;
      acc3 += l23.y * r3;
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
