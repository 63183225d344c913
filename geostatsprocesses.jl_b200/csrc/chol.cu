// Blocked Cholesky on the device: replaces cholesky(Symmetric(.)).L and the
// triangular solves of lusim.jl:92,98-103.  Recursive right-looking formulation over 128x128
// tiles: every flop outside the 128-wide diagonal blocks runs in the DMMA tile GEMM (gemm.cuh);
// diagonal blocks are factorised AND inverted by one CTA (potrf_diag_kernel) so that every
// triangular solve becomes a tile GEMM with the explicit inverse.
#include "chol.h"

#include <cstdlib>
#include <vector>

#include "gemm.cuh"

namespace gsp {

constexpr int DB = 128;   // diagonal block
constexpr int DLD = 129;  // padded smem leading dimension (row-major S[r][c])
constexpr int DPW = 16;   // inner panel width
constexpr int DIAG_SMEM = (DB * DLD + DPW * (DPW + 1) + DPW) * 8;

// columns J.. of one 128x16 panel, one matrix row per thread (x = that row's 16 panel entries, in registers).
// Left-looking: the pivot thread finalises its diagonal entry and publishes its row, everyone below
// finishes its entry of column J.  One barrier per column.
template <int J>
GSP_DEV void diag_panel_cols(double (&x)[DPW], int i, int c0, bool active, double* P, double* rinv, int* info, int gbase) {
  if constexpr (J < DPW) {
    if (active && i == c0 + J) {
      double d = x[J];
#pragma unroll
      for (int k = 0; k < J; ++k) d -= x[k] * x[k];
      if (!(d > 0.0)) atomicCAS(info, 0, gbase + c0 + J + 1);
      const double s = sqrt(d);
      x[J] = s;
#pragma unroll
      for (int k = 0; k <= J; ++k) P[J * (DPW + 1) + k] = x[k];
      rinv[J] = 1.0 / s;
    }
    __syncthreads();
    if (active && i > c0 + J) {
      double v = x[J];
#pragma unroll
      for (int k = 0; k < J; ++k) v -= x[k] * P[J * (DPW + 1) + k];
      x[J] = v * rinv[J];
    }
    diag_panel_cols<J + 1>(x, i, c0, active, P, rinv, info, gbase);
  }
}

// X21 = -inv(C) * B * inv(A) for the pair of diagonal sub-blocks A = S[o:o+h, o:o+h], C = S[o+h:o+2h, o+h:o+2h]
// (both already replaced by their inverses) and B = S[o+h:o+2h, o:o+h]; in place.  All 256 threads call this
// with the same h; `pairs` pairs are processed at once (pairs * h == 64).
template <int H>
GSP_DEV void diag_inverse_level(double* S, int tid) {
  constexpr int PAIRS = DB / (2 * H);
  constexpr int TPP = 256 / PAIRS;          // threads per pair
  constexpr int CG = (H * H) / TPP;          // outputs per thread (a run of CG columns of one row)
  constexpr int CGS = CG < H ? CG : H;       // columns per thread in a row
  static_assert(CG >= 1 && CG <= H, "tiling");
  const int pr = tid / TPP, lt = tid - pr * TPP;
  const int o = pr * 2 * H;
  const int i = lt % H;                      // row within B
  const int j0 = (lt / H) * CGS;             // first output column
  double* Bm = S + (o + H) * DLD + o;        // B[i][k]   = Bm[i*DLD + k]
  const double* Am = S + o * DLD + o;        // invA[k][j]
  const double* Cm = S + (o + H) * DLD + o + H;  // invC[i][k]
  double acc[CGS];
  // T = B * invA  (invA lower triangular: k >= j)
#pragma unroll
  for (int c = 0; c < CGS; ++c) acc[c] = 0.0;
  for (int k = j0; k < H; ++k) {
    const double b = Bm[i * DLD + k];
#pragma unroll
    for (int c = 0; c < CGS; ++c)
      if (k >= j0 + c) acc[c] += b * Am[k * DLD + j0 + c];
  }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < CGS; ++c) Bm[i * DLD + j0 + c] = acc[c];
  __syncthreads();
  // X21 = -invC * T  (invC lower triangular: k <= i)
#pragma unroll
  for (int c = 0; c < CGS; ++c) acc[c] = 0.0;
  for (int k = 0; k <= i; ++k) {
    const double cv = Cm[i * DLD + k];
#pragma unroll
    for (int c = 0; c < CGS; ++c) acc[c] += cv * Bm[k * DLD + j0 + c];
  }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < CGS; ++c) Bm[i * DLD + j0 + c] = -acc[c];
  __syncthreads();
}

// Factor A[blk,blk] (128x128, lower) in place -> L (strict upper zeroed), write inv(L) to invD.
__global__ void __launch_bounds__(256, 1) potrf_diag_kernel(double* __restrict__ A, long long lda, long long blk,
                                                            double* __restrict__ invD, int* __restrict__ info) {
  GSP_DYN_SMEM(smem);
  double* S = reinterpret_cast<double*>(smem);   // [DB][DLD]
  double* P = S + DB * DLD;                      // [DPW][DPW+1] pivot rows of the current panel
  double* rinv = P + DPW * (DPW + 1);            // [DPW]
  const int tid = threadIdx.x;
  double* Ab = A + blk * DB * (lda + 1);

  for (int idx = tid; idx < DB * DB; idx += 256) {
    const int r = idx & (DB - 1), c = idx >> 7;
    S[r * DLD + c] = (r >= c) ? Ab[r + (long long)c * lda] : 0.0;
  }
  __syncthreads();

  for (int p = 0; p < DB / DPW; ++p) {
    const int c0 = p * DPW;
    // (a) panel factorisation, one row per thread
    {
      const int i = tid;
      const bool active = tid < DB && i >= c0;
      double x[DPW];
#pragma unroll
      for (int k = 0; k < DPW; ++k) x[k] = active ? S[i * DLD + c0 + k] : 0.0;
      diag_panel_cols<0>(x, i, c0, active, P, rinv, info, (int)(blk * DB));
      if (active) {
#pragma unroll
        for (int k = 0; k < DPW; ++k) S[i * DLD + c0 + k] = (c0 + k <= i) ? x[k] : 0.0;
      }
    }
    __syncthreads();
    // (b) trailing update of the lower triangle right of the panel (rank-16)
    {
      const int ti = tid & 15, tk = tid >> 4;
      for (int a = p + 1; a < DB / DPW; ++a) {
        const int i = ti + DPW * a;
        double ra[DPW];
#pragma unroll
        for (int c = 0; c < DPW; ++c) ra[c] = S[i * DLD + c0 + c];
        for (int b = p + 1; b <= a; ++b) {
          const int k = tk + DPW * b;
          if (i < k) continue;
          const double* sk = S + k * DLD + c0;
          double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
          for (int c = 0; c < DPW; c += 2) {
            acc0 += ra[c] * sk[c];
            acc1 += ra[c + 1] * sk[c + 1];
          }
          S[i * DLD + k] -= acc0 + acc1;
        }
      }
    }
    __syncthreads();
  }

  // write L (upper triangle explicitly zero)
  for (int idx = tid; idx < DB * DB; idx += 256) {
    const int r = idx & (DB - 1), c = idx >> 7;
    Ab[r + (long long)c * lda] = (r >= c) ? S[r * DLD + c] : 0.0;
  }
  __syncthreads();

  // inverse of the lower-triangular block, blocked: 16x16 diagonal blocks first (one per warp, one column per lane) ...
  {
    const int w = tid >> 5, l = tid & 31;
    const int o = w * DPW;
    double x[DPW];
    if (l < DPW) {
#pragma unroll
      for (int i = 0; i < DPW; ++i) {
        double sacc = (i == l) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < i; ++k) sacc -= S[(o + i) * DLD + o + k] * x[k];
        x[i] = sacc / S[(o + i) * DLD + o + i];
      }
    }
    __syncthreads();
    if (l < DPW) {
#pragma unroll
      for (int i = 0; i < DPW; ++i) S[(o + i) * DLD + o + l] = x[i];  // column l of the inverse (zero above the diagonal)
    }
    __syncthreads();
  }
  // ... then X21 = -inv(C) B inv(A) level by level (16 -> 32 -> 64 -> 128)
  diag_inverse_level<16>(S, tid);
  diag_inverse_level<32>(S, tid);
  diag_inverse_level<64>(S, tid);

  double* Xo = invD + blk * DB * DB;
  for (int idx = tid; idx < DB * DB; idx += 256) {
    const int r = idx & (DB - 1), c = idx >> 7;
    Xo[r + c * DB] = (r >= c) ? S[r * DLD + c] : 0.0;
  }
}

// y = L11^{-1} z for the leading nb x nb blocks (forward substitution by 128-blocks using invD):
// y_b = invD_b * (z_b - sum_{j<b} L[b][j] y_j).  One CTA; z is overwritten by y.
__global__ void __launch_bounds__(256, 1) trsv_blocks_kernel(const double* __restrict__ L, long long ld,
                                                             const double* __restrict__ invD, int nblocks,
                                                             double* __restrict__ z) {
  __shared__ double part[256];
  __shared__ double rb[DB];
  const int tid = threadIdx.x, r = tid & (DB - 1), half = tid >> 7;
  for (int b = 0; b < nblocks; ++b) {
    // residual r_b = z_b - L[b, 0:b*128] y
    const long long row = (long long)b * DB + r;
    const int kn = b * DB;
    double s = 0.0;
    for (int k = half; k < kn; k += 2) s += L[row + (long long)k * ld] * z[k];
    part[tid] = s;
    __syncthreads();
    if (tid < DB) rb[tid] = z[row] - (part[tid] + part[tid + DB]);
    __syncthreads();
    const double* X = invD + (long long)b * DB * DB;
    double t = 0.0;
    for (int k = half; k <= r; k += 2) t += X[r + k * DB] * rb[k];
    part[tid] = t;
    __syncthreads();
    if (tid < DB) z[row] = part[tid] + part[tid + DB];
    __syncthreads();
  }
}

// d2[i] = sum_{k<kn} L[row0 + i][k] * y[k]   (lusim.jl:102: A21 * (L11 \ z1))
__global__ void __launch_bounds__(256) gemv_rows_kernel(const double* __restrict__ L, long long ld, long long row0, long long nrows,
                                                        int kn, const double* __restrict__ y, double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  const double* p = L + row0 + i;
  double s0 = 0.0, s1 = 0.0;
  int k = 0;
  for (; k + 1 < kn; k += 2) {
    s0 += p[(long long)k * ld] * y[k];
    s1 += p[(long long)(k + 1) * ld] * y[k + 1];
  }
  if (k < kn) s0 += p[(long long)k * ld] * y[k];
  out[i] = s0 + s1;
}

// ------------------------------------------------------------------ host orchestration
namespace {

struct Chol {
  cudaStream_t st;       // main (high priority): diagonal blocks, panels, leading part of every update
  cudaStream_t* side;    // look-ahead streams, one per recursion depth
  int nside;
  double* A;
  long long ld;
  double* invD;
  int* info;
  cudaError_t err = cudaSuccess;
  std::vector<cudaEvent_t> events;

  double* at(int br, int bc) const { return A + (long long)br * DB + (long long)bc * DB * ld; }

  void check(cudaError_t e) {
    if (err == cudaSuccess && e != cudaSuccess) err = e;
  }
  cudaEvent_t record(cudaStream_t s) {
    cudaEvent_t ev;
    check(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    check(cudaEventRecord(ev, s));
    events.push_back(ev);
    return ev;
  }

  // C[mt x nt blocks at (cr, cc)] -= A[(ar, ac), K blocks] * B[(br, bc), K blocks]^T   (lower tiles only if tri)
  // On a look-ahead stream the K range is cut into slices so that no CTA holds an SM for long:
  // pending high-priority CTAs of the main stream then get an SM within one slice.
  void update(cudaStream_t s, bool sliced, int cr, int cc, int mt, int nt, int ar, int ac, int br, int bc, int kb, bool tri) {
    if (mt <= 0 || nt <= 0 || kb <= 0) return;
    static int slice = -1;  // GSP_CHOL_SLICE: K blocks per launch on look-ahead streams
    if (slice < 0) {
      const char* env = getenv("GSP_CHOL_SLICE");
      slice = env ? atoi(env) : 0;  // default: no slicing (measured best on B200: 70.4 ms vs 73.1 ms at C3)
      if (slice < 1) slice = 1 << 20;
    }
    const int step = sliced ? slice : kb;
    for (int k0 = 0; k0 < kb; k0 += step) {
      const int kk = (kb - k0 < step) ? kb - k0 : step;
      GemmArgs g{};
      g.A = at(ar, ac + k0); g.lda = ld;
      g.B = at(br, bc + k0); g.ldb = ld;
      g.C = at(cr, cc); g.ldc = ld;
      g.mt = mt; g.nt = nt; g.K = kk * DB; g.tri = tri ? 1 : 0;
      check(launch_gemm<GEMM_SUB, false>(s, g));
    }
  }

  // X * L[c0:c0+nc, c0:c0+nc]^T = A[r0:r0+nr, c0:c0+nc]   (in place)
  void trsm(int r0, int nr, int c0, int nc) {
    if (nr <= 0 || nc <= 0) return;
    if (nc == 1) {
      GemmArgs g{};
      g.A = at(r0, c0); g.lda = ld;
      g.B = invD + (long long)c0 * DB * DB; g.ldb = DB;
      g.C = at(r0, c0); g.ldc = ld;
      g.mt = nr; g.nt = 1; g.K = DB;
      check(launch_gemm<GEMM_SET, false>(st, g));
      return;
    }
    const int c1 = nc / 2;
    trsm(r0, nr, c0, c1);
    update(st, false, r0, c0 + c1, nr, nc - c1, r0, c0, c0 + c1, c0, c1, false);
    trsm(r0, nr, c0 + c1, nc - c1);
  }

  // Cholesky of the n diagonal blocks starting at o.  `pend`: event after which the second half [o + n/2, o + n) of the
  // block rows/cols is up to date (nullptr: already valid on the main stream).  The first half is always valid on entry.
  void potrf(int o, int n, cudaEvent_t pend, int depth) {
    if (n <= 0) return;
    if (n == 1) {
      auto kfn = potrf_diag_kernel;
      check(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM));
      ProfScope prof_("potrf_diag", st);
      GSP_LAUNCH(kfn, dim3(1), dim3(256), (size_t)DIAG_SMEM, st, A, ld, (long long)o, invD, info);
      g_launches++;
      check(cudaGetLastError());
      return;
    }
    const int n1 = n / 2, n2 = n - n1;
    potrf(o, n1, nullptr, depth + 1);
    if (pend) check(cudaStreamWaitEvent(st, pend, 0));
    trsm(o + n1, n2, o, n1);
    const int p = o + n1;  // first block of A22
    if (n2 == 1 || depth >= nside || n < 8) {
      update(st, false, p, p, n2, n2, p, o, p, o, n1, true);
      potrf(p, n2, nullptr, depth + 1);
      return;
    }
    // look-ahead: the leading half of A22 (what the next level factors first) is updated on the main stream,
    // the rest of the trailing update runs on a low-priority stream concurrently with that factorisation
    const int m1 = n2 / 2, m2 = n2 - m1;
    cudaEvent_t panel_done = record(st);
    update(st, false, p, p, m1, m1, p, o, p, o, n1, true);
    cudaStream_t sd = side[depth];
    check(cudaStreamWaitEvent(sd, panel_done, 0));
    update(sd, true, p + m1, p, m2, m1, p + m1, o, p, o, n1, false);
    update(sd, true, p + m1, p + m1, m2, m2, p + m1, o, p + m1, o, n1, true);
    cudaEvent_t rest_done = record(sd);
    potrf(p, n2, rest_done, depth + 1);
  }
};

}  // namespace

cudaError_t chol_factor(cudaStream_t st, cudaStream_t* side, int nside, double* A, long long ld, int nblocks, double* invD, int* info) {
  Chol c{st, side, nside, A, ld, invD, info};
  cudaError_t e = cudaMemsetAsync(info, 0, sizeof(int), st);
  if (e != cudaSuccess) return e;
  static int lookahead = -1;  // GSP_CHOL_LOOKAHEAD=0 disables the side streams (A/B measurements)
  if (lookahead < 0) {
    const char* env = getenv("GSP_CHOL_LOOKAHEAD");
    lookahead = (env && env[0] == '0') ? 0 : 1;
  }
  if (!lookahead) c.nside = 0;
  c.potrf(0, nblocks, nullptr, 0);
  // every side stream was joined into `st` by the events waited on above, except possibly none: nothing left pending
  for (cudaEvent_t ev : c.events) cudaEventDestroy(ev);
  return c.err;
}

cudaError_t chol_forward_solve(cudaStream_t st, const double* L, long long ld, const double* invD, int nblocks, double* z) {
  if (nblocks <= 0) return cudaSuccess;
  ProfScope prof_("trsv_blocks", st);
  GSP_LAUNCH(trsv_blocks_kernel, dim3(1), dim3(256), 0, st, L, ld, invD, nblocks, z);
  g_launches++;
  return cudaGetLastError();
}

cudaError_t chol_gemv_rows(cudaStream_t st, const double* L, long long ld, long long row0, long long nrows, int kn,
                           const double* y, double* out) {
  if (nrows <= 0) return cudaSuccess;
  ProfScope prof_("gemv_rows", st);
  GSP_LAUNCH(gemv_rows_kernel, dim3((unsigned)((nrows + 255) / 256)), dim3(256), 0, st, L, ld, row0, nrows, kn, y, out);
  g_launches++;
  return cudaGetLastError();
}

cudaError_t sample_gemm(cudaStream_t st, const double* L22, long long ld, int mt, const double* Wp, long long ldw, int nt,
                        double* Z, long long ldz, const double* d2, const long long* sinds, double addmu, long long Ns,
                        long long R) {
  GemmArgs g{};
  g.A = L22; g.lda = ld;
  g.B = Wp; g.ldb = ldw;
  g.C = Z; g.ldc = ldz;
  g.mt = mt; g.nt = nt; g.K = mt * DB; g.klimit = 1;
  g.d2 = d2; g.sinds = sinds; g.addmu = addmu; g.Ns = Ns; g.R = R;
  return launch_gemm<GEMM_SAMPLE, false>(st, g);
}

}  // namespace gsp
