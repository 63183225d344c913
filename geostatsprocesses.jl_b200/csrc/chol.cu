// Blocked Cholesky on the device: replaces cholesky(Symmetric(.)).L and the
// triangular solves of lusim.jl:92,98-103.  Recursive right-looking formulation over 128x128
// tiles: every flop outside the 128-wide diagonal blocks runs in the DMMA tile GEMM (gemm.cuh);
// diagonal blocks are factorised AND inverted by one CTA (potrf_diag_kernel) so that every
// triangular solve becomes a tile GEMM with the explicit inverse.
#include "chol.h"

#include "gemm.cuh"

namespace gsp {

constexpr int DB = 128;   // diagonal block
constexpr int DLD = 129;  // padded smem leading dimension (row-major S[r][c])
constexpr int DPW = 16;   // inner panel width
constexpr int DIAG_SMEM = (DB * DLD + DB + DPW * (DPW + 1) + DPW) * 8;

// Factor A[blk,blk] (128x128, lower) in place -> L (strict upper zeroed), write inv(L) to invD.
__global__ void __launch_bounds__(256, 1) potrf_diag_kernel(double* __restrict__ A, long long lda, long long blk,
                                                            double* __restrict__ invD, int* __restrict__ info) {
  GSP_DYN_SMEM(smem);
  double* S = reinterpret_cast<double*>(smem);   // [DB][DLD]
  double* rowbuf = S + DB * DLD;                 // [DB]
  double* D16 = rowbuf + DB;                     // [DPW][DPW+1]
  double* rdiag = D16 + DPW * (DPW + 1);         // [DPW]
  const int tid = threadIdx.x;
  double* Ab = A + blk * DB * (lda + 1);

  for (int idx = tid; idx < DB * DB; idx += 256) {
    const int r = idx & (DB - 1), c = idx >> 7;
    S[r * DLD + c] = (r >= c) ? Ab[r + (long long)c * lda] : 0.0;
  }
  __syncthreads();

  for (int p = 0; p < DB / DPW; ++p) {
    const int c0 = p * DPW;
    // (a) 16x16 diagonal block in registers of warp 0, one row per lane
    if (tid < 32) {
      const int l = tid;
      double row[DPW];
#pragma unroll
      for (int k = 0; k < DPW; ++k) row[k] = (l < DPW) ? S[(c0 + l) * DLD + c0 + k] : 0.0;
#pragma unroll
      for (int j = 0; j < DPW; ++j) {
        const double piv = __shfl_sync(0xffffffffu, row[j], j);
        if (!(piv > 0.0) && l == 0) atomicCAS(info, 0, (int)(blk * DB + c0 + j + 1));
        const double s = sqrt(piv);
        const double rinv = 1.0 / s;
        if (l == j) {
          row[j] = s;
          rdiag[j] = rinv;
        } else if (l > j) {
          row[j] *= rinv;
        }
#pragma unroll
        for (int k = j + 1; k < DPW; ++k) {
          const double lkj = __shfl_sync(0xffffffffu, row[j], k);
          if (l >= k) row[k] -= row[j] * lkj;
        }
      }
      if (l < DPW) {
#pragma unroll
        for (int k = 0; k < DPW; ++k) {
          const double v = (k <= l) ? row[k] : 0.0;
          S[(c0 + l) * DLD + c0 + k] = v;
          D16[l * (DPW + 1) + k] = v;
        }
      }
    }
    __syncthreads();
    // (b) panel rows below the diagonal block: x * D^T = b, one row per thread
    if (tid < DB && tid >= c0 + DPW) {
      double x[DPW];
      double* srow = S + tid * DLD + c0;
#pragma unroll
      for (int c = 0; c < DPW; ++c) {
        double v = srow[c];
#pragma unroll
        for (int k = 0; k < c; ++k) v -= x[k] * D16[c * (DPW + 1) + k];
        x[c] = v * rdiag[c];
      }
#pragma unroll
      for (int c = 0; c < DPW; ++c) srow[c] = x[c];
    }
    __syncthreads();
    // (c) trailing update of the lower triangle right of the panel (rank-16)
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

  // in-place inverse of the lower-triangular block, row by row:
  // X[i][c] = -(sum_{k=c}^{i-1} L[i][k] X[k][c]) / L[i][i],  X[i][i] = 1 / L[i][i]
  for (int i = 0; i < DB; ++i) {
    if (tid <= i) rowbuf[tid] = S[i * DLD + tid];
    __syncthreads();
    if (tid < i) {
      const int c = tid;
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int k = c;
      for (; k + 3 < i; k += 4) {
        s0 += rowbuf[k] * S[k * DLD + c];
        s1 += rowbuf[k + 1] * S[(k + 1) * DLD + c];
        s2 += rowbuf[k + 2] * S[(k + 2) * DLD + c];
        s3 += rowbuf[k + 3] * S[(k + 3) * DLD + c];
      }
      for (; k < i; ++k) s0 += rowbuf[k] * S[k * DLD + c];
      S[i * DLD + c] = -((s0 + s1) + (s2 + s3)) / rowbuf[i];
    } else if (tid == i) {
      S[i * DLD + i] = 1.0 / rowbuf[i];
    }
    __syncthreads();
  }
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
  cudaStream_t st;
  double* A;
  long long ld;
  double* invD;
  int* info;
  cudaError_t err = cudaSuccess;

  double* at(int br, int bc) const { return A + (long long)br * DB + (long long)bc * DB * ld; }

  void check(cudaError_t e) {
    if (err == cudaSuccess && e != cudaSuccess) err = e;
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
    GemmArgs g{};
    g.A = at(r0, c0); g.lda = ld;
    g.B = at(c0 + c1, c0); g.ldb = ld;
    g.C = at(r0, c0 + c1); g.ldc = ld;
    g.mt = nr; g.nt = nc - c1; g.K = c1 * DB;
    check(launch_gemm<GEMM_SUB, false>(st, g));
    trsm(r0, nr, c0 + c1, nc - c1);
  }

  void potrf(int o, int n) {
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
    potrf(o, n1);
    trsm(o + n1, n2, o, n1);
    GemmArgs g{};
    g.A = at(o + n1, o); g.lda = ld;
    g.B = at(o + n1, o); g.ldb = ld;
    g.C = at(o + n1, o + n1); g.ldc = ld;
    g.mt = n2; g.nt = n2; g.K = n1 * DB; g.tri = 1;
    check(launch_gemm<GEMM_SUB, false>(st, g));
    potrf(o + n1, n2);
  }
};

}  // namespace

cudaError_t chol_factor(cudaStream_t st, double* A, long long ld, int nblocks, double* invD, int* info) {
  Chol c{st, A, ld, invD, info};
  cudaError_t e = cudaMemsetAsync(info, 0, sizeof(int), st);
  if (e != cudaSuccess) return e;
  c.potrf(0, nblocks);
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
  return launch_gemm<GEMM_SAMPLE, true>(st, g);
}

}  // namespace gsp
