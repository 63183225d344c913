// Device Cholesky / triangular helpers (chol.cu) and assembly launchers (assemble.cu).
#pragma once
#include <vector>

#include "cov.cuh"

namespace gsp {

// A: (nblocks*128)^2 column-major, lower triangle read; on return lower = L, diagonal blocks have
// zero strict-upper part.  invD: nblocks * 128*128 doubles (inverse of every diagonal block of L).
// *info (device int): 0 or 1-based index of the first non-positive pivot.
// `side[nside]`: low-priority streams used for look-ahead (the bulk of every trailing update runs there while
// the next diagonal block / panel proceeds on `st`); everything is joined back into `st` before returning.
// `work` (optional, chol_work_doubles(nblocks) doubles - 0 unless GSP_CHOL_PANELS=1, an experiment measured slower): room for the inverses of the aligned 2-, 4- and 8-block diagonal panels; with it
// every triangular solve against a panel of <= 8 blocks is ONE tile GEMM with the panel's inverse (+ a copy) instead of a recursion of 2 nc - 1 launches.
size_t chol_work_doubles(int nblocks);
cudaError_t chol_factor(cudaStream_t st, cudaStream_t* side, int nside, double* A, long long ld, int nblocks, double* invD, int* info,
                        double* work = nullptr);
// multi-GPU variant: every device holds a full (nblocks*128)^2 buffer `A` with the matrix assembled; panels of PB blocks are
// owned cyclically, factored by their owner and pushed peer-to-peer into the same place on all devices (chol.cu).
struct MgDev {
  int dev;
  cudaStream_t main, upd, copy;
  double* A;
  double* invD;
  int* info;
};
cudaError_t chol_factor_mg(const std::vector<MgDev>& devs, long long ld, int nblocks, int PB);
// z[0 : nblocks*128] <- L^{-1} z  for the leading nblocks diagonal blocks
cudaError_t chol_forward_solve(cudaStream_t st, const double* L, long long ld, const double* invD, int nblocks, double* z);
// out[i] = sum_{k<kn} L[row0+i][k] y[k]
cudaError_t chol_gemv_rows(cudaStream_t st, const double* L, long long ld, long long row0, long long nrows, int kn,
                           const double* y, double* out);
// Z[sinds[i] + r*ldz] = d2[i] + addmu + sum_k L22[i][k] Wt[r + k*ldw]   (lusim.jl:162; Wt = realization-major noise)
cudaError_t sample_gemm(cudaStream_t st, const double* L22, long long ld, int mt, const double* Wp, long long ldw, int nt,
                        double* Z, long long ldz, const double* d2, const long long* sinds, double addmu, long long Ns,
                        long long R);

void launch_assemble(cudaStream_t st, const CovDev& m, const DomDev& drow, const DomDev& dcol, const long long* rowmap,
                     const long long* colmap, long long nrow, long long ncol, double* out, long long ld, bool lower_only);
void launch_cov_to_center(cudaStream_t st, int sms, const CovDev& m, const DomDev& d, long long eref, double* out);

}  // namespace gsp
