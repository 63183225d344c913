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
cudaError_t chol_factor(cudaStream_t st, cudaStream_t* side, int nside, double* A, long long ld, int nblocks, double* invD, int* info);

// the other devices' buffers a finished block is multicast into (peer-mapped pointers; n = 0 on a single device)
struct DiagPeers {
  double* A[7];
  double* invD[7];
  int n;
};
// Panel factorization over the G >= 1 devices of a context with row-panel ownership (chol.cu).  Every device holds a full
// (nblocks*128)^2 buffer `A` in which the block rows it owns (chol_dist_owned_rows) are assembled; on return every device holds the
// whole factor and all block inverses.  `rows`: device scratch of nblocks ints, `flags`: chol_dist_flag_ints() ints.  main / aux: high-priority streams, upd: low priority.
struct DistDev {
  int dev;
  cudaStream_t main, aux, upd;
  double* A;
  double* invD;
  int* info;
  int* rows;
  int* flags;  // chol_dist_flag_ints() ints: inter-CTA flags of the fused diagonal-square kernel
};
int chol_dist_flag_ints();
void chol_dist_owned_rows(int nblocks, int PB, int G, int g, std::vector<int>* rows);
cudaError_t chol_factor_dist(const std::vector<DistDev>& devs, long long ld, int nblocks, int PB);
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
                     const long long* colmap, long long nrow, long long ncol, double* out, long long ld, bool lower_only, long long row_base = 0);
void launch_cov_to_center(cudaStream_t st, int sms, const CovDev& m, const DomDev& d, long long eref, double* out);

}  // namespace gsp
