// Tiled pairwise covariance assembly: replaces `_pairwise` (src/utils.jl:50-62 ->
// GeoStatsFunctions.pairwise) as called from lusim.jl:88,95,96 and fftsim.jl:86.
// HBM-write-bound: 8 bytes per matrix entry; coordinates are staged once per tile in shared
// memory, every thread owns 4 consecutive rows x 4 columns and stores 16-byte vectors.
#include "cov.cuh"

namespace gsp {

constexpr int AT = 64;  // tile edge

// out[p + q*ld] = C(x_row(p) - x_col(q)), p < nrow, q < ncol.
// rowmap/colmap: 0-based element index per matrix row/col, or -1 = padding (identity row/col);
// nullptr = identity map.  lower_only: skip tiles strictly above the diagonal.
__global__ void __launch_bounds__(256) assemble_kernel(CovDev m, DomDev drow, DomDev dcol, const long long* __restrict__ rowmap,
                                                       const long long* __restrict__ colmap, long long nrow, long long ncol,
                                                       double* __restrict__ out, long long ld, int lower_only, int vec_ok,
                                                       long long row_base) {
  // row_base: global index of matrix row 0 of this launch (a device assembling only a band of rows passes rowmap / out already
  // offset to the band); it only enters the "is this the diagonal" decisions
  const long long r0 = (long long)blockIdx.x * AT, c0 = (long long)blockIdx.y * AT;
  if (lower_only && row_base + r0 + AT <= c0) return;
  __shared__ double xr[3][AT], xc[3][AT];
  __shared__ int padr[AT], padc[AT];
  const int tid = threadIdx.x;
  if (tid < 2 * AT) {
    const bool isrow = tid < AT;
    const int l = isrow ? tid : tid - AT;
    const long long g = (isrow ? r0 : c0) + l;
    const long long lim = isrow ? nrow : ncol;
    const long long* map = isrow ? rowmap : colmap;
    long long e = -1;
    if (g < lim) e = map ? map[g] : g;
    double x = 0, y = 0, z = 0;
    if (e >= 0) centroid(isrow ? drow : dcol, e, x, y, z);
    if (isrow) {
      xr[0][l] = x; xr[1][l] = y; xr[2][l] = z; padr[l] = e < 0;
    } else {
      xc[0][l] = x; xc[1][l] = y; xc[2][l] = z; padc[l] = e < 0;
    }
  }
  __syncthreads();
  const int tx = tid & 15, ty = tid >> 4;
  const int lr = 4 * tx;
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    const int lc = ty + 16 * cc;
    const long long q = c0 + lc;
    if (q >= ncol) continue;
    double v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long p = r0 + lr + i;
      if (padr[lr + i] || padc[lc])
        v[i] = (row_base + p == q) ? 1.0 : 0.0;
      else
        v[i] = cov_eval(m, xr[0][lr + i] - xc[0][lc], xr[1][lr + i] - xc[1][lc], xr[2][lr + i] - xc[2][lc]);
    }
    double* dst = out + (r0 + lr) + q * ld;
    if (vec_ok && r0 + lr + 3 < nrow) {
      *reinterpret_cast<double2*>(dst) = make_double2(v[0], v[1]);
      *reinterpret_cast<double2*>(dst + 2) = make_double2(v[2], v[3]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (r0 + lr + i < nrow) dst[i] = v[i];
    }
  }
}

void launch_assemble(cudaStream_t st, const CovDev& m, const DomDev& drow, const DomDev& dcol, const long long* rowmap,
                     const long long* colmap, long long nrow, long long ncol, double* out, long long ld, bool lower_only, long long row_base) {
  dim3 grid((unsigned)((nrow + AT - 1) / AT), (unsigned)((ncol + AT - 1) / AT));
  int vec_ok = (ld % 2 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  ProfScope prof_("assemble", st);
  GSP_LAUNCH(assemble_kernel, grid, dim3(256), 0, st, m, drow, dcol, rowmap, colmap, nrow, ncol, out, ld, lower_only ? 1 : 0,
             vec_ok, row_base);
  g_launches++;
}

// covariance from one reference element to every element of a grid (fftsim.jl:84-86): out[e] = C(x_e - x_ref)
__global__ void __launch_bounds__(256) cov_to_center_kernel(CovDev m, DomDev d, long long eref, double* __restrict__ out) {
  double cx, cy, cz;
  centroid(d, eref, cx, cy, cz);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < d.nelems; e += stride) {
    double x, y, z;
    centroid(d, e, x, y, z);
    out[e] = cov_eval(m, cx - x, cy - y, cz - z);
  }
}

void launch_cov_to_center(cudaStream_t st, int sms, const CovDev& m, const DomDev& d, long long eref, double* out) {
  long long blocks = (d.nelems + 255) / 256;
  if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
  ProfScope prof_("cov_to_center", st);
  GSP_LAUNCH(cov_to_center_kernel, dim3((unsigned)blocks), dim3(256), 0, st, m, d, eref, out);
  g_launches++;
}

}  // namespace gsp
