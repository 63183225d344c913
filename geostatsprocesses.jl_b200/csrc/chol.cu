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

#ifndef GSP_CHOL_MG_RESERVE_DEFAULT
#define GSP_CHOL_MG_RESERVE_DEFAULT 0
#endif
#ifndef GSP_CHOL_RESERVE_DEFAULT
#define GSP_CHOL_RESERVE_DEFAULT 8  // measured on B200: C3 66.4 -> 64.6 ms, 32k nodes 394 -> 385 ms (tools/gpu_cholreserve.py); 16-24 the same, 48 worse
#endif
constexpr int DB = 128;   // diagonal block
constexpr int DLD = 132;  // smem leading dimension of the COLUMN-major block S[c * DLD + r]: DMMA fragment loads of both kinds
                          // (8 consecutive rows x 4 columns, 4 consecutive rows x 8 columns) are bank-conflict-free for DLD = 4 mod 16
constexpr int DPW = 16;   // inner panel width
constexpr int DVL = 20;   // leading dimension of the 16x16 inverse blocks (transposed), = 4 mod 16
constexpr int DIAG_SMEM = (DB * DLD + (DB / DPW) * DPW * DVL) * 8;

// 16x16 diagonal block of a panel, factorised by ONE warp in registers (lane l < 16 owns row l), right-looking,
// no block barriers: per column one shuffle broadcasts the pivot, rsqrt gives 1/L_jj (kept by every lane in rd[]), the
// remaining columns are updated with shuffled L_kj.  Fully unrolled through the template recursion (static register indexing).
template <int J>
GSP_DEV void diag16_cols(double (&row)[DPW], double (&rd)[DPW], int l, int* info, int gpos) {
  if constexpr (J < DPW) {
    const double piv = __shfl_sync(0xffffffffu, row[J], J);
    if (!(piv > 0.0) && l == 0) atomicCAS(info, 0, gpos + J + 1);
    const double rinv = rsqrt(piv);
    rd[J] = rinv;
    if (l == J) {
      row[J] = piv * rinv;
    } else if (l > J) {
      row[J] *= rinv;
    }
#pragma unroll
    for (int k = J + 1; k < DPW; ++k) {
      const double lkj = __shfl_sync(0xffffffffu, row[J], k);
      if (l >= k) row[k] -= row[J] * lkj;
    }
    diag16_cols<J + 1>(row, rd, l, info, gpos);
  }
}

// inverse of that 16x16 factor, same warp: lane j solves L x = e_j by forward substitution (x[i] = inv[i][j]), the entries of L
// come from their owner lanes by shuffle
template <int I>
GSP_DEV void diag16_inverse(const double (&row)[DPW], const double (&rd)[DPW], double (&x)[DPW], int l) {
  if constexpr (I < DPW) {
    double s0 = (l == I) ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
    for (int k = 0; k < I; ++k) {
      const double lik = __shfl_sync(0xffffffffu, row[k], I);
      if (k & 1)
        s1 -= lik * x[k];
      else
        s0 -= lik * x[k];
    }
    x[I] = (s0 + s1) * rd[I];
    diag16_inverse<I + 1>(row, rd, x, l);
  }
}

// DMMA fragment addressing on the column-major block: element (r, c) lives at S[c * DLD + r]
//   A fragment (8 rows x 4 k):  lane holds M[r0 + lane/4][k0 + lane%4]
//   B fragment (4 k x 8 cols):  lane holds M[k0 + lane%4][c0 + lane/4]          (row-type: B = a sub-block itself)
//   B^T fragment:               lane holds M[c0 + lane/4][k0 + lane%4]          (B[k][n] = M[n][k]: same addressing as A)
//   C fragment (8 x 8):         lane holds M[r0 + lane/4][c0 + 2*(lane%4) + {0,1}]
GSP_DEV double frag_a(const double* S, int r0, int k0, int lane) { return S[(k0 + (lane & 3)) * DLD + r0 + (lane >> 2)]; }
GSP_DEV double frag_b(const double* S, int k0, int c0, int lane) { return S[(c0 + (lane >> 2)) * DLD + k0 + (lane & 3)]; }

// Factor A[blk,blk] (128x128, lower) in place -> L (strict upper zeroed), write inv(L) to invD.
// One CTA, 8 warps.  Left-looking over 16-wide panels; everything but the 16x16 diagonal blocks runs on DMMA 8x8x4 tiles out of
// shared memory:  (U) panel -= L[:, :c0] * L[c0:c0+16, :c0]^T   (F) warp 0 factors and inverts the 16x16 block in registers
// (S) rows below = panel * inv(L16)^T.  The inverse of the whole block is then assembled from the eight 16x16 inverses by
// X21 = -inv(C) * B * inv(A) over 16 -> 32 -> 64 wide halves, again on DMMA tiles (one 8-row block per warp and level).
__global__ void __launch_bounds__(256, 1) potrf_diag_kernel(double* __restrict__ A, long long lda, long long blk,
                                                            double* __restrict__ invD, int* __restrict__ info,
                                                            double* __restrict__ inv2, long long ld2) {
  GSP_DYN_SMEM(smem);
  double* S = reinterpret_cast<double*>(smem);   // [DB cols][DLD]
  double* Dv = S + DB * DLD;                     // [8 panels][16 k][DVL]: Dv[p][k * DVL + n] = inv(L16_p)[n][k]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  double* Ab = A + blk * DB * (lda + 1);

  // load the lower triangle (zeros above): 16 independent loads in flight per thread - a single CTA is latency-bound here.
  // (One bulk async copy per column was measured SLOWER on the B200: 87 vs 65 us per block; 128 small, 16-byte-aligned copies
  // from one SM serialise in the TMA unit.)
  for (int base = 0; base < DB * DB; base += 16 * 256) {
    double v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int idx = base + u * 256 + tid;
      const int r = idx & (DB - 1), c = idx >> 7;
      v[u] = (r >= c) ? Ab[r + (long long)c * lda] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int idx = base + u * 256 + tid;
      S[(idx >> 7) * DLD + (idx & (DB - 1))] = v[u];
    }
  }
  __syncthreads();

  for (int p = 0; p < DB / DPW; ++p) {
    const int c0 = p * DPW;
    // (U) left-looking update of the panel's rows c0..127 with the finished columns [0, c0).  Warp 0 takes the 16 diagonal rows
    // and goes straight on to (F); warps 1-7 share the rows below meanwhile (they read rows c0..c0+15 only in columns < c0,
    // which (F) does not touch).
    if (p > 0) {
      const int mb0 = c0 / 8 + (warp == 0 ? 0 : 1 + warp), mbstep = warp == 0 ? 1 : 7, mbend = warp == 0 ? c0 / 8 + 2 : DB / 8;
      for (int mb = mb0; mb < mbend; mb += mbstep) {
        const int r0 = mb * 8;
        double acc[2][2];
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
          const double* cp = S + (c0 + nb * 8 + 2 * (lane & 3)) * DLD + r0 + (lane >> 2);
          acc[nb][0] = cp[0];
          acc[nb][1] = cp[DLD];
        }
        for (int k0 = 0; k0 < c0; k0 += 4) {
          const double a = -frag_a(S, r0, k0, lane);
          dmma884(acc[0][0], acc[0][1], a, frag_a(S, c0, k0, lane));       // B[k][n] = L[c0 + n][k0 + k]
          dmma884(acc[1][0], acc[1][1], a, frag_a(S, c0 + 8, k0, lane));
        }
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
          double* cp = S + (c0 + nb * 8 + 2 * (lane & 3)) * DLD + r0 + (lane >> 2);
          cp[0] = acc[nb][0];
          cp[DLD] = acc[nb][1];
        }
      }
      __syncwarp();
    }
    // (F) 16x16 diagonal block: factor + inverse in the registers of warp 0
    if (warp == 0) {
      const int l = lane;
      double row[DPW], rd[DPW], x[DPW];
#pragma unroll
      for (int k = 0; k < DPW; ++k) row[k] = (l < DPW) ? S[(c0 + k) * DLD + c0 + l] : 0.0;
      diag16_cols<0>(row, rd, l, info, (int)(blk * DB) + c0);
      diag16_inverse<0>(row, rd, x, l);
      if (l < DPW) {
        double* dv = Dv + p * DPW * DVL + l * DVL;
#pragma unroll
        for (int k = 0; k < DPW; ++k) {
          S[(c0 + k) * DLD + c0 + l] = (k <= l) ? row[k] : 0.0;
          dv[k] = (k >= l) ? x[k] : 0.0;   // inv[k][l], zero above the diagonal
        }
      }
    }
    __syncthreads();
    // (S) rows below the diagonal block: X = P * inv(L16)^T, in place (a warp only touches its own rows)
    {
      const double* dv = Dv + p * DPW * DVL;
      for (int mb = (c0 + DPW) / 8 + warp; mb < DB / 8; mb += 8) {
        const int r0 = mb * 8;
        double af[4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) af[ks] = frag_a(S, r0, c0 + 4 * ks, lane);
        double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
          for (int nb = 0; nb < 2; ++nb) {
            // B[k][n] = inv(L16)[n][k] = dv[k * DVL + n]
            const double bf = dv[(4 * ks + (lane & 3)) * DVL + nb * 8 + (lane >> 2)];
            dmma884(acc[nb][0], acc[nb][1], af[ks], bf);
          }
        }
        __syncwarp();
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
          double* cp = S + (c0 + nb * 8 + 2 * (lane & 3)) * DLD + r0 + (lane >> 2);
          cp[0] = acc[nb][0];
          cp[DLD] = acc[nb][1];
        }
      }
    }
    __syncthreads();
  }

  // write L (upper triangle explicitly zero)
  for (int idx = tid; idx < DB * DB; idx += 256) {
    const int r = idx & (DB - 1), c = idx >> 7;
    Ab[r + (long long)c * lda] = (r >= c) ? S[c * DLD + r] : 0.0;
  }
  __syncthreads();

  // inverse: the 16x16 diagonal inverses go in place ...
  for (int idx = tid; idx < (DB / DPW) * DPW * DPW; idx += 256) {
    const int p = idx >> 8, j = (idx >> 4) & 15, i = idx & 15;   // element inv_p[i][j]
    S[(p * DPW + j) * DLD + p * DPW + i] = Dv[p * DPW * DVL + j * DVL + i];
  }
  __syncthreads();
  // ... then X21 = -inv(C) B inv(A) level by level; the B blocks of a level have 64 rows in total: one 8-row block per warp
  for (int H = DPW; H < DB; H *= 2) {
    const int per = H / 8;                       // 8-row blocks per pair
    const int o = (warp / per) * 2 * H;          // the pair's origin
    const int i0 = (warp % per) * 8;             // this warp's rows within B / X21
    const int rB = o + H + i0;                   // global row of the block
    const int nblk = H / 8;
    double acc[8][2];
    {
      // T = B * inv(A): T[i][n] = sum_{k >= n} B[i][k] invA[k][n]; own rows only -> in place
      double af[16];
#pragma unroll
      for (int ks = 0; ks < 16; ++ks) af[ks] = (ks < H / 4) ? frag_a(S, rB, o + 4 * ks, lane) : 0.0;
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        acc[nb][0] = acc[nb][1] = 0.0;
        if (nb < nblk) {
#pragma unroll
          for (int ks = 0; ks < 16; ++ks)
            if (ks < H / 4 && ks >= 2 * nb) dmma884(acc[nb][0], acc[nb][1], af[ks], frag_b(S, o + 4 * ks, o + nb * 8, lane));
        }
      }
      __syncwarp();
#pragma unroll
      for (int nb = 0; nb < 8; ++nb)
        if (nb < nblk) {
          double* cp = S + (o + nb * 8 + 2 * (lane & 3)) * DLD + rB + (lane >> 2);
          cp[0] = acc[nb][0];
          cp[DLD] = acc[nb][1];
        }
    }
    __syncthreads();
    {
      // X21 = -inv(C) * T: X[i][n] = -sum_{k <= i} invC[i][k] T[k][n]; every warp reads all of T before anyone overwrites it
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) acc[nb][0] = acc[nb][1] = 0.0;
      for (int k0 = 0; k0 < i0 + 8; k0 += 4) {
        const double a = -frag_a(S, rB, o + H + k0, lane);
#pragma unroll
        for (int nb = 0; nb < 8; ++nb)
          if (nb < nblk) dmma884(acc[nb][0], acc[nb][1], a, frag_b(S, o + H + k0, o + nb * 8, lane));
      }
    }
    __syncthreads();
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
      if (nb < nblk) {
        double* cp = S + (o + nb * 8 + 2 * (lane & 3)) * DLD + rB + (lane >> 2);
        cp[0] = acc[nb][0];
        cp[DLD] = acc[nb][1];
      }
    __syncthreads();
  }

  double* Xo = invD + blk * DB * DB;
  for (int idx = tid; idx < DB * DB; idx += 256) {
    const int r = idx & (DB - 1), c = idx >> 7;
    const double v = (r >= c) ? S[c * DLD + r] : 0.0;
    Xo[r + c * DB] = v;
    if (inv2) inv2[r + c * ld2] = v;  // second copy on the diagonal of the panel-inverse workspace (chol_factor)
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
  int side_ctas = 0;  // > 0: look-ahead GEMMs run as persistent grids of this many CTAs, the other SMs stay free for the main stream
  // panel inverses (chol_factor with a workspace): group g = blocks [8g, 8g + 8) owns a dense 1024 x 1024 slot; the inverse of an
  // aligned s-block panel (s = 1, 2, 4, 8) is the s*128-square sub-matrix on the slot's diagonal (ld = PIL).  built[k] marks the
  // panels of size 2^k that are complete.
  static constexpr int PIG = 8;
  static constexpr long long PIL = (long long)PIG * DB;
  double* linv = nullptr;
  double* tmp = nullptr;   // (PIG/2 * 128)^2 scratch of the level builds
  double* xbuf = nullptr;  // nb_total x PIG blocks: result of a panel solve before it is copied over the panel
  int nb_total = 0;
  std::vector<unsigned char> built[4];

  double* linv_at(int blk_row, int blk_col) const {  // both inside the same 8-group
    const int grp = blk_row / PIG;
    return linv + (long long)grp * PIL * PIL + (long long)(blk_row % PIG) * DB + (long long)(blk_col % PIG) * DB * PIL;
  }
  static int lg2(int s) { return s == 1 ? 0 : (s == 2 ? 1 : (s == 4 ? 2 : (s == 8 ? 3 : -1))); }
  bool panel_ready(int c0, int nc) const {
    const int k = lg2(nc);
    return linv && k >= 0 && c0 % nc == 0 && (size_t)(c0 / nc) < built[k].size() && built[k][c0 / nc];
  }
  // inverse of the aligned s-block panel at o from its two halves: X21 = -inv(C) * (B * inv(A))
  void build_panel_inverse(int o, int s) {
    const int k = lg2(s), h = s / 2;
    if (!linv || k < 1 || o % s != 0 || !panel_ready(o, h) || !panel_ready(o + h, h)) return;
    GemmArgs g{};
    g.A = at(o + h, o); g.lda = ld;
    g.B = linv_at(o, o); g.ldb = PIL;            // K x N: inv(A)[k][j]
    g.C = tmp; g.ldc = (long long)h * DB;
    g.mt = h; g.nt = h; g.K = h * DB;
    check(launch_gemm<GEMM_SET, true>(st, g));
    GemmArgs g2{};
    g2.A = linv_at(o + h, o + h); g2.lda = PIL;  // inv(C)
    g2.B = tmp; g2.ldb = (long long)h * DB;      // K x N: T
    g2.C = linv_at(o + h, o); g2.ldc = PIL;      // zero on entry (workspace cleared by chol_factor)
    g2.mt = h; g2.nt = h; g2.K = h * DB;
    check(launch_gemm<GEMM_SUB, true>(st, g2));
    built[k][o / s] = 1;
  }

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
      g.max_ctas = sliced ? side_ctas : 0;
      check(launch_gemm<GEMM_SUB, false>(s, g));
    }
  }

  // X * L[c0:c0+nc, c0:c0+nc]^T = A[r0:r0+nr, c0:c0+nc]   (in place)
  void trsm(int r0, int nr, int c0, int nc) {
    if (nr <= 0 || nc <= 0) return;
    if (nc > 1 && panel_ready(c0, nc)) {
      GemmArgs g{};
      g.A = at(r0, c0); g.lda = ld;
      g.B = linv_at(c0, c0); g.ldb = PIL;   // N x K: inv(L)[j][k], lower triangular
      g.C = xbuf; g.ldc = (long long)nr * DB;
      g.mt = nr; g.nt = nc; g.K = nc * DB; g.strip = 1;
      check(launch_gemm<GEMM_SET, false>(st, g));
      check(cudaMemcpy2DAsync(at(r0, c0), (size_t)ld * sizeof(double), xbuf, (size_t)nr * DB * sizeof(double), (size_t)nr * DB * sizeof(double),
                              (size_t)nc * DB, cudaMemcpyDeviceToDevice, st));
      return;
    }
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
      double* inv2 = linv ? linv_at(o, o) : nullptr;
      GSP_LAUNCH(kfn, dim3(1), dim3(256), (size_t)DIAG_SMEM, st, A, ld, (long long)o, invD, info, inv2, PIL);
      g_launches++;
      check(cudaGetLastError());
      if (linv) built[0][o] = 1;
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
      if (n <= PIG) build_panel_inverse(o, n);
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
    if (n <= PIG) build_panel_inverse(o, n);
  }
};

}  // namespace

// GSP_CHOL_PANELS=1 turns the panel-inverse solves on.  Measured on the B200 (session 3): C3 factorization 70.0 ms against 64.7 ms for
// the recursive solves, 32k nodes 387 against 384 ms (an in-place variant, one CTA per 64-row strip walking the column tiles right to
// left, was worse still: 73.9 ms) - the recursion's many small launches run back to back and spread over more CTAs than one
// triangular-K GEMM per panel, so the saved launches buy nothing.  Off by default.
static bool chol_panels_enabled() {
  static int use_panels = -1;
  if (use_panels < 0) {
    const char* env = getenv("GSP_CHOL_PANELS");
    use_panels = (env && env[0] == '1') ? 1 : 0;
  }
  return use_panels != 0;
}

size_t chol_work_doubles(int nblocks) {
  if (!chol_panels_enabled()) return 0;
  const size_t groups = (size_t)(nblocks + Chol::PIG - 1) / Chol::PIG;
  return groups * (size_t)Chol::PIL * Chol::PIL + (size_t)(Chol::PIG / 2 * DB) * (Chol::PIG / 2 * DB) +
         (size_t)nblocks * DB * Chol::PIL;
}

cudaError_t chol_factor(cudaStream_t st, cudaStream_t* side, int nside, double* A, long long ld, int nblocks, double* invD, int* info,
                        double* work) {
  Chol c{st, side, nside, A, ld, invD, info};
  cudaError_t e = cudaMemsetAsync(info, 0, sizeof(int), st);
  if (e != cudaSuccess) return e;
  if (work && chol_panels_enabled() && nblocks >= 4) {
    const size_t groups = (size_t)(nblocks + Chol::PIG - 1) / Chol::PIG;
    c.linv = work;
    c.tmp = work + groups * (size_t)Chol::PIL * Chol::PIL;
    c.xbuf = c.tmp + (size_t)(Chol::PIG / 2 * DB) * (Chol::PIG / 2 * DB);
    c.nb_total = nblocks;
    for (int k = 0; k < 4; ++k) c.built[k].assign((size_t)(nblocks >> k) + 1, 0);
    // the level builds and the strip kernels read whole tiles of the (triangular) inverses: everything not written must be zero
    e = cudaMemsetAsync(work, 0, groups * (size_t)Chol::PIL * Chol::PIL * sizeof(double), st);
    if (e != cudaSuccess) return e;
  }
  static int lookahead = -1;  // GSP_CHOL_LOOKAHEAD=0 disables the side streams (A/B measurements)
  if (lookahead < 0) {
    const char* env = getenv("GSP_CHOL_LOOKAHEAD");
    lookahead = (env && env[0] == '0') ? 0 : 1;
  }
  if (!lookahead) c.nside = 0;
  // GSP_CHOL_RESERVE=r: SMs kept free of look-ahead work.  A one-CTA-per-tile look-ahead GEMM fills every SM for the length of
  // a tile (1.2 ms at K = 8192), and the short kernels of the critical path queue behind it whatever their stream priority.
  static int reserve = -1;
  if (reserve < 0) {
    const char* env = getenv("GSP_CHOL_RESERVE");
    reserve = env ? atoi(env) : GSP_CHOL_RESERVE_DEFAULT;
    if (reserve < 0) reserve = 0;
  }
  if (reserve > 0) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    c.side_ctas = sms - reserve > 8 ? sms - reserve : 8;
  }
  c.potrf(0, nblocks, nullptr, 0);
  // every side stream was joined into `st` by the events waited on above, except possibly none: nothing left pending
  for (cudaEvent_t ev : c.events) cudaEventDestroy(ev);
  return c.err;
}

// ------------------------------------------------------------------ multi-GPU factorization
// 1-D block-cyclic panels of PB 128-blocks over the G devices of a context, every device holding the full
// matrix buffer.  Step q: owner(q) = q mod G factors panel q (diagonal potrf + TRSM of the rows below), the
// panel is pushed device-to-device (cudaMemcpy2DAsync over NVLink peer access) INTO THE SAME POSITION of every
// other device's matrix, and every device applies it to the panels it owns.  Look-ahead: the owner of
// panel q+1 updates that panel first on its main stream, factors it and starts the next broadcast while the
// bulk updates of step q still run on the update streams.  Because panels land in place, every device ends
// with the complete factor L: the all-gather needed for the realization-sharded sampling is free.
cudaError_t chol_factor_mg(const std::vector<MgDev>& devs, long long ld, int nblocks, int PB) {
  const int G = (int)devs.size();
  const int Q = (nblocks + PB - 1) / PB;
  cudaError_t err = cudaSuccess;
  auto check = [&](cudaError_t e) {
    if (err == cudaSuccess && e != cudaSuccess) err = e;
  };
  // GSP_CHOL_MG_RESERVE=r: the bulk updates on the update streams run as persistent grids that leave r SMs to the main stream, where the
  // next panel is updated, factored and sent (see chol_factor's GSP_CHOL_RESERVE)
  static int mg_reserve = -1;
  if (mg_reserve < 0) {
    const char* env = getenv("GSP_CHOL_MG_RESERVE");
    mg_reserve = env ? atoi(env) : GSP_CHOL_MG_RESERVE_DEFAULT;
    if (mg_reserve < 0) mg_reserve = 0;
  }
  std::vector<Chol> ch;
  ch.reserve(G);
  for (int g = 0; g < G; ++g) {
    check(cudaSetDevice(devs[g].dev));
    ch.push_back(Chol{devs[g].main, nullptr, 0, devs[g].A, ld, devs[g].invD, devs[g].info});
    if (mg_reserve > 0) {
      int sms = 0;
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, devs[g].dev);
      ch.back().side_ctas = sms - mg_reserve > 8 ? sms - mg_reserve : 8;
    }
    check(cudaMemsetAsync(devs[g].info, 0, sizeof(int), devs[g].main));
  }
  std::vector<cudaEvent_t> evs;
  auto record = [&](int g, cudaStream_t s) {
    cudaEvent_t ev;
    check(cudaSetDevice(devs[g].dev));
    check(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    check(cudaEventRecord(ev, s));
    evs.push_back(ev);
    return ev;
  };
  std::vector<cudaEvent_t> upd_done(G, nullptr);   // last event of device g's update stream
  std::vector<cudaEvent_t> have(G, nullptr);       // panel q is present on device g (factored there or received)
  const size_t esz = sizeof(double);
  for (int q = 0; q < Q && err == cudaSuccess; ++q) {
    const int o = q % G;
    const int r0 = q * PB;
    const int nq = (nblocks - r0 < PB) ? nblocks - r0 : PB;
    const int below = nblocks - r0 - nq;
    // ---- factor panel q on its owner.  No extra wait: the panel's last update (by panel q-1) ran on this main stream as
    // the look-ahead update of step q-1, and that update already waited for the update stream's earlier work on the panel;
    // the bulk updates of step q-1 touch other panels and keep running underneath.
    check(cudaSetDevice(devs[o].dev));
    ch[o].potrf(r0, nq, nullptr, 0);
    ch[o].trsm(r0 + nq, below, r0, nq);
    check(ch[o].err);
    cudaEvent_t ready = record(o, devs[o].main);
    // ---- broadcast: rows >= r0*128 of the panel's columns, plus the inverses of its diagonal blocks
    for (int g = 0; g < G; ++g) {
      if (g == o) {
        have[g] = ready;
        continue;
      }
      check(cudaSetDevice(devs[g].dev));
      check(cudaStreamWaitEvent(devs[g].copy, ready, 0));
      const size_t off = (size_t)r0 * DB * (size_t)(ld + 1);
      check(cudaMemcpy2DAsync(devs[g].A + off, (size_t)ld * esz, devs[o].A + off, (size_t)ld * esz, (size_t)(nblocks - r0) * DB * esz,
                              (size_t)nq * DB, cudaMemcpyDefault, devs[g].copy));
      check(cudaMemcpyAsync(devs[g].invD + (size_t)r0 * DB * DB, devs[o].invD + (size_t)r0 * DB * DB, (size_t)nq * DB * DB * esz,
                            cudaMemcpyDefault, devs[g].copy));
      have[g] = record(g, devs[g].copy);
    }
    if (below <= 0) break;
    // ---- trailing updates with panel q on every device, for the panels it owns
    for (int g = 0; g < G; ++g) {
      check(cudaSetDevice(devs[g].dev));
      bool any_upd = false;
      for (int q2 = q + 1; q2 < Q; ++q2) {
        if (q2 % G != g) continue;
        const int c0 = q2 * PB;
        const int n2 = (nblocks - c0 < PB) ? nblocks - c0 : PB;
        const bool lookahead = (q2 == q + 1);
        cudaStream_t s = lookahead ? devs[g].main : devs[g].upd;
        if (lookahead) {
          check(cudaStreamWaitEvent(s, have[g], 0));
          if (upd_done[g]) check(cudaStreamWaitEvent(s, upd_done[g], 0));
        } else if (!any_upd) {
          check(cudaStreamWaitEvent(s, have[g], 0));
          any_upd = true;
        }
        ch[g].st = s;  // Chol::update launches on the stream it is given; keep st consistent for error paths
        ch[g].update(s, !lookahead, c0, c0, nblocks - c0, n2, c0, r0, c0, r0, nq, false);
        ch[g].st = devs[g].main;
        check(ch[g].err);
      }
      if (any_upd) upd_done[g] = record(g, devs[g].upd);
    }
  }
  // join: every stream of every device has finished before the caller continues on the main streams
  for (int g = 0; g < G; ++g) {
    check(cudaSetDevice(devs[g].dev));
    if (upd_done[g]) check(cudaStreamWaitEvent(devs[g].main, upd_done[g], 0));
    if (have[g]) check(cudaStreamWaitEvent(devs[g].main, have[g], 0));
  }
  for (int g = 0; g < G; ++g) {
    check(cudaSetDevice(devs[g].dev));
    check(cudaStreamSynchronize(devs[g].main));
    check(cudaStreamSynchronize(devs[g].copy));
    check(cudaStreamSynchronize(devs[g].upd));
  }
  for (cudaEvent_t ev : evs) cudaEventDestroy(ev);
  return err;
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
