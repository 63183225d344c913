// Blocked Cholesky on the device: replaces cholesky(Symmetric(.)).L and the
// triangular solves of lusim.jl:92,98-103.  Recursive right-looking formulation over 128x128
// tiles: every flop outside the 128-wide diagonal blocks runs in the DMMA tile GEMM (gemm.cuh);
// diagonal blocks are factorised AND inverted by one CTA (potrf_diag_kernel) so that every
// triangular solve becomes a tile GEMM with the explicit inverse.
#include "chol.h"

#include <algorithm>
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
GSP_DEV void potrf_diag_body(double* A, long long lda, long long blk, double* invD, int* info, unsigned char* smem) {
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
  }
}

__global__ void __launch_bounds__(256, 1) potrf_diag_kernel(double* __restrict__ A, long long lda, long long blk,
                                                            double* __restrict__ invD, int* __restrict__ info) {
  GSP_DYN_SMEM(smem);
  potrf_diag_body(A, lda, blk, invD, info, smem);
}

// ------------------------------------------------------------------ fused diagonal square (the panel head of chol_factor_dist)
// Cholesky of the nq x nq-block square at block (blk0, blk0) in ONE launch: 2 CTAs per block row (64-row strips), each walks the
// columns of its row - S: X_ij = A_ij inv(L_jj)^T, U: A_il -= X_ij X_lj^T for j < l <= i - and the upper CTA of a row then factors
// and inverts its diagonal block (potrf_diag_body).  CTAs synchronise through release / acquire flags in global memory (the whole
// square is 2 MB and L2-resident): P[j] = diagonal block j done, S[i][j] = strips of X_ij done, H[i] = lower strip of row i final.
// 12 dependent launches (4 diagonal blocks, 4 solves, 4 updates at PB = 4, 30-40 us each next to the wide updates) become one
// kernel whose critical path is the four diagonal blocks plus six 10 us strip products.
constexpr int SQ_MAXB = 8;                                   // blocks per square side
constexpr int SQ_LDA = 64 + 4, SQ_LDB = DB + 4;              // strip and operand tiles in shared memory, k-major: T[k * LD + r]
constexpr int SQ_SMEM = (DB * SQ_LDA + DB * SQ_LDB) * 8;     // 204.8 KB (potrf_diag_body overlays it with its 155 KB)
constexpr int SQ_FLAGS = SQ_MAXB + SQ_MAXB * SQ_MAXB + SQ_MAXB;

GSP_DEV void sq_wait(const int* f, int v) {
  if (threadIdx.x == 0)
    while (ld_acquire_gpu(f) < v) spin_pause();
  __syncthreads();
}
GSP_DEV void sq_signal(int* f) {
  __syncthreads();
  if (threadIdx.x == 0) {
#ifndef GSP_EMU
    __threadfence();
#endif
    red_release_gpu_add(f, 1);
  }
}

// C(64 x 128) = Aop(64 x 128) * B(128 x 128)^T  (SUB: C -= ...), column-major operands in global memory, one CTA of 8 warps;
// warp w owns rows 8w..8w+7.  C may alias Aop (the strip is staged in shared memory before anything is stored).
template <bool SUB>
GSP_DEV void sq_strip_gemm(double* C, long long ldc, const double* Aop, long long lda, const double* B, long long ldb, unsigned char* smem) {
  double* As = reinterpret_cast<double*>(smem);
  double* Bs = As + DB * SQ_LDA;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lk = lane & 3;
  __syncthreads();  // the previous task's readers of the tiles are done
  for (int base = 0; base < 64 * DB; base += 8 * 256) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = base + u * 256 + tid;
      v[u] = Aop[(idx & 63) + (long long)(idx >> 6) * lda];
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = base + u * 256 + tid;
      As[(idx >> 6) * SQ_LDA + (idx & 63)] = v[u];
    }
  }
  for (int base = 0; base < DB * DB; base += 16 * 256) {
    double v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int idx = base + u * 256 + tid;
      v[u] = B[(idx & (DB - 1)) + (long long)(idx >> 7) * ldb];
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int idx = base + u * 256 + tid;
      Bs[(idx >> 7) * SQ_LDB + (idx & (DB - 1))] = v[u];
    }
  }
  double acc[16][2];
  double* crow = C + 8 * warp + lr + (long long)(2 * lk) * ldc;
#pragma unroll
  for (int nb = 0; nb < 16; ++nb) {
    acc[nb][0] = SUB ? crow[(long long)(nb * 8) * ldc] : 0.0;
    acc[nb][1] = SUB ? crow[(long long)(nb * 8 + 1) * ldc] : 0.0;
  }
  __syncthreads();
  for (int k0 = 0; k0 < DB; k0 += 4) {
    const double av = As[(k0 + lk) * SQ_LDA + 8 * warp + lr];
    const double a = SUB ? -av : av;
    const double* bp = Bs + (k0 + lk) * SQ_LDB + lr;
#pragma unroll
    for (int nb = 0; nb < 16; ++nb) dmma884(acc[nb][0], acc[nb][1], a, bp[nb * 8]);
  }
#pragma unroll
  for (int nb = 0; nb < 16; ++nb) {
    crow[(long long)(nb * 8) * ldc] = acc[nb][0];
    crow[(long long)(nb * 8 + 1) * ldc] = acc[nb][1];
  }
}

__global__ void __launch_bounds__(256, 1) potrf_square_kernel(double* A, long long lda, long long blk0, int nq, double* invD, int* info,
                                                              int* flags) {
  GSP_DYN_SMEM(smem);
  const int i = blockIdx.x >> 1, h = blockIdx.x & 1;
  int* Pf = flags;
  int* Sf = flags + SQ_MAXB;
  int* Hf = flags + SQ_MAXB + SQ_MAXB * SQ_MAXB;
  const long long rbase = (blk0 + i) * DB + 64 * h;
  for (int j = 0; j < i; ++j) {
    double* Xij = A + rbase + (blk0 + j) * DB * lda;
    sq_wait(&Pf[j], 1);
    sq_strip_gemm<false>(Xij, lda, Xij, lda, invD + (blk0 + j) * DB * DB, DB, smem);
    sq_signal(&Sf[i * SQ_MAXB + j]);
    for (int l = j + 1; l <= i; ++l) {
      sq_wait(&Sf[l * SQ_MAXB + j], 2);
      sq_strip_gemm<true>(A + rbase + (blk0 + l) * DB * lda, lda, Xij, lda, A + (blk0 + l) * DB + (blk0 + j) * DB * lda, lda, smem);
    }
  }
  if (h == 1) {
    sq_signal(&Hf[i]);
    return;
  }
  sq_wait(&Hf[i], 1);
  potrf_diag_body(A, lda, blk0 + i, invD, info, smem);
  sq_signal(&Pf[i]);
}

// Distributed factorization: the lower blocks of a finished square and the inverses of its diagonal blocks go to the other devices.
// One CTA per 16 columns of a block (a single CTA storing 1.8 MB over NVLink from the factorizing kernel itself cost 30 us per
// diagonal block on the critical path).
__global__ void __launch_bounds__(256) push_square_kernel(const double* __restrict__ A, long long lda, long long blk0, int nq,
                                                          const double* __restrict__ invD, const GSP_GRID_CONSTANT DiagPeers peers) {
  const int item = blockIdx.x / 8, part = blockIdx.x % 8, tid = threadIdx.x;
  long long off0, ldx;
  const double* src;
  bool inv = item < nq;
  if (inv) {
    off0 = (blk0 + item) * DB * DB;
    ldx = DB;
    src = invD;
  } else {
    int t = item - nq, bi = 0;
    while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
    const int bj = t - bi * (bi + 1) / 2;
    off0 = (blk0 + bi) * DB + (blk0 + bj) * DB * lda;
    ldx = lda;
    src = A;
  }
  for (int idx = tid; idx < 16 * 64; idx += 256) {
    const int c = part * 16 + (idx >> 6), r2 = (idx & 63) * 2;
    const long long off = off0 + r2 + (long long)c * ldx;
    const double2 v = *reinterpret_cast<const double2*>(src + off);
    for (int p = 0; p < peers.n; ++p) *reinterpret_cast<double2*>((inv ? peers.invD[p] : peers.A[p]) + off) = v;
  }
}

// blocks (r0 + i, c0 + j), i < nr, of the matrix -> the same position on the other devices; one CTA per 16 columns of a block
__global__ void __launch_bounds__(256) push_rect_kernel(const double* __restrict__ A, long long lda, long long r0, int nr, long long c0,
                                                        const GSP_GRID_CONSTANT DiagPeers peers) {
  const int item = blockIdx.x / 8, part = blockIdx.x % 8, tid = threadIdx.x;
  const long long off0 = (r0 + item % nr) * DB + (c0 + item / nr) * DB * lda;
  for (int idx = tid; idx < 16 * 64; idx += 256) {
    const int c = part * 16 + (idx >> 6), r2 = (idx & 63) * 2;
    const long long off = off0 + r2 + (long long)c * lda;
    const double2 v = *reinterpret_cast<const double2*>(A + off);
    for (int p = 0; p < peers.n; ++p) *reinterpret_cast<double2*>(peers.A[p] + off) = v;
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
  cudaStream_t* side;    // look-ahead streams, one per recursion depth (recursive single-device algorithm)
  int nside;
  double* A;
  long long ld;
  double* invD;
  int* info;
  cudaError_t err = cudaSuccess;
  std::vector<cudaEvent_t> events;
  int side_ctas = 0;  // > 0: look-ahead GEMMs run as persistent grids of this many CTAs, the other SMs stay free for the main stream
  DiagPeers peers{};  // distributed factorization: the other devices' matrices / inverse arrays (final blocks are multicast there)

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
  void set_peers(GemmArgs& g, long long off) const {
    g.npeer = peers.n;
    for (int p = 0; p < peers.n; ++p) g.Cpeer[p] = peers.A[p] + off;
  }

  // C[mt x nt blocks at (cr, cc)] -= A[(ar, ac), K blocks] * B[(br, bc), K blocks]^T   (lower tiles only if tri)
  void update(cudaStream_t s, bool side_grid, int cr, int cc, int mt, int nt, int ar, int ac, int br, int bc, int kb, bool tri) {
    if (mt <= 0 || nt <= 0 || kb <= 0) return;
    GemmArgs g{};
    g.A = at(ar, ac); g.lda = ld;
    g.B = at(br, bc); g.ldb = ld;
    g.C = at(cr, cc); g.ldc = ld;
    g.mt = mt; g.nt = nt; g.K = kb * DB; g.tri = tri ? 1 : 0;
    g.max_ctas = side_grid ? side_ctas : 0;
    GSP_DEP_ACCESS(A, ar, ar + mt, ac, ac + kb, false);
    GSP_DEP_ACCESS(A, br, br + nt, bc, bc + kb, false);
    GSP_DEP_ACCESS(A, cr, cr + mt, cc, cc + nt, true);
    check(launch_gemm<GEMM_SUB, false>(s, g));
  }

  // X * L[c0:c0+nc, c0:c0+nc]^T = A[r0:r0+nr, c0:c0+nc]   (in place; the leaves multiply by the explicit inverses of the diagonal blocks)
  void trsm(int r0, int nr, int c0, int nc) {
    if (nr <= 0 || nc <= 0) return;
    if (nc == 1) {
      GemmArgs g{};
      g.A = at(r0, c0); g.lda = ld;
      g.B = invD + (long long)c0 * DB * DB; g.ldb = DB;
      g.C = at(r0, c0); g.ldc = ld;
      g.mt = nr; g.nt = 1; g.K = DB;
      set_peers(g, (long long)r0 * DB + (long long)c0 * DB * ld);
      GSP_DEP_ACCESS(invD, c0, c0 + 1, 0, 1, false);
      GSP_DEP_ACCESS(A, r0, r0 + nr, c0, c0 + 1, true);
      for (int p = 0; p < peers.n; ++p) GSP_DEP_ACCESS(peers.A[p], r0, r0 + nr, c0, c0 + 1, true);
      check(launch_gemm<GEMM_SET, false>(st, g));
      return;
    }
    const int c1 = nc / 2;
    trsm(r0, nr, c0, c1);
    update(st, false, r0, c0 + c1, nr, nc - c1, r0, c0, c0 + c1, c0, c1, false);
    trsm(r0, nr, c0 + c1, nc - c1);
  }

  // ---- the same two operations on a LIST of block rows (distributed factorization: the rows this device owns below a panel).
  // rows_dev: device array of nrows ascending block-row indices.
  // C[rows, ccol0 + j] -= A[rows, kcol0 : kcol0 + kb) * L[ccol0 + j, kcol0 : kcol0 + kb)^T,  j < ncolblk; stair: only j with ccol0 + j <= row
  void update_rows(cudaStream_t s, const int* rows_dev, int nrows, long long valid_tiles, int ccol0, int ncolblk, int kcol0, int kb, bool stair) {
    if (nrows <= 0 || ncolblk <= 0 || kb <= 0 || valid_tiles <= 0) return;
    GemmArgs g{};
    g.A = A + (long long)kcol0 * DB * ld; g.lda = ld;
    g.B = at(ccol0, kcol0); g.ldb = ld;
    g.C = A + (long long)ccol0 * DB * ld; g.ldc = ld;
    g.mt = nrows; g.nt = ncolblk; g.K = kb * DB;
    g.rows = rows_dev; g.stair = stair ? 1 : 0; g.colblk0 = ccol0;
#ifdef GSP_EMU
    for (int i = 0; i < nrows; ++i) {  // (device memory is host memory in the emulator)
      const int r = rows_dev[i], c1 = stair ? std::min(ccol0 + ncolblk, r + 1) : ccol0 + ncolblk;
      GSP_DEP_ACCESS(A, r, r + 1, kcol0, kcol0 + kb, false);
      GSP_DEP_ACCESS(A, r, r + 1, ccol0, c1, true);
    }
    GSP_DEP_ACCESS(A, ccol0, ccol0 + ncolblk, kcol0, kcol0 + kb, false);
#endif
    check(launch_gemm<GEMM_SUB, false>(s, g, valid_tiles));
  }
  void trsm_rows(cudaStream_t s, const int* rows_dev, int nrows, int c0, int nc, bool multicast = true) {
    if (nrows <= 0 || nc <= 0) return;
    if (nc == 1) {
      GemmArgs g{};
      g.A = A + (long long)c0 * DB * ld; g.lda = ld;
      g.B = invD + (long long)c0 * DB * DB; g.ldb = DB;
      g.C = A + (long long)c0 * DB * ld; g.ldc = ld;
      g.mt = nrows; g.nt = 1; g.K = DB;
      g.rows = rows_dev;
      if (multicast) set_peers(g, (long long)c0 * DB * ld);
#ifdef GSP_EMU
      GSP_DEP_ACCESS(invD, c0, c0 + 1, 0, 1, false);
      for (int i = 0; i < nrows; ++i) {
        GSP_DEP_ACCESS(A, rows_dev[i], rows_dev[i] + 1, c0, c0 + 1, true);
        for (int p = 0; p < g.npeer; ++p) GSP_DEP_ACCESS(peers.A[p], rows_dev[i], rows_dev[i] + 1, c0, c0 + 1, true);
      }
#endif
      check(launch_gemm<GEMM_SET, false>(s, g));
      return;
    }
    const int c1 = nc / 2;
    trsm_rows(s, rows_dev, nrows, c0, c1, multicast);
    update_rows(s, rows_dev, nrows, (long long)nrows * (nc - c1), c0 + c1, nc - c1, c0, c1, false);
    trsm_rows(s, rows_dev, nrows, c0 + c1, nc - c1, multicast);
  }
  // blocks [r0, r0 + nr) x [c0, c0 + nc) of the matrix to the other devices (many CTAs: one SM's peer stores are slow)
  void push_rect(cudaStream_t s, int r0, int nr, int c0, int nc) {
    if (peers.n <= 0 || nr <= 0 || nc <= 0) return;
    GSP_DEP_ACCESS(A, r0, r0 + nr, c0, c0 + nc, false);
    for (int p = 0; p < peers.n; ++p) GSP_DEP_ACCESS(peers.A[p], r0, r0 + nr, c0, c0 + nc, true);
    ProfScope prof_("push_rect", s);
    GSP_LAUNCH(push_rect_kernel, dim3((unsigned)(8 * nr * nc)), dim3(256), 0, s, (const double*)A, ld, (long long)r0, nr, (long long)c0, peers);
    g_launches++;
    check(cudaGetLastError());
  }

  // the whole diagonal square [o, o + n) in one launch (potrf_square_kernel), then its blocks and inverses to the other devices
  void square(int o, int n, int* flags) {
    if (n <= 0) return;
    if (n == 1) {
      potrf(o, 1, nullptr, 0);
    } else {
      check(cudaMemsetAsync(flags, 0, SQ_FLAGS * sizeof(int), st));
      auto kfn = potrf_square_kernel;
      check(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, SQ_SMEM));
      GSP_DEP_ACCESS(A, o, o + n, o, o + n, true);
      GSP_DEP_ACCESS(invD, o, o + n, 0, 1, true);
      ProfScope prof_("potrf_square", st);
      GSP_LAUNCH_COOP(kfn, dim3((unsigned)(2 * n)), dim3(256), (size_t)SQ_SMEM, st, A, ld, (long long)o, n, invD, info, flags);
      g_launches++;
      check(cudaGetLastError());
    }
    if (peers.n > 0) {
      GSP_DEP_ACCESS(A, o, o + n, o, o + n, false);
      GSP_DEP_ACCESS(invD, o, o + n, 0, 1, false);
      for (int p = 0; p < peers.n; ++p) {
        GSP_DEP_ACCESS(peers.A[p], o, o + n, o, o + n, true);
        GSP_DEP_ACCESS(peers.invD[p], o, o + n, 0, 1, true);
      }
      ProfScope prof_("push_square", st);
      GSP_LAUNCH(push_square_kernel, dim3((unsigned)(8 * (n + n * (n + 1) / 2))), dim3(256), 0, st, (const double*)A, ld, (long long)o, n,
                 (const double*)invD, peers);
      g_launches++;
      check(cudaGetLastError());
    }
  }

  // Cholesky of the n diagonal blocks starting at o.  `pend`: event after which the second half [o + n/2, o + n) of the
  // block rows/cols is up to date (nullptr: already valid on the main stream).  The first half is always valid on entry.
  void potrf(int o, int n, cudaEvent_t pend, int depth) {
    if (n <= 0) return;
    if (n == 1) {
      auto kfn = potrf_diag_kernel;
      check(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM));
      GSP_DEP_ACCESS(A, o, o + 1, o, o + 1, true);
      GSP_DEP_ACCESS(invD, o, o + 1, 0, 1, true);
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

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

}  // namespace

cudaError_t chol_factor(cudaStream_t st, cudaStream_t* side, int nside, double* A, long long ld, int nblocks, double* invD, int* info) {
  Chol c{st, side, nside, A, ld, invD, info};
  cudaError_t e = cudaMemsetAsync(info, 0, sizeof(int), st);
  if (e != cudaSuccess) return e;
  static const int lookahead = env_int("GSP_CHOL_LOOKAHEAD", 1);  // 0 disables the side streams (A/B measurements)
  if (!lookahead) c.nside = 0;
  // GSP_CHOL_RESERVE=r: SMs kept free of look-ahead work.  A one-CTA-per-tile look-ahead GEMM fills every SM for the length of
  // a tile (1.2 ms at K = 8192), and the short kernels of the critical path queue behind it whatever their stream priority.
  static const int reserve = std::max(0, env_int("GSP_CHOL_RESERVE", GSP_CHOL_RESERVE_DEFAULT));
  if (reserve > 0) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    c.side_ctas = sms - reserve > 8 ? sms - reserve : 8;
  }
  c.potrf(0, nblocks, nullptr, 0);
  for (cudaEvent_t ev : c.events) cudaEventDestroy(ev);
  return c.err;
}

// ------------------------------------------------------------------ distributed (and single-device panel) factorization
// Right-looking over column panels of PB 128-blocks with ROW-PANEL ownership: block rows [p PB, (p+1) PB) belong to device
// owner(p) (boustrophedon over the G devices: 0..G-1, G-1..0, ... - the work of a row grows with its index, the snake keeps every
// device's share within a panel of the mean).  Every device holds a full-size buffer but assembles and updates ONLY its own rows;
// finished blocks are written straight into the same position of every other device's buffer from the epilogue of the kernel
// that produces them (peer-mapped stores over NVLink: potrf_diag_kernel and the TRSM leaf GEMM), so when the last panel is done
// every device holds the whole factor L - exactly what the realization-sharded sampling needs - without a separate broadcast.
// Step q (column panel q, owner o = owner(q), next owner o1 = owner(q+1)):
//   D(q)      on o:   Cholesky of the PB x PB diagonal square (needs only o's own rows); blocks + inverses go to every device [main]
//   T_first   on o1:  o1 solves its rows of panel q+1 against the square, then LA_sq: the square of panel q+1 updated with
//                     them, then D(q+1) - this is the whole critical path, and it stays on the main streams of two devices   [main]
//   T(q)      on all: each device solves ITS other rows below the square and multicasts them from the GEMM epilogue         [aux]
//   LA(q)     on all: own rows x column panel q+1 updated with panel q                                                        [aux]
//   near(q), far(q) on all: own rows x columns of panel q+2 / of the panels >= q+3 updated with panel q  [low-priority stream]
// LA(q+1) waits for near(q) only, which is queued before far(q): a look-ahead of depth two - the critical path never waits for
// the wide update of the current or the previous step.  The panel solve and the updates are split G ways by construction; the
// wide updates are launches of short (K = PB*128) tiles that the high-priority kernels pre-empt at tile granularity.
int chol_dist_flag_ints() { return SQ_FLAGS; }

static int dist_owner(int p, int G) {
  const int m = p % (2 * G);
  return m < G ? m : 2 * G - 1 - m;
}

void chol_dist_owned_rows(int nblocks, int PB, int G, int g, std::vector<int>* rows) {
  rows->clear();
  const int Q = (nblocks + PB - 1) / PB;
  for (int p = 0; p < Q; ++p)
    if (dist_owner(p, G) == g)
      for (int b = p * PB; b < std::min(nblocks, (p + 1) * PB); ++b) rows->push_back(b);
}

cudaError_t chol_factor_dist(const std::vector<DistDev>& devs, long long ld, int nblocks, int PB) {
  const int G = (int)devs.size();
  const int Q = (nblocks + PB - 1) / PB;
  cudaError_t err = cudaSuccess;
  auto check = [&](cudaError_t e) {
    if (err == cudaSuccess && e != cudaSuccess) err = e;
  };
  std::vector<Chol> ch;
  ch.reserve(G);
  std::vector<std::vector<int>> rows(G);
  for (int g = 0; g < G; ++g) {
    check(cudaSetDevice(devs[g].dev));
    ch.push_back(Chol{devs[g].main, nullptr, 0, devs[g].A, ld, devs[g].invD, devs[g].info});
    for (int h = 0; h < G; ++h)
      if (h != g) {
        ch[g].peers.A[ch[g].peers.n] = devs[h].A;
        ch[g].peers.invD[ch[g].peers.n] = devs[h].invD;
        ch[g].peers.n++;
      }
    check(cudaMemsetAsync(devs[g].info, 0, sizeof(int), devs[g].main));
    chol_dist_owned_rows(nblocks, PB, G, g, &rows[g]);
    if (!rows[g].empty())
      check(cudaMemcpyAsync(devs[g].rows, rows[g].data(), rows[g].size() * sizeof(int), cudaMemcpyHostToDevice, devs[g].main));
  }
  if (g_prof.on) {
    std::vector<int> ids;
    for (auto& d : devs) ids.push_back(d.dev);
    g_prof.mark_reference(ids);
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
  // index of the first own row >= block b, per device
  auto first_at = [&](int g, int b) { return (int)(std::lower_bound(rows[g].begin(), rows[g].end(), b) - rows[g].begin()); };
  // tiles (own row r, column block c) with c0 <= c < c1 and c <= r
  auto stair_tiles = [&](int g, int i0, int c0, int c1) {
    long long t = 0;
    for (size_t i = (size_t)i0; i < rows[g].size(); ++i) t += std::max(0, std::min(c1, rows[g][i] + 1) - c0);
    return t;
  };
  // the callers' earlier work on the main streams (assembly of the own rows) precedes everything on the other two streams
  for (int g = 0; g < G; ++g) {
    cudaEvent_t e = record(g, devs[g].main);
#ifdef GSP_EMU
    if (getenv("GSP_DEPCHECK_SELFTEST")) continue;  // test-only: drop this dependency - the checker must then report the race
#endif
    check(cudaStreamWaitEvent(devs[g].aux, e, 0));
    check(cudaStreamWaitEvent(devs[g].upd, e, 0));
  }
  // events of the current step (per device): rows solved (T on the aux stream; the next owner's leading rows on its main stream),
  // look-ahead update done, near / far parts of the wide update done
  std::vector<cudaEvent_t> evT(G, nullptr), evLA(G, nullptr), evNear(G, nullptr);
  cudaEvent_t evTfirst = nullptr, evD = nullptr;
  for (int q = 0; q < Q && err == cudaSuccess; ++q) {
    const int o = dist_owner(q, G);
    const int c0 = q * PB;
    const int nq = std::min(PB, nblocks - c0);
    const int cn = c0 + nq;  // first block below / right of the panel
    // ---- D(q): the diagonal square on its owner (up to date: LA_sq(q-1) ran on this stream), then its diagonal blocks and inverses
    // go to the other devices
    check(cudaSetDevice(devs[o].dev));
    static const int fused = env_int("GSP_CHOL_FUSED_SQUARE", 1);  // 0: the recursion of separate launches (A/B)
    if (fused && nq <= SQ_MAXB) {
      ch[o].square(c0, nq, devs[o].flags);
    } else {
      ch[o].potrf(c0, nq, nullptr, 0);
      if (ch[o].peers.n > 0) {
        GSP_DEP_ACCESS(devs[o].A, c0, c0 + nq, c0, c0 + nq, false);
        GSP_DEP_ACCESS(devs[o].invD, c0, c0 + nq, 0, 1, false);
        for (int p = 0; p < ch[o].peers.n; ++p) {
          GSP_DEP_ACCESS(ch[o].peers.A[p], c0, c0 + nq, c0, c0 + nq, true);
          GSP_DEP_ACCESS(ch[o].peers.invD[p], c0, c0 + nq, 0, 1, true);
        }
        ProfScope prof_("push_square", devs[o].main);
        GSP_LAUNCH(push_square_kernel, dim3((unsigned)(8 * (nq + nq * (nq + 1) / 2))), dim3(256), 0, devs[o].main, (const double*)devs[o].A, ld,
                   (long long)c0, nq, (const double*)devs[o].invD, ch[o].peers);
        g_launches++;
        check(cudaGetLastError());
      }
    }
    check(ch[o].err);
    evD = (G > 1) ? record(o, devs[o].main) : nullptr;
    if (cn >= nblocks) break;
    const int o1 = dist_owner(q + 1, G);
    const int n1 = std::min(PB, nblocks - cn);      // width of panel q+1
    const int cf = std::min(nblocks, cn + n1 + PB); // first block column of the far update (panel q+3 on)
    // ---- T(q): every device solves its own rows below the square and multicasts them.  The rows of panel q+1 (on o1) are on the
    // critical path - T_first, LA_sq, D(q+1) on o1's main stream -, everything else runs on the aux streams.
    evTfirst = nullptr;
    for (int g = 0; g < G; ++g) {
      const int i0 = first_at(g, cn);
      int ia = i0;
      check(cudaSetDevice(devs[g].dev));
      cudaStream_t sm = devs[g].main, sa = devs[g].aux;
      if (g == o1) {
        ia = first_at(g, cn + n1);
        if (g != o && evD) check(cudaStreamWaitEvent(sm, evD, 0));
        if (evLA[g]) check(cudaStreamWaitEvent(sm, evLA[g], 0));   // LA(q-1) updated these rows in column panel q (aux stream)
        // solved locally on the main stream (LA_sq and D(q+1) only need them here); the copy for the other devices' look-ahead
        // updates leaves from the aux stream, off the critical path
        ch[g].trsm_rows(sm, devs[g].rows + i0, ia - i0, c0, nq, false);
        cudaEvent_t solved = record(g, sm);
        check(cudaStreamWaitEvent(sa, solved, 0));
        ch[g].push_rect(sa, cn, n1, c0, nq);
        evTfirst = record(g, sa);
      }
      const int m = (int)rows[g].size() - ia;
      if (m > 0) {
        if (evD) check(cudaStreamWaitEvent(sa, evD, 0));  // (same device: orders the aux stream after D on the main stream)
        else if (g == o) check(cudaStreamWaitEvent(sa, record(g, sm), 0));
        ch[g].trsm_rows(sa, devs[g].rows + ia, m, c0, nq);
      }
      evT[g] = record(g, sa);
      check(ch[g].err);
    }
    // ---- LA(q): own rows x column panel q+1 with panel q;  near(q) / far(q): the columns of panel q+2 / of the panels from q+3 on
    for (int g = 0; g < G; ++g) {
      check(cudaSetDevice(devs[g].dev));
      const int i0 = first_at(g, cn);
      if ((int)rows[g].size() - i0 <= 0) continue;
      cudaStream_t sm = devs[g].main, sa = devs[g].aux, su = devs[g].upd;
      int ia = i0;
      if (g == o1) {
        // the square of panel q+1 first, on the main stream: D(q+1) follows immediately.  Column panel q+1 was last touched by near(q-1).
        ia = first_at(g, cn + n1);
        if (evNear[g]) check(cudaStreamWaitEvent(sm, evNear[g], 0));
        ch[g].update_rows(sm, devs[g].rows + i0, ia - i0, stair_tiles(g, i0, cn, cn + n1) - stair_tiles(g, ia, cn, cn + n1), cn, n1, c0, nq, true);
      }
      if ((int)rows[g].size() - ia > 0) {
        // needs o1's rows of panel q (T_first) besides the own rows (aux stream order)
        if (evTfirst) check(cudaStreamWaitEvent(sa, evTfirst, 0));
        if (evNear[g]) check(cudaStreamWaitEvent(sa, evNear[g], 0));
        ch[g].update_rows(sa, devs[g].rows + ia, (int)rows[g].size() - ia, stair_tiles(g, ia, cn, cn + n1), cn, n1, c0, nq, true);
        evLA[g] = record(g, sa);
      } else {
        evLA[g] = nullptr;
      }
      if (cn + n1 < nblocks) {
        // near: besides the own rows it needs the rows of panel q+2 in panel q (solved on owner(q+2)'s aux stream)
        check(cudaStreamWaitEvent(su, evT[g], 0));
        if (q + 2 < Q) check(cudaStreamWaitEvent(su, evT[dist_owner(q + 2, G)], 0));
        const int ib = first_at(g, cn + n1);
        ch[g].update_rows(su, devs[g].rows + ib, (int)rows[g].size() - ib, stair_tiles(g, ib, cn + n1, cf), cn + n1, cf - cn - n1, c0, nq, true);
        evNear[g] = record(g, su);
        if (cf < nblocks) {
          // far: needs every device's rows of panel q
          for (int h = 0; h < G; ++h) check(cudaStreamWaitEvent(su, evT[h], 0));
          const int ic = first_at(g, cf);
          ch[g].update_rows(su, devs[g].rows + ic, (int)rows[g].size() - ic, stair_tiles(g, ic, cf, nblocks), cf, nblocks - cf, c0, nq, true);
        }
      } else {
        evNear[g] = nullptr;
      }
      check(ch[g].err);
    }
  }
  // join: every stream of every device has finished (and every peer store has landed) before the caller continues
  for (int g = 0; g < G; ++g) {
    check(cudaSetDevice(devs[g].dev));
    check(cudaStreamSynchronize(devs[g].main));
    check(cudaStreamSynchronize(devs[g].aux));
    check(cudaStreamSynchronize(devs[g].upd));
  }
  for (cudaEvent_t ev : evs) cudaEventDestroy(ev);
  for (auto& c : ch)
    for (cudaEvent_t ev : c.events) cudaEventDestroy(ev);
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
