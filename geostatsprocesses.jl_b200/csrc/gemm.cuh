// FP64 tensor-core (DMMA) tile GEMM used by the blocked Cholesky (TRSM-by-inverse, SYRK/GEMM
// trailing updates: lusim.jl:98-103) and by the realization product L22 * W (lusim.jl:162,164).
//
// One CTA = one 128x128 tile of C.  Warp-specialised: warp 8 is the producer, it streams 16-deep
// K-chunks of A and B into a 4-stage shared-memory ring with bulk async copies (TMA engine,
// cp.async.bulk -> SASS UBLKCP) completing on mbarriers; warps 0-7 are consumers, each owning a
// 64x32 sub-tile = 8x4 DMMA.8x8x4 accumulators in registers.  Shared-memory rows are padded by 4
// doubles (ld = 4 mod 16) which makes every 8-byte fragment load bank-conflict-free.
// All matrices are column-major with extents padded to multiples of 128 (K to 16) by the callers.
#pragma once
#include <cstdlib>

#include "common.h"

namespace gsp {

constexpr int GT = 128;        // block granularity of all callers (matrices are padded to multiples of 128)
constexpr int GKC = 16;        // K chunk per stage
constexpr int GSTAGES = 4;
constexpr int GLDBK = GKC + 4; // smem leading dim of [n][k] operand tiles (20 = 4 mod 16)
#define GSP_MAX_PEERS 7        // one context drives at most 8 devices

enum GemmMode { GEMM_SET = 0, GEMM_SUB = 1, GEMM_SAMPLE = 2 };

// tile shapes: TM x TN per CTA, WM x WN consumer warps, each warp owns (TM/WM) x (TN/WN)
template <int TM, int TN, int WM, int WN>
struct GemmCfg {
  static constexpr int LDA = TM + 4;  // = 4 mod 16: conflict-free 8-byte fragment loads
  static constexpr int LDB = TN + 4;
  static constexpr int A_BYTES = GKC * LDA * 8;
  static constexpr int B_BYTES = (TN * GLDBK * 8 > GKC * LDB * 8) ? TN * GLDBK * 8 : GKC * LDB * 8;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SMEM_BYTES = GSTAGES * STAGE_BYTES + 2 * GSTAGES * 16 + 128;
  static constexpr int CWARPS = WM * WN;
  static constexpr int THREADS = (CWARPS + 1) * 32;  // + 1 producer warp
  static constexpr int FM = TM / WM / 8, FN = TN / WN / 8;  // DMMA fragments per warp
};

struct GemmArgs {
  const double* A;   // M x K, A[i + k*lda]
  long long lda;
  const double* B;   // !BK: N x K, B[j + k*ldb] (C = A * B^T);  BK: K x N, B[k + j*ldb] (C = A * B)
  long long ldb;
  double* C;         // M x N, C[i + j*ldc]
  long long ldc;
  int mt, nt;        // tiles along M and N (in units of TM / TN)
  int K;             // multiple of GKC
  int tri;           // 1: only tiles with ti >= tj (TM == TN)
  int klimit;        // 1: A is lower triangular -> k < (ti+1)*TM only
  // GEMM_SAMPLE epilogue: Z[sinds[i] + r*ldz] = acc + d2[i] + addmu for i < Ns, r < R
  const double* d2;
  const long long* sinds;
  double addmu;
  long long Ns, R;
  long long ntiles;  // filled by launch_gemm_cfg
  int max_ctas;      // > 0: persistent grid of at most this many CTAs walking the tiles (look-ahead streams leave SMs free)
  // ---- distributed factorization (chol.cu, chol_factor_dist): a device works on the block rows it owns
  const int* rows;   // device array or nullptr: block row (128-row units) of every group of 128 rows of A and C (tile row ti covers
                     // rows rows[ti*TM/128]*128 + (ti*TM)%128 ...); nullptr: rows are contiguous from the base pointers
  int stair;         // 1: skip tiles above the diagonal of the GLOBAL matrix: keep tile (ti, tj) iff colblk0*128 + tj*TN <= first row of ti
  int colblk0;       // global block column of tile column 0 of C (stair only)
  int npeer;         // GEMM_SET: the finished tile is ALSO stored at the same offsets of Cpeer[0..npeer) - the matrices of the other
  double* Cpeer[GSP_MAX_PEERS];  // devices (peer-mapped over NVLink): the panel multicast is fused into the TRSM epilogue
};

template <int MODE, bool BK, int TM, int TN, int WM, int WN>
__global__ void __launch_bounds__(GemmCfg<TM, TN, WM, WN>::THREADS, 1) gemm_dmma_kernel(GemmArgs g) {
  using C_ = GemmCfg<TM, TN, WM, WN>;
  constexpr int LDA = C_::LDA, LDB = C_::LDB, FM = C_::FM, FN = C_::FN, CW = C_::CWARPS;
  GSP_DYN_SMEM(smem);
  unsigned char* stage_base = smem;
  mbar_t* full = reinterpret_cast<mbar_t*>(smem + GSTAGES * C_::STAGE_BYTES);
  mbar_t* empty = full + GSTAGES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < GSTAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], CW);
    }
    fence_mbar_init();
  }
  __syncthreads();

  // tile t -> (ti, tj, chunks of K)
  auto decode = [&](long long t, int& ti, int& tj, int& nchunks) {
    if (g.tri) {
      int i = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
      while ((long long)(i + 1) * (i + 2) / 2 <= t) ++i;
      while ((long long)i * (i + 1) / 2 > t) --i;
      ti = i;
      tj = (int)(t - (long long)i * (i + 1) / 2);
    } else if (g.klimit) {
      // lower-triangular A: row tile ti costs (ti+1) chunks -> heaviest rows first, columns fastest (LPT order)
      tj = (int)(t % g.nt);
      ti = g.mt - 1 - (int)(t / g.nt);
    } else {
      ti = (int)(t % g.mt);
      tj = (int)(t / g.mt);
    }
    int kend = g.K;
    if (g.klimit) {
      long long lim = (long long)(ti + 1) * TM;
      if (lim < kend) kend = (int)lim;
    }
    nchunks = kend / GKC;
  };
  // first row of A / C tile row ti (row lists: the block rows a device owns in the distributed factorization)
  auto rowoff = [&](int ti) -> long long {
    if (!g.rows) return (long long)ti * TM;
    constexpr int PER = GT / TM;
    return (long long)g.rows[ti / PER] * GT + (long long)(ti % PER) * TM;
  };
  auto skip = [&](int ti, int tj) -> bool { return g.stair && (long long)g.colblk0 * GT + (long long)tj * TN > rowoff(ti); };
  // The grid is either one CTA per tile or (max_ctas) a persistent one walking the tiles; `cg` numbers the K chunks of all the
  // tiles a CTA processes, so the ring and its mbarrier phases simply keep running across tiles.

  if (warp == CW) {
    // ------------------------------------------------------------ producer
    long long cg = 0;
    for (long long t = blockIdx.x; t < g.ntiles; t += gridDim.x) {
    int ti, tj, nchunks;
    decode(t, ti, tj, nchunks);
    if (skip(ti, tj)) continue;
    const double* Abase = g.A + rowoff(ti);
    const double* Bbase = BK ? g.B + (long long)tj * TN * g.ldb : g.B + (long long)tj * TN;
    for (int c = 0; c < nchunks; ++c, ++cg) {
      const int s = (int)(cg % GSTAGES);
      const uint32_t ph = (uint32_t)((cg / GSTAGES) & 1);
      mbar_wait(&empty[s], ph ^ 1u);
      unsigned char* sa = stage_base + s * C_::STAGE_BYTES;
      unsigned char* sb = sa + C_::A_BYTES;
      const long long k0 = (long long)c * GKC;
      if (lane == 0) mbar_arrive_expect_tx(&full[s], (uint32_t)(GKC * (TM + TN) * 8));
      __syncwarp();
      if (lane < GKC) {
        bulk_g2s(sa + lane * LDA * 8, Abase + (k0 + lane) * g.lda, TM * 8, &full[s]);
      } else if (!BK) {
        const int l = lane - GKC;
        bulk_g2s(sb + l * LDB * 8, Bbase + (k0 + l) * g.ldb, TN * 8, &full[s]);
      }
      if (BK) {
#pragma unroll
        for (int i = 0; i < TN / 32; ++i) {
          const int n = lane + 32 * i;
          bulk_g2s(sb + n * GLDBK * 8, Bbase + k0 + (long long)n * g.ldb, GKC * 8, &full[s]);
        }
      }
    }
    }
    return;
  }

  // -------------------------------------------------------------- consumers
  const int wm = warp % WM, wn = warp / WM;
  const int lr = lane >> 2, lk = lane & 3;
  long long cg = 0;
  for (long long t = blockIdx.x; t < g.ntiles; t += gridDim.x) {
  int ti, tj, nchunks;
  decode(t, ti, tj, nchunks);
  if (skip(ti, tj)) continue;
  double acc[FM][FN][2];
  const long long row0 = rowoff(ti) + wm * (TM / WM) + lr;
  const long long col0 = (long long)tj * TN + wn * (TN / WN) + 2 * lk;
  if (MODE == GEMM_SUB) {
    // C -= A B^T: the accumulators START at C and the A fragments are negated, so the tile's C is read here - all loads in flight
    // at once, under the latency of the first operand chunk - and the epilogue is a plain store.  (A read-modify-write epilogue
    // serialises on memory latency: its loads cannot be hoisted over the preceding stores; measured 30 us per tile on the B200,
    // 40 % of a K = 512 tile.)
#pragma unroll
    for (int a = 0; a < FM; ++a)
#pragma unroll
      for (int b = 0; b < FN; ++b) {
        const double* p0 = g.C + (row0 + a * 8) + (col0 + b * 8) * g.ldc;
        acc[a][b][0] = p0[0];
        acc[a][b][1] = p0[g.ldc];
      }
  } else {
#pragma unroll
    for (int a = 0; a < FM; ++a)
#pragma unroll
      for (int b = 0; b < FN; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  }

  for (int c = 0; c < nchunks; ++c, ++cg) {
    const int s = (int)(cg % GSTAGES);
    const uint32_t ph = (uint32_t)((cg / GSTAGES) & 1);
    mbar_wait(&full[s], ph);
    const double* sa = reinterpret_cast<const double*>(stage_base + s * C_::STAGE_BYTES);
    const double* sb = reinterpret_cast<const double*>(stage_base + s * C_::STAGE_BYTES + C_::A_BYTES);
#pragma unroll
    for (int k4 = 0; k4 < GKC / 4; ++k4) {
      const int k = k4 * 4 + lk;
      double af[FM], bf[FN];
#pragma unroll
      for (int a = 0; a < FM; ++a) af[a] = (MODE == GEMM_SUB) ? -sa[k * LDA + wm * (TM / WM) + a * 8 + lr] : sa[k * LDA + wm * (TM / WM) + a * 8 + lr];
#pragma unroll
      for (int b = 0; b < FN; ++b) {
        const int n = wn * (TN / WN) + b * 8 + lr;
        bf[b] = BK ? sb[n * GLDBK + k] : sb[k * LDB + n];
      }
#pragma unroll
      for (int a = 0; a < FM; ++a)
#pragma unroll
        for (int b = 0; b < FN; ++b) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }

  // -------------------------------------------------------------- epilogue
  if (MODE == GEMM_SET && g.npeer > 0) {
    // Multicast epilogue (distributed factorization: the solved panel rows go into every device's matrix).  The tile is staged in
    // the operand ring - free now: one tile per CTA (launch_gemm_cfg guarantees it when npeer > 0), every chunk consumed once the
    // consumer warps meet at the barrier - and leaves as whole columns: 16 bytes per lane, TM*8 contiguous bytes per store
    // instruction and destination, which is what NVLink wants (the fragment layout would give 64-byte pieces).
    constexpr int LDS = TM + 4;
    static_assert(TN * LDS * 8 <= GSTAGES * C_::STAGE_BYTES, "staging tile must fit the operand ring");
    double* S = reinterpret_cast<double*>(stage_base);
    named_bar_sync(1, CW * 32);
    const int r0s = wm * (TM / WM) + lr, c0s = wn * (TN / WN) + 2 * lk;
#pragma unroll
    for (int a = 0; a < FM; ++a)
#pragma unroll
      for (int b = 0; b < FN; ++b) {
        S[(c0s + b * 8) * LDS + r0s + a * 8] = acc[a][b][0];
        S[(c0s + b * 8 + 1) * LDS + r0s + a * 8] = acc[a][b][1];
      }
    named_bar_sync(1, CW * 32);
    const long long rbase = rowoff(ti), cbase = (long long)tj * TN;
    for (int idx = warp * 32 + lane; idx < TN * (TM / 2); idx += CW * 32) {  // consecutive lanes: consecutive row pairs of a column
      const int c = idx / (TM / 2), r2 = 2 * (idx % (TM / 2));
      const double2 v = *reinterpret_cast<const double2*>(S + c * LDS + r2);
      const long long off = rbase + r2 + (cbase + c) * g.ldc;
      *reinterpret_cast<double2*>(g.C + off) = v;
      for (int p = 0; p < g.npeer; ++p) *reinterpret_cast<double2*>(g.Cpeer[p] + off) = v;
    }
    continue;
  }
#pragma unroll
  for (int a = 0; a < FM; ++a) {
    const long long i = row0 + a * 8;
#pragma unroll
    for (int b = 0; b < FN; ++b) {
      const long long j = col0 + b * 8;
      if (MODE == GEMM_SET) {
        g.C[i + j * g.ldc] = acc[a][b][0];
        g.C[i + (j + 1) * g.ldc] = acc[a][b][1];
      } else if (MODE == GEMM_SUB) {
        g.C[i + j * g.ldc] = acc[a][b][0];
        g.C[i + (j + 1) * g.ldc] = acc[a][b][1];
      } else {
        if (i < g.Ns) {
          const long long zi = g.sinds[i];
          const double add = g.d2[i] + g.addmu;
          if (j < g.R) g.C[zi + j * g.ldc] = acc[a][b][0] + add;
          if (j + 1 < g.R) g.C[zi + (j + 1) * g.ldc] = acc[a][b][1] + add;
        }
      }
    }
  }
  }
}

template <int MODE, bool BK, int TM, int TN, int WM, int WN>
inline cudaError_t launch_gemm_cfg(cudaStream_t st, const GemmArgs& g0) {
  GemmArgs g = g0;
  using C_ = GemmCfg<TM, TN, WM, WN>;
  auto kfn = gemm_dmma_kernel<MODE, BK, TM, TN, WM, WN>;
  // per-device attribute, cheap host-side call: set it on every launch (multi-device contexts)
  cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, C_::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  long long tiles = g.tri ? (long long)g.mt * (g.mt + 1) / 2 : (long long)g.mt * g.nt;
  if (tiles <= 0 || g.K <= 0) return cudaSuccess;
  g.ntiles = tiles;
  static int force_ctas = -1;  // GSP_GEMM_MAX_CTAS (tests): every GEMM runs as a persistent grid of at most this many CTAs
  if (force_ctas < 0) {
    const char* env = getenv("GSP_GEMM_MAX_CTAS");
    force_ctas = env ? atoi(env) : 0;
  }
  if (force_ctas > 0) g.max_ctas = force_ctas;
  if (g.npeer > 0) g.max_ctas = 0;  // the multicast epilogue stages its tile in the operand ring: one tile per CTA
  const long long grid = (g.max_ctas > 0 && g.max_ctas < tiles) ? g.max_ctas : tiles;
  ProfScope prof_(MODE == GEMM_SAMPLE ? "gemm_dmma_sample" : (MODE == GEMM_SET ? "gemm_dmma_trsm" : "gemm_dmma_update"), st);
  GSP_LAUNCH(kfn, dim3((unsigned)grid), dim3(C_::THREADS), (size_t)C_::SMEM_BYTES, st, g);
  g_launches++;
  return cudaGetLastError();
}

// `g.mt`, `g.nt` are given in 128-blocks.  Problems with too few 128x128 tiles to fill the GPU are
// retiled to 64-row (in-place TRSM: the C tile must span all 128 columns) or 64x64 tiles: 2-4x more CTAs
// on the latency-bound small GEMMs of the factorisation's critical path.
// `work_tiles` (>= 0): number of 128x128 tiles that are actually computed when the stair predicate skips part of the mt x nt grid.
template <int MODE, bool BK>
inline cudaError_t launch_gemm(cudaStream_t st, const GemmArgs& g0, long long work_tiles = -1) {
  const long long tiles128 = work_tiles >= 0 ? work_tiles : (g0.tri ? (long long)g0.mt * (g0.mt + 1) / 2 : (long long)g0.mt * g0.nt);
  static long long kSmall = -1;  // GSP_GEMM_SMALL_TILES overrides the retiling threshold (tests force either path)
  if (kSmall < 0) {
    const char* env = getenv("GSP_GEMM_SMALL_TILES");
    kSmall = env ? atoll(env) : 100;
  }
  if (MODE == GEMM_SET && !BK && g0.npeer > 0 && tiles128 < 48) {
    // multicast solves: the peer stores of one SM run at ~10 GB/s over NVLink, so the rows are spread over 4x more CTAs
    GemmArgs g = g0;
    g.mt = 4 * g0.mt;
    return launch_gemm_cfg<MODE, BK, 32, 128, 1, 8>(st, g);
  }
  if (MODE == GEMM_SET && !BK && tiles128 < kSmall) {
    GemmArgs g = g0;
    g.mt = 2 * g0.mt;
    return launch_gemm_cfg<MODE, BK, 64, 128, 2, 4>(st, g);
  }
  if (MODE == GEMM_SUB && !BK && tiles128 < kSmall) {
    GemmArgs g = g0;
    g.mt = 2 * g0.mt;
    g.nt = 2 * g0.nt;
    return launch_gemm_cfg<MODE, BK, 64, 64, 2, 2>(st, g);
  }
  return launch_gemm_cfg<MODE, BK, 128, 128, 2, 4>(st, g0);
}

}  // namespace gsp
