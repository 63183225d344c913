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
#include "common.h"

namespace gsp {

constexpr int GT = 128;        // C tile edge
constexpr int GKC = 16;        // K chunk per stage
constexpr int GSTAGES = 4;
constexpr int GLDA = GT + 4;   // smem leading dim of [k][m] operand tiles (132 = 4 mod 16)
constexpr int GLDBK = GKC + 4; // smem leading dim of [n][k] operand tiles (20 = 4 mod 16)
constexpr int G_A_BYTES = GKC * GLDA * 8;                         // 16896
constexpr int G_B_BYTES = (GT * GLDBK * 8 > G_A_BYTES) ? GT * GLDBK * 8 : G_A_BYTES;  // 20480
constexpr int G_STAGE_BYTES = G_A_BYTES + G_B_BYTES;              // 37376
constexpr int G_SMEM_BYTES = GSTAGES * G_STAGE_BYTES + 2 * GSTAGES * 16 + 128;
constexpr int G_THREADS = 288;  // 8 consumer warps + 1 producer warp

enum GemmMode { GEMM_SET = 0, GEMM_SUB = 1, GEMM_SAMPLE = 2 };

struct GemmArgs {
  const double* A;   // M x K, A[i + k*lda]
  long long lda;
  const double* B;   // !BK: N x K, B[j + k*ldb] (C = A * B^T);  BK: K x N, B[k + j*ldb] (C = A * B)
  long long ldb;
  double* C;         // M x N, C[i + j*ldc]
  long long ldc;
  int mt, nt;        // tiles along M and N
  int K;             // multiple of GKC
  int tri;           // 1: only tiles with ti >= tj (mt == nt)
  int klimit;        // 1: A is lower triangular -> k < (ti+1)*GT only
  // GEMM_SAMPLE epilogue: Z[sinds[i] + r*ldz] = acc + d2[i] + addmu for i < Ns, r < R
  const double* d2;
  const long long* sinds;
  double addmu;
  long long Ns, R;
};

template <int MODE, bool BK>
__global__ void __launch_bounds__(G_THREADS, 1) gemm_dmma_kernel(GemmArgs g) {
  GSP_DYN_SMEM(smem);
  // carve: stages first (16-byte aligned), then barriers
  unsigned char* stage_base = smem;
  mbar_t* full = reinterpret_cast<mbar_t*>(smem + GSTAGES * G_STAGE_BYTES);
  mbar_t* empty = full + GSTAGES;

  // tile coordinates
  int ti, tj;
  {
    const long long t = blockIdx.x;
    if (g.tri) {
      int i = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
      while ((long long)(i + 1) * (i + 2) / 2 <= t) ++i;
      while ((long long)i * (i + 1) / 2 > t) --i;
      ti = i;
      tj = (int)(t - (long long)i * (i + 1) / 2);
    } else {
      ti = (int)(t % g.mt);
      tj = (int)(t / g.mt);
      if (g.klimit) ti = g.mt - 1 - ti;  // heaviest row tiles first
    }
  }
  int kend = g.K;
  if (g.klimit) {
    long long lim = (long long)(ti + 1) * GT;
    if (lim < kend) kend = (int)lim;
  }
  const int nchunks = kend / GKC;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < GSTAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 8);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == 8) {
    // ------------------------------------------------------------ producer
    const double* Abase = g.A + (long long)ti * GT;
    const double* Bbase = BK ? g.B + (long long)tj * GT * g.ldb : g.B + (long long)tj * GT;
    for (int c = 0; c < nchunks; ++c) {
      const int s = c % GSTAGES;
      const uint32_t ph = (uint32_t)((c / GSTAGES) & 1);
      mbar_wait(&empty[s], ph ^ 1u);
      unsigned char* sa = stage_base + s * G_STAGE_BYTES;
      unsigned char* sb = sa + G_A_BYTES;
      const long long k0 = (long long)c * GKC;
      if (lane == 0) mbar_arrive_expect_tx(&full[s], (uint32_t)(2 * GKC * GT * 8));
      __syncwarp();
      if (lane < GKC) {
        bulk_g2s(sa + lane * GLDA * 8, Abase + (k0 + lane) * g.lda, GT * 8, &full[s]);
      } else if (!BK) {
        const int l = lane - GKC;
        bulk_g2s(sb + l * GLDA * 8, Bbase + (k0 + l) * g.ldb, GT * 8, &full[s]);
      }
      if (BK) {
#pragma unroll
        for (int i = 0; i < GT / 32; ++i) {
          const int n = lane + 32 * i;
          bulk_g2s(sb + n * GLDBK * 8, Bbase + k0 + (long long)n * g.ldb, GKC * 8, &full[s]);
        }
      }
    }
    return;
  }

  // -------------------------------------------------------------- consumers
  const int wm = warp & 1, wn = warp >> 1;  // 2 x 4 warps -> 64 x 32 sub-tiles
  const int lr = lane >> 2, lk = lane & 3;
  double acc[8][4][2];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

  for (int c = 0; c < nchunks; ++c) {
    const int s = c % GSTAGES;
    const uint32_t ph = (uint32_t)((c / GSTAGES) & 1);
    mbar_wait(&full[s], ph);
    const double* sa = reinterpret_cast<const double*>(stage_base + s * G_STAGE_BYTES);
    const double* sb = reinterpret_cast<const double*>(stage_base + s * G_STAGE_BYTES + G_A_BYTES);
#pragma unroll
    for (int k4 = 0; k4 < GKC / 4; ++k4) {
      const int k = k4 * 4 + lk;
      double af[8], bf[4];
#pragma unroll
      for (int a = 0; a < 8; ++a) af[a] = sa[k * GLDA + wm * 64 + a * 8 + lr];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int n = wn * 32 + b * 8 + lr;
        bf[b] = BK ? sb[n * GLDBK + k] : sb[k * GLDA + n];
      }
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }

  // -------------------------------------------------------------- epilogue
  const long long row0 = (long long)ti * GT + wm * 64 + lr;
  const long long col0 = (long long)tj * GT + wn * 32 + 2 * lk;
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const long long i = row0 + a * 8;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const long long j = col0 + b * 8;
      if (MODE == GEMM_SET) {
        g.C[i + j * g.ldc] = acc[a][b][0];
        g.C[i + (j + 1) * g.ldc] = acc[a][b][1];
      } else if (MODE == GEMM_SUB) {
        double* p0 = g.C + i + j * g.ldc;
        double* p1 = p0 + g.ldc;
        *p0 -= acc[a][b][0];
        *p1 -= acc[a][b][1];
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

template <int MODE, bool BK>
inline cudaError_t launch_gemm(cudaStream_t st, const GemmArgs& g) {
  auto kfn = gemm_dmma_kernel<MODE, BK>;
  // per-device attribute, cheap host-side call: set it on every launch (multi-device contexts)
  cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  long long tiles = g.tri ? (long long)g.mt * (g.mt + 1) / 2 : (long long)g.mt * g.nt;
  if (tiles <= 0 || g.K <= 0) return cudaSuccess;
  ProfScope prof_(MODE == GEMM_SAMPLE ? "gemm_dmma_sample" : (MODE == GEMM_SET ? "gemm_dmma_trsm" : "gemm_dmma_update"), st);
  GSP_LAUNCH(kfn, dim3((unsigned)tiles), dim3(G_THREADS), (size_t)G_SMEM_BYTES, st, g);
  g_launches++;
  return cudaGetLastError();
}

}  // namespace gsp
