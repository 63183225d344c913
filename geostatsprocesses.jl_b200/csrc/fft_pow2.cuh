// Register-resident power-of-two FFT passes for FFTSIM (the fast path for extents 16..4096).
//
// Every thread keeps SL = max-radix complex values ("slots") of one line in registers for the whole
// transform; a stage of radix R < SL does SL/R butterflies per thread, so all threads stay busy in
// every stage.  Stages are Stockham autosort; between stages the line is exchanged through shared
// memory exactly once, and the last forward stage leaves frequency f = jb + r*(N/R_last) in the same
// slot that the first stage of the inverse transform (reversed radix order) wants, so the spectral
// multiply P = s*F*W/|W| (fftsim.jl:125) and the turn-around to the inverse happen in registers.
// Strided (y/z) passes bundle B adjacent kx so that every global access is a contiguous run of
// B*16 bytes and every shared-memory access is conflict-free; x passes put lanes along the row and
// pad the row by one element per R0 to break the stride-R0 conflict of the first exchange.
#pragma once
#include "fft_kernels.cuh"
#include "rng.cuh"

namespace gsp {

// GSP_P2_R8 = 1 (experiment): extents 128 and 256 on 8 register slots per thread (radix 8 x 4 x 4 / 8 x 8 x 4, two exchanges) instead of
// 16 (radix 16 x 8 / 16 x 16, one exchange): half the registers per thread, twice the threads per line.
#ifndef GSP_P2_R8
#define GSP_P2_R8 0
#endif
#ifndef GSP_P2_B256
#define GSP_P2_B256 8  // kx per bundle of the strided passes for extents <= 512
#endif
GSP_HD constexpr int p2_stages(int N) {
  return N <= 16 ? 1 : ((N == 512 || N >= 1024 || (GSP_P2_R8 && (N == 128 || N == 256))) ? 3 : 2);
}
GSP_HD constexpr int p2_radix_fwd(int N, int s) {
  switch (N) {
    case 2: return 2;
    case 4: return 4;
    case 8: return 8;
    case 16: return 16;
    case 32: return s == 0 ? 16 : 2;
    case 64: return 8;
    case 128: return GSP_P2_R8 ? (s == 0 ? 8 : 4) : (s == 0 ? 16 : 8);
    case 256: return GSP_P2_R8 ? (s < 2 ? 8 : 4) : 16;
    case 512: return 8;
    case 1024: return s < 2 ? 16 : 4;
    case 2048: return s < 2 ? 16 : 8;
    case 4096: return 16;
    default: return 0;
  }
}
GSP_HD constexpr bool p2_supported(int N) { return N >= 16 && N <= 4096 && (N & (N - 1)) == 0; }
GSP_HD constexpr int p2_radix(int N, bool inv, int s) { return p2_radix_fwd(N, inv ? p2_stages(N) - 1 - s : s); }
GSP_HD constexpr int p2_ns(int N, bool inv, int s) {
  int ns = 1;
  for (int i = 0; i < s; ++i) ns *= p2_radix(N, inv, i);
  return ns;
}
GSP_HD constexpr int p2_slots(int N) {
  int m = 0;
  for (int s = 0; s < p2_stages(N); ++s) m = p2_radix_fwd(N, s) > m ? p2_radix_fwd(N, s) : m;
  return m;
}
GSP_HD constexpr int p2_log2(int x) { return x <= 1 ? 0 : 1 + p2_log2(x / 2); }

// position (along the line) of slot (q, r) among the INPUTS of stage s
template <int N, bool INV, int S_>
GSP_DEV int p2_in_pos(int t, int q, int r) {
  constexpr int R = p2_radix(N, INV, S_);
  constexpr int TPL = N / p2_slots(N);
  return (t + TPL * q) + r * (N / R);
}
// position of slot (q, r) among the OUTPUTS of stage s
template <int N, bool INV, int S_>
GSP_DEV int p2_out_pos(int t, int q, int r) {
  constexpr int R = p2_radix(N, INV, S_);
  constexpr int Ns = p2_ns(N, INV, S_);
  constexpr int TPL = N / p2_slots(N);
  const int jb = t + TPL * q;
  const int k = jb & (Ns - 1);
  return (jb - k) * R + k + r * Ns;
}

// Per-stage twiddle tables (x passes): stage s >= 1 stores w[r][k] = exp(-2*pi*i*r*k/(Ns*R)), k fastest, so that lanes with
// consecutive k read consecutive 16-byte entries (the strided table tw[(r*k)*step] costs up to 8-way bank conflicts there).
GSP_HD constexpr int p2_stw_offset(int N, bool inv, int s) {
  int off = 0;
  for (int i = 1; i < s; ++i) off += p2_radix(N, inv, i) * p2_ns(N, inv, i);
  return off;
}
GSP_HD constexpr int p2_stw_size(int N, bool inv) { return p2_stw_offset(N, inv, p2_stages(N)); }

// twiddle + butterflies of stage s on the register slots.  TWS > 0: tw[i * TWS] = exp(-2*pi*i*i/N) (strided table);
// TWS == 0: tw = per-stage tables (p2_stw_offset layout).  Both in shared memory.
template <int N, bool INV, int S_, int TWS>
GSP_DEV void p2_stage(cplx* v, int t, const cplx* tw) {
  constexpr int R = p2_radix(N, INV, S_);
  constexpr int Ns = p2_ns(N, INV, S_);
  constexpr int SL = p2_slots(N);
  constexpr int TPL = N / SL;
  constexpr int Q = SL / R;
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    if constexpr (Ns > 1) {
      const int k = (t + TPL * q) & (Ns - 1);
#pragma unroll
      for (int r = 1; r < R; ++r) {
        cplx w = (TWS == 0) ? tw[p2_stw_offset(N, INV, S_) + r * Ns + k] : tw[(r * k) * (N / (Ns * R)) * (TWS == 0 ? 1 : TWS)];
        if (INV) w.im = -w.im;
        v[q * R + r] = cmul(v[q * R + r], w);
      }
    }
    dft_pow2<R, INV>(v + q * R);
  }
}

// exchange through shared memory: outputs of stage s -> inputs of stage s+1.  Lay(m) = element index.
template <int N, bool INV, int S_, class Lay>
GSP_DEV void p2_exchange(cplx* v, int t, cplx* buf, Lay lay) {
  constexpr int SL = p2_slots(N);
  constexpr int R = p2_radix(N, INV, S_);
  constexpr int R2 = p2_radix(N, INV, S_ + 1);
  __syncthreads();  // previous readers of buf are done
#pragma unroll
  for (int q = 0; q < SL / R; ++q)
#pragma unroll
    for (int r = 0; r < R; ++r) buf[lay(p2_out_pos<N, INV, S_>(t, q, r))] = v[q * R + r];
  __syncthreads();
#pragma unroll
  for (int q = 0; q < SL / R2; ++q)
#pragma unroll
    for (int r = 0; r < R2; ++r) v[q * R2 + r] = buf[lay(p2_in_pos<N, INV, S_ + 1>(t, q, r))];
}

struct P2NoHook {
  GSP_DEV void operator()() const {}
};

// `after_last_exchange` runs once every thread's slots are back in registers for the final stage: from there on the transform no
// longer touches `buf` (single-stage passes start the next item's load at that point)
template <int N, bool INV, int S_, int TWS, class Lay, class Hook>
GSP_DEV void p2_run_from(cplx* v, int t, cplx* buf, Lay lay, const cplx* tw, Hook after_last_exchange) {
  p2_stage<N, INV, S_, TWS>(v, t, tw);
  if constexpr (S_ + 1 < p2_stages(N)) {
    p2_exchange<N, INV, S_>(v, t, buf, lay);
    if constexpr (S_ + 2 == p2_stages(N)) after_last_exchange();
    p2_run_from<N, INV, S_ + 1, TWS>(v, t, buf, lay, tw, after_last_exchange);
  }
}

// Full transform of one line held in registers.
//   in : slot (q, r) = x[p2_in_pos<N,INV,0>(t,q,r)]
//   out: slot (q, r) = X[p2_in_pos<N,!INV,0>(t,q,r)]  (natural order; same mapping as the opposite direction's input)
template <int N, bool INV, int TWS, class Lay, class Hook = P2NoHook>
GSP_DEV void p2_fft(cplx* v, int t, cplx* buf, Lay lay, const cplx* tw, Hook after_last_exchange = Hook()) {
  p2_run_from<N, INV, 0, TWS>(v, t, buf, lay, tw, after_last_exchange);
}

// Two-stage transforms (N <= 256) with the stage-1 twiddles of THIS thread in registers: they depend on (t, q, r) only, so a
// persistent kernel loads them once instead of once per item (the x passes are bound by the shared-memory pipe, not by registers).
template <int N>
struct P2RegTw {
  static constexpr bool OK = p2_stages(N) == 2;
  static constexpr int R1F = p2_radix(N, false, 1), R1I = p2_radix(N, true, 1);
  static constexpr int NF = OK ? (p2_slots(N) / R1F) * (R1F - 1) : 1;  // forward: Q * (R - 1) entries
  static constexpr int NI = OK ? (p2_slots(N) / R1I) * (R1I - 1) : 1;
};
// stw: per-stage table of direction INV in shared memory (p2_stw_offset layout); out[q * (R-1) + r-1], conjugated for INV
template <int N, bool INV, int NT>
GSP_DEV void p2_load_regtw(cplx (&out)[NT], int t, const cplx* stw) {
  constexpr int R = p2_radix(N, INV, 1), Ns = p2_ns(N, INV, 1), SL = p2_slots(N), TPL = N / SL, Q = SL / R;
  static_assert(NT == Q * (R - 1), "register twiddle count");
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int k = (t + TPL * q) & (Ns - 1);
#pragma unroll
    for (int r = 1; r < R; ++r) {
      cplx w = stw[p2_stw_offset(N, INV, 1) + r * Ns + k];
      if (INV) w.im = -w.im;
      out[q * (R - 1) + r - 1] = w;
    }
  }
}
template <int N, bool INV, int NT, class Lay>
GSP_DEV void p2_fft_regtw(cplx* v, int t, cplx* buf, Lay lay, const cplx (&rtw)[NT]) {
  constexpr int R = p2_radix(N, INV, 1), SL = p2_slots(N), Q = SL / R;
  p2_stage<N, INV, 0, 0>(v, t, nullptr);  // stage 0: no twiddles
  p2_exchange<N, INV, 0>(v, t, buf, lay);
#pragma unroll
  for (int q = 0; q < Q; ++q) {
#pragma unroll
    for (int r = 1; r < R; ++r) v[q * R + r] = cmul(v[q * R + r], rtw[q * (R - 1) + r - 1]);
    dft_pow2<R, INV>(v + q * R);
  }
}

struct BundleLay {  // strided passes: [position][b], b fastest
  int B, b;
  GSP_DEV int operator()(int m) const { return m * B + b; }
};
template <int SH>
struct RowLay {  // x passes: one padded row per line, one extra element every 2^SH
  int base;
  GSP_DEV int operator()(int m) const { return base + m + (m >> SH); }
};

enum { P2_FWD = 1, P2_MUL = 2, P2_INV = 4 };

GSP_HD constexpr int p2_bundle(int N) { return N <= 512 ? GSP_P2_B256 : (N <= 2048 ? 4 : 2); }

// All three pass kernels are PERSISTENT and software-pipelined: a CTA walks over its work items
// (line bundles / row groups); while it transforms item i out of shared-memory stage i&1, the TMA
// engine (cp.async.bulk, completing on an mbarrier) is already filling stage (i+1)&1 with item i+1.
// Results leave straight from registers with 16-byte streaming stores.

// ------------------------------------------------------------------------------------------------
// strided pass (y or z axis) over the half spectrum, in place.  One item = B adjacent kx times the
// whole line.  FLAGS: FWD only / INV only / FWD|MUL|INV (last axis: spectral multiply fused in).
// F is stored with an even leading dimension (hxF) so that its 8-byte rows are 16-byte aligned (TMA global strides).
// STAGES = 2: the next item is prefetched while the current one is transformed (fewer, fatter CTAs);
// STAGES = 1: no prefetch, half the shared memory -> more resident CTAs overlap each other's phases instead.
template <int N, int B, int FLAGS, int STAGES>
struct StridedCfg {
  static constexpr int SL = p2_slots(N);
  static constexpr int TPL = N / SL;
  static constexpr int THREADS = TPL * B;
  static constexpr bool MUL = (FLAGS & P2_MUL) != 0;
  static constexpr size_t IN_BYTES = (size_t)N * B * sizeof(cplx);
  static constexpr size_t F_BYTES = 0;  // F goes global -> registers (issued before the forward transform): keeps 3 CTAs per SM
  static constexpr int STWF = p2_stw_size(N, false), STWI = p2_stw_size(N, true);  // per-stage twiddle tables
  // long lines: the per-stage tables would not fit next to the data -> plain strided table exp(-2*pi*i*t/N) instead
  static constexpr bool STW_OK = (size_t)(STWF + STWI) * sizeof(cplx) <= 40 * 1024;
  static constexpr size_t TW_BYTES = ((size_t)(STW_OK ? STWF + STWI : N) * sizeof(cplx) + 127) / 128 * 128;
  static constexpr size_t SMEM = TW_BYTES + STAGES * (IN_BYTES + F_BYTES) + 2 * sizeof(mbar_t) + 16;
#ifndef GSP_STRIDED_MINB
#define GSP_STRIDED_MINB 4
#endif
  // register cap for that many CTAs per SM (8-slot experiment: 24 warps per SM)
  static constexpr int MINB = (GSP_P2_R8 && SL == 8 && STAGES == 1 && THREADS <= 256) ? 768 / THREADS
                                                                                      : ((STAGES == 1 && THREADS <= 128) ? GSP_STRIDED_MINB
                                                                                         : ((STAGES == 1 && THREADS == 256) ? 2 : 1));
};

template <int N, int B, int FLAGS, int STAGES>
__global__ void __launch_bounds__(StridedCfg<N, B, FLAGS, STAGES>::THREADS, StridedCfg<N, B, FLAGS, STAGES>::MINB)
    p2_strided_kernel(const GSP_GRID_CONSTANT TensorMap tmH, int line_axis,
                                                                                     cplx* __restrict__ H, const cplx* __restrict__ twg,
                                                                                     const cplx* __restrict__ stwfg, const cplx* __restrict__ stwig,
                                                                                     long long es, int hx, int nbundles, long long nunits,
                                                                                     long long other_stride, const double* __restrict__ Fh,
                                                                                     long long esF, long long other_strideF, double s, int bx0,
                                                                                     long long ubeg, const double* __restrict__ Fperm) {
  // work units [ubeg, nunits): unit = o * nbundles + (bx - bx0), i.e. `nbundles` kx bundles starting at bundle bx0 for every
  // index o along the remaining axis.  Sub-ranges (slabs of planes or groups of kx bundles) keep a slab resident in L2.
  using C = StridedCfg<N, B, FLAGS, STAGES>;
  constexpr int SL = C::SL, TPU = C::THREADS;
  constexpr bool MUL = C::MUL;
  GSP_DYN_SMEM(smem);
  constexpr int TWS = C::STW_OK ? 0 : 1;
  cplx* stwf = reinterpret_cast<cplx*>(smem);                   // per-stage twiddles, forward radix order (or the plain table)
  cplx* stwi = C::STW_OK ? stwf + C::STWF : stwf;               // ... inverse radix order
  unsigned char* stage0 = smem + C::TW_BYTES;
  mbar_t* full = reinterpret_cast<mbar_t*>(stage0 + STAGES * (C::IN_BYTES + C::F_BYTES));
  const int tid = threadIdx.x;
  const int b = tid % B, t = tid / B;
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_mbar_init();
  }
  if constexpr (C::STW_OK) {
    if ((FLAGS & P2_FWD) != 0)
      for (int i = tid; i < C::STWF; i += TPU) stwf[i] = stwfg[i];
    if ((FLAGS & P2_INV) != 0)
      for (int i = tid; i < C::STWI; i += TPU) stwi[i] = stwig[i];
  } else {
    for (int i = tid; i < N; i += TPU) stwf[i] = twg[i];
  }
  __syncthreads();

  auto stage_in = [&](int sg) { return reinterpret_cast<cplx*>(stage0 + (size_t)sg * (C::IN_BYTES + C::F_BYTES)); };
  // one TMA tensor copy per item and operand: box = (B kx as 2B doubles) x (whole line), zero-filled past hx
  auto issue = [&](long long unit, int sg) {
    if (tid == 0) {
      // the stage was last touched through the generic proxy (exchanges overlay it); the block barrier that ended that
      // iteration ordered those accesses before this thread, the proxy fence orders them before the async-proxy fill
      fence_proxy_async();
      const long long o = unit / nbundles;
      const int bx = bx0 + (int)(unit - o * nbundles);
      const int c1 = line_axis == 1 ? 0 : (int)o, c2 = line_axis == 1 ? (int)o : 0;
      mbar_arrive_expect_tx(&full[sg], (uint32_t)(C::IN_BYTES + C::F_BYTES));
      constexpr int LBOX = N < 256 ? N : 256;  // TMA boxes are limited to 256 elements per dimension
#pragma unroll
      for (int k = 0; k < N; k += LBOX) {
        tma_load_3d(stage_in(sg) + (size_t)k * B, &tmH, bx * 2 * B, line_axis == 1 ? k : c1, line_axis == 1 ? c2 : k, &full[sg]);
      }
    }
  };

  // STAGES == 1 and GSP_STRIDED_EARLY: the one stage is refilled as soon as the LAST exchange of the item has been read back into
  // registers - the final radix stage and the stores of item i then overlap the load of item i + 1 (the exposed load wait is
  // 22 % of the warp samples of these kernels, profiles/r01_s3_ncu_full_summary.csv).  Measured on the B200 at 256^3: no change
  // (279.7 vs 279.8 us per realization, every pass within 1 us) - the other 3 CTAs of the SM already fill that wait.  Off.
#ifndef GSP_STRIDED_EARLY
#define GSP_STRIDED_EARLY 0
#endif
  constexpr bool EARLY = STAGES == 1 && GSP_STRIDED_EARLY != 0;
  constexpr bool HAS_EX = p2_stages(N) > 1;
  long long unit = ubeg + blockIdx.x;
  if ((STAGES == 2 || EARLY) && unit < nunits) issue(unit, 0);
  for (int it = 0; unit < nunits; unit += gridDim.x, ++it) {
    const int cur = STAGES == 2 ? (it & 1) : 0;
    if (STAGES == 2) {
      if (unit + gridDim.x < nunits) issue(unit + gridDim.x, cur ^ 1);
    } else if (!EARLY) {
      issue(unit, 0);
    }
    auto refill = [&]() {
      if constexpr (EARLY) {
        __syncthreads();  // every thread has read its slots out of the stage
        if (unit + gridDim.x < nunits) issue(unit + gridDim.x, 0);
      }
    };
    const long long o = unit / nbundles;
    const int bx = bx0 + (int)(unit - o * nbundles);
    const bool valid = bx * B + b < hx;
    const long long base = o * other_stride + (long long)bx * B + b;
    cplx* buf = stage_in(cur);
    const BundleLay lay{B, b};
    // F of this thread's 16 frequencies: read-only, coalesced 64-byte runs, in flight during the forward transform
    double fv[MUL ? SL : 1];
    if constexpr (MUL) {
      constexpr int RI = p2_radix(N, true, 0);
      if (Fperm) {
        // item-major copy of F (p2_permute_F_kernel): slot j of all the item's threads is one contiguous run
        const long long item = o * ((hx + B - 1) / B) + bx;
        const double* fp = Fperm + item * (long long)(SL * TPU) + tid;
#pragma unroll
        for (int j = 0; j < SL; ++j) fv[j] = __ldg(fp + j * TPU);
      } else {
        const double* fp = Fh + o * other_strideF + (long long)bx * B + b;
#pragma unroll
        for (int q = 0; q < SL / RI; ++q)
#pragma unroll
          for (int r = 0; r < RI; ++r) fv[q * RI + r] = valid ? __ldg(fp + (long long)p2_in_pos<N, true, 0>(t, q, r) * esF) : 0.0;
      }
    }
    mbar_wait(&full[cur], (uint32_t)(STAGES == 2 ? ((it >> 1) & 1) : (it & 1)));
    constexpr bool FIRST_INV = (FLAGS & P2_FWD) == 0;
    constexpr int R0 = p2_radix(N, FIRST_INV, 0);
    cplx v[SL];
#pragma unroll
    for (int q = 0; q < SL / R0; ++q)
#pragma unroll
      for (int r = 0; r < R0; ++r) v[q * R0 + r] = buf[lay(p2_in_pos<N, FIRST_INV, 0>(t, q, r))];
    if constexpr (!HAS_EX) refill();
    constexpr bool LAST_IS_FWD = (FLAGS & P2_INV) == 0;
    if constexpr ((FLAGS & P2_FWD) != 0) {
      if constexpr (LAST_IS_FWD && HAS_EX)
        p2_fft<N, false, TWS>(v, t, buf, lay, stwf, refill);
      else
        p2_fft<N, false, TWS>(v, t, buf, lay, stwf);
    }
    if constexpr (MUL) {
      // slot (q, r) holds frequency f = p2_in_pos<N, true, 0>(t, q, r): P = s*F*W/|W|, angle(0) = 0 (fftsim.jl:125)
      constexpr int RI = p2_radix(N, true, 0);
#pragma unroll
      for (int q = 0; q < SL / RI; ++q)
#pragma unroll
        for (int r = 0; r < RI; ++r) {
          const double f1 = s * fv[q * RI + r];
          const cplx w = v[q * RI + r];
          const double m2 = w.re * w.re + w.im * w.im;
          if (m2 > 0.0) {
            const double g = f1 * rsqrt(m2);
            v[q * RI + r] = cplx{g * w.re, g * w.im};
          } else {
            v[q * RI + r] = cplx{f1, 0.0};
          }
        }
    }
    if constexpr ((FLAGS & P2_INV) != 0) {
      if constexpr (HAS_EX)
        p2_fft<N, true, TWS>(v, t, buf, lay, stwi, refill);
      else
        p2_fft<N, true, TWS>(v, t, buf, lay, stwi);
    }
    constexpr bool LAST_INV = (FLAGS & P2_INV) != 0;
    constexpr int RO = p2_radix(N, !LAST_INV, 0);
    if (valid) {
#pragma unroll
      for (int q = 0; q < SL / RO; ++q)
#pragma unroll
        for (int r = 0; r < RO; ++r) {
          const int m = p2_in_pos<N, !LAST_INV, 0>(t, q, r);
          st_stream2(reinterpret_cast<double*>(H + base + (long long)m * es), make_double2(v[q * RO + r].re, v[q * RO + r].im));
        }
    }
    if constexpr (!EARLY) __syncthreads();
  }
}

// F in the order the fused forward-multiply-inverse pass consumes it: Fperm[(item * SL + j) * TPU + tid] = F of slot j of thread tid
// of item = o * nball + bx (zero past hx).  One CTA of TPU threads per item, built once per plan: the pass then reads SL contiguous
// runs of TPU doubles per item instead of N strided 64-byte runs.
template <int N, int B>
__global__ void __launch_bounds__(StridedCfg<N, B, P2_FWD | P2_MUL | P2_INV, 1>::THREADS)
    p2_permute_F_kernel(const double* __restrict__ Fh, long long esF, long long other_strideF, int hx, double* __restrict__ Fperm) {
  using C = StridedCfg<N, B, P2_FWD | P2_MUL | P2_INV, 1>;
  constexpr int SL = C::SL, TPU = C::THREADS, RI = p2_radix(N, true, 0);
  const int nball = (hx + B - 1) / B;
  const long long item = blockIdx.x;
  const long long o = item / nball;
  const int bx = (int)(item - o * nball);
  const int tid = threadIdx.x, b = tid % B, t = tid / B;
  const bool valid = bx * B + b < hx;
  const double* fp = Fh + o * other_strideF + (long long)bx * B + b;
  double* out = Fperm + item * (long long)(SL * TPU) + tid;
#pragma unroll
  for (int q = 0; q < SL / RI; ++q)
#pragma unroll
    for (int r = 0; r < RI; ++r) out[(q * RI + r) * TPU] = valid ? fp[(long long)p2_in_pos<N, true, 0>(t, q, r) * esF] : 0.0;
}

// ------------------------------------------------------------------------------------------------
// x-axis passes, even nx = 2*HN (packed real <-> complex).  One item = ROWS consecutive rows.
#ifndef GSP_X_STAGES
#define GSP_X_STAGES 2
#endif
#ifndef GSP_X_REGTW
#define GSP_X_REGTW 1  // forward x pass: per-thread twiddles in registers (two-stage transforms): 61.8 -> 58.7 us at 256^3
#endif
#ifndef GSP_XINV_REGTW
#define GSP_XINV_REGTW 0  // inverse x pass: measured slower with them (51.7 -> 53.3 us, 230 registers)
#endif
template <int HN, bool INV>
struct XCfg {
  static constexpr int STAGES = HN >= 4096 ? 1 : GSP_X_STAGES;  // 2: prefetch the next row group; 1: no prefetch (also: nx = 8192 only fits once)
  static constexpr int NX = 2 * HN, HX = HN + 1;
  static constexpr int SL = p2_slots(HN);
  static constexpr int TPL = HN / SL;                       // threads per row
  static constexpr int ROWS = (TPL >= 128) ? 1 : 128 / TPL; // rows per item
  static constexpr int THREADS = TPL * ROWS;
  static constexpr int SH = p2_log2(p2_radix(HN, INV, 0));  // pad one element every R0: first exchange conflict-free
  static constexpr int ROWLEN = HN + 1 + ((HN + 1) >> SH);
  static constexpr size_t IN_BYTES = INV ? (size_t)ROWS * HX * sizeof(cplx) : (size_t)ROWS * NX * sizeof(double);
  static constexpr size_t EX_BYTES = (size_t)ROWS * ROWLEN * sizeof(cplx);
  // the exchange buffer overlays the (already consumed) input stage: 2 stages + twiddles = 74.5 KB at nx = 256 -> 3 CTAs per SM
  static constexpr size_t STAGE_BYTES = ((IN_BYTES > EX_BYTES ? IN_BYTES : EX_BYTES) + 127) / 128 * 128;
  static constexpr bool STW_OK = (size_t)p2_stw_size(HN, INV) * sizeof(cplx) <= 24 * 1024;
  static constexpr int STW = STW_OK ? p2_stw_size(HN, INV) : 0;  // per-stage twiddle entries (conflict-free layout); 0: use tw with stride 2
  static constexpr size_t TW_BYTES = ((size_t)(NX + STW) * sizeof(cplx) + 127) / 128 * 128;
  static constexpr size_t SMEM = TW_BYTES + STAGES * STAGE_BYTES + 2 * sizeof(mbar_t) + 16;
#ifndef GSP_X_MINB2
#define GSP_X_MINB2 1  // CTAs per SM the register allocation of the 2-stage x kernels must allow
#endif
  static constexpr int MINB = (GSP_P2_R8 && SL == 8 && THREADS <= 128) ? 5 : ((STAGES == 1 && THREADS <= 128) ? 4 : ((THREADS <= 128) ? GSP_X_MINB2 : 1));
};

// noise source of the forward x pass when no array is injected: uniforms from the counter RNG (rng.cuh), generated straight into
// the transform's registers - the 8 N bytes of noise are never written to or read from HBM.  Element pair p of realization
// `real` is Philox(seed, stream 0, real, p), exactly what rng_fill_kernel would have stored at in[2p], in[2p + 1].
struct XRng {
  unsigned long long seed;
  long long first_real;  // realization of row `row_base`
  long long row_base;    // index, within that realization sequence, of the first row of this launch
  long long rows_per_real;
};

template <int HN, bool RNG>
__global__ void __launch_bounds__(XCfg<HN, false>::THREADS, XCfg<HN, false>::MINB) p2_xfwd_kernel(const double* __restrict__ in, cplx* __restrict__ H,
                                                                          const cplx* __restrict__ twg, const cplx* __restrict__ stwg,
                                                                          long long nrows, XRng rng) {
  using C = XCfg<HN, false>;
  constexpr int NX = C::NX, HX = C::HX, SL = C::SL, TPL = C::TPL;
  GSP_DYN_SMEM(smem);
  cplx* tw = reinterpret_cast<cplx*>(smem);  // exp(-2*pi*i*t/NX), t < NX (untangling)
  cplx* stw = tw + NX;                         // per-stage twiddles of the length-HN transform
  unsigned char* stage0 = smem + C::TW_BYTES;
  mbar_t* full = reinterpret_cast<mbar_t*>(stage0 + C::STAGES * C::STAGE_BYTES);
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_mbar_init();
  }
  for (int i = tid; i < NX; i += C::THREADS) tw[i] = twg[i];
  for (int i = tid; i < C::STW; i += C::THREADS) stw[i] = stwg[i];
  __syncthreads();
  const long long ngroups = (nrows + C::ROWS - 1) / C::ROWS;
  auto issue = [&](long long g, int sg) {
    if (!RNG && tid == 0) {
      fence_proxy_async();  // see p2_strided_kernel: generic-proxy accesses of the stage precede the async-proxy fill
      const long long r0 = g * C::ROWS;
      const long long nv = (nrows - r0 < C::ROWS) ? nrows - r0 : C::ROWS;
      const uint32_t bytes = (uint32_t)(nv * NX * sizeof(double));
      mbar_arrive_expect_tx(&full[sg], bytes);
      bulk_g2s(stage0 + (size_t)sg * C::STAGE_BYTES, in + r0 * NX, bytes, &full[sg]);
    }
  };
  const int rl = tid / TPL, t = tid - rl * TPL;
  const RowLay<C::SH> lay{rl * C::ROWLEN};
  // this thread's twiddles, once for all its items: stage 1 of the length-HN transform and the untangling factors w^f
  constexpr bool REGTW = P2RegTw<HN>::OK && C::STW_OK && GSP_X_REGTW != 0;
  cplx rtw[P2RegTw<HN>::NF];
  cplx utw[REGTW ? SL : 1];
  if constexpr (REGTW) {
    p2_load_regtw<HN, false>(rtw, t, stw);
    constexpr int RIu = p2_radix(HN, true, 0);
#pragma unroll
    for (int q = 0; q < SL / RIu; ++q)
#pragma unroll
      for (int r = 0; r < RIu; ++r) utw[q * RIu + r] = tw[p2_in_pos<HN, true, 0>(t, q, r)];
  }
  long long g = blockIdx.x;
  if (C::STAGES == 2 && g < ngroups) issue(g, 0);
  for (int it = 0; g < ngroups; g += gridDim.x, ++it) {
    const int cur = C::STAGES == 2 ? (it & 1) : 0;
    if (C::STAGES == 2) {
      if (g + gridDim.x < ngroups) issue(g + gridDim.x, cur ^ 1);
    } else {
      issue(g, 0);
    }
    const long long row = g * C::ROWS + rl;
    const bool valid = row < nrows;
    cplx* ex = reinterpret_cast<cplx*>(stage0 + (size_t)cur * C::STAGE_BYTES);  // overlays the input once it is in registers
    const cplx* src = ex + (size_t)rl * HN;
    constexpr int R0 = p2_radix(HN, false, 0);
    cplx v[SL];
    if constexpr (RNG) {
      const long long grow = rng.row_base + row;
      const long long rr = grow / rng.rows_per_real;
      const unsigned long long real = (unsigned long long)(rng.first_real + rr);
      const unsigned long long pair0 = (unsigned long long)(grow - rr * rng.rows_per_real) * HN;
#pragma unroll
      for (int q = 0; q < SL / R0; ++q)
#pragma unroll
        for (int r = 0; r < R0; ++r) {
          double u0 = 0.0, u1 = 0.0;
          if (valid) philox_uniform2(rng.seed, 0u, real, pair0 + (unsigned long long)p2_in_pos<HN, false, 0>(t, q, r), u0, u1);
          v[q * R0 + r] = cplx{u0, u1};
        }
    } else {
      mbar_wait(&full[cur], (uint32_t)(C::STAGES == 2 ? ((it >> 1) & 1) : (it & 1)));
#pragma unroll
      for (int q = 0; q < SL / R0; ++q)
#pragma unroll
        for (int r = 0; r < R0; ++r) v[q * R0 + r] = src[p2_in_pos<HN, false, 0>(t, q, r)];
    }
    if constexpr (REGTW)
      p2_fft_regtw<HN, false>(v, t, ex, lay, rtw);
    else
      p2_fft<HN, false, C::STW_OK ? 0 : 2>(v, t, ex, lay, C::STW_OK ? stw : tw);
    // untangle: X[f] = E + w^f * O with E = (Z[f] + conj Z[h-f])/2, O = -i (Z[f] - conj Z[h-f])/2
    constexpr int RI = p2_radix(HN, true, 0);
    __syncthreads();
#pragma unroll
    for (int q = 0; q < SL / RI; ++q)
#pragma unroll
      for (int r = 0; r < RI; ++r) ex[lay(p2_in_pos<HN, true, 0>(t, q, r))] = v[q * RI + r];
    __syncthreads();
    if (valid) {
      cplx* dst = H + row * HX;
#pragma unroll
      for (int q = 0; q < SL / RI; ++q)
#pragma unroll
        for (int r = 0; r < RI; ++r) {
          const int f = p2_in_pos<HN, true, 0>(t, q, r);
          const cplx zk = v[q * RI + r];
          const cplx zc = cconj(ex[lay((HN - f) & (HN - 1))]);
          const cplx e = cplx{0.5 * (zk.re + zc.re), 0.5 * (zk.im + zc.im)};
          const cplx d = csub(zk, zc);
          const cplx od = cplx{0.5 * d.im, -0.5 * d.re};
          const cplx o = cadd(e, cmul(REGTW ? utw[REGTW ? q * RI + r : 0] : tw[f], od));
          st_stream2(reinterpret_cast<double*>(dst + f), make_double2(o.re, o.im));
          if (f == 0) st_stream2(reinterpret_cast<double*>(dst + HN), make_double2(zk.re - zk.im, 0.0));
        }
    }
    __syncthreads();
  }
}

// x-axis inverse pass: half spectrum rows -> real rows; out = scale * (unnormalised inverse DFT) + mu
template <int HN>
__global__ void __launch_bounds__(XCfg<HN, true>::THREADS, XCfg<HN, true>::MINB) p2_xinv_kernel(const cplx* __restrict__ H, double* __restrict__ out,
                                                                         const cplx* __restrict__ twg, const cplx* __restrict__ stwg,
                                                                         long long nrows, double scale, double mu) {
  using C = XCfg<HN, true>;
  constexpr int NX = C::NX, HX = C::HX, SL = C::SL, TPL = C::TPL;
  GSP_DYN_SMEM(smem);
  cplx* tw = reinterpret_cast<cplx*>(smem);
  cplx* stw = tw + NX;
  unsigned char* stage0 = smem + C::TW_BYTES;
  mbar_t* full = reinterpret_cast<mbar_t*>(stage0 + C::STAGES * C::STAGE_BYTES);
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_mbar_init();
  }
  for (int i = tid; i < NX; i += C::THREADS) tw[i] = twg[i];
  for (int i = tid; i < C::STW; i += C::THREADS) stw[i] = stwg[i];
  __syncthreads();
  const long long ngroups = (nrows + C::ROWS - 1) / C::ROWS;
  auto issue = [&](long long g, int sg) {
    if (tid == 0) {
      fence_proxy_async();
      const long long r0 = g * C::ROWS;
      const long long nv = (nrows - r0 < C::ROWS) ? nrows - r0 : C::ROWS;
      const uint32_t bytes = (uint32_t)(nv * HX * sizeof(cplx));
      mbar_arrive_expect_tx(&full[sg], bytes);
      bulk_g2s(stage0 + (size_t)sg * C::STAGE_BYTES, H + r0 * HX, bytes, &full[sg]);
    }
  };
  const int rl = tid / TPL, t = tid - rl * TPL;
  const RowLay<C::SH> lay{rl * C::ROWLEN};
  constexpr bool REGTW = P2RegTw<HN>::OK && C::STW_OK && GSP_XINV_REGTW != 0;  // see p2_xfwd_kernel
  cplx rtw[P2RegTw<HN>::NI];
  cplx utw[REGTW ? SL : 1];
  if constexpr (REGTW) {
    p2_load_regtw<HN, true>(rtw, t, stw);
    constexpr int R0u = p2_radix(HN, true, 0);
#pragma unroll
    for (int q = 0; q < SL / R0u; ++q)
#pragma unroll
      for (int r = 0; r < R0u; ++r) utw[q * R0u + r] = cconj(tw[p2_in_pos<HN, true, 0>(t, q, r)]);
  }
  long long g = blockIdx.x;
  if (C::STAGES == 2 && g < ngroups) issue(g, 0);
  for (int it = 0; g < ngroups; g += gridDim.x, ++it) {
    const int cur = C::STAGES == 2 ? (it & 1) : 0;
    if (C::STAGES == 2) {
      if (g + gridDim.x < ngroups) issue(g + gridDim.x, cur ^ 1);
    } else {
      issue(g, 0);
    }
    const long long row = g * C::ROWS + rl;
    const bool valid = row < nrows;
    cplx* ex = reinterpret_cast<cplx*>(stage0 + (size_t)cur * C::STAGE_BYTES);  // overlays the input once it is in registers
    const cplx* X = ex + (size_t)rl * HX;
    mbar_wait(&full[cur], (uint32_t)(C::STAGES == 2 ? ((it >> 1) & 1) : (it & 1)));
    constexpr int R0 = p2_radix(HN, true, 0);
    cplx v[SL];
#pragma unroll
    for (int q = 0; q < SL / R0; ++q)
#pragma unroll
      for (int r = 0; r < R0; ++r) {
        const int m = p2_in_pos<HN, true, 0>(t, q, r);
        const cplx xk = X[m];
        const cplx xc = cconj(X[HN - m]);
        const cplx sm = cadd(xk, xc);
        const cplx d = csub(xk, xc);
        const cplx tt = cmul(REGTW ? utw[REGTW ? q * R0 + r : 0] : cconj(tw[m]), d);
        v[q * R0 + r] = cplx{sm.re - tt.im, sm.im + tt.re};
      }
    if constexpr (REGTW)
      p2_fft_regtw<HN, true>(v, t, ex, lay, rtw);
    else
      p2_fft<HN, true, C::STW_OK ? 0 : 2>(v, t, ex, lay, C::STW_OK ? stw : tw);
    constexpr int RO = p2_radix(HN, false, 0);
    if (valid) {
      double* dst = out + row * NX;
#pragma unroll
      for (int q = 0; q < SL / RO; ++q)
#pragma unroll
        for (int r = 0; r < RO; ++r) {
          const int j = p2_in_pos<HN, false, 0>(t, q, r);
          st_stream2(dst + 2 * j, make_double2(v[q * RO + r].re * scale + mu, v[q * RO + r].im * scale + mu));
        }
    }
    __syncthreads();
  }
}

}  // namespace gsp
