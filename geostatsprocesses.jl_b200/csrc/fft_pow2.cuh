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

namespace gsp {

GSP_HD constexpr int p2_stages(int N) { return N <= 16 ? 1 : ((N == 512 || N >= 1024) ? 3 : 2); }
GSP_HD constexpr int p2_radix_fwd(int N, int s) {
  switch (N) {
    case 2: return 2;
    case 4: return 4;
    case 8: return 8;
    case 16: return 16;
    case 32: return s == 0 ? 16 : 2;
    case 64: return 8;
    case 128: return s == 0 ? 16 : 8;
    case 256: return 16;
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

// twiddle + butterflies of stage s on the register slots.  tw[i * TWS] = exp(-2*pi*i*i/N) (shared memory)
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
        cplx w = tw[(r * k) * (N / (Ns * R)) * TWS];
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

template <int N, bool INV, int S_, int TWS, class Lay>
GSP_DEV void p2_run_from(cplx* v, int t, cplx* buf, Lay lay, const cplx* tw) {
  p2_stage<N, INV, S_, TWS>(v, t, tw);
  if constexpr (S_ + 1 < p2_stages(N)) {
    p2_exchange<N, INV, S_>(v, t, buf, lay);
    p2_run_from<N, INV, S_ + 1, TWS>(v, t, buf, lay, tw);
  }
}

// Full transform of one line held in registers.
//   in : slot (q, r) = x[p2_in_pos<N,INV,0>(t,q,r)]
//   out: slot (q, r) = X[p2_in_pos<N,!INV,0>(t,q,r)]  (natural order; same mapping as the opposite direction's input)
template <int N, bool INV, int TWS, class Lay>
GSP_DEV void p2_fft(cplx* v, int t, cplx* buf, Lay lay, const cplx* tw) {
  p2_run_from<N, INV, 0, TWS>(v, t, buf, lay, tw);
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

GSP_HD constexpr int p2_bundle(int N) { return N <= 512 ? 8 : (N <= 2048 ? 4 : 2); }
GSP_HD constexpr int p2_units(int N) {
  return ((N / p2_slots(N)) * p2_bundle(N) >= 256) ? 1 : 256 / ((N / p2_slots(N)) * p2_bundle(N));
}

// ------------------------------------------------------------------------------------------------
// strided pass (y or z axis) over the half spectrum, in place.  One "unit" = B adjacent kx times the
// whole line; a CTA holds U units.  FLAGS: FWD only / INV only / FWD|MUL|INV (last axis, fused).
template <int N, int B, int U, int FLAGS>
__global__ void __launch_bounds__((N / p2_slots(N)) * B * U) p2_strided_kernel(cplx* __restrict__ H, const cplx* __restrict__ twg, long long es,
                                                                              int hx, int nbundles, long long nunits, long long other_stride,
                                                                              const double* __restrict__ Fh, double s) {
  constexpr int SL = p2_slots(N);
  constexpr int TPL = N / SL;
  constexpr int TPU = TPL * B;
  GSP_DYN_SMEM(smem);
  cplx* tw = reinterpret_cast<cplx*>(smem);
  cplx* bufs = tw + N;
  for (int i = threadIdx.x; i < N; i += TPU * U) tw[i] = twg[i];
  const int u = threadIdx.x / TPU;
  const int lt = threadIdx.x - u * TPU;
  const int b = lt % B, t = lt / B;
  cplx* buf = bufs + (size_t)u * N * B;
  const long long unit = (long long)blockIdx.x * U + u;
  const bool live = unit < nunits;
  const long long o = live ? unit / nbundles : 0;
  const int bx = live ? (int)(unit - o * nbundles) : 0;
  const bool valid = live && (bx * B + b < hx);
  const long long base = o * other_stride + (long long)bx * B + b;
  const BundleLay lay{B, b};
  constexpr bool FIRST_INV = (FLAGS & P2_FWD) == 0;
  constexpr int R0 = p2_radix(N, FIRST_INV, 0);
  cplx v[SL];
#pragma unroll
  for (int q = 0; q < SL / R0; ++q)
#pragma unroll
    for (int r = 0; r < R0; ++r) {
      const int m = p2_in_pos<N, FIRST_INV, 0>(t, q, r);
      cplx x{0.0, 0.0};
      if (valid) {
        const double2 d = ld_stream2(reinterpret_cast<const double*>(H + base + (long long)m * es));
        x = cplx{d.x, d.y};
      }
      v[q * R0 + r] = x;
    }
  __syncthreads();  // twiddle table ready
  if constexpr ((FLAGS & P2_FWD) != 0) p2_fft<N, false, 1>(v, t, buf, lay, tw);
  if constexpr ((FLAGS & P2_MUL) != 0) {
    // slot (q, r) now holds frequency f = p2_in_pos<N, true, 0>(t, q, r)
    constexpr int RI = p2_radix(N, true, 0);
#pragma unroll
    for (int q = 0; q < SL / RI; ++q)
#pragma unroll
      for (int r = 0; r < RI; ++r) {
        const int f = p2_in_pos<N, true, 0>(t, q, r);
        const double fv = valid ? s * Fh[base + (long long)f * es] : 0.0;
        const cplx w = v[q * RI + r];
        const double m2 = w.re * w.re + w.im * w.im;
        if (m2 > 0.0) {
          const double g = fv * rsqrt(m2);
          v[q * RI + r] = cplx{g * w.re, g * w.im};
        } else {
          v[q * RI + r] = cplx{fv, 0.0};
        }
      }
  }
  if constexpr ((FLAGS & P2_INV) != 0) p2_fft<N, true, 1>(v, t, buf, lay, tw);
  // results: slot (q, r) of the LAST executed direction's opposite input mapping
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
}

// ------------------------------------------------------------------------------------------------
// x-axis forward pass, even nx = 2*HN: real rows -> half spectrum rows (packed real-to-complex).
template <int HN, bool INV>
struct XCfg {
  static constexpr int SL = p2_slots(HN);
  static constexpr int TPL = HN / SL;                       // threads per row
  static constexpr int ROWS = (TPL >= 256) ? 1 : 256 / TPL; // rows per CTA
  static constexpr int THREADS = TPL * ROWS;
  static constexpr int SH = p2_log2(p2_radix(HN, INV, 0));  // pad one element every R0: first exchange conflict-free
  static constexpr int ROWLEN = HN + 1 + ((HN + 1) >> SH) + 1;  // padded row (holds k = 0..HN)
  static constexpr size_t SMEM = (size_t)(2 * HN + ROWS * ROWLEN) * sizeof(cplx);
};

template <int HN>
__global__ void __launch_bounds__(XCfg<HN, false>::THREADS) p2_xfwd_kernel(const double* __restrict__ in, cplx* __restrict__ H,
                                                                   const cplx* __restrict__ twg, long long nrows) {
  using C = XCfg<HN, false>;
  constexpr int NX = 2 * HN, HX = HN + 1, SL = C::SL, TPL = C::TPL;
  GSP_DYN_SMEM(smem);
  cplx* tw = reinterpret_cast<cplx*>(smem);  // exp(-2*pi*i*t/NX), t < NX
  cplx* bufs = tw + NX;
  for (int i = threadIdx.x; i < NX; i += C::THREADS) tw[i] = twg[i];
  const int rl = threadIdx.x / TPL, t = threadIdx.x - rl * TPL;
  const long long row = (long long)blockIdx.x * C::ROWS + rl;
  const bool valid = row < nrows;
  const RowLay<C::SH> lay{rl * C::ROWLEN};
  constexpr int R0 = p2_radix(HN, false, 0);
  cplx v[SL];
  const double* src = in + row * NX;
#pragma unroll
  for (int q = 0; q < SL / R0; ++q)
#pragma unroll
    for (int r = 0; r < R0; ++r) {
      const int m = p2_in_pos<HN, false, 0>(t, q, r);
      cplx x{0.0, 0.0};
      if (valid) {
        const double2 d = ld_stream2(src + 2 * m);
        x = cplx{d.x, d.y};
      }
      v[q * R0 + r] = x;
    }
  __syncthreads();
  p2_fft<HN, false, 2>(v, t, bufs, lay, tw);
  // untangle: X[f] = E + w^f * O with E = (Z[f] + conj Z[h-f])/2, O = -i (Z[f] - conj Z[h-f])/2
  constexpr int RI = p2_radix(HN, true, 0);
  __syncthreads();
#pragma unroll
  for (int q = 0; q < SL / RI; ++q)
#pragma unroll
    for (int r = 0; r < RI; ++r) bufs[lay(p2_in_pos<HN, true, 0>(t, q, r))] = v[q * RI + r];
  __syncthreads();
  if (valid) {
    cplx* dst = H + row * HX;
#pragma unroll
    for (int q = 0; q < SL / RI; ++q)
#pragma unroll
      for (int r = 0; r < RI; ++r) {
        const int f = p2_in_pos<HN, true, 0>(t, q, r);
        const cplx zk = v[q * RI + r];
        const cplx zc = cconj(bufs[lay((HN - f) & (HN - 1))]);
        const cplx e = cplx{0.5 * (zk.re + zc.re), 0.5 * (zk.im + zc.im)};
        const cplx d = csub(zk, zc);
        const cplx od = cplx{0.5 * d.im, -0.5 * d.re};
        const cplx o = cadd(e, cmul(tw[f], od));
        st_stream2(reinterpret_cast<double*>(dst + f), make_double2(o.re, o.im));
        if (f == 0) st_stream2(reinterpret_cast<double*>(dst + HN), make_double2(zk.re - zk.im, 0.0));
      }
  }
}

// x-axis inverse pass: half spectrum rows -> real rows; out = scale * (unnormalised inverse DFT) + mu
template <int HN>
__global__ void __launch_bounds__(XCfg<HN, true>::THREADS) p2_xinv_kernel(const cplx* __restrict__ H, double* __restrict__ out,
                                                                   const cplx* __restrict__ twg, long long nrows, double scale, double mu) {
  using C = XCfg<HN, true>;
  constexpr int NX = 2 * HN, HX = HN + 1, SL = C::SL, TPL = C::TPL;
  GSP_DYN_SMEM(smem);
  cplx* tw = reinterpret_cast<cplx*>(smem);
  cplx* bufs = tw + NX;
  for (int i = threadIdx.x; i < NX; i += C::THREADS) tw[i] = twg[i];
  const int rl = threadIdx.x / TPL, t = threadIdx.x - rl * TPL;
  const long long row = (long long)blockIdx.x * C::ROWS + rl;
  const bool valid = row < nrows;
  const RowLay<C::SH> lay{rl * C::ROWLEN};
  const cplx* src = H + row * HX;
  // stage the row (k = 0..HN) in shared memory: the pre-processing pairs k with HN-k
  for (int k = t; k < HX; k += TPL) {
    cplx x{0.0, 0.0};
    if (valid) {
      const double2 d = ld_stream2(reinterpret_cast<const double*>(src + k));
      x = cplx{d.x, d.y};
    }
    bufs[lay(k)] = x;
  }
  __syncthreads();
  constexpr int R0 = p2_radix(HN, true, 0);
  cplx v[SL];
#pragma unroll
  for (int q = 0; q < SL / R0; ++q)
#pragma unroll
    for (int r = 0; r < R0; ++r) {
      const int m = p2_in_pos<HN, true, 0>(t, q, r);
      const cplx xk = bufs[lay(m)];
      const cplx xc = cconj(bufs[lay(HN - m)]);
      const cplx sm = cadd(xk, xc);
      const cplx d = csub(xk, xc);
      const cplx tt = cmul(cconj(tw[m]), d);
      v[q * R0 + r] = cplx{sm.re - tt.im, sm.im + tt.re};
    }
  p2_fft<HN, true, 2>(v, t, bufs, lay, tw);
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
}

}  // namespace gsp
