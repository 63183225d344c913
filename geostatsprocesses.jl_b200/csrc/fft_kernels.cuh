// Radix-staged FP64 FFT building blocks for FFTSIM: replaces FFTW's fft/ifft at
// fftsim.jl:90,125,128 together with the elementwise passes fused around them.
//
// A "line bundle" is B lines of length n held in shared memory as buf[j*B + b] (j = position on
// the line, b = line within the bundle).  Stages are Stockham autosort (ping-pong between two
// buffers), radix 16/8/4/2 done entirely in registers (recursive DIT with literal twiddles) and
// radix 3/5/7/11/13 by a small O(R^2) DFT; any other prime factor p of the extent becomes a stage of radix p evaluated one output at
// a time (O(p) each, fft_stage_generic) - slower, but every extent FFTW accepts in the reference works here too.
// Inter-stage twiddles come from a table exp(-2*pi*i*t/n) computed on the host in long double.
// Lanes run fastest over b, so every shared-memory access of a stage is a contiguous run of 16-byte
// elements; x-axis passes use an odd B which also makes the global<->shared transposition
// conflict-free.
#pragma once
#include "common.h"

namespace gsp {

constexpr int FFT_MAX_STAGES = 16;

struct LinePlan {
  int n;                       // complex transform length
  int nst;
  int radix[FFT_MAX_STAGES];
  const cplx* tw;              // tw[t * tw_stride] = exp(-2*pi*i*t/n)
  int tw_stride;
};

// ---- literal twiddles for the in-register power-of-two butterflies: exp(-2*pi*i*k/16), k = 0..7
GSP_DEV double c16(int k) {
  switch (k) {
    case 0: return 1.0;
    case 1: return 0.92387953251128674;
    case 2: return 0.70710678118654752;
    case 3: return 0.38268343236508977;
    case 4: return 0.0;
    case 5: return -0.38268343236508977;
    case 6: return -0.70710678118654752;
    default: return -0.92387953251128674;
  }
}
GSP_DEV double s16(int k) {
  switch (k) {
    case 0: return 0.0;
    case 1: return 0.38268343236508977;
    case 2: return 0.70710678118654752;
    case 3: return 0.92387953251128674;
    case 4: return 1.0;
    case 5: return 0.92387953251128674;
    case 6: return 0.70710678118654752;
    default: return 0.38268343236508977;
  }
}

// a * exp(-/+ 2*pi*i*K/R), K < R/2, R in {2,4,8,16}
template <int R, int K, bool INV>
GSP_DEV cplx mul_w(cplx a) {
  constexpr int idx = K * (16 / R);
  if constexpr (idx == 0) {
    return a;
  } else if constexpr (idx == 4) {
    return INV ? cplx{-a.im, a.re} : cplx{a.im, -a.re};
  } else {
    const double c = c16(idx), s = s16(idx);
    return INV ? cplx{a.re * c - a.im * s, a.im * c + a.re * s} : cplx{a.re * c + a.im * s, a.im * c - a.re * s};
  }
}

template <int R, bool INV, int K>
GSP_DEV void bfly_rec(const cplx* e, const cplx* o, cplx* v) {
  if constexpr (K < R / 2) {
    const cplx t = mul_w<R, K, INV>(o[K]);
    v[K] = cadd(e[K], t);
    v[K + R / 2] = csub(e[K], t);
    bfly_rec<R, INV, K + 1>(e, o, v);
  }
}

// in-register DFT of size R (power of two), natural order in and out
template <int R, bool INV>
GSP_DEV void dft_pow2(cplx* v) {
  if constexpr (R == 2) {
    const cplx a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  } else {
    cplx e[R / 2], o[R / 2];
#pragma unroll
    for (int k = 0; k < R / 2; ++k) {
      e[k] = v[2 * k];
      o[k] = v[2 * k + 1];
    }
    dft_pow2<R / 2, INV>(e);
    dft_pow2<R / 2, INV>(o);
    bfly_rec<R, INV, 0>(e, o, v);
  }
}

// small odd-radix DFT through the twiddle table: w_R^m = tw[m * step], step = (n/R)*tw_stride
template <int R, bool INV>
GSP_DEV void dft_odd(cplx* v, const cplx* __restrict__ tw, int step) {
  cplx out[R];
#pragma unroll
  for (int q = 0; q < R; ++q) {
    cplx acc = v[0];
#pragma unroll
    for (int r = 1; r < R; ++r) {
      cplx w = tw[((r * q) % R) * step];
      if (INV) w.im = -w.im;
      acc = cadd(acc, cmul(v[r], w));
    }
    out[q] = acc;
  }
#pragma unroll
  for (int q = 0; q < R; ++q) v[q] = out[q];
}

template <int R, bool INV>
GSP_DEV void dft_any(cplx* v, const cplx* __restrict__ tw, int step) {
  if constexpr (R == 2 || R == 4 || R == 8 || R == 16)
    dft_pow2<R, INV>(v);
  else
    dft_odd<R, INV>(v, tw, step);
}

// one Stockham stage of radix R over a bundle: in -> out.  Ns = product of earlier radices.
template <int R, bool INV>
GSP_DEV void fft_stage(const cplx* __restrict__ in, cplx* __restrict__ out, int n, int Ns, int B, const cplx* __restrict__ tw,
                       int tw_stride) {
  const int T = n / R;
  const int items = T * B;
  const int twk = (n / (Ns * R)) * tw_stride;  // table step of exp(-2*pi*i/(Ns*R))
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int j = it / B, b = it - j * B;
    const int k = j % Ns;
    cplx v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = in[(j + r * T) * B + b];
    if (Ns > 1) {
#pragma unroll
      for (int r = 1; r < R; ++r) {
        cplx w = tw[(r * k) * twk];
        if (INV) w.im = -w.im;
        v[r] = cmul(v[r], w);
      }
    }
    dft_any<R, INV>(v, tw, (n / R) * tw_stride);
    const int j0 = (j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) out[(j0 + r * Ns) * B + b] = v[r];
  }
}

// Stockham stage of ANY radix R (prime factors > 13): one output per work item.  With m = k + r*Ns (k = j mod Ns) the twiddled
// butterfly collapses to y[(j-k)*R + m] = sum_q x[j + q*n/R] * w_M^(q*m), M = Ns*R, and w_M = tw[(n/M) * tw_stride]; the exponent
// q*m mod M is carried incrementally in integers, so the accuracy is that of the table.
template <bool INV>
GSP_DEV void fft_stage_generic(const cplx* __restrict__ in, cplx* __restrict__ out, int n, int R, int Ns, int B, const cplx* __restrict__ tw,
                               int tw_stride) {
  const int T = n / R;
  const int M = Ns * R;
  const int twm = (n / M) * tw_stride;
  const int items = n * B;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int o = it / B, b = it - o * B;
    const int r = o / T, j = o - r * T;
    const int k = j % Ns;
    const int m = k + r * Ns;
    double are = 0.0, aim = 0.0;
    int idx = 0;
    for (int q = 0; q < R; ++q) {
      cplx w = tw[idx * twm];
      if (INV) w.im = -w.im;
      const cplx x = in[(j + q * T) * B + b];
      are += x.re * w.re - x.im * w.im;
      aim += x.re * w.im + x.im * w.re;
      idx += m;
      if (idx >= M) idx -= M;
    }
    out[((j - k) * R + m) * B + b] = cplx{are, aim};
  }
}

// full transform of a bundle; returns the buffer holding the result (a or b).  Ends synchronised.
template <bool INV>
GSP_DEV cplx* fft_bundle(const LinePlan& lp, cplx* a, cplx* b, int B) {
  int Ns = 1;
  cplx* in = a;
  cplx* out = b;
  for (int s = 0; s < lp.nst; ++s) {
    const int R = lp.radix[s];
    switch (R) {
      case 16: fft_stage<16, INV>(in, out, lp.n, Ns, B, lp.tw, lp.tw_stride); break;
      case 8: fft_stage<8, INV>(in, out, lp.n, Ns, B, lp.tw, lp.tw_stride); break;
      case 4: fft_stage<4, INV>(in, out, lp.n, Ns, B, lp.tw, lp.tw_stride); break;
      case 2: fft_stage<2, INV>(in, out, lp.n, Ns, B, lp.tw, lp.tw_stride); break;
      case 3: fft_stage<3, INV>(in, out, lp.n, Ns, B, lp.tw, lp.tw_stride); break;
      case 5: fft_stage<5, INV>(in, out, lp.n, Ns, B, lp.tw, lp.tw_stride); break;
      case 7: fft_stage<7, INV>(in, out, lp.n, Ns, B, lp.tw, lp.tw_stride); break;
      case 11: fft_stage<11, INV>(in, out, lp.n, Ns, B, lp.tw, lp.tw_stride); break;
      case 13: fft_stage<13, INV>(in, out, lp.n, Ns, B, lp.tw, lp.tw_stride); break;
      default: fft_stage_generic<INV>(in, out, lp.n, R, Ns, B, lp.tw, lp.tw_stride); break;
    }
    Ns *= R;
    __syncthreads();
    cplx* t = in;
    in = out;
    out = t;
  }
  return in;
}

}  // namespace gsp
