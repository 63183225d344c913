// Counter-based on-device noise (Philox4x32-10).  Replaces randn (lusim.jl:160) / rand (fftsim.jl:124)
// in throughput mode; parity runs inject the reference's noise arrays instead.
// Counter = (pair index lo, pair index hi, realization, stream), key = seed: the value of element e
// of realization r does not depend on how realizations are sharded over GPUs or chunks.
#pragma once
#include "common.h"

namespace gsp {

GSP_DEV uint4 philox4x32_10(uint4 c, uint2 k) {
  const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const unsigned hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

// two uniforms in [0,1) with 53 random bits each
GSP_DEV void philox_uniform2(unsigned long long seed, unsigned stream, unsigned long long real, unsigned long long pair,
                             double& u0, double& u1) {
  // realization index folded as (real_lo in c.z, real_hi ^ stream<<16 in c.w)
  uint4 c = make_uint4((unsigned)pair, (unsigned)(pair >> 32), (unsigned)real, ((unsigned)(real >> 32) & 0xffffu) | (stream << 16));
  uint2 k = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
  const uint4 o = philox4x32_10(c, k);
  const unsigned long long a = ((unsigned long long)o.y << 32) | o.x;
  const unsigned long long b = ((unsigned long long)o.w << 32) | o.z;
  u0 = (double)(a >> 11) * (1.0 / 9007199254740992.0);
  u1 = (double)(b >> 11) * (1.0 / 9007199254740992.0);
}

// two standard normals (Box-Muller on the pair above)
GSP_DEV void philox_normal2(unsigned long long seed, unsigned stream, unsigned long long real, unsigned long long pair,
                            double& n0, double& n1) {
  double u0, u1;
  philox_uniform2(seed, stream, real, pair, u0, u1);
  const double r = sqrt(-2.0 * log(1.0 - u0));  // 1-u0 in (0,1]
  double s, c;
  sincospi(2.0 * u1, &s, &c);
  n0 = r * c;
  n1 = r * s;
}

// tr == 0: out[e + r*ld];  tr != 0 (realization-major): out[r + e*ld];  e < n, r < R;  normal != 0 -> N(0,1) else U[0,1)
static __global__ void __launch_bounds__(256) rng_fill_kernel(double* __restrict__ out, long long n, long long ld, long long R,
                                                       unsigned long long seed, unsigned stream, unsigned long long first_real,
                                                       int normal, int tr) {
  const long long npairs = (n + 1) / 2;
  const long long total = npairs * R;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const long long r = tr ? t % R : t / npairs;
    const long long p = tr ? t / R : t - r * npairs;
    double a, b;
    if (normal)
      philox_normal2(seed, stream, first_real + (unsigned long long)r, (unsigned long long)p, a, b);
    else
      philox_uniform2(seed, stream, first_real + (unsigned long long)r, (unsigned long long)p, a, b);
    if (tr) {
      double* o = out + r + 2 * p * ld;
      o[0] = a;
      if (2 * p + 1 < n) o[ld] = b;
    } else {
      double* o = out + r * ld + 2 * p;
      o[0] = a;
      if (2 * p + 1 < n) o[1] = b;
    }
  }
}

inline cudaError_t launch_rng_fill(cudaStream_t st, int sms, double* out, long long n, long long ld, long long R,
                                   unsigned long long seed, unsigned stream, unsigned long long first_real, bool normal,
                                   bool transposed = false) {
  long long total = ((n + 1) / 2) * R;
  if (total <= 0) return cudaSuccess;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)sms * 32) blocks = (long long)sms * 32;
  ProfScope prof_("rng_fill", st);
  GSP_LAUNCH(rng_fill_kernel, dim3((unsigned)blocks), dim3(256), 0, st, out, n, ld, R, seed, stream, first_real, normal ? 1 : 0,
             transposed ? 1 : 0);
  g_launches++;
  return cudaGetLastError();
}

}  // namespace gsp
