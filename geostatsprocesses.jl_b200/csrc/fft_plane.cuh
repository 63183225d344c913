// Fused x+y passes of the 3-D FFTSIM pipeline: one persistent kernel transforms every z-plane along
// x AND y, exchanging the intermediate half-spectrum plane through L2 instead of HBM.
//
//   forward  (INV = false): x items (ROWS real rows -> half-spectrum rows) then y items (B kx columns, in place)
//   inverse  (INV = true ): y items (in place)                              then x items (rows -> real field, +mu)
//
// Work items are numbered plane by plane, the dependent kind lagging `lag` planes behind the producing kind
// (lag is chosen by the host to exceed the window of items in flight, so a dependency is almost always already
// satisfied when its item is claimed).  Items are claimed DYNAMICALLY, in order, from a global counter: every
// dependency points to a lower item number, the lowest unfinished item never waits, so the walk cannot deadlock
// whatever the residency of the grid (several lanes = several of these kernels share the SMs).  A producing
// item bumps its plane's counter (release) when its stores are done; a dependent item is only fetched (TMA /
// bulk copy) once its plane counter has reached the number of producing items of a plane (acquire).
// The plane written by the producers (<= 0.6 MB) is still in the 126 MB L2 when the consumers read it:
// HBM sees one read of the input and one write of the result per plane - 2 passes per transform, not 3.
//
// Second version (round 1, session 3).  The first one (2 shared-memory stages + separate exchange buffer =
// 107 KB, 2 CTAs = 8 warps per SM, strided twiddle tables, static item ownership) halved the DRAM traffic but
// ran slower than the separate passes (profiles/r01_fused_plane_notes.md).  This one follows what made the
// stand-alone passes fast: ONE stage that the exchange buffer overlays (44 KB -> 4 CTAs = 16 warps per SM
// overlap each other's load / transform / store phases), conflict-free per-stage twiddle tables, and the
// Philox noise generated straight into the x items' registers when no noise array is injected.
#pragma once
#include "fft_pow2.cuh"

namespace gsp {

constexpr int PLANE_THREADS = 128;

template <int HN, int NY, bool INV>
struct PlaneCfg {
  using XC = XCfg<HN, INV>;
  static constexpr int NX = 2 * HN, HX = HN + 1;
  static constexpr int B = p2_bundle(NY);
  static constexpr int SLY = p2_slots(NY), TPLY = NY / SLY, TPU = TPLY * B;
  static constexpr int U = (TPU <= PLANE_THREADS) ? PLANE_THREADS / TPU : 1;  // kx bundles per y item
  static constexpr bool OK = XC::THREADS == PLANE_THREADS && TPU <= PLANE_THREADS && (PLANE_THREADS % TPU) == 0 && NY <= 256 &&
                             (NY % XC::ROWS) == 0 && XC::STW_OK;
  static constexpr int STWX = p2_stw_size(HN, INV), STWY = p2_stw_size(NY, INV);
  static constexpr size_t XIN = XC::IN_BYTES;
  static constexpr size_t YIN = (size_t)U * NY * B * sizeof(cplx);
  static constexpr size_t STAGE_RAW = XIN > YIN ? (XIN > XC::EX_BYTES ? XIN : XC::EX_BYTES) : (YIN > XC::EX_BYTES ? YIN : XC::EX_BYTES);
  static constexpr size_t STAGE = (STAGE_RAW + 127) / 128 * 128;
  static constexpr size_t TW_BYTES = ((size_t)(NX + STWX + STWY) * sizeof(cplx) + 127) / 128 * 128;
  static constexpr size_t SMEM = TW_BYTES + STAGE + sizeof(mbar_t) + 2 * sizeof(long long) + 16;
  static constexpr int XI = NY / XC::ROWS;         // x items per plane
  static constexpr int NBUN = (HX + B - 1) / B;    // kx bundles per plane
  static constexpr int YI = (NBUN + U - 1) / U;    // y items per plane
  static constexpr int PB = XI + YI;               // items per plane block
};

// sync[0] = next item to claim, sync[1 + plane] = finished producer items of the plane; zeroed by the host before every launch.
template <int HN, int NY, bool INV, bool RNG>
__global__ void __launch_bounds__(PLANE_THREADS, 4) p2_plane_kernel(const GSP_GRID_CONSTANT TensorMap tmHy, const double* __restrict__ in,
                                                                    double* __restrict__ out, cplx* __restrict__ H,
                                                                    const cplx* __restrict__ twxg, const cplx* __restrict__ stwxg,
                                                                    const cplx* __restrict__ stwyg, int nz, int lag, int* __restrict__ sync,
                                                                    double scale, double mu, XRng rng) {
  using C = PlaneCfg<HN, NY, INV>;
  using XC = typename C::XC;
  constexpr int NX = C::NX, HX = C::HX, B = C::B, U = C::U, XI = C::XI, YI = C::YI, PB = C::PB;
  constexpr int SLX = XC::SL, SLY = C::SLY;
  static_assert(!(RNG && INV), "the noise source belongs to the forward kernel");
  GSP_DYN_SMEM(smem);
  cplx* twx = reinterpret_cast<cplx*>(smem);   // exp(-2*pi*i*t/NX): untangling of the packed real transform
  cplx* stwx = twx + NX;                       // per-stage twiddles of the length-HN transform
  cplx* stwy = stwx + C::STWX;                 // per-stage twiddles of the length-NY transform
  unsigned char* stg = smem + C::TW_BYTES;     // the one stage: item input, then (overlaid) the exchange buffer
  mbar_t* full = reinterpret_cast<mbar_t*>(stg + C::STAGE);
  long long* slot = reinterpret_cast<long long*>(full + 1);  // claimed item numbers, two alternating slots
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(full, 1);
    fence_mbar_init();
  }
  for (int i = tid; i < NX; i += PLANE_THREADS) twx[i] = twxg[i];
  for (int i = tid; i < C::STWX; i += PLANE_THREADS) stwx[i] = stwxg[i];
  for (int i = tid; i < C::STWY; i += PLANE_THREADS) stwy[i] = stwyg[i];

  const long long nitems = (long long)(nz + lag) * PB;
  int* cnt = sync + 1;
  constexpr int need = INV ? YI : XI;  // producer items of one plane
  uint32_t phase = 0;

  for (int k = 0;; ++k) {
    // ---- claim the next item (the barrier also ends the previous item: its stage is free, its counter is bumped)
    if (tid == 0) slot[k & 1] = (long long)atomicAdd(sync, 1);
    __syncthreads();
    const long long item = slot[k & 1];
    if (item >= nitems) break;
    const int blk = (int)(item / PB), r = (int)(item - (long long)blk * PB);
    const bool first_kind = INV ? (r < YI) : (r < XI);   // the producing kind comes first in a block
    const bool is_x = INV ? !first_kind : first_kind;
    const int plane = first_kind ? blk : blk - lag;
    const int idx = first_kind ? r : r - (INV ? YI : XI);
    if (plane < 0 || plane >= nz) continue;              // padding items at both ends (uniform across the CTA)
    const bool producer = first_kind;
    const bool loads = !(RNG && is_x);

    // ---- fetch
    if (tid == 0 && loads) {
      if (!producer) {
        while (ld_acquire_gpu(cnt + plane) < need) spin_pause();
      }
      // generic-proxy accesses (this CTA's exchanges in the stage; other CTAs' stores to the plane) precede the async-proxy copy
      fence_proxy_async_all();
      if (is_x) {
        const long long row0 = (long long)plane * NY + (long long)idx * XC::ROWS;
        const uint32_t bytes = (uint32_t)XC::IN_BYTES;
        mbar_arrive_expect_tx(full, bytes);
        if (INV)
          bulk_g2s(stg, H + row0 * HX, bytes, full);
        else
          bulk_g2s(stg, in + row0 * NX, bytes, full);
      } else {
        mbar_arrive_expect_tx(full, (uint32_t)C::YIN);
#pragma unroll
        for (int u = 0; u < U; ++u) tma_load_3d(stg + (size_t)u * NY * B * sizeof(cplx), &tmHy, (idx * U + u) * 2 * B, 0, plane, full);
      }
    }
    if (loads) {
      mbar_wait(full, phase);
      phase ^= 1;
    }

    // ---- transform and store
    if (is_x) {
      const int rl = tid / XC::TPL, t = tid - rl * XC::TPL;
      const long long row = (long long)plane * NY + (long long)idx * XC::ROWS + rl;
      cplx* ex = reinterpret_cast<cplx*>(stg);
      const RowLay<XC::SH> lay{rl * XC::ROWLEN};
      cplx v[SLX];
      if constexpr (!INV) {
        constexpr int R0 = p2_radix(HN, false, 0);
        if constexpr (RNG) {
          const long long grow = rng.row_base + row;
          const long long rr = grow / rng.rows_per_real;
          const unsigned long long real = (unsigned long long)(rng.first_real + rr);
          const unsigned long long pair0 = (unsigned long long)(grow - rr * rng.rows_per_real) * HN;
#pragma unroll
          for (int q = 0; q < SLX / R0; ++q)
#pragma unroll
            for (int r2 = 0; r2 < R0; ++r2) {
              double u0, u1;
              philox_uniform2(rng.seed, 0u, real, pair0 + (unsigned long long)p2_in_pos<HN, false, 0>(t, q, r2), u0, u1);
              v[q * R0 + r2] = cplx{u0, u1};
            }
        } else {
          const cplx* src = ex + (size_t)rl * HN;
#pragma unroll
          for (int q = 0; q < SLX / R0; ++q)
#pragma unroll
            for (int r2 = 0; r2 < R0; ++r2) v[q * R0 + r2] = src[p2_in_pos<HN, false, 0>(t, q, r2)];
        }
        p2_fft<HN, false, 0>(v, t, ex, lay, stwx);
        // untangle: X[f] = E + w^f * O with E = (Z[f] + conj Z[h-f])/2, O = -i (Z[f] - conj Z[h-f])/2
        constexpr int RI = p2_radix(HN, true, 0);
        __syncthreads();
#pragma unroll
        for (int q = 0; q < SLX / RI; ++q)
#pragma unroll
          for (int r2 = 0; r2 < RI; ++r2) ex[lay(p2_in_pos<HN, true, 0>(t, q, r2))] = v[q * RI + r2];
        __syncthreads();
        double2* dst = reinterpret_cast<double2*>(H + row * HX);
#pragma unroll
        for (int q = 0; q < SLX / RI; ++q)
#pragma unroll
          for (int r2 = 0; r2 < RI; ++r2) {
            const int f = p2_in_pos<HN, true, 0>(t, q, r2);
            const cplx zk = v[q * RI + r2];
            const cplx zc = cconj(ex[lay((HN - f) & (HN - 1))]);
            const cplx e = cplx{0.5 * (zk.re + zc.re), 0.5 * (zk.im + zc.im)};
            const cplx d = csub(zk, zc);
            const cplx od = cplx{0.5 * d.im, -0.5 * d.re};
            const cplx o = cadd(e, cmul(twx[f], od));
            dst[f] = make_double2(o.re, o.im);  // plain store: the plane is re-read from L2 by the y items
            if (f == 0) dst[HN] = make_double2(zk.re - zk.im, 0.0);
          }
      } else {
        constexpr int R0 = p2_radix(HN, true, 0);
        const cplx* X = ex + (size_t)rl * HX;
#pragma unroll
        for (int q = 0; q < SLX / R0; ++q)
#pragma unroll
          for (int r2 = 0; r2 < R0; ++r2) {
            const int m = p2_in_pos<HN, true, 0>(t, q, r2);
            const cplx xk = X[m];
            const cplx xc = cconj(X[HN - m]);
            const cplx sm = cadd(xk, xc);
            const cplx d = csub(xk, xc);
            const cplx tt = cmul(cconj(twx[m]), d);
            v[q * R0 + r2] = cplx{sm.re - tt.im, sm.im + tt.re};
          }
        p2_fft<HN, true, 0>(v, t, ex, lay, stwx);
        constexpr int RO = p2_radix(HN, false, 0);
        double* dst = out + row * NX;
#pragma unroll
        for (int q = 0; q < SLX / RO; ++q)
#pragma unroll
          for (int r2 = 0; r2 < RO; ++r2) {
            const int j = p2_in_pos<HN, false, 0>(t, q, r2);
            st_stream2(dst + 2 * j, make_double2(v[q * RO + r2].re * scale + mu, v[q * RO + r2].im * scale + mu));
          }
      }
    } else {
      const int u = tid / C::TPU, lt = tid - u * C::TPU;
      const int b = lt % B, t = lt / B;
      const int bx = idx * U + u;
      const bool valid = bx < C::NBUN && bx * B + b < HX;
      cplx* buf = reinterpret_cast<cplx*>(stg) + (size_t)u * NY * B;
      const BundleLay lay{B, b};
      constexpr int R0 = p2_radix(NY, INV, 0);
      cplx v[SLY];
#pragma unroll
      for (int q = 0; q < SLY / R0; ++q)
#pragma unroll
        for (int r2 = 0; r2 < R0; ++r2) v[q * R0 + r2] = buf[lay(p2_in_pos<NY, INV, 0>(t, q, r2))];
      p2_fft<NY, INV, 0>(v, t, buf, lay, stwy);
      constexpr int RO = p2_radix(NY, !INV, 0);
      if (valid) {
        cplx* Hcol = H + (long long)plane * NY * HX + (long long)bx * B + b;
#pragma unroll
        for (int q = 0; q < SLY / RO; ++q)
#pragma unroll
          for (int r2 = 0; r2 < RO; ++r2) {
            const int m = p2_in_pos<NY, !INV, 0>(t, q, r2);
            double* p = reinterpret_cast<double*>(Hcol + (long long)m * HX);
            // forward: final result of this kernel (streamed out); inverse: re-read by the x items of this kernel (keep in L2)
            if (!INV)
              st_stream2(p, make_double2(v[q * RO + r2].re, v[q * RO + r2].im));
            else
              *reinterpret_cast<double2*>(p) = make_double2(v[q * RO + r2].re, v[q * RO + r2].im);
          }
      }
    }

    // ---- publish: every thread's plane stores become visible to the consumers' async-proxy reads, then one release per item
    if (producer) {
      fence_proxy_async_all();
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        red_release_gpu_add(cnt + plane, 1);
      }
    }
  }
}

}  // namespace gsp
