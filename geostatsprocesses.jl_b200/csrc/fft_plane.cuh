// Fused x+y passes of the 3-D FFTSIM pipeline: one persistent kernel transforms every z-plane along
// x AND y, exchanging the intermediate half-spectrum plane through L2 instead of HBM.
//
//   forward  (INV = false): x items (ROWS real rows -> half-spectrum rows) then y items (B kx columns, in place)
//   inverse  (INV = true ): y items (in place)                              then x items (rows -> real field, +mu)
//
// Work items are numbered plane by plane, the dependent kind lagging `lag` planes behind the producing kind
// (lag is chosen by the host to exceed the window of items that are in flight or being prefetched at any time,
// so a dependency is almost always already satisfied when its item is fetched).
// CTA c owns items c, c+G, c+2G, ... and walks them in order; a producing item bumps its plane's counter
// (release) when its stores are done, a dependent item is only fetched (TMA / bulk copy into the
// other shared-memory stage) once its plane counter has reached the expected value (acquire).  Every
// dependency points to a lower item number, so with all CTAs co-resident the walk cannot deadlock.
// The plane written by the producers (<= 0.6 MB) is still in the 126 MB L2 when the consumers read it:
// HBM sees one read of the input and one write of the result per plane - 2 passes per transform, not 3.
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
                             (NY % XC::ROWS) == 0;
  static constexpr size_t XIN = XC::IN_BYTES;
  static constexpr size_t YIN = (size_t)U * NY * B * sizeof(cplx);
  static constexpr size_t STAGE = ((XIN > YIN ? XIN : YIN) + 127) / 128 * 128;
  static constexpr size_t TW_BYTES = (size_t)(NX + NY) * sizeof(cplx);
  static constexpr size_t SMEM = TW_BYTES + 2 * STAGE + XC::EX_BYTES + 2 * sizeof(mbar_t) + 16;
  static constexpr int XI = NY / XC::ROWS;  // x items per plane
};

// ---- item bodies (same arithmetic as the stand-alone pass kernels of fft_pow2.cuh)
template <int HN>
GSP_DEV void plane_xfwd_item(const cplx* srcrow, cplx* ex, const cplx* tw, cplx* dst, bool valid, int t, RowLay<XCfg<HN, false>::SH> lay) {
  constexpr int SL = p2_slots(HN);
  constexpr int R0 = p2_radix(HN, false, 0);
  cplx v[SL];
#pragma unroll
  for (int q = 0; q < SL / R0; ++q)
#pragma unroll
    for (int r = 0; r < R0; ++r) v[q * R0 + r] = srcrow[p2_in_pos<HN, false, 0>(t, q, r)];
  p2_fft<HN, false, 2>(v, t, ex, lay, tw);
  constexpr int RI = p2_radix(HN, true, 0);
  __syncthreads();
#pragma unroll
  for (int q = 0; q < SL / RI; ++q)
#pragma unroll
    for (int r = 0; r < RI; ++r) ex[lay(p2_in_pos<HN, true, 0>(t, q, r))] = v[q * RI + r];
  __syncthreads();
  if (valid) {
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
        const cplx o = cadd(e, cmul(tw[f], od));
        reinterpret_cast<double2*>(dst)[f] = make_double2(o.re, o.im);  // plain store: the plane is re-read from L2
        if (f == 0) reinterpret_cast<double2*>(dst)[HN] = make_double2(zk.re - zk.im, 0.0);
      }
  }
}

template <int HN>
GSP_DEV void plane_xinv_item(const cplx* X, cplx* ex, const cplx* tw, double* dst, bool valid, int t, RowLay<XCfg<HN, true>::SH> lay,
                             double scale, double mu) {
  constexpr int SL = p2_slots(HN);
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
      const cplx tt = cmul(cconj(tw[m]), d);
      v[q * R0 + r] = cplx{sm.re - tt.im, sm.im + tt.re};
    }
  p2_fft<HN, true, 2>(v, t, ex, lay, tw);
  constexpr int RO = p2_radix(HN, false, 0);
  if (valid) {
#pragma unroll
    for (int q = 0; q < SL / RO; ++q)
#pragma unroll
      for (int r = 0; r < RO; ++r) {
        const int j = p2_in_pos<HN, false, 0>(t, q, r);
        st_stream2(dst + 2 * j, make_double2(v[q * RO + r].re * scale + mu, v[q * RO + r].im * scale + mu));
      }
  }
}

template <int NY, bool INV>
GSP_DEV void plane_y_item(cplx* buf, const cplx* tw, cplx* Hcol, long long es, bool valid, int t, int b, bool streaming) {
  constexpr int SL = p2_slots(NY), B = p2_bundle(NY);
  const BundleLay lay{B, b};
  constexpr int R0 = p2_radix(NY, INV, 0);
  cplx v[SL];
#pragma unroll
  for (int q = 0; q < SL / R0; ++q)
#pragma unroll
    for (int r = 0; r < R0; ++r) v[q * R0 + r] = buf[lay(p2_in_pos<NY, INV, 0>(t, q, r))];
  p2_fft<NY, INV, 1>(v, t, buf, lay, tw);
  constexpr int RO = p2_radix(NY, !INV, 0);
  if (valid) {
#pragma unroll
    for (int q = 0; q < SL / RO; ++q)
#pragma unroll
      for (int r = 0; r < RO; ++r) {
        const int m = p2_in_pos<NY, !INV, 0>(t, q, r);
        double* p = reinterpret_cast<double*>(Hcol + (long long)m * es);
        if (streaming)
          st_stream2(p, make_double2(v[q * RO + r].re, v[q * RO + r].im));
        else
          *reinterpret_cast<double2*>(p) = make_double2(v[q * RO + r].re, v[q * RO + r].im);
      }
  }
}

// cnt[plane] counts finished producer items of the plane; `epoch` (1, 2, ...) makes the counters reusable across
// launches without resetting them: the dependent kind waits for cnt[plane] >= epoch * (producer items per plane).
template <int HN, int NY, bool INV>
__global__ void __launch_bounds__(PLANE_THREADS) p2_plane_kernel(const GSP_GRID_CONSTANT TensorMap tmHy, const double* __restrict__ in,
                                                                 double* __restrict__ out, cplx* __restrict__ H, const cplx* __restrict__ twxg,
                                                                 const cplx* __restrict__ twyg, int nz, int lag, int* __restrict__ cnt,
                                                                 int epoch, double scale, double mu) {
  using C = PlaneCfg<HN, NY, INV>;
  using XC = typename C::XC;
  constexpr int NX = C::NX, HX = C::HX, B = C::B, U = C::U, XI = C::XI;
  constexpr int NBUN = (HX + B - 1) / B;          // kx bundles per plane
  constexpr int YI = (NBUN + U - 1) / U;          // y items per plane
  constexpr int PB = XI + YI;                     // items per plane block
  GSP_DYN_SMEM(smem);
  cplx* twx = reinterpret_cast<cplx*>(smem);
  cplx* twy = twx + NX;
  unsigned char* stage0 = smem + C::TW_BYTES;
  cplx* ex = reinterpret_cast<cplx*>(stage0 + 2 * C::STAGE);
  mbar_t* full = reinterpret_cast<mbar_t*>(stage0 + 2 * C::STAGE + XC::EX_BYTES);
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_mbar_init();
  }
  for (int i = tid; i < NX; i += PLANE_THREADS) twx[i] = twxg[i];
  for (int i = tid; i < NY; i += PLANE_THREADS) twy[i] = twyg[i];
  __syncthreads();

  const long long nitems = (long long)(nz + lag) * PB;
  // item -> (is_x, plane, index within the kind); returns false for the padding items at both ends
  auto decode = [&](long long item, bool& is_x, int& plane, int& idx) -> bool {
    const int blk = (int)(item / PB), r = (int)(item - (long long)blk * PB);
    const bool first_kind = INV ? (r < YI) : (r < XI);   // producing kind comes first in a block
    is_x = INV ? !first_kind : first_kind;
    plane = first_kind ? blk : blk - lag;
    idx = first_kind ? r : r - (INV ? YI : XI);
    return plane >= 0 && plane < nz;
  };
  auto next_valid = [&](long long item) -> long long {
    bool ix; int pl, id;
    while (item < nitems && !decode(item, ix, pl, id)) item += gridDim.x;
    return item;
  };
  const int need = epoch * (INV ? YI : XI);  // producer items of one plane, accumulated over launches
  // thread 0 only: fetch `item` into stage sg (dependent kinds first wait for their plane)
  auto issue = [&](long long item, int sg, bool blocking) -> bool {
    bool is_x; int plane, idx;
    decode(item, is_x, plane, idx);
    const bool dependent = INV ? is_x : !is_x;
    if (dependent) {
      if (blocking) {
        while (ld_acquire_gpu(cnt + plane) < need) spin_pause();
      } else if (ld_acquire_gpu(cnt + plane) < need) {
        return false;
      }
      fence_proxy_async();  // the plane was written through the generic proxy, the copy below reads it through the async proxy
    }
    unsigned char* dst = stage0 + (size_t)sg * C::STAGE;
    if (is_x) {
      const long long row0 = (long long)plane * NY + (long long)idx * XC::ROWS;
      const uint32_t bytes = (uint32_t)XC::IN_BYTES;
      mbar_arrive_expect_tx(&full[sg], bytes);
      if (INV)
        bulk_g2s(dst, H + row0 * HX, bytes, &full[sg]);
      else
        bulk_g2s(dst, in + row0 * NX, bytes, &full[sg]);
    } else {
      mbar_arrive_expect_tx(&full[sg], (uint32_t)C::YIN);
#pragma unroll
      for (int u = 0; u < U; ++u)
        tma_load_3d(dst + (size_t)u * NY * B * sizeof(cplx), &tmHy, (idx * U + u) * 2 * B, 0, plane, &full[sg]);
    }
    return true;
  };

  long long cur = next_valid(blockIdx.x);
  if (cur < nitems && tid == 0) issue(cur, 0, true);
  for (int k = 0; cur < nitems; ++k) {
    const int sg = k & 1;
    const long long nxt = next_valid(cur + gridDim.x);
    bool issued_next = true;
    if (nxt < nitems && tid == 0) issued_next = issue(nxt, sg ^ 1, false);
    bool is_x; int plane, idx;
    decode(cur, is_x, plane, idx);
    unsigned char* stg = stage0 + (size_t)sg * C::STAGE;
    mbar_wait(&full[sg], (uint32_t)((k >> 1) & 1));
    if (is_x) {
      const int rl = tid / XC::TPL, t = tid - rl * XC::TPL;
      const long long row = (long long)plane * NY + (long long)idx * XC::ROWS + rl;
      const RowLay<XC::SH> lay{rl * XC::ROWLEN};
      if constexpr (INV)
        plane_xinv_item<HN>(reinterpret_cast<const cplx*>(stg) + (size_t)rl * HX, ex, twx, out + row * NX, true, t, lay, scale, mu);
      else
        plane_xfwd_item<HN>(reinterpret_cast<const cplx*>(stg) + (size_t)rl * HN, ex, twx, H + row * HX, true, t, lay);
    } else {
      const int u = tid / C::TPU, lt = tid - u * C::TPU;
      const int b = lt % B, t = lt / B;
      const int bx = idx * U + u;
      const bool valid = bx < NBUN && bx * B + b < HX;
      cplx* buf = reinterpret_cast<cplx*>(stg) + (size_t)u * NY * B;
      // forward: final result of this kernel (streamed out); inverse: re-read by the x items of this kernel (keep in L2)
      plane_y_item<NY, INV>(buf, twy, H + (long long)plane * NY * HX + (long long)bx * B + b, HX, valid, t, b, !INV);
    }
    const bool producer = INV ? !is_x : is_x;
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      // the barrier orders every thread's stores before this release (cumulativity): one GPU-scope release per item
      if (producer) {
        __threadfence();
        red_release_gpu_add(cnt + plane, 1);
      }
      if (nxt < nitems && !issued_next) issue(nxt, sg ^ 1, true);
    }
    cur = nxt;
  }
}

}  // namespace gsp
