// NearestInit on the device for CartesianGrid domains: replaces `initialize` + `NearestInit`
// (src/processes/field.jl:43-58, src/initialization/nearest.jl:12-34) for one variable.  The reference builds a KD-tree over all N
// element centroids (16.7 M Points at 256^3) and fills N-sized value / mask buffers to find where <= a few thousand data fall;
// on a regular grid the nearest centroid is index arithmetic, so the work is O(nd):
//   snap     datum i -> node floor((x - origin) / spacing), clamped to the grid (= the nearest centroid); NaN value = missing, skipped
//   winner   later data overwrite earlier ones (nearest.jl:27-31): atomicMax of the datum index per node
//   compact  winning (node, value) pairs sorted by node = findall(mask) (lusim.jl:71, fftsim.jl:104) and the values in that order
// Output crosses the boundary as (dinds 1-based ascending, z1): exactly what gsp_lu_plan_create / gsp_fft_plan_condition take.
#include <algorithm>
#include <climits>
#include <vector>

#include "common.h"
#include "cov.cuh"

namespace gsp {

constexpr long long KEY_NONE = LLONG_MAX;

// node[i] = 0-based linear index of the element nearest to datum i, or -1 when its value is missing (NaN)
__global__ void __launch_bounds__(256) snap_kernel(DomDev d, long long nd, const double* __restrict__ X, const double* __restrict__ v,
                                                   long long* __restrict__ node, int* __restrict__ winner) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nd) return;
  long long lin = -1;
  if (v[i] == v[i]) {
    lin = 0;
    long long stride = 1;
    for (int a = 0; a < d.dim; ++a) {
      long long c = (long long)floor((X[i * d.dim + a] - d.origin[a]) / d.spacing[a]);
      c = c < 0 ? 0 : (c >= d.dims[a] ? d.dims[a] - 1 : c);
      lin += c * stride;
      stride *= d.dims[a];
    }
    atomicMax(&winner[lin], (int)i);
  }
  node[i] = lin;
}

// key[i] = node of datum i if it is the last datum falling on that node, KEY_NONE otherwise (padding up to n2 too)
__global__ void __launch_bounds__(256) select_kernel(long long nd, long long n2, const long long* __restrict__ node, const int* __restrict__ winner,
                                                     const double* __restrict__ v, long long* __restrict__ key, double* __restrict__ val,
                                                     int* __restrict__ count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n2) return;
  long long k = KEY_NONE;
  double x = 0.0;
  if (i < nd && node[i] >= 0 && winner[node[i]] == (int)i) {
    k = node[i];
    x = v[i];
    atomicAdd(count, 1);
  }
  key[i] = k;
  val[i] = x;
}

// one compare-exchange stage of a bitonic sort of n2 = 2^m (key, value) pairs, ascending
__global__ void __launch_bounds__(256) bitonic_stage_kernel(long long n2, long long k, long long j, long long* __restrict__ key, double* __restrict__ val) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n2) return;
  const long long p = i ^ j;
  if (p <= i) return;
  const bool up = (i & k) == 0;
  const long long a = key[i], b = key[p];
  if ((a > b) == up) {
    key[i] = b;
    key[p] = a;
    const double t = val[i];
    val[i] = val[p];
    val[p] = t;
  }
}

}  // namespace gsp

using namespace gsp;

extern "C" int gsp_nearest_init(gsp_ctx* ctx, const gsp_domain* grid, int64_t nd, const double* dcoords, const double* dvals, int64_t* dinds,
                                double* z1, int64_t* count) {
  if (!ctx) return -1;
  std::lock_guard<std::mutex> lk(ctx->mu);
  DomDev dd;
  GSP_TRY(make_dom_dev(ctx, grid, 2, &dd));
  if (dd.kind != 1) return set_err(ctx, -2, "gsp_nearest_init needs a CartesianGrid (kind 1): other domains are searched by the host glue");
  if (nd < 0 || nd > INT_MAX) return set_err(ctx, -3, "nd out of range");
  if (!count) return set_err(ctx, -8, "count is NULL");
  *count = 0;
  if (nd == 0) return GSP_OK;
  if (!dcoords || !dvals || !dinds || !z1) return set_err(ctx, -4, "NULL array");
  DevCtx& dc = ctx->devs[0];
  cudaSetDevice(dc.dev);
  cudaStream_t st = dc.stream;
  long long n2 = 1;
  while (n2 < nd) n2 *= 2;
  DevBuf X, v, node, winner, key, val, cnt;
  GSP_CUDA_OK(ctx, X.alloc(dc.dev, (size_t)nd * dd.dim * sizeof(double), st));
  GSP_CUDA_OK(ctx, v.alloc(dc.dev, (size_t)nd * sizeof(double), st));
  GSP_CUDA_OK(ctx, node.alloc(dc.dev, (size_t)nd * sizeof(long long), st));
  GSP_CUDA_OK(ctx, winner.alloc(dc.dev, (size_t)dd.nelems * sizeof(int), st));
  GSP_CUDA_OK(ctx, key.alloc(dc.dev, (size_t)n2 * sizeof(long long), st));
  GSP_CUDA_OK(ctx, val.alloc(dc.dev, (size_t)n2 * sizeof(double), st));
  GSP_CUDA_OK(ctx, cnt.alloc(dc.dev, sizeof(int), st));
  GSP_CUDA_OK(ctx, cudaMemcpyAsync(X.p, dcoords, (size_t)nd * dd.dim * sizeof(double), cudaMemcpyHostToDevice, st));
  GSP_CUDA_OK(ctx, cudaMemcpyAsync(v.p, dvals, (size_t)nd * sizeof(double), cudaMemcpyHostToDevice, st));
  GSP_CUDA_OK(ctx, cudaMemsetAsync(winner.p, 0xFF, (size_t)dd.nelems * sizeof(int), st));  // -1
  GSP_CUDA_OK(ctx, cudaMemsetAsync(cnt.p, 0, sizeof(int), st));
  const unsigned gb = (unsigned)((nd + 255) / 256), gb2 = (unsigned)((n2 + 255) / 256);
  {
    ProfScope prof_("nearest_init", st);
    GSP_LAUNCH(snap_kernel, dim3(gb), dim3(256), 0, st, dd, (long long)nd, (const double*)X.as<double>(), (const double*)v.as<double>(),
               node.as<long long>(), winner.as<int>());
    GSP_LAUNCH(select_kernel, dim3(gb2), dim3(256), 0, st, (long long)nd, n2, (const long long*)node.as<long long>(), (const int*)winner.as<int>(),
               (const double*)v.as<double>(), key.as<long long>(), val.as<double>(), cnt.as<int>());
    g_launches += 2;
    for (long long k = 2; k <= n2; k *= 2)
      for (long long j = k / 2; j >= 1; j /= 2) {
        GSP_LAUNCH(bitonic_stage_kernel, dim3(gb2), dim3(256), 0, st, n2, k, j, key.as<long long>(), val.as<double>());
        g_launches++;
      }
  }
  GSP_CUDA_OK(ctx, cudaGetLastError());
  int c = 0;
  GSP_CUDA_OK(ctx, cudaMemcpyAsync(&c, cnt.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  GSP_CUDA_OK(ctx, cudaStreamSynchronize(st));
  if (c > 0) {
    std::vector<long long> hk((size_t)c);
    GSP_CUDA_OK(ctx, cudaMemcpyAsync(hk.data(), key.p, (size_t)c * sizeof(long long), cudaMemcpyDeviceToHost, st));
    GSP_CUDA_OK(ctx, cudaMemcpyAsync(z1, val.p, (size_t)c * sizeof(double), cudaMemcpyDeviceToHost, st));
    GSP_CUDA_OK(ctx, cudaStreamSynchronize(st));
    for (int i = 0; i < c; ++i) dinds[i] = (int64_t)hk[(size_t)i] + 1;  // 1-based like findall(mask)
  }
  *count = c;
  return GSP_OK;
}
