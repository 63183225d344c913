// Micro-benchmark (development tool): what does HBM deliver for the strided 16-byte-element access
// patterns of the y/z FFT passes, as a function of the contiguous run length (B * 16 bytes)?
#include <cstdio>
#include <cuda_runtime.h>
struct c16 { double re, im; };
// item = (bundle bx, other o): N positions x B contiguous elements; in-place style read+write of every element once
template <int N, int B>
__global__ void strided_rw(c16* __restrict__ H, long long es, int hx, int nbundles, long long nunits, long long other_stride) {
  const int TPU = (N / 16) * B;
  const int b = threadIdx.x % B, t = threadIdx.x / B;
  for (long long unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
    const long long o = unit / nbundles; const int bx = (int)(unit - o * nbundles);
    if (bx * B + b >= hx) continue;
    const long long base = o * other_stride + (long long)bx * B + b;
    double2 v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = *reinterpret_cast<const double2*>(H + base + (long long)(t + (N / 16) * r) * es);
#pragma unroll
    for (int r = 0; r < 16; ++r) { v[r].x += 1.0; *reinterpret_cast<double2*>(H + base + (long long)(t + (N / 16) * r) * es) = v[r]; }
  }
  (void)TPU;
}
__global__ void contiguous_rw(double2* __restrict__ H, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double2 v = H[i]; v.x += 1.0; H[i] = v;
  }
}
template <int N, int B>
float run(c16* H, long long es, int hx, long long nother, long long other_stride, int grid) {
  int nbundles = (hx + B - 1) / B; long long nunits = (long long)nbundles * nother;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 2; ++w) strided_rw<N, B><<<grid, (N / 16) * B>>>(H, es, hx, nbundles, nunits, other_stride);
  cudaEventRecord(e0);
  for (int w = 0; w < 5; ++w) strided_rw<N, B><<<grid, (N / 16) * B>>>(H, es, hx, nbundles, nunits, other_stride);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 5;
}
int main() {
  const int hx = 129, ny = 256, nz = 256; const long long nh = (long long)hx * ny * nz;
  c16* H; cudaMalloc(&H, nh * 16 * 2); cudaMemset(H, 0, nh * 16 * 2);
  const double gb = 2.0 * nh * 16 / 1e9;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  contiguous_rw<<<148 * 8, 256>>>((double2*)H, nh);
  cudaEventRecord(e0); for (int w = 0; w < 5; ++w) contiguous_rw<<<148 * 8, 256>>>((double2*)H, nh); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  printf("contiguous rw: %.1f us  %.0f GB/s\n", ms * 1e3, gb / ms * 1e3);
  for (int grid : {148 * 2, 148 * 4, 148 * 8, 148 * 16}) {
    float a = run<256, 4>(H, (long long)hx * ny, hx, ny, hx, grid);
    float b = run<256, 8>(H, (long long)hx * ny, hx, ny, hx, grid);
    float c = run<256, 16>(H, (long long)hx * ny, hx, ny, hx, grid);
    float d = run<256, 32>(H, (long long)hx * ny, hx, ny, hx, grid);
    printf("z-pattern grid %5d: B=4 %.1f us (%.0f GB/s)  B=8 %.1f us (%.0f)  B=16 %.1f us (%.0f)  B=32 %.1f us (%.0f)\n", grid, a * 1e3, gb / a * 1e3,
           b * 1e3, gb / b * 1e3, c * 1e3, gb / c * 1e3, d * 1e3, gb / d * 1e3);
    a = run<256, 4>(H, hx, hx, nz, (long long)hx * ny, grid);
    b = run<256, 8>(H, hx, hx, nz, (long long)hx * ny, grid);
    c = run<256, 16>(H, hx, hx, nz, (long long)hx * ny, grid);
    d = run<256, 32>(H, hx, hx, nz, (long long)hx * ny, grid);
    printf("y-pattern grid %5d: B=4 %.1f us (%.0f GB/s)  B=8 %.1f us (%.0f)  B=16 %.1f us (%.0f)  B=32 %.1f us (%.0f)\n", grid, a * 1e3, gb / a * 1e3,
           b * 1e3, gb / b * 1e3, c * 1e3, gb / c * 1e3, d * 1e3, gb / d * 1e3);
  }
  // padded leading dimension hx = 136 (128-byte aligned rows) for comparison
  {
    const int hp = 136; 
    float b = run<256, 8>(H, (long long)hp * ny, 129, ny, hp, 148 * 8);
    float c = run<256, 16>(H, (long long)hp * ny, 129, ny, hp, 148 * 8);
    float b2 = run<256, 8>(H, hp, 129, nz, (long long)hp * ny, 148 * 8);
    float c2 = run<256, 16>(H, hp, 129, nz, (long long)hp * ny, 148 * 8);
    printf("ld=136 aligned rows: z B=8 %.1f us B=16 %.1f us | y B=8 %.1f us B=16 %.1f us\n", b * 1e3, c * 1e3, b2 * 1e3, c2 * 1e3);
  }
  return 0;
}
