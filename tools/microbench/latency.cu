// Dependent-issue latencies that bound the single-CTA q x q solve of the polar tail (tail.cu): FP64 FMA chain, rsqrt
// variants, double shuffles, CTA barrier at 512 threads, shared-memory round trip.   nvcc -arch=sm_100a -O3 latency.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double rsqrt_fast(double x) {
  double y = static_cast<double>(rsqrtf(static_cast<float>(x)));
  const double hx = 0.5 * x;
  y = y * fma(-hx * y, y, 1.5);
  y = y * fma(-hx * y, y, 1.5);
  return y;
}

template <int WHICH>
__global__ void k_lat(double* out, long long* cyc, int iters, double seed) {
  __shared__ double sh[64];
  double x = seed + threadIdx.x * 1e-9, y = 1.0000001;
  if (threadIdx.x < 64) sh[threadIdx.x] = x;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (WHICH == 0) { x = fma(x, y, 1e-9); }                                   // DFMA chain
    if (WHICH == 1) { x = rsqrt(x + 1.5); }                                     // library rsqrt (+ DADD)
    if (WHICH == 2) { x = rsqrt_fast(x + 1.5); }                                // FP32 seed + 2 Newton (+ DADD)
    if (WHICH == 3) { x += __shfl_xor_sync(0xffffffffu, x, 1); }                // double shuffle + DADD
    if (WHICH == 4) { __syncthreads(); }                                        // CTA barrier
    if (WHICH == 5) { sh[threadIdx.x & 63] = x; __syncthreads(); x = sh[(threadIdx.x + 1) & 63] + 1e-9; __syncthreads(); }
    if (WHICH == 6) { x = sqrt(x + 1.5); }
    if (WHICH == 7) { x = 1.0 / (x + 1.5); }
    if (WHICH == 8) { x = x * y; }                                              // DMUL chain
    if (WHICH == 9) { x = x + y; }                                              // DADD chain
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = x;
}

template <int WHICH>
void run(const char* name, int threads) {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * sizeof(double));
  cudaMalloc(&cyc, sizeof(long long));
  const int iters = 2000;
  k_lat<WHICH><<<1, threads>>>(out, cyc, iters, 1.25);
  k_lat<WHICH><<<1, threads>>>(out, cyc, iters, 1.25);
  long long c = 0;
  cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  printf("%-44s threads %4d: %7.1f cycles per iteration\n", name, threads, static_cast<double>(c) / iters);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("DFMA dependent chain", 32);
  run<8>("DMUL dependent chain", 32);
  run<9>("DADD dependent chain", 32);
  run<1>("rsqrt(double) + DADD", 32);
  run<2>("rsqrtf seed + 2 Newton + DADD", 32);
  run<6>("sqrt(double) + DADD", 32);
  run<7>("1/x (double) + DADD", 32);
  run<3>("double __shfl_xor + DADD", 32);
  run<4>("__syncthreads", 512);
  run<4>("__syncthreads", 128);
  run<5>("STS + bar + LDS + DADD + bar", 512);
  run<0>("DFMA dependent chain, 16 warps resident", 512);
  return 0;
}
