// FP64 pipe microbenchmark for sm_100a: dependent-issue latency and per-warp / per-SM throughput of DFMA,
// as a function of independent chains per warp (ILP) and warps per SM (TLP).   nvcc -arch=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, long long *cyc, int n, double a, double b) {
  double v[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) v[i] = threadIdx.x * 1e-3 + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < n; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) v[i] = fma(v[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ILP>
void run(int warps, int blocks_per_sm) {
  double *out; long long *cyc;
  int nb = 148 * blocks_per_sm, n = 4096;
  cudaMalloc(&out, sizeof(double) * nb * warps * 32); cudaMalloc(&cyc, sizeof(long long) * nb);
  k<ILP><<<nb, warps * 32>>>(out, cyc, n, 0.999, 1e-3);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  double per = (double)h / ((double)n * ILP);
  printf("ILP %2d warps/CTA %2d CTAs/SM %d : %.2f cycles per DFMA per warp -> %.3f warp-DFMA/cycle/SM\n", ILP, warps, blocks_per_sm, per,
         warps * blocks_per_sm / per);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<1>(1, 1); run<2>(1, 1); run<4>(1, 1); run<8>(1, 1); run<16>(1, 1);
  run<1>(4, 1); run<2>(4, 1); run<4>(4, 1); run<8>(4, 1);
  run<1>(8, 1); run<2>(8, 1); run<4>(8, 1); run<8>(8, 1);
  run<1>(16, 1); run<2>(16, 1); run<4>(16, 1);
  run<1>(32, 1); run<2>(32, 1);
  return 0;
}
