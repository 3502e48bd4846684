// FP64 pipe / issue-port microbenchmark for sm_100a: does an integer (or FP32, or LDS) instruction issue in the
// shadow of a DFMA?  Per loop iteration: 8 independent DFMA + NI independent IMAD (or FFMA / LDS).
//   nvcc -arch=sm_100a -O3 -o fp64_mix fp64_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int NI, int KIND>
__global__ void k(double *out, long long *cyc, int n, double a, double b, int ia, float fa) {
  __shared__ float sh[1024];
  double v[8];
  int iv[16];
  float fv[16];
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
  for (int i = 0; i < 16; i++) { iv[i] = threadIdx.x + i; fv[i] = threadIdx.x * 0.5f + i; }
  sh[threadIdx.x % 1024] = 1.0f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < n; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      v[i] = fma(v[i], a, b);
#pragma unroll
      for (int j = 0; j < NI; j++) {
        const int s = (i * NI + j) % 16;
        if (KIND == 0) iv[s] = iv[s] * ia + it;                         // IMAD
        if (KIND == 1) fv[s] = fmaf(fv[s], fa, 1.0f);                   // FFMA
        if (KIND == 2) fv[s] += sh[(threadIdx.x + s * 32 + it) & 1023];  // LDS + FADD
        if (KIND == 3) iv[s] = (iv[s] > it) ? iv[s] - ia : iv[s] + 3;   // ISETP + SEL-ish
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += v[i];
#pragma unroll
  for (int i = 0; i < 16; i++) s += iv[i] + fv[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int NI, int KIND>
void run(int warps) {
  double *out; long long *cyc;
  int nb = 148, n = 4096;
  cudaMalloc(&out, sizeof(double) * nb * warps * 32); cudaMalloc(&cyc, sizeof(long long) * nb);
  k<NI, KIND><<<nb, warps * 32>>>(out, cyc, n, 0.999, 1e-3, 3, 0.999f);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  const char *names[] = {"IMAD", "FFMA", "LDS+FADD", "ISETP+SEL"};
  double per = (double)h / ((double)n * 8);
  printf("%-10s x%d per DFMA, %2d warps/SM (%d per SMSP): %.2f cycles per (DFMA + %d other) per warp; SMSP cycles per group %.2f\n",
         names[KIND], NI, warps, warps / 4, per, NI, per / (warps / 4));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0, 0>(8); run<1, 0>(8); run<2, 0>(8); run<3, 0>(8);
  run<0, 0>(12); run<1, 0>(12); run<2, 0>(12); run<3, 0>(12);
  run<1, 1>(8); run<2, 1>(8); run<1, 1>(12); run<2, 1>(12);
  run<1, 2>(8); run<2, 2>(8); run<1, 2>(12);
  run<1, 3>(8); run<2, 3>(8); run<1, 3>(12);
  run<1, 0>(4); run<2, 0>(4); run<1, 0>(16); run<2, 0>(16);
  return 0;
}
