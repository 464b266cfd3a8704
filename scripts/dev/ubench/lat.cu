// dev microbenchmark: dependent-issue latencies on this GPU (cycles per op in a single warp)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, const double* in, int n) {
  extern __shared__ double sm[];
  const int lane = threadIdx.x;
  for (int i = lane; i < 1024; i += 32) sm[i] = in[i];
  __syncwarp();
  double a = in[lane], b = in[lane + 32], c = in[lane + 64];
  long long t0, t1;
  // dependent DFMA chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < 1024; i++) a = fma(a, b, c);
  t1 = clock64(); if (lane == 0) cyc[0] = t1 - t0;
  // dependent DADD
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < 1024; i++) a = a + b;
  t1 = clock64(); if (lane == 0) cyc[1] = t1 - t0;
  // dependent FFMA
  float fa = (float)a, fb = (float)b, fc = (float)c;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < 1024; i++) fa = fmaf(fa, fb, fc);
  t1 = clock64(); if (lane == 0) cyc[2] = t1 - t0;
  a += fa;
  // dependent shuffle of a double (2 SHFL)
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < 1024; i++) a = __shfl_sync(0xffffffffu, a, (lane + 1) & 31);
  t1 = clock64(); if (lane == 0) cyc[3] = t1 - t0;
  // dependent shared load (pointer chase through indices stored as doubles)
  int idx = lane;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < 1024; i++) idx = ((int)sm[idx]) & 1023;
  t1 = clock64(); if (lane == 0) cyc[4] = t1 - t0;
  // generic-pointer load from shared (compiler cannot prove the space)
  double* volatile gp = sm; double* g = gp;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < 1024; i++) idx = ((int)g[idx]) & 1023;
  t1 = clock64(); if (lane == 0) cyc[5] = t1 - t0;
  // shfl + dfma chain (trsv step)
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < 1024; i++) { double y = __shfl_sync(0xffffffffu, a, i & 31); a = fma(-b, y, a); }
  t1 = clock64(); if (lane == 0) cyc[6] = t1 - t0;
  // fp64 division chain
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < 256; i++) a = c / (a + 1.5);
  t1 = clock64(); if (lane == 0) cyc[7] = t1 - t0;
  // __syncwarp chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < 1024; i++) { __syncwarp(); }
  t1 = clock64(); if (lane == 0) cyc[8] = t1 - t0;
  // store + syncwarp + load (smem round trip between lanes)
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < 1024; i++) { sm[lane] = a; __syncwarp(); a = sm[(lane + 1) & 31] + 1.0; __syncwarp(); }
  t1 = clock64(); if (lane == 0) cyc[9] = t1 - t0;
  // global (L2-resident) dependent load
  const double* gi = in;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < 256; i++) idx = ((int)__ldcg(gi + idx)) & 1023;
  t1 = clock64(); if (lane == 0) cyc[10] = t1 - t0;
  out[lane] = a + idx;
}
int main() {
  double h[1024]; for (int i = 0; i < 1024; i++) h[i] = (double)((i * 7 + 3) & 1023);
  double *in, *out; long long* cyc;
  cudaMalloc(&in, sizeof(h)); cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 16 * 8);
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  k<<<1, 32, 8192>>>(out, cyc, in, 1024); k<<<1, 32, 8192>>>(out, cyc, in, 1024);
  long long c[16]; cudaMemcpy(c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  const char* nm[] = {"DFMA dep", "DADD dep", "FFMA dep", "SHFL(double) dep", "LDS dep (+cvt)", "LD generic->shared dep (+cvt)", "SHFL+DFMA (trsv step)", "DDIV dep (+dadd)", "syncwarp", "STS+sync+LDS+DADD+sync", "LDG.cg L2 dep"};
  int cnt[] = {1024,1024,1024,1024,1024,1024,1024,256,1024,1024,256};
  for (int i = 0; i < 11; i++) printf("%-34s %7.1f cycles/op\n", nm[i], (double)c[i] / cnt[i]);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
