// dev microbenchmark: cost of taken branches / divergent regions / calls for a single warp
#include <cstdio>
#include <cuda_runtime.h>
__device__ __noinline__ double callee(double a, double b) { return fma(a, b, 1.0); }
__global__ void k(double* out, long long* cyc, const int* flags, int n) {
  const int lane = threadIdx.x;
  double a = 1.0 + lane, b = 0.5;
  long long t0, t1;
  // 0: straight-line baseline: 8 dependent DFMA per iteration, loop unrolled x1
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 1024; i++) { a = fma(a, b, 1.0); a = fma(a, b, 1.0); a = fma(a, b, 1.0); a = fma(a, b, 1.0); }
  t1 = clock64(); if (lane == 0) cyc[0] = t1 - t0;
  // 1: same work, plus a UNIFORM taken forward branch every iteration (flags[i] is 1)
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 1024; i++) {
    a = fma(a, b, 1.0); a = fma(a, b, 1.0);
    if (flags[i & 63]) { a = fma(a, b, 1.0); a = fma(a, b, 1.0); } else { a = a * 3.0 + b; b = b * 0.5; a += b; a = a * a; }
  }
  t1 = clock64(); if (lane == 0) cyc[1] = t1 - t0;
  // 2: divergent region: only one lane does the work (if lane == i % 32)
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 1024; i++) {
    a = fma(a, b, 1.0); a = fma(a, b, 1.0);
    if (lane == (i & 31)) { a = fma(a, b, 1.0); a = fma(a, b, 1.0); }
  }
  t1 = clock64(); if (lane == 0) cyc[2] = t1 - t0;
  // 3: same with select instead of branch
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 1024; i++) {
    a = fma(a, b, 1.0); a = fma(a, b, 1.0);
    double c = fma(a, b, 1.0); c = fma(c, b, 1.0);
    a = (lane == (i & 31)) ? c : a;
  }
  t1 = clock64(); if (lane == 0) cyc[3] = t1 - t0;
  // 4: function call per iteration
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 1024; i++) { a = callee(a, b); a = fma(a, b, 1.0); a = fma(a, b, 1.0); a = fma(a, b, 1.0); }
  t1 = clock64(); if (lane == 0) cyc[4] = t1 - t0;
  // 5: inner CW_FOR-like loop (one iteration for lanes < n) inside
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 1024; i++) {
    a = fma(a, b, 1.0); a = fma(a, b, 1.0);
    for (int j = lane; j < n; j += 32) { a = fma(a, b, 1.0); a = fma(a, b, 1.0); }
  }
  t1 = clock64(); if (lane == 0) cyc[5] = t1 - t0;
  // 6: uniform outer loop + predicated body (new CW_FOR)
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 1024; i++) {
    a = fma(a, b, 1.0); a = fma(a, b, 1.0);
    for (int jb = 0; jb < n; jb += 32) for (int j = jb + lane, o = 1; o; o = 0) if (j < n) { a = fma(a, b, 1.0); a = fma(a, b, 1.0); }
  }
  t1 = clock64(); if (lane == 0) cyc[6] = t1 - t0;
  // 7: plain `if (lane < n)` region
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 1024; i++) {
    a = fma(a, b, 1.0); a = fma(a, b, 1.0);
    if (lane < n) { a = fma(a, b, 1.0); a = fma(a, b, 1.0); }
  }
  t1 = clock64(); if (lane == 0) cyc[7] = t1 - t0;
  // 8: divergent CW_FOR with a bigger body (shared memory traffic: cannot be predicated away cheaply)
  extern __shared__ double sm[];
  sm[lane] = a; sm[lane + 32] = b; __syncwarp();
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 1024; i++) {
    a = fma(a, b, 1.0); a = fma(a, b, 1.0);
    for (int j = lane; j < n; j += 32) { double s = sm[j]; for (int k = 0; k < 4; k++) s = fma(s, sm[32 + ((j + k) & 31)], 1.0); sm[j] = s; }
    __syncwarp();
  }
  t1 = clock64(); if (lane == 0) cyc[8] = t1 - t0;
  // 9: same body, uniform loop + predicate
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 1024; i++) {
    a = fma(a, b, 1.0); a = fma(a, b, 1.0);
    for (int jb = 0; jb < n; jb += 32) for (int j = jb + lane, o = 1; o; o = 0) if (j < n) { double s = sm[j]; for (int k = 0; k < 4; k++) s = fma(s, sm[32 + ((j + k) & 31)], 1.0); sm[j] = s; }
    __syncwarp();
  }
  t1 = clock64(); if (lane == 0) cyc[9] = t1 - t0;
  // 10: same body, no guard at all (all 32 lanes)
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 1024; i++) {
    a = fma(a, b, 1.0); a = fma(a, b, 1.0);
    { const int j = lane; double s = sm[j]; for (int k = 0; k < 4; k++) s = fma(s, sm[32 + ((j + k) & 31)], 1.0); sm[j] = s; }
    __syncwarp();
  }
  t1 = clock64(); if (lane == 0) cyc[10] = t1 - t0;
  out[lane] = a + b + sm[lane];
}
int main() {
  int h[64]; for (int i = 0; i < 64; i++) h[i] = 1;
  int* f; double* out; long long* cyc;
  cudaMalloc(&f, sizeof(h)); cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 16 * 8);
  cudaMemcpy(f, h, sizeof(h), cudaMemcpyHostToDevice);
  k<<<1, 32, 1024>>>(out, cyc, f, 24); k<<<1, 32, 1024>>>(out, cyc, f, 24);
  long long c[16]; cudaMemcpy(c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  const char* nm[] = {"4 dep DFMA (baseline)", "4 DFMA + uniform taken branch (+LDG flag)", "4 DFMA, 2 of them in a 1-lane divergent region", "same with select", "3 DFMA + call(DFMA)", "4 DFMA, 2 in a CW_FOR (n=24)", "4 DFMA, 2 in uniform-loop+predicate", "4 DFMA, 2 in if (lane < n)", "2 DFMA + divergent CW_FOR smem body", "2 DFMA + uniform-loop predicated smem body", "2 DFMA + unguarded smem body"};
  for (int i = 0; i < 11; i++) printf("%-52s %7.1f cycles/iter\n", nm[i], (double)c[i] / 1024);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
