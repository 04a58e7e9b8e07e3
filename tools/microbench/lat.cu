// Latency microbenchmarks (one warp): dependent DFMA chain, rcp sequence, LDS round trip, FP64 issue rate.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double fast_rcp(double x) {
  double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0); y = fma(y, e, y); e = fma(-x, y, 1.0); y = fma(y, e, y); return y;
}
__global__ void k(double *out, long long *t, double a, double b, int iters, int nwarps_active) {
  __shared__ double sm[64][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = 0; i < 64; i++) sm[i][lane] = 1.0 + i;
  __syncthreads();
  if (warp >= nwarps_active) return;
  double x = a + lane;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) { x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); }
  long long t1 = clock64();
  double y = 1.5 + lane;
  for (int i = 0; i < iters; i++) { y = fast_rcp(y) + 1.25; }
  long long t2 = clock64();
  double z = 0; int idx = lane & 1;
  for (int i = 0; i < iters; i++) { z += sm[idx][lane]; idx = ((int)z) & 63; }
  long long t3 = clock64();
  double w0 = x, w1 = y, w2 = z, w3 = 1, w4 = 2, w5 = 3, w6 = 4, w7 = 5;
  for (int i = 0; i < iters; i++) {
    w0 = fma(w0, a, b); w1 = fma(w1, a, b); w2 = fma(w2, a, b); w3 = fma(w3, a, b);
    w4 = fma(w4, a, b); w5 = fma(w5, a, b); w6 = fma(w6, a, b); w7 = fma(w7, a, b);
  }
  long long t4 = clock64();
  double d = 3.0 + lane;
  for (int i = 0; i < iters; i++) { d = 1.0 / d + 1.25; }
  long long t5 = clock64();
  double s = 3.0 + lane;
  for (int i = 0; i < iters; i++) { s = sqrt(s) + 1.25; }
  long long t6 = clock64();
  if (lane == 0) { t[warp * 8 + 0] = t1 - t0; t[warp * 8 + 1] = t2 - t1; t[warp * 8 + 2] = t3 - t2; t[warp * 8 + 3] = t4 - t3; t[warp*8+4] = t5 - t4; t[warp*8+5] = t6 - t5; }
  out[threadIdx.x] = x + y + z + w0 + w1 + w2 + w3 + w4 + w5 + w6 + w7 + d + s;
}
int main() {
  double *out; long long *t; cudaMalloc(&out, 8 * 1024); cudaMalloc(&t, 8 * 8 * 32);
  const int iters = 4096;
  for (int nw : {1, 4, 8, 16}) {
    k<<<1, 512>>>(out, t, 0.999999, 1e-7, iters, nw); cudaDeviceSynchronize();
    long long h[256]; cudaMemcpy(h, t, sizeof h, cudaMemcpyDeviceToHost);
    printf("active warps %2d: dep DFMA %.1f cyc | rcp(NR2)+add %.1f cyc | LDS dep %.1f cyc | 8 indep DFMA %.2f cyc/instr | IEEE div+add %.1f | IEEE sqrt+add %.1f\n", nw,
           h[0] / (4.0 * iters), h[1] / (double)iters, h[2] / (double)iters, h[3] / (8.0 * iters), h[4] / (double)iters, h[5] / (double)iters);
  }
  return 0;
}
