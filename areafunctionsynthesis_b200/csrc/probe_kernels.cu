// probe_kernels.cu -- FP64 pipe peak probe.  MEASURED_PEAKS.json carries HBM and bf16 tensor peaks
// only; both hot kernels of this library are FP64-pipe bound, so the roofline denominator is
// measured with independent DFMA chains (8 per thread, 1024 resident threads per SM).
#include "afs_internal.h"

namespace afs {

__global__ void __launch_bounds__(256) fp64_probe_kernel(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456) out[0] = s;  // never true: keeps the chains alive
}

cudaError_t launch_fp64_probe(double *d_out, int ctas, int iters, cudaStream_t stream) {
  fp64_probe_kernel<<<ctas, 256, 0, stream>>>(d_out, iters, 0.9999999, 1e-9);
  return cudaGetLastError();
}

}  // namespace afs
