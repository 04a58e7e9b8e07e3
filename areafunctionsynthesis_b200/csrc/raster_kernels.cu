// raster_kernels.cu -- K4 `area_function_raster_kernel`: OneDimAreaFunction parameter sets -> tube frames on the
// device.
//
// Replaces OneDimAreaFunction::calculateArea / calculateOneDimTubeFunction (reference OneDimAreaFunction.cpp:23-138)
// for batches: 40 equal-length pharynx/mouth sections, each taking the MINIMUM of the raised-cosine area function
// A(x) sampled every 1/100 section width.  The sample position x ACCUMULATES over the whole tract in the reference
// (`x_cm += stepSize` inside `while (x_cm < (i+1)*secWidth)`), which makes the sampling discontinuous in rounding;
// the kernel repeats exactly that sequence of additions and comparisons (they are exact IEEE operations), so the
// sample positions, the number of samples per section and the articulator labels are the reference's.  Only
// pow()/cos() differ from glibc by an ulp or two, hence the areas agree to ~1e-15 relative, not bitwise.
//
// Mapping: the work per parameter set is 4000 (pow, cos) evaluations, embarrassingly parallel over (set, section).
// One thread per (set, section): it first replays the cheap running-x additions of the sections before its own (no
// area evaluation), then samples its own section.  A CTA = 40 threads x 4 sets; parameters are read through the
// read-only path (16 doubles per set), the frame fields are written once.
#include <cfloat>

#include "afs_internal.h"

namespace afs {

namespace {
enum { LLAR, ALAR, POWLAR, XP, AP, POWP, XC, AC, POWC, XA, AA, POWA, XIN, AIN, LVT, ALIP };

// OneDimAreaFunction::calculateArea (OneDimAreaFunction.cpp:23-61)
__device__ __forceinline__ double area_at(const double *q, double x) {
  double a;
  if (x <= q[LLAR]) a = q[ALAR];
  else if (x <= q[XP]) a = (q[AP] + q[ALAR]) / 2 + (q[AP] - q[ALAR]) / 2 * cos(kPi * pow((q[XP] - x) / (q[XP] - q[LLAR]), q[POWLAR]));
  else if (x <= q[XC]) a = (q[AC] + q[AP]) / 2 + (q[AC] - q[AP]) / 2 * cos(kPi * pow((q[XC] - x) / (q[XC] - q[XP]), q[POWP]));
  else if (x <= q[XA]) a = (q[AA] + q[AC]) / 2 + (q[AA] - q[AC]) / 2 * cos(kPi * pow((q[XA] - x) / (q[XA] - q[XC]), q[POWC]));
  else if (x <= q[XIN]) a = (q[AIN] + q[AA]) / 2 + (q[AIN] - q[AA]) / 2 * cos(kPi * pow((q[XIN] - x) / (q[XIN] - q[XA]), q[POWA]));
  else a = q[ALIP];
  return a < 0.0 ? 0.0 : a;
}
}  // namespace

constexpr int RASTER_SETS_PER_CTA = 4;

__global__ void __launch_bounds__(AFS_NUM_PM_SECTIONS *RASTER_SETS_PER_CTA)
area_function_raster_kernel(int n_sets, const double *__restrict__ params, afs_tube_frame *__restrict__ frames) {
  const int sec = threadIdx.x % AFS_NUM_PM_SECTIONS;
  const int set = blockIdx.x * RASTER_SETS_PER_CTA + threadIdx.x / AFS_NUM_PM_SECTIONS;
  if (set >= n_sets) return;
  double q[16];
#pragma unroll
  for (int i = 0; i < 16; i++) q[i] = __ldg(params + (size_t)set * 16 + i);
  const int n = AFS_NUM_PM_SECTIONS;
  const double width = __ddiv_rn(q[LVT], (double)n);
  const double step = __dmul_rn(width, 0.01);
  // replay the running sample position of the sections before this one (OneDimAreaFunction.cpp:94-116)
  double x = 0.0;
  for (int i = 0; i < sec; i++) {
    const double end = __dmul_rn((double)(i + 1), width);
    while (x < end) x = __dadd_rn(x, step);
  }
  double amin = DBL_MAX;
  const double end = __dmul_rn((double)(sec + 1), width);
  while (x < end) {
    const double a = area_at(q, x);
    if (a < amin) amin = a;
    x = __dadd_rn(x, step);
  }
  afs_tube_frame *f = frames + set;
  f->length_cm[sec] = width;
  f->area_cm2[sec] = amin < kMinArea ? kMinArea : amin;  // Tube::setPharynxMouthGeometry clamps on the way in (Tube.cpp:337)
  f->laterality[sec] = 0.0;
  int art;  // articulator from the sample position left after the scan (OneDimAreaFunction.cpp:118-134)
  if (x <= q[XP]) art = AFS_OTHER_ARTICULATOR;
  else if (x <= q[XIN] && __dadd_rn(x, width) < q[XIN]) art = AFS_TONGUE;
  else if (x <= q[XIN]) art = AFS_LOWER_INCISORS;
  else art = AFS_LOWER_LIP;
  f->articulator[sec] = art;
  if (sec == 0) f->incisor_pos_cm = q[XIN];
}

// every frame := the template (the rasteriser only writes the pharynx / mouth fields and the incisor position)
__global__ void fill_frames_kernel(size_t n_words, uint64_t *__restrict__ frames, const afs_tube_frame tpl) {
  constexpr size_t W = sizeof(afs_tube_frame) / 8;
  const uint64_t *t = reinterpret_cast<const uint64_t *>(&tpl);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (size_t)gridDim.x * blockDim.x) frames[i] = t[i % W];
}
cudaError_t launch_fill_frames(size_t n, afs_tube_frame *d_frames, const afs_tube_frame &tpl, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  static_assert(sizeof(afs_tube_frame) % 8 == 0, "frame is a whole number of 64-bit words");
  fill_frames_kernel<<<148 * 8, 256, 0, stream>>>(n * (sizeof(afs_tube_frame) / 8), reinterpret_cast<uint64_t *>(d_frames), tpl);
  return cudaGetLastError();
}

cudaError_t launch_area_function_raster(int n_sets, const double *d_params, afs_tube_frame *d_frames, cudaStream_t stream) {
  if (n_sets <= 0) return cudaSuccess;
  const int ctas = (n_sets + RASTER_SETS_PER_CTA - 1) / RASTER_SETS_PER_CTA;
  area_function_raster_kernel<<<ctas, AFS_NUM_PM_SECTIONS * RASTER_SETS_PER_CTA, 0, stream>>>(n_sets, d_params, d_frames);
  return cudaGetLastError();
}

}  // namespace afs
