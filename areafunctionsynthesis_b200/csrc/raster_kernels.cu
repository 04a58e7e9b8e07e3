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
// Mapping: lane = parameter set (32 sets per CTA), warp w of RASTER_WARPS owns the sections w, w + RASTER_WARPS, ...
// A thread walks the running sample position from 0 to the end of its last section (cheap exact additions and
// comparisons) and evaluates A(x) only inside its own sections.  All lanes of a warp are in the same section at the
// same time, and A(x) is evaluated branch-free (the segment only selects the operands of ONE pow + cos), so the
// five-way segment choice costs no divergence.  Parameters are read through the read-only path (16 doubles per
// set), the frame fields are written once.
#include <cfloat>

#include "afs_internal.h"

namespace afs {

namespace {
enum { LLAR, ALAR, POWLAR, XP, AP, POWP, XC, AC, POWC, XA, AA, POWA, XIN, AIN, LVT, ALIP };

// OneDimAreaFunction::calculateArea (OneDimAreaFunction.cpp:23-61): the raised-cosine segment that contains x only
// selects the operands; the arithmetic is the reference's expression, operation by operation.
__device__ __forceinline__ double area_at(const double *q, double x) {
  const bool s0 = x <= q[LLAR], s1 = x <= q[XP], s2 = x <= q[XC], s3 = x <= q[XA], s4 = x <= q[XIN];
  // segment [x0, x1] with areas A0 -> A1 and exponent p
  const double x1 = s1 ? q[XP] : s2 ? q[XC] : s3 ? q[XA] : q[XIN];
  const double x0 = s1 ? q[LLAR] : s2 ? q[XP] : s3 ? q[XC] : q[XA];
  const double A1 = s1 ? q[AP] : s2 ? q[AC] : s3 ? q[AA] : q[AIN];
  const double A0 = s1 ? q[ALAR] : s2 ? q[AP] : s3 ? q[AC] : q[AA];
  const double pw = s1 ? q[POWLAR] : s2 ? q[POWP] : s3 ? q[POWC] : q[POWA];
  double a = (A1 + A0) / 2 + (A1 - A0) / 2 * cos(kPi * pow((x1 - x) / (x1 - x0), pw));
  a = s0 ? q[ALAR] : (s4 ? a : q[ALIP]);
  return a < 0.0 ? 0.0 : a;
}
}  // namespace

constexpr int RASTER_WARPS = 5;  // 8 sections per thread
static_assert(AFS_NUM_PM_SECTIONS % RASTER_WARPS == 0, "sections are dealt round robin to the warps");

__global__ void __launch_bounds__(32 * RASTER_WARPS)
area_function_raster_kernel(int n_sets, const double *__restrict__ params, afs_tube_frame *__restrict__ frames) {
  const int w = threadIdx.x >> 5;
  const int set_raw = blockIdx.x * 32 + (threadIdx.x & 31);
  const bool active = set_raw < n_sets;
  const int set = active ? set_raw : n_sets - 1;
  double q[16];
#pragma unroll
  for (int i = 0; i < 16; i++) q[i] = __ldg(params + (size_t)set * 16 + i);
  const int n = AFS_NUM_PM_SECTIONS;
  const double width = __ddiv_rn(q[LVT], (double)n);
  const double step = __dmul_rn(width, 0.01);
  afs_tube_frame *f = frames + set;
  // the running sample position of the reference (OneDimAreaFunction.cpp:94-116): `x += step` inside
  // `while (x < (i + 1) * width)`, never reset between sections
  double x = 0.0;
  for (int sec = 0; sec <= n - RASTER_WARPS + w; sec++) {
    const double end = __dmul_rn((double)(sec + 1), width);
    if (sec % RASTER_WARPS != w) {  // (warp-uniform) not mine: only advance the position
      while (x < end) x = __dadd_rn(x, step);
      continue;
    }
    double amin = DBL_MAX;
    while (x < end) {
      const double a = area_at(q, x);
      if (a < amin) amin = a;
      x = __dadd_rn(x, step);
    }
    if (!active) continue;
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

// Synthesizer::interpolateParameters (Synthesizer.cpp:1286-1294) along the time base of playTargetSequence
// (:1299-1422), for a batch of sequences that share the timing: params[b][i][j] from the four target shapes of
// sequence b and the per-sample (segment, cosine) of the time base (afs_host.cpp: target_sequence_timebase).
// Same expression as the reference, no FMA contraction in this translation unit => the same bits.
__global__ void target_params_kernel(const int n_seq, const int ns, const double *__restrict__ shapes, const int32_t *__restrict__ seg,
                                     const double *__restrict__ cosv, double *__restrict__ params) {
  const size_t total = (size_t)n_seq * (size_t)ns * 16;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx & 15);
    const size_t bi = idx >> 4;
    const int i = (int)(bi % (size_t)ns), b = (int)(bi / (size_t)ns);
    const double *sh = shapes + (size_t)b * 64;
    const int s = seg[i];
    double v;
    if ((s & 1) == 0) v = sh[(s / 2) * 16 + j];
    else {
      const double p0 = sh[((s - 1) / 2) * 16 + j], p1 = sh[((s + 1) / 2) * 16 + j];
      v = (p1 - p0) / 2 * cosv[i] + (p1 + p0) / 2;
    }
    params[idx] = v;
  }
}
cudaError_t launch_target_params(int n_seq, int ns, const double *d_shapes, const int32_t *d_seg, const double *d_cosv,
                                 double *d_params, cudaStream_t stream) {
  if (n_seq <= 0 || ns <= 0) return cudaSuccess;
  target_params_kernel<<<148 * 8, 256, 0, stream>>>(n_seq, ns, d_shapes, d_seg, d_cosv, d_params);
  return cudaGetLastError();
}
// rows 1 .. n-1 of a [n][row_doubles] array = row 0 (the glottis trajectory is the same for every sequence)
__global__ void broadcast_row_kernel(const size_t n, const size_t row_doubles, double *__restrict__ a) {
  const size_t total = (n - 1) * row_doubles;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x)
    a[row_doubles + idx] = a[idx % row_doubles];
}
cudaError_t launch_broadcast_row(size_t n, size_t row_doubles, double *d_a, cudaStream_t stream) {
  if (n <= 1 || row_doubles == 0) return cudaSuccess;
  broadcast_row_kernel<<<148 * 8, 256, 0, stream>>>(n, row_doubles, d_a);
  return cudaGetLastError();
}

cudaError_t launch_area_function_raster(int n_sets, const double *d_params, afs_tube_frame *d_frames, cudaStream_t stream) {
  if (n_sets <= 0) return cudaSuccess;
  const int ctas = (n_sets + 31) / 32;
  area_function_raster_kernel<<<ctas, 32 * RASTER_WARPS, 0, stream>>>(n_sets, d_params, d_frames);
  return cudaGetLastError();
}

}  // namespace afs
