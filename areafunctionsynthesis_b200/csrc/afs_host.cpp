// afs_host.cpp -- host-side helpers of the boundary: filter design and the
// parametric area-function rasteriser.  Neither is on the device hot path;
// they produce INPUTS of it (filter coefficients, tube frames).
#include <cfloat>
#include <cmath>
#include <cstring>

#include "../../include/afs_b200.h"

namespace afs {

// Bin- and sample-dependent (tube-independent) factors of TlModel::getImpulseResponse, evaluated with the
// host libm in the reference's own expressions so that the device result can be bit-identical:
//   twiddles  the (ur, ui) sequence of complexFFT (Dsp.cpp:270-300): stage with half-length le2 owns
//             entries [le2-1, 2*le2-1), entry j-1 being (sr + i*si)^(j-1) by repeated multiplication;
//   radiation getRadiationCharacteristic (TlModel.cpp:685-697) at omega = 2 pi f0 i (TlModel.cpp:507-513);
//   window    getImpulseResponseWindow (TlModel.cpp:70-80).
void impulse_response_tables(double f0_hz, int length_exponent, int num_freq, double *tw_re, double *tw_im,
                             double *rad_re, double *rad_im, double *window) {
  const int N = 1 << length_exponent;
  for (int l = 1; l <= length_exponent; l++) {
    const int le2 = (1 << l) / 2;
    double ur = 1.0, ui = 0.0;
    const double sr = cos(M_PI / (double)le2), si = -sin(M_PI / (double)le2);
    for (int j = 1; j <= le2; j++) {
      tw_re[le2 - 1 + j - 1] = ur;
      tw_im[le2 - 1 + j - 1] = ui;
      const double tr = ur;
      ur = tr * sr - ui * si;
      ui = tr * si + ui * sr;
    }
  }
  for (int i = 0; i < num_freq; i++) {
    const double omega = 2.0 * M_PI * f0_hz * (double)i;
    const double hr = 0.0, hi = (1.14e-3 * omega) / (4.0 * M_PI * 25.0);
    const double angle = (omega * 25.0) / 3.5e4;
    const double cr = cos(angle), ci = -sin(angle);
    rad_re[i] = hr * cr - hi * ci;  // g++ inline std::complex product
    rad_im[i] = hr * ci + hi * cr;
  }
  for (int i = 0; i < N; i++) window[i] = 0.5 - 0.5 * cos(M_PI + (M_PI * (double)i) / (double)(N - 1));
}

// Recursive Chebyshev type-I low-pass design, 0.5 % pass-band ripple, by
// cascading two-pole stages (bilinear transform + low-pass-to-low-pass warp),
// the classic table-driven recipe the reference uses in
// IirFilter::createChebyshev (IirFilter.cpp:286-432).  Output convention:
// y[n] = sum_i a[i] x[n-i] + sum_{i>=1} b[i] y[n-i], unit DC gain.
void chebyshev_lowpass(double cutoff_ratio, int num_poles, double *a_out, double *b_out) {
  enum { MAXO = 32 };
  const double ripple_percent = 0.5;
  double a[MAXO + 3], b[MAXO + 3], ta[MAXO + 3], tb[MAXO + 3];
  if (num_poles & 1) num_poles++;
  if (num_poles > MAXO - 2) num_poles = MAXO - 2;
  for (int i = 0; i < MAXO + 3; i++) a[i] = b[i] = 0.0;
  a[2] = 1.0;
  b[2] = 1.0;

  for (int stage = 1; stage <= num_poles / 2; stage++) {
    // pole on the unit circle ...
    const double ang = M_PI / (2.0 * num_poles) + (M_PI * (stage - 1)) / (double)num_poles;
    double pr = -cos(ang), pi = sin(ang);
    // ... warped onto the Chebyshev ellipse
    {
      const double g = 100.0 / (100.0 - ripple_percent);
      const double es = sqrt(g * g - 1.0);
      const double vx = (1.0 / (double)num_poles) * log((1.0 / es) + sqrt((1.0 / (es * es)) + 1));
      double kx = (1.0 / (double)num_poles) * log((1.0 / es) + sqrt((1.0 / (es * es)) - 1));
      kx = 0.5 * (exp(kx) + exp(-kx));
      pr = pr * (0.5 * (exp(vx) - exp(-vx))) / kx;
      pi = pi * (0.5 * (exp(vx) + exp(-vx))) / kx;
    }
    // s -> z (bilinear, prototype cutoff 1 rad/s)
    const double t = 2.0 * tan(0.5);
    const double w = 2.0 * M_PI * cutoff_ratio;
    const double m = pr * pr + pi * pi;
    double d = 4.0 - 4.0 * pr * t + m * t * t;
    const double x0 = (t * t) / d, x1 = (2.0 * t * t) / d, x2 = (t * t) / d;
    const double y1 = (8.0 - 2.0 * m * t * t) / d;
    const double y2 = (-4.0 - 4.0 * pr * t - m * t * t) / d;
    // low-pass -> low-pass frequency warp
    const double k = sin(0.5 - 0.5 * w) / sin(0.5 + 0.5 * w);
    d = 1.0 + y1 * k - y2 * k * k;
    const double A0 = (x0 - x1 * k + x2 * k * k) / d;
    const double A1 = (-2.0 * x0 * k + x1 + x1 * k * k - 2.0 * x2 * k) / d;
    const double A2 = (x0 * k * k - x1 * k + x2) / d;
    const double B1 = (2.0 * k + y1 + y1 * k * k - 2.0 * y2 * k) / d;
    const double B2 = (-(k * k) - y1 * k + y2) / d;
    // fold the stage into the cascade polynomials
    memcpy(ta, a, sizeof a);
    memcpy(tb, b, sizeof b);
    for (int i = 2; i <= MAXO; i++) {
      a[i] = A0 * ta[i] + A1 * ta[i - 1] + A2 * ta[i - 2];
      b[i] = tb[i] - B1 * tb[i - 1] - B2 * tb[i - 2];
    }
  }
  b[2] = 0.0;
  for (int i = 0; i <= MAXO - 2; i++) { a[i] = a[i + 2]; b[i] = -b[i + 2]; }
  double sa = 0.0, sb = 0.0;
  for (int i = 0; i <= MAXO - 2; i++) { sa += a[i]; sb += b[i]; }
  const double gain = sa / (1.0 - sb);
  for (int i = 0; i <= MAXO - 2; i++) a[i] /= gain;
  for (int i = 0; i <= num_poles; i++) { a_out[i] = a[i]; b_out[i] = b[i]; }
}

namespace {
enum { LLAR, ALAR, POWLAR, XP, AP, POWP, XC, AC, POWC, XA, AA, POWA, XIN, AIN, LVT, ALIP };

// Stone/Marxen/Birkholz 2018 parametric area function A(x): raised-cosine
// segments between control points (OneDimAreaFunction::calculateArea,
// OneDimAreaFunction.cpp:23-61).
double area_at(const double *q, double x) {
  double a;
  if (x <= q[LLAR]) a = q[ALAR];
  else if (x <= q[XP]) a = (q[AP] + q[ALAR]) / 2 + (q[AP] - q[ALAR]) / 2 * cos(M_PI * pow((q[XP] - x) / (q[XP] - q[LLAR]), q[POWLAR]));
  else if (x <= q[XC]) a = (q[AC] + q[AP]) / 2 + (q[AC] - q[AP]) / 2 * cos(M_PI * pow((q[XC] - x) / (q[XC] - q[XP]), q[POWP]));
  else if (x <= q[XA]) a = (q[AA] + q[AC]) / 2 + (q[AA] - q[AC]) / 2 * cos(M_PI * pow((q[XA] - x) / (q[XA] - q[XC]), q[POWC]));
  else if (x <= q[XIN]) a = (q[AIN] + q[AA]) / 2 + (q[AIN] - q[AA]) / 2 * cos(M_PI * pow((q[XIN] - x) / (q[XIN] - q[XA]), q[POWA]));
  else a = q[ALIP];
  return a < 0.0 ? 0.0 : a;
}
}  // namespace
}  // namespace afs

extern "C" {

int afs_chebyshev_lowpass(double cutoff_ratio, int num_poles, double *a, double *b) {
  if (!a || !b || num_poles < 2 || num_poles > 16 || (num_poles & 1) || !(cutoff_ratio > 0.0) || !(cutoff_ratio < 0.5))
    return AFS_ERR_BAD_ARGUMENT;
  afs::chebyshev_lowpass(cutoff_ratio, num_poles, a, b);
  return AFS_OK;
}

// OneDimAreaFunction::calculateOneDimTubeFunction (OneDimAreaFunction.cpp:75-138):
// 40 equal-length sections, each taking the MINIMUM of A(x) sampled every
// 1/100 section width (running x accumulates, as in the reference), and the
// articulator label derived from the sample position left after the scan.
int afs_area_function_to_frame(const double params[16], afs_tube_frame *frame) {
  if (!params || !frame) return AFS_ERR_BAD_ARGUMENT;
  const double *q = params;
  const int n = AFS_NUM_PM_SECTIONS;
  const double width = q[afs::LVT] / n;
  const double step = width * 0.01;
  double x = 0.0;
  for (int i = 0; i < n; i++) {
    double amin = DBL_MAX;
    while (x < (i + 1) * width) {
      const double a = afs::area_at(q, x);
      if (a < amin) amin = a;
      x += step;
    }
    frame->length_cm[i] = width;
    // Tube::setPharynxMouthGeometry clamps on the way in (Tube.cpp:337)
    frame->area_cm2[i] = amin < 0.1e-2 ? 0.1e-2 : amin;
    frame->laterality[i] = 0.0;
    int art;
    if (x <= q[afs::XP]) art = AFS_OTHER_ARTICULATOR;
    else if (x <= q[afs::XIN] && x + width < q[afs::XIN]) art = AFS_TONGUE;
    else if (x <= q[afs::XIN]) art = AFS_LOWER_INCISORS;
    else art = AFS_LOWER_LIP;
    frame->articulator[i] = art;
  }
  frame->incisor_pos_cm = q[afs::XIN];
  return AFS_OK;
}

}  // extern "C"
