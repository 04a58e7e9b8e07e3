// afs_host.cpp -- host-side helpers of the boundary: filter design and the
// parametric area-function rasteriser.  Neither is on the device hot path;
// they produce INPUTS of it (filter coefficients, tube frames).
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdint>
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

// ---------------------------------------------------------------------------------------------------
// LF glottal pulse train of Synthesizer::synthesizeVowelLf (Synthesizer.cpp:645-801): LfPulse::getPulse with
// getEpsilon / getAlpha / getB (LfPulse.cpp:37-245), the F0 / amplitude contours (TimeFunction::getValue,
// TimeFunction.cpp:77-133) and the "flutter".  Host code: ~100 pulses per vowel, libm in the reference's own
// expressions, so the device convolution starts from bit-identical input.
// ---------------------------------------------------------------------------------------------------
namespace {
struct TfNode { double time, value; };
double tf_value(const TfNode *nodes, int N, double t) {  // TimeFunction::getValue
  if (t < nodes[0].time) return nodes[0].value;
  if (t >= nodes[N - 1].time) return nodes[N - 1].value;
  int left = 0, right = N - 1;
  while (right - left > 1) {
    const int center = left + (right - left) / 2;
    if (nodes[center].time < t) left = center; else right = center;
  }
  double deno = nodes[right].time - nodes[left].time;
  if (deno < 0.000000001) deno = 0.000000001;
  const double a = (t - nodes[left].time) / deno;
  return nodes[left].value + a * (nodes[right].value - nodes[left].value);
}
// One period of the LF model on the unit interval (LfPulse.cpp:37-245).  The reference solves two scalar equations
// per period -- the return-phase constant by Newton's method, the growth constant of the open phase by bracketing +
// false position -- and then evaluates the two closed-form branches of the flow.  The numerical procedures are the
// reference's (same iterates in the same order => same bits); they are organised here around the residual of each
// equation.
struct ReturnPhase {   // 1 - exp(-eps c) = eps ta, c = 1 - te
  double c, ta;
  ReturnPhase(double ta_in, double te) : c(1.0 - te), ta(ta_in) {
    if (c < 0.001) c = 0.001;
    if (ta < 0.0001) ta = 0.0001;
    const double ta_max = c - 0.00001;
    if (ta > ta_max) ta = ta_max;
  }
  double residual(double eps) const { return 1.0 - exp(-eps * c) - eps * ta; }
  double slope(double eps) const { return c * exp(-eps * c) - ta; }
  // Newton from 1 / ta until the residual BEFORE the update is within 1e-5 (the reference's step counter never
  // counts, LfPulse.cpp:139-146; the guard only protects against NaN input)
  double solve() const {
    double eps = 1.0 / ta;
    for (int guard = 0; guard < 100000; guard++) {
      const double f = residual(eps);
      eps = eps - f / slope(eps);
      if (!(fabs(f) > 0.00001)) break;
    }
    return eps;
  }
};
struct OpenPhase {     // area balance of the pulse as a function of the growth constant
  double w, w2, s_te, c_te, te, tail_end, tail_te;
  OpenPhase(double tp, double te_in, double ta, double eps) : te(te_in) {
    w = M_PI / tp;
    w2 = w * w;
    s_te = sin(w * te);
    c_te = cos(w * te);
    const double lead = s_te * exp(eps * te);
    const double at_end = (lead * (-exp(-eps) / eps - exp(-eps))) / (eps * ta);
    const double at_te = (lead * (-exp(-eps * te) / eps - te * exp(-eps))) / (eps * ta);
    tail_end = at_end; tail_te = at_te;  // (two terms: the reference adds them one after the other)
  }
  double operator()(double a) const {
    const double q = w2 + a * a;
    return w / q + exp(a * te) * ((a * s_te - w * c_te) / q + tail_end - tail_te);
  }
  // bracket [0, n] with unit steps (at most 20), then at most 20 false-position updates
  double solve() const {
    double lo = 0.0, hi = 0.0, f_hi = 0.0;
    for (int n = 0; n < 20; n++) {
      hi += 1.0;
      f_hi = (*this)(hi);
      if (!(f_hi >= 0.0)) break;
    }
    if (f_hi >= 0) return 0.0;
    double root = 0.0;
    for (int n = 0; n < 20; n++) {
      const double f_lo = (*this)(lo);
      f_hi = (*this)(hi);
      root = lo - (f_lo * (hi - lo)) / (f_hi - f_lo);
      const double f_root = (*this)(root);
      (f_root < 0.0 ? hi : lo) = root;
      if (!(fabs(f_root) > 0.00001)) break;
    }
    return root;
  }
};
// LfPulse::getPulse(s, numSamples, false): one period of the glottal FLOW
void lf_pulse(double AMP, double OQ, double SQ, double TL, int numSamples, double *x) {
  const double te = OQ, tp = (te * SQ) / (1.0 + SQ);
  double ta = TL < 0.01 ? 0.01 : TL;
  if (ta > 1.0 - te) ta = 1.0 - te;
  const double eps = ReturnPhase(ta, te).solve();
  const double alpha = OpenPhase(tp, te, ta, eps).solve();
  const double w = 3.1415926 / tp;  // (sic: the flow branches use the truncated constant, LfPulse.cpp:81,214)
  const double den = w * w + alpha * alpha;
  auto rise = [&](double t) { return exp(alpha * t) * (alpha * sin(w * t) - w * cos(w * t)) + w; };
  const double B = AMP / (rise(tp) / den);  // LfPulse::getB
  const double flow_te = (B * rise(te)) / den;
  const double k = (B * exp(alpha * te) * sin(w * te) * exp(eps * te)) / (eps * ta);
  auto decay = [&](double t) { return k * (-exp(-eps * t) / eps - t * exp(-eps * 1.0)); };
  const double decay_te = decay(te);
  for (int i = 0; i < numSamples; i++) {
    const double t = (double)i / (double)numSamples;
    x[i] = t <= te ? (B * rise(t)) / den : flow_te + decay(t) - decay_te;
  }
}
}  // namespace

int lf_vowel_length(int is_long, double sampling_rate_hz) {
  return (int)(((is_long ? 650.0 : 350.0) / 1000.0) * sampling_rate_hz);
}

// Glottal flow of one vowel: flow[length] (the reference's 2048-sample ring buffer, unrolled: what it reads back at
// most 511 samples behind the write position is what was written, zeros before the first pulse at sample 10).
// Returns 0, or -1 when a pulse would not fit the reference's ring buffer (F0 below ~SR/1024).
int lf_vowel_flow(const afs_lf_pulse *lf, int is_long, double sampling_rate_hz, double *flow) {
  const double mx = lf->f0_hz;
  const TfNode f0_long[4] = {{0.0, 0.9 * mx}, {300.0, 1.00 * mx}, {450.0, 0.8 * mx}, {600.0, 0.7 * mx}};
  const TfNode amp_long[4] = {{0.0, 0.0}, {40.0, 500.0}, {400.0, 450.0}, {600.0, 0.0}};
  const TfNode f0_short[4] = {{0.0, 0.9 * mx}, {125.0, 1.0 * mx}, {126.0, 1.0 * mx}, {300.0, 0.82 * mx}};
  const TfNode amp_short[4] = {{0.0, 0.0}, {20.0, 500.0}, {200.0, 450.0}, {300.0, 0.0}};
  const TfNode *f0n = is_long ? f0_long : f0_short, *ampn = is_long ? amp_long : amp_short;
  const int length = lf_vowel_length(is_long, sampling_rate_hz);
  for (int i = 0; i < length; i++) flow[i] = 0.0;
  double pulse[1024];
  int next = 10;
  for (int i = 0; i < length; i++) {
    if (i != next) continue;
    const double t_s = (double)i / sampling_rate_hz, t_ms = t_s * 1000.0;
    const double AMP = tf_value(ampn, 4, t_ms);
    double F0 = tf_value(f0n, 4, t_ms);
    F0 += 0.5 * (F0 / 100.0) * (sin(2.0 * M_PI * 12.7 * t_s) + sin(2.0 * M_PI * 7.1 * t_s) + sin(2.0 * M_PI * 4.7 * t_s));
    const int n = (int)(sampling_rate_hz / F0);
    if (n < 1 || n > 1024) return -1;
    lf_pulse(AMP, lf->oq, lf->sq, lf->tl, n, pulse);
    for (int k = 0; k < n && i + k < length; k++) flow[i + k] = pulse[k];
    next += n;
  }
  return 0;
}


// ---------------------------------------------------------------------------------------------------
// Trajectories of Synthesizer::playTargetSequence (Synthesizer.cpp:1299-1422, interpolateParameters :1286-1294):
// per output sample the 16 area-function parameters (cosine interpolation between four target shapes) and the
// glottis control parameters (F0 contour 100 -> 115 -> 105 -> 80 Hz, lung pressure 8000 dPa with 50 ms silence +
// 50 ms cosine fade-in and the reference's fade-out expression, operator-precedence slip included).  Host libm,
// the reference's own expressions (SAMPLING_RATE is an int there: every product below converts it the same way).
// ---------------------------------------------------------------------------------------------------
int target_sequence_length(const double *stationary_s, const double *transition_s, double sampling_rate_hz) {
  const double total = stationary_s[0] + transition_s[0] + stationary_s[1] + transition_s[1] + stationary_s[2] + transition_s[2] + stationary_s[3];
  const double n = sampling_rate_hz * total;
  return n > 0.0 && n < 2.0e9 ? (int)n : 0;
}

// The time base is shared by all sequences of a call: segment and cosine factor of every sample are computed once
// (`seg`, `cosv`; also what the reference evaluates 16 times per sample with the same argument).
void target_sequence_timebase(const double *stationary_s, const double *transition_s, const double *glottis6, double SR,
                              int numSamples, int *seg_out, double *cosv, double *glottis) {
  double gp[AFS_NUM_GLOTTIS_CONTROL];
  for (int i = 0; i < AFS_NUM_GLOTTIS_CONTROL; i++) gp[i] = glottis6[i];
  const double f0_Hz[4] = {100, 115, 105, 80};
  gp[1] = 8000.0;  // Glottis::PRESSURE (sensorDataToGlottisParams, Synthesizer.cpp:889-921)
  const double steady = gp[1];
  const double b[7] = {
      stationary_s[0],
      stationary_s[0] + transition_s[0],
      stationary_s[0] + transition_s[0] + stationary_s[1],
      stationary_s[0] + transition_s[0] + stationary_s[1] + transition_s[1],
      stationary_s[0] + transition_s[0] + stationary_s[1] + transition_s[1] + stationary_s[2],
      stationary_s[0] + transition_s[0] + stationary_s[1] + transition_s[1] + stationary_s[2] + transition_s[2],
      stationary_s[0] + transition_s[0] + stationary_s[1] + transition_s[1] + stationary_s[2] + transition_s[2] + stationary_s[3],
  };
  const double total = b[6];
  for (int i = 0; i < numSamples; i++) {
    if (i < 0.1 * SR) {
      if (i < 0.05 * SR) gp[1] = 0.0;
      else gp[1] = steady / 2 * cos((0.1 * SR - i) / (0.05 * SR) * M_PI) + steady / 2;
    }
    if (i < b[1] * SR) gp[0] = (f0_Hz[0] + f0_Hz[1]) / 2 + (f0_Hz[1] - f0_Hz[0]) / 2 * cos((b[1] * SR - i) / (b[1] * SR) * M_PI);
    else if (i < b[3] * SR) gp[0] = (f0_Hz[2] + f0_Hz[1]) / 2 + (f0_Hz[2] - f0_Hz[1]) / 2 * cos((b[3] * SR - i) / ((b[3] - b[1]) * SR) * M_PI);
    else gp[0] = (f0_Hz[3] + f0_Hz[2]) / 2 + (f0_Hz[3] - f0_Hz[2]) / 2 * cos((b[6] * SR - i) / ((b[6] - b[3]) * SR) * M_PI);
    // 0, 2, 4, 6: stationary shape seg / 2; 1, 3, 5: transition between shapes (seg - 1) / 2 and (seg + 1) / 2; -1: keep
    int seg = -1;
    for (int k = 0; k < 7; k++) if (i <= b[k] * SR) { seg = k; break; }
    double c = 0.0;
    if (seg >= 0 && (seg & 1)) {
      const double t0 = b[seg - 1] * SR, t1 = b[seg] * SR, tX = i;
      c = cos((t1 - tX) / (t1 - t0) * M_PI);
    }
    // (seg -1 = keep the last values: past the end of the last stationary part, which rounding can put one sample
    //  before the end.  Passed on as the previous sample's (segment, cosine): the same values, and the interpolation
    //  needs no state)
    if (seg < 0 && i > 0) { seg = seg_out[i - 1]; c = cosv[i - 1]; }
    if (seg < 0) seg = 0;
    seg_out[i] = seg; cosv[i] = c;
    if (i > (total - 0.1) * SR) gp[1] = -steady / 2 * cos((total * SR - i) / (total - 0.1 * SR) * M_PI) + steady / 2;
    for (int k = 0; k < AFS_NUM_GLOTTIS_CONTROL; k++) glottis[(size_t)i * AFS_NUM_GLOTTIS_CONTROL + k] = gp[k];
  }
}

// interpolateParameters (Synthesizer.cpp:1286-1294) along the time base, one sequence
void target_sequence_params(const double *shapes4x16, const int *seg, const double *cosv, int numSamples, double *params) {
  const double *sh[4] = {shapes4x16, shapes4x16 + 16, shapes4x16 + 32, shapes4x16 + 48};
  double cur[16];
  for (int j = 0; j < 16; j++) cur[j] = sh[0][j];
  for (int i = 0; i < numSamples; i++) {
    const int s = seg[i];
    if (s >= 0) {
      if ((s & 1) == 0) { for (int j = 0; j < 16; j++) cur[j] = sh[s / 2][j]; }
      else {
        const double *p0 = sh[(s - 1) / 2], *p1 = sh[(s + 1) / 2];
        const double c = cosv[i];
        for (int j = 0; j < 16; j++) cur[j] = (p1[j] - p0[j]) / 2 * c + (p1[j] + p0[j]) / 2;
      }
    }
    for (int j = 0; j < 16; j++) params[(size_t)i * 16 + j] = cur[j];
  }
}

void target_sequence_trajectory(const double *shapes4x16, const double *stationary_s, const double *transition_s,
                                const double *glottis6, double SR, int numSamples, double *params, double *glottis) {
  int *seg = new int[(size_t)numSamples];
  double *cosv = new double[(size_t)numSamples];
  target_sequence_timebase(stationary_s, transition_s, glottis6, SR, numSamples, seg, cosv, glottis);
  target_sequence_params(shapes4x16, seg, cosv, numSamples, params);
  delete[] seg;
  delete[] cosv;
}

}  // namespace afs

extern "C" {

int afs_target_sequence_length(const double stationary_s[4], const double transition_s[3], double sampling_rate_hz) {
  if (!stationary_s || !transition_s || !(sampling_rate_hz > 0.0)) return 0;
  return afs::target_sequence_length(stationary_s, transition_s, sampling_rate_hz);
}

int afs_target_sequence_trajectory(const double *shapes, const double stationary_s[4], const double transition_s[3],
                                   const double glottis_control[AFS_NUM_GLOTTIS_CONTROL], double sampling_rate_hz,
                                   double *params, double *glottis) {
  if (!shapes || !stationary_s || !transition_s || !glottis_control || !params || !glottis || !(sampling_rate_hz > 0.0))
    return AFS_ERR_BAD_ARGUMENT;
  const int n = afs::target_sequence_length(stationary_s, transition_s, sampling_rate_hz);
  if (n < 1) return AFS_ERR_BAD_ARGUMENT;
  afs::target_sequence_trajectory(shapes, stationary_s, transition_s, glottis_control, sampling_rate_hz, n, params, glottis);
  return AFS_OK;
}

int afs_lf_glottal_flow(const afs_lf_pulse *pulse, int is_long_vowel, double sampling_rate_hz, double *flow) {
  if (!pulse || !flow || !(sampling_rate_hz > 0.0) || !(pulse->f0_hz > 0.0) || !(pulse->oq > 0.0) || !(pulse->oq < 1.0) ||
      !(pulse->sq > 0.0))
    return AFS_ERR_BAD_ARGUMENT;
  return afs::lf_vowel_flow(pulse, is_long_vowel, sampling_rate_hz, flow) ? AFS_ERR_BAD_ARGUMENT : AFS_OK;
}

int afs_audio_to_int16(const double *audio, int64_t n, int16_t *pcm) {
  if (n < 0 || (n > 0 && (!audio || !pcm))) return AFS_ERR_BAD_ARGUMENT;
  for (int64_t i = 0; i < n; i++) {
    const double v = audio[i];
    int16_t s = (int16_t)(int)(v * 32767.0);  // Signal16::setValue((double)newSignal[i] * SHRT_MAX)
    if (v > 1.0) s = 32767;
    if (v < -1.0) s = -32768;
    if (!(v == v)) s = 0;  // NaN: undefined in the reference, silence here
    pcm[i] = s;
  }
  return AFS_OK;
}

int afs_write_wav_int16(const char *path, const int16_t *pcm, int64_t n_frames, int n_channels, int sample_rate_hz) {
  if (!path || n_frames < 0 || n_channels < 1 || n_channels > 64 || sample_rate_hz < 1 || (n_frames > 0 && !pcm))
    return AFS_ERR_BAD_ARGUMENT;
  const uint64_t data_bytes = (uint64_t)n_frames * (uint64_t)n_channels * 2u;
  if (data_bytes > 0xffffffffu - 36u) return AFS_ERR_BAD_ARGUMENT;  // RIFF sizes are 32 bit
  FILE *fh = fopen(path, "wb");
  if (!fh) return AFS_ERR_BAD_ARGUMENT;
  unsigned char h[44];
  auto le32 = [&](int at, uint32_t v) { for (int i = 0; i < 4; i++) h[at + i] = (unsigned char)(v >> (8 * i)); };
  auto le16 = [&](int at, uint32_t v) { h[at] = (unsigned char)v; h[at + 1] = (unsigned char)(v >> 8); };
  memcpy(h, "RIFF", 4); le32(4, 36u + (uint32_t)data_bytes); memcpy(h + 8, "WAVEfmt ", 8);
  le32(16, 16); le16(20, 1); le16(22, (uint32_t)n_channels); le32(24, (uint32_t)sample_rate_hz);
  le32(28, (uint32_t)sample_rate_hz * (uint32_t)n_channels * 2u); le16(32, (uint32_t)n_channels * 2u); le16(34, 16);
  memcpy(h + 36, "data", 4); le32(40, (uint32_t)data_bytes);
  bool ok = fwrite(h, 1, 44, fh) == 44;
  if (ok && data_bytes) ok = fwrite(pcm, 1, (size_t)data_bytes, fh) == (size_t)data_bytes;  // little-endian host
  ok = (fclose(fh) == 0) && ok;
  return ok ? AFS_OK : AFS_ERR_BAD_ARGUMENT;
}

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
