// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, not product code.
//
// extern "C" wrapper around the UNMODIFIED reference classes (compiled from
// /root/reference/src/Backend where they lie, see oracle/Makefile) so that
// tests and bench.py's cpu_baseline leg can run the real reference.
// The driver loop below restates Synthesizer::synthesizeSignalTds
// (Synthesizer.cpp:515-639) because Synthesizer itself needs wxWidgets and
// PortAudio; everything it calls is the reference's own code.
#include "../include/afs_b200.h"

#include "TdsModel.h"
#include "TlModel.h"
#include "TriangularGlottis.h"
#include "OneDimAreaFunction.h"
#include "IirFilter.h"
#include "Tube.h"
#include "LfPulse.h"
#include "TimeFunction.h"

#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static void frame_to_tube(const afs_tube_frame *f, Tube *t) {
  Tube::Articulator art[Tube::NUM_PHARYNX_MOUTH_SECTIONS];
  for (int i = 0; i < Tube::NUM_PHARYNX_MOUTH_SECTIONS; i++) art[i] = (Tube::Articulator)f->articulator[i];
  t->setPharynxMouthGeometry(f->length_cm, f->area_cm2, art, f->laterality, f->incisor_pos_cm);
  t->setGlottisGeometry(f->glottis_length_cm, f->glottis_area_cm2);
  t->setVelumOpening(f->velum_area_cm2);
  t->setAspirationStrength(f->aspiration_dB);
}

extern "C" {

int ref_sampling_rate() { return SAMPLING_RATE; }
int ref_tds_sr_factor() { return TDS_SR_FACTOR; }

// OneDimAreaFunction::setParameters + calculateOneDimTubeFunction -> frame
void ref_area_function_to_frame(const double *params16, afs_tube_frame *frame) {
  OneDimAreaFunction af;
  OneDimAreaFunction::Param p[OneDimAreaFunction::NUM_AF_PARAMS];
  af.getParameters(p);
  for (int i = 0; i < OneDimAreaFunction::NUM_AF_PARAMS; i++) p[i].x = params16[i];
  af.setParameters(p);
  Tube t;
  af.calculateOneDimTubeFunction(&t);
  for (int i = 0; i < Tube::NUM_PHARYNX_MOUTH_SECTIONS; i++) {
    // NB: Tube clamps the area to MIN_AREA_CM2 on the way in (Tube.cpp:337).
    frame->area_cm2[i] = t.pharynxMouthSection[i].area_cm2;
    frame->length_cm[i] = t.pharynxMouthSection[i].length_cm;
    frame->laterality[i] = t.pharynxMouthSection[i].laterality;
    frame->articulator[i] = (int)t.pharynxMouthSection[i].articulator;
  }
  frame->incisor_pos_cm = t.teethPosition_cm;
}

// Dump all 93 sections of the Tube a frame expands to (area,length,volume,pos,Mw,Bw,Kw).
void ref_frame_sections(const afs_tube_frame *frame, double *out93x7) {
  Tube t;
  frame_to_tube(frame, &t);
  for (int i = 0; i < Tube::NUM_SECTIONS; i++) {
    Tube::Section *s = t.section[i];
    double *o = out93x7 + 7 * i;
    o[0] = s->area_cm2; o[1] = s->length_cm; o[2] = s->volume_cm3; o[3] = s->pos_cm;
    o[4] = s->wallMass_cgs; o[5] = s->wallResistance_cgs; o[6] = s->wallStiffness_cgs;
  }
}

static void set_tl_options(TlModel *tl, const afs_tl_options *o) {
  tl->options.radiation = (TlModel::RadiationType)o->radiation;
  tl->options.boundaryLayer = o->boundary_layer;
  tl->options.heatConduction = o->heat_conduction;
  tl->options.softWalls = o->soft_walls;
  tl->options.hagenResistance = o->hagen_resistance;
  tl->options.innerLengthCorrections = o->inner_length_corrections;
  tl->options.lumpedElements = o->lumped_elements;
  tl->options.paranasalSinuses = o->paranasal_sinuses;
  tl->options.piriformFossa = o->piriform_fossa;
  tl->options.staticPressureDrops = o->static_pressure_drops;
  tl->setLungPressure(o->lung_pressure_dPa);
}

// TlModel::getSpectrum on a fresh TlModel per tube (a reused
// object carries hidden lung-pressure state).  spectrum_length picks the grid:
// f0 = SAMPLING_RATE / spectrum_length; returns numFreq written per tube.
int ref_tl_spectrum(const afs_tl_options *opts, int n_tubes, const afs_tube_frame *tubes,
                    int type, int section, int spectrum_length, int max_bins,
                    double *re, double *im, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  std::vector<int> nf(n_threads, 0);
  auto work = [&](int tid) {
    TlModel *tl = new TlModel();  // 6 MB object: heap
    ComplexSignal s;
    for (int b = tid; b < n_tubes; b += n_threads) {
      set_tl_options(tl, opts);
      frame_to_tube(&tubes[b], &tl->tube);
      tl->getSpectrum((TlModel::SpectrumType)type, &s, spectrum_length, section);
      int n = 0;
      // numFreq is private; recover it as in TlModel.cpp:131-143.
      double f0 = (double)SAMPLING_RATE / (double)spectrum_length;
      n = (int)(10100.0 / f0);
      if (n > 2048) n = 2048;  // MAX_NUM_FREQ of the patched build
      if (n >= spectrum_length / 2) n = spectrum_length / 2 - 1;
      if (n > max_bins) n = max_bins;
      for (int i = 0; i < n; i++) {
        re[(size_t)b * max_bins + i] = s.re[i];
        im[(size_t)b * max_bins + i] = s.im[i];
      }
      nf[tid] = n;
    }
    delete tl;
  };
  std::vector<std::thread> th;
  for (int t = 1; t < n_threads; t++) th.emplace_back(work, t);
  work(0);
  for (auto &t : th) t.join();
  return nf[0];
}

// TlModel::getImpulseResponse (TlModel.cpp:90-111) on a fresh TlModel per tube; out[n][1 << length_exponent].
void ref_tl_impulse_response(const afs_tl_options *opts, int n_tubes, const afs_tube_frame *tubes,
                             int length_exponent, double *out) {
  const int N = 1 << length_exponent;
  for (int b = 0; b < n_tubes; b++) {
    TlModel *tl = new TlModel();
    set_tl_options(tl, opts);
    frame_to_tube(&tubes[b], &tl->tube);
    Signal h;
    tl->getImpulseResponse(&h, length_exponent);
    for (int i = 0; i < N; i++) out[(size_t)b * N + i] = h.x[i];
    delete tl;
  }
}

// ---------------------------------------------------------------------------------------------------
// Data::getVowelFormants / getFormantError / optimizeFormantsVowel (Frontend/Data.cpp:106-135, 149-196, 203-487).
// Data.cpp is wx-bound (progress dialog, printf, events), so the greedy coordinate search is restated here line by
// line over the reference's own OneDimAreaFunction and TlModel; dialogs, printing and the GUI refresh are left out.
// ---------------------------------------------------------------------------------------------------
static bool ref_get_vowel_formants(TlModel *tlModel, OneDimAreaFunction *vtAf, double &F1, double &F2, double &F3,
                                   double &minArea_cm2) {
  const int MAX_FORMANTS = 3;
  double formantFreq[MAX_FORMANTS], formantBw[MAX_FORMANTS];
  int numFormants;
  bool frictionNoise, isClosure, isNasal;
  F1 = F2 = F3 = 0.0;
  minArea_cm2 = 0.0;
  Tube tubeFunction;
  vtAf->calculateOneDimTubeFunction(&tubeFunction);
  tlModel->tube = tubeFunction;
  tlModel->tube.setGlottisArea(0.0);
  minArea_cm2 = 10000.0;
  for (int i = Tube::FIRST_PHARYNX_SECTION; i < Tube::NUM_PHARYNX_MOUTH_SECTIONS; i++)  // sic: sections 25..39
    if (tlModel->tube.section[i]->area_cm2 < minArea_cm2) minArea_cm2 = tubeFunction.section[i]->area_cm2;
  tlModel->getFormants(formantFreq, formantBw, numFormants, MAX_FORMANTS, frictionNoise, isClosure, isNasal);
  if (numFormants < MAX_FORMANTS) return false;
  F1 = formantFreq[0]; F2 = formantFreq[1]; F3 = formantFreq[2];
  return true;
}

static double ref_formant_error(double c1, double c2, double c3, double t1, double t2, double t3) {
  const double EPSILON = 1.0;
  double e = 0.0;
  if (t3 > EPSILON) {
    e = (1.0 - c1 / t1) * (1.0 - c1 / t1) + (1.0 - c2 / t2) * (1.0 - c2 / t2) + (1.0 - c3 / t3) * (1.0 - c3 / t3);
    e /= 3.0;
  } else {
    e = (1.0 - c1 / t1) * (1.0 - c1 / t1) + (1.0 - c2 / t2) * (1.0 - c2 / t2);
    e /= 2.0;
  }
  e = sqrt(e);
  e = e * 100.0;
  return e;
}

// params16: in/out.  out3 = initial error, final error, number of runs.  trace (nullable) [max_runs][2]: changed
// parameter (-1: none) and error after every run.  formants_out (nullable) [3]: F1..F3 of the final parameters.
void ref_optimize_formants_vowel(const afs_tl_options *opts, double *params16, const double *target3,
                                 double maxParamChange_cm, double minAdvisedArea_cm2, const int32_t *paramFixed16,
                                 int max_runs, double *out3, double *trace, double *formants_out) {
  typedef OneDimAreaFunction AF;
  AF afObj;
  AF *vtAf = &afObj;
  TlModel *tlModel = new TlModel();
  set_tl_options(tlModel, opts);
  {
    AF::Param p[AF::NUM_AF_PARAMS];
    vtAf->getParameters(p);
    for (int i = 0; i < AF::NUM_AF_PARAMS; i++) p[i].x = params16[i];
    vtAf->setParameters(p);
  }
  const double targetF1 = target3[0], targetF2 = target3[1], targetF3 = target3[2];
  const int MAX_RUNS = max_runs;
  double F1, F2, F3, minArea_cm2;
  double changeStep[AF::NUM_AF_PARAMS];
  int stepsTaken[AF::NUM_AF_PARAMS];
  double currParamValue, bestError, currError, newError, bestParamChange;
  int bestParam, i;

  ref_get_vowel_formants(tlModel, vtAf, F1, F2, F3, minArea_cm2);
  double initialError = ref_formant_error(F1, F2, F3, targetF1, targetF2, targetF3);

  const double STEP_SIZE_CM = 0.05;
  int maxSteps = (int)((maxParamChange_cm / STEP_SIZE_CM) + 0.5);
  if (maxSteps < 1) maxSteps = 1;
  if (maxSteps > 40) maxSteps = 40;
  for (i = 0; i < AF::NUM_AF_PARAMS; i++) { changeStep[i] = 0.0; stepsTaken[i] = 0; }
  changeStep[AF::LLAR_CM] = STEP_SIZE_CM; changeStep[AF::XP_CM] = STEP_SIZE_CM; changeStep[AF::XC_CM] = STEP_SIZE_CM;
  changeStep[AF::XA_CM] = STEP_SIZE_CM; changeStep[AF::XIN_CM] = STEP_SIZE_CM; changeStep[AF::LVT_CM] = STEP_SIZE_CM;
  changeStep[AF::ALAR_CM2] = STEP_SIZE_CM * STEP_SIZE_CM; changeStep[AF::AP_CM2] = STEP_SIZE_CM * STEP_SIZE_CM;
  changeStep[AF::AC_CM2] = STEP_SIZE_CM * STEP_SIZE_CM; changeStep[AF::AA_CM2] = STEP_SIZE_CM * STEP_SIZE_CM;
  changeStep[AF::AIN_CM2] = STEP_SIZE_CM * STEP_SIZE_CM; changeStep[AF::ALIP_CM2] = STEP_SIZE_CM * STEP_SIZE_CM;
  changeStep[AF::POWLAR] = 0.1; changeStep[AF::POWP] = 0.1; changeStep[AF::POWC] = 0.1; changeStep[AF::POWA] = 0.1;
  for (i = 0; i < AF::NUM_AF_PARAMS; i++) if (paramFixed16 && paramFixed16[i]) changeStep[i] = 0.0;

  bool paramChanged = false;
  int runCounter = 0;
  AF::Param currentParameters[AF::NUM_AF_PARAMS];
  do {
    paramChanged = false;
    ref_get_vowel_formants(tlModel, vtAf, F1, F2, F3, minArea_cm2);
    currError = ref_formant_error(F1, F2, F3, targetF1, targetF2, targetF3);
    bestError = currError; bestParam = -1; bestParamChange = 0.0;
    vtAf->getParameters(currentParameters);
    for (i = 0; i < AF::NUM_AF_PARAMS; i++) {
      currParamValue = currentParameters[i].x;
      if (changeStep[i] > 0.0) {
        if (stepsTaken[i] < maxSteps) {
          currentParameters[i].x = currParamValue + changeStep[i];
          if (i == AF::XP_CM && currentParameters[i].x > currentParameters[AF::XC_CM].x) { changeStep[i] = currentParameters[AF::XC_CM].x - currentParameters[i].x; currentParameters[i].x = currentParameters[AF::XC_CM].x; }
          if (i == AF::XC_CM && currentParameters[i].x > currentParameters[AF::XA_CM].x) { changeStep[i] = currentParameters[AF::XA_CM].x - currentParameters[i].x; currentParameters[i].x = currentParameters[AF::XA_CM].x; }
          if (i == AF::XA_CM && currentParameters[i].x > currentParameters[AF::XIN_CM].x) { changeStep[i] = currentParameters[AF::XIN_CM].x - currentParameters[i].x; currentParameters[i].x = currentParameters[AF::XIN_CM].x; }
          if (i == AF::XIN_CM && currentParameters[i].x > currentParameters[AF::LVT_CM].x) { changeStep[i] = currentParameters[AF::LVT_CM].x - currentParameters[i].x; currentParameters[i].x = currentParameters[AF::LVT_CM].x; }
          vtAf->setParameters(currentParameters);
          ref_get_vowel_formants(tlModel, vtAf, F1, F2, F3, minArea_cm2);
          if (minArea_cm2 >= minAdvisedArea_cm2) {
            newError = ref_formant_error(F1, F2, F3, targetF1, targetF2, targetF3);
            if (newError < bestError) { bestError = newError; bestParam = i; bestParamChange = changeStep[i]; }
          }
        }
        if (stepsTaken[i] > -maxSteps) {
          currentParameters[i].x = currParamValue - changeStep[i];
          if (i == AF::XP_CM && currentParameters[i].x < currentParameters[AF::LLAR_CM].x) { changeStep[i] = currentParameters[AF::LLAR_CM].x - currentParameters[i].x; currentParameters[i].x = currentParameters[AF::LLAR_CM].x; }
          if (i == AF::XC_CM && currentParameters[i].x < currentParameters[AF::XP_CM].x) { changeStep[i] = currentParameters[AF::XP_CM].x - currentParameters[i].x; currentParameters[i].x = currentParameters[AF::XP_CM].x; }
          if (i == AF::XA_CM && currentParameters[i].x < currentParameters[AF::XC_CM].x) { changeStep[i] = currentParameters[AF::XC_CM].x - currentParameters[i].x; currentParameters[i].x = currentParameters[AF::XC_CM].x; }
          if (i == AF::XIN_CM && currentParameters[i].x < currentParameters[AF::XA_CM].x) { changeStep[i] = currentParameters[AF::XA_CM].x - currentParameters[i].x; currentParameters[i].x = currentParameters[AF::XA_CM].x; }
          if (i == AF::LVT_CM && currentParameters[i].x < currentParameters[AF::XIN_CM].x) { changeStep[i] = currentParameters[AF::XIN_CM].x - currentParameters[i].x; currentParameters[i].x = currentParameters[AF::XIN_CM].x; }
          vtAf->setParameters(currentParameters);
          ref_get_vowel_formants(tlModel, vtAf, F1, F2, F3, minArea_cm2);
          if (minArea_cm2 >= minAdvisedArea_cm2) {
            newError = ref_formant_error(F1, F2, F3, targetF1, targetF2, targetF3);
            if (newError < bestError) { bestError = newError; bestParam = i; bestParamChange = -changeStep[i]; }
          }
        }
        currentParameters[i].x = currParamValue;
        vtAf->setParameters(currentParameters);
      }
    }
    if ((bestParam != -1) && (bestError < currError)) {
      currentParameters[bestParam].x += bestParamChange;
      vtAf->setParameters(currentParameters);
      if (bestParamChange > 0.0) stepsTaken[bestParam]++; else stepsTaken[bestParam]--;
      paramChanged = true;
    }
    if (trace) { trace[2 * runCounter] = paramChanged ? (double)bestParam : -1.0; trace[2 * runCounter + 1] = bestError; }
    runCounter++;
  } while (paramChanged && (runCounter < MAX_RUNS));

  ref_get_vowel_formants(tlModel, vtAf, F1, F2, F3, minArea_cm2);
  double finalError = ref_formant_error(F1, F2, F3, targetF1, targetF2, targetF3);
  vtAf->getParameters(currentParameters);
  for (i = 0; i < AF::NUM_AF_PARAMS; i++) params16[i] = currentParameters[i].x;
  out3[0] = initialError; out3[1] = finalError; out3[2] = (double)runCounter;
  if (formants_out) { formants_out[0] = F1; formants_out[1] = F2; formants_out[2] = F3; }
  delete tlModel;
}

// ---------------------------------------------------------------------------------------------------
// Synthesizer::synthesizeVowelLf (Synthesizer.cpp:645-801): LF glottal pulses convolved with the vocal tract impulse
// response.  Synthesizer is wx/PortAudio-bound, so the loop is restated line by line over the reference's own
// TlModel, LfPulse, TimeFunction, IirFilter and Signal; the audio ring buffer becomes a plain output array.
// lf5 = AMP, F0, OQ, SQ, TL (LfPulse.h:25-29).  Returns the number of samples written.
// ---------------------------------------------------------------------------------------------------
int ref_lf_vowel_length(int isLongVowel) {
  double duration_ms = isLongVowel ? 650.0 : 350.0;
  return (int)((duration_ms / 1000.0) * (double)SAMPLING_RATE);
}

int ref_synthesize_vowel_lf(const afs_tl_options *opts, const afs_tube_frame *tubeFrame, const double *lf5,
                            int isLongVowel, int16_t *out, double *pressure_out, double *flow_out) {
  TlModel *tlModel = new TlModel();
  set_tl_options(tlModel, opts);
  frame_to_tube(tubeFrame, &tlModel->tube);
  LfPulse lfPulse;
  lfPulse.AMP = lf5[0]; lfPulse.F0 = lf5[1]; lfPulse.OQ = lf5[2]; lfPulse.SQ = lf5[3]; lfPulse.TL = lf5[4];

  const int NUM_F0_NODES = 4;
  const int NUM_AMP_NODES = 4;
  const int BUFFER_LENGTH = 2048;
  const int BUFFER_MASK = 2047;
  TimeFunction ampTimeFunction;
  TimeFunction f0TimeFunction;
  double duration_ms;
  Signal singlePulse;
  Signal pulseSignal(BUFFER_LENGTH);
  Signal pressureSignal(BUFFER_LENGTH);
  int i, k;
  int nextPulsePos = 10;
  int pulseLength;
  double t_s, t_ms;
  double sum;
  double filteredValue;

  if (isLongVowel) {
    duration_ms = 650.0;
    double max = lfPulse.F0;
    TimeFunction::Node f0[NUM_F0_NODES] = {{0.0, 0.9 * max}, {300.0, 1.00 * max}, {450.0, 0.8 * max}, {600.0, 0.7 * max}};
    TimeFunction::Node amp[NUM_AMP_NODES] = {{0.0, 0.0}, {40.0, 500.0}, {400.0, 450.0}, {600.0, 0.0}};
    ampTimeFunction.setNodes(amp, NUM_AMP_NODES);
    f0TimeFunction.setNodes(f0, NUM_F0_NODES);
  } else {
    duration_ms = 350.0;
    double max = lfPulse.F0;
    TimeFunction::Node f0[NUM_F0_NODES] = {{0.0, 0.9 * max}, {125.0, 1.0 * max}, {126.0, 1.0 * max}, {300.0, 0.82 * max}};
    TimeFunction::Node amp[NUM_AMP_NODES] = {{0.0, 0.0}, {20.0, 500.0}, {200.0, 450.0}, {300.0, 0.0}};
    ampTimeFunction.setNodes(amp, NUM_AMP_NODES);
    f0TimeFunction.setNodes(f0, NUM_F0_NODES);
  }
  int length = (int)((duration_ms / 1000.0) * (double)SAMPLING_RATE);

  const double NUM_LOWPASS_POLES = 6;
  IirFilter filter;
  filter.createChebyshev((double)SYNTHETIC_SPEECH_BANDWIDTH_HZ / (double)SAMPLING_RATE, false, (int)NUM_LOWPASS_POLES);

  const int IMPULSE_RESPONSE_EXPONENT = 9;
  const int IMPULSE_RESPONSE_LENGTH = 1 << IMPULSE_RESPONSE_EXPONENT;
  Signal impulseResponse(IMPULSE_RESPONSE_LENGTH);
  tlModel->getImpulseResponse(&impulseResponse, IMPULSE_RESPONSE_EXPONENT);
  impulseResponse *= 80.0 / lfPulse.F0;

  for (i = 0; i < length; i++) {
    t_s = (double)i / (double)SAMPLING_RATE;
    t_ms = t_s * 1000.0;
    if (i == nextPulsePos) {
      lfPulse.AMP = ampTimeFunction.getValue(t_ms);
      lfPulse.F0 = f0TimeFunction.getValue(t_ms);
      lfPulse.F0 += 0.5 * (lfPulse.F0 / 100.0) * (sin(2.0 * M_PI * 12.7 * t_s) + sin(2.0 * M_PI * 7.1 * t_s) + sin(2.0 * M_PI * 4.7 * t_s));
      pulseLength = (int)((double)SAMPLING_RATE / lfPulse.F0);
      lfPulse.getPulse(singlePulse, pulseLength, false);
      for (k = 0; k < pulseLength; k++) pulseSignal.x[(i + k) & BUFFER_MASK] = singlePulse.getValue(k);
      nextPulsePos += pulseLength;
    }
    sum = 0.0;
    for (k = 0; k < IMPULSE_RESPONSE_LENGTH; k++) sum += impulseResponse.x[k] * pulseSignal.x[(i - k) & BUFFER_MASK];
    pressureSignal.x[i & BUFFER_MASK] = sum;
    filteredValue = 2000.0 * filter.getOutputSample(pressureSignal.getValue(i));
    out[i] = (short)filteredValue;
    if (pressure_out) pressure_out[i] = sum;
    if (flow_out) flow_out[i] = pulseSignal.x[i & BUFFER_MASK];
  }
  delete tlModel;
  return length;
}

// ---------------------------------------------------------------------------------------------------
// Synthesizer::playTargetSequence (Synthesizer.cpp:1299-1422) with interpolateParameters (:1286-1294): four target
// shapes, cosine-interpolated parameter trajectories, F0 contour 100 -> 115 -> 105 -> 80 Hz, lung-pressure fade in /
// out, ONE synthesizeSignalTds(tube, glottisParams, 1, out) call per output sample.  Synthesizer is wx/PortAudio-
// bound: the trajectory is restated line by line (operator-precedence slips included), the tube of every sample comes
// from the reference's own OneDimAreaFunction, the driver loop is tds_one above.  glottis6: the control parameters of
// the glottis object (FREQUENCY and PRESSURE are overwritten).  audio [numSamples]; sample 0 is the priming call
// (synthesizeSignalTds returns without writing, Synthesizer.cpp:522-532) and reads 0 here.
// ---------------------------------------------------------------------------------------------------
struct RefProbes { int n; const int32_t *section; double *pressure, *inflow, *outflow; };  // per step, [n][steps]
static void tds_one(const afs_tds_options *o, int n_frames, const int32_t *frame_len, const afs_tube_frame *frames,
                    const double *glottis, unsigned seed, double *audio, double *flow, double *state_dump, int dump_step,
                    const RefProbes *probes = nullptr);

int ref_target_sequence_length(const double *stationary_s, const double *transition_s) {
  double totalTime_s = stationary_s[0] + transition_s[0] + stationary_s[1] + transition_s[1] + stationary_s[2] + transition_s[2] + stationary_s[3];
  int numSamples = SAMPLING_RATE * totalTime_s;
  return numSamples;
}

// params_out (nullable) [numSamples][16], glottis_out (nullable) [numSamples][6]: the trajectories
int ref_play_target_sequence(const afs_tds_options *opts, const double *shapes4x16, const double *stationary_s,
                             const double *transition_s, const double *glottis6, uint32_t seed, double *audio,
                             double *params_out, double *glottis_out) {
  typedef OneDimAreaFunction AF;
  double glottisParams[AFS_NUM_GLOTTIS_CONTROL];
  for (int i = 0; i < AFS_NUM_GLOTTIS_CONTROL; i++) glottisParams[i] = glottis6[i];
  double f0_Hz[4] = {100, 115, 105, 80};
  glottisParams[Glottis::PRESSURE] = 8000.0;  // sensorDataToGlottisParams (Synthesizer.cpp:889-921), no F0 wobble at t = 0
  double steadyStateLungPressure = glottisParams[Glottis::PRESSURE];
  double boundary_s[7] = {
      stationary_s[0],
      stationary_s[0] + transition_s[0],
      stationary_s[0] + transition_s[0] + stationary_s[1],
      stationary_s[0] + transition_s[0] + stationary_s[1] + transition_s[1],
      stationary_s[0] + transition_s[0] + stationary_s[1] + transition_s[1] + stationary_s[2],
      stationary_s[0] + transition_s[0] + stationary_s[1] + transition_s[1] + stationary_s[2] + transition_s[2],
      stationary_s[0] + transition_s[0] + stationary_s[1] + transition_s[1] + stationary_s[2] + transition_s[2] + stationary_s[3],
  };
  double totalTime_s = boundary_s[6];
  int numSamples = SAMPLING_RATE * totalTime_s;
  std::vector<afs_tube_frame> frames((size_t)numSamples);
  std::vector<double> gl((size_t)numSamples * AFS_NUM_GLOTTIS_CONTROL);
  AF vtAreaFunction;
  AF::Param currentParams[AF::NUM_AF_PARAMS];
  vtAreaFunction.getParameters(currentParams);
  auto interpolateParameters = [&](const double *p0, const double *p1, double t0, double t1, double tX) {
    for (int i = 0; i < AF::NUM_AF_PARAMS; i++) currentParams[i].x = (p1[i] - p0[i]) / 2 * cos((t1 - tX) / (t1 - t0) * M_PI) + (p1[i] + p0[i]) / 2;
  };
  for (int i = 0; i < numSamples; i++) {
    if (i < (double)0.1 * (double)SAMPLING_RATE) {
      if (i < (double)0.05 * (double)SAMPLING_RATE) glottisParams[Glottis::PRESSURE] = 0.0;
      else glottisParams[Glottis::PRESSURE] = steadyStateLungPressure / 2 * cos((0.1 * SAMPLING_RATE - i) / (0.05 * SAMPLING_RATE) * M_PI) + steadyStateLungPressure / 2;
    }
    if (i < (double)boundary_s[1] * SAMPLING_RATE)
      glottisParams[Glottis::FREQUENCY] = (f0_Hz[0] + f0_Hz[1]) / 2 + (f0_Hz[1] - f0_Hz[0]) / 2 * cos((boundary_s[1] * SAMPLING_RATE - i) / (boundary_s[1] * SAMPLING_RATE) * M_PI);
    else if (i < (double)boundary_s[3] * SAMPLING_RATE)
      glottisParams[Glottis::FREQUENCY] = (f0_Hz[2] + f0_Hz[1]) / 2 + (f0_Hz[2] - f0_Hz[1]) / 2 * cos((boundary_s[3] * SAMPLING_RATE - i) / ((boundary_s[3] - boundary_s[1]) * SAMPLING_RATE) * M_PI);
    else
      glottisParams[Glottis::FREQUENCY] = (f0_Hz[3] + f0_Hz[2]) / 2 + (f0_Hz[3] - f0_Hz[2]) / 2 * cos((boundary_s[6] * SAMPLING_RATE - i) / ((boundary_s[6] - boundary_s[3]) * SAMPLING_RATE) * M_PI);
    const double *sh[4] = {shapes4x16, shapes4x16 + 16, shapes4x16 + 32, shapes4x16 + 48};
    if (i <= boundary_s[0] * SAMPLING_RATE) { for (int j = 0; j < AF::NUM_AF_PARAMS; j++) currentParams[j].x = sh[0][j]; }
    else if (i <= boundary_s[1] * SAMPLING_RATE) interpolateParameters(sh[0], sh[1], boundary_s[0] * SAMPLING_RATE, boundary_s[1] * SAMPLING_RATE, i);
    else if (i <= boundary_s[2] * SAMPLING_RATE) { for (int j = 0; j < AF::NUM_AF_PARAMS; j++) currentParams[j].x = sh[1][j]; }
    else if (i <= boundary_s[3] * SAMPLING_RATE) interpolateParameters(sh[1], sh[2], boundary_s[2] * SAMPLING_RATE, boundary_s[3] * SAMPLING_RATE, i);
    else if (i <= boundary_s[4] * SAMPLING_RATE) { for (int j = 0; j < AF::NUM_AF_PARAMS; j++) currentParams[j].x = sh[2][j]; }
    else if (i <= boundary_s[5] * SAMPLING_RATE) interpolateParameters(sh[2], sh[3], boundary_s[4] * SAMPLING_RATE, boundary_s[5] * SAMPLING_RATE, i);
    else if (i <= boundary_s[6] * SAMPLING_RATE) { for (int j = 0; j < AF::NUM_AF_PARAMS; j++) currentParams[j].x = sh[3][j]; }
    if (i > (double)(totalTime_s - 0.1) * (double)SAMPLING_RATE)
      glottisParams[Glottis::PRESSURE] = -steadyStateLungPressure / 2 * cos((totalTime_s * SAMPLING_RATE - i) / (totalTime_s - 0.1 * SAMPLING_RATE) * M_PI) + steadyStateLungPressure / 2;

    vtAreaFunction.setParameters(currentParams);
    Tube tubeFunction;
    vtAreaFunction.calculateOneDimTubeFunction(&tubeFunction);
    tubeFunction.teethPosition_cm = currentParams[AF::XIN_CM].x;
    afs_tube_frame &f = frames[i];
    memset(&f, 0, sizeof f);
    for (int k = 0; k < Tube::NUM_PHARYNX_MOUTH_SECTIONS; k++) {
      f.area_cm2[k] = tubeFunction.pharynxMouthSection[k].area_cm2;
      f.length_cm[k] = tubeFunction.pharynxMouthSection[k].length_cm;
      f.laterality[k] = tubeFunction.pharynxMouthSection[k].laterality;
      f.articulator[k] = (int)tubeFunction.pharynxMouthSection[k].articulator;
    }
    f.incisor_pos_cm = tubeFunction.teethPosition_cm;
    f.velum_area_cm2 = 0.0;  // a fresh Tube: nasal port closed
    f.aspiration_dB = tubeFunction.aspirationStrength_dB;
    f.glottis_area_cm2[0] = f.glottis_area_cm2[1] = 0.1; f.glottis_length_cm[0] = f.glottis_length_cm[1] = 0.3;
    for (int c = 0; c < AFS_NUM_GLOTTIS_CONTROL; c++) gl[(size_t)i * AFS_NUM_GLOTTIS_CONTROL + c] = glottisParams[c];
    if (params_out) for (int j = 0; j < 16; j++) params_out[(size_t)i * 16 + j] = currentParams[j].x;
    if (glottis_out) for (int c = 0; c < AFS_NUM_GLOTTIS_CONTROL; c++) glottis_out[(size_t)i * AFS_NUM_GLOTTIS_CONTROL + c] = glottisParams[c];
  }
  std::vector<int32_t> frame_len((size_t)(numSamples > 1 ? numSamples - 1 : 0), 1);
  audio[0] = 0.0;
  if (numSamples > 1)
    tds_one(opts, numSamples, frame_len.data(), frames.data(), gl.data(), seed, audio + 1, nullptr, nullptr, 0);
  return numSamples;
}

// TlModel::getFormants on a fresh TlModel per tube.
void ref_tl_formants(const afs_tl_options *opts, int n_tubes, const afs_tube_frame *tubes,
                     int max_formants, double *freq, double *bw, int32_t *num, int32_t *flags) {
  TlModel *tl = new TlModel();
  for (int b = 0; b < n_tubes; b++) {
    delete tl;
    tl = new TlModel();
    set_tl_options(tl, opts);
    frame_to_tube(&tubes[b], &tl->tube);
    int n = 0;
    bool friction = false, closure = false, nasal = false;
    tl->getFormants(freq + (size_t)b * max_formants, bw + (size_t)b * max_formants, n, max_formants,
                    friction, closure, nasal);
    num[b] = n;
    flags[b] = (friction ? 1 : 0) | (closure ? 2 : 0) | (nasal ? 4 : 0);
  }
  delete tl;
}

// One utterance through the reference TdsModel + TriangularGlottis, driver
// loop restated from Synthesizer.cpp:515-639 (outputFlow[] starts at 0; the
// reference leaves it uninitialised).  Not thread safe when noise is on
// (libc rand()).
static void tds_one(const afs_tds_options *o, int n_frames, const int32_t *frame_len,
                    const afs_tube_frame *frames, const double *glottis, unsigned seed,
                    double *audio, double *flow, double *state_dump, int dump_step, const RefProbes *probes) {
  long total_steps = 0;
  for (int k = 0; k + 1 < n_frames; k++) total_steps += (long)frame_len[k] * TDS_SR_FACTOR;
  TdsModel *tds = new TdsModel();
  TriangularGlottis *gl = new TriangularGlottis();
  tds->options.turbulenceLosses = o->turbulence_losses;
  tds->options.softWalls = o->soft_walls;
  tds->options.generateNoiseSources = o->generate_noise_sources;
  tds->options.radiationFromSkin = o->radiation_from_skin;
  tds->options.piriformFossa = o->piriform_fossa;
  tds->options.innerLengthCorrections = o->inner_length_corrections;
  tds->options.transvelarCoupling = o->transvelar_coupling;
  tds->options.glottisLossOption = o->glottis_loss_option;
  tds->options.flowSeparationAreaRatio = o->flow_separation_area_ratio;
  for (int i = 0; i < AFS_NUM_GLOTTIS_STATIC; i++) gl->staticParam[i].x = o->glottis_static[i];
  tds->resetMotion();
  gl->resetMotion();
  if (o->generate_noise_sources) srand(seed);

  IirFilter outputPressureFilter;
  outputPressureFilter.createChebyshev(7000.0 / (double)TDS_SAMPLING_RATE, false, 8);
  outputPressureFilter.resetBuffers();
  double prevFlow = 0.0;

  Tube tube, prevTube, newTube;
  frame_to_tube(&frames[0], &prevTube);
  const double *prevG = glottis;
  long outPos = 0, flowPos = 0;
  double length_cm[2], area_cm2[2], pressure_dPa[4];

  for (int k = 1; k < n_frames; k++) {
    frame_to_tube(&frames[k], &newTube);
    const double *newG = glottis + (size_t)k * AFS_NUM_GLOTTIS_CONTROL;
    int numNewSamples = frame_len[k - 1] * TDS_SR_FACTOR;
    if (numNewSamples < 1) numNewSamples = 1;
    for (int i = 0; i < numNewSamples; i++) {
      double ratio = (double)i / (double)numNewSamples;
      double ratio1 = 1.0 - ratio;
      tube.interpolate(&prevTube, &newTube, ratio);
      for (int j = 0; j < AFS_NUM_GLOTTIS_CONTROL; j++) gl->controlParam[j].x = ratio1 * prevG[j] + ratio * newG[j];
      gl->calcGeometry();
      gl->getTubeData(length_cm, area_cm2);
      tube.setGlottisGeometry(length_cm, area_cm2);
      tube.setAspirationStrength(gl->getAspirationStrength_dB());
      tds->setTube(&tube, tds->getSampleIndex() != 0);
      tds->setFlowSource(o->flow_source_cm3_s, o->flow_source_section);
      tds->setPressureSource(gl->controlParam[Glottis::PRESSURE].x, Tube::FIRST_TRACHEA_SECTION);
      pressure_dPa[0] = tds->getSectionPressure(Tube::LAST_TRACHEA_SECTION);
      pressure_dPa[1] = tds->getSectionPressure(Tube::LOWER_GLOTTIS_SECTION);
      pressure_dPa[2] = tds->getSectionPressure(Tube::UPPER_GLOTTIS_SECTION);
      pressure_dPa[3] = tds->getSectionPressure(Tube::FIRST_PHARYNX_SECTION);
      gl->incTime(1.0 / (double)TDS_SAMPLING_RATE, pressure_dPa);
      double u = tds->proceedTimeStep();
      int pos = tds->getSampleIndex();
      if (flow) flow[flowPos] = u;
      if (probes) {  // what a caller of the reference reads after every proceedTimeStep (TdsModel.h:279-280)
        for (int q = 0; q < probes->n; q++) {
          double in = 0.0, out = 0.0;
          tds->getSectionFlow(probes->section[q], in, out);
          if (probes->pressure) probes->pressure[q * total_steps + flowPos] = tds->getSectionPressure(probes->section[q]);
          if (probes->inflow) probes->inflow[q * total_steps + flowPos] = in;
          if (probes->outflow) probes->outflow[q * total_steps + flowPos] = out;
        }
      }
      flowPos++;
      if (state_dump && pos == dump_step) {
        // [0..92] pressure, [93..185] pressureRate, [186..278] wallCurrent,
        // [279..375] current magnitude, [376..472] magnitudeRate
        for (int s = 0; s < 93; s++) {
          state_dump[s] = tds->tubeSection[s].pressure;
          state_dump[93 + s] = tds->tubeSection[s].pressureRate;
          state_dump[186 + s] = tds->tubeSection[s].wallCurrent;
        }
        for (int c = 0; c < 97; c++) {
          state_dump[279 + c] = tds->branchCurrent[c].magnitude;
          state_dump[376 + c] = tds->branchCurrent[c].magnitudeRate;
        }
      }
      double outP = (u - prevFlow) / tds->timeStep;
      prevFlow = u;
      double filt = outputPressureFilter.getOutputSample(outP);
      if ((pos % TDS_SR_FACTOR) == 0) {
        double s = filt * 0.004;
        audio[i / TDS_SR_FACTOR + outPos] = s / SHRT_MAX;
      }
    }
    outPos += frame_len[k - 1];
    prevTube = newTube;
    prevG = newG;
  }
  delete tds;
  delete gl;
}

// Batch of utterances; n_threads > 1 only legal with noise off.
void ref_tds_synthesize(const afs_tds_options *opts, int n_tracts, int n_frames, const int32_t *frame_len,
                        const afs_tube_frame *tubes, const double *glottis, const uint32_t *seeds,
                        double *audio, double *flow, int n_threads) {
  long n_out = 0;
  for (int k = 0; k + 1 < n_frames; k++) n_out += frame_len[k];
  if (n_threads < 1 || opts->generate_noise_sources) n_threads = 1;
  auto work = [&](int tid) {
    for (int b = tid; b < n_tracts; b += n_threads) {
      tds_one(opts, n_frames, frame_len, tubes + (size_t)b * n_frames,
              glottis + (size_t)b * n_frames * AFS_NUM_GLOTTIS_CONTROL, seeds ? seeds[b] : 1u,
              audio + (size_t)b * n_out, flow ? flow + (size_t)b * n_out * TDS_SR_FACTOR : nullptr, nullptr, 0);
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < n_threads; t++) th.emplace_back(work, t);
  work(0);
  for (auto &t : th) t.join();
}

// Single utterance with taps: pressure / inflow / outflow [n_probes][steps] of the given tube sections.
void ref_tds_probed(const afs_tds_options *opts, int n_frames, const int32_t *frame_len, const afs_tube_frame *tubes,
                    const double *glottis, uint32_t seed, double *audio, int n_probes, const int32_t *sections,
                    double *pressure, double *inflow, double *outflow) {
  RefProbes pr{n_probes, sections, pressure, inflow, outflow};
  tds_one(opts, n_frames, frame_len, tubes, glottis, seed, audio, nullptr, nullptr, 0, &pr);
}

// TlModel::getMostConstrictedSection / getMeanFlow (TlModel.cpp:204-246) for a batch of frames.
void ref_tl_mean_flow(int n, const afs_tube_frame *frames, double lung_pressure_dPa, double *flow, int32_t *section) {
  TlModel *tl = new TlModel();
  for (int i = 0; i < n; i++) {
    frame_to_tube(&frames[i], &tl->tube);
    if (section) section[i] = tl->getMostConstrictedSection();
    if (flow) flow[i] = tl->getMeanFlow(lung_pressure_dPa);
  }
  delete tl;
}

// Single utterance with a state dump after simulation step `dump_step` (1-based).
void ref_tds_state_after(const afs_tds_options *opts, int n_frames, const int32_t *frame_len,
                         const afs_tube_frame *tubes, const double *glottis, int dump_step,
                         double *audio, double *flow, double *state473) {
  tds_one(opts, n_frames, frame_len, tubes, glottis, 1u, audio, flow, state473, dump_step);
}

// IirFilter::createChebyshev (low-pass) coefficients.
void ref_chebyshev_lowpass(double cutoff_ratio, int num_poles, double *a, double *b) {
  IirFilter f;
  f.createChebyshev(cutoff_ratio, false, num_poles);
  for (int i = 0; i <= num_poles; i++) { a[i] = f.a[i]; b[i] = f.b[i]; }
}

}  // extern "C"
