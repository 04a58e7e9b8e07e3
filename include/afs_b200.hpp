// afs_b200.hpp -- header-only C++17 facade over the C ABI (afs_b200.h) that keeps
// the reference's class vocabulary for the hot path, so that callers written
// against the VocalTractLab backend classes of Simon-Stone/AreaFunctionSynthesis
// (`Tube`, `TlModel`, `Synthesizer::synthesizeSignalTds`) can switch with
// mechanical edits.  Batched: every object drives a whole batch of tubes /
// utterances on one B200.  Errors surface as afs::Error (the reference prints
// and carries on, TdsModel.cpp:259,1832,2267).
//
// Reference interfaces mirrored (paths under src/Backend/):
//   afs::Tube            Tube.h:14-170        setPharynxMouthGeometry, setGlottisGeometry, setGlottisArea,
//                                             setVelumOpening, setAspirationStrength, interpolate
//   afs::TlModel         TlModel.h:24-131     public `options`, `tube`; getSpectrum, getFormants
//   afs::TdsSynthesizer  Synthesizer.h:161    synthesizeSignalTds(newTube, newGlottisParams, numNewSamples, out)
//                        + TdsModel.h:257-280 options, resetMotion, setFlowSource, getSectionPressure / getSectionFlow
//   afs::DeviceGroup     (no counterpart: the reference is single-threaded) several B200s of one box, one process
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "afs_b200.h"

namespace afs {

struct Error : std::runtime_error {
  int status;
  Error(int s, const std::string &what) : std::runtime_error(what), status(s) {}
};
inline void check(int status, const afs_ctx *ctx = nullptr) {
  if (status == AFS_OK) return;
  std::string msg = afs_status_string(status);
  if (ctx && afs_last_cuda_error(ctx)[0]) msg += std::string(" [") + afs_last_cuda_error(ctx) + "]";
  throw Error(status, msg);
}

// One CUDA device.  Replaces `new TdsModel()` / `new TlModel()` (Data.cpp:78-88, Synthesizer.cpp:48).
class Device {
 public:
  explicit Device(int ordinal = 0) { check(afs_create(&ctx_, ordinal)); }
  ~Device() { afs_destroy(ctx_); }
  Device(const Device &) = delete;
  Device &operator=(const Device &) = delete;
  afs_ctx *ctx() const { return ctx_; }

 private:
  afs_ctx *ctx_ = nullptr;
};

// Several devices of one box (afs_group): a batch is split into contiguous shards, one host thread per device.
class DeviceGroup {
 public:
  DeviceGroup() { check(afs_group_create(&g_, nullptr, 0)); }  // every visible device
  explicit DeviceGroup(const std::vector<int> &ordinals) { check(afs_group_create(&g_, ordinals.data(), (int)ordinals.size())); }
  ~DeviceGroup() { afs_group_destroy(g_); }
  DeviceGroup(const DeviceGroup &) = delete;
  DeviceGroup &operator=(const DeviceGroup &) = delete;
  afs_group *group() const { return g_; }
  int size() const { return afs_group_size(g_); }

 private:
  afs_group *g_ = nullptr;
};

// The dynamic part of the reference Tube as a value type (a reference Tube holds pointers into itself and cannot
// be copied).  Same setter names and argument meaning as Tube.h.
class Tube {
 public:
  enum Articulator { VOCAL_FOLDS, TONGUE, LOWER_INCISORS, LOWER_LIP, OTHER_ARTICULATOR, NUM_ARTICULATORS };
  static constexpr int NUM_PHARYNX_MOUTH_SECTIONS = AFS_NUM_PM_SECTIONS;
  static constexpr int NUM_GLOTTIS_SECTIONS = AFS_NUM_GLOTTIS_SECTIONS;
  static constexpr int FIRST_PHARYNX_SECTION = AFS_FIRST_PHARYNX_SECTION;
  static constexpr double MIN_AREA_CM2 = 0.1e-2;

  afs_tube_frame frame;

  Tube() { afs_tube_frame_default(&frame); }
  void setPharynxMouthGeometry(const double *length_cm, const double *area_cm2, const Articulator *articulator,
                               const double *laterality, double incisorPos_cm) {
    for (int i = 0; i < NUM_PHARYNX_MOUTH_SECTIONS; i++) {
      frame.length_cm[i] = length_cm[i];
      frame.area_cm2[i] = area_cm2[i] < MIN_AREA_CM2 ? MIN_AREA_CM2 : area_cm2[i];  // Tube.cpp:337
      frame.articulator[i] = (int32_t)articulator[i];
      frame.laterality[i] = laterality[i];
    }
    frame.incisor_pos_cm = incisorPos_cm;
  }
  void setGlottisGeometry(const double *length_cm, const double *area_cm2) {
    for (int i = 0; i < 2; i++) { frame.glottis_length_cm[i] = length_cm[i]; frame.glottis_area_cm2[i] = area_cm2[i]; }
  }
  void setGlottisArea(double area_cm2) {  // Tube.cpp:388-394
    const double l[2] = {0.3, 0.3}, a[2] = {area_cm2, area_cm2};
    setGlottisGeometry(l, a);
  }
  void setVelumOpening(double area_cm2) { frame.velum_area_cm2 = area_cm2; }
  void setAspirationStrength(double dB) { frame.aspiration_dB = dB; }
  double getVelumOpening_cm2() const { return frame.velum_area_cm2 < MIN_AREA_CM2 ? MIN_AREA_CM2 : frame.velum_area_cm2; }
  // Tube::interpolate (Tube.cpp:438-505) for the fields a frame carries.
  void interpolate(const Tube *left, const Tube *right, double ratio) {
    const double r1 = 1.0 - ratio;
    for (int i = 0; i < NUM_PHARYNX_MOUTH_SECTIONS; i++) {
      const double a = r1 * left->frame.area_cm2[i] + ratio * right->frame.area_cm2[i];
      frame.area_cm2[i] = a < MIN_AREA_CM2 ? MIN_AREA_CM2 : a;
      frame.length_cm[i] = r1 * left->frame.length_cm[i] + ratio * right->frame.length_cm[i];
      frame.articulator[i] = left->frame.articulator[i];
      frame.laterality[i] = r1 * left->frame.laterality[i] + ratio * right->frame.laterality[i];
    }
    frame.incisor_pos_cm = r1 * left->frame.incisor_pos_cm + ratio * right->frame.incisor_pos_cm;
    frame.velum_area_cm2 = r1 * left->getVelumOpening_cm2() + ratio * right->getVelumOpening_cm2();
    frame.aspiration_dB = r1 * left->frame.aspiration_dB + ratio * right->frame.aspiration_dB;
  }
};

// OneDimAreaFunction::setParameters + calculateOneDimTubeFunction (OneDimAreaFunction.cpp:75-138, 356-368).
inline void calculateOneDimTubeFunction(const double params[16], Tube *tube) {
  check(afs_area_function_to_frame(params, &tube->frame));
}

struct ComplexSpectrum {  // the non-negative-frequency half of the reference's ComplexSignal (Signal.h:61-92)
  int numFreq = 0;
  std::vector<double> re, im;  // [n_tubes][numFreq]
};

// Batched TlModel.  `options` and `tubes` are public like TlModel::options / TlModel::tube.
class TlModel {
 public:
  enum SpectrumType { INPUT_IMPEDANCE, OUTPUT_IMPEDANCE, FLOW_SOURCE_TF, PRESSURE_SOURCE_TF, RADIATION };
  afs_tl_options options;
  std::vector<Tube> tubes;

  explicit TlModel(Device &dev) : dev_(dev) { afs_tl_options_default(&options); }
  void setLungPressure(double lungPressure_dPa) { options.lung_pressure_dPa = lungPressure_dPa; }
  // TlModel::getMostConstrictedSection (TlModel.cpp:204-217) for every tube.
  std::vector<int32_t> getMostConstrictedSection() {
    std::vector<afs_tube_frame> f(tubes.size());
    for (size_t i = 0; i < tubes.size(); i++) f[i] = tubes[i].frame;
    std::vector<int32_t> sec(f.size(), 0);
    check(afs_tl_mean_flow(dev_.ctx(), (int)f.size(), f.data(), 0.0, nullptr, sec.data()), dev_.ctx());
    return sec;
  }
  // TlModel::getMeanFlow(lungPressure_dPa) (TlModel.cpp:226-246) for every tube; like the reference it also sets the
  // lung pressure of the model.
  std::vector<double> getMeanFlow(double lungPressure_dPa) {
    std::vector<afs_tube_frame> f(tubes.size());
    for (size_t i = 0; i < tubes.size(); i++) f[i] = tubes[i].frame;
    std::vector<double> flow(f.size(), 0.0);
    check(afs_tl_mean_flow(dev_.ctx(), (int)f.size(), f.data(), lungPressure_dPa, flow.data(), nullptr), dev_.ctx());
    setLungPressure(lungPressure_dPa);
    return flow;
  }

  // TlModel::getSpectrum (TlModel.cpp:123-197) for every tube; the bin grid is explicit: bin i at i*f0_hz.
  void getSpectrum(SpectrumType type, ComplexSpectrum *spectrum, double f0_hz, int numFreq, int section) {
    std::vector<afs_tube_frame> f(tubes.size());
    for (size_t i = 0; i < tubes.size(); i++) f[i] = tubes[i].frame;
    spectrum->numFreq = numFreq;
    spectrum->re.assign(tubes.size() * (size_t)numFreq, 0.0);
    spectrum->im.assign(tubes.size() * (size_t)numFreq, 0.0);
    check(afs_tl_sweep(dev_.ctx(), &options, (int)f.size(), f.data(), (int)type, section, f0_hz, numFreq,
                       spectrum->re.data(), spectrum->im.data()), dev_.ctx());
  }
  // Reference grid: spectrumLength points at `samplingRate` (reference: 22050), numFreq as in TlModel.cpp:125-148
  // without the MAX_NUM_FREQ cap.
  void getSpectrum(SpectrumType type, ComplexSpectrum *spectrum, int spectrumLength, int section,
                   double samplingRate = 22050.0) {
    const double f0 = samplingRate / (double)spectrumLength;
    int n = (int)(10100.0 / f0);
    if (n >= spectrumLength / 2) n = spectrumLength / 2 - 1;
    getSpectrum(type, spectrum, f0, n, section);
  }
  // TlModel::getFormants (TlModel.cpp:280-461) for every tube.
  void getFormants(std::vector<double> *formantFreq, std::vector<double> *formantBW, std::vector<int32_t> *numFormants,
                   int MAX_FORMANTS, std::vector<int32_t> *flags /* bit0 friction, bit1 closure, bit2 nasal */) {
    std::vector<afs_tube_frame> f(tubes.size());
    for (size_t i = 0; i < tubes.size(); i++) f[i] = tubes[i].frame;
    formantFreq->assign(f.size() * (size_t)MAX_FORMANTS, 0.0);
    formantBW->assign(f.size() * (size_t)MAX_FORMANTS, 0.0);
    numFormants->assign(f.size(), 0);
    flags->assign(f.size(), 0);
    check(afs_tl_formants(dev_.ctx(), &options, (int)f.size(), f.data(), 22050.0, 2048, MAX_FORMANTS,
                          formantFreq->data(), formantBW->data(), numFormants->data(), flags->data()), dev_.ctx());
  }
  // TlModel::getImpulseResponse (TlModel.cpp:90-111) for every tube: impulseResponse[n][1 << lengthExponent].
  void getImpulseResponse(std::vector<double> *impulseResponse, int lengthExponent, double samplingRate = 22050.0) {
    std::vector<afs_tube_frame> f(tubes.size());
    for (size_t i = 0; i < tubes.size(); i++) f[i] = tubes[i].frame;
    impulseResponse->assign(f.size() << lengthExponent, 0.0);
    check(afs_tl_impulse_response(dev_.ctx(), &options, (int)f.size(), f.data(), samplingRate, lengthExponent,
                                  impulseResponse->data()), dev_.ctx());
  }

  Device &device() { return dev_; }

 private:
  Device &dev_;
};

// Data::optimizeFormantsVowel (Frontend/Data.cpp:203-487) for a batch of area-function parameter sets
// (params[n][16] in/out, targets[n][3]); returns the final errors in percent.
inline std::vector<double> optimizeFormantsVowel(TlModel &tlModel, std::vector<double> *params, const std::vector<double> &targets,
                                                 double maxParamChange_cm, double minAdvisedArea_cm2,
                                                 const bool paramFixed[16] = nullptr, int maxRuns = 100) {
  const int n = (int)(params->size() / 16);
  int32_t fixed[16];
  for (int i = 0; i < 16; i++) fixed[i] = paramFixed ? (paramFixed[i] ? 1 : 0) : 0;
  std::vector<double> finalError((size_t)n, 0.0);
  check(afs_optimize_formants_vowel(tlModel.device().ctx(), &tlModel.options, n, params->data(), targets.data(), maxParamChange_cm,
                                    minAdvisedArea_cm2, fixed, maxRuns, 22050.0, 2048, nullptr, finalError.data(), nullptr,
                                    nullptr, nullptr), tlModel.device().ctx());
  return finalError;
}

// Synthesizer::synthesizeVowelLf (Synthesizer.cpp:645-801) for a batch of vowels: the tubes of `tlModel`, one LF pulse
// parameter set per tube; returns int16 audio [n][afs_lf_vowel_length(isLongVowel, samplingRate)].
inline std::vector<int16_t> synthesizeVowelsLf(TlModel &tlModel, const std::vector<afs_lf_pulse> &lfPulses, bool isLongVowel,
                                               double samplingRate = 22050.0) {
  std::vector<afs_tube_frame> f(tlModel.tubes.size());
  for (size_t i = 0; i < f.size(); i++) f[i] = tlModel.tubes[i].frame;
  std::vector<int16_t> audio(f.size() * (size_t)afs_lf_vowel_length(isLongVowel ? 1 : 0, samplingRate), 0);
  check(afs_synthesize_vowels_lf(tlModel.device().ctx(), &tlModel.options, (int)f.size(), f.data(), lfPulses.data(),
                                 isLongVowel ? 1 : 0, samplingRate, audio.data(), nullptr), tlModel.device().ctx());
  return audio;
}

// Synthesizer::playTargetSequence (Synthesizer.cpp:1299-1422) for a batch of sequences sharing the timing:
// targetShapes[n][4][16]; returns audio [n][afs_target_sequence_length(...)] in [-1, 1].
inline std::vector<double> playTargetSequences(Device &dev, const afs_tds_options &options, const std::vector<double> &targetShapes,
                                               const double stationary_s[4], const double transition_s[3],
                                               const double glottisControl[AFS_NUM_GLOTTIS_CONTROL], double samplingRate = 22050.0,
                                               int tdsSrFactor = 2 /* 1 = the reference as shipped, see TdsSynthesizer */) {
  const int n = (int)(targetShapes.size() / 64);
  std::vector<double> audio((size_t)n * (size_t)afs_target_sequence_length(stationary_s, transition_s, samplingRate), 0.0);
  check(afs_play_target_sequences(dev.ctx(), &options, n, targetShapes.data(), stationary_s, transition_s, glottisControl,
                                  samplingRate, tdsSrFactor, nullptr, audio.data(), nullptr), dev.ctx());
  return audio;
}

// Batched replacement of the Synthesizer's TDS driver: the caller queues control frames exactly like successive
// calls of Synthesizer::synthesizeSignalTds(newTube, newGlottisParams, numNewSamples, out) (Synthesizer.cpp:515-639)
// -- the first call primes -- and flush() runs the whole batch of utterances in one kernel launch; flushAndContinue()
// does the same and keeps the models alive for the frames that follow (utterances of any length, piece by piece).
class TdsSynthesizer {
 public:
  afs_tds_options options;  // TdsModel::options + TriangularGlottis static parameters
  double samplingRate = 22050.0;
  // Simulation rate = tdsSrFactor * samplingRate.  The reference ships TDS_SR_FACTOR = 1 (Constants.h:25: TDS at
  // 22.05 kHz); 2 -- a simulation at 44.1 kHz, the configuration BASELINE.json is quoted on -- is the default HERE.
  // Set it to 1 for the audio of an unmodified reference build.
  int tdsSrFactor = 2;

  TdsSynthesizer(Device &dev, int numUtterances) : dev_(&dev), n_(numUtterances) { afs_tds_options_default(&options); }
  // the same over several devices (contiguous shards of the utterances)
  TdsSynthesizer(DeviceGroup &group, int numUtterances) : group_(&group), n_(numUtterances) { afs_tds_options_default(&options); }
  ~TdsSynthesizer() {
    if (stream_) afs_tds_stream_destroy(stream_);
    if (gstream_) afs_group_tds_stream_destroy(gstream_);
  }
  TdsSynthesizer(const TdsSynthesizer &) = delete;
  TdsSynthesizer &operator=(const TdsSynthesizer &) = delete;
  // TdsModel::resetMotion + a driver that has not seen a tube yet (Synthesizer::reset)
  void resetMotion() {
    clearQueue();
    continuing_ = false;
    if (stream_) afs_tds_stream_reset(stream_);
    if (gstream_) afs_group_tds_stream_reset(gstream_);
  }
  // TdsModel::setFlowSource(flow_cm3_s, section) (TdsModel.cpp:629-633); section -1 switches it off.
  void setFlowSource(double flow_cm3_s, int section) { options.flow_source_cm3_s = flow_cm3_s; options.flow_source_section = section; }
  // Taps for the next flush(): TdsModel::getSectionPressure / getSectionFlow (TdsModel.cpp:1715-1734) of these tube
  // sections after every simulation step; read them with sectionPressure() / sectionInflow() / sectionOutflow().
  void setProbeSections(const std::vector<int> &sections) { probeSections_ = sections; }
  const std::vector<double> &sectionPressure() const { return probeP_; }  // [n][sections][steps]
  const std::vector<double> &sectionInflow() const { return probeIn_; }
  const std::vector<double> &sectionOutflow() const { return probeOut_; }

  // One control frame for ALL utterances: newTubes[n], newGlottisParams[n][6].  The first call primes (returns
  // without producing samples, Synthesizer.cpp:522-532); later calls append numNewSamples output samples.
  void synthesizeSignalTds(const Tube *newTubes, const double *newGlottisParams, int numNewSamples) {
    for (int b = 0; b < n_; b++) tubes_.push_back(newTubes[b].frame);
    glottis_.insert(glottis_.end(), newGlottisParams, newGlottisParams + (size_t)n_ * AFS_NUM_GLOTTIS_CONTROL);
    if (frames_ > 0 || continuing_) frameLen_.push_back(numNewSamples < 1 ? 1 : numNewSamples);
    frames_++;
  }
  // Runs the queued frames and KEEPS the models: the next synthesizeSignalTds calls continue from the state and the
  // tube these frames ended in, exactly like further calls on the reference's Synthesizer object (afs_tds_stream_*,
  // afs_group_tds_stream_*; `options`, the rates and the number of utterances are fixed by the first call).  The audio of an
  // utterance synthesised in pieces equals the audio of one flush() of all its frames.
  void flushAndContinue(std::vector<double> *audio, std::vector<double> *flow = nullptr, std::vector<uint32_t> *status = nullptr) {
    if (frames_ < 1) throw Error(AFS_ERR_BAD_ARGUMENT, "no control frames queued");
    if (!probeSections_.empty()) throw Error(AFS_ERR_UNSUPPORTED, "taps are served by flush() only");
    if (dev_ && !stream_) check(afs_tds_stream_create(dev_->ctx(), &options, n_, samplingRate * tdsSrFactor, tdsSrFactor, nullptr, &stream_), dev_->ctx());
    if (group_ && !gstream_) check(afs_group_tds_stream_create(group_->group(), &options, n_, samplingRate * tdsSrFactor, tdsSrFactor, nullptr, &gstream_));
    std::vector<afs_tube_frame> t;
    std::vector<double> g;
    transposeQueue(&t, &g);
    std::vector<int32_t> len;
    if (!continuing_) len.push_back(0);  // the priming frame
    len.insert(len.end(), frameLen_.begin(), frameLen_.end());
    size_t n_out = 0;
    for (int v : frameLen_) n_out += (size_t)v;
    audio->assign((size_t)n_ * n_out, 0.0);
    if (flow) flow->assign((size_t)n_ * n_out * tdsSrFactor, 0.0);
    if (status) status->assign((size_t)n_, 0u);
    if (group_) check(afs_group_tds_stream_synthesize(gstream_, frames_, len.data(), t.data(), g.data(), audio->data(),
                                                      flow ? flow->data() : nullptr, status ? status->data() : nullptr));
    else check(afs_tds_stream_synthesize(stream_, frames_, len.data(), t.data(), g.data(), audio->data(), flow ? flow->data() : nullptr,
                                         status ? status->data() : nullptr), dev_->ctx());
    clearQueue();
    continuing_ = true;
  }
  // Checkpoint / resume of the models kept by flushAndContinue (afs_tds_stream_save / _load): restore() accepts the
  // blob of a synthesizer with the same number of utterances, rates, options and devices.
  std::vector<unsigned char> checkpoint() {
    if (!stream_ && !gstream_) throw Error(AFS_ERR_BAD_ARGUMENT, "no live models: call flushAndContinue first");
    std::vector<unsigned char> blob(stream_ ? afs_tds_stream_checkpoint_bytes(stream_) : afs_group_tds_stream_checkpoint_bytes(gstream_));
    if (stream_) check(afs_tds_stream_save(stream_, blob.data(), blob.size()), dev_->ctx());
    else check(afs_group_tds_stream_save(gstream_, blob.data(), blob.size()));
    return blob;
  }
  void restore(const std::vector<unsigned char> &blob) {
    clearQueue();
    if (dev_ && !stream_) check(afs_tds_stream_create(dev_->ctx(), &options, n_, samplingRate * tdsSrFactor, tdsSrFactor, nullptr, &stream_), dev_->ctx());
    if (group_ && !gstream_) check(afs_group_tds_stream_create(group_->group(), &options, n_, samplingRate * tdsSrFactor, tdsSrFactor, nullptr, &gstream_));
    if (stream_) check(afs_tds_stream_load(stream_, blob.data(), blob.size()), dev_->ctx());
    else check(afs_group_tds_stream_load(gstream_, blob.data(), blob.size()));
    continuing_ = stream_ ? afs_tds_stream_primed(stream_) > 0 : afs_group_tds_stream_samples(gstream_) > 0;
  }
  // Runs the queued frames; audio[n][sum(numNewSamples)] in [-1, 1]; flow (optional) per simulation step.
  void flush(std::vector<double> *audio, std::vector<double> *flow = nullptr, std::vector<uint32_t> *status = nullptr) {
    if (frames_ < 2) throw Error(AFS_ERR_BAD_ARGUMENT, "need at least two control frames");
    size_t n_out = 0;
    for (int v : frameLen_) n_out += (size_t)v;
    if (continuing_) throw Error(AFS_ERR_BAD_ARGUMENT, "models kept by flushAndContinue: go on with it, or resetMotion() first");
    std::vector<afs_tube_frame> t;
    std::vector<double> g;
    transposeQueue(&t, &g);
    audio->assign((size_t)n_ * n_out, 0.0);
    if (flow) flow->assign((size_t)n_ * n_out * tdsSrFactor, 0.0);
    if (status) status->assign((size_t)n_, 0u);
    if (group_) {
      if (!probeSections_.empty()) throw Error(AFS_ERR_UNSUPPORTED, "taps are a single-device feature");
      check(afs_group_tds_synthesize(group_->group(), &options, n_, frames_, frameLen_.data(), t.data(), g.data(),
                                     samplingRate * tdsSrFactor, tdsSrFactor, nullptr, audio->data(),
                                     flow ? flow->data() : nullptr, status ? status->data() : nullptr));
    } else {
      afs_tds_probes pr;
      pr.n_probes = (int32_t)probeSections_.size();
      if (pr.n_probes > AFS_MAX_TDS_PROBES) throw Error(AFS_ERR_BAD_ARGUMENT, "too many probe sections");
      const size_t pn = (size_t)n_ * probeSections_.size() * n_out * (size_t)tdsSrFactor;
      probeP_.assign(pn, 0.0); probeIn_.assign(pn, 0.0); probeOut_.assign(pn, 0.0);
      for (int q = 0; q < pr.n_probes; q++) pr.section[q] = probeSections_[(size_t)q];
      pr.pressure_dPa = probeP_.data(); pr.inflow_cm3_s = probeIn_.data(); pr.outflow_cm3_s = probeOut_.data();
      check(afs_tds_synthesize_probed(dev_->ctx(), &options, n_, frames_, frameLen_.data(), t.data(), g.data(),
                                      samplingRate * tdsSrFactor, tdsSrFactor, nullptr, audio->data(),
                                      flow ? flow->data() : nullptr, status ? status->data() : nullptr,
                                      pr.n_probes ? &pr : nullptr), dev_->ctx());
    }
    clearQueue();
  }

 private:
  void clearQueue() { tubes_.clear(); glottis_.clear(); frameLen_.clear(); frames_ = 0; }
  // queue order is frame-major; the ABI wants [utterance][frame]
  void transposeQueue(std::vector<afs_tube_frame> *t, std::vector<double> *g) const {
    t->resize((size_t)n_ * frames_);
    g->resize((size_t)n_ * frames_ * AFS_NUM_GLOTTIS_CONTROL);
    for (int k = 0; k < frames_; k++)
      for (int b = 0; b < n_; b++) {
        (*t)[(size_t)b * frames_ + k] = tubes_[(size_t)k * n_ + b];
        for (int c = 0; c < AFS_NUM_GLOTTIS_CONTROL; c++)
          (*g)[((size_t)b * frames_ + k) * AFS_NUM_GLOTTIS_CONTROL + c] = glottis_[((size_t)k * n_ + b) * AFS_NUM_GLOTTIS_CONTROL + c];
      }
  }
  afs_tds_stream *stream_ = nullptr;
  afs_group_tds_stream *gstream_ = nullptr;
  bool continuing_ = false;  // the models carry state and a previous tube (flushAndContinue)
  Device *dev_ = nullptr;
  DeviceGroup *group_ = nullptr;
  std::vector<int> probeSections_;
  std::vector<double> probeP_, probeIn_, probeOut_;
  int n_;
  int frames_ = 0;
  std::vector<afs_tube_frame> tubes_;
  std::vector<double> glottis_;
  std::vector<int32_t> frameLen_;
};

}  // namespace afs
