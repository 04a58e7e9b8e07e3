/*
 * afs_b200.h -- C ABI of the B200-native area-function-to-speech hot path.
 *
 * This is the drop-in boundary for the acoustic back end of
 * Simon-Stone/AreaFunctionSynthesis (VocalTractLab classes).  The reference has
 * no FFI of its own; the boundary is the public C++
 * class API used by its two callers, restated here as batched, POD-only entry
 * points.  Every entry point cites the reference interface it replaces
 * (paths relative to the reference repository root, src/Backend/...).
 *
 * Conventions
 *   - plain pointers and sizes only, no C++/torch types;
 *   - every function returns an afs_status (0 = success); nothing prints,
 *     nothing throws;
 *   - all buffers are caller-owned; "_device" variants take device pointers
 *     and a cudaStream_t passed as void* and are asynchronous on that stream;
 *   - there is NO CPU fallback: without a CUDA device afs_create fails with
 *     AFS_ERR_NO_DEVICE and every compute entry point fails with
 *     AFS_ERR_BAD_CONTEXT.
 */
#ifndef AFS_B200_H
#define AFS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AFS_VERSION 200 /* 0.2.0 */

/* ---- tube topology (Tube.h:53-79) ---------------------------------------- */
#define AFS_NUM_TRACHEA_SECTIONS 23
#define AFS_NUM_GLOTTIS_SECTIONS 2
#define AFS_NUM_PHARYNX_SECTIONS 16
#define AFS_NUM_MOUTH_SECTIONS 24
#define AFS_NUM_PM_SECTIONS 40 /* pharynx + mouth: the dynamic part */
#define AFS_NUM_NOSE_SECTIONS 19
#define AFS_NUM_FOSSA_SECTIONS 5
#define AFS_NUM_SINUS_SECTIONS 4
#define AFS_NUM_SECTIONS 93
#define AFS_NUM_CURRENTS 97 /* TdsModel.h:37 */

#define AFS_FIRST_TRACHEA_SECTION 0
#define AFS_LOWER_GLOTTIS_SECTION 23
#define AFS_UPPER_GLOTTIS_SECTION 24
#define AFS_FIRST_PHARYNX_SECTION 25
#define AFS_FIRST_MOUTH_SECTION 41
#define AFS_FIRST_NOSE_SECTION 65
#define AFS_FIRST_FOSSA_SECTION 84
#define AFS_FIRST_SINUS_SECTION 89

/* Tube::Articulator (Tube.h:23-31) */
enum afs_articulator {
  AFS_VOCAL_FOLDS = 0,
  AFS_TONGUE = 1,
  AFS_LOWER_INCISORS = 2,
  AFS_LOWER_LIP = 3,
  AFS_OTHER_ARTICULATOR = 4
};

/* TlModel::SpectrumType (TlModel.h:40-43) */
enum afs_spectrum_type {
  AFS_INPUT_IMPEDANCE = 0,
  AFS_OUTPUT_IMPEDANCE = 1,
  AFS_FLOW_SOURCE_TF = 2,
  AFS_PRESSURE_SOURCE_TF = 3,
  AFS_RADIATION = 4
};

/* TlModel::RadiationType (TlModel.h:31-38) */
enum afs_radiation_type {
  AFS_NO_RADIATION = 0,
  AFS_PISTONINSPHERE_RADIATION = 1,
  AFS_PISTONINWALL_RADIATION = 2,
  AFS_PARALLEL_RADIATION = 3
};

/* How the TL sweep evaluates the chain products of TlModel::prepareCalculations (TlModel.cpp:484-676).
 *   REFERENCE: the reference's operations in the reference's order, no FMA contraction: spectra are
 *              BIT-IDENTICAL to TlModel::getSpectrum.
 *   FAST:      the same products as ladder recurrences (series/shunt updates, shared wall-admittance
 *              reciprocals, FMA): ~6x fewer FP64 instructions.  Accuracy against the reference's own arithmetic,
 *              complex relative error per bin (measured on 4096 x 2048 points, tests/test_tl_gpu.py): <= 1e-9 on
 *              99.9 % of the bins, <= 1e-8 (measured maximum 2.9e-9) on ALL bins; magnitude and phase separately
 *              stay below the same bounds.  The bins above 1e-9 are those where the reference itself loses digits
 *              (the cancellation in inverse(P[section]) * P[40], TlModel.cpp:1336): evaluated with a 64-bit
 *              mantissa the formulas agree with FAST to 1e-12 there (tools/tl_truth.py).  I.e. the north-star bar
 *              of 1e-9 holds against the exact value of the reference's formulas, not against their double-precision
 *              rounding.  Wave-propagation mode (lumped_elements = 0) and the RADIATION spectrum are always
 *              evaluated in REFERENCE arithmetic. */
enum afs_tl_arithmetic { AFS_TL_ARITHMETIC_REFERENCE = 0, AFS_TL_ARITHMETIC_FAST = 1 };

/* TdsModel::GlottisLossOptions / SolverType (TdsModel.h:60-73) */
enum afs_glottis_loss { AFS_STANDARD_ENTRANCE_LOSS = 0, AFS_ENTRANCE_LOSS_VAN_DEN_BERG = 1, AFS_VARIABLE_ENTRANCE_LOSS = 2 };

typedef enum afs_status {
  AFS_OK = 0,
  AFS_ERR_NO_DEVICE = 1,      /* no CUDA device / driver: there is no CPU fallback */
  AFS_ERR_BAD_CONTEXT = 2,
  AFS_ERR_BAD_ARGUMENT = 3,
  AFS_ERR_UNSUPPORTED = 4,    /* option of the reference not implemented on the device path */
  AFS_ERR_CUDA = 5,           /* a CUDA call failed; see afs_last_cuda_error */
  AFS_ERR_OUT_OF_MEMORY = 6
} afs_status;

/*
 * One control frame of the area function: the part of a reference `Tube`
 * that changes during speech (Tube.h:33-51, 100-118; written by
 * Tube::setPharynxMouthGeometry Tube.cpp:323, setVelumOpening :402,
 * setGlottisGeometry :358, setAspirationStrength :425).  The static cavities
 * (trachea, nose, sinuses, piriform fossa) are those of Tube::Tube()
 * (Tube.cpp:46-260) with the default lengths (Tube.cpp:29-32).
 *
 * A `Tube` cannot cross a C ABI (it holds pointers into itself, SURVEY.md
 * App. D), hence this POD.  1184 bytes.
 */
typedef struct afs_tube_frame {
  double area_cm2[AFS_NUM_PM_SECTIONS];
  double length_cm[AFS_NUM_PM_SECTIONS];
  double laterality[AFS_NUM_PM_SECTIONS];
  double incisor_pos_cm;       /* Tube::teethPosition_cm */
  double velum_area_cm2;       /* argument of Tube::setVelumOpening */
  double aspiration_dB;        /* Tube::aspirationStrength_dB (TL ignores it) */
  double glottis_area_cm2[2];  /* TL only: TDS takes these from the glottis model */
  double glottis_length_cm[2]; /* TL only */
  double reserved;             /* keeps the int block 8-byte aligned; must be 0 */
  int32_t articulator[AFS_NUM_PM_SECTIONS]; /* enum afs_articulator */
} afs_tube_frame;

/* TriangularGlottis control parameters (TriangularGlottis.h:27-38, defaults
 * TriangularGlottis.cpp:17-26): f0 Hz, pressure dPa, lower/upper rest
 * displacement cm, arytenoid area cm^2, aspiration strength dB. */
#define AFS_NUM_GLOTTIS_CONTROL 6
/* TriangularGlottis static parameters (TriangularGlottis.h:40-61, defaults
 * TriangularGlottis.cpp:38-56). */
#define AFS_NUM_GLOTTIS_STATIC 16

/* TlModel::Options (TlModel.h:47-59); defaults of TlModel::TlModel()
 * (TlModel.cpp:30-39) via afs_tl_options_default. */
typedef struct afs_tl_options {
  int32_t radiation; /* enum afs_radiation_type */
  int32_t boundary_layer;
  int32_t heat_conduction;
  int32_t soft_walls;
  int32_t hagen_resistance;
  int32_t inner_length_corrections;
  int32_t lumped_elements; /* 0: wave propagation with cosh/sinh (TlModel.cpp:896-929), 1e-9 parity instead of bitwise */
  int32_t paranasal_sinuses;
  int32_t piriform_fossa;
  int32_t static_pressure_drops;
  /* TlModel keeps the lung pressure as hidden state (TlModel.cpp:253-260,
   * SURVEY.md App. D); here it is explicit.  Default 0. */
  double lung_pressure_dPa;
  /* enum afs_tl_arithmetic.  Default AFS_TL_ARITHMETIC_REFERENCE. */
  int32_t arithmetic;
  int32_t reserved; /* must be 0 */
} afs_tl_options;

/* TdsModel::Options (TdsModel.h:75-87); defaults of TdsModel::TdsModel()
 * (TdsModel.cpp:35-44) via afs_tds_options_default.  With
 * generate_noise_sources every tract draws from its own glibc-compatible
 * random() stream (the reference shares libc rand(), TdsModel.cpp:1690-1696),
 * seeded by `seeds[tract]` (1 when no seeds are given, like an unseeded
 * process).  A noise source whose amplitude has decayed below 1e-13 with no constriction feeding it is switched
 * off (amplitude 0); the reference lets it decay for ever.  A source is silent below 1e-3 (TdsModel.cpp:1642-
 * 1660), so the samples are the same. */
typedef struct afs_tds_options {
  int32_t turbulence_losses;
  int32_t soft_walls;
  int32_t generate_noise_sources;
  int32_t radiation_from_skin;
  int32_t piriform_fossa;
  int32_t inner_length_corrections;
  int32_t transvelar_coupling;   /* Dang et al. velum filters (TdsModel.cpp:476-533, 970-980) */
  int32_t glottis_loss_option;   /* enum afs_glottis_loss (VARIABLE: Fulcher 2011, TdsModel.cpp:1019-1092) */
  double flow_separation_area_ratio;
  /* TriangularGlottis static parameters; afs_tds_options_default fills the
   * reference defaults. */
  double glottis_static[AFS_NUM_GLOTTIS_STATIC];
  /* TdsModel::setFlowSource(flow_cm3_s, section) (TdsModel.cpp:629-633): a constant volume-velocity source
   * added to the source term of one tube section (TdsModel.cpp:996-1008).  section -1 = none, the state of a
   * fresh TdsModel (TdsModel.cpp:84) and the default here.  (TdsModel::setPressureSource: the lung pressure is
   * the PRESSURE entry of the glottis control parameters, applied at the first trachea section like
   * Synthesizer.cpp:597 does.) */
  double flow_source_cm3_s;
  int32_t flow_source_section;
  int32_t reserved; /* must be 0 */
} afs_tds_options;

/* Opaque: one CUDA device, its streams and buffers.  Threading / streams: a context serves one host thread at a time.
 * The host-buffer entry points run on the context's own stream and return synchronised.  The *_device entry points
 * are asynchronous on the caller's stream and keep their temporaries in a stream-ordered pool (allocated and freed
 * on that stream), so calls on DIFFERENT streams of one context may be in flight together; outputs are valid once
 * the caller's stream has reached the point of the call. */
typedef struct afs_ctx afs_ctx;

/* ---- life cycle ----------------------------------------------------------- */
int afs_version(void);
const char *afs_status_string(int status);
/* Replaces `new TdsModel()` / `new TlModel()` (Data.cpp:78-88,
 * Synthesizer.cpp:48): binds one CUDA device. */
int afs_create(afs_ctx **ctx, int device_ordinal);
int afs_destroy(afs_ctx *ctx);
/* Text of the last CUDA error seen by this context ("" if none). */
const char *afs_last_cuda_error(const afs_ctx *ctx);
/* Number of kernel launches issued by this context so far (bench.py's
 * `gpu_launches` claim is counted here, not guessed). */
int64_t afs_launch_count(const afs_ctx *ctx);

void afs_tl_options_default(afs_tl_options *opts);   /* TlModel.cpp:30-47 */
void afs_tds_options_default(afs_tds_options *opts); /* TdsModel.cpp:35-44 */
/* A frame equal to a freshly constructed reference Tube (Tube.cpp:46-73,
 * 265-315): 40 sections of 4 cm^2 x 0.4 cm, velum 0, glottis 0.1 cm^2. */
void afs_tube_frame_default(afs_tube_frame *frame);

/* ---- TlModel: transfer-function sweep ------------------------------------- */
/*
 * Replaces TlModel::getSpectrum (TlModel.cpp:123-197) incl.
 * prepareCalculations (:484-676) and getFlowSourceTF / getPressureSourceTF /
 * getInputImpedance / getOutputImpedance (:987-1469) for a batch of tubes.
 * The reference derives the bin grid from its compile-time SAMPLING_RATE and
 * caps it at MAX_NUM_FREQ = 1024 (TlModel.h:89, TlModel.cpp:125-148); here
 * (f0_hz, num_freq) are arguments: bin i is at i*f0_hz, i = 0..num_freq-1,
 * num_freq >= 2 (bin 0 reports the value of bin 1, TlModel.cpp:992,1293).
 * Output: re/im [n_tubes][num_freq] (the non-negative-frequency half of the
 * reference's ComplexSignal; Dsp.cpp:128-138 mirrors it).
 * Host buffers; H2D/D2H copies happen inside the call.
 */
int afs_tl_sweep(afs_ctx *ctx, const afs_tl_options *opts, int n_tubes,
                 const afs_tube_frame *tubes, int spectrum_type, int section,
                 double f0_hz, int num_freq, double *re, double *im);
/* Same, device-resident tubes and outputs, asynchronous on `stream`. */
int afs_tl_sweep_device(afs_ctx *ctx, const afs_tl_options *opts, int n_tubes,
                        const afs_tube_frame *d_tubes, int spectrum_type,
                        int section, double f0_hz, int num_freq, double *d_re,
                        double *d_im, void *stream);

/*
 * Replaces TlModel::getFormants (TlModel.cpp:280-461) for a batch of tubes:
 * FLOW_SOURCE_TF at FIRST_PHARYNX_SECTION on the grid f0 = sampling_rate /
 * spectrum_length (reference: 22050 / 2048), peak picking between 150 Hz and
 * 7 kHz with parabolic interpolation and -3 dB bandwidths.
 * Outputs per tube: formant_hz/bandwidth_hz [n_tubes][max_formants],
 * num_formants [n_tubes], flags [n_tubes] (bit0 frictionNoise, bit1
 * isClosure, bit2 isNasal).
 */
int afs_tl_formants(afs_ctx *ctx, const afs_tl_options *opts, int n_tubes,
                    const afs_tube_frame *tubes, double sampling_rate_hz,
                    int spectrum_length, int max_formants, double *formant_hz,
                    double *bandwidth_hz, int32_t *num_formants, int32_t *flags);

/*
 * Replaces TlModel::getMostConstrictedSection (TlModel.cpp:204-217) and TlModel::getMeanFlow(lungPressure_dPa)
 * (TlModel.cpp:226-246) for a batch of tubes: the narrowest pharynx / mouth section (tube section index 25..64;
 * of equally narrow sections the last one) and sqrt(2 P / (rho (1/A_g^2 + 1/A_c^2))) with A_g the lower glottis
 * area of the frame and A_c the area of that section, both clamped to 0.001 cm^2.  Either output may be null.
 * (The reference's getMeanFlow also stores the pressure in the model; here it is afs_tl_options.lung_pressure_dPa.)
 */
int afs_tl_mean_flow(afs_ctx *ctx, int n_tubes, const afs_tube_frame *tubes, double lung_pressure_dPa,
                     double *mean_flow_cm3_s, int32_t *most_constricted_section);

/*
 * Replaces TlModel::getImpulseResponse (TlModel.cpp:90-111) for a batch of tubes: the flow-source transfer
 * function at FIRST_PHARYNX_SECTION times the radiation characteristic on the grid f0 = sampling_rate_hz / N,
 * N = 1 << length_exponent (bins as in TlModel::getSpectrum, TlModel.cpp:125-148, without the MAX_NUM_FREQ
 * cap), negative frequencies mirrored (Dsp.cpp:128-138), inverse FFT (complexIFFT, Dsp.cpp:228-326, same
 * radix-2 butterflies and the same recursively multiplied twiddle factors) and the rising half-cosine window
 * of getImpulseResponseWindow (TlModel.cpp:70-80).  impulse_response [n_tubes][N].  4 <= length_exponent <= 11.
 * With REFERENCE arithmetic the result is bit-identical to the reference (reference call:
 * getImpulseResponse(&signal, 9) at 22050 Hz, Synthesizer.cpp:719-724).
 */
int afs_tl_impulse_response(afs_ctx *ctx, const afs_tl_options *opts, int n_tubes,
                            const afs_tube_frame *tubes, double sampling_rate_hz,
                            int length_exponent, double *impulse_response);
/* d_tubes / d_impulse_response are device pointers; asynchronous on `stream`. */
int afs_tl_impulse_response_device(afs_ctx *ctx, const afs_tl_options *opts, int n_tubes,
                                   const afs_tube_frame *d_tubes, double sampling_rate_hz,
                                   int length_exponent, double *d_impulse_response, void *stream);

/*
 * Replaces Data::optimizeFormantsVowel with Data::getVowelFormants / getFormantError (Frontend/Data.cpp:203-487,
 * 149-196, 106-135) -- the reference's only caller of repeated TL sweeps -- for a batch of independent problems.
 * Greedy coordinate search: every run evaluates the current parameter set and up to 32 neighbours (each free
 * parameter one step up / down: 0.05 cm for positions, 0.0025 cm^2 for areas, 0.1 for exponents, at most
 * round(max_param_change_cm / 0.05) in [1, 40] steps from the start, control points kept in order exactly as the
 * reference does, including its in-place change of the step size when a neighbour is clamped), takes the
 * neighbour with the smallest formant error (strictly smaller than the current one, minimum area of tube
 * sections 25..39 >= min_advised_area_cm2) and stops when nothing improves or after max_runs runs (reference:
 * 100).  The candidates of ALL problems of a run are evaluated in one batch on the device: rasteriser (K4) ->
 * TL sweep (glottis closed, TlModel::getFormants grid sampling_rate_hz / spectrum_length = 22050 / 2048) ->
 * formant picking (3 formants; fewer than 3 count as F1 = F2 = F3 = 0 like the reference).
 * Hidden state of the reference, kept: TlModel::getFormants leaves lungPressure_dPa = 8000 in the (persistent)
 * Data::tlModel (TlModel.cpp:427, 253-260), so every evaluation after the first runs with 8000 dPa; only the
 * initial error is evaluated with opts->lung_pressure_dPa (0 = a fresh TlModel).
 *   params        [n][16] in/out, order of OneDimAreaFunction::ParamIndex (OneDimAreaFunction.h:34-43)
 *   target_hz     [n][3]  (F3 <= 1: only F1, F2 count)
 *   param_fixed   [16] nullable (non-zero: parameter is not changed)
 *   initial_error, final_error [n] nullable (percent), runs [n] nullable, formants_hz [n][3] nullable (final)
 *   trace         [n][max_runs][2] nullable: parameter changed by every run (-1: none) and the error reached
 */
int afs_optimize_formants_vowel(afs_ctx *ctx, const afs_tl_options *opts, int n_problems, double *params,
                                const double *target_hz, double max_param_change_cm, double min_advised_area_cm2,
                                const int32_t *param_fixed, int max_runs, double sampling_rate_hz, int spectrum_length,
                                double *initial_error, double *final_error, int32_t *runs, double *formants_hz,
                                double *trace);

/*
 * LF glottal pulse parameters: the public members of LfPulse (LfPulse.h:25-29; defaults of
 * LfPulse::resetParams, LfPulse.cpp:24-31, via afs_lf_pulse_default).
 */
typedef struct afs_lf_pulse {
  double amp_cm3_s; /* AMP */
  double f0_hz;     /* F0  */
  double oq;        /* OQ  open quotient */
  double sq;        /* SQ  speed quotient */
  double tl;        /* TL  spectral tilt */
} afs_lf_pulse;
void afs_lf_pulse_default(afs_lf_pulse *p);
/* Samples of one vowel: (int)(0.65 * rate) long, (int)(0.35 * rate) short (Synthesizer.cpp:667-715). */
int afs_lf_vowel_length(int is_long_vowel, double sampling_rate_hz);
/*
 * Replaces Synthesizer::synthesizeVowelLf(tlModel, lfPulse, startPos, isLongVowel) (Synthesizer.cpp:645-801) for a
 * batch of vowels: impulse response of every tube (afs_tl_impulse_response, length 512) scaled by 80 / F0, the LF
 * pulse train with the reference's F0 / amplitude contours and flutter (LfPulse.cpp:37-245, TimeFunction.cpp:77-
 * 133; host side, ~100 pulses per vowel), the 512-tap convolution, the Chebyshev-6 low-pass at 8 kHz, x 2000, cast
 * to int16.  audio [n][length] (length = afs_lf_vowel_length); pressure (nullable) [n][length]: the convolution
 * output before the filter.  With REFERENCE arithmetic the int16 samples are identical to the reference's.
 * AFS_ERR_BAD_ARGUMENT when a pulse would not fit the reference's 2048-sample ring buffer (F0 below rate / 1024).
 */
/* Host helper (no device work): the glottal flow (LF pulse train) of one vowel, flow [afs_lf_vowel_length]:
 * LfPulse::getPulse per period with the F0 / amplitude contours and flutter of Synthesizer.cpp:667-771. */
int afs_lf_glottal_flow(const afs_lf_pulse *pulse, int is_long_vowel, double sampling_rate_hz, double *flow);
int afs_synthesize_vowels_lf(afs_ctx *ctx, const afs_tl_options *opts, int n_vowels, const afs_tube_frame *tubes,
                             const afs_lf_pulse *pulses, int is_long_vowel, double sampling_rate_hz, int16_t *audio,
                             double *pressure);

/*
 * Replaces Synthesizer::playTargetSequence(targetShape[4], stationary_s[4], transition_s[3]) (Synthesizer.cpp:
 * 1299-1422) for a batch of target sequences that share the timing: per OUTPUT sample the 16 area-function
 * parameters are cosine-interpolated between four target shapes (interpolateParameters, :1286-1294), F0 follows
 * the reference's contour 100 -> 115 -> 105 -> 80 Hz, the lung pressure (8000 dPa) its fade-in / fade-out, the
 * parameter set is rasterised (OneDimAreaFunction, on the device: K4) and ONE synthesizeSignalTds(tube,
 * glottisParams, 1, out) call is made per sample -- i.e. a control frame per output sample for the TDS kernel.
 *   shapes [n][4][16] (OneDimAreaFunction::ParamIndex order); glottis_control [6]: control parameters of the
 *   glottis object (FREQUENCY and PRESSURE are overwritten by the contour); audio [n][num_samples], num_samples =
 *   afs_target_sequence_length (the reference's int truncation); sample 0 is the priming call of the reference
 *   (it returns without writing, Synthesizer.cpp:522-532) and is 0 here.  seeds (nullable) [n] as in afs_tds_synthesize.
 * Host helpers (no device work): afs_target_sequence_length, afs_target_sequence_trajectory (params
 * [num_samples][16], glottis [num_samples][6] of one sequence).
 */
int afs_target_sequence_length(const double stationary_s[4], const double transition_s[3], double sampling_rate_hz);
int afs_target_sequence_trajectory(const double *shapes, const double stationary_s[4], const double transition_s[3],
                                   const double glottis_control[AFS_NUM_GLOTTIS_CONTROL], double sampling_rate_hz,
                                   double *params, double *glottis);
int afs_play_target_sequences(afs_ctx *ctx, const afs_tds_options *opts, int n_sequences, const double *shapes,
                              const double stationary_s[4], const double transition_s[3],
                              const double glottis_control[AFS_NUM_GLOTTIS_CONTROL], double sampling_rate_hz,
                              int sr_factor, const uint32_t *seeds, double *audio, uint32_t *status);

/* ---- TdsModel + TriangularGlottis + driver loop ---------------------------- */
/*
 * Replaces, for a batch of independent utterances ("tracts"), the per-sample
 * loop of Synthesizer::synthesizeSignalTds (Synthesizer.cpp:515-639): for
 * every audio sample Tube::interpolate (Tube.cpp:438-505), glottis parameter
 * lerp + TriangularGlottis::calcGeometry/incTime (TriangularGlottis.cpp:154-
 * 397), TdsModel::setTube/setPressureSource/proceedTimeStep (TdsModel.cpp:
 * 552-711: prepareTimeStep :718, calcMatrix :1785, linear solve :2231,
 * updateVariables :2046), flow differentiation and the Chebyshev-8 output
 * low-pass (Synthesizer.cpp:52,614-627).
 *
 * Control frames: tract b has n_frames frames tubes[b*n_frames + k] and
 * glottis[(b*n_frames + k)*6 ...].  Frame 0 primes the driver (the reference's
 * first call only stores prevTube, Synthesizer.cpp:522-532); frame k >= 1
 * produces frame_len[k-1] output samples interpolating linearly from frame
 * k-1 to frame k (ratio i/N, right end excluded), exactly like one reference
 * call synthesizeSignalTds(tube_k, glottis_k, frame_len[k-1], out).
 * The simulation runs at fs_tds_hz = sr_factor * output rate (reference:
 * TDS_SAMPLING_RATE = TDS_SR_FACTOR * SAMPLING_RATE, Constants.h:22-26).  The
 * reference SHIPS TDS_SR_FACTOR = 1 (a simulation at 22.05 kHz); sr_factor = 2
 * with fs_tds_hz = 44100 is the configuration BASELINE.json is quoted on and
 * the default of the Python binding and of the C++ facade -- pass sr_factor =
 * 1, fs_tds_hz = 22050 for the audio of an unmodified reference build.
 *
 * Outputs: audio [n_tracts][n_out], n_out = sum(frame_len) (range ~[-1,1]);
 * flow (nullable) [n_tracts][n_out*sr_factor], the radiated flow returned by
 * TdsModel::proceedTimeStep per simulation step; status (nullable)
 * [n_tracts]: bit0 = non-finite value seen (the reference prints and carries
 * on, TdsModel.cpp:2267); bit1 (device-buffer entry points only, all tracts at
 * once) = the record pool of the noise-source pre-pass was too small and the
 * audio is invalid -- the host-buffer entry points repeat the call with a
 * larger pool themselves, callers of the *_device variants raise it with
 * afs_debug_set("tds.noise_run_pool_factor").
 * seeds (nullable) [n_tracts]: per-tract srand() seeds, used only with
 * generate_noise_sources.
 */
int afs_tds_synthesize(afs_ctx *ctx, const afs_tds_options *opts, int n_tracts,
                       int n_frames, const int32_t *frame_len,
                       const afs_tube_frame *tubes, const double *glottis,
                       double fs_tds_hz, int sr_factor, const uint32_t *seeds,
                       double *audio, double *flow, uint32_t *status);
/* Same with device-resident tubes/glottis/outputs (frame_len stays on the
 * host), asynchronous on `stream`. */
int afs_tds_synthesize_device(afs_ctx *ctx, const afs_tds_options *opts,
                              int n_tracts, int n_frames,
                              const int32_t *frame_len,
                              const afs_tube_frame *d_tubes,
                              const double *d_glottis, double fs_tds_hz,
                              int sr_factor, const uint32_t *d_seeds,
                              double *d_audio, double *d_flow,
                              uint32_t *d_status, void *stream);

/*
 * Batched taps into the running simulation: TdsModel::getSectionPressure(section) (TdsModel.cpp:1725-1734) and
 * TdsModel::getSectionFlow(section, inflow, outflow) (TdsModel.cpp:1715-1719), i.e. the pressure in the centre of
 * a tube section and the volume velocities into / out of it, as a caller of the reference reads them after every
 * proceedTimeStep.  Up to AFS_MAX_TDS_PROBES sections (tube section indices 0..92) per call; outputs
 * [n_tracts][n_probes][n_steps] with n_steps = n_out * sr_factor simulation steps; each of the three may be null.
 */
#define AFS_MAX_TDS_PROBES 8
typedef struct afs_tds_probes {
  int32_t n_probes;
  int32_t section[AFS_MAX_TDS_PROBES];
  double *pressure_dPa;
  double *inflow_cm3_s;
  double *outflow_cm3_s;
} afs_tds_probes;
/* afs_tds_synthesize / afs_tds_synthesize_device with probes (host resp. device output pointers). */
int afs_tds_synthesize_probed(afs_ctx *ctx, const afs_tds_options *opts, int n_tracts, int n_frames,
                              const int32_t *frame_len, const afs_tube_frame *tubes, const double *glottis,
                              double fs_tds_hz, int sr_factor, const uint32_t *seeds, double *audio, double *flow,
                              uint32_t *status, const afs_tds_probes *probes);
int afs_tds_synthesize_probed_device(afs_ctx *ctx, const afs_tds_options *opts, int n_tracts, int n_frames,
                                     const int32_t *frame_len, const afs_tube_frame *d_tubes, const double *d_glottis,
                                     double fs_tds_hz, int sr_factor, const uint32_t *d_seeds, double *d_audio,
                                     double *d_flow, uint32_t *d_status, const afs_tds_probes *d_probes, void *stream);

/*
 * Chunked synthesis: a batch of utterances whose models live ACROSS calls, the way the reference's Synthesizer is
 * used (Synthesizer::synthesizeSignalTds is called once per control frame and continues from the state the TdsModel,
 * the TriangularGlottis and the driver -- prevTube, prevGlottisParams, the output filter, the last flow sample -- were
 * left in, Synthesizer.cpp:515-639; SURVEY.md section 5 "checkpoint / resume": state in / state out per tract).
 * The recurrent state of every tract (pressures, currents, wall currents and their rates, noise sources and the
 * tract's random() stream, glottis displacements, filters: ~15 KB) stays in device memory between calls; utterances
 * of any length are synthesised piece by piece in bounded memory, and the audio of a chunked utterance is the audio
 * of the same frames given to afs_tds_synthesize in one call (tests/test_tds_stream_gpu.py: bit-identical with the
 * noise sources off, within 1e-9 -- measured 7e-10 -- with them on, for every chunking and option set).
 *   afs_tds_stream_create      n_tracts fresh models (`new TdsModel`, `new TriangularGlottis`, Synthesizer::init);
 *                              seeds nullable, as in afs_tds_synthesize (used by the first chunk only).
 *   afs_tds_stream_synthesize  n_frames >= 1 further control frames per tract, tubes [n_tracts][n_frames], glottis
 *                              [n_tracts][n_frames][6]: frame k is one reference call synthesizeSignalTds(tube_k,
 *                              glottis_k, frame_len[k], out) -- frame_len[k] output samples interpolated from the
 *                              PREVIOUS frame (of this call or of the last one) to frame k.  Like the reference, the
 *                              very first frame of a fresh stream only primes (its frame_len is ignored, no samples).
 *                              audio [n_tracts][n_out], n_out = sum of the frame_len that produced samples; flow
 *                              (nullable) [n_tracts][n_out * sr_factor]; status (nullable) as in afs_tds_synthesize.
 *                              A call that fails leaves the stream in the state it had before the call.
 *   afs_tds_stream_reset       TdsModel::resetMotion + a driver that has not been primed.
 *   afs_tds_stream_samples     output samples per tract produced so far.
 *   afs_tds_stream_save/load   checkpoint / resume: the whole stream (state, last frames, counters) as one host blob of
 *                              afs_tds_stream_checkpoint_bytes bytes; load accepts a blob written by a stream of the
 *                              same batch size, rates and options (else AFS_ERR_BAD_ARGUMENT).
 * A stream belongs to its context (same threading rule); destroy it before the context.  Taps (afs_tds_probes) and the
 * state dump are served by the one-call entry points only.  The audio of a chunk leaves the device while its kernel
 * runs, like in afs_tds_synthesize; pass page-locked buffers (afs_host_alloc) for the copies to overlap.
 */
typedef struct afs_tds_stream afs_tds_stream;
int afs_tds_stream_create(afs_ctx *ctx, const afs_tds_options *opts, int n_tracts, double fs_tds_hz, int sr_factor,
                          const uint32_t *seeds, afs_tds_stream **stream);
int afs_tds_stream_destroy(afs_tds_stream *stream);
int afs_tds_stream_reset(afs_tds_stream *stream);
int64_t afs_tds_stream_samples(const afs_tds_stream *stream);
/* 1 once the stream holds a previous frame (the next frame produces samples), 0 before the priming frame. */
int afs_tds_stream_primed(const afs_tds_stream *stream);
int afs_tds_stream_synthesize(afs_tds_stream *stream, int n_frames, const int32_t *frame_len, const afs_tube_frame *tubes,
                              const double *glottis, double *audio, double *flow, uint32_t *status);
size_t afs_tds_stream_checkpoint_bytes(const afs_tds_stream *stream);
int afs_tds_stream_save(afs_tds_stream *stream, void *blob, size_t bytes);
int afs_tds_stream_load(afs_tds_stream *stream, const void *blob, size_t bytes);

/* Diagnostics: the next afs_tds_synthesize* calls copy the recurrent state of
 * tract 0 right after simulation step `step` (1-based; 0 disables) into
 * host_out473: [0..92] section pressure, [93..185] pressure rate, [186..278]
 * wall current, [279..375] branch current, [376..472] branch-current rate
 * (TdsModel.h:141-143,166-167,184). */
int afs_tds_set_state_dump(afs_ctx *ctx, int step, double *host_out473);

/* Test hooks, not tuning knobs of the product: "tds.noise_scan_bytes" (device budget of the per-step noise pre-pass
 * records, used when the control frames are a few steps apart; 0 = derived from the free device memory: tests force
 * several waves with a small value); "tds.noise_run_pool_factor" (records of the run-record pool of the noise pre-pass per
 * piece of a control-frame interval, default 32: tests force the overflow path with 2); "tds.d2h_slices" (time slices
 * of the overlapped device-to-host copy of afs_tds_synthesize, 0 = default 16, 1 = one copy after the kernel);
 * "tds.small_batch_max" (largest batch served by the one-tract-per-CTA kernel tds_small_kernel; -1 = default, three
 * tracts per SM; 0 = never: tests run both kernels on the same input); "tds.spread_sms" (spread a TDS batch over
 * this many CTAs instead of packing 28 tracts into every CTA, 0 = default: tests and measurements of the packing).  Unknown keys: AFS_ERR_BAD_ARGUMENT. */
int afs_debug_set(afs_ctx *ctx, const char *key, int64_t value);

/* Measures the FP64 FMA peak of the bound device (independent DFMA chains,
 * best of 4), in TFLOP/s: the roofline denominator of both kernels
 * (MEASURED_PEAKS.json has no FP64 figure). */
int afs_probe_fp64_tflops(afs_ctx *ctx, double *tflops);

/* ---- several devices of one box, one process ------------------------------------ */
/*
 * The units of the path (utterances, area functions) are independent (Synthesizer.cpp:557-629 keeps one TdsModel
 * per utterance): a group shards a batch into contiguous ranges [r*B/G, (r+1)*B/G) over its G
 * devices, runs every shard through the single-device entry point of the same name on its own host thread and
 * stream, and lets every device write its rows of the caller's output arrays directly -- there is no data-path
 * collective and no gather step: the outputs of all devices land in ONE host array.  Results are those of the
 * single-device call, bit for bit, whatever the number of devices.
 *   afs_group_create: device_ordinals null / n_devices <= 0 = every visible device.
 *   afs_group_context: the afs_ctx of device `index` (for the *_device entry points).
 * Pinned buffers (afs_host_alloc) let the copies of the eight devices run at full PCIe rate beside the kernels.
 */
typedef struct afs_group afs_group;
int afs_group_create(afs_group **group, const int *device_ordinals, int n_devices);
int afs_group_destroy(afs_group *group);
int afs_group_size(const afs_group *group);
afs_ctx *afs_group_context(afs_group *group, int index);
/* [lo, hi) of the batch of n_units that device `index` takes (sizes differ by at most one). */
int afs_group_shard(const afs_group *group, int n_units, int index, int *lo, int *hi);
int afs_group_tds_synthesize(afs_group *group, const afs_tds_options *opts, int n_tracts, int n_frames,
                             const int32_t *frame_len, const afs_tube_frame *tubes, const double *glottis,
                             double fs_tds_hz, int sr_factor, const uint32_t *seeds, double *audio, double *flow,
                             uint32_t *status);
int afs_group_tl_sweep(afs_group *group, const afs_tl_options *opts, int n_tubes, const afs_tube_frame *tubes,
                       int spectrum_type, int section, double f0_hz, int num_freq, double *re, double *im);
/* Chunked synthesis (afs_tds_stream_*) over the devices of a group: every device keeps the models of its shard;
 * arguments and results as for the single-device stream, bit for bit. */
typedef struct afs_group_tds_stream afs_group_tds_stream;
int afs_group_tds_stream_create(afs_group *group, const afs_tds_options *opts, int n_tracts, double fs_tds_hz, int sr_factor,
                                const uint32_t *seeds, afs_group_tds_stream **stream);
int afs_group_tds_stream_destroy(afs_group_tds_stream *stream);
int afs_group_tds_stream_reset(afs_group_tds_stream *stream);
int afs_group_tds_stream_synthesize(afs_group_tds_stream *stream, int n_frames, const int32_t *frame_len,
                                    const afs_tube_frame *tubes, const double *glottis, double *audio, double *flow,
                                    uint32_t *status);
int64_t afs_group_tds_stream_samples(const afs_group_tds_stream *stream);
/* checkpoint / resume: the blobs of the per-device streams back to back (a blob belongs to one sharding) */
size_t afs_group_tds_stream_checkpoint_bytes(const afs_group_tds_stream *stream);
int afs_group_tds_stream_save(afs_group_tds_stream *stream, void *blob, size_t bytes);
int afs_group_tds_stream_load(afs_group_tds_stream *stream, const void *blob, size_t bytes);
/* Page-locked host memory (cudaHostAlloc, portable across the devices of a group). */
int afs_host_alloc(void **ptr, size_t bytes);
int afs_host_free(void *ptr);

/* ---- host-side helpers of the boundary (no device work) -------------------- */
/* IirFilter::createChebyshev low-pass (IirFilter.cpp:286-432): recursion
 * coefficients a[0..num_poles], b[0..num_poles] (b[0] = 0) for
 * y[n] = sum a_i x[n-i] + sum b_i y[n-i].  num_poles even, <= 16. */
int afs_chebyshev_lowpass(double cutoff_ratio, int num_poles, double *a, double *b);

/* The int16 conversion of Synthesizer::synthesizeSegment (Synthesizer.cpp:955-965): sample * SHRT_MAX truncated
 * towards zero, clipped to SHRT_MAX / SHRT_MIN outside [-1, 1].  n samples. */
int afs_audio_to_int16(const double *audio, int64_t n, int16_t *pcm);
/* On-disk replacement of the reference's PortAudio ring buffer (Synthesizer.cpp:955-973, 1429-1483): a
 * canonical 16-bit PCM RIFF/WAVE file, `n_frames` sample frames of `n_channels` interleaved channels. */
int afs_write_wav_int16(const char *path, const int16_t *pcm, int64_t n_frames, int n_channels, int sample_rate_hz);
/* OneDimAreaFunction::calculateOneDimTubeFunction (OneDimAreaFunction.cpp:
 * 23-138): 16 parameters (Llar, Alar, powLar, xp, Ap, powP, xc, Ac, powC, xa,
 * Aa, powA, xin, Ain, Lvt, Alip) -> the 40 pharynx/mouth sections of `frame`
 * (other fields untouched). Input generator for benches; host code. */
int afs_area_function_to_frame(const double params[16], afs_tube_frame *frame);
/* The same for a batch on the device (K4 area_function_raster_kernel; SURVEY.md
 * 8f rank 1): params [n][16] -> frames [n].  Replaces n calls of
 * OneDimAreaFunction::calculateOneDimTubeFunction (OneDimAreaFunction.cpp:75-138,
 * as called from Data.cpp / the parametric-tract sweeps).  Sample positions,
 * samples per section and articulator labels are the reference's exactly;
 * areas agree to ~1e-15 relative (device pow/cos vs glibc).  Fills the 40
 * pharynx/mouth sections, their articulators and the incisor position; the
 * other fields of `frames` are left as they are (initialise them with
 * afs_tube_frame_default).  The host variant copies frames in and out. */
int afs_area_functions_to_frames(afs_ctx *ctx, int n, const double *params, afs_tube_frame *frames);
int afs_area_functions_to_frames_device(afs_ctx *ctx, int n, const double *d_params, afs_tube_frame *d_frames,
                                        void *stream);

#ifdef __cplusplus
}
#endif
#endif /* AFS_B200_H */
