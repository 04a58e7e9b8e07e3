// afs_internal.h -- launch-parameter structs and launcher prototypes shared by
// the C-ABI layer (afs_api.cu) and the kernel translation units.
#pragma once
#include <vector>

#include "afs_common.cuh"

namespace afs {

// ---- TL -------------------------------------------------------------------
struct TlParams {
  afs_tl_options o;
  int n_tubes, type, section, num_freq;
  double f0;
  double rs_factor, cp_factor, gp_factor;  // TlModel.cpp:14-19, evaluated on the host with libm
  double sinus_gp[4];                      // 4 pi (3V/4pi)^(2/3) / 65000 per sinus (TlModel.cpp:858,863), host libm pow
};
struct FormantParams {
  int n_tubes, num_freq, max_formants;
  double f0;
};
cudaError_t launch_tl_sweep(const TlParams &p, const afs_tube_frame *d_tubes, double *d_re, double *d_im,
                            cudaStream_t stream);
cudaError_t launch_tl_sweep_fast(const TlParams &p, const afs_tube_frame *d_tubes, double *d_re, double *d_im,
                                 cudaStream_t stream);  // tl_fast_kernels.cu (afs_tl_options.arithmetic = FAST)
cudaError_t launch_formant_pick(const FormantParams &p, const afs_tube_frame *d_tubes, const double *d_re,
                                const double *d_im, double *d_freq, double *d_bw, int32_t *d_num, int32_t *d_flags,
                                cudaStream_t stream);

cudaError_t launch_tl_mean_flow(int n, const afs_tube_frame *d_tubes, double lung_pressure_dPa, double *d_flow,
                                int32_t *d_section, cudaStream_t stream);

// impulse response: tables [N-1] twiddles (re, im), [num_freq] radiation characteristic (re, im), [N] window, all
// computed on the host with libm exactly as the reference does (afs_host.cpp: impulse_response_tables)
struct IrTables {
  const double *tw_re, *tw_im, *rad_re, *rad_im, *window;
};
cudaError_t launch_impulse_response(int n_tubes, int length_exponent, int num_freq, const double *d_re,
                                    const double *d_im, const IrTables &t, double *d_out, cudaStream_t stream);
void impulse_response_tables(double f0_hz, int length_exponent, int num_freq, double *tw_re, double *tw_im,
                             double *rad_re, double *rad_im, double *window);

// LF vowel synthesis (Synthesizer::synthesizeVowelLf): convolution + output filter on the device (tl_kernels.cu),
// pulse train on the host (afs_host.cpp)
cudaError_t launch_lf_convolve(int n, int length, const double *d_h512, const double *d_scale, const double *d_flow,
                               double *d_pressure, cudaStream_t stream);
cudaError_t launch_lf_filter(int n, int length, const double *d_ab14, const double *d_pressure, int16_t *d_audio,
                             cudaStream_t stream);
int lf_vowel_length(int is_long, double sampling_rate_hz);
int lf_vowel_flow(const afs_lf_pulse *lf, int is_long, double sampling_rate_hz, double *flow);

// ---- TDS ------------------------------------------------------------------
struct TdsParams;  // tds_kernels.cu
struct TdsLaunch {
  std::vector<int32_t> step_bounds;  // [n_frames]: first simulation step of every frame interval, then the total
  std::vector<unsigned char> params; // opaque TdsParams blob
  int n_tracts = 0, n_frames = 0;
};
int tds_prepare_launch(const afs_tds_options *opts, int n_tracts, int n_frames, const int32_t *frame_len,
                       double fs_tds_hz, int sr_factor, int spread_ctas, TdsLaunch *L);
cudaError_t launch_tds(const TdsLaunch &L, int tract0, int n_wave, const int32_t *d_bounds, const afs_tube_frame *d_tubes,
                       const double *d_glottis, const uint32_t *d_seeds, double *d_audio, double *d_flow,
                       uint32_t *d_status, double *d_dump, int dump_step, void *d_nzc, const void *d_runs,
                       const uint32_t *d_run_ctl, int run_items, const afs_tds_probes *d_probes, unsigned int *d_progress,
                       int progress_every, const double *d_state_in, double *d_state_out, cudaStream_t stream);
// chunked synthesis: doubles of recurrent state per tract (d_state_in: null = a fresh model; d_state_out: null = dropped)
size_t tds_state_words();
// noise pre-pass in runs mode (control frames far apart)
bool tds_noise_runs_mode(const TdsLaunch &L);
size_t tds_noise_run_bytes();
// items4: {interval, first step, last step, next first step} per piece; iv_first: first piece of every interval, then the total
void tds_noise_run_items(const TdsLaunch &L, std::vector<int32_t> *items4, std::vector<int32_t> *iv_first);
cudaError_t launch_noise_runs(const TdsLaunch &L, int n_items, const void *d_items, const int32_t *d_iv_first,
                              const int32_t *d_bounds, const afs_tube_frame *d_tubes, void *d_runs, uint32_t pool_records,
                              uint32_t *d_ctl, cudaStream_t stream);
int tds_total_steps(const TdsLaunch &L);
// small batches: one tract per CTA, one tube section per lane (tds_small_kernel)
bool tds_small_supported(const TdsLaunch &L);
cudaError_t launch_tds_small(const TdsLaunch &L, const int32_t *d_bounds, const afs_tube_frame *d_tubes, const double *d_glottis,
                             double *d_audio, double *d_flow, uint32_t *d_status, cudaStream_t stream);
size_t tds_noise_scan_bytes_per_cta(const TdsLaunch &L);  // per-step records; 0 without noise sources or in runs mode
int tds_lanes_per_cta(const TdsLaunch &L);
size_t tds_smem_bytes();

// ---- area-function rasteriser -----------------------------------------------
cudaError_t launch_area_function_raster(int n_sets, const double *d_params, afs_tube_frame *d_frames, cudaStream_t stream);
cudaError_t launch_fill_frames(size_t n, afs_tube_frame *d_frames, const afs_tube_frame &tpl, cudaStream_t stream);
cudaError_t launch_target_params(int n_seq, int ns, const double *d_shapes, const int32_t *d_seg, const double *d_cosv,
                                 double *d_params, cudaStream_t stream);
cudaError_t launch_broadcast_row(size_t n, size_t row_doubles, double *d_a, cudaStream_t stream);
int target_sequence_length(const double *stationary_s, const double *transition_s, double sampling_rate_hz);
void target_sequence_timebase(const double *stationary_s, const double *transition_s, const double *glottis6, double SR,
                              int numSamples, int *seg_out, double *cosv, double *glottis);
void target_sequence_params(const double *shapes4x16, const int *seg, const double *cosv, int numSamples, double *params);
void target_sequence_trajectory(const double *shapes4x16, const double *stationary_s, const double *transition_s,
                                const double *glottis6, double SR, int numSamples, double *params, double *glottis);

// ---- probes ------------------------------------------------------------------
cudaError_t launch_fp64_probe(double *d_out, int ctas, int iters, cudaStream_t stream);

// ---- host helpers (afs_host.cpp) -------------------------------------------
void chebyshev_lowpass(double cutoff_ratio, int num_poles, double *a, double *b);

}  // namespace afs
