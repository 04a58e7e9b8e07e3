/* oracle/afs_oracle.h -- TEST INFRASTRUCTURE (see afs_oracle.c). */
#ifndef AFS_ORACLE_H
#define AFS_ORACLE_H
#include "../include/afs_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* One expanded section of a reference Tube (Tube.h:33-51). */
typedef struct orc_section {
  double pos, area, length, volume, Mw, Bw, Kw, laterality;
  int articulator;
} orc_section;

void orc_expand_frame(const afs_tube_frame *f, orc_section sec[AFS_NUM_SECTIONS]);

int orc_tl_spectrum(const afs_tl_options *opts, int n_tubes, const afs_tube_frame *tubes,
                    int type, int section, double f0_hz, int num_freq, double *re, double *im);
int orc_tl_formants(const afs_tl_options *opts, int n_tubes, const afs_tube_frame *tubes,
                    double sampling_rate_hz, int spectrum_length, int max_formants,
                    double *freq, double *bw, int32_t *num, int32_t *flags);

int orc_tl_impulse_response(const afs_tl_options *opts, int n_tubes, const afs_tube_frame *tubes,
                            double sampling_rate_hz, int length_exponent, double *out);

void orc_chebyshev_lowpass(double cutoff_ratio, int num_poles, double *a, double *b);

int orc_tds_synthesize(const afs_tds_options *opts, int n_tracts, int n_frames,
                       const int32_t *frame_len, const afs_tube_frame *tubes,
                       const double *glottis, double fs_tds_hz, int sr_factor,
                       const uint32_t *seeds, double *audio, double *flow,
                       double *state473, int dump_step);

int orc_tds_synthesize_probed(const afs_tds_options *opts, int n_tracts, int n_frames, const int32_t *frame_len,
                              const afs_tube_frame *tubes, const double *glottis, double fs_tds_hz, int sr_factor,
                              const uint32_t *seeds, double *audio, int n_probes, const int32_t *sections,
                              double *pressure, double *inflow, double *outflow);
int orc_tl_mean_flow(int n, const afs_tube_frame *frames, double lung_pressure_dPa, double *flow, int32_t *section);

void orc_area_function_to_frame(const double params[16], afs_tube_frame *frame);

#ifdef __cplusplus
}
#endif
#endif
