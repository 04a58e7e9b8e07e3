// afs_api.cu -- the C ABI of include/afs_b200.h: context, argument checks,
// host<->device staging, kernel launches.  No compute happens on the host and
// there is no CPU fallback: without a CUDA device afs_create fails.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "afs_internal.h"

using namespace afs;

struct afs_ctx {
  int device = -1;
  cudaStream_t stream = nullptr;
  int64_t launches = 0;
  char last_error[256] = {0};
  // grow-only device scratch
  void *scratch[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t scratch_bytes[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int sm_count = 0;
  int dump_step = 0;          // diagnostics, see afs_tds_set_state_dump
  double *dump_host = nullptr;
};

namespace {

int cuda_fail(afs_ctx *ctx, cudaError_t e, const char *what) {
  if (ctx) snprintf(ctx->last_error, sizeof ctx->last_error, "%s: %s", what, cudaGetErrorString(e));
  return e == cudaErrorMemoryAllocation ? AFS_ERR_OUT_OF_MEMORY : AFS_ERR_CUDA;
}
#define CU(call)                                                 \
  do {                                                           \
    cudaError_t e__ = (call);                                    \
    if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call);   \
  } while (0)

int bind(afs_ctx *ctx) {
  if (!ctx || ctx->device < 0) return AFS_ERR_BAD_CONTEXT;
  CU(cudaSetDevice(ctx->device));
  return AFS_OK;
}

int scratch(afs_ctx *ctx, int slot, size_t bytes, void **out) {
  if (ctx->scratch_bytes[slot] < bytes) {
    if (ctx->scratch[slot]) { CU(cudaFree(ctx->scratch[slot])); ctx->scratch[slot] = nullptr; ctx->scratch_bytes[slot] = 0; }
    CU(cudaMalloc(&ctx->scratch[slot], bytes));
    ctx->scratch_bytes[slot] = bytes;
  }
  *out = ctx->scratch[slot];
  return AFS_OK;
}

int fill_tl_params(const afs_tl_options *o, int n_tubes, int type, int section, double f0, int num_freq, TlParams *p) {
  if (!o || n_tubes < 0 || num_freq < 2 || !(f0 > 0.0)) return AFS_ERR_BAD_ARGUMENT;  // bin 0 aliases bin 1
  if (type < AFS_INPUT_IMPEDANCE || type > AFS_RADIATION) return AFS_ERR_BAD_ARGUMENT;
  if (section < -1 || section >= AFS_NUM_SECTIONS) return AFS_ERR_BAD_ARGUMENT;
  if (o->radiation < AFS_NO_RADIATION || o->radiation > AFS_PARALLEL_RADIATION) return AFS_ERR_BAD_ARGUMENT;
  if (o->arithmetic != AFS_TL_ARITHMETIC_REFERENCE && o->arithmetic != AFS_TL_ARITHMETIC_FAST) return AFS_ERR_BAD_ARGUMENT;
  if (o->reserved != 0) return AFS_ERR_BAD_ARGUMENT;
  p->o = *o;
  p->n_tubes = n_tubes; p->type = type; p->section = section; p->num_freq = num_freq; p->f0 = f0;
  // TlModel.cpp:14-19 -- same expressions, host libm
  p->rs_factor = sqrt(kRho * kMu * 0.5);
  p->cp_factor = 1.0 / (kRho * kC * kC);
  p->gp_factor = ((kKappa - 1.0) * sqrt((0.5 * kHeatCond) / (kSpecHeat * kRho))) / (kRho * kC * kC);
  static const double sinus_vol[4] = {11.3, 6.8, 33.0, 6.2};  // Tube.cpp:159
  for (int i = 0; i < 4; i++) {
    const double inner = 4.0 * M_PI * pow((3.0 * sinus_vol[i]) / (4.0 * M_PI), 2.0 / 3.0);  // TlModel.cpp:858
    p->sinus_gp[i] = inner / 65000.0;
  }
  return AFS_OK;
}

}  // namespace

extern "C" {

int afs_version(void) { return AFS_VERSION; }

const char *afs_status_string(int s) {
  switch (s) {
    case AFS_OK: return "ok";
    case AFS_ERR_NO_DEVICE: return "no CUDA device (there is no CPU fallback)";
    case AFS_ERR_BAD_CONTEXT: return "bad context";
    case AFS_ERR_BAD_ARGUMENT: return "bad argument";
    case AFS_ERR_UNSUPPORTED: return "option not supported on the device path";
    case AFS_ERR_CUDA: return "CUDA error";
    case AFS_ERR_OUT_OF_MEMORY: return "out of device memory";
  }
  return "unknown status";
}

int afs_create(afs_ctx **out, int device_ordinal) {
  if (!out) return AFS_ERR_BAD_ARGUMENT;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return AFS_ERR_NO_DEVICE; }
  if (device_ordinal < 0 || device_ordinal >= n) return AFS_ERR_BAD_ARGUMENT;
  afs_ctx *ctx = new (std::nothrow) afs_ctx();
  if (!ctx) return AFS_ERR_OUT_OF_MEMORY;
  ctx->device = device_ordinal;
  cudaError_t e = cudaSetDevice(device_ordinal);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device_ordinal);
  if (e != cudaSuccess) { delete ctx; cudaGetLastError(); return AFS_ERR_CUDA; }
  *out = ctx;
  return AFS_OK;
}

int afs_destroy(afs_ctx *ctx) {
  if (!ctx) return AFS_OK;
  if (ctx->device >= 0) {
    cudaSetDevice(ctx->device);
    for (int i = 0; i < 8; i++) if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
  }
  delete ctx;
  return AFS_OK;
}

const char *afs_last_cuda_error(const afs_ctx *ctx) { return ctx ? ctx->last_error : ""; }
int64_t afs_launch_count(const afs_ctx *ctx) { return ctx ? ctx->launches : 0; }

void afs_tl_options_default(afs_tl_options *o) {
  if (!o) return;
  o->radiation = AFS_PARALLEL_RADIATION;
  o->boundary_layer = 1; o->heat_conduction = 0; o->soft_walls = 1; o->hagen_resistance = 0;
  o->inner_length_corrections = 0; o->lumped_elements = 1; o->paranasal_sinuses = 1; o->piriform_fossa = 0;
  o->static_pressure_drops = 1; o->lung_pressure_dPa = 0.0;
  o->arithmetic = AFS_TL_ARITHMETIC_REFERENCE; o->reserved = 0;
}

void afs_tds_options_default(afs_tds_options *o) {
  if (!o) return;
  static const double gs[AFS_NUM_GLOTTIS_STATIC] = {1.3, 0.24, 0.06, 0.12, 0.03, 0.1, 0.6, 80000.0, 8000.0,
                                                    240000.0, 24000.0, 25000.0, 0.05, 0.01, 129.0, 125.51};
  o->turbulence_losses = 1; o->soft_walls = 1; o->generate_noise_sources = 1; o->radiation_from_skin = 1;
  o->piriform_fossa = 0; o->inner_length_corrections = 1; o->transvelar_coupling = 0;
  o->glottis_loss_option = AFS_STANDARD_ENTRANCE_LOSS; o->flow_separation_area_ratio = 1.0;
  for (int i = 0; i < AFS_NUM_GLOTTIS_STATIC; i++) o->glottis_static[i] = gs[i];
}

void afs_tube_frame_default(afs_tube_frame *f) {
  if (!f) return;
  memset(f, 0, sizeof *f);
  for (int i = 0; i < AFS_NUM_PM_SECTIONS; i++) {
    f->area_cm2[i] = 4.0; f->length_cm[i] = 16.0 / (double)AFS_NUM_PM_SECTIONS; f->articulator[i] = AFS_OTHER_ARTICULATOR;
  }
  f->incisor_pos_cm = 15.0; f->aspiration_dB = -40.0;
  f->glottis_area_cm2[0] = f->glottis_area_cm2[1] = 0.1;
  f->glottis_length_cm[0] = f->glottis_length_cm[1] = 0.3;
}

// ---------------------------------------------------------------------------
// Area-function rasteriser (K4)
// ---------------------------------------------------------------------------
int afs_area_functions_to_frames_device(afs_ctx *ctx, int n, const double *d_params, afs_tube_frame *d_frames,
                                        void *stream) {
  int rc = bind(ctx);
  if (rc) return rc;
  if (n < 0) return AFS_ERR_BAD_ARGUMENT;
  if (n == 0) return AFS_OK;
  if (!d_params || !d_frames) return AFS_ERR_BAD_ARGUMENT;
  CU(launch_area_function_raster(n, d_params, d_frames, (cudaStream_t)stream));
  ctx->launches++;
  return AFS_OK;
}

int afs_area_functions_to_frames(afs_ctx *ctx, int n, const double *params, afs_tube_frame *frames) {
  int rc = bind(ctx);
  if (rc) return rc;
  if (n < 0) return AFS_ERR_BAD_ARGUMENT;
  if (n == 0) return AFS_OK;
  if (!params || !frames) return AFS_ERR_BAD_ARGUMENT;
  void *d_p, *d_f;
  const size_t pb = (size_t)n * 16 * sizeof(double), fb = (size_t)n * sizeof(afs_tube_frame);
  if ((rc = scratch(ctx, 1, pb, &d_p))) return rc;
  if ((rc = scratch(ctx, 0, fb, &d_f))) return rc;
  CU(cudaMemcpyAsync(d_p, params, pb, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(d_f, frames, fb, cudaMemcpyHostToDevice, ctx->stream));
  rc = afs_area_functions_to_frames_device(ctx, n, (const double *)d_p, (afs_tube_frame *)d_f, ctx->stream);
  if (rc) return rc;
  CU(cudaMemcpyAsync(frames, d_f, fb, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return AFS_OK;
}

int afs_probe_fp64_tflops(afs_ctx *ctx, double *tflops) {
  int rc = bind(ctx);
  if (rc) return rc;
  if (!tflops) return AFS_ERR_BAD_ARGUMENT;
  void *d_out;
  if ((rc = scratch(ctx, 6, 473 * sizeof(double), &d_out))) return rc;
  const int ctas = ctx->sm_count * 8, iters = 1 << 15;
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {  // first rep is the warm-up
    CU(cudaEventRecord(e0, ctx->stream));
    CU(launch_fp64_probe((double *)d_out, ctas, iters, ctx->stream));
    ctx->launches++;
    CU(cudaEventRecord(e1, ctx->stream));
    CU(cudaEventSynchronize(e1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double flops = 2.0 * 8.0 * (double)iters * 256.0 * (double)ctas;
  *tflops = flops / ((double)best * 1e-3) / 1e12;
  return AFS_OK;
}

// ---------------------------------------------------------------------------
// TL
// ---------------------------------------------------------------------------
int afs_tl_sweep_device(afs_ctx *ctx, const afs_tl_options *opts, int n_tubes, const afs_tube_frame *d_tubes,
                        int spectrum_type, int section, double f0_hz, int num_freq, double *d_re, double *d_im,
                        void *stream) {
  int rc = bind(ctx);
  if (rc) return rc;
  TlParams p;
  rc = fill_tl_params(opts, n_tubes, spectrum_type, section, f0_hz, num_freq, &p);
  if (rc) return rc;
  if (n_tubes == 0) return AFS_OK;
  if (!d_tubes || !d_re || !d_im) return AFS_ERR_BAD_ARGUMENT;
  if (n_tubes > 65535) {  // grid.y limit: split
    for (int b = 0; b < n_tubes; b += 65535) {
      const int nb = n_tubes - b < 65535 ? n_tubes - b : 65535;
      rc = afs_tl_sweep_device(ctx, opts, nb, d_tubes + b, spectrum_type, section, f0_hz, num_freq,
                               d_re + (size_t)b * num_freq, d_im + (size_t)b * num_freq, stream);
      if (rc) return rc;
    }
    return AFS_OK;
  }
  if (p.o.arithmetic == AFS_TL_ARITHMETIC_FAST && p.o.lumped_elements && spectrum_type != AFS_RADIATION)
    CU(launch_tl_sweep_fast(p, d_tubes, d_re, d_im, (cudaStream_t)stream));
  else
    CU(launch_tl_sweep(p, d_tubes, d_re, d_im, (cudaStream_t)stream));
  ctx->launches++;
  return AFS_OK;
}

int afs_tl_sweep(afs_ctx *ctx, const afs_tl_options *opts, int n_tubes, const afs_tube_frame *tubes,
                 int spectrum_type, int section, double f0_hz, int num_freq, double *re, double *im) {
  int rc = bind(ctx);
  if (rc) return rc;
  if (n_tubes < 0 || num_freq < 1) return AFS_ERR_BAD_ARGUMENT;
  if (n_tubes == 0) return AFS_OK;
  if (!tubes || !re || !im) return AFS_ERR_BAD_ARGUMENT;
  void *d_tubes, *d_re, *d_im;
  const size_t nb = (size_t)n_tubes * num_freq * sizeof(double);
  if ((rc = scratch(ctx, 0, (size_t)n_tubes * sizeof(afs_tube_frame), &d_tubes))) return rc;
  if ((rc = scratch(ctx, 1, nb, &d_re))) return rc;
  if ((rc = scratch(ctx, 2, nb, &d_im))) return rc;
  CU(cudaMemcpyAsync(d_tubes, tubes, (size_t)n_tubes * sizeof(afs_tube_frame), cudaMemcpyHostToDevice, ctx->stream));
  rc = afs_tl_sweep_device(ctx, opts, n_tubes, (const afs_tube_frame *)d_tubes, spectrum_type, section, f0_hz,
                           num_freq, (double *)d_re, (double *)d_im, ctx->stream);
  if (rc) return rc;
  CU(cudaMemcpyAsync(re, d_re, nb, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(im, d_im, nb, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return AFS_OK;
}

int afs_tl_formants(afs_ctx *ctx, const afs_tl_options *opts, int n_tubes, const afs_tube_frame *tubes,
                    double sampling_rate_hz, int spectrum_length, int max_formants, double *formant_hz,
                    double *bandwidth_hz, int32_t *num_formants, int32_t *flags) {
  int rc = bind(ctx);
  if (rc) return rc;
  if (n_tubes < 0 || spectrum_length < 4 || max_formants < 1 || max_formants > 32 || !(sampling_rate_hz > 0))
    return AFS_ERR_BAD_ARGUMENT;
  if (n_tubes == 0) return AFS_OK;
  if (!tubes || !formant_hz || !bandwidth_hz || !num_formants || !flags) return AFS_ERR_BAD_ARGUMENT;
  // bin grid of TlModel::getSpectrum (TlModel.cpp:125-148) without the MAX_NUM_FREQ cap
  const double f0 = sampling_rate_hz / (double)spectrum_length;
  int nf = (int)(10100.0 / f0);
  if (nf >= spectrum_length / 2) nf = spectrum_length / 2 - 1;
  void *d_tubes, *d_re, *d_im, *d_f, *d_b, *d_n, *d_fl;
  const size_t nb = (size_t)n_tubes * nf * sizeof(double), fb = (size_t)n_tubes * max_formants * sizeof(double);
  if ((rc = scratch(ctx, 0, (size_t)n_tubes * sizeof(afs_tube_frame), &d_tubes))) return rc;
  if ((rc = scratch(ctx, 1, nb, &d_re))) return rc;
  if ((rc = scratch(ctx, 2, nb, &d_im))) return rc;
  if ((rc = scratch(ctx, 3, fb, &d_f))) return rc;
  if ((rc = scratch(ctx, 4, fb, &d_b))) return rc;
  if ((rc = scratch(ctx, 5, (size_t)n_tubes * sizeof(int32_t), &d_n))) return rc;
  if ((rc = scratch(ctx, 6, (size_t)n_tubes * sizeof(int32_t), &d_fl))) return rc;
  CU(cudaMemcpyAsync(d_tubes, tubes, (size_t)n_tubes * sizeof(afs_tube_frame), cudaMemcpyHostToDevice, ctx->stream));
  rc = afs_tl_sweep_device(ctx, opts, n_tubes, (const afs_tube_frame *)d_tubes, AFS_FLOW_SOURCE_TF,
                           AFS_FIRST_PHARYNX_SECTION, f0, nf, (double *)d_re, (double *)d_im, ctx->stream);
  if (rc) return rc;
  FormantParams fp;
  fp.n_tubes = n_tubes; fp.num_freq = nf; fp.max_formants = max_formants; fp.f0 = f0;
  CU(launch_formant_pick(fp, (const afs_tube_frame *)d_tubes, (const double *)d_re, (const double *)d_im,
                         (double *)d_f, (double *)d_b, (int32_t *)d_n, (int32_t *)d_fl, ctx->stream));
  ctx->launches++;
  CU(cudaMemcpyAsync(formant_hz, d_f, fb, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(bandwidth_hz, d_b, fb, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(num_formants, d_n, (size_t)n_tubes * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(flags, d_fl, (size_t)n_tubes * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return AFS_OK;
}

int afs_tl_impulse_response_device(afs_ctx *ctx, const afs_tl_options *opts, int n_tubes,
                                   const afs_tube_frame *d_tubes, double sampling_rate_hz, int length_exponent,
                                   double *d_out, void *stream) {
  int rc = bind(ctx);
  if (rc) return rc;
  if (n_tubes < 0 || length_exponent < 4 || length_exponent > 11 || !(sampling_rate_hz > 0)) return AFS_ERR_BAD_ARGUMENT;
  if (n_tubes == 0) return AFS_OK;
  if (!d_tubes || !d_out) return AFS_ERR_BAD_ARGUMENT;
  // bin grid of TlModel::getSpectrum (TlModel.cpp:125-148) without the MAX_NUM_FREQ cap
  const int N = 1 << length_exponent;
  const double f0 = sampling_rate_hz / (double)N;
  int nf = (int)(10100.0 / f0);
  if (nf >= N / 2) nf = N / 2 - 1;
  if (nf < 2) return AFS_ERR_BAD_ARGUMENT;
  void *d_re, *d_im, *d_tab;
  const size_t nb = (size_t)n_tubes * nf * sizeof(double);
  const size_t ntab = (size_t)(2 * (N - 1) + 2 * nf + N);
  if ((rc = scratch(ctx, 1, nb, &d_re))) return rc;
  if ((rc = scratch(ctx, 2, nb, &d_im))) return rc;
  if ((rc = scratch(ctx, 7, ntab * sizeof(double), &d_tab))) return rc;
  std::vector<double> tab(ntab);
  double *tw_re = tab.data(), *tw_im = tw_re + (N - 1), *rad_re = tw_im + (N - 1), *rad_im = rad_re + nf,
         *window = rad_im + nf;
  impulse_response_tables(f0, length_exponent, nf, tw_re, tw_im, rad_re, rad_im, window);
  cudaStream_t st = (cudaStream_t)stream;
  // pageable source: the copy is staged before the call returns, `tab` may go out of scope
  CU(cudaMemcpyAsync(d_tab, tab.data(), ntab * sizeof(double), cudaMemcpyHostToDevice, st));
  rc = afs_tl_sweep_device(ctx, opts, n_tubes, d_tubes, AFS_FLOW_SOURCE_TF, AFS_FIRST_PHARYNX_SECTION, f0, nf,
                           (double *)d_re, (double *)d_im, stream);
  if (rc) return rc;
  IrTables t;
  const double *base = (const double *)d_tab;
  t.tw_re = base; t.tw_im = base + (N - 1); t.rad_re = base + 2 * (N - 1); t.rad_im = t.rad_re + nf;
  t.window = t.rad_im + nf;
  CU(launch_impulse_response(n_tubes, length_exponent, nf, (const double *)d_re, (const double *)d_im, t, d_out, st));
  ctx->launches++;
  return AFS_OK;
}

int afs_tl_impulse_response(afs_ctx *ctx, const afs_tl_options *opts, int n_tubes, const afs_tube_frame *tubes,
                            double sampling_rate_hz, int length_exponent, double *impulse_response) {
  int rc = bind(ctx);
  if (rc) return rc;
  if (n_tubes < 0 || length_exponent < 4 || length_exponent > 11) return AFS_ERR_BAD_ARGUMENT;
  if (n_tubes == 0) return AFS_OK;
  if (!tubes || !impulse_response) return AFS_ERR_BAD_ARGUMENT;
  void *d_tubes, *d_out;
  const size_t ob = (size_t)n_tubes * ((size_t)1 << length_exponent) * sizeof(double);
  if ((rc = scratch(ctx, 0, (size_t)n_tubes * sizeof(afs_tube_frame), &d_tubes))) return rc;
  if ((rc = scratch(ctx, 3, ob, &d_out))) return rc;
  CU(cudaMemcpyAsync(d_tubes, tubes, (size_t)n_tubes * sizeof(afs_tube_frame), cudaMemcpyHostToDevice, ctx->stream));
  rc = afs_tl_impulse_response_device(ctx, opts, n_tubes, (const afs_tube_frame *)d_tubes, sampling_rate_hz,
                                      length_exponent, (double *)d_out, ctx->stream);
  if (rc) return rc;
  CU(cudaMemcpyAsync(impulse_response, d_out, ob, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return AFS_OK;
}

// ---------------------------------------------------------------------------
// TDS
// ---------------------------------------------------------------------------
int afs_tds_synthesize_device(afs_ctx *ctx, const afs_tds_options *opts, int n_tracts, int n_frames,
                              const int32_t *frame_len, const afs_tube_frame *d_tubes, const double *d_glottis,
                              double fs_tds_hz, int sr_factor, const uint32_t *d_seeds, double *d_audio,
                              double *d_flow, uint32_t *d_status, void *stream) {
  int rc = bind(ctx);
  if (rc) return rc;
  if (!opts || n_tracts < 0 || n_frames < 2 || !frame_len || sr_factor < 1 || !(fs_tds_hz > 0))
    return AFS_ERR_BAD_ARGUMENT;
  if (opts->glottis_loss_option < AFS_STANDARD_ENTRANCE_LOSS || opts->glottis_loss_option > AFS_VARIABLE_ENTRANCE_LOSS)
    return AFS_ERR_BAD_ARGUMENT;
  for (int k = 0; k + 1 < n_frames; k++) if (frame_len[k] < 1) return AFS_ERR_BAD_ARGUMENT;
  if (n_tracts == 0) return AFS_OK;
  if (!d_tubes || !d_glottis || !d_audio) return AFS_ERR_BAD_ARGUMENT;

  TdsLaunch L;
  rc = tds_prepare_launch(opts, n_tracts, n_frames, frame_len, fs_tds_hz, sr_factor, ctx->sm_count, &L);
  if (rc) return rc;
  // per-launch frame boundaries live in device scratch slot 7
  void *d_bounds;
  if ((rc = scratch(ctx, 7, sizeof(int32_t) * (size_t)(n_frames + 1), &d_bounds))) return rc;
  CU(cudaMemcpyAsync(d_bounds, L.step_bounds.data(), sizeof(int32_t) * (size_t)n_frames, cudaMemcpyHostToDevice,
                     (cudaStream_t)stream));
  void *d_dump = nullptr;
  if (ctx->dump_step > 0 && ctx->dump_host) {
    if ((rc = scratch(ctx, 6, 473 * sizeof(double), &d_dump))) return rc;
    CU(cudaMemsetAsync(d_dump, 0, 473 * sizeof(double), (cudaStream_t)stream));
  }
  CU(launch_tds(L, (const int32_t *)d_bounds, d_tubes, d_glottis, d_seeds, d_audio, d_flow, d_status,
                (double *)d_dump, ctx->dump_step, (cudaStream_t)stream));
  ctx->launches++;
  if (d_dump) {
    CU(cudaMemcpyAsync(ctx->dump_host, d_dump, 473 * sizeof(double), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CU(cudaStreamSynchronize((cudaStream_t)stream));
  }
  return AFS_OK;
}

int afs_tds_set_state_dump(afs_ctx *ctx, int step, double *host_out473) {
  if (!ctx) return AFS_ERR_BAD_CONTEXT;
  ctx->dump_step = step;
  ctx->dump_host = host_out473;
  return AFS_OK;
}

int afs_tds_synthesize(afs_ctx *ctx, const afs_tds_options *opts, int n_tracts, int n_frames,
                       const int32_t *frame_len, const afs_tube_frame *tubes, const double *glottis,
                       double fs_tds_hz, int sr_factor, const uint32_t *seeds, double *audio, double *flow,
                       uint32_t *status) {
  int rc = bind(ctx);
  if (rc) return rc;
  if (n_tracts < 0 || n_frames < 2 || !frame_len || sr_factor < 1) return AFS_ERR_BAD_ARGUMENT;
  if (n_tracts == 0) return AFS_OK;
  if (!tubes || !glottis || !audio) return AFS_ERR_BAD_ARGUMENT;
  size_t n_out = 0;
  for (int k = 0; k + 1 < n_frames; k++) { if (frame_len[k] < 1) return AFS_ERR_BAD_ARGUMENT; n_out += (size_t)frame_len[k]; }
  const size_t tb = (size_t)n_tracts * n_frames * sizeof(afs_tube_frame);
  const size_t gb = (size_t)n_tracts * n_frames * AFS_NUM_GLOTTIS_CONTROL * sizeof(double);
  const size_t ab = (size_t)n_tracts * n_out * sizeof(double);
  const size_t fb = ab * (size_t)sr_factor;
  void *d_tubes, *d_gl, *d_audio, *d_flow = nullptr, *d_status, *d_seeds = nullptr;
  if ((rc = scratch(ctx, 0, tb, &d_tubes))) return rc;
  if ((rc = scratch(ctx, 1, gb, &d_gl))) return rc;
  if ((rc = scratch(ctx, 2, ab, &d_audio))) return rc;
  if (flow && (rc = scratch(ctx, 3, fb, &d_flow))) return rc;
  if ((rc = scratch(ctx, 4, (size_t)n_tracts * sizeof(uint32_t), &d_status))) return rc;
  if (seeds) {
    if ((rc = scratch(ctx, 5, (size_t)n_tracts * sizeof(uint32_t), &d_seeds))) return rc;
    CU(cudaMemcpyAsync(d_seeds, seeds, (size_t)n_tracts * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  }
  CU(cudaMemcpyAsync(d_tubes, tubes, tb, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(d_gl, glottis, gb, cudaMemcpyHostToDevice, ctx->stream));
  rc = afs_tds_synthesize_device(ctx, opts, n_tracts, n_frames, frame_len, (const afs_tube_frame *)d_tubes,
                                 (const double *)d_gl, fs_tds_hz, sr_factor, (const uint32_t *)d_seeds,
                                 (double *)d_audio, (double *)d_flow, (uint32_t *)d_status, ctx->stream);
  if (rc) return rc;
  CU(cudaMemcpyAsync(audio, d_audio, ab, cudaMemcpyDeviceToHost, ctx->stream));
  if (flow) CU(cudaMemcpyAsync(flow, d_flow, fb, cudaMemcpyDeviceToHost, ctx->stream));
  if (status) CU(cudaMemcpyAsync(status, d_status, (size_t)n_tracts * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return AFS_OK;
}

}  // extern "C"
