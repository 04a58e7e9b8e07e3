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
  cudaStream_t copy_stream = nullptr;  // device->host copies of finished chunks beside the next chunk's kernel
  cudaEvent_t chunk_done[2] = {nullptr, nullptr};
  int64_t launches = 0;
  char last_error[256] = {0};
  // grow-only device scratch
  void *scratch[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t scratch_bytes[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  int sm_count = 0;
  int dump_step = 0;          // diagnostics, see afs_tds_set_state_dump
  double *dump_host = nullptr;
  // Stream-ordered pool for what the *_device entry points need beyond the caller's buffers: allocated and freed on
  // the caller's stream, so calls on different streams never share a buffer (the grow-only `scratch` slots above
  // belong to the host-buffer entry points, which run on ctx->stream and return synchronised).
  cudaMemPool_t pool = nullptr;
  // progress marks of a running TDS launch, in mapped host memory (see afs_tds_synthesize_probed)
  unsigned int *progress_h = nullptr, *progress_d = nullptr;
  // test hooks (afs_debug_set)
  int64_t dbg_noise_scan_bytes = 0;
  int dbg_tds_spread_sms = 0;      // > 0: spread a TDS batch over this many CTAs (default: 28 tracts per CTA, as few CTAs as possible)
  int dbg_small_batch_max = -1;  // largest batch served by tds_small_kernel (-1: 3 tracts per SM, 0: never)
  int dbg_d2h_slices = 0;
  int run_pool_factor = 32;   // records of the noise pre-pass pool per piece (see tds_run_device); grows on overflow
};
static constexpr int kMaxProgressCtas = 4096;

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

// Stream-ordered temporary: valid for work enqueued on `stream` until pool_free on the same stream.
int pool_alloc(afs_ctx *ctx, size_t bytes, cudaStream_t stream, void **out) {
  CU(cudaMallocFromPoolAsync(out, bytes ? bytes : 1, ctx->pool, stream));
  return AFS_OK;
}
void pool_free(void *p, cudaStream_t stream) {
  if (p) cudaFreeAsync(p, stream);
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
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
  for (int i = 0; i < 2 && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&ctx->chunk_done[i], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device_ordinal);
  if (e == cudaSuccess) {
    cudaMemPoolProps props;
    memset(&props, 0, sizeof props);
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device_ordinal;
    e = cudaMemPoolCreate(&ctx->pool, &props);
    if (e == cudaSuccess) {  // keep freed blocks: the same sizes come back call after call
      unsigned long long keep = ~0ull;
      e = cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  if (e == cudaSuccess) e = cudaHostAlloc((void **)&ctx->progress_h, kMaxProgressCtas * sizeof(unsigned int), cudaHostAllocMapped);
  if (e == cudaSuccess) e = cudaHostGetDevicePointer((void **)&ctx->progress_d, ctx->progress_h, 0);
  if (e != cudaSuccess) { afs_destroy(ctx); cudaGetLastError(); return AFS_ERR_CUDA; }
  *out = ctx;
  return AFS_OK;
}

int afs_destroy(afs_ctx *ctx) {
  if (!ctx) return AFS_OK;
  if (ctx->device >= 0) {
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < 10; i++) if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (int i = 0; i < 2; i++) if (ctx->chunk_done[i]) cudaEventDestroy(ctx->chunk_done[i]);
    if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
    if (ctx->progress_h) cudaFreeHost(ctx->progress_h);
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
  o->flow_source_cm3_s = 0.0; o->flow_source_section = -1; o->reserved = 0;  // TdsModel.cpp:84
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
  // Chunks of area functions: the spectra of chunk i travel to the host (copy stream) while chunk i + 1 is swept.
  // With pinned caller buffers the 16 B per bin of output leave at PCIe rate beside the kernel; pageable buffers are
  // staged by the driver and gain nothing, but lose nothing either.
  const int chunk = n_tubes >= 1024 ? (n_tubes + 7) / 8 : n_tubes;
  int k = 0;
  cudaError_t e = cudaSuccess;
  rc = AFS_OK;
  for (int b = 0; b < n_tubes && rc == AFS_OK && e == cudaSuccess; b += chunk, k++) {
    const int nb_t = n_tubes - b < chunk ? n_tubes - b : chunk;
    const size_t off = (size_t)b * num_freq, bytes = (size_t)nb_t * num_freq * sizeof(double);
    rc = afs_tl_sweep_device(ctx, opts, nb_t, (const afs_tube_frame *)d_tubes + b, spectrum_type, section, f0_hz,
                             num_freq, (double *)d_re + off, (double *)d_im + off, ctx->stream);
    if (rc) break;
    cudaEvent_t ev = ctx->chunk_done[k & 1];
    if ((e = cudaEventRecord(ev, ctx->stream)) != cudaSuccess) break;
    if ((e = cudaStreamWaitEvent(ctx->copy_stream, ev, 0)) != cudaSuccess) break;
    if ((e = cudaMemcpyAsync(re + off, (double *)d_re + off, bytes, cudaMemcpyDeviceToHost, ctx->copy_stream)) != cudaSuccess) break;
    e = cudaMemcpyAsync(im + off, (double *)d_im + off, bytes, cudaMemcpyDeviceToHost, ctx->copy_stream);
  }
  // both streams are drained on EVERY path: earlier chunks may still be on their way into the caller's buffers
  const cudaError_t e1 = cudaStreamSynchronize(ctx->copy_stream), e2 = cudaStreamSynchronize(ctx->stream);
  if (rc) return rc;
  if (e != cudaSuccess) return cuda_fail(ctx, e, "afs_tl_sweep: chunk copy");
  if (e1 != cudaSuccess) return cuda_fail(ctx, e1, "cudaStreamSynchronize(copy stream)");
  if (e2 != cudaSuccess) return cuda_fail(ctx, e2, "cudaStreamSynchronize");
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

int afs_tl_mean_flow(afs_ctx *ctx, int n_tubes, const afs_tube_frame *tubes, double lung_pressure_dPa,
                     double *mean_flow_cm3_s, int32_t *most_constricted_section) {
  int rc = bind(ctx);
  if (rc) return rc;
  if (n_tubes < 0) return AFS_ERR_BAD_ARGUMENT;
  if (n_tubes == 0) return AFS_OK;
  if (!tubes) return AFS_ERR_BAD_ARGUMENT;
  void *d_tubes, *d_f, *d_s;
  if ((rc = scratch(ctx, 0, (size_t)n_tubes * sizeof(afs_tube_frame), &d_tubes))) return rc;
  if ((rc = scratch(ctx, 3, (size_t)n_tubes * sizeof(double), &d_f))) return rc;
  if ((rc = scratch(ctx, 5, (size_t)n_tubes * sizeof(int32_t), &d_s))) return rc;
  CU(cudaMemcpyAsync(d_tubes, tubes, (size_t)n_tubes * sizeof(afs_tube_frame), cudaMemcpyHostToDevice, ctx->stream));
  CU(launch_tl_mean_flow(n_tubes, (const afs_tube_frame *)d_tubes, lung_pressure_dPa, (double *)d_f, (int32_t *)d_s, ctx->stream));
  ctx->launches++;
  if (mean_flow_cm3_s) CU(cudaMemcpyAsync(mean_flow_cm3_s, d_f, (size_t)n_tubes * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (most_constricted_section)
    CU(cudaMemcpyAsync(most_constricted_section, d_s, (size_t)n_tubes * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
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
  void *d_re = nullptr, *d_im = nullptr, *d_tab = nullptr;
  const size_t nb = (size_t)n_tubes * nf * sizeof(double);
  const size_t ntab = (size_t)(2 * (N - 1) + 2 * nf + N);
  cudaStream_t st = (cudaStream_t)stream;
  // stream-ordered temporaries (see afs_ctx::pool): calls on different streams do not share them
  auto release = [&]() { pool_free(d_re, st); pool_free(d_im, st); pool_free(d_tab, st); };
  if ((rc = pool_alloc(ctx, nb, st, &d_re)) || (rc = pool_alloc(ctx, nb, st, &d_im)) ||
      (rc = pool_alloc(ctx, ntab * sizeof(double), st, &d_tab))) { release(); return rc; }
  std::vector<double> tab(ntab);
  double *tw_re = tab.data(), *tw_im = tw_re + (N - 1), *rad_re = tw_im + (N - 1), *rad_im = rad_re + nf,
         *window = rad_im + nf;
  impulse_response_tables(f0, length_exponent, nf, tw_re, tw_im, rad_re, rad_im, window);
  // pageable source: the copy is staged before the call returns, `tab` may go out of scope
  cudaError_t e = cudaMemcpyAsync(d_tab, tab.data(), ntab * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) { release(); return cuda_fail(ctx, e, "cudaMemcpyAsync(impulse-response tables)"); }
  rc = afs_tl_sweep_device(ctx, opts, n_tubes, d_tubes, AFS_FLOW_SOURCE_TF, AFS_FIRST_PHARYNX_SECTION, f0, nf,
                           (double *)d_re, (double *)d_im, stream);
  if (rc) { release(); return rc; }
  IrTables t;
  const double *base = (const double *)d_tab;
  t.tw_re = base; t.tw_im = base + (N - 1); t.rad_re = base + 2 * (N - 1); t.rad_im = t.rad_re + nf;
  t.window = t.rad_im + nf;
  e = launch_impulse_response(n_tubes, length_exponent, nf, (const double *)d_re, (const double *)d_im, t, d_out, st);
  release();
  if (e != cudaSuccess) return cuda_fail(ctx, e, "launch_impulse_response");
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
// Target sequences: Synthesizer::playTargetSequence, batched
// ---------------------------------------------------------------------------
int afs_play_target_sequences(afs_ctx *ctx, const afs_tds_options *opts, int n, const double *shapes,
                              const double stationary_s[4], const double transition_s[3],
                              const double glottis_control[AFS_NUM_GLOTTIS_CONTROL], double sampling_rate_hz,
                              int sr_factor, const uint32_t *seeds, double *audio, uint32_t *status) {
  int rc = bind(ctx);
  if (rc) return rc;
  if (!opts || n < 0 || !stationary_s || !transition_s || !glottis_control || !(sampling_rate_hz > 0.0) || sr_factor < 1)
    return AFS_ERR_BAD_ARGUMENT;
  if (n == 0) return AFS_OK;
  if (!shapes || !audio) return AFS_ERR_BAD_ARGUMENT;
  const int ns = target_sequence_length(stationary_s, transition_s, sampling_rate_hz);
  if (ns < 2) return AFS_ERR_BAD_ARGUMENT;
  const size_t nf = (size_t)n * (size_t)ns;  // control frames = parameter sets
  if (nf > 0x7fffffffull) return AFS_ERR_BAD_ARGUMENT;
  {  // the frames (1184 B per output sample and sequence) stay on the device; they have to fit
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    const size_t need = nf * (sizeof(afs_tube_frame) + 22 * sizeof(double)) + (size_t)n * ns * sizeof(double);
    if (need > free_b + ctx->scratch_bytes[0] + ctx->scratch_bytes[1] + ctx->scratch_bytes[2]) return AFS_ERR_OUT_OF_MEMORY;
  }
  // host: the time base (shared by all sequences: per-sample segment, cosine and glottis parameters, ~70 B per
  // sample); device: the parameter trajectories of every sequence (n x ns x 16 doubles never cross PCIe)
  std::vector<double> gl((size_t)ns * AFS_NUM_GLOTTIS_CONTROL), cosv((size_t)ns);
  std::vector<int> seg((size_t)ns);
  target_sequence_timebase(stationary_s, transition_s, glottis_control, sampling_rate_hz, ns, seg.data(), cosv.data(), gl.data());
  void *d_frames, *d_params, *d_gl, *d_audio, *d_status, *d_seeds = nullptr, *d_tb;
  if ((rc = scratch(ctx, 0, nf * sizeof(afs_tube_frame), &d_frames))) return rc;
  if ((rc = scratch(ctx, 1, nf * 16 * sizeof(double), &d_params))) return rc;
  if ((rc = scratch(ctx, 2, nf * AFS_NUM_GLOTTIS_CONTROL * sizeof(double), &d_gl))) return rc;
  if ((rc = scratch(ctx, 3, (size_t)n * (ns - 1) * sizeof(double), &d_audio))) return rc;
  if ((rc = scratch(ctx, 4, (size_t)n * sizeof(uint32_t), &d_status))) return rc;
  if ((rc = scratch(ctx, 9, (size_t)n * 64 * sizeof(double) + (size_t)ns * (sizeof(double) + sizeof(int32_t)), &d_tb))) return rc;
  double *d_shapes = (double *)d_tb, *d_cosv = d_shapes + (size_t)n * 64;
  int32_t *d_seg = (int32_t *)(d_cosv + ns);
  if (seeds) {
    if ((rc = scratch(ctx, 5, (size_t)n * sizeof(uint32_t), &d_seeds))) return rc;
    CU(cudaMemcpyAsync(d_seeds, seeds, (size_t)n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  }
  CU(cudaMemcpyAsync(d_shapes, shapes, (size_t)n * 64 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(d_cosv, cosv.data(), (size_t)ns * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(d_seg, seg.data(), (size_t)ns * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(d_gl, gl.data(), (size_t)ns * AFS_NUM_GLOTTIS_CONTROL * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CU(launch_target_params(n, ns, d_shapes, d_seg, d_cosv, (double *)d_params, ctx->stream));
  CU(launch_broadcast_row((size_t)n, (size_t)ns * AFS_NUM_GLOTTIS_CONTROL, (double *)d_gl, ctx->stream));
  ctx->launches += 2;
  afs_tube_frame tpl;  // a fresh Tube (nasal port closed; the glottis sections come from the glottis model)
  afs_tube_frame_default(&tpl);
  CU(launch_fill_frames(nf, (afs_tube_frame *)d_frames, tpl, ctx->stream));
  CU(launch_area_function_raster((int)nf, (const double *)d_params, (afs_tube_frame *)d_frames, ctx->stream));
  ctx->launches += 2;
  std::vector<int32_t> frame_len((size_t)ns - 1, 1);  // one synthesizeSignalTds(tube, glottisParams, 1, out) per sample
  rc = afs_tds_synthesize_device(ctx, opts, n, ns, frame_len.data(), (const afs_tube_frame *)d_frames, (const double *)d_gl,
                                 sampling_rate_hz * sr_factor, sr_factor, (const uint32_t *)d_seeds, (double *)d_audio, nullptr,
                                 (uint32_t *)d_status, ctx->stream);
  if (rc) return rc;
  // sample 0: the priming call produces nothing
  for (int b = 0; b < n; b++) audio[(size_t)b * ns] = 0.0;
  CU(cudaMemcpy2DAsync(audio + 1, (size_t)ns * sizeof(double), d_audio, (size_t)(ns - 1) * sizeof(double),
                       (size_t)(ns - 1) * sizeof(double), (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  if (status) CU(cudaMemcpyAsync(status, d_status, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return AFS_OK;
}

// ---------------------------------------------------------------------------
// Formant optimiser: Data::optimizeFormantsVowel, batched
// ---------------------------------------------------------------------------
namespace {
enum { LLAR, ALAR, POWLAR, XP, AP, POWP, XC, AC, POWC, XA, AA, POWA, XIN, AIN, LVT, ALIP };  // OneDimAreaFunction.h:34-43

// Data::getFormantError (Data.cpp:106-135)
double formant_error(const double *c, const double *t) {
  double e;
  if (t[2] > 1.0) {
    e = (1.0 - c[0] / t[0]) * (1.0 - c[0] / t[0]) + (1.0 - c[1] / t[1]) * (1.0 - c[1] / t[1]) +
        (1.0 - c[2] / t[2]) * (1.0 - c[2] / t[2]);
    e /= 3.0;
  } else {
    e = (1.0 - c[0] / t[0]) * (1.0 - c[0] / t[0]) + (1.0 - c[1] / t[1]) * (1.0 - c[1] / t[1]);
    e /= 2.0;
  }
  e = sqrt(e);
  return e * 100.0;
}

// Data::getVowelFormants (Data.cpp:149-196) for a batch of parameter sets, everything on the device: F [n][3]
// (zeros when fewer than three formants were found) and the minimum area of tube sections 25..39 (sic: the loop
// of the reference runs from FIRST_PHARYNX_SECTION to NUM_PHARYNX_MOUTH_SECTIONS - 1).
int vowel_formants(afs_ctx *ctx, const afs_tl_options *opts, int n, const double *params, double sr, int N,
                   std::vector<afs_tube_frame> &frames, double *F, double *min_area) {
  const double f0 = sr / (double)N;
  int nf = (int)(10100.0 / f0);
  if (nf >= N / 2) nf = N / 2 - 1;
  if (nf < 2) return AFS_ERR_BAD_ARGUMENT;
  int rc;
  void *d_tubes, *d_re, *d_im, *d_f, *d_b, *d_n, *d_fl, *d_p;
  const size_t nb = (size_t)n * nf * sizeof(double), fb = (size_t)n * 3 * sizeof(double);
  if ((rc = scratch(ctx, 0, (size_t)n * sizeof(afs_tube_frame), &d_tubes))) return rc;
  if ((rc = scratch(ctx, 1, nb, &d_re))) return rc;
  if ((rc = scratch(ctx, 2, nb, &d_im))) return rc;
  if ((rc = scratch(ctx, 3, fb, &d_f))) return rc;
  if ((rc = scratch(ctx, 4, fb, &d_b))) return rc;
  if ((rc = scratch(ctx, 5, (size_t)n * sizeof(int32_t), &d_n))) return rc;
  if ((rc = scratch(ctx, 6, (size_t)n * sizeof(int32_t), &d_fl))) return rc;
  if ((rc = scratch(ctx, 7, (size_t)n * 16 * sizeof(double), &d_p))) return rc;
  afs_tube_frame tpl;  // fresh Tube + setGlottisArea(0.0) (Data.cpp:165-172)
  afs_tube_frame_default(&tpl);
  tpl.glottis_area_cm2[0] = tpl.glottis_area_cm2[1] = 0.0;
  frames.assign((size_t)n, tpl);
  CU(cudaMemcpyAsync(d_tubes, frames.data(), (size_t)n * sizeof(afs_tube_frame), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(d_p, params, (size_t)n * 16 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CU(launch_area_function_raster(n, (const double *)d_p, (afs_tube_frame *)d_tubes, ctx->stream));
  ctx->launches++;
  rc = afs_tl_sweep_device(ctx, opts, n, (const afs_tube_frame *)d_tubes, AFS_FLOW_SOURCE_TF, AFS_FIRST_PHARYNX_SECTION,
                           f0, nf, (double *)d_re, (double *)d_im, ctx->stream);
  if (rc) return rc;
  FormantParams fp;
  fp.n_tubes = n; fp.num_freq = nf; fp.max_formants = 3; fp.f0 = f0;
  CU(launch_formant_pick(fp, (const afs_tube_frame *)d_tubes, (const double *)d_re, (const double *)d_im, (double *)d_f,
                         (double *)d_b, (int32_t *)d_n, (int32_t *)d_fl, ctx->stream));
  ctx->launches++;
  std::vector<int32_t> num((size_t)n);
  CU(cudaMemcpyAsync(F, d_f, fb, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(num.data(), d_n, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(frames.data(), d_tubes, (size_t)n * sizeof(afs_tube_frame), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < n; i++) {
    if (num[i] < 3) F[3 * i] = F[3 * i + 1] = F[3 * i + 2] = 0.0;
    double m = 10000.0;
    for (int k = 0; k < AFS_NUM_PHARYNX_SECTIONS - 1; k++) if (frames[i].area_cm2[k] < m) m = frames[i].area_cm2[k];
    min_area[i] = m;
  }
  return AFS_OK;
}

struct OptProblem {
  double cur[16], step[16];
  int taken[16];
  bool active;
  int runs;
};
struct OptCand { int problem, param; double change; };  // param -1: the current configuration
}  // namespace

int afs_optimize_formants_vowel(afs_ctx *ctx, const afs_tl_options *opts, int n_problems, double *params,
                                const double *target_hz, double max_param_change_cm, double min_advised_area_cm2,
                                const int32_t *param_fixed, int max_runs, double sampling_rate_hz, int spectrum_length,
                                double *initial_error, double *final_error, int32_t *runs, double *formants_hz,
                                double *trace) {
  int rc = bind(ctx);
  if (rc) return rc;
  if (n_problems < 0 || max_runs < 1 || spectrum_length < 4 || !(sampling_rate_hz > 0) || !opts) return AFS_ERR_BAD_ARGUMENT;
  if (n_problems == 0) return AFS_OK;
  if (!params || !target_hz) return AFS_ERR_BAD_ARGUMENT;
  const double STEP = 0.05;  // Data.cpp:258
  int max_steps = (int)((max_param_change_cm / STEP) + 0.5);
  if (max_steps < 1) max_steps = 1;
  if (max_steps > 40) max_steps = 40;
  std::vector<OptProblem> pr((size_t)n_problems);
  for (int b = 0; b < n_problems; b++) {
    OptProblem &q = pr[b];
    memcpy(q.cur, params + 16 * (size_t)b, sizeof q.cur);
    for (int i = 0; i < 16; i++) {
      const bool pos = i == LLAR || i == XP || i == XC || i == XA || i == XIN || i == LVT;
      const bool pw = i == POWLAR || i == POWP || i == POWC || i == POWA;
      q.step[i] = pos ? STEP : pw ? 0.1 : STEP * STEP;  // Data.cpp:277-295
      if (param_fixed && param_fixed[i]) q.step[i] = 0.0;
      q.taken[i] = 0;
    }
    q.active = true;
    q.runs = 0;
  }
  std::vector<OptCand> cand;
  std::vector<double> cp, F, ma;
  std::vector<afs_tube_frame> frames;
  // Hidden state of the reference: TlModel::getFormants ends with getMeanFlow(8000.0) (TlModel.cpp:427), which
  // leaves lungPressure_dPa = 8000 in the model (:253-260).  Data::tlModel is one persistent
  // object, so every evaluation AFTER the first one runs with 8000 dPa (static pressure drops, :526-541); the
  // first one -- the initial error, Data.cpp:224-246 -- sees the state the model was handed over in (opts).
  afs_tl_options o_rest = *opts;
  o_rest.lung_pressure_dPa = 8000.0;
  {
    F.resize((size_t)n_problems * 3); ma.resize((size_t)n_problems);
    if ((rc = vowel_formants(ctx, opts, n_problems, params, sampling_rate_hz, spectrum_length, frames, F.data(), ma.data()))) return rc;
    if (initial_error)
      for (int b = 0; b < n_problems; b++) initial_error[b] = formant_error(&F[3 * (size_t)b], target_hz + 3 * (size_t)b);
  }
  opts = &o_rest;
  for (int run = 0; run < max_runs; run++) {
    cand.clear(); cp.clear();
    for (int b = 0; b < n_problems; b++) {
      OptProblem &q = pr[b];
      if (!q.active) continue;
      cand.push_back({b, -1, 0.0});
      cp.insert(cp.end(), q.cur, q.cur + 16);
      double c[16];
      memcpy(c, q.cur, sizeof c);
      for (int i = 0; i < 16; i++) {  // Data.cpp:337-431: the neighbours, generated in the reference's order
        const double v = c[i];
        if (!(q.step[i] > 0.0)) continue;
        if (q.taken[i] < max_steps) {
          c[i] = v + q.step[i];
          const int up = i == XP ? XC : i == XC ? XA : i == XA ? XIN : i == XIN ? LVT : -1;
          if (up >= 0 && c[i] > c[up]) { q.step[i] = c[up] - c[i]; c[i] = c[up]; }
          cand.push_back({b, i, q.step[i]});
          cp.insert(cp.end(), c, c + 16);
        }
        if (q.taken[i] > -max_steps) {
          c[i] = v - q.step[i];
          const int dn = i == XP ? LLAR : i == XC ? XP : i == XA ? XC : i == XIN ? XA : i == LVT ? XIN : -1;
          if (dn >= 0 && c[i] < c[dn]) { q.step[i] = c[dn] - c[i]; c[i] = c[dn]; }
          cand.push_back({b, i, -q.step[i]});
          cp.insert(cp.end(), c, c + 16);
        }
        c[i] = v;
      }
    }
    if (cand.empty()) break;
    const int nc = (int)cand.size();
    F.resize((size_t)nc * 3); ma.resize((size_t)nc);
    if ((rc = vowel_formants(ctx, opts, nc, cp.data(), sampling_rate_hz, spectrum_length, frames, F.data(), ma.data()))) return rc;
    for (int k = 0; k < nc;) {
      const int b = cand[k].problem;
      OptProblem &q = pr[b];
      const double *t = target_hz + 3 * (size_t)b;
      const double curr = formant_error(&F[3 * (size_t)k], t);
      double best = curr, best_change = 0.0;
      int best_param = -1;
      for (k++; k < nc && cand[k].problem == b; k++) {
        if (ma[k] >= min_advised_area_cm2) {
          const double e = formant_error(&F[3 * (size_t)k], t);
          if (e < best) { best = e; best_param = cand[k].param; best_change = cand[k].change; }
        }
      }
      bool changed = false;
      if (best_param != -1 && best < curr) {
        q.cur[best_param] += best_change;
        if (best_change > 0.0) q.taken[best_param]++; else q.taken[best_param]--;
        changed = true;
      }
      if (trace) {
        trace[((size_t)b * max_runs + run) * 2] = changed ? (double)best_param : -1.0;
        trace[((size_t)b * max_runs + run) * 2 + 1] = best;
      }
      q.runs = run + 1;
      q.active = changed;
    }
  }
  // final formants and error (Data.cpp:476-482)
  cp.clear();
  for (int b = 0; b < n_problems; b++) cp.insert(cp.end(), pr[b].cur, pr[b].cur + 16);
  F.resize((size_t)n_problems * 3); ma.resize((size_t)n_problems);
  if ((rc = vowel_formants(ctx, opts, n_problems, cp.data(), sampling_rate_hz, spectrum_length, frames, F.data(), ma.data()))) return rc;
  for (int b = 0; b < n_problems; b++) {
    memcpy(params + 16 * (size_t)b, pr[b].cur, sizeof pr[b].cur);
    if (final_error) final_error[b] = formant_error(&F[3 * (size_t)b], target_hz + 3 * (size_t)b);
    if (runs) runs[b] = pr[b].runs;
    if (formants_hz) memcpy(formants_hz + 3 * (size_t)b, &F[3 * (size_t)b], 3 * sizeof(double));
  }
  return AFS_OK;
}

// ---------------------------------------------------------------------------
// LF vowel synthesis: Synthesizer::synthesizeVowelLf, batched
// ---------------------------------------------------------------------------
void afs_lf_pulse_default(afs_lf_pulse *p) {  // LfPulse::resetParams (LfPulse.cpp:24-31)
  if (!p) return;
  p->amp_cm3_s = 300.0; p->f0_hz = 120.0; p->oq = 0.5; p->sq = 3.0; p->tl = 0.02;
}

int afs_lf_vowel_length(int is_long_vowel, double sampling_rate_hz) { return lf_vowel_length(is_long_vowel, sampling_rate_hz); }

int afs_synthesize_vowels_lf(afs_ctx *ctx, const afs_tl_options *opts, int n, const afs_tube_frame *tubes,
                             const afs_lf_pulse *pulses, int is_long_vowel, double sampling_rate_hz, int16_t *audio,
                             double *pressure) {
  int rc = bind(ctx);
  if (rc) return rc;
  if (n < 0 || !(sampling_rate_hz > 16000.0) || !opts) return AFS_ERR_BAD_ARGUMENT;  // 8 kHz low-pass below Nyquist
  if (n == 0) return AFS_OK;
  if (!tubes || !pulses || !audio) return AFS_ERR_BAD_ARGUMENT;
  const int length = lf_vowel_length(is_long_vowel, sampling_rate_hz);
  if (length < 1) return AFS_ERR_BAD_ARGUMENT;
  // host: pulse trains, filter coefficients, scale factors
  std::vector<double> flow((size_t)n * length), scale((size_t)n), ab(14, 0.0);
  for (int v = 0; v < n; v++) {
    if (!(pulses[v].f0_hz > 0.0) || !(pulses[v].oq > 0.0) || !(pulses[v].oq < 1.0) || !(pulses[v].sq > 0.0)) return AFS_ERR_BAD_ARGUMENT;
    if (lf_vowel_flow(&pulses[v], is_long_vowel, sampling_rate_hz, flow.data() + (size_t)v * length)) return AFS_ERR_BAD_ARGUMENT;
    scale[v] = 80.0 / pulses[v].f0_hz;
  }
  {
    double a[33], b[33];
    chebyshev_lowpass(8000.0 / sampling_rate_hz, 6, a, b);  // SYNTHETIC_SPEECH_BANDWIDTH_HZ (Constants.h:28)
    for (int i = 0; i < 7; i++) { ab[i] = a[i]; ab[7 + i] = b[i]; }
  }
  void *d_tubes, *d_h, *d_flow, *d_pr, *d_audio, *d_misc;
  const size_t lb = (size_t)n * length;
  if ((rc = scratch(ctx, 0, (size_t)n * sizeof(afs_tube_frame), &d_tubes))) return rc;
  if ((rc = scratch(ctx, 3, (size_t)n * 512 * sizeof(double), &d_h))) return rc;
  if ((rc = scratch(ctx, 4, lb * sizeof(double), &d_flow))) return rc;
  if ((rc = scratch(ctx, 5, lb * sizeof(double), &d_pr))) return rc;
  if ((rc = scratch(ctx, 6, lb * sizeof(int16_t), &d_audio))) return rc;
  CU(cudaMemcpyAsync(d_tubes, tubes, (size_t)n * sizeof(afs_tube_frame), cudaMemcpyHostToDevice, ctx->stream));
  rc = afs_tl_impulse_response_device(ctx, opts, n, (const afs_tube_frame *)d_tubes, sampling_rate_hz, 9, (double *)d_h,
                                      ctx->stream);
  if (rc) return rc;
  if ((rc = scratch(ctx, 7, ((size_t)n + 14) * sizeof(double), &d_misc))) return rc;
  double *d_scale = (double *)d_misc, *d_ab = d_scale + n;
  CU(cudaMemcpyAsync(d_flow, flow.data(), lb * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(d_scale, scale.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(d_ab, ab.data(), 14 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CU(launch_lf_convolve(n, length, (const double *)d_h, d_scale, (const double *)d_flow, (double *)d_pr, ctx->stream));
  CU(launch_lf_filter(n, length, d_ab, (const double *)d_pr, (int16_t *)d_audio, ctx->stream));
  ctx->launches += 2;
  CU(cudaMemcpyAsync(audio, d_audio, lb * sizeof(int16_t), cudaMemcpyDeviceToHost, ctx->stream));
  if (pressure) CU(cudaMemcpyAsync(pressure, d_pr, lb * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return AFS_OK;
}

// ---------------------------------------------------------------------------
// TDS
// ---------------------------------------------------------------------------
}  // extern "C"

namespace {

// One call of the time-domain path on device buffers, asynchronous on `stream`.  `d_progress` (nullable, with
// progress_every > 0): the kernel publishes per-CTA progress marks there; only possible when the batch runs as ONE
// launch (*single_launch tells).
int tds_run_device(afs_ctx *ctx, const afs_tds_options *opts, int n_tracts, int n_frames, const int32_t *frame_len,
                   const afs_tube_frame *d_tubes, const double *d_glottis, double fs_tds_hz, int sr_factor,
                   const uint32_t *d_seeds, double *d_audio, double *d_flow, uint32_t *d_status,
                   const afs_tds_probes *d_probes, cudaStream_t stream, unsigned int *d_progress, int progress_every,
                   int *ctas_out /* [2]: CTAs of the single launch (0: several launches), tracts per CTA */,
                   const double *d_state_in = nullptr, double *d_state_out = nullptr /* chunked synthesis, see afs_tds_stream */) {
  int rc;
  if (!opts || n_tracts < 0 || n_frames < 2 || !frame_len || sr_factor < 1 || !(fs_tds_hz > 0))
    return AFS_ERR_BAD_ARGUMENT;
  if (opts->glottis_loss_option < AFS_STANDARD_ENTRANCE_LOSS || opts->glottis_loss_option > AFS_VARIABLE_ENTRANCE_LOSS)
    return AFS_ERR_BAD_ARGUMENT;
  for (int k = 0; k + 1 < n_frames; k++) if (frame_len[k] < 1) return AFS_ERR_BAD_ARGUMENT;
  if (d_probes) {
    if (d_probes->n_probes < 0 || d_probes->n_probes > AFS_MAX_TDS_PROBES) return AFS_ERR_BAD_ARGUMENT;
    for (int q = 0; q < d_probes->n_probes; q++)
      if (d_probes->section[q] < 0 || d_probes->section[q] >= AFS_NUM_SECTIONS) return AFS_ERR_BAD_ARGUMENT;
  }
  if (ctas_out) *ctas_out = 0;
  if (n_tracts == 0) return AFS_OK;
  if (!d_tubes || !d_glottis || !d_audio) return AFS_ERR_BAD_ARGUMENT;

  TdsLaunch L;
  rc = tds_prepare_launch(opts, n_tracts, n_frames, frame_len, fs_tds_hz, sr_factor,
                          ctx->dbg_tds_spread_sms, &L);
  if (rc) return rc;
  // everything below the caller's buffers is a stream-ordered temporary (see afs_ctx::pool)
  void *d_bounds = nullptr, *d_dump = nullptr, *d_nzc = nullptr, *d_runs = nullptr, *d_off = nullptr, *d_items = nullptr;
  auto release = [&]() {
    pool_free(d_bounds, stream); pool_free(d_dump, stream); pool_free(d_nzc, stream);
    pool_free(d_runs, stream); pool_free(d_off, stream); pool_free(d_items, stream);
  };
  if ((rc = pool_alloc(ctx, sizeof(int32_t) * (size_t)(n_frames + 1), stream, &d_bounds))) return rc;
  // (pageable source: staged before the call returns)
  cudaError_t e = cudaMemcpyAsync(d_bounds, L.step_bounds.data(), sizeof(int32_t) * (size_t)n_frames, cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) { release(); return cuda_fail(ctx, e, "cudaMemcpyAsync(step bounds)"); }
  // Small batches: tds_batch_kernel advances up to 28 tracts per SM in lock step and takes the same time for one; a
  // batch that leaves most of its lanes empty runs on tds_small_kernel (one tract per CTA, one tube section per lane).
  {
    const int small_max = ctx->dbg_small_batch_max >= 0 ? ctx->dbg_small_batch_max : 3 * ctx->sm_count;  // (three resident CTAs per SM)
    if (n_tracts <= small_max && tds_small_supported(L) && !(d_probes && d_probes->n_probes > 0) &&
        !(ctx->dump_step > 0 && ctx->dump_host) && !d_state_in && !d_state_out) {
      e = launch_tds_small(L, (const int32_t *)d_bounds, d_tubes, d_glottis, d_audio, d_flow, d_status, stream);
      release();
      if (e != cudaSuccess) return cuda_fail(ctx, e, "launch_tds_small");
      ctx->launches++;
      return AFS_OK;
    }
  }
  if (ctx->dump_step > 0 && ctx->dump_host) {
    if ((rc = pool_alloc(ctx, 473 * sizeof(double), stream, &d_dump))) { release(); return rc; }
    cudaMemsetAsync(d_dump, 0, 473 * sizeof(double), stream);
  }
  // Noise sources, control frames far apart: one record per run of steps with the same constrictions.  Two passes of
  // the bisection kernel -- count, then write at the offsets derived from the counts (the one host round trip of
  // the call: 4 B per tract each way).
  // Noise sources, control frames far apart: one record per run of steps with the same constrictions, found by
  // bisection (tds_noise_runs_kernel), kept as linked lists in one pool.  The pool is sized from the number of
  // pieces (every piece starts with a record) times run_pool_factor; a pool that turns out too small is reported
  // as status bit 1 of every tract (the host-buffer entry point then repeats the call with a larger one).
  int run_items = 0;
  if (opts->generate_noise_sources && tds_noise_runs_mode(L)) {
    std::vector<int32_t> items, iv_first;
    tds_noise_run_items(L, &items, &iv_first);
    run_items = (int)(items.size() / 4);
    items.insert(items.end(), iv_first.begin(), iv_first.end());  // one upload: the pieces, then the per-interval table
    const size_t n_slots = (size_t)n_tracts * (size_t)run_items;
    size_t pool = n_slots * (size_t)ctx->run_pool_factor + 4096;
    if (pool > 0xfffffff0ull) pool = 0xfffffff0ull;
    if (n_slots >= pool) { release(); return AFS_ERR_OUT_OF_MEMORY; }
    if ((rc = pool_alloc(ctx, sizeof(int32_t) * items.size(), stream, &d_items))) { release(); return rc; }
    if ((rc = pool_alloc(ctx, 2 * sizeof(uint32_t), stream, &d_off))) { release(); return rc; }
    if ((rc = pool_alloc(ctx, tds_noise_run_bytes() * pool, stream, &d_runs))) { release(); return rc; }
    const uint32_t ctl0[2] = {(uint32_t)n_slots, 0u};
    e = cudaMemcpyAsync(d_items, items.data(), sizeof(int32_t) * items.size(), cudaMemcpyHostToDevice, stream);  // (pageable: staged)
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_off, ctl0, sizeof ctl0, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess)
      e = launch_noise_runs(L, run_items, d_items, (const int32_t *)d_items + (size_t)run_items * 4, (const int32_t *)d_bounds,
                            d_tubes, d_runs, (uint32_t)pool, (uint32_t *)d_off, stream);
    if (e != cudaSuccess) { release(); return cuda_fail(ctx, e, "noise pre-pass"); }
    ctx->launches++;
  }
  // Noise sources, control frames a few steps apart: one record per tract-step (48 B, 32 lanes per CTA); a batch whose
  // records would not fit the budget runs as several waves of CTAs that reuse the buffer.
  const size_t per_cta = tds_noise_scan_bytes_per_cta(L);
  const int lpc = tds_lanes_per_cta(L);
  const int ctas = (n_tracts + lpc - 1) / lpc;
  int wave_ctas = ctas;
  if (per_cta) {
    size_t free_b = 0, total_b = 0;
    unsigned long long cached = 0;
    cudaMemGetInfo(&free_b, &total_b);
    cudaMemPoolGetAttribute(ctx->pool, cudaMemPoolAttrReservedMemCurrent, &cached);
    size_t budget = (size_t)((free_b + (size_t)cached) / 10 * 6);  // 60 % of what this context could get
    if (budget > ((size_t)12 << 30)) budget = (size_t)12 << 30;
    if (ctx->dbg_noise_scan_bytes > 0) budget = (size_t)ctx->dbg_noise_scan_bytes;  // (tests force several waves)
    if (per_cta > budget) { release(); return AFS_ERR_OUT_OF_MEMORY; }  // a single utterance of hours
    size_t fit = budget / per_cta;
    if (fit >= (size_t)ctx->sm_count) fit -= fit % (size_t)ctx->sm_count;  // whole waves of one CTA per SM
    if ((size_t)wave_ctas > fit) wave_ctas = (int)fit;
    if ((rc = pool_alloc(ctx, per_cta * (size_t)wave_ctas, stream, &d_nzc))) { release(); return rc; }
  }
  const bool single = wave_ctas >= ctas && ctas <= kMaxProgressCtas;  // (one launch, and a mark per CTA fits)
  if (ctas_out) *ctas_out = single ? ctas : 0;
  if (ctas_out && single) ctas_out[1] = lpc;
  for (int c0 = 0; c0 < ctas; c0 += wave_ctas) {
    const int t0 = c0 * lpc;
    const int nw = (c0 + wave_ctas >= ctas) ? n_tracts - t0 : wave_ctas * lpc;
    e = launch_tds(L, t0, nw, (const int32_t *)d_bounds, d_tubes, d_glottis, d_seeds, d_audio, d_flow, d_status,
                   c0 == 0 ? (double *)d_dump : nullptr, ctx->dump_step, d_nzc, d_runs, (const uint32_t *)d_off, run_items, d_probes,
                   single ? d_progress : nullptr, single ? progress_every : 0, d_state_in, d_state_out, stream);
    if (e != cudaSuccess) { release(); return cuda_fail(ctx, e, "launch_tds"); }
    ctx->launches += per_cta ? 2 : 1;
  }
  if (d_dump) {
    cudaMemcpyAsync(ctx->dump_host, d_dump, 473 * sizeof(double), cudaMemcpyDeviceToHost, stream);
    cudaStreamSynchronize(stream);
  }
  release();
  return AFS_OK;
}

}  // namespace

extern "C" {

int afs_tds_synthesize_device(afs_ctx *ctx, const afs_tds_options *opts, int n_tracts, int n_frames,
                              const int32_t *frame_len, const afs_tube_frame *d_tubes, const double *d_glottis,
                              double fs_tds_hz, int sr_factor, const uint32_t *d_seeds, double *d_audio,
                              double *d_flow, uint32_t *d_status, void *stream) {
  return afs_tds_synthesize_probed_device(ctx, opts, n_tracts, n_frames, frame_len, d_tubes, d_glottis, fs_tds_hz, sr_factor,
                                          d_seeds, d_audio, d_flow, d_status, nullptr, stream);
}

int afs_tds_synthesize_probed_device(afs_ctx *ctx, const afs_tds_options *opts, int n_tracts, int n_frames,
                                     const int32_t *frame_len, const afs_tube_frame *d_tubes, const double *d_glottis,
                                     double fs_tds_hz, int sr_factor, const uint32_t *d_seeds, double *d_audio,
                                     double *d_flow, uint32_t *d_status, const afs_tds_probes *d_probes, void *stream) {
  int rc = bind(ctx);
  if (rc) return rc;
  return tds_run_device(ctx, opts, n_tracts, n_frames, frame_len, d_tubes, d_glottis, fs_tds_hz, sr_factor, d_seeds, d_audio,
                        d_flow, d_status, d_probes, (cudaStream_t)stream, nullptr, 0, nullptr);
}

int afs_tds_set_state_dump(afs_ctx *ctx, int step, double *host_out473) {
  if (!ctx) return AFS_ERR_BAD_CONTEXT;
  ctx->dump_step = step;
  ctx->dump_host = host_out473;
  return AFS_OK;
}

int afs_debug_set(afs_ctx *ctx, const char *key, int64_t value) {
  if (!ctx) return AFS_ERR_BAD_CONTEXT;
  if (!key) return AFS_ERR_BAD_ARGUMENT;
  if (!strcmp(key, "tds.noise_scan_bytes")) ctx->dbg_noise_scan_bytes = value;
  else if (!strcmp(key, "tds.d2h_slices")) ctx->dbg_d2h_slices = (int)value;
  else if (!strcmp(key, "tds.small_batch_max")) ctx->dbg_small_batch_max = (int)value;
  else if (!strcmp(key, "tds.spread_sms")) ctx->dbg_tds_spread_sms = (int)value;
  else if (!strcmp(key, "tds.noise_run_pool_factor")) ctx->run_pool_factor = value < 2 ? 2 : value > 4096 ? 4096 : (int)value;
  else return AFS_ERR_BAD_ARGUMENT;
  return AFS_OK;
}

int afs_tds_synthesize(afs_ctx *ctx, const afs_tds_options *opts, int n_tracts, int n_frames,
                       const int32_t *frame_len, const afs_tube_frame *tubes, const double *glottis,
                       double fs_tds_hz, int sr_factor, const uint32_t *seeds, double *audio, double *flow,
                       uint32_t *status) {
  return afs_tds_synthesize_probed(ctx, opts, n_tracts, n_frames, frame_len, tubes, glottis, fs_tds_hz, sr_factor, seeds,
                                   audio, flow, status, nullptr);
}

// Host buffers.  The audio leaves the device WHILE the kernel runs: the kernel publishes, per CTA, how many
// simulation steps are complete (progress marks in mapped host memory, every 1/16 of the utterance); as soon as all
// CTAs have passed a mark, the finished time slice of every utterance -- a [n_tracts] x [slice] column block of the
// [n_tracts][n_out] array -- is copied on the second stream.  What is left after the kernel is the last slice.
int afs_tds_synthesize_probed(afs_ctx *ctx, const afs_tds_options *opts, int n_tracts, int n_frames,
                              const int32_t *frame_len, const afs_tube_frame *tubes, const double *glottis,
                              double fs_tds_hz, int sr_factor, const uint32_t *seeds, double *audio, double *flow,
                              uint32_t *status, const afs_tds_probes *probes) {
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
  // taps: device copies of whichever of the three outputs the caller asked for, back to back in scratch slot 9
  afs_tds_probes dp;
  const afs_tds_probes *use = nullptr;
  size_t pb = 0;
  if (probes && probes->n_probes > 0) {
    if (probes->n_probes > AFS_MAX_TDS_PROBES) return AFS_ERR_BAD_ARGUMENT;
    dp = *probes;
    pb = (size_t)n_tracts * (size_t)probes->n_probes * n_out * (size_t)sr_factor * sizeof(double);
    const int n_arr = (probes->pressure_dPa ? 1 : 0) + (probes->inflow_cm3_s ? 1 : 0) + (probes->outflow_cm3_s ? 1 : 0);
    void *d_p = nullptr;
    if (n_arr && (rc = scratch(ctx, 9, pb * n_arr, &d_p))) return rc;
    unsigned char *at = (unsigned char *)d_p;
    dp.pressure_dPa = dp.inflow_cm3_s = dp.outflow_cm3_s = nullptr;
    if (probes->pressure_dPa) { dp.pressure_dPa = (double *)at; at += pb; }
    if (probes->inflow_cm3_s) { dp.inflow_cm3_s = (double *)at; at += pb; }
    if (probes->outflow_cm3_s) { dp.outflow_cm3_s = (double *)at; at += pb; }
    use = &dp;
  }
  // time slices: worth it from a few MB of audio on; the progress marks are in simulation steps
  const size_t total_steps = n_out * (size_t)sr_factor;
  int n_slices = ab >= ((size_t)8 << 20) && n_out >= 64 ? 16 : 1;
  if (ctx->dbg_d2h_slices > 0) n_slices = ctx->dbg_d2h_slices;
  const size_t slice_out = (n_out + (size_t)n_slices - 1) / (size_t)n_slices;
  const int every = n_slices > 1 && total_steps < 0x7fffffffull ? (int)(slice_out * (size_t)sr_factor) : 0;
  if (every) memset(ctx->progress_h, 0, kMaxProgressCtas * sizeof(unsigned int));
  int grid[2] = {0, 0};
  rc = tds_run_device(ctx, opts, n_tracts, n_frames, frame_len, (const afs_tube_frame *)d_tubes, (const double *)d_gl,
                      fs_tds_hz, sr_factor, (const uint32_t *)d_seeds, (double *)d_audio, (double *)d_flow,
                      (uint32_t *)d_status, use, ctx->stream, every ? ctx->progress_d : nullptr, every, grid);
  if (rc) { cudaStreamSynchronize(ctx->stream); return rc; }
  const int ctas = grid[0], lpc = grid[1];
  const volatile unsigned int *mark = ctx->progress_h;
  // waits until the CTAs [c0, c1) have all published `need` steps; false when the kernel is over (or failed) first
  auto wait_for = [&](int c0, int c1, unsigned int need) {
    for (int spin = 0;; spin++) {
      bool ready = true;
      for (int c = c0; c < c1; c++) if (mark[c] < need) { ready = false; break; }
      if (ready) return true;
      if ((spin & 1023) == 1023 && cudaStreamQuery(ctx->stream) != cudaErrorNotReady) return false;
    }
  };
  size_t copied = 0;       // output samples per utterance already on their way to the host (time slices) ...
  size_t rows_copied = 0;  // ... or whole utterances (batches of more CTAs than SMs: the CTAs finish wave after wave)
  cudaError_t ce = cudaSuccess;
  if (every && ctas > 0 && ctas <= ctx->sm_count) {
    while (ce == cudaSuccess && copied + slice_out < n_out) {  // (the last slice goes after the kernel)
      if (!wait_for(0, ctas, (unsigned int)((copied + slice_out) * (size_t)sr_factor))) break;
      ce = cudaMemcpy2DAsync(audio + copied, n_out * sizeof(double), (const double *)d_audio + copied, n_out * sizeof(double),
                             slice_out * sizeof(double), (size_t)n_tracts, cudaMemcpyDeviceToHost, ctx->copy_stream);
      copied += slice_out;
    }
  } else if (every && ctas > ctx->sm_count) {
    for (int c0 = 0; ce == cudaSuccess && c0 + ctx->sm_count < ctas; c0 += ctx->sm_count) {  // (the last wave goes after the kernel)
      if (!wait_for(c0, c0 + ctx->sm_count, (unsigned int)total_steps)) break;
      const size_t r1 = (size_t)(c0 + ctx->sm_count) * (size_t)lpc;
      ce = cudaMemcpyAsync(audio + rows_copied * n_out, (const double *)d_audio + rows_copied * n_out,
                           (r1 - rows_copied) * n_out * sizeof(double), cudaMemcpyDeviceToHost, ctx->copy_stream);
      rows_copied = r1;
    }
  }
  if (ce != cudaSuccess) { cudaStreamSynchronize(ctx->stream); cudaStreamSynchronize(ctx->copy_stream); return cuda_fail(ctx, ce, "device-to-host copy of finished audio"); }
  cudaError_t e = cudaSuccess;
  if (copied == 0) e = cudaMemcpyAsync(audio + rows_copied * n_out, (const double *)d_audio + rows_copied * n_out,
                                       ((size_t)n_tracts - rows_copied) * n_out * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  else e = cudaMemcpy2DAsync(audio + copied, n_out * sizeof(double), (const double *)d_audio + copied, n_out * sizeof(double),
                             (n_out - copied) * sizeof(double), (size_t)n_tracts, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess && flow) e = cudaMemcpyAsync(flow, d_flow, fb, cudaMemcpyDeviceToHost, ctx->stream);
  std::vector<uint32_t> status_tmp;
  uint32_t *status_h = status;
  if (!status_h && opts->generate_noise_sources) { status_tmp.resize((size_t)n_tracts); status_h = status_tmp.data(); }
  if (e == cudaSuccess && status_h) e = cudaMemcpyAsync(status_h, d_status, (size_t)n_tracts * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess && use) {
    if (dp.pressure_dPa) e = cudaMemcpyAsync(probes->pressure_dPa, dp.pressure_dPa, pb, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && dp.inflow_cm3_s) e = cudaMemcpyAsync(probes->inflow_cm3_s, dp.inflow_cm3_s, pb, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && dp.outflow_cm3_s) e = cudaMemcpyAsync(probes->outflow_cm3_s, dp.outflow_cm3_s, pb, cudaMemcpyDeviceToHost, ctx->stream);
  }
  // both streams are drained on every path: the caller owns the host buffers again when this returns
  const cudaError_t e1 = cudaStreamSynchronize(ctx->copy_stream), e2 = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "device-to-host copies");
  if (e1 != cudaSuccess) return cuda_fail(ctx, e1, "cudaStreamSynchronize(copy stream)");
  if (e2 != cudaSuccess) return cuda_fail(ctx, e2, "cudaStreamSynchronize");
  // status bit 1: the record pool of the noise pre-pass was too small (tds_run_device) -- once more with a larger one
  if (status_h && opts->generate_noise_sources && (status_h[0] & 2u) && ctx->run_pool_factor < 4096) {
    ctx->run_pool_factor *= 8;
    return afs_tds_synthesize_probed(ctx, opts, n_tracts, n_frames, frame_len, tubes, glottis, fs_tds_hz, sr_factor, seeds,
                                     audio, flow, status, probes);
  }
  return AFS_OK;
}

}  // extern "C"

// ---- chunked synthesis: afs_tds_stream -------------------------------------------------------------------------------
// A batch of TdsModel + TriangularGlottis + driver objects that live across calls, like the reference's Synthesizer
// (one synthesizeSignalTds call per control frame, Synthesizer.cpp:515-639).  The recurrent state of every tract stays
// in device memory between calls (two buffers: a chunk reads one and writes the other, so a failed chunk leaves the
// stream where it was), next to the last control frame of every tract (prevTube / prevGlottisParams).
struct afs_tds_stream {
  afs_ctx *ctx = nullptr;
  afs_tds_options opts;
  int n = 0, sr_factor = 1;
  double fs = 0.0;
  bool primed = false;     // the first frame has been stored (Synthesizer.cpp:522-532)
  bool has_state = false;  // at least one chunk has run
  int cur = 0;             // which state buffer is current
  int64_t samples = 0;     // output samples produced so far (per tract)
  double *d_state[2] = {nullptr, nullptr};
  afs_tube_frame *d_prev_tubes = nullptr;
  double *d_prev_gl = nullptr;
  uint32_t *d_seeds = nullptr;
};

namespace {
constexpr uint64_t kStreamMagic = 0x3153445453464100ull;  // "\0AFSTDS1"
struct StreamHeader {
  uint64_t magic;
  int32_t n, sr_factor, primed, has_state;
  int64_t samples;
  uint64_t state_words;
  double fs;
  afs_tds_options opts;
};
size_t stream_state_bytes(const afs_tds_stream *s) { return (size_t)s->n * tds_state_words() * sizeof(double); }
size_t stream_blob_bytes(const afs_tds_stream *s) {
  return sizeof(StreamHeader) + (size_t)s->n * (sizeof(afs_tube_frame) + AFS_NUM_GLOTTIS_CONTROL * sizeof(double)) + stream_state_bytes(s);
}
// Device-to-host copy of a chunk's audio WHILE its kernel runs (the scheme of afs_tds_synthesize_probed: per-CTA progress
// marks in mapped host memory, finished time slices / CTA waves copied on the second stream).  Returns what is already
// on its way: output samples per tract (*copied) or whole tracts (*rows_copied); the caller copies the rest behind the kernel.
cudaError_t copy_finished_audio(afs_ctx *ctx, double *audio, const double *d_audio, size_t n_tracts, size_t n_out, int sr_factor,
                                int every, size_t slice_out, int ctas, int lpc, size_t *copied, size_t *rows_copied) {
  *copied = 0; *rows_copied = 0;
  const volatile unsigned int *mark = ctx->progress_h;
  auto wait_for = [&](int c0, int c1, unsigned int need) {
    for (int spin = 0;; spin++) {
      bool ready = true;
      for (int c = c0; c < c1; c++) if (mark[c] < need) { ready = false; break; }
      if (ready) return true;
      if ((spin & 1023) == 1023 && cudaStreamQuery(ctx->stream) != cudaErrorNotReady) return false;
    }
  };
  cudaError_t ce = cudaSuccess;
  const size_t total_steps = n_out * (size_t)sr_factor;
  if (every && ctas > 0 && ctas <= ctx->sm_count) {
    while (ce == cudaSuccess && *copied + slice_out < n_out) {
      if (!wait_for(0, ctas, (unsigned int)((*copied + slice_out) * (size_t)sr_factor))) break;
      ce = cudaMemcpy2DAsync(audio + *copied, n_out * sizeof(double), d_audio + *copied, n_out * sizeof(double),
                             slice_out * sizeof(double), n_tracts, cudaMemcpyDeviceToHost, ctx->copy_stream);
      *copied += slice_out;
    }
  } else if (every && ctas > ctx->sm_count) {
    for (int c0 = 0; ce == cudaSuccess && c0 + ctx->sm_count < ctas; c0 += ctx->sm_count) {
      if (!wait_for(c0, c0 + ctx->sm_count, (unsigned int)total_steps)) break;
      const size_t r1 = (size_t)(c0 + ctx->sm_count) * (size_t)lpc;
      ce = cudaMemcpyAsync(audio + *rows_copied * n_out, d_audio + *rows_copied * n_out, (r1 - *rows_copied) * n_out * sizeof(double),
                           cudaMemcpyDeviceToHost, ctx->copy_stream);
      *rows_copied = r1;
    }
  }
  return ce;
}

void stream_release(afs_tds_stream *s) {
  cudaFree(s->d_state[0]); cudaFree(s->d_state[1]); cudaFree(s->d_prev_tubes); cudaFree(s->d_prev_gl); cudaFree(s->d_seeds);
  delete s;
}
}  // namespace

extern "C" {

int afs_tds_stream_create(afs_ctx *ctx, const afs_tds_options *opts, int n_tracts, double fs_tds_hz, int sr_factor,
                          const uint32_t *seeds, afs_tds_stream **out) {
  int rc = bind(ctx);
  if (rc) return rc;
  if (!opts || !out || n_tracts < 1 || sr_factor < 1 || !(fs_tds_hz > 0)) return AFS_ERR_BAD_ARGUMENT;
  if (opts->glottis_loss_option < AFS_STANDARD_ENTRANCE_LOSS || opts->glottis_loss_option > AFS_VARIABLE_ENTRANCE_LOSS ||
      opts->reserved != 0 || opts->flow_source_section < -1 || opts->flow_source_section >= AFS_NUM_SECTIONS)
    return AFS_ERR_BAD_ARGUMENT;
  afs_tds_stream *s = new (std::nothrow) afs_tds_stream;
  if (!s) return AFS_ERR_OUT_OF_MEMORY;
  s->ctx = ctx; s->opts = *opts; s->n = n_tracts; s->sr_factor = sr_factor; s->fs = fs_tds_hz;
  cudaError_t e = cudaMalloc((void **)&s->d_state[0], stream_state_bytes(s));
  if (e == cudaSuccess) e = cudaMalloc((void **)&s->d_state[1], stream_state_bytes(s));
  if (e == cudaSuccess) e = cudaMalloc((void **)&s->d_prev_tubes, (size_t)n_tracts * sizeof(afs_tube_frame));
  if (e == cudaSuccess) e = cudaMalloc((void **)&s->d_prev_gl, (size_t)n_tracts * AFS_NUM_GLOTTIS_CONTROL * sizeof(double));
  if (e == cudaSuccess && seeds) {
    e = cudaMalloc((void **)&s->d_seeds, (size_t)n_tracts * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemcpy(s->d_seeds, seeds, (size_t)n_tracts * sizeof(uint32_t), cudaMemcpyHostToDevice);
  }
  if (e != cudaSuccess) { stream_release(s); return cuda_fail(ctx, e, "afs_tds_stream_create"); }
  *out = s;
  return AFS_OK;
}

int afs_tds_stream_destroy(afs_tds_stream *s) {
  if (!s) return AFS_OK;
  int rc = bind(s->ctx);
  if (rc) return rc;
  cudaStreamSynchronize(s->ctx->stream);
  stream_release(s);
  return AFS_OK;
}

int afs_tds_stream_reset(afs_tds_stream *s) {
  if (!s) return AFS_ERR_BAD_ARGUMENT;
  s->primed = false; s->has_state = false; s->samples = 0;
  return AFS_OK;
}

int64_t afs_tds_stream_samples(const afs_tds_stream *s) { return s ? s->samples : -1; }
int afs_tds_stream_primed(const afs_tds_stream *s) { return s ? (s->primed ? 1 : 0) : -1; }

int afs_tds_stream_synthesize(afs_tds_stream *s, int n_frames, const int32_t *frame_len, const afs_tube_frame *tubes,
                              const double *glottis, double *audio, double *flow, uint32_t *status) {
  if (!s) return AFS_ERR_BAD_ARGUMENT;
  afs_ctx *ctx = s->ctx;
  int rc = bind(ctx);
  if (rc) return rc;
  if (n_frames < 1 || !frame_len || !tubes || !glottis) return AFS_ERR_BAD_ARGUMENT;
  const int k0 = s->primed ? 0 : 1;  // frames that only prime
  const int m = n_frames - k0;       // frames that produce samples
  size_t n_out = 0;
  for (int k = k0; k < n_frames; k++) { if (frame_len[k] < 1) return AFS_ERR_BAD_ARGUMENT; n_out += (size_t)frame_len[k]; }
  if (m > 0 && !audio) return AFS_ERR_BAD_ARGUMENT;
  const size_t n = (size_t)s->n, fsz = sizeof(afs_tube_frame), gsz = AFS_NUM_GLOTTIS_CONTROL * sizeof(double);
  // every error path drains both streams before it returns: the caller owns its buffers again
  auto fail = [&](cudaError_t err, const char *what) {
    cudaStreamSynchronize(ctx->stream); cudaStreamSynchronize(ctx->copy_stream);
    return cuda_fail(ctx, err, what);
  };
  cudaError_t e = cudaSuccess;
  if (!s->primed) {  // prevTube = the first frame, no samples (Synthesizer.cpp:522-532)
    e = cudaMemcpy2DAsync(s->d_prev_tubes, fsz, tubes, (size_t)n_frames * fsz, fsz, n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpy2DAsync(s->d_prev_gl, gsz, glottis, (size_t)n_frames * gsz, gsz, n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return fail(e, "afs_tds_stream_synthesize (priming frame)");
    if (m == 0) s->primed = true;  // (with more frames in this call: once their chunk is in)
  }
  if (m == 0) return AFS_OK;
  // the chunk's frames on the device: [tract][prev, new 0 .. m-1]
  const size_t ab = n * n_out * sizeof(double), fb = ab * (size_t)s->sr_factor;
  void *d_tubes, *d_gl, *d_audio, *d_flow = nullptr, *d_status;
  if ((rc = scratch(ctx, 0, n * (size_t)(m + 1) * fsz, &d_tubes))) return rc;
  if ((rc = scratch(ctx, 1, n * (size_t)(m + 1) * gsz, &d_gl))) return rc;
  if ((rc = scratch(ctx, 2, ab, &d_audio))) return rc;
  if (flow && (rc = scratch(ctx, 3, fb, &d_flow))) return rc;
  if ((rc = scratch(ctx, 4, n * sizeof(uint32_t), &d_status))) return rc;
  const size_t dp_t = (size_t)(m + 1) * fsz, dp_g = (size_t)(m + 1) * gsz;
  e = cudaMemcpy2DAsync(d_tubes, dp_t, s->d_prev_tubes, fsz, fsz, n, cudaMemcpyDeviceToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpy2DAsync(d_gl, dp_g, s->d_prev_gl, gsz, gsz, n, cudaMemcpyDeviceToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpy2DAsync((unsigned char *)d_tubes + fsz, dp_t, tubes + k0, (size_t)n_frames * fsz, (size_t)m * fsz, n,
                                              cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpy2DAsync((unsigned char *)d_gl + gsz, dp_g, glottis + (size_t)k0 * AFS_NUM_GLOTTIS_CONTROL,
                                              (size_t)n_frames * gsz, (size_t)m * gsz, n, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) return fail(e, "afs_tds_stream_synthesize (control frames)");
  std::vector<uint32_t> st((size_t)s->n);
  // finished time slices of the audio leave the device while the kernel runs (as in afs_tds_synthesize)
  const size_t total_steps = n_out * (size_t)s->sr_factor;
  int n_slices = ab >= ((size_t)8 << 20) && n_out >= 64 ? 16 : 1;
  if (ctx->dbg_d2h_slices > 0) n_slices = ctx->dbg_d2h_slices;
  const size_t slice_out = (n_out + (size_t)n_slices - 1) / (size_t)n_slices;
  const int every = n_slices > 1 && total_steps < 0x7fffffffull ? (int)(slice_out * (size_t)s->sr_factor) : 0;
  size_t copied = 0, rows_copied = 0;
  for (;;) {
    if (every) memset(ctx->progress_h, 0, kMaxProgressCtas * sizeof(unsigned int));
    int grid[2] = {0, 0};
    rc = tds_run_device(ctx, &s->opts, s->n, m + 1, frame_len + k0, (const afs_tube_frame *)d_tubes, (const double *)d_gl, s->fs,
                        s->sr_factor, s->d_seeds, (double *)d_audio, (double *)d_flow, (uint32_t *)d_status, nullptr, ctx->stream,
                        every ? ctx->progress_d : nullptr, every, grid, s->has_state ? s->d_state[s->cur] : nullptr,
                        s->d_state[s->cur ^ 1]);
    if (rc) { cudaStreamSynchronize(ctx->stream); cudaStreamSynchronize(ctx->copy_stream); return rc; }
    e = copy_finished_audio(ctx, audio, (const double *)d_audio, n, n_out, s->sr_factor, every, slice_out, grid[0], grid[1],
                                        &copied, &rows_copied);
    if (e == cudaSuccess) e = cudaMemcpyAsync(st.data(), d_status, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
    const cudaError_t e1 = cudaStreamSynchronize(ctx->copy_stream), e2 = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = e1 != cudaSuccess ? e1 : e2;
    if (e != cudaSuccess) return cuda_fail(ctx, e, "afs_tds_stream_synthesize");
    // status bit 1: the record pool of the noise pre-pass was too small -- the chunk is repeated from the same state
    // (whatever was copied already is overwritten by the repeat)
    if (s->opts.generate_noise_sources && (st[0] & 2u) && ctx->run_pool_factor < 4096) { ctx->run_pool_factor *= 8; continue; }
    break;
  }
  // the chunk is in: its last frame is the next prevTube, the state buffers change roles
  if (copied == 0) e = cudaMemcpyAsync(audio + rows_copied * n_out, (const double *)d_audio + rows_copied * n_out,
                                       (n - rows_copied) * n_out * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  else e = cudaMemcpy2DAsync(audio + copied, n_out * sizeof(double), (const double *)d_audio + copied, n_out * sizeof(double),
                             (n_out - copied) * sizeof(double), n, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess && flow) e = cudaMemcpyAsync(flow, d_flow, fb, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return fail(e, "afs_tds_stream_synthesize (device-to-host copies)");
  // (the last frame becomes prevTube only now: a call that failed above leaves the stream as it was)
  e = cudaMemcpy2DAsync(s->d_prev_tubes, fsz, (unsigned char *)d_tubes + (size_t)m * fsz, dp_t, fsz, n, cudaMemcpyDeviceToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpy2DAsync(s->d_prev_gl, gsz, (unsigned char *)d_gl + (size_t)m * gsz, dp_g, gsz, n, cudaMemcpyDeviceToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return fail(e, "afs_tds_stream_synthesize (last control frame)");
  if (status) memcpy(status, st.data(), n * sizeof(uint32_t));
  s->cur ^= 1; s->has_state = true; s->primed = true; s->samples += (int64_t)n_out;
  return AFS_OK;
}

size_t afs_tds_stream_checkpoint_bytes(const afs_tds_stream *s) { return s ? stream_blob_bytes(s) : 0; }

int afs_tds_stream_save(afs_tds_stream *s, void *blob, size_t bytes) {
  if (!s || !blob || bytes < stream_blob_bytes(s)) return AFS_ERR_BAD_ARGUMENT;
  afs_ctx *ctx = s->ctx;
  int rc = bind(ctx);
  if (rc) return rc;
  StreamHeader h;
  memset(&h, 0, sizeof h);
  h.magic = kStreamMagic; h.n = s->n; h.sr_factor = s->sr_factor; h.primed = s->primed; h.has_state = s->has_state;
  h.samples = s->samples; h.state_words = tds_state_words(); h.fs = s->fs; h.opts = s->opts;
  unsigned char *at = (unsigned char *)blob;
  memcpy(at, &h, sizeof h); at += sizeof h;
  const size_t tb = (size_t)s->n * sizeof(afs_tube_frame), gb = (size_t)s->n * AFS_NUM_GLOTTIS_CONTROL * sizeof(double);
  memset(at, 0, tb + gb + stream_state_bytes(s));
  cudaError_t e = cudaSuccess;
  if (s->primed) {
    e = cudaMemcpyAsync(at, s->d_prev_tubes, tb, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(at + tb, s->d_prev_gl, gb, cudaMemcpyDeviceToHost, ctx->stream);
  }
  if (e == cudaSuccess && s->has_state)
    e = cudaMemcpyAsync(at + tb + gb, s->d_state[s->cur], stream_state_bytes(s), cudaMemcpyDeviceToHost, ctx->stream);
  const cudaError_t es = cudaStreamSynchronize(ctx->stream);  // (drained on every path)
  if (e == cudaSuccess) e = es;
  return e == cudaSuccess ? AFS_OK : cuda_fail(ctx, e, "afs_tds_stream_save");
}

int afs_tds_stream_load(afs_tds_stream *s, const void *blob, size_t bytes) {
  if (!s || !blob || bytes < stream_blob_bytes(s)) return AFS_ERR_BAD_ARGUMENT;
  afs_ctx *ctx = s->ctx;
  int rc = bind(ctx);
  if (rc) return rc;
  StreamHeader h;
  const unsigned char *at = (const unsigned char *)blob;
  memcpy(&h, at, sizeof h); at += sizeof h;
  // a checkpoint belongs to one batch size, rate, option set and build of the kernels
  if (h.magic != kStreamMagic || h.n != s->n || h.sr_factor != s->sr_factor || h.fs != s->fs ||
      h.state_words != tds_state_words() || memcmp(&h.opts, &s->opts, sizeof h.opts) != 0)
    return AFS_ERR_BAD_ARGUMENT;
  const size_t tb = (size_t)s->n * sizeof(afs_tube_frame), gb = (size_t)s->n * AFS_NUM_GLOTTIS_CONTROL * sizeof(double);
  cudaError_t e = cudaMemcpyAsync(s->d_prev_tubes, at, tb, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(s->d_prev_gl, at + tb, gb, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(s->d_state[s->cur], at + tb + gb, stream_state_bytes(s), cudaMemcpyHostToDevice, ctx->stream);
  const cudaError_t es = cudaStreamSynchronize(ctx->stream);  // (drained on every path)
  if (e == cudaSuccess) e = es;
  if (e != cudaSuccess) { s->primed = false; s->has_state = false; return cuda_fail(ctx, e, "afs_tds_stream_load"); }  // (half-loaded: a fresh stream)
  s->primed = h.primed != 0; s->has_state = h.has_state != 0; s->samples = h.samples;
  return AFS_OK;
}

}  // extern "C"
