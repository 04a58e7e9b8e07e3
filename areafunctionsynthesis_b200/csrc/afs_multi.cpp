// afs_multi.cpp -- several devices of one box behind the C ABI (afs_group_*, include/afs_b200.h).
//
// The reference runs one TdsModel / TlModel per unit and the units never talk to each other
// (Synthesizer.cpp:557-629, TlModel.cpp:123-197), so the multi-device layer is sharding and nothing else:
// contiguous shards, one host thread + context + stream per device, every device copying its rows straight into
// the caller's arrays.  No collective sits on the data path.
#include <cuda_runtime.h>

#include <cstring>
#include <new>
#include <thread>
#include <vector>

#include "../../include/afs_b200.h"

struct afs_group {
  std::vector<afs_ctx *> ctx;
};

namespace {

void shard_of(int n_units, int n_dev, int r, int *lo, int *hi) {
  const int base = n_units / n_dev, extra = n_units % n_dev;  // earlier devices take the extras
  *lo = r * base + (r < extra ? r : extra);
  *hi = *lo + base + (r < extra ? 1 : 0);
}

// fn(r, lo, hi) on one thread per device with a non-empty shard; the first failure is the result
template <class F>
int for_each_shard(afs_group *g, int n_units, F fn) {
  const int G = (int)g->ctx.size();
  std::vector<int> rc((size_t)G, AFS_OK);
  std::vector<std::thread> th;
  for (int r = 0; r < G; r++) {
    int lo, hi;
    shard_of(n_units, G, r, &lo, &hi);
    if (hi > lo) th.emplace_back([&rc, &fn, r, lo, hi] { rc[(size_t)r] = fn(r, lo, hi); });
  }
  for (auto &t : th) t.join();
  for (int r = 0; r < G; r++) if (rc[(size_t)r] != AFS_OK) return rc[(size_t)r];
  return AFS_OK;
}

}  // namespace

extern "C" {

int afs_group_create(afs_group **out, const int *device_ordinals, int n_devices) {
  if (!out) return AFS_ERR_BAD_ARGUMENT;
  *out = nullptr;
  int visible = 0;
  if (cudaGetDeviceCount(&visible) != cudaSuccess || visible <= 0) { cudaGetLastError(); return AFS_ERR_NO_DEVICE; }
  std::vector<int> ids;
  if (!device_ordinals || n_devices <= 0) for (int d = 0; d < visible; d++) ids.push_back(d);
  else ids.assign(device_ordinals, device_ordinals + n_devices);
  for (size_t i = 0; i < ids.size(); i++)
    for (size_t j = 0; j < i; j++) if (ids[i] == ids[j]) return AFS_ERR_BAD_ARGUMENT;
  afs_group *g = new (std::nothrow) afs_group();
  if (!g) return AFS_ERR_OUT_OF_MEMORY;
  for (int d : ids) {
    afs_ctx *c = nullptr;
    const int rc = afs_create(&c, d);
    if (rc != AFS_OK) { afs_group_destroy(g); return rc; }
    g->ctx.push_back(c);
  }
  *out = g;
  return AFS_OK;
}

int afs_group_destroy(afs_group *g) {
  if (!g) return AFS_OK;
  for (afs_ctx *c : g->ctx) afs_destroy(c);
  delete g;
  return AFS_OK;
}

int afs_group_size(const afs_group *g) { return g ? (int)g->ctx.size() : 0; }

afs_ctx *afs_group_context(afs_group *g, int index) {
  return (g && index >= 0 && index < (int)g->ctx.size()) ? g->ctx[(size_t)index] : nullptr;
}

int afs_group_shard(const afs_group *g, int n_units, int index, int *lo, int *hi) {
  if (!g || g->ctx.empty()) return AFS_ERR_BAD_CONTEXT;
  if (n_units < 0 || index < 0 || index >= (int)g->ctx.size() || !lo || !hi) return AFS_ERR_BAD_ARGUMENT;
  shard_of(n_units, (int)g->ctx.size(), index, lo, hi);
  return AFS_OK;
}

int afs_group_tds_synthesize(afs_group *g, const afs_tds_options *opts, int n_tracts, int n_frames,
                             const int32_t *frame_len, const afs_tube_frame *tubes, const double *glottis,
                             double fs_tds_hz, int sr_factor, const uint32_t *seeds, double *audio, double *flow,
                             uint32_t *status) {
  if (!g || g->ctx.empty()) return AFS_ERR_BAD_CONTEXT;
  if (n_tracts < 0 || n_frames < 2 || !frame_len || sr_factor < 1) return AFS_ERR_BAD_ARGUMENT;
  if (n_tracts == 0) return AFS_OK;
  if (!tubes || !glottis || !audio) return AFS_ERR_BAD_ARGUMENT;
  size_t n_out = 0;
  for (int k = 0; k + 1 < n_frames; k++) { if (frame_len[k] < 1) return AFS_ERR_BAD_ARGUMENT; n_out += (size_t)frame_len[k]; }
  return for_each_shard(g, n_tracts, [&](int r, int lo, int hi) {
    const size_t o = (size_t)lo;
    return afs_tds_synthesize(g->ctx[(size_t)r], opts, hi - lo, n_frames, frame_len, tubes + o * (size_t)n_frames,
                              glottis + o * (size_t)n_frames * AFS_NUM_GLOTTIS_CONTROL, fs_tds_hz, sr_factor,
                              seeds ? seeds + o : nullptr, audio + o * n_out,
                              flow ? flow + o * n_out * (size_t)sr_factor : nullptr, status ? status + o : nullptr);
  });
}

int afs_group_tl_sweep(afs_group *g, const afs_tl_options *opts, int n_tubes, const afs_tube_frame *tubes,
                       int spectrum_type, int section, double f0_hz, int num_freq, double *re, double *im) {
  if (!g || g->ctx.empty()) return AFS_ERR_BAD_CONTEXT;
  if (n_tubes < 0 || num_freq < 1) return AFS_ERR_BAD_ARGUMENT;
  if (n_tubes == 0) return AFS_OK;
  if (!tubes || !re || !im) return AFS_ERR_BAD_ARGUMENT;
  return for_each_shard(g, n_tubes, [&](int r, int lo, int hi) {
    const size_t o = (size_t)lo * (size_t)num_freq;
    return afs_tl_sweep(g->ctx[(size_t)r], opts, hi - lo, tubes + lo, spectrum_type, section, f0_hz, num_freq, re + o, im + o);
  });
}

// ---- chunked synthesis over the group: one afs_tds_stream per device, each holding its shard's models ----------------
struct afs_group_tds_stream {
  afs_group *g = nullptr;
  int n = 0, sr_factor = 1;
  bool primed = false;
  std::vector<afs_tds_stream *> part;  // by device; null where the shard is empty
};

int afs_group_tds_stream_create(afs_group *g, const afs_tds_options *opts, int n_tracts, double fs_tds_hz, int sr_factor,
                                const uint32_t *seeds, afs_group_tds_stream **out) {
  if (!g || g->ctx.empty()) return AFS_ERR_BAD_CONTEXT;
  if (!out || !opts || n_tracts < 1 || sr_factor < 1) return AFS_ERR_BAD_ARGUMENT;
  afs_group_tds_stream *s = new (std::nothrow) afs_group_tds_stream();
  if (!s) return AFS_ERR_OUT_OF_MEMORY;
  s->g = g; s->n = n_tracts; s->sr_factor = sr_factor;
  s->part.assign(g->ctx.size(), nullptr);
  const int rc = for_each_shard(g, n_tracts, [&](int r, int lo, int hi) {
    return afs_tds_stream_create(g->ctx[(size_t)r], opts, hi - lo, fs_tds_hz, sr_factor, seeds ? seeds + lo : nullptr, &s->part[(size_t)r]);
  });
  if (rc != AFS_OK) { afs_group_tds_stream_destroy(s); return rc; }
  *out = s;
  return AFS_OK;
}

int afs_group_tds_stream_destroy(afs_group_tds_stream *s) {
  if (!s) return AFS_OK;
  for (afs_tds_stream *p : s->part) afs_tds_stream_destroy(p);
  delete s;
  return AFS_OK;
}

int afs_group_tds_stream_reset(afs_group_tds_stream *s) {
  if (!s) return AFS_ERR_BAD_ARGUMENT;
  for (afs_tds_stream *p : s->part) if (p) afs_tds_stream_reset(p);
  s->primed = false;
  return AFS_OK;
}

int afs_group_tds_stream_synthesize(afs_group_tds_stream *s, int n_frames, const int32_t *frame_len, const afs_tube_frame *tubes,
                                    const double *glottis, double *audio, double *flow, uint32_t *status) {
  if (!s) return AFS_ERR_BAD_ARGUMENT;
  if (n_frames < 1 || !frame_len || !tubes || !glottis) return AFS_ERR_BAD_ARGUMENT;
  size_t n_out = 0;
  for (int k = s->primed ? 0 : 1; k < n_frames; k++) { if (frame_len[k] < 1) return AFS_ERR_BAD_ARGUMENT; n_out += (size_t)frame_len[k]; }
  if (n_out > 0 && !audio) return AFS_ERR_BAD_ARGUMENT;
  // (a device that fails leaves ITS shard where it was; the caller should reset the stream after an error)
  const int rc = for_each_shard(s->g, s->n, [&](int r, int lo, int hi) {
    const size_t o = (size_t)lo;
    (void)hi;
    return afs_tds_stream_synthesize(s->part[(size_t)r], n_frames, frame_len, tubes + o * (size_t)n_frames,
                                     glottis + o * (size_t)n_frames * AFS_NUM_GLOTTIS_CONTROL, audio ? audio + o * n_out : nullptr,
                                     flow ? flow + o * n_out * (size_t)s->sr_factor : nullptr, status ? status + o : nullptr);
  });
  if (rc == AFS_OK) s->primed = true;
  return rc;
}

// checkpoint / resume of a group stream: the blobs of its per-device streams back to back, in device order
size_t afs_group_tds_stream_checkpoint_bytes(const afs_group_tds_stream *s) {
  size_t total = 0;
  if (s) for (const afs_tds_stream *p : s->part) if (p) total += afs_tds_stream_checkpoint_bytes(p);
  return total;
}

int afs_group_tds_stream_save(afs_group_tds_stream *s, void *blob, size_t bytes) {
  if (!s || !blob || bytes < afs_group_tds_stream_checkpoint_bytes(s)) return AFS_ERR_BAD_ARGUMENT;
  unsigned char *at = (unsigned char *)blob;
  for (afs_tds_stream *p : s->part) {
    if (!p) continue;
    const size_t b = afs_tds_stream_checkpoint_bytes(p);
    const int rc = afs_tds_stream_save(p, at, b);
    if (rc != AFS_OK) return rc;
    at += b;
  }
  return AFS_OK;
}

int afs_group_tds_stream_load(afs_group_tds_stream *s, const void *blob, size_t bytes) {
  if (!s || !blob || bytes < afs_group_tds_stream_checkpoint_bytes(s)) return AFS_ERR_BAD_ARGUMENT;
  const unsigned char *at = (const unsigned char *)blob;
  for (afs_tds_stream *p : s->part) {
    if (!p) continue;
    const size_t b = afs_tds_stream_checkpoint_bytes(p);
    const int rc = afs_tds_stream_load(p, at, b);  // (refuses a blob of another sharding: the batch sizes differ)
    if (rc != AFS_OK) { afs_group_tds_stream_reset(s); return rc; }
    at += b;
  }
  s->primed = false;
  for (const afs_tds_stream *p : s->part) if (p && afs_tds_stream_primed(p) > 0) s->primed = true;
  return AFS_OK;
}

int64_t afs_group_tds_stream_samples(const afs_group_tds_stream *s) {
  if (s) for (const afs_tds_stream *p : s->part) if (p) return afs_tds_stream_samples(p);
  return -1;
}

int afs_host_alloc(void **ptr, size_t bytes) {
  if (!ptr) return AFS_ERR_BAD_ARGUMENT;
  *ptr = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return AFS_ERR_NO_DEVICE; }
  const cudaError_t e = cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable);
  if (e != cudaSuccess) { cudaGetLastError(); return e == cudaErrorMemoryAllocation ? AFS_ERR_OUT_OF_MEMORY : AFS_ERR_CUDA; }
  return AFS_OK;
}

int afs_host_free(void *ptr) {
  if (!ptr) return AFS_OK;
  return cudaFreeHost(ptr) == cudaSuccess ? AFS_OK : AFS_ERR_CUDA;
}

}  // extern "C"
