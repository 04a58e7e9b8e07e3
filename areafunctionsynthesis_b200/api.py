"""ctypes bindings of include/afs_b200.h."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NUM_PM = 40
NUM_GLOTTIS_CONTROL = 6

# afs_tube_frame (include/afs_b200.h)
FRAME_DTYPE = np.dtype(
    [
        ("area_cm2", "<f8", (NUM_PM,)),
        ("length_cm", "<f8", (NUM_PM,)),
        ("laterality", "<f8", (NUM_PM,)),
        ("incisor_pos_cm", "<f8"),
        ("velum_area_cm2", "<f8"),
        ("aspiration_dB", "<f8"),
        ("glottis_area_cm2", "<f8", (2,)),
        ("glottis_length_cm", "<f8", (2,)),
        ("reserved", "<f8"),
        ("articulator", "<i4", (NUM_PM,)),
    ]
)
assert FRAME_DTYPE.itemsize == 1184

SPECTRUM_TYPES = {"input_impedance": 0, "output_impedance": 1, "flow_source_tf": 2, "pressure_source_tf": 3,
                  "radiation": 4}
FIRST_PHARYNX_SECTION = 25
TL_ARITHMETIC_REFERENCE, TL_ARITHMETIC_FAST = 0, 1  # enum afs_tl_arithmetic


class AfsError(RuntimeError):
    pass


class TlOptions(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "radiation", "boundary_layer", "heat_conduction", "soft_walls", "hagen_resistance",
        "inner_length_corrections", "lumped_elements", "paranasal_sinuses", "piriform_fossa",
        "static_pressure_drops")] + [("lung_pressure_dPa", C.c_double), ("arithmetic", C.c_int32),
                                     ("reserved", C.c_int32)]


class TdsOptions(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "turbulence_losses", "soft_walls", "generate_noise_sources", "radiation_from_skin",
        "piriform_fossa", "inner_length_corrections", "transvelar_coupling", "glottis_loss_option")] + [
        ("flow_separation_area_ratio", C.c_double), ("glottis_static", C.c_double * 16),
        ("flow_source_cm3_s", C.c_double), ("flow_source_section", C.c_int32), ("reserved", C.c_int32)]


MAX_TDS_PROBES = 8


class TdsProbes(C.Structure):
    """afs_tds_probes: taps into the simulation (TdsModel::getSectionPressure / getSectionFlow)."""
    _fields_ = [("n_probes", C.c_int32), ("section", C.c_int32 * MAX_TDS_PROBES), ("pressure_dPa", C.c_void_p),
                ("inflow_cm3_s", C.c_void_p), ("outflow_cm3_s", C.c_void_p)]


_lib = None


def library_path():
    # (AFS_LIB: an instrumented build of the same library, profiling tools only)
    return os.environ.get("AFS_LIB") or os.path.join(HERE, "libafs_b200.so")


def load_library():
    """Load libafs_b200.so; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise AfsError(f"{path} is missing: build it with `python -m areafunctionsynthesis_b200.build` "
                       "(there is no CPU fallback)")
    lib = C.CDLL(path)
    lib.afs_status_string.restype = C.c_char_p
    lib.afs_last_cuda_error.restype = C.c_char_p
    lib.afs_last_cuda_error.argtypes = [C.c_void_p]
    lib.afs_launch_count.restype = C.c_int64
    lib.afs_launch_count.argtypes = [C.c_void_p]
    lib.afs_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    lib.afs_destroy.argtypes = [C.c_void_p]
    vp, dbl, i32 = C.c_void_p, C.c_double, C.c_int
    lib.afs_tl_sweep.argtypes = [vp, vp, i32, vp, i32, i32, dbl, i32, vp, vp]
    lib.afs_tl_sweep_device.argtypes = [vp, vp, i32, vp, i32, i32, dbl, i32, vp, vp, vp]
    lib.afs_tl_formants.argtypes = [vp, vp, i32, vp, dbl, i32, i32, vp, vp, vp, vp]
    lib.afs_optimize_formants_vowel.argtypes = [vp, vp, i32, vp, vp, dbl, dbl, vp, i32, dbl, i32, vp, vp, vp, vp, vp]
    lib.afs_synthesize_vowels_lf.argtypes = [vp, vp, i32, vp, vp, i32, dbl, vp, vp]
    lib.afs_target_sequence_length.argtypes = [vp, vp, dbl]
    lib.afs_target_sequence_trajectory.argtypes = [vp, vp, vp, vp, dbl, vp, vp]
    lib.afs_play_target_sequences.argtypes = [vp, vp, i32, vp, vp, vp, vp, dbl, i32, vp, vp, vp]
    lib.afs_lf_vowel_length.argtypes = [i32, dbl]
    lib.afs_lf_glottal_flow.argtypes = [vp, i32, dbl, vp]
    lib.afs_tl_impulse_response.argtypes = [vp, vp, i32, vp, dbl, i32, vp]
    lib.afs_tl_impulse_response_device.argtypes = [vp, vp, i32, vp, dbl, i32, vp, vp]
    lib.afs_tds_synthesize.argtypes = [vp, vp, i32, i32, vp, vp, vp, dbl, i32, vp, vp, vp, vp]
    lib.afs_tds_synthesize_device.argtypes = [vp, vp, i32, i32, vp, vp, vp, dbl, i32, vp, vp, vp, vp, vp]
    lib.afs_tds_synthesize_probed.argtypes = [vp, vp, i32, i32, vp, vp, vp, dbl, i32, vp, vp, vp, vp, vp]
    lib.afs_tds_synthesize_probed_device.argtypes = [vp, vp, i32, i32, vp, vp, vp, dbl, i32, vp, vp, vp, vp, vp, vp]
    lib.afs_tl_mean_flow.argtypes = [vp, i32, vp, dbl, vp, vp]
    lib.afs_tds_stream_create.argtypes = [vp, vp, i32, dbl, i32, vp, C.POINTER(C.c_void_p)]
    lib.afs_tds_stream_destroy.argtypes = [vp]
    lib.afs_tds_stream_reset.argtypes = [vp]
    lib.afs_tds_stream_samples.argtypes = [vp]
    lib.afs_tds_stream_samples.restype = C.c_int64
    lib.afs_tds_stream_synthesize.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp]
    lib.afs_tds_stream_checkpoint_bytes.argtypes = [vp]
    lib.afs_tds_stream_checkpoint_bytes.restype = C.c_size_t
    lib.afs_tds_stream_save.argtypes = [vp, vp, C.c_size_t]
    lib.afs_tds_stream_load.argtypes = [vp, vp, C.c_size_t]
    lib.afs_group_tds_stream_create.argtypes = [vp, vp, i32, dbl, i32, vp, C.POINTER(C.c_void_p)]
    lib.afs_group_tds_stream_destroy.argtypes = [vp]
    lib.afs_group_tds_stream_reset.argtypes = [vp]
    lib.afs_group_tds_stream_synthesize.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp]
    lib.afs_group_tds_stream_samples.argtypes = [vp]
    lib.afs_group_tds_stream_samples.restype = C.c_int64
    lib.afs_group_tds_stream_checkpoint_bytes.argtypes = [vp]
    lib.afs_group_tds_stream_checkpoint_bytes.restype = C.c_size_t
    lib.afs_group_tds_stream_save.argtypes = [vp, vp, C.c_size_t]
    lib.afs_group_tds_stream_load.argtypes = [vp, vp, C.c_size_t]
    lib.afs_tds_stream_primed.argtypes = [vp]
    lib.afs_tds_set_state_dump.argtypes = [vp, i32, vp]
    lib.afs_debug_set.argtypes = [vp, C.c_char_p, C.c_int64]
    lib.afs_group_create.argtypes = [C.POINTER(C.c_void_p), vp, i32]
    lib.afs_group_destroy.argtypes = [vp]
    lib.afs_group_size.argtypes = [vp]
    lib.afs_group_context.argtypes = [vp, i32]
    lib.afs_group_context.restype = C.c_void_p
    lib.afs_group_shard.argtypes = [vp, i32, i32, vp, vp]
    lib.afs_group_tds_synthesize.argtypes = [vp, vp, i32, i32, vp, vp, vp, dbl, i32, vp, vp, vp, vp]
    lib.afs_group_tl_sweep.argtypes = [vp, vp, i32, vp, i32, i32, dbl, i32, vp, vp]
    lib.afs_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    lib.afs_host_free.argtypes = [vp]
    lib.afs_probe_fp64_tflops.argtypes = [vp, vp]
    lib.afs_chebyshev_lowpass.argtypes = [dbl, i32, vp, vp]
    lib.afs_audio_to_int16.argtypes = [vp, C.c_int64, vp]
    lib.afs_write_wav_int16.argtypes = [C.c_char_p, vp, C.c_int64, i32, i32]
    lib.afs_area_function_to_frame.argtypes = [vp, vp]
    lib.afs_area_functions_to_frames.argtypes = [vp, i32, vp, vp]
    lib.afs_area_functions_to_frames_device.argtypes = [vp, i32, vp, vp, vp]
    _lib = lib
    return lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if hasattr(a, "data_ptr"):  # torch tensor
        return C.c_void_p(a.data_ptr())
    return a


def _check(lib, rc, ctx=None):
    if rc != 0:
        msg = lib.afs_status_string(rc).decode()
        if ctx is not None:
            extra = lib.afs_last_cuda_error(ctx).decode()
            if extra:
                msg += f" [{extra}]"
        raise AfsError(f"afs status {rc}: {msg}")


def tl_options_default():
    o = TlOptions()
    load_library().afs_tl_options_default(C.byref(o))
    return o


def tds_options_default(noise=True):
    """TdsModel::TdsModel() defaults (TdsModel.cpp:35-44); `noise=False` switches generateNoiseSources off."""
    o = TdsOptions()
    load_library().afs_tds_options_default(C.byref(o))
    if not noise:
        o.generate_noise_sources = 0
    return o


def default_frames(n):
    lib = load_library()
    f = np.zeros(n, dtype=FRAME_DTYPE)
    if n:
        lib.afs_tube_frame_default(_ptr(f[:1]))
        f[:] = f[0]
    return f


def area_function_to_frames(params, frames=None):
    """OneDimAreaFunction parameters [N,16] -> frames (pharynx/mouth part filled)."""
    lib = load_library()
    p = np.ascontiguousarray(np.atleast_2d(params), dtype=np.float64)
    out = default_frames(len(p)) if frames is None else frames
    for i in range(len(p)):
        _check(lib, lib.afs_area_function_to_frame(_ptr(p[i]), _ptr(out[i:i + 1])))
    return out


def target_sequence_trajectory(shapes4x16, stationary_s, transition_s, glottis_control, sampling_rate_hz=22050.0):
    """Host helper: the per-sample trajectories of Synthesizer::playTargetSequence: (params [n, 16], glottis [n, 6])."""
    lib = load_library()
    sh = np.ascontiguousarray(shapes4x16, dtype=np.float64).reshape(4, 16)
    st = np.ascontiguousarray(stationary_s, dtype=np.float64); tr = np.ascontiguousarray(transition_s, dtype=np.float64)
    g6 = np.ascontiguousarray(glottis_control, dtype=np.float64)
    n = int(lib.afs_target_sequence_length(_ptr(st), _ptr(tr), float(sampling_rate_hz)))
    prm = np.zeros((n, 16)); gl = np.zeros((n, NUM_GLOTTIS_CONTROL))
    rc = lib.afs_target_sequence_trajectory(_ptr(sh), _ptr(st), _ptr(tr), _ptr(g6), float(sampling_rate_hz), _ptr(prm), _ptr(gl))
    if rc:
        raise AfsError(f"afs_target_sequence_trajectory: status {rc}")
    return prm, gl


def lf_glottal_flow(lf5, is_long=True, sampling_rate_hz=22050.0):
    """Host helper: the LF pulse train of one vowel (lf5 = AMP, F0, OQ, SQ, TL)."""
    lib = load_library()
    lf = np.ascontiguousarray(lf5, dtype=np.float64)
    n = int(lib.afs_lf_vowel_length(int(bool(is_long)), float(sampling_rate_hz)))
    flow = np.zeros(n)
    rc = lib.afs_lf_glottal_flow(_ptr(lf), int(bool(is_long)), float(sampling_rate_hz), _ptr(flow))
    if rc:
        raise AfsError(f"afs_lf_glottal_flow: status {rc}")
    return flow


def audio_to_int16(audio):
    """Synthesizer.cpp:955-965: double [-1, 1] -> int16 (truncation, clipping)."""
    a = np.ascontiguousarray(audio, dtype=np.float64)
    out = np.empty(a.shape, dtype=np.int16)
    rc = load_library().afs_audio_to_int16(_ptr(a), a.size, _ptr(out))
    if rc:
        raise AfsError(f"afs_audio_to_int16: status {rc}")
    return out


def write_wav_int16(path, pcm, sample_rate_hz, n_channels=1):
    p = np.ascontiguousarray(pcm, dtype=np.int16)
    rc = load_library().afs_write_wav_int16(os.fsencode(path), _ptr(p), p.size // n_channels, int(n_channels), int(sample_rate_hz))
    if rc:
        raise AfsError(f"afs_write_wav_int16: status {rc}")


def chebyshev_lowpass(cutoff_ratio, poles):
    lib = load_library()
    a = np.zeros(poles + 1)
    b = np.zeros(poles + 1)
    _check(lib, lib.afs_chebyshev_lowpass(cutoff_ratio, poles, _ptr(a), _ptr(b)))
    return a, b


def pinned_empty(shape, dtype=np.float64):
    """numpy array in page-locked host memory (afs_host_alloc); keep the returned array alive while it is in use."""
    lib = load_library()
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    _check(lib, lib.afs_host_alloc(C.byref(p), n))
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    a = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[p.value] = (buf, lib)
    return a


_PINNED = {}


def pinned_free(a):
    addr = a.ctypes.data
    buf, lib = _PINNED.pop(addr)
    lib.afs_host_free(C.c_void_p(addr))


class Group:
    """Several devices of one box in one process (afs_group): contiguous shards, one host thread per device."""

    def __init__(self, devices=None):
        self.lib = load_library()
        h = C.c_void_p()
        ids = None if devices is None else np.ascontiguousarray(devices, dtype=np.int32)
        _check(self.lib, self.lib.afs_group_create(C.byref(h), _ptr(ids), 0 if ids is None else len(ids)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.afs_group_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return int(self.lib.afs_group_size(self.h))

    def shard(self, n_units, index):
        lo, hi = C.c_int(), C.c_int()
        _check(self.lib, self.lib.afs_group_shard(self.h, int(n_units), int(index), C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def tds_synthesize(self, frames, glottis, frame_len, fs_tds_hz=44100.0, sr_factor=2, opts=None, seeds=None, audio=None,
                       want_flow=False):
        opts = opts or tds_options_default()
        frames = np.ascontiguousarray(frames, dtype=FRAME_DTYPE)
        B, F = frames.shape
        glottis = np.ascontiguousarray(glottis, dtype=np.float64).reshape(B, F, NUM_GLOTTIS_CONTROL)
        frame_len = np.ascontiguousarray(frame_len, dtype=np.int32)
        n_out = int(frame_len.sum())
        if audio is None:
            audio = np.empty((B, n_out))
        flow = np.empty((B, n_out * sr_factor)) if want_flow else None
        status = np.zeros(B, dtype=np.uint32)
        s = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint32)
        _check(self.lib, self.lib.afs_group_tds_synthesize(self.h, C.byref(opts), B, F, _ptr(frame_len), _ptr(frames), _ptr(glottis),
                                                           float(fs_tds_hz), int(sr_factor), _ptr(s), _ptr(audio), _ptr(flow),
                                                           _ptr(status)))
        return audio, flow, status

    def tl_sweep(self, frames, f0_hz, num_freq, spectrum_type=2, section=FIRST_PHARYNX_SECTION, opts=None):
        opts = opts or tl_options_default()
        frames = np.ascontiguousarray(frames, dtype=FRAME_DTYPE)
        n = len(frames)
        re = np.empty((n, num_freq)); im = np.empty((n, num_freq))
        _check(self.lib, self.lib.afs_group_tl_sweep(self.h, C.byref(opts), n, _ptr(frames), int(spectrum_type), int(section),
                                                     float(f0_hz), int(num_freq), _ptr(re), _ptr(im)))
        return re, im


class TdsStream:
    """afs_tds_stream: a batch of utterances whose models live across calls (chunked synthesis, checkpoint / resume)."""

    def __init__(self, ctx, n_tracts, fs_tds_hz=44100.0, sr_factor=2, opts=None, seeds=None):
        """ctx: a Context, or a Group (afs_group_tds_stream_*: every device keeps the models of its shard)."""
        self.ctx, self.lib, self.n, self.sr_factor = ctx, ctx.lib, int(n_tracts), int(sr_factor)
        self.opts = opts or tds_options_default()
        s = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint32)
        self.h = C.c_void_p()
        self._group = isinstance(ctx, Group)
        self._err = None if self._group else ctx.h
        pre = "afs_group_tds_stream_" if self._group else "afs_tds_stream_"
        self._f = {k: getattr(self.lib, pre + k) for k in ("create", "destroy", "reset", "synthesize", "samples", "checkpoint_bytes",
                                                           "save", "load")}
        _check(self.lib, self._f["create"](ctx.h, C.byref(self.opts), self.n, float(fs_tds_hz), self.sr_factor,
                                           _ptr(s), C.byref(self.h)), self._err)
        self._primed = False

    def close(self):
        if self.h:
            if getattr(self.ctx, "h", None):  # (a stream must not outlive its context: nothing to release then)
                self._f["destroy"](self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        _check(self.lib, self._f["reset"](self.h), self._err)
        self._primed = False

    @property
    def samples(self):
        return int(self._f["samples"](self.h))

    def synthesize(self, frames, glottis, frame_len, want_flow=False):
        """frames [B, F], glottis [B, F, 6], frame_len [F]: frame k yields frame_len[k] samples interpolated from the
        previous frame; the first frame of a fresh stream only primes.  Returns (audio [B, n_out], flow or None, status)."""
        frames = np.ascontiguousarray(frames, dtype=FRAME_DTYPE)
        B, F = frames.shape
        assert B == self.n
        glottis = np.ascontiguousarray(glottis, dtype=np.float64).reshape(B, F, NUM_GLOTTIS_CONTROL)
        frame_len = np.ascontiguousarray(frame_len, dtype=np.int32)
        assert frame_len.shape == (F,)
        n_out = int(frame_len[0 if self._primed else 1:].sum())
        audio = np.empty((B, n_out))
        flow = np.empty((B, n_out * self.sr_factor)) if want_flow else None
        status = np.zeros(B, dtype=np.uint32)
        _check(self.lib, self._f["synthesize"](self.h, F, _ptr(frame_len), _ptr(frames), _ptr(glottis), _ptr(audio),
                                               _ptr(flow), _ptr(status)), self._err)
        self._primed = True
        return audio, flow, status

    def save(self):
        blob = np.empty(int(self._f["checkpoint_bytes"](self.h)), dtype=np.uint8)
        _check(self.lib, self._f["save"](self.h, _ptr(blob), blob.size), self._err)
        return blob

    def load(self, blob):
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        _check(self.lib, self._f["load"](self.h, _ptr(blob), blob.size), self._err)
        self._primed = bool(np.frombuffer(blob[16:20].tobytes(), dtype=np.int32)[0])  # StreamHeader.primed (of the first shard)


class Context:
    """One CUDA device (afs_ctx)."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        _check(self.lib, self.lib.afs_create(C.byref(h), int(device)))
        self.h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "h", None):
            self.lib.afs_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def probe_fp64_tflops(self):
        out = C.c_double()
        _check(self.lib, self.lib.afs_probe_fp64_tflops(self.h, C.byref(out)), self.h)
        return out.value

    @property
    def launch_count(self):
        return int(self.lib.afs_launch_count(self.h))

    # ---- TL ----------------------------------------------------------------
    def tl_sweep(self, frames, f0_hz, num_freq, spectrum_type="flow_source_tf", section=FIRST_PHARYNX_SECTION,
                 opts=None):
        """Host buffers in/out (H2D + kernel + D2H).  Returns (re, im) [n, num_freq]."""
        opts = opts or tl_options_default()
        frames = np.ascontiguousarray(frames, dtype=FRAME_DTYPE)
        n = len(frames)
        re = np.empty((n, num_freq))
        im = np.empty((n, num_freq))
        t = SPECTRUM_TYPES[spectrum_type] if isinstance(spectrum_type, str) else int(spectrum_type)
        _check(self.lib, self.lib.afs_tl_sweep(self.h, C.byref(opts), n, _ptr(frames), t, int(section),
                                               float(f0_hz), int(num_freq), _ptr(re), _ptr(im)), self.h)
        return re, im

    def tl_sweep_device(self, d_frames, n, f0_hz, num_freq, d_re, d_im, stream, spectrum_type=2,
                        section=FIRST_PHARYNX_SECTION, opts=None):
        """Device-resident variant: d_* are torch CUDA tensors (or raw pointers); async on `stream`."""
        opts = opts or tl_options_default()
        _check(self.lib, self.lib.afs_tl_sweep_device(self.h, C.byref(opts), int(n), _ptr(d_frames),
                                                      int(spectrum_type), int(section), float(f0_hz), int(num_freq),
                                                      _ptr(d_re), _ptr(d_im), C.c_void_p(stream)), self.h)

    def optimize_formants_vowel(self, params, targets_hz, max_param_change_cm=1.0, min_advised_area_cm2=0.0,
                                param_fixed=None, max_runs=100, sampling_rate_hz=22050.0, spectrum_length=2048, opts=None):
        """Data::optimizeFormantsVowel for a batch of problems: params [n, 16], targets_hz [n, 3].
        Returns dict(params, initial_error, final_error, runs, formants_hz, trace[n, max_runs, 2])."""
        opts = opts or tl_options_default()
        p = np.array(params, dtype=np.float64).reshape(-1, 16).copy()
        n = len(p)
        t = np.ascontiguousarray(targets_hz, dtype=np.float64).reshape(n, 3)
        fx = None if param_fixed is None else np.ascontiguousarray(param_fixed, dtype=np.int32)
        e0 = np.zeros(n); e1 = np.zeros(n); runs = np.zeros(n, dtype=np.int32); fo = np.zeros((n, 3))
        trace = np.full((n, int(max_runs), 2), np.nan)
        _check(self.lib, self.lib.afs_optimize_formants_vowel(
            self.h, C.byref(opts), n, _ptr(p), _ptr(t), float(max_param_change_cm), float(min_advised_area_cm2), _ptr(fx),
            int(max_runs), float(sampling_rate_hz), int(spectrum_length), _ptr(e0), _ptr(e1), _ptr(runs), _ptr(fo),
            _ptr(trace)), self.h)
        return dict(params=p, initial_error=e0, final_error=e1, runs=runs, formants_hz=fo, trace=trace)

    def play_target_sequences(self, shapes, stationary_s, transition_s, glottis_control, sampling_rate_hz=22050.0, sr_factor=2,
                              opts=None, seeds=None):
        """Synthesizer::playTargetSequence for a batch: shapes [n, 4, 16].  Returns (audio [n, num_samples], status [n])."""
        opts = opts or tds_options_default()
        sh = np.ascontiguousarray(shapes, dtype=np.float64).reshape(-1, 4, 16)
        n = len(sh)
        st = np.ascontiguousarray(stationary_s, dtype=np.float64); tr = np.ascontiguousarray(transition_s, dtype=np.float64)
        g6 = np.ascontiguousarray(glottis_control, dtype=np.float64)
        ns = int(self.lib.afs_target_sequence_length(_ptr(st), _ptr(tr), float(sampling_rate_hz)))
        audio = np.zeros((n, ns)); status = np.zeros(n, dtype=np.uint32)
        s = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint32)
        _check(self.lib, self.lib.afs_play_target_sequences(self.h, C.byref(opts), n, _ptr(sh), _ptr(st), _ptr(tr), _ptr(g6),
                                                            float(sampling_rate_hz), int(sr_factor), _ptr(s), _ptr(audio),
                                                            _ptr(status)), self.h)
        return audio, status

    def synthesize_vowels_lf(self, frames, lf_params, is_long=True, sampling_rate_hz=22050.0, opts=None, want_pressure=False):
        """Synthesizer::synthesizeVowelLf for a batch: lf_params [n, 5] = AMP, F0, OQ, SQ, TL.  Returns int16 [n, length]
        (and the unfiltered pressure when asked)."""
        opts = opts or tl_options_default()
        frames = np.ascontiguousarray(frames, dtype=FRAME_DTYPE)
        n = len(frames)
        lf = np.ascontiguousarray(lf_params, dtype=np.float64).reshape(n, 5)
        length = int(self.lib.afs_lf_vowel_length(int(bool(is_long)), float(sampling_rate_hz)))
        audio = np.zeros((n, length), dtype=np.int16)
        pr = np.zeros((n, length)) if want_pressure else None
        _check(self.lib, self.lib.afs_synthesize_vowels_lf(self.h, C.byref(opts), n, _ptr(frames), _ptr(lf), int(bool(is_long)),
                                                           float(sampling_rate_hz), _ptr(audio), _ptr(pr)), self.h)
        return (audio, pr) if want_pressure else audio

    def tl_impulse_response(self, frames, sampling_rate_hz=22050.0, length_exponent=9, opts=None):
        """TlModel::getImpulseResponse for a batch: returns [n, 2**length_exponent]."""
        opts = opts or tl_options_default()
        frames = np.ascontiguousarray(frames, dtype=FRAME_DTYPE)
        n = len(frames)
        out = np.empty((n, 1 << int(length_exponent)))
        _check(self.lib, self.lib.afs_tl_impulse_response(self.h, C.byref(opts), n, _ptr(frames), float(sampling_rate_hz),
                                                          int(length_exponent), _ptr(out)), self.h)
        return out

    def tl_mean_flow(self, frames, lung_pressure_dPa):
        """TlModel::getMeanFlow / getMostConstrictedSection for a batch: (flow [n], section [n])."""
        frames = np.ascontiguousarray(frames, dtype=FRAME_DTYPE)
        n = len(frames)
        flow = np.zeros(n); sec = np.zeros(n, dtype=np.int32)
        _check(self.lib, self.lib.afs_tl_mean_flow(self.h, n, _ptr(frames), float(lung_pressure_dPa), _ptr(flow), _ptr(sec)), self.h)
        return flow, sec

    def tl_formants(self, frames, sampling_rate_hz=22050.0, spectrum_length=2048, max_formants=8, opts=None):
        opts = opts or tl_options_default()
        frames = np.ascontiguousarray(frames, dtype=FRAME_DTYPE)
        n = len(frames)
        f = np.zeros((n, max_formants)); bw = np.zeros((n, max_formants))
        num = np.zeros(n, dtype=np.int32); flags = np.zeros(n, dtype=np.int32)
        _check(self.lib, self.lib.afs_tl_formants(self.h, C.byref(opts), n, _ptr(frames), float(sampling_rate_hz),
                                                  int(spectrum_length), int(max_formants), _ptr(f), _ptr(bw),
                                                  _ptr(num), _ptr(flags)), self.h)
        return f, bw, num, flags

    # ---- area-function rasteriser (K4) -------------------------------------
    def area_functions_to_frames(self, params, frames=None):
        """OneDimAreaFunction parameters [N,16] -> frames on the device (host buffers in and out)."""
        p = np.ascontiguousarray(np.atleast_2d(params), dtype=np.float64)
        out = default_frames(len(p)) if frames is None else np.ascontiguousarray(frames, dtype=FRAME_DTYPE)
        _check(self.lib, self.lib.afs_area_functions_to_frames(self.h, len(p), _ptr(p), _ptr(out)), self.h)
        return out

    def area_functions_to_frames_device(self, d_params, d_frames, n, stream):
        _check(self.lib, self.lib.afs_area_functions_to_frames_device(self.h, int(n), _ptr(d_params), _ptr(d_frames),
                                                                      C.c_void_p(stream)), self.h)

    # ---- TDS ---------------------------------------------------------------
    def set_state_dump(self, step, out473=None):
        """Diagnostics: dump tract 0's state after simulation step `step` (1-based) on the next calls."""
        self._dump = out473 if step else None
        _check(self.lib, self.lib.afs_tds_set_state_dump(self.h, int(step), _ptr(self._dump)), self.h)

    def debug_set(self, key, value):
        """Test hooks of the library (afs_debug_set)."""
        _check(self.lib, self.lib.afs_debug_set(self.h, key.encode(), int(value)), self.h)

    def tds_synthesize(self, frames, glottis, frame_len, fs_tds_hz=44100.0, sr_factor=2, opts=None, seeds=None,
                       want_flow=False):
        """frames [B, F] FRAME_DTYPE, glottis [B, F, 6], frame_len [F-1] output samples per interval.
        Host buffers (H2D + kernel + D2H).  Returns (audio [B, n_out], flow or None, status [B]).
        Defaults: a simulation at 44.1 kHz (sr_factor 2), the BASELINE configuration; the reference as shipped runs at
        fs_tds_hz = 22050, sr_factor = 1 (Constants.h:25)."""
        opts = opts or tds_options_default()
        frames = np.ascontiguousarray(frames, dtype=FRAME_DTYPE)
        B, F = frames.shape
        glottis = np.ascontiguousarray(glottis, dtype=np.float64).reshape(B, F, NUM_GLOTTIS_CONTROL)
        frame_len = np.ascontiguousarray(frame_len, dtype=np.int32)
        assert frame_len.shape == (F - 1,)
        n_out = int(frame_len.sum())
        audio = np.empty((B, n_out))
        flow = np.empty((B, n_out * sr_factor)) if want_flow else None
        status = np.zeros(B, dtype=np.uint32)
        s = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint32)
        _check(self.lib, self.lib.afs_tds_synthesize(self.h, C.byref(opts), B, F, _ptr(frame_len), _ptr(frames),
                                                     _ptr(glottis), float(fs_tds_hz), int(sr_factor), _ptr(s),
                                                     _ptr(audio), _ptr(flow), _ptr(status)), self.h)
        return audio, flow, status

    def tds_synthesize_probed(self, frames, glottis, frame_len, sections, fs_tds_hz=44100.0, sr_factor=2, opts=None, seeds=None,
                              want=("pressure", "inflow", "outflow")):
        """afs_tds_synthesize_probed: returns (audio, status, dict(pressure / inflow / outflow -> [B, n_probes, steps]))."""
        opts = opts or tds_options_default()
        frames = np.ascontiguousarray(frames, dtype=FRAME_DTYPE)
        B, F = frames.shape
        glottis = np.ascontiguousarray(glottis, dtype=np.float64).reshape(B, F, NUM_GLOTTIS_CONTROL)
        frame_len = np.ascontiguousarray(frame_len, dtype=np.int32)
        n_out = int(frame_len.sum())
        audio = np.empty((B, n_out)); status = np.zeros(B, dtype=np.uint32)
        s = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint32)
        pr = TdsProbes()
        pr.n_probes = len(sections)
        for q, sc in enumerate(sections[:MAX_TDS_PROBES]):
            pr.section[q] = int(sc)
        taps = {k: np.zeros((B, len(sections), n_out * sr_factor)) for k in want}
        pr.pressure_dPa = taps["pressure"].ctypes.data if "pressure" in taps else None
        pr.inflow_cm3_s = taps["inflow"].ctypes.data if "inflow" in taps else None
        pr.outflow_cm3_s = taps["outflow"].ctypes.data if "outflow" in taps else None
        _check(self.lib, self.lib.afs_tds_synthesize_probed(self.h, C.byref(opts), B, F, _ptr(frame_len), _ptr(frames), _ptr(glottis),
                                                            float(fs_tds_hz), int(sr_factor), _ptr(s), _ptr(audio), None,
                                                            _ptr(status), C.byref(pr)), self.h)
        return audio, status, taps

    def tds_synthesize_device(self, d_frames, d_glottis, n_tracts, n_frames, frame_len, d_audio, stream,
                              fs_tds_hz=44100.0, sr_factor=2, opts=None, d_seeds=None, d_flow=None, d_status=None):
        opts = opts or tds_options_default()
        frame_len = np.ascontiguousarray(frame_len, dtype=np.int32)
        _check(self.lib, self.lib.afs_tds_synthesize_device(
            self.h, C.byref(opts), int(n_tracts), int(n_frames), _ptr(frame_len), _ptr(d_frames), _ptr(d_glottis),
            float(fs_tds_hz), int(sr_factor), _ptr(d_seeds), _ptr(d_audio), _ptr(d_flow), _ptr(d_status),
            C.c_void_p(stream)), self.h)
