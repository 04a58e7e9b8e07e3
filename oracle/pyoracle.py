"""ctypes bindings for the oracle (TEST INFRASTRUCTURE -- never imported by the product).

Two checkers live here:
  * `Ref(sr)`      -- oracle/_ref/libafs_ref_sr{22,44}.so: the UNMODIFIED reference classes
                      compiled from /root/reference (oracle/Makefile, oracle/ref_harness.cpp);
  * `Oracle()`     -- oracle/liboracle.so: the plain-C restatement (oracle/afs_oracle.c).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

NUM_PM = 40
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
    ],
    align=False,
)
assert FRAME_DTYPE.itemsize == 1184


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


GLOTTIS_STATIC_DEFAULT = (1.3, 0.24, 0.06, 0.12, 0.03, 0.1, 0.6, 80000.0, 8000.0, 240000.0, 24000.0,
                          25000.0, 0.05, 0.01, 129.0, 125.51)  # TriangularGlottis.cpp:38-56


def tl_options_default():
    """TlModel::TlModel() defaults (TlModel.cpp:30-47)."""
    return TlOptions(3, 1, 0, 1, 0, 0, 1, 1, 0, 1, 0.0)


def tds_options_default(noise=False):
    """TdsModel::TdsModel() defaults (TdsModel.cpp:35-44); noise selectable."""
    o = TdsOptions(1, 1, 1 if noise else 0, 1, 0, 1, 0, 0, 1.0)
    for i, v in enumerate(GLOTTIS_STATIC_DEFAULT):
        o.glottis_static[i] = v
    o.flow_source_cm3_s, o.flow_source_section = 0.0, -1  # TdsModel.cpp:84
    return o


def default_frames(n):
    """Frames equal to a freshly constructed reference Tube (Tube.cpp:46-73, 265-315)."""
    f = np.zeros(n, dtype=FRAME_DTYPE)
    f["area_cm2"] = 4.0
    f["length_cm"] = 16.0 / 40.0
    f["incisor_pos_cm"] = 15.0
    f["aspiration_dB"] = -40.0
    f["glottis_area_cm2"] = 0.1
    f["glottis_length_cm"] = 0.3
    f["articulator"] = 4
    return f


def _ptr(a, t=C.c_void_p):
    return a.ctypes.data_as(t) if a is not None else None


def build(target):
    subprocess.run(["make", "-s", "-C", HERE, target], check=True, stdout=subprocess.DEVNULL)


class Ref:
    """The compiled reference (sr = 22: as shipped; sr = 44: TDS_SR_FACTOR = 2)."""

    def __init__(self, sr=44):
        path = os.path.join(HERE, "_ref", f"libafs_ref_sr{sr}.so")
        if not os.path.exists(path):
            if os.path.isdir("/root/reference"):
                build("ref")
            else:
                raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.lib.ref_tl_spectrum.restype = C.c_int
        self.sampling_rate = self.lib.ref_sampling_rate()
        self.sr_factor = self.lib.ref_tds_sr_factor()

    def area_function_to_frame(self, params16, frame=None):
        p = np.ascontiguousarray(params16, dtype=np.float64)
        out = default_frames(1) if frame is None else frame
        self.lib.ref_area_function_to_frame(_ptr(p), _ptr(out))
        return out

    def frame_sections(self, frame):
        out = np.zeros((93, 7))
        self.lib.ref_frame_sections(_ptr(frame), _ptr(out))
        return out

    def tl_spectrum(self, frames, spectrum_length, opts=None, type_=2, section=25, threads=1):
        opts = opts or tl_options_default()
        n = len(frames)
        max_bins = 2048
        re = np.zeros((n, max_bins))
        im = np.zeros((n, max_bins))
        nf = self.lib.ref_tl_spectrum(C.byref(opts), n, _ptr(frames), type_, section, spectrum_length,
                                      max_bins, _ptr(re), _ptr(im), threads)
        return re[:, :nf].copy(), im[:, :nf].copy()

    def optimize_formants_vowel(self, params16, targets, max_param_change_cm=1.0, min_advised_area_cm2=0.0,
                                param_fixed=None, max_runs=100, opts=None):
        """Data::optimizeFormantsVowel (Data.cpp:203-487) restated over the reference classes (ref_harness.cpp).
        Returns (params, initial_error, final_error, runs, trace[runs, 2], formants[3])."""
        opts = opts or tl_options_default()
        p = np.array(params16, dtype=np.float64).copy()
        t = np.ascontiguousarray(targets, dtype=np.float64)
        fx = np.zeros(16, dtype=np.int32) if param_fixed is None else np.ascontiguousarray(param_fixed, dtype=np.int32)
        out3 = np.zeros(3); trace = np.zeros((max_runs, 2)); fo = np.zeros(3)
        self.lib.ref_optimize_formants_vowel(C.byref(opts), _ptr(p), _ptr(t), C.c_double(max_param_change_cm),
                                             C.c_double(min_advised_area_cm2), _ptr(fx), int(max_runs), _ptr(out3),
                                             _ptr(trace), _ptr(fo))
        runs = int(out3[2])
        return p, out3[0], out3[1], runs, trace[:runs], fo

    def play_target_sequence(self, shapes4x16, stationary_s, transition_s, glottis6, opts=None, seed=1):
        """Synthesizer::playTargetSequence restated over the reference classes (ref_harness.cpp).
        Returns (audio [numSamples], params [numSamples, 16], glottis [numSamples, 6])."""
        opts = opts or tds_options_default()
        sh = np.ascontiguousarray(shapes4x16, dtype=np.float64).reshape(4, 16)
        st = np.ascontiguousarray(stationary_s, dtype=np.float64); tr = np.ascontiguousarray(transition_s, dtype=np.float64)
        g6 = np.ascontiguousarray(glottis6, dtype=np.float64)
        n = self.lib.ref_target_sequence_length(_ptr(st), _ptr(tr))
        audio = np.zeros(n); prm = np.zeros((n, 16)); gl = np.zeros((n, 6))
        m = self.lib.ref_play_target_sequence(C.byref(opts), _ptr(sh), _ptr(st), _ptr(tr), _ptr(g6), C.c_uint32(seed), _ptr(audio),
                                              _ptr(prm), _ptr(gl))
        assert m == n
        return audio, prm, gl

    def synthesize_vowel_lf(self, frame, lf5, is_long, opts=None):
        """Synthesizer::synthesizeVowelLf restated over the reference classes: (int16 audio, float64 pressure)."""
        opts = opts or tl_options_default()
        n = self.lib.ref_lf_vowel_length(int(is_long))
        out = np.zeros(n, dtype=np.int16); pr = np.zeros(n); fl = np.zeros(n)
        lf = np.ascontiguousarray(lf5, dtype=np.float64)
        m = self.lib.ref_synthesize_vowel_lf(C.byref(opts), _ptr(np.ascontiguousarray(frame)), _ptr(lf), int(is_long), _ptr(out),
                                             _ptr(pr), _ptr(fl))
        assert m == n
        self.last_flow = fl
        return out, pr

    def tl_impulse_response(self, frames, length_exponent=9, opts=None):
        opts = opts or tl_options_default()
        n = len(frames)
        out = np.zeros((n, 1 << length_exponent))
        self.lib.ref_tl_impulse_response(C.byref(opts), n, _ptr(np.ascontiguousarray(frames)), length_exponent, _ptr(out))
        return out

    def tl_formants(self, frames, opts=None, max_formants=8):
        opts = opts or tl_options_default()
        n = len(frames)
        f = np.zeros((n, max_formants)); bw = np.zeros((n, max_formants))
        num = np.zeros(n, dtype=np.int32); flags = np.zeros(n, dtype=np.int32)
        self.lib.ref_tl_formants(C.byref(opts), n, _ptr(frames), max_formants, _ptr(f), _ptr(bw), _ptr(num), _ptr(flags))
        return f, bw, num, flags

    def tds_synthesize(self, frames, glottis, frame_len, opts=None, seeds=None, want_flow=True, threads=1):
        """frames [B, F] FRAME_DTYPE, glottis [B, F, 6], frame_len [F-1] (output-rate samples)."""
        opts = opts or tds_options_default()
        B, F = frames.shape
        frame_len = np.ascontiguousarray(frame_len, dtype=np.int32)
        n_out = int(frame_len.sum())
        audio = np.zeros((B, n_out))
        flow = np.zeros((B, n_out * self.sr_factor)) if want_flow else None
        g = np.ascontiguousarray(glottis, dtype=np.float64)
        s = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint32)
        self.lib.ref_tds_synthesize(C.byref(opts), B, F, _ptr(frame_len), _ptr(np.ascontiguousarray(frames)), _ptr(g),
                                    _ptr(s), _ptr(audio), _ptr(flow), threads)
        return audio, flow

    def tds_probed(self, frames, glottis, frame_len, sections, opts=None, seed=1):
        """One utterance (frames [F]) with TdsModel::getSectionPressure / getSectionFlow read after every step:
        (audio, pressure, inflow, outflow), the last three [n_probes, steps]."""
        opts = opts or tds_options_default()
        F = frames.shape[0]
        frame_len = np.ascontiguousarray(frame_len, dtype=np.int32)
        n_out = int(frame_len.sum()); steps = n_out * self.sr_factor
        sec = np.ascontiguousarray(sections, dtype=np.int32)
        audio = np.zeros(n_out); pr = np.zeros((len(sec), steps)); fi = np.zeros_like(pr); fo = np.zeros_like(pr)
        g = np.ascontiguousarray(glottis, dtype=np.float64)
        self.lib.ref_tds_probed(C.byref(opts), F, _ptr(frame_len), _ptr(np.ascontiguousarray(frames)), _ptr(g), C.c_uint32(seed),
                                _ptr(audio), len(sec), _ptr(sec), _ptr(pr), _ptr(fi), _ptr(fo))
        return audio, pr, fi, fo

    def tl_mean_flow(self, frames, lung_pressure_dPa):
        n = len(frames)
        flow = np.zeros(n); sec = np.zeros(n, dtype=np.int32)
        self.lib.ref_tl_mean_flow(n, _ptr(np.ascontiguousarray(frames)), C.c_double(lung_pressure_dPa), _ptr(flow), _ptr(sec))
        return flow, sec

    def tds_state_after(self, frames, glottis, frame_len, step, opts=None):
        opts = opts or tds_options_default()
        F = frames.shape[0]
        frame_len = np.ascontiguousarray(frame_len, dtype=np.int32)
        n_out = int(frame_len.sum())
        audio = np.zeros(n_out); flow = np.zeros(n_out * self.sr_factor); st = np.zeros(473)
        g = np.ascontiguousarray(glottis, dtype=np.float64)
        self.lib.ref_tds_state_after(C.byref(opts), F, _ptr(frame_len), _ptr(np.ascontiguousarray(frames)), _ptr(g),
                                     step, _ptr(audio), _ptr(flow), _ptr(st))
        return audio, flow, st

    def chebyshev_lowpass(self, cutoff_ratio, poles):
        a = np.zeros(poles + 1); b = np.zeros(poles + 1)
        self.lib.ref_chebyshev_lowpass(C.c_double(cutoff_ratio), poles, _ptr(a), _ptr(b))
        return a, b


def load_default_params(path="/root/reference/Default.params"):
    """The 51 named shapes of Default.params (format: MainWindow.cpp:193-237)."""
    names, rows = [], []
    with open(path) as fh:
        for line in fh.read().splitlines()[2:]:
            parts = [p for p in line.split("  ") if p.strip()]
            if len(parts) >= 17:
                names.append(parts[0].strip())
                rows.append([float(x) for x in parts[1:17]])
    return names, np.array(rows)


class Oracle:
    """The plain-C restatement (oracle/afs_oracle.c)."""

    def __init__(self):
        path = os.path.join(HERE, "liboracle.so")
        src = os.path.join(HERE, "afs_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build("oracle")
        self.lib = C.CDLL(path)

    def expand_frame(self, frame):
        dt = np.dtype([("v", "<f8", (8,)), ("articulator", "<i4"), ("_pad", "<i4")])
        out = np.zeros(93, dtype=dt)
        self.lib.orc_expand_frame(_ptr(frame), _ptr(out))
        return out  # v = pos, area, length, volume, Mw, Bw, Kw, laterality

    def tl_spectrum(self, frames, f0_hz, num_freq, opts=None, type_=2, section=25):
        opts = opts or tl_options_default()
        n = len(frames)
        re = np.zeros((n, num_freq)); im = np.zeros((n, num_freq))
        rc = self.lib.orc_tl_spectrum(C.byref(opts), n, _ptr(np.ascontiguousarray(frames)), type_, section,
                                      C.c_double(f0_hz), num_freq, _ptr(re), _ptr(im))
        assert rc == 0, rc
        return re, im

    def tl_impulse_response(self, frames, sampling_rate=22050.0, length_exponent=9, opts=None):
        opts = opts or tl_options_default()
        n = len(frames)
        out = np.zeros((n, 1 << length_exponent))
        rc = self.lib.orc_tl_impulse_response(C.byref(opts), n, _ptr(np.ascontiguousarray(frames)),
                                              C.c_double(sampling_rate), length_exponent, _ptr(out))
        assert rc == 0, rc
        return out

    def tl_formants(self, frames, opts=None, sampling_rate=22050.0, spectrum_length=2048, max_formants=8):
        opts = opts or tl_options_default()
        n = len(frames)
        f = np.zeros((n, max_formants)); bw = np.zeros((n, max_formants))
        num = np.zeros(n, dtype=np.int32); flags = np.zeros(n, dtype=np.int32)
        rc = self.lib.orc_tl_formants(C.byref(opts), n, _ptr(np.ascontiguousarray(frames)), C.c_double(sampling_rate),
                                      spectrum_length, max_formants, _ptr(f), _ptr(bw), _ptr(num), _ptr(flags))
        assert rc == 0, rc
        return f, bw, num, flags

    def tds_synthesize(self, frames, glottis, frame_len, fs_tds_hz=44100.0, sr_factor=2, opts=None, want_flow=True,
                       dump_step=0, seeds=None):
        opts = opts or tds_options_default()
        frames = np.ascontiguousarray(frames)
        B, F = frames.shape
        frame_len = np.ascontiguousarray(frame_len, dtype=np.int32)
        n_out = int(frame_len.sum())
        audio = np.zeros((B, n_out))
        flow = np.zeros((B, n_out * sr_factor)) if want_flow else None
        st = np.zeros(473) if dump_step else None
        g = np.ascontiguousarray(glottis, dtype=np.float64)
        sd = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint32)
        rc = self.lib.orc_tds_synthesize(C.byref(opts), B, F, _ptr(frame_len), _ptr(frames), _ptr(g),
                                         C.c_double(fs_tds_hz), sr_factor, _ptr(sd), _ptr(audio), _ptr(flow), _ptr(st),
                                         dump_step)
        assert rc == 0, rc
        if dump_step:
            return audio, flow, st
        return audio, flow

    def tds_synthesize_probed(self, frames, glottis, frame_len, sections, fs_tds_hz=44100.0, sr_factor=2, opts=None, seeds=None):
        """(audio [B, n_out], pressure, inflow, outflow [B, n_probes, steps])."""
        opts = opts or tds_options_default()
        frames = np.ascontiguousarray(frames)
        B, F = frames.shape
        frame_len = np.ascontiguousarray(frame_len, dtype=np.int32)
        n_out = int(frame_len.sum()); steps = n_out * sr_factor
        sec = np.ascontiguousarray(sections, dtype=np.int32)
        audio = np.zeros((B, n_out)); pr = np.zeros((B, len(sec), steps)); fi = np.zeros_like(pr); fo = np.zeros_like(pr)
        g = np.ascontiguousarray(glottis, dtype=np.float64)
        sd = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint32)
        rc = self.lib.orc_tds_synthesize_probed(C.byref(opts), B, F, _ptr(frame_len), _ptr(frames), _ptr(g), C.c_double(fs_tds_hz),
                                                sr_factor, _ptr(sd), _ptr(audio), len(sec), _ptr(sec), _ptr(pr), _ptr(fi), _ptr(fo))
        assert rc == 0, rc
        return audio, pr, fi, fo

    def tl_mean_flow(self, frames, lung_pressure_dPa):
        n = len(frames)
        flow = np.zeros(n); sec = np.zeros(n, dtype=np.int32)
        rc = self.lib.orc_tl_mean_flow(n, _ptr(np.ascontiguousarray(frames)), C.c_double(lung_pressure_dPa), _ptr(flow), _ptr(sec))
        assert rc == 0, rc
        return flow, sec

    def chebyshev_lowpass(self, cutoff_ratio, poles):
        a = np.zeros(poles + 1); b = np.zeros(poles + 1)
        self.lib.orc_chebyshev_lowpass(C.c_double(cutoff_ratio), poles, _ptr(a), _ptr(b))
        return a, b
