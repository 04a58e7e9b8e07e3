#!/usr/bin/env python
"""bench.py -- headline benchmark of the area-function-to-speech hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric (BASELINE.json): TDS simulated audio-seconds per wall-second at batch 4096 (headline) and TL
transfer functions per second at 2048 bins (reported in the same JSON line under "tl").

A "step" is one pass of the hot path over one batch of synthetic input:
  TDS  4096 independent utterances per GPU, each a 1 s vowel-to-vowel transition simulated at 44.1 kHz
       with the nasal branch open (velum U[0,1] cm^2 per endpoint) and the subglottal tract (always
       part of the 93-section network), the reference's DEFAULT options (TdsModel.cpp:35-44, i.e.
       generateNoiseSources = true; every utterance owns a seeded random() stream) -- BASELINE configs[3]
       ("workload": "tds_nasal_subglottal_b4096").  N > 1: every rank gets its own 4096 utterances
       ("scaling": "weak", no data-path collective); the same JSON line carries `strong`: BASELINE configs[3]
       as written, 4096 utterances IN TOTAL sharded over the N ranks.
  TL   4096 random parametric area functions x 2048 bins -- BASELINE configs[1].
Other BASELINE configs in the same line: `c1_default_a` (configs[0]: Default.params /a/, one TL spectrum + 1 s TDS,
reference on one core beside the batch-of-one latency here), `oral_b1024` (configs[2]), `dataset_sweep` (configs[4]:
65 536 parameter sets in total sharded over the N ranks).
`parity` (N = 1): the audio of the first utterances of the ACTUAL bench batch against the compiled reference run on the
host cores, 1 s each, noise sources off and on (the reference seeded with srand(seed_b) per utterance).

`value` is device-resident throughput (inputs already in HBM), timed with CUDA events on the launching
stream; `e2e` is the same metric through the host-buffer C-ABI call (H2D + kernel + D2H inside the timed
region).  `roofline` is against the FP64 FMA peak measured live by afs_probe_fp64_tflops (both kernels are
FP64-pipe bound scalar recurrences) with W_tds = 1.0e4 flop per tract-step and W_tl = 8.0e3
flop per (area function, bin).  `cpu_baseline` times the compiled reference (oracle/_ref) on the host cores.

Further objects of the same JSON line (rank 0; the small legs only at N = 1): `noise_sources_on` (the TDS batch with
the reference's default noise sources), `gather` (NCCL all-gather of the audio after the compute),
`tl.fast_arithmetic` (afs_tl_options.arithmetic = FAST, with its error against the bit-identical output of this run),
`chunked` (the headline batch in five pieces through afs_tds_stream_*: models kept on the device between calls),
`raster` (K4), `formant_optimizer`, `lf_vowels`, `target_sequences`, `dataset_sweep` (one GPU's share of BASELINE
configs[4]) -- each with the reference's CPU figure beside it where the reference has that path.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W_TDS = 1.0e4  # algorithmic FP64 flop per tract-step, noise sources off
W_TDS_NOISE = 1.06e4  # ... with the noise sources (+0.6 k)
W_TL = 8.0e3   # algorithmic FP64 flop per (area function, bin)
FP64_NOMINAL_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12  # 148 SMs x 64 DFMA/clk x 2 flop x 1.965 GHz = 37.2
W_TL_FAST = 4.4e3  # FP64 flop per (area function, bin) the ladder formulation of K2f executes (84 sections x 43 + sinuses + algebra)
B_TDS = 4096
T_SEC = 1.0  # overridden by --utt-seconds (profiling only)
FS = 44100.0
SR_FACTOR = 2
B_TL, K_TL = 4096, 2048
F0_TL = 22050.0 / 8192
VOWELS = ["a:", "e:", "i:", "o:", "u:", "E:", "2:", "y:", "I", "E", "a", "O", "9", "Y", "U", "@"]


def make_tds_batch(n, seed):
    """Vowel-to-vowel transitions between the 16 vowel shapes of Default.params (frames stored as a
    fixture under tests/golden/), velum U[0,1] cm^2 per endpoint, F0 U[90,220] Hz gliding, P = 8000 dPa."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "default_shapes.npz"))
    names = [str(x) for x in g["names"]]
    idx = np.array([names.index(v) for v in VOWELS])
    rng = np.random.default_rng(seed)
    a = rng.integers(0, len(idx), size=n)
    b = (a + 1 + rng.integers(0, len(idx) - 1, size=n)) % len(idx)
    fr = g["frames"][np.stack([idx[a], idx[b]], axis=1)].copy()
    fr["velum_area_cm2"] = rng.random((n, 2))
    gl = np.tile(np.array([120.0, 8000.0, 0.01, 0.01, 0.0, -40.0]), (n, 2, 1))
    gl[:, :, 0] = rng.uniform(90.0, 220.0, size=(n, 2))
    return fr, gl


def make_tl_batch(n, seed):
    import areafunctionsynthesis_b200 as afs
    rng = np.random.default_rng(seed)
    lo = np.array([-0.2, 0.67, 0.015, 1.92, 1.82, 0.37, 6.36, 0.087, 0.31, 10.97, 0.092, 0.0, 16.33, 1.05, 16.53, 0.107])
    hi = np.array([2.63, 2.05, 5.22, 6.43, 11.55, 6.96, 16.44, 2.50, 5.21, 18.82, 14.33, 9.33, 20.45, 6.12, 21.52, 4.40])
    p = lo + (hi - lo) * rng.random((n, 16))
    p[:, [0, 3, 6, 9, 12, 14]] = np.sort(p[:, [0, 3, 6, 9, 12, 14]], axis=1)
    p[:, 0] = np.maximum(p[:, 0], 0.0)
    f = afs.area_function_to_frames(p)
    f["glottis_area_cm2"] = 0.0
    f["velum_area_cm2"] = np.where(rng.random(n) < 0.5, rng.random(n), 0.0)
    return f


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.th.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


# ---------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation (oracle/_ref), all host cores
# ---------------------------------------------------------------------------------------------------
def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def bench_seeds(n, rank=0):
    """srand() seed of utterance b of rank r's bench batch."""
    return (np.arange(1, n + 1, dtype=np.int64) + 7919 * rank).astype(np.uint32)


def _ref_worker(job):
    """One worker PROCESS of the reference pool: the compiled reference (oracle/_ref) on a slice of utterances, one
    after the other on one core.  Processes, not threads: with generateNoiseSources = true the reference draws from
    libc rand() (TdsModel.cpp:1690-1696), which is per-process state."""
    fr_bytes, gl, seeds, n_out, noise, want_audio = job
    from oracle.pyoracle import FRAME_DTYPE, Oracle, Ref, tds_options_default
    fr = np.frombuffer(fr_bytes, dtype=FRAME_DTYPE).reshape(len(gl), -1)
    opts = tds_options_default(noise=noise)
    try:
        ref, kind = Ref(44), "reference"
    except (FileNotFoundError, OSError):
        ref, kind = Oracle(), "port"
    t0 = time.time()
    if kind == "reference":
        audio, _ = ref.tds_synthesize(fr, gl, [n_out], opts=opts, seeds=seeds, want_flow=False, threads=1)
    else:
        audio, _ = ref.tds_synthesize(fr, gl, [n_out], FS, SR_FACTOR, opts=opts, seeds=seeds, want_flow=False)
    t1 = time.time()
    return t0, t1, kind, (audio if want_audio else None)


class RefPool:
    """Persistent pool of worker processes (spawned: nothing of CUDA / NCCL is inherited)."""

    def __init__(self, workers):
        import multiprocessing as mp
        from concurrent.futures import ProcessPoolExecutor
        self.workers = workers
        self.ex = ProcessPoolExecutor(max_workers=workers, mp_context=mp.get_context("spawn"))

    def run(self, fr, gl, seeds, n_out, noise, want_audio=False):
        """Utterances split into `workers` contiguous slices.  Returns (audio-s/s, wall s, kind, audio or None)."""
        n = len(fr)
        cuts = np.linspace(0, n, min(self.workers, n) + 1).astype(int)
        jobs = [(np.ascontiguousarray(fr[lo:hi]).tobytes(), np.ascontiguousarray(gl[lo:hi]),
                 np.ascontiguousarray(seeds[lo:hi]), n_out, bool(noise), want_audio) for lo, hi in zip(cuts[:-1], cuts[1:])]
        res = list(self.ex.map(_ref_worker, jobs))
        wall = max(r[1] for r in res) - min(r[0] for r in res)
        audio = np.concatenate([r[3] for r in res]) if want_audio else None
        return n * n_out * SR_FACTOR / FS / wall, wall, res[0][2], audio

    def close(self):
        self.ex.shutdown()


def cpu_reference_tl(n_threads, n_tubes):
    from oracle.pyoracle import Oracle, Ref
    f = make_tl_batch(n_tubes, 20260001)
    try:
        ref = Ref(22)
        t0 = time.perf_counter()
        ref.tl_spectrum(f, 8192, threads=n_threads)
        return n_tubes / (time.perf_counter() - t0), "reference", n_threads
    except (FileNotFoundError, OSError):
        orc = Oracle()
        t0 = time.perf_counter()
        orc.tl_spectrum(f[:8], F0_TL, K_TL)
        return 8 / (time.perf_counter() - t0), "port", 1


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    utt = 0.25  # seconds of audio per utterance per step: bounded sample of the 1 s workload
    n_out = int(round(utt * FS / SR_FACTOR))
    fr, gl = make_tds_batch(cores, 20260003)
    seeds = bench_seeds(cores)
    pool = RefPool(cores)
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt, kind, _ = pool.run(fr, gl, seeds, n_out, noise=True)
        if i >= args.warmup:
            vals.append((v, dt))
    pool.close()
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([d for _, d in vals]) * 1e3)
    sample = (f"{cores} utterances x {utt} s @44.1 kHz per step, one per core ({cores} worker processes, {kind}), "
              "default options (noise sources on)")
    line = {"impl": "reference", "metric": "tds_audio_seconds_per_second", "value": value,
            "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "tds_nasal_subglottal_b4096", "batch_per_gpu": B_TDS, "utterance_s": T_SEC,
                       "fs_tds_hz": FS, "noise_sources": True, "sample": sample},
            "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit_json_line(line)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def traffic_of(key):
    """DRAM bytes per launch of the dominant kernel from this round's `ncu --set full` capture (profiles/traffic.json,
    written by tools/ncu_summary.py from the .ncu-rep); null when the kernel has not been captured since it changed."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[key]
        return {"traffic": t["bytes"], "traffic_source": t["source"]}
    except Exception:
        return {"traffic": None, "traffic_source": "no capture of the current kernel"}


def config1_default_a(ctx):
    """BASELINE configs[0]: Default.params /a:/ -- one TL transfer function (N = 8192, K = 2048) and 1 s of TDS audio at
    44.1 kHz, the reference on ONE core beside the batch-of-one latency of the device path (host API, copies included)."""
    import areafunctionsynthesis_b200 as afs
    from oracle.pyoracle import Ref, tds_options_default as ref_opts
    g = np.load(os.path.join(ROOT, "tests", "golden", "default_shapes.npz"))
    k = [str(x) for x in g["names"]].index("a:")
    f = g["frames"][k:k + 1].copy()
    f["glottis_area_cm2"] = 0.0
    out = {"workload": "Default.params a: -- FLOW_SOURCE_TF 2048 bins + 1 s TDS at 44.1 kHz (F0 120 Hz, 8000 dPa)"}
    try:
        r22, r44 = Ref(22), Ref(44)
    except (FileNotFoundError, OSError):
        return None
    t0 = time.perf_counter(); rre, rim = r22.tl_spectrum(f, 8192); t_ref_tl = time.perf_counter() - t0
    ctx.tl_sweep(f, F0_TL, K_TL)
    t0 = time.perf_counter(); re, im = ctx.tl_sweep(f, F0_TL, K_TL); t_tl = time.perf_counter() - t0
    fr2 = np.stack([g["frames"][k], g["frames"][k]])[None]
    gl2 = np.tile(np.array([120.0, 8000.0, 0.01, 0.01, 0.0, -40.0]), (1, 2, 1))
    n_out = int(round(FS / SR_FACTOR))
    sd = np.array([1234], dtype=np.uint32)
    tds = {}
    for name, noise in (("noise_sources_on", True), ("noise_sources_off", False)):
        t0 = time.perf_counter()
        ra, _ = r44.tds_synthesize(fr2, gl2, [n_out], opts=ref_opts(noise=noise), seeds=sd, want_flow=False)
        t_ref = time.perf_counter() - t0
        o = afs.tds_options_default(noise=noise)
        ctx.tds_synthesize(fr2, gl2, [n_out], FS, SR_FACTOR, opts=o, seeds=sd)
        t0 = time.perf_counter()
        a, _, _ = ctx.tds_synthesize(fr2, gl2, [n_out], FS, SR_FACTOR, opts=o, seeds=sd)
        t_dev = time.perf_counter() - t0
        tds[name] = {"reference_s": t_ref, "device_s": t_dev, "rel_err": float(np.max(np.abs(a - ra)) / np.max(np.abs(ra)))}
    nb = rre.shape[1]
    out["tl"] = {"reference_s": t_ref_tl, "device_s": t_tl, "bins_compared": int(nb),
                 "bit_identical": bool(np.array_equal(re[:, :nb], rre) and np.array_equal(im[:, :nb], rim))}
    out["tds"] = tds
    out["cores"] = 1
    return out


def run_single_process(args):
    """`--single-process`: N GPUs of the box from ONE process through the C ABI (afs_group_*, csrc/afs_multi.cpp) -- what
    a C++ caller of the drop-in boundary gets: contiguous shards, one host thread + stream per device, every device
    copying its rows straight into the caller's (pinned) arrays.  Host buffers in, host buffers out: this is an e2e
    number by construction; timed with the wall clock around the call."""
    import areafunctionsynthesis_b200 as afs
    n_dev = args.gpus
    grp = afs.Group(list(range(n_dev)))
    B = n_dev * B_TDS
    n_out = int(round(T_SEC * FS / SR_FACTOR))
    fr, gl = make_tds_batch(B, 20260003)
    seeds = bench_seeds(B)
    audio = afs.pinned_empty((B, n_out))
    out = {}
    for name, noise in (("noise_sources_on", True), ("noise_sources_off", False)):
        o = afs.tds_options_default(noise=noise)
        for _ in range(max(1, args.warmup)):
            grp.tds_synthesize(fr, gl, [n_out], FS, SR_FACTOR, opts=o, seeds=seeds, audio=audio)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            _, _, st = grp.tds_synthesize(fr, gl, [n_out], FS, SR_FACTOR, opts=o, seeds=seeds, audio=audio)
        dt = (time.perf_counter() - t0) / args.steps
        assert not st.any()
        out[name] = {"value": B * T_SEC / dt, "unit": "audio-s/s", "ms_per_step": dt * 1e3}
    f_tl = make_tl_batch(n_dev * B_TL, 20260001)
    grp.tl_sweep(f_tl, F0_TL, K_TL)
    t0 = time.perf_counter()
    for _ in range(3):
        grp.tl_sweep(f_tl, F0_TL, K_TL)
    tl_dt = (time.perf_counter() - t0) / 3
    line = {"metric": "tds_audio_seconds_per_second", "value": out["noise_sources_on"]["value"], "unit": "audio-s/s", "n_gpus": n_dev,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": out["noise_sources_on"]["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "tds_nasal_subglottal_b4096", "batch_per_gpu": B_TDS, "utterance_s": T_SEC, "fs_tds_hz": FS,
                       "noise_sources": True, "parallelism": f"ONE process, afs_group over {n_dev} GPU(s): contiguous shards, one host "
                                                             "thread per device, no collective; host buffers (pageable inputs, pinned output)"},
            "e2e": {"value": out["noise_sources_on"]["value"], "unit": "audio-s/s",
                    "h2d_bytes_per_step": int(fr.nbytes + gl.nbytes + seeds.nbytes), "d2h_bytes_per_step": int(audio.nbytes)},
            "noise_sources_off": out["noise_sources_off"],
            "tl": {"metric": "tl_spectra_per_second_2048_bins", "value": n_dev * B_TL / tl_dt, "unit": "spectra/s",
                   "note": "afs_group_tl_sweep, pageable host buffers"},
            "single_process": True}
    emit_json_line(line)
    afs.pinned_free(audio)
    grp.close()


def run_ours(args):
    import torch
    import torch.distributed as dist

    import areafunctionsynthesis_b200 as afs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = afs.Context(local)
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn, steps, warmup):
        """W untimed warm-up steps, then K steps, each bracketed by CUDA events on the launching stream; L2 is
        flushed (256 MiB write, untimed) between steps.  Returns total ms (max over ranks)."""
        for _ in range(warmup):
            fn()
        barrier()
        tot = 0.0
        for _ in range(steps):
            flush_buf.fill_(1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            tot += e0.elapsed_time(e1)
        barrier()
        return max_over_ranks(tot)

    # ---------------- TDS: headline (reference-default options: noise sources ON) -----------------------
    n_out = int(round(T_SEC * FS / SR_FACTOR))
    total_steps = n_out * SR_FACTOR
    fr, gl = make_tds_batch(B_TDS, 20260003 + rank)
    seeds = bench_seeds(B_TDS, rank)
    frame_len = np.array([n_out], dtype=np.int32)
    d_fr = torch.from_numpy(fr.view(np.uint8).reshape(B_TDS, -1)).to(dev)
    d_gl = torch.from_numpy(gl).to(dev)
    d_seeds = torch.from_numpy(seeds.view(np.int32)).to(dev)
    d_audio = torch.empty(B_TDS, n_out, dtype=torch.float64, device=dev)       # noise sources on
    d_audio_off = torch.empty(B_TDS, n_out, dtype=torch.float64, device=dev)   # noise sources off
    d_status = torch.zeros(B_TDS, dtype=torch.int32, device=dev)
    noisy = afs.tds_options_default(noise=True)    # TdsModel::TdsModel() defaults (TdsModel.cpp:35-44)
    quiet = afs.tds_options_default(noise=False)

    def tds_dev(n=B_TDS, opts=noisy, out=d_audio):
        ctx.tds_synthesize_device(d_fr, d_gl, n, 2, frame_len, out, stream, FS, SR_FACTOR, opts=opts,
                                  d_seeds=d_seeds, d_status=d_status)

    fp64_peak = ctx.probe_fp64_tflops()
    sampler = ClockSampler(local)
    for _ in range(args.warmup):
        tds_dev()
    barrier()
    l0 = ctx.launch_count
    sampler.start()
    ms_total = timed(tds_dev, args.steps, 0)
    clocks = sampler.stop()
    launches = ctx.launch_count - l0
    ms_step = ms_total / args.steps
    audio_s_per_step = world * B_TDS * T_SEC
    value = audio_s_per_step / (ms_step * 1e-3)
    assert int(d_status.sum().item()) == 0, "non-finite values in the TDS batch"
    tds_tflops = W_TDS_NOISE * B_TDS * total_steps / (ms_step * 1e-3) / 1e12  # this rank's kernels (pre-pass + time loop)

    # same batch with generateNoiseSources = false
    off_steps = max(1, min(args.steps, 3))
    off_ms = timed(lambda: tds_dev(opts=quiet, out=d_audio_off), off_steps, 1) / off_steps
    assert int(d_status.sum().item()) == 0, "non-finite values in the TDS batch (noise sources off)"
    off_tflops = W_TDS * B_TDS * total_steps / (off_ms * 1e-3) / 1e12

    # e2e: pinned host buffers -> C-ABI host call (H2D + kernels + D2H inside the timed region)
    h_fr = torch.from_numpy(fr.view(np.uint8).reshape(B_TDS, -1).copy()).pin_memory()
    h_gl = torch.from_numpy(gl.copy()).pin_memory()
    h_seeds = torch.from_numpy(seeds.view(np.int32).copy()).pin_memory()
    h_audio = torch.empty(B_TDS, n_out, dtype=torch.float64).pin_memory()
    h_status = torch.zeros(B_TDS, dtype=torch.int32).pin_memory()

    def tds_host(opts):
        rc = ctx.lib.afs_tds_synthesize(ctx.h, afs.api.C.byref(opts), B_TDS, 2, frame_len.ctypes.data,
                                        h_fr.data_ptr(), h_gl.data_ptr(), FS, SR_FACTOR, h_seeds.data_ptr(), h_audio.data_ptr(),
                                        None, h_status.data_ptr())
        assert rc == 0, rc

    def e2e_of(opts, steps):
        tds_host(opts)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            tds_host(opts)
        barrier()
        return audio_s_per_step / max_over_ranks((time.perf_counter() - t0) / steps)

    e2e_off = e2e_of(quiet, off_steps)
    assert torch.equal(h_audio, d_audio_off.cpu()), "host-path audio differs from the device-resident path (noise off)"
    e2e_value = e2e_of(noisy, args.steps)
    assert torch.equal(h_audio, d_audio.cpu()), "host-path audio differs from the device-resident path"

    # ---------------- chunked synthesis: the same batch in five pieces, the models kept on the device ---------
    # (afs_tds_stream_*: what a caller of the reference's Synthesizer does -- one synthesizeSignalTds call per control
    #  frame on a live object.  The intermediate control frames are the linear interpolation of the two end frames,
    #  i.e. the same trajectory up to rounding; host buffers, H2D + kernels + D2H per chunk inside the timed region.)
    N_CHUNKS = 5
    chunked = None
    if n_out % N_CHUNKS == 0:
        def lerp_frames(r):
            f = fr[:, 0].copy()
            for name in ("area_cm2", "length_cm", "laterality", "incisor_pos_cm", "velum_area_cm2", "aspiration_dB"):
                a, b = fr[name][:, 0], fr[name][:, 1]
                if name in ("area_cm2", "velum_area_cm2"):  # (Tube::setPharynxMouthGeometry / setVelumOpening clamp on the way in)
                    a, b = np.maximum(a, 1e-3), np.maximum(b, 1e-3)
                f[name] = (1.0 - r) * a + r * b
            return f, (1.0 - r) * gl[:, 0] + r * gl[:, 1]
        pieces = [lerp_frames(k / N_CHUNKS) for k in range(N_CHUNKS + 1)]
        h_cf = [torch.from_numpy(np.ascontiguousarray(f).view(np.uint8).reshape(B_TDS, -1).copy()).pin_memory() for f, _ in pieces]
        h_cg = [torch.from_numpy(np.ascontiguousarray(g)).pin_memory() for _, g in pieces]
        h_ca = [torch.empty(B_TDS, n_out // N_CHUNKS, dtype=torch.float64).pin_memory() for _ in range(N_CHUNKS)]
        st_h = afs.api.C.c_void_p()
        rc = ctx.lib.afs_tds_stream_create(ctx.h, afs.api.C.byref(noisy), B_TDS, FS, SR_FACTOR, h_seeds.data_ptr(), afs.api.C.byref(st_h))
        assert rc == 0, rc
        one_len = np.array([n_out // N_CHUNKS], dtype=np.int32)

        def run_chunks():
            ctx.lib.afs_tds_stream_reset(st_h)
            for k in range(N_CHUNKS + 1):  # (piece 0 primes)
                rc = ctx.lib.afs_tds_stream_synthesize(st_h, 1, one_len.ctypes.data, h_cf[k].data_ptr(), h_cg[k].data_ptr(),
                                                       h_ca[k - 1].data_ptr() if k else None, None, h_status.data_ptr())
                assert rc == 0, rc
        run_chunks()
        barrier()
        t0 = time.perf_counter()
        run_chunks()
        barrier()
        ch_dt = max_over_ranks(time.perf_counter() - t0)
        ctx.lib.afs_tds_stream_destroy(st_h)
        a_ch = torch.cat(h_ca, dim=1)
        ch_rel = float(((a_ch - h_audio).abs().amax(dim=1) / h_audio.abs().amax(dim=1)).max().item())
        chunked = {"workload": f"the headline batch in {N_CHUNKS} pieces of {T_SEC / N_CHUNKS:g} s through afs_tds_stream_* (models and "
                               "state resident on the device between calls), host buffers",
                   "value": audio_s_per_step / ch_dt, "unit": "audio-s/s", "chunks": N_CHUNKS,
                   "vs_one_call_e2e": audio_s_per_step / ch_dt / e2e_value,
                   "max_rel_vs_one_call": ch_rel}
        assert ch_rel <= 1e-6, ch_rel
        del h_cf, h_cg, h_ca, a_ch

    # the only collective of the path: one all-gather of the outputs, after the compute; timed separately
    digest = d_audio.abs().mean(dim=1)
    gather_ms = 0.0
    if world > 1:
        all_audio = torch.empty(world * B_TDS, n_out, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(all_audio, d_audio)  # warm-up (NCCL channel setup)
        torch.cuda.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.all_gather_into_tensor(all_audio, d_audio)
        e1.record()
        e1.synchronize()
        gather_ms = max_over_ranks(e0.elapsed_time(e1))
        digest = all_audio.abs().mean(dim=1)
        assert torch.equal(all_audio[rank * B_TDS:(rank + 1) * B_TDS], d_audio)
        del all_audio
    digest_mean = float(digest.mean().item())

    # ---------------- BASELINE configs[3] as written: 4096 utterances IN TOTAL, sharded over the ranks -----
    # (strong scaling; contiguous shards, no data-path collective.  Rank r runs the first hi - lo utterances of its own
    #  batch: same work as a slice of one global batch.)
    from areafunctionsynthesis_b200.sharding import shard_bounds
    lo_s, hi_s = shard_bounds(B_TDS, world)[rank]
    n_shard = hi_s - lo_s
    strong_steps = max(1, min(args.steps, 3))
    strong = {}
    for name, o, out in (("noise_sources_on", noisy, d_audio), ("noise_sources_off", quiet, d_audio_off)):
        ms = timed(lambda: tds_dev(n=n_shard, opts=o, out=out), strong_steps, 1) / strong_steps
        strong[name] = {"value": B_TDS * T_SEC / (ms * 1e-3), "unit": "audio-s/s", "ms_per_step": ms}
    assert int(d_status[:n_shard].sum().item()) == 0

    # ---------------- BASELINE configs[2]: 1024 oral vowel-to-vowel transitions per GPU ---------------------
    B_ORAL = 1024
    fr_o, gl_o = make_tds_batch(B_ORAL, 20260002 + rank)
    fr_o["velum_area_cm2"] = 0.0
    d_fr_o = torch.from_numpy(fr_o.view(np.uint8).reshape(B_ORAL, -1)).to(dev)
    d_gl_o = torch.from_numpy(gl_o).to(dev)
    oral = {}
    for name, o in (("noise_sources_on", noisy), ("noise_sources_off", quiet)):
        def oral_dev():
            ctx.tds_synthesize_device(d_fr_o, d_gl_o, B_ORAL, 2, frame_len, d_audio_off, stream, FS, SR_FACTOR, opts=o,
                                      d_seeds=d_seeds, d_status=d_status)
        ms = timed(oral_dev, strong_steps, 1) / strong_steps
        oral[name] = {"value": world * B_ORAL * T_SEC / (ms * 1e-3), "unit": "audio-s/s", "ms_per_step": ms}
    assert int(d_status[:B_ORAL].sum().item()) == 0
    tds_dev(opts=quiet, out=d_audio_off)  # (d_audio_off served as scratch above: restore the noise-off audio for `parity`)
    torch.cuda.synchronize()

    # ---------------- TL: second metric -----------------------------------------------------------
    f_tl = make_tl_batch(B_TL, 20260001 + rank)
    d_ftl = torch.from_numpy(f_tl.view(np.uint8).reshape(B_TL, -1)).to(dev)
    d_re = torch.empty(B_TL, K_TL, dtype=torch.float64, device=dev)
    d_im = torch.empty_like(d_re)

    def tl_dev():
        ctx.tl_sweep_device(d_ftl, B_TL, F0_TL, K_TL, d_re, d_im, stream)

    tl_steps = max(args.steps, 5)
    tl_ms = timed(tl_dev, tl_steps, 3) / tl_steps
    tl_value = world * B_TL / (tl_ms * 1e-3)
    tl_tflops = W_TL * B_TL * K_TL / (tl_ms * 1e-3) / 1e12
    # e2e: pinned host buffers -> C-ABI host call (H2D + kernel + D2H, the D2H of a chunk beside the next chunk's kernel)
    h_ftl = torch.from_numpy(f_tl.view(np.uint8).reshape(B_TL, -1).copy()).pin_memory()
    h_re = torch.empty(B_TL, K_TL, dtype=torch.float64).pin_memory()
    h_im = torch.empty(B_TL, K_TL, dtype=torch.float64).pin_memory()

    def tl_host(o):
        rc = ctx.lib.afs_tl_sweep(ctx.h, afs.api.C.byref(o), B_TL, h_ftl.data_ptr(), 2, 25, F0_TL, K_TL, h_re.data_ptr(),
                                  h_im.data_ptr())
        assert rc == 0, rc

    def tl_e2e_of(o):
        tl_host(o)
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            tl_host(o)
        barrier()
        return world * B_TL / max_over_ranks((time.perf_counter() - t0) / 3)

    tl_e2e = tl_e2e_of(afs.tl_options_default())
    assert torch.equal(h_re, d_re.cpu()) and torch.equal(h_im, d_im.cpu()), "host-path spectra differ from the device-resident path"

    # same sweep with afs_tl_options.arithmetic = FAST (K2f: ladder recurrences, 1e-9 class instead of bit-identical)
    fast = afs.tl_options_default(); fast.arithmetic = afs.TL_ARITHMETIC_FAST
    d_re_f = torch.empty_like(d_re); d_im_f = torch.empty_like(d_re)

    def tl_fast_dev():
        ctx.tl_sweep_device(d_ftl, B_TL, F0_TL, K_TL, d_re_f, d_im_f, stream, opts=fast)

    tlf_ms = timed(tl_fast_dev, tl_steps, 3) / tl_steps
    tlf_tflops = W_TL_FAST * B_TL * K_TL / (tlf_ms * 1e-3) / 1e12
    h_ref, h_fast = torch.complex(d_re, d_im), torch.complex(d_re_f, d_im_f)
    rel = ((h_fast - h_ref).abs() / h_ref.abs()).flatten()
    mag = ((h_fast.abs() - h_ref.abs()).abs() / h_ref.abs()).flatten()
    ph = torch.angle(h_fast * h_ref.conj()).abs().flatten()
    tlf_err = {"max_rel": float(rel.max()), "p999_rel": float(torch.quantile(rel[: 1 << 23], 0.999)),
               "magnitude": {"max_rel": float(mag.max()), "p999_rel": float(torch.quantile(mag[: 1 << 23], 0.999))},
               "phase_rad": {"max": float(ph.max()), "p999": float(torch.quantile(ph[: 1 << 23], 0.999))},
               "bar": "1e-9 for 99.9 % of the points, 1e-8 everywhere (the points above 1e-9 are where the reference's own "
                      "inverse(P[section]) * P[40] cancels: tools/tl_truth.py)",
               "against": "the bit-identical kernel's output of this run; see tests/test_tl_gpu.py::"
                          "test_fast_arithmetic_against_extended_precision for which of the two is nearer to the formulas"}
    del mag, ph
    del h_ref, h_fast, rel
    tlf_e2e = tl_e2e_of(fast)

    # ---------------- rasteriser: parameter sets -> frames on the device --------------
    B_RA = 65536
    rng = np.random.default_rng(20260004 + rank)
    lo = np.array([-0.2, 0.67, 0.015, 1.92, 1.82, 0.37, 6.36, 0.087, 0.31, 10.97, 0.092, 0.0, 16.33, 1.05, 16.53, 0.107])
    hi = np.array([2.63, 2.05, 5.22, 6.43, 11.55, 6.96, 16.44, 2.50, 5.21, 18.82, 14.33, 9.33, 20.45, 6.12, 21.52, 4.40])
    pr = lo + (hi - lo) * rng.random((B_RA, 16))
    pr[:, [0, 3, 6, 9, 12, 14]] = np.sort(pr[:, [0, 3, 6, 9, 12, 14]], axis=1)
    pr[:, 0] = np.maximum(pr[:, 0], 0.0)
    d_pr = torch.from_numpy(pr).to(dev)
    d_fra = torch.from_numpy(afs.default_frames(B_RA).view(np.uint8).reshape(B_RA, -1)).to(dev)

    def raster_dev():
        ctx.area_functions_to_frames_device(d_pr, d_fra, B_RA, stream)

    ra_ms = timed(raster_dev, 5, 3) / 5
    ra_value = world * B_RA / (ra_ms * 1e-3)
    ra_cpu = None
    if world == 1 and not args.no_cpu_baseline:
        t0 = time.perf_counter()
        host_fr = afs.area_function_to_frames(pr[:2048])
        ra_cpu = 2048 / (time.perf_counter() - t0)
        dev_fr = np.frombuffer(d_fra[:2048].cpu().numpy().tobytes(), dtype=afs.FRAME_DTYPE)
        assert np.array_equal(dev_fr["articulator"], host_fr["articulator"])
        assert np.max(np.abs(dev_fr["area_cm2"] - host_fr["area_cm2"]) / host_fr["area_cm2"]) <= 1e-12

    # ---------------- BASELINE configs[4] (C5): 65 536 parameter sets IN TOTAL sharded over the ranks ----------
    # parameter sets -> frames (K4) -> 2048-bin spectra (K2) + 0.5 s TDS audio at 44.1 kHz (K1, default options)
    from areafunctionsynthesis_b200.sweep import dataset_sweep
    lo_d, hi_d = shard_bounds(B_RA, world)[rank]
    n_ds = hi_d - lo_d
    ds_seeds = bench_seeds(n_ds, rank)
    ds = dataset_sweep(ctx, pr[:min(n_ds, 4096)], want_formants=False, tds_opts=noisy, seeds=ds_seeds[:min(n_ds, 4096)])  # warm-up
    del ds
    barrier()
    t0 = time.perf_counter()
    ds = dataset_sweep(ctx, pr[:n_ds], want_formants=False, tds_opts=noisy, seeds=ds_seeds)
    barrier()
    ds_s = max_over_ranks(time.perf_counter() - t0)
    assert int(ds["status"].sum().item()) == 0
    del ds
    torch.cuda.empty_cache()
    dataset = {"metric": "parameter_sets_per_second", "value": B_RA / ds_s, "unit": "sets/s", "seconds": ds_s,
               "config": {"workload": f"{B_RA} parameter sets in total, {n_ds} on this rank (BASELINE configs[4]) -> frames (K4) -> "
                                      "2048-bin spectra (K2, bit-identical arithmetic) + 0.5 s TDS audio at 44.1 kHz (K1, default "
                                      "options: noise sources on), all stages on the device; wall clock incl. the H2D of the parameters"}}

    # ---------------- formant optimiser and LF vowel synthesis (small legs, host API) ------
    extra = {}
    if world == 1:
        g = np.load(os.path.join(ROOT, "tests", "golden", "default_shapes.npz"))
        vowels = [i for i, nm in enumerate(g["names"]) if "(" not in str(nm)][:16]
        n_opt, runs_opt = 64, 10
        p0 = g["params"][[vowels[i % len(vowels)] for i in range(n_opt)]]
        tg = np.tile(np.array([[650.0, 1300.0, 2600.0], [300.0, 2100.0, 2900.0], [500.0, 1500.0, 2500.0], [380.0, 700.0, 2500.0]]), (n_opt // 4, 1))
        ctx.optimize_formants_vowel(p0, tg, 1.0, 0.0, None, 1)  # warm-up at full size (device scratch is grow-only)
        t0 = time.perf_counter()
        ro = ctx.optimize_formants_vowel(p0, tg, 1.0, 0.0, None, runs_opt)
        dt = time.perf_counter() - t0
        evals = int(ro["runs"].sum()) * 33 + 2 * n_opt
        extra["formant_optimizer"] = {"metric": "optimiser_runs_per_second", "value": float(ro["runs"].sum()) / dt, "unit": "runs/s",
                                      "config": {"workload": f"{n_opt} vowels x <= {runs_opt} greedy runs (33 candidates each), 938-bin formant grid"},
                                      "candidate_evaluations_per_second": evals / dt, "cpu_baseline": None}
        n_lf = 1024
        fr_lf = g["frames"][[vowels[i % len(vowels)] for i in range(n_lf)]]
        lf = np.tile(np.array([[300.0, 120.0, 0.5, 3.0, 0.02]]), (n_lf, 1)); lf[:, 1] = np.linspace(90.0, 220.0, n_lf)
        ctx.synthesize_vowels_lf(fr_lf, lf, True)  # warm-up at full size
        t0 = time.perf_counter()
        a_lf = ctx.synthesize_vowels_lf(fr_lf, lf, True)
        dt = time.perf_counter() - t0
        extra["lf_vowels"] = {"metric": "lf_vowels_per_second", "value": n_lf / dt, "unit": "vowels/s (0.65 s each, 22.05 kHz)",
                              "config": {"workload": f"{n_lf} long vowels, host API incl. pulse trains and copies"}, "cpu_baseline": None}
        n_ts = 256
        gt = np.load(os.path.join(ROOT, "tests", "golden", "target_sequence.npz"))
        sh_ts = np.tile(gt["shapes"], (n_ts // 3 + 1, 1, 1))[:n_ts]
        ctx.play_target_sequences(sh_ts, gt["stationary"], gt["transition"], gt["glottis"], 22050.0, 2, opts=quiet)  # warm-up
        t0 = time.perf_counter()
        a_ts, st_ts = ctx.play_target_sequences(sh_ts, gt["stationary"], gt["transition"], gt["glottis"], 22050.0, 2, opts=quiet)
        dt = time.perf_counter() - t0
        ts_s = a_ts.shape[1] / 22050.0
        assert not st_ts.any()
        extra["target_sequences"] = {"metric": "target_sequence_audio_seconds_per_second", "value": n_ts * ts_s / dt, "unit": "audio-s/s",
                                     "config": {"workload": f"{n_ts} sequences of 4 target shapes x {ts_s:.2f} s, one control frame per output sample "
                                                            "(parameter sets rasterised on the device), noise sources off, host API"},
                                     "cpu_baseline": None}
        if not args.no_cpu_baseline:
            try:
                from oracle.pyoracle import Ref, tds_options_default as ref_tds_opts
                t0 = time.perf_counter()
                ra, _, _ = Ref(44).play_target_sequence(gt["shapes"][0], gt["stationary"], gt["transition"], gt["glottis"],
                                                        opts=ref_tds_opts(noise=False))
                extra["target_sequences"]["cpu_baseline"] = {"value": ts_s / (time.perf_counter() - t0), "unit": "audio-s/s", "cores": 1,
                                                             "kind": "reference", "sample": "1 sequence"}
                assert np.max(np.abs(ra - a_ts[0])) <= 1e-6 * np.max(np.abs(ra)), "target sequence differs from the reference"
            except (FileNotFoundError, OSError):
                pass
        if not args.no_cpu_baseline:
            try:
                from oracle.pyoracle import Ref
                r22 = Ref(22)
                t0 = time.perf_counter()
                _, _, _, rr, _, _ = r22.optimize_formants_vowel(p0[0], tg[0], 1.0, 0.0, None, 2)
                extra["formant_optimizer"]["cpu_baseline"] = {"value": rr / (time.perf_counter() - t0), "unit": "runs/s", "cores": 1,
                                                              "kind": "reference", "sample": "1 vowel x 2 runs (Data.cpp loop over the reference classes)"}
                t0 = time.perf_counter()
                ref_a = [r22.synthesize_vowel_lf(fr_lf[i:i + 1], lf[i], 1)[0] for i in range(8)]
                extra["lf_vowels"]["cpu_baseline"] = {"value": 8 / (time.perf_counter() - t0), "unit": "vowels/s", "cores": 1,
                                                      "kind": "reference", "sample": "8 long vowels"}
                assert all(np.array_equal(ref_a[i], a_lf[i]) for i in range(8)), "LF vowels differ from the reference"
            except (FileNotFoundError, OSError):
                pass

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- CPU baseline, parity of the bench batch, BASELINE configs[0] (rank 0, N = 1 only) -------------
    cpu = None
    tl_cpu = None
    parity = None
    c1 = None
    if world == 1 and not args.no_cpu_baseline:
        cores = host_threads()
        pool = RefPool(cores)
        # the compiled reference on the FIRST utterances of this very batch (same frames, glottis parameters and seeds),
        # full length; its audio is the parity check of the headline configuration, its wall time the CPU baseline
        n_par = min(4 * cores, 64)
        v_on, dt_on, kind, ref_on = pool.run(fr[:n_par], gl[:n_par], seeds[:n_par], n_out, noise=True, want_audio=True)
        v_off, dt_off, _, ref_off = pool.run(fr[:n_par], gl[:n_par], seeds[:n_par], n_out, noise=False, want_audio=True)
        pool.close()
        cpu = {"value": v_on, "unit": "audio-s/s", "cores": cores, "kind": kind,
               "sample": f"{n_par} utterances x {T_SEC} s @44.1 kHz: the first {n_par} of the bench batch, default options (noise sources "
                         f"on, srand(seed_b) per utterance), {cores} worker processes, {dt_on:.1f} s wall",
               "noise_sources_off": {"value": v_off, "unit": "audio-s/s", "seconds": dt_off}}

        def rel_errs(dev_audio, ref):
            a = dev_audio[:len(ref)].cpu().numpy()
            return np.max(np.abs(a - ref), axis=1) / np.max(np.abs(ref), axis=1)

        e_on, e_off = rel_errs(d_audio, ref_on), rel_errs(d_audio_off, ref_off)
        parity = {"bar": 1e-6, "measure": "max_t |audio - reference| / max_t |reference| per utterance, 1 s at 44.1 kHz",
                  "against": kind, "n": int(n_par),
                  "noise_sources_on": {"max_rel": float(e_on.max()), "median_rel": float(np.median(e_on))},
                  "noise_sources_off": {"max_rel": float(e_off.max()), "median_rel": float(np.median(e_off))},
                  "ok": bool(e_on.max() <= 1e-6 and e_off.max() <= 1e-6)}
        tv, tkind, tcores = cpu_reference_tl(cores, 4 * cores)
        tl_cpu = {"value": tv, "unit": "spectra/s", "cores": tcores, "kind": tkind,
                  "sample": f"{4 * cores} area functions x 2048 bins"}
        c1 = config1_default_a(ctx)

    peaks = load_peaks()
    line = {
        "metric": "tds_audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "tds_nasal_subglottal_b4096", "batch_per_gpu": B_TDS, "utterance_s": T_SEC,
                   "fs_tds_hz": FS, "sr_factor": SR_FACTOR, "control_frames": 2, "noise_sources": True,
                   "options": "TdsModel::TdsModel() defaults (TdsModel.cpp:35-44)",
                   "parallelism": f"independent utterances sharded over {world} GPU(s), no data-path collective",
                   "l2": "flushed between timed steps (256 MiB write, untimed)"},
        "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": int(h_fr.numel() + h_gl.numel() * 8 + h_seeds.numel() * 4),
                "d2h_bytes_per_step": int(h_audio.numel() * 8 + h_status.numel() * 4)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "fp64", "achieved": tds_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": tds_tflops / fp64_peak, **traffic_of("tds_noise_on"),
                     "kernel": "tds_batch_kernel<noise> (+ tds_noise_runs_kernel: the 3.5 ms pre-pass is inside the timed step)",
                     "flop_per_tract_step": W_TDS_NOISE,
                     "peak_source": "afs_probe_fp64_tflops, measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
                     "peak_nominal": FP64_NOMINAL_TFLOPS, "frac_of_nominal": tds_tflops / FP64_NOMINAL_TFLOPS,
                     "hbm_gbs_algorithmic": B_TDS * n_out * 8 / (ms_step * 1e-3) / 1e9,
                     "hbm_gbs_peak_measured": peaks.get("hbm_gbs")},
        "cpu_baseline": cpu,
        "output_digest": digest_mean,
        "gather": {"collective": "all_gather of the audio [B, n_out] f64 over NCCL, after the compute", "ms": gather_ms,
                   "bytes_per_rank": int(B_TDS * n_out * 8),
                   "value_including_gather": audio_s_per_step / ((ms_step + gather_ms) * 1e-3)},
        **extra,
        "raster": {"metric": "area_function_parameter_sets_per_second", "value": ra_value, "unit": "sets/s", "ms_per_step": ra_ms,
                   "config": {"workload": "onedim_area_function_to_frames_65536", "sets_per_gpu": B_RA},
                   "cpu_baseline": None if ra_cpu is None else {"value": ra_cpu, "unit": "sets/s", "cores": 1, "kind": "port",
                                                                "sample": "2048 sets, host rasteriser of the C ABI (bit-identical to the reference on Default.params)"}},
        "noise_sources_off": {"value": audio_s_per_step / (off_ms * 1e-3), "unit": "audio-s/s", "ms_per_step": off_ms,
                              "steps": off_steps, "e2e": {"value": e2e_off, "unit": "audio-s/s"},
                              "roofline": {"bound": "fp64", "achieved": off_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
                                           "frac": off_tflops / fp64_peak, "flop_per_tract_step": W_TDS,
                                           "kernel": "tds_batch_kernel", **traffic_of("tds_noise_off")}},
        "parity": parity,
        "strong": {"workload": "BASELINE configs[3] as written: 4096 utterances in total, contiguous shards over the ranks",
                   "scaling": "strong", "utterances_this_rank": int(n_shard), **strong},
        "chunked": chunked,
        "oral_b1024": {"workload": "BASELINE configs[2]: 1024 oral vowel-to-vowel transitions x 1 s per GPU (velum closed)",
                       "scaling": "weak", **oral},
        "dataset_sweep": dataset,
        "c1_default_a": c1,
        "tl": {"metric": "tl_spectra_per_second_2048_bins", "value": tl_value, "unit": "spectra/s", "ms_per_step": tl_ms,
               "config": {"workload": "tl_sweep_4096x2048", "area_functions_per_gpu": B_TL, "bins": K_TL},
               "e2e": {"value": tl_e2e, "unit": "spectra/s", "h2d_bytes_per_step": int(f_tl.nbytes),
                       "d2h_bytes_per_step": int(2 * B_TL * K_TL * 8)},
               "roofline": {"bound": "fp64", "achieved": tl_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
                            "frac": tl_tflops / fp64_peak, "kernel": "tl_sweep_kernel", "flop_per_bin": W_TL,
                            "hbm_gbs_algorithmic": B_TL * K_TL * 16 / (tl_ms * 1e-3) / 1e9},
               "cpu_baseline": tl_cpu,
               "fast_arithmetic": {"value": world * B_TL / (tlf_ms * 1e-3), "unit": "spectra/s", "ms_per_step": tlf_ms,
                                   "e2e": {"value": tlf_e2e, "unit": "spectra/s"},
                                   "roofline": {"bound": "fp64", "achieved": tlf_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
                                                "frac": tlf_tflops / fp64_peak, "kernel": "tl_sweep_fast_kernel",
                                                "flop_per_bin": W_TL_FAST,
                                                "note": "flops the ladder formulation executes (DESIGN.md K2f), not the 8.0e3 of the reference-like formulation"},
                                   "error": tlf_err}},
    }
    emit_json_line(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def quiet_stdout():
    """Native libraries write to fd 1 (NCCL prints its version banner there): point fd 1 at stderr until the ONE JSON
    line is due, so that stdout carries nothing else."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit_json_line(line):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--single-process", action="store_true",
                    help="N GPUs from one process through afs_group_* (the C-ABI multi-device layer) instead of one rank per GPU")
    ap.add_argument("--utt-seconds", type=float, default=1.0, help="utterance length; != 1.0 is for profiling only")
    args = ap.parse_args()
    global T_SEC
    T_SEC = args.utt_seconds
    quiet_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.single_process:
        run_single_process(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
