"""The C++ facade (include/afs_b200.hpp) compiled as a reference-style caller against libafs_b200.so."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

EXE = os.path.join(ROOT, "tests", "cpp", "facade_test")


def build_exe():
    from areafunctionsynthesis_b200.build import build_library
    build_library()
    lib_dir = os.path.join(ROOT, "areafunctionsynthesis_b200")
    src = os.path.join(ROOT, "tests", "cpp", "facade_test.cpp")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(lib_dir, "libafs_b200.so"))):
        subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), src, "-L", lib_dir,
                        "-lafs_b200", f"-Wl,-rpath,{lib_dir}", "-o", EXE], check=True)
    return EXE


def test_facade_compiles_and_refuses_to_run_without_a_device():
    import torch
    exe = build_exe()
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    if not torch.cuda.is_available():
        assert r.stdout.startswith("error 1 "), r.stdout  # AFS_ERR_NO_DEVICE: no CPU fallback


@pytest.mark.gpu
def test_facade_matches_ctypes_path(ctx, shapes):
    exe = build_exe()
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = dict((ln.split()[0], ln.split()[1:]) for ln in r.stdout.strip().splitlines())
    fr = shapes["frames"][[0, 0]].copy()
    fr["velum_area_cm2"][1] = 0.5
    f, bw, num, flags = ctx.tl_formants(fr)
    assert int(lines["formants"][0]) == num[0]
    np.testing.assert_allclose([float(x) for x in lines["formants"][1:4]], f[0, :3], atol=1e-6)
    assert int(lines["formants"][5]) == 1 and abs(float(lines["formants"][6]) - f[1, 0]) < 1e-6
    re, im = ctx.tl_sweep(fr[:1], 22050.0 / 2048, 938)
    assert int(lines["spectrum"][0]) == 938 and float(lines["spectrum"][1]) == re[0, 0]
    h = ctx.tl_impulse_response(fr, 22050.0, 9)
    assert int(lines["impulse"][0]) == h.size and float(lines["impulse"][1]) == h[0, 100] and float(lines["impulse"][2]) == h[1, 100]
    o = ctx.optimize_formants_vowel(shapes["params"][0], [650.0, 1300.0, 2600.0], 1.0, 0.0, None, 3)
    np.testing.assert_allclose([float(x) for x in lines["optimizer"]], [o["final_error"][0], o["params"][0, 2], o["params"][0, 8]], rtol=1e-9)
    import areafunctionsynthesis_b200 as afs
    pcm = ctx.synthesize_vowels_lf(fr, [[300.0, 120.0, 0.5, 3.0, 0.02]] * 2, False)
    assert int(lines["lfvowels"][0]) == pcm.size and int(lines["lfvowels"][1]) == int((pcm.astype(np.int64) ** 2).sum())
    p0 = shapes["params"][0]
    sh = np.stack([p0 * np.where(np.arange(16) == 7, 1.0 + 0.1 * k, 1.0) for k in range(4)])
    ta, _ = ctx.play_target_sequences(sh, [0.02, 0.01, 0.01, 0.02], [0.02, 0.02, 0.02], [120, 8000, 0.01, 0.01, 0, -40], 22050.0, 2,
                                      opts=afs.tds_options_default(noise=False))
    assert int(lines["targets"][0]) == ta.size
    assert abs(float(lines["targets"][1]) - float((ta ** 2).sum())) <= 1e-9 * float((ta ** 2).sum())
    gl = np.array([[[120, 8000, 0.01, 0.01, 0, -40]] * 2, [[150, 8000, 0.01, 0.01, 0, -40]] * 2], dtype=float)
    a, _, st = ctx.tds_synthesize(np.stack([fr, fr], axis=1), gl, [1102], 44100.0, 2)
    assert int(lines["audio"][0]) == a.size and int(lines["audio"][3]) == 0
    assert abs(float(lines["audio"][1]) - float((a ** 2).sum())) <= 1e-9 * float((a ** 2).sum())
    # setFlowSource + taps, getMeanFlow, the device group
    o = afs.tds_options_default(); o.flow_source_section = 30; o.flow_source_cm3_s = 50.0
    a2, _, taps = ctx.tds_synthesize_probed(np.stack([fr, fr], axis=1), gl, [200], [24, 40], opts=o)
    assert int(lines["probes"][0]) == taps["pressure"].size
    ep, ei = float((taps["pressure"] ** 2).sum()), float((taps["inflow"] * taps["outflow"]).sum())
    assert abs(float(lines["probes"][1]) - ep) <= 1e-9 * ep and abs(float(lines["probes"][2]) - ei) <= 1e-9 * abs(ei)
    mf, mc = ctx.tl_mean_flow(fr[:1], 8000.0)
    assert float(lines["meanflow"][0]) == mf[0] and int(lines["meanflow"][1]) == mc[0]
    a3, _, _ = ctx.tds_synthesize(np.stack([fr, fr], axis=1), gl, [1102], 44100.0, 2)  # default options, default seed
    # two pieces of 551 samples on models kept across calls == one piece of 1102 (constant frames)
    assert int(lines["stream"][0]) == a3.size
    assert abs(float(lines["stream"][1]) - float((a3 ** 2).sum())) <= 1e-9 * float((a3 ** 2).sum())
    import torch
    assert int(lines["group"][0]) == torch.cuda.device_count() and int(lines["group"][1]) == a3.size
    assert abs(float(lines["group"][2]) - float((a3 ** 2).sum())) <= 1e-12 * float((a3 ** 2).sum())
