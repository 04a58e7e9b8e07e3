"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/afs_b200.h
declares, refuses to compute without a device, and its host helpers match the reference."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import areafunctionsynthesis_b200 as afs
from conftest import ROOT, load_golden


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "afs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(afs_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_in_tree_and_exports_the_header():
    from areafunctionsynthesis_b200.build import build_library
    build_library()
    lib = afs.load_library()
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/afs_b200.h but not exported"
    assert lib.afs_version() == 200


def test_struct_layouts_match_the_header():
    assert afs.FRAME_DTYPE.itemsize == 1184
    assert C.sizeof(afs.TlOptions) == 56
    assert C.sizeof(afs.TdsOptions) == 8 * 4 + 8 + 16 * 8 + 8 + 4 + 4
    assert C.sizeof(afs.TdsProbes) == 4 + 8 * 4 + 4 + 3 * 8  # n_probes, sections, padding, three pointers
    o = afs.tl_options_default()
    assert (o.radiation, o.boundary_layer, o.soft_walls, o.lumped_elements, o.paranasal_sinuses,
            o.piriform_fossa, o.static_pressure_drops, o.lung_pressure_dPa) == (3, 1, 1, 1, 1, 0, 1, 0.0)
    t = afs.tds_options_default()
    assert (t.turbulence_losses, t.soft_walls, t.generate_noise_sources, t.radiation_from_skin, t.piriform_fossa,
            t.inner_length_corrections) == (1, 1, 1, 1, 0, 1)  # TdsModel.cpp:35-44
    assert t.glottis_static[14] == 129.0 and t.glottis_static[15] == 125.51
    assert (t.flow_source_section, t.flow_source_cm3_s, t.reserved) == (-1, 0.0, 0)  # TdsModel.cpp:84
    f = afs.default_frames(2)
    assert np.all(f["area_cm2"] == 4.0) and np.all(f["articulator"] == 4) and np.all(f["glottis_area_cm2"] == 0.1)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(afs.AfsError, match="no CUDA device"):
        afs.Context(0)
    lib = afs.load_library()
    # compute entry points refuse a null context instead of computing on the host
    rc = lib.afs_tl_sweep(None, None, 1, None, 2, 25, 10.0, 16, None, None)
    assert rc == 2  # AFS_ERR_BAD_CONTEXT


def test_chebyshev_designer_matches_reference():
    g = load_golden("chebyshev.npz")
    a, b = afs.chebyshev_lowpass(7000.0 / 44100.0, 8)
    assert np.array_equal(a, g["a44"]) and np.array_equal(b, g["b44"])
    a, b = afs.chebyshev_lowpass(7000.0 / 22050.0, 8)
    assert np.array_equal(a, g["a22"]) and np.array_equal(b, g["b22"])


def test_area_function_rasteriser_matches_reference(shapes):
    n = len(shapes["names"])
    mine = afs.area_function_to_frames(shapes["params"])
    ref = shapes["frames"][:n]
    for k in ("area_cm2", "length_cm", "articulator", "incisor_pos_cm", "laterality"):
        assert np.array_equal(mine[k], ref[k]), k


def test_bad_arguments_are_reported_not_printed():
    lib = afs.load_library()
    assert lib.afs_chebyshev_lowpass(0.2, 7, None, None) == 3
    assert lib.afs_area_function_to_frame(None, None) == 3
    assert b"bad argument" in lib.afs_status_string(3)


def test_int16_conversion_and_wav_writer(tmp_path):
    """Synthesizer.cpp:955-965 (sample * SHRT_MAX truncated, clipped outside [-1, 1]) and the on-disk output format."""
    import wave
    x = np.array([0.0, 0.5, -0.5, 1.0, -1.0, 1.5, -1.5, 0.99999, -0.99999, 1e-6, -3.0517578125e-05, 0.25000763])
    pcm = afs.audio_to_int16(x)
    want = np.trunc(x * 32767.0)
    want[x > 1.0] = 32767; want[x < -1.0] = -32768
    assert np.array_equal(pcm, want.astype(np.int16))
    rng = np.random.default_rng(0)
    stereo = (rng.standard_normal((1000, 2)) * 8000).astype(np.int16)
    path = str(tmp_path / "out.wav")
    afs.write_wav_int16(path, stereo, 22050, n_channels=2)
    with wave.open(path, "rb") as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (2, 2, 22050, 1000)
        assert np.array_equal(np.frombuffer(w.readframes(1000), dtype="<i2").reshape(1000, 2), stereo)
    with pytest.raises(afs.AfsError):
        afs.write_wav_int16(str(tmp_path / "no_such_dir" / "x.wav"), stereo, 22050)


def test_lf_pulse_train_matches_reference():
    """Host side of afs_synthesize_vowels_lf: LfPulse::getPulse / TimeFunction::getValue / flutter
    (LfPulse.cpp:37-245, TimeFunction.cpp:77-133, Synthesizer.cpp:667-771) -- bit-identical to the pulse signal of the
    reference run that produced tests/golden/lf_vowels.npz."""
    g = load_golden("lf_vowels.npz")
    for is_long in (1, 0):
        for k in range(len(g["lf"])):
            flow = afs.lf_glottal_flow(g["lf"][k], bool(is_long))
            assert np.array_equal(flow[::4], g[f"flow_{is_long}"][k]), (is_long, k)
    with pytest.raises(afs.AfsError):
        afs.lf_glottal_flow([300.0, 5.0, 0.5, 3.0, 0.02])  # pulse longer than the reference's ring buffer


def test_target_sequence_trajectory_matches_reference():
    """Host side of afs_play_target_sequences: the per-sample parameter / glottis trajectories of
    Synthesizer::playTargetSequence (Synthesizer.cpp:1286-1411), bit-identical to the reference run behind
    tests/golden/target_sequence.npz."""
    g = load_golden("target_sequence.npz")
    prm, gl = afs.target_sequence_trajectory(g["shapes"][0], g["stationary"], g["transition"], g["glottis"])
    assert len(prm) == g["audio_noise0"].shape[1] == 9702
    assert np.array_equal(prm[::7], g["params"]) and np.array_equal(gl[::7], g["glottis_traj"])
    assert gl[0, 1] == 0.0 and gl[3000, 1] > 7999.0 and abs(gl[0, 0] - 100.0) < 1e-9  # silence first, F0 contour from 100 Hz
