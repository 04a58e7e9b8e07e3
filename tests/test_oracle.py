"""CPU tests: the oracle restatement (oracle/afs_oracle.c) against the reference's outputs.

The reference has no golden vectors of its own; the fixtures under tests/golden/ were produced by
the unmodified reference compiled from /root/reference (tools/make_golden.py).  Where the compiled
reference is available (build container, or oracle/_ref shipped to the GPU box) it is also run live.
"""
import numpy as np
import pytest

from conftest import have_ref, load_golden

F0_938 = 22050.0 / 2048
F0_2048 = 22050.0 / 8192


def test_frame_expansion_matches_reference_tube(oracle, shapes):
    if not have_ref():
        pytest.skip("compiled reference not available")
    from oracle.pyoracle import Ref
    r = Ref(22)
    for i in (0, 7, 60, 101):
        sec = r.frame_sections(shapes["frames"][i:i + 1])
        e = oracle.expand_frame(shapes["frames"][i:i + 1])["v"]
        assert np.array_equal(sec[:, 0], e[:, 1]) and np.array_equal(sec[:, 1], e[:, 2])
        assert np.array_equal(sec[:, 2], e[:, 3]) and np.array_equal(sec[:, 3], e[:, 0])
        assert np.array_equal(sec[:, 4:7], e[:, 4:7])


def test_tl_spectra_bit_identical_to_golden(oracle, shapes):
    g = load_golden("tl_reference.npz")
    fr = shapes["frames"]
    re, im = oracle.tl_spectrum(fr[g["sel938"]], F0_938, 938)
    assert np.array_equal(re, g["re938"]) and np.array_equal(im, g["im938"])
    re, im = oracle.tl_spectrum(fr[g["sel2048"]], F0_2048, 2048)
    assert np.array_equal(re, g["re2048"]) and np.array_equal(im, g["im2048"])


def test_tl_bin0_aliases_bin1(oracle, shapes):
    re, im = oracle.tl_spectrum(shapes["frames"][:1], F0_2048, 2048)
    assert re[0, 0] == re[0, 1] and im[0, 0] == im[0, 1]
    # SURVEY section 4 anchor for /a/
    assert re[0, 0] == 1.0003095278623926 and abs(im[0, 0] - -0.0025227522178727) < 1e-15


def test_tl_other_spectrum_types_and_options(oracle, shapes):
    from oracle.pyoracle import tl_options_default
    g = load_golden("tl_reference.npz")
    n = len(shapes["names"])
    fr = shapes["frames"][[0, n + 30]]
    for t, sec in ((0, 30), (1, 30), (3, 30), (3, -1), (2, 50), (0, 70), (4, 0)):
        re, im = oracle.tl_spectrum(fr, F0_938, 938, type_=t, section=sec)
        assert np.array_equal(re, g[f"t{t}_s{sec}_re"]) and np.array_equal(im, g[f"t{t}_s{sec}_im"]), (t, sec)
    o = tl_options_default()
    o.piriform_fossa = 1; o.inner_length_corrections = 1; o.lung_pressure_dPa = 8000.0
    o.hagen_resistance = 1; o.heat_conduction = 1
    re, im = oracle.tl_spectrum(shapes["frames"][[0, 5, n + 30]], F0_938, 938, opts=o)
    assert np.array_equal(re, g["re_opts"]) and np.array_equal(im, g["im_opts"])


def test_formants_bit_identical_to_golden(oracle, shapes):
    g = load_golden("tl_reference.npz")
    f, bw, num, flags = oracle.tl_formants(shapes["frames"])
    assert np.array_equal(num, g["num_formants"]) and np.array_equal(flags, g["flags"])
    assert np.array_equal(f, g["formant_hz"]) and np.array_equal(bw, g["bandwidth_hz"])
    # SURVEY section 4 anchors: /a/ oral
    np.testing.assert_allclose(f[0, :3], [705.947857, 1246.117959, 2679.313681], atol=1e-6)


def test_impulse_response_bit_identical_to_golden(oracle, shapes):
    """TlModel::getImpulseResponse (spectrum product, mirror, complexIFFT with recursive twiddles, window)."""
    from oracle.pyoracle import tl_options_default
    g = load_golden("tl_impulse.npz")
    fr = shapes["frames"]
    assert np.array_equal(oracle.tl_impulse_response(fr[g["sel9"]], 22050.0, 9), g["h9"])
    assert np.array_equal(oracle.tl_impulse_response(fr[g["sel11"]], 22050.0, 11), g["h11"])
    o = tl_options_default(); o.piriform_fossa = 1; o.radiation = 1
    assert np.array_equal(oracle.tl_impulse_response(fr[g["sel9"][:4]], 22050.0, 9, opts=o), g["h9_opts"])


def test_tl_live_reference_random_tubes(oracle):
    if not have_ref():
        pytest.skip("compiled reference not available")
    from conftest import random_frames
    from oracle.pyoracle import Ref
    fr = random_frames(24, seed=11)
    re, im = Ref(22).tl_spectrum(fr, 8192)
    re2, im2 = oracle.tl_spectrum(fr, F0_2048, 2048)
    assert np.array_equal(re, re2) and np.array_equal(im, im2)


# ---------------------------------------------------------------------------------------------
# TDS restatement vs reference fixtures (bit-identical flow and audio, noise sources off)
# ---------------------------------------------------------------------------------------------
def test_tl_wave_propagation_mode_live_reference(oracle, shapes):
    """lumpedElements = false (TlModel.cpp:896-929): glibc csqrt/ccosh/csinh on both sides -> bit-identical."""
    if not have_ref():
        pytest.skip("compiled reference not available")
    from oracle.pyoracle import Ref, tl_options_default
    o = tl_options_default(); o.lumped_elements = 0; o.heat_conduction = 1
    fr = shapes["frames"][[0, 5, 20, 60, 90]]
    for t, sec in ((2, 25), (0, 30), (3, -1)):
        re, im = oracle.tl_spectrum(fr, 22050.0 / 2048, 938, type_=t, section=sec, opts=o)
        rr, ri = Ref(22).tl_spectrum(fr, 2048, type_=t, section=sec, opts=o)
        assert np.array_equal(re, rr[:, :938]) and np.array_equal(im, ri[:, :938])


def test_tds_static_vowel_bit_identical(oracle):
    g = load_golden("tds_reference.npz")
    a, u = oracle.tds_synthesize(g["static_frames"], g["static_glottis"], [2205], 44100.0, 2)
    assert np.array_equal(u, g["static_flow44"]) and np.array_equal(a, g["static_audio44"])
    a, u = oracle.tds_synthesize(g["static_frames"], g["static_glottis"], [2205], 22050.0, 1)
    assert np.array_equal(u, g["static_flow22"]) and np.array_equal(a, g["static_audio22"])


def test_tds_transition_velum_and_options_bit_identical(oracle):
    from oracle.pyoracle import tds_options_default
    g = load_golden("tds_reference.npz")
    a, u = oracle.tds_synthesize(g["trans_frames"], g["trans_glottis"], g["trans_len"], 44100.0, 2)
    assert np.array_equal(u, g["trans_flow44"]) and np.array_equal(a, g["trans_audio44"])
    o = tds_options_default()
    o.piriform_fossa = 1; o.glottis_loss_option = 1
    a, u = oracle.tds_synthesize(g["nasal_frames"], g["static_glottis"], [1102], 44100.0, 2, opts=o)
    assert np.array_equal(u, g["nasal_flow44"]) and np.array_equal(a, g["nasal_audio44"])


def test_tds_noise_sources_bit_identical(oracle):
    """calcNoiseSources/calcNoiseSample (TdsModel.cpp:1188-1708) with the glibc rand() stream restated."""
    from oracle.pyoracle import tds_options_default
    g = load_golden("tds_reference.npz")
    on = tds_options_default(noise=True)
    a, u = oracle.tds_synthesize(g["noise_frames"], g["noise_glottis"], [3308], 44100.0, 2, opts=on,
                                 seeds=g["noise_seeds"])
    assert np.array_equal(u, g["noise_flow44"]) and np.array_equal(a, g["noise_audio44"])
    # the fixture really exercises the sources: noise off gives a different waveform
    a0, _ = oracle.tds_synthesize(g["noise_frames"], g["noise_glottis"], [3308], 44100.0, 2)
    assert np.abs(a0 - a).max() > 1e-3 * np.abs(a).max()


def test_tds_transvelar_and_variable_loss_bit_identical(oracle):
    from oracle.pyoracle import tds_options_default
    g = load_golden("tds_reference.npz")
    o = tds_options_default()
    o.transvelar_coupling = 1; o.glottis_loss_option = 2
    a, u = oracle.tds_synthesize(g["opt_frames"], g["opt_glottis"], [2205], 44100.0, 2, opts=o)
    assert np.array_equal(u, g["opt_flow44"]) and np.array_equal(a, g["opt_audio44"])


def test_tds_state_dump_matches_reference(oracle):
    g = load_golden("tds_reference.npz")
    for step in (1, 2, 500):
        _, _, st = oracle.tds_synthesize(g["static_frames"], g["static_glottis"], [2205], 44100.0, 2, dump_step=step)
        assert np.array_equal(st, g[f"static_state_{step}"]), step


def test_tds_live_reference_random_transitions(oracle, shapes):
    if not have_ref():
        pytest.skip("compiled reference not available")
    from oracle.pyoracle import Ref
    rng = np.random.default_rng(3)
    n = len(shapes["names"])
    idx = rng.integers(0, n, size=(3, 2))
    fr = shapes["frames"][idx]  # [3 tracts, 2 frames], incl. consonant closures
    fr["velum_area_cm2"] = rng.random((3, 2)) * 0.6
    gl = np.tile(np.array([120.0, 8000.0, 0.01, 0.01, 0.0, -40.0]), (3, 2, 1))
    gl[:, :, 0] = rng.uniform(90, 220, size=(3, 2))
    a, u = oracle.tds_synthesize(fr, gl, [1500], 44100.0, 2)
    ar, ur = Ref(44).tds_synthesize(fr, gl, [1500])
    assert np.array_equal(u, ur) and np.array_equal(a, ar)


def test_chebyshev_restatement(oracle):
    g = load_golden("chebyshev.npz")
    a, b = oracle.chebyshev_lowpass(7000.0 / 44100.0, 8)
    assert np.array_equal(a, g["a44"]) and np.array_equal(b, g["b44"])


def test_flow_source_probes_and_mean_flow_match_live_reference(oracle, shapes):
    """Pins the oracle's restatement of TdsModel::setFlowSource (TdsModel.cpp:629-633, 996-1008), of the taps
    getSectionPressure / getSectionFlow (:1715-1734) and of TlModel::getMeanFlow / getMostConstrictedSection
    (TlModel.cpp:204-246) to the compiled reference, bit for bit."""
    if not have_ref():
        pytest.skip("compiled reference not available")
    from oracle.pyoracle import Ref, tds_options_default
    r = Ref(44)
    n = len(shapes["names"])
    fr = np.stack([shapes["frames"][0], shapes["frames"][n + 8]])  # /a:/ -> /i:/ with nasal coupling
    gl = np.array([[120.0, 8000.0, 0.01, 0.01, 0.0, -40.0], [150.0, 8000.0, 0.01, 0.01, 0.0, -40.0]])
    sections = [0, 22, 24, 28, 40, 64, 66, 83, 84, 90][:8]
    for src_sec, amp in ((-1, 0.0), (30, 50.0), (5, -20.0), (70, 10.0)):
        o = tds_options_default(noise=False); o.piriform_fossa = 1
        o.flow_source_section = src_sec; o.flow_source_cm3_s = amp
        ra, rp, ri, ro = r.tds_probed(fr, gl, [300], sections, opts=o)
        oa, op, oi, oo = oracle.tds_synthesize_probed(fr[None], gl[None], [300], sections, opts=o)
        assert np.array_equal(ra, oa[0]), src_sec
        assert np.array_equal(rp, op[0]) and np.array_equal(ri, oi[0]) and np.array_equal(ro, oo[0]), src_sec
    # the source really acts
    o0 = tds_options_default(noise=False); o1 = tds_options_default(noise=False)
    o1.flow_source_section = 30; o1.flow_source_cm3_s = 50.0
    a0, _ = oracle.tds_synthesize(fr[None], gl[None], [300], opts=o0)
    a1, _ = oracle.tds_synthesize(fr[None], gl[None], [300], opts=o1)
    assert np.max(np.abs(a0 - a1)) > 1e-3 * np.max(np.abs(a0))
    f = shapes["frames"].copy()
    f["glottis_area_cm2"][::3] = 0.0
    for P in (8000.0, 0.0, 1234.5):
        rf, rs = Ref(22).tl_mean_flow(f, P)
        of, os_ = oracle.tl_mean_flow(f, P)
        assert np.array_equal(rf, of) and np.array_equal(rs, os_)
