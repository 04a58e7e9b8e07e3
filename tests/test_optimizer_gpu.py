"""GPU parity of the batched formant optimiser (afs_optimize_formants_vowel) against golden runs of
Data::optimizeFormantsVowel restated over the unmodified reference classes (tools/make_golden_optimizer.py).

The search is discrete (which neighbour wins a run), so the bar has two parts: the SEQUENCE of changed parameters is
identical, and errors / formants / final parameters agree to 1e-7 relative (the device rasteriser's pow/cos differ
from glibc by an ulp, areas to ~1e-15, formants to ~1e-10)."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def test_optimizer_matches_reference_runs(ctx):
    g = load_golden("formant_optimizer.npz")
    n = len(g["start"])
    max_runs = g["trace"].shape[1]
    # problems differ in max change / min area / fixed parameters: these are per call, so run them one by one ...
    res = []
    for i in range(n):
        r = ctx.optimize_formants_vowel(g["start"][i], g["targets"][i], float(g["max_change"][i]), float(g["min_area"][i]),
                                        g["fixed"][i], max_runs)
        res.append(r)
        runs = int(g["runs"][i])
        assert int(r["runs"][0]) == runs, i
        assert np.array_equal(r["trace"][0, :runs, 0], g["trace"][i, :runs, 0]), i      # same parameter changed in every run
        np.testing.assert_allclose(r["trace"][0, :runs, 1], g["trace"][i, :runs, 1], rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(r["initial_error"][0], g["e0"][i], rtol=1e-9)
        np.testing.assert_allclose(r["final_error"][0], g["e1"][i], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(r["params"][0], g["params"][i], rtol=0, atol=1e-12)
        np.testing.assert_allclose(r["formants_hz"][0], g["formants"][i], rtol=1e-9)
        assert r["final_error"][0] <= r["initial_error"][0]
    # ... and those that share the call arguments as ONE batch: identical to the single runs
    same = [i for i in range(n) if g["max_change"][i] == 1.0 and g["min_area"][i] == 0.0 and not g["fixed"][i].any()]
    assert len(same) >= 2
    r = ctx.optimize_formants_vowel(g["start"][same], g["targets"][same], 1.0, 0.0, None, max_runs)
    for k, i in enumerate(same):
        assert np.array_equal(r["params"][k], res[i]["params"][0])
        assert r["runs"][k] == res[i]["runs"][0] and r["final_error"][k] == res[i]["final_error"][0]


def test_optimizer_edge_cases(ctx, shapes):
    import areafunctionsynthesis_b200 as afs
    p = shapes["params"][[0, 8]]
    t = np.array([[650.0, 1300.0, 2600.0], [300.0, 2100.0, 2900.0]])
    # everything fixed: one run, nothing changes
    r = ctx.optimize_formants_vowel(p, t, 1.0, 0.0, np.ones(16, dtype=np.int32), 5)
    assert np.array_equal(r["params"], p) and np.all(r["runs"] == 1)
    # the initial error is evaluated in the state the TlModel is handed over in, everything later with the 8000 dPa
    # TlModel::getFormants leaves behind (TlModel.cpp:427): equal only when the model was "used" before
    assert np.all(r["initial_error"] != r["final_error"])
    o = afs.tl_options_default(); o.lung_pressure_dPa = 8000.0
    r = ctx.optimize_formants_vowel(p, t, 1.0, 0.0, np.ones(16, dtype=np.int32), 5, opts=o)
    assert np.array_equal(r["initial_error"], r["final_error"])
    # run limit
    r = ctx.optimize_formants_vowel(p, t, 1.0, 0.0, None, 3)
    assert np.all(r["runs"] <= 3) and np.all(r["final_error"] < r["initial_error"])
    # empty batch, bad arguments
    r = ctx.optimize_formants_vowel(np.zeros((0, 16)), np.zeros((0, 3)))
    assert r["params"].shape == (0, 16)
    with pytest.raises(afs.AfsError):
        ctx.optimize_formants_vowel(p, t, max_runs=0)
