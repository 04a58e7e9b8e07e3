"""BASELINE.json configs[4] on a subset: parameter sets -> frames -> TL spectra + formants -> TDS
audio, every stage on the device (areafunctionsynthesis_b200.sweep.dataset_sweep), checked stage by stage against the
host rasteriser + oracle."""
import numpy as np
import pytest

import areafunctionsynthesis_b200 as afs
from areafunctionsynthesis_b200.sweep import dataset_sweep
from test_raster_gpu import _random_params

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("noise", [False, True])
def test_dataset_sweep_subset_against_oracle(ctx, oracle, noise):
    from oracle.pyoracle import tds_options_default as oracle_opts
    n, K, T_S = 24, 2048, 0.05
    p = _random_params(n, 20260004)
    seeds = np.arange(100, 100 + n, dtype=np.uint32)
    out = dataset_sweep(ctx, p, num_freq=K, utterance_s=T_S, tds_opts=afs.tds_options_default(noise=noise), seeds=seeds)
    # stage 1: frames
    host = afs.area_function_to_frames(p)
    host["glottis_area_cm2"] = 0.0
    dev = np.frombuffer(out["frames"].cpu().numpy().tobytes(), dtype=afs.FRAME_DTYPE)
    assert np.array_equal(dev["articulator"], host["articulator"]) and np.array_equal(dev["length_cm"], host["length_cm"])
    assert np.max(np.abs(dev["area_cm2"] - host["area_cm2"]) / host["area_cm2"]) <= 1e-12
    # stage 2: spectra (1e-9 relative magnitude: the frames differ by ulps) and formants
    ro, io = oracle.tl_spectrum(host, 22050.0 / 8192, K)
    mag, mago = np.hypot(out["re"].cpu().numpy(), out["im"].cpu().numpy()), np.hypot(ro, io)
    assert np.max(np.abs(mag - mago) / mago) <= 1e-9
    fo, bwo, numo, _ = oracle.tl_formants(host)
    assert np.array_equal(out["num_formants"], numo)
    assert np.allclose(out["formant_hz"], fo, rtol=1e-9, atol=1e-9)
    # stage 3: audio of the static tubes
    n_out = int(round(T_S * 44100.0 / 2))
    fr2 = np.stack([host, host], axis=1)
    gl = np.tile(np.array([120.0, 8000.0, 0.01, 0.01, 0.0, -40.0]), (n, 2, 1))
    ao, _ = oracle.tds_synthesize(fr2, gl, [n_out], 44100.0, 2, opts=oracle_opts(noise=noise), seeds=seeds, want_flow=False)
    a = out["audio"].cpu().numpy()
    assert not out["status"].cpu().numpy().any()
    for b in range(n):
        assert np.max(np.abs(a[b] - ao[b])) / np.max(np.abs(ao[b])) <= 1e-6, b
