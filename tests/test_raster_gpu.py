"""GPU parity tests of K4 area_function_raster_kernel (OneDimAreaFunction parameter sets -> tube frames).

Checker: the host rasteriser `afs_area_function_to_frame`, which is pinned bit-identical to the reference's
OneDimAreaFunction::calculateOneDimTubeFunction on all 51 Default.params shapes (tests/test_abi.py against
tests/golden/default_shapes.npz).  Bar: section lengths, articulator labels and incisor position identical (they come
from exact IEEE additions / comparisons that the kernel repeats); areas within 1e-12 relative (device pow/cos differ
from glibc by an ulp or two; measured ~1e-15).
"""
import numpy as np
import pytest

import areafunctionsynthesis_b200 as afs
from conftest import load_golden

pytestmark = pytest.mark.gpu

AREA_TOL = 1e-12


def _random_params(n, seed):
    rng = np.random.default_rng(seed)
    lo = np.array([-0.2, 0.67, 0.015, 1.92, 1.82, 0.37, 6.36, 0.087, 0.31, 10.97, 0.092, 0.0, 16.33, 1.05, 16.53, 0.107])
    hi = np.array([2.63, 2.05, 5.22, 6.43, 11.55, 6.96, 16.44, 2.50, 5.21, 18.82, 14.33, 9.33, 20.45, 6.12, 21.52, 4.40])
    p = lo + (hi - lo) * rng.random((n, 16))
    p[:, [0, 3, 6, 9, 12, 14]] = np.sort(p[:, [0, 3, 6, 9, 12, 14]], axis=1)
    p[:, 0] = np.maximum(p[:, 0], 0.0)
    return p


def _check(dev, host):
    assert np.array_equal(dev["length_cm"], host["length_cm"])
    assert np.array_equal(dev["articulator"], host["articulator"])
    assert np.array_equal(dev["incisor_pos_cm"], host["incisor_pos_cm"])
    assert np.array_equal(dev["laterality"], host["laterality"])
    rel = np.abs(dev["area_cm2"] - host["area_cm2"]) / host["area_cm2"]
    assert rel.max() <= AREA_TOL, rel.max()
    for name in ("velum_area_cm2", "aspiration_dB", "glottis_area_cm2", "glottis_length_cm"):  # untouched fields
        assert np.array_equal(dev[name], host[name]), name


def test_default_params_shapes(ctx):
    g = load_golden("default_shapes.npz")
    p = g["params"]  # the 51 shapes of Default.params
    dev = ctx.area_functions_to_frames(p)
    host = afs.area_function_to_frames(p)
    _check(dev, host)
    # and against the frames produced by the reference itself (fixture)
    ref = g["frames"][:len(p)]
    assert np.array_equal(dev["articulator"], ref["articulator"]) and np.array_equal(dev["length_cm"], ref["length_cm"])
    assert np.max(np.abs(dev["area_cm2"] - ref["area_cm2"]) / ref["area_cm2"]) <= AREA_TOL


def test_random_parameter_sets_and_ragged_batches(ctx):
    for n in (1, 3, 4, 5, 257):  # CTAs hold four sets: partial CTAs, single set
        p = _random_params(n, 20260001 + n)
        _check(ctx.area_functions_to_frames(p), afs.area_function_to_frames(p))
    assert len(ctx.area_functions_to_frames(np.zeros((0, 16)))) == 0  # empty batch
    # the other fields of caller-provided frames survive
    fr = afs.default_frames(7)
    fr["velum_area_cm2"] = np.arange(7) * 0.1
    fr["glottis_area_cm2"] = 0.25
    p = _random_params(7, 5)
    out = ctx.area_functions_to_frames(p, fr.copy())
    assert np.array_equal(out["velum_area_cm2"], fr["velum_area_cm2"]) and np.all(out["glottis_area_cm2"] == 0.25)


def test_device_frames_feed_the_tl_sweep(ctx, oracle):
    """End to end on the device side of the boundary: rasterise -> TL sweep; the spectra of device-made frames agree
    with those of host-made frames within the north-star TL tolerance (1e-9), although the frames differ by ulps."""
    p = _random_params(24, 77)
    dev = ctx.area_functions_to_frames(p); host = afs.area_function_to_frames(p)
    for f in (dev, host):
        f["glottis_area_cm2"] = 0.0
    f0 = 22050.0 / 8192
    re, im = ctx.tl_sweep(dev, f0, 2048)
    ro, io = oracle.tl_spectrum(host, f0, 2048)
    mag, mago = np.hypot(re, im), np.hypot(ro, io)
    assert np.max(np.abs(mag - mago) / mago) <= 1e-9
