import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def shapes():
    g = load_golden("default_shapes.npz")
    return {"names": [str(x) for x in g["names"]], "params": g["params"], "frames": g["frames"]}


@pytest.fixture(scope="session")
def ctx():
    """CUDA context through the C ABI; the GPU tests fail loudly if the extension is missing."""
    import areafunctionsynthesis_b200 as afs
    c = afs.Context(0)
    yield c
    c.close()


def have_ref():
    return os.path.isdir("/root/reference") or os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libafs_ref_sr44.so"))


def random_frames(n, seed, nasal_fraction=0.5, glottis_area=0.0):
    """Random parametric area functions in the SURVEY 8d C2 ranges -> frames via the PRODUCT's host
    rasteriser (inputs only; both sides of every comparison consume the same frames)."""
    import areafunctionsynthesis_b200 as afs
    rng = np.random.default_rng(seed)
    lo = np.array([-0.2, 0.67, 0.015, 1.92, 1.82, 0.37, 6.36, 0.087, 0.31, 10.97, 0.092, 0.0, 16.33, 1.05, 16.53, 0.107])
    hi = np.array([2.63, 2.05, 5.22, 6.43, 11.55, 6.96, 16.44, 2.50, 5.21, 18.82, 14.33, 9.33, 20.45, 6.12, 21.52, 4.40])
    p = lo + (hi - lo) * rng.random((n, 16))
    xs = np.sort(p[:, [0, 3, 6, 9, 12, 14]], axis=1)  # monotone x coordinates (OneDimAreaFunction.cpp:193-212)
    p[:, [0, 3, 6, 9, 12, 14]] = xs
    p[:, 0] = np.maximum(p[:, 0], 0.0)
    f = afs.area_function_to_frames(p)
    f["glottis_area_cm2"] = glottis_area
    nasal = rng.random(n) < nasal_fraction
    f["velum_area_cm2"] = np.where(nasal, rng.random(n), 0.0)
    return f
