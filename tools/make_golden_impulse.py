"""Generate tests/golden/tl_impulse.npz from the UNMODIFIED reference (oracle/_ref): TlModel::getImpulseResponse
(TlModel.cpp:90-111) for Default.params shapes.  Runs only in the build container (needs /root/reference).

    python tools/make_golden_impulse.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Ref, tl_options_default  # noqa: E402


def main():
    g = np.load(os.path.join(ROOT, "tests", "golden", "default_shapes.npz"))
    n = len(g["names"])
    sel9 = np.r_[0:8, n + 10:n + 14]   # oral vowels + nasalised shapes
    sel11 = np.r_[0:2, n + 30:n + 31]
    r22 = Ref(22)
    h9 = r22.tl_impulse_response(g["frames"][sel9], 9)     # the call of Synthesizer::synthesizeVowelLf (Synthesizer.cpp:719-724)
    h11 = r22.tl_impulse_response(g["frames"][sel11], 11)
    o = tl_options_default(); o.piriform_fossa = 1; o.radiation = 1
    h9o = r22.tl_impulse_response(g["frames"][sel9[:4]], 9, opts=o)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tl_impulse.npz"), sel9=sel9, h9=h9, sel11=sel11, h11=h11, h9_opts=h9o)
    print("wrote tl_impulse.npz", h9.shape, h11.shape, float(np.abs(h9).max()))


if __name__ == "__main__":
    main()
