"""Generate tests/golden/lf_vowels.npz: Synthesizer::synthesizeVowelLf (Synthesizer.cpp:645-801) restated over the
UNMODIFIED reference classes (oracle/ref_harness.cpp: ref_synthesize_vowel_lf).  Needs /root/reference.

    python tools/make_golden_lf.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Ref  # noqa: E402


def main():
    g = np.load(os.path.join(ROOT, "tests", "golden", "default_shapes.npz"))
    n = len(g["names"])
    sel = np.array([0, 4, 8, 12, 16, n + 20])
    lf = np.array([[300.0, 120.0, 0.5, 3.0, 0.02],     # LfPulse defaults
                   [300.0, 95.0, 0.6, 2.0, 0.05],
                   [300.0, 210.0, 0.4, 2.5, 0.01],
                   [300.0, 150.0, 0.7, 1.5, 0.10],
                   [300.0, 180.0, 0.55, 3.5, 0.03],
                   [300.0, 110.0, 0.5, 3.0, 0.02]])
    r = Ref(22)
    out = {"sel": sel, "lf": lf}
    for is_long in (1, 0):
        a, fl = [], []
        for k, i in enumerate(sel):
            a.append(r.synthesize_vowel_lf(g["frames"][i:i + 1], lf[k], is_long))
            fl.append(r.last_flow.copy())
        out[f"flow_{is_long}"] = np.stack(fl)[:, ::4].copy()  # glottal flow (pulse train), every 4th sample, exact values
        out[f"audio_{is_long}"] = np.stack([x[0] for x in a])
        out[f"pressure_{is_long}"] = np.stack([x[1] for x in a])[:, ::16].copy()  # every 16th sample, exact values
        print(is_long, out[f"audio_{is_long}"].shape, out[f"audio_{is_long}"].min(), out[f"audio_{is_long}"].max())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "lf_vowels.npz"), **out)


if __name__ == "__main__":
    main()
