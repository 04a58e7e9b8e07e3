"""Generate tests/golden/formant_optimizer.npz: Data::optimizeFormantsVowel (Frontend/Data.cpp:203-487) restated over
the UNMODIFIED reference classes (oracle/ref_harness.cpp: ref_optimize_formants_vowel).  Needs /root/reference.

    python tools/make_golden_optimizer.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Ref  # noqa: E402

MAX_RUNS = 100


def main():
    g = np.load(os.path.join(ROOT, "tests", "golden", "default_shapes.npz"))
    names = [str(x) for x in g["names"]]
    P = g["params"]
    idx = {n: i for i, n in enumerate(names)}
    fixed = np.zeros(16, dtype=np.int32); fixed[[0, 1, 2, 14]] = 1   # larynx tube and tract length fixed
    problems = [  # (shape, targets, max change cm, min advised area, fixed)
        ("a:", [650.0, 1300.0, 2600.0], 1.0, 0.0, None),
        ("i:", [300.0, 2100.0, 2900.0], 0.5, 0.15, None),
        ("u:", [320.0, 750.0, 0.0], 2.0, 0.0, fixed),          # two-formant target
        ("@", [500.0, 1500.0, 2500.0], 0.3, 0.0, None),
        ("E:", [560.0, 1900.0, 2700.0], 1.0, 0.3, fixed),
        ("o:", [380.0, 700.0, 2500.0], 1.0, 0.0, None),
    ]
    r = Ref(22)
    out = dict(start=[], targets=[], max_change=[], min_area=[], fixed=[], params=[], e0=[], e1=[], runs=[], trace=[], formants=[])
    for nm, t, mc, ma, fx in problems:
        p, e0, e1, runs, trace, fo = r.optimize_formants_vowel(P[idx[nm]], t, mc, ma, fx, MAX_RUNS)
        tr = np.full((MAX_RUNS, 2), np.nan); tr[:runs] = trace
        for k, v in zip(out, (P[idx[nm]], t, mc, ma, np.zeros(16, dtype=np.int32) if fx is None else fx, p, e0, e1, runs, tr, fo)):
            out[k].append(v)
        print(nm, t, "error %.4f -> %.4f in %d runs, F = %s" % (e0, e1, runs, fo), flush=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "formant_optimizer.npz"), **{k: np.array(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
