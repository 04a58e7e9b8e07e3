"""Generate tests/golden/target_sequence.npz: Synthesizer::playTargetSequence (Synthesizer.cpp:1299-1422) restated over
the UNMODIFIED reference classes (oracle/ref_harness.cpp: ref_play_target_sequence; TDS at 44.1 kHz = the sr44 build).
Needs /root/reference.

    python tools/make_golden_targets.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Ref, tds_options_default  # noqa: E402

G6 = [120.0, 8000.0, 0.01, 0.01, 0.0, -40.0]
STATIONARY = [0.06, 0.04, 0.04, 0.06]
TRANSITION = [0.08, 0.08, 0.08]


def main():
    g = np.load(os.path.join(ROOT, "tests", "golden", "default_shapes.npz"))
    idx = {str(n): i for i, n in enumerate(g["names"])}
    seqs = [["a:", "i:", "u:", "@"], ["e:", "o:", "a:", "y:"], ["@", "(a)d(a):", "a:", "E:"]]
    shapes = np.stack([g["params"][[idx[s] for s in q]] for q in seqs])
    r = Ref(44)
    out = dict(shapes=shapes, stationary=np.array(STATIONARY), transition=np.array(TRANSITION), glottis=np.array(G6))
    for noise in (0, 1):
        res = [r.play_target_sequence(shapes[k], STATIONARY, TRANSITION, G6, opts=tds_options_default(noise=bool(noise)), seed=11 + k)
               for k in range(len(seqs))]
        out[f"audio_noise{noise}"] = np.stack([x[0] for x in res])
    out["params"] = res[0][1][::7].copy()     # trajectories of the first sequence, every 7th sample, exact values
    out["glottis_traj"] = res[0][2][::7].copy()
    out["seeds"] = np.array([11, 12, 13], dtype=np.uint32)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "target_sequence.npz"), **out)
    print(out["audio_noise0"].shape, np.abs(out["audio_noise0"]).max(), np.abs(out["audio_noise1"]).max())


if __name__ == "__main__":
    main()
