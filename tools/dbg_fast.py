import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import areafunctionsynthesis_b200 as afs
from conftest import random_frames
g = np.load('/root/repo/tests/golden/default_shapes.npz'); n = len(g['names'])
fr = np.concatenate([g['frames'][[0, 9, n + 30]], random_frames(13, seed=7)])
ctx = afs.Context(0)
variants = (dict(), dict(piriform_fossa=1, inner_length_corrections=1, heat_conduction=1, hagen_resistance=1),
            dict(soft_walls=0, paranasal_sinuses=0, radiation=1), dict(radiation=2, lung_pressure_dPa=8000.0),
            dict(radiation=0, boundary_layer=0, static_pressure_drops=0))
for kw in variants:
    a, b = afs.tl_options_default(), afs.tl_options_default(); b.arithmetic = 1
    for k, v in kw.items(): setattr(a, k, v); setattr(b, k, v)
    for t, sec in ((2, 25), (2, 0), (2, 24), (2, 33), (2, 40), (2, 41), (2, 50), (2, 64), (2, 65)):
        r1, i1 = ctx.tl_sweep(fr, 22050.0/2048, 938, spectrum_type=t, section=sec, opts=a)
        r2, i2 = ctx.tl_sweep(fr, 22050.0/2048, 938, spectrum_type=t, section=sec, opts=b)
        h1, h2 = r1 + 1j * i1, r2 + 1j * i2
        rel = np.abs(h2 - h1) / np.abs(h1)
        w = np.unravel_index(np.argmax(rel), rel.shape)
        print(kw, t, sec, f"{rel.max():.3e} at tube {w[0]} bin {w[1]} |H|={abs(h1[w]):.3e} max|H|={np.abs(h1[w[0]]).max():.3e}")
