"""Small invocations of every kernel for smoke / memcheck purposes (scratch tool; compute-sanitizer is closed on the GPU pool):
   python tools/sanitize_small.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import areafunctionsynthesis_b200 as afs
from conftest import random_frames
ctx = afs.Context(0)
fr = random_frames(5, seed=3)
fast = afs.tl_options_default(); fast.arithmetic = 1
for o in (afs.tl_options_default(), fast):
    for nf in (2, 129, 700):
        for t, sec in ((2, 25), (2, 50), (0, 30), (3, -1)):
            ctx.tl_sweep(fr, 10.0, nf, spectrum_type=t, section=sec, opts=o)
    o.piriform_fossa = 1; o.inner_length_corrections = 1
    ctx.tl_sweep(fr, 10.0, 300, opts=o); ctx.tl_sweep(fr, 10.0, 300, section=28, opts=o); ctx.tl_sweep(fr, 10.0, 300, section=0, opts=o)
    ctx.tl_formants(fr, opts=o)
    ctx.tl_impulse_response(fr, 22050.0, 9, opts=o); ctx.tl_impulse_response(fr, 22050.0, 4, opts=o); ctx.tl_impulse_response(fr, 44100.0, 11, opts=o)
lf = np.tile([300.0, 120.0, 0.5, 3.0, 0.02], (5, 1))
ctx.synthesize_vowels_lf(fr, lf, True); ctx.synthesize_vowels_lf(fr, lf, False)
g = np.load(os.path.join(ROOT, "tests", "golden", "default_shapes.npz"))
ctx.optimize_formants_vowel(g["params"][[0, 8, 16]], [[650, 1300, 2600]] * 3, 1.0, 0.0, None, 2)
gl = np.array([[[120.0, 8000.0, 0.01, 0.01, 0.0, -40.0]] * 2] * 3)
for noise in (False, True):
    ctx.tds_synthesize(np.stack([g["frames"][[0, 4]]] * 3), gl, [64], 44100.0, 2, opts=afs.tds_options_default(noise=noise))
print("sanitize_small: done, launches", ctx.launch_count)
