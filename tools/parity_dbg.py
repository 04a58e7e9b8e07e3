"""Scratch: where does the noise-on audio of the bench batch leave the reference?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import areafunctionsynthesis_b200 as afs
from bench import make_tds_batch, bench_seeds
from oracle.pyoracle import Ref, Oracle, tds_options_default as ropts
B = int(os.environ.get("B", "4096")); NP = int(os.environ.get("NP", "16")); T = float(os.environ.get("T", "1.0"))
n_out = int(round(T * 22050))
fr, gl = make_tds_batch(B, 20260003)
seeds = bench_seeds(B)
ctx = afs.Context(0)
on = afs.tds_options_default(noise=True)
a_all, _, _ = ctx.tds_synthesize(fr, gl, [n_out], 44100.0, 2, opts=on, seeds=seeds)
a_sub, _, _ = ctx.tds_synthesize(fr[:NP], gl[:NP], [n_out], 44100.0, 2, opts=on, seeds=seeds[:NP])
ref, _ = Ref(44).tds_synthesize(fr[:NP], gl[:NP], [n_out], opts=ropts(noise=True), seeds=seeds[:NP], want_flow=False)
orc, _ = Oracle().tds_synthesize(fr[:NP], gl[:NP], [n_out], 44100.0, 2, opts=ropts(noise=True), seeds=seeds[:NP], want_flow=False)
def first_bad(a, b):
    e = np.abs(a - b) / np.max(np.abs(b))
    bad = np.nonzero(e > 1e-6)[0]
    return (int(bad[0]) if len(bad) else -1), float(e.max())
print("oracle vs reference:", [first_bad(orc[b], ref[b]) for b in range(NP)])
for b in range(NP):
    print(b, "batch-of-%d vs ref" % B, first_bad(a_all[b], ref[b]), "| batch-of-%d vs ref" % NP, first_bad(a_sub[b], ref[b]),
          "| all vs sub equal:", bool(np.array_equal(a_all[b], a_sub[b])))
