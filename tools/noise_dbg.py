import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import areafunctionsynthesis_b200 as afs
from oracle.pyoracle import Oracle, tds_options_default as oracle_opts
g = np.load("tests/golden/default_shapes.npz")
names = [str(x) for x in g["names"]]; frames = g["frames"]
G0 = np.array([120.0, 8000.0, 0.01, 0.01, 0.0, -40.0])
n = len(names)
rng = np.random.default_rng(20260004)
B = 36
idx = rng.integers(0, n, size=(B, 3))
fr = frames[idx].copy()
fr["velum_area_cm2"] = np.where(rng.random((B, 3)) < 0.3, rng.random((B, 3)) * 0.5, 0.0)
gl = np.tile(G0, (B, 3, 1))
gl[:, :, 0] = rng.uniform(90, 220, size=(B, 3))
gl[:, :, 2:4] = rng.uniform(0.0, 0.04, size=(B, 3, 1))
gl[:, :, 5] = rng.uniform(-40.0, -10.0, size=(B, 3))
seeds = rng.integers(0, 2**32, size=B, dtype=np.uint64).astype(np.uint32)
fl = [900, 601]
ctx = afs.Context(0); orc = Oracle()
a, u, st = ctx.tds_synthesize(fr, gl, fl, 44100.0, 2, opts=afs.tds_options_default(noise=True), seeds=seeds, want_flow=True)
ao, uo = orc.tds_synthesize(fr, gl, fl, 44100.0, 2, opts=oracle_opts(noise=True), seeds=seeds)
a0, u0 = orc.tds_synthesize(fr, gl, fl, 44100.0, 2, opts=oracle_opts(noise=False))
for b in range(B):
    e = np.abs(u[b] - uo[b]); sc = np.abs(uo[b]).max()
    bad = np.nonzero(e > 1e-8 * sc)[0]
    print(b, [names[i] for i in idx[b]], f"rel {e.max()/sc:.2e}", "first>1e-8 at", (int(bad[0]) if len(bad) else -1),
          f"noise effect {np.abs(uo[b]-u0[b]).max()/sc:.1e}")
for b in (3, 8):
    print("tract", b)
    for t in range(46, 70):
        print(t, f"{u[b][t]:+.10e} {uo[b][t]:+.10e} {u0[b][t]:+.10e}  d_gpu_orc {u[b][t]-uo[b][t]:+.2e}  d_noise {uo[b][t]-u0[b][t]:+.2e}")
print("noise OFF on both sides:")
aq, uq, _ = ctx.tds_synthesize(fr, gl, fl, 44100.0, 2, opts=afs.tds_options_default(noise=False), want_flow=True)
for b in range(B):
    e = np.abs(uq[b] - u0[b]); sc = np.abs(u0[b]).max()
    if e.max()/sc > 2e-9: print(b, f"rel {e.max()/sc:.2e}", "rest", gl[b,:,2], "f0", gl[b,:,0])
