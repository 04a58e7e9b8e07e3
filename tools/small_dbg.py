import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import areafunctionsynthesis_b200 as afs
from bench import make_tds_batch
T = 4410; B = 1024
ctx = afs.Context(0)
fr, gl = make_tds_batch(B, 20260003)
o = afs.tds_options_default(noise=False)
res = {}
for name, mx in (("small", 1 << 30), ("batch", 0)):
    ctx.debug_set("tds.small_batch_max", mx)
    a, _, st = ctx.tds_synthesize(fr, gl, [T // 2], 44100.0, 2, opts=o)
    res[name] = (a, st)
for name in res:
    a, st = res[name]
    badrows = np.nonzero(~np.isfinite(a).all(axis=1))[0]
    print(name, "non-finite rows", badrows[:20], "status nonzero", np.nonzero(st)[0][:20])
a, b = res["small"][0], res["batch"][0]
ok = np.isfinite(a).all(axis=1) & np.isfinite(b).all(axis=1)
err = np.max(np.abs(a[ok] - b[ok]), axis=1) / np.max(np.abs(b[ok]), axis=1)
print("max err over finite rows", err.max(), "argmax", np.nonzero(ok)[0][err.argmax()])
r = np.nonzero(~np.isfinite(a).all(axis=1))[0]
if len(r):
    t = r[0]
    print("first bad tract", t, "first bad sample", np.nonzero(~np.isfinite(a[t]))[0][:3], "velum", fr[t]["velum_area_cm2"], "min areas", fr[t]["area_cm2"].min(axis=1), "F0", gl[t, :, 0])
