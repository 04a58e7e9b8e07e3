import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import areafunctionsynthesis_b200 as afs
from conftest import random_frames
fr = random_frames(64, seed=7)
ctx = afs.Context(0)
a, b = afs.tl_options_default(), afs.tl_options_default(); b.arithmetic = 1
a.piriform_fossa = 1; b.piriform_fossa = 1
f0 = 22050.0 / 2048
sec = 33
r1, i1 = ctx.tl_sweep(fr, f0, 938, section=sec, opts=a)
r2, i2 = ctx.tl_sweep(fr, f0, 938, section=sec, opts=b)
h1, h2 = r1 + 1j * i1, r2 + 1j * i2
rel = (np.abs(h2 - h1) / np.abs(h1)).flatten()
top = np.argsort(rel)[-12:][::-1]
tb, bb = top // 938, top % 938
print("max rel", rel.max())
np.savez("/root/repo/gpurun_out/tl_fast_worst_fossa.npz", frames=fr[tb].view(np.uint8), tubes=tb, bins=bb, f0=f0, section=sec, fossa=1,
         ref_re=r1[tb, bb], ref_im=i1[tb, bb], fast_re=r2[tb, bb], fast_im=i2[tb, bb])
