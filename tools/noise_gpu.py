import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np
import areafunctionsynthesis_b200 as afs
g = np.load('/root/repo/tests/golden/tds_reference.npz')
ctx = afs.Context(0)
o = afs.tds_options_default(); o.generate_noise_sources = 1
a, u, st = ctx.tds_synthesize(g["noise_frames"], g["noise_glottis"], [3308], 44100.0, 2, opts=o, seeds=g["noise_seeds"], want_flow=True)
ru, ra = g["noise_flow44"], g["noise_audio44"]
for b in range(3):
    e = np.abs(u[b]-ru[b]); 
    print(b, "flow rel", e.max()/np.abs(ru[b]).max(), "first bad step", int(np.argmax(e > 1e-6*np.abs(ru[b]).max())), "audio rel", np.abs(a[b]-ra[b]).max()/np.abs(ra[b]).max(), st[b])
o2 = afs.tds_options_default(); o2.generate_noise_sources = 0
a0, u0, _ = ctx.tds_synthesize(g["noise_frames"], g["noise_glottis"], [3308], 44100.0, 2, opts=o2, want_flow=True)
print("noise on vs off (ref):", np.abs(u0-ru).max()/np.abs(ru).max())
