"""Quick device-side timing of tl_sweep (scratch tool, not the bench contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import areafunctionsynthesis_b200 as afs
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import random_frames

B, K = 4096, 2048
ctx = afs.Context(0)
opts = afs.tl_options_default(); opts.arithmetic = int(os.environ.get("FAST", "0"))
fr = random_frames(B, seed=20260001)
d_fr = torch.from_numpy(fr.view(np.uint8).reshape(B, -1)).cuda()
d_re = torch.empty(B, K, dtype=torch.float64, device="cuda"); d_im = torch.empty_like(d_re)
s = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    ctx.tl_sweep_device(d_fr, B, 22050.0 / 8192, K, d_re, d_im, s, opts=opts)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ctx.tl_sweep_device(d_fr, B, 22050.0 / 8192, K, d_re, d_im, s, opts=opts)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"tl_sweep {B}x{K}: {ms:.3f} ms -> {B / ms * 1e3:.0f} spectra/s")
t = time.time(); re, im = ctx.tl_sweep(fr, 22050.0 / 8192, K, opts=opts); print("host e2e s:", time.time() - t)
print("checksum", float(d_re.sum()), float(re.sum()))
