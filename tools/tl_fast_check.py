"""Fast-arithmetic TL sweep vs the bit-identical kernel on the bench workload (scratch tool)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import areafunctionsynthesis_b200 as afs
from conftest import random_frames

B, K = 4096, 2048
ctx = afs.Context(0)
fr = random_frames(B, seed=20260001)
d_fr = torch.from_numpy(fr.view(np.uint8).reshape(B, -1)).cuda()
mk = lambda: torch.empty(B, K, dtype=torch.float64, device="cuda")
re0, im0, re1, im1 = mk(), mk(), mk(), mk()
s = torch.cuda.current_stream().cuda_stream
fast = afs.tl_options_default(); fast.arithmetic = afs.TL_ARITHMETIC_FAST
f0 = 22050.0 / 8192
ctx.tl_sweep_device(d_fr, B, f0, K, re0, im0, s)
for _ in range(3):
    ctx.tl_sweep_device(d_fr, B, f0, K, re1, im1, s, opts=fast)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ctx.tl_sweep_device(d_fr, B, f0, K, re1, im1, s, opts=fast)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"tl_sweep fast {B}x{K}: {ms:.3f} ms -> {B / ms * 1e3:.0f} spectra/s")
h0 = torch.complex(re0, im0); h1 = torch.complex(re1, im1)
rel = (h1 - h0).abs() / h0.abs()
print("max |dH|/|H|:", float(rel.max()), " max rel magnitude err:", float(((h1.abs() - h0.abs()).abs() / h0.abs()).max()),
      " max phase err rad:", float((h1 / h0).angle().abs().max()), " min |H|:", float(h0.abs().min()))
flat = rel.flatten()
top = torch.topk(flat, 24).indices.cpu().numpy()
tb, bb = top // K, top % K
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez(os.path.join(ROOT, "gpurun_out", "tl_fast_worst.npz"), frames=fr[tb].view(np.uint8), tubes=tb, bins=bb, f0=f0,
         ref_re=re0.cpu().numpy()[tb, bb], ref_im=im0.cpu().numpy()[tb, bb],
         fast_re=re1.cpu().numpy()[tb, bb], fast_im=im1.cpu().numpy()[tb, bb])
print("99.99 pct |dH|/|H|:", float(torch.quantile(rel.flatten()[:10_000_000], 0.9999)))
for name, o in (("fossa+ilc+heat+hagen", dict(piriform_fossa=1, inner_length_corrections=1, heat_conduction=1, hagen_resistance=1)),
                ("no walls no sinuses sphere", dict(soft_walls=0, paranasal_sinuses=0, radiation=1)),
                ("wall radiation, lung pressure", dict(radiation=2, lung_pressure_dPa=8000.0))):
    a = afs.tl_options_default(); b = afs.tl_options_default(); b.arithmetic = 1
    for k, v in o.items(): setattr(a, k, v); setattr(b, k, v)
    for t, sec in ((2, 25), (2, 50), (2, 65), (0, 30), (1, 30), (3, 30), (3, -1), (2, 0), (0, 70)):
        n = 256
        ctx.tl_sweep_device(d_fr, n, f0, K, re0, im0, s, spectrum_type=t, section=sec, opts=a)
        ctx.tl_sweep_device(d_fr, n, f0, K, re1, im1, s, spectrum_type=t, section=sec, opts=b)
        torch.cuda.synchronize()
        h0 = torch.complex(re0[:n], im0[:n]); h1 = torch.complex(re1[:n], im1[:n])
        print(f"{name:32s} type {t} section {sec:3d}: max |dH|/|H| {float(((h1 - h0).abs() / h0.abs()).max()):.3e}"
              f"  max |dH|/max|H| {float((h1 - h0).abs().max() / h0.abs().max()):.3e}")
