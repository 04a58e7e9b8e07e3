"""Per-phase clock64 breakdown of tds_batch_kernel (needs a build with AFS_TDS_TIMING=1)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import areafunctionsynthesis_b200 as afs
from bench import make_tds_batch
B, T = int(os.environ.get("B", "4096")), int(os.environ.get("T", "2205"))
ctx = afs.Context(0)
if os.environ.get('SPREAD'): ctx.debug_set('tds.spread_sms', int(os.environ['SPREAD']))
if os.environ.get('SLICES'): ctx.debug_set('tds.d2h_slices', int(os.environ['SLICES']))
fr, gl = make_tds_batch(B, 1)
buf = np.zeros(473); ctx.set_state_dump(10**9, buf)
noise = os.environ.get('NOISE', '0') == '1'
a, _, _ = ctx.tds_synthesize(fr, gl, [T // 2 + 1], 44100.0, 2, opts=afs.tds_options_default(noise=noise))
steps = (T // 2 + 1) * 2
names = ["top/stage", "passA", "B1 wait", "elim/aux", "reduced", "B5 wait", "back", "B6 wait"]
roles = ["GP", "P1", "P2", "M0", "M1", "M2", "M3", "NV", "T0", "T1", "N2", "T2", "F", "N1", "X1", "X2"]
t = buf[:128].reshape(16, 8) / steps
print("cycles per step, per warp (lane 0 of CTA 0):")
print("warp " + " ".join(f"{n[:10]:>10s}" for n in names) + "      total")
for w in range(16):
    print(f"{roles[w]:>4s} " + " ".join(f"{v:10.0f}" for v in t[w]) + f" {t[w].sum():10.0f}")
sw = buf[136:200].reshape(16, 4) / steps
print("static warps, swarp_eliminate: first row | sinus leaves | chain | publish   (X2: noise_stage slots | decays | draws | refill)")
for w in range(16):
    if sw[w].sum() > 0:
        print(f"{roles[w]:>4s} " + " ".join(f"{v:8.0f}" for v in sw[w]))

x1 = buf[200:336] / steps; x2 = buf[336:472] / steps
nzx = x2[x2 > 0]
print("cycles per step over the first 136 CTAs: X1 pass A min %.0f median %.0f max %.0f | whole step (%d CTAs) min %.0f median %.0f p90 %.0f max %.0f" % (
    x1.min(), np.median(x1), x1.max(), len(nzx), nzx.min(), np.median(nzx), np.percentile(nzx, 90), nzx.max()))
