"""Mean arrival rank of the 16 warps of a mid-grid CTA at three barrier sites (needs a build with -DAFS_TDS_ORDER)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import areafunctionsynthesis_b200 as afs
from bench import make_tds_batch
B, T = 4096, int(os.environ.get("T", "22050"))
ctx = afs.Context(0)
fr, gl = make_tds_batch(B, 1)
roles = ["GP", "P1", "P2", "M0", "M1", "M2", "M3", "NV", "T0", "T1", "N2", "T2", "F", "N1", "X1", "X2"]
for noise in (False, True):
    buf = np.zeros(473); ctx.set_state_dump(10**9, buf)
    ctx.tds_synthesize(fr, gl, [T // 2], 44100.0, 2, opts=afs.tds_options_default(noise=noise))
    r = buf[:48].reshape(16, 3)
    print(f"noise={noise}: mean arrival rank (0 = first, 15 = last) at: end of pass A | end of step | barrier 1 (before the reduced system)")
    for w in np.argsort(-r[:, 0]): print(f"  {roles[w]:>3s} {r[w,0]:5.1f} {r[w,1]:5.1f} {r[w,2]:5.1f}")
