"""One launch of the TL sweep on the bench batch (profiling target for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import areafunctionsynthesis_b200 as afs
from bench import make_tl_batch, B_TL, K_TL, F0_TL
ctx = afs.Context(0)
f = make_tl_batch(B_TL, 20260001)
d_f = torch.from_numpy(f.view(np.uint8).reshape(B_TL, -1)).cuda()
d_re = torch.empty(B_TL, K_TL, dtype=torch.float64, device="cuda"); d_im = torch.empty_like(d_re)
o = afs.tl_options_default()
if os.environ.get("FAST") == "1": o.arithmetic = afs.TL_ARITHMETIC_FAST
ctx.tl_sweep_device(d_f, B_TL, F0_TL, K_TL, d_re, d_im, torch.cuda.current_stream().cuda_stream, opts=o)
torch.cuda.synchronize()
print("done")
