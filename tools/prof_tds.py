"""One launch of the TDS path on the bench batch (profiling target for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import areafunctionsynthesis_b200 as afs
from bench import make_tds_batch
B = 4096; T = int(os.environ.get("T", "44100")); noise = os.environ.get("NOISE", "1") == "1"
ctx = afs.Context(0)
fr, gl = make_tds_batch(B, 20260003)
d_fr = torch.from_numpy(fr.view(np.uint8).reshape(B, -1)).cuda(); d_gl = torch.from_numpy(gl).cuda()
d_a = torch.empty(B, T // 2, dtype=torch.float64, device="cuda")
d_seeds = torch.arange(1, B + 1, dtype=torch.int32, device="cuda")
ctx.tds_synthesize_device(d_fr, d_gl, B, 2, [T // 2], d_a, torch.cuda.current_stream().cuda_stream,
                          opts=afs.tds_options_default(noise=noise), d_seeds=d_seeds)
torch.cuda.synchronize()
print("done", float(d_a.abs().mean()))
