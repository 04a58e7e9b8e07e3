import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import areafunctionsynthesis_b200 as afs
from bench import make_tds_batch
B, T = 4096, 2205
ctx = afs.Context(0)
fr, gl = make_tds_batch(B, 20260003)
d_fr = torch.from_numpy(fr.view(np.uint8).reshape(B, -1)).cuda(); d_gl = torch.from_numpy(gl).cuda()
d_a = torch.empty(B, T // 2, dtype=torch.float64, device="cuda")
d_seeds = torch.arange(1, B + 1, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
o = afs.tds_options_default(noise=True)
ctx.tds_synthesize_device(d_fr, d_gl, B, 2, [T // 2], d_a, st, opts=o, d_seeds=d_seeds); torch.cuda.synchronize()
