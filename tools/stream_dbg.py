import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import areafunctionsynthesis_b200 as afs
from bench import make_tds_batch
B, T = 4096, 4410
ctx = afs.Context(0)
fr, gl = make_tds_batch(B, 20260003)
d_fr = torch.from_numpy(fr.view(np.uint8).reshape(B, -1)).cuda(); d_gl = torch.from_numpy(gl).cuda()
d_a = torch.empty(B, T // 2, dtype=torch.float64, device="cuda")
d_seeds = torch.arange(1, B + 1, dtype=torch.int32, device="cuda")
side = torch.cuda.Stream()
for name, stream in (("default", torch.cuda.current_stream()), ("side", side)):
    with torch.cuda.stream(stream):
        st = stream.cuda_stream
        for noise in (False, True):
            o = afs.tds_options_default(noise=noise)
            for rep in range(3):
                ctx.tds_synthesize_device(d_fr, d_gl, B, 2, [T // 2], d_a, st, opts=o, d_seeds=d_seeds); torch.cuda.synchronize()
            ts = []
            for rep in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); ctx.tds_synthesize_device(d_fr, d_gl, B, 2, [T // 2], d_a, st, opts=o, d_seeds=d_seeds); e1.record(stream); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            print(name, "noise", noise, ["%.2f" % t for t in ts])
