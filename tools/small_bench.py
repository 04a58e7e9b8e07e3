"""Scratch: tds_small_kernel vs tds_batch_kernel on small batches (noise sources off), device-resident."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import areafunctionsynthesis_b200 as afs
from bench import make_tds_batch
T = int(os.environ.get("T", "4410"))
ctx = afs.Context(0)
st = torch.cuda.current_stream().cuda_stream
o = afs.tds_options_default(noise=False)
for B in [int(x) for x in os.environ.get("BS", "1,64,148,296,512,1024,2048").split(",")]:
    fr, gl = make_tds_batch(B, 20260003)
    d_fr = torch.from_numpy(fr.view(np.uint8).reshape(B, -1)).cuda(); d_gl = torch.from_numpy(gl).cuda()
    res = {}
    for name, mx in (("small", 1 << 30), ("batch", 0)):
        ctx.debug_set("tds.small_batch_max", mx)
        d_a = torch.zeros(B, T // 2, dtype=torch.float64, device="cuda")
        l0 = ctx.launch_count
        ctx.tds_synthesize_device(d_fr, d_gl, B, 2, [T // 2], d_a, st, opts=o); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ctx.tds_synthesize_device(d_fr, d_gl, B, 2, [T // 2], d_a, st, opts=o); e1.record(); torch.cuda.synchronize()
        res[name] = (e0.elapsed_time(e1), d_a.cpu().numpy())
    a, b = res["small"][1], res["batch"][1]
    err = float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
    print(f"B={B:5d} T={T}: small {res['small'][0]:8.2f} ms ({res['small'][0]*1e3/T:6.2f} us/step)  batch {res['batch'][0]:8.2f} ms  "
          f"speed-up {res['batch'][0]/res['small'][0]:5.2f}  rel diff {err:.2e}")
ctx.debug_set("tds.small_batch_max", -1)
