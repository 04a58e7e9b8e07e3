"""Scratch: cost of the chunked path (afs_tds_stream_*) next to the one-call path, and of the OPT instantiation of K1."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import areafunctionsynthesis_b200 as afs
from bench import make_tds_batch
B = int(os.environ.get("B", "4096")); T = int(os.environ.get("T", "44100")); NCH = int(os.environ.get("NCH", "5"))
ctx = afs.Context(0)
fr, gl = make_tds_batch(B, 20260003)
d_fr = torch.from_numpy(fr.view(np.uint8).reshape(B, -1)).cuda(); d_gl = torch.from_numpy(gl).cuda()
d_a = torch.empty(B, T // 2, dtype=torch.float64, device="cuda")
d_seeds = torch.arange(1, B + 1, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for noise in (False, True):
    for opt in (False, True):
        o = afs.tds_options_default(noise=noise)
        if opt: o.flow_source_section = 30; o.flow_source_cm3_s = 0.0
        ctx.tds_synthesize_device(d_fr, d_gl, B, 2, [T // 2], d_a, st, opts=o, d_seeds=d_seeds); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ctx.tds_synthesize_device(d_fr, d_gl, B, 2, [T // 2], d_a, st, opts=o, d_seeds=d_seeds); e1.record(); torch.cuda.synchronize()
        print(f"noise={noise} OPT kernel={opt}: {e0.elapsed_time(e1):.1f} ms")
# stream, chunk by chunk (pinned buffers)
o = afs.tds_options_default(noise=True)
n_out = T // 2 // NCH
h_f = [torch.from_numpy(np.ascontiguousarray(fr[:, k % 2]).view(np.uint8).reshape(B, -1).copy()).pin_memory() for k in range(NCH + 1)]
h_g = [torch.from_numpy(np.ascontiguousarray(gl[:, k % 2])).pin_memory() for k in range(NCH + 1)]
h_a = torch.empty(B, n_out, dtype=torch.float64).pin_memory()
h_s = torch.zeros(B, dtype=torch.int32).pin_memory()
C = afs.api.C
sh = C.c_void_p()
assert ctx.lib.afs_tds_stream_create(ctx.h, C.byref(o), B, 44100.0, 2, None, C.byref(sh)) == 0
ln = np.array([n_out], dtype=np.int32)
for rep in range(2):
    ctx.lib.afs_tds_stream_reset(sh)
    ts = []
    for k in range(NCH + 1):
        t0 = time.perf_counter()
        assert ctx.lib.afs_tds_stream_synthesize(sh, 1, ln.ctypes.data, h_f[k].data_ptr(), h_g[k].data_ptr(), h_a.data_ptr() if k else None, None, h_s.data_ptr()) == 0
        ts.append((time.perf_counter() - t0) * 1e3)
    print("stream chunks (ms):", " ".join(f"{t:.1f}" for t in ts), f"| one chunk of the kernel alone would be {T / NCH} steps")
