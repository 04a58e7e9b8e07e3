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
if os.environ.get("STREAM") == "1":  # chunked synthesis: two chunks of 0.2 s, the second one resumes from the state of the first
    st = afs.TdsStream(ctx, B, 44100.0, 2, opts=afs.tds_options_default(noise=noise), seeds=np.arange(1, B + 1, dtype=np.uint32))
    st.synthesize(fr[:, :1], gl[:, :1], [0])
    for k in (1, 0):
        a, _, _ = st.synthesize(fr[:, k:k + 1], gl[:, k:k + 1], [4410])
    st.close()
    print("done", float(np.abs(a).mean()))
    sys.exit(0)
ctx.tds_synthesize_device(d_fr, d_gl, B, 2, [T // 2], d_a, torch.cuda.current_stream().cuda_stream,
                          opts=afs.tds_options_default(noise=noise), d_seeds=d_seeds)
torch.cuda.synchronize()
print("done", float(d_a.abs().mean()))
