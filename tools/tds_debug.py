import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import areafunctionsynthesis_b200 as afs
from oracle.pyoracle import Oracle
g = np.load("tests/golden/tds_reference.npz")
ctx = afs.Context(0); o = Oracle()
names = [("p", 0, 93), ("pr", 93, 186), ("wc", 186, 279), ("u", 279, 376), ("ur", 376, 473)]
for step in (1, 2, 3, 10, 500):
    st = np.zeros(473); ctx.set_state_dump(step, st)
    ctx.tds_synthesize(g["static_frames"], g["static_glottis"], [2205], 44100.0, 2, opts=afs.tds_options_default(noise=False))
    _, _, ref = o.tds_synthesize(g["static_frames"], g["static_glottis"], [2205], 44100.0, 2, dump_step=step)
    out = []
    for nm, a, b in names:
        sc = np.max(np.abs(ref[a:b])) + 1e-300
        e = np.abs(st[a:b] - ref[a:b]) / sc
        out.append(f"{nm}:{e.max():.2e}@{e.argmax()}")
    print(step, " ".join(out))
ctx.set_state_dump(0)
a, u, s = ctx.tds_synthesize(g["static_frames"], g["static_glottis"], [2205], 44100.0, 2, want_flow=True, opts=afs.tds_options_default(noise=False))
print("static flow rel err", np.max(np.abs(u - g["static_flow44"])) / np.max(np.abs(g["static_flow44"])), "audio", np.max(np.abs(a - g["static_audio44"])) / np.max(np.abs(g["static_audio44"])), s)
a, u, s = ctx.tds_synthesize(g["trans_frames"], g["trans_glottis"], g["trans_len"], 44100.0, 2, want_flow=True, opts=afs.tds_options_default(noise=False))
print("trans flow rel err", np.max(np.abs(u - g["trans_flow44"])) / np.max(np.abs(g["trans_flow44"])), "audio", np.max(np.abs(a - g["trans_audio44"])) / np.max(np.abs(g["trans_audio44"])))
# timing
import torch
B = int(os.environ.get("B", "4096")); T = int(os.environ.get("T", "4410"))
fr = np.repeat(g["trans_frames"][:, :2], B, axis=0); gl = np.repeat(g["trans_glottis"][:, :2], B, axis=0)
d_fr = torch.from_numpy(fr.view(np.uint8).reshape(B, -1)).cuda(); d_gl = torch.from_numpy(gl).cuda()
d_a = torch.empty(B, T // 2, dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for noise in (False, True):
    o = afs.tds_options_default(noise=noise)
    ctx.tds_synthesize_device(d_fr, d_gl, B, 2, [T // 2], d_a, st, opts=o); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.tds_synthesize_device(d_fr, d_gl, B, 2, [T // 2], d_a, st, opts=o); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"noise={noise} B={B} T={T}: {ms:.1f} ms -> {ms*1e3/T:.2f} us/step, {B*T/44100/(ms/1e3):.0f} audio-s/s")
