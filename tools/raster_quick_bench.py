"""Quick device-side timing of area_function_raster_kernel (scratch tool, not the bench contract)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import areafunctionsynthesis_b200 as afs
B = 65536
rng = np.random.default_rng(20260004)
lo = np.array([-0.2, 0.67, 0.015, 1.92, 1.82, 0.37, 6.36, 0.087, 0.31, 10.97, 0.092, 0.0, 16.33, 1.05, 16.53, 0.107])
hi = np.array([2.63, 2.05, 5.22, 6.43, 11.55, 6.96, 16.44, 2.50, 5.21, 18.82, 14.33, 9.33, 20.45, 6.12, 21.52, 4.40])
pr = lo + (hi - lo) * rng.random((B, 16))
pr[:, [0, 3, 6, 9, 12, 14]] = np.sort(pr[:, [0, 3, 6, 9, 12, 14]], axis=1)
pr[:, 0] = np.maximum(pr[:, 0], 0.0)
ctx = afs.Context(0)
d_pr = torch.from_numpy(pr).cuda()
d_fr = torch.from_numpy(afs.default_frames(B).view(np.uint8).reshape(B, -1)).cuda()
st = torch.cuda.current_stream().cuda_stream
for _ in range(2): ctx.area_functions_to_frames_device(d_pr, d_fr, B, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): ctx.area_functions_to_frames_device(d_pr, d_fr, B, st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"raster {B} sets: {ms:.3f} ms -> {B / ms * 1e3:.0f} sets/s")
host = afs.area_function_to_frames(pr[:512])
dev = np.frombuffer(d_fr[:512].cpu().numpy().tobytes(), dtype=afs.FRAME_DTYPE)
print("articulators equal:", np.array_equal(dev["articulator"], host["articulator"]), " lengths equal:", np.array_equal(dev["length_cm"], host["length_cm"]),
      " max rel area err:", float(np.max(np.abs(dev["area_cm2"] - host["area_cm2"]) / host["area_cm2"])))
