import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import areafunctionsynthesis_b200 as afs
from bench import make_tds_batch, bench_seeds
B, n_out = 4096, 22050
ctx = afs.Context(0)
fr, gl = make_tds_batch(B, 20260003); seeds = bench_seeds(B)
h_fr = torch.from_numpy(fr.view(np.uint8).reshape(B, -1).copy()).pin_memory()
h_gl = torch.from_numpy(gl.copy()).pin_memory()
h_seeds = torch.from_numpy(seeds.view(np.int32).copy()).pin_memory()
h_audio = torch.empty(B, n_out, dtype=torch.float64).pin_memory()
h_status = torch.zeros(B, dtype=torch.int32).pin_memory()
fl = np.array([n_out], dtype=np.int32)
def host(opts):
    rc = ctx.lib.afs_tds_synthesize(ctx.h, afs.api.C.byref(opts), B, 2, fl.ctypes.data, h_fr.data_ptr(), h_gl.data_ptr(), 44100.0, 2,
                                    h_seeds.data_ptr(), h_audio.data_ptr(), None, h_status.data_ptr())
    assert rc == 0
for factor in (1, 16):
    ctx.debug_set("tds.d2h_slices", factor)
    for noise, optk in ((False, False), (False, True), (True, False)):
        o = afs.tds_options_default(noise=noise)
        if optk:
            o.flow_source_section = 30; o.flow_source_cm3_s = 0.0
        ts = []
        for rep in range(4):
            t0 = time.perf_counter(); host(o); ts.append((time.perf_counter() - t0) * 1e3)
        print("slices", factor, "noise", noise, "opt kernel", optk, ["%.1f" % t for t in ts], "status", int(h_status.max()))
