"""Parametric-tract dataset sweep (BASELINE.json configs[4]) with every stage on the device:

    parameter sets [N,16] --K4 raster--> tube frames --K2 tl_sweep--> spectra [N,K] --K3--> formants
                                                   \\--K1 tds_batch--> audio [N, n_out]

The reference does this per parameter set on the CPU (OneDimAreaFunction::calculateOneDimTubeFunction ->
TlModel::getSpectrum / getFormants -> Synthesizer::synthesizeSignalTds).  Here the frames never leave the GPU
between the stages; torch only owns the device buffers.  Units are independent: shard the parameter sets over
ranks with sharding.shard_bounds, no collective on the data path.
"""
import numpy as np

from .api import FRAME_DTYPE, default_frames, tds_options_default, tl_options_default


def dataset_sweep(ctx, params, num_freq=2048, f0_hz=22050.0 / 8192, utterance_s=0.5, fs_tds_hz=44100.0, sr_factor=2,
                  glottis=(120.0, 8000.0, 0.01, 0.01, 0.0, -40.0), tl_opts=None, tds_opts=None, seeds=None,
                  want_formants=True):
    """params [N,16] (host) -> dict(frames, re, im, formant_hz, bandwidth_hz, num_formants, audio, status) as torch
    CUDA tensors (frames as bytes [N,1184]).  TL: glottis closed (Data.cpp:167-172); TDS: static tube, constant
    glottis control parameters."""
    import torch
    p = np.ascontiguousarray(np.atleast_2d(params), dtype=np.float64)
    n = len(p)
    dev = torch.device("cuda", ctx.device)
    stream = torch.cuda.current_stream(dev).cuda_stream
    d_p = torch.from_numpy(p).to(dev)
    base = default_frames(1)
    base["glottis_area_cm2"] = 0.0
    d_fr = torch.from_numpy(base.view(np.uint8).reshape(1, -1)).to(dev).repeat(n, 1).contiguous()
    ctx.area_functions_to_frames_device(d_p, d_fr, n, stream)
    out = {"frames": d_fr}
    d_re = torch.empty(n, num_freq, dtype=torch.float64, device=dev)
    d_im = torch.empty_like(d_re)
    ctx.tl_sweep_device(d_fr, n, f0_hz, num_freq, d_re, d_im, stream, opts=tl_opts or tl_options_default())
    out["re"], out["im"] = d_re, d_im
    if want_formants:
        torch.cuda.synchronize(dev)
        fr_host = np.frombuffer(d_fr.cpu().numpy().tobytes(), dtype=FRAME_DTYPE)
        f, bw, num, flags = ctx.tl_formants(fr_host, opts=tl_opts)
        out.update(formant_hz=f, bandwidth_hz=bw, num_formants=num, formant_flags=flags)
    n_out = int(round(utterance_s * fs_tds_hz / sr_factor))
    d_fr2 = d_fr.view(n, 1, -1).expand(n, 2, d_fr.shape[1]).contiguous()
    d_gl = torch.tensor(glottis, dtype=torch.float64, device=dev).view(1, 1, 6).expand(n, 2, 6).contiguous()
    d_audio = torch.empty(n, n_out, dtype=torch.float64, device=dev)
    d_status = torch.zeros(n, dtype=torch.int32, device=dev)
    d_seeds = None if seeds is None else torch.from_numpy(np.ascontiguousarray(seeds, dtype=np.uint32).view(np.int32)).to(dev)
    ctx.tds_synthesize_device(d_fr2, d_gl, n, 2, [n_out], d_audio, stream, fs_tds_hz, sr_factor,
                              opts=tds_opts or tds_options_default(), d_seeds=d_seeds, d_status=d_status)
    torch.cuda.synchronize(dev)
    out["audio"], out["status"] = d_audio, d_status
    return out
