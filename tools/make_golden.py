"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref).

Runs only in the build container (needs /root/reference); the fixtures it writes
are committed so that the GPU box -- which has no reference sources -- can check
both the oracle restatement and the CUDA path against reference outputs.

    python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Ref, default_frames, load_default_params, tds_options_default, tl_options_default  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
VOWELS = ["a:", "e:", "i:", "o:", "u:", "E:", "2:", "y:", "I", "E", "a", "O", "9", "Y", "U", "@"]


def main():
    os.makedirs(OUT, exist_ok=True)
    r22, r44 = Ref(22), Ref(44)
    names, P = load_default_params()
    n = len(P)

    # --- the 51 Default.params shapes as tube frames (oral, glottis closed) + nasal copies
    frames = default_frames(2 * n)
    for i, p in enumerate(P):
        r22.area_function_to_frame(p, frames[i:i + 1])
        r22.area_function_to_frame(p, frames[n + i:n + i + 1])
    frames["glottis_area_cm2"][:n] = 0.0
    frames["velum_area_cm2"][n:] = np.linspace(0.0, 1.0, n)
    np.savez_compressed(os.path.join(OUT, "default_shapes.npz"), names=np.array(names), params=P, frames=frames)

    # --- TL: formants of all 102 frames (TlModel::getFormants: N = 2048 -> 938 bins)
    f, bw, num, flags = r22.tl_formants(frames)
    # --- TL: spectra. 938-bin grid for 16 frames, 2048-bin grid (patched cap) for 8 frames
    sel938 = np.r_[0:8, n:n + 8]
    sel2048 = np.r_[0:4, n + 20:n + 24]
    re938, im938 = r22.tl_spectrum(frames[sel938], 2048)
    re2048, im2048 = r22.tl_spectrum(frames[sel2048], 8192)
    # other spectrum types on /a/ and a nasalised shape
    other = {}
    for t, sec in ((0, 30), (1, 30), (3, 30), (3, -1), (2, 50), (0, 70), (4, 0)):
        re, im = r22.tl_spectrum(frames[[0, n + 30]], 2048, type_=t, section=sec)
        other[f"t{t}_s{sec}_re"] = re
        other[f"t{t}_s{sec}_im"] = im
    o = tl_options_default()
    o.piriform_fossa = 1; o.inner_length_corrections = 1; o.lung_pressure_dPa = 8000.0
    o.hagen_resistance = 1; o.heat_conduction = 1
    re_o, im_o = r22.tl_spectrum(frames[[0, 5, n + 30]], 2048, opts=o)
    np.savez_compressed(os.path.join(OUT, "tl_reference.npz"), formant_hz=f, bandwidth_hz=bw, num_formants=num,
                        flags=flags, sel938=sel938, re938=re938, im938=im938, sel2048=sel2048, re2048=re2048,
                        im2048=im2048, re_opts=re_o, im_opts=im_o, **other)

    # --- TDS: reference audio/flow, noise off (gate config)
    idx = {nm: i for i, nm in enumerate(names)}
    g0 = np.array([120.0, 8000.0, 0.01, 0.01, 0.0, -40.0])
    tds = {}
    # (1) static /a/, 0.1 s at TDS 44.1 kHz and at 22.05 kHz (as shipped)
    fa = frames[idx["a:"]:idx["a:"] + 1].copy()
    fr = np.stack([fa, fa], axis=1)
    gl = np.stack([g0, g0])[None]
    a44, u44 = r44.tds_synthesize(fr, gl, [2205])
    a22, u22 = r22.tds_synthesize(fr, gl, [2205])
    tds.update(static_frames=fr, static_glottis=gl, static_audio44=a44, static_flow44=u44, static_audio22=a22,
               static_flow22=u22)
    # (2) /a/ -> /i/ with F0 glide, 3 control frames, 0.1 s + 0.05 s, velum opening on the way
    fi = frames[idx["i:"]:idx["i:"] + 1].copy()
    fu = frames[idx["u:"]:idx["u:"] + 1].copy()
    fu["velum_area_cm2"] = 0.4
    fr3 = np.stack([fa, fi, fu], axis=1)
    gl3 = np.array([[[110.0, 8000.0, 0.01, 0.01, 0.0, -40.0], [180.0, 9000.0, 0.012, 0.011, 0.02, -40.0],
                     [140.0, 7000.0, 0.02, 0.02, 0.0, -40.0]]])
    a3, u3 = r44.tds_synthesize(fr3, gl3, [2205, 1102])
    tds.update(trans_frames=fr3, trans_glottis=gl3, trans_len=np.array([2205, 1102], dtype=np.int32), trans_audio44=a3,
               trans_flow44=u3)
    # (3) piriform fossa on, van-den-Berg entrance loss, nasal vowel
    o = tds_options_default()
    o.piriform_fossa = 1; o.glottis_loss_option = 1
    fn = frames[n + 40:n + 41].copy()
    frn = np.stack([fn, fn], axis=1)
    an, un = r44.tds_synthesize(frn, gl, [1102], opts=o)
    tds.update(nasal_frames=frn, nasal_audio44=an, nasal_flow44=un)
    # (4) state right after steps 1, 2 and 500 of the static run (single-step debugging anchors)
    for step in (1, 2, 500):
        _, _, st = r44.tds_state_after(fr[0], gl[0], [2205], step)
        tds[f"static_state_{step}"] = st
    # (5) noise sources on (TdsModel.cpp:1188-1708; libc rand() seeded per utterance): /a/ -> /s/ (tongue
    #     constriction + teeth source), /a/ -> /f/ (lip source), static /S/ with strong aspiration
    on = tds_options_default(noise=True)
    gn = np.array([[[120.0, 8000.0, 0.01, 0.01, 0.0, -40.0], [120.0, 8000.0, 0.03, 0.03, 0.05, -10.0]]] * 3)
    frn3 = np.stack([np.stack([fa[0], frames[idx["s"]]]), np.stack([fa[0], frames[idx["f"]]]),
                     np.stack([frames[idx["S"]], frames[idx["S"]]])])
    seeds = np.array([1234, 99, 20260002], dtype=np.uint32)
    ans, uns = r44.tds_synthesize(frn3, gn, [3308], opts=on, seeds=seeds)
    tds.update(noise_frames=frn3, noise_glottis=gn, noise_seeds=seeds, noise_audio44=ans, noise_flow44=uns)
    # (6) non-default options: transvelar coupling + Fulcher variable entrance loss (TdsModel.cpp:476-533, 966-999,
    #     1019-1092), nasalised /a/ -> /i/
    oo = tds_options_default()
    oo.transvelar_coupling = 1; oo.glottis_loss_option = 2
    fo = np.stack([fa[0], fi[0]])[None].copy()
    fo["velum_area_cm2"] = 0.3
    ao_, uo_ = r44.tds_synthesize(fo, gl3[:, :2], [2205], opts=oo)
    tds.update(opt_frames=fo, opt_glottis=gl3[:, :2], opt_audio44=ao_, opt_flow44=uo_)
    np.savez_compressed(os.path.join(OUT, "tds_reference.npz"), **tds)

    # --- filter design
    a, b = r44.chebyshev_lowpass(7000.0 / 44100.0, 8)
    a2, b2 = r22.chebyshev_lowpass(7000.0 / 22050.0, 8)
    np.savez_compressed(os.path.join(OUT, "chebyshev.npz"), a44=a, b44=b, a22=a2, b22=b2)
    for fn_ in sorted(os.listdir(OUT)):
        print(fn_, os.path.getsize(os.path.join(OUT, fn_)))


if __name__ == "__main__":
    main()
