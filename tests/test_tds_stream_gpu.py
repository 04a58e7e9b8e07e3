"""GPU tests of the chunked synthesis (afs_tds_stream_*): models that live across calls, like the reference's
Synthesizer (one synthesizeSignalTds call per control frame, Synthesizer.cpp:515-639; SURVEY.md section 5: state in /
state out per tract).

Properties checked: (i) the audio of a chunked utterance equals the audio of the same control frames given to
afs_tds_synthesize in one call (bar 1e-9 relative; measured: bit-identical with the noise sources off, 7e-10 with
them on; the measured difference is printed), whatever the chunking, for every option set and with
the noise sources on (the tract's random() stream and its noise-source filters are part of the state); (ii) the
chunked audio meets the 1e-6 bar against the oracle, which synthesises the utterance in one piece; (iii) a saved
stream continues bit for bit like the stream it was saved from.
"""
import numpy as np
import pytest

from test_tds_gpu import _vowel_transitions, rel_err

pytestmark = pytest.mark.gpu

CHUNK_TOL = 1e-9


def _utterances(shapes, n, seed, n_frames, nasal=True):
    """n utterances of n_frames control frames each: random Default.params shapes, F0 and velum varying per frame."""
    rng = np.random.default_rng(seed)
    names = shapes["names"]
    fr = shapes["frames"][rng.integers(0, len(names), size=(n, n_frames))].copy()
    if nasal:
        fr["velum_area_cm2"] = rng.random((n, n_frames)) * (rng.random((n, 1)) < 0.6)
    gl = np.tile(np.array([120.0, 8000.0, 0.01, 0.01, 0.0, -40.0]), (n, n_frames, 1))
    gl[:, :, 0] = rng.uniform(90.0, 220.0, size=(n, n_frames))
    gl[:, :, 5] = rng.uniform(-40.0, -20.0, size=(n, n_frames))
    return fr, gl


def _chunked(ctx, fr, gl, frame_len, cuts, fs, srf, opts, seeds=None, want_flow=False, via_checkpoint=False):
    """frames 1.. given to a stream in pieces ending at the frame indices `cuts` (frame 0 primes with the first piece)."""
    import areafunctionsynthesis_b200 as afs
    st = afs.TdsStream(ctx, len(fr), fs, srf, opts=opts, seeds=seeds)
    audio, flow, status = [], [], np.zeros(len(fr), dtype=np.uint32)
    lo = 0
    for hi in list(cuts) + [fr.shape[1]]:
        if hi <= lo:
            continue
        fl = ([0] if lo == 0 else []) + list(frame_len[max(lo - 1, 0):hi - 1])
        a, u, s = st.synthesize(fr[:, lo:hi], gl[:, lo:hi], fl, want_flow=want_flow)
        audio.append(a); flow.append(u); status |= s
        lo = hi
        if via_checkpoint:  # continue on a NEW stream from the saved blob
            blob = st.save()
            st.close()
            st = afs.TdsStream(ctx, len(fr), fs, srf, opts=opts, seeds=None)
            st.load(blob)
    assert st.samples == int(np.sum(frame_len))
    st.close()
    return np.concatenate(audio, axis=1), (np.concatenate(flow, axis=1) if want_flow else None), status


@pytest.mark.parametrize("noise", [False, True])
def test_chunked_equals_one_call_and_oracle(ctx, oracle, shapes, noise):
    """Six control frames, 0.12 s, nasal tracts, more tracts than one CTA holds; chunk boundaries after every frame and
    after frames 2 and 3 only."""
    import areafunctionsynthesis_b200 as afs
    from oracle.pyoracle import tds_options_default as oracle_opts
    n, F = 34, 6
    fr, gl = _utterances(shapes, n, 31 + noise, F)
    frame_len = [441, 300, 1, 1100, 804]
    seeds = np.arange(100, 100 + n, dtype=np.uint32)
    o = afs.tds_options_default(noise=noise)
    ctx.debug_set("tds.small_batch_max", 0)
    try:
        a1, u1, s1 = ctx.tds_synthesize(fr, gl, frame_len, 44100.0, 2, opts=o, seeds=seeds, want_flow=True)
    finally:
        ctx.debug_set("tds.small_batch_max", -1)
    assert not s1.any()
    errs = {}
    for cuts in ([2, 3, 4, 5], [3, 4], [1]):
        a2, u2, s2 = _chunked(ctx, fr, gl, frame_len, cuts, 44100.0, 2, o, seeds=seeds, want_flow=True)
        assert not s2.any()
        assert a2.shape == a1.shape and u2.shape == u1.shape
        errs[str(cuts)] = max(max(rel_err(a2[b], a1[b]), rel_err(u2[b], u1[b])) for b in range(n))
    # (measured: 0 with the noise sources off, 5e-10 .. 7e-10 with them on -- also when the whole utterance is ONE chunk,
    #  so it is not the chunking: the chunked path runs its own instantiation of the kernel (state in / state out), and
    #  the amplitudes of the noise sources follow the state to rounding level only)
    print(f"chunked vs one call, noise {'on' if noise else 'off'}: max rel difference per chunking {errs}")
    assert max(errs.values()) <= CHUNK_TOL, errs
    sel = [0, 9, 33]
    ao, uo = oracle.tds_synthesize(fr[sel], gl[sel], frame_len, 44100.0, 2, opts=oracle_opts(noise=noise), seeds=seeds[sel])
    for k, b in enumerate(sel):
        assert rel_err(a2[b], ao[k]) <= 1e-6 and rel_err(u2[b], uo[k]) <= 1e-6, b


@pytest.mark.parametrize("noise", [False, True])
@pytest.mark.parametrize("srf", [1, 3])
def test_chunks_with_odd_step_counts(ctx, shapes, srf, noise):
    """The step-parity buffers of the kernel (glottis geometry, filter histories of the non-default options) across
    chunks of an odd number of simulation steps; transvelar coupling + variable entrance loss + fossa + a flow source on
    (the OPT x STATE instantiations of the kernel)."""
    import areafunctionsynthesis_b200 as afs
    n, F = 5, 5
    fr, gl = _utterances(shapes, n, 77, F)
    frame_len = [7, 1, 33, 260]
    o = afs.tds_options_default(noise=noise)
    o.transvelar_coupling = 1; o.glottis_loss_option = 2; o.piriform_fossa = 1
    o.flow_source_section = 45; o.flow_source_cm3_s = 20.0
    fs = 22050.0 * srf
    a1, u1, _ = ctx.tds_synthesize(fr, gl, frame_len, fs, srf, opts=o, want_flow=True)
    a2, u2, _ = _chunked(ctx, fr, gl, frame_len, [2, 3, 4], fs, srf, o, want_flow=True)
    e = max(rel_err(a2, a1), rel_err(u2, u1))
    print(f"odd chunks, sr_factor {srf}, noise {'on' if noise else 'off'}: max rel difference {e:.2e}")
    assert e <= CHUNK_TOL


def test_checkpoint_resume_and_reset(ctx, shapes):
    """save -> new stream -> load continues bit for bit; reset gives a fresh model; a blob of another batch is refused."""
    import areafunctionsynthesis_b200 as afs
    n, F = 3, 4
    fr, gl = _utterances(shapes, n, 5, F)
    frame_len = [500, 400, 300]
    o = afs.tds_options_default(noise=True)
    seeds = np.array([7, 8, 9], dtype=np.uint32)
    a_ref, _, _ = _chunked(ctx, fr, gl, frame_len, [2, 3], 44100.0, 2, o, seeds=seeds)
    a_ckp, _, _ = _chunked(ctx, fr, gl, frame_len, [2, 3], 44100.0, 2, o, seeds=seeds, via_checkpoint=True)
    assert np.array_equal(a_ref, a_ckp)
    st = afs.TdsStream(ctx, n, 44100.0, 2, opts=o, seeds=seeds)
    a, _, _ = st.synthesize(fr, gl, [0] + frame_len)
    assert np.array_equal(a, _chunked(ctx, fr, gl, frame_len, [], 44100.0, 2, o, seeds=seeds)[0])
    st.reset()
    assert st.samples == 0
    b, _, _ = st.synthesize(fr, gl, [0] + frame_len)
    assert np.array_equal(a, b)
    blob = st.save()
    other = afs.TdsStream(ctx, n + 1, 44100.0, 2, opts=o)
    with pytest.raises(afs.AfsError):
        other.load(blob)
    other.close(); st.close()


def test_long_utterance_in_bounded_memory(ctx, shapes):
    """10 s of audio for 64 tracts, one second per call: the device never holds more than one chunk; every tract stays
    finite and keeps phonating (the glottis state carries over, so the F0 contour is continuous)."""
    import areafunctionsynthesis_b200 as afs
    n = 64
    fr, gl = _utterances(shapes, n, 11, 11, nasal=False)
    fr = shapes["frames"][np.array([0, 4, 8, 12, 16])[np.random.default_rng(3).integers(0, 5, size=(n, 11))]].copy()  # a: e: i: o: u:
    o = afs.tds_options_default(noise=True)
    st = afs.TdsStream(ctx, n, 44100.0, 2, opts=o)
    st.synthesize(fr[:, :1], gl[:, :1], [0])
    rms = []
    for k in range(1, 11):
        a, _, s = st.synthesize(fr[:, k:k + 1], gl[:, k:k + 1], [22050])
        assert not s.any() and np.isfinite(a).all()
        rms.append(np.sqrt(np.mean(a[:, 2205:] ** 2, axis=1)))
    assert st.samples == 220500
    rms = np.array(rms)
    assert (rms[1:] > 1e-3).all(), rms.min()
    st.close()


def test_chunk_repeated_after_pool_overflow(ctx, shapes):
    """The record pool of the noise pre-pass turns out too small in the middle of a stream: the chunk is repeated with
    a larger pool FROM THE SAME STATE (two state buffers), the audio is that of an undisturbed stream."""
    import areafunctionsynthesis_b200 as afs
    n = len(shapes["names"])
    rng = np.random.default_rng(5)
    B = 6400
    fr = np.tile(shapes["frames"][rng.integers(0, n, size=(40, 3))], (160, 1))
    gl = np.tile(np.array([120.0, 8000.0, 0.01, 0.01, 0.0, -15.0]), (B, 3, 1))
    seeds = np.tile(np.arange(1, 41, dtype=np.uint32), 160)
    on = afs.tds_options_default(noise=True)
    ref, _, _ = _chunked(ctx, fr, gl, [400, 400], [2], 44100.0, 2, on, seeds=seeds)
    st = afs.TdsStream(ctx, B, 44100.0, 2, opts=on, seeds=seeds)
    a1, _, s1 = st.synthesize(fr[:, :2], gl[:, :2], [0, 400])
    ctx.debug_set("tds.noise_run_pool_factor", 2)   # (the chunk below needs more records than 2 per piece + 4096)
    l0 = ctx.launch_count
    try:
        a2, _, s2 = st.synthesize(fr[:, 2:], gl[:, 2:], [400])
    finally:
        ctx.debug_set("tds.noise_run_pool_factor", 32)
    assert ctx.launch_count - l0 >= 4, "pre-pass + time loop at least twice: the overflow must have happened"
    st.close()
    assert not s1.any() and not s2.any()
    assert np.array_equal(np.concatenate([a1, a2], axis=1), ref)
    assert np.array_equal(ref[:40], ref[-40:])
