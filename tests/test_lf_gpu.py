"""GPU parity of the batched LF vowel synthesis (afs_synthesize_vowels_lf) against golden runs of
Synthesizer::synthesizeVowelLf restated over the unmodified reference classes (tools/make_golden_lf.py).
Bar: the int16 samples are IDENTICAL and the convolution output (pressure) is bit-identical -- impulse response,
pulse train (host libm, reference expressions), tap order and the Chebyshev recursion all repeat the reference."""
import numpy as np
import pytest

from conftest import load_golden, random_frames

pytestmark = pytest.mark.gpu


def test_lf_vowels_identical_to_reference(ctx, shapes):
    g = load_golden("lf_vowels.npz")
    fr = shapes["frames"][g["sel"]]
    for is_long in (1, 0):
        audio, pr = ctx.synthesize_vowels_lf(fr, g["lf"], bool(is_long), want_pressure=True)
        assert audio.dtype == np.int16 and audio.shape == g[f"audio_{is_long}"].shape
        assert np.array_equal(pr[:, ::16], g[f"pressure_{is_long}"]), is_long
        assert np.array_equal(audio, g[f"audio_{is_long}"]), is_long


def test_lf_vowels_batch_properties(ctx, shapes):
    import areafunctionsynthesis_b200 as afs
    g = load_golden("lf_vowels.npz")
    # a large ragged batch: every vowel depends on its own inputs only (order / split invariance)
    fr = random_frames(300, seed=5)
    rng = np.random.default_rng(5)
    lf = np.stack([np.full(300, 300.0), rng.uniform(80, 250, 300), rng.uniform(0.35, 0.75, 300), rng.uniform(1.5, 3.5, 300),
                   rng.uniform(0.0, 0.12, 300)], axis=1)
    a = ctx.synthesize_vowels_lf(fr, lf, True)
    perm = rng.permutation(300)
    b = ctx.synthesize_vowels_lf(fr[perm], lf[perm], True)
    assert np.array_equal(a[perm], b)
    c = np.concatenate([ctx.synthesize_vowels_lf(fr[:7], lf[:7], True), ctx.synthesize_vowels_lf(fr[7:], lf[7:], True)])
    assert np.array_equal(a, c)
    assert np.abs(a).max() > 1000 and a[:, :10].max() == 0  # silence before the first pulse at sample 10
    # fast arithmetic: a couple of LSB at most
    o = afs.tl_options_default(); o.arithmetic = afs.TL_ARITHMETIC_FAST
    d = ctx.synthesize_vowels_lf(fr, lf, True, opts=o)
    assert np.max(np.abs(d.astype(np.int32) - a.astype(np.int32))) <= 1
    # empty batch / bad arguments
    assert ctx.synthesize_vowels_lf(fr[:0], lf[:0], True).shape[0] == 0
    bad = lf[:2].copy(); bad[0, 1] = 5.0  # pulse longer than the reference's ring buffer
    with pytest.raises(afs.AfsError):
        ctx.synthesize_vowels_lf(fr[:2], bad, True)
    with pytest.raises(afs.AfsError):
        ctx.synthesize_vowels_lf(fr[:2], lf[:2], True, sampling_rate_hz=8000.0)
