"""GPU parity of afs_play_target_sequences against golden runs of Synthesizer::playTargetSequence restated over the
unmodified reference classes (tools/make_golden_targets.py): one control frame per output sample, parameter sets
rasterised on the device.  Bar: 1e-6 relative max-abs (the north star's TDS bar); the device rasteriser's areas differ
from glibc's by an ulp (pow / cos), everything else follows the reference."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("noise", [0, 1])
def test_target_sequences_match_reference(ctx, noise):
    import areafunctionsynthesis_b200 as afs
    g = load_golden("target_sequence.npz")
    audio, status = ctx.play_target_sequences(g["shapes"], g["stationary"], g["transition"], g["glottis"], 22050.0, 2,
                                              opts=afs.tds_options_default(noise=bool(noise)), seeds=g["seeds"])
    ref = g[f"audio_noise{noise}"]
    assert audio.shape == ref.shape and not status.any()
    assert np.all(audio[:, 0] == 0.0)
    for k in range(len(ref)):
        err = np.max(np.abs(audio[k] - ref[k])) / np.max(np.abs(ref[k]))
        assert err <= 1e-6, (k, err)


def test_target_sequences_batch_properties(ctx):
    import areafunctionsynthesis_b200 as afs
    g = load_golden("target_sequence.npz")
    o = afs.tds_options_default(noise=False)
    sh = np.concatenate([g["shapes"], g["shapes"][::-1], g["shapes"][[1]]])
    st, tr = [0.03, 0.02, 0.02, 0.03], [0.03, 0.03, 0.03]
    a, s = ctx.play_target_sequences(sh, st, tr, g["glottis"], 22050.0, 2, opts=o)
    assert not s.any() and np.array_equal(a[0], a[5]) and np.array_equal(a[1], a[4]) and np.array_equal(a[1], a[6])
    b, _ = ctx.play_target_sequences(sh[2:3], st, tr, g["glottis"], 22050.0, 2, opts=o)
    assert np.array_equal(b[0], a[2])
    assert ctx.play_target_sequences(sh[:0], st, tr, g["glottis"])[0].shape[0] == 0
    with pytest.raises(afs.AfsError):
        ctx.play_target_sequences(sh, [0, 0, 0, 0], [0, 0, 0], g["glottis"])
