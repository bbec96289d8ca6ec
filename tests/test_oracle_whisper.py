"""CPU: the Whisper oracle (oracle/whisper_oracle.py) and the host-side slicing mirror against the golden vectors produced by
the vendored reference itself (tests/golden/make_whisper_golden.py)."""
import os

import numpy as np
import pytest

from helpers import GOLD, WHISPER_TINY, seeded_whisper_state, synthetic_speech

G = np.load(os.path.join(GOLD, "whisper_golden.npz"))
CASES = (("win52", 52 * 320, 0), ("win20", 20 * 320, 1), ("odd", 9999, 2))


@pytest.mark.parametrize("name,n,seed", CASES)
def test_log_mel_matches_reference(name, n, seed):
    from oracle import whisper_oracle as O
    mel = O.log_mel(synthetic_speech(n, seed))
    ref = G[name + "_mel"]
    assert mel.shape == ref.shape
    assert np.abs(mel - ref).max() < 2e-4          # fp32 FFT vs torch.stft: last-bit differences through log10


def test_filterbank_is_the_reference_asset():
    from oracle import whisper_oracle as O
    from mere_fusion_b200.whisper_pack import whisper_filters
    assert np.array_equal(O.mel_filters(), whisper_filters())
    p = "/root/reference/musetalk/whisper/whisper/assets/mel_filters.npz"
    if os.path.exists(p):                          # build container only
        ref = np.load(p)["mel_80"]
        assert np.abs(ref - O.mel_filters()).max() < 4e-9


@pytest.mark.parametrize("name,n,seed", CASES[:2])
def test_audio2feat_matches_reference(name, n, seed):
    from oracle import whisper_oracle as O
    feat = O.audio2feat(seeded_whisper_state(7), synthetic_speech(n, seed), WHISPER_TINY)
    ref = G[name + "_feat"].astype(np.float32)
    assert feat.shape == ref.shape
    err = np.abs(feat - ref)
    assert err.max() < 2e-2 and err.mean() < 1e-3  # golden stored as fp16 (|x| up to ~10 -> 4e-3 rounding)


def test_slicing_matches_reference():
    from oracle import whisper_oracle as O
    from mere_fusion_b200 import whisper as Wh
    feat = G["win52_feat"].astype(np.float32)
    for mod in (O, Wh):
        idx = np.array([mod.get_sliced_feature(feat, i + 5, [2, 2], 25)[1] for i in range(16)], np.int32)
        assert np.array_equal(idx, G["win52_chunk_idx"])
        chunks = mod.feature2chunks(feature_array=feat, fps=25, batch_size=16, start=5)
        assert len(chunks) == 16 and chunks[0].shape == (50, 384)
        assert np.allclose([c.sum() for c in chunks], G["win52_chunks_sum"], rtol=1e-6)
    # clamping at both ends (audio2feature.py:36-38)
    assert O.get_sliced_feature(feat, 0, [2, 2], 25)[1] == [0, 0, 0, 0, 0, 1, 2, 3, 4, 5]
    assert O.get_sliced_feature(feat, 25, [2, 2], 25)[1] == [46, 47, 48, 49, 50, 51, 51, 51, 51, 51]
