"""GPU: mf_whisper_features (log-mel + Whisper encoder + embedding gather, one C-ABI call) against the golden features of the
vendored reference (tests/golden/whisper_golden.npz, fp32 CPU run of musetalk/whisper) and against the oracle."""
import os

import numpy as np
import pytest

from helpers import GOLD, WHISPER_TINY, seeded_whisper_state, synthetic_speech

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

G = np.load(os.path.join(GOLD, "whisper_golden.npz"))


@pytest.fixture(scope="module")
def engine():
    from mere_fusion_b200.whisper import WhisperEngine
    return WhisperEngine(seeded_whisper_state(7), WHISPER_TINY)


def _rel(got, ref):
    return float(np.linalg.norm(got - ref) / np.linalg.norm(ref))


@pytest.mark.parametrize("name,n,seed", [("win52", 52 * 320, 0), ("win20", 20 * 320, 1), ("odd", 9999, 2)])
def test_features_match_reference_golden(engine, name, n, seed):
    audio = torch.from_numpy(synthetic_speech(n, seed)).cuda()
    got = engine.features(audio).cpu().numpy()
    ref = G[name + "_feat"].astype(np.float32)
    assert got.shape == ref.shape
    # row 0 of axis 1 is conv2 + GELU + positional embedding (no attention yet): tight
    assert _rel(got[:, 0], ref[:, 0]) < 6e-3
    # bf16 activations / fp32 accumulation through 4 residual attention blocks vs the fp32 reference (the reference's own GPU
    # path is fp16): stated tolerance 2 % relative L2 per embedding level, max abs error below 3 % of the largest value (~5 bf16 ulps at |x| ~ 9)
    for j in range(ref.shape[1]):
        assert _rel(got[:, j], ref[:, j]) < 2e-2, (j, _rel(got[:, j], ref[:, j]))
    assert np.abs(got - ref).max() < 0.03 * np.abs(ref).max()
    print(f"whisper {name}: rel L2 per level {[round(_rel(got[:, j], ref[:, j]), 5) for j in range(ref.shape[1])]}, launches {engine.last_launches}")


def test_features_replay_and_resize(engine):
    """graph replay with a different sample count (grid of the log-mel kernel and T change) and back"""
    a52 = torch.from_numpy(synthetic_speech(52 * 320, 0)).cuda()
    a20 = torch.from_numpy(synthetic_speech(20 * 320, 1)).cuda()
    r1 = engine.features(a52).clone()
    r2 = engine.features(a20).clone()
    r3 = engine.features(a52).clone()
    torch.cuda.synchronize()
    assert r1.shape == (52, 5, 384) and r2.shape == (20, 5, 384)
    assert torch.equal(r1, r3)
    assert _rel(r2.cpu().numpy(), G["win20_feat"].astype(np.float32)) < 2e-2


def test_audio2feature_mirror(engine):
    """Audio2Feature.audio2feat + feature2chunks exactly as MuseASR.run_step calls them (museasr.py:26-27)"""
    from mere_fusion_b200.whisper import Audio2Feature
    a2f = Audio2Feature(engine=engine)
    feat = a2f.audio2feat(synthetic_speech(52 * 320, 0))
    assert feat.shape == (52, 5, 384) and feat.dtype == np.float32
    chunks = a2f.feature2chunks(feature_array=feat, fps=25, batch_size=16, start=5)
    assert len(chunks) == 16 and chunks[0].shape == (50, 384)
    ref = G["win52_feat"].astype(np.float32)
    assert _rel(np.stack(chunks), np.stack(a2f.feature2chunks(feature_array=ref, fps=25, batch_size=16, start=5))) < 2e-2


def test_device_chunks_equal_host_chunks(engine):
    """audio2chunks_device (gather on the GPU) == feature2chunks(audio2feat(...)) (the reference's host path)"""
    from mere_fusion_b200.whisper import Audio2Feature
    a2f = Audio2Feature(engine=engine)
    audio = synthetic_speech(52 * 320, 0)
    host = np.stack(a2f.feature2chunks(feature_array=a2f.audio2feat(audio), fps=25.0, batch_size=16, start=5.0)).astype(np.float16)
    dev = a2f.audio2chunks_device(audio, fps=25.0, batch_size=16, start=5.0)
    assert dev.shape == (16, 50, 384) and dev.dtype == torch.float16
    assert np.array_equal(dev.cpu().numpy(), host)


def test_rejects_bad_arguments(engine):
    from mere_fusion_b200._lib import MfError
    with pytest.raises(MfError):
        engine.features(torch.zeros(100, device="cuda"))                 # shorter than the reflect padding
    with pytest.raises(MfError):
        engine.features(torch.zeros(3001 * 160, device="cuda"))          # more than one 30 s segment
