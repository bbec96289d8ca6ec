"""GPU: mf_wav2vec2_logits (waveform normalisation + conv feature encoder + positional conv + transformer + lm_head, one C-ABI
call) against golden logits of HF transformers' Wav2Vec2ForCTC (tests/golden/make_wav2vec2_golden.py: fp32 CPU run, seeded
weights), for a small config of identical structure and for the XLSR-53-large shape NerfASR loads."""
import os

import numpy as np
import pytest

from helpers import GOLD, W2V_SMALL, W2V_XLSR53, seeded_w2v_state, synthetic_speech

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

G = np.load(os.path.join(GOLD, "wav2vec2_golden.npz"))


def _rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.mark.parametrize("name,cfg,seed", [("small", W2V_SMALL, 21), ("xlsr53", W2V_XLSR53, 22)])
def test_logits_match_hf_golden(name, cfg, seed):
    from mere_fusion_b200.wav2vec2 import Wav2Vec2Engine
    eng = Wav2Vec2Engine(seeded_w2v_state(seed, cfg), cfg)
    audio = torch.from_numpy(synthetic_speech(8960, seed)).cuda()
    got = eng.logits(audio).cpu().numpy()
    ref = G[name + "_logits"]
    assert got.shape == ref.shape == (27, 44)
    # bf16 activations / fp32 accumulation through 7 conv + 24 transformer layers vs fp32: stated tolerance 3 % relative L2,
    # and the per-frame argmax (what a CTC decoder would read) agrees on at least 25 of 27 frames
    r = _rel(got, ref)
    assert r < 3e-2, r
    assert int((got.argmax(1) == ref.argmax(1)).sum()) >= 25
    got2 = eng.logits(audio).cpu().numpy()
    assert np.array_equal(got, got2)                       # replay is bit-reproducible
    print(f"wav2vec2 {name}: rel L2 {r:.4f}, launches {eng.last_launches}")


def test_nerfasr_with_the_gpu_acoustic_model():
    """NerfASR.run_step -> feature_fn on the GPU engine: rows [l : T - r + 1] of the window's logits land in the ring"""
    from mere_fusion_b200.plugin.nerfasr import NerfASR
    from mere_fusion_b200.wav2vec2 import Wav2Vec2Engine
    from test_plugin_cpu import make_opt
    eng = Wav2Vec2Engine(seeded_w2v_state(21, W2V_SMALL), W2V_SMALL)
    asr = NerfASR(make_opt(), None, feature_fn=eng.feature_fn, device="cuda")
    wav = synthetic_speech(28 * 320 * 2, 3)
    for i in range(56):
        asr.put_audio_frame(wav[i * 320:(i + 1) * 320])
    asr.warm_up()
    for _ in range(16):
        asr.run_step()
    f = asr.get_next_feat()
    assert f.shape == (8, 44, 16) and f.is_cuda and bool(torch.isfinite(f).all()) and float(f.abs().max()) > 0
    import ctypes
    from mere_fusion_b200._lib import lib
    out = torch.empty(27, 44, device="cuda")
    bad = lib().mf_wav2vec2_logits(eng.ctx.handle, ctypes.c_void_p(out.data_ptr()), 4000, ctypes.c_void_p(out.data_ptr()), None)
    assert bad == -1 and b"8960" in lib().mf_last_error(eng.ctx.handle)          # wrong window length is refused


@pytest.mark.parametrize("name,cfg,seed,B", [("small", W2V_SMALL, 21, 3), ("xlsr53", W2V_XLSR53, 22, 4)])
def test_batched_windows_match_single_windows(name, cfg, seed, B):
    """mf_wav2vec2_logits_batch: the windows of B sessions in one pass give each session the logits of its own window (per-window
    normalisation statistics, no leakage across the batch through the 128-tap positional conv or the attention); window 0 is the
    golden window"""
    from mere_fusion_b200._lib import MfError
    from mere_fusion_b200.wav2vec2 import Wav2Vec2Engine
    eng = Wav2Vec2Engine(seeded_w2v_state(seed, cfg), cfg, max_batch=B)
    wins = np.stack([synthetic_speech(8960, seed)] + [synthetic_speech(8960, 50 + i) * (0.5 + i) for i in range(B - 1)])
    audio = torch.from_numpy(wins).cuda()
    got = eng.logits_batch(audio).cpu().numpy()
    assert got.shape == (B, 27, 44)
    ref0 = G[name + "_logits"]
    assert _rel(got[0], ref0) < 3e-2 and int((got[0].argmax(1) == ref0.argmax(1)).sum()) >= 25
    for i in range(B):
        solo = eng.logits(audio[i]).cpu().numpy()
        # batch size changes tile shapes / split-K: same operands, another summation order
        assert _rel(got[i], solo) < 1.5e-2, (i, _rel(got[i], solo))
        assert int((got[i].argmax(1) == solo.argmax(1)).sum()) >= 25
    assert _rel(got[1], got[0]) > 0.05                      # the windows really differ
    assert np.array_equal(eng.logits_batch(audio).cpu().numpy(), got)          # replay: bit-identical
    with pytest.raises(MfError):
        eng.logits_batch(torch.zeros(B + 1, 8960, device="cuda"))
