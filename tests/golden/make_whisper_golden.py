"""Golden features of the REFERENCE (vendored) Whisper path: musetalk/whisper/audio2feature.py:99-112 audio2feat ->
whisper/transcribe.py -> whisper/audio.py log_mel_spectrogram + whisper/model.py AudioEncoder, imported from /root/reference
and run on CPU (fp32) in the build container, on the seeded weights / audio of tests/helpers -> whisper_golden.npz.

`ffmpeg` (file decoding only, whisper/audio.py:5) and `soundfile` (audio2feature.py:3) are absent from this image and unused
on the ndarray path: they are stubbed.  ./models/whisper/tiny.pt is external: weights are seeded."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
for m in ("ffmpeg", "soundfile"):
    sys.modules.setdefault(m, types.ModuleType(m))
from helpers import WHISPER_TINY, seeded_whisper_state, synthetic_speech      # noqa: E402
from musetalk.whisper.whisper.audio import log_mel_spectrogram                  # noqa: E402
from musetalk.whisper.whisper.model import ModelDimensions, Whisper             # noqa: E402
from musetalk.whisper.audio2feature import Audio2Feature                        # noqa: E402

dims = ModelDimensions(n_vocab=51865, n_text_ctx=448, n_text_state=384, n_text_head=6, n_text_layer=4, **WHISPER_TINY)
model = Whisper(dims).eval()
sd = {k: torch.from_numpy(v) for k, v in seeded_whisper_state(7).items()}
print(model.encoder.load_state_dict(sd, strict=True))           # key names + shapes match the reference encoder

a2f = Audio2Feature.__new__(Audio2Feature)                       # skip load_model(path): no checkpoint file
a2f.model = model
out = {}
for name, n, seed in (("win52", 52 * 320, 0), ("win20", 20 * 320, 1), ("odd", 9999, 2)):
    audio = synthetic_speech(n, seed)
    with torch.no_grad():
        feat = a2f.audio2feat(audio)
    mel = log_mel_spectrogram(audio).numpy()
    print(name, "mel", mel.shape, "feat", feat.shape, feat.dtype, float(np.abs(feat).mean()), float(feat.std()))
    out[name + "_feat"] = feat.astype(np.float16)
    out[name + "_mel"] = mel.astype(np.float32)
# the slicing arithmetic (audio2feature.py:16-45, 82-97) on the 52-row feature, fps 25, batch 16, start 5 (museasr.py:27)
feat = out["win52_feat"].astype(np.float32)
chunks = a2f.feature2chunks(feature_array=feat, fps=25, batch_size=16, start=5)
out["win52_chunk_idx"] = np.array([a2f.get_sliced_feature(feat, i + 5, [2, 2], 25)[1] for i in range(16)], np.int32)
out["win52_chunks_sum"] = np.array([c.sum() for c in chunks], np.float64)
print("chunks", len(chunks), chunks[0].shape, out["win52_chunk_idx"][0], out["win52_chunk_idx"][-1])
np.savez_compressed(os.path.join(HERE, "whisper_golden.npz"), **out)
print(os.path.getsize(os.path.join(HERE, "whisper_golden.npz")) / 1e3, "KB")
