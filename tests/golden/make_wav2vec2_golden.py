"""Golden CTC logits of the THIRD-PARTY reference for NerfASR.__frame_to_text (nerfasr.py:128-143): HF transformers'
Wav2Vec2ForCTC (the class AutoModelForCTC resolves to for cpierse/wav2vec2-large-xlsr-53-esperanto) with the feature extractor's
do_normalize, run on CPU in fp32 in the build container (transformers 5.5) on the seeded weights / audio of tests/helpers
-> wav2vec2_golden.npz.  The checkpoint itself is external (no network): weights are seeded; strict load checks names + shapes."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from helpers import W2V_SMALL, W2V_XLSR53, seeded_w2v_state, synthetic_speech      # noqa: E402
from transformers import Wav2Vec2Config, Wav2Vec2FeatureExtractor, Wav2Vec2ForCTC   # noqa: E402

out = {}
for name, c, seed in (("small", W2V_SMALL, 21), ("xlsr53", W2V_XLSR53, 22)):
    cfg = Wav2Vec2Config(vocab_size=c["vocab"], hidden_size=c["hidden"], num_hidden_layers=c["layers"], num_attention_heads=c["heads"],
                         intermediate_size=c["inter"], feat_extract_norm="layer", do_stable_layer_norm=True, conv_bias=True,
                         conv_dim=c["conv_dim"], conv_stride=c["conv_stride"], conv_kernel=c["conv_kernel"],
                         num_conv_pos_embeddings=c["pos_k"], num_conv_pos_embedding_groups=c["pos_groups"], hidden_dropout=0.0,
                         attention_dropout=0.0, activation_dropout=0.0, feat_proj_dropout=0.0, final_dropout=0.0, layerdrop=0.0)
    model = Wav2Vec2ForCTC(cfg).eval()
    print(name, model.load_state_dict({k: torch.from_numpy(v) for k, v in seeded_w2v_state(seed, c).items()}, strict=True))
    fe = Wav2Vec2FeatureExtractor(feature_size=1, sampling_rate=16000, padding_value=0.0, do_normalize=True, return_attention_mask=True)
    audio = synthetic_speech(8960, seed)
    inputs = fe(audio, sampling_rate=16000, return_tensors="pt", padding=True)       # nerfasr.py:131
    with torch.no_grad():
        logits = model(inputs.input_values).logits[0].numpy()                        # nerfasr.py:134-138
    print(name, logits.shape, float(np.abs(logits).mean()), float(logits.std()))
    out[name + "_logits"] = logits.astype(np.float32)
np.savez_compressed(os.path.join(HERE, "wav2vec2_golden.npz"), **out)
print(os.path.getsize(os.path.join(HERE, "wav2vec2_golden.npz")) / 1e3, "KB")
