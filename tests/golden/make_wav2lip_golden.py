"""Golden output of the REFERENCE Wav2Lip nn.Module (imported from /root/reference; run in the
build container) on the seeded weights / inputs of tests/helpers -> wav2lip_golden_b2.npz."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
from helpers import seeded_wav2lip_state, wav2lip_inputs          # noqa: E402
from wav2lip.models import Wav2Lip                                # noqa: E402  (the reference module)

model = Wav2Lip().eval()
missing = model.load_state_dict(seeded_wav2lip_state(2), strict=True)   # key names + shapes match the reference
print(missing)
mel, faces = wav2lip_inputs(2)
img = faces.copy()
masked = img.copy()
masked[:, img.shape[1] // 2:] = 0                                  # lipreal.py:110-113
x = np.concatenate((masked, img), axis=3) / 255.
x = torch.FloatTensor(np.transpose(x, (0, 3, 1, 2)))
with torch.no_grad():
    pred = model(torch.FloatTensor(mel), x)
pred = pred.cpu().numpy().transpose(0, 2, 3, 1)
print("pred", pred.shape, pred.min(), pred.max(), pred.mean(), pred.std())
np.savez_compressed(os.path.join(HERE, "wav2lip_golden_b2.npz"), pred=pred.astype(np.float32))
print(os.path.getsize(os.path.join(HERE, "wav2lip_golden_b2.npz")) / 1e3, "KB")
