"""quick device-side timing of mf_wav2lip_forward at B=16 (not the bench)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import seeded_wav2lip_state, wav2lip_inputs
from mere_fusion_b200.wav2lip import Wav2LipEngine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = int(sys.argv[2]) if len(sys.argv) > 2 else 50
S = int(sys.argv[3]) if len(sys.argv) > 3 else 96          # 96: reference generator; 256: the extended one (configs[1])
eng = Wav2LipEngine(seeded_wav2lip_state(2, face_hw=S), max_batch=B, face_hw=S)
mel, faces = wav2lip_inputs(B, S=S)
mel, faces = torch.from_numpy(mel).cuda(), torch.from_numpy(faces).cuda()
out = torch.empty_like(faces)
for _ in range(5):
    eng.forward(mel, faces, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    eng.forward(mel, faces, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
fl = eng.flops_per_frame * B
print(f"S={S} B={B}: {ms:.3f} ms/batch = {B*1000/ms:.0f} frames/s; {fl/ms/1e9:.1f} TFLOP/s algorithmic; launches {eng.last_launches}", flush=True)
