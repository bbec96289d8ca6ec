"""wav2vec2 engine: time per window (CUDA events), for launch-list captures"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import W2V_XLSR53, seeded_w2v_state, synthetic_speech
from mere_fusion_b200.wav2vec2 import Wav2Vec2Engine
sd = seeded_w2v_state(22, W2V_XLSR53)
eng = Wav2Vec2Engine(sd, W2V_XLSR53, device=0, max_batch=1)
a = torch.from_numpy(synthetic_speech(8960, 3)).cuda()
for _ in range(5):
    out = eng.logits(a)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    out = eng.logits(a)
e1.record(); torch.cuda.synchronize()
print(f"B=1: {e0.elapsed_time(e1) / 20:.3f} ms per window, launches {eng.last_launches}", flush=True)
import ctypes
from mere_fusion_b200._lib import lib
ts = (ctypes.c_ulonglong * 11)()
fn = lib().mf_debug_w2v_phase_ns
if fn(eng.ctx.handle, ts, 11) == 0:
    t = [int(v) for v in ts]
    names = ["P1 LN1+QKV", "barrier", "P2 attention", "barrier", "P3 out-proj", "barrier", "P4 LN2+FFN1", "barrier", "P5 FFN2", "barrier"]
    print("layer 1, CTA 0 (us): " + ", ".join(f"{n} {(t[i + 1] - t[i]) / 1e3:.1f}" for i, n in enumerate(names)) + f"; layer {(t[10] - t[0]) / 1e3:.1f}", flush=True)
eng4 = Wav2Vec2Engine(sd, W2V_XLSR53, device=0, max_batch=4)
a4 = torch.from_numpy(np.stack([synthetic_speech(8960, i) for i in range(4)])).cuda()
for _ in range(3):
    eng4.logits_batch(a4)
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    eng4.logits_batch(a4)
e1.record(); torch.cuda.synchronize()
print(f"B=4: {e0.elapsed_time(e1) / 20:.3f} ms per pass", flush=True)
print("checksum", float(out.float().abs().sum()))
