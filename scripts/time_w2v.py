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
ts = (ctypes.c_ulonglong * 24)()
fn = lib().mf_debug_w2v_phase_ns
if fn(eng.ctx.handle, ts, 24) == 0:
    t = [int(v) for v in ts]
    names = ["P1 LN1+QKV", "barrier", "P2 attention", "barrier", "P3 out-proj", "barrier", "P4 LN2+FFN1", "barrier", "P5 FFN2", "barrier"]
    print("layer 1, CTA 0 (us): " + ", ".join(f"{n} {(t[i + 1] - t[i]) / 1e3:.1f}" for i, n in enumerate(names)) + f"; layer {(t[10] - t[0]) / 1e3:.1f}", flush=True)
    print(f"   inside LN1: x loaded + first reduction {(t[16] - t[0]) / 1e3:.1f}, rest of the row + remote stores {(t[17] - t[16]) / 1e3:.1f}, cluster sync {(t[15] - t[17]) / 1e3:.1f}")
    print(f"   inside P1: LN (thread 0's rows) {(t[15] - t[0]) / 1e3:.1f}, rest of LN + W wait {(t[11] - t[15]) / 1e3:.1f}, MMA {(t[12] - t[11]) / 1e3:.1f}, reduce + epilogue {(t[1] - t[12]) / 1e3:.1f};  "
          f"inside P5: prologue {(t[13] - t[8]) / 1e3:.1f}, chunk loop {(t[14] - t[13]) / 1e3:.1f}, reduce + epilogue {(t[9] - t[14]) / 1e3:.1f}", flush=True)
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
