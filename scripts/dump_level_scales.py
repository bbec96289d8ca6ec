"""GPU box: dump the device-evaluated grid level scales -> gpurun_out/ernerf_level_scales.json
(committed as tests/golden/ernerf_level_scales.json)."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mere_fusion_b200._lib import Context, lib
ctx = Context(0)
out = {}
for name, base, desired, L in (("head", 64, 512, 12), ("torso", 16, 2048, 16)):
    S = float(np.log2(np.exp2(np.log2(desired / base) / (L - 1))))
    buf = (ctypes.c_float * L)()
    assert lib().mf_grid_level_scales(ctx.handle, S, base, L, buf) == 0
    out[name] = [float(v) for v in buf]
json.dump(out, open("gpurun_out/ernerf_level_scales.json", "w"), indent=1)
print(out)
