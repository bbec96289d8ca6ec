"""device-side timing of mf_musetalk_forward, full MuseTalk-v1 configuration, random weights (not the bench)"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import musetalk_oracle as M
from mere_fusion_b200.musetalk import MuseTalkEngine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
small = len(sys.argv) > 2 and sys.argv[2] == "small"
u, v = M.small_cfgs() if small else (M.UNET_CFG, M.VAE_CFG)
t0 = time.time()
usd = M.seeded_state(M.unet_param_shapes(u), 5); vsd = M.seeded_state(M.vae_decoder_param_shapes(v), 6)
t1 = time.time()
eng = MuseTalkEngine(usd, vsd, u, v, max_batch=B)
print(f"weights {t1-t0:.1f}s, pack+load {time.time()-t1:.1f}s, GFLOP/frame unet {eng.unet_flops/1e9:.1f} vae {eng.vae_flops/1e9:.1f}, mem {torch.cuda.memory_allocated()/1e9:.1f} GB torch", flush=True)
rng = np.random.default_rng(10)
lat = torch.from_numpy((rng.standard_normal((B, 8, 32, 32)) * 0.9).astype(np.float16)).cuda()
wh = torch.from_numpy(rng.standard_normal((B, 50, 384)).astype(np.float16)).cuda()
out = torch.empty(B, 256, 256, 3, dtype=torch.uint8, device="cuda")
for _ in range(3):
    eng.forward(lat, wh, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = int(os.environ.get('N_ITERS', '10'))
e0.record()
for _ in range(n):
    eng.forward(lat, wh, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"B={B}: {ms:.2f} ms/batch = {B*1000/ms:.1f} frames/s; {eng.flops_per_frame*B/ms/1e9:.1f} TFLOP/s algorithmic; launches {eng.last_launches}; out mean {out.float().mean().item():.1f}", flush=True)
free, tot = torch.cuda.mem_get_info(); print(f"device memory used {(tot-free)/1e9:.1f} GB")
