"""quick device-side timing of mf_ernerf_render (not the bench)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import ernerf_inputs, load_ernerf_fixture
from mere_fusion_b200.ernerf import ErnerfRenderer
sd, md = load_ernerf_fixture()
ren = ErnerfRenderer(sd, md)
for H in (450, 512):
    ins = [ernerf_inputs(f, H, H) for f in range(8)]
    auds = [torch.from_numpy(i[2]).cuda() for i in ins]
    out = torch.empty(H, H, 3, dtype=torch.uint8, device="cuda")
    for k in range(5):
        ren.render(ins[k % 8][0], ins[k % 8][1], H, H, auds[k % 8], ins[k % 8][3], out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 40
    e0.record()
    for k in range(n):
        ren.render(ins[k % 8][0], ins[k % 8][1], H, H, auds[k % 8], ins[k % 8][3], out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    o, dbg = ren.render(ins[0][0], ins[0][1], H, H, auds[0], ins[0][3], debug=True)
    torch.cuda.synchronize()
    print(f"H={H}: {ms:.3f} ms/frame = {1000/ms:.1f} fps; rounds:", dbg["round_info"].cpu().numpy()[:8].tolist(), flush=True)
    import cv2
    cv2.imwrite(f"gpurun_out/ernerf_{H}.png", o.cpu().numpy()[..., ::-1])
