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

# ---- batched render (mf_ernerf_render_batch): F sessions of the same avatar model per pass
H = 512
rens = [ErnerfRenderer(blob=ren.blob, cfg=ren.cfg) for _ in range(4)]
ins = [ernerf_inputs(f, H, H) for f in range(16)]
auds = [torch.from_numpy(i[2]).cuda() for i in ins]
for F in (1, 2, 3, 4):
    outs = [torch.empty(H, H, 3, dtype=torch.uint8, device="cuda") for _ in range(F)]

    def go(k):
        fs = [dict(pose=ins[(k * F + j) % 16][0], intrinsics=ins[(k * F + j) % 16][1], H=H, W=H, auds=auds[(k * F + j) % 16],
                   eye=ins[(k * F + j) % 16][3]) for j in range(F)]
        ErnerfRenderer.render_batch(rens[:F], fs, outs=outs)
    for k in range(5):
        go(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 40
    e0.record()
    for k in range(n):
        go(k)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"batch F={F}: {ms:.3f} ms/pass = {ms / F:.3f} ms/frame = {1000 * F / ms:.1f} fps", flush=True)
