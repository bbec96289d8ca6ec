"""per-kernel device times of one ErNeRF frame (CUDA events between the launches are not available through the C ABI, so this
times the frame and prints k_head's own event time): quick A/B of experiment builds (MF_B200_LIB)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import ernerf_inputs, load_ernerf_fixture
from mere_fusion_b200.ernerf import ErnerfRenderer
sd, md = load_ernerf_fixture()
ren = ErnerfRenderer(sd, md)
H = 512
ins = [ernerf_inputs(f, H, H) for f in range(8)]
auds = [torch.from_numpy(i[2]).cuda() for i in ins]
out = torch.empty(H, H, 3, dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ren.profile(True)
for flushed in (False, True):
    tot, head = [], []
    for k in range(40):
        if flushed:
            flush.fill_(k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ren.render(ins[k % 8][0], ins[k % 8][1], H, H, auds[k % 8], ins[k % 8][3], out=out)
        e1.record(); torch.cuda.synchronize()
        if k >= 8:
            tot.append(e0.elapsed_time(e1)); head.append(ren.last_head_ms()[0])
    print(f"flush={flushed}: frame {np.mean(tot)*1e3:.1f} us, k_head {np.mean(head)*1e3:.1f} us, rest {np.mean(tot)*1e3 - np.mean(head)*1e3:.1f} us", flush=True)
