"""diagnostic (GPU box): which fp32 evaluation order reproduces torch's get_rays (utils.py:255-341) bit for bit?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch

def f32(x): return x.to(torch.float32)
def fma(a, b, c): return f32(a.double() * b.double() + c.double())
def mul(a, b): return f32(a.double() * b.double())
def add(a, b): return f32(a.double() + b.double())

for (H, W, fl) in ((450, 450, 1200.0), (512, 512, 1200 * 512 / 450), (360, 640, 900.0), (333, 517, 1234.5)):
    cx, cy = W / 2.0, H / 2.0
    col = torch.arange(W, device="cuda", dtype=torch.float32); row = torch.arange(H, device="cuda", dtype=torch.float32)
    i = (col[None, :].expand(H, W).reshape(-1) + 0.5)
    j = (row[:, None].expand(H, W).reshape(-1) + 0.5)
    xs = (i - cx) / fl; ys = (j - cy) / fl; zs = torch.ones_like(xs)
    dirs_t = torch.stack((xs, ys, zs), -1)[None]
    tn = torch.norm(dirs_t, dim=-1, keepdim=True).view(-1)
    one = torch.ones_like(xs)
    x2, y2 = mul(xs, xs), mul(ys, ys)
    c = {"(x2+1)+y2": add(add(x2, one), y2), "(x2+y2)+1": add(add(x2, y2), one), "x2+(y2+1)": add(x2, add(y2, one)),
         "fma(y,y,fma(x,x,1))": fma(ys, ys, fma(xs, xs, one)), "fma(x,x,fma(y,y,1))": fma(xs, xs, fma(ys, ys, one)),
         "fma(y,y,x2+1)": fma(ys, ys, add(x2, one)), "fma(x,x,y2+1)": fma(xs, xs, add(y2, one)),
         "x2+fma(y,y,1)": add(x2, fma(ys, ys, one)), "y2+fma(x,x,1)": add(y2, fma(xs, xs, one)),
         "fma(y,y,x2)+1": add(fma(ys, ys, x2), one), "fma(x,x,y2)+1": add(fma(xs, xs, y2), one),
         "double": f32(xs.double() ** 2 + ys.double() ** 2 + 1.0)}
    for k, v in c.items():
        print(H, W, k, "mismatch:", int((torch.sqrt(v) != tn).sum()), " (sqrt in double:", int((f32(torch.sqrt(v.double())) != tn).sum()), ")")
    dsq = f32(torch.sqrt(xs.double() ** 2 + ys.double() ** 2 + 1.0))
    print(H, W, "all-double then round:", int((dsq != tn).sum()))
