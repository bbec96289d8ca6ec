"""Microbenchmark of the implicit-GEMM conv kernel on the layer shapes that dominate MuseTalk / Wav2Lip (not the bench):
one-op programs, the conv timed alone with CUDA events on its stream (mf_wav2lip_profile).  MF_CONV_DBG modes (skip A
loads / B loads / MMA) locate the bound.  usage: bench_conv.py [modes, e.g. 0,1,2,4]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from mere_fusion_b200._lib import check, lib
from mere_fusion_b200.convnet_pack import ProgramBuilder
from mere_fusion_b200.wav2lip import ConvNet
import ctypes

SHAPES = [  # (name, cin, cout, H, k, B, ups)
    ("vae 128->128 @256", 128, 128, 256, 3, 16, 0),
    ("vae 256->256 @128", 256, 256, 128, 3, 16, 0),
    ("vae 512->512 @64", 512, 512, 64, 3, 16, 0),
    ("vae 512->512 @32", 512, 512, 32, 3, 16, 0),
    ("vae up 256->256 @128->256", 256, 256, 128, 3, 16, 1),
    ("unet 320->320 @32", 320, 320, 32, 3, 16, 0),
    ("unet 640->640 @16", 640, 640, 16, 3, 16, 0),
    ("unet 1280->1280 @8", 1280, 1280, 8, 3, 16, 0),
    ("unet 2560->1280 @8", 2560, 1280, 8, 3, 16, 0),
    ("unet lin 1280->1280 @8 (1x1)", 1280, 1280, 8, 1, 16, 0),
    ("unet lin 320->320 @32 (1x1)", 320, 320, 32, 1, 16, 0),
    ("unet lin 640->640 @16 (1x1)", 640, 640, 16, 1, 16, 0),
    ("unet 1280->1280 @4", 1280, 1280, 4, 3, 16, 0),
    ("w2l 512->512 @1 (1x1)", 512, 512, 1, 1, 16, 0),
    ("w2l 512->512 @3", 512, 512, 3, 3, 16, 0),
    ("w2l 64->64 @96", 64, 64, 96, 3, 16, 0),
    ("w2l 128->128 @48", 128, 128, 48, 3, 16, 0),
    ("w2l256 64->64 @256", 64, 64, 256, 3, 16, 0),
    ("w2l256 64->64 @256 +res", 64, 64, 256, 3, 16, 0),
    ("w2l256 128->32 @256", 128, 32, 256, 3, 16, 0),
    ("w2l256 128->128 @128", 128, 128, 128, 3, 16, 0),
]
modes = [int(m) for m in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["0"])]
forces = os.environ.get("FORCES", "").split(";") if os.environ.get("FORCES") else [None]   # e.g. FORCES="256,1;128,2;64,1"
only = sys.argv[2] if len(sys.argv) > 2 else None
g = torch.Generator().manual_seed(0)
for name, cin, cout, H, k, B, ups in SHAPES:
    if only and only not in name:
        continue
    w = (torch.randn(cout, cin, k, k, generator=g) / np.sqrt(cin * k * k)).numpy()
    row = []
    for mode, force in [(m, f) for m in modes for f in forces]:
        os.environ["MF_CONV_DBG"] = str(mode)
        if force:
            os.environ["MF_CONV_FORCE"] = force
        pb = ProgramBuilder(B)
        a, b = pb.buffer(H, H, cin), pb.buffer(H << ups, H << ups, cout)
        pb.conv(a, 0, b, 0, w, np.zeros(cout, np.float32), padding=k // 2, relu=False, ups=ups,
                res=(a, 0) if "+res" in name else None)
        net = ConvNet(pb.finish(), max_batch=B)
        x = torch.randn(B, H, H, cin, generator=g)
        h = net.ctx.handle
        check(h, lib().mf_wav2lip_profile(h, 0), "profile")
        ms = []
        for it in range(6):
            net.debug_run(a, x, b, (B, H << ups, H << ups, cout))
            v = ctypes.c_float()
            check(h, lib().mf_wav2lip_last_op_ms(h, ctypes.byref(v)), "last_op_ms")
            ms.append(v.value)
        t = float(np.median(ms[2:]))
        fl = pb.flops_per_sample * B
        row.append(f"m{mode}{'[' + force + ']' if force else ''}: {t * 1e3:8.1f} us {fl / t / 1e9:7.1f} TF/s")
        del net
    print(f"{name:34s} " + " | ".join(row), flush=True)
