import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, cv2, torch
from mere_fusion_b200._lib import Context, lib
from oracle.paste_oracle import resize_linear_u8, paste_cv2
ctx = Context(0)
rng = np.random.default_rng(7)
H, W, S, n = 300, 420, 96, 5
frames = rng.integers(0, 256, (n, H, W, 3), dtype=np.uint8)
boxes = [(10, 202, 20, 212), (0, 96, 0, 96), (50, 98, 60, 108), (3, 300, 1, 420), (100, 170, 200, 250), (120, 217, 33, 128), (7, 200, 300, 420)]
B = len(boxes)
faces = rng.integers(0, 256, (B, S, S, 3), dtype=np.uint8)
rows = np.array([(i % n,) + b for i, b in enumerate(boxes)], np.int32)
print(rows.flags['C_CONTIGUOUS'], rows.dtype, rows.shape)
d_frames, d_faces = torch.from_numpy(frames).cuda(), torch.from_numpy(faces).cuda()
out = torch.empty(B, H, W, 3, dtype=torch.uint8, device="cuda")
rc = lib().mf_paste_resize_u8(ctx.handle, ctypes.c_void_p(d_frames.data_ptr()), n, H, W, ctypes.c_void_p(d_faces.data_ptr()),
                              S, B, rows.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), ctypes.c_void_p(out.data_ptr()), None)
torch.cuda.synchronize()
got = out.cpu().numpy()
for i, b in enumerate(boxes):
    exp = paste_cv2(frames[i % n], faces[i], b)
    d = got[i] != exp
    y1, y2, x1, x2 = b
    inside = np.zeros((H, W), bool); inside[y1:y2, x1:x2] = True
    dm = d.any(-1)
    print(i, b, "rc", rc, "mism total", int(dm.sum()), "inside", int((dm & inside).sum()), "outside", int((dm & ~inside).sum()))
    if dm.any():
        ys, xs = np.nonzero(dm)
        print("    ", list(zip(ys[:6], xs[:6])), got[i][ys[0], xs[0]], exp[ys[0], xs[0]])
        o = resize_linear_u8(faces[i], x2 - x1, y2 - y1)
        c = cv2.resize(faces[i], (x2 - x1, y2 - y1))
        print("     oracle-cv2", int((o != c).sum()), "cuda-oracle", int((got[i][y1:y2, x1:x2] != o).sum()))
