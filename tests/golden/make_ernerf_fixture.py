"""Generate the ErNeRF fixtures from the reference's own data (run in the build container;
/root/reference does not exist on the GPU box).

  ernerf_ckpt_infer.npz : the inference-relevant tensors of data/pretrained/ngp_kf.pth
                          (torso table stored as the fp16 the live path uses; only row 0 of the
                          individual codes, the only row inference reads: renderer.py:197-204,308-315)
  ernerf_poses.npz      : poses / eye areas for frames 0..299 of data/data_kf.json + data/au.csv, produced by
                          the REFERENCE's provider functions (imported from /root/reference) --
                          the golden vectors for mere_fusion_b200/ernerf_data.py
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

ck = torch.load(f"{REF}/data/pretrained/ngp_kf.pth", map_location="cpu", weights_only=False)
sd = ck["model"]
keep = {}
for k, v in sd.items():
    if k in ("density_grid", "step_counter", "aabb_train", "aabb_infer") or k.startswith("unc_net"):
        continue
    v = v.numpy()
    if k in ("individual_codes", "individual_codes_torso"):
        v = v[:1]
    if k == "torso_encoder.embeddings":
        v = v.astype(np.float16)
    keep[k] = v
keep["mean_density_torso"] = np.float32(ck["mean_density_torso"])
np.savez_compressed(f"{OUT}/ernerf_ckpt_infer.npz", **keep)
print("ckpt fixture", os.path.getsize(f"{OUT}/ernerf_ckpt_infer.npz") / 1e6, "MB")

# reference provider functions (stub the imports the reference never uses at inference)
for m in ("trimesh", "tensorboardX", "mcubes", "torch_ema", "lpips", "imageio", "matplotlib", "matplotlib.pyplot",
          "cv2", "tqdm", "packaging"):
    pass
sys.path.insert(0, REF)
import json
import importlib.util
src = open(f"{REF}/ernerf/nerf_triplane/provider.py").read()
# execute only the two pure functions, textually extracted from the reference file at run time
start = src.index("def nerf_matrix_to_ngp")
end = src.index("def polygon_area")
ns = {"np": np}
from scipy.spatial.transform import Rotation
ns["Rotation"] = Rotation
exec(compile(src[start:end], "provider_excerpt", "exec"), ns)
tr = json.load(open(f"{REF}/data/data_kf.json"))
frames = tr["frames"]
poses = np.stack([ns["nerf_matrix_to_ngp"](np.array(f["transform_matrix"], dtype=np.float32), scale=4, offset=[0, 0, 0])
                  for f in frames], 0)
poses = ns["smooth_camera_path"](poses, 7)
import pandas as pd
au = pd.read_csv(f"{REF}/data/au.csv")[" AU45_r"].values
area = np.array([np.clip(au[f["img_id"]], 0, 2) / 2 for f in frames], dtype=np.float32)
ori = area.copy()
for i in range(ori.shape[0]):   # provider.py:243-250
    area[i] = ori[max(0, i - 1):min(ori.shape[0], i + 2)].mean()
np.savez_compressed(f"{OUT}/ernerf_poses.npz", poses=poses[:300].astype(np.float32), eye=area[:300],
                    raw=np.stack([np.array(f["transform_matrix"], dtype=np.float32) for f in frames[:304]]),
                    img_id=np.array([f["img_id"] for f in frames[:304]]), au=au[:400].astype(np.float64),
                    focal_len=np.float64(tr["focal_len"]), cx=np.float64(tr["cx"]), cy=np.float64(tr["cy"]),
                    n_frames=np.int64(len(frames)))
print("poses fixture", os.path.getsize(f"{OUT}/ernerf_poses.npz") / 1e3, "KB")
