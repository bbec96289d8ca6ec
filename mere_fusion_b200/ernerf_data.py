"""Host-side pose / eye provider for ErNeRF sessions: the inference half of
ernerf/nerf_triplane/provider.py (NeRFDataset_Test, :84-341).

Everything here is load-time or O(1) per frame; rays and bg_coords are generated inside the
CUDA kernels, so unlike the reference no per-frame [H*W, 3] tensors are built on the host.
"""
import json

import numpy as np


def nerf_matrix_to_ngp(pose, scale=0.33, offset=(0, 0, 0)):
    """provider.py:19-26"""
    return np.array([
        [pose[1, 0], -pose[1, 1], -pose[1, 2], pose[1, 3] * scale + offset[0]],
        [pose[2, 0], -pose[2, 1], -pose[2, 2], pose[2, 3] * scale + offset[1]],
        [pose[0, 0], -pose[0, 1], -pose[0, 2], pose[0, 3] * scale + offset[2]],
        [0, 0, 0, 1],
    ], dtype=np.float32)


def smooth_camera_path(poses, kernel_size=5):
    """provider.py:29-45: window mean of translations, chordal mean of rotations"""
    from scipy.spatial.transform import Rotation
    N = poses.shape[0]
    K = kernel_size // 2
    trans = poses[:, :3, 3].copy()
    rots = poses[:, :3, :3].copy()
    for i in range(N):
        start = max(0, i - K)
        end = min(N, i + K + 1)
        poses[i, :3, 3] = trans[start:end].mean(0)
        poses[i, :3, :3] = Rotation.from_matrix(rots[start:end]).mean().as_matrix()
    return poses


def smooth_eye_area(area):
    """provider.py:243-250 ("naive 5 window average" -- the window is 3)"""
    ori = area.copy()
    out = area.copy()
    for i in range(ori.shape[0]):
        out[i] = ori[max(0, i - 1):min(ori.shape[0], i + 2)].mean()
    return out


def mirror_index(size, index):
    """provider.py:276-283 / basereal.py:133-139: ping-pong replay"""
    turn = index // size
    res = index % size
    return res if turn % 2 == 0 else size - res - 1


class ErnerfPoseProvider:
    """poses [N,4,4] fp32, eye_area [N] fp32, intrinsics (fx, fy, cx, cy), H, W."""

    def __init__(self, transform, au_blink=None, scale=4.0, offset=(0, 0, 0), smooth_path=True,
                 smooth_path_window=7, exp_eye=True, smooth_eye=True, data_range=(0, -1), downscale=1, bg_img="white",
                 torso_imgs=""):
        if isinstance(transform, str):
            with open(transform, "r") as f:
                transform = json.load(f)
        self.H = int(transform["cy"]) * 2 // downscale
        self.W = int(transform["cx"]) * 2 // downscale
        frames = transform["frames"]
        end = len(frames) if data_range[1] == -1 else data_range[1]
        frames = frames[data_range[0]:end]
        poses = [nerf_matrix_to_ngp(np.array(f["transform_matrix"], dtype=np.float32), scale, offset) for f in frames]
        self.poses = np.stack(poses, 0)
        if smooth_path:
            self.poses = smooth_camera_path(self.poses, smooth_path_window)
        self.poses = self.poses.astype(np.float32)
        self.eye_area = None
        if exp_eye and au_blink is not None:
            area = np.array([np.clip(au_blink[f["img_id"]], 0, 2) / 2 for f in frames], dtype=np.float32)
            self.eye_area = smooth_eye_area(area) if smooth_eye else area
        fl = transform["focal_len"]
        self.intrinsics = np.array([fl, fl, transform["cx"] / downscale, transform["cy"] / downscale])
        self.index = 0
        if torso_imgs != "":
            raise NotImplementedError("opt.torso_imgs (per-frame torso composites as background, provider.py:316-328) is not supported by "
                                      "the fused renderer: run with the torso model (torso_imgs='')")
        self.bg_img = load_bg_img(bg_img, self.H, self.W)

    def __len__(self):
        return self.poses.shape[0]

    def get(self, index):
        """collate (provider.py:285-341) minus the ray tensors"""
        i = mirror_index(len(self), index)
        eye = float(self.eye_area[i]) if self.eye_area is not None else None
        return i, self.poses[i], eye

    def __iter__(self):
        return self

    def __next__(self):
        out = self.get(self.index)
        self.index += 1
        return out


def load_bg_img(bg_img, H, W):
    """provider.py:203-214: 'white' -> None (the renderer's default), 'black' -> zeros, else an image file read as RGB in
    [0,1], resized with INTER_AREA when its size differs.  Returns fp32 [H,W,3] or None."""
    if bg_img is None or (isinstance(bg_img, str) and bg_img == "white"):
        return None
    if isinstance(bg_img, str) and bg_img == "black":
        return np.zeros((H, W, 3), np.float32)
    if isinstance(bg_img, str):
        import cv2
        img = cv2.imread(bg_img, cv2.IMREAD_UNCHANGED)
        if img is None:
            raise FileNotFoundError(bg_img)
        if img.shape[0] != H or img.shape[1] != W:
            img = cv2.resize(img, (W, H), interpolation=cv2.INTER_AREA)
        img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
        return img.astype(np.float32) / 255
    return np.asarray(bg_img, np.float32).reshape(H, W, 3)


def load_au_blink(path):
    """provider.py:145-147: column ' AU45_r' of the OpenFace csv"""
    import pandas as pd
    return pd.read_csv(path)[" AU45_r"].values
