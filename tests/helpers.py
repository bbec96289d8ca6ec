"""Shared test helpers: fixtures, synthetic inputs of SURVEY.md section 8(d), PSNR."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_ernerf_fixture():
    z = np.load(os.path.join(GOLD, "ernerf_ckpt_infer.npz"))
    sd = {k: z[k] for k in z.files if k != "mean_density_torso"}
    return sd, float(z["mean_density_torso"])


def load_pose_fixture():
    return np.load(os.path.join(GOLD, "ernerf_poses.npz"))


def ernerf_inputs(frame, H, W):
    """config 4 of SURVEY.md 8(d): real poses, intrinsics scaled to the render size,
    auds ~ N(0,1) [8,44,16] with seed 30+frame, eye from au.csv"""
    pf = load_pose_fixture()
    pose = pf["poses"][frame].astype(np.float32)
    fl = float(pf["focal_len"]) * H / (2 * float(pf["cy"]))
    intr = (fl, fl, W / 2.0, H / 2.0)
    auds = np.random.default_rng(30 + frame).standard_normal((8, 44, 16)).astype(np.float32)
    return pose, intr, auds, float(pf["eye"][frame])


def psnr(a, b, peak=1.0):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    mse = np.mean((a - b) ** 2)
    return 99.0 if mse == 0 else 10 * np.log10(peak * peak / mse)
