"""Loader for the REFERENCE's own compiled ErNeRF CUDA extensions (oracle/_ref/, built by
oracle/build_ref.py from the sources under /root/reference).  Test infrastructure only."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ("_raymarching_face", "_grid_encoder", "_sh_encoder", "_freqencoder")


def available():
    return all(os.path.exists(os.path.join(ROOT, "oracle", "_ref", n, n + ".so")) for n in NAMES)


def load():
    import torch  # noqa: F401  (libtorch must be resident before the pybind modules load)
    mods = {}
    for n in NAMES:
        spec = importlib.util.spec_from_file_location(n, os.path.join(ROOT, "oracle", "_ref", n, n + ".so"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        mods[n] = m
    return mods


# ------------------------------------------------------------------------------------------------
# The reference's OWN full render (GPU box): NeRFNetwork + Trainer.test_gui_with_data on the staged,
# unmodified reference Python (oracle/_ref/py, written by oracle/build_ref.py::stage_python) and
# the reference's own compiled kernels (oracle/_ref/*.so).  Ground truth for the frame-level parity
# test and the timed `reference_cuda` baseline of bench.py.  Test / baseline infrastructure only.
# ------------------------------------------------------------------------------------------------
PY = os.path.join(ROOT, "oracle", "_ref", "py")
# the names the reference wrappers import (grid.py:9-12 etc.) -> the names its backend.py files build
ALIASES = {"_raymarching_face": "_raymarching_face", "_gridencoder": "_grid_encoder",
           "_shencoder": "_sh_encoder", "_freqencoder": "_freqencoder"}
STUBS = ("trimesh", "tensorboardX", "mcubes", "torch_ema", "lpips", "imageio", "matplotlib", "matplotlib.pyplot")


def render_available():
    return available() and os.path.exists(os.path.join(PY, "ernerf", "nerf_triplane", "renderer.py"))


def live_opt(**over):
    """the options that reach the reference render in the live app: argparse defaults of app.py:560-690
    plus the ernerf overrides of app.py:355-371"""
    import argparse
    o = argparse.Namespace(
        pose="", au="", torso_imgs="", O=False, data_range=[0, -1], workspace="", seed=0, ckpt="",
        num_rays=4096 * 16, cuda_ray=True, max_steps=16, num_steps=16, upsample_steps=0, update_extra_interval=16,
        max_ray_batch=4096, warmup_step=10000, amb_aud_loss=1, amb_eye_loss=1, unc_loss=1, lambda_amb=1e-4,
        fp16=True, bg_img="white", fbg=False, exp_eye=True, fix_eye=-1, smooth_eye=True, torso_shrink=0.8,
        color_space="srgb", preload=0, bound=1, scale=4, offset=[0, 0, 0], dt_gamma=1 / 256, min_near=0.05,
        density_thresh=10, density_thresh_torso=0.01, patch_size=1, init_lips=False, finetune_lips=False,
        smooth_lips=True, torso=True, head_ckpt="", gui=False, W=450, H=450, radius=3.35, fovy=21.24, max_spp=1,
        att=2, aud="", emb=False, ind_dim=4, ind_num=1, ind_dim_torso=8, amb_dim=2, part=False, part2=False,
        train_camera=False, smooth_path=True, smooth_path_window=7, asr=True, asr_wav="", asr_play=False,
        asr_model="cpierse/wav2vec2-large-xlsr-53-esperanto", asr_save_feats=False, fps=50, l=10, m=8, r=10,
        fullbody=False, test=True, test_train=False, customopt=[], transport="rtc", model="ernerf")
    for k, v in over.items():
        setattr(o, k, v)
    return o


def _import_reference():
    import sys
    import types

    import torch.nn as nn
    mods = load()
    for want, have in ALIASES.items():
        sys.modules[want] = mods[have]
    for name in STUBS:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["lpips"], "LPIPS"):
        class LPIPS(nn.Module):                       # utils.py:663-666 builds one unconditionally; never called at inference
            def __init__(self, net="alex"):
                super().__init__()
        sys.modules["lpips"].LPIPS = LPIPS
    if not hasattr(sys.modules["torch_ema"], "ExponentialMovingAverage"):
        sys.modules["torch_ema"].ExponentialMovingAverage = object
    if "matplotlib" in sys.modules and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if PY not in sys.path:
        sys.path.insert(0, PY)
    from ernerf.nerf_triplane import network, provider, utils
    return network, provider, utils


class ReferenceErnerf:
    """The reference objects app.py:372-392 builds -- NeRFNetwork, Trainer, NeRFDataset_Test -- on the checkpoint
    fixture (tests/golden/ernerf_ckpt_infer.npz: the inference tensors of data/pretrained/ngp_kf.pth) and the pose /
    AU fixture written back into the on-disk formats the reference loader reads (transforms json + au.csv)."""

    def __init__(self, H=450, W=450, device="cuda", n_frames=290, tmpdir=None, **opt_over):
        import json
        import tempfile

        import numpy as np
        import torch
        from helpers import load_ernerf_fixture, load_pose_fixture
        network, provider, utils = _import_reference()
        self.utils = utils
        sd, md = load_ernerf_fixture()
        pf = load_pose_fixture()
        tmp = tmpdir or tempfile.mkdtemp(prefix="mf_ref_")
        # the loader derives H, W from cx, cy (provider.py:109-110); scale the intrinsics to the render size
        # exactly as SURVEY 8(d) config 4 does (fl = 1200 * H / 450)
        fl = float(pf["focal_len"]) * H / (2 * float(pf["cy"]))
        frames = [{"img_id": int(pf["img_id"][i]), "aud_id": int(pf["img_id"][i]),
                   "transform_matrix": pf["raw"][i].astype(np.float64).tolist()} for i in range(len(pf["raw"]))]
        with open(os.path.join(tmp, "transforms.json"), "w") as f:
            json.dump({"focal_len": fl, "cx": W / 2.0, "cy": H / 2.0, "frames": frames}, f)
        au = pf["au"]
        with open(os.path.join(tmp, "au.csv"), "w") as f:
            f.write("frame, AU45_r\n")
            for i, v in enumerate(au):
                f.write(f"{i}, {float(v)!r}\n")
        self.opt = live_opt(pose=os.path.join(tmp, "transforms.json"), au=os.path.join(tmp, "au.csv"), W=W, H=H, **opt_over)
        self.device = torch.device(device)
        model = network.NeRFNetwork(self.opt)
        state = {k: torch.from_numpy(np.asarray(v)).float() if np.asarray(v).dtype in (np.float16, np.float32)
                 else torch.from_numpy(np.asarray(v)) for k, v in sd.items()}
        missing, unexpected = model.load_state_dict(state, strict=False)
        # the fixture leaves out exactly what inference never reads (make_ernerf_fixture.py)
        assert not unexpected, unexpected
        assert all(k.startswith("unc_net") or k in ("density_grid", "step_counter", "aabb_train", "aabb_infer")
                   for k in missing), missing
        model.mean_density_torso = md                 # utils.py:1510-1511
        self.trainer = utils.Trainer("ngp", self.opt, model, device=self.device, workspace=None,
                                     criterion=torch.nn.MSELoss(reduction="none"), fp16=self.opt.fp16, metrics=[],
                                     use_checkpoint="scratch")
        self.model = model
        self.dataset = provider.NeRFDataset_Test(self.opt, device=self.device)
        self.n_frames = n_frames
        self.H, self.W = H, W

    def data(self, frame, auds):
        """what NeRFDataset_Test.collate + NeRFReal.test_step (nerfreal.py:72-79) hand to the trainer"""
        import torch
        d = self.dataset.collate([frame])
        d["auds"] = torch.as_tensor(auds, dtype=torch.float32, device=self.device)
        return d

    def render(self, frame, auds, outW=None, outH=None):
        """fp32 [outH,outW,3] in [0,1]: Trainer.test_gui_with_data (utils.py:1191-1223)"""
        out = self.trainer.test_gui_with_data(self.data(frame, auds), outW or self.W, outH or self.H)
        return out["image"]

    def reset(self):
        self.model.enc_a = None
