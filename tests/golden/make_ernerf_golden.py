"""Golden vectors for the NETWORK GLUE of the ErNeRF oracle (oracle/ernerf_oracle.py), produced by the REFERENCE's own
nn.Modules (AudioNet, AudioAttNet, MLP, NeRFNetwork.encode_audio / .density, the colour head of .forward) imported from
/root/reference and loaded with the real checkpoint data/pretrained/ngp_kf.pth.  Run in the build container only
(/root/reference does not exist on the GPU box):

    python tests/golden/make_ernerf_golden.py     ->  tests/golden/ernerf_glue_golden.npz

Everything that can run without the CUDA-only encoders runs here on CPU: the encoders' outputs (enc_x, the SH vector) are
seeded inputs.  Two variants per output: `*_f32` = the modules in plain fp32, `*_ac` = under torch.autocast("cpu", fp16),
the CPU counterpart of the `torch.cuda.amp.autocast` the live path runs under (utils.py:1200).  The GPU-side pin of the
same glue (rounding points included) is tests/test_ernerf_reference_render_gpu.py.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import ref_ernerf  # noqa: E402

REF = "/root/reference"
ref_ernerf.PY = REF                      # import the reference where it lies
network, provider, utils = ref_ernerf._import_reference()
opt = ref_ernerf.live_opt(ind_num=10000)
model = network.NeRFNetwork(opt)
ck = torch.load(f"{REF}/data/pretrained/ngp_kf.pth", map_location="cpu", weights_only=False)
print(model.load_state_dict(ck["model"], strict=True))
model.eval()

rng = np.random.default_rng(77)
M = 96
out = {}
auds = rng.standard_normal((3, 8, 44, 16)).astype(np.float32)            # three consecutive attention windows
emb_std = float(ck["model"]["encoder_xy.embeddings"].std())
enc_x = (rng.standard_normal((M, 36)) * emb_std).astype(np.float32)
d = rng.standard_normal((M, 3)).astype(np.float32)
d /= np.linalg.norm(d, axis=1, keepdims=True)
eye = np.float32(0.37)
out.update(auds=auds, enc_x=enc_x, dirs=d, eye=eye)


def run(tag, ctx):
    with torch.no_grad(), ctx:
        enc = [model.encode_audio(torch.from_numpy(a)) for a in auds]     # network.py:222-237
        out[f"enc_a_{tag}"] = np.stack([e.float().numpy() for e in enc])  # [3,1,32]
        enc_a = enc[0].float()
        r = model.density(None, enc_a, torch.tensor([[eye]]), torch.from_numpy(enc_x))   # network.py:280-308
        out[f"sigma_{tag}"] = r["sigma"].float().numpy()
        out[f"geo_{tag}"] = r["geo_feat"].float().numpy()
        out[f"amb_aud_{tag}"] = r["ambient_aud"].float().numpy()
        out[f"eye_att_{tag}"] = r["ambient_eye"].float().numpy()
        # colour head exactly as network.py:264-272 writes it, on a seeded stand-in for the SH vector
        enc_d = torch.from_numpy(sh)
        h = torch.cat([enc_d, r["geo_feat"], model.individual_codes[0][None].repeat(M, 1)], dim=-1)
        hc = model.color_net(h)
        out[f"color_{tag}"] = (torch.sigmoid(hc) * (1 + 2 * 0.001) - 0.001).float().numpy()
        # torso MLPs on seeded inputs (the freq / tiled-grid encoders are CUDA only)
        out[f"deform_{tag}"] = model.torso_deform_net(torch.from_numpy(th)).float().numpy()
        out[f"torso_{tag}"] = model.torso_net(torch.from_numpy(tt)).float().numpy()


sh = rng.uniform(-1, 1, (M, 16)).astype(np.float32)
th = rng.uniform(-1, 1, (M, 84)).astype(np.float32)
tt = rng.uniform(-1, 1, (M, 116)).astype(np.float32)
out.update(sh=sh, torso_deform_in=th, torso_in=tt)
import contextlib
run("f32", contextlib.nullcontext())
run("ac", torch.autocast("cpu", dtype=torch.float16))
np.savez_compressed(os.path.join(HERE, "ernerf_glue_golden.npz"), **out)
for k, v in out.items():
    print(k, np.asarray(v).shape, np.asarray(v).dtype)
