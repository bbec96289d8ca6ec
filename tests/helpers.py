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


# ------------------------------------------------------------------------------------------------
# Wav2Lip: seeded weights (the reference ships no checkpoint: ./models/wav2lip.pth is external)
# ------------------------------------------------------------------------------------------------
def wav2lip_param_shapes():
    """name -> shape for every tensor of the reference Wav2Lip generator state_dict
    (wav2lip/models/wav2lip.py:12-85); verified against the reference module by
    tests/golden/make_wav2lip_golden.py (strict load)."""
    shapes = {}

    def block(prefix, cin, cout, k, transpose=False):
        kh, kw = (k, k) if isinstance(k, int) else k
        shapes[prefix + ".conv_block.0.weight"] = (cin, cout, kh, kw) if transpose else (cout, cin, kh, kw)
        shapes[prefix + ".conv_block.0.bias"] = (cout,)
        for n in ("weight", "bias", "running_mean", "running_var"):
            shapes[prefix + ".conv_block.1." + n] = (cout,)
        shapes[prefix + ".conv_block.1.num_batches_tracked"] = ()

    enc = [[(6, 16, 7)], [(16, 32, 3), (32, 32, 3), (32, 32, 3)], [(32, 64, 3)] + [(64, 64, 3)] * 3,
           [(64, 128, 3)] + [(128, 128, 3)] * 2, [(128, 256, 3)] + [(256, 256, 3)] * 2, [(256, 512, 3), (512, 512, 3)],
           [(512, 512, 3), (512, 512, 1)]]
    for i, blk in enumerate(enc):
        for j, (ci, co, k) in enumerate(blk):
            block(f"face_encoder_blocks.{i}.{j}", ci, co, k)
    aud = [(1, 32, 3), (32, 32, 3), (32, 32, 3), (32, 64, 3), (64, 64, 3), (64, 64, 3), (64, 128, 3), (128, 128, 3),
           (128, 128, 3), (128, 256, 3), (256, 256, 3), (256, 512, 3), (512, 512, 1)]
    for j, (ci, co, k) in enumerate(aud):
        block(f"audio_encoder.{j}", ci, co, k)
    dec = [[(512, 512, 1, False)], [(1024, 512, 3, True), (512, 512, 3, False)],
           [(1024, 512, 3, True)] + [(512, 512, 3, False)] * 2, [(768, 384, 3, True)] + [(384, 384, 3, False)] * 2,
           [(512, 256, 3, True)] + [(256, 256, 3, False)] * 2, [(320, 128, 3, True)] + [(128, 128, 3, False)] * 2,
           [(160, 64, 3, True)] + [(64, 64, 3, False)] * 2]
    for i, blk in enumerate(dec):
        for j, (ci, co, k, t) in enumerate(blk):
            block(f"face_decoder_blocks.{i}.{j}", ci, co, k, transpose=t)
    block("output_block.0", 80, 32, 3)
    shapes["output_block.1.weight"] = (3, 32, 1, 1)
    shapes["output_block.1.bias"] = (3,)
    return shapes


def seeded_wav2lip_state(seed=2, face_hw=96):
    """deterministic (numpy RNG, key order) weights with NON-trivial BatchNorm running statistics --
    eval-mode BN is part of the arithmetic.  face_hw=256: the extended generator (not in the reference)"""
    import torch
    rng = np.random.default_rng(seed)
    sd = {}
    if face_hw == 96:
        shapes = wav2lip_param_shapes()
    else:
        from mere_fusion_b200.wav2lip_pack import wav2lip_param_shapes as pack_shapes
        shapes = pack_shapes(face_hw)
    for name, shp in shapes.items():
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.tensor(100, dtype=torch.long)
        elif name.endswith("conv_block.0.weight") or name == "output_block.1.weight":
            if "face_decoder" in name and len(shp) == 4 and ".0.conv_block" in name and shp[0] > shp[1] and shp[2] in (3, 4) \
                    and not name.startswith("face_decoder_blocks.0"):
                fan_in = shp[0] * shp[2] * shp[3] / 4.0          # ConvTranspose2d: ~1/4 of the taps hit a pixel
            else:
                fan_in = shp[1] * shp[2] * shp[3]
            if "face_decoder_blocks.1.0" in name:
                fan_in = shp[0]
            v = rng.standard_normal(shp) * np.sqrt(1.0 / fan_in)
            sd[name] = torch.from_numpy(v.astype(np.float32))
        elif name.endswith("running_var"):
            sd[name] = torch.from_numpy(rng.uniform(0.6, 1.6, shp).astype(np.float32))
        elif name.endswith("running_mean"):
            sd[name] = torch.from_numpy((rng.standard_normal(shp) * 0.2).astype(np.float32))
        elif name.endswith("conv_block.1.weight"):
            sd[name] = torch.from_numpy(rng.uniform(0.7, 1.3, shp).astype(np.float32))
        else:   # conv biases, BN beta
            sd[name] = torch.from_numpy((rng.standard_normal(shp) * 0.1).astype(np.float32))
    return sd


def wav2lip_inputs(B, mel_seed=3, face_seed=4, S=96):
    """config 2 of SURVEY.md 8(d): mel ~ N(0,1) clipped to +-4 [B,1,80,16]; faces u8 uniform [B,S,S,3]"""
    mel = np.clip(np.random.default_rng(mel_seed).standard_normal((B, 1, 80, 16)), -4, 4).astype(np.float32)
    faces = np.random.default_rng(face_seed).integers(0, 256, (B, S, S, 3), dtype=np.uint8)
    return mel, faces


# ------------------------------------------------------------------------------------------------
# Whisper encoder (tiny dims): seeded weights (./models/whisper/tiny.pt is external) and the synthetic audio of
# SURVEY.md 8(d) config 1; golden features come from the vendored reference (tests/golden/make_whisper_golden.py)
# ------------------------------------------------------------------------------------------------
WHISPER_TINY = dict(n_mels=80, n_audio_ctx=1500, n_audio_state=384, n_audio_head=6, n_audio_layer=4)


def whisper_encoder_shapes(d=WHISPER_TINY):
    D, M = d["n_audio_state"], d["n_mels"]
    s = {"conv1.weight": (D, M, 3), "conv1.bias": (D,), "conv2.weight": (D, D, 3), "conv2.bias": (D,),
         "positional_embedding": (d["n_audio_ctx"], D)}
    for i in range(d["n_audio_layer"]):
        p = f"blocks.{i}."
        for n in ("query", "key", "value", "out"):
            s[p + f"attn.{n}.weight"] = (D, D)
            if n != "key":
                s[p + f"attn.{n}.bias"] = (D,)
        for n in ("attn_ln", "mlp_ln"):
            s[p + n + ".weight"] = (D,)
            s[p + n + ".bias"] = (D,)
        s[p + "mlp.0.weight"], s[p + "mlp.0.bias"] = (4 * D, D), (4 * D,)
        s[p + "mlp.2.weight"], s[p + "mlp.2.bias"] = (D, 4 * D), (D,)
    s["ln_post.weight"], s["ln_post.bias"] = (D,), (D,)
    return s


def seeded_whisper_state(seed=7, d=WHISPER_TINY):
    """numpy arrays keyed like AudioEncoder.state_dict(); attention weights are scaled up a little so that the softmax
    is not uniform (a uniform softmax would hide indexing mistakes)"""
    rng = np.random.default_rng(seed)
    sd = {}
    for name, shp in whisper_encoder_shapes(d).items():
        if name == "positional_embedding":
            inc = np.log(10000) / (shp[1] // 2 - 1)
            inv = np.exp(-inc * np.arange(shp[1] // 2))
            t = np.arange(shp[0])[:, None] * inv[None, :]
            sd[name] = np.concatenate([np.sin(t), np.cos(t)], axis=1).astype(np.float32)
        elif name.endswith("ln.weight") or name == "ln_post.weight":
            sd[name] = rng.uniform(0.7, 1.3, shp).astype(np.float32)
        elif name.endswith(".bias"):
            sd[name] = (rng.standard_normal(shp) * 0.1).astype(np.float32)
        else:
            fan_in = int(np.prod(shp[1:]))
            gain = 2.5 if (".query." in name or ".key." in name) else 1.0
            sd[name] = (rng.standard_normal(shp) * gain / np.sqrt(fan_in)).astype(np.float32)
    return sd


def synthetic_speech(n, seed=0):
    """SURVEY.md 8(d) config 1 waveform: 0.3 sin(2 pi 220 t) (0.5 + 0.5 sin(2 pi 3 t)) + 0.01 N(0,1), 16 kHz"""
    t = np.arange(n) / 16000.0
    x = 0.3 * np.sin(2 * np.pi * 220 * t) * (0.5 + 0.5 * np.sin(2 * np.pi * 3 * t)) + 0.01 * np.random.default_rng(seed).standard_normal(n)
    return x.astype(np.float32)


# ------------------------------------------------------------------------------------------------
# wav2vec2 CTC (HF Wav2Vec2ForCTC, layer-norm feature encoder + stable-layer-norm transformer): seeded weights for the
# XLSR-53-large shape NerfASR loads (nerfasr.py:44-45; the checkpoint is external) and for a small config of identical structure
# ------------------------------------------------------------------------------------------------
W2V_XLSR53 = dict(vocab=44, hidden=1024, layers=24, heads=16, inter=4096, conv_dim=(512,) * 7, conv_stride=(5, 2, 2, 2, 2, 2, 2),
                  conv_kernel=(10, 3, 3, 3, 3, 2, 2), pos_k=128, pos_groups=16, eps=1e-5)
W2V_SMALL = dict(W2V_XLSR53, hidden=256, layers=3, heads=4, inter=512, conv_dim=(64,) * 7, pos_groups=4)


def w2v_param_shapes(c):
    D, I, V = c["hidden"], c["inter"], c["vocab"]
    s = {"wav2vec2.masked_spec_embed": (D,)}
    cin = 1
    for i, (co, k) in enumerate(zip(c["conv_dim"], c["conv_kernel"])):
        p = f"wav2vec2.feature_extractor.conv_layers.{i}."
        s[p + "conv.weight"], s[p + "conv.bias"] = (co, cin, k), (co,)
        s[p + "layer_norm.weight"], s[p + "layer_norm.bias"] = (co,), (co,)
        cin = co
    s["wav2vec2.feature_projection.layer_norm.weight"], s["wav2vec2.feature_projection.layer_norm.bias"] = (cin,), (cin,)
    s["wav2vec2.feature_projection.projection.weight"], s["wav2vec2.feature_projection.projection.bias"] = (D, cin), (D,)
    pc = "wav2vec2.encoder.pos_conv_embed.conv."
    s[pc + "bias"] = (D,)
    s[pc + "parametrizations.weight.original0"] = (1, 1, c["pos_k"])
    s[pc + "parametrizations.weight.original1"] = (D, D // c["pos_groups"], c["pos_k"])
    s["wav2vec2.encoder.layer_norm.weight"], s["wav2vec2.encoder.layer_norm.bias"] = (D,), (D,)
    for i in range(c["layers"]):
        p = f"wav2vec2.encoder.layers.{i}."
        for n in ("k", "v", "q", "out"):
            s[p + f"attention.{n}_proj.weight"], s[p + f"attention.{n}_proj.bias"] = (D, D), (D,)
        for n in ("layer_norm", "final_layer_norm"):
            s[p + n + ".weight"], s[p + n + ".bias"] = (D,), (D,)
        s[p + "feed_forward.intermediate_dense.weight"], s[p + "feed_forward.intermediate_dense.bias"] = (I, D), (I,)
        s[p + "feed_forward.output_dense.weight"], s[p + "feed_forward.output_dense.bias"] = (D, I), (D,)
    s["lm_head.weight"], s["lm_head.bias"] = (V, D), (V,)
    return s


def seeded_w2v_state(seed, c):
    rng = np.random.default_rng(seed)
    sd = {}
    for name, shp in w2v_param_shapes(c).items():
        if name.endswith("layer_norm.weight"):
            sd[name] = rng.uniform(0.7, 1.3, shp).astype(np.float32)
        elif name.endswith(".bias") or name.endswith("masked_spec_embed"):
            sd[name] = (rng.standard_normal(shp, dtype=np.float32) * 0.1)
        elif name.endswith("original0"):
            sd[name] = rng.uniform(0.5, 1.5, shp).astype(np.float32)
        else:
            fan_in = int(np.prod(shp[1:]))
            gain = 2.0 if (".q_proj." in name or ".k_proj." in name) else 1.0
            sd[name] = rng.standard_normal(shp, dtype=np.float32) * np.float32(gain / np.sqrt(fan_in))
    return sd
