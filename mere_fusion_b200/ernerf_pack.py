"""Pack a reference ErNeRF checkpoint (`ckpt['model']`) into the single
device-resident blob libmf_b200.so consumes (csrc/ernerf_layout.h).

Host-side, run once per avatar.  The blob is plain bytes so rank 0 can build it and
`torch.distributed.broadcast` it to the other GPUs.  The loader is strict: every
tensor shape is checked against the architecture of ernerf/nerf_triplane/network.py:93-163.
"""
import ctypes
import struct

import numpy as np

from ._lib import MfErnerfCfg, lib, MF_ERNERF_HEAD_LEVELS, MF_ERNERF_TORSO_LEVELS

MAGIC = 0x3242464D
(ID_HEAD_PLANES, ID_BITFIELD, ID_TORSO_TABLE, ID_TORSO_DENSITY, ID_HEAD_MLP, ID_TORSO_MLP, ID_AUDIO, ID_MISC,
 ID_TORSO_CONST) = range(1, 10)

_LAYOUT_NAMES = ["H_AUD1", "H_AUD2", "H_EYE1", "H_SIG1", "H_SIG2", "H_SIG3", "H_COL1", "H_COL2", "H_EYE2",
                 "H_HALFS", "H_COLBIAS_BYTES", "H_BYTES", "T_DEF1", "T_DEF2", "T_DEF3", "T_TOR1", "T_TOR2",
                 "T_TOR3", "T_HALFS", "T_BYTES"]


def blob_layout():
    buf = (ctypes.c_int32 * 32)()
    n = lib().mf_ernerf_blob_layout(buf, 32)
    assert n == len(_LAYOUT_NAMES)
    return dict(zip(_LAYOUT_NAMES, list(buf)[:n]))


def _np(v):
    return v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)


def _put(img, off, W, stride):
    """store W[n][k] (fp32) as fp16 rows of `stride` halfs at half-offset `off`"""
    n, k = W.shape
    view = img[off:off + n * stride].reshape(n, stride)
    view[:, :k] = W.astype(np.float16)


def build_blob(entries, kind):
    """entries: {id: bytes-like}; returns np.uint8 blob with 256-byte aligned payloads"""
    ids = sorted(entries)
    hdr = 16 + 24 * len(ids)
    off = (hdr + 255) // 256 * 256
    table = []
    for i in ids:
        b = entries[i]
        table.append((i, off, len(b)))
        off = (off + len(b) + 255) // 256 * 256
    blob = np.zeros(off, np.uint8)
    head = struct.pack("<IIII", MAGIC, kind, 1, len(ids))
    for i, o, n in table:
        head += struct.pack("<IIQQ", i, 0, o, n)
    blob[:len(head)] = np.frombuffer(head, np.uint8)
    for i, o, n in table:
        blob[o:o + n] = np.frombuffer(entries[i], np.uint8)
    return blob


def pack_ernerf(sd, mean_density_torso, opt=None):
    """sd: state_dict (torch tensors or numpy).  opt: overrides of the live options that reach
    run_cuda (app.py:355-371, :598-613).  Returns (blob uint8 ndarray, MfErnerfCfg)."""
    sd = {k: _np(v) for k, v in sd.items()}
    o = dict(bound=1.0, min_near=0.05, dt_gamma=1 / 256, max_steps=16, T_thresh=1e-4,
             density_thresh_torso=0.01, torso_shrink=0.8, smooth_lips=True)
    o.update(opt or {})
    L = blob_layout()
    f32, f16 = np.float32, np.float16

    def need(name, shape):
        if name not in sd or tuple(sd[name].shape) != tuple(shape):
            raise ValueError(f"pack_ernerf: {name} is {None if name not in sd else sd[name].shape}, expected {shape}")
        return sd[name].astype(f32)

    A = sd["audio_net.encoder_conv.0.weight"].shape[1]
    head_off = sd["encoder_xy.offsets"].astype(np.int32)
    torso_off = sd["torso_encoder.offsets"].astype(np.int32)
    if head_off.shape != (MF_ERNERF_HEAD_LEVELS + 1,) or torso_off.shape != (MF_ERNERF_TORSO_LEVELS + 1,):
        raise ValueError("pack_ernerf: unexpected number of grid levels")
    rows = int(head_off[-1])
    planes = np.stack([need(f"encoder_{p}.embeddings", (rows, 1))[:, 0] for p in ("xy", "yz", "xz")])
    for p in ("yz", "xz"):
        if not np.array_equal(sd[f"encoder_{p}.offsets"], sd["encoder_xy.offsets"]):
            raise ValueError("pack_ernerf: plane offsets differ")
    G = 128
    bitfield = sd["density_bitfield"].astype(np.uint8)
    if bitfield.shape != (G ** 3 // 8,):
        raise ValueError(f"pack_ernerf: density_bitfield {bitfield.shape}: only cascade=1, grid 128 supported")
    torso_table = need("torso_encoder.embeddings", (int(torso_off[-1]), 2)).astype(f16)
    torso_density = need("density_grid_torso", (G * G,))

    # ---- head MLP image
    img = np.zeros(L["H_BYTES"] // 2, f16)
    _put(img, L["H_AUD1"], need("aud_ch_att_net.net.0.weight", (64, 36)), 56)
    _put(img, L["H_AUD2"], need("aud_ch_att_net.net.1.weight", (32, 64)), 72)
    _put(img, L["H_EYE1"], need("eye_att_net.net.0.weight", (16, 36)), 56)
    img[L["H_EYE2"]:L["H_EYE2"] + 16] = need("eye_att_net.net.1.weight", (1, 16))[0].astype(f16)
    s1 = need("sigma_net.net.0.weight", (64, 69))
    s1p = np.zeros((64, 80), f32)
    s1p[:, 0:36] = s1[:, 0:36]      # enc_x
    s1p[:, 36] = s1[:, 68]          # e = eye * eye_att
    s1p[:, 48:80] = s1[:, 36:68]    # enc_w
    _put(img, L["H_SIG1"], s1p, 88)
    _put(img, L["H_SIG2"], need("sigma_net.net.1.weight", (64, 64)), 72)
    s3 = need("sigma_net.net.2.weight", (65, 64))
    s3p = np.zeros((72, 64), f32)
    s3p[0:64] = s3[1:65]            # geo_feat
    s3p[64] = s3[0]                 # sigma logit
    _put(img, L["H_SIG3"], s3p, 72)
    c1 = need("color_net.net.0.weight", (64, 84))
    c1p = np.zeros((64, 80), f32)
    c1p[:, 0:64] = c1[:, 16:80]     # geo_feat
    c1p[:, 64:80] = c1[:, 0:16]     # SH
    _put(img, L["H_COL1"], c1p, 88)
    c2p = np.zeros((8, 64), f32)
    c2p[0:3] = need("color_net.net.1.weight", (3, 64))
    _put(img, L["H_COL2"], c2p, 72)
    ind = need("individual_codes", sd["individual_codes"].shape)[0]
    if ind.shape != (4,):
        raise ValueError("pack_ernerf: ind_dim must be 4")
    colbias = (c1[:, 80:84].astype(f16).astype(f32) @ ind.astype(f16).astype(f32)).astype(f32)
    head_mlp = img.tobytes()[:L["H_COLBIAS_BYTES"]] + colbias.tobytes()
    assert len(head_mlp) == L["H_BYTES"]

    # ---- torso MLP image
    timg = np.zeros(L["T_HALFS"], f16)
    d1 = need("torso_deform_net.net.0.weight", (32, 84))
    _put(timg, L["T_DEF1"], d1[:, 0:34], 56)
    _put(timg, L["T_DEF2"], need("torso_deform_net.net.1.weight", (32, 32)), 40)
    d3 = np.zeros((8, 32), f32)
    d3[0:2] = need("torso_deform_net.net.2.weight", (2, 32))
    _put(timg, L["T_DEF3"], d3, 40)
    t1 = need("torso_net.net.0.weight", (32, 116))
    _put(timg, L["T_TOR1"], t1[:, 0:66], 88)
    _put(timg, L["T_TOR2"], need("torso_net.net.1.weight", (32, 32)), 40)
    t3 = np.zeros((8, 32), f32)
    t3[0:4] = need("torso_net.net.2.weight", (4, 32))
    _put(timg, L["T_TOR3"], t3, 40)
    torso_const = np.stack([d1[:, 34:84], t1[:, 66:116]]).astype(f16)   # [2][32][50]

    # ---- audio nets, module order
    parts = []
    for i, shp in zip((0, 2, 4, 6), ((32, A, 3), (32, 32, 3), (64, 32, 3), (64, 64, 3))):
        parts += [need(f"audio_net.encoder_conv.{i}.weight", shp), need(f"audio_net.encoder_conv.{i}.bias", shp[:1])]
    parts += [need("audio_net.encoder_fc1.0.weight", (64, 64)), need("audio_net.encoder_fc1.0.bias", (64,)),
              need("audio_net.encoder_fc1.2.weight", (32, 64)), need("audio_net.encoder_fc1.2.bias", (32,))]
    for i, shp in zip((0, 2, 4, 6, 8), ((16, 32, 3), (8, 16, 3), (4, 8, 3), (2, 4, 3), (1, 2, 3))):
        parts += [need(f"audio_att_net.attentionConvNet.{i}.weight", shp),
                  need(f"audio_att_net.attentionConvNet.{i}.bias", shp[:1])]
    parts += [need("audio_att_net.attentionNet.0.weight", (8, 8)), need("audio_att_net.attentionNet.0.bias", (8,))]
    audio = np.concatenate([p.reshape(-1) for p in parts]).astype(f16)

    ind_t = need("individual_codes_torso", sd["individual_codes_torso"].shape)[0]
    if ind_t.shape != (8,):
        raise ValueError("pack_ernerf: ind_dim_torso must be 8")
    misc = np.concatenate([need("anchor_points", (3, 4)).reshape(-1), ind, ind_t]).astype(f32)

    blob = build_blob({ID_HEAD_PLANES: planes.astype(f32).tobytes(), ID_BITFIELD: bitfield.tobytes(),
                       ID_TORSO_TABLE: torso_table.tobytes(), ID_TORSO_DENSITY: torso_density.tobytes(),
                       ID_HEAD_MLP: head_mlp, ID_TORSO_MLP: timg.tobytes(), ID_AUDIO: audio.tobytes(),
                       ID_MISC: misc.tobytes(), ID_TORSO_CONST: torso_const.tobytes()}, kind=1)

    cfg = MfErnerfCfg()
    cfg.bound = o["bound"]
    cfg.min_near = o["min_near"]
    cfg.dt_gamma = o["dt_gamma"]
    cfg.T_thresh = o["T_thresh"]
    cfg.density_thresh_torso = min(o["density_thresh_torso"], float(mean_density_torso))  # renderer.py:325
    cfg.torso_shrink = o["torso_shrink"]
    cfg.max_steps = int(o["max_steps"])
    cfg.cascade = 1 + int(np.ceil(np.log2(o["bound"])))
    cfg.grid_size = G
    cfg.smooth_lips = 1 if o["smooth_lips"] else 0
    # gridencoder/grid.py:98-99: per_level_scale = exp2(log2(desired / base) / (L - 1)); S = log2(scale) as fp32
    cfg.head_log2_scale = float(np.log2(np.exp2(np.log2(512 * o["bound"] / 64) / (MF_ERNERF_HEAD_LEVELS - 1))))
    cfg.head_base = 64
    cfg.torso_log2_scale = float(np.log2(np.exp2(np.log2(2048 / 16) / (MF_ERNERF_TORSO_LEVELS - 1))))
    cfg.torso_base = 16
    for i in range(MF_ERNERF_HEAD_LEVELS + 1):
        cfg.head_offsets[i] = int(head_off[i])
    for i in range(MF_ERNERF_TORSO_LEVELS + 1):
        cfg.torso_offsets[i] = int(torso_off[i])
    cfg.audio_in_dim = int(A)
    return blob, cfg


def load_checkpoint(path):
    """ernerf/nerf_triplane/utils.py:1479-1511: `model` state_dict + mean_density_torso"""
    import torch
    ck = torch.load(path, map_location="cpu", weights_only=False)
    return ck["model"], float(ck.get("mean_density_torso", 0.0))
