"""Pack the Whisper audio encoder used by MuseTalk (vendored reference: musetalk/whisper/whisper/model.py:131-171,
loaded from ./models/whisper/tiny.pt by whisper/__init__.py load_model) into a conv-net program for csrc/wav2lip.cu.

  conv1 (k3, pad 1) + GELU -> conv2 (k3, stride 2, pad 1) + GELU + positional embedding -> n_layer x
  [x += attn(ln(x)); x += mlp(ln(x))], keeping the input of the first block and the output of every block
  (forward(include_embeddings=True)).  ln_post only feeds `audio_features`, which audio2feat discards: not computed.

The log-mel front-end (whisper/audio.py:92-125) runs inside the same program (k_logmel_* in convnet_ops.cuh); its 80 x 201
filterbank is generated here (librosa.filters.mel(sr=16000, n_fft=400, n_mels=80), Slaney) and matches the reference's
assets/mel_filters.npz to 1 ulp.
"""
import numpy as np

from .audio_mel import mel_filterbank
from .convnet_pack import ACT_GELU, ProgramBuilder

TINY_DIMS = dict(n_mels=80, n_audio_ctx=1500, n_audio_state=384, n_audio_head=6, n_audio_layer=4)


def _np(v):
    return v.detach().cpu().float().numpy() if hasattr(v, "detach") else np.asarray(v, np.float32)


def sinusoids(length, channels, max_timescale=10000):
    """model.py:49-55, evaluated in float32 like torch does"""
    inc = np.log(max_timescale) / (channels // 2 - 1)
    inv = np.exp((-inc * np.arange(channels // 2)).astype(np.float32)).astype(np.float32)
    t = np.arange(length, dtype=np.float32)[:, None] * inv[None, :]
    return np.concatenate([np.sin(t), np.cos(t)], axis=1).astype(np.float32)


def whisper_filters():
    return mel_filterbank(sr=16000, n_fft=400, n_mels=80, fmin=0.0, fmax=8000.0)


def pack_whisper(sd, dims=TINY_DIMS):
    """sd: state_dict of the whole Whisper model or of its encoder (keys with or without the `encoder.` prefix).
    Returns (blob, ProgramBuilder)."""
    sd = {k: _np(v) for k, v in sd.items()}
    if any(k.startswith("encoder.") for k in sd):
        sd = {k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}
    D, T, Hh, L, M = dims["n_audio_state"], dims["n_audio_ctx"], dims["n_audio_head"], dims["n_audio_layer"], dims["n_mels"]
    assert M == 80 and D % Hh == 0 and D % 16 == 0

    def need(name, shape):
        if name not in sd or tuple(sd[name].shape) != tuple(shape):
            raise ValueError(f"pack_whisper: {name} is {None if name not in sd else sd[name].shape}, expected {tuple(shape)}")
        return sd[name]

    pb = ProgramBuilder(1)
    mel = pb.buffer(2 * T, 1, M)
    pos = need("positional_embedding", (T, D)) if "positional_embedding" in sd else sinusoids(T, D)
    posb = pb.buffer(T, 1, D, init=pos.reshape(T, 1, D))
    c1 = pb.buffer(2 * T, 1, D)
    pb.conv(mel, 0, c1, 0, need("conv1.weight", (D, M, 3))[:, :, :, None], need("conv1.bias", (D,)), padding=(1, 0), relu=ACT_GELU)
    x = pb.buffer(T, 1, D)
    pb.conv(c1, 0, x, 0, need("conv2.weight", (D, D, 3))[:, :, :, None], need("conv2.bias", (D,)), stride=(2, 1), padding=(1, 0),
            relu=ACT_GELU, res=(posb, 0), res_after_act=True)
    embeds = [x]
    ln, qkv, ao, hid = pb.buffer(T, 1, D), pb.buffer(T, 1, 3 * D), pb.buffer(T, 1, D), pb.buffer(T, 1, 4 * D)
    for i in range(L):
        p = f"blocks.{i}."
        pb.layer_norm(x, ln, need(p + "attn_ln.weight", (D,)), need(p + "attn_ln.bias", (D,)))
        wq = np.concatenate([need(p + "attn.query.weight", (D, D)), need(p + "attn.key.weight", (D, D)), need(p + "attn.value.weight", (D, D))])
        bq = np.concatenate([need(p + "attn.query.bias", (D,)), np.zeros(D, np.float32), need(p + "attn.value.bias", (D,))])
        pb.linear(ln, 0, qkv, 0, wq, bq)
        # (q * dh^-1/4) (k * dh^-1/4)^T == q k^T * dh^-1/2  (model.py:84-88)
        pb.attention((qkv, 0), (qkv, D), (qkv, 2 * D), (ao, 0), Hh, D // Hh)
        x1 = pb.buffer(T, 1, D)
        pb.linear(ao, 0, x1, 0, need(p + "attn.out.weight", (D, D)), need(p + "attn.out.bias", (D,)), res=(x, 0))
        pb.layer_norm(x1, ln, need(p + "mlp_ln.weight", (D,)), need(p + "mlp_ln.bias", (D,)))
        pb.linear(ln, 0, hid, 0, need(p + "mlp.0.weight", (4 * D, D)), need(p + "mlp.0.bias", (4 * D,)), act=ACT_GELU)
        x2 = pb.buffer(T, 1, D)
        pb.linear(hid, 0, x2, 0, need(p + "mlp.2.weight", (D, 4 * D)), need(p + "mlp.2.bias", (D,)), res=(x1, 0))
        embeds.append(x2)
        x = x2
    filt = pb._tensor(np.ascontiguousarray(whisper_filters(), np.float32).tobytes())
    pb.aux = [len(embeds)] + embeds + [filt]
    pb.hdr.update(in_face_buf=mel, in_mel_buf=-1, face_hw=0, mel_h=0, mel_w=-2, out_hw=0)
    pb.n_embeds = len(embeds)
    return pb.finish(), pb
