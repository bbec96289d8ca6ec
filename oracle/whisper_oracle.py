"""oracle/whisper_oracle.py -- CPU restatement of the MuseTalk audio-feature path.  TEST INFRASTRUCTURE ONLY (imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg; nothing under mere_fusion_b200/ may import it).

Follows the vendored reference:
  log_mel        musetalk/whisper/whisper/audio.py:92-125   (torch.stft n_fft 400 hop 160 periodic Hann, centre/reflect,
                                                             drop the last frame, |.|^2, 80 x 201 filterbank, log10,
                                                             max(., max - 8), (x + 4) / 4)
  encoder        musetalk/whisper/whisper/model.py:143-171  (conv1/conv2 + GELU, + positional embedding, residual attention
                                                             blocks model.py:112-128, attention :84-96)
  audio2feat     musetalk/whisper/audio2feature.py:99-112 + whisper/transcribe.py:85-128 (zero-pad the log-mel to 3000 frames,
                                                             keep the input of block 0 and every block output, first
                                                             int((end - start) / 2) rows)
  slicing        musetalk/whisper/audio2feature.py:16-45, 82-97

PINNED: tests/test_oracle_whisper.py checks every function here against tests/golden/whisper_golden.npz, which was produced
by importing the reference modules themselves (tests/golden/make_whisper_golden.py).
"""
import numpy as np
import torch
import torch.nn.functional as F

N_FFT, HOP, N_FRAMES = 400, 160, 3000


def mel_filters():
    """librosa.filters.mel(sr=16000, n_fft=400, n_mels=80) restated (Slaney scale + area normalisation); the reference ships
    the same matrix as assets/mel_filters.npz (agrees to 1 ulp)"""
    def hz_to_mel(f):
        f = np.asarray(f, np.float64)
        return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-10) / 1000.0) / (np.log(6.4) / 27.0), f / (200.0 / 3))

    def mel_to_hz(m):
        m = np.asarray(m, np.float64)
        return np.where(m >= 15.0, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), (200.0 / 3) * m)

    fft = np.linspace(0, 8000.0, 201)
    mf = mel_to_hz(np.linspace(hz_to_mel(0.0), hz_to_mel(8000.0), 82))
    ramps = np.subtract.outer(mf, fft)
    w = np.zeros((80, 201))
    for i in range(80):
        w[i] = np.maximum(0, np.minimum(-ramps[i] / (mf[i + 1] - mf[i]), ramps[i + 2] / (mf[i + 2] - mf[i + 1])))
    w *= (2.0 / (mf[2:82] - mf[:80]))[:, None]
    return w.astype(np.float32)


def log_mel(audio):
    """audio.py:92-125 in numpy: float32 [n] -> float32 [80, n // 160]"""
    a = np.asarray(audio, np.float32)
    n = len(a)
    pad = np.pad(a, N_FFT // 2, mode="reflect")
    nf = 1 + n // HOP
    idx = np.arange(N_FFT)[None, :] + HOP * np.arange(nf)[:, None]
    win = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N_FFT) / N_FFT)).astype(np.float32)
    spec = np.fft.rfft((pad[idx] * win).astype(np.float32), axis=1)            # [nf, 201]
    mag = (np.abs(spec[:-1]) ** 2).astype(np.float32).T                         # drop the last frame
    m = mel_filters() @ mag
    ls = np.log10(np.maximum(m, 1e-10))
    ls = np.maximum(ls, ls.max() - 8.0)
    return ((ls + 4.0) / 4.0).astype(np.float32)


def encoder_embeddings(sd, mel, dims):
    """model.py:143-171 with include_embeddings=True, fp32: mel [80, 3000] -> [n_layer + 1, 1500, D]"""
    W = {k: torch.as_tensor(np.asarray(v, np.float32)) for k, v in sd.items()}
    H = dims["n_audio_head"]
    x = torch.as_tensor(mel, dtype=torch.float32)[None]
    x = F.gelu(F.conv1d(x, W["conv1.weight"], W["conv1.bias"], padding=1))
    x = F.gelu(F.conv1d(x, W["conv2.weight"], W["conv2.bias"], stride=2, padding=1))
    x = x.permute(0, 2, 1) + W["positional_embedding"]
    embs = [x[0].numpy().copy()]
    D = x.shape[-1]
    scale = (D // H) ** -0.25
    for i in range(dims["n_audio_layer"]):
        p = f"blocks.{i}."
        h = F.layer_norm(x, (D,), W[p + "attn_ln.weight"], W[p + "attn_ln.bias"])
        q = F.linear(h, W[p + "attn.query.weight"], W[p + "attn.query.bias"])
        k = F.linear(h, W[p + "attn.key.weight"])
        v = F.linear(h, W[p + "attn.value.weight"], W[p + "attn.value.bias"])
        T = q.shape[1]
        q = q.view(1, T, H, -1).permute(0, 2, 1, 3) * scale
        k = k.view(1, T, H, -1).permute(0, 2, 3, 1) * scale
        v = v.view(1, T, H, -1).permute(0, 2, 1, 3)
        a = (torch.softmax(q @ k, dim=-1) @ v).permute(0, 2, 1, 3).flatten(start_dim=2)
        x = x + F.linear(a, W[p + "attn.out.weight"], W[p + "attn.out.bias"])
        h = F.layer_norm(x, (D,), W[p + "mlp_ln.weight"], W[p + "mlp_ln.bias"])
        x = x + F.linear(F.gelu(F.linear(h, W[p + "mlp.0.weight"], W[p + "mlp.0.bias"])), W[p + "mlp.2.weight"], W[p + "mlp.2.bias"])
        embs.append(x[0].numpy().copy())
    return np.stack(embs)


def audio2feat(sd, audio, dims):
    """audio2feature.py:99-112 for one segment: float32 [n] -> float32 [int(n_frames / 2), n_layer + 1, D]"""
    mel = log_mel(audio)
    nf = mel.shape[1]
    assert nf <= N_FRAMES, "one 30 s segment"
    seg = np.zeros((80, N_FRAMES), np.float32)
    seg[:, :nf] = mel                                                           # pad_or_trim: zeros AFTER normalisation
    with torch.no_grad():
        e = encoder_embeddings(sd, seg, dims)                                   # [5, 1500, D]
    return e.transpose(1, 0, 2)[:int(nf / 2)]


def get_sliced_feature(feature_array, vid_idx, audio_feat_length=(2, 2), fps=25):
    """audio2feature.py:16-45"""
    length = len(feature_array)
    c = int(vid_idx * 50 / fps)
    idx = [min(length - 1, max(0, i)) for i in range(c - audio_feat_length[0] * 2, c + (audio_feat_length[1] + 1) * 2)]
    return np.concatenate([feature_array[i] for i in idx], axis=0).reshape(-1, 384), idx


def feature2chunks(feature_array, fps, batch_size, audio_feat_length=(2, 2), start=0):
    """audio2feature.py:82-97"""
    return [get_sliced_feature(feature_array, i + start, audio_feat_length, fps)[0] for i in range(batch_size)]
