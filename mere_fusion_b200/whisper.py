"""WhisperEngine + Audio2Feature: the MuseTalk audio-feature path on the GPU.

Replaces Audio2Feature.audio2feat -> Whisper.transcribe -> log_mel_spectrogram + AudioEncoder.forward(include_embeddings=True)
(musetalk/whisper/audio2feature.py:99-112, whisper/transcribe.py:85-128, whisper/audio.py:92-125, whisper/model.py:143-171)
with ONE C-ABI call (mf_whisper_features): log-mel, zero padding to the 30 s context, the encoder, and the
[T, n_layer + 1, n_state] embedding gather all happen on the device.  The index arithmetic of get_sliced_feature /
feature2chunks (audio2feature.py:16-45, 82-97) is mirrored exactly on the host (it is a handful of integers)."""
import ctypes
import threading

import numpy as np
import torch

from ._lib import check, lib
from .wav2lip import ConvNet, _ptr
from .whisper_pack import TINY_DIMS, pack_whisper


class WhisperEngine(ConvNet):
    def __init__(self, state_dict=None, dims=TINY_DIMS, device=0, blob=None):
        self.dims = dims
        if blob is None:
            blob, pb = pack_whisper(state_dict, dims)
            self.flops_per_call = pb.flops_per_sample
        self.n_embeds = dims["n_audio_layer"] + 1
        self.n_state = dims["n_audio_state"]
        self._host_lock = threading.Lock()
        self._last_done = None
        super().__init__(blob, 1, device)

    def features(self, audio, T=None, out=None, stream=None):
        """audio: cuda fp32 [n] (16 kHz) -> cuda fp32 [T, n_layer + 1, n_state]; T defaults to int(n_frames / 2), the
        rows audio2feat keeps (audio2feature.py:106-109)"""
        assert audio.is_cuda and audio.dtype == torch.float32 and audio.is_contiguous() and audio.dim() == 1
        n = int(audio.shape[0])
        if T is None:
            T = int((n // 160) / 2)
        fixed = out is None
        if fixed:
            # the executor re-captures its CUDA graph whenever the caller's pointers change: without an `out` the program writes into a
            # per-T buffer that stays put and the caller gets a copy (one small D2D kernel instead of a possible capture per window)
            outs = self.__dict__.setdefault("_outs", {})
            out = outs.get(T)
            if out is None:
                out = outs[T] = torch.empty((T, self.n_embeds, self.n_state), dtype=torch.float32, device=self.device)
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        # one audio processor may serve several MuseReal sessions (threads, each on its own stream): the context and its
        # activation workspace are single-user, so calls are serialised on the host (lock) AND on the device (event chain)
        with self._host_lock:
            if self._last_done is not None:
                s.wait_event(self._last_done)
            check(self.ctx.handle, lib().mf_whisper_features(self.ctx.handle, _ptr(audio), n, _ptr(out), T, ctypes.c_void_p(s.cuda_stream)),
                  "mf_whisper_features")
            if fixed:
                with torch.cuda.stream(s):
                    out = out.clone()
            ev = torch.cuda.Event()
            ev.record(s)
            self._last_done = ev
        return out


def get_sliced_feature(feature_array, vid_idx, audio_feat_length=(2, 2), fps=25):
    """audio2feature.py:16-45"""
    length = len(feature_array)
    center_idx = int(vid_idx * 50 / fps)
    left_idx = center_idx - audio_feat_length[0] * 2
    right_idx = center_idx + (audio_feat_length[1] + 1) * 2
    selected_idx = [min(length - 1, max(0, idx)) for idx in range(left_idx, right_idx)]
    selected_feature = np.concatenate([feature_array[i] for i in selected_idx], axis=0).reshape(-1, 384)
    return selected_feature, selected_idx


def feature2chunks(feature_array, fps, batch_size, audio_feat_length=(2, 2), start=0):
    """audio2feature.py:82-97"""
    return [get_sliced_feature(feature_array, i + start, audio_feat_length, fps)[0] for i in range(batch_size)]


class Audio2Feature:
    """drop-in for musetalk/whisper/audio2feature.py:Audio2Feature on the GPU engine"""

    def __init__(self, whisper_model_type="tiny", model_path="./models/whisper/tiny.pt", state_dict=None, device=0, engine=None):
        self.whisper_model_type = whisper_model_type
        if engine is None:
            if state_dict is None:
                ck = torch.load(model_path, map_location="cpu")          # whisper/__init__.py load_model
                state_dict = ck["model_state_dict"]
            engine = WhisperEngine(state_dict, device=device)
        self.engine = engine
        self._pin = None
        self._h2d_done = None

    get_sliced_feature = staticmethod(get_sliced_feature)
    feature2chunks = staticmethod(feature2chunks)

    def chunk_indices(self, n_rows, fps, batch_size, audio_feat_length=(2, 2), start=0):
        """row indices feature2chunks would select (audio2feature.py:16-45, 82-97): int64 [batch_size, 10]"""
        idx = []
        for i in range(batch_size):
            c = int((i + start) * 50 / fps)
            idx.append([min(n_rows - 1, max(0, j)) for j in range(c - audio_feat_length[0] * 2, c + (audio_feat_length[1] + 1) * 2)])
        return np.asarray(idx, np.int64)

    def audio2chunks_device(self, audio, fps, batch_size, start=0, audio_dev=None):
        """audio2feat + feature2chunks without leaving the GPU: float32 waveform -> cuda fp16 [batch_size, 50, 384] (the gather
        is index plumbing on device memory; the indices are the reference's integer arithmetic, computed on the host)"""
        if audio_dev is None:
            a = np.ascontiguousarray(audio, np.float32)
            self._stage(a.size)
            if self._h2d_done is not None:
                self._h2d_done.synchronize()       # the previous window's async H2D must have read the pinned buffer
            self._pin[:a.size].copy_(torch.from_numpy(a))
            audio_dev = self._dev[:a.size]
            audio_dev.copy_(self._pin[:a.size], non_blocking=True)
            if self._h2d_done is None:
                self._h2d_done = torch.cuda.Event()
            self._h2d_done.record(torch.cuda.current_stream(self.engine.device))
        feat = self.engine.features(audio_dev)                                      # [T, 5, 384] fp32
        key = (int(feat.shape[0]), float(fps), int(batch_size), float(start))
        if getattr(self, "_idx_key", None) != key:
            self._idx = torch.from_numpy(self.chunk_indices(feat.shape[0], fps, batch_size, start=start).reshape(-1)).to(feat.device)
            self._idx_key = key
        return feat.index_select(0, self._idx).reshape(batch_size, -1, feat.shape[2]).to(torch.float16)

    def _stage(self, n):
        if self._pin is None or self._pin.numel() < n:
            self._pin = torch.empty(max(n, 16640), dtype=torch.float32).pin_memory()
            self._dev = torch.empty(self._pin.numel(), dtype=torch.float32, device=self.engine.device)

    def audio2feat(self, audio):
        """float32 waveform (16 kHz) -> np.float32 [T, 5, 384]; one 30 s segment (the live window is 0.33 s)"""
        a = np.ascontiguousarray(audio, np.float32)
        self._stage(a.size)
        self._pin[:a.size].copy_(torch.from_numpy(a))
        d = self._dev[:a.size]
        d.copy_(self._pin[:a.size], non_blocking=True)
        out = self.engine.features(d)
        return out.cpu().numpy()
