"""Wav2LipEngine: torch tensors in/out around mf_wav2lip_forward (include/mf_b200.h).
Replaces `model(mel_batch, img_batch)` plus the batch build and x255 of lipreal.py:108-126."""
import contextlib
import ctypes

import numpy as np
import torch

from ._lib import Context, check, lib
from .wav2lip_pack import pack_wav2lip


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class ConvNet:
    """a loaded conv-net program (blob from convnet_pack.ProgramBuilder.finish())"""

    def __init__(self, blob, max_batch=16, device=0):
        self.device = torch.device("cuda", device)
        self.ctx = Context(device)
        if isinstance(blob, np.ndarray):
            blob = torch.from_numpy(blob)
        self.blob = blob.to(self.device)
        self.max_batch = max_batch
        check(self.ctx.handle, lib().mf_wav2lip_load(self.ctx.handle, _ptr(self.blob), self.blob.numel(), max_batch),
              "mf_wav2lip_load")

    def debug_set(self, buf, x_nhwc):
        x = x_nhwc.contiguous().float().to(self.device)
        s = torch.cuda.current_stream(self.device)
        check(self.ctx.handle, lib().mf_convnet_debug_set(self.ctx.handle, buf, _ptr(x), x.shape[0], ctypes.c_void_p(s.cuda_stream)),
              "mf_convnet_debug_set")
        torch.cuda.synchronize()

    def debug_run(self, in_buf, x_nhwc, out_buf, out_shape):
        x = x_nhwc.contiguous().float().to(self.device)
        out = torch.empty(out_shape, dtype=torch.float32, device=self.device)
        s = torch.cuda.current_stream(self.device)
        check(self.ctx.handle, lib().mf_convnet_debug_run(self.ctx.handle, in_buf, _ptr(x), out_buf, _ptr(out),
                                                          x.shape[0], ctypes.c_void_p(s.cuda_stream)), "mf_convnet_debug_run")
        return out

    @property
    def last_launches(self):
        return lib().mf_wav2lip_last_launches(self.ctx.handle)

    def profile_op(self, op_index):
        check(self.ctx.handle, lib().mf_wav2lip_profile(self.ctx.handle, int(op_index)), "mf_wav2lip_profile")

    def last_op_ms(self):
        ms = ctypes.c_float()
        check(self.ctx.handle, lib().mf_wav2lip_last_op_ms(self.ctx.handle, ctypes.byref(ms)), "mf_wav2lip_last_op_ms")
        return ms.value


class Wav2LipEngine(ConvNet):
    def __init__(self, state_dict=None, max_batch=16, device=0, blob=None, face_hw=96):
        self.face_hw = face_hw
        self.flops_per_frame = None
        if blob is None:
            blob, pb = pack_wav2lip(state_dict, nominal_batch=max_batch, face_hw=face_hw)
            self.flops_per_frame = pb.flops_per_sample
            self.n_ops = len(pb.ops)
        super().__init__(blob, max_batch, device)

    def forward(self, mel, faces, out=None, out_f32=None, stream=None):
        """mel: cuda fp32 [B,1,80,16]; faces: cuda u8 [B,S,S,3] BGR -> u8 [B,S,S,3]"""
        B = int(faces.shape[0])
        assert mel.is_cuda and faces.is_cuda and mel.dtype == torch.float32 and faces.dtype == torch.uint8
        assert mel.is_contiguous() and faces.is_contiguous() and mel.shape[0] == B
        if out is None:
            out = torch.empty_like(faces)
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        check(self.ctx.handle, lib().mf_wav2lip_forward(self.ctx.handle, _ptr(mel), _ptr(faces), _ptr(out), _ptr(out_f32),
                                                        B, ctypes.c_void_p(s.cuda_stream)), "mf_wav2lip_forward")
        return out

    def forward_host(self, mel_pinned, faces_pinned, out_pinned):
        """end-to-end with HOST buffers (pinned): H2D of mel + faces, forward, D2H of the u8 frames"""
        st = getattr(self, "_stage", None)
        if st is None or st[0].shape != mel_pinned.shape:
            st = self._stage = (torch.empty(mel_pinned.shape, dtype=torch.float32, device=self.device),
                                torch.empty(faces_pinned.shape, dtype=torch.uint8, device=self.device),
                                torch.empty(faces_pinned.shape, dtype=torch.uint8, device=self.device))
        st[0].copy_(mel_pinned, non_blocking=True)
        st[1].copy_(faces_pinned, non_blocking=True)
        self.forward(st[0], st[1], out=st[2])
        out_pinned.copy_(st[2], non_blocking=True)
        return out_pinned



class MelFrontEnd:
    """Wav2Lip mel + chunk slicing on the GPU (mf_wav2lip_mel_chunks): replaces audio.melspectrogram + the slicing loop of
    LipASR.run_step (wav2lip/audio.py:45-51, lipasr.py:24-35) for one window; the chunk start columns are the reference's
    integer arithmetic, computed on the host."""

    def __init__(self, engine):
        from .audio_mel import mel_filterbank
        # a context of its own (mf_wav2lip_mel_chunks needs no loaded program): the front-end runs on the ASR thread while the engine's
        # context is used by the inference thread, and a context is single-user (include/mf_b200.h)
        from ._lib import Context
        self.device = engine.device
        self.ctx = Context(self.device.index if self.device.index is not None else 0)
        self._lock = contextlib.nullcontext()
        self.filters = torch.from_numpy(np.ascontiguousarray(mel_filterbank(), np.float32)).to(self.device)
        self._pin = None
        self._h2d_done = None

    @staticmethod
    def chunk_starts(n_chunks, left_size, right_size, fps, n_cols):
        """lipasr.py:24-35: chunk i starts at int(left * 80 / 50 + i * 160 / fps), clamped so that 16 columns fit"""
        left = max(0, left_size * 80 / 50)
        mult = 80. * 2 / fps
        starts, i = [], 0
        while i < (n_chunks - left_size - right_size) / 2:
            s0 = int(left + i * mult)
            starts.append(n_cols - 16 if s0 + 16 > n_cols else s0)
            i += 1
        return np.asarray(starts, np.int32)

    def chunks(self, audio, n_chunks, left_size, right_size, fps):
        """float32 waveform of the whole window -> cuda fp32 [B, 1, 80, 16]"""
        a = np.ascontiguousarray(audio, np.float32)
        if self._pin is None or self._pin.numel() < a.size:
            self._pin = torch.empty(max(a.size, 16640), dtype=torch.float32).pin_memory()
            self._dev = torch.empty(self._pin.numel(), dtype=torch.float32, device=self.device)
        if self._h2d_done is not None:
            self._h2d_done.synchronize()
        self._pin[:a.size].copy_(torch.from_numpy(a))
        d = self._dev[:a.size]
        d.copy_(self._pin[:a.size], non_blocking=True)
        if self._h2d_done is None:
            self._h2d_done = torch.cuda.Event()
        s = torch.cuda.current_stream(self.device)
        self._h2d_done.record(s)
        starts = self.chunk_starts(n_chunks, left_size, right_size, fps, a.size // 200 + 1)
        out = torch.empty((len(starts), 1, 80, 16), dtype=torch.float32, device=self.device)
        with self._lock:
            check(self.ctx.handle, lib().mf_wav2lip_mel_chunks(self.ctx.handle, _ptr(d), a.size, _ptr(self.filters),
                                                               starts.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), len(starts), _ptr(out),
                                                               ctypes.c_void_p(s.cuda_stream)), "mf_wav2lip_mel_chunks")
        return out
