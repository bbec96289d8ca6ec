"""Wav2Vec2Engine: the CTC acoustic model of the ErNeRF audio path on the GPU.  Replaces `processor(frame) -> model(input_values)
.logits` of NerfASR.__frame_to_text (nerfasr.py:128-143) with ONE C-ABI call (mf_wav2vec2_logits): waveform normalisation, the
7-layer conv feature encoder, the positional conv, the transformer and the lm_head all run on the device."""
import ctypes

import numpy as np
import torch

from ._lib import check, lib
from .wav2lip import ConvNet, _ptr
from .wav2vec2_pack import XLSR53_CFG, pack_wav2vec2


class Wav2Vec2Engine(ConvNet):
    def __init__(self, state_dict=None, cfg=XLSR53_CFG, n_samples=8960, device=0, blob=None, n_frames=None, max_batch=1):
        """max_batch > 1: `logits_batch` runs the windows of several sessions in one pass (one engine per GPU)"""
        self.n_samples, self.vocab = n_samples, cfg["vocab"]
        if blob is None:
            # one window per pass: the transformer layers run as ONE persistent kernel (csrc/w2v_stack.cuh); an engine that batches the
            # windows of several sessions (max_batch > 1) keeps the op-by-op program, whose GEMMs take any number of rows
            blob, pb = pack_wav2vec2(state_dict, cfg, n_samples, fused_stack=max_batch == 1)
            self.flops_per_call = pb.flops_per_sample
            n_frames = pb.n_frames
        self.n_frames = n_frames
        super().__init__(blob, max_batch, device)
        self._pin = torch.empty(n_samples, dtype=torch.float32).pin_memory()
        self._dev = torch.empty(n_samples, dtype=torch.float32, device=self.device)
        self._h2d_done = None
        # the executor replays a CUDA graph that is re-captured whenever the caller's pointers change: without an `out` the program
        # writes into this fixed buffer and the caller gets a copy (one tiny D2D kernel instead of a capture + instantiate per window)
        self._out = torch.empty((max_batch, self.n_frames, self.vocab), dtype=torch.float32, device=self.device)

    def logits(self, audio, out=None, stream=None):
        """audio: cuda fp32 [n_samples] -> cuda fp32 [n_frames, vocab]"""
        assert audio.is_cuda and audio.dtype == torch.float32 and audio.is_contiguous() and audio.numel() == self.n_samples
        dst = out if out is not None else self._out[0]
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        check(self.ctx.handle, lib().mf_wav2vec2_logits(self.ctx.handle, _ptr(audio), self.n_samples, _ptr(dst), ctypes.c_void_p(s.cuda_stream)),
              "mf_wav2vec2_logits")
        if out is not None:
            return out
        with torch.cuda.stream(s):
            return dst.clone()

    def logits_batch(self, audio, out=None, stream=None):
        """audio: cuda fp32 [B, n_samples] (B <= max_batch) -> cuda fp32 [B, n_frames, vocab]"""
        assert audio.is_cuda and audio.dtype == torch.float32 and audio.is_contiguous() and audio.dim() == 2 and audio.shape[1] == self.n_samples
        B = int(audio.shape[0])
        dst = out if out is not None else self._out[:B]
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        check(self.ctx.handle, lib().mf_wav2vec2_logits_batch(self.ctx.handle, _ptr(audio), self.n_samples, B, _ptr(dst),
                                                              ctypes.c_void_p(s.cuda_stream)), "mf_wav2vec2_logits_batch")
        if out is not None:
            return out
        with torch.cuda.stream(s):
            return dst.clone()

    def feature_fn(self, frame):
        """NerfASR's `feature_fn(float32[n_samples]) -> [T, audio_dim]` (device tensor)"""
        a = np.ascontiguousarray(frame, np.float32)
        if self._h2d_done is not None:
            self._h2d_done.synchronize()
        self._pin.copy_(torch.from_numpy(a))
        self._dev.copy_(self._pin, non_blocking=True)
        if self._h2d_done is None:
            self._h2d_done = torch.cuda.Event()
        self._h2d_done.record(torch.cuda.current_stream(self.device))
        return self.logits(self._dev)


def engine_from_hf(model, n_samples=8960, device=0):
    """build the engine from a loaded HF `Wav2Vec2ForCTC` (what AutoModelForCTC.from_pretrained(opt.asr_model) returns for the
    default cpierse/wav2vec2-large-xlsr-53-esperanto, nerfasr.py:44-45).  Only the architecture family the live default uses is
    supported (layer-norm feature encoder, stable-layer-norm transformer, no adapter); anything else is refused, not emulated."""
    c = model.config
    if getattr(c, "model_type", "") != "wav2vec2" or c.feat_extract_norm != "layer" or not c.do_stable_layer_norm or \
            getattr(c, "add_adapter", False) or c.feat_extract_activation != "gelu" or c.hidden_act != "gelu" or not c.conv_bias:
        raise ValueError("mere_fusion_b200.wav2vec2: unsupported acoustic-model architecture (expected the XLSR-53 wav2vec2 family)")
    cfg = dict(vocab=c.vocab_size, hidden=c.hidden_size, layers=c.num_hidden_layers, heads=c.num_attention_heads, inter=c.intermediate_size,
               conv_dim=tuple(c.conv_dim), conv_stride=tuple(c.conv_stride), conv_kernel=tuple(c.conv_kernel),
               pos_k=c.num_conv_pos_embeddings, pos_groups=c.num_conv_pos_embedding_groups, eps=c.layer_norm_eps)
    return Wav2Vec2Engine(model.state_dict(), cfg, n_samples=n_samples, device=device)
