"""NerfASR -- mirror of /root/reference/nerfasr.py:15-151: the 32-slot logits ring and the
[8, audio_dim, 16] attention window.

The acoustic model (a 315 M-parameter HF wav2vec2 CTC head, nerfasr.py:40-45,128-143) is injected as
`feature_fn(float32[n_samples]) -> tensor [T, audio_dim]` (logits of one window).  When none is given, the checkpoint named by
opt.asr_model is loaded through transformers (weights only) and run on the sm_100a engine (mere_fusion_b200.wav2vec2:
mf_wav2vec2_logits); architectures outside the XLSR-53 wav2vec2 family (HuBERT, deepspeech) are refused, not emulated.
"""
import queue

import numpy as np

from .baseasr import BaseASR


def _gpu_feature_fn(opt, device):
    """AutoModelForCTC.from_pretrained(opt.asr_model) only supplies the weights; the forward pass is mf_wav2vec2_logits"""
    import torch
    from transformers import AutoModelForCTC
    from ..wav2vec2 import engine_from_hf
    model = AutoModelForCTC.from_pretrained(opt.asr_model)
    n_samples = (opt.l + opt.m + opt.r) * (16000 // opt.fps)
    dev = torch.device(device)
    engine = engine_from_hf(model, n_samples=n_samples, device=dev.index or 0)
    del model
    return engine.feature_fn


class NerfASR(BaseASR):
    def __init__(self, opt, parent, feature_fn=None, device=None):
        super().__init__(opt, parent)
        import torch
        self.device = device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu")
        if "esperanto" in self.opt.asr_model:
            self.audio_dim = 44
        elif "deepspeech" in self.opt.asr_model:
            self.audio_dim = 29
        elif "hubert" in self.opt.asr_model:
            self.audio_dim = 1024
        else:
            self.audio_dim = 32
        self.context_size = opt.m
        self.stride_left_size = opt.l
        self.stride_right_size = opt.r
        if self.stride_left_size > 0:                                   # nerfasr.py:35-36
            self.frames.extend([np.zeros(self.chunk, dtype=np.float32)] * self.stride_left_size)
        self.feature_fn = feature_fn if feature_fn is not None else _gpu_feature_fn(opt, self.device)
        self.feat_buffer_size = 4
        self.feat_buffer_idx = 0
        self.feat_queue = torch.zeros(self.feat_buffer_size * self.context_size, self.audio_dim, dtype=torch.float32,
                                      device=self.device)
        self.front = self.feat_buffer_size * self.context_size - 8
        self.tail = 8
        self.att_feats = [torch.zeros(self.audio_dim, 16, dtype=torch.float32, device=self.device)] * 4
        self.warm_up_steps = self.context_size + self.stride_left_size + self.stride_right_size

    def get_audio_frame(self):
        """nerfasr.py:60-73: non-blocking variant"""
        try:
            frame = self.queue.get(block=False)
            type = 0
        except queue.Empty:
            if self.parent and self.parent.curr_state > 1:
                frame = self.parent.get_audio_stream(self.parent.curr_state)
                type = self.parent.curr_state
            else:
                frame = np.zeros(self.chunk, dtype=np.float32)
                type = 1
        return frame, type

    def _window(self):
        import torch
        if self.front < self.tail:
            feat = self.feat_queue[self.front:self.tail]
        else:
            feat = torch.cat([self.feat_queue[self.front:], self.feat_queue[:self.tail]], dim=0)
        self.front = (self.front + 2) % self.feat_queue.shape[0]
        self.tail = (self.tail + 2) % self.feat_queue.shape[0]
        return feat

    def get_next_feat(self):
        """nerfasr.py:75-103 -> [8, audio_dim, 16] (att > 0) or [1, audio_dim, 16]"""
        import torch
        if self.opt.att > 0:
            while len(self.att_feats) < 8:
                self.att_feats.append(self._window().permute(1, 0))
            att_feat = torch.stack(self.att_feats, dim=0)
            self.att_feats = self.att_feats[1:]
        else:
            att_feat = self._window().permute(1, 0).unsqueeze(0)
        return att_feat

    def run_step(self):
        """nerfasr.py:105-124"""
        frame, type = self.get_audio_frame()
        self.frames.append(frame)
        self.output_queue.put((frame, type))
        if len(self.frames) < self.stride_left_size + self.context_size + self.stride_right_size:
            return
        inputs = np.concatenate(self.frames)
        self.frames = self.frames[-(self.stride_left_size + self.stride_right_size):]
        feats = self._frame_to_text(inputs)
        start = self.feat_buffer_idx * self.context_size
        end = start + feats.shape[0]
        self.feat_queue[start:end] = feats
        self.feat_buffer_idx = (self.feat_buffer_idx + 1) % self.feat_buffer_size

    def _frame_to_text(self, frame):
        """nerfasr.py:128-143: logits of the window, left/right stride rows cut"""
        import torch
        logits = torch.as_tensor(self.feature_fn(frame), dtype=torch.float32, device=self.device)
        left = max(0, self.stride_left_size)
        right = min(logits.shape[0], logits.shape[0] - self.stride_right_size + 1)
        return logits[left:right]

    def warm_up(self):
        for _ in range(self.warm_up_steps):
            self.run_step()
