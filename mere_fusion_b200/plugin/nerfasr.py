"""NerfASR -- mirror of /root/reference/nerfasr.py:15-151: the 32-slot logits ring and the
[8, audio_dim, 16] attention window.

The acoustic model (a 315 M-parameter HF wav2vec2 CTC head, nerfasr.py:40-45,128-143) is injected as
`feature_fn(float32[n_samples]) -> tensor [T, audio_dim]` (logits of one window).  When none is given, the checkpoint named by
opt.asr_model is loaded through transformers (weights only) and run on the sm_100a engine (mere_fusion_b200.wav2vec2:
mf_wav2vec2_logits); architectures outside the XLSR-53 wav2vec2 family (HuBERT, deepspeech) are refused, not emulated.
"""
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


AUDIO_DIMS = (("esperanto", 44), ("deepspeech", 29), ("hubert", 1024))       # by acoustic-model family; anything else: 32


class NerfASR(BaseASR):
    """One chunk per step; every `m` chunks the (l + m + r)-chunk window goes through the acoustic model and its middle rows land in a
    ring of 4 * m logit rows; `get_next_feat` reads a 16-row window of the ring (advancing 2 rows = one video frame) and, with
    attention, stacks the last 8 of them: [8, audio_dim, 16]."""
    poll_timeout = None                # never wait for audio: the render loop paces itself (nerfasr.py:60-73)
    RING_SEGMENTS = 4
    WINDOW_ROWS = 16

    def __init__(self, opt, parent, feature_fn=None, device=None):
        super().__init__(opt, parent)
        import torch
        self.device = device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu")
        self.audio_dim = next((d for key, d in AUDIO_DIMS if key in self.opt.asr_model), 32)
        self.context_size = opt.m
        self.frames.extend(np.zeros(self.chunk, dtype=np.float32) for _ in range(max(0, self.stride_left_size)))   # left padding
        self.feature_fn = feature_fn if feature_fn is not None else _gpu_feature_fn(opt, self.device)
        self.feat_buffer_size = self.RING_SEGMENTS
        self.feat_buffer_idx = 0
        rows = self.feat_buffer_size * self.context_size
        self.feat_queue = torch.zeros(rows, self.audio_dim, dtype=torch.float32, device=self.device)
        self.front, self.tail = rows - self.WINDOW_ROWS // 2, self.WINDOW_ROWS // 2      # the window straddles the ring's seam at start
        self.att_feats = [torch.zeros(self.audio_dim, self.WINDOW_ROWS, dtype=torch.float32, device=self.device)] * 4
        self.warm_up_steps = self.context_size + self._context()

    def _ring_window(self):
        """rows [front, tail) of the ring (wrapping), transposed to [audio_dim, 16]; then both ends move on by one video frame"""
        import torch
        ring, n = self.feat_queue, self.feat_queue.shape[0]
        rows = ring[self.front:self.tail] if self.front < self.tail else torch.cat((ring[self.front:], ring[:self.tail]))
        self.front, self.tail = (self.front + 2) % n, (self.tail + 2) % n
        return rows.permute(1, 0)

    def get_next_feat(self):
        import torch
        if self.opt.att <= 0:
            return self._ring_window().unsqueeze(0)
        while len(self.att_feats) < 8:
            self.att_feats.append(self._ring_window())
        stacked = torch.stack(self.att_feats, dim=0)
        del self.att_feats[0]
        return stacked

    def run_step(self):
        self._pull(1)
        wave = self._window(at_least=self._context() + self.context_size)
        if wave is None:
            return
        self._keep_context()
        rows = self._middle_logits(wave)
        at = self.feat_buffer_idx * self.context_size
        self.feat_queue[at:at + rows.shape[0]] = rows
        self.feat_buffer_idx = (self.feat_buffer_idx + 1) % self.feat_buffer_size

    def _middle_logits(self, wave):
        """logits of the window without its look-behind rows and all but one of its look-ahead rows (nerfasr.py:139-142)"""
        import torch
        logits = torch.as_tensor(self.feature_fn(wave), dtype=torch.float32, device=self.device)
        T = logits.shape[0]
        return logits[max(0, self.stride_left_size):min(T, T - self.stride_right_size + 1)]

    def warm_up(self):
        for _ in range(self.warm_up_steps):
            self.run_step()
