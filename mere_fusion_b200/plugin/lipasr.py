"""LipASR -- mirror of /root/reference/lipasr.py:12-37: 2*batch chunks per step, mel over the whole
l + 2B + r window, B mel chunks of [80, 16]."""
import numpy as np

from .. import audio_mel
from .baseasr import BaseASR


class LipASR(BaseASR):
    def run_step(self):
        for _ in range(self.batch_size * 2):
            frame, type = self.get_audio_frame()
            self.frames.append(frame)
            self.output_queue.put((frame, type))
        if len(self.frames) <= self.stride_left_size + self.stride_right_size:
            return
        inputs = np.concatenate(self.frames)
        fe = getattr(self.parent, "mel_front_end", None)
        if fe is not None:
            # mel + chunk slicing on the GPU: a cuda fp32 [B,1,80,16] tensor goes through feat_queue instead of B numpy chunks
            self.feat_queue.put(fe.chunks(inputs, len(self.frames), self.stride_left_size, self.stride_right_size, self.fps))
        else:
            mel = audio_mel.melspectrogram(inputs)
            self.feat_queue.put(mel_chunks(mel, len(self.frames), self.stride_left_size, self.stride_right_size, self.fps))
        self.frames = self.frames[-(self.stride_left_size + self.stride_right_size):]


def mel_chunks(mel, n_frames, left_size, right_size, fps):
    """lipasr.py:24-35: chunk i starts at int(left*80/50 + i * 160/fps), 16 columns, clamped to the tail"""
    left = max(0, left_size * 80 / 50)
    mel_idx_multiplier = 80. * 2 / fps
    mel_step_size = 16
    i = 0
    chunks = []
    while i < (n_frames - left_size - right_size) / 2:
        start_idx = int(left + i * mel_idx_multiplier)
        if start_idx + mel_step_size > len(mel[0]):
            chunks.append(mel[:, len(mel[0]) - mel_step_size:])
        else:
            chunks.append(mel[:, start_idx: start_idx + mel_step_size])
        i += 1
    return chunks
