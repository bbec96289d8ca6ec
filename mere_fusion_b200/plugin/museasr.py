"""MuseASR -- mirror of /root/reference/museasr.py:10-29: 2 * batch chunks per step, Whisper features over the whole
l + 2B + r window (one mf_whisper_features call on the GPU), B chunks of [50, 384] starting at video index l / 2."""
import numpy as np

from .baseasr import BaseASR


class MuseASR(BaseASR):
    def __init__(self, opt, parent, audio_processor):
        super().__init__(opt, parent)
        self.audio_processor = audio_processor
        self.device_chunks = True

    def run_step(self):
        for _ in range(self.batch_size * 2):
            audio_frame, type = self.get_audio_frame()
            self.frames.append(audio_frame)
            self.output_queue.put((audio_frame, type))
        if len(self.frames) <= self.stride_left_size + self.stride_right_size:
            return
        inputs = np.concatenate(self.frames)
        if self.device_chunks and hasattr(self.audio_processor, "audio2chunks_device"):
            # features and the [B, 50, 384] chunk gather stay on the GPU (same rows as feature2chunks; tests compare the two)
            whisper_chunks = self.audio_processor.audio2chunks_device(inputs, fps=self.fps / 2, batch_size=self.batch_size,
                                                                      start=self.stride_left_size / 2)
        else:
            whisper_feature = self.audio_processor.audio2feat(inputs)
            whisper_chunks = self.audio_processor.feature2chunks(feature_array=whisper_feature, fps=self.fps / 2,
                                                                 batch_size=self.batch_size, start=self.stride_left_size / 2)
        self.feat_queue.put(whisper_chunks)
        self.frames = self.frames[-(self.stride_left_size + self.stride_right_size):]
