"""MuseASR -- MuseTalk's audio features (/root/reference/museasr.py:10-29): every step takes 2 * batch chunks, runs Whisper over the
whole l + 2B + r window (one mf_whisper_features call) and hands the inference loop B chunks of [50, 384]; video frame i of the
batch is centred l / 2 frames into the window."""
from .baseasr import BaseASR


class MuseASR(BaseASR):
    def __init__(self, opt, parent, audio_processor):
        super().__init__(opt, parent)
        self.audio_processor = audio_processor
        self.device_chunks = True          # keep features and the [B, 50, 384] gather on the GPU when the processor can

    def run_step(self):
        self._pull(2 * self.batch_size)
        wave = self._window()
        if wave is None:
            return
        ap = self.audio_processor
        where = dict(fps=self.fps / 2, batch_size=self.batch_size, start=self.stride_left_size / 2)
        if self.device_chunks and hasattr(ap, "audio2chunks_device"):
            chunks = ap.audio2chunks_device(wave, **where)        # same rows as feature2chunks (tests compare the two)
        else:
            chunks = ap.feature2chunks(feature_array=ap.audio2feat(wave), **where)
        self.feat_queue.put(chunks)
        self._keep_context()
