"""LipASR -- Wav2Lip's audio features (/root/reference/lipasr.py:12-37): every step takes 2 * batch chunks, computes the mel of the
whole l + 2B + r window and hands the inference loop B slices of [80, 16]."""
from .. import audio_mel
from .baseasr import BaseASR

MEL_COLS_PER_CHUNK = 80 / 50          # 80 mel frames per second at 50 chunks per second
MEL_STEP = 16                         # columns per video frame (wav2lip hparams: syncnet_mel_step_size)


def mel_chunk_starts(n_frames, left_size, right_size, fps, n_cols):
    """first mel column of each of the (n_frames - l - r) / 2 video frames of a window: the look-behind chunks are skipped, a video
    frame advances 160 / fps columns, and a slice that would run past the end is moved back to the last MEL_STEP columns"""
    n = max(0, -(-(n_frames - left_size - right_size) // 2))
    first, step = max(0, left_size * MEL_COLS_PER_CHUNK), 80. * 2 / fps
    return [min(int(first + i * step), n_cols - MEL_STEP) for i in range(n)]


def mel_chunks(mel, n_frames, left_size, right_size, fps):
    return [mel[:, s:s + MEL_STEP] for s in mel_chunk_starts(n_frames, left_size, right_size, fps, mel.shape[1])]


class LipASR(BaseASR):
    def run_step(self):
        self._pull(2 * self.batch_size)
        wave = self._window()
        if wave is None:
            return
        front_end = getattr(self.parent, "mel_front_end", None)
        if front_end is not None:      # mel + slicing on the GPU: one cuda fp32 [B, 1, 80, 16] tensor instead of B numpy slices
            feats = front_end.chunks(wave, len(self.frames), self.stride_left_size, self.stride_right_size, self.fps)
        else:
            feats = mel_chunks(audio_mel.melspectrogram(wave), len(self.frames), self.stride_left_size, self.stride_right_size, self.fps)
        self.feat_queue.put(feats)
        self._keep_context()
