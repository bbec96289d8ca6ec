"""BaseASR -- the audio side of the plugin surface (/root/reference/baseasr.py:9-63).

Public contract kept for the orchestration and for subclasses written against the reference: constructor `(opt, parent)`,
attributes `fps, sample_rate, chunk, queue, output_queue, feat_queue, batch_size, frames, stride_left_size, stride_right_size`,
methods `put_audio_frame, get_audio_frame, get_audio_out, warm_up, pause_talk, run_step, get_next_feat`, and the
`(chunk, type)` pairs on `output_queue` (type 0 = speech, 1 = silence, > 1 = custom audio state of the parent).

Built differently from the reference where it does not show: the queues are thread queues (inference runs in this process, not
in an mp.Process), and the per-step bookkeeping that every subclass of the reference repeats inline -- pull chunks, echo them to
the output side, test for enough context, cut the window, keep the l + r context -- lives here once (`_pull`, `_window`,
`_keep_context`).
"""
import queue as _queue
from queue import Queue

import numpy as np

SPEECH, SILENCE = 0, 1


class BaseASR:
    poll_timeout = 0.01          # how long get_audio_frame waits for a chunk before it substitutes one (None: do not wait)

    def __init__(self, opt, parent=None):
        self.opt, self.parent = opt, parent
        self.fps = opt.fps                                   # chunks per second: 50 -> 20 ms
        self.sample_rate = 16000
        self.chunk = self.sample_rate // self.fps            # samples per chunk (320)
        self.batch_size = opt.batch_size
        self.stride_left_size, self.stride_right_size = opt.l, opt.r
        self.queue = Queue()                                 # chunks in
        self.output_queue = Queue()                          # (chunk, type) out, aligned with the features
        self.feat_queue = Queue(2)                           # feature batches for the inference loop
        self.frames = []                                     # the sliding window, chunk by chunk

    # ---- intake ------------------------------------------------------------------------------------------
    def put_audio_frame(self, audio_chunk):
        """one 20 ms chunk of 16 kHz float32 PCM"""
        self.queue.put(audio_chunk)

    def pause_talk(self):
        with self.queue.mutex:
            self.queue.queue.clear()

    def _substitute(self):
        """what plays when nothing was said: the parent's custom audio for its current state, else silence"""
        state = self.parent.curr_state if self.parent else 0
        if state > SILENCE:
            return self.parent.get_audio_stream(state), state
        return np.zeros(self.chunk, dtype=np.float32), SILENCE

    def get_audio_frame(self):
        try:
            if self.poll_timeout is None:
                return self.queue.get_nowait(), SPEECH
            return self.queue.get(timeout=self.poll_timeout), SPEECH
        except _queue.Empty:
            return self._substitute()

    # ---- the step bookkeeping shared by the heads ---------------------------------------------------------
    def _pull(self, n):
        """n chunks into the window, each echoed to the output side with its type"""
        for _ in range(n):
            item = self.get_audio_frame()
            self.frames.append(item[0])
            self.output_queue.put(item)

    def _context(self):
        return self.stride_left_size + self.stride_right_size

    def _window(self, at_least=None):
        """the window as one waveform, or None while it holds no more than the look-behind / look-ahead context"""
        need = self._context() + 1 if at_least is None else at_least
        return np.concatenate(self.frames) if len(self.frames) >= need else None

    def _keep_context(self):
        self.frames = self.frames[-self._context():]

    # ---- output side -------------------------------------------------------------------------------------
    def get_audio_out(self):
        return self.output_queue.get()

    def warm_up(self):
        """fill the context, then drop the look-behind part from the output side so that audio out lines up with the features"""
        self._pull(self._context())
        for _ in range(self.stride_left_size):
            self.output_queue.get()

    def run_step(self):
        pass

    def get_next_feat(self, block, timeout):
        return self.feat_queue.get(block, timeout)
