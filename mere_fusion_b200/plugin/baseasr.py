"""BaseASR -- mirror of /root/reference/baseasr.py:9-63.

Same attributes and (chunk, type) contract: type 0 = speech, 1 = silence, >1 = custom audio.
`output_queue` / `feat_queue` are thread queues here (the reference uses mp.Queue because its
inference runs in a child process; ours is in-process) with the same put/get API and depths.
"""
import queue
from queue import Queue

import numpy as np


class BaseASR:
    def __init__(self, opt, parent=None):
        self.opt = opt
        self.parent = parent
        self.fps = opt.fps                               # 20 ms per chunk at fps = 50
        self.sample_rate = 16000
        self.chunk = self.sample_rate // self.fps        # 320 samples
        self.queue = Queue()
        self.output_queue = Queue()
        self.batch_size = opt.batch_size
        self.frames = []
        self.stride_left_size = opt.l
        self.stride_right_size = opt.r
        self.feat_queue = Queue(2)

    def pause_talk(self):
        self.queue.queue.clear()

    def put_audio_frame(self, audio_chunk):              # 16 kHz, 20 ms PCM float32[320]
        self.queue.put(audio_chunk)

    def get_audio_frame(self):
        """baseasr.py:36-48: next chunk, else custom-audio slice, else silence"""
        try:
            frame = self.queue.get(block=True, timeout=0.01)
            type = 0
        except queue.Empty:
            if self.parent and self.parent.curr_state > 1:
                frame = self.parent.get_audio_stream(self.parent.curr_state)
                type = self.parent.curr_state
            else:
                frame = np.zeros(self.chunk, dtype=np.float32)
                type = 1
        return frame, type

    def get_audio_out(self):
        return self.output_queue.get()

    def warm_up(self):
        """baseasr.py:53-59: prefill l + r chunks, drop l of them from the output side"""
        for _ in range(self.stride_left_size + self.stride_right_size):
            audio_frame, type = self.get_audio_frame()
            self.frames.append(audio_frame)
            self.output_queue.put((audio_frame, type))
        for _ in range(self.stride_left_size):
            self.output_queue.get()

    def run_step(self):
        pass

    def get_next_feat(self, block, timeout):
        return self.feat_queue.get(block, timeout)
