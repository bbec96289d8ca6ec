"""BaseReal resolution.

Inside a mere-fusion deployment the reference's own `basereal.BaseReal` is importable (its repo
root is on sys.path, with PyAV / soundfile / the TTS clients installed) and is used UNCHANGED as
the base class, so TTS, custom idle video/audio and mp4 recording behave exactly as before
(basereal.py:32-154).  Outside of one (tests, bench, this image) a local stand-in with the same
attributes and the pure-logic methods is used; TTS and recording then raise if called.
"""
import glob
import os
from queue import Queue

try:                                                      # deployment: the untouched reference class
    from basereal import BaseReal as _RefBaseReal         # type: ignore
    BaseReal = _RefBaseReal
    USING_REFERENCE_BASE = True
except Exception:                                         # noqa: BLE001 (any missing dependency of the reference)
    USING_REFERENCE_BASE = False

    class _NoTTS:
        def render(self, quit_event):
            pass

        def put_msg_txt(self, msg):
            raise RuntimeError("no TTS backend in this environment (reference ttsreal.py not importable)")

        def pause_talk(self):
            pass

    class BaseReal:
        def __init__(self, opt):
            self.opt = opt
            self.sample_rate = 16000
            self.chunk = self.sample_rate // opt.fps
            self.tts = _NoTTS()
            self.recording = False
            self.recordq_video = Queue()
            self.recordq_audio = Queue()
            self.curr_state = 0
            self.custom_img_cycle = {}
            self.custom_audio_cycle = {}
            self.custom_audio_index = {}
            self.custom_index = {}
            self.custom_opt = {}
            self._loadcustom()

        def _loadcustom(self):
            """basereal.py:59-68"""
            for item in getattr(self.opt, "customopt", []) or []:
                import cv2
                import soundfile as sf
                lst = glob.glob(os.path.join(item["imgpath"], "*.[jpJP][pnPN]*[gG]"))
                lst = sorted(lst, key=lambda x: int(os.path.splitext(os.path.basename(x))[0]))
                self.custom_img_cycle[item["audiotype"]] = [cv2.imread(p) for p in lst]
                self.custom_audio_cycle[item["audiotype"]], _ = sf.read(item["audiopath"], dtype="float32")
                self.custom_audio_index[item["audiotype"]] = 0
                self.custom_index[item["audiotype"]] = 0
                self.custom_opt[item["audiotype"]] = item

        def init_customindex(self):
            self.curr_state = 0
            for key in self.custom_audio_index:
                self.custom_audio_index[key] = 0
            for key in self.custom_index:
                self.custom_index[key] = 0

        def start_recording(self, path):
            raise RuntimeError("recording needs PyAV and the reference BaseReal (basereal.py:77-131)")

        def stop_recording(self):
            self.recording = False

        def mirror_index(self, size, index):
            turn = index // size
            res = index % size
            return res if turn % 2 == 0 else size - res - 1

        def get_audio_stream(self, audiotype):
            idx = self.custom_audio_index[audiotype]
            stream = self.custom_audio_cycle[audiotype][idx:idx + self.chunk]
            self.custom_audio_index[audiotype] += self.chunk
            if self.custom_audio_index[audiotype] >= self.custom_audio_cycle[audiotype].shape[0]:
                self.curr_state = 1
            return stream

        def set_curr_state(self, audiotype, reinit):
            self.curr_state = audiotype
            if reinit:
                self.custom_audio_index[audiotype] = 0
                self.custom_index[audiotype] = 0
