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

    def _numbered_images(directory):
        """the images of a directory in the order of their integer file names"""
        import cv2
        paths = glob.glob(os.path.join(directory, "*.[jpJP][pnPN]*[gG]"))
        return [cv2.imread(p) for p in sorted(paths, key=lambda p: int(os.path.splitext(os.path.basename(p))[0]))]

    class BaseReal:
        """stand-in with the attributes the plugin classes and the orchestration read (`opt, tts, curr_state, recording, recordq_*,
        custom_*` dictionaries keyed by audiotype) and the pure-logic methods of basereal.py:59-75,133-154"""

        def __init__(self, opt):
            self.opt = opt
            self.sample_rate = 16000
            self.chunk = self.sample_rate // opt.fps
            self.tts = _NoTTS()
            self.recording = False
            self.recordq_video, self.recordq_audio = Queue(), Queue()
            self.curr_state = 0                      # 0 speech / 1 silence / > 1: a custom idle clip is playing
            self.custom_img_cycle, self.custom_audio_cycle = {}, {}
            self.custom_audio_index, self.custom_index, self.custom_opt = {}, {}, {}
            for item in getattr(opt, "customopt", None) or ():
                self._add_custom(item)

        def _add_custom(self, item):
            """one entry of the custom-video config: frames from imgpath, mono float32 audio from audiopath, both rewound"""
            import soundfile as sf
            kind = item["audiotype"]
            self.custom_img_cycle[kind] = _numbered_images(item["imgpath"])
            self.custom_audio_cycle[kind] = sf.read(item["audiopath"], dtype="float32")[0]
            self.custom_audio_index[kind] = self.custom_index[kind] = 0
            self.custom_opt[kind] = item

        def init_customindex(self):
            self.curr_state = 0
            for positions in (self.custom_audio_index, self.custom_index):
                positions.update(dict.fromkeys(positions, 0))

        def start_recording(self, path):
            raise RuntimeError("recording needs PyAV and the reference BaseReal (basereal.py:77-131)")

        def stop_recording(self):
            self.recording = False

        def mirror_index(self, size, index):
            """ping-pong over `size` items: 0 .. size-1, size-1 .. 0, 0 .. (basereal.py:133-139)"""
            lap, pos = divmod(index, size)
            return pos if lap % 2 == 0 else size - 1 - pos

        def get_audio_stream(self, audiotype):
            """the next chunk of the custom clip; at its end the state falls back to silence"""
            clip, at = self.custom_audio_cycle[audiotype], self.custom_audio_index[audiotype]
            self.custom_audio_index[audiotype] = at + self.chunk
            if at + self.chunk >= clip.shape[0]:
                self.curr_state = 1
            return clip[at:at + self.chunk]

        def set_curr_state(self, audiotype, reinit):
            self.curr_state = audiotype
            if reinit:
                self.custom_audio_index[audiotype] = self.custom_index[audiotype] = 0
