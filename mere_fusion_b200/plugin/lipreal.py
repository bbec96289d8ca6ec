"""LipReal -- drop-in for /root/reference/lipreal.py:144-250 on the sm_100a Wav2Lip engine.

What changes relative to the reference:
  * `inference()` (lipreal.py:75-141) runs in a thread of this process instead of an
    mp.Process: no pickling of frames through mp.Queue, one CUDA context per GPU;
  * the model call, the batch build (mask / concat / /255) and the x255 are one C-ABI call
    (mf_wav2lip_forward); resize + paste (lipreal.py:207-214) is one more (mf_paste_resize_u8) on
    the device-resident avatar, so the host only receives finished u8 full frames;
  * everything the orchestration sees is unchanged: constructor LipReal(opt), put_msg_txt,
    put_audio_frame, pause_talk, render(quit_event, loop, audio_track, video_track), one
    VideoFrame(bgr24) + exactly two AudioFrame(s16, mono, 16 kHz, 320 samples) per video frame.
"""
import asyncio
import contextlib
import copy
import ctypes
import glob
import os
import pickle
import queue
import time
from queue import Queue
from threading import Event, Thread

import numpy as np

from .basereal import BaseReal
from .frames import AudioFrame, VideoFrame
from .lipasr import LipASR


def mirror_index(size, index):
    """lipreal.py:65-72"""
    turn = index // size
    res = index % size
    return res if turn % 2 == 0 else size - res - 1


class Avatar:
    """coords.pkl + full_imgs/ + face_imgs/ (wav2lip/genavatar.py:101-125), or the same in memory"""

    def __init__(self, frame_list_cycle, face_list_cycle, coord_list_cycle):
        self.frame_list_cycle = frame_list_cycle
        self.face_list_cycle = face_list_cycle
        self.coord_list_cycle = coord_list_cycle

    @classmethod
    def load(cls, avatar_path):
        import cv2

        def read_dir(d):
            lst = glob.glob(os.path.join(d, "*.[jpJP][pnPN]*[gG]"))
            lst = sorted(lst, key=lambda x: int(os.path.splitext(os.path.basename(x))[0]))
            return [cv2.imread(p) for p in lst]

        with open(os.path.join(avatar_path, "coords.pkl"), "rb") as f:
            coords = pickle.load(f)
        return cls(read_dir(os.path.join(avatar_path, "full_imgs")), read_dir(os.path.join(avatar_path, "face_imgs")), coords)


_NOLOCK = contextlib.nullcontext()


RING = 4     # pinned result buffers in flight: res_frame_queue holds <= 2 batches, one more is being written and one frame of a
             # fourth may still be inside process_frames when its slot comes round again


class _Pasted:
    """a full frame that already carries the pasted face (GPU paste path): a view into a pinned ring slot plus the CUDA event
    of the device->host copy that fills it -- the consumer (process_frames) waits for the event, not the inference thread"""
    __slots__ = ("_frame", "_event")

    def __init__(self, frame, event=None):
        self._frame = frame
        self._event = event

    @property
    def frame(self):
        if self._event is not None:
            self._event.synchronize()
            self._event = None
        return self._frame


class LipReal(BaseReal):
    def __init__(self, opt, engine=None, avatar=None, state_dict=None, device=0, paste="gpu", mel="gpu"):
        super().__init__(opt)
        self.W = opt.W
        self.H = opt.H
        self.fps = opt.fps
        self.avatar_id = opt.avatar_id
        self.avatar_path = f"./data/avatars/{self.avatar_id}"
        self.batch_size = opt.batch_size
        self.idx = 0
        self.res_frame_queue = Queue(self.batch_size * 2)
        self.avatar = avatar if avatar is not None else Avatar.load(self.avatar_path)
        self.frame_list_cycle = self.avatar.frame_list_cycle
        self.coord_list_cycle = self.avatar.coord_list_cycle
        self.face_list_cycle = self.avatar.face_list_cycle
        self.device = device
        self.engine = engine if engine is not None else self._load_engine(state_dict)
        self.paste = paste
        self._dev = None
        self.mel_front_end = None
        if mel == "gpu" and hasattr(self.engine, "ctx"):
            from ..wav2lip import MelFrontEnd
            self.mel_front_end = MelFrontEnd(self.engine)
        self.asr = LipASR(opt, self)
        self.asr.warm_up()
        self.render_event = Event()
        self.infer_frames = 0
        self._render_alive = False

    # lipreal.py:42-53: checkpoint["state_dict"], "module." prefix stripped (by the packer)
    def _load_engine(self, state_dict):
        import torch
        from ..wav2lip import Wav2LipEngine
        if state_dict is None:
            ck = torch.load("./models/wav2lip.pth", map_location="cpu")
            state_dict = ck["state_dict"]
        return Wav2LipEngine(state_dict, max_batch=self.batch_size, device=self.device)

    def put_msg_txt(self, msg):
        self.tts.put_msg_txt(msg)

    def put_audio_frame(self, audio_chunk):
        self.asr.put_audio_frame(audio_chunk)

    def pause_talk(self):
        self.tts.pause_talk()
        self.asr.pause_talk()

    # ------------------------------------------------------------------------------------------
    def _device_state(self):
        """avatar resident on the GPU + pinned staging buffers (GPU paste path)"""
        if self._dev is None:
            import torch
            dev = self.engine.device
            if hasattr(self.avatar, "device_tensors"):           # packed avatar (avatar_pack.DeviceAvatar): one upload, device views
                t = self.avatar.device_tensors(dev)
                faces, frames = t["faces"], t["frames"]
            else:
                faces = torch.from_numpy(np.stack(self.face_list_cycle)).to(dev)
                frames = torch.from_numpy(np.stack(self.frame_list_cycle)).to(dev)
            B, S = self.batch_size, faces.shape[1]
            Hf, Wf = frames.shape[1:3]
            self._dev = dict(faces=faces, frames=frames, S=S,
                             mel_pin=torch.empty((B, 1, 80, 16), dtype=torch.float32).pin_memory(),
                             mel=torch.empty((B, 1, 80, 16), dtype=torch.float32, device=dev),
                             sel=torch.empty((B, S, S, 3), dtype=torch.uint8, device=dev),
                             pred=torch.empty((B, S, S, 3), dtype=torch.uint8, device=dev),
                             out=torch.empty((B, Hf, Wf, 3), dtype=torch.uint8, device=dev),
                             out_pin=[torch.empty((B, Hf, Wf, 3), dtype=torch.uint8).pin_memory() for _ in range(RING)] if self.paste == "gpu" else None,
                             pred_pin=torch.empty((B, S, S, 3), dtype=torch.uint8).pin_memory(), slot=0)
        return self._dev

    def infer_batch(self, mel_batch, index):
        """one pass of the hot path for `batch_size` frames starting at avatar index `index`:
        mel chunks (list of [80,16]) -> list of per-frame results for res_frame_queue"""
        import torch
        from .._lib import check, lib
        d = self._device_state()
        B = self.batch_size
        length = len(self.face_list_cycle)
        idxs = [mirror_index(length, index + i) for i in range(B)]
        dev = d["out"].device
        with torch.cuda.device(dev):                             # this thread may never have selected the session's GPU
            s = torch.cuda.current_stream(dev)
            if torch.is_tensor(mel_batch):                       # device-resident chunks from LipASR (GPU mel front-end)
                d["mel"].copy_(mel_batch, non_blocking=True)
            else:
                d["mel_pin"].copy_(torch.from_numpy(np.asarray(mel_batch, dtype=np.float32).reshape(B, 1, 80, 16)))
                d["mel"].copy_(d["mel_pin"], non_blocking=True)
            torch.index_select(d["faces"], 0, torch.as_tensor(idxs, device=dev), out=d["sel"])
            self.engine.forward(d["mel"], d["sel"], out=d["pred"])
            if self.paste == "gpu":
                rows = np.empty((B, 5), np.int32)
                for i, k in enumerate(idxs):
                    y1, y2, x1, x2 = self.coord_list_cycle[k]
                    rows[i] = (k, y1, y2, x1, x2)
                fr = d["frames"]
                with getattr(self.engine, "lock", _NOLOCK):      # a shared engine (scheduler.SharedEngine): one caller at a time on its context
                    check(self.engine.ctx.handle,
                          lib().mf_paste_resize_u8(self.engine.ctx.handle, ctypes.c_void_p(fr.data_ptr()), fr.shape[0], fr.shape[1],
                                                   fr.shape[2], ctypes.c_void_p(d["pred"].data_ptr()), d["S"], B,
                                                   rows.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                                   ctypes.c_void_p(d["out"].data_ptr()), ctypes.c_void_p(s.cuda_stream)),
                          "mf_paste_resize_u8")
                # event + pinned ring: no host wait here and no per-frame copy; process_frames waits on the event of the slot it reads
                pin = d["out_pin"][d["slot"] % RING]
                d["slot"] += 1
                pin.copy_(d["out"], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s)
                full = pin.numpy()
                return [_Pasted(full[i], ev) for i in range(B)]
            d["pred_pin"].copy_(d["pred"], non_blocking=True)
            s.synchronize()
        pred = d["pred_pin"].numpy()
        return [pred[i].copy() for i in range(B)]

    def _put_result(self, item, quit_event):
        """res_frame_queue.put that gives up once the session quits (the consumer thread is gone by then)"""
        while not quit_event.is_set():
            try:
                self.res_frame_queue.put(item, timeout=0.1)
                return
            except queue.Full:
                pass

    def inference(self, quit_event):
        """lipreal.py:75-141, in-process"""
        length = len(self.face_list_cycle)
        index = 0
        count, counttime = 0, 0.0
        while not quit_event.is_set():
            if not self.render_event.is_set():
                time.sleep(0.01)
                continue
            try:
                mel_batch = self.asr.feat_queue.get(block=True, timeout=1)
            except queue.Empty:
                continue
            is_all_silence = True
            audio_frames = []
            for _ in range(self.batch_size * 2):
                frame, type = self.asr.output_queue.get()
                audio_frames.append((frame, type))
                if type == 0:
                    is_all_silence = False
            if is_all_silence:
                for i in range(self.batch_size):
                    self._put_result((None, mirror_index(length, index), audio_frames[i * 2:i * 2 + 2]), quit_event)
                    index = index + 1
            else:
                t = time.perf_counter()
                results = self.infer_batch(mel_batch, index)
                counttime += time.perf_counter() - t
                count += self.batch_size
                self.infer_frames += self.batch_size
                if count >= 100:
                    print(f"------actual avg infer fps:{count / counttime:.4f}")
                    count, counttime = 0, 0.0
                for i, res_frame in enumerate(results):
                    self._put_result((res_frame, mirror_index(length, index), audio_frames[i * 2:i * 2 + 2]), quit_event)
                    index = index + 1
        # the render loop may be blocked in feat_queue.put (depth 2) when the quit event arrives: keep draining until it is out
        while self._render_alive:
            try:
                self.asr.feat_queue.get(timeout=0.05)
            except queue.Empty:
                pass

    def _emit(self, coro, loop):
        if loop is not None:
            asyncio.run_coroutine_threadsafe(coro, loop)
        else:                                      # synchronous fake track (tests / bench)
            try:
                coro.send(None)
            except StopIteration:
                pass

    def process_frames(self, quit_event, loop=None, audio_track=None, video_track=None):
        """lipreal.py:191-230"""
        while not quit_event.is_set():
            try:
                res_frame, idx, audio_frames = self.res_frame_queue.get(block=True, timeout=1)
            except queue.Empty:
                continue
            if audio_frames[0][1] != 0 and audio_frames[1][1] != 0:      # both chunks non-speech: idle frame
                audiotype = audio_frames[0][1]
                if self.custom_index.get(audiotype) is not None:
                    mirindex = self.mirror_index(len(self.custom_img_cycle[audiotype]), self.custom_index[audiotype])
                    combine_frame = self.custom_img_cycle[audiotype][mirindex]
                    self.custom_index[audiotype] += 1
                else:
                    combine_frame = self.frame_list_cycle[idx]
            elif isinstance(res_frame, _Pasted):
                combine_frame = res_frame.frame
            else:
                import cv2
                bbox = self.coord_list_cycle[idx]
                combine_frame = copy.deepcopy(self.frame_list_cycle[idx])
                y1, y2, x1, x2 = bbox
                try:
                    res_frame = cv2.resize(res_frame.astype(np.uint8), (x2 - x1, y2 - y1))
                except Exception:                  # noqa: BLE001 (same recovery as the reference)
                    continue
                combine_frame[y1:y2, x1:x2] = res_frame
            new_frame = VideoFrame.from_ndarray(combine_frame, format="bgr24")
            self._emit(video_track._queue.put(new_frame), loop)
            if self.recording:
                self.recordq_video.put(new_frame)
            for audio_frame in audio_frames:
                frame, type = audio_frame
                frame = (frame * 32767).astype(np.int16)
                new_frame = AudioFrame(format="s16", layout="mono", samples=frame.shape[0])
                new_frame.planes[0].update(frame.tobytes())
                new_frame.sample_rate = 16000
                self._emit(audio_track._queue.put(new_frame), loop)
                if self.recording:
                    self.recordq_audio.put(new_frame)

    def render(self, quit_event, loop=None, audio_track=None, video_track=None):
        """lipreal.py:232-250"""
        self.tts.render(quit_event)
        self.init_customindex()
        self._render_alive = True
        process_thread = Thread(target=self.process_frames, args=(quit_event, loop, audio_track, video_track), daemon=True)
        process_thread.start()
        infer_thread = Thread(target=self.inference, args=(quit_event,), daemon=True)
        infer_thread.start()
        self.render_event.set()
        while not quit_event.is_set():
            self.asr.run_step()
            if video_track._queue.qsize() >= 5:
                time.sleep(0.04 * video_track._queue.qsize() * 0.8)
        self.render_event.clear()
        self._render_alive = False
        process_thread.join()
        infer_thread.join()
