"""MuseReal -- drop-in for /root/reference/musereal.py:124-290 on the sm_100a MuseTalk engine.

What changes relative to the reference:
  * `inference()` (musereal.py:52-120) runs in a thread of this process instead of an mp.Process;
  * Whisper features: one C-ABI call per batch on the GPU (mf_whisper_features) behind the unchanged
    Audio2Feature.audio2feat / feature2chunks interface;
  * pe(...) -> unet.model(...).sample -> vae.decode_latents(...) is one C-ABI call (mf_musetalk_forward) on the
    device-resident latent cycle; resize + mask blending (musereal.py:240-248, blending.py:103-125) is one more
    (mf_paste_blend_u8, bit-exact vs cv2) on the device-resident avatar frames and masks, so the host only
    receives finished u8 full frames;
  * everything the orchestration sees is unchanged: MuseReal(opt), put_msg_txt, put_audio_frame, pause_talk,
    render(quit_event, loop, audio_track, video_track), one VideoFrame(bgr24) + exactly two AudioFrame(s16, mono,
    16 kHz, 320 samples) per video frame.
"""
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
from .lipreal import _NOLOCK, RING, LipReal, _Pasted, mirror_index
from .museasr import MuseASR


class MuseAvatar:
    """full_imgs/ + coords.pkl + latents.pt + mask/ + mask_coords.pkl (musetalk/mere_musetalk.py:250-317), or in memory.
    coords are (x1, y1, x2, y2); mask crop boxes are (x_s, y_s, x_e, y_e); masks are BGR u8 of the crop-box size."""

    def __init__(self, frame_list_cycle, coord_list_cycle, input_latent_list_cycle, mask_list_cycle, mask_coords_list_cycle):
        self.frame_list_cycle = frame_list_cycle
        self.coord_list_cycle = coord_list_cycle
        self.input_latent_list_cycle = input_latent_list_cycle
        self.mask_list_cycle = mask_list_cycle
        self.mask_coords_list_cycle = mask_coords_list_cycle

    @classmethod
    def load(cls, avatar_path):
        import cv2
        import torch

        def read_dir(d):
            lst = glob.glob(os.path.join(d, "*.[jpJP][pnPN]*[gG]"))
            lst = sorted(lst, key=lambda x: int(os.path.splitext(os.path.basename(x))[0]))
            return [cv2.imread(p) for p in lst]

        with open(os.path.join(avatar_path, "coords.pkl"), "rb") as f:
            coords = pickle.load(f)
        with open(os.path.join(avatar_path, "mask_coords.pkl"), "rb") as f:
            mask_coords = pickle.load(f)
        latents = torch.load(os.path.join(avatar_path, "latents.pt"), map_location="cpu")
        return cls(read_dir(os.path.join(avatar_path, "full_imgs")), coords, latents, read_dir(os.path.join(avatar_path, "mask")),
                   mask_coords)


class MuseReal(BaseReal):
    def __init__(self, opt, engine=None, audio_processor=None, avatar=None, unet_sd=None, vae_sd=None, whisper_sd=None,
                 device=0, paste="gpu"):
        super().__init__(opt)
        self.W = opt.W
        self.H = opt.H
        self.fps = opt.fps
        self.avatar_id = opt.avatar_id
        self.bbox_shift = getattr(opt, "bbox_shift", 0)
        self.avatar_path = f"./data/avatars/{self.avatar_id}"
        self.batch_size = opt.batch_size
        self.idx = 0
        self.res_frame_queue = Queue(self.batch_size * 2)
        self.device = device
        self.avatar = avatar if avatar is not None else MuseAvatar.load(self.avatar_path)
        self.frame_list_cycle = self.avatar.frame_list_cycle
        self.coord_list_cycle = self.avatar.coord_list_cycle
        self.input_latent_list_cycle = self.avatar.input_latent_list_cycle
        self.mask_list_cycle = self.avatar.mask_list_cycle
        self.mask_coords_list_cycle = self.avatar.mask_coords_list_cycle
        self.engine = engine if engine is not None else self._load_engine(unet_sd, vae_sd)
        self.audio_processor = audio_processor if audio_processor is not None else self._load_audio_model(whisper_sd)
        self.paste = paste
        self._dev = None
        self.asr = MuseASR(opt, self, self.audio_processor)
        self.asr.warm_up()
        self.render_event = Event()
        self.infer_frames = 0
        self._render_alive = False

    # musetalk/utils/utils.py:70-75 load_all_model: diffusers-format state dicts
    def _load_engine(self, unet_sd, vae_sd):
        import torch
        from ..musetalk import MuseTalkEngine
        if unet_sd is None:
            unet_sd = torch.load("./models/musetalk/pytorch_model.bin", map_location="cpu")
        if vae_sd is None:
            vae_sd = torch.load("./models/sd-vae-ft-mse/diffusion_pytorch_model.bin", map_location="cpu")
        return MuseTalkEngine(unet_sd, vae_sd, max_batch=self.batch_size, device=self.device)

    def _load_audio_model(self, whisper_sd):
        from ..whisper import Audio2Feature
        return Audio2Feature(model_path="./models/whisper/tiny.pt", state_dict=whisper_sd, device=self.device)

    def put_msg_txt(self, msg):
        self.tts.put_msg_txt(msg)

    def put_audio_frame(self, audio_chunk):
        self.asr.put_audio_frame(audio_chunk)

    def pause_talk(self):
        self.tts.pause_talk()
        self.asr.pause_talk()

    # ------------------------------------------------------------------------------------------
    def _device_state(self):
        """avatar (frames, latents, masks) resident on the GPU + pinned staging buffers"""
        if self._dev is None:
            import torch
            dev = self.engine.device
            if hasattr(self.avatar, "device_tensors"):           # packed avatar (avatar_pack.DeviceAvatar): one upload, device views
                t = self.avatar.device_tensors(dev)
                frames, lat, masks, offs = t["frames"], t["latents"], t["masks"], t["mask_off"]
            else:
                frames, lat, masks, offs = self._upload_avatar(dev)
            B = self.batch_size
            Hf, Wf = frames.shape[1:3]
            self._dev = dict(frames=frames, latents=lat, masks=masks, mask_off=offs,
                             wh_pin=torch.empty((B, 50, 384), dtype=torch.float16).pin_memory(),
                             wh=torch.empty((B, 50, 384), dtype=torch.float16, device=dev),
                             sel=torch.empty((B,) + tuple(lat.shape[1:]), dtype=torch.float16, device=dev),
                             pred=torch.empty((B, 256, 256, 3), dtype=torch.uint8, device=dev),
                             out=torch.empty((B, Hf, Wf, 3), dtype=torch.uint8, device=dev),
                             out_pin=[torch.empty((B, Hf, Wf, 3), dtype=torch.uint8).pin_memory() for _ in range(RING)] if self.paste == "gpu" else None,
                             pred_pin=torch.empty((B, 256, 256, 3), dtype=torch.uint8).pin_memory(), slot=0)
        return self._dev

    def _upload_avatar(self, dev):
        import torch
        frames = torch.from_numpy(np.stack(self.frame_list_cycle)).to(dev)
        lat = torch.cat([torch.as_tensor(np.asarray(l)) if not torch.is_tensor(l) else l for l in self.input_latent_list_cycle], dim=0)
        lat = lat.to(device=dev, dtype=torch.float16).contiguous()              # [n, 8, 32, 32] (musereal.py:103)
        offs, parts, off = [], [], 0
        for m, (xs, ys, xe, ye) in zip(self.mask_list_cycle, self.mask_coords_list_cycle):
            m = np.ascontiguousarray(m, np.uint8)
            if m.ndim == 2:
                m = np.repeat(m[:, :, None], 3, axis=2)
            assert m.shape == (ye - ys, xe - xs, 3), "mask size must equal its crop box (blending.py:109-121)"
            offs.append(off)
            parts.append(m.reshape(-1))
            off += m.size
        masks = torch.from_numpy(np.concatenate(parts)).to(dev)
        return frames, lat, masks, offs

    def infer_batch(self, whisper_chunks, index):
        """one pass of the hot path for `batch_size` frames starting at avatar index `index`"""
        import torch
        from .._lib import check, lib
        d = self._device_state()
        B = self.batch_size
        length = len(self.input_latent_list_cycle)
        idxs = [mirror_index(length, index + i) for i in range(B)]
        dev = d["out"].device
        with torch.cuda.device(dev):                              # this thread may never have selected the session's GPU
            s = torch.cuda.current_stream(dev)
            if torch.is_tensor(whisper_chunks):                   # device-resident chunks from MuseASR (already fp16)
                d["wh"].copy_(whisper_chunks, non_blocking=True)
            else:
                d["wh_pin"].copy_(torch.from_numpy(np.stack(whisper_chunks).astype(np.float16)))  # .to(dtype=half), musereal.py:99-101
                d["wh"].copy_(d["wh_pin"], non_blocking=True)
            torch.index_select(d["latents"], 0, torch.as_tensor(idxs, device=dev), out=d["sel"])
            self.engine.forward(d["sel"], d["wh"], out=d["pred"])
            if self.paste == "gpu":
                rows = np.empty((B, 9), np.int32)
                moff = np.empty(B, np.int64)
                for i, k in enumerate(idxs):
                    x1, y1, x2, y2 = self.coord_list_cycle[k]
                    xs, ys, xe, ye = self.mask_coords_list_cycle[k]
                    rows[i] = (k, y1, y2, x1, x2, ys, ye, xs, xe)
                    moff[i] = d["mask_off"][k]
                fr = d["frames"]
                with getattr(self.engine, "lock", _NOLOCK):       # a shared engine (scheduler.SharedEngine): one caller at a time on its context
                    check(self.engine.ctx.handle,
                          lib().mf_paste_blend_u8(self.engine.ctx.handle, ctypes.c_void_p(fr.data_ptr()), fr.shape[0], fr.shape[1], fr.shape[2],
                                                  ctypes.c_void_p(d["pred"].data_ptr()), 256, B,
                                                  rows.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), ctypes.c_void_p(d["masks"].data_ptr()),
                                                  d["masks"].numel(), moff.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                                  ctypes.c_void_p(d["out"].data_ptr()), ctypes.c_void_p(s.cuda_stream)),
                          "mf_paste_blend_u8")
                # event + pinned ring (plugin/lipreal.py): process_frames waits on the slot's event, no per-frame host copy
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

    def inference(self, quit_event):
        """musereal.py:52-120, in-process"""
        length = len(self.input_latent_list_cycle)
        index = 0
        count, counttime = 0, 0.0
        while not quit_event.is_set():
            if not self.render_event.is_set():
                time.sleep(0.01)
                continue
            try:
                whisper_chunks = self.asr.feat_queue.get(block=True, timeout=1)
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
                results = self.infer_batch(whisper_chunks, index)
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

    _emit = LipReal._emit
    _put_result = LipReal._put_result

    def process_frames(self, quit_event, loop=None, audio_track=None, video_track=None):
        """musereal.py:222-264"""
        while not quit_event.is_set():
            try:
                res_frame, idx, audio_frames = self.res_frame_queue.get(block=True, timeout=1)
            except queue.Empty:
                continue
            if audio_frames[0][1] != 0 and audio_frames[1][1] != 0:      # both chunks non-speech: full image only
                audiotype = audio_frames[0][1]
                if self.custom_index.get(audiotype) is not None:
                    mirindex = self.mirror_index(len(self.custom_img_cycle[audiotype]), self.custom_index[audiotype])
                    combine_frame = self.custom_img_cycle[audiotype][mirindex]
                    self.custom_index[audiotype] += 1
                else:
                    combine_frame = self.frame_list_cycle[idx]
            elif isinstance(res_frame, _Pasted):
                combine_frame = res_frame.frame
            else:                                                         # host blend, as the reference does it
                import cv2
                bbox = self.coord_list_cycle[idx]
                ori_frame = copy.deepcopy(self.frame_list_cycle[idx])
                x1, y1, x2, y2 = bbox
                try:
                    res_frame = cv2.resize(res_frame.astype(np.uint8), (x2 - x1, y2 - y1))
                except Exception:                                         # noqa: BLE001 (same recovery as the reference)
                    continue
                combine_frame = get_image_blending(ori_frame, res_frame, bbox, self.mask_list_cycle[idx], self.mask_coords_list_cycle[idx])
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
        """musereal.py:267-290"""
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
            if video_track._queue.qsize() >= 1.5 * self.opt.batch_size:
                time.sleep(0.04 * video_track._queue.qsize() * 0.8)
        self.render_event.clear()
        self._render_alive = False
        process_thread.join()
        infer_thread.join()


def get_image_blending(image, face, face_box, mask_array, crop_box):
    """musetalk/utils/blending.py:103-125 on the host with cv2 (paste="cpu" mode: the check for the GPU kernel)"""
    import cv2
    body = image
    x, y, x1, y1 = face_box
    x_s, y_s, x_e, y_e = crop_box
    face_large = copy.deepcopy(body[y_s:y_e, x_s:x_e])
    face_large[y - y_s:y1 - y_s, x - x_s:x1 - x_s] = face
    mask_image = cv2.cvtColor(mask_array, cv2.COLOR_BGR2GRAY)
    mask_image = (mask_image / 255).astype(np.float32)
    body[y_s:y_e, x_s:x_e] = cv2.blendLinear(face_large, body[y_s:y_e, x_s:x_e], mask_image, 1 - mask_image)
    return body
