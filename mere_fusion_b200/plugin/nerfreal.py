"""NeRFReal -- drop-in for /root/reference/nerfreal.py:34-156 on the fused sm_100a ErNeRF renderer.

Reference:  NeRFReal(opt, trainer, data_loader); test_step = next(loader) [get_rays on the host side,
~15 torch ops] -> asr.get_next_feat -> trainer.test_gui_with_data [~400 launches, <= 16 host syncs]
-> (image * 255).astype(uint8) -> VideoFrame(rgb24).
Here:       `trainer` is an ErnerfRenderer (or a reference Trainer, from whose model the checkpoint is
taken), `data_loader` an ErnerfPoseProvider (or the reference loader, whose poses / eye areas /
intrinsics are taken over); test_step is one mf_ernerf_render call + one D2H of the u8 frame.
"""
import asyncio
import time

import numpy as np

from .basereal import BaseReal
from .frames import AudioFrame, VideoFrame
from .nerfasr import NerfASR


def _adapt_provider(data_loader):
    from ..ernerf_data import ErnerfPoseProvider
    if isinstance(data_loader, ErnerfPoseProvider):
        return data_loader
    ds = getattr(data_loader, "_data", None)            # reference NeRFDataset_Test behind a torch DataLoader
    if ds is None:
        raise TypeError("data_loader must be an ErnerfPoseProvider or the reference test loader")
    prov = ErnerfPoseProvider.__new__(ErnerfPoseProvider)
    prov.poses = ds.poses.detach().cpu().numpy().astype(np.float32)
    prov.eye_area = ds.eye_area.detach().cpu().numpy().reshape(-1) if getattr(ds, "eye_area", None) is not None else None
    prov.intrinsics = np.asarray(ds.intrinsics)
    prov.H, prov.W, prov.index = ds.H, ds.W, 0
    # background (provider.py:203-214,235-238,330-332): 'white' / 'black' / an image file, [H,W,3] in [0,1]
    if getattr(getattr(ds, "opt", None), "torso_imgs", "") != "":
        raise NotImplementedError("opt.torso_imgs (per-frame torso composites as background, provider.py:316-328) is not supported by "
                                  "the fused renderer: run with the torso model (torso_imgs='')")
    bg = getattr(ds, "bg_img", None)
    if bg is not None:
        bg = np.asarray(bg.detach().float().cpu().numpy() if hasattr(bg, "detach") else bg, np.float32).reshape(ds.H, ds.W, 3)
        if np.all(bg == 1.0):
            bg = None                                   # the kernel's default
    prov.bg_img = bg
    return prov


def _adapt_renderer(trainer, opt, device):
    from ..ernerf import ErnerfRenderer
    if isinstance(trainer, ErnerfRenderer) or isinstance(getattr(trainer, "renderer", None), ErnerfRenderer):
        return trainer                                  # a renderer, or the scheduler's batched proxy around one (scheduler.ErnerfBatcher)
    model = getattr(trainer, "model", None)             # reference Trainer: take the loaded weights over
    if model is None:
        raise TypeError("trainer must be an ErnerfRenderer or the reference Trainer")
    o = dict(bound=opt.bound, min_near=opt.min_near, dt_gamma=opt.dt_gamma, max_steps=opt.max_steps,
             density_thresh_torso=opt.density_thresh_torso, torso_shrink=opt.torso_shrink, smooth_lips=opt.smooth_lips)
    return ErnerfRenderer(model.state_dict(), float(getattr(model, "mean_density_torso", 0.0)), o, device=device)


class NeRFReal(BaseReal):
    def __init__(self, opt, trainer, data_loader, debug=True, feature_fn=None, device=0):
        super().__init__(opt)
        self.W = opt.W
        self.H = opt.H
        self.provider = _adapt_provider(data_loader)
        self.renderer = _adapt_renderer(trainer, opt, device)
        self.trainer = trainer
        self.data_loader = data_loader
        if getattr(opt, "fullbody", False):
            import glob
            import os
            import cv2
            lst = glob.glob(os.path.join(self.opt.fullbody_img, "*.[jpJP][pnPN]*[gG]"))
            lst = sorted(lst, key=lambda x: int(os.path.splitext(os.path.basename(x))[0]))
            self.fullbody_list_cycle = [cv2.imread(p) for p in lst[:len(self.provider)]]
        self.asr = NerfASR(opt, self, feature_fn=feature_fn, device=f"cuda:{device}" if self._has_cuda() else "cpu")
        self.asr.warm_up()
        self._out = None                                # two-slot ring of device / pinned frames (_render_async)
        self._pin = None
        self._bg = None

    @staticmethod
    def _has_cuda():
        import torch
        return torch.cuda.is_available()

    def put_msg_txt(self, msg):
        self.tts.put_msg_txt(msg)

    def put_audio_frame(self, audio_chunk):
        self.asr.put_audio_frame(audio_chunk)

    def pause_talk(self):
        self.tts.pause_talk()
        self.asr.pause_talk()

    def _emit(self, coro, loop):
        if loop is not None:
            asyncio.run_coroutine_threadsafe(coro, loop)
        else:
            try:
                coro.send(None)
            except StopIteration:
                pass

    def _render_async(self, pose, eye, auds):
        """Trainer.test_gui_with_data (utils.py:1191-1223) + nerfreal.py:110, enqueued only: the frame is rendered into one of two device
        buffers and copied to its pinned twin on a copy stream; returns (pinned u8 tensor, event that completes with the copy)."""
        import torch
        dev = self.renderer.device
        if self._out is None:
            self._out = [torch.empty((self.H, self.W, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
            self._pin = [torch.empty((self.H, self.W, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
            self._done = [torch.cuda.Event() for _ in range(2)]
            self._drawn = [torch.cuda.Event() for _ in range(2)]
            self._copy = torch.cuda.Stream(dev)
            self._slot = 0
        p = self.provider
        fix_eye = getattr(self.opt, "fix_eye", -1)
        eye = fix_eye if (self.opt.exp_eye and fix_eye >= 0) else eye            # utils.py:937-940
        if self._bg is None and getattr(p, "bg_img", None) is not None:        # data['bg_color'] (provider.py:330-332), uploaded once
            self._bg = torch.from_numpy(np.ascontiguousarray(p.bg_img, np.float32).reshape(-1, 3)).to(dev, torch.float16)
        k = self._slot
        self._slot ^= 1
        with torch.cuda.device(dev):                     # this thread may never have selected the session's GPU
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(self._done[k])                # the slot's previous frame has left the device buffer (no-op the first time)
            self.renderer.render(pose, p.intrinsics, p.H, p.W, auds.contiguous(), eye if eye is not None else 0.0,
                                 out=self._out[k], outH=self.H, outW=self.W, bg_color=self._bg)
            self._drawn[k].record(cur)
            self._copy.wait_event(self._drawn[k])
            with torch.cuda.stream(self._copy):
                self._pin[k].copy_(self._out[k], non_blocking=True)
                self._done[k].record(self._copy)
        return self._pin[k], self._done[k]

    def render_image(self, pose, eye, auds):
        """one frame, synchronously: [H, W, 3] u8 RGB ndarray (a view of a pinned buffer that is reused two frames later)"""
        pin, done = self._render_async(pose, eye, auds)
        done.synchronize()
        return pin.numpy()

    def _step(self, loop, audio_track):
        """nerfreal.py:70-108: one step up to the image -- the two audio frames are emitted, the video frame is returned as a handle
        (kind, payload, provider index): ("image", ndarray) or ("pending", (pinned tensor, event))."""
        index, pose, eye = next(self.provider)
        auds = self.asr.get_next_feat()
        audiotype1 = audiotype2 = 0
        for i in range(2):
            frame, type = self.asr.get_audio_out()
            if i == 0:
                audiotype1 = type
            else:
                audiotype2 = type
            frame = (frame * 32767).astype(np.int16)
            new_frame = AudioFrame(format="s16", layout="mono", samples=frame.shape[0])
            new_frame.planes[0].update(frame.tobytes())
            new_frame.sample_rate = 16000
            self._emit(audio_track._queue.put(new_frame), loop)
        if audiotype1 != 0 and audiotype2 != 0 and self.custom_index.get(audiotype1) is not None:
            import cv2
            mirindex = self.mirror_index(len(self.custom_img_cycle[audiotype1]), self.custom_index[audiotype1])
            image = cv2.cvtColor(self.custom_img_cycle[audiotype1][mirindex], cv2.COLOR_BGR2RGB)
            self.custom_index[audiotype1] += 1
            return ("image", image, index)
        return ("pending", self._render_async(pose, eye, auds), index)

    def _finish(self, handle, loop, video_track):
        """nerfreal.py:108-127: wait for the frame if it is still in flight, full-body composite, VideoFrame, video queue"""
        kind, payload, index = handle
        if kind == "pending":
            pin, done = payload
            done.synchronize()
            image = pin.numpy()                          # VideoFrame.from_ndarray copies (PyAV semantics): the pinned slot is free again
            if getattr(self.opt, "fullbody", False):
                import cv2
                image_fullbody = cv2.cvtColor(self.fullbody_list_cycle[index], cv2.COLOR_BGR2RGB)
                sx, sy = self.opt.fullbody_offset_x, self.opt.fullbody_offset_y
                image_fullbody[sy:sy + image.shape[0], sx:sx + image.shape[1]] = image
                image = image_fullbody
        else:
            image = payload
        new_frame = VideoFrame.from_ndarray(image, format="rgb24")
        self._emit(video_track._queue.put(new_frame), loop)

    def test_step(self, loop=None, audio_track=None, video_track=None):
        """nerfreal.py:70-127, one step start to finish (the reference's own granularity)"""
        self._finish(self._step(loop, audio_track), loop, video_track)

    def render(self, quit_event, loop=None, audio_track=None, video_track=None):
        """nerfreal.py:129-156 (webrtc transport).  Same steps in the same order; the video frame of step k is handed to the track while
        step k + 1 is already on the GPU -- but only while more audio is waiting: with an empty input queue (a live 25 fps session, a
        closed-loop client) the frame is finished at once, so nothing is ever held back behind the queue's 10 ms time-out."""
        self.init_customindex()
        count, totaltime = 0, 0.0
        self.tts.render(quit_event)
        pending = None
        while not quit_event.is_set():
            t = time.perf_counter()
            for _ in range(2):
                self.asr.run_step()
            handle = self._step(loop, audio_track)
            if pending is not None:
                self._finish(pending, loop, video_track)
            pending = handle
            if self.asr.queue.empty():
                self._finish(pending, loop, video_track)
                pending = None
            totaltime += time.perf_counter() - t
            count += 1
            if count == 100:
                print(f"------actual avg infer fps:{count / totaltime:.4f}")
                count, totaltime = 0, 0.0
            if video_track._queue.qsize() >= 5:
                time.sleep(0.04 * video_track._queue.qsize() * 0.8)
        if pending is not None:
            self._finish(pending, loop, video_track)
