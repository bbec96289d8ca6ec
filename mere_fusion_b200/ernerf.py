"""ErnerfRenderer: torch tensors in/out around mf_ernerf_render (include/mf_b200.h).

Replaces Trainer.test_gui_with_data + NeRFRenderer.run_cuda/run_torso
(ernerf/nerf_triplane/utils.py:1191-1223, renderer.py:158-352) for one session on one GPU.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import Context, MfErnerfDebug, MfErnerfFrame, check, lib
from .ernerf_pack import pack_ernerf


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class ErnerfRenderer:
    def __init__(self, state_dict=None, mean_density_torso=0.0, opt=None, device=0, blob=None, cfg=None):
        """Either (state_dict, mean_density_torso) or a pre-packed (blob, cfg) -- e.g. a blob
        broadcast from rank 0."""
        self.device = torch.device("cuda", device)
        self.ctx = Context(device)
        if blob is None:
            blob, cfg = pack_ernerf(state_dict, mean_density_torso, opt)
        if isinstance(blob, np.ndarray):
            blob = torch.from_numpy(blob)
        self.blob = blob.to(self.device)          # must outlive the context
        self.cfg = cfg
        check(self.ctx.handle, lib().mf_ernerf_load(self.ctx.handle, _ptr(self.blob), self.blob.numel(),
                                                     ctypes.byref(cfg)), "mf_ernerf_load")

    def reset(self):
        check(self.ctx.handle, lib().mf_ernerf_reset_state(self.ctx.handle), "mf_ernerf_reset_state")

    def encode_audio(self, auds, out=None, stream=None):
        """the audio half of a frame (encode_audio + EMA, advances the session's audio state): cuda fp32 [8, A, 16] -> cuda fp32 [32].
        With render(..., enc_a=that) a rank can render any subset of a session's frames bit-identically to the in-order stream
        (mere_fusion_b200.dist.render_stream_shard)."""
        if out is None:
            out = torch.empty(32, dtype=torch.float32, device=self.device)
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        check(self.ctx.handle, lib().mf_ernerf_encode_audio(self.ctx.handle, _ptr(auds), _ptr(out), ctypes.c_void_p(s.cuda_stream)),
              "mf_ernerf_encode_audio")
        return out

    def profile(self, enable=True):
        check(self.ctx.handle, lib().mf_ernerf_profile(self.ctx.handle, int(enable)), "mf_ernerf_profile")

    def last_head_ms(self, want_ms=True):
        """(duration of the last k_head launch in ms, samples of THIS session's frame it processed); synchronises.  In a batched
        render the launch is timed on the first session only: ask the others with want_ms=False (-> 0.0, samples)."""
        ms, n = ctypes.c_float(), ctypes.c_int64()
        check(self.ctx.handle, lib().mf_ernerf_last_head_ms(self.ctx.handle, ctypes.byref(ms) if want_ms else None, ctypes.byref(n)),
              "mf_ernerf_last_head_ms")
        return ms.value, n.value

    def render_host(self, pose, intrinsics, H, W, auds_pinned, eye, out_pinned, stage=None, **kw):
        """end-to-end call with HOST buffers: pinned fp32 auds in, pinned u8 frame out; both copies are
        enqueued on the current stream around the render (what NeRFReal.test_step does per frame)."""
        if stage is None:
            stage = self._stage = getattr(self, "_stage", None) or {}
        key = (tuple(auds_pinned.shape), tuple(out_pinned.shape))
        if key not in stage:
            stage[key] = (torch.empty(auds_pinned.shape, dtype=torch.float32, device=self.device),
                          torch.empty(out_pinned.shape, dtype=torch.uint8, device=self.device))
        d_auds, d_out = stage[key]
        d_auds.copy_(auds_pinned, non_blocking=True)
        self.render(pose, intrinsics, H, W, d_auds, eye, out=d_out, **kw)
        out_pinned.copy_(d_out, non_blocking=True)
        return out_pinned

    def render_host_async(self, pose, intrinsics, H, W, auds_pinned, eye, outH=None, outW=None, **kw):
        """Pipelined end-to-end call with HOST buffers: pinned fp32 auds in, pinned u8 frame out through a two-slot ring.  The
        device -> host copy of this frame runs on a side stream and overlaps the NEXT call's render; before returning, the current
        stream is made to wait for the PREVIOUS call's copy, so a caller that brackets calls on its stream sees every copy inside
        some call's interval.  Returns (pinned u8 tensor, event): the frame is complete once the event has completed (it is
        overwritten two calls later)."""
        oH, oW = int(outH or H), int(outW or W)
        p = getattr(self, "_pipe", None)
        if p is None or p["shape"] != (oH, oW) or p["auds"].shape != auds_pinned.shape:
            p = self._pipe = dict(shape=(oH, oW), k=0, copy=torch.cuda.Stream(self.device),
                                  auds=torch.empty(auds_pinned.shape, dtype=torch.float32, device=self.device),
                                  dev=[torch.empty((oH, oW, 3), dtype=torch.uint8, device=self.device) for _ in range(2)],
                                  pin=[torch.empty((oH, oW, 3), dtype=torch.uint8).pin_memory() for _ in range(2)],
                                  rendered=[torch.cuda.Event() for _ in range(2)], copied=[torch.cuda.Event() for _ in range(2)], used=[False, False])
        s = p["k"] & 1
        p["k"] += 1
        cur = torch.cuda.current_stream(self.device)
        if p["used"][s]:
            cur.wait_event(p["copied"][s])                 # the slot's previous frame has left the device buffer
        p["auds"].copy_(auds_pinned, non_blocking=True)
        self.render(pose, intrinsics, H, W, p["auds"], eye, out=p["dev"][s], outH=oH, outW=oW, **kw)
        p["rendered"][s].record(cur)
        p["copy"].wait_event(p["rendered"][s])
        with torch.cuda.stream(p["copy"]):
            p["pin"][s].copy_(p["dev"][s], non_blocking=True)
            p["copied"][s].record(p["copy"])
        p["used"][s] = True
        if p["used"][s ^ 1]:
            cur.wait_event(p["copied"][s ^ 1])             # the previous frame's copy, overlapped with this frame's render
        return p["pin"][s], p["copied"][s]

    @property
    def last_launches(self):
        return lib().mf_ernerf_last_launches(self.ctx.handle)

    @staticmethod
    def render_batch(renderers, frames, outs=None, stream=None):
        """One pass for up to 4 sessions of the same avatar model (mf_ernerf_render_batch): renderers[i] renders
        frames[i] = dict(pose, intrinsics, H, W, auds | enc_a, eye[, outH, outW, bg_color]) into outs[i].  Same images
        as renderers[i].render(**frames[i]), bit for bit; the fused head kernel is launched once for the batch."""
        n = len(renderers)
        assert n == len(frames) and n >= 1
        dev = renderers[0].device
        arr = (MfErnerfFrame * n)()
        keep = []
        if outs is None:
            outs = [None] * n
        outs = list(outs)
        for i, (r, f) in enumerate(zip(renderers, frames)):
            pose = np.ascontiguousarray(np.asarray(f["pose"], np.float32).reshape(16))
            keep.append(pose)
            fr = arr[i]
            fr.pose = pose.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
            fr.fx, fr.fy, fr.cx, fr.cy = [float(v) for v in f["intrinsics"]]
            fr.H, fr.W = int(f["H"]), int(f["W"])
            fr.auds = _ptr(f.get("auds"))
            fr.enc_a = _ptr(f.get("enc_a"))
            fr.eye = float(f.get("eye", 0.25))
            fr.bg_color = _ptr(f.get("bg_color"))
            fr.rays_o = fr.rays_d = fr.bg_coords = None
            fr.n_rays = 0
            fr.outH, fr.outW = int(f.get("outH") or f["H"]), int(f.get("outW") or f["W"])
            fr.out_image_f32 = _ptr(f.get("out_f32"))
            if outs[i] is None:
                outs[i] = torch.empty((fr.outH, fr.outW, 3), dtype=torch.uint8, device=dev)
        ctxs = (ctypes.c_void_p * n)(*[r.ctx.handle for r in renderers])
        optr = (ctypes.c_void_p * n)(*[o.data_ptr() for o in outs])
        s = stream if stream is not None else torch.cuda.current_stream(dev)
        rc = lib().mf_ernerf_render_batch(ctxs, arr, optr, n, ctypes.c_void_p(s.cuda_stream))
        check(renderers[0].ctx.handle, rc, "mf_ernerf_render_batch")
        return outs

    def render(self, pose, intrinsics, H, W, auds=None, eye=0.25, out=None, outH=None, outW=None, enc_a=None,
               bg_color=None, rays_o=None, rays_d=None, bg_coords=None, out_f32=None, debug=False, stream=None):
        """pose: 16 floats (host, row-major cam2world); auds: cuda fp32 [8, A, 16];
        returns u8 cuda tensor [outH, outW, 3] RGB (and a dict of debug tensors when debug=True)."""
        pose = np.ascontiguousarray(np.asarray(pose, np.float32).reshape(16))
        fr = MfErnerfFrame()
        fr.pose = pose.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
        fr.fx, fr.fy, fr.cx, fr.cy = [float(v) for v in intrinsics]
        fr.H, fr.W = int(H), int(W)
        fr.auds = _ptr(auds)
        fr.enc_a = _ptr(enc_a)
        fr.eye = float(eye)
        fr.bg_color = _ptr(bg_color)
        explicit = rays_o is not None
        fr.rays_o, fr.rays_d, fr.bg_coords = _ptr(rays_o), _ptr(rays_d), _ptr(bg_coords)
        fr.n_rays = int(rays_o.shape[0]) if explicit else 0
        oH, oW = (1, fr.n_rays) if explicit else (int(outH or H), int(outW or W))
        fr.outH, fr.outW = oH, oW
        if out is None:
            out = torch.empty((oH, oW, 3), dtype=torch.uint8, device=self.device)
        fr.out_image_f32 = _ptr(out_f32)
        dbg = None
        dbg_t = None
        if debug:
            N = fr.n_rays if explicit else H * W
            f32 = dict(dtype=torch.float32, device=self.device)
            dbg_t = dict(nears=torch.zeros(N, **f32), fars=torch.zeros(N, **f32),
                         round_info=torch.zeros((17, 4), dtype=torch.int32, device=self.device),
                         weights_sum=torch.zeros(N, **f32), image_head=torch.zeros((N, 3), **f32),
                         enc_a=torch.zeros(32, **f32), torso_mask=torch.zeros(N, dtype=torch.uint8, device=self.device))
            dbg = MfErnerfDebug(*[_ptr(dbg_t[k]) for k in ("nears", "fars", "round_info", "weights_sum",
                                                           "image_head", "enc_a", "torso_mask")])
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        rc = lib().mf_ernerf_render(self.ctx.handle, ctypes.byref(fr), _ptr(out),
                                    ctypes.byref(dbg) if dbg is not None else None, ctypes.c_void_p(s.cuda_stream))
        check(self.ctx.handle, rc, "mf_ernerf_render")
        return (out, dbg_t) if debug else out
