#!/usr/bin/env python
"""bench.py -- face frames/sec at 512x512 through the B200-native hot path.

  python bench.py --gpus N --steps K --warmup W [--workload ernerf] [--impl reference]

Contract (one JSON line on rank 0): see the repository task statement.  A "step" is one pass of
the hot path over one batch: for ErNeRF one 512x512 frame (262 144 rays marched, shaded,
composited, torso-blended, written as u8 RGB).

Timing: W untimed warm-up steps, then K timed steps.  Every timed step is bracketed by CUDA
events on the launching stream; between steps the L2 is flushed by writing a 256 MiB buffer
(outside the events) -- the grid tables (7 MB) would otherwise stay L2-resident.  The K-step
region as a whole is bracketed by barrier + synchronize; the per-rank time is the sum of the
per-step device times and the reported time is the MAX over ranks.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "face frames/sec at 512x512"
H = W = 512
RAYS = H * W
BYTES_PER_SAMPLE = 576      # 3 planes x 12 levels x 4 corners x 4 B gathered per sample
FLOP_PER_SAMPLE = 46368     # 23 184 MAC


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf=d["bf16_tflops"], tf_sus=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf=1590.0, tf_sus=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """samples SM clocks / throttle reasons through NVML during the timed region"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_ev = threading.Event()
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self.stop_ev.is_set():
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def result(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def oracle_sample_fps(n_side, frames=1):
    """CPU port (oracle) on a bounded sample: an n_side x n_side sub-grid of the 512x512 ray grid,
    full pipeline; returns (equivalent full frames/s, seconds per sample, cores)."""
    from helpers import ernerf_inputs, load_ernerf_fixture
    from oracle.ernerf_oracle import ErnerfOracle
    sd, md = load_ernerf_fixture()
    orc = ErnerfOracle(sd, md)
    ts = []
    for f in range(frames):
        pose, intr, auds, eye = ernerf_inputs(f, n_side, n_side)
        t0 = time.perf_counter()
        orc.render_frame(pose, intr, n_side, n_side, auds, eye)
        ts.append(time.perf_counter() - t0)
    t = float(np.median(ts))
    return (n_side * n_side / RAYS) / t, t, os.cpu_count()


def config1_cpu_plumbing(seconds=10.0):
    """BASELINE configs[0] / SURVEY 8(d) config 1: the Wav2Lip 96x96 plugin path on the HOST only -- LipASR windows + numpy mel, the fp32
    restatement of the reference generator (oracle, all host cores), cv2 resize + paste, fake tracks -- for a 10 s clip (500 chunks ->
    250 frames).  A reported CPU baseline (kind "port": librosa / PyAV are absent, the mirrors of lipreal.py / lipasr.py stand in); it
    never touches the GPU library."""
    import threading
    import torch
    from helpers import seeded_wav2lip_state, synthetic_speech
    from oracle import wav2lip_oracle as O
    from test_plugin_cpu import FakeTrack, _fake_avatar, make_opt
    from mere_fusion_b200.plugin.lipreal import LipReal, mirror_index
    torch.set_num_threads(os.cpu_count())
    sd = seeded_wav2lip_state(2)

    class HostLipReal(LipReal):
        def infer_batch(self, mel_batch, index):
            n = len(self.face_list_cycle)
            faces = np.stack([self.face_list_cycle[mirror_index(n, index + i)] for i in range(self.batch_size)])
            mel = np.asarray(mel_batch, np.float32).reshape(self.batch_size, 1, 80, 16)
            pred, _ = O.infer(sd, mel, faces)
            return [pred[i] * 255. for i in range(self.batch_size)]          # lipreal.py:126

    real = HostLipReal(make_opt(), engine=object(), avatar=_fake_avatar(), paste="cpu", mel="host")
    n_chunks = int(seconds * 50)
    wav = synthetic_speech(n_chunks * 320, 0)
    for i in range(n_chunks):
        real.put_audio_frame(wav[i * 320:(i + 1) * 320])
    quit_event = threading.Event()
    vt, at = FakeTrack(), FakeTrack()
    want = n_chunks // 2
    t0 = time.perf_counter()
    th = threading.Thread(target=real.render, args=(quit_event, None, at, vt), daemon=True)
    th.start()
    while len(vt._queue.items) < want and time.perf_counter() - t0 < 300:
        time.sleep(0.005)
    dt = time.perf_counter() - t0
    quit_event.set()
    th.join(timeout=30)
    nv, na = len(vt._queue.items), len(at._queue.items)
    return {"value": want / dt, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{seconds:.0f} s clip ({n_chunks} chunks -> {want} frames at 512x512) through LipReal.render on the host: numpy mel, fp32 oracle of the "
                      f"reference generator, cv2 paste, fake tracks; {dt:.2f} s, {nv} video / {na} audio frames emitted "
                      "(BASELINE configs[0]; the fake video track never fills, so no back-pressure sleep is taken)"}


def run_reference(args):
    os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count() or 1))
    return _run_reference(args)


def _run_reference(args):
    """--impl reference: the reference has NO CPU renderer for ErNeRF (renderer.py:664 always
    dispatches to run_cuda), so the CPU arm is the oracle port (kind "port") on all host cores,
    each step a bounded sub-grid sample of the 512x512 workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from helpers import ernerf_inputs, load_ernerf_fixture
    from oracle.ernerf_oracle import ErnerfOracle
    sd, md = load_ernerf_fixture()
    orc = ErnerfOracle(sd, md)
    # a CONSTANT sample (the same 96x96 sub-grid of the 512x512 ray grid every run, all host cores): the sample used to be sized
    # from a first-estimate rate, which made this arm drift 2x between boxes
    n_side = 96
    os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count()))
    for w in range(args.warmup):
        orc.render_frame(*ernerf_inputs(w, n_side, n_side)[:2], n_side, n_side, *ernerf_inputs(w, n_side, n_side)[2:])
    t0 = time.perf_counter()
    for k in range(args.steps):
        pose, intr, auds, eye = ernerf_inputs(k, n_side, n_side)
        orc.render_frame(pose, intr, n_side, n_side, auds, eye)
    dt = time.perf_counter() - t0
    fps = (args.steps * n_side * n_side / RAYS) / dt
    sample = f"{n_side}x{n_side} sub-grid of the 512x512 ray grid per step ({n_side * n_side} of {RAYS} rays), full pipeline, scaled to full frames"
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
            "config": {"workload": "ernerf_512x512_fullframe", "note": "reference has no CPU renderer; oracle port"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=sys.__stdout__, flush=True)


def reference_cuda_leg(dev, n_frames=24, warmup=4):
    """SURVEY 8(d) config 4 baseline (2), "the number to beat": the REFERENCE'S OWN CUDA render on this same GPU -- the unmodified
    reference Python (oracle/_ref/py, staged by oracle/build_ref.py) + the reference's own compiled kernels (oracle/_ref/*.so),
    Trainer.test_gui_with_data (utils.py:1191-1223) -> (image*255).astype(u8) (nerfreal.py:109-110), same checkpoint / poses /
    audio windows / 512x512 as our arm.  Timed the way the reference runs it: one frame per call, host clock around the call (the
    call itself ends in .cpu().numpy(), i.e. it synchronises), plus CUDA events for the device-side share.  A baseline leg: it
    may import the reference; nothing of it is on the product path."""
    import torch
    import ref_ernerf
    from helpers import ernerf_inputs
    if not ref_ernerf.render_available():
        return {"unavailable": "oracle/_ref (reference kernels + staged reference python) not built"}
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        ref = ref_ernerf.ReferenceErnerf(H=H, W=W, device=str(dev))
    ins = [ernerf_inputs(f % 290, H, W) for f in range(n_frames + warmup)]
    host, devt = [], []
    for k, (pose, intr, auds, eye) in enumerate(ins):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        a.record()
        img = ref.render(k % 290, auds)
        u8 = (img * 255).astype(np.uint8)
        b.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        if k >= warmup:
            host.append(t1 - t0)
            devt.append(a.elapsed_time(b) * 1e-3)
    assert u8.shape == (H, W, 3)
    t = float(np.median(host))
    return {"value": 1.0 / t, "unit": "frames/s", "ms_per_frame": t * 1e3, "ms_per_frame_mean": float(np.mean(host)) * 1e3,
            "ms_per_frame_cuda_events": float(np.median(devt)) * 1e3, "frames": n_frames, "kind": "reference",
            "what": "unmodified reference NeRFNetwork + Trainer.test_gui_with_data + its own CUDA extensions compiled for sm_100 "
                    "(oracle/_ref), fp16 autocast, 512x512, same checkpoint / poses / audio windows; host buffers in (auds) and "
                    "out (u8 frame), median of per-frame host times (the reference synchronises every frame)"}


def _ref_py():
    p = os.path.join(ROOT, "oracle", "_ref", "py")
    if p not in sys.path:
        sys.path.insert(0, p)
    return p


def _event_ms(fn, K, warm, flush):
    """median device time of fn() over K calls (CUDA events on the current stream, L2 flushed outside the events)"""
    import torch
    for _ in range(warm):
        fn()
    ms = []
    for k in range(K):
        flush.fill_(k & 0xff)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))


def torch_gpu_wav2lip(dev, flush, B=16):
    """SURVEY 2b "cuDNN/cuBLAS on B200 are the bar to beat for the dense heads": the REFERENCE's own Wav2Lip nn.Module
    (wav2lip/models/wav2lip.py:87-125, staged unmodified in oracle/_ref/py) on this same GPU, same seeded weights / inputs / batch.
    (a) fp32 exactly as lipreal.py:108-126 runs it (model.to(device), no autocast), (b) the strongest library setting: bf16,
    channels_last, cudnn.benchmark.  net_ms = network only (inputs resident, CUDA events); e2e = the reference's own batch loop:
    numpy batch build -> H2D -> net -> .cpu().numpy().transpose * 255 -> cv2 resize + paste of 16 frames (lipreal.py:207-214)."""
    import cv2
    import torch
    from helpers import seeded_wav2lip_state, wav2lip_inputs
    if not os.path.isdir(os.path.join(_ref_py(), "wav2lip", "models")):
        return {"unavailable": "oracle/_ref/py/wav2lip/models not staged"}
    from wav2lip.models import Wav2Lip
    model = Wav2Lip().eval()
    model.load_state_dict(seeded_wav2lip_state(2), strict=True)
    model = model.to(dev)
    mel_np, faces = wav2lip_inputs(B)

    def build_batch():                                  # lipreal.py:108-116
        img = faces.copy()
        masked = img.copy()
        masked[:, img.shape[1] // 2:] = 0
        x = np.concatenate((masked, img), axis=3) / 255.
        return torch.FloatTensor(np.transpose(x, (0, 3, 1, 2))), torch.FloatTensor(mel_np)

    x_h, mel_h = build_batch()
    x32, mel32 = x_h.to(dev), mel_h.to(dev)
    out = {}
    with torch.no_grad():
        out["fp32_net_ms"] = _event_ms(lambda: model(mel32, x32), 20, 5, flush)
        rng = np.random.default_rng(1)
        full = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(B)]

        def e2e():
            x, m = build_batch()
            pred = model(m.to(dev), x.to(dev))
            pred = pred.cpu().numpy().transpose(0, 2, 3, 1) * 255.
            for i in range(B):                          # lipreal.py:207-214 (process_frames; a second thread in the reference)
                f = full[i].copy()
                f[176:368, 160:352] = cv2.resize(pred[i].astype(np.uint8), (192, 192))
        for _ in range(3):
            e2e()
        ts = []
        for _ in range(10):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e2e()
            ts.append(time.perf_counter() - t0)
        out["fp32_e2e_ms"] = float(np.median(ts)) * 1e3
        prev = torch.backends.cudnn.benchmark
        torch.backends.cudnn.benchmark = True
        m16 = model.to(dtype=torch.bfloat16, memory_format=torch.channels_last)
        x16 = x32.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        mel16 = mel32.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        out["bf16_channels_last_net_ms"] = _event_ms(lambda: m16(mel16, x16), 20, 8, flush)
        g = torch.cuda.CUDAGraph()                      # the best a PyTorch user can do about launch latency
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(3):
                m16(mel16, x16)
        torch.cuda.current_stream().wait_stream(s)
        try:
            with torch.cuda.graph(g):
                m16(mel16, x16)
            out["bf16_channels_last_cudagraph_net_ms"] = _event_ms(g.replay, 20, 3, flush)
        except Exception as e:                          # noqa: BLE001
            out["bf16_channels_last_cudagraph_net_ms"] = None
            out["cudagraph_error"] = repr(e)[:200]
        torch.backends.cudnn.benchmark = prev
    out.update(batch=B, what="reference Wav2Lip nn.Module (unmodified, oracle/_ref/py) through PyTorch/cuDNN on the same GPU; ms per batch of 16 frames")
    return out


def torch_gpu_musetalk(dev, flush, B=16):
    """The same-box PyTorch GPU baseline for MuseTalk: diffusers is absent, so the library arm is the oracle
    restatement of UNet2DConditionModel + AutoencoderKL.decode run through PyTorch (cuDNN convs, cuBLAS GEMMs, SDPA attention) in
    fp16 as the reference runs it (musereal.py:60-62), same seeded weights and input shapes as our arm.  Network only."""
    import torch
    from oracle import musetalk_oracle as M
    u, v = M.UNET_CFG, M.VAE_CFG
    usd = {k: t.to(dev, torch.float16) for k, t in M.seeded_state(M.unet_param_shapes(u), 5).items()}
    vsd = {k: t.to(dev, torch.float16) for k, t in M.seeded_state(M.vae_decoder_param_shapes(v), 6).items()}
    rng = np.random.default_rng(11)
    lat = torch.from_numpy((rng.standard_normal((B, 8, 32, 32)) * 0.18215 * 5).astype(np.float16)).to(dev)
    wh = torch.from_numpy(rng.standard_normal((B, 50, 384)).astype(np.float16)).to(dev)
    M.USE_SDPA = True
    prev = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = True
    out = {}
    try:
        out["fp16_net_ms"] = _event_ms(lambda: M.infer_device(usd, vsd, lat, wh, u, v), 5, 3, flush)
        usd_c = {k: (t.contiguous(memory_format=torch.channels_last) if t.dim() == 4 else t) for k, t in usd.items()}
        vsd_c = {k: (t.contiguous(memory_format=torch.channels_last) if t.dim() == 4 else t) for k, t in vsd.items()}
        lat_c = lat.contiguous(memory_format=torch.channels_last)
        out["fp16_channels_last_net_ms"] = _event_ms(lambda: M.infer_device(usd_c, vsd_c, lat_c, wh, u, v), 5, 3, flush)
    finally:
        M.USE_SDPA = False
        torch.backends.cudnn.benchmark = prev
    out.update(batch=B, what="PyTorch restatement of the diffusers SD-1.x UNet (t=0) + sd-vae-ft-mse decoder (oracle/musetalk_oracle.py) on the same "
                             "GPU, fp16, cuDNN benchmark mode, SDPA attention; ms per batch of 16 frames, network only (diffusers itself is absent)")
    return out


def torch_gpu_whisper(dev, flush):
    """the vendored reference Whisper (musetalk/whisper, staged unmodified) on the same GPU: Audio2Feature.audio2feat on one
    52-chunk window (audio2feature.py:99-112 -> transcribe -> log_mel + AudioEncoder, fp16 as transcribe defaults), host audio in,
    numpy features out -- the call MuseASR.run_step makes per batch (museasr.py:23-27)."""
    import types
    import torch
    from helpers import WHISPER_TINY, seeded_whisper_state, synthetic_speech
    if not os.path.isdir(os.path.join(_ref_py(), "musetalk", "whisper")):
        return {"unavailable": "oracle/_ref/py/musetalk/whisper not staged"}
    for m in ("ffmpeg", "soundfile"):
        sys.modules.setdefault(m, types.ModuleType(m))
    from musetalk.whisper.audio2feature import Audio2Feature
    from musetalk.whisper.whisper.model import ModelDimensions, Whisper
    dims = ModelDimensions(n_vocab=51865, n_text_ctx=448, n_text_state=384, n_text_head=6, n_text_layer=4, **WHISPER_TINY)
    model = Whisper(dims).eval()
    model.encoder.load_state_dict({k: torch.from_numpy(v) for k, v in seeded_whisper_state(7).items()}, strict=True)
    model = model.to(dev)
    a2f = Audio2Feature.__new__(Audio2Feature)
    a2f.model = model
    audio = synthetic_speech(52 * 320, 100)
    with torch.no_grad():
        for _ in range(3):
            feat = a2f.audio2feat(audio)
        ts = []
        for _ in range(10):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            feat = a2f.audio2feat(audio)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
    return {"ms_per_window": float(np.median(ts)) * 1e3, "feature_shape": list(feat.shape),
            "what": "vendored reference Whisper-tiny (unmodified, oracle/_ref/py) Audio2Feature.audio2feat on one 52-chunk window, host audio in / "
                    "numpy out, on the same GPU (the reference pads every window to 30 s before the encoder, transcribe.py:108)"}


class _TimedQueue:
    """the aiortc track queue of the fake tracks: every put is time-stamped; never reports a backlog"""

    def __init__(self):
        self.t = []

    async def put(self, x):
        self.t.append(time.perf_counter())

    def qsize(self):
        return 0


class _TimedTrack:
    def __init__(self):
        self._queue = _TimedQueue()


def plugin_drive(real, B, wav, n_tp, n_lat):
    """What the DROP-IN delivers (reference loops lipreal.py:232-250, musereal.py:267-290, nerfreal.py:129-156):
    the plugin object itself -- put_audio_frame -> asr.run_step -> inference -> process_frames -> VideoFrame + 2 AudioFrame on the
    (fake) tracks, its own threads and queues -- driven (1) at full speed, all audio queued up front: frames/s out of
    video_track._queue; (2) closed loop, one engine step in flight: p50 of [put_audio_frame of the last chunk of a step ->
    the last video frame of that step on the track] (SURVEY 8(d) latency definition; excludes the reference's fixed look-ahead).
    B = video frames per engine step (16 Lip / Muse, 1 ErNeRF)."""
    vt, at = _TimedTrack(), _TimedTrack()
    quit_event = threading.Event()
    real.asr.poll_timeout = 30.0                 # wait for audio instead of substituting silence: every frame below is a speech frame
    cps = 2 * B                                  # chunks per step
    pos = [0]

    def feed(n):
        for _ in range(n):
            i = pos[0] % (len(wav) // 320)
            real.put_audio_frame(wav[i * 320:(i + 1) * 320])
            pos[0] += 1

    feed(cps * (n_tp + 3))
    th = threading.Thread(target=real.render, args=(quit_event, None, at, vt), daemon=True)
    t_start = time.perf_counter()
    th.start()
    want = B * (n_tp + 2)
    while len(vt._queue.t) < want and time.perf_counter() - t_start < 120:
        time.sleep(0.002)
    tv = list(vt._queue.t)
    fps = (want - 2 * B) / (tv[want - 1] - tv[2 * B - 1]) if len(tv) >= want else None
    # drain what is left of the throughput phase
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 10:
        n0 = len(vt._queue.t)
        time.sleep(0.1)
        if len(vt._queue.t) == n0:
            break
    lats = []
    for _ in range(n_lat):
        base = len(vt._queue.t)
        feed(cps)
        t_put = time.perf_counter()
        while len(vt._queue.t) < base + B and time.perf_counter() - t_put < 20:
            time.sleep(0.0002)
        if len(vt._queue.t) >= base + B:
            lats.append((vt._queue.t[base + B - 1] - t_put) * 1e3)
    quit_event.set()
    feed(cps * 2)                                # unblock a run_step that waits for audio
    th.join(timeout=20)
    nv, na = len(vt._queue.t), len(at._queue.t)
    return {"frames_per_s": fps, "p50_last_chunk_to_last_frame_of_step_ms": float(np.median(lats)) if lats else None,
            "frames_per_step": B, "steps_full_speed": n_tp, "steps_closed_loop": len(lats), "video_frames": nv, "audio_frames": na,
            "audio_per_video": na / max(nv, 1)}


def plugin_legs(args, dev, local, shared, ernerf_blob, ernerf_cfg, w2v):
    """heads.<head>.plugin: LipReal / MuseReal / NeRFReal constructed as app.py does (fake avatar / tracks), see plugin_drive"""
    import torch
    from helpers import load_pose_fixture, synthetic_speech
    from test_plugin_cpu import _fake_avatar, _fake_muse_avatar, make_opt
    out = {}
    wav = synthetic_speech(160000, 0)
    try:
        if "wav2lip_blob" in shared:
            from mere_fusion_b200.plugin.lipreal import LipReal
            from mere_fusion_b200.wav2lip import Wav2LipEngine
            real = LipReal(make_opt(), engine=Wav2LipEngine(blob=shared["wav2lip_blob"], max_batch=16, device=local), avatar=_fake_avatar(), paste="gpu")
            out["wav2lip"] = plugin_drive(real, 16, wav, 40, 15)
    except Exception as e:                                   # noqa: BLE001
        out["wav2lip"] = {"error": repr(e)[:300]}
    try:
        if "musetalk_engine" in shared:
            from mere_fusion_b200.plugin.musereal import MuseReal
            real = MuseReal(make_opt(), engine=shared["musetalk_engine"], audio_processor=shared["a2f"], avatar=_fake_muse_avatar(), paste="gpu")
            out["musetalk"] = plugin_drive(real, 16, wav, 12, 8)
    except Exception as e:                                   # noqa: BLE001
        out["musetalk"] = {"error": repr(e)[:300]}
    try:
        from mere_fusion_b200.ernerf import ErnerfRenderer
        from mere_fusion_b200.ernerf_data import ErnerfPoseProvider
        from mere_fusion_b200.plugin.nerfreal import NeRFReal
        pf = load_pose_fixture()
        tr = dict(cx=W / 2.0, cy=H / 2.0, focal_len=float(pf["focal_len"]) * H / (2 * float(pf["cy"])),
                  frames=[dict(transform_matrix=pf["raw"][i].tolist(), img_id=int(pf["img_id"][i])) for i in range(290)])
        au = np.zeros(int(pf["img_id"][:290].max()) + 1)
        au[:min(len(au), len(pf["au"]))] = pf["au"][:len(au)]
        for name, fn in (("ernerf_synthetic_logits", lambda fr: torch.zeros(((len(fr) - 400) // 320 + 1, 44))),
                         ("ernerf_with_wav2vec2", w2v.feature_fn if w2v is not None else None)):
            if fn is None:
                continue
            ren = ErnerfRenderer(blob=ernerf_blob, cfg=ernerf_cfg, device=local)
            real = NeRFReal(make_opt(W=W, H=H), ren, ErnerfPoseProvider(tr, au), feature_fn=fn, device=local)
            out[name] = plugin_drive(real, 1, wav, 300, 60)
    except Exception as e:                                   # noqa: BLE001
        out["ernerf"] = {"error": repr(e)[:300]}
    return out


class PipelinedD2H:
    """two device result buffers + two pinned slots filled by a copy stream: the device -> host copy of step k overlaps the compute of
    step k + 1, and the main stream waits for the copy of step k - 1 at the end of step k (so every copy lies inside some step's
    timed interval and every result is in host memory before the timed region ends)"""

    def __init__(self, dev, like):
        import torch
        self.torch, self.dev = torch, dev
        self.copy = torch.cuda.Stream(dev)
        self.buf = [torch.empty_like(like) for _ in range(2)]
        self.pin = [torch.empty(like.shape, dtype=like.dtype).pin_memory() for _ in range(2)]
        self.done = [torch.cuda.Event() for _ in range(2)]
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.used = [False, False]
        self.k = 0

    def dst(self):
        """the device buffer this step writes: the copy that read it two steps ago must have finished"""
        s = self.k & 1
        if self.used[s]:
            self.torch.cuda.current_stream(self.dev).wait_event(self.done[s])
        return self.buf[s]

    def push(self):
        torch = self.torch
        s = self.k & 1
        self.k += 1
        cur = torch.cuda.current_stream(self.dev)
        self.ready[s].record(cur)
        self.copy.wait_event(self.ready[s])
        with torch.cuda.stream(self.copy):
            self.pin[s].copy_(self.buf[s], non_blocking=True)
            self.done[s].record(self.copy)
        self.used[s] = True
        if self.used[s ^ 1]:
            cur.wait_event(self.done[s ^ 1])               # the previous step's copy, overlapped with this step's compute
        return self.pin[s]


def p50_latency_ms(step_host, n=30):
    """p50 audio-chunk -> frame: host clock from the call that receives the last chunk's features to the finished u8 frames
    in pinned host memory (excludes the reference's fixed look-ahead)"""
    import torch
    lat = []
    for k in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step_host(k)
        torch.cuda.synchronize()
        lat.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(lat))


def packed_on_all_ranks(pack_fn, rank, world, dev):
    """rank 0 packs the checkpoint once, ONE NCCL broadcast hands the blob to the peers (plus a small pickled
    metadata record); pack_fn() -> (blob ndarray, meta dict)"""
    import torch
    import torch.distributed as dist
    from mere_fusion_b200.dist import broadcast_bytes
    blob, meta = pack_fn() if rank == 0 else (None, None)
    if world == 1:
        return torch.from_numpy(blob).to(dev), meta
    t = broadcast_bytes(blob, src=0, device=dev)
    box = [meta]
    dist.broadcast_object_list(box, src=0, device=dev)
    return t, box[0]


def wav2lip_leg(args, dev, local, rank, world, flush, timed_fn, pk, S=96, shared=None):
    """BASELINE configs[1]-shaped leg.  S = 96: the architecture the reference actually ships (96x96 crop);
    S = 256: the 256x256 EXTENSION of SURVEY 8(d) config 2 (ii) (not a reference net; parity vs our own fp32 restatement).
    16 frames per step: mel windows -> Wav2Lip (tcgen05 implicit-GEMM convs, bf16) -> cv2-exact resize + paste into
    the 512x512 avatar frame.  Returns the sub-object reported under "heads"."""
    import ctypes
    import torch
    from helpers import seeded_wav2lip_state, wav2lip_inputs
    from mere_fusion_b200._lib import check, lib
    from mere_fusion_b200.wav2lip import Wav2LipEngine
    B = 16

    def pack():
        from mere_fusion_b200.wav2lip_pack import pack_wav2lip
        blob, pb = pack_wav2lip(seeded_wav2lip_state(2, face_hw=S), nominal_batch=B, face_hw=S)
        return blob, dict(flops=pb.flops_per_sample, n_ops=len(pb.ops))

    blob, meta = packed_on_all_ranks(pack, rank, world, dev)
    eng = Wav2LipEngine(blob=blob, max_batch=B, device=local, face_hw=S)
    eng.flops_per_frame, eng.n_ops = meta["flops"], meta["n_ops"]
    if shared is not None:
        shared["wav2lip_blob"] = blob
    rng = np.random.default_rng(1)
    n_av = 25
    frames = torch.from_numpy(rng.integers(0, 256, (n_av, H, W, 3), dtype=np.uint8)).to(dev)
    faces_all = torch.from_numpy(rng.integers(0, 256, (n_av, S, S, 3), dtype=np.uint8)).to(dev)
    mels = [torch.from_numpy(wav2lip_inputs(B, mel_seed=100 + rank * 16 + i)[0]) for i in range(8)]
    mel_dev = [m.to(dev) for m in mels]
    mel_pin = [m.pin_memory() for m in mels]
    mel_stage = torch.empty_like(mel_dev[0])
    sel = torch.empty((B, S, S, 3), dtype=torch.uint8, device=dev)
    pred = torch.empty_like(sel)
    out = torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev)
    out_pin = torch.empty((B, H, W, 3), dtype=torch.uint8).pin_memory()
    rows_all = []
    for k in range(8):
        idxs = [(k * B + i) % n_av for i in range(B)]
        rows_all.append((torch.as_tensor(idxs, device=dev), np.array([(j, 176, 368, 160, 352) for j in idxs], np.int32)))
    h = eng.ctx.handle

    def core(k, mel, n=B, dst=None):
        idx_t, rows = rows_all[k % 8]
        torch.index_select(faces_all, 0, idx_t[:n], out=sel[:n])
        eng.forward(mel[:n], sel[:n], out=pred[:n])
        s = torch.cuda.current_stream(dev)
        check(h, lib().mf_paste_resize_u8(h, ctypes.c_void_p(frames.data_ptr()), n_av, H, W, ctypes.c_void_p(pred.data_ptr()), S, n,
                                          rows.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), ctypes.c_void_p((dst if dst is not None else out).data_ptr()),
                                          ctypes.c_void_p(s.cuda_stream)), "mf_paste_resize_u8")

    def step_host_b1(k):                                   # the latency metric "also with B=1"
        mel_stage[:1].copy_(mel_pin[k % 8][:1], non_blocking=True)
        core(k, mel_stage, 1)
        out_pin[:1].copy_(out[:1], non_blocking=True)

    def step(k):
        core(k, mel_dev[k % 8])

    def step_host(k):
        mel_stage.copy_(mel_pin[k % 8], non_blocking=True)
        core(k, mel_stage)
        out_pin.copy_(out, non_blocking=True)

    pipe = PipelinedD2H(dev, out)

    def step_host_pipelined(k):                            # the D2H of step k (12.6 MB) overlaps the compute of step k + 1
        mel_stage.copy_(mel_pin[k % 8], non_blocking=True)
        core(k, mel_stage, dst=pipe.dst())
        pipe.push()

    K = max(20, args.steps // 4)
    tot, per, _ = timed_fn(step, K, args.warmup)
    e2e, _, _ = timed_fn(step_host_pipelined, K, args.warmup)
    e2e_serial, _, _ = timed_fn(step_host, K, args.warmup)
    p50 = p50_latency_ms(step_host)
    p50_b1 = p50_latency_ms(step_host_b1)
    # dominant kernel by time share: the two 64->64 3x3 convs at SxS (last decoder block), one of them timed live
    op = eng.n_ops - 4
    eng.profile_op(op)
    ms = []
    for k in range(10):
        flush.fill_(k)
        step(k)
        ms.append(eng.last_op_ms())
    eng.profile_op(-1)
    flop = 2 * 64 * 64 * 9 * S * S * B
    m = float(np.mean(ms)) * 1e-3
    # ---- cross-session batching: four sessions' 16-frame requests as ONE pass of a shared engine
    # (scheduler.SharedEngine) against the same four requests run one after the other on this engine
    coalesce = None
    if S == 96:
        from mere_fusion_b200.scheduler import SharedEngine
        NS = 4
        sh = SharedEngine(Wav2LipEngine(blob=blob, max_batch=B * NS, device=local, face_hw=S), threaded=False)
        sels = [torch.empty((B, S, S, 3), dtype=torch.uint8, device=dev) for _ in range(NS)]
        preds = [torch.empty_like(sels[0]) for _ in range(NS)]
        outs = [torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev) for _ in range(NS)]

        def paste(k, j):
            idx_t, rows = rows_all[(k + j) % 8]
            st = torch.cuda.current_stream(dev)
            check(h, lib().mf_paste_resize_u8(h, ctypes.c_void_p(frames.data_ptr()), n_av, H, W, ctypes.c_void_p(preds[j].data_ptr()), S, B,
                                              rows.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), ctypes.c_void_p(outs[j].data_ptr()),
                                              ctypes.c_void_p(st.cuda_stream)), "mf_paste_resize_u8")

        def step_coalesced(k):
            reqs = []
            for j in range(NS):
                torch.index_select(faces_all, 0, rows_all[(k + j) % 8][0], out=sels[j])
                reqs.append(sh.submit(mel_dev[(k + j) % 8], sels[j], out=preds[j]))
            sh.flush()
            for j, rq in enumerate(reqs):
                sh.wait(rq)
                paste(k, j)

        def step_sequential(k):
            for j in range(NS):
                torch.index_select(faces_all, 0, rows_all[(k + j) % 8][0], out=sels[j])
                eng.forward(mel_dev[(k + j) % 8], sels[j], out=preds[j])
                paste(k, j)

        tc, _, _ = timed_fn(step_coalesced, K, args.warmup)
        ts, _, _ = timed_fn(step_sequential, K, args.warmup)
        coalesce = {"sessions": NS, "frames_per_step": NS * B, "coalesced_ms_per_step": tc / K, "sequential_ms_per_step": ts / K,
                    "coalesced_frames_per_s": world * K * NS * B / (tc / 1e3), "sequential_frames_per_s": world * K * NS * B / (ts / 1e3),
                    "engine_calls_per_step": 1}
        sh.shutdown()
    wl = ("wav2lip_96x96_B16 -> paste into 512x512 (reference architecture; the 256x256 net of configs[1] does not exist in the reference, SURVEY M2)"
          if S == 96 else
          "wav2lip_256x256_B16 -> paste into 512x512 (BASELINE configs[1]; EXTENDED generator of SURVEY 8(d) config 2 (ii) / section 7 step 4 -- "
          "not a reference architecture, random weights, parity vs our own fp32 restatement only)")
    return {"workload": wl, "gflop_per_frame": eng.flops_per_frame / 1e9, "cross_session_batching": coalesce,
            "value": world * K * B / (tot / 1e3), "unit": "frames/s", "ms_per_step": tot / K, "frames_per_step": B,
            "e2e": {"value": world * K * B / (e2e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": int(mel_pin[0].numel() * 4),
                    "d2h_bytes_per_step": int(out_pin.numel()), "serial_value": world * K * B / (e2e_serial / 1e3),
                    "note": "device->host copy of step k on a copy stream, overlapped with step k+1 (two pinned slots); serial_value = one stream"},
            "p50_chunk_to_frame_ms": p50, "p50_chunk_to_frame_ms_B1": p50_b1,
            "gpu_launches_per_step": eng.last_launches + 1, "dtype": "bf16",
            "algorithmic_tflops": eng.flops_per_frame * B / (tot / K * 1e-3) / 1e12,
            "roofline": {"kernel": f"k_conv_tma<2> (last face_decoder_block 3x3 64->64 @{S}x{S}, B=16)", "bound": "tensor",
                         "achieved": flop / m / 1e12, "peak": pk["tf"], "unit": "TFLOP/s", "frac": flop / m / 1e12 / pk["tf"],
                         "ms_per_launch": m * 1e3, "traffic": None, "peak_source": pk["src"] + " burst"}}


def musetalk_leg(args, dev, local, rank, world, flush, timed_fn, pk, shared=None):
    """BASELINE configs[2]-shaped leg: MuseTalk v1 architecture (SD-1.x UNet + sd-vae-ft-mse decoder, random weights), 16
    frames per step: 52-chunk audio window -> Whisper-tiny features (log-mel + encoder on the GPU) -> [16,50,384] chunks ->
    PE + UNet (t = 0) + VAE decode -> cv2-exact resize + mask blend into the 512x512 avatar frames."""
    import ctypes
    import struct
    import torch
    from helpers import WHISPER_TINY, seeded_whisper_state, synthetic_speech
    from oracle import musetalk_oracle as M          # parameter shapes + seeded weights only (no oracle compute on this path)
    from mere_fusion_b200._lib import check, lib
    from mere_fusion_b200.musetalk import MuseTalkEngine
    from mere_fusion_b200.whisper import Audio2Feature, WhisperEngine
    B = 16
    u, v = M.UNET_CFG, M.VAE_CFG

    def pack_m():
        from mere_fusion_b200.musetalk_pack import pack_musetalk
        blob, pb = pack_musetalk(M.seeded_state(M.unet_param_shapes(u), 5), M.seeded_state(M.vae_decoder_param_shapes(v), 6), u, v,
                                 nominal_batch=B)
        op = next(i for i, rec in enumerate(pb.ops) if struct.unpack("<28i", rec[:112])[14:18] == (9, 128, 1152, 128))
        return blob, dict(flops=pb.flops_per_sample, unet=pb.unet_flops, vae=pb.vae_flops, op=op)

    def pack_w():
        from mere_fusion_b200.whisper_pack import pack_whisper
        blob, pb = pack_whisper(seeded_whisper_state(7), WHISPER_TINY)
        return blob, dict(flops=pb.flops_per_sample)

    blob_m, meta_m = packed_on_all_ranks(pack_m, rank, world, dev)
    blob_w, meta_w = packed_on_all_ranks(pack_w, rank, world, dev)
    eng = MuseTalkEngine(blob=blob_m, max_batch=B, device=local)
    eng.flops_per_frame, eng.unet_flops, eng.vae_flops = meta_m["flops"], meta_m["unet"], meta_m["vae"]
    wh = WhisperEngine(blob=blob_w, dims=WHISPER_TINY, device=local)
    wh.flops_per_call = meta_w["flops"]
    a2f = Audio2Feature(engine=wh)
    if shared is not None:
        shared["musetalk_engine"], shared["a2f"] = eng, a2f
    rng = np.random.default_rng(11 + rank)
    n_av = 12
    frames = torch.from_numpy(rng.integers(0, 200, (n_av, H, W, 3), dtype=np.uint8)).to(dev)
    lat_all = torch.from_numpy((rng.standard_normal((n_av, 8, 32, 32)) * 0.18215 * 5).astype(np.float16)).to(dev)
    boxes, crops, moffs, masks, off = [], [], [], [], 0
    for i in range(n_av):
        x1, y1 = 150 + i, 140 + 2 * i
        x2, y2 = x1 + 200 + i, y1 + 210
        xs, ys, xe, ye = x1 - 40, y1 - 30, x2 + 35, y2 + 45
        yy, xx = np.mgrid[0:ye - ys, 0:xe - xs]
        ramp = np.clip(255 - 3 * np.hypot(yy - (ye - ys) / 2, xx - (xe - xs) / 2) + 200, 0, 255).astype(np.uint8)
        m = np.repeat(ramp[:, :, None], 3, axis=2)
        boxes.append((y1, y2, x1, x2)); crops.append((ys, ye, xs, xe)); moffs.append(off)
        masks.append(m.reshape(-1)); off += m.size
    masks_d = torch.from_numpy(np.concatenate(masks)).to(dev)
    audio = [torch.from_numpy(synthetic_speech(52 * 320, 100 + rank * 8 + i)) for i in range(8)]
    audio_dev = [a.to(dev) for a in audio]
    audio_pin = [a.pin_memory() for a in audio]
    audio_stage = torch.empty_like(audio_dev[0])
    sel = torch.empty((B, 8, 32, 32), dtype=torch.float16, device=dev)
    pred = torch.empty((B, 256, 256, 3), dtype=torch.uint8, device=dev)
    out = torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev)
    out_pin = torch.empty((B, H, W, 3), dtype=torch.uint8).pin_memory()
    rows_all = []
    for k in range(8):
        idxs = [(k * B + i) % n_av for i in range(B)]
        rows = np.array([(j,) + boxes[j] + crops[j] for j in idxs], np.int32)
        rows_all.append((torch.as_tensor(idxs, device=dev), rows, np.array([moffs[j] for j in idxs], np.int64)))
    h = eng.ctx.handle

    def core(k, audio_t, n=B):
        idx_t, rows, mo = rows_all[k % 8]
        chunks = a2f.audio2chunks_device(None, fps=25.0, batch_size=B, start=5.0, audio_dev=audio_t)   # MuseASR.run_step (museasr.py:26-27)
        torch.index_select(lat_all, 0, idx_t[:n], out=sel[:n])
        eng.forward(sel[:n], chunks[:n], out=pred[:n])
        st = torch.cuda.current_stream(dev)
        check(h, lib().mf_paste_blend_u8(h, ctypes.c_void_p(frames.data_ptr()), n_av, H, W, ctypes.c_void_p(pred.data_ptr()), 256, n,
                                         rows.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), ctypes.c_void_p(masks_d.data_ptr()),
                                         masks_d.numel(), mo.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                         ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(st.cuda_stream)), "mf_paste_blend_u8")

    def step(k):
        core(k, audio_dev[k % 8])

    def step_host(k):
        audio_stage.copy_(audio_pin[k % 8], non_blocking=True)
        core(k, audio_stage)
        out_pin.copy_(out, non_blocking=True)

    def step_host_b1(k):                                   # the latency metric "also with B=1" (Whisper window included)
        audio_stage.copy_(audio_pin[k % 8], non_blocking=True)
        core(k, audio_stage, 1)
        out_pin[:1].copy_(out[:1], non_blocking=True)

    K = max(10, args.steps // 10)
    tot, per, _ = timed_fn(step, K, args.warmup)
    e2e, _, _ = timed_fn(step_host, K, args.warmup)
    p50 = p50_latency_ms(step_host, 15)
    launches = eng.last_launches + a2f.engine.last_launches + 1
    p50_b1 = p50_latency_ms(step_host_b1, 15)
    # dominant kernel by time share: the 128 -> 128 3x3 convs of the last VAE up block at 256x256 (k_conv_tma), one of them timed live
    eng.profile_op(meta_m["op"])
    ms = []
    for k in range(5):
        flush.fill_(k)
        step(k)
        ms.append(eng.last_op_ms())
    eng.profile_op(-1)
    flop = 2 * 128 * 128 * 9 * 256 * 256 * B
    m = float(np.mean(ms)) * 1e-3
    fl_step = eng.flops_per_frame * B + a2f.engine.flops_per_call
    return {"workload": "musetalk_v1 (SD-1.x UNet t=0 + sd-vae-ft-mse decoder) 256x256 B16 + whisper-tiny features per batch -> blend into 512x512 "
                        "(BASELINE configs[2]; random weights: parity unpinned, DESIGN.md 6)",
            "value": world * K * B / (tot / 1e3), "unit": "frames/s", "ms_per_step": tot / K, "frames_per_step": B,
            "e2e": {"value": world * K * B / (e2e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": int(audio_pin[0].numel() * 4),
                    "d2h_bytes_per_step": int(out_pin.numel())},
            "p50_chunk_to_frame_ms": p50, "p50_chunk_to_frame_ms_B1": p50_b1,
            "gpu_launches_per_step": int(launches), "dtype": "bf16",
            "algorithmic_tflops": fl_step / (tot / K * 1e-3) / 1e12,
            "gflop_per_frame": {"unet": eng.unet_flops / 1e9, "vae_decoder": eng.vae_flops / 1e9, "whisper_per_batch": a2f.engine.flops_per_call / 1e9},
            "roofline": {"kernel": "k_conv_tma<2> (cta_group::2; vae decoder up_blocks.3 resnet conv 3x3 128->128 @256x256, B=16)", "bound": "tensor",
                         "achieved": flop / m / 1e12, "peak": pk["tf"], "unit": "TFLOP/s", "frac": flop / m / 1e12 / pk["tf"],
                         "ms_per_launch": m * 1e3, "traffic": None, "peak_source": pk["src"] + " burst"}}


def musetalk_streams_leg(args, dev, local, rank, world, timed_fn, shared, n_streams=8):
    """BASELINE configs[2] as written: 8 concurrent MuseTalk streams on one GPU.  Every stream is a session (own avatar latents,
    own audio window) served by the GPU's shared engine (scheduler.SharedEngine: the sessions' 16-frame requests of a step are
    coalesced into passes of the one resident UNet + VAE); one step = every stream advances 16 frames."""
    import torch
    from helpers import synthetic_speech
    from mere_fusion_b200.scheduler import SharedEngine
    B = 16
    eng, a2f = shared["musetalk_engine"], shared["a2f"]
    sh = SharedEngine(eng, threaded=False)
    rng = np.random.default_rng(77 + rank)
    lat = [torch.from_numpy((rng.standard_normal((B, 8, 32, 32)) * 0.18215 * 5).astype(np.float16)).to(dev) for _ in range(n_streams)]
    audio = [torch.from_numpy(synthetic_speech(52 * 320, 400 + rank * 8 + i)).to(dev) for i in range(n_streams)]
    preds = [torch.empty((B, 256, 256, 3), dtype=torch.uint8, device=dev) for _ in range(n_streams)]

    def step(k):
        reqs = []
        for i in range(n_streams):
            chunks = a2f.audio2chunks_device(None, fps=25.0, batch_size=B, start=5.0, audio_dev=audio[(i + k) % n_streams])
            reqs.append(sh.submit(lat[i], chunks, out=preds[i]))
        sh.flush()
        for rq in reqs:
            sh.wait(rq)

    K = max(3, args.steps // 30)
    tot, _, _ = timed_fn(step, K, 3)
    sh.shutdown()
    fps = world * K * n_streams * B / (tot / 1e3)
    return {"workload": f"{n_streams} concurrent MuseTalk streams per GPU on {world} GPU(s) (BASELINE configs[2]), 16 frames per stream and step, "
                        "Whisper window + UNet + VAE per stream through the shared engine (no paste)",
            "value": fps, "unit": "frames/s (aggregate)", "ms_per_step": tot / K, "frames_per_step_per_gpu": n_streams * B,
            "fps_per_stream": fps / (world * n_streams), "realtime_streams_capacity_25fps": fps / 25.0}


def mixed_leg(args, dev, local, rank, world, flush, timed_fn, shared, ernerf_blob, ernerf_cfg):
    """BASELINE configs[4] / SURVEY 8(d) config 5: 64 concurrent mixed sessions on 8 GPUs = 8 sessions per GPU, heads round-robin by
    session id (22 ErNeRF + 21 MuseTalk + 21 Wav2Lip in total, every GPU hosts all three heads).  This rank runs ITS 8 sessions
    (round-robin, dist.shard): at --gpus 1 the line is the per-GPU slice, at --gpus 8 the full config.
    One step = every session advances 16 video frames (0.64 s of video): ErNeRF 16 passes, each rendering one frame of every ErNeRF
    session of the GPU (mf_ernerf_render_batch; per-session contexts keep the EMA state); MuseTalk one Whisper window + one 16-frame UNet/VAE pass + blend per session on the shared
    engine; Wav2Lip one 16-frame pass per session, the same-GPU sessions COALESCED into one launch sequence by
    scheduler.SharedEngine, + paste.  value = sessions x 16 frames / step time, aggregate over ranks."""
    import ctypes
    import torch
    from helpers import ernerf_inputs, synthetic_speech, wav2lip_inputs
    from mere_fusion_b200._lib import check, lib
    from mere_fusion_b200.ernerf import ErnerfRenderer
    from mere_fusion_b200.dist import mixed_sessions, shard
    from mere_fusion_b200.scheduler import SharedEngine
    from mere_fusion_b200.wav2lip import Wav2LipEngine
    B, PER_GPU = 16, 8
    n_total = PER_GPU * world
    heads_all = [h for h, _ in mixed_sessions(n_total)]          # interleaved by head (22 + 21 + 21 at 64 sessions)
    # session -> GPU: round-robin (dist.shard), which gives every GPU all three heads
    mine = shard(list(enumerate(heads_all)), world, rank)
    n_lip = sum(1 for _, h in mine if h == "wav2lip")
    lip_eng = SharedEngine(Wav2LipEngine(blob=shared["wav2lip_blob"], max_batch=B * max(1, n_lip), device=local), threaded=False)
    muse_eng, a2f = shared["musetalk_engine"], shared["a2f"]
    hL, hM = lip_eng.ctx.handle, muse_eng.ctx.handle
    sessions = []
    nerf_sessions = []                          # the GPU's ErNeRF sessions render their frames together (ErnerfRenderer.render_batch)
    pins = []                                   # (pinned host result, device result) per session for the e2e arm
    h2d = 0
    for sid, head in mine:
        rng = np.random.default_rng(1000 + sid)
        out = torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev)
        out_pin = torch.empty((B, H, W, 3), dtype=torch.uint8).pin_memory()
        pins.append((out_pin, out))
        if head == "ernerf":
            ren = ErnerfRenderer(blob=ernerf_blob, cfg=ernerf_cfg, device=local)
            nerf_sessions.append(dict(ren=ren, sid=sid, out=out))
            ins = [ernerf_inputs((sid * 37 + f) % 300, H, W) for f in range(B)]
            auds_pin = torch.from_numpy(np.stack([i[2] for i in ins])).pin_memory()
            auds_dev = auds_pin.to(dev)
            stage = torch.empty_like(auds_dev)
            h2d += auds_pin.numel() * 4 + B * 84

            nerf_sessions[-1].update(ins=ins, auds_dev=auds_dev, auds_pin=auds_pin, stage=stage)
        elif head == "musetalk":
            n_av = 12
            frames = torch.from_numpy(rng.integers(0, 200, (n_av, H, W, 3), dtype=np.uint8)).to(dev)
            lat_all = torch.from_numpy((rng.standard_normal((n_av, 8, 32, 32)) * 0.18215 * 5).astype(np.float16)).to(dev)
            boxes, crops, moffs, masks, off = [], [], [], [], 0
            for i in range(n_av):
                x1, y1 = 150 + i, 140 + 2 * i
                x2, y2 = x1 + 200 + i, y1 + 210
                xs, ys, xe, ye = x1 - 40, y1 - 30, x2 + 35, y2 + 45
                m = np.full((ye - ys, xe - xs, 3), 180, np.uint8)
                boxes.append((y1, y2, x1, x2)); crops.append((ys, ye, xs, xe)); moffs.append(off)
                masks.append(m.reshape(-1)); off += m.size
            masks_d = torch.from_numpy(np.concatenate(masks)).to(dev)
            audio_pin = torch.from_numpy(synthetic_speech(52 * 320, 300 + sid)).pin_memory()
            audio_dev = audio_pin.to(dev)
            audio_stage = torch.empty_like(audio_dev)
            sel = torch.empty((B, 8, 32, 32), dtype=torch.float16, device=dev)
            pred = torch.empty((B, 256, 256, 3), dtype=torch.uint8, device=dev)
            h2d += audio_pin.numel() * 4

            def run(k, host, frames=frames, lat_all=lat_all, boxes=boxes, crops=crops, moffs=moffs, masks_d=masks_d, audio_dev=audio_dev,
                    audio_pin=audio_pin, audio_stage=audio_stage, sel=sel, pred=pred, out=out, n_av=n_av):
                src = audio_dev
                if host:
                    audio_stage.copy_(audio_pin, non_blocking=True)
                    src = audio_stage
                idxs = [(k * B + i) % n_av for i in range(B)]
                rows = np.array([(j,) + boxes[j] + crops[j] for j in idxs], np.int32)
                mo = np.array([moffs[j] for j in idxs], np.int64)
                chunks = a2f.audio2chunks_device(None, fps=25.0, batch_size=B, start=5.0, audio_dev=src)
                torch.index_select(lat_all, 0, torch.as_tensor(idxs, device=dev), out=sel)
                muse_eng.forward(sel, chunks, out=pred)
                st = torch.cuda.current_stream(dev)
                check(hM, lib().mf_paste_blend_u8(hM, ctypes.c_void_p(frames.data_ptr()), n_av, H, W, ctypes.c_void_p(pred.data_ptr()), 256, B,
                                                  rows.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), ctypes.c_void_p(masks_d.data_ptr()),
                                                  masks_d.numel(), mo.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                                  ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(st.cuda_stream)), "mf_paste_blend_u8")
            sessions.append((head, run, None))
        else:
            n_av = 25
            frames = torch.from_numpy(rng.integers(0, 256, (n_av, H, W, 3), dtype=np.uint8)).to(dev)
            faces_all = torch.from_numpy(rng.integers(0, 256, (n_av, 96, 96, 3), dtype=np.uint8)).to(dev)
            mel_pin = torch.from_numpy(wav2lip_inputs(B, mel_seed=500 + sid)[0]).pin_memory()
            mel_dev = mel_pin.to(dev)
            mel_stage = torch.empty_like(mel_dev)
            sel = torch.empty((B, 96, 96, 3), dtype=torch.uint8, device=dev)
            pred = torch.empty_like(sel)
            h2d += mel_pin.numel() * 4

            def submit(k, host, faces_all=faces_all, mel_dev=mel_dev, mel_pin=mel_pin, mel_stage=mel_stage, sel=sel, pred=pred, n_av=n_av):
                src = mel_dev
                if host:
                    mel_stage.copy_(mel_pin, non_blocking=True)
                    src = mel_stage
                idxs = [(k * B + i) % n_av for i in range(B)]
                torch.index_select(faces_all, 0, torch.as_tensor(idxs, device=dev), out=sel)
                return lip_eng.submit(src, sel, out=pred), idxs

            def finish(req, idxs, frames=frames, pred=pred, out=out, n_av=n_av):
                lip_eng.wait(req)
                rows = np.array([(j, 176, 368, 160, 352) for j in idxs], np.int32)
                st = torch.cuda.current_stream(dev)
                check(hL, lib().mf_paste_resize_u8(hL, ctypes.c_void_p(frames.data_ptr()), n_av, H, W, ctypes.c_void_p(pred.data_ptr()), 96, B,
                                                   rows.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), ctypes.c_void_p(out.data_ptr()),
                                                   ctypes.c_void_p(st.cuda_stream)), "mf_paste_resize_u8")
            sessions.append((head, submit, finish))

    def nerf_tick(k, host):
        """16 frames of every ErNeRF session of this GPU: frame f of all sessions in one batched pass (<= 4 sessions per pass)"""
        for ns in nerf_sessions:
            if host:
                ns["stage"].copy_(ns["auds_pin"], non_blocking=True)
        for f in range(B):
            for c0 in range(0, len(nerf_sessions), 4):
                grp = nerf_sessions[c0:c0 + 4]
                frs = []
                for ns in grp:
                    p, intr, _, eye = ns["ins"][(f + k) % B]
                    src = ns["stage"] if host else ns["auds_dev"]
                    frs.append(dict(pose=p, intrinsics=intr, H=H, W=W, auds=src[(f + k) % B], eye=eye))
                ErnerfRenderer.render_batch([ns["ren"] for ns in grp], frs, outs=[ns["out"][f] for ns in grp])

    def tick(k, host):
        lip = []
        if nerf_sessions:
            nerf_tick(k, host)
        for head, fn, fin in sessions:
            if head == "wav2lip":
                lip.append((fn(k, host), fin))
            else:
                fn(k, host)
        lip_eng.flush()                                            # ONE coalesced Wav2Lip pass for this GPU's sessions
        for (req, idxs), fin in lip:
            fin(req, idxs)
        if host:
            for out_pin, out in pins:
                out_pin.copy_(out, non_blocking=True)

    K = max(5, args.steps // 25)
    lip_b0 = lip_eng.batches
    tot, per, _ = timed_fn(lambda k: tick(k, False), K, 3)
    lip_calls = (lip_eng.batches - lip_b0) / (K + 3)
    e2e, _, _ = timed_fn(lambda k: tick(k, True), K, 3)
    counts = {h: sum(1 for _, hh in mine if hh == h) for h in ("ernerf", "musetalk", "wav2lip")}
    frames_step = PER_GPU * B
    return {"workload": f"{n_total} concurrent mixed sessions on {world} GPU(s), 8 per GPU, heads round-robin by session id "
                        f"(all ranks: {heads_all.count('ernerf')} ErNeRF + {heads_all.count('musetalk')} MuseTalk + {heads_all.count('wav2lip')} Wav2Lip; "
                        "BASELINE configs[4] is this at 8 GPUs); one step = 16 frames per session",
            "sessions_on_rank0": counts, "value": world * K * frames_step / (tot / 1e3), "unit": "frames/s (aggregate, all sessions)",
            "ms_per_step": tot / K, "frames_per_step_per_gpu": frames_step,
            "realtime_sessions_capacity": world * K * frames_step / (tot / 1e3) / 25.0,
            "e2e": {"value": world * K * frames_step / (e2e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(sum(p.numel() for p, _ in pins))},
            "wav2lip_engine_calls_per_step": lip_calls, "wav2lip_sessions_coalesced": n_lip}


def main():
    sys.stdout = sys.stderr   # ONE JSON line on stdout: everything else (the plugins' own fps prints, library chatter) goes to stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native")
    ap.add_argument("--workload", default="ernerf")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-wav2lip", action="store_true")
    ap.add_argument("--no-musetalk", action="store_true")
    ap.add_argument("--no-asr", action="store_true")
    ap.add_argument("--no-mixed", action="store_true")
    ap.add_argument("--no-reference-gpu", action="store_true")
    ap.add_argument("--no-plugin", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from helpers import ernerf_inputs, load_ernerf_fixture
    from mere_fusion_b200.ernerf import ErnerfRenderer
    from mere_fusion_b200.ernerf_pack import pack_ernerf
    from mere_fusion_b200._lib import MfErnerfCfg

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- weights: rank 0 packs the checkpoint, one NCCL broadcast hands the blob to the peers
    import ctypes
    if rank == 0:
        sd, md = load_ernerf_fixture()
        blob_np, cfg = pack_ernerf(sd, md)
        blob = torch.from_numpy(blob_np).to(dev)
        meta = torch.tensor([blob.numel()], dtype=torch.int64, device=dev)
        cfg_t = torch.frombuffer(bytearray(bytes(cfg)), dtype=torch.uint8).to(dev)
    else:
        meta = torch.zeros(1, dtype=torch.int64, device=dev)
        cfg_t = torch.zeros(ctypes.sizeof(MfErnerfCfg), dtype=torch.uint8, device=dev)
    if world > 1:
        dist.broadcast(meta, 0)
        if rank != 0:
            blob = torch.empty(int(meta.item()), dtype=torch.uint8, device=dev)
        dist.broadcast(blob, 0)
        dist.broadcast(cfg_t, 0)
        cfg = MfErnerfCfg.from_buffer_copy(cfg_t.cpu().numpy().tobytes())
    ren = ErnerfRenderer(blob=blob, cfg=cfg, device=local)

    # ---- synthetic inputs: each rank renders its own stream of frames
    n_in = 16
    ins = [ernerf_inputs((rank * 37 + f) % 300, H, W) for f in range(n_in)]
    auds_dev = [torch.from_numpy(i[2]).to(dev) for i in ins]
    auds_pin = [torch.from_numpy(i[2]).pin_memory() for i in ins]
    out = torch.empty(H, W, 3, dtype=torch.uint8, device=dev)
    out_pin = torch.empty(H, W, 3, dtype=torch.uint8).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step(k):
        p, intr, _, eye = ins[k % n_in]
        ren.render(p, intr, H, W, auds_dev[k % n_in], eye, out=out)

    def step_host(k):                                # serial: H2D, render, D2H on one stream (what the latency metric uses)
        p, intr, _, eye = ins[k % n_in]
        ren.render_host(p, intr, H, W, auds_pin[k % n_in], eye, out_pin)

    def step_host_pipelined(k):                      # throughput: the D2H of frame k overlaps the render of frame k + 1 (side stream)
        p, intr, _, eye = ins[k % n_in]
        ren.render_host_async(p, intr, H, W, auds_pin[k % n_in], eye)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, K, Wm):
        for k in range(Wm):
            fn(k)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        barrier()
        t0 = time.perf_counter()
        for k in range(K):
            flush.fill_(k & 0xff)                     # L2 flush, outside the events
            evs[k][0].record()
            fn(k)
            evs[k][1].record()
        barrier()
        wall = time.perf_counter() - t0
        per = [a.elapsed_time(b) for a, b in evs]
        tot = torch.tensor([sum(per)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        return float(tot.item()), per, wall

    sampler = ClockSampler(local)
    sampler.start()
    total_ms, per, wall = timed(step, args.steps, args.warmup)
    launches = ren.last_launches * args.steps
    e2e_ms, e2e_per, _ = timed(step_host_pipelined, args.steps, args.warmup)
    e2e_serial_ms, _, _ = timed(step_host, args.steps, args.warmup)
    sampler.stop_ev.set()
    sampler.join(timeout=1.0)

    # p50 audio-chunk -> frame: host clock from the call with the last chunk's features to the u8
    # frame being complete in pinned host memory (excludes the reference's fixed look-ahead)
    lat = []
    for k in range(min(50, args.steps)):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step_host(k)
        torch.cuda.synchronize()
        lat.append((time.perf_counter() - t0) * 1e3)

    # ---- the literal "2048 rays / frame" reading -- a fixed subset of the 512x512 grid (every 128th
    # pixel) as explicit rays (utils.py:255-341: pixel centres -> camera directions -> world), same pipeline, device-resident
    def ray_subset(pose, intr):
        pose = np.asarray(pose, np.float32).reshape(4, 4)
        fx, fy, cx, cy = intr
        idx = np.arange(0, RAYS, 128)
        row, col = idx // W, idx % W
        d = np.stack([(col + 0.5 - cx) / fx, (row + 0.5 - cy) / fy, np.ones(len(idx))], -1).astype(np.float32)
        d /= np.linalg.norm(d, axis=-1, keepdims=True)
        rd = (d @ pose[:3, :3].T).astype(np.float32)
        ro = np.broadcast_to(pose[:3, 3], rd.shape).astype(np.float32)
        bgc = np.stack([np.linspace(-1, 1, H, dtype=np.float32)[row], np.linspace(-1, 1, W, dtype=np.float32)[col]], -1)
        return [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (ro, rd, bgc)]

    subs = [ray_subset(i[0], i[1]) for i in ins]
    out_sub = torch.empty(1, RAYS // 128, 3, dtype=torch.uint8, device=dev)

    def step_2048(k):
        p, intr, _, eye = ins[k % n_in]
        ro, rd, bgc = subs[k % n_in]
        ren.render(p, intr, H, W, auds_dev[k % n_in], eye, out=out_sub, rays_o=ro, rays_d=rd, bg_coords=bgc)

    sub_ms, _, _ = timed(step_2048, args.steps, args.warmup)

    # ---- roofline of the dominant kernel (k_head), timed live with events on its stream
    ren.profile(True)
    head_ms, head_samples = [], []
    for k in range(min(30, args.steps)):
        flush.fill_(k & 0xff)
        step(k)
        ms, n = ren.last_head_ms()
        head_ms.append(ms)
        head_samples.append(n)
    ren.profile(False)

    # ---- batched sessions (mf_ernerf_render_batch): FB sessions of the same avatar model, one frame each per pass -- what a GPU
    # that hosts several ErNeRF sessions runs (scheduler.ErnerfBatcher); images are bit-identical to single renders
    FB = 4
    rens_b = [ren] + [ErnerfRenderer(blob=blob, cfg=cfg, device=local) for _ in range(FB - 1)]
    outs_b = [torch.empty(H, W, 3, dtype=torch.uint8, device=dev) for _ in range(FB)]
    outs_b_pin = torch.empty(FB, H, W, 3, dtype=torch.uint8).pin_memory()
    auds_b_stage = [torch.empty_like(auds_dev[0]) for _ in range(FB)]

    def batch_frames(k, src):
        return [dict(pose=ins[(k * FB + j) % n_in][0], intrinsics=ins[(k * FB + j) % n_in][1], H=H, W=W, auds=src(k, j),
                     eye=ins[(k * FB + j) % n_in][3]) for j in range(FB)]

    def step_batch(k):
        ErnerfRenderer.render_batch(rens_b, batch_frames(k, lambda k, j: auds_dev[(k * FB + j) % n_in]), outs=outs_b)

    def step_batch_host(k):
        for j in range(FB):
            auds_b_stage[j].copy_(auds_pin[(k * FB + j) % n_in], non_blocking=True)
        ErnerfRenderer.render_batch(rens_b, batch_frames(k, lambda k, j: auds_b_stage[j]), outs=outs_b)
        for j in range(FB):
            outs_b_pin[j].copy_(outs_b[j], non_blocking=True)

    Kb = max(20, args.steps // 4)
    batch_ms, _, _ = timed(step_batch, Kb, args.warmup)
    batch_e2e_ms, _, _ = timed(step_batch_host, Kb, args.warmup)
    for r_ in rens_b:
        r_.profile(True)
    bh_ms, bh_samples = [], []
    for k in range(min(20, args.steps)):
        flush.fill_(k & 0xff)
        step_batch(k)
        ms, n = rens_b[0].last_head_ms()
        bh_ms.append(ms)
        bh_samples.append(n + sum(r_.last_head_ms(want_ms=False)[1] for r_ in rens_b[1:]))
    for r_ in rens_b:
        r_.profile(False)
    ren.reset()

    # ---- the acoustic model behind NerfASR.run_step (wav2vec2 XLSR-53 shape, random weights): one 28-chunk window every 8 chunks
    # (= every 4 video frames, nerfasr.py:105-124); reported beside the render, not inside its step (SURVEY 8(d) config 4 feeds
    # synthetic logit windows)
    asr = None
    w2v = None
    if not args.no_asr:
        from helpers import W2V_XLSR53, seeded_w2v_state, synthetic_speech
        from mere_fusion_b200.wav2vec2 import Wav2Vec2Engine

        def pack_a(fused):
            from mere_fusion_b200.wav2vec2_pack import pack_wav2vec2
            b, pb_ = pack_wav2vec2(seeded_w2v_state(22, W2V_XLSR53), W2V_XLSR53, fused_stack=fused)
            return b, dict(flops=pb_.flops_per_sample, frames=pb_.n_frames)

        # one window per call: the transformer layers as ONE persistent kernel (csrc/w2v_stack.cuh); the engine that batches the
        # windows of several sessions keeps the op-by-op program (its GEMMs take any number of rows)
        blob_a, meta_a = packed_on_all_ranks(lambda: pack_a(True), rank, world, dev)
        w2v = Wav2Vec2Engine(blob=blob_a, cfg=W2V_XLSR53, device=local, n_frames=meta_a["frames"], max_batch=1)
        wins = [synthetic_speech(8960, 200 + rank * 8 + i) for i in range(8)]
        asr_ms, _, _ = timed(lambda k: w2v.feature_fn(wins[k % 8]), max(20, args.steps // 4), args.warmup)
        launches_1 = w2v.last_launches
        blob_b, meta_b = packed_on_all_ranks(lambda: pack_a(False), rank, world, dev)
        w2v_b = Wav2Vec2Engine(blob=blob_b, cfg=W2V_XLSR53, device=local, n_frames=meta_b["frames"], max_batch=4)
        # four sessions' windows in one pass (mf_wav2vec2_logits_batch, scheduler.AsrBatcher), host windows in
        win_pin = torch.from_numpy(np.stack(wins[:4])).pin_memory()
        win_dev = torch.empty((4, 8960), dtype=torch.float32, device=dev)

        def asr_batch(k):
            win_dev.copy_(win_pin, non_blocking=True)
            w2v_b.logits_batch(win_dev)
        asr_b_ms, _, _ = timed(asr_batch, max(20, args.steps // 4), args.warmup)
        asr = {"workload": "wav2vec2 XLSR-53-large CTC (315 M parameters, random weights), one 8960-sample window per call, host window in",
               "ms_per_window": asr_ms / max(20, args.steps // 4), "ms_per_video_frame_amortised": asr_ms / max(20, args.steps // 4) / 4,
               "gpu_launches_per_window": launches_1, "gflop_per_window": meta_a["flops"] / 1e9,
               "weight_streaming_bound_ms": 630e6 / (peaks()["hbm"] * 1e9) * 1e3,
               "frac_of_weight_streaming_bound": (630e6 / (peaks()["hbm"] * 1e9) * 1e3) / (asr_ms / max(20, args.steps // 4)),
               "batched_4_sessions": {"ms_per_pass": asr_b_ms / max(20, args.steps // 4), "ms_per_window": asr_b_ms / max(20, args.steps // 4) / 4,
                                      "ms_per_video_frame_amortised": asr_b_ms / max(20, args.steps // 4) / 16}}
        del w2v_b, blob_b

    # ---- SURVEY 8(e) single-stream scaling: ONE session's frames sharded round-robin over the ranks; every rank follows the
    # session's audio state (mf_ernerf_encode_audio on every frame) and renders only its own frames with the feature passed
    # explicitly (bit-identical to the in-order stream: tests/test_ernerf_gpu.py).  K frames per rank and pass.
    ren_s = ErnerfRenderer(blob=blob, cfg=cfg, device=local)
    ins_s = [ernerf_inputs(f % 290, H, W) for f in range(16 * world)]
    auds_s = [torch.from_numpy(i[2]).to(dev) for i in ins_s]
    enc_s = torch.empty(32, dtype=torch.float32, device=dev)

    def step_sharded(k):
        for j in range(world):                      # frames k*world .. k*world + world - 1 of the session; this rank owns one
            i = (k * world + j) % len(ins_s)
            ren_s.encode_audio(auds_s[i], out=enc_s)
            if j == rank:
                p_, intr_, _, eye_ = ins_s[i]
                ren_s.render(p_, intr_, H, W, None, eye_, out=out, enc_a=enc_s)

    shard_ms, _, _ = timed(step_sharded, args.steps, args.warmup)
    del ren_s

    value = world * args.steps / (total_ms / 1e3)
    e2e_value = world * args.steps / (e2e_ms / 1e3)
    heads = {}
    shared = {}
    if not args.no_wav2lip:
        heads["wav2lip"] = wav2lip_leg(args, dev, local, rank, world, flush, timed, peaks(), shared=shared)
        heads["wav2lip_256"] = wav2lip_leg(args, dev, local, rank, world, flush, timed, peaks(), S=256)
    if not args.no_musetalk:
        heads["musetalk"] = musetalk_leg(args, dev, local, rank, world, flush, timed, peaks(), shared=shared)
    if not args.no_musetalk and "musetalk_engine" in shared:
        try:
            heads["musetalk_8streams"] = musetalk_streams_leg(args, dev, local, rank, world, timed, shared)
        except Exception as e:                           # noqa: BLE001
            heads["musetalk_8streams"] = {"error": repr(e)[:300]}
    if not args.no_mixed and "wav2lip_blob" in shared and "musetalk_engine" in shared:
        heads["mixed_sessions"] = mixed_leg(args, dev, local, rank, world, flush, timed, shared, blob, cfg)
    plugin = None
    if not args.no_plugin and rank == 0:
        plugin = plugin_legs(args, dev, local, shared, blob, cfg, w2v)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    hm = float(np.mean(head_ms))
    hs = float(np.mean(head_samples))
    ach_gbs = hs * BYTES_PER_SAMPLE / (hm * 1e-3) / 1e9
    ach_tf = hs * FLOP_PER_SAMPLE / (hm * 1e-3) / 1e12
    # DRAM traffic + L1 sector rate of k_head from the newest `ncu --set full` capture of this kernel (scripts/gpu/ncu_head.sh writes
    # the json; copied to profiles/ with the summary of the same capture)
    import glob
    traffic, l1_view = None, None
    caps = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_k_head_traffic.json")))
    if caps:
        cap = json.load(open(caps[-1]))
        traffic = cap.get("dram_bytes_per_launch")
        sec, us = cap.get("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"), cap.get("gpu__time_duration.sum")
        if sec and us:
            clk = us * 1e-6 * 1.965e9                     # ncu runs with --clock-control none: boost clock
            l1_view = {"l1_global_load_sectors_per_launch": sec, "sectors_per_clk_per_sm": sec / clk / 148, "peak_sectors_per_clk_per_sm": 4.0,
                       "frac": sec / clk / 148 / 4.0, "l1_hit_pct": cap.get("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": cap.get("lts__t_sector_hit_rate.pct"),
                       "shared_load_wavefronts_pct_of_peak": cap.get("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed"),
                       "capture": os.path.basename(caps[-1]),
                       "note": "the tables are L2/L1-resident (DRAM traffic ~2.7 MB per launch): this is the ceiling the gathers actually face; "
                               "the HBM figure above is the conservative stand-in SURVEY 8(d) prescribes"}
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
        "config": {"workload": "ernerf_512x512_fullframe (BASELINE configs[3]; 262144 rays/frame, 1 frame per step per GPU)",
                   "checkpoint": "tests/golden/ernerf_ckpt_infer.npz (reference data/pretrained/ngp_kf.pth)",
                   "inputs": "real poses (data_kf.json), N(0,1) audio windows [8,44,16], white background",
                   "l2": "flushed between steps (256 MiB write), flush outside the per-step CUDA events",
                   "parallelism": f"{world} independent frame streams (one process per GPU), every head's packed blob NCCL-broadcast from rank 0 at init, no collective on the frame path"},
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(auds_pin[0].numel() * 4 + 64 + 20),
                "d2h_bytes_per_step": int(out_pin.numel()), "ms_per_step": e2e_ms / args.steps,
                "serial_value": world * args.steps / (e2e_serial_ms / 1e3),
                "note": "ErnerfRenderer.render_host_async: pinned auds in, pinned u8 frame out; the device->host copy of frame k runs on a copy "
                        "stream and overlaps the render of frame k+1, the stream waits for the copy of frame k-1 inside step k, every frame is "
                        "in host memory before the timed region ends; serial_value = H2D, render, D2H back to back on one stream (render_host)"},
        "p50_chunk_to_frame_ms": float(np.median(lat)),
        "rays_2048_per_frame": {"value": world * args.steps / (sub_ms / 1e3), "unit": "frames/s", "ms_per_step": sub_ms / args.steps,
                                "note": "SURVEY 8(d) config 4 (i): 2048 explicit rays (every 128th pixel of the 512x512 grid) per frame"},
        "ernerf_batched_sessions": {
            "workload": f"{FB} ErNeRF sessions of one avatar model per GPU, one 512x512 frame each per pass (mf_ernerf_render_batch: ONE k_head "
                        "launch per pass, the sessions' hit lists form one queue); bit-identical to single renders",
            "value": world * Kb * FB / (batch_ms / 1e3), "unit": "frames/s", "ms_per_pass": batch_ms / Kb, "ms_per_frame": batch_ms / Kb / FB,
            "e2e": {"value": world * Kb * FB / (batch_e2e_ms / 1e3), "unit": "frames/s",
                    "h2d_bytes_per_step": int(FB * (auds_pin[0].numel() * 4 + 84)), "d2h_bytes_per_step": int(outs_b_pin.numel())},
            "roofline": {"kernel": "k_head (4 frames per launch)", "bound": "hbm",
                         "achieved": float(np.mean(bh_samples)) * BYTES_PER_SAMPLE / (float(np.mean(bh_ms)) * 1e-3) / 1e9,
                         "peak": peaks()["hbm"], "unit": "GB/s",
                         "frac": float(np.mean(bh_samples)) * BYTES_PER_SAMPLE / (float(np.mean(bh_ms)) * 1e-3) / 1e9 / peaks()["hbm"],
                         "ms_per_launch": float(np.mean(bh_ms)), "samples_per_launch": float(np.mean(bh_samples)),
                         "tensor_view_tflops": float(np.mean(bh_samples)) * FLOP_PER_SAMPLE / (float(np.mean(bh_ms)) * 1e-3) / 1e12}},
        "ernerf_single_stream_sharded": {
            "workload": f"ONE ErNeRF session, frames round-robin over {world} rank(s), audio state followed on every rank (SURVEY 8e)",
            "value": world * args.steps / (shard_ms / 1e3), "unit": "frames/s", "ms_per_pass": shard_ms / args.steps},
        "plugin_level": plugin,
        "nerfasr_acoustic_model": asr,
        "gpu_launches": int(launches),
        "kernels_per_step": ["k_setup (8 audio-window CTAs + the ray pass)", "k_head", "k_torso_compose"],
        "clocks": sampler.result(),
        "roofline": {"kernel": "k_head", "bound": "hbm", "achieved": ach_gbs, "peak": pk["hbm"], "unit": "GB/s",
                     "frac": ach_gbs / pk["hbm"], "traffic": traffic, "l1_view": l1_view, "peak_source": pk["src"] + " burst",
                     "ms_per_launch": hm, "samples_per_launch": hs, "algorithmic_bytes_per_sample": BYTES_PER_SAMPLE,
                     "share_of_step": hm / (total_ms / args.steps),
                     "tensor_view": {"achieved": ach_tf, "peak": pk["tf"], "unit": "TFLOP/s", "frac": ach_tf / pk["tf"],
                                     "flop_per_sample": FLOP_PER_SAMPLE},
                     "note": "gathers are served from L1/L2 (tables 1.96 MB): the HBM figure is the conservative stand-in SURVEY 8(d) prescribes"},
        "wall_s_timed_region": wall,
        "heads": heads,
    }
    if not args.no_reference_gpu:
        for key, fn in (("wav2lip", lambda: torch_gpu_wav2lip(dev, flush)), ("musetalk", lambda: torch_gpu_musetalk(dev, flush)),
                        ("musetalk_whisper", lambda: torch_gpu_whisper(dev, flush))):
            hk = key.split("_")[0]
            if hk not in heads:
                continue
            try:
                r = fn()
            except Exception as e:                       # noqa: BLE001 -- a side baseline must not cost the bench line
                r = {"error": repr(e)[:300]}
            heads[hk]["torch_gpu" if key == hk else "whisper_torch_gpu"] = r
        try:
            line["reference_cuda"] = reference_cuda_leg(dev)
            if "value" in line["reference_cuda"]:
                line["reference_cuda"]["ours_e2e_over_reference"] = e2e_value / world / line["reference_cuda"]["value"]
        except Exception as e:                           # noqa: BLE001 -- a side baseline must not cost the bench line
            line["reference_cuda"] = {"error": repr(e)}
    if not args.no_cpu_baseline:
        if "wav2lip" in heads:
            import torch as _t
            from helpers import seeded_wav2lip_state as _sw, wav2lip_inputs as _wi
            from oracle import wav2lip_oracle as _O
            _t.set_num_threads(os.cpu_count())
            _m, _f = _wi(16)
            _sd = _sw(2)
            _O.infer(_sd, _m[:2], _f[:2])
            t0 = time.perf_counter()
            _O.infer(_sd, _m, _f)
            dt = time.perf_counter() - t0
            heads["wav2lip"]["cpu_baseline"] = {"value": 16 / dt, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                                                "sample": f"one batch of 16 frames through the fp32 PyTorch oracle (pinned on the reference nn.Module's golden output), {dt:.2f} s, network only (no paste)"}
            try:
                heads["wav2lip"]["cpu_plumbing_config1"] = config1_cpu_plumbing()
            except Exception as e:                       # noqa: BLE001 -- a reported side baseline must not cost the bench line
                heads["wav2lip"]["cpu_plumbing_config1"] = {"error": repr(e)}
        if "wav2lip_256" in heads:
            _m, _f = _wi(4, S=256)
            _sd = _sw(2, face_hw=256)
            _O.infer(_sd, _m[:1], _f[:1])
            t0 = time.perf_counter()
            _O.infer(_sd, _m, _f)
            dt = time.perf_counter() - t0
            heads["wav2lip_256"]["cpu_baseline"] = {"value": 4 / dt, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                                                    "sample": f"4 frames through the fp32 PyTorch restatement of the extended 256x256 generator, {dt:.2f} s, network only "
                                                              "(no reference implementation of this net exists)"}
        if "musetalk" in heads:
            import torch as _t
            from helpers import WHISPER_TINY as _WT, seeded_whisper_state as _sws, synthetic_speech as _ss
            from oracle import musetalk_oracle as _M, whisper_oracle as _WO
            _t.set_num_threads(os.cpu_count())
            _u, _v = _M.UNET_CFG, _M.VAE_CFG
            _usd, _vsd = _M.seeded_state(_M.unet_param_shapes(_u), 5), _M.seeded_state(_M.vae_decoder_param_shapes(_v), 6)
            _rng = np.random.default_rng(0)
            _lat = (_rng.standard_normal((1, 8, 32, 32)) * 0.9).astype(np.float32)
            _wh = _rng.standard_normal((1, 50, 384)).astype(np.float32)
            t0 = time.perf_counter()
            _M.infer(_usd, _vsd, _lat, _wh, _u, _v)
            t_net = time.perf_counter() - t0
            t0 = time.perf_counter()
            _WO.audio2feat(_sws(7), _ss(52 * 320, 0), _WT)
            t_wh = time.perf_counter() - t0
            heads["musetalk"]["cpu_baseline"] = {"value": 1.0 / (t_net + t_wh / 16), "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                                                 "sample": f"ONE frame through the fp32 PyTorch restatement of UNet + VAE decoder ({t_net:.2f} s) plus 1/16 of one "
                                                           f"52-chunk Whisper window through the oracle ({t_wh:.2f} s per window); diffusers is absent, so no reference CPU run exists"}
        fps, t, cores = oracle_sample_fps(128, frames=3)
        line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": f"128x128 sub-grid of the 512x512 ray grid (16384 of {RAYS} rays), full pipeline, "
                                          f"median of 3 ({t:.2f} s each), scaled to full frames; the reference has no CPU renderer"}
    print(json.dumps(line), file=sys.__stdout__, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
