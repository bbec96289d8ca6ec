"""Session scheduler: in-process replacement for the reference's one
`mp.Process` + pickled `mp.Queue` per session (lipreal.py:169-172, musereal.py:161-164; one model object per
`session_id`, app.py:45,87-93,331-345).

Two jobs, nothing else:
  * placement -- session -> GPU, least-loaded with a deterministic tie-break (8e: sessions are independent, no
    collective on the frame path; the only inter-GPU step is the weight broadcast at init, dist.py);
  * cross-session batching -- sessions of the same head on one GPU share ONE engine (one copy of the weights, one
    CUDA graph per batch size) and requests that arrive within a short window are coalesced into ONE
    mf_wav2lip_forward / mf_musetalk_forward launch sequence.  The conv executor is latency-bound at 16 frames
    (Wav2Lip: 68 launches of 10-25 us), so 2-4 coalesced sessions cost little more than one.

`SharedEngine` keeps the engine's call shape (`forward(a, b, out=..., out_f32=None)`, `.ctx`, `.device`), so a
`LipReal(opt, engine=shared)` / `MuseReal(opt, engine=shared)` session needs no change.  All C-ABI calls on the
shared context are issued by one dispatcher (a context is not thread-safe, include/mf_b200.h); `lock` serialises
the callers' own calls on that context (the paste kernels).
"""
import threading
import time
from collections import defaultdict

HEADS = ("ernerf", "musetalk", "wav2lip")


def mixed_session_heads(n_sessions):
    """heads interleaved (64 sessions -> 22 ErNeRF + 21 MuseTalk + 21 Wav2Lip), same order as
    dist.mixed_sessions"""
    from .dist import mixed_sessions
    return [h for h, _ in mixed_sessions(n_sessions)]


class Placement:
    """session -> GPU.  Least-loaded GPU by a per-head cost (ms of GPU time per 25 fps second of video, measured:
    profiles/), ties to the lowest index; `round_robin=True` reproduces config 5's fixed `session % n_gpu` layout."""

    COST = {"ernerf": 0.55 * 25, "musetalk": 20.5 / 16 * 25, "wav2lip": 1.0 / 16 * 25}

    def __init__(self, n_gpus, round_robin=False, max_sessions_per_gpu=None):
        assert n_gpus >= 1
        self.n_gpus = n_gpus
        self.round_robin = round_robin
        self.cap = max_sessions_per_gpu
        self.load = [0.0] * n_gpus
        self.sessions = {}                       # session_id -> (gpu, head)
        self._count = 0
        self._lock = threading.Lock()

    def place(self, session_id, head):
        if head not in self.COST:
            raise ValueError(f"unknown head {head!r}")
        with self._lock:
            if session_id in self.sessions:
                return self.sessions[session_id][0]
            n_on = [sum(1 for g, _ in self.sessions.values() if g == i) for i in range(self.n_gpus)]
            free = [i for i in range(self.n_gpus) if self.cap is None or n_on[i] < self.cap]
            if not free:
                raise RuntimeError("Maximum number of sessions reached")          # app.py:75-77
            if self.round_robin:
                g = self._count % self.n_gpus
                if g not in free:
                    g = free[0]
            else:
                g = min(free, key=lambda i: (self.load[i], i))
            self._count += 1
            self.load[g] += self.COST[head]
            self.sessions[session_id] = (g, head)
            return g

    def release(self, session_id):
        with self._lock:
            g, head = self.sessions.pop(session_id)
            self.load[g] -= self.COST[head]
            if not any(gg == g for gg, _ in self.sessions.values()):
                self.load[g] = 0.0                                                # no drift from float subtraction
            return g

    def on_gpu(self, g):
        with self._lock:
            return [(sid, h) for sid, (gg, h) in self.sessions.items() if gg == g]


class _Request:
    __slots__ = ("a", "b", "out", "out_f32", "n", "ready", "done", "error", "done_event")

    def __init__(self, a, b, out, out_f32):
        self.a, self.b, self.out, self.out_f32 = a, b, out, out_f32
        self.n = int(a.shape[0])
        self.ready = None                      # CUDA event: inputs complete on the caller's stream
        self.done = threading.Event()
        self.done_event = None                 # CUDA event: outputs complete on the dispatcher's stream
        self.error = None


class _Ticket:
    """one queued request of ErnerfBatcher / AsrBatcher: `a` = who / what (renderer, audio window), `b` = the frame arguments"""
    __slots__ = ("a", "b", "out", "ready", "done", "error", "done_event")

    def __init__(self, a, b=None, out=None):
        self.a, self.b, self.out = a, b, out
        self.ready = None
        self.done = threading.Event()
        self.done_event = None
        self.error = None


class SharedEngine:
    """one conv-net engine (Wav2LipEngine / MuseTalkEngine: `forward(a, b, out=, out_f32=)`, `.max_batch`) shared by
    the same-head sessions of a GPU, with request coalescing.

    forward() may be called from any number of session threads.  Requests are collected for at most `window_ms`
    after the first one arrives (or until `max_batch` frames are waiting) and run as one engine call on the
    concatenated batch; every caller's stream then waits for the batch's completion event, so the call is
    asynchronous exactly like the engine's own forward().  Frame i of a caller's output depends only on frame i
    of its inputs (eval-mode BatchNorm / GroupNorm per item), so coalescing does not change results beyond the
    batch-size dependent split-K summation order documented in tests/test_wav2lip_gpu.py.
    """

    def __init__(self, engine, window_ms=1.0, threaded=True):
        import torch
        self._torch = torch
        self.engine = engine
        self.ctx = getattr(engine, "ctx", None)
        self.device = getattr(engine, "device", None)
        self.max_batch = engine.max_batch
        self.window_s = window_ms * 1e-3
        self.lock = threading.RLock()            # callers' own C-ABI calls on the shared context (paste)
        self._pending = []
        self._cv = threading.Condition()
        self._stop = False
        self.batches = 0                         # engine calls issued
        self.requests = 0                        # forward() calls served
        self.frames = 0
        self._stage = {}
        self._cuda = bool(self.device is not None and getattr(self.device, "type", "cpu") == "cuda")
        self._stream = torch.cuda.Stream(self.device) if self._cuda else None
        self._thread = None
        if threaded:
            self._thread = threading.Thread(target=self._loop, name="mf-shared-engine", daemon=True)
            self._thread.start()

    # ---- the engine's own surface -----------------------------------------------------------------------------
    @property
    def last_launches(self):
        return self.engine.last_launches

    def __getattr__(self, name):                 # flops_per_frame, face_hw, ... of the wrapped engine
        if name in ("engine", "_torch"):
            raise AttributeError(name)
        return getattr(self.engine, name)

    def forward(self, a, b, out=None, out_f32=None, stream=None):
        req = self.submit(a, b, out, out_f32)
        if self._thread is None:
            self.flush()
        return self.wait(req)

    # ---- split submit / flush / wait: a single-threaded driver (bench.py config 5) uses the same coalescing path
    def submit(self, a, b, out=None, out_f32=None):
        torch = self._torch
        if int(a.shape[0]) > self.max_batch:
            raise ValueError(f"SharedEngine: batch {int(a.shape[0])} > max_batch {self.max_batch}")
        if out is None and out_f32 is None:
            out = self._default_out(a, b)
        req = _Request(a, b, out, out_f32)
        if self._cuda:
            req.ready = torch.cuda.Event()
            req.ready.record(torch.cuda.current_stream(self.device))
        with self._cv:
            if self._stop:
                raise RuntimeError("SharedEngine is shut down")
            self._pending.append(req)
            self._cv.notify_all()
        return req

    def wait(self, req):
        req.done.wait()
        if req.error is not None:
            raise req.error
        if self._cuda:
            cur = self._torch.cuda.current_stream(self.device)
            cur.wait_event(req.done_event)
            for t in (req.out, req.out_f32):     # allocated under the dispatcher's stream, consumed on the caller's: keep the allocator from
                if t is not None and hasattr(t, "record_stream"):   # recycling the block for the next batch while the caller still reads it
                    t.record_stream(cur)
        return req.out if req.out is not None else req.out_f32

    def flush(self):
        """run everything that is pending now (in arrival order, greedily packed into batches of <= max_batch)"""
        with self._cv:
            todo, self._pending = self._pending, []
        while todo:
            take, n = [], 0
            while todo and n + todo[0].n <= self.max_batch:
                n += todo[0].n
                take.append(todo.pop(0))
            self._run(take)

    def shutdown(self):
        with self._cv:
            self._stop = True
            self._cv.notify_all()
        if self._thread is not None:
            self._thread.join(timeout=5.0)
        self.flush()

    # ---- internals ----------------------------------------------------------------------------------------------
    def _default_out(self, a, b):
        torch = self._torch
        if b.dtype == torch.uint8:               # Wav2Lip: faces u8 [B,S,S,3] -> u8 [B,S,S,3]
            return torch.empty_like(b)
        return torch.empty((int(a.shape[0]), 256, 256, 3), dtype=torch.uint8, device=a.device)   # MuseTalk

    def _loop(self):
        while True:
            with self._cv:
                while not self._pending and not self._stop:
                    self._cv.wait()
                if self._stop and not self._pending:
                    return
                deadline = time.perf_counter() + self.window_s
                while sum(r.n for r in self._pending) < self.max_batch and not self._stop:
                    left = deadline - time.perf_counter()
                    if left <= 0:
                        break
                    self._cv.wait(left)
            self.flush()

    def _staging(self, key, like, n):
        t = self._stage.get(key)
        shape = (self.max_batch,) + tuple(like.shape[1:])
        if t is None or tuple(t.shape) != shape or t.dtype != like.dtype:
            t = self._stage[key] = self._torch.empty(shape, dtype=like.dtype, device=like.device)
        return t[:n]

    def _run(self, reqs):
        torch = self._torch
        try:
            ctxmgr = torch.cuda.stream(self._stream) if self._cuda else _null()
            with self.lock, ctxmgr:
                if self._cuda:
                    for r in reqs:
                        self._stream.wait_event(r.ready)
                if len(reqs) == 1:
                    r = reqs[0]
                    self.engine.forward(r.a, r.b, out=r.out, out_f32=r.out_f32)
                else:
                    n = sum(r.n for r in reqs)
                    A = self._staging("a", reqs[0].a, n)
                    Bt = self._staging("b", reqs[0].b, n)
                    torch.cat([r.a for r in reqs], 0, out=A)
                    torch.cat([r.b for r in reqs], 0, out=Bt)
                    want_u8 = any(r.out is not None for r in reqs)
                    want_f32 = any(r.out_f32 is not None for r in reqs)
                    O = self._staging("o", next(r.out for r in reqs if r.out is not None), n) if want_u8 else None
                    F = self._staging("f", next(r.out_f32 for r in reqs if r.out_f32 is not None), n) if want_f32 else None
                    self.engine.forward(A, Bt, out=O, out_f32=F)
                    o = 0
                    for r in reqs:
                        if r.out is not None:
                            r.out.copy_(O[o:o + r.n], non_blocking=True)
                        if r.out_f32 is not None:
                            r.out_f32.copy_(F[o:o + r.n], non_blocking=True)
                        o += r.n
                if self._cuda:
                    ev = torch.cuda.Event()
                    ev.record(self._stream)
                    for r in reqs:
                        r.done_event = ev
            self.batches += 1
            self.requests += len(reqs)
            self.frames += sum(r.n for r in reqs)
        except Exception as e:                   # noqa: BLE001 -- handed to every caller of this batch
            for r in reqs:
                r.error = e
        for r in reqs:
            r.done.set()


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _BatchedRenderer:
    """what a NeRFReal session holds instead of its ErnerfRenderer when the GPU's sessions are batched: same `render(...)`
    call, `.device`, `.ctx`, `.reset()`; the frame is rendered by ErnerfBatcher together with the other sessions' frames"""

    def __init__(self, batcher, renderer):
        self._b, self.renderer = batcher, renderer
        self.device, self.ctx = renderer.device, renderer.ctx

    def __getattr__(self, name):
        if name in ("renderer", "_b"):
            raise AttributeError(name)
        return getattr(self.renderer, name)

    def render(self, pose, intrinsics, H, W, auds=None, eye=0.25, out=None, outH=None, outW=None, enc_a=None, bg_color=None,
               out_f32=None, **unsupported):
        if unsupported:                      # explicit rays / debug taps: not batched, straight to the session's own renderer
            return self.renderer.render(pose, intrinsics, H, W, auds, eye, out=out, outH=outH, outW=outW, enc_a=enc_a,
                                        bg_color=bg_color, out_f32=out_f32, **unsupported)
        frame = dict(pose=pose, intrinsics=intrinsics, H=H, W=W, auds=auds, enc_a=enc_a, eye=eye, outH=outH, outW=outW,
                     bg_color=bg_color, out_f32=out_f32)
        req = self._b.submit(self.renderer, frame, out)
        if self._b._thread is None:
            self._b.flush()
        return self._b.wait(req)


class ErnerfBatcher:
    """ErNeRF sessions of one GPU that render the same avatar model: frames requested within `window_ms` of each other (at most
    one per session, at most 4) go through ONE mf_ernerf_render_batch pass -- the fused head kernel's round barriers are shared
    and the sessions fill each other's round tails (0.51 -> 0.41 ms per 512x512 frame at 4 sessions).  Images are bit-identical
    to unbatched renders; every session keeps its own context (EMA state, workspace)."""

    MAX_FRAMES = 4

    def __init__(self, device=None, window_ms=1.0, threaded=True):
        import torch
        self._torch = torch
        self.device = device
        self.window_s = window_ms * 1e-3
        self._cuda = device is not None and getattr(device, "type", "cpu") == "cuda"
        self._stream = torch.cuda.Stream(device) if self._cuda else None
        self._pending = []
        self._cv = threading.Condition()
        self._stop = False
        self.batches = 0
        self.frames = 0
        self._thread = None
        if threaded:
            self._thread = threading.Thread(target=self._loop, name="mf-ernerf-batcher", daemon=True)
            self._thread.start()

    def wrap(self, renderer):
        return _BatchedRenderer(self, renderer)

    def submit(self, renderer, frame, out=None):
        torch = self._torch
        req = _Ticket(renderer, frame, out)
        if self._cuda:
            req.ready = torch.cuda.Event()
            req.ready.record(torch.cuda.current_stream(self.device))
        with self._cv:
            if self._stop:
                raise RuntimeError("ErnerfBatcher is shut down")
            self._pending.append(req)
            self._cv.notify_all()
        return req

    def wait(self, req):
        req.done.wait()
        if req.error is not None:
            raise req.error
        if self._cuda:
            cur = self._torch.cuda.current_stream(self.device)
            cur.wait_event(req.done_event)
            if req.out is not None and hasattr(req.out, "record_stream"):
                req.out.record_stream(cur)
        return req.out

    @staticmethod
    def _model_key(renderer):
        """mf_ernerf_render_batch requires contexts loaded from the SAME device blob (pointer-identical tables): sessions whose
        renderers hold different copies of a model are never put into one pass"""
        blob = getattr(renderer, "blob", None)
        return blob.data_ptr() if hasattr(blob, "data_ptr") else 0

    def _take(self):
        """next batch: arrival order, one frame per session, at most MAX_FRAMES, all of the first request's model"""
        take, rest, seen, key = [], [], set(), None
        for r in self._pending:
            k = self._model_key(r.a)
            if len(take) < self.MAX_FRAMES and id(r.a) not in seen and (key is None or k == key):
                take.append(r)
                seen.add(id(r.a))
                key = k
            else:
                rest.append(r)
        self._pending = rest
        return take

    def flush(self):
        while True:
            with self._cv:
                take = self._take()
            if not take:
                return
            self._run(take)

    def shutdown(self):
        with self._cv:
            self._stop = True
            self._cv.notify_all()
        if self._thread is not None:
            self._thread.join(timeout=5.0)
        self.flush()

    def _loop(self):
        while True:
            with self._cv:
                while not self._pending and not self._stop:
                    self._cv.wait()
                if self._stop and not self._pending:
                    return
                deadline = time.perf_counter() + self.window_s
                while len({id(r.a) for r in self._pending}) < self.MAX_FRAMES and not self._stop:
                    left = deadline - time.perf_counter()
                    if left <= 0:
                        break
                    self._cv.wait(left)
            self.flush()

    def _run(self, reqs):
        torch = self._torch
        try:
            with (torch.cuda.stream(self._stream) if self._cuda else _null()):
                if self._cuda:
                    for r in reqs:
                        self._stream.wait_event(r.ready)
                rens = [r.a for r in reqs]
                outs = type(rens[0]).render_batch(rens, [r.b for r in reqs], outs=[r.out for r in reqs])
                for r, o in zip(reqs, outs):
                    r.out = o
                if self._cuda:
                    ev = torch.cuda.Event()
                    ev.record(self._stream)
                    for r in reqs:
                        r.done_event = ev
            self.batches += 1
            self.frames += len(reqs)
        except Exception as e:                   # noqa: BLE001 -- handed to every caller of this batch
            for r in reqs:
                r.error = e
        for r in reqs:
            r.done.set()


class AsrBatcher:
    """The wav2vec2 acoustic model behind NerfASR.run_step (nerfasr.py:105-143: one 28-chunk window every 8 chunks per session) shared by
    the ErNeRF sessions of a GPU: `feature_fn(frame)` calls that arrive within `window_ms` run as ONE mf_wav2vec2_logits_batch pass
    over the 630 MB of weights (Wav2Vec2Engine(max_batch=n)).  `feature_fn` keeps NerfASR's contract: float32[n_samples] host
    window in, device tensor [T, vocab] out."""

    def __init__(self, engine, window_ms=2.0, threaded=True):
        import torch
        self._torch = torch
        self.engine = engine
        self.max_batch = engine.max_batch
        self.device = getattr(engine, "device", None)
        self.window_s = window_ms * 1e-3
        self._cuda = self.device is not None and getattr(self.device, "type", "cpu") == "cuda"
        self._stream = torch.cuda.Stream(self.device) if self._cuda else None
        self._pending = []
        self._cv = threading.Condition()
        self._stop = False
        self.batches = 0
        self.windows = 0
        n = engine.n_samples
        self._pin = torch.empty((self.max_batch, n), dtype=torch.float32)
        if self._cuda:
            self._pin = self._pin.pin_memory()
        self._dev = torch.empty((self.max_batch, n), dtype=torch.float32, device=self.device) if self._cuda else self._pin
        self._thread = None
        if threaded:
            self._thread = threading.Thread(target=self._loop, name="mf-asr-batcher", daemon=True)
            self._thread.start()

    def feature_fn(self, frame):
        req = _Ticket(frame)
        with self._cv:
            if self._stop:
                raise RuntimeError("AsrBatcher is shut down")
            self._pending.append(req)
            self._cv.notify_all()
        if self._thread is None:
            self.flush()
        req.done.wait()
        if req.error is not None:
            raise req.error
        if self._cuda:
            self._torch.cuda.current_stream(self.device).wait_event(req.done_event)
        return req.out

    def flush(self):
        while True:
            with self._cv:
                take, self._pending = self._pending[:self.max_batch], self._pending[self.max_batch:]
            if not take:
                return
            self._run(take)

    def shutdown(self):
        with self._cv:
            self._stop = True
            self._cv.notify_all()
        if self._thread is not None:
            self._thread.join(timeout=5.0)
        self.flush()

    def _loop(self):
        while True:
            with self._cv:
                while not self._pending and not self._stop:
                    self._cv.wait()
                if self._stop and not self._pending:
                    return
                deadline = time.perf_counter() + self.window_s
                while len(self._pending) < self.max_batch and not self._stop:
                    left = deadline - time.perf_counter()
                    if left <= 0:
                        break
                    self._cv.wait(left)
            self.flush()

    def _run(self, reqs):
        torch = self._torch
        try:
            import numpy as np
            with (torch.cuda.stream(self._stream) if self._cuda else _null()):
                if self._cuda:
                    self._stream.synchronize()               # the pinned staging buffer of the previous pass has been consumed
                B = len(reqs)
                for i, r in enumerate(reqs):
                    self._pin[i].copy_(torch.from_numpy(np.ascontiguousarray(r.a, np.float32)))
                if self._cuda:
                    self._dev[:B].copy_(self._pin[:B], non_blocking=True)
                out = self.engine.logits_batch(self._dev[:B])
                ev = None
                if self._cuda:
                    ev = torch.cuda.Event()
                    ev.record(self._stream)
                for i, r in enumerate(reqs):
                    r.out, r.done_event = out[i], ev
            self.batches += 1
            self.windows += len(reqs)
        except Exception as e:                   # noqa: BLE001 -- handed to every caller of this batch
            for r in reqs:
                r.error = e
        for r in reqs:
            r.done.set()


class SessionScheduler:
    """placement + one SharedEngine per (GPU, head, model key).  `factory(device_index, max_batch)` builds the engine
    the first time a head is used on a GPU (weights: one blob per GPU, NCCL-broadcast by the caller when multi-process)."""

    def __init__(self, n_gpus=1, sessions_per_engine=4, batch_size=16, window_ms=1.0, round_robin=False, max_sessions_per_gpu=None,
                 threaded=True):
        self.placement = Placement(n_gpus, round_robin=round_robin, max_sessions_per_gpu=max_sessions_per_gpu)
        self.sessions_per_engine = sessions_per_engine
        self.batch_size = batch_size
        self.window_ms = window_ms
        self.threaded = threaded
        self._engines = {}
        self._refs = defaultdict(int)
        self._lock = threading.Lock()

    def open(self, session_id, head, factory=None, key=None):
        """-> (gpu index, engine).  Wav2Lip / MuseTalk: the GPU's SharedEngine for that head.  ErNeRF: the session's own renderer
        (private context: per-session EMA state and workspace) behind the GPU's ErnerfBatcher, which renders the frames that
        several sessions request at the same time in one mf_ernerf_render_batch pass."""
        g = self.placement.place(session_id, head)
        if factory is None:
            return g, None
        if head == "ernerf":                  # factory(gpu, 1) -> the session's own ErnerfRenderer (contexts share the device blob)
            kb = (g, "ernerf", key)
            with self._lock:
                if kb not in self._engines:
                    import torch
                    dev = torch.device("cuda", g) if torch.cuda.is_available() else None
                    self._engines[kb] = ErnerfBatcher(dev, window_ms=self.window_ms, threaded=self.threaded)
                self._refs[kb] += 1
                return g, self._engines[kb].wrap(factory(g, 1))
        k = (g, head, key)
        with self._lock:
            if k not in self._engines:
                # Wav2Lip is launch-latency bound at 16 frames: room for several sessions' batches in one pass.  MuseTalk fills the GPU
                # at 16 frames (tensor-bound, ~0.5 GB of activations per frame): its sessions share the weights and take turns
                eng = factory(g, self.batch_size * (self.sessions_per_engine if head == "wav2lip" else 1))
                self._engines[k] = SharedEngine(eng, window_ms=self.window_ms, threaded=self.threaded)
            self._refs[k] += 1
            return g, self._engines[k]

    def close(self, session_id, key=None):
        g, head = self.placement.sessions[session_id]
        self.placement.release(session_id)
        k = (g, head, key)
        with self._lock:
            if k in self._engines:
                self._refs[k] -= 1
                if self._refs[k] <= 0:
                    self._engines.pop(k).shutdown()
                    self._refs.pop(k, None)

    def engines(self):
        return dict(self._engines)

    def shutdown(self):
        with self._lock:
            for e in self._engines.values():
                e.shutdown()
            self._engines.clear()
            self._refs.clear()
