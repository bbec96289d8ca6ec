"""Host logic of the session scheduler: placement and cross-session request coalescing,
with a fake engine so that no GPU is needed.  The GPU side (coalesced Wav2Lip sessions == separate sessions) is
tests/test_scheduler_gpu.py."""
import threading

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from mere_fusion_b200.scheduler import HEADS, Placement, SessionScheduler, SharedEngine, mixed_session_heads


class FakeEngine:
    """frame i of the output depends on frame i of the inputs only, like the real engines"""

    def __init__(self, max_batch=64):
        self.max_batch = max_batch
        self.calls = []
        self.last_launches = 7

    def forward(self, a, b, out=None, out_f32=None, stream=None):
        assert a.shape[0] == b.shape[0] <= self.max_batch
        self.calls.append(int(a.shape[0]))
        v = a.reshape(a.shape[0], -1).sum(1)
        if out is not None:
            out.copy_((b.to(torch.int64) + v.to(torch.int64)[:, None, None, None]).clamp(0, 255).to(torch.uint8))
        if out_f32 is not None:
            out_f32.copy_(b.float() * 2 + v[:, None, None, None])
        return out


def _inputs(n, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randint(0, 5, (n, 1, 4, 4), generator=g).float(),
            torch.randint(0, 200, (n, 6, 6, 3), generator=g, dtype=torch.uint8))


def test_mixed_session_heads_config5():
    h = mixed_session_heads(64)
    assert (h.count("ernerf"), h.count("musetalk"), h.count("wav2lip")) == (22, 21, 21)
    pl = Placement(8, round_robin=True)
    for i, head in enumerate(h):
        pl.place(f"s{i}", head)
    for g in range(8):
        on = pl.on_gpu(g)
        assert len(on) == 8 and {hd for _, hd in on} == set(HEADS)                            # every GPU hosts all three heads


def test_placement_least_loaded_and_release():
    pl = Placement(2)
    assert pl.place("a", "musetalk") == 0
    assert pl.place("b", "wav2lip") == 1
    assert pl.place("c", "wav2lip") == 1            # GPU 1 still lighter than one MuseTalk session
    assert pl.place("a", "musetalk") == 0           # idempotent
    pl.release("a")
    assert pl.place("d", "ernerf") == 0
    with pytest.raises(ValueError):
        pl.place("x", "hubert")
    capped = Placement(1, max_sessions_per_gpu=2)
    capped.place(1, "wav2lip"), capped.place(2, "wav2lip")
    with pytest.raises(RuntimeError):               # app.py:75-77 "Maximum number of sessions reached"
        capped.place(3, "wav2lip")


def test_flush_coalesces_and_splits_correctly():
    eng = FakeEngine(max_batch=40)
    sh = SharedEngine(eng, threaded=False)
    ref = FakeEngine()
    reqs, want = [], []
    for s, n in enumerate((16, 16, 5, 16, 3)):
        a, b = _inputs(n, s)
        o = torch.empty_like(b)
        f = torch.empty(b.shape, dtype=torch.float32) if s % 2 == 0 else None
        reqs.append((sh.submit(a, b, out=o, out_f32=f), o, f))
        ro, rf = torch.empty_like(b), torch.empty(b.shape, dtype=torch.float32)
        ref.forward(a, b, out=ro, out_f32=rf)
        want.append((ro, rf))
    sh.flush()
    assert eng.calls == [37, 19]                    # greedy packing in arrival order, never above max_batch
    for (r, o, f), (ro, rf) in zip(reqs, want):
        assert sh.wait(r) is o
        assert torch.equal(o, ro)
        if f is not None:
            assert torch.equal(f, rf)
    assert (sh.batches, sh.requests, sh.frames) == (2, 5, 56)
    with pytest.raises(ValueError):
        sh.submit(*_inputs(41, 9))


def test_threaded_sessions_share_launches():
    eng = FakeEngine(max_batch=64)
    sh = SharedEngine(eng, window_ms=200.0)          # wide window: all four callers land in one batch
    outs, barrier = {}, threading.Barrier(4)

    def session(i):
        a, b = _inputs(16, 100 + i)
        barrier.wait()
        outs[i] = (sh.forward(a, b), a, b)

    th = [threading.Thread(target=session, args=(i,)) for i in range(4)]
    [t.start() for t in th]
    [t.join(timeout=30) for t in th]
    sh.shutdown()
    assert eng.calls == [64]                         # max_batch reached -> dispatched before the window expires
    ref = FakeEngine()
    for i in range(4):
        o, a, b = outs[i]
        ro = torch.empty_like(b)
        ref.forward(a, b, out=ro)
        assert torch.equal(o, ro)
    with pytest.raises(RuntimeError):
        sh.submit(*_inputs(1, 0))


def test_engine_errors_reach_every_caller():
    class Broken(FakeEngine):
        def forward(self, *a, **k):
            raise RuntimeError("mf_wav2lip_forward: -2")

    sh = SharedEngine(Broken(), threaded=False)
    r1, r2 = sh.submit(*_inputs(2, 1)), sh.submit(*_inputs(3, 2))
    sh.flush()
    for r in (r1, r2):
        with pytest.raises(RuntimeError, match="mf_wav2lip_forward"):
            sh.wait(r)


def test_scheduler_shares_one_engine_per_gpu_and_head():
    built = []

    def factory(g, max_batch):
        built.append((g, max_batch))
        return FakeEngine(max_batch)

    sc = SessionScheduler(n_gpus=2, sessions_per_engine=4, batch_size=16, round_robin=True, threaded=False)
    g0, e0 = sc.open("s0", "wav2lip", factory)
    g1, e1 = sc.open("s1", "wav2lip", factory)
    g2, e2 = sc.open("s2", "wav2lip", factory)
    g3, e3 = sc.open("s3", "ernerf")
    assert (g0, g1, g2, g3) == (0, 1, 0, 1) and e0 is e2 and e0 is not e1 and e3 is None
    assert built == [(0, 64), (1, 64)]
    sc.close("s0")
    assert (0, "wav2lip", None) in sc.engines()      # s2 still uses it
    sc.close("s2")
    assert (0, "wav2lip", None) not in sc.engines()
    sc.shutdown()


class FakeRenderer:
    calls = []

    def __init__(self, i):
        self.i, self.device, self.ctx = i, None, None

    @staticmethod
    def render_batch(rens, frames, outs=None):
        FakeRenderer.calls.append([r.i for r in rens])
        return [(r.i, f["eye"]) for r, f in zip(rens, frames)]


def test_ernerf_batcher_one_frame_per_session_at_most_four():
    from mere_fusion_b200.scheduler import ErnerfBatcher
    FakeRenderer.calls = []
    b = ErnerfBatcher(None, threaded=False)
    rs = [FakeRenderer(i) for i in range(6)]
    order = [0, 1, 0, 2, 3, 4, 5, 1]
    reqs = [b.submit(rs[i], dict(eye=j)) for j, i in enumerate(order)]
    b.flush()
    assert FakeRenderer.calls == [[0, 1, 2, 3], [0, 4, 5, 1]]      # a session's second frame waits for the next pass (EMA order)
    assert [b.wait(r) for r in reqs] == [(i, j) for j, i in enumerate(order)]
    assert (b.batches, b.frames) == (2, 8)
    proxy = b.wrap(rs[2])
    assert proxy.render(None, None, 8, 8, eye=0.5) == (2, 0.5) and proxy.i == 2      # same call shape as ErnerfRenderer.render
    b.shutdown()
    with pytest.raises(RuntimeError):
        b.submit(rs[0], dict(eye=0))


def test_scheduler_hands_ernerf_sessions_a_batched_renderer():
    sc = SessionScheduler(n_gpus=1, threaded=False)
    g, r0 = sc.open("n0", "ernerf", factory=lambda gpu, _: FakeRenderer(10))
    g, r1 = sc.open("n1", "ernerf", factory=lambda gpu, _: FakeRenderer(11))
    assert r0._b is r1._b and r0.renderer.i == 10 and r1.renderer.i == 11
    FakeRenderer.calls = []
    q0, q1 = r0._b.submit(r0.renderer, dict(eye=1)), r1._b.submit(r1.renderer, dict(eye=2))
    r0._b.flush()
    assert FakeRenderer.calls == [[10, 11]] and r0._b.wait(q1) == (11, 2)
    sc.close("n0"), sc.close("n1")
    assert not sc.engines()


def test_asr_batcher_coalesces_windows():
    from mere_fusion_b200.scheduler import AsrBatcher

    class FakeAsr:
        max_batch, n_samples, device = 4, 64, None
        calls = []

        def logits_batch(self, audio):
            FakeAsr.calls.append(int(audio.shape[0]))
            return audio[:, :6].reshape(audio.shape[0], 3, 2) * 2.0

    b = AsrBatcher(FakeAsr(), window_ms=300.0)
    outs, gate = {}, threading.Barrier(4)

    def session(i):
        frame = np.full(64, float(i), np.float32)
        gate.wait()
        outs[i] = b.feature_fn(frame)

    th = [threading.Thread(target=session, args=(i,)) for i in range(4)]
    [t.start() for t in th]
    [t.join(timeout=30) for t in th]
    b.shutdown()
    assert FakeAsr.calls == [4]                                   # one pass for the four sessions' windows
    for i in range(4):
        assert outs[i].shape == (3, 2) and float(outs[i][0, 0]) == 2.0 * i
    b2 = AsrBatcher(FakeAsr(), threaded=False)
    assert float(b2.feature_fn(np.ones(64, np.float32))[1, 1]) == 2.0 and (b2.batches, b2.windows) == (1, 1)


def test_musetalk_sessions_share_weights_but_not_a_bigger_batch():
    built = []
    sc = SessionScheduler(n_gpus=1, sessions_per_engine=4, batch_size=16, threaded=False)
    g, e0 = sc.open("m0", "musetalk", lambda gpu, mb: built.append(mb) or FakeEngine(mb))
    g, e1 = sc.open("m1", "musetalk", lambda gpu, mb: built.append(mb) or FakeEngine(mb))
    assert e0 is e1 and built == [16]
    ra, rb = e0.submit(*_inputs(16, 1)), e0.submit(*_inputs(16, 2))
    e0.flush()
    assert e0.engine.calls == [16, 16]              # two passes of 16, one set of weights
    sc.shutdown()
