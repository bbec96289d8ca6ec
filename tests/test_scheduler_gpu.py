"""GPU side of the session scheduler: same-GPU Wav2Lip sessions coalesced into ONE engine pass give
each session what it would have got alone (to the batch-size dependent split-K summation order, see
tests/test_wav2lip_gpu.py), through the C ABI, from the main thread and from concurrent session threads."""
import threading

import numpy as np
import pytest

from helpers import psnr, seeded_wav2lip_state, wav2lip_inputs

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def shared():
    from mere_fusion_b200.scheduler import SharedEngine
    from mere_fusion_b200.wav2lip import Wav2LipEngine
    sh = SharedEngine(Wav2LipEngine(seeded_wav2lip_state(2), max_batch=48, device=0), window_ms=50.0)
    yield sh
    sh.shutdown()


def _session_inputs(i, n=16):
    mel, faces = wav2lip_inputs(n, mel_seed=70 + i, face_seed=80 + i)
    return mel, faces


def test_coalesced_sessions_match_oracle_and_solo(shared):
    from oracle import wav2lip_oracle as O
    sd = seeded_wav2lip_state(2)
    ins = [_session_inputs(i) for i in range(3)]
    f32 = [torch.empty(16, 96, 96, 3, device="cuda") for _ in ins]
    b0 = shared.batches
    reqs = [shared.submit(torch.from_numpy(m).cuda(), torch.from_numpy(f).cuda(), out=torch.empty(16, 96, 96, 3, dtype=torch.uint8, device="cuda"),
                          out_f32=o) for (m, f), o in zip(ins, f32)]
    shared.flush()
    outs = [shared.wait(r) for r in reqs]
    torch.cuda.synchronize()
    assert shared.batches == b0 + 1                               # three sessions, ONE launch sequence
    for (m, f), o32, o8 in zip(ins, f32, outs):
        pred, u8 = O.infer(sd, m, f)
        assert psnr(o32.cpu().numpy(), pred) >= 38.0
        assert np.abs(o8.cpu().numpy().astype(int) - u8.astype(int)).mean() < 1.5
        solo = torch.empty(16, 96, 96, 3, device="cuda")
        shared.engine.forward(torch.from_numpy(m).cuda(), torch.from_numpy(f).cuda(), out_f32=solo)
        torch.cuda.synchronize()
        assert psnr(solo.cpu().numpy(), o32.cpu().numpy()) > 50.0


def test_concurrent_session_threads(shared):
    ins = [_session_inputs(10 + i) for i in range(3)]
    res, errs = {}, []
    gate = threading.Barrier(3)

    def session(i):
        try:
            with torch.cuda.stream(torch.cuda.Stream()):
                m, f = ins[i]
                md, fd = torch.from_numpy(m).cuda(), torch.from_numpy(f).cuda()
                o32 = torch.empty(16, 96, 96, 3, device="cuda")
                gate.wait()
                shared.forward(md, fd, out_f32=o32)
                torch.cuda.current_stream().synchronize()
                res[i] = o32.cpu().numpy()
        except Exception as e:                                    # noqa: BLE001
            errs.append(e)

    b0, r0 = shared.batches, shared.requests
    th = [threading.Thread(target=session, args=(i,)) for i in range(3)]
    [t.start() for t in th]
    [t.join(timeout=120) for t in th]
    assert not errs, errs
    assert shared.requests == r0 + 3 and shared.batches - b0 <= 3
    for i in range(3):
        m, f = ins[i]
        solo = torch.empty(16, 96, 96, 3, device="cuda")
        shared.engine.forward(torch.from_numpy(m).cuda(), torch.from_numpy(f).cuda(), out_f32=solo)
        torch.cuda.synchronize()
        assert psnr(solo.cpu().numpy(), res[i]) > 50.0


def test_two_lipreal_sessions_on_one_shared_engine():
    """plugin level: two LipReal sessions (threads) driven concurrently through ONE SharedEngine and packed avatars emit the
    frames a session on its own engine emits (pasted region equal to the batch-size dependent rounding, rest identical)"""
    from test_plugin_cpu import _fake_avatar, clip_10s, make_opt
    from test_plugin_gpu import _run
    from mere_fusion_b200.avatar_pack import DeviceAvatar, pack_lip_avatar
    from mere_fusion_b200.plugin.lipreal import LipReal
    from mere_fusion_b200.scheduler import SessionScheduler
    from mere_fusion_b200.wav2lip import Wav2LipEngine
    sd = seeded_wav2lip_state(2)
    wav = clip_10s()
    chunks = [wav[i * 320:(i + 1) * 320] for i in range(140)]
    solo = LipReal(make_opt(), engine=Wav2LipEngine(sd, max_batch=16, device=0), avatar=_fake_avatar())
    v_ref, _ = _run(solo, 48, chunks)
    ref = [f.to_ndarray().copy() for f in v_ref[:48]]

    sched = SessionScheduler(n_gpus=1, sessions_per_engine=2, batch_size=16, window_ms=5.0)
    packed = pack_lip_avatar(_fake_avatar())
    reals, res = [], {}
    for sid in ("a", "b"):
        g, eng = sched.open(sid, "wav2lip", factory=lambda gpu, mb: Wav2LipEngine(sd, max_batch=mb, device=gpu))
        assert g == 0 and eng.max_batch == 32
        reals.append(LipReal(make_opt(), engine=eng, avatar=DeviceAvatar(packed), device=g))
    assert reals[0].engine is reals[1].engine

    def drive(i):
        v, a = _run(reals[i], 48, chunks)
        res[i] = ([f.to_ndarray().copy() for f in v[:48]], len(v), len(a))

    th = [threading.Thread(target=drive, args=(i,)) for i in range(2)]
    [t.start() for t in th]
    [t.join(timeout=300) for t in th]
    shared = reals[0].engine
    assert shared.requests >= 2 * 3 and shared.batches <= shared.requests
    for i in range(2):
        frames, nv, na = res[i]
        assert len(frames) == 48 and abs(na - 2 * nv) <= 2
        for k, (f, r) in enumerate(zip(frames, ref)):
            outside = np.ones(f.shape[:2], bool)
            outside[176:368, 160:352] = False
            assert np.array_equal(f[outside], r[outside]), f"session {i} frame {k}"
            assert np.abs(f.astype(int) - r.astype(int)).max() <= 8 and np.abs(f.astype(int) - r.astype(int)).mean() < 0.1
    sched.close("a"), sched.close("b")
    assert not sched.engines()


def test_ernerf_batcher_threads_bit_identical():
    """three ErNeRF sessions (threads, own streams, own contexts on one shared blob) rendering 4 consecutive frames each through
    scheduler.ErnerfBatcher get exactly the images they get alone -- the EMA of the audio feature is per-session state"""
    from helpers import ernerf_inputs, load_ernerf_fixture
    from mere_fusion_b200.ernerf import ErnerfRenderer
    from mere_fusion_b200.scheduler import SessionScheduler
    sd, md = load_ernerf_fixture()
    base = ErnerfRenderer(sd, md, device=0)
    H = 96
    n_sess, n_frames = 3, 4
    ins = [[ernerf_inputs(10 * s + f, H, H) for f in range(n_frames)] for s in range(n_sess)]
    solo = []
    for s in range(n_sess):
        r = ErnerfRenderer(blob=base.blob, cfg=base.cfg, device=0)
        row = []
        for f in range(n_frames):
            p, intr, auds, eye = ins[s][f]
            row.append(r.render(p, intr, H, H, torch.from_numpy(auds).cuda(), eye).clone())
        solo.append(row)
    torch.cuda.synchronize()

    sched = SessionScheduler(n_gpus=1, window_ms=20.0)
    proxies = [sched.open(f"n{s}", "ernerf", factory=lambda g, _: ErnerfRenderer(blob=base.blob, cfg=base.cfg, device=g))[1] for s in range(n_sess)]
    batcher = proxies[0]._b
    res, errs, gate = {}, [], threading.Barrier(n_sess)

    def session(s):
        try:
            with torch.cuda.stream(torch.cuda.Stream()):
                got = []
                gate.wait()
                for f in range(n_frames):
                    p, intr, auds, eye = ins[s][f]
                    got.append(proxies[s].render(p, intr, H, H, torch.from_numpy(auds).cuda(), eye).clone())
                torch.cuda.current_stream().synchronize()
                res[s] = got
        except Exception as e:                                    # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=session, args=(s,)) for s in range(n_sess)]
    [t.start() for t in th]
    [t.join(timeout=120) for t in th]
    assert not errs, errs
    assert batcher.frames == n_sess * n_frames and batcher.batches < n_sess * n_frames      # some passes carried several sessions
    for s in range(n_sess):
        for f in range(n_frames):
            assert torch.equal(res[s][f], solo[s][f]), f"session {s} frame {f}"
    for s in range(n_sess):
        sched.close(f"n{s}")


def test_two_nerfreal_sessions_behind_the_batcher():
    """plugin level: two NeRFReal sessions whose renderers come from SessionScheduler.open(..., "ernerf") stream frames concurrently;
    the frames equal those of a session on a plain ErnerfRenderer fed the same audio features"""
    from helpers import load_ernerf_fixture, load_pose_fixture
    from test_plugin_cpu import clip_10s, make_opt
    from test_plugin_gpu import _run
    from mere_fusion_b200.ernerf import ErnerfRenderer
    from mere_fusion_b200.ernerf_data import ErnerfPoseProvider
    from mere_fusion_b200.plugin.nerfreal import NeRFReal
    from mere_fusion_b200.scheduler import SessionScheduler
    sd, md = load_ernerf_fixture()
    pf = load_pose_fixture()
    tr = dict(cx=float(pf["cx"]), cy=float(pf["cy"]), focal_len=float(pf["focal_len"]),
              frames=[dict(transform_matrix=pf["raw"][i].tolist(), img_id=int(pf["img_id"][i])) for i in range(40)])
    au = np.zeros(int(pf["img_id"][:40].max()) + 1)
    au[:min(len(au), len(pf["au"]))] = pf["au"][:len(au)]
    base = ErnerfRenderer(sd, md, device=0)
    wav = clip_10s()
    chunks = [wav[i * 320:(i + 1) * 320] for i in range(100)]

    def feature_fn_for(seed):
        rng = np.random.default_rng(seed)
        return lambda frame: torch.from_numpy(rng.standard_normal(((len(frame) - 400) // 320 + 1, 44)).astype(np.float32))

    def make(renderer, seed):
        return NeRFReal(make_opt(W=128, H=128), renderer, ErnerfPoseProvider(tr, au), feature_fn=feature_fn_for(seed), device=0)

    n = 24
    ref = []
    for s in range(2):
        v, a = _run(make(ErnerfRenderer(blob=base.blob, cfg=base.cfg, device=0), 100 + s), n, chunks)
        ref.append([f.to_ndarray().copy() for f in v[:n]])
    sched = SessionScheduler(n_gpus=1, window_ms=5.0)
    reals = [make(sched.open(f"n{s}", "ernerf", factory=lambda g, _: ErnerfRenderer(blob=base.blob, cfg=base.cfg, device=g))[1], 100 + s)
             for s in range(2)]
    res = {}

    def drive(s):
        v, a = _run(reals[s], n, chunks)
        res[s] = ([f.to_ndarray().copy() for f in v[:n]], len(v), len(a))

    th = [threading.Thread(target=drive, args=(s,)) for s in range(2)]
    [t.start() for t in th]
    [t.join(timeout=300) for t in th]
    batcher = reals[0].renderer._b
    assert batcher.frames >= 2 * n
    for s in range(2):
        frames, nv, na = res[s]
        assert len(frames) == n and abs(na - 2 * nv) <= 2
        for k, (f, r) in enumerate(zip(frames, ref[s])):
            assert np.array_equal(f, r), f"session {s} frame {k}"
    sched.close("n0"), sched.close("n1")


def test_asr_batcher_threads_on_the_gpu_engine():
    """three sessions' NerfASR feature_fn calls through scheduler.AsrBatcher (one mf_wav2vec2_logits_batch pass) return each session
    the logits of its own window"""
    from helpers import W2V_SMALL, seeded_w2v_state, synthetic_speech
    from mere_fusion_b200.scheduler import AsrBatcher
    from mere_fusion_b200.wav2vec2 import Wav2Vec2Engine
    eng = Wav2Vec2Engine(seeded_w2v_state(21, W2V_SMALL), W2V_SMALL, max_batch=4)
    wins = [synthetic_speech(8960, 30 + i) * (1 + i) for i in range(3)]
    solo = [eng.feature_fn(w).cpu().numpy() for w in wins]
    b = AsrBatcher(eng, window_ms=200.0)
    res, errs, gate = {}, [], threading.Barrier(3)

    def session(i):
        try:
            with torch.cuda.stream(torch.cuda.Stream()):
                gate.wait()
                out = b.feature_fn(wins[i])
                torch.cuda.current_stream().synchronize()
                res[i] = out.cpu().numpy()
        except Exception as e:                                    # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=session, args=(i,)) for i in range(3)]
    [t.start() for t in th]
    [t.join(timeout=120) for t in th]
    b.shutdown()
    assert not errs, errs
    assert b.windows == 3 and b.batches <= 2
    for i in range(3):
        rel = float(np.linalg.norm(res[i] - solo[i]) / np.linalg.norm(solo[i]))
        assert res[i].shape == (27, 44) and rel < 1.5e-2, (i, rel)
