"""CPU tests of the host-side plugin mirror (BaseASR / LipASR / LipReal / NerfASR plumbing):
BASELINE config 1 -- Wav2Lip 96x96, 25 fps, mel feats, a 10 s clip, no GPU."""
import os
import sys
import threading
import time
import types

import numpy as np
import pytest

from helpers import load_pose_fixture

REF = "/root/reference"
try:                                   # before any test stubs librosa (transformers probes for it at import time)
    import transformers.audio_utils as _TU
except Exception:                      # noqa: BLE001
    _TU = None


def _stub_module(name):
    """an empty stand-in for an absent third-party module (with a spec, so importlib.util.find_spec keeps working)"""
    import importlib.machinery
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        sys.modules[name] = m


def make_opt(**kw):
    o = types.SimpleNamespace(fps=50, l=10, m=8, r=10, batch_size=16, W=450, H=450, avatar_id="test", tts="none",
                              customopt=[], att=2, asr_model="cpierse/wav2vec2-large-xlsr-53-esperanto", exp_eye=True,
                              fix_eye=-1, fullbody=False, transport="rtc")
    o.__dict__.update(kw)
    return o


def clip_10s():
    """SURVEY.md 8(d) config 1 audio: 0.3 sin(2pi 220 t)(0.5 + 0.5 sin(2pi 3 t)) + 0.01 N(0,1), seed 0"""
    t = np.arange(160000) / 16000.0
    rng = np.random.default_rng(0)
    return (0.3 * np.sin(2 * np.pi * 220 * t) * (0.5 + 0.5 * np.sin(2 * np.pi * 3 * t)) + 0.01 * rng.standard_normal(160000)).astype(np.float32)


class FakeQueue:
    def __init__(self):
        self.items = []

    async def put(self, x):
        self.items.append(x)

    def qsize(self):
        return 0


class FakeTrack:
    def __init__(self):
        self._queue = FakeQueue()


def test_mel_shape_range_and_filterbank():
    from mere_fusion_b200 import audio_mel
    mel = audio_mel.melspectrogram(clip_10s()[:16640])
    assert mel.shape == (80, 16640 // 200 + 1)                 # [80, 84] for the 52-chunk window
    assert mel.min() >= -4.0 and mel.max() <= 4.0 and mel.std() > 0.5
    fb = audio_mel.mel_filterbank()
    assert fb.shape == (80, 401) and fb.dtype == np.float32 and (fb >= 0).all()
    peaks = fb.argmax(1)
    assert (np.diff(peaks) >= 0).all() and peaks[0] >= 2 and peaks[-1] <= 380     # 55 Hz .. 7600 Hz on a 20 Hz grid
    # silence maps to the floor: 20 log10(1e-5) - 20 = -120 dB -> clipped to -4
    assert np.allclose(audio_mel.melspectrogram(np.zeros(3200, np.float32)), -4.0)
    # a pure tone lights up the band that contains it
    tone = np.sin(2 * np.pi * 1000 * np.arange(16000) / 16000).astype(np.float32)
    band = audio_mel.melspectrogram(tone)[:, 20:60].mean(1).argmax()
    assert abs(int(fb[band].argmax()) * 20 - 1000) <= 60


def test_mel_chunk_slicing_matches_reference_lipasr():
    """the queue / window / slicing logic against the reference LipASR itself (librosa stubbed, the mel
    function injected), when the reference tree is present"""
    if not os.path.isdir(REF):
        pytest.skip("reference tree not present")
    from mere_fusion_b200 import audio_mel
    from mere_fusion_b200.plugin.lipasr import LipASR
    for name in ("librosa", "librosa.filters"):
        _stub_module(name)
    sys.path.insert(0, REF)
    try:
        import lipasr as ref_lipasr
        from wav2lip import audio as ref_audio
    finally:
        sys.path.remove(REF)
    ref_audio.melspectrogram = audio_mel.melspectrogram
    opt = make_opt()
    wav = clip_10s()
    outs = []
    for cls in (ref_lipasr.LipASR, LipASR):
        asr = cls(opt, None)
        for i in range(100):                                    # 2 s of audio, then silence fill
            asr.put_audio_frame(wav[i * 320:(i + 1) * 320])
        asr.warm_up()
        chunks = []
        for _ in range(4):
            asr.run_step()
            chunks.append(asr.feat_queue.get())
        types_ = []
        while True:
            try:
                types_.append(asr.output_queue.get(timeout=0.05)[1])
            except Exception:
                break
        outs.append((chunks, types_, len(asr.frames)))
    (c0, t0, n0), (c1, t1, n1) = outs
    assert n0 == n1 == 20 and t0 == t1
    assert len(c0) == len(c1) == 4
    for a, b in zip(c0, c1):
        assert len(a) == len(b) == 16
        for x, y in zip(a, b):
            assert x.shape == (80, 16) and np.array_equal(x, y)


def _fake_avatar(n=25, H=512, W=512):
    from mere_fusion_b200.plugin.lipreal import Avatar
    rng = np.random.default_rng(1)
    frames, faces, coords = [], [], []
    for i in range(n):
        f = rng.integers(0, 200, (H, W, 3), dtype=np.uint8)
        f[0, 0] = i                                             # index marker
        frames.append(f)
        faces.append(rng.integers(0, 256, (96, 96, 3), dtype=np.uint8))
        coords.append((176, 368, 160, 352))                     # (y1, y2, x1, x2), SURVEY 8(d) config 1
    return Avatar(frames, faces, coords)


def test_config1_lipreal_plumbing_10s_clip():
    from mere_fusion_b200.plugin.lipreal import LipReal, mirror_index

    class HostLipReal(LipReal):
        """the plumbing with a host stand-in for the engine: crop := 255 - face"""

        def infer_batch(self, mel_batch, index):
            assert len(mel_batch) == self.batch_size and mel_batch[0].shape == (80, 16)
            n = len(self.face_list_cycle)
            return [(255 - self.face_list_cycle[mirror_index(n, index + i)]).astype(np.float32) for i in range(self.batch_size)]

    opt = make_opt()
    real = HostLipReal(opt, engine=object(), avatar=_fake_avatar(), paste="cpu")
    wav = clip_10s()
    for i in range(500):
        real.put_audio_frame(wav[i * 320:(i + 1) * 320])
    quit_event = threading.Event()
    vt, at = FakeTrack(), FakeTrack()
    th = threading.Thread(target=real.render, args=(quit_event, None, at, vt), daemon=True)
    th.start()
    t0 = time.time()
    while len(vt._queue.items) < 272 and time.time() - t0 < 120:
        time.sleep(0.05)
    quit_event.set()
    th.join(timeout=30)
    nv, na = len(vt._queue.items), len(at._queue.items)
    assert nv >= 272
    assert abs(na - 2 * nv) <= 2                                # exactly two audio frames per video frame
    import cv2
    speech_frames = 0
    for k in range(250):
        fr = vt._queue.items[k].to_ndarray()
        assert fr.shape == (512, 512, 3) and fr.dtype == np.uint8
        idx = mirror_index(25, k)
        assert fr[0, 0, 0] == idx                               # ping-pong replay 0..24,24..0
        a0, a1 = at._queue.items[2 * k], at._queue.items[2 * k + 1]
        assert a0.samples == 320 and a0.sample_rate == 16000
        face = real.face_list_cycle[idx]
        expect = cv2.resize((255 - face), (192, 192))
        if np.array_equal(fr[176:368, 160:352], expect):
            speech_frames += 1
        else:
            assert np.array_equal(fr, real.frame_list_cycle[idx])   # idle frame: the untouched full frame
    # warm_up() ran on an empty queue (20 silent chunks, the first 10 dropped from the output side,
    # baseasr.py:53-59): output = 10 silent chunks (5 idle frames), then the 500 speech chunks
    assert speech_frames == 245
    pcm = np.concatenate([a.to_ndarray() for a in at._queue.items[:480]])
    ref = (wav * 32767).astype(np.int16)
    assert not pcm[:3200].any()
    assert np.array_equal(pcm[3200:3200 + 320 * 400], ref[:320 * 400])


def test_nerfasr_ring_and_window():
    import torch
    from mere_fusion_b200.plugin.nerfasr import NerfASR
    opt = make_opt()
    calls = []

    def feature_fn(frame):      # wav2vec2 geometry: 25 ms receptive field, 20 ms hop -> 27 rows for 28 chunks
        calls.append(len(frame))
        n = (len(frame) - 400) // 320 + 1
        base = float(len(calls))
        return torch.arange(n, dtype=torch.float32)[:, None].repeat(1, 44) + 100 * base

    asr = NerfASR(opt, None, feature_fn=feature_fn, device="cpu")
    asr.warm_up()                                               # 28 steps
    assert calls == [] or all(c == 28 * 320 for c in calls)
    for _ in range(16):
        asr.run_step()
    assert all(c == 28 * 320 for c in calls) and len(calls) >= 2
    # rows written per call: logits[l : T - r + 1] = rows 10..17 (nerfasr.py:139-141), at ring offset idx * m
    f = asr.get_next_feat()
    assert f.shape == (8, 44, 16)
    f2 = asr.get_next_feat()
    assert torch.equal(f2[:-1], f[1:])                           # the attention window slides by one
    assert asr.front == (32 - 8 + 2 * 5) % 32 and asr.tail == (8 + 2 * 5) % 32   # 4 + 1 windows consumed, 2 rows each


def test_pose_provider_matches_reference_golden():
    from mere_fusion_b200.ernerf_data import ErnerfPoseProvider, mirror_index
    pf = load_pose_fixture()
    tr = dict(cx=float(pf["cx"]), cy=float(pf["cy"]), focal_len=float(pf["focal_len"]),
              frames=[dict(transform_matrix=pf["raw"][i].tolist(), img_id=int(pf["img_id"][i])) for i in range(304)])
    au = np.zeros(int(pf["img_id"][:304].max()) + 1)
    au[:len(pf["au"])] = pf["au"][:len(au)]
    prov = ErnerfPoseProvider(tr, au)
    assert prov.H == 450 and prov.W == 450
    np.testing.assert_allclose(prov.poses[:296], pf["poses"][:296], rtol=0, atol=1e-6)
    np.testing.assert_allclose(prov.eye_area[:296], pf["eye"][:296], rtol=0, atol=1e-7)
    assert [mirror_index(3, i) for i in range(8)] == [0, 1, 2, 2, 1, 0, 0, 1]


# ------------------------------------------------------------------------------------------------
# MuseTalk plumbing (config 3 host side)
# ------------------------------------------------------------------------------------------------
class FakeAudioProcessor:
    """stands in for Audio2Feature: feature row t = mean |audio| of 20 ms chunk t, recognisable per row"""

    def audio2feat(self, audio):
        n = int((len(audio) // 160) / 2)
        rows = np.abs(np.asarray(audio[:n * 320], np.float32)).reshape(n, 320).mean(1)
        return np.broadcast_to(rows[:, None, None], (n, 5, 384)).astype(np.float32).copy()

    def feature2chunks(self, feature_array, fps, batch_size, audio_feat_length=[2, 2], start=0):
        from mere_fusion_b200.whisper import feature2chunks
        return feature2chunks(feature_array, fps, batch_size, audio_feat_length, start)


def test_museasr_matches_reference_museasr():
    """queue / window / slicing logic against the reference MuseASR itself (soundfile / ffmpeg stubbed, a fake audio processor
    injected in both), when the reference tree is present"""
    if not os.path.isdir(REF):
        pytest.skip("reference tree not present")
    from mere_fusion_b200.plugin.museasr import MuseASR
    for name in ("soundfile", "ffmpeg"):
        _stub_module(name)
    sys.path.insert(0, REF)
    try:
        import museasr as ref_museasr
        from musetalk.whisper.audio2feature import Audio2Feature as RefA2F
    finally:
        sys.path.remove(REF)
    ref_proc = RefA2F.__new__(RefA2F)
    ref_proc.audio2feat = FakeAudioProcessor().audio2feat            # the model call only; slicing stays the reference's
    opt = make_opt()
    wav = clip_10s()
    outs = []
    for cls, proc in ((ref_museasr.MuseASR, ref_proc), (MuseASR, FakeAudioProcessor())):
        asr = cls(opt, None, proc)
        for i in range(100):
            asr.put_audio_frame(wav[i * 320:(i + 1) * 320])
        asr.warm_up()
        chunks = []
        for _ in range(4):
            asr.run_step()
            chunks.append(asr.feat_queue.get())
        types_ = []
        while True:
            try:
                types_.append(asr.output_queue.get(timeout=0.05)[1])
            except Exception:
                break
        outs.append((chunks, types_, len(asr.frames)))
    (c0, t0, n0), (c1, t1, n1) = outs
    assert n0 == n1 == 20 and t0 == t1 and len(c0) == len(c1) == 4
    for a, b in zip(c0, c1):
        assert len(a) == len(b) == 16
        for x, y in zip(a, b):
            assert x.shape == (50, 384) and np.array_equal(x, y)


def _fake_muse_avatar(n=12, H=512, W=512):
    from mere_fusion_b200.plugin.musereal import MuseAvatar
    rng = np.random.default_rng(11)
    frames, coords, lat, masks, mcoords = [], [], [], [], []
    for i in range(n):
        f = rng.integers(0, 200, (H, W, 3), dtype=np.uint8)
        f[0, 0] = i
        frames.append(f)
        x1, y1 = 150 + i, 140 + 2 * i
        coords.append((x1, y1, x1 + 200 + i, y1 + 210))                         # (x1, y1, x2, y2), musereal.py:238
        xs, ys, xe, ye = x1 - 40, y1 - 30, x1 + 200 + i + 35, y1 + 210 + 45      # crop box around the face box
        mcoords.append((xs, ys, xe, ye))
        m = np.zeros((ye - ys, xe - xs, 3), np.uint8)
        yy, xx = np.mgrid[0:ye - ys, 0:xe - xs]
        ramp = np.clip(255 - 3 * np.hypot(yy - (ye - ys) / 2, xx - (xe - xs) / 2) + 200, 0, 255).astype(np.uint8)   # soft-edged blob
        m[:] = ramp[:, :, None]
        masks.append(m)
        lat.append((rng.standard_normal((1, 8, 32, 32)) * 0.18215 * 5).astype(np.float32))
    return MuseAvatar(frames, coords, lat, masks, mcoords)


def test_musereal_plumbing_host_blend():
    """MuseReal plumbing with a host stand-in for the engine: frame count, 2 audio frames per video frame, mirror indices,
    and the paste = get_image_blending (the reference's cv2 arithmetic) for speech frames"""
    from mere_fusion_b200.plugin.musereal import MuseReal
    from mere_fusion_b200.plugin.lipreal import mirror_index
    from oracle.paste_oracle import blend_cv2
    av = _fake_muse_avatar()

    class HostMuseReal(MuseReal):
        def infer_batch(self, whisper_chunks, index):
            assert len(whisper_chunks) == self.batch_size and whisper_chunks[0].shape == (50, 384)
            n = len(self.input_latent_list_cycle)
            return [np.full((256, 256, 3), 10 * mirror_index(n, index + i) + 5, np.float32) for i in range(self.batch_size)]

    real = HostMuseReal(make_opt(), engine=object(), audio_processor=FakeAudioProcessor(), avatar=av, paste="cpu")
    wav = clip_10s()
    for i in range(200):
        real.put_audio_frame(wav[i * 320:(i + 1) * 320])
    quit_event = threading.Event()
    vt, at = FakeTrack(), FakeTrack()
    th = threading.Thread(target=real.render, args=(quit_event, None, at, vt), daemon=True)
    th.start()
    t0 = time.time()
    while len(vt._queue.items) < 112 and time.time() - t0 < 120:
        time.sleep(0.05)
    quit_event.set()
    th.join(timeout=30)
    nv, na = len(vt._queue.items), len(at._queue.items)
    assert nv >= 112 and abs(na - 2 * nv) <= 2
    speech = 0
    for k in range(100):
        fr = vt._queue.items[k].to_ndarray()
        idx = mirror_index(12, k)
        assert fr[0, 0, 0] == idx
        face = np.full((256, 256, 3), 10 * idx + 5, np.uint8)
        expect = blend_cv2(av.frame_list_cycle[idx], face, av.coord_list_cycle[idx], av.mask_list_cycle[idx], av.mask_coords_list_cycle[idx])
        if np.array_equal(fr, expect):
            speech += 1
        else:
            assert np.array_equal(fr, av.frame_list_cycle[idx])
    assert speech == 95                                        # 5 idle frames from the silent warm-up, then speech


def test_blend_restatement_equals_cv2():
    from oracle.paste_oracle import blend_cv2, blend_numpy
    av = _fake_muse_avatar(4)
    rng = np.random.default_rng(5)
    for i in range(4):
        face = rng.integers(0, 256, (256, 256, 3), dtype=np.uint8)
        mask = av.mask_list_cycle[i].copy()
        mask[..., 1] = rng.integers(0, 256, mask.shape[:2], dtype=np.uint8)        # exercise the BGR2GRAY weights
        a = blend_cv2(av.frame_list_cycle[i], face, av.coord_list_cycle[i], mask, av.mask_coords_list_cycle[i])
        b = blend_numpy(av.frame_list_cycle[i], face, av.coord_list_cycle[i], mask, av.mask_coords_list_cycle[i])
        assert np.array_equal(a, b)


def test_mel_front_end_against_independent_librosa_compatible_implementation():
    """librosa itself is absent (requirements.txt:6, unpinned), so the restatement in audio_mel.py is cross-checked against an
    independent implementation written to reproduce librosa (transformers.audio_utils: slaney mel filterbank, centred STFT
    with constant padding, periodic Hann): filterbanks to 1e-8, STFT magnitudes to 1e-5 relative, the whole melspectrogram to
    1e-4.  Not the reference's own output (none exists offline) -- DESIGN.md keeps the row "parity unpinned"."""
    if _TU is None:
        pytest.skip("transformers.audio_utils not importable")
    tu = _TU
    from mere_fusion_b200 import audio_mel
    from mere_fusion_b200.whisper_pack import whisper_filters
    fb = tu.mel_filter_bank(num_frequency_bins=401, num_mel_filters=80, min_frequency=55, max_frequency=7600, sampling_rate=16000,
                            norm="slaney", mel_scale="slaney")
    assert np.abs(fb.T - audio_mel.mel_filterbank()).max() < 1e-8
    fbw = tu.mel_filter_bank(num_frequency_bins=201, num_mel_filters=80, min_frequency=0, max_frequency=8000, sampling_rate=16000,
                             norm="slaney", mel_scale="slaney")
    assert np.abs(fbw.T - whisper_filters()).max() < 1e-8
    wav = clip_10s()[:16640]
    pre = np.append(wav[0], wav[1:] - 0.97 * wav[:-1]).astype(np.float64)            # lfilter([1, -0.97], [1], wav)
    S = tu.spectrogram(pre, tu.window_function(800, "hann"), frame_length=800, hop_length=200, fft_length=800, power=1.0,
                       center=True, pad_mode="constant")
    D = np.abs(audio_mel.stft(pre))
    assert S.shape == D.shape == (401, 84) and np.abs(S - D).max() < 1e-5 * np.abs(D).max()
    mel_ref = 20 * np.log10(np.maximum(np.exp(-100 / 20 * np.log(10)), fb.T @ S)) - 20
    mel_ref = np.clip(8.0 * ((mel_ref + 100) / 100) - 4.0, -4.0, 4.0)
    assert np.abs(audio_mel.melspectrogram(wav) - mel_ref).max() < 1e-4


def test_mel_chunk_starts_agree_between_host_and_gpu_front_ends():
    """the host slicing (plugin.lipasr.mel_chunk_starts) and the start columns handed to mf_wav2lip_mel_chunks
    (wav2lip.MelFrontEnd.chunk_starts) are the same integers, and both equal the loop of lipasr.py:24-35 written out"""
    from mere_fusion_b200.plugin.lipasr import mel_chunk_starts
    from mere_fusion_b200.wav2lip import MelFrontEnd

    def reference_loop(n_frames, l, r, fps, n_cols):
        left = max(0, l * 80 / 50)
        mult = 80. * 2 / fps
        out, i = [], 0
        while i < (n_frames - l - r) / 2:
            s = int(left + i * mult)
            out.append(n_cols - 16 if s + 16 > n_cols else s)
            i += 1
        return out

    for fps in (50, 25, 40):
        for l, r in ((10, 10), (0, 0), (3, 7)):
            for B in (1, 4, 16):
                n_frames = l + 2 * B + r
                for n_cols in (n_frames * 320 // 200 + 1, 40, 17):
                    want = reference_loop(n_frames, l, r, fps, n_cols)
                    assert mel_chunk_starts(n_frames, l, r, fps, n_cols) == want
                    assert MelFrontEnd.chunk_starts(n_frames, l, r, fps, n_cols).tolist() == want
    assert mel_chunk_starts(15, 10, 10, 50, 80) == []          # not more than the context: no video frame yet


def test_ernerf_background_image_is_carried_over(tmp_path):
    """opt.bg_img (provider.py:203-214): 'white' = the renderer's default, 'black', or an image file (RGB, /255, INTER_AREA resize);
    opt.torso_imgs is refused loudly (the fused renderer always runs the torso model)"""
    import cv2
    from mere_fusion_b200.ernerf_data import ErnerfPoseProvider, load_bg_img
    assert load_bg_img("white", 8, 8) is None
    assert np.array_equal(load_bg_img("black", 4, 6), np.zeros((4, 6, 3), np.float32))
    img = np.random.default_rng(0).integers(0, 256, (20, 30, 3), dtype=np.uint8)
    p = str(tmp_path / "bg.png")
    cv2.imwrite(p, img)
    got = load_bg_img(p, 20, 30)
    assert np.array_equal(got, cv2.cvtColor(img, cv2.COLOR_BGR2RGB).astype(np.float32) / 255)
    small = load_bg_img(p, 10, 15)
    assert np.array_equal(small, cv2.cvtColor(cv2.resize(img, (15, 10), interpolation=cv2.INTER_AREA), cv2.COLOR_BGR2RGB).astype(np.float32) / 255)
    tr = dict(cx=4.0, cy=4.0, focal_len=10.0, frames=[dict(transform_matrix=np.eye(4).tolist(), img_id=0)] * 3)
    assert ErnerfPoseProvider(tr, None, bg_img="black").bg_img.shape == (8, 8, 3)
    with pytest.raises(NotImplementedError):
        ErnerfPoseProvider(tr, None, torso_imgs="data/torso")


def test_nerfreal_render_hands_frames_over_in_order_and_never_holds_one_back():
    """NeRFReal.render (nerfreal.py:129-156): the video frame of step k is finished while step k + 1 is in flight -- but only while more
    audio is waiting; with an empty input queue it is finished at once.  Every step yields exactly one video frame and two audio frames,
    in order, and the frame that is still pending when the loop is told to quit is flushed.  The GPU part is replaced by a stand-in."""
    from mere_fusion_b200.ernerf_data import ErnerfPoseProvider
    from mere_fusion_b200.plugin.nerfreal import NeRFReal

    class Ev:
        def __init__(self, log, k):
            self.log, self.k = log, k

        def synchronize(self):
            self.log.append(("sync", self.k))

    class Pin:
        def __init__(self, k):
            self.k = k

        def numpy(self):
            return np.full((8, 8, 3), self.k % 256, np.uint8)

    tr = dict(cx=4.0, cy=4.0, focal_len=10.0, frames=[dict(transform_matrix=np.eye(4).tolist(), img_id=0)] * 4)
    real = NeRFReal.__new__(NeRFReal)
    from mere_fusion_b200.plugin.basereal import BaseReal
    opt = make_opt(W=8, H=8)
    BaseReal.__init__(real, opt)
    real.W = real.H = 8
    real.provider = ErnerfPoseProvider(tr, None)
    log = []
    state = dict(k=0, waiting=6)      # audio chunks "waiting" in the input queue: 6, then none

    class Asr:
        class Q:
            def empty(self_q):
                return state["waiting"] <= 0
        queue = Q()

        def run_step(self):
            state["waiting"] -= 1

        def get_next_feat(self):
            return None

        def get_audio_out(self):
            return np.zeros(320, np.float32), 0

    real.asr = Asr()

    def fake_render_async(pose, eye, auds):
        k = state["k"]
        state["k"] += 1
        log.append(("issue", k))
        return Pin(k), Ev(log, k)
    real._render_async = fake_render_async
    quit_event = threading.Event()
    audio_track, video_track = FakeTrack(), FakeTrack()
    orig_put = video_track._queue.put

    async def put_and_count(x):
        await orig_put(x)
        log.append(("video", len(video_track._queue.items) - 1))
        if len(video_track._queue.items) == 5:
            quit_event.set()
    video_track._queue.put = put_and_count
    real.render(quit_event, None, audio_track, video_track)
    vids = video_track._queue.items
    assert len(vids) == state["k"] and len(audio_track._queue.items) == 2 * state["k"]            # one video + two audio frames per step, none lost
    assert [int(v.to_ndarray()[0, 0, 0]) for v in vids] == list(range(len(vids)))                   # in order
    ev = [e for e in log if e[0] != "sync"]
    # while audio was waiting (steps 0..2: two chunks per step) frame k went out only after step k + 1 had been issued ...
    assert ev[:5] == [("issue", 0), ("issue", 1), ("video", 0), ("issue", 2), ("video", 1)]
    # ... and once the queue was empty every frame was finished inside its own step
    tail = ev[5:]
    assert tail[:2] == [("video", 2), ("issue", 3)] or tail[:1] == [("video", 2)]
    for a, b in zip(tail, tail[1:]):
        if a[0] == "issue" and a[1] >= 3:
            assert b == ("video", a[1])
