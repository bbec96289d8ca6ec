"""GPU: paste-back kernel bit-exact vs cv2, and the LipReal / NeRFReal plugin objects end to end
through the C ABI with fake tracks."""
import ctypes
import threading
import time

import numpy as np
import pytest

from helpers import load_ernerf_fixture, load_pose_fixture, seeded_wav2lip_state
from test_plugin_cpu import FakeTrack, _fake_avatar, clip_10s, make_opt

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def test_paste_resize_bit_exact_vs_cv2():
    from mere_fusion_b200._lib import Context, lib
    from oracle.paste_oracle import paste_cv2
    ctx = Context(0)
    rng = np.random.default_rng(7)
    H, W, S, n = 300, 420, 96, 5
    frames = rng.integers(0, 256, (n, H, W, 3), dtype=np.uint8)
    boxes = [(10, 202, 20, 212), (0, 96, 0, 96), (50, 98, 60, 108), (3, 300, 1, 420), (100, 170, 200, 250),
             (120, 217, 33, 128), (7, 200, 300, 420)]       # 2x up, identity, exact 2x down (area), ragged, non-integer down
    B = len(boxes)
    faces = rng.integers(0, 256, (B, S, S, 3), dtype=np.uint8)
    rows = np.array([(i % n,) + b for i, b in enumerate(boxes)], np.int32)
    d_frames, d_faces = torch.from_numpy(frames).cuda(), torch.from_numpy(faces).cuda()
    out = torch.empty(B, H, W, 3, dtype=torch.uint8, device="cuda")
    rc = lib().mf_paste_resize_u8(ctx.handle, ctypes.c_void_p(d_frames.data_ptr()), n, H, W, ctypes.c_void_p(d_faces.data_ptr()),
                                  S, B, rows.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), ctypes.c_void_p(out.data_ptr()), None)
    assert rc == 0
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    for i, b in enumerate(boxes):
        assert np.array_equal(got[i], paste_cv2(frames[i % n], faces[i], b)), f"box {b}"
    bad = np.array([[0, 10, 400, 0, 50]], np.int32)           # y2 beyond the frame
    assert lib().mf_paste_resize_u8(ctx.handle, ctypes.c_void_p(d_frames.data_ptr()), n, H, W, ctypes.c_void_p(d_faces.data_ptr()),
                                    S, 1, bad.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), ctypes.c_void_p(out.data_ptr()), None) == -1


def _run(real, n_frames, chunks, timeout=120):
    quit_event = threading.Event()
    vt, at = FakeTrack(), FakeTrack()
    for c in chunks:
        real.put_audio_frame(c)
    th = threading.Thread(target=real.render, args=(quit_event, None, at, vt), daemon=True)
    th.start()
    t0 = time.time()
    while len(vt._queue.items) < n_frames and time.time() - t0 < timeout:
        time.sleep(0.02)
    quit_event.set()
    th.join(timeout=30)
    return vt._queue.items, at._queue.items


def test_lipreal_gpu_paste_equals_cpu_paste():
    """the same session rendered with the GPU paste kernel and with the reference's cv2 paste on the host
    gives identical frames; speech frames differ from the idle avatar frame inside the box only"""
    from mere_fusion_b200.plugin.lipreal import LipReal
    from mere_fusion_b200.wav2lip import Wav2LipEngine
    eng = Wav2LipEngine(seeded_wav2lip_state(2), max_batch=16, device=0)
    wav = clip_10s()
    chunks = [wav[i * 320:(i + 1) * 320] for i in range(200)]
    outs = []
    for mode in ("gpu", "cpu"):
        real = LipReal(make_opt(), engine=eng, avatar=_fake_avatar(), paste=mode)
        v, a = _run(real, 96, chunks)
        outs.append([f.to_ndarray().copy() for f in v[:96]])
        assert abs(len(a) - 2 * len(v)) <= 2
    for k, (g, c) in enumerate(zip(*outs)):
        assert np.array_equal(g, c), f"frame {k}"
    av = _fake_avatar()
    changed = 0
    for k in range(5, 96):
        idx = k if k < 25 else (49 - k if k < 50 else (k - 50 if k < 75 else 99 - k))
        f = outs[0][k]
        base = av.frame_list_cycle[idx]
        outside = np.ones((512, 512), bool)
        outside[176:368, 160:352] = False
        assert np.array_equal(f[outside], base[outside])
        changed += int(not np.array_equal(f, base))
    assert changed >= 85


def test_nerfreal_end_to_end():
    from mere_fusion_b200.ernerf import ErnerfRenderer
    from mere_fusion_b200.ernerf_data import ErnerfPoseProvider
    from mere_fusion_b200.plugin.nerfreal import NeRFReal
    sd, md = load_ernerf_fixture()
    pf = load_pose_fixture()
    tr = dict(cx=float(pf["cx"]), cy=float(pf["cy"]), focal_len=float(pf["focal_len"]),
              frames=[dict(transform_matrix=pf["raw"][i].tolist(), img_id=int(pf["img_id"][i])) for i in range(40)])
    au = np.zeros(int(pf["img_id"][:40].max()) + 1)
    au[:min(len(au), len(pf["au"]))] = pf["au"][:len(au)]
    prov = ErnerfPoseProvider(tr, au)
    ren = ErnerfRenderer(sd, md, device=0)
    rng = np.random.default_rng(0)

    def feature_fn(frame):
        return torch.from_numpy(rng.standard_normal(((len(frame) - 400) // 320 + 1, 44)).astype(np.float32))

    opt = make_opt(W=256, H=256)
    real = NeRFReal(opt, ren, prov, feature_fn=feature_fn, device=0)
    wav = clip_10s()
    v, a = _run(real, 30, [wav[i * 320:(i + 1) * 320] for i in range(100)])
    assert len(v) >= 30 and abs(len(a) - 2 * len(v)) <= 2
    img = v[10].to_ndarray()
    assert img.shape == (256, 256, 3) and img.dtype == np.uint8
    assert img[:20].mean() > 250 and 40 < img[100:200, 80:180].mean() < 230     # white background, a face in the middle
    assert not np.array_equal(v[10].to_ndarray(), v[20].to_ndarray())           # pose / audio advance


def test_paste_blend_bit_exact_vs_cv2():
    """mf_paste_blend_u8 == cv2.resize + get_image_blending (cvtColor + blendLinear) of the reference, bit for bit"""
    from mere_fusion_b200._lib import Context, lib
    from oracle.paste_oracle import blend_cv2
    from test_plugin_cpu import _fake_muse_avatar
    ctx = Context(0)
    av = _fake_muse_avatar(6)
    rng = np.random.default_rng(9)
    n, H, W, S = 6, 512, 512, 256
    B = 8
    faces = rng.integers(0, 256, (B, S, S, 3), dtype=np.uint8)
    masks, offs, off = [], [], 0
    for i in range(n):
        m = av.mask_list_cycle[i].copy()
        if i % 2:
            m[..., 1] = rng.integers(0, 256, m.shape[:2], dtype=np.uint8)
            m[..., 2] = rng.integers(0, 256, m.shape[:2], dtype=np.uint8)
        av.mask_list_cycle[i] = m
        offs.append(off)
        masks.append(m.reshape(-1))
        off += m.size
    rows = np.empty((B, 9), np.int32)
    moff = np.empty(B, np.int64)
    for i in range(B):
        k = (i * 5) % n
        x1, y1, x2, y2 = av.coord_list_cycle[k]
        xs, ys, xe, ye = av.mask_coords_list_cycle[k]
        rows[i] = (k, y1, y2, x1, x2, ys, ye, xs, xe)
        moff[i] = offs[k]
    d_frames = torch.from_numpy(np.stack(av.frame_list_cycle)).cuda()
    d_faces = torch.from_numpy(faces).cuda()
    d_masks = torch.from_numpy(np.concatenate(masks)).cuda()
    out = torch.empty(B, H, W, 3, dtype=torch.uint8, device="cuda")

    def call(r, mo, b):
        return lib().mf_paste_blend_u8(ctx.handle, ctypes.c_void_p(d_frames.data_ptr()), n, H, W, ctypes.c_void_p(d_faces.data_ptr()), S, b,
                                       r.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), ctypes.c_void_p(d_masks.data_ptr()), d_masks.numel(),
                                       mo.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), ctypes.c_void_p(out.data_ptr()), None)

    assert call(rows, moff, B) == 0
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    for i in range(B):
        k = int(rows[i, 0])
        ref = blend_cv2(av.frame_list_cycle[k], faces[i], av.coord_list_cycle[k], av.mask_list_cycle[k], av.mask_coords_list_cycle[k])
        assert np.array_equal(got[i], ref), f"item {i}: {(got[i] != ref).sum()} bytes differ"
    bad = rows[:1].copy()
    bad[0, 5] = bad[0, 1] + 1                                     # crop box starts below the face box
    assert call(bad, moff[:1], 1) == -1
    far = moff[:1].copy()
    far[0] = d_masks.numel() - 10                                 # mask would run past the buffer
    assert call(rows[:1], far, 1) == -1


def test_musereal_end_to_end_gpu_blend_equals_cpu_blend():
    """MuseReal on the GPU engines (Whisper features -> UNet + VAE -> blend) with fake tracks: the GPU blend path and the
    reference's host cv2 path give identical frames; speech frames differ from the avatar frame inside the crop box only"""
    from helpers import WHISPER_TINY, seeded_whisper_state
    from mere_fusion_b200.musetalk import MuseTalkEngine
    from mere_fusion_b200.plugin.musereal import MuseReal
    from mere_fusion_b200.plugin.lipreal import mirror_index
    from mere_fusion_b200.whisper import Audio2Feature, WhisperEngine
    from oracle import musetalk_oracle as M
    from test_plugin_cpu import _fake_muse_avatar
    u, v = M.small_cfgs()
    eng = MuseTalkEngine(M.seeded_state(M.unet_param_shapes(u), 5), M.seeded_state(M.vae_decoder_param_shapes(v), 6), u, v, max_batch=16)
    a2f = Audio2Feature(engine=WhisperEngine(seeded_whisper_state(7), WHISPER_TINY))
    wav = clip_10s()
    chunks = [wav[i * 320:(i + 1) * 320] for i in range(200)]
    outs = []
    for mode in ("gpu", "cpu"):
        real = MuseReal(make_opt(), engine=eng, audio_processor=a2f, avatar=_fake_muse_avatar(), paste=mode)
        vfr, afr = _run(real, 64, chunks)
        outs.append([f.to_ndarray().copy() for f in vfr[:64]])
        assert abs(len(afr) - 2 * len(vfr)) <= 2
    for k, (g, c) in enumerate(zip(*outs)):
        assert np.array_equal(g, c), f"frame {k}"
    av = _fake_muse_avatar()
    changed = 0
    for k in range(5, 64):
        idx = mirror_index(12, k)
        f, base = outs[0][k], av.frame_list_cycle[idx]
        xs, ys, xe, ye = av.mask_coords_list_cycle[idx]
        outside = np.ones((512, 512), bool)
        outside[ys:ye, xs:xe] = False
        assert np.array_equal(f[outside], base[outside])
        changed += int(not np.array_equal(f, base))
    assert changed >= 55


def test_wav2lip_mel_front_end_gpu_vs_host():
    """mf_wav2lip_mel_chunks (preemphasis + 800-point DFT + mel + dB + normalise + chunk slicing on the GPU) against the numpy
    restatement + the reference's slicing loop (audio_mel.melspectrogram, lipasr.mel_chunks); also the tail clamp"""
    from mere_fusion_b200 import audio_mel
    from mere_fusion_b200.plugin.lipasr import mel_chunks
    from mere_fusion_b200.wav2lip import ConvNet, MelFrontEnd
    from mere_fusion_b200.convnet_pack import ProgramBuilder
    pb = ProgramBuilder(1)
    a, b = pb.buffer(4, 4, 64), pb.buffer(4, 4, 64)
    pb.conv(a, 0, b, 0, np.zeros((64, 64, 1, 1), np.float32))
    fe = MelFrontEnd(ConvNet(pb.finish(), max_batch=1))          # any loaded context will do
    wav = clip_10s()
    for n_chunks, B in ((52, 16), (36, 8), (22, 1)):
        audio = wav[3000:3000 + n_chunks * 320]
        got = fe.chunks(audio, n_chunks, 10, 10, 50).cpu().numpy()
        mel = audio_mel.melspectrogram(audio)
        ref = np.stack(mel_chunks(mel, n_chunks, 10, 10, 50)).astype(np.float32)[:, None]
        assert got.shape == ref.shape == (B, 1, 80, 16)
        assert np.abs(got - ref).max() < 5e-3, np.abs(got - ref).max()
    # a window whose last chunk would run past the mel: clamped to the last 16 columns (lipasr.py:31-32)
    starts = MelFrontEnd.chunk_starts(52, 10, 10, 50, 60)
    assert starts.max() == 60 - 16 and len(starts) == 16
    assert list(MelFrontEnd.chunk_starts(52, 10, 10, 50, 84)[:4]) == [16, 19, 22, 25]
