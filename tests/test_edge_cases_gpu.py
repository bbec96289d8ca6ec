"""Edge cases of the hot path through the C ABI: empty / ragged / degenerate inputs (the reference's
own tests cover none of these, the kernels' index arithmetic does have corners): a camera that sees nothing, a frame
whose size is not a multiple of anything, ray lists shorter than a warp, empty batches, tiny and full-frame paste boxes."""
import ctypes

import numpy as np
import pytest

from helpers import ernerf_inputs, load_ernerf_fixture, psnr, seeded_wav2lip_state, wav2lip_inputs

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def env():
    from mere_fusion_b200._lib import lib
    from mere_fusion_b200.ernerf import ErnerfRenderer
    from oracle.ernerf_oracle import ErnerfOracle
    sd, md = load_ernerf_fixture()
    ren = ErnerfRenderer(sd, md, device=0)
    orc = ErnerfOracle(sd, md)
    for name, S, base, L in (("head_scales", ren.cfg.head_log2_scale, 64, 12), ("torso_scales", ren.cfg.torso_log2_scale, 16, 16)):
        buf = (ctypes.c_float * L)()
        assert lib().mf_grid_level_scales(ren.ctx.handle, S, base, L, buf) == 0
        setattr(orc, name, np.array(list(buf), np.float32))
    return ren, orc


def _render_pair(ren, orc, pose, intr, H, W, auds, eye):
    ren.reset()
    orc.enc_a_prev = None
    dbg_o = {}
    img_ref, u8_ref = orc.render_frame(pose, intr, H, W, auds, eye, debug=dbg_o)
    f32 = torch.empty(H, W, 3, device="cuda")
    a = torch.from_numpy(auds).cuda()
    out, dbg = ren.render(pose, intr, H, W, a, eye, out_f32=f32, debug=True)
    torch.cuda.synchronize()
    return f32.cpu().numpy(), out.cpu().numpy(), {k: v.cpu().numpy() for k, v in dbg.items()}, img_ref, u8_ref, dbg_o


def test_camera_that_sees_nothing(env):
    """the box is behind the camera: nothing is marched (the rays die in the first composite, raymarching.cu:2193), the frame is
    torso over background"""
    ren, orc = env
    H = W = 64
    pose, intr, auds, eye = ernerf_inputs(0, H, W)
    pose = pose.copy()
    fwd = pose[:3, 2] / np.linalg.norm(pose[:3, 2])                                  # camera z axis in world space (utils.py:300-304)
    pose[:3, 3] = 10.0 * fwd                                                          # far outside the box, looking away from it
    got, u8, dbg, ref, u8_ref, dbg_o = _render_pair(ren, orc, pose, intr, H, W, auds, eye)
    # the reference's slab test (raymarching.cu:91-145) does not reject a box BEHIND the ray: such rays keep near = min_near >
    # far < 0 and simply march nothing; rays that miss the slabs get near = far = FLT_MAX.  Same values on both sides:
    miss_o, miss_g = dbg_o["nears"] > 1e30, dbg["nears"] > 1e30
    assert (miss_o != miss_g).sum() <= 1 and (dbg_o["fars"][~miss_o] < 0).all()
    both = ~miss_o & ~miss_g
    np.testing.assert_allclose(dbg["fars"][both], dbg_o["fars"][both], rtol=1e-5, atol=1e-6)
    ri = dbg["round_info"]
    assert int(ri[0, 0]) == H * W and int(ri[0, 2]) == 0 and int(ri[1, 0]) == 0          # all alive, nothing emitted, nobody survives
    assert dbg_o["rounds"] == [(H * W, 1, 0)]
    assert (dbg["weights_sum"] == 0).all()
    assert np.array_equal(dbg["torso_mask"].astype(bool), dbg_o["torso_mask"])
    assert psnr(got, ref) >= 40.0
    assert (got.reshape(-1, 3)[dbg["torso_mask"] == 0] == 1.0).all()                     # bare background where there is no torso


@pytest.mark.parametrize("H,W", [(40, 72), (33, 17), (2, 50)])
def test_ragged_frame_sizes(env, H, W):
    """non-square frames whose ray count is not a multiple of the 32-sample tile (last tile partly empty, odd row length)"""
    ren, orc = env
    pose, intr, auds, eye = ernerf_inputs(5, 64, 64)
    fx = intr[0] * max(H, W) / 64
    intr = (fx, fx, W / 2.0, H / 2.0)
    got, u8, dbg, ref, u8_ref, dbg_o = _render_pair(ren, orc, pose, intr, H, W, auds, eye)
    assert got.shape == (H, W, 3)
    assert ((dbg["nears"] < 1e30) != (dbg_o["nears"] < 1e30)).sum() <= 1
    ri = dbg["round_info"]
    assert (int(ri[0, 0]), int(ri[0, 3])) == (H * W, 1)
    # in-kernel ray generation vs numpy: a few rays differ in the last bit of a direction and step into a neighbouring cell
    assert abs(int(ri[0, 2]) - dbg_o["rounds"][0][2]) <= max(4, dbg_o["rounds"][0][2] // 200)
    assert psnr(got, ref) >= 40.0
    assert np.percentile(np.abs(u8.astype(int) - u8_ref.astype(int)), 99) <= 2


@pytest.mark.parametrize("n", [1, 31, 33])
def test_ray_lists_shorter_than_a_tile(env, n):
    from oracle import ernerf_oracle as O
    ren, orc = env
    H = 64
    pose, intr, auds, eye = ernerf_inputs(2, H, H)
    ro, rd = O.get_rays(pose, intr, H, H)
    sub = (np.arange(n) * 97 + H * H // 2) % (H * H)                                   # rays through the head
    ro, rd = np.ascontiguousarray(ro[sub]), np.ascontiguousarray(rd[sub])
    bgc = np.ascontiguousarray(O.get_bg_coords(H, H)[sub])
    ren.reset()
    orc.enc_a_prev = None
    dbg_o = {}
    ref = orc.run_cuda(ro, rd, auds, bgc, pose, eye, np.ones((n, 3), np.float16), debug=dbg_o)
    keep = [torch.from_numpy(a).cuda() for a in (ro, rd, bgc, auds)]
    f32 = torch.empty(1, n, 3, device="cuda")
    out, dbg = ren.render(pose, intr, H, H, keep[3], eye, rays_o=keep[0], rays_d=keep[1], bg_coords=keep[2], out_f32=f32, debug=True)
    torch.cuda.synchronize()
    assert out.shape == (1, n, 3)
    assert np.array_equal(dbg["nears"].cpu().numpy(), dbg_o["nears"]) and np.array_equal(dbg["fars"].cpu().numpy(), dbg_o["fars"])
    ri = dbg["round_info"].cpu().numpy()
    assert (int(ri[0, 0]), int(ri[0, 3]), int(ri[0, 2])) == dbg_o["rounds"][0]
    assert np.abs(f32.cpu().numpy().reshape(-1, 3) - ref).max() < 2e-2


def test_empty_and_oversized_batches_are_refused():
    from mere_fusion_b200._lib import MfError, lib
    from mere_fusion_b200.wav2lip import Wav2LipEngine
    eng = Wav2LipEngine(seeded_wav2lip_state(2), max_batch=2, device=0)
    mel, faces = wav2lip_inputs(3)
    md, fd = torch.from_numpy(mel).cuda(), torch.from_numpy(faces).cuda()
    out = torch.empty_like(fd)
    h = eng.ctx.handle
    P = lambda t: ctypes.c_void_p(t.data_ptr())                                        # noqa: E731
    assert lib().mf_wav2lip_forward(h, P(md), P(fd), P(out), None, 0, None) == -1      # empty batch
    assert lib().mf_wav2lip_forward(h, P(md), P(fd), P(out), None, 3, None) == -1      # more than max_batch
    assert lib().mf_wav2lip_forward(h, None, P(fd), P(out), None, 1, None) == -1       # null input
    assert lib().mf_musetalk_forward(h, P(md), P(fd), P(out), None, 1, None) != 0      # a Wav2Lip program is not a MuseTalk program
    with pytest.raises(MfError):
        eng.forward(md, fd)
    eng.forward(md[:2], fd[:2], out=out[:2])                                           # and the context still works afterwards
    torch.cuda.synchronize()
    assert int(out[:2].float().std()) > 0


def test_paste_degenerate_boxes_bit_exact():
    """1-pixel, 2x3 and full-frame boxes, boxes touching every border: cv2.resize's border handling (x fraction zeroed, y index
    clamped) is where a restatement goes wrong first"""
    from mere_fusion_b200._lib import Context, lib
    from oracle.paste_oracle import paste_cv2
    ctx = Context(0)
    rng = np.random.default_rng(17)
    H, W, S, n = 97, 131, 96, 2
    frames = rng.integers(0, 256, (n, H, W, 3), dtype=np.uint8)
    boxes = [(0, 1, 0, 1), (H - 1, H, W - 1, W), (10, 12, 20, 23), (0, H, 0, W), (0, H, 5, 6), (40, 41, 0, W), (0, 96, 0, 96),
             (H - 96, H, W - 96, W), (3, 3 + 191, 1, 1 + 95)][:9]
    boxes = [b for b in boxes if b[1] <= H and b[3] <= W]
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
    empty = np.array([[0, 5, 5, 2, 9]], np.int32)                                       # y2 == y1: empty box
    assert lib().mf_paste_resize_u8(ctx.handle, ctypes.c_void_p(d_frames.data_ptr()), n, H, W, ctypes.c_void_p(d_faces.data_ptr()),
                                    S, 1, empty.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), ctypes.c_void_p(out.data_ptr()), None) == -1
    assert lib().mf_paste_resize_u8(ctx.handle, ctypes.c_void_p(d_frames.data_ptr()), n, H, W, ctypes.c_void_p(d_faces.data_ptr()),
                                    S, 0, rows.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), ctypes.c_void_p(out.data_ptr()), None) in (0, -1)
