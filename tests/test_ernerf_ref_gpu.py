"""GPU parity against the REFERENCE's own CUDA kernels (oracle/_ref, compiled unchanged from
/root/reference/ernerf/*/src for sm_100).  This is what pins the C oracle and the sm_100a
kernels to the reference: integer-valued outputs and fp32 paths are compared bit-exactly."""
import ctypes
import json
import os

import numpy as np
import pytest

import ref_ernerf
from helpers import GOLD, ernerf_inputs, load_ernerf_fixture

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def env():
    if not ref_ernerf.available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    from mere_fusion_b200._lib import Context, lib
    from oracle.ernerf_oracle import ErnerfOracle
    sd, md = load_ernerf_fixture()
    orc = ErnerfOracle(sd, md)
    gold = json.load(open(os.path.join(GOLD, "ernerf_level_scales.json")))   # device-evaluated exp2f
    orc.head_scales = np.array(gold["head"], np.float32)
    orc.torso_scales = np.array(gold["torso"], np.float32)
    return dict(sd=sd, ref=ref_ernerf.load(), lib=lib(), ctx=Context(0), orc=orc)


def P(t):
    return ctypes.c_void_p(t.data_ptr())


_KEEP = []


def cu(a):
    """host array -> cuda tensor, kept alive until the end of the test (kernels are asynchronous
    and only see raw pointers)"""
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    _KEEP.append(t)
    return t


@pytest.fixture(autouse=True)
def _release():
    yield
    torch.cuda.synchronize()
    _KEEP.clear()


def _setup(H=128, frame=0):
    from oracle import ernerf_oracle as O
    pose, intr, _, _ = ernerf_inputs(frame, H, H)
    ro, rd = O.get_rays(pose, intr, H, H)
    return cu(ro), cu(rd)


def test_near_far_and_march_vs_reference(env):
    rm = env["ref"]["_raymarching_face"]
    L, ctx = env["lib"], env["ctx"]
    ro, rd = _setup()
    N = ro.shape[0]
    aabb = cu(np.array([-1, -0.5, -1, 1, 0.5, 1], np.float32))
    n_r, f_r = torch.empty(N, device="cuda"), torch.empty(N, device="cuda")
    rm.near_far_from_aabb(ro, rd, aabb, N, 0.05, n_r, f_r)
    n_m, f_m = torch.empty(N, device="cuda"), torch.empty(N, device="cuda")
    assert L.mf_near_far_from_aabb(ctx.handle, P(ro), P(rd), P(aabb), N, 0.05, P(n_m), P(f_m), None) == 0
    torch.cuda.synchronize()
    assert torch.equal(n_r, n_m) and torch.equal(f_r, f_m)
    # and the C oracle
    from oracle import ernerf_oracle as O
    n_o, f_o = O.near_far_from_aabb(ro.cpu().numpy(), rd.cpu().numpy(), aabb.cpu().numpy(), 0.05)
    assert np.array_equal(n_o, n_r.cpu().numpy()) and np.array_equal(f_o, f_r.cpu().numpy())

    bit = cu(np.ascontiguousarray(env["sd"]["density_bitfield"], np.uint8))
    alive = torch.nonzero(n_r < 1e30).flatten().to(torch.int32)
    n_alive = alive.numel()
    assert n_alive > 1000
    for n_step in (1, 4, 8, 16):
        M = n_alive * n_step
        outs = []
        for which in ("ref", "mine"):
            xyzs, dirs, deltas = torch.zeros(M, 3, device="cuda"), torch.zeros(M, 3, device="cuda"), torch.zeros(M, 2, device="cuda")
            noises = torch.zeros(n_alive, device="cuda")
            if which == "ref":
                rm.march_rays(n_alive, n_step, alive, n_r, ro, rd, 1.0, 1 / 256, 16, 1, 128, bit, n_r, f_r, xyzs, dirs, deltas, noises)
            else:
                assert L.mf_march_rays(ctx.handle, n_alive, n_step, P(alive), P(n_r), P(ro), P(rd), 1.0, 1 / 256, 16, 1, 128,
                                       P(bit), P(n_r), P(f_r), P(xyzs), P(dirs), P(deltas), P(noises), None) == 0
            torch.cuda.synchronize()
            outs.append((xyzs, dirs, deltas))
        for a, b in zip(*outs):
            assert torch.equal(a, b)
        x_o, d_o, dl_o = O.march_rays(n_alive, n_step, alive.cpu().numpy(), n_r.cpu().numpy(), ro.cpu().numpy(),
                                      rd.cpu().numpy(), 1.0, bit.cpu().numpy(), 1, 128, n_r.cpu().numpy(),
                                      f_r.cpu().numpy(), -1, 1 / 256, 16)
        assert np.array_equal(x_o, outs[0][0].cpu().numpy())
        assert np.array_equal(dl_o, outs[0][2].cpu().numpy())


def test_composite_vs_reference(env):
    rm = env["ref"]["_raymarching_face"]
    L, ctx = env["lib"], env["ctx"]
    rng = np.random.default_rng(11)
    n_alive, n_step, N = 5000, 5, 9000
    alive = rng.permutation(N)[:n_alive].astype(np.int32)
    sig = np.exp(rng.standard_normal(n_alive * n_step) * 2 + 1).astype(np.float32)
    rgb = rng.random((n_alive * n_step, 3)).astype(np.float32)
    deltas = np.zeros((n_alive * n_step, 2), np.float32)
    deltas[:, 0] = 0.0270632939
    deltas[:, 1] = rng.random(n_alive * n_step) + 1
    deltas.reshape(n_alive, n_step, 2)[rng.random(n_alive) < 0.3, 3:, :] = 0
    amb = rng.random(n_alive * n_step).astype(np.float32)
    state0 = [rng.random(N).astype(np.float32) * 0.5, rng.random(N).astype(np.float32), rng.random((N, 3)).astype(np.float32),
              rng.random(N).astype(np.float32)]
    res = []
    for which in ("ref", "mine"):
        a, (ws, dp, im, rt) = cu(alive), [cu(s) for s in state0]
        z1, z2, z3 = torch.zeros(N, device="cuda"), torch.zeros(N, device="cuda"), torch.zeros(N, device="cuda")
        args = (cu(sig), cu(rgb), cu(deltas), cu(amb), cu(amb), cu(amb))
        if which == "ref":
            rm.composite_rays_triplane(n_alive, n_step, 1e-4, a, rt, *args, ws, dp, im, z1, z2, z3)
        else:
            assert L.mf_composite_rays_triplane(ctx.handle, n_alive, n_step, 1e-4, P(a), P(rt), *[P(t) for t in args],
                                                P(ws), P(dp), P(im), P(z1), P(z2), P(z3), None) == 0
        torch.cuda.synchronize()
        res.append((a, rt, ws, dp, im, z1, z2, z3))
    for x, y in zip(*res):
        assert torch.equal(x, y)


def test_encoders_vs_reference(env):
    ge, she, fe = env["ref"]["_grid_encoder"], env["ref"]["_sh_encoder"], env["ref"]["_freqencoder"]
    L, ctx, orc = env["lib"], env["ctx"], env["orc"]
    rng = np.random.default_rng(12)
    B = 20011
    x = cu(rng.random((B, 2)).astype(np.float32))
    # head plane, fp32, hash
    emb = cu(env["sd"]["encoder_yz.embeddings"].astype(np.float32))
    off = cu(env["sd"]["encoder_yz.offsets"].astype(np.int32))
    S = float(np.log2(orc.hs))
    o_r, o_m = torch.empty(12, B, 1, device="cuda"), torch.empty(12, B, 1, device="cuda")
    ge.grid_encode_forward(x, emb, off, o_r, B, 2, 1, 12, S, 64, None, 0, False)
    assert L.mf_grid_encode_forward(ctx.handle, P(x), P(emb), P(off), P(o_m), B, 2, 1, 12, S, 64, 0, 0, 0, None) == 0
    torch.cuda.synchronize()
    assert torch.equal(o_r, o_m)
    from oracle import ernerf_oracle as O
    o_o = O.grid_encode(x.cpu().numpy(), emb.cpu().numpy(), off.cpu().numpy(), orc.hs, 64, 0, scales=orc.head_scales)
    assert np.array_equal(o_o, o_r.cpu().numpy().transpose(1, 0, 2).reshape(B, 12))
    # torso, fp16, tiled
    emb = cu(env["sd"]["torso_encoder.embeddings"].astype(np.float16))
    off = cu(env["sd"]["torso_encoder.offsets"].astype(np.int32))
    S = float(np.log2(orc.ts))
    o_r = torch.empty(16, B, 2, device="cuda", dtype=torch.float16)
    o_m = torch.empty(16, B, 2, device="cuda", dtype=torch.float16)
    ge.grid_encode_forward(x, emb, off, o_r, B, 2, 2, 16, S, 16, None, 1, False)
    assert L.mf_grid_encode_forward(ctx.handle, P(x), P(emb), P(off), P(o_m), B, 2, 2, 16, S, 16, 1, 0, 1, None) == 0
    torch.cuda.synchronize()
    assert torch.equal(o_r, o_m)
    o_o = O.grid_encode(x.cpu().numpy(), emb.cpu().numpy(), off.cpu().numpy(), orc.ts, 16, 1, half=True, scales=orc.torso_scales)
    assert np.array_equal(o_o, o_r.cpu().numpy().transpose(1, 0, 2).reshape(B, 32))
    # SH degree 4
    d = rng.standard_normal((B, 3)).astype(np.float32)
    d = cu(d / np.linalg.norm(d, axis=1, keepdims=True))
    s_r, s_m = torch.empty(B, 16, device="cuda"), torch.empty(B, 16, device="cuda")
    she.sh_encode_forward(d, s_r, B, 3, 4, None)
    assert L.mf_sh_encode_forward(ctx.handle, P(d), P(s_m), B, 3, 4, None) == 0
    torch.cuda.synchronize()
    assert torch.equal(s_r, s_m)
    np.testing.assert_allclose(O.sh_encode4(d.cpu().numpy()), s_r.cpu().numpy(), rtol=0, atol=5e-7)
    # frequency
    for D, deg in ((2, 8), (6, 3)):
        xi = cu(rng.random((4097, D)).astype(np.float32) * 2 - 1)
        C = D + D * deg * 2
        f_r, f_m = torch.empty(4097, C, device="cuda"), torch.empty(4097, C, device="cuda")
        fe.freq_encode_forward(xi, 4097, D, deg, C, f_r)
        assert L.mf_freq_encode_forward(ctx.handle, P(xi), 4097, D, deg, C, P(f_m), None) == 0
        torch.cuda.synchronize()
        assert torch.equal(f_r, f_m)
        np.testing.assert_allclose(O.freq_encode(xi.cpu().numpy(), deg), f_r.cpu().numpy(), rtol=0, atol=2e-4)
