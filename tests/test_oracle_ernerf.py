"""The network glue of oracle/ernerf_oracle.py against golden vectors produced by the reference's own nn.Modules
(tests/golden/make_ernerf_golden.py: AudioNet / AudioAttNet / MLP / NeRFNetwork.density / colour head, real checkpoint).

`*_ac` goldens ran under torch.autocast("cpu", fp16) -- the same rounding points as the cuda autocast of the live path
(SURVEY N7); the oracle must agree to fp16 resolution (differences come from the accumulation order inside the fp16
matmuls only).  `*_f32` goldens are the modules in plain fp32: they bound the whole fp16 emulation error."""
import os

import numpy as np
import pytest

from helpers import GOLD, load_ernerf_fixture
from oracle.ernerf_oracle import ErnerfOracle, affine16, mlp16, sigmoid16


@pytest.fixture(scope="module")
def G():
    return np.load(os.path.join(GOLD, "ernerf_glue_golden.npz"))


@pytest.fixture(scope="module")
def orc():
    sd, md = load_ernerf_fixture()
    return ErnerfOracle(sd, md)


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def test_encode_audio_vs_reference_modules(G, orc):
    for i in range(3):
        e = orc.encode_audio(G["auds"][i])
        assert e.shape == (1, 32) and e.dtype == np.float32
        assert rel(e, G["enc_a_ac"][i]) < 2e-3, rel(e, G["enc_a_ac"][i])
        assert rel(e, G["enc_a_f32"][i]) < 1e-2


def test_enc_a_smoothing_is_the_reference_recurrence(G):
    sd, md = load_ernerf_fixture()
    o = ErnerfOracle(sd, md)
    prev = None
    for i in range(3):                      # renderer.py:190-194 with the reference modules' own enc_a
        cur = G["enc_a_ac"][i]
        want = cur if prev is None else np.float32(0.35) * prev + np.float32(1 - 0.35) * cur
        got = o.smooth_enc_a(o.encode_audio(G["auds"][i]))
        assert rel(got, want) < 2e-3
        prev = want


def test_density_vs_reference_module(G, orc):
    enc_a = G["enc_a_ac"][0]
    sigma, geo, aud, eye_att = orc.density_from_enc(G["enc_x"], enc_a, float(G["eye"]))
    assert rel(geo, G["geo_ac"]) < 3e-3
    assert rel(np.log(sigma), np.log(G["sigma_ac"])) < 3e-3        # sigma = exp(h0): compare h0
    assert rel(eye_att, G["eye_att_ac"]) < 2e-3
    nrm = np.linalg.norm(aud.astype(np.float32), axis=-1, keepdims=True)
    assert rel(nrm, G["amb_aud_ac"]) < 3e-3
    assert rel(geo, G["geo_f32"]) < 2e-2 and rel(np.log(sigma), np.log(G["sigma_f32"])) < 2e-2


def test_color_head_vs_reference_module(G, orc):
    col = orc.color_from_enc(G["sh"], G["geo_ac"].astype(np.float16))
    assert np.abs(col.astype(np.float32) - G["color_ac"]).max() <= 2e-3      # fp16 values in (0,1): 1-2 ulp
    assert np.abs(col.astype(np.float32) - G["color_f32"]).max() <= 6e-3


def test_torso_mlps_vs_reference_modules(G, orc):
    dx = mlp16(G["torso_deform_in"], orc.w("torso_deform_net", 3))
    assert rel(dx, G["deform_ac"]) < 3e-3 and rel(dx, G["deform_f32"]) < 2e-2
    h = mlp16(G["torso_in"], orc.w("torso_net", 3))
    assert rel(h, G["torso_ac"]) < 3e-3 and rel(h, G["torso_f32"]) < 2e-2
    a = affine16(sigmoid16(h[:, :1]))
    assert a.dtype == np.float16 and float(a.min()) >= -0.00101 and float(a.max()) <= 1.0011
