"""CPU checks of the packers' program structure (no GPU): the parallel-region tags the executor turns into fork-joins
(convnet_pack.ProgramBuilder.par; csrc/wav2lip.cu launch_direct), and the LayerNorm folding of the fused wav2vec2 stack."""
import struct

import numpy as np

from helpers import W2V_SMALL, seeded_w2v_state, seeded_wav2lip_state


def _ops(pb):
    return [struct.unpack("<28i", rec[:112]) for rec in pb.ops]


def test_wav2lip_program_runs_the_audio_encoder_beside_the_face_encoder():
    """audio_encoder (wav2lip.py:38-55) and face_encoder_blocks (:13-36) meet only in the decoder: two branches of ONE region, in
    front of an untagged decoder; no op of one branch writes a buffer the other touches"""
    from mere_fusion_b200.wav2lip_pack import pack_wav2lip
    _, pb = pack_wav2lip(seeded_wav2lip_state(1), nominal_batch=2, face_hw=96)
    tags = [(o[27] >> 8) & 0xff for o in _ops(pb)]
    n1, n2 = tags.count(1), tags.count(2)
    assert n1 == 13 and n2 >= 14                                           # 13 audio-encoder layers, the face-encoder layers
    assert tags == [1] * n1 + [2] * n2 + [0] * (len(tags) - n1 - n2)       # one contiguous region, then the decoder
    ops = _ops(pb)
    outs = {t: {(o[2], o[3]) for o, tg in zip(ops, tags) if tg == t} for t in (1, 2)}
    ins = {t: {(o[0], o[1]) for o, tg in zip(ops, tags) if tg == t} | {(o[4], o[5]) for o, tg in zip(ops, tags) if tg == t and o[4] >= 0}
           for t in (1, 2)}
    assert not ({b for b, _ in outs[1]} & ({b for b, _ in outs[2]} | {b for b, _ in ins[2]}))
    assert not ({b for b, _ in outs[2]} & {b for b, _ in ins[1]})


def test_wav2vec2_program_tags_the_positional_conv_groups_and_folds_layernorm():
    from mere_fusion_b200.wav2vec2_pack import pack_wav2vec2
    sd = seeded_w2v_state(3, W2V_SMALL)
    _, pb = pack_wav2vec2(sd, W2V_SMALL, 8960, fused_stack=True)
    ops = _ops(pb)
    tags = [(o[27] >> 8) & 0xff for o in ops]
    G = W2V_SMALL["pos_groups"]
    region = [t for t in tags if t]
    assert region == list(range(1, G + 1))                                 # one branch per group, contiguous
    i0 = tags.index(1)
    cg = W2V_SMALL["hidden"] // G
    assert [ops[i0 + g][1] for g in range(G)] == [g * cg for g in range(G)] == [ops[i0 + g][3] for g in range(G)]   # disjoint channel slices
    # the fused stack's image: LayerNorm affine folded into Wqkv / bqkv (exact algebra in fp64, then one rounding)
    stack = [o for o in ops if o[25] == 5]
    assert len(stack) == 1
    D, I, L = W2V_SMALL["hidden"], W2V_SMALL["inter"], W2V_SMALL["layers"]
    img = np.frombuffer(pb.tensors[stack[0][22]], np.uint8)
    per = (9 * D + I) * 4 + (4 * D * D + 2 * D * I) * 2
    assert img.size == L * per
    vec = img[:(9 * D + I) * 4].view(np.float32)
    assert np.all(vec[:D] == 1) and np.all(vec[D:2 * D] == 0)              # gamma / beta slots are neutral
    p = "wav2vec2.encoder.layers.0."
    wq = np.concatenate([sd[p + f"attention.{n}_proj.weight"] for n in ("q", "k", "v")]).astype(np.float64)
    bq = np.concatenate([sd[p + f"attention.{n}_proj.bias"] for n in ("q", "k", "v")]).astype(np.float64)
    g, b = sd[p + "layer_norm.weight"].astype(np.float64), sd[p + "layer_norm.bias"].astype(np.float64)
    assert np.allclose(vec[2 * D:5 * D], bq + wq @ b, rtol=1e-6, atol=1e-6)
    wbits = img[(9 * D + I) * 4:(9 * D + I) * 4 + 3 * D * D * 2].view(np.uint16).astype(np.uint32) << 16
    wpacked = wbits.view(np.float32).reshape(3 * D, D)
    ref = (wq * g[None, :]).astype(np.float32)
    assert np.allclose(wpacked, ref, rtol=2 ** -7, atol=1e-6)                # bf16 rounding of the folded weight
