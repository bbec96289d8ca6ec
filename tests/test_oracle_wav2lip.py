"""CPU: the Wav2Lip oracle restatement against the golden output of the REFERENCE nn.Module, and
the packer's geometry / FLOP bookkeeping."""
import os

import numpy as np

from helpers import GOLD, seeded_wav2lip_state, wav2lip_inputs


def test_oracle_matches_reference_module_golden():
    from oracle import wav2lip_oracle as O
    gold = np.load(os.path.join(GOLD, "wav2lip_golden_b2.npz"))["pred"]
    mel, faces = wav2lip_inputs(2)
    pred, u8 = O.infer(seeded_wav2lip_state(2), mel, faces)
    assert pred.shape == gold.shape == (2, 96, 96, 3)
    np.testing.assert_allclose(pred, gold, rtol=0, atol=2e-5)
    assert 0.2 < gold.std() < 0.3                      # the seeded net is not saturated


def test_oracle_mask_and_channel_order():
    from oracle import wav2lip_oracle as O
    faces = np.full((1, 96, 96, 3), 255, np.uint8)
    x = O.build_batch(faces).numpy()
    assert x.shape == (1, 6, 96, 96)
    assert x[0, :3, :48].min() == 1.0 and x[0, :3, 48:].max() == 0.0 and x[0, 3:].min() == 1.0


def test_packer_program_geometry():
    import struct
    from mere_fusion_b200.wav2lip_pack import pack_wav2lip
    # packing is pure host code but queries nothing from the library: runs without a GPU
    blob, pb = pack_wav2lip(seeded_wav2lip_state(2))
    assert pb.flops_per_sample == 7933968384          # 7.934 GFLOP / frame
    assert len(pb.ops) == 51 + 5 * 3                  # 51 convs; each stride-2 ConvTranspose = 4 parity classes
    magic, kind, ver, n = struct.unpack("<IIII", blob[:16].tobytes())
    assert magic == 0x3242464D and kind == 2 and n == 1 + 3 * len(pb.ops)
    cat_c = [pb.buffers[i][2] for i in range(7)]
    # wav2lip.py:57-81 skip-concat widths 80, 160, 320, 512, 768, 1024, 1024; the two that are not multiples of 64 carry zero pad
    # channels (never written, zero weights) so that the layers reading them run on the TMA-fed kernel
    assert cat_c == [128, 192, 320, 512, 768, 1024, 1024]
    _, pb256 = pack_wav2lip(seeded_wav2lip_state(2, face_hw=256), face_hw=256)
    assert len(pb256.ops) == 74 and abs(pb256.flops_per_sample / 1e9 - 55.879) < 1e-3      # ~56 GFLOP/frame
