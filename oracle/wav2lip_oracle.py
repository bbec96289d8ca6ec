"""oracle/wav2lip_oracle.py -- CPU restatement (PyTorch fp32, functional) of the reference's
Wav2Lip generator and of the batch build / post-processing around it.

TEST INFRASTRUCTURE ONLY: never imported by mere_fusion_b200; used by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.

Follows (paths relative to /root/reference):
  Conv2d / Conv2dTranspose blocks      wav2lip/models/conv.py:5-19,33-44 (conv -> BatchNorm(eval) -> [+x] -> ReLU)
  layer list and forward with skips    wav2lip/models/wav2lip.py:12-85, 87-125
  batch build (mask lower half, 6ch)   lipreal.py:108-122
  post (x255, uint8 truncation)        lipreal.py:126, 209

Parity pin: tests/test_oracle_wav2lip.py checks this restatement against a golden output produced
by the REFERENCE nn.Module itself (imported from /root/reference in the build container by
tests/golden/make_wav2lip_golden.py) on the seeded weights of tests/helpers.seeded_wav2lip_state.
No reference test or checkpoint pins the numerics further (./models/wav2lip.pth is not shipped).
"""
import numpy as np
import torch
import torch.nn.functional as F

FACE_ENC = [[(3, 1, 3, False)],
            [(3, 2, 1, False), (3, 1, 1, True), (3, 1, 1, True)],
            [(3, 2, 1, False), (3, 1, 1, True), (3, 1, 1, True), (3, 1, 1, True)],
            [(3, 2, 1, False), (3, 1, 1, True), (3, 1, 1, True)],
            [(3, 2, 1, False), (3, 1, 1, True), (3, 1, 1, True)],
            [(3, 2, 1, False), (3, 1, 1, True)],
            [(3, 1, 0, False), (1, 1, 0, False)]]
FACE_ENC[0] = [(7, 1, 3, False)]
AUDIO_ENC = [((1, 1), 1, False), ((1, 1), 1, True), ((1, 1), 1, True), ((3, 1), 1, False), ((1, 1), 1, True),
             ((1, 1), 1, True), ((3, 3), 1, False), ((1, 1), 1, True), ((1, 1), 1, True), ((3, 2), 1, False),
             ((1, 1), 1, True), ((1, 1), 0, False), ((1, 1), 0, False)]
# decoder: ('c', stride, pad) conv | ('t', stride, pad, out_pad) transpose | ('r',) residual 3x3
FACE_DEC = [[("c", 1, 0)], [("t", 1, 0, 0), ("r",)], [("t", 2, 1, 1), ("r",), ("r",)], [("t", 2, 1, 1), ("r",), ("r",)],
            [("t", 2, 1, 1), ("r",), ("r",)], [("t", 2, 1, 1), ("r",), ("r",)], [("t", 2, 1, 1), ("r",), ("r",)]]


# 256x256 EXTENSION (not a reference architecture): one more stride-2 512-channel stage on each
# side and a 4x4 bottleneck.  PARITY UNPINNED by construction -- this restatement is the only oracle for that net.
FACE_ENC_256 = FACE_ENC[:6] + [[(3, 2, 1, False), (3, 1, 1, True)], [(4, 1, 0, False), (1, 1, 0, False)]]
FACE_DEC_256 = [FACE_DEC[0], [("t", 1, 0, 0), ("r",)]] + [[("t", 2, 1, 1), ("r",), ("r",)]] * 6
ARCHS = {96: (FACE_ENC, FACE_DEC), 256: (FACE_ENC_256, FACE_DEC_256)}


def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.0, 1e-5)


def conv_block(sd, prefix, x, stride, padding, residual):
    """conv.py:5-19"""
    out = _bn(sd, prefix + ".conv_block.1", F.conv2d(x, sd[prefix + ".conv_block.0.weight"], sd[prefix + ".conv_block.0.bias"],
                                                    stride, padding))
    if residual:
        out = out + x
    return F.relu(out)


def convT_block(sd, prefix, x, stride, padding, output_padding):
    """conv.py:33-44"""
    out = F.conv_transpose2d(x, sd[prefix + ".conv_block.0.weight"], sd[prefix + ".conv_block.0.bias"], stride, padding,
                             output_padding)
    return F.relu(_bn(sd, prefix + ".conv_block.1", out))


def wav2lip_forward(sd, mel, img):
    """wav2lip.py:87-125 for 4-D inputs.  mel [B,1,80,16], img [B,6,96,96] fp32 -> [B,3,96,96] in (0,1)
    (a 256x256 img selects the extension above)"""
    FACE_ENC, FACE_DEC = ARCHS[int(img.shape[-1])]
    x = mel
    for j, (st, p, r) in enumerate(AUDIO_ENC):
        x = conv_block(sd, f"audio_encoder.{j}", x, st, p, r)
    audio_embedding = x
    feats = []
    x = img
    for i, blk in enumerate(FACE_ENC):
        for j, (k, s, p, r) in enumerate(blk):
            x = conv_block(sd, f"face_encoder_blocks.{i}.{j}", x, s, p, r)
        feats.append(x)
    x = audio_embedding
    for i, blk in enumerate(FACE_DEC):
        for j, spec in enumerate(blk):
            pre = f"face_decoder_blocks.{i}.{j}"
            if spec[0] == "c":
                x = conv_block(sd, pre, x, spec[1], spec[2], False)
            elif spec[0] == "t":
                x = convT_block(sd, pre, x, spec[1], spec[2], spec[3])
            else:
                x = conv_block(sd, pre, x, 1, 1, True)
        x = torch.cat((x, feats.pop()), dim=1)
    x = conv_block(sd, "output_block.0", x, 1, 1, False)
    x = F.conv2d(x, sd["output_block.1.weight"], sd["output_block.1.bias"])
    return torch.sigmoid(x)


def build_batch(faces_u8):
    """lipreal.py:108-122: faces u8 [B,S,S,3] BGR -> fp32 [B,6,S,S]"""
    img = np.asarray(faces_u8)
    masked = img.copy()
    masked[:, img.shape[1] // 2:] = 0
    x = np.concatenate((masked, img), axis=3) / 255.
    return torch.from_numpy(np.transpose(x, (0, 3, 1, 2)).astype(np.float32))


def infer(sd, mel, faces_u8):
    """lipreal.py:119-126,209: returns (pred fp32 [B,S,S,3] in (0,1), frames u8 [B,S,S,3])"""
    with torch.no_grad():
        pred = wav2lip_forward(sd, torch.as_tensor(mel, dtype=torch.float32), build_batch(faces_u8))
    pred = pred.cpu().numpy().transpose(0, 2, 3, 1)
    return pred, (pred * 255.).astype(np.uint8)
