"""Pack a reference Wav2Lip state_dict (wav2lip/models/wav2lip.py:12-85, loaded as in
lipreal.py:42-53) into the conv-net program blob of csrc/wav2lip.cu.

torch.cat((x, feats[-1]), dim=1) (wav2lip.py:107) is realised by letting the decoder block and
the matching encoder block write into disjoint channel ranges of one "cat" buffer.
"""
import numpy as np

from .convnet_pack import ProgramBuilder

# (cout, k, stride, pad, residual) per Conv2d of the reference
FACE_ENC = [[(16, 7, 1, 3, False)],
            [(32, 3, 2, 1, False), (32, 3, 1, 1, True), (32, 3, 1, 1, True)],
            [(64, 3, 2, 1, False), (64, 3, 1, 1, True), (64, 3, 1, 1, True), (64, 3, 1, 1, True)],
            [(128, 3, 2, 1, False), (128, 3, 1, 1, True), (128, 3, 1, 1, True)],
            [(256, 3, 2, 1, False), (256, 3, 1, 1, True), (256, 3, 1, 1, True)],
            [(512, 3, 2, 1, False), (512, 3, 1, 1, True)],
            [(512, 3, 1, 0, False), (512, 1, 1, 0, False)]]
AUDIO_ENC = [(32, 3, (1, 1), 1, False), (32, 3, (1, 1), 1, True), (32, 3, (1, 1), 1, True),
             (64, 3, (3, 1), 1, False), (64, 3, (1, 1), 1, True), (64, 3, (1, 1), 1, True),
             (128, 3, (3, 3), 1, False), (128, 3, (1, 1), 1, True), (128, 3, (1, 1), 1, True),
             (256, 3, (3, 2), 1, False), (256, 3, (1, 1), 1, True),
             (512, 3, (1, 1), 0, False), (512, 1, (1, 1), 0, False)]
# decoder blocks: first layer + residual convs; ('c' conv | 't' transpose, cout, k, stride, pad, out_pad)
FACE_DEC = [[("c", 512, 1, 1, 0, 0)],
            [("t", 512, 3, 1, 0, 0), ("r", 512)],
            [("t", 512, 3, 2, 1, 1), ("r", 512), ("r", 512)],
            [("t", 384, 3, 2, 1, 1), ("r", 384), ("r", 384)],
            [("t", 256, 3, 2, 1, 1), ("r", 256), ("r", 256)],
            [("t", 128, 3, 2, 1, 1), ("r", 128), ("r", 128)],
            [("t", 64, 3, 2, 1, 1), ("r", 64), ("r", 64)]]


# ---- the 256x256 EXTENSION (BASELINE configs[1]).  NOT a reference architecture: the
# reference generator only accepts 96x96 (a 256 input fails at the first skip-concat, wav2lip.py:107).  Same block vocabulary,
# one more stride-2 stage at 512 channels on each side, a 4x4 bottleneck conv / transposed conv instead of 3x3; the audio
# encoder and the output block (80 -> 32 -> 3) are unchanged.  Parity for this net is against our own fp32 restatement only.
FACE_ENC_256 = [[(16, 7, 1, 3, False)],
                [(32, 3, 2, 1, False), (32, 3, 1, 1, True), (32, 3, 1, 1, True)],                       # 128
                [(64, 3, 2, 1, False), (64, 3, 1, 1, True), (64, 3, 1, 1, True), (64, 3, 1, 1, True)],  # 64
                [(128, 3, 2, 1, False), (128, 3, 1, 1, True), (128, 3, 1, 1, True)],                    # 32
                [(256, 3, 2, 1, False), (256, 3, 1, 1, True), (256, 3, 1, 1, True)],                    # 16
                [(512, 3, 2, 1, False), (512, 3, 1, 1, True)],                                          # 8
                [(512, 3, 2, 1, False), (512, 3, 1, 1, True)],                                          # 4
                [(512, 4, 1, 0, False), (512, 1, 1, 0, False)]]                                         # 1
FACE_DEC_256 = [[("c", 512, 1, 1, 0, 0)],
                [("t", 512, 4, 1, 0, 0), ("r", 512)],                                                   # 4
                [("t", 512, 3, 2, 1, 1), ("r", 512), ("r", 512)],                                       # 8
                [("t", 512, 3, 2, 1, 1), ("r", 512), ("r", 512)],                                       # 16
                [("t", 384, 3, 2, 1, 1), ("r", 384), ("r", 384)],                                       # 32
                [("t", 256, 3, 2, 1, 1), ("r", 256), ("r", 256)],                                       # 64
                [("t", 128, 3, 2, 1, 1), ("r", 128), ("r", 128)],                                       # 128
                [("t", 64, 3, 2, 1, 1), ("r", 64), ("r", 64)]]                                          # 256
ARCHS = {96: (FACE_ENC, FACE_DEC), 256: (FACE_ENC_256, FACE_DEC_256)}


def wav2lip_param_shapes(face_hw=96):
    """name -> shape of every state_dict tensor of the generator for `face_hw` (96: the reference module,
    wav2lip/models/wav2lip.py:12-85; 256: the extension above), same key naming"""
    enc, dec = ARCHS[face_hw]
    shapes = {}

    def block(prefix, cin, cout, k, transpose=False):
        shapes[prefix + ".conv_block.0.weight"] = (cin, cout, k, k) if transpose else (cout, cin, k, k)
        shapes[prefix + ".conv_block.0.bias"] = (cout,)
        for n in ("weight", "bias", "running_mean", "running_var"):
            shapes[prefix + ".conv_block.1." + n] = (cout,)
        shapes[prefix + ".conv_block.1.num_batches_tracked"] = ()

    cin, enc_c = 6, []
    for i, blk in enumerate(enc):
        for j, (cout, k, s, p, r) in enumerate(blk):
            block(f"face_encoder_blocks.{i}.{j}", cin, cout, k)
            cin = cout
        enc_c.append(cin)
    cin = 1
    for j, (cout, k, st, p, r) in enumerate(AUDIO_ENC):
        block(f"audio_encoder.{j}", cin, cout, k)
        cin = cout
    for i, blk in enumerate(dec):
        for j, spec in enumerate(blk):
            cout = spec[1]
            k = spec[2] if spec[0] != "r" else 3
            block(f"face_decoder_blocks.{i}.{j}", cin, cout, k, transpose=spec[0] == "t")
            cin = cout
        cin += enc_c[len(enc) - 1 - i]
    block("output_block.0", cin, 32, 3)
    shapes["output_block.1.weight"] = (3, 32, 1, 1)
    shapes["output_block.1.bias"] = (3,)
    return shapes


def _np(v):
    return v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)


def strip_module_prefix(sd):
    """lipreal.py:47-50"""
    return {k.replace("module.", ""): v for k, v in sd.items()}


def pack_wav2lip(sd, nominal_batch=16, face_hw=96):
    sd = {k: _np(v) for k, v in strip_module_prefix(sd).items()}

    def block(prefix):
        cb = prefix + ".conv_block"
        return (sd[cb + ".0.weight"], sd[cb + ".0.bias"],
                (sd[cb + ".1.weight"], sd[cb + ".1.bias"], sd[cb + ".1.running_mean"], sd[cb + ".1.running_var"]))

    if face_hw not in ARCHS:
        raise ValueError(f"pack_wav2lip: no generator architecture for {face_hw}x{face_hw} crops (have {sorted(ARCHS)})")
    FACE_ENC, FACE_DEC = ARCHS[face_hw]
    n = len(FACE_ENC)
    pb = ProgramBuilder(nominal_batch)
    S = face_hw
    # spatial size of every encoder stage (96, 48, 24, 12, 6, 3, 1 for 96x96; 256, 128, ..., 4, 1 for 256x256)
    sizes = [S]
    for blk in FACE_ENC[1:]:
        _, k, s, p, _ = blk[0]
        sizes.append((sizes[-1] + 2 * p - k) // s + 1)
    enc_c = [blk[-1][0] for blk in FACE_ENC]                      # 16, 32, 64, 128, 256, 512, 512
    dec_c = [blk[0][1] for blk in FACE_DEC]                       # 512, 512, 512, 384, 256, 128, 64
    # cat buffer i holds [decoder block (n - 1 - i) output | encoder block i output] at resolution sizes[i].  Widths that are not
    # a multiple of 64 (80 and 160 channels at the two largest resolutions) get zero pad channels up to the next multiple: the
    # layers that read the whole buffer then run on the TMA-fed kernel (Cin % 64 == 0) with zero weights for the pad channels --
    # measured on the 256 net: 80 -> 32 @256^2 469 us and the four 160 -> 64 parity convs 382 us on the cp.async im2col kernel
    def cat_width(c):
        return c if c < 64 else (c + 63) // 64 * 64
    cat = [pb.buffer(sizes[i], sizes[i], cat_width(dec_c[n - 1 - i] + enc_c[i])) for i in range(n)]
    cat_off = [dec_c[n - 1 - i] for i in range(n)]

    in_face = pb.buffer(S, S, 8)
    in_mel = pb.buffer(80, 16, 8)
    pb.hdr.update(in_face_buf=in_face, in_mel_buf=in_mel, face_hw=S, mel_h=80, mel_w=16, out_hw=S)

    # ---- audio encoder (wav2lip.py:38-55).  It meets the face encoder only in the decoder: the two are the branches of one parallel
    # region (convnet_pack.ProgramBuilder.par), 13 small layers that used to run in front of the face encoder now run beside it
    pb.par = 1
    cur, coff, H, W = in_mel, 0, 80, 16
    for j, (cout, k, st, p, resid) in enumerate(AUDIO_ENC):
        w, b, bn = block(f"audio_encoder.{j}")
        Ho, Wo = (H + 2 * p - k) // st[0] + 1, (W + 2 * p - k) // st[1] + 1
        out = pb.buffer(Ho, Wo, cout)
        pb.conv(cur, coff, out, 0, w, b, bn, stride=st, padding=p, res=(cur, coff) if resid else None)
        cur, coff, H, W = out, 0, Ho, Wo
    audio_emb = cur                                               # [B, 1, 1, 512]

    # ---- face encoder (wav2lip.py:13-36); the last conv of block i lands in cat[i][..., cat_off[i]:]
    pb.par = 2
    cur, coff = in_face, 0
    for i, blk in enumerate(FACE_ENC):
        for j, (cout, k, s, p, resid) in enumerate(blk):
            w, b, bn = block(f"face_encoder_blocks.{i}.{j}")
            last = j == len(blk) - 1
            if last:
                out, ooff = cat[i], cat_off[i]
            else:
                Hi = pb.buffers[cur][0]
                Ho = (Hi + 2 * p - k) // s + 1
                # the 32-channel intermediates are allocated 64 wide (zero pad channels, as for the cat buffers above) so that the
                # stride-1 32 -> 32 residual convs of the second block read Cin = 64 and run on the TMA-fed kernel
                out, ooff = pb.buffer(Ho, Ho, 64 if cout == 32 else cout), 0
            wide_in = s == 1 and coff == 0 and w.shape[1] == 32 and pb.buffers[cur][2] == 64
            pb.conv(cur, coff, out, ooff, w, b, bn, stride=s, padding=p, res=(cur, coff) if resid else None,
                    cin_align=64 if wide_in else 8)
            cur, coff = out, ooff

    # ---- face decoder (wav2lip.py:57-81,102-112)
    pb.par = 0
    cur, coff = audio_emb, 0
    for i, blk in enumerate(FACE_DEC):
        tgt = cat[n - 1 - i]
        for j, spec in enumerate(blk):
            last = j == len(blk) - 1
            Ht = pb.buffers[tgt][0]
            if last:
                out, ooff = tgt, 0
            else:
                out, ooff = pb.buffer(Ht, Ht, spec[1]), 0
            w, b, bn = block(f"face_decoder_blocks.{i}.{j}")
            align = 64 if (j == 0 and i > 0 and pb.buffers[cur][2] % 64 == 0) else 8   # first layer of a block reads a whole cat buffer
            if spec[0] == "c":
                pb.conv(cur, coff, out, ooff, w, b, bn, stride=spec[3], padding=spec[4], cin_align=align)
            elif spec[0] == "t":
                pb.conv_transpose(cur, coff, out, ooff, w, b, bn, stride=spec[3], padding=spec[4], output_padding=spec[5],
                                  cin_align=align)
            else:
                pb.conv(cur, coff, out, ooff, w, b, bn, stride=1, padding=1, res=(cur, coff))
            cur, coff = out, ooff
        cur, coff = tgt, 0                                        # x = cat(x, feats[-1]): the whole cat buffer

    # ---- output block (wav2lip.py:83-85)
    w, b, bn = block("output_block.0")
    o1 = pb.buffer(S, S, 64)                                      # 32 channels + zero pad: the 1x1 head reads Cin = 64 (TMA kernel)
    pb.conv(cat[0], 0, o1, 0, w, b, bn, stride=1, padding=1, cin_align=64 if pb.buffers[cat[0]][2] % 64 == 0 else 8)
    pb.conv(o1, 0, -1, 0, sd["output_block.1.weight"], sd["output_block.1.bias"], None, stride=1, padding=0,
            relu=False, mode=1, cin_align=64)
    return pb.finish(), pb
