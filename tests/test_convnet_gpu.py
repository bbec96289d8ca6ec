"""GPU unit tests of the tcgen05 implicit-GEMM conv executor (csrc/wav2lip.cu) through the C ABI,
one op at a time, against torch.nn.functional on the bf16-rounded operands (fp32 math).
Tolerance: bf16 output rounding (2^-8 relative) plus fp32 accumulation-order noise."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
F = torch.nn.functional


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def run_case(B, Hin, Win, cin, cout, k, stride, pad, transpose=False, out_pad=0, residual=False, in_coff=0, in_C=None,
             out_coff=0, out_C=None, bn=True, relu=True, seed=0):
    from mere_fusion_b200.convnet_pack import ProgramBuilder
    from mere_fusion_b200.wav2lip import ConvNet
    g = torch.Generator().manual_seed(seed)
    in_C = in_C or (cin + 7) // 8 * 8
    x = torch.randn(B, Hin, Win, in_C, generator=g)
    if transpose:
        w = torch.randn(cin, cout, k, k, generator=g) / np.sqrt(cin * k * k / 4)
    else:
        w = torch.randn(cout, cin, k, k, generator=g) / np.sqrt(cin * k * k)
    bias = torch.randn(cout, generator=g) * 0.1
    bnp = None
    if bn:
        bnp = (torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1,
               torch.randn(cout, generator=g) * 0.2, torch.rand(cout, generator=g) + 0.5)
    xin = bf(x)[..., in_coff:in_coff + cin].permute(0, 3, 1, 2)
    if transpose:
        y = F.conv_transpose2d(xin, bf(w), None, stride, pad, out_pad)
    else:
        y = F.conv2d(xin, bf(w), None, stride, pad)
    Hout, Wout = y.shape[2:]
    pb = ProgramBuilder(nominal_batch=B)
    ib = pb.buffer(Hin, Win, in_C)
    out_C = out_C or cout
    ob = pb.buffer(Hout, Wout, out_C)
    npw = w.numpy()
    nbn = None if bnp is None else tuple(t.numpy() for t in bnp)
    if transpose:
        pb.conv_transpose(ib, in_coff, ob, out_coff, npw, bias.numpy(), nbn, stride=stride, padding=pad,
                          output_padding=out_pad, relu=relu)
    else:
        pb.conv(ib, in_coff, ob, out_coff, npw, bias.numpy(), nbn, stride=stride, padding=pad,
                res=(ib, in_coff) if residual else None, relu=relu)
    # reference epilogue in fp32: BatchNorm(eval) folded exactly like the packer does
    from mere_fusion_b200.convnet_pack import bn_fold
    sc, sh = bn_fold(bias.numpy(), nbn, cout)
    y = y * torch.from_numpy(sc).view(1, -1, 1, 1) + torch.from_numpy(sh).view(1, -1, 1, 1)
    if residual:
        y = y + xin
    if relu:
        y = F.relu(y)
    net = ConvNet(pb.finish(), max_batch=B)
    got = net.debug_run(ib, x, ob, (B, Hout, Wout, out_C)).cpu()
    torch.cuda.synchronize()
    got_c = got[..., out_coff:out_coff + cout].permute(0, 3, 1, 2)
    err = (got_c - y).abs()
    tol = 1e-2 + 1e-2 * y.abs()
    assert bool((err <= tol).all()), f"max err {err.max().item():.4f} at |y| max {y.abs().max().item():.3f}"
    # untouched channels of a wider (concat) buffer stay zero
    if out_C != cout:
        mask = torch.ones(out_C, dtype=torch.bool)
        mask[out_coff:out_coff + cout] = False
        assert float(got[..., mask].abs().max()) == 0.0
    return float(err.max())


def test_gemm_1x1():
    run_case(2, 16, 16, 64, 64, 1, 1, 0)             # plain GEMM: M=512, K=64, N=64


def test_gemm_1x1_bn_variants():
    for cout in (16, 32, 128, 256):
        run_case(3, 10, 10, 128, cout, 1, 1, 0, seed=cout)   # ragged M = 300


def test_conv3x3_pad_residual():
    run_case(2, 24, 24, 64, 64, 3, 1, 1, residual=True)


def test_conv3x3_stride2_and_rect_stride():
    run_case(2, 48, 48, 32, 64, 3, 2, 1)
    run_case(2, 80, 16, 32, 64, 3, (3, 1), 1)
    run_case(2, 9, 6, 128, 256, 3, (3, 2), 1)


def test_conv7x7_padded_input_channels():
    run_case(1, 96, 96, 6, 16, 7, 1, 3)              # Cin 6 -> 8, 49 taps, K = 392 -> 448


def test_conv_single_channel_audio_first_layer():
    run_case(2, 80, 16, 1, 32, 3, 1, 1)              # Cin 1 -> 8


def test_conv3x3_valid_to_1x1_and_big_k():
    run_case(4, 3, 3, 512, 512, 3, 1, 0)             # K = 4608, M = 4


def test_conv_transpose_stride2():
    run_case(2, 6, 6, 768, 384, 3, 2, 1, transpose=True, out_pad=1)
    run_case(1, 48, 48, 160, 64, 3, 2, 1, transpose=True, out_pad=1)


def test_conv_transpose_stride1_nopad():
    run_case(3, 1, 1, 1024, 512, 3, 1, 0, transpose=True)


def test_concat_channel_offsets():
    # read channels [64:80) of an 80-wide buffer, write channels [128:160) of a 160-wide one
    run_case(2, 96, 96, 16, 32, 3, 2, 1, in_coff=64, in_C=80, out_coff=128, out_C=160)


def test_no_bn_no_relu():
    run_case(2, 12, 12, 80, 32, 3, 1, 1, bn=False, relu=False)


def test_conv3x3_row_halo_tiles():
    """single-row 128-pixel tiles (TW = 128): the three dx taps of a filter row are descriptors into ONE loaded row segment of
    130 pixels (start address + 0 / 128 / 256 B); widths of one and two tiles, a ragged width, channel offsets, a residual, CTA
    pairs on the larger one"""
    run_case(2, 6, 128, 64, 64, 3, 1, 1, seed=11)
    run_case(1, 5, 256, 128, 128, 3, 1, 1, residual=True, seed=12)
    run_case(2, 4, 160, 64, 32, 3, 1, 1, seed=13)                              # 2 tiles per row, the second one ragged
    run_case(1, 3, 128, 64, 48, 3, 1, 1, in_coff=64, in_C=192, out_coff=16, out_C=96, seed=14)
    run_case(16, 40, 128, 64, 64, 3, 1, 1, bn=False, relu=False, seed=15)      # 640 row tiles: CTA pairs (cta_group::2)
    run_case(2, 6, 128, 64, 64, (1, 3)[1], 1, (0, 1)[1], seed=16)
    # 256 output channels with fewer row tiles than SMs (no CTA pair): three 256-row B tiles per stage would not fit the ring
    # twice -- the planner must take a narrower BN (MuseTalk VAE 256 -> 256 @128x128 at B = 1 failed to plan before)
    run_case(1, 100, 128, 256, 256, 3, 1, 1, residual=True, seed=17)
    run_case(1, 64, 128, 512, 256, 3, 1, 1, seed=18)


def test_parallel_region_branches_are_checked_and_equal_the_serial_program():
    """ProgramBuilder.par: two branches of one region (here: the two halves of a 128-channel 3x3 conv as independent 64-channel convs,
    each followed by a second conv of its own branch) give the result of the untagged program bit for bit; branches that write the
    same output channels are refused at load (MfError), not raced."""
    from mere_fusion_b200._lib import MfError
    from mere_fusion_b200.convnet_pack import ProgramBuilder
    from mere_fusion_b200.wav2lip import ConvNet
    g = torch.Generator().manual_seed(5)
    B, H, C = 2, 12, 64
    x = torch.randn(B, H, H, 2 * C, generator=g)
    ws = [(torch.randn(C, C, 3, 3, generator=g) / 24).numpy() for _ in range(4)]
    bs = [(torch.randn(C, generator=g) * 0.1).numpy() for _ in range(4)]

    def build(tagged, clash=False):
        pb = ProgramBuilder(nominal_batch=B)
        ib, mid, ob = pb.buffer(H, H, 2 * C), pb.buffer(H, H, 2 * C), pb.buffer(H, H, 2 * C)
        for br in range(2):
            pb.par = (br + 1) if tagged else 0
            pb.conv(ib, br * C, mid, br * C, ws[2 * br], bs[2 * br], None, stride=1, padding=1)
            pb.conv(mid, br * C, ob, 0 if clash else br * C, ws[2 * br + 1], bs[2 * br + 1], None, stride=1, padding=1)
        pb.par = 0
        return pb, ib, ob
    outs = []
    for tagged in (False, True):
        pb, ib, ob = build(tagged)
        net = ConvNet(pb.finish(), max_batch=B)
        outs.append(net.debug_run(ib, x, ob, (B, H, H, 2 * C)).cpu())
        torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1]) and float(outs[0].abs().max()) > 0
    pb, ib, ob = build(True, clash=True)
    with pytest.raises(MfError, match="different branches of a parallel region"):
        ConvNet(pb.finish(), max_batch=B)
