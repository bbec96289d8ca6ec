"""Program builder for the sm_100a conv-net executor (csrc/wav2lip.cu).

A network is described as a list of implicit-GEMM conv ops over NHWC bf16 buffers; the builder
turns PyTorch-layout weights (+ BatchNorm statistics) into the padded K-major weight matrices,
per-channel scale/shift vectors and op records the C side consumes.  Pure host code (numpy).
"""
import struct

import numpy as np

from .ernerf_pack import build_blob

CONV_MAX_TAPS = 128
CONV_BK = 64
ID_PROGRAM = 1
ID_AUX = 2
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
ID_FIRST_TENSOR = 16
SM_COUNT = 148


def f32_to_bf16_bits(a):
    """round-to-nearest-even fp32 -> bf16, returned as uint16"""
    a = np.ascontiguousarray(a, np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    u = u + 0x7FFF + ((u >> 16) & 1)
    return (u >> 16).astype(np.uint16)


def bn_fold(conv_bias, bn, cout, eps=1e-5):
    """Conv2d bias + eval-mode BatchNorm2d -> per-channel (scale, shift) applied to the raw conv sum
    (wav2lip/models/conv.py:8-11)"""
    bias = np.zeros(cout, np.float32) if conv_bias is None else np.asarray(conv_bias, np.float32)
    if bn is None:
        return np.ones(cout, np.float32), bias
    gamma, beta, mean, var = [np.asarray(t, np.float32) for t in bn]
    scale = gamma / np.sqrt(var + np.float32(eps))
    return scale.astype(np.float32), ((bias - mean) * scale + beta).astype(np.float32)


class ProgramBuilder:
    def __init__(self, nominal_batch=16):
        self.buffers = []
        self.ops = []
        self.tensors = {}
        self.buffer_init = {}
        self.aux = None                # optional int32 list stored as blob entry ID_AUX (program-kind specific)
        self.next_id = ID_FIRST_TENSOR
        self.nominal_batch = nominal_batch
        self.hdr = dict(in_face_buf=-1, in_mel_buf=-1, face_hw=0, mel_h=0, mel_w=0, out_hw=0)
        self.flops_per_sample = 0      # algorithmic: 2 * MACs of the original (unpadded) layers
        self.decompose_ups = True      # run upsample + 3x3 conv as four 2x2-tap parity convs

    def buffer(self, H, W, C, init=None):
        """init: optional constant content [H, W, C] (fp32, rounded to bf16), replicated over the batch at load; such a
        buffer must never be an op output"""
        assert C % 8 == 0
        self.buffers.append((H, W, C))
        if init is not None:
            init = np.asarray(init, np.float32)
            assert init.shape == (H, W, C)
            self.buffer_init[len(self.buffers) - 1] = self._tensor(f32_to_bf16_bits(init).tobytes())
        return len(self.buffers) - 1

    def _tensor(self, data):
        i = self.next_id
        self.next_id += 1
        self.tensors[i] = data
        return i

    def _pick_bn(self, cout, M):
        bn = 128
        while bn > 16 and (cout % bn != 0 and cout < bn):
            bn //= 2
        if cout <= 16:
            return 16
        bn = min(bn, 128)
        while cout % bn != 0 and bn > 16:
            bn //= 2
        # small-M layers are weight-bandwidth bound: spread the weight reads over more CTAs
        while bn > 32 and ((M + 127) // 128) * ((cout + bn - 1) // bn) < SM_COUNT:
            bn //= 2
        return bn

    def _emit(self, in_buf, in_coff, cin_pad, out_buf, out_coff, Wm, taps, scale, shift, Mh, Mw, oy0, ox0, osy, osx,
              isy, isx, res, relu, mode, cout, ups=0, flags=0):
        ntaps = len(taps)
        assert ntaps <= CONV_MAX_TAPS
        if mode == 0 and cout % 16 != 0:              # the bf16 epilogue stores 16 channels at a time: pad with zero rows
            pad = (cout + 15) // 16 * 16
            Wm = np.concatenate([Wm, np.zeros((pad - cout, Wm.shape[1]), np.float32)])
            scale = np.concatenate([scale, np.ones(pad - cout, np.float32)])
            shift = np.concatenate([shift, np.zeros(pad - cout, np.float32)])
            cout = pad
        M = self.nominal_batch * Mh * Mw
        bn = self._pick_bn(cout, M)
        cout_pad = (cout + bn - 1) // bn * bn
        K = ntaps * cin_pad
        kpad = (K + CONV_BK - 1) // CONV_BK * CONV_BK
        Wp = np.zeros((cout_pad, kpad), np.float32)
        Wp[:cout, :K] = Wm
        sc = np.zeros(cout_pad, np.float32)
        sh = np.zeros(cout_pad, np.float32)
        sc[:cout] = scale
        sh[:cout] = shift
        # stored re-tiled as [Cout_pad / 16][Kpad / 64][16][64]: every 16-row x 64-k piece of a B tile is one contiguous 2 KB
        # run in HBM (a row-major [Cout][K] matrix would be fetched as 128-byte pieces 2 * K bytes apart, one DRAM page each:
        # the weight-streaming 4x4 / 8x8 UNet layers measured 1.2 TB/s that way)
        Wt = Wp.reshape(cout_pad // 16, 16, kpad // CONV_BK, CONV_BK).transpose(0, 2, 1, 3)
        w_id = self._tensor(f32_to_bf16_bits(np.ascontiguousarray(Wt)).tobytes())
        s_id = self._tensor(sc.tobytes())
        h_id = self._tensor(sh.tobytes())
        dy = [t[0] for t in taps] + [0] * (CONV_MAX_TAPS - ntaps)
        dx = [t[1] for t in taps] + [0] * (CONV_MAX_TAPS - ntaps)
        rb, rc = res if res is not None else (-1, 0)
        if np.all(sc[:cout] == 1.0):
            flags |= 2                                  # unit scale: the epilogue adds the bias only
        flags |= (int(getattr(self, "par", 0)) & 0xff) << 8   # branch of a parallel region (see `par` below)
        rec = struct.pack("<28i", in_buf, in_coff, out_buf, out_coff, rb, rc, Mh, Mw, oy0, ox0, osy, osx, isy, isx,
                          ntaps, cin_pad, kpad, cout, cout_pad, bn, int(relu), mode, w_id, s_id, h_id, 0, ups, flags)
        rec += struct.pack(f"<{CONV_MAX_TAPS}b", *dy) + struct.pack(f"<{CONV_MAX_TAPS}b", *dx)
        self.ops.append(rec)
        return len(self.ops) - 1

    def conv(self, in_buf, in_coff, out_buf, out_coff, weight, bias=None, bn=None, stride=1, padding=0, res=None,
             relu=True, mode=0, ups=0, extra_shift=None, wscale=1.0, res_after_act=False, cin_align=8):
        """nn.Conv2d (+BatchNorm2d, +residual, +activation).  weight [Cout, Cin, kh, kw].  relu: False / True / ACT_GELU.
        res_after_act: out = act(conv) + res instead of act(conv + res).  ups = 1: the conv runs on the
        nearest-neighbour 2x upsampling of the input buffer (F.interpolate(scale_factor=2) folded into the gather).
        extra_shift: per-channel constant added after the conv (e.g. a time-embedding projection).
        cin_align = 64: read the input channels up to the next multiple of 64 with ZERO weights for the extra ones (they must exist
        in the buffer and hold finite values -- buffers are zero-filled at load and pad channels are never written), which makes a
        stride-1 layer eligible for the TMA-fed kernel (Cin % 64 == 0) at the price of the padded MACs."""
        weight = np.asarray(weight, np.float32)
        cout, cin, kh, kw = weight.shape
        sy, sx = (stride, stride) if np.isscalar(stride) else stride
        py, px = (padding, padding) if np.isscalar(padding) else padding
        Hin, Win, _ = self.buffers[in_buf]
        Hin, Win = Hin << ups, Win << ups
        Hout, Wout = (Hin + 2 * py - kh) // sy + 1, (Win + 2 * px - kw) // sx + 1
        if mode == 0:
            assert self.buffers[out_buf][:2] == (Hout, Wout), (self.buffers[out_buf], Hout, Wout)
        cin_pad = (cin + cin_align - 1) // cin_align * cin_align
        assert in_coff + cin_pad <= self.buffers[in_buf][2], (in_coff, cin_pad, self.buffers[in_buf])
        taps = [(ky - py, kx - px) for ky in range(kh) for kx in range(kw)]
        Wm = np.zeros((cout, len(taps), cin_pad), np.float32)
        Wm[:, :, :cin] = weight.transpose(0, 2, 3, 1).reshape(cout, kh * kw, cin) * np.float32(wscale)
        scale, shift = bn_fold(bias, bn, cout)
        if extra_shift is not None:
            shift = (shift + np.asarray(extra_shift, np.float32) * scale).astype(np.float32)
        self.flops_per_sample += 2 * cout * cin * kh * kw * Hout * Wout
        if ups == 1 and (kh, kw, sy, sx, py, px) == (3, 3, 1, 1, 1, 1) and cin % CONV_BK == 0 and mode == 0 and self.decompose_ups:
            # nearest-2x upsampling followed by a 3x3 conv == four 2x2-tap convs on the LOW-resolution input, one per output
            # parity class, with the taps that read the same input pixel summed (4/9 of the MACs, exact in real arithmetic):
            #   parity 0: rows (2y-1, 2y, 2y+1) -> input rows (y-1, y, y): taps {-1: k0, 0: k1+k2};  parity 1: {0: k0+k1, +1: k2}
            sets = {0: [(-1, [0]), (0, [1, 2])], 1: [(0, [0, 1]), (1, [2])]}
            w = weight * np.float32(wscale)
            ids = []
            for cy in (0, 1):
                for cx in (0, 1):
                    ptaps, cols = [], []
                    for dy, kys in sets[cy]:
                        for dx, kxs in sets[cx]:
                            ptaps.append((dy, dx))
                            cols.append(sum(w[:, :, ky, kx] for ky in kys for kx in kxs))          # [Cout, Cin]
                    Wp = np.stack(cols, 1).reshape(cout, -1)
                    ids.append(self._emit(in_buf, in_coff, cin_pad, out_buf, out_coff, Wp, ptaps, scale, shift, Hin >> 1, Win >> 1,
                                          cy, cx, 2, 2, 1, 1, res, relu, mode, cout, ups=0, flags=int(bool(res_after_act))))
            return ids
        return self._emit(in_buf, in_coff, cin_pad, out_buf, out_coff, Wm.reshape(cout, -1), taps, scale, shift, Hout,
                          Wout, 0, 0, 1, 1, sy, sx, res, relu, mode, cout, ups=ups, flags=int(bool(res_after_act)))

    def linear(self, in_buf, in_coff, out_buf, out_coff, weight, bias=None, res=None, act=False, mode=0):
        """nn.Linear on a token buffer = 1x1 conv.  weight [Cout, Cin].  mode = 3: fp32 [tokens, Cout] to the caller's buffer"""
        w = np.asarray(weight, np.float32)
        return self.conv(in_buf, in_coff, out_buf, out_coff, w[:, :, None, None], bias, None, res=res, relu=act, mode=mode)

    # Parallel regions.  While `self.par` is non-zero every conv op that is emitted carries it in flags bits 8..15: a maximal run of
    # tagged ops is a REGION, the ops of one tag are a BRANCH (executed in order), and the packer promises that no op touches the
    # output channels of an op of another branch of its region (checked at load).  The executor's captured graph runs the branches of
    # a region side by side (fork-join): the groups of a grouped conv, or two sub-networks that only meet later.
    par = 0

    def conv1d_same(self, in_buf, in_coff, out_buf, out_coff, weight, bias, left_pad, act=False, res=None, res_after_act=False):
        """nn.Conv1d over the token axis of [T, 1, C] buffers with `left_pad` zeros in front and as many behind as needed to
        keep T outputs (an even kernel with padding k/2 followed by dropping the last output: HF Wav2Vec2SamePadLayer).
        weight [Cout, Cin, k]"""
        w = np.asarray(weight, np.float32)
        cout, cin, k = w.shape
        T = self.buffers[in_buf][0]
        assert self.buffers[in_buf][1] == 1 and self.buffers[out_buf][:2] == (T, 1) and cin % 8 == 0
        taps = [(j - left_pad, 0) for j in range(k)]
        Wm = w.transpose(0, 2, 1).reshape(cout, k * cin)
        scale, shift = bn_fold(bias, None, cout)
        self.flops_per_sample += 2 * cout * cin * k * T
        return self._emit(in_buf, in_coff, cin, out_buf, out_coff, Wm, taps, scale, shift, T, 1, 0, 0, 1, 1, 1, 1, res, act, 0, cout,
                          flags=int(bool(res_after_act)))

    def _misc(self, kind, in_buf, out_buf, **f):
        v = dict(in_coff=0, out_coff=0, res_buf=-1, res_coff=0, Mh=0, Mw=0, ntaps=0, Cin=0, Kpad=0, relu=0, s_id=-1, h_id=-1)
        v.update(f)
        rec = struct.pack("<28i", in_buf, v["in_coff"], out_buf, v["out_coff"], v["res_buf"], v["res_coff"], v["Mh"], v["Mw"],
                          0, 0, 1, 1, 1, 1, v["ntaps"], v["Cin"], v["Kpad"], 0, 0, 0, v["relu"], 0, -1, v["s_id"], v["h_id"],
                          kind, 0, 0)
        rec += bytes(2 * CONV_MAX_TAPS)
        self.ops.append(rec)
        return len(self.ops) - 1

    @staticmethod
    def _fbits(x):
        return struct.unpack("<i", struct.pack("<f", float(x)))[0]

    def group_norm(self, in_buf, out_buf, gamma, beta, groups=32, eps=1e-5, silu=False, in_coff=0):
        """input = channels [in_coff, in_coff + C) of in_buf (C = len(gamma)); output buffer is dense [H, W, C]"""
        C = len(gamma)
        assert self.buffers[out_buf] == self.buffers[in_buf][:2] + (C,) and in_coff + C <= self.buffers[in_buf][2]
        slot = getattr(self, "_gn_slots", 0)
        self._gn_slots = slot + 1
        return self._misc(1, in_buf, out_buf, in_coff=in_coff, Cin=C, ntaps=groups, relu=int(silu), Kpad=self._fbits(eps), Mh=slot,
                          s_id=self._tensor(np.asarray(gamma, np.float32).tobytes()),
                          h_id=self._tensor(np.asarray(beta, np.float32).tobytes()))

    def layer_norm(self, in_buf, out_buf, gamma, beta, eps=1e-5, act=ACT_NONE):
        """act = ACT_GELU: LayerNorm followed by exact GELU in the same kernel (wav2vec2 feature encoder)"""
        C = self.buffers[in_buf][2]
        assert self.buffers[out_buf] == self.buffers[in_buf] and len(gamma) == C and act in (ACT_NONE, ACT_GELU)
        return self._misc(2, in_buf, out_buf, Cin=C, Kpad=self._fbits(eps), relu=int(act),
                          s_id=self._tensor(np.asarray(gamma, np.float32).tobytes()),
                          h_id=self._tensor(np.asarray(beta, np.float32).tobytes()))

    def attention(self, q, k, v, out, heads, dim_head):
        """q, k, v, out: (buffer, channel offset).  softmax(q k^T / sqrt(dim_head)) v per head"""
        nq = self.buffers[q[0]][0] * self.buffers[q[0]][1]
        nk = self.buffers[k[0]][0] * self.buffers[k[0]][1]
        self.flops_per_sample += 2 * 2 * heads * nq * nk * dim_head
        return self._misc(3, q[0], out[0], in_coff=q[1], out_coff=out[1], res_buf=k[0], res_coff=k[1], Mh=v[0], Mw=v[1],
                          ntaps=heads, Cin=dim_head, Kpad=self._fbits(dim_head ** -0.5))

    def transformer_stack(self, in_buf, out_buf, image, D, inter, heads, layers, eps, flops):
        """op kind 5 (csrc/w2v_stack.cuh): `layers` pre-LN transformer layers as ONE persistent kernel; `image` = the packed weights
        of all layers (bytes, layout in w2v_stack.cuh)"""
        assert self.buffers[in_buf] == self.buffers[out_buf] and self.buffers[in_buf][2] == D
        self.flops_per_sample += flops
        rec = struct.pack("<28i", in_buf, 0, out_buf, 0, -1, 0, layers, inter, 0, 0, 1, 1, 1, 1, heads, D, self._fbits(eps), 0, 0, 0, 0, 0,
                          self._tensor(image), -1, -1, 5, 0, 0)
        rec += bytes(2 * CONV_MAX_TAPS)
        self.ops.append(rec)
        return len(self.ops) - 1

    def geglu(self, in_buf, out_buf):
        Hd = self.buffers[out_buf][2]
        assert self.buffers[in_buf][2] == 2 * Hd
        return self._misc(4, in_buf, out_buf, Cin=Hd)

    def conv_transpose(self, in_buf, in_coff, out_buf, out_coff, weight, bias=None, bn=None, stride=1, padding=0,
                       output_padding=0, relu=True, cin_align=8):
        """nn.ConvTranspose2d (+BatchNorm2d +ReLU).  weight [Cin, Cout, kh, kw].  out[oy] gathers
        in[(oy + p - ky) / s] for the taps where the division is exact: one op per output-parity
        class, each with only its live taps."""
        weight = np.asarray(weight, np.float32)
        cin, cout, kh, kw = weight.shape
        s, p = stride, padding
        Hin, Win, _ = self.buffers[in_buf]
        Hout = (Hin - 1) * s - 2 * p + kh + output_padding
        Wout = (Win - 1) * s - 2 * p + kw + output_padding
        assert self.buffers[out_buf][:2] == (Hout, Wout), (self.buffers[out_buf], Hout, Wout)
        cin_pad = (cin + cin_align - 1) // cin_align * cin_align              # zero weights for the pad channels (see conv)
        assert cin_pad % 8 == 0 and in_coff + cin_pad <= self.buffers[in_buf][2]
        scale, shift = bn_fold(bias, bn, cout)
        self.flops_per_sample += 2 * cout * cin * kh * kw * Hin * Win
        ids = []
        for cy in range(s):
            for cx in range(s):
                ty = [(ky, (cy + p - ky) // s) for ky in range(kh) if (cy + p - ky) % s == 0]
                tx = [(kx, (cx + p - kx) // s) for kx in range(kw) if (cx + p - kx) % s == 0]
                Mh = (Hout - cy + s - 1) // s
                Mw = (Wout - cx + s - 1) // s
                if not ty or not tx or Mh <= 0 or Mw <= 0:
                    raise NotImplementedError("conv_transpose: parity class without taps needs a bias-only op")
                taps, cols = [], []
                for ky, dy in ty:
                    for kx, dx in tx:
                        taps.append((dy, dx))
                        col = np.zeros((cout, cin_pad), np.float32)
                        col[:, :cin] = weight[:, :, ky, kx].T        # [Cout, Cin]
                        cols.append(col)
                Wm = np.stack(cols, 1).reshape(cout, -1)
                ids.append(self._emit(in_buf, in_coff, cin_pad, out_buf, out_coff, Wm, taps, scale, shift, Mh, Mw, cy, cx,
                                      s, s, 1, 1, None, relu, 0, cout))
        return ids

    def finish(self):
        h = self.hdr
        prog = struct.pack("<8i", len(self.buffers), len(self.ops), h["in_face_buf"], h["in_mel_buf"], h["face_hw"],
                           h["mel_h"], h["mel_w"], h["out_hw"])
        for i, (H, W, C) in enumerate(self.buffers):
            prog += struct.pack("<4i", H, W, C, self.buffer_init.get(i, 0))
        prog += b"".join(self.ops)
        entries = {ID_PROGRAM: prog}
        if self.aux is not None:
            entries[ID_AUX] = np.asarray(self.aux, np.int32).tobytes()
        entries.update(self.tensors)
        return build_blob(entries, kind=2)
