"""Pack a HF `Wav2Vec2ForCTC` state_dict (the acoustic model behind NerfASR.__frame_to_text, nerfasr.py:40-45,128-143:
`AutoModelForCTC.from_pretrained(opt.asr_model)`, default cpierse/wav2vec2-large-xlsr-53-esperanto, 44 CTC labels) into a
conv-net program for csrc/wav2lip.cu.  Third-party architecture: transformers' modeling_wav2vec2.py (layer-norm feature
encoder + stable-layer-norm transformer, XLSR-53 large).

  waveform -> zero mean / unit variance (Wav2Vec2FeatureExtractor do_normalize) -> conv0 (k 10, stride 5, fp32, in the input
  kernel) -> [LayerNorm over channels + GELU] -> 6 x [conv k3/k2 stride 2 -> LayerNorm + GELU] -> LayerNorm -> Linear 512 -> D
  -> x + GELU(grouped positional conv k 128, 16 groups, weight-norm folded) -> n x [x += attn(LN(x)); x += ffn(LN(x))] -> LN
  -> lm_head (fp32 logits [T, vocab])

The program is built for ONE window length (n_samples = (l + m + r) * 320 = 8960 for the live defaults -> 27 frames).
"""
import numpy as np

from .convnet_pack import ACT_GELU, ProgramBuilder, f32_to_bf16_bits

XLSR53_CFG = dict(vocab=44, hidden=1024, layers=24, heads=16, inter=4096, conv_dim=(512,) * 7, conv_stride=(5, 2, 2, 2, 2, 2, 2),
                  conv_kernel=(10, 3, 3, 3, 3, 2, 2), pos_k=128, pos_groups=16, eps=1e-5)


def _np(v):
    return v.detach().cpu().float().numpy() if hasattr(v, "detach") else np.asarray(v, np.float32)


def frame_counts(n_samples, cfg=XLSR53_CFG):
    T, out = n_samples, []
    for k, s in zip(cfg["conv_kernel"], cfg["conv_stride"]):
        T = (T - k) // s + 1
        out.append(T)
    return out


def pack_wav2vec2(sd, cfg=XLSR53_CFG, n_samples=8960, fused_stack=True):
    sd = {k: _np(v) for k, v in sd.items()}
    D, L, H, I, V = cfg["hidden"], cfg["layers"], cfg["heads"], cfg["inter"], cfg["vocab"]
    cd, ck, cs = cfg["conv_dim"], cfg["conv_kernel"], cfg["conv_stride"]
    eps = cfg["eps"]
    assert D % H == 0 and (D // H) % 8 == 0 and D % cfg["pos_groups"] == 0 and (D // cfg["pos_groups"]) % 64 == 0

    def need(name, shape):
        if name not in sd or tuple(sd[name].shape) != tuple(shape):
            raise ValueError(f"pack_wav2vec2: {name} is {None if name not in sd else sd[name].shape}, expected {tuple(shape)}")
        return sd[name]

    Ts = frame_counts(n_samples, cfg)
    pb = ProgramBuilder(1)
    fe = "wav2vec2.feature_extractor.conv_layers."
    # conv0 runs in fp32 inside the input kernel (the waveform is never rounded to bf16)
    w0 = need(fe + "0.conv.weight", (cd[0], 1, ck[0])).reshape(cd[0], ck[0])
    b0 = need(fe + "0.conv.bias", (cd[0],))
    conv0_id = pb._tensor(np.concatenate([w0.reshape(-1), b0]).astype(np.float32).tobytes())
    pb.flops_per_sample += 2 * cd[0] * ck[0] * Ts[0]
    c = pb.buffer(Ts[0], 1, cd[0])
    in_buf = c
    a = pb.buffer(Ts[0], 1, cd[0])
    pb.layer_norm(c, a, need(fe + "0.layer_norm.weight", (cd[0],)), need(fe + "0.layer_norm.bias", (cd[0],)), 1e-5, act=ACT_GELU)
    for i in range(1, len(cd)):
        c = pb.buffer(Ts[i], 1, cd[i])
        pb.conv(a, 0, c, 0, need(fe + f"{i}.conv.weight", (cd[i], cd[i - 1], ck[i]))[:, :, :, None], need(fe + f"{i}.conv.bias", (cd[i],)),
                stride=(cs[i], 1), padding=0, relu=False)
        a = pb.buffer(Ts[i], 1, cd[i])
        pb.layer_norm(c, a, need(fe + f"{i}.layer_norm.weight", (cd[i],)), need(fe + f"{i}.layer_norm.bias", (cd[i],)), 1e-5, act=ACT_GELU)
    T = Ts[-1]
    # feature projection
    n0 = pb.buffer(T, 1, cd[-1])
    pb.layer_norm(a, n0, need("wav2vec2.feature_projection.layer_norm.weight", (cd[-1],)),
                  need("wav2vec2.feature_projection.layer_norm.bias", (cd[-1],)), eps)
    h = pb.buffer(T, 1, D)
    pb.linear(n0, 0, h, 0, need("wav2vec2.feature_projection.projection.weight", (D, cd[-1])),
              need("wav2vec2.feature_projection.projection.bias", (D,)))
    # positional conv embedding: weight_norm(dim=2): w = g * v / ||v|| over (out, in) per kernel position
    pc = "wav2vec2.encoder.pos_conv_embed.conv."
    G, K = cfg["pos_groups"], cfg["pos_k"]
    cg = D // G
    if pc + "parametrizations.weight.original0" in sd:
        g_, v_ = need(pc + "parametrizations.weight.original0", (1, 1, K)), need(pc + "parametrizations.weight.original1", (D, cg, K))
    else:                                                                  # older torch weight_norm naming
        g_, v_ = need(pc + "weight_g", (1, 1, K)), need(pc + "weight_v", (D, cg, K))
    wpos = (v_ * (g_ / np.sqrt((v_.astype(np.float64) ** 2).sum(axis=(0, 1), keepdims=True))).astype(np.float32)).astype(np.float32)
    bpos = need(pc + "bias", (D,))
    x = pb.buffer(T, 1, D)
    for g in range(G):   # x = h + gelu(conv(h)): the residual is added after the activation
        pb.par = 1 + g % 16                                    # the G groups are independent: branches of one parallel region
        pb.conv1d_same(h, g * cg, x, g * cg, wpos[g * cg:(g + 1) * cg], bpos[g * cg:(g + 1) * cg], left_pad=K // 2, act=ACT_GELU,
                       res=(h, g * cg), res_after_act=True)
    pb.par = 0
    import os
    fused = fused_stack and os.environ.get("MF_W2V_FUSED", "1") != "0" and D % 64 == 0 and (I % 512 == 0 or (I <= 1024 and I % 64 == 0)) and max(3 * D, I) <= 4096 and T <= 32 and (D // H) <= 128
    ln = pb.buffer(T, 1, D)
    if fused:
        # the L transformer layers as ONE persistent kernel (csrc/w2v_stack.cuh): per layer fp32 vectors, then bf16 row-major matrices
        parts = []
        for i in range(L):
            p = f"wav2vec2.encoder.layers.{i}."
            wq = np.concatenate([need(p + f"attention.{n}_proj.weight", (D, D)) for n in ("q", "k", "v")])
            bq = np.concatenate([need(p + f"attention.{n}_proj.bias", (D,)) for n in ("q", "k", "v")])
            # the LayerNorm affine is folded into the linear layer it feeds (fp64 at pack time): LN(x) W^T + b =
            # ((x - mean) rstd) (W diag(g))^T + (b + W beta); the kernel normalises only.  The image keeps the g / beta slots (unit / zero).
            g1, be1 = need(p + "layer_norm.weight", (D,)).astype(np.float64), need(p + "layer_norm.bias", (D,)).astype(np.float64)
            g2, be2 = need(p + "final_layer_norm.weight", (D,)).astype(np.float64), need(p + "final_layer_norm.bias", (D,)).astype(np.float64)
            w1 = need(p + "feed_forward.intermediate_dense.weight", (I, D))
            bq = (bq.astype(np.float64) + wq.astype(np.float64) @ be1).astype(np.float32)
            b1 = (need(p + "feed_forward.intermediate_dense.bias", (I,)).astype(np.float64) + w1.astype(np.float64) @ be2).astype(np.float32)
            wq = (wq.astype(np.float64) * g1[None, :]).astype(np.float32)
            w1 = (w1.astype(np.float64) * g2[None, :]).astype(np.float32)
            vec = np.concatenate([np.ones(D, np.float32), np.zeros(D, np.float32), bq,
                                  need(p + "attention.out_proj.bias", (D,)), np.ones(D, np.float32),
                                  np.zeros(D, np.float32), b1,
                                  need(p + "feed_forward.output_dense.bias", (D,))]).astype(np.float32)
            assert vec.size == 9 * D + I
            parts.append(vec.tobytes())
            for wm in (wq, need(p + "attention.out_proj.weight", (D, D)), w1,
                       need(p + "feed_forward.output_dense.weight", (D, I))):
                parts.append(f32_to_bf16_bits(np.ascontiguousarray(wm, np.float32)).tobytes())
        x2 = pb.buffer(T, 1, D)
        flops = L * (2 * T * (4 * D * D + 2 * D * I) + 2 * 2 * H * T * T * (D // H))
        pb.transformer_stack(x, x2, b"".join(parts), D, I, H, L, eps, flops)
        x = x2
    else:
        qkv, ao, hid = pb.buffer(T, 1, 3 * D), pb.buffer(T, 1, D), pb.buffer(T, 1, I)
        for i in range(L):
            p = f"wav2vec2.encoder.layers.{i}."
            pb.layer_norm(x, ln, need(p + "layer_norm.weight", (D,)), need(p + "layer_norm.bias", (D,)), eps)
            wq = np.concatenate([need(p + f"attention.{n}_proj.weight", (D, D)) for n in ("q", "k", "v")])
            bq = np.concatenate([need(p + f"attention.{n}_proj.bias", (D,)) for n in ("q", "k", "v")])
            pb.linear(ln, 0, qkv, 0, wq, bq)
            pb.attention((qkv, 0), (qkv, D), (qkv, 2 * D), (ao, 0), H, D // H)
            x1 = pb.buffer(T, 1, D)
            pb.linear(ao, 0, x1, 0, need(p + "attention.out_proj.weight", (D, D)), need(p + "attention.out_proj.bias", (D,)), res=(x, 0))
            pb.layer_norm(x1, ln, need(p + "final_layer_norm.weight", (D,)), need(p + "final_layer_norm.bias", (D,)), eps)
            pb.linear(ln, 0, hid, 0, need(p + "feed_forward.intermediate_dense.weight", (I, D)),
                      need(p + "feed_forward.intermediate_dense.bias", (I,)), act=ACT_GELU)
            x2 = pb.buffer(T, 1, D)
            pb.linear(hid, 0, x2, 0, need(p + "feed_forward.output_dense.weight", (D, I)), need(p + "feed_forward.output_dense.bias", (D,)),
                      res=(x1, 0))
            x = x2
    pb.layer_norm(x, ln, need("wav2vec2.encoder.layer_norm.weight", (D,)), need("wav2vec2.encoder.layer_norm.bias", (D,)), eps)
    pb.linear(ln, 0, -1, 0, need("lm_head.weight", (V, D)), need("lm_head.bias", (V,)), mode=3)
    pb.aux = [n_samples, T, V, conv0_id, cd[0], ck[0], cs[0]]
    pb.hdr.update(in_face_buf=in_buf, in_mel_buf=-1, face_hw=0, mel_h=0, mel_w=-3, out_hw=0)
    pb.n_frames, pb.vocab = T, V
    return pb.finish(), pb
