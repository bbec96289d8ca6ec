"""Pack the MuseTalk frame path (diffusers-format UNet2DConditionModel + AutoencoderKL decoder state_dicts, as
loaded by musetalk/utils/utils.py:70-75) into one conv-net program for csrc/wav2lip.cu.

Single denoising step at t = 0 (musereal.py:59,105-107): the time embedding is a constant, so every ResnetBlock's
`time_emb_proj(silu(temb))` is folded into its conv1 bias at pack time.  `latents / scaling_factor` (vae.py:101) is
folded into post_quant_conv.  Skip connections / torch.cat are channel ranges of shared buffers.  The loader is
strict: every tensor must have the shape of the SD-1.x / sd-vae-ft-mse architecture for the given config.

PARITY UNPINNED (diffusers, model json and checkpoints are all absent from the reference repo; see
oracle/musetalk_oracle.py).
"""
import numpy as np

from .convnet_pack import ProgramBuilder

UNET_CFG = dict(block_out=(320, 640, 1280, 1280), layers=2, heads=8, ctx_dim=384, in_ch=8, out_ch=4, groups=32,
                attn=(True, True, True, False), eps=1e-5)
VAE_CFG = dict(block_out=(128, 256, 512, 512), layers=2, latent_ch=4, out_ch=3, groups=32, eps=1e-6,
               scaling_factor=0.18215)


def _np(v):
    return v.detach().cpu().float().numpy() if hasattr(v, "detach") else np.asarray(v, np.float32)


def _silu(x):
    return x / (1.0 + np.exp(-x))


class _Pool:
    """scratch buffers by shape; `get` hands out a free one, `put` returns it"""

    def __init__(self, pb):
        self.pb, self.free = pb, {}

    def get(self, H, W, C):
        lst = self.free.setdefault((H, W, C), [])
        return lst.pop() if lst else self.pb.buffer(H, W, C)

    def put(self, *bufs):
        for b in bufs:
            self.free.setdefault(tuple(self.pb.buffers[b]), []).append(b)


def pack_musetalk(unet_sd, vae_sd, ucfg=UNET_CFG, vcfg=VAE_CFG, nominal_batch=16, latent_hw=32, ctx_len=50):
    U = {k: _np(v) for k, v in unet_sd.items()}
    V = {k: _np(v) for k, v in vae_sd.items()}
    pb = ProgramBuilder(nominal_batch)
    pool = _Pool(pb)
    bo, L, heads, G = ucfg["block_out"], ucfg["layers"], ucfg["heads"], ucfg["groups"]
    S = latent_hw

    def need(sd, name, shape):
        if name not in sd or tuple(sd[name].shape) != tuple(shape):
            raise ValueError(f"pack_musetalk: {name} is {None if name not in sd else sd[name].shape}, expected {tuple(shape)}")
        return sd[name]

    # ---- time embedding at t = 0 (get_timestep_embedding with flip_sin_to_cos: [cos | sin] = [1 | 0])
    t_emb = np.concatenate([np.ones(bo[0] // 2, np.float32), np.zeros(bo[0] // 2, np.float32)])
    temb_dim = 4 * bo[0]
    h1 = need(U, "time_embedding.linear_1.weight", (temb_dim, bo[0])) @ t_emb + need(U, "time_embedding.linear_1.bias", (temb_dim,))
    temb = need(U, "time_embedding.linear_2.weight", (temb_dim, temb_dim)) @ _silu(h1) + need(U, "time_embedding.linear_2.bias", (temb_dim,))
    temb_act = _silu(temb).astype(np.float32)

    in_lat = pb.buffer(S, S, 8)
    in_ctx = pb.buffer(ctx_len, 1, ucfg["ctx_dim"])
    pb.hdr.update(in_face_buf=in_lat, in_mel_buf=in_ctx, face_hw=ucfg["in_ch"], mel_h=0, mel_w=-1, out_hw=S * 8)

    def resnet(sd, p, x, xoff, cin, cout, H, out, ooff, eps, groups, with_temb):
        """ResnetBlock2D: x = channels [xoff, xoff+cin) of buffer x -> channels [ooff, ooff+cout) of buffer out"""
        t1 = pool.get(H, H, cin)
        pb.group_norm(x, t1, need(sd, p + ".norm1.weight", (cin,)), need(sd, p + ".norm1.bias", (cin,)), groups, eps, True, in_coff=xoff)
        h = pool.get(H, H, cout)
        extra = None
        if with_temb:
            extra = need(sd, p + ".time_emb_proj.weight", (cout, temb_dim)) @ temb_act + need(sd, p + ".time_emb_proj.bias", (cout,))
        pb.conv(t1, 0, h, 0, need(sd, p + ".conv1.weight", (cout, cin, 3, 3)), need(sd, p + ".conv1.bias", (cout,)), padding=1,
                relu=False, extra_shift=extra)
        pool.put(t1)
        t2 = pool.get(H, H, cout)
        pb.group_norm(h, t2, need(sd, p + ".norm2.weight", (cout,)), need(sd, p + ".norm2.bias", (cout,)), groups, eps, True)
        pool.put(h)
        if cin != cout:
            sc = pool.get(H, H, cout)
            pb.conv(x, xoff, sc, 0, need(sd, p + ".conv_shortcut.weight", (cout, cin, 1, 1)), need(sd, p + ".conv_shortcut.bias", (cout,)),
                    relu=False)
            res = (sc, 0)
        else:
            sc, res = None, (x, xoff)
        pb.conv(t2, 0, out, ooff, need(sd, p + ".conv2.weight", (cout, cout, 3, 3)), need(sd, p + ".conv2.bias", (cout,)), padding=1,
                res=res, relu=False)
        pool.put(t2)
        if sc is not None:
            pool.put(sc)

    def transformer(p, x, xoff, C, H, out, ooff):
        """Transformer2DModel (conv projections, one BasicTransformerBlock, GEGLU)"""
        ctx_dim = ucfg["ctx_dim"]
        dh = C // heads
        t = p + ".transformer_blocks.0"
        g = pool.get(H, H, C)
        pb.group_norm(x, g, need(U, p + ".norm.weight", (C,)), need(U, p + ".norm.bias", (C,)), G, 1e-6, False, in_coff=xoff)
        h = pool.get(H, H, C)
        pb.conv(g, 0, h, 0, need(U, p + ".proj_in.weight", (C, C, 1, 1)), need(U, p + ".proj_in.bias", (C,)), relu=False)
        n = g                                                           # reuse as the LayerNorm output
        # self-attention
        pb.layer_norm(h, n, need(U, t + ".norm1.weight", (C,)), need(U, t + ".norm1.bias", (C,)))
        qkv = pool.get(H, H, 3 * C)
        wqkv = np.concatenate([need(U, t + ".attn1.to_q.weight", (C, C)), need(U, t + ".attn1.to_k.weight", (C, C)),
                               need(U, t + ".attn1.to_v.weight", (C, C))])
        pb.linear(n, 0, qkv, 0, wqkv)
        ao = pool.get(H, H, C)
        pb.attention((qkv, 0), (qkv, C), (qkv, 2 * C), (ao, 0), heads, dh)
        pool.put(qkv)
        h2 = pool.get(H, H, C)
        pb.linear(ao, 0, h2, 0, need(U, t + ".attn1.to_out.0.weight", (C, C)), need(U, t + ".attn1.to_out.0.bias", (C,)), res=(h, 0))
        pool.put(h)
        # cross-attention over the 50 audio tokens
        pb.layer_norm(h2, n, need(U, t + ".norm2.weight", (C,)), need(U, t + ".norm2.bias", (C,)))
        qb = pool.get(H, H, C)
        pb.linear(n, 0, qb, 0, need(U, t + ".attn2.to_q.weight", (C, C)))
        kv = pool.get(ctx_len, 1, 2 * C)
        wkv = np.concatenate([need(U, t + ".attn2.to_k.weight", (C, ctx_dim)), need(U, t + ".attn2.to_v.weight", (C, ctx_dim))])
        pb.linear(in_ctx, 0, kv, 0, wkv)
        pb.attention((qb, 0), (kv, 0), (kv, C), (ao, 0), heads, dh)
        pool.put(qb, kv)
        h3 = pool.get(H, H, C)
        pb.linear(ao, 0, h3, 0, need(U, t + ".attn2.to_out.0.weight", (C, C)), need(U, t + ".attn2.to_out.0.bias", (C,)), res=(h2, 0))
        pool.put(h2, ao)
        # GEGLU feed-forward
        pb.layer_norm(h3, n, need(U, t + ".norm3.weight", (C,)), need(U, t + ".norm3.bias", (C,)))
        f1 = pool.get(H, H, 8 * C)
        pb.linear(n, 0, f1, 0, need(U, t + ".ff.net.0.proj.weight", (8 * C, C)), need(U, t + ".ff.net.0.proj.bias", (8 * C,)))
        f2 = pool.get(H, H, 4 * C)
        pb.geglu(f1, f2)
        pool.put(f1)
        h4 = pool.get(H, H, C)
        pb.linear(f2, 0, h4, 0, need(U, t + ".ff.net.2.weight", (C, 4 * C)), need(U, t + ".ff.net.2.bias", (C,)), res=(h3, 0))
        pool.put(f2, h3, n)
        pb.conv(h4, 0, out, ooff, need(U, p + ".proj_out.weight", (C, C, 1, 1)), need(U, p + ".proj_out.bias", (C,)), res=(x, xoff),
                relu=False)
        pool.put(h4)

    # ---- plan the skip connections: every tensor pushed on the down path lands directly in the concat buffer of
    #      the up-path resnet that pops it (torch.cat([x, skip], dim=1): x first)
    skip_shapes = [(bo[0], S)]
    ch, H = bo[0], S
    for i, co in enumerate(bo):
        for j in range(L):
            ch = co
            skip_shapes.append((ch, H))
        if i != len(bo) - 1:
            H //= 2
            skip_shapes.append((ch, H))
    up_plan = []                                                         # (cx, cskip, cout, H) per up resnet, in order
    st = list(skip_shapes)
    for i, co in enumerate(reversed(bo)):
        for j in range(L + 1):
            cs, Hs = st.pop()
            assert Hs == H
            up_plan.append((ch, cs, co, H))
            ch = co
        if i != len(bo) - 1:
            H *= 2
    cat = [pb.buffer(Hc, Hc, cx + cs) for (cx, cs, co, Hc) in up_plan]
    # skip k (push order) is popped by up resnet (n_up - 1 - k)
    n_up = len(up_plan)

    def skip_target(k):
        r = n_up - 1 - k
        return cat[r], up_plan[r][0]

    # ---- down path
    k = 0
    cur, coff = skip_target(k)
    pb.conv(in_lat, 0, cur, coff, need(U, "conv_in.weight", (bo[0], ucfg["in_ch"], 3, 3)), need(U, "conv_in.bias", (bo[0],)),
            padding=1, relu=False)
    k += 1
    ch, H = bo[0], S
    for i, co in enumerate(bo):
        for j in range(L):
            dst, doff = skip_target(k)
            k += 1
            if ucfg["attn"][i]:
                r = pool.get(H, H, co)
                resnet(U, f"down_blocks.{i}.resnets.{j}", cur, coff, ch, co, H, r, 0, ucfg["eps"], G, True)
                transformer(f"down_blocks.{i}.attentions.{j}", r, 0, co, H, dst, doff)
                pool.put(r)
            else:
                resnet(U, f"down_blocks.{i}.resnets.{j}", cur, coff, ch, co, H, dst, doff, ucfg["eps"], G, True)
            cur, coff, ch = dst, doff, co
        if i != len(bo) - 1:
            dst, doff = skip_target(k)
            k += 1
            pb.conv(cur, coff, dst, doff, need(U, f"down_blocks.{i}.downsamplers.0.conv.weight", (co, co, 3, 3)),
                    need(U, f"down_blocks.{i}.downsamplers.0.conv.bias", (co,)), stride=2, padding=1, relu=False)
            H //= 2
            cur, coff = dst, doff
    # ---- mid block
    m1 = pool.get(H, H, ch)
    resnet(U, "mid_block.resnets.0", cur, coff, ch, ch, H, m1, 0, ucfg["eps"], G, True)
    m2 = pool.get(H, H, ch)
    transformer("mid_block.attentions.0", m1, 0, ch, H, m2, 0)
    pool.put(m1)
    resnet(U, "mid_block.resnets.1", m2, 0, ch, ch, H, cat[0], 0, ucfg["eps"], G, True)      # x part of the first concat
    pool.put(m2)
    # ---- up path
    r_idx = 0
    rattn = list(reversed(ucfg["attn"]))
    lat_out = pb.buffer(S, S, 16)
    for i, co in enumerate(reversed(bo)):
        for j in range(L + 1):
            cx, cs, cout, Hc = up_plan[r_idx]
            last_of_block = j == L
            last = r_idx == n_up - 1
            # where does this resnet (+ attention) write?  next concat's x part, or a scratch before an upsampler / the head
            if not last_of_block:
                dst, doff = cat[r_idx + 1], 0
            else:
                dst, doff = pool.get(Hc, Hc, cout), 0
            if rattn[i]:
                r = pool.get(Hc, Hc, cout)
                resnet(U, f"up_blocks.{i}.resnets.{j}", cat[r_idx], 0, cx + cs, cout, Hc, r, 0, ucfg["eps"], G, True)
                transformer(f"up_blocks.{i}.attentions.{j}", r, 0, cout, Hc, dst, doff)
                pool.put(r)
            else:
                resnet(U, f"up_blocks.{i}.resnets.{j}", cat[r_idx], 0, cx + cs, cout, Hc, dst, doff, ucfg["eps"], G, True)
            r_idx += 1
            if last_of_block:
                if not last:
                    pb.conv(dst, 0, cat[r_idx], 0, need(U, f"up_blocks.{i}.upsamplers.0.conv.weight", (cout, cout, 3, 3)),
                            need(U, f"up_blocks.{i}.upsamplers.0.conv.bias", (cout,)), padding=1, relu=False, ups=1)
                    pool.put(dst)
                else:
                    g = pool.get(Hc, Hc, cout)
                    pb.group_norm(dst, g, need(U, "conv_norm_out.weight", (cout,)), need(U, "conv_norm_out.bias", (cout,)), G,
                                  ucfg["eps"], True)
                    pb.conv(g, 0, lat_out, 0, need(U, "conv_out.weight", (ucfg["out_ch"], cout, 3, 3)),
                            need(U, "conv_out.bias", (ucfg["out_ch"],)), padding=1, relu=False)
                    pool.put(g, dst)
    unet_flops = pb.flops_per_sample

    # ---- VAE decoder (AutoencoderKL.decode): latents / scaling_factor -> post_quant_conv -> Decoder
    vbo, VL, VG, veps, lc = vcfg["block_out"], vcfg["layers"], vcfg["groups"], vcfg["eps"], vcfg["latent_ch"]
    top = vbo[-1]
    z = pb.buffer(S, S, 16)
    pb.conv(lat_out, 0, z, 0, need(V, "post_quant_conv.weight", (lc, lc, 1, 1)), need(V, "post_quant_conv.bias", (lc,)), relu=False,
            wscale=1.0 / vcfg["scaling_factor"])
    x = pool.get(S, S, top)
    pb.conv(z, 0, x, 0, need(V, "decoder.conv_in.weight", (top, lc, 3, 3)), need(V, "decoder.conv_in.bias", (top,)), padding=1, relu=False)
    y = pool.get(S, S, top)
    resnet(V, "decoder.mid_block.resnets.0", x, 0, top, top, S, y, 0, veps, VG, False)
    a = "decoder.mid_block.attentions.0"
    g = pool.get(S, S, top)
    pb.group_norm(y, g, need(V, a + ".group_norm.weight", (top,)), need(V, a + ".group_norm.bias", (top,)), VG, veps, False)
    qkv = pool.get(S, S, 3 * top)
    pb.linear(g, 0, qkv, 0, np.concatenate([need(V, a + ".to_q.weight", (top, top)), need(V, a + ".to_k.weight", (top, top)),
                                            need(V, a + ".to_v.weight", (top, top))]),
              np.concatenate([need(V, a + ".to_q.bias", (top,)), need(V, a + ".to_k.bias", (top,)), need(V, a + ".to_v.bias", (top,))]))
    pb.attention((qkv, 0), (qkv, top), (qkv, 2 * top), (g, 0), 1, top)
    pool.put(qkv)
    pb.linear(g, 0, x, 0, need(V, a + ".to_out.0.weight", (top, top)), need(V, a + ".to_out.0.bias", (top,)), res=(y, 0))
    pool.put(g)
    resnet(V, "decoder.mid_block.resnets.1", x, 0, top, top, S, y, 0, veps, VG, False)
    pool.put(x)
    cur, ch, H = y, top, S
    for i, co in enumerate(reversed(vbo)):
        for j in range(VL + 1):
            nxt = pool.get(H, H, co)
            resnet(V, f"decoder.up_blocks.{i}.resnets.{j}", cur, 0, ch, co, H, nxt, 0, veps, VG, False)
            pool.put(cur)
            cur, ch = nxt, co
        if i != len(vbo) - 1:
            nxt = pool.get(2 * H, 2 * H, co)
            pb.conv(cur, 0, nxt, 0, need(V, f"decoder.up_blocks.{i}.upsamplers.0.conv.weight", (co, co, 3, 3)),
                    need(V, f"decoder.up_blocks.{i}.upsamplers.0.conv.bias", (co,)), padding=1, relu=False, ups=1)
            pool.put(cur)
            cur, H = nxt, 2 * H
    g = pool.get(H, H, ch)
    pb.group_norm(cur, g, need(V, "decoder.conv_norm_out.weight", (ch,)), need(V, "decoder.conv_norm_out.bias", (ch,)), VG, veps, True)
    pb.conv(g, 0, -1, 0, need(V, "decoder.conv_out.weight", (vcfg["out_ch"], ch, 3, 3)), need(V, "decoder.conv_out.bias", (vcfg["out_ch"],)),
            padding=1, relu=False, mode=2)
    assert H == pb.hdr["out_hw"]
    pb.unet_flops = unet_flops
    pb.vae_flops = pb.flops_per_sample - unet_flops
    return pb.finish(), pb
