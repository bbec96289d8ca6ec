"""oracle/musetalk_oracle.py -- CPU restatement (PyTorch fp32, functional) of the MuseTalk frame path:
PositionalEncoding -> diffusers UNet2DConditionModel (single step, t = 0) -> AutoencoderKL.decode ->
decode_latents post-processing.

TEST INFRASTRUCTURE ONLY: never imported by mere_fusion_b200.

PARITY UNPINNED.  The arithmetic lives in a third-party dependency that is absent from the reference
repository, from this image and from the GPU box: `diffusers` (requirements.txt:19, unpinned), classes
UNet2DConditionModel (musetalk/models/unet.py:6,37) and AutoencoderKL (musetalk/models/vae.py:1,24); the
model json (./models/musetalk/musetalk.json) and both checkpoints are external too.  This file restates the
PUBLISHED algorithm of those classes for the configuration MuseTalk v1 uses (SD-1.x UNet: block_out
320/640/1280/1280, 2 layers per block, 8 heads, cross-attention dim 384, in 8 / out 4 channels; sd-vae-ft-mse
decoder: 128/256/512/512, mid attention) with the diffusers state_dict key names, so a real checkpoint would
load strictly.  What the reference itself pins, and what the tests anchor on:
  shapes        latents [B,8,32,32] -> [B,4,32,32] (vae.py:118-121), context [B,50,384] (audio2feature.py:44),
                timesteps = [0] (musereal.py:59), output u8 [B,256,256,3] BGR (vae.py:96-108)
  PE            musetalk/models/unet.py:12-27 (restated exactly)
  post          (image / 2 + 0.5).clamp(0, 1) -> * 255 -> round -> uint8 -> RGB->BGR (vae.py:102-107)
  scaling       latents / 0.18215 before decode (vae.py:101, scaling_factor of sd-vae-ft-mse)
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

UNET_CFG = dict(block_out=(320, 640, 1280, 1280), layers=2, heads=8, ctx_dim=384, in_ch=8, out_ch=4, groups=32,
                attn=(True, True, True, False), eps=1e-5)
VAE_CFG = dict(block_out=(128, 256, 512, 512), layers=2, latent_ch=4, out_ch=3, groups=32, eps=1e-6,
               scaling_factor=0.18215)


def small_cfgs():
    """same structure, 1/5 .. 1/4 width: for tests that must finish in seconds"""
    u = dict(UNET_CFG, block_out=(64, 128, 256, 256))
    v = dict(VAE_CFG, block_out=(32, 64, 128, 128))
    return u, v


# ----------------------------------------------------------------------------------------------
# parameter shapes (diffusers key names)
# ----------------------------------------------------------------------------------------------
def _resnet_shapes(p, cin, cout, temb):
    s = {f"{p}.norm1.weight": (cin,), f"{p}.norm1.bias": (cin,), f"{p}.conv1.weight": (cout, cin, 3, 3),
         f"{p}.conv1.bias": (cout,), f"{p}.norm2.weight": (cout,), f"{p}.norm2.bias": (cout,),
         f"{p}.conv2.weight": (cout, cout, 3, 3), f"{p}.conv2.bias": (cout,)}
    if temb:
        s[f"{p}.time_emb_proj.weight"] = (cout, temb)
        s[f"{p}.time_emb_proj.bias"] = (cout,)
    if cin != cout:
        s[f"{p}.conv_shortcut.weight"] = (cout, cin, 1, 1)
        s[f"{p}.conv_shortcut.bias"] = (cout,)
    return s


def _transformer_shapes(p, c, ctx):
    t = f"{p}.transformer_blocks.0"
    s = {f"{p}.norm.weight": (c,), f"{p}.norm.bias": (c,), f"{p}.proj_in.weight": (c, c, 1, 1), f"{p}.proj_in.bias": (c,),
         f"{p}.proj_out.weight": (c, c, 1, 1), f"{p}.proj_out.bias": (c,)}
    for n in ("norm1", "norm2", "norm3"):
        s[f"{t}.{n}.weight"] = (c,)
        s[f"{t}.{n}.bias"] = (c,)
    for a, kd in (("attn1", c), ("attn2", ctx)):
        s[f"{t}.{a}.to_q.weight"] = (c, c)
        s[f"{t}.{a}.to_k.weight"] = (c, kd)
        s[f"{t}.{a}.to_v.weight"] = (c, kd)
        s[f"{t}.{a}.to_out.0.weight"] = (c, c)
        s[f"{t}.{a}.to_out.0.bias"] = (c,)
    s[f"{t}.ff.net.0.proj.weight"] = (8 * c, c)
    s[f"{t}.ff.net.0.proj.bias"] = (8 * c,)
    s[f"{t}.ff.net.2.weight"] = (c, 4 * c)
    s[f"{t}.ff.net.2.bias"] = (c,)
    return s


def unet_param_shapes(cfg=UNET_CFG):
    bo, L, ctx = cfg["block_out"], cfg["layers"], cfg["ctx_dim"]
    temb = 4 * bo[0]
    s = {"conv_in.weight": (bo[0], cfg["in_ch"], 3, 3), "conv_in.bias": (bo[0],),
         "time_embedding.linear_1.weight": (temb, bo[0]), "time_embedding.linear_1.bias": (temb,),
         "time_embedding.linear_2.weight": (temb, temb), "time_embedding.linear_2.bias": (temb,)}
    ch = bo[0]
    skips = [ch]
    for i, co in enumerate(bo):
        for j in range(L):
            s.update(_resnet_shapes(f"down_blocks.{i}.resnets.{j}", ch, co, temb))
            ch = co
            if cfg["attn"][i]:
                s.update(_transformer_shapes(f"down_blocks.{i}.attentions.{j}", co, ctx))
            skips.append(ch)
        if i != len(bo) - 1:
            s[f"down_blocks.{i}.downsamplers.0.conv.weight"] = (co, co, 3, 3)
            s[f"down_blocks.{i}.downsamplers.0.conv.bias"] = (co,)
            skips.append(ch)
    s.update(_resnet_shapes("mid_block.resnets.0", ch, ch, temb))
    s.update(_transformer_shapes("mid_block.attentions.0", ch, ctx))
    s.update(_resnet_shapes("mid_block.resnets.1", ch, ch, temb))
    rev = list(reversed(bo))
    rattn = list(reversed(cfg["attn"]))
    for i, co in enumerate(rev):
        for j in range(L + 1):
            sk = skips.pop()
            s.update(_resnet_shapes(f"up_blocks.{i}.resnets.{j}", ch + sk, co, temb))
            ch = co
            if rattn[i]:
                s.update(_transformer_shapes(f"up_blocks.{i}.attentions.{j}", co, ctx))
        if i != len(rev) - 1:
            s[f"up_blocks.{i}.upsamplers.0.conv.weight"] = (co, co, 3, 3)
            s[f"up_blocks.{i}.upsamplers.0.conv.bias"] = (co,)
    s["conv_norm_out.weight"] = (bo[0],)
    s["conv_norm_out.bias"] = (bo[0],)
    s["conv_out.weight"] = (cfg["out_ch"], bo[0], 3, 3)
    s["conv_out.bias"] = (cfg["out_ch"],)
    return s


def vae_decoder_param_shapes(cfg=VAE_CFG):
    bo, L, lc = cfg["block_out"], cfg["layers"], cfg["latent_ch"]
    top = bo[-1]
    s = {"post_quant_conv.weight": (lc, lc, 1, 1), "post_quant_conv.bias": (lc,),
         "decoder.conv_in.weight": (top, lc, 3, 3), "decoder.conv_in.bias": (top,)}
    s.update(_resnet_shapes("decoder.mid_block.resnets.0", top, top, 0))
    a = "decoder.mid_block.attentions.0"
    s[f"{a}.group_norm.weight"] = (top,)
    s[f"{a}.group_norm.bias"] = (top,)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        s[f"{a}.{n}.weight"] = (top, top)
        s[f"{a}.{n}.bias"] = (top,)
    s.update(_resnet_shapes("decoder.mid_block.resnets.1", top, top, 0))
    ch = top
    for i, co in enumerate(reversed(bo)):
        for j in range(L + 1):
            s.update(_resnet_shapes(f"decoder.up_blocks.{i}.resnets.{j}", ch, co, 0))
            ch = co
        if i != len(bo) - 1:
            s[f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"] = (co, co, 3, 3)
            s[f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"] = (co,)
    s["decoder.conv_norm_out.weight"] = (ch,)
    s["decoder.conv_norm_out.bias"] = (ch,)
    s["decoder.conv_out.weight"] = (cfg["out_ch"], ch, 3, 3)
    s["decoder.conv_out.bias"] = (cfg["out_ch"],)
    return s


def seeded_state(shapes, seed):
    """deterministic weights (numpy RNG in key order): fan-in scaled, residual-branch outputs damped so that
    a random network stays in a sane numeric range through ~60 residual additions"""
    rng = np.random.default_rng(seed)
    sd = {}
    for name, shp in shapes.items():
        if name.endswith(".weight") and len(shp) >= 2:
            fan_in = int(np.prod(shp[1:]))
            gain = 1.0
            if any(t in name for t in ("conv2.", "to_out.0", "ff.net.2", "proj_out")):
                gain = 0.4
            v = rng.standard_normal(shp) * (gain / math.sqrt(fan_in))
        elif name.endswith(".weight"):                      # norm gamma
            v = rng.uniform(0.8, 1.2, shp)
        else:                                               # biases / norm beta
            v = rng.standard_normal(shp) * 0.05
        sd[name] = torch.from_numpy(np.asarray(v, np.float32))
    return sd


# ----------------------------------------------------------------------------------------------
# forward
# ----------------------------------------------------------------------------------------------
def positional_encoding(x):
    """musetalk/models/unet.py:12-27, d_model = 384"""
    b, seq_len, d_model = x.shape
    pe = torch.zeros(seq_len, d_model)
    position = torch.arange(0, seq_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return x + pe.unsqueeze(0).to(device=x.device, dtype=x.dtype)


USE_SDPA = False     # the GPU baseline leg of bench.py sets this: diffusers' AttnProcessor2_0 calls F.scaled_dot_product_attention


def timestep_embedding_t0(dim):
    """diffusers get_timestep_embedding(timesteps=[0], dim, flip_sin_to_cos=True, downscale_freq_shift=0):
    emb = t * exp(...) = 0 -> [sin, cos] = [0, 1] -> flipped to [cos | sin] = [1 .. 1 | 0 .. 0]"""
    half = dim // 2
    return torch.cat([torch.ones(1, half), torch.zeros(1, half)], dim=-1)


def _gn(sd, p, x, groups, eps):
    return F.group_norm(x, groups, sd[p + ".weight"], sd[p + ".bias"], eps)


def resnet(sd, p, x, temb, groups, eps):
    """diffusers ResnetBlock2D (output_scale_factor 1, dropout 0, time_embedding_norm 'default')"""
    h = F.conv2d(F.silu(_gn(sd, p + ".norm1", x, groups, eps)), sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    if temb is not None and (p + ".time_emb_proj.weight") in sd:
        h = h + F.linear(F.silu(temb), sd[p + ".time_emb_proj.weight"], sd[p + ".time_emb_proj.bias"])[:, :, None, None]
    h = F.conv2d(F.silu(_gn(sd, p + ".norm2", h, groups, eps)), sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    if (p + ".conv_shortcut.weight") in sd:
        x = F.conv2d(x, sd[p + ".conv_shortcut.weight"], sd[p + ".conv_shortcut.bias"])
    return x + h


def attention(sd, p, x, ctx, heads, bias_qkv=False):
    """diffusers Attention: q/k/v projections, softmax(q k^T / sqrt(d)) v, output projection"""
    B, N, C = x.shape
    d = C // heads

    def lin(name, t):
        return F.linear(t, sd[f"{p}.{name}.weight"], sd.get(f"{p}.{name}.bias") if bias_qkv else None)

    q = lin("to_q", x).view(B, N, heads, d).transpose(1, 2)
    k = lin("to_k", ctx).view(B, ctx.shape[1], heads, d).transpose(1, 2)
    v = lin("to_v", ctx).view(B, ctx.shape[1], heads, d).transpose(1, 2)
    if USE_SDPA:
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, N, C)
    else:
        a = torch.softmax(q @ k.transpose(-1, -2) * (d ** -0.5), dim=-1)
        o = (a @ v).transpose(1, 2).reshape(B, N, C)
    return F.linear(o, sd[f"{p}.to_out.0.weight"], sd[f"{p}.to_out.0.bias"])


def transformer2d(sd, p, x, ctx, heads, groups):
    """diffusers Transformer2DModel (conv projections) with one BasicTransformerBlock (GEGLU feed-forward)"""
    B, C, H, W = x.shape
    res = x
    h = F.conv2d(F.group_norm(x, groups, sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-6), sd[p + ".proj_in.weight"],
                 sd[p + ".proj_in.bias"])
    h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
    t = p + ".transformer_blocks.0"
    n = F.layer_norm(h, (C,), sd[t + ".norm1.weight"], sd[t + ".norm1.bias"], 1e-5)
    h = h + attention(sd, t + ".attn1", n, n, heads)
    n = F.layer_norm(h, (C,), sd[t + ".norm2.weight"], sd[t + ".norm2.bias"], 1e-5)
    h = h + attention(sd, t + ".attn2", n, ctx, heads)
    n = F.layer_norm(h, (C,), sd[t + ".norm3.weight"], sd[t + ".norm3.bias"], 1e-5)
    g = F.linear(n, sd[t + ".ff.net.0.proj.weight"], sd[t + ".ff.net.0.proj.bias"])
    a, gate = g.chunk(2, dim=-1)
    h = h + F.linear(a * F.gelu(gate), sd[t + ".ff.net.2.weight"], sd[t + ".ff.net.2.bias"])
    h = h.reshape(B, H, W, C).permute(0, 3, 1, 2)
    return F.conv2d(h, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"]) + res


def unet_forward(sd, latents, ctx, cfg=UNET_CFG):
    """UNet2DConditionModel.forward(sample, timestep=0, encoder_hidden_states).sample"""
    bo, L, heads, G, eps = cfg["block_out"], cfg["layers"], cfg["heads"], cfg["groups"], cfg["eps"]
    temb = F.linear(F.silu(F.linear(timestep_embedding_t0(bo[0]).to(latents), sd["time_embedding.linear_1.weight"],
                                    sd["time_embedding.linear_1.bias"])),
                    sd["time_embedding.linear_2.weight"], sd["time_embedding.linear_2.bias"])
    x = F.conv2d(latents, sd["conv_in.weight"], sd["conv_in.bias"], padding=1)
    skips = [x]
    for i in range(len(bo)):
        for j in range(L):
            x = resnet(sd, f"down_blocks.{i}.resnets.{j}", x, temb, G, eps)
            if cfg["attn"][i]:
                x = transformer2d(sd, f"down_blocks.{i}.attentions.{j}", x, ctx, heads, G)
            skips.append(x)
        if i != len(bo) - 1:
            x = F.conv2d(x, sd[f"down_blocks.{i}.downsamplers.0.conv.weight"], sd[f"down_blocks.{i}.downsamplers.0.conv.bias"],
                         stride=2, padding=1)
            skips.append(x)
    x = resnet(sd, "mid_block.resnets.0", x, temb, G, eps)
    x = transformer2d(sd, "mid_block.attentions.0", x, ctx, heads, G)
    x = resnet(sd, "mid_block.resnets.1", x, temb, G, eps)
    rattn = list(reversed(cfg["attn"]))
    for i in range(len(bo)):
        for j in range(L + 1):
            x = torch.cat([x, skips.pop()], dim=1)
            x = resnet(sd, f"up_blocks.{i}.resnets.{j}", x, temb, G, eps)
            if rattn[i]:
                x = transformer2d(sd, f"up_blocks.{i}.attentions.{j}", x, ctx, heads, G)
        if i != len(bo) - 1:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = F.conv2d(x, sd[f"up_blocks.{i}.upsamplers.0.conv.weight"], sd[f"up_blocks.{i}.upsamplers.0.conv.bias"], padding=1)
    x = F.silu(_gn(sd, "conv_norm_out", x, G, eps))
    return F.conv2d(x, sd["conv_out.weight"], sd["conv_out.bias"], padding=1)


def vae_decode(sd, z, cfg=VAE_CFG):
    """AutoencoderKL.decode(z).sample: post_quant_conv -> Decoder"""
    bo, L, G, eps = cfg["block_out"], cfg["layers"], cfg["groups"], cfg["eps"]
    x = F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    x = F.conv2d(x, sd["decoder.conv_in.weight"], sd["decoder.conv_in.bias"], padding=1)
    x = resnet(sd, "decoder.mid_block.resnets.0", x, None, G, eps)
    a = "decoder.mid_block.attentions.0"
    B, C, H, W = x.shape
    h = F.group_norm(x, G, sd[a + ".group_norm.weight"], sd[a + ".group_norm.bias"], eps)
    h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
    h = attention(sd, a, h, h, 1, bias_qkv=True)
    x = x + h.reshape(B, H, W, C).permute(0, 3, 1, 2)
    x = resnet(sd, "decoder.mid_block.resnets.1", x, None, G, eps)
    for i in range(len(bo)):
        for j in range(L + 1):
            x = resnet(sd, f"decoder.up_blocks.{i}.resnets.{j}", x, None, G, eps)
        if i != len(bo) - 1:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = F.conv2d(x, sd[f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"],
                         sd[f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"], padding=1)
    x = F.silu(_gn(sd, "decoder.conv_norm_out", x, G, eps))
    return F.conv2d(x, sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"], padding=1)


def infer(unet_sd, vae_sd, latents, whisper, ucfg=UNET_CFG, vcfg=VAE_CFG):
    """musereal.py:99-108 + vae.py:96-108: returns (pred_latents, image fp32 [B,256,256,3] RGB in [0,1], u8 BGR)"""
    with torch.no_grad():
        ctx = positional_encoding(torch.as_tensor(whisper, dtype=torch.float32))
        pred = unet_forward(unet_sd, torch.as_tensor(latents, dtype=torch.float32), ctx, ucfg)
        img = vae_decode(vae_sd, pred / vcfg["scaling_factor"], vcfg)
        img = (img / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1)
    u8 = (img.numpy() * 255).round().astype("uint8")[..., ::-1]
    return pred.numpy(), img.numpy(), np.ascontiguousarray(u8)


def infer_device(unet_sd, vae_sd, latents, whisper, ucfg=UNET_CFG, vcfg=VAE_CFG):
    """the same arithmetic with every tensor where the caller put it (device, dtype): the same-box PyTorch GPU baseline of
    bench.py (`heads.musetalk.torch_gpu`; the reference runs pe / unet / vae in fp16 on the GPU, musereal.py:60-62).
    Returns the u8 BGR frames as a device tensor (vae.py:102-107)."""
    with torch.no_grad():
        pred = unet_forward(unet_sd, latents, positional_encoding(whisper), ucfg)
        img = vae_decode(vae_sd, pred / vcfg["scaling_factor"], vcfg)
        img = (img / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1).float()
        return (img * 255).round().to(torch.uint8).flip(-1)
