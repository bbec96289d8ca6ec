"""Map an ncu launch list of scripts/time_musetalk.py (gpu__time_duration per launch) onto the program's ops and print the
time / achieved rate per distinct op shape (analysis helper, not the bench).  usage: map_launches.py launches.csv"""
import csv
import struct
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
sys.path.insert(0, __file__.rsplit("/", 2)[0] + "/tests")
import mere_fusion_b200.convnet_pack as cp                      # noqa: E402
from mere_fusion_b200.musetalk_pack import pack_musetalk        # noqa: E402
from oracle import musetalk_oracle as M                         # noqa: E402


def fake(self, data):                                           # skip the weight bytes: only the op list is needed
    i = self.next_id
    self.next_id += 1
    self.tensors[i] = b""
    return i


cp.ProgramBuilder._tensor = fake
cp.f32_to_bf16_bits = lambda a: np.zeros(1, np.uint16)
u, v = M.UNET_CFG, M.VAE_CFG
usd = {k: np.zeros(s, np.float32) for k, s in M.unet_param_shapes(u).items()}
vsd = {k: np.zeros(s, np.float32) for k, s in M.vae_decoder_param_shapes(v).items()}
blob, pb = pack_musetalk(usd, vsd, u, v, nominal_batch=16)
ops = [struct.unpack("<28i", rec[:112]) for rec in pb.ops]
rows = []
for row in csv.DictReader([l for l in open(sys.argv[1]) if not l.startswith("==")]):
    try:
        t = float(row["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    if row["Metric Unit"] in ("ns", "nsecond"):
        t /= 1000
    rows.append((row["Kernel Name"].replace("void ", "")[:14], t))
start = next(i for i, r in enumerate(rows) if r[0].startswith("k_prep_latents"))
nxt = next(i for i, r in enumerate(rows) if i > start and r[0].startswith("k_prep_latents")) if sum(r[0].startswith("k_prep_latents") for r in rows) > 1 else None
n_launch = (nxt - start) if nxt else int(sys.argv[3])
# the forward is periodic in the launch list: positions past the end of the capture come from the previous forward
one = [rows[start + j] if start + j < len(rows) else rows[start + j - n_launch] for j in range(n_launch)]
idx, B, out = 2, 16, []
for f in ops:
    (in_buf, in_coff, out_buf, out_coff, rb, rc, Mh, Mw, oy0, ox0, osy, osx, isy, isx, ntaps, Cin, Kpad, Cout, Cout_pad, BN, relu, mode,
     w, s_, h, kind, ups, flags) = f
    name = one[idx][0]
    n = {0: 1, 1: 1 if name.startswith("k_gn_small") else 2, 2: 1, 3: 1 if name.startswith("k_flash") else 3, 4: 1}[kind]
    t = sum(x[1] for x in one[idx:idx + n])
    idx += n
    if kind == 0:
        fl = 2 * Cout * Cin * ntaps * Mh * Mw * B
        out.append((t, f"conv Cin{Cin} Cout{Cout} taps{ntaps} M{Mh}x{Mw} os{osy} {name}", fl / t / 1e6, "TF/s"))
    elif kind == 1:
        H, W, C = pb.buffers[out_buf]
        out.append((t, f"GN C{C} {H}x{W}", 3 * H * W * C * 2 * B / t / 1e3, "GB/s (3x tensor)"))
    elif kind == 3:
        nq, nk = pb.buffers[in_buf][0] * pb.buffers[in_buf][1], pb.buffers[rb][0] * pb.buffers[rb][1]
        out.append((t, f"attn heads{ntaps} dh{Cin} nq{nq} nk{nk}", 4 * ntaps * Cin * nq * nk * B / t / 1e6, "TF/s"))
    elif kind == 2:
        H, W, C = pb.buffers[out_buf]
        out.append((t, f"LN C{C} {H}x{W}", 2 * H * W * C * 2 * B / t / 1e3, "GB/s"))
    else:
        H, W, C = pb.buffers[out_buf]
        out.append((t, f"GEGLU C{C} {H}x{W}", 3 * H * W * C * 2 * B / t / 1e3, "GB/s"))
agg = {}
for t, d, r, unit in out:
    a = agg.setdefault(d, [0, 0.0, r, unit])
    a[0] += 1
    a[1] += t
print(f"one forward: {sum(x[1] for x in one):.0f} us over {n_launch} launches")
for d, (n, t, r, unit) in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{t:9.1f} us x{n:2d}  {d:60s} {r:8.0f} {unit}")
