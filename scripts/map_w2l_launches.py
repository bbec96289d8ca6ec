"""Map an ncu launch list of scripts/time_wav2lip.py (gpu__time_duration per launch) onto the Wav2Lip program's ops and print the time /
achieved rate per op (analysis helper, not the bench).  usage: map_w2l_launches.py launches.csv [face_hw]"""
import csv
import struct
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import mere_fusion_b200.convnet_pack as cp                      # noqa: E402


def fake(self, data):                                           # skip the weight bytes: only the op list is needed
    i = self.next_id
    self.next_id += 1
    self.tensors[i] = b""
    return i


cp.ProgramBuilder._tensor = fake
cp.f32_to_bf16_bits = lambda a: np.zeros(1, np.uint16)
from mere_fusion_b200.wav2lip_pack import pack_wav2lip, wav2lip_param_shapes   # noqa: E402

S = int(sys.argv[2]) if len(sys.argv) > 2 else 96
B = 16
sd = {k: np.zeros(s, np.float32) for k, s in wav2lip_param_shapes(S).items()}
blob, pb = pack_wav2lip(sd, nominal_batch=B, face_hw=S)
ops = [struct.unpack("<28i", rec[:112]) for rec in pb.ops]
rows = []
for row in csv.DictReader([l for l in open(sys.argv[1]) if not l.startswith("==")]):
    try:
        t = float(row["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    if row["Metric Unit"] in ("ns", "nsecond"):
        t /= 1000
    rows.append((row["Kernel Name"].replace("void ", "")[:22], t))
starts = [i for i, r in enumerate(rows) if r[0].startswith("k_prep_face")]
st = starts[-2]
fw = rows[st:starts[-1]]
print(f"one forward: {len(fw)} launches, {sum(t for _, t in fw):.1f} us (ncu: cold caches, serialised)")
for i, o in enumerate(ops):
    n, t = fw[2 + i]
    Mh, Mw, ntaps, cin, cout = o[6], o[7], o[14], o[15], o[17]
    fl = 2 * B * Mh * Mw * ntaps * cin * cout
    print(f"{i:3d} {n:22s} {t:8.1f} us  M={Mh}x{Mw} taps={ntaps} cin={cin} cout={cout} bn={o[19]} stride={o[12]}  {fl / t / 1e6:8.1f} TF/s (padded MACs)")
