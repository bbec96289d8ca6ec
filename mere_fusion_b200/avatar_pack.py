"""Avatar packer: the reference's per-avatar directory -> ONE blob that is uploaded with a
single host->device copy (or NCCL-broadcast like the weight blobs) and sliced into device views.

On-disk formats read (written by the reference's offline prep tools, which stay out of scope):
  Wav2Lip  (wav2lip/genavatar.py:101-125, loaded by lipreal.py:154-167):
      full_imgs/%08d.png   BGR full frames            face_imgs/%08d.png   BGR 96x96 face crops
      coords.pkl           list of (y1, y2, x1, x2)   (genavatar.py:96)
  MuseTalk (musetalk/mere_musetalk.py:250-317, loaded by musereal.py:165-179 and :55 `torch.load(latents)`):
      full_imgs/%08d.png   coords.pkl  list of (x1, y1, x2, y2)      latents.pt   list of [1,8,32,32] tensors
      mask/%08d.png        mask_coords.pkl  list of (x_s, y_s, x_e, y_e) crop boxes (blending.py:73-101)

What the reference does per session with these: `cv2.imread` of every PNG at start-up (seconds, per session),
then per output frame a `copy.deepcopy` of the full frame + a CPU paste (lipreal.py:207-214, musereal.py:240-248).
Here the directory is decoded ONCE by `pack_*_avatar` into a flat blob (`save_blob`; `load_blob` memory-maps it
back, no PNG decode), `DeviceAvatar` uploads it in one copy, and the paste kernels (mf_paste_resize_u8 /
mf_paste_blend_u8) read frames / masks straight from the device views -- no per-frame host copy of a full frame.

Blob layout = the container of ernerf_pack.build_blob (magic 'MFB2', kind 3, 256-byte aligned entries):
  1 META   utf-8 JSON {"head", "n", "H", "W", "S", ...}
  2 FRAMES u8  [n, H, W, 3] BGR          3 FACES  u8 [n, S, S, 3] BGR (Wav2Lip)
  4 COORDS i32 [n, 4] as stored by the reference (order differs per head, see above)
  5 LATENTS f16 [n, 8, 32, 32] (MuseTalk; musereal.py:103 casts to half before the UNet)
  6 MASKS  u8  the BGR masks back to back     7 MASK_OFF i64 [n] byte offset of each mask inside MASKS
  8 MASK_COORDS i32 [n, 4]
"""
import json
import struct

import numpy as np

from .ernerf_pack import MAGIC, build_blob

KIND_AVATAR = 3
E_META, E_FRAMES, E_FACES, E_COORDS, E_LATENTS, E_MASKS, E_MASK_OFF, E_MASK_COORDS = 1, 2, 3, 4, 5, 6, 7, 8


def _stack_u8(imgs, what):
    if len(imgs) == 0:
        raise ValueError(f"avatar has no {what}")
    shp = imgs[0].shape
    for i, im in enumerate(imgs):
        if im is None:
            raise ValueError(f"{what}[{i}] could not be decoded")
        if im.shape != shp or im.dtype != np.uint8:
            raise ValueError(f"{what}[{i}] is {im.dtype}{im.shape}, expected uint8{shp}: all {what} must share one size")
    return np.ascontiguousarray(np.stack(imgs))


def pack_lip_avatar(avatar):
    """avatar: plugin.lipreal.Avatar (frame_list_cycle, face_list_cycle, coord_list_cycle) or a directory path"""
    if isinstance(avatar, str):
        from .plugin.lipreal import Avatar
        avatar = Avatar.load(avatar)
    frames = _stack_u8(avatar.frame_list_cycle, "full_imgs")
    faces = _stack_u8(avatar.face_list_cycle, "face_imgs")
    coords = np.asarray(avatar.coord_list_cycle, np.int64)
    n = frames.shape[0]
    if faces.shape[0] != n or coords.shape != (n, 4):
        raise ValueError(f"avatar: {n} full_imgs, {faces.shape[0]} face_imgs, coords {coords.shape}")
    if faces.shape[1] != faces.shape[2]:
        raise ValueError("face crops must be square")
    H, W = frames.shape[1:3]
    y1, y2, x1, x2 = coords.T
    if (y1 < 0).any() or (x1 < 0).any() or (y2 > H).any() or (x2 > W).any() or (y2 <= y1).any() or (x2 <= x1).any():
        raise ValueError("coords.pkl: a face box lies outside its frame")
    meta = dict(head="wav2lip", n=int(n), H=int(H), W=int(W), S=int(faces.shape[1]), coords="y1,y2,x1,x2")
    return build_blob({E_META: json.dumps(meta).encode(), E_FRAMES: frames.tobytes(), E_FACES: faces.tobytes(),
                       E_COORDS: coords.astype(np.int32).tobytes()}, kind=KIND_AVATAR)


def pack_muse_avatar(avatar):
    """avatar: plugin.musereal.MuseAvatar or a directory path"""
    if isinstance(avatar, str):
        from .plugin.musereal import MuseAvatar
        avatar = MuseAvatar.load(avatar)
    frames = _stack_u8(avatar.frame_list_cycle, "full_imgs")
    n = frames.shape[0]
    H, W = frames.shape[1:3]
    coords = np.asarray(avatar.coord_list_cycle, np.int64)
    mcoords = np.asarray(avatar.mask_coords_list_cycle, np.int64)
    lats = [np.asarray(l.detach().cpu().float().numpy() if hasattr(l, "detach") else l, np.float32) for l in avatar.input_latent_list_cycle]
    lat = np.concatenate([l.reshape((-1,) + l.shape[-3:]) for l in lats], axis=0)
    if coords.shape != (n, 4) or mcoords.shape != (n, 4) or lat.shape[0] != n or len(avatar.mask_list_cycle) != n:
        raise ValueError(f"avatar: {n} full_imgs, coords {coords.shape}, mask_coords {mcoords.shape}, {lat.shape[0]} latents, "
                         f"{len(avatar.mask_list_cycle)} masks")
    if lat.shape[1:] != (8, 32, 32):
        raise ValueError(f"latents are {lat.shape[1:]}, expected (8, 32, 32) (vae.py:118-121)")
    parts, offs, off = [], [], 0
    for i, (m, (xs, ys, xe, ye)) in enumerate(zip(avatar.mask_list_cycle, mcoords)):
        m = np.ascontiguousarray(m, np.uint8)
        if m.ndim == 2:
            m = np.repeat(m[:, :, None], 3, axis=2)
        if m.shape != (ye - ys, xe - xs, 3):
            raise ValueError(f"mask[{i}] is {m.shape}, its crop box is {(ye - ys, xe - xs)} (blending.py:109-121)")
        offs.append(off)
        parts.append(m.reshape(-1))
        off += m.size
    meta = dict(head="musetalk", n=int(n), H=int(H), W=int(W), S=256, coords="x1,y1,x2,y2", mask_coords="xs,ys,xe,ye")
    return build_blob({E_META: json.dumps(meta).encode(), E_FRAMES: frames.tobytes(), E_COORDS: coords.astype(np.int32).tobytes(),
                       E_LATENTS: lat.astype(np.float16).tobytes(), E_MASKS: np.concatenate(parts).tobytes(),
                       E_MASK_OFF: np.asarray(offs, np.int64).tobytes(), E_MASK_COORDS: mcoords.astype(np.int32).tobytes()},
                      kind=KIND_AVATAR)


def save_blob(path, blob):
    np.asarray(blob, np.uint8).tofile(path)


def load_blob(path):
    """memory-mapped: nothing is decoded or copied until the upload"""
    return np.memmap(path, dtype=np.uint8, mode="r")


def parse_blob(blob):
    """host view -> {entry id: (offset, nbytes)}; raises on anything that is not an avatar blob"""
    head = bytes(np.asarray(blob[:16]))
    if len(head) < 16:
        raise ValueError("avatar blob: truncated header")
    magic, kind, version, n = struct.unpack("<IIII", head)
    if magic != MAGIC or kind != KIND_AVATAR or version != 1 or n > 64:
        raise ValueError(f"not an avatar blob (magic {magic:#x}, kind {kind}, version {version})")
    table = bytes(np.asarray(blob[16:16 + 24 * n]))
    out = {}
    for i in range(n):
        eid, _, off, nb = struct.unpack_from("<IIQQ", table, 24 * i)
        if off + nb > len(blob):
            raise ValueError("avatar blob: entry beyond the end of the blob")
        out[eid] = (off, nb)
    return out


class DeviceAvatar:
    """One packed avatar: host views for the plugin's CPU-side paths (idle frames, recording) and device views for
    the engines.  Exposes the attribute names of plugin.lipreal.Avatar / plugin.musereal.MuseAvatar, so
    `LipReal(opt, avatar=DeviceAvatar(...))` / `MuseReal(opt, avatar=...)` work unchanged."""

    def __init__(self, blob):
        self.blob = blob if isinstance(blob, np.ndarray) else np.asarray(blob)
        ent = self.entries = parse_blob(self.blob)
        o, nb = ent[E_META]
        self.meta = json.loads(bytes(self.blob[o:o + nb]).decode())
        m = self.meta
        n, H, W = m["n"], m["H"], m["W"]
        self.head = m["head"]
        self.frames = self._view(E_FRAMES, np.uint8, (n, H, W, 3))
        self.frame_list_cycle = list(self.frames)                      # views, no copy
        self.coord_list_cycle = [tuple(int(v) for v in r) for r in self._view(E_COORDS, np.int32, (n, 4))]
        if self.head == "wav2lip":
            self.faces = self._view(E_FACES, np.uint8, (n, m["S"], m["S"], 3))
            self.face_list_cycle = list(self.faces)
        else:
            self.latents = self._view(E_LATENTS, np.float16, (n, 8, 32, 32))
            self.input_latent_list_cycle = [self.latents[i:i + 1] for i in range(n)]
            self.mask_off = [int(v) for v in self._view(E_MASK_OFF, np.int64, (n,))]
            self.mask_coords_list_cycle = [tuple(int(v) for v in r) for r in self._view(E_MASK_COORDS, np.int32, (n, 4))]
            o, nb = ent[E_MASKS]
            masks = self.blob[o:o + nb]
            self.mask_list_cycle = []
            for off, (xs, ys, xe, ye) in zip(self.mask_off, self.mask_coords_list_cycle):
                self.mask_list_cycle.append(masks[off:off + (ye - ys) * (xe - xs) * 3].reshape(ye - ys, xe - xs, 3))
        self._dev = {}

    @classmethod
    def load(cls, path):
        return cls(load_blob(path))

    def _view(self, eid, dtype, shape):
        o, nb = self.entries[eid]
        want = int(np.prod(shape)) * np.dtype(dtype).itemsize
        if nb != want:
            raise ValueError(f"avatar blob entry {eid}: {nb} bytes, expected {want}")
        return self.blob[o:o + nb].view(dtype).reshape(shape)

    def device_tensors(self, device, blob_on_device=None):
        """-> dict of device views (frames, faces | latents, masks) of ONE uploaded copy of the blob.  `blob_on_device`:
        a uint8 cuda tensor that already holds the blob (e.g. received by dist.broadcast_bytes)."""
        import torch
        key = str(device)
        if key not in self._dev:
            if blob_on_device is None:
                import warnings
                with warnings.catch_warnings():                  # a memory-mapped blob is read-only; it is only read (uploaded)
                    warnings.simplefilter("ignore", UserWarning)
                    host = torch.from_numpy(np.ascontiguousarray(self.blob))
                blob_on_device = host.to(device, non_blocking=False)
            assert blob_on_device.dtype == torch.uint8 and blob_on_device.numel() == len(self.blob)
            m = self.meta

            def dv(eid, dtype, shape):
                o, nb = self.entries[eid]
                return blob_on_device[o:o + nb].view(dtype).reshape(shape)

            d = dict(blob=blob_on_device, frames=dv(E_FRAMES, torch.uint8, (m["n"], m["H"], m["W"], 3)))
            if self.head == "wav2lip":
                d["faces"] = dv(E_FACES, torch.uint8, (m["n"], m["S"], m["S"], 3))
            else:
                d["latents"] = dv(E_LATENTS, torch.float16, (m["n"], 8, 32, 32))
                o, nb = self.entries[E_MASKS]
                d["masks"] = blob_on_device[o:o + nb]
                d["mask_off"] = self.mask_off
            self._dev[key] = d
        return self._dev[key]
