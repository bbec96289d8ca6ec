"""Avatar packer: the reference's on-disk avatar directories (wav2lip/genavatar.py:101-125,
musetalk/mere_musetalk.py:250-317) -> one blob -> the same lists the reference's loaders produce
(lipreal.py:154-167, musereal.py:165-179), byte for byte."""
import os
import pickle

import numpy as np
import pytest

from test_plugin_cpu import _fake_avatar, _fake_muse_avatar

cv2 = pytest.importorskip("cv2")
torch = pytest.importorskip("torch")

from mere_fusion_b200 import avatar_pack as AP


def _write_lip_dir(root, av):
    os.makedirs(root / "full_imgs"), os.makedirs(root / "face_imgs")
    for i, (f, c) in enumerate(zip(av.frame_list_cycle, av.face_list_cycle)):
        cv2.imwrite(str(root / "full_imgs" / f"{i:08d}.png"), f)            # genavatar.py:108 video2imgs, :121
        cv2.imwrite(str(root / "face_imgs" / f"{i:08d}.png"), c)
    with open(root / "coords.pkl", "wb") as fh:
        pickle.dump(av.coord_list_cycle, fh)                                # genavatar.py:125


def _write_muse_dir(root, av):
    os.makedirs(root / "full_imgs"), os.makedirs(root / "mask")
    for i, (f, m) in enumerate(zip(av.frame_list_cycle, av.mask_list_cycle)):
        cv2.imwrite(str(root / "full_imgs" / f"{str(i).zfill(8)}.png"), f)  # mere_musetalk.py:303
        cv2.imwrite(str(root / "mask" / f"{str(i).zfill(8)}.png"), m)       # :306
    with open(root / "mask_coords.pkl", "wb") as fh:
        pickle.dump(av.mask_coords_list_cycle, fh)
    with open(root / "coords.pkl", "wb") as fh:
        pickle.dump(av.coord_list_cycle, fh)
    torch.save([torch.from_numpy(l) for l in av.input_latent_list_cycle], root / "latents.pt")   # :316


def test_lip_avatar_directory_round_trip(tmp_path):
    from mere_fusion_b200.plugin.lipreal import Avatar
    src = _fake_avatar(n=7, H=120, W=160)
    src.coord_list_cycle = [(10 + i, 90 + i, 20, 110) for i in range(7)]
    _write_lip_dir(tmp_path / "av", src)
    ref = Avatar.load(str(tmp_path / "av"))                                 # the reference's loading order (sorted by int(stem))
    blob = AP.pack_lip_avatar(str(tmp_path / "av"))
    AP.save_blob(tmp_path / "av.mfav", blob)
    av = AP.DeviceAvatar.load(tmp_path / "av.mfav")
    assert av.head == "wav2lip" and av.meta["n"] == 7 and av.meta["S"] == 96
    assert av.coord_list_cycle == [tuple(c) for c in ref.coord_list_cycle]
    for a, b in zip(av.frame_list_cycle, ref.frame_list_cycle):
        assert np.array_equal(a, b)
    for a, b in zip(av.face_list_cycle, ref.face_list_cycle):
        assert np.array_equal(a, b)
    t = av.device_tensors("cpu")                                            # the same slicing the GPU upload uses
    assert t["frames"].shape == (7, 120, 160, 3) and np.array_equal(t["faces"].numpy(), np.stack(ref.face_list_cycle))
    assert t["frames"].data_ptr() - t["blob"].data_ptr() == av.entries[AP.E_FRAMES][0]     # views of ONE buffer, no copies
    assert av.device_tensors("cpu") is t


def test_muse_avatar_directory_round_trip(tmp_path):
    from mere_fusion_b200.plugin.musereal import MuseAvatar
    src = _fake_muse_avatar(n=5, H=512, W=512)
    _write_muse_dir(tmp_path / "mv", src)
    ref = MuseAvatar.load(str(tmp_path / "mv"))
    av = AP.DeviceAvatar(AP.pack_muse_avatar(str(tmp_path / "mv")))
    assert av.head == "musetalk" and av.coord_list_cycle == [tuple(c) for c in ref.coord_list_cycle]
    assert av.mask_coords_list_cycle == [tuple(c) for c in ref.mask_coords_list_cycle]
    for a, b in zip(av.frame_list_cycle, ref.frame_list_cycle):
        assert np.array_equal(a, b)
    for a, b in zip(av.mask_list_cycle, ref.mask_list_cycle):
        assert np.array_equal(a, b)
    for a, b in zip(av.input_latent_list_cycle, ref.input_latent_list_cycle):
        assert a.shape == (1, 8, 32, 32) and np.array_equal(a, b.numpy().astype(np.float16))   # musereal.py:103 .half()
    t = av.device_tensors("cpu")
    off = av.mask_off[3]
    xs, ys, xe, ye = av.mask_coords_list_cycle[3]
    assert np.array_equal(t["masks"][off:off + (ye - ys) * (xe - xs) * 3].numpy().reshape(ye - ys, xe - xs, 3), ref.mask_list_cycle[3])
    assert t["latents"].dtype == torch.float16 and t["latents"].shape == (5, 8, 32, 32)


def test_packer_refuses_inconsistent_avatars(tmp_path):
    av = _fake_avatar(n=3, H=64, W=64)
    av.coord_list_cycle = [(0, 96, 0, 32)] * 3                              # box taller than the frame
    with pytest.raises(ValueError, match="outside"):
        AP.pack_lip_avatar(av)
    av = _fake_avatar(n=3, H=64, W=64)
    av.coord_list_cycle = [(0, 32, 0, 32)] * 3
    av.frame_list_cycle[1] = av.frame_list_cycle[1][:60]
    with pytest.raises(ValueError, match="share one size"):
        AP.pack_lip_avatar(av)
    mv = _fake_muse_avatar(n=3)
    mv.mask_list_cycle[2] = mv.mask_list_cycle[2][:-1]
    with pytest.raises(ValueError, match="crop box"):
        AP.pack_muse_avatar(mv)
    blob = AP.pack_muse_avatar(_fake_muse_avatar(n=2))
    bad = blob.copy()
    bad[4] = 2                                                              # kind: a conv-net program, not an avatar
    with pytest.raises(ValueError, match="not an avatar blob"):
        AP.DeviceAvatar(bad)


def test_lipreal_accepts_a_packed_avatar():
    """plumbing: LipReal's host-side paths (idle frames, cv2 paste) read the packed avatar's views"""
    from test_plugin_cpu import make_opt
    from mere_fusion_b200.plugin.lipreal import LipReal

    class Eng:
        max_batch = 16
    packed = AP.DeviceAvatar(AP.pack_lip_avatar(_fake_avatar()))
    real = LipReal(make_opt(), engine=Eng(), avatar=packed, mel="host")
    assert len(real.frame_list_cycle) == 25 and real.coord_list_cycle[0] == (176, 368, 160, 352)
    assert np.array_equal(real.frame_list_cycle[3], _fake_avatar().frame_list_cycle[3])
