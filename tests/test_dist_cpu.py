"""world_size-2 gloo tests of the N > 1 path: weight-blob broadcast from rank 0 and session sharding."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mere_fusion_b200.dist import broadcast_bytes, mixed_sessions, shard
    from helpers import seeded_wav2lip_state
    blob = None
    if rank == 0:
        # a real packed blob (single conv layer program) so the header survives the trip
        from mere_fusion_b200.convnet_pack import ProgramBuilder
        pb = ProgramBuilder(2)
        a, b = pb.buffer(8, 8, 16), pb.buffer(8, 8, 32)
        pb.conv(a, 0, b, 0, np.random.default_rng(0).standard_normal((32, 16, 3, 3)).astype(np.float32), padding=1)
        blob = pb.finish()
    t = broadcast_bytes(blob, src=0)
    import struct
    magic, kind, ver, n = struct.unpack("<IIII", t[:16].numpy().tobytes())
    mine = shard(mixed_sessions(64))
    # a packed avatar travels the same way: rank 0 packs, every rank slices its copy into views
    from mere_fusion_b200.avatar_pack import DeviceAvatar, pack_lip_avatar
    from test_plugin_cpu import _fake_avatar
    av_blob = pack_lip_avatar(_fake_avatar(n=3)) if rank == 0 else None
    at = broadcast_bytes(av_blob, src=0)
    av = DeviceAvatar(at.numpy())
    views = av.device_tensors("cpu", blob_on_device=at)
    assert views["frames"].shape == (3, 512, 512, 3) and views["frames"].data_ptr() - at.data_ptr() == av.entries[2][0]
    av_sum = int(views["frames"].sum().item()) + int(views["faces"].sum().item()) + sum(sum(c) for c in av.coord_list_cycle)
    q.put((rank, int(t.numel()), int(t.sum().item()), magic, kind, [m[0] for m in mine], av_sum))
    dist.barrier()
    dist.destroy_process_group()


def test_blob_broadcast_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, n0, s0, m0, k0, h0, a0), (r1, n1, s1, m1, k1, h1, a1) = res
    assert a0 == a1 > 0                                  # both ranks see the same avatar frames / faces / boxes
    assert n0 == n1 > 0 and s0 == s1 and m0 == m1 == 0x3242464D and k0 == k1 == 2
    assert len(h0) + len(h1) == 64 and abs(len(h0) - len(h1)) <= 1
    for h in (h0, h1):                                   # every GPU hosts all three heads
        assert {"ernerf", "musetalk", "wav2lip"} <= set(h)


def test_shard_is_a_partition():
    from mere_fusion_b200.dist import mixed_sessions, shard
    s = mixed_sessions(64)
    assert len(s) == 64 and sorted(x[0] for x in s).count("ernerf") == 22
    parts = [shard(s, 8, r) for r in range(8)]
    assert sorted(sum(parts, [])) == sorted(s) and all(len(p) == 8 for p in parts)


def _shard_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mere_fusion_b200.dist import render_stream_shard, stream_shard_plan

    class HostRenderer:
        """stands in for ErnerfRenderer on the CPU: the same two calls, the EMA of renderer.py:190-194 as the only state"""
        def __init__(self):
            self.prev = None

        def encode_audio(self, auds):
            raw = auds.mean(dim=(0, 2))[:32]
            self.prev = raw if self.prev is None else 0.35 * self.prev + 0.65 * raw
            return self.prev.clone()

        def render(self, pose, intr, H, W, auds, eye, enc_a=None):
            assert auds is None and enc_a is not None                   # sharded frames carry the feature explicitly
            return (enc_a.sum() * 1000 + float(pose) + eye).reshape(1)

    g = torch.Generator().manual_seed(5)
    frames = [(float(i), None, 4, 4, torch.randn(8, 44, 16, generator=g), 0.25) for i in range(11)]
    part = render_stream_shard(HostRenderer(), frames, world, rank)
    assert sorted(part) == stream_shard_plan(len(frames), world, rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, {k: float(v) for k, v in part.items()})      # test-only: the product exchanges nothing
    if rank == 0:
        merged = {}
        for d in gathered:
            merged.update(d)
        inorder = render_stream_shard(HostRenderer(), frames, 1, 0)
        q.put((sorted(merged) == list(range(len(frames))), all(abs(merged[i] - float(inorder[i])) == 0.0 for i in merged)))
    dist.barrier()
    dist.destroy_process_group()


def test_single_stream_frame_sharding_world2():
    """one session's frames round-robin over 2 ranks with the audio state followed on every rank: the union of the
    ranks' frames is the in-order stream (host logic on gloo; the GPU version is tests/test_ernerf_gpu.py)"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    complete, equal = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert complete and equal
