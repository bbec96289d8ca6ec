"""Multi-GPU plumbing: one process per GPU, sessions / frame batches are independent,
so the only collective is ONE broadcast of each packed weight blob from rank 0 at start-up
(NCCL over NVLink on the GPU box, gloo in the CPU tests).  Nothing is exchanged on the frame path.
"""
import numpy as np
import torch
import torch.distributed as dist


def broadcast_bytes(data, src=0, device=None):
    """data: uint8 tensor / ndarray / bytes on rank `src` (ignored elsewhere) -> uint8 tensor on every rank"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        t = data if isinstance(data, torch.Tensor) else torch.frombuffer(bytearray(bytes(data)), dtype=torch.uint8) \
            if not isinstance(data, np.ndarray) else torch.from_numpy(data)
        return t.to(device) if device is not None else t
    rank = dist.get_rank()
    backend = dist.get_backend()
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu"))
    if rank == src:
        if isinstance(data, np.ndarray):
            t = torch.from_numpy(data)
        elif isinstance(data, torch.Tensor):
            t = data
        else:
            t = torch.frombuffer(bytearray(bytes(data)), dtype=torch.uint8)
        t = t.to(dev).contiguous().view(torch.uint8)
        n = torch.tensor([t.numel()], dtype=torch.int64, device=dev)
    else:
        n = torch.zeros(1, dtype=torch.int64, device=dev)
    dist.broadcast(n, src)
    if rank != src:
        t = torch.empty(int(n.item()), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src)
    return t


def shard(items, world=None, rank=None):
    """round-robin session -> GPU assignment: item i is served by rank i % world"""
    world = world if world is not None else (dist.get_world_size() if dist.is_initialized() else 1)
    rank = rank if rank is not None else (dist.get_rank() if dist.is_initialized() else 0)
    return [it for i, it in enumerate(items) if i % world == rank]


def mixed_sessions(n=64):
    """BASELINE config 5: 22 ErNeRF + 21 MuseTalk + 21 Wav2Lip sessions, interleaved by head so that every
    GPU hosts all three heads under round-robin sharding"""
    heads = ["ernerf"] * 22 + ["musetalk"] * 21 + ["wav2lip"] * 21
    order = []
    pools = {h: [i for i, x in enumerate(heads) if x == h] for h in ("ernerf", "musetalk", "wav2lip")}
    while any(pools.values()):
        for h in ("ernerf", "musetalk", "wav2lip"):
            if pools[h]:
                order.append((h, pools[h].pop(0)))
    return order[:n]


def stream_shard_plan(n_frames, world, rank):
    """SURVEY 8(e), single-stream scaling: frame i of ONE session is rendered by rank i % world (frame batches round-robin, the
    audio window replicated); returns this rank's frame indices"""
    return [i for i in range(n_frames) if i % world == rank]


def render_stream_shard(renderer, frames, world, rank, render_kw=None):
    """One session's frames [(pose, intrinsics, H, W, auds cuda tensor, eye), ...] sharded over `world` ranks: EVERY rank follows
    the session's audio state (renderer.encode_audio on every frame: 8 small CTAs), and renders only the frames of
    stream_shard_plan with the smoothed feature passed explicitly -- so the EMA, the one per-session recurrence
    (renderer.py:190-194), is carried as an input and the frames are bit-identical to the in-order stream.  Returns {frame index:
    u8 cuda image}; nothing is exchanged between the ranks."""
    mine = set(stream_shard_plan(len(frames), world, rank))
    out = {}
    for i, (pose, intr, H, W, auds, eye) in enumerate(frames):
        enc_a = renderer.encode_audio(auds)
        if i in mine:
            out[i] = renderer.render(pose, intr, H, W, None, eye, enc_a=enc_a, **(render_kw or {}))
    return out
