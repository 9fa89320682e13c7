"""Frame sharding of a clip over the GPUs of one box (BASELINE.json configs[3]).

Frames are independent units of the reprojection path (each frame has its own world->chassis pose;
vertices, extrinsics and intrinsics are read-only and small), so a clip of F' renderable frames is
split into contiguous blocks, one per rank, rendered without any communication, and — only if the
caller wants every frame on every rank — assembled with ONE all-gather of the uint8 frames
(``torch.distributed``: NCCL over NVLink on GPUs, gloo in the CPU tests).  The reference has no
multi-process path at all; the single-process result is the oracle for this one:
the gathered tensor must equal the one-GPU render bit for bit.
"""
from __future__ import annotations

import numpy as np


def frame_block(n_frames, rank, world_size):
    """Contiguous block [lo, hi) of rank ``rank``: ceil(n/world) frames each, the tail ranks may get
    fewer (or none).  Contiguous blocks make the all-gather output already frame-ordered."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank {rank} / world size {world_size}")
    per = -(-int(n_frames) // world_size) if n_frames > 0 else 0
    lo = min(rank * per, n_frames)
    return lo, min(lo + per, n_frames)


def block_size(n_frames, world_size):
    return -(-int(n_frames) // world_size) if n_frames > 0 else 0


def gather_frames(local_frames, n_frames, group=None):
    """All-gather the per-rank frame blocks into the whole clip, on every rank.

    local_frames  torch uint8 [n_local, ...] — this rank's block (``frame_block``), any device
    n_frames      total number of frames of the clip
    Blocks are padded to the common block size for the collective and the padding is dropped.
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = block_size(n_frames, world)
    lo, hi = frame_block(n_frames, rank, world)
    if local_frames.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} holds {local_frames.shape[0]} frames, its block has {hi - lo}")
    tail = tuple(local_frames.shape[1:])
    send = local_frames
    if hi - lo < per:
        send = torch.zeros((per,) + tail, dtype=local_frames.dtype, device=local_frames.device)
        send[:hi - lo] = local_frames
    out = torch.empty((world * per,) + tail, dtype=local_frames.dtype, device=local_frames.device)
    dist.all_gather_into_tensor(out, send.contiguous(), group=group)
    return out[:n_frames]


def gather_records(records, n, group=None):
    """All-gather the ranks' overlay records (the sparse output of cama_clip_render): ~4 % of the bytes of the
    dense frames.  Two collectives: the record counts, then the records padded to the largest count.

    records  torch int32 [>= n, words] — this rank's records, any device
    -> (torch int32 [world, n_max, words] on that device, list of the world's counts)
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    counts = torch.zeros(world, dtype=torch.int64, device=records.device)
    mine = torch.tensor([int(n)], dtype=torch.int64, device=records.device)
    dist.all_gather_into_tensor(counts, mine, group=group)
    counts = [int(c) for c in counts.tolist()]
    n_max = max(max(counts), 1)
    words = int(records.shape[1])
    if records.shape[0] >= n_max:
        send = records[:n_max]                    # (rows past n are never looked at by the receiver)
    else:
        send = torch.zeros((n_max, words), dtype=records.dtype, device=records.device)
        send[:n] = records[:n]
    out = torch.empty((world, n_max, words), dtype=records.dtype, device=records.device)
    dist.all_gather_into_tensor(out.view(world * n_max, words), send.contiguous(), group=group)
    return out, counts


def render_sharded(reproject, dataset, gather=True, group=None, mode="auto"):
    """Render this rank's block of the clip's frames; optionally assemble every block on every rank.

    gather  False: only this rank's block; True / "dense": one all-gather of the uint8 frames (north_star's
            collective: NVLink-bound, the ranks exchange every frame byte); "sparse": the ranks render the sparse
            output, all-gather the lit-chunk records and rebuild the dense frames locally with
            cama_overlay_expand — same bytes in HBM at the end, ~25x fewer over NVLink (blank backgrounds).
    -> (image_idx list of the frames returned, torch uint8 [n, C, H, W, 3] on the rank's GPU)
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    idx, w2c = reproject.frame_poses(dataset)
    lo, hi = frame_block(len(idx), rank, world)
    if gather == "sparse" and world > 1:
        r, res = reproject.renderer, reproject.resident(dataset)
        w2c_dev = torch.from_numpy(np.ascontiguousarray(w2c[lo:hi], dtype=np.float32).reshape(-1, 16)).to(reproject.rt.device)
        records, n, fmt = r.render_overlay(res, w2c_dev, mode=mode)
        everyone, counts = gather_records(records, n, group=group)
        frames = torch.empty((len(idx), r.n_cams, r.height, r.width, 3), dtype=torch.uint8, device=reproject.rt.device)
        for peer in range(world):                  # a rank's chunk indices are relative to its own block
            p_lo, p_hi = frame_block(len(idx), peer, world)
            if p_hi > p_lo:
                r.expand_overlay(everyone[peer], counts[peer], fmt, res.palette, p_hi - p_lo, out=frames[p_lo:p_hi])
        return idx, frames
    local = reproject.render_device(dataset, w2c=np.ascontiguousarray(w2c[lo:hi]), mode=mode)
    if not gather or world == 1:
        return idx[lo:hi], local
    return idx, gather_frames(local, len(idx), group=group)
