"""Frame sharding + the all-gather of rendered frames, world size 2 and 3, gloo on CPU.

The render itself needs a GPU (tests/test_gpu_parity.py::test_frame_sharding_is_exact proves that
blocks rendered separately equal the whole clip); here the host-side partition and the collective
are exercised with stand-in frames."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cama_b200 import shard


def test_frame_block_partition():
    for n in (0, 1, 5, 39, 40, 41, 320):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard.frame_block(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))          # contiguous, ordered
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) == shard.block_size(n, world) and all(s >= 0 for s in sizes)
    with pytest.raises(ValueError):
        shard.frame_block(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_frames, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        whole = torch.from_numpy(rng.integers(0, 256, size=(n_frames, 2, 6, 8, 3), dtype=np.uint8))   # every rank: same "clip"
        lo, hi = shard.frame_block(n_frames, rank, world)
        got = shard.gather_frames(whole[lo:hi].clone(), n_frames)
        ok = bool((got == whole).all()) and tuple(got.shape) == tuple(whole.shape)
        try:                                       # a block of the wrong size is refused before any communication
            shard.gather_frames(torch.zeros((hi - lo + 1, 2, 6, 8, 3), dtype=torch.uint8), n_frames)
            ok = False
        except ValueError:
            pass
        results[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_frames", [(2, 5), (2, 40), (3, 7), (3, 2)])
def test_gather_frames_gloo(world, n_frames):
    ctx = mp.get_context("spawn")
    with ctx.Manager() as manager:
        results = manager.dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, results)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert dict(results) == {r: True for r in range(world)}


def _records_worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ok = True
        for words in (3, 8):
            counts = [5 + 3 * r for r in range(world)]                  # ragged: every rank has another count
            mine = torch.arange(counts[rank] * words, dtype=torch.int32).reshape(-1, words) + 1000 * rank
            padded = torch.cat([mine, torch.full((4, words), -1, dtype=torch.int32)])     # capacity > count, like the device buffer
            got, got_counts = shard.gather_records(padded, counts[rank])
            ok = ok and got_counts == counts and tuple(got.shape) == (world, max(counts), words)
            for r in range(world):
                want = torch.arange(counts[r] * words, dtype=torch.int32).reshape(-1, words) + 1000 * r
                ok = ok and bool((got[r, :counts[r]] == want).all())
        results[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gather_records_gloo(world):
    """The sparse all-gather's collective part: ragged record lists of every rank arrive intact on every rank."""
    ctx = mp.get_context("spawn")
    with ctx.Manager() as manager:
        results = manager.dict()
        port = _free_port()
        procs = [ctx.Process(target=_records_worker, args=(r, world, port, results)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
        assert all(p.exitcode == 0 for p in procs)
        assert dict(results) == {r: True for r in range(world)}


def test_peer_mailbox_layout():
    """[parities][world sources] slots of header | records, 256-byte aligned, disjoint."""
    for world, cap, rec in ((2, 1000, 12), (8, 2187629, 12), (3, 0, 32), (8, 5, 32)):
        slot, total, offset = shard.slot_layout(world, cap, rec)
        assert slot % 256 == 0 and slot >= 256 + cap * rec and total == shard.PARITIES * world * slot
        starts = sorted(offset(p, s) for p in range(shard.PARITIES) for s in range(world))
        assert starts == [k * slot for k in range(shard.PARITIES * world)]


# ---------------------------------------------------------------------------------------------- SiteAssembler: one decision on all ranks
class _StubRuntime:
    device = torch.device("cpu")


class _StubRenderer:
    n_cams, height, width = 6, 540, 960

    def __init__(self, rank):
        self.rank, self.last_stats = rank, None

    def render_overlay(self, res, w2c_dev, mode="auto"):
        n_frames = int(w2c_dev.shape[0])
        if n_frames == 0:                                    # (a rank with an empty block renders nothing and has no statistics)
            return None, 0, 1
        self.last_stats = {"record_capacity": 8192, "record_capacity_needed": 5000 + 1000 * self.rank, "lists_per_image": 34}
        return None, 1000 * (self.rank + 1), 1


class _StubReproject:
    def __init__(self, rank, n_frames):
        self.renderer, self.rt, self.n = _StubRenderer(rank), _StubRuntime(), n_frames

    def resident(self, dataset):
        return object()

    def frame_poses(self, dataset):
        return list(range(1, self.n + 1)), np.zeros((self.n, 16), np.float32)


def _assembler_worker(rank, world, port, n_frames, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        made = {}

        class _StubExchange:                                 # stands in for ListExchange / PeerExchange (both need CUDA IPC)
            available, error = True, None

            def __init__(self, *args, **kwargs):
                made["args"] = args

        saved = shard.ListExchange, shard.PeerExchange
        shard.ListExchange = shard.PeerExchange = _StubExchange
        try:
            asm = shard.SiteAssembler(_StubReproject(rank, n_frames), "nuscenes", exchange="lists")
        finally:
            shard.ListExchange, shard.PeerExchange = saved
        # args of ListExchange: (rt, renderer, res, n_frames_total, capacity, lists_per_image)
        results[rank] = (asm.kind, asm.lo, asm.hi, made["args"][3], made["args"][4], made["args"][5])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_frames", [(3, 2), (2, 5)])
def test_site_assembler_decides_collectively(world, n_frames):
    """Every rank must build the same exchange with the same list capacity — also a rank whose frame block is empty
    (3 ranks, 2 frames) and therefore has no statistics of its own: a rank deciding alone would wait for peers that took
    the other path."""
    ctx = mp.get_context("spawn")
    with ctx.Manager() as manager:
        results = manager.dict()
        port = _free_port()
        procs = [ctx.Process(target=_assembler_worker, args=(r, world, port, n_frames, results)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=120)
        assert all(p.exitcode == 0 for p in procs)
        got = [results[r] for r in range(world)]
    assert {g[0] for g in got} == {"lists"}
    assert [(g[1], g[2]) for g in got] == [shard.frame_block(n_frames, r, world) for r in range(world)]
    ranks_with_frames = [r for r in range(world) if shard.frame_block(n_frames, r, world)[1] > shard.frame_block(n_frames, r, world)[0]]
    needed = 5000 + 1000 * max(ranks_with_frames)
    assert len({g[3:] for g in got}) == 1 and got[0][3] == n_frames and got[0][4] == max(8192, int(needed * 1.1) + 256) and got[0][5] == 34
