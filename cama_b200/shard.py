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

import ctypes

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
            output, all-gather the lit-chunk records (NCCL) and rebuild the dense frames locally with
            cama_overlay_expand — same bytes in HBM at the end, ~25x fewer over NVLink (blank backgrounds);
            "peer": the same exchange without NCCL and without the host — the raster mirrors its records into
            the peers' memory while it runs (SiteAssembler / PeerExchange below); falls back to "sparse" on a box
            without a peer-to-peer path.
    -> (image_idx list of the frames returned, torch uint8 [n, C, H, W, 3] on the rank's GPU)
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    idx, w2c = reproject.frame_poses(dataset)
    lo, hi = frame_block(len(idx), rank, world)
    if gather == "peer" and world > 1:
        cache = reproject.__dict__.setdefault("_site_assemblers", {})
        key = (dataset, id(group), mode)
        if key not in cache:
            cache[key] = SiteAssembler(reproject, dataset, group=group, mode=mode)
        if cache[key].available:
            frames = cache[key].step()
            code = cache[key].exchange.status_code()
            if code:
                raise RuntimeError(f"peer assembly failed with status {code} (1: a peer's step timed out, 2: a slot overflowed)")
            return idx, frames
        gather = "sparse"
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


# ---------------------------------------------------------------------------------------------- peer exchange
def _MODES_BINNED(mode):
    return mode in ("auto", "binned")


PARITIES = 4        # slots a source cycles through in every mailbox (see PeerExchange.render_and_assemble)


def slot_layout(world_size, capacity_records, record_bytes, header_bytes=256, parities=PARITIES):
    """Byte layout of a rank's mailbox: [parities][world sources] slots of header | records.
    -> (slot_bytes, total_bytes, offset(parity, source))"""
    slot = -(-(header_bytes + int(capacity_records) * int(record_bytes)) // 256) * 256
    return slot, parities * world_size * slot, (lambda parity, source: (parity * world_size + source) * slot)


class PeerExchange:
    """The sparse output of a frame-sharded clip, exchanged between the GPUs of one box through peer memory.

    Every rank owns a mailbox (``cama_peer_alloc``) that all ranks map (CUDA IPC handles travel through
    ``torch.distributed.all_gather_object``, once).  Per step a rank renders its frame block with the sparse output
    pointed at its own slot and mirrored into the same slot on every peer (the raster's flushes cross NVLink while it
    computes), publishes the slot, and expands the slots of all ranks into dense frames.  Nothing on the data path
    goes through the host or through NCCL.  ``available`` is False when the box has no peer-to-peer path (the caller
    then falls back to ``render_sharded(gather="sparse")``: NCCL all-gather of the records).
    """

    def __init__(self, rt, capacity_records, fmt, group=None, devices=None):
        import torch
        import torch.distributed as dist
        from . import _native as N
        self.rt, self.group, self.fmt = rt, group, fmt
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > N.MAX_PEERS:
            raise ValueError(f"at most {N.MAX_PEERS} ranks")
        self.record_bytes = N.OVERLAY_RECORD_BYTES[fmt]
        self.capacity = int(capacity_records)
        self.slot_bytes, self.total_bytes, self.offset = slot_layout(self.world, self.capacity, self.record_bytes, N.PEER_HEADER_BYTES)
        self.step = 0
        self.status = torch.zeros(1, dtype=torch.int32, device=rt.device)
        self.count = torch.zeros(4, dtype=torch.int32, device=rt.device)
        self._expanded = {}                         # step -> event recorded after its expand (render_and_assemble)
        self.base = [None] * self.world             # mailbox of every rank, as mapped into this process
        self._own = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(N.PEER_HANDLE_BYTES)
        error = None
        try:
            N.check(N.lib().cama_peer_alloc(rt.ctx, self.total_bytes, ctypes.byref(self._own), handle))
        except N.CamaError as exc:                  # (every rank still takes part in the collectives below)
            error = str(exc)
        device_index = rt.device.index if devices is None else devices[self.rank]
        mine = {"handle": bytes(handle.raw), "device": int(device_index), "error": error}
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        if all(e["error"] is None for e in everyone):
            for r, e in enumerate(everyone):
                if r == self.rank:
                    self.base[r] = self._own.value
                    continue
                ptr = ctypes.c_void_p()
                code = N.lib().cama_peer_open(rt.ctx, e["device"], e["handle"], ctypes.byref(ptr))
                if code != N.CAMA_OK:
                    error = N.lib().cama_last_error().decode("utf-8", "replace")
                    break
                self.base[r] = ptr.value
        else:
            error = error or "a peer could not allocate its mailbox"
        verdicts = [None] * self.world
        dist.all_gather_object(verdicts, error, group=group)
        self.error = next((v for v in verdicts if v), None)
        self.available = self.error is None
        if not self.available:
            self.close()

    def close(self):
        from . import _native as N
        for r, ptr in enumerate(self.base):
            if ptr is not None and r != self.rank:
                N.lib().cama_peer_close(self.rt.ctx, ptr)
        self.base = [None] * self.world
        if self._own.value:
            N.lib().cama_peer_free(self.rt.ctx, self._own)
            self._own = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:                            # (interpreter shutdown: the driver frees the memory anyway)
            pass

    def slot(self, owner, parity, source):
        """device pointer (in this process) of the slot of `source`'s records in `owner`'s mailbox"""
        return self.base[owner] + self.offset(parity, source)

    def render_and_assemble(self, renderer, res, w2c_dev, frame_lo, n_frames_total, out, mode="auto", render_stream=None):
        """One step: render this rank's block (frames [frame_lo, frame_lo + len(w2c_dev)) of the clip), exchange, expand.

        out            torch uint8 [n_frames_total, C, H, W, 3] on this device: all frames of the clip afterwards
        render_stream  optional torch stream for the render + publish; zero-fill and expand stay on the current stream.
                       The two streams form a two-stage pipeline over consecutive steps: the render of step s+1 only
                       waits for the expand of step s-1, so it runs while step s is still being filled and expanded
                       (inside one step the two do not overlap usefully: the fill saturates the memory system and the
                       latency-bound render stretches by what it gains).  Four slot parities make that safe: a rank's
                       render of step s+4 — the next writer of the slots of step s on every peer — waits for its own
                       expand of step s+2, which has seen every peer's step s+2, which those peers rendered after
                       their expand of step s.
        Asynchronous; check ``status_code()`` after synchronising.
        """
        import torch
        from . import _native as N
        rt = self.rt
        self.step += 1
        parity = self.step % PARITIES
        main = torch.cuda.current_stream()
        hdr = N.PEER_HEADER_BYTES
        own_slot = self.slot(self.rank, parity, self.rank)
        overlay = {"records_ptr": own_slot + hdr, "count_ptr": self.count.data_ptr(), "capacity": self.capacity, "fmt": self.fmt,
                   "mirrors": [self.slot(r, parity, self.rank) + hdr for r in range(self.world) if r != self.rank],
                   "image_base": frame_lo * renderer.n_cams}
        headers = (ctypes.c_void_p * self.world)(*[self.slot(r, parity, self.rank) for r in range(self.world)])

        def render_and_publish():
            if int(w2c_dev.shape[0]) > 0:
                saved, renderer.geometry_ctas_per_sm = renderer.geometry_ctas_per_sm, (3 if render_stream is not None else renderer.geometry_ctas_per_sm)
                try:                                  # (three geometry CTAs per SM: the fill / expand of the step before runs beside them)
                    renderer.enqueue_overlay(res, w2c_dev, overlay, mode=mode)
                finally:
                    renderer.geometry_ctas_per_sm = saved
            else:
                self.count.zero_()
            N.check(N.lib().cama_peer_publish(rt.ctx, self.count.data_ptr(), self.step, headers, self.world, rt.stream()))

        if render_stream is not None:
            before = self._expanded.pop(self.step - 2, None)
            if before is not None:
                render_stream.wait_event(before)     # (not the expand of step s-1: that one runs beside this render)
            else:
                render_stream.wait_stream(main)      # first steps: everything enqueued so far
            with torch.cuda.stream(render_stream):
                render_and_publish()
                published = torch.cuda.Event()
                published.record(render_stream)
            N.check(N.lib().cama_frames_clear(rt.ctx, rt.ptr(out), out.numel(), rt.stream()))
            main.wait_event(published)
        else:
            N.check(N.lib().cama_frames_clear(rt.ctx, rt.ptr(out), out.numel(), rt.stream()))
            render_and_publish()
        slots = (ctypes.c_void_p * self.world)(*[self.slot(self.rank, parity, r) for r in range(self.world)])
        pal_dev = scratch = None
        if self.fmt == N.OVERLAY_PALETTE:
            pal_dev = self._palette_dev(res)
            scratch = rt.scratch("palette32", 1024)
        N.check(N.lib().cama_peer_expand(rt.ctx, slots, self.world, self.rank, self.step, self.capacity, self.fmt, rt.ptr(pal_dev), rt.ptr(scratch),
                                         rt.ptr(out), int(n_frames_total), renderer.n_cams, renderer.height, renderer.width, 0,
                                         self.status.data_ptr(), rt.stream()))
        if render_stream is not None:
            done = torch.cuda.Event()
            done.record(main)
            self._expanded[self.step] = done
            self._expanded.pop(self.step - 3, None)
        return out

    def reassemble(self, renderer, res, n_frames_total, out):
        """Zero-fill + expand of the slots of the LAST published step again (no render, no exchange): what the assembly
        alone costs — every frame byte of the clip written once, plus the lit chunks (bench.py's roofline figure)."""
        from . import _native as N
        rt = self.rt
        parity = self.step % PARITIES
        N.check(N.lib().cama_frames_clear(rt.ctx, rt.ptr(out), out.numel(), rt.stream()))
        slots = (ctypes.c_void_p * self.world)(*[self.slot(self.rank, parity, r) for r in range(self.world)])
        pal_dev = scratch = None
        if self.fmt == N.OVERLAY_PALETTE:
            pal_dev = self._palette_dev(res)
            scratch = rt.scratch("palette32", 1024)
        N.check(N.lib().cama_peer_expand(rt.ctx, slots, self.world, self.rank, self.step, self.capacity, self.fmt, rt.ptr(pal_dev), rt.ptr(scratch),
                                         rt.ptr(out), int(n_frames_total), renderer.n_cams, renderer.height, renderer.width, 0,
                                         self.status.data_ptr(), rt.stream()))
        return out

    def _palette_dev(self, res):
        cached = getattr(res, "_palette_dev", None)
        if cached is None:
            cached = res._palette_dev = self.rt.to_device(np.ascontiguousarray(res.palette, dtype=np.uint8))
        return cached

    def status_code(self):
        """0 ok, 1 a peer's step timed out, 2 a slot overflowed (synchronises)."""
        return int(self.status.item())


class ListExchange:
    """A frame-sharded clip assembled on every rank by exchanging the CENTRE RECORDS and rastering everything everywhere.

    Every rank ends a step with every frame of the clip in its HBM; those bytes have to be written by somebody on that
    GPU, and the raster is the kernel that writes them at ~0.75 of the HBM peak anyway.  So only the geometry is sharded:
    a rank runs ``cama_clip_render(phases=GEOMETRY)`` on its frame block (centre records into the per-band lists of its
    own list array); ``cama_peer_publish_lists`` copies the filled part of those lists to the same place of every peer's
    array with wide peer stores over NVLink (4 bytes per visible point: a third of the bytes of the lit-chunk records of
    PeerExchange), then the list lengths and the step number; after ``cama_peer_wait`` every rank runs ``phases=RASTER``
    over ALL frames.  No zero-fill, no expand, no NCCL, no host round trip.  Two parities of list arrays (one stream per rank:
    a rank reaches the geometry of step s+2 — the next writer of step s's arrays on its peers — only after it has seen
    their step s+1, which they published after rastering step s).
    """

    def __init__(self, rt, renderer, res, n_frames_total, capacity, lists_per_image, group=None):
        import torch
        import torch.distributed as dist
        from . import _native as N
        self.rt, self.group = rt, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > N.MAX_PEERS:
            raise ValueError(f"at most {N.MAX_PEERS} ranks")
        self.capacity = -(-int(capacity) // 4) * 4        # (16-byte lists: cama_peer_publish_lists copies them with wide stores)
        self.n_frames = int(n_frames_total)
        self.lists_per_frame = renderer.n_cams * int(lists_per_image)
        self.n_lists = self.n_frames * self.lists_per_frame
        hdr = N.PEER_HEADER_BYTES
        self.cursor_off = self.world * hdr
        self.records_off = -(-(self.cursor_off + self.n_lists * 4) // 256) * 256
        self.parity_bytes = -(-(self.records_off + self.n_lists * self.capacity * 4) // 256) * 256
        self.total_bytes = 2 * self.parity_bytes
        self.step = 0
        self.status = torch.zeros(1, dtype=torch.int32, device=rt.device)
        self.base = [None] * self.world
        self._own = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(N.PEER_HANDLE_BYTES)
        error = None
        try:
            N.check(N.lib().cama_peer_alloc(rt.ctx, self.total_bytes, ctypes.byref(self._own), handle))
        except N.CamaError as exc:
            error = str(exc)
        mine = {"handle": bytes(handle.raw), "device": int(rt.device.index), "error": error}
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        if all(e["error"] is None for e in everyone):
            for r, e in enumerate(everyone):
                if r == self.rank:
                    self.base[r] = self._own.value
                    continue
                ptr = ctypes.c_void_p()
                if N.lib().cama_peer_open(rt.ctx, e["device"], e["handle"], ctypes.byref(ptr)) != N.CAMA_OK:
                    error = N.lib().cama_last_error().decode("utf-8", "replace")
                    break
                self.base[r] = ptr.value
        else:
            error = error or "a peer could not allocate its list arrays"
        verdicts = [None] * self.world
        dist.all_gather_object(verdicts, error, group=group)
        self.error = next((v for v in verdicts if v), None)
        self.available = self.error is None
        if not self.available:
            self.close()

    close = PeerExchange.close
    __del__ = PeerExchange.__del__

    def render_and_assemble(self, renderer, res, w2c_dev, frame_lo, out, mode="binned", marks=None):
        """One step (asynchronous): geometry of this rank's frames [frame_lo, frame_lo + len(w2c_dev)) into everybody's
        lists, hand-off, raster of all frames into ``out`` (torch uint8 [n_frames_total, C, H, W, 3]).
        ``marks``: optional list that receives five timing events (start, geometry, published, arrived, rastered)."""
        import torch
        from . import _native as N
        rt = self.rt

        def mark():
            if marks is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record(torch.cuda.current_stream())
                marks.append(e)

        mark()
        self.step += 1
        at = (self.step & 1) * self.parity_bytes
        hdr = N.PEER_HEADER_BYTES
        n_local = int(w2c_dev.shape[0])
        own = self.base[self.rank] + at
        peers = [r for r in range(self.world) if r != self.rank]
        if n_local:
            renderer.enqueue_phase(res, w2c_dev, n_local,
                                   {"phases": N.PHASE_GEOMETRY, "records_ptr": own + self.records_off, "cursor_ptr": own + self.cursor_off,
                                    "frame_base": frame_lo, "frames": self.n_frames}, self.capacity, mode=mode)
        mark()
        peer_records = (ctypes.c_void_p * max(len(peers), 1))(*[self.base[r] + at + self.records_off for r in peers])
        peer_cursors = (ctypes.c_void_p * max(len(peers), 1))(*[self.base[r] + at + self.cursor_off for r in peers])
        headers = (ctypes.c_void_p * self.world)(*[self.base[r] + at + self.rank * hdr for r in range(self.world)])
        N.check(N.lib().cama_peer_publish_lists(rt.ctx, own + self.records_off, own + self.cursor_off, self.capacity,
                                                frame_lo * self.lists_per_frame, n_local * self.lists_per_frame,
                                                peer_records, peer_cursors, len(peers), self.step, headers, self.world, rt.stream()))
        mark()
        arrived = (ctypes.c_void_p * self.world)(*[own + r * hdr for r in range(self.world)])
        N.check(N.lib().cama_peer_wait(rt.ctx, arrived, self.world, self.step, 0, self.status.data_ptr(), rt.stream()))
        mark()
        self._last_raster = renderer.enqueue_phase(res, None, self.n_frames,
                                                   {"phases": N.PHASE_RASTER, "records_ptr": own + self.records_off, "cursor_ptr": own + self.cursor_off,
                                                    "frame_base": 0, "frames": self.n_frames}, self.capacity, out=out, mode=mode)
        mark()
        return out

    def reraster(self, renderer, res, out, mode="binned"):
        """The raster phase of the last step again (complete lists, no geometry, no exchange): what writing every frame of
        the clip on this rank costs — bench.py's roofline figure."""
        from . import _native as N
        own = self.base[self.rank] + (self.step & 1) * self.parity_bytes
        renderer.enqueue_phase(res, None, self.n_frames,
                               {"phases": N.PHASE_RASTER, "records_ptr": own + self.records_off, "cursor_ptr": own + self.cursor_off,
                                "frame_base": 0, "frames": self.n_frames}, self.capacity, out=out, mode=mode)
        return out

    def status_code(self):
        """0 ok, 1 a peer's step timed out, 2 a record list overflowed its capacity in the last step (synchronises)."""
        from . import _native as N
        code = int(self.status.item())
        last = getattr(self, "_last_raster", None)
        if code == 0 and last is not None:
            desc, ws = last
            stats = N.ClipStats()
            rc = N.lib().cama_clip_stats_read(self.rt.ctx, ctypes.byref(desc), self.rt.ptr(ws), self.rt.stream(), ctypes.byref(stats))
            if rc == N.CAMA_E_CAPACITY:
                code = 2
            else:
                N.check(rc)
            self.last_stats = {f: getattr(stats, f) for f, _ in N.ClipStats._fields_}
        return code


class SiteAssembler:
    """A frame-sharded clip assembled on every rank, step after step (BASELINE.json configs[3]).

    Construction is collective: sizes the record slots from one checked sparse render per rank (max over ranks,
    25 % headroom), builds the PeerExchange.  ``step()`` enqueues render + exchange + expand of this rank's frame
    block and returns the tensor that holds ALL frames of the clip afterwards (the same tensor every step)."""

    def __init__(self, reproject, dataset, group=None, mode="auto", overlap_clear=True, exchange="auto"):
        """exchange="lists": centre records exchanged, every rank rasters every frame (ListExchange);
        "chunks": lit-chunk records exchanged, zero-fill + expand on every rank (PeerExchange);
        "auto": lists up to 4 ranks, chunks beyond.  Measured on the config-3 site [B200], ms per assembled site, lists /
        chunks: 2 GPUs 1.006 / 1.175, 4 GPUs 0.957 / 0.986, 8 GPUs 1.092 / 1.000 — the list push sends 4 B per visible point
        to every peer after the geometry (0.33 ms for 7 peers at the ~300 GB/s it reaches), the chunk records leave the
        raster while it computes."""
        import torch
        import torch.distributed as dist
        self.rp, self.dataset, self.mode = reproject, dataset, mode
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        r, rt = reproject.renderer, reproject.rt
        self.res = reproject.resident(dataset)
        self.idx, w2c = reproject.frame_poses(dataset)
        self.n_frames = len(self.idx)
        self.lo, self.hi = frame_block(self.n_frames, self.rank, self.world)
        self.w2c_dev = torch.from_numpy(np.ascontiguousarray(w2c[self.lo:self.hi], dtype=np.float32).reshape(-1, 16)).to(rt.device)
        _, n, fmt = r.render_overlay(self.res, self.w2c_dev, mode=mode)          # settles the centre-record lists, counts the lit chunks
        stats = (r.last_stats or {}) if self.hi > self.lo else {}          # (a rank with an empty block has rendered nothing)
        most = torch.tensor([int(n), int(stats.get("record_capacity", 0)), int(stats.get("record_capacity_needed", 0)), int(stats.get("lists_per_image", 0))],
                            dtype=torch.int64, device=rt.device)
        dist.all_reduce(most, op=dist.ReduceOp.MAX, group=group)           # every rank decides from the same numbers
        if exchange == "auto":
            exchange = "lists" if self.world <= 4 else "chunks"
        self.kind = exchange
        self.frames = None
        self.render_stream = None
        if exchange == "lists" and _MODES_BINNED(mode) and int(most[3].item()) > 0:
            capacity = max(int(most[1].item()), int(int(most[2].item()) * 1.1) + 256)       # every rank: the same list capacity
            self.exchange = ListExchange(rt, r, self.res, self.n_frames, capacity, int(most[3].item()), group=group)
        else:
            self.kind = "chunks"
            self.exchange = PeerExchange(rt, int(int(most[0].item()) * 1.25) + 4096, fmt, group=group)
            self.render_stream = torch.cuda.Stream(device=rt.device, priority=-1) if overlap_clear else None
        self.available = self.exchange.available

    def step(self, out=None):
        import torch
        r = self.rp.renderer
        if out is None:
            if self.frames is None:
                self.frames = torch.empty((self.n_frames, r.n_cams, r.height, r.width, 3), dtype=torch.uint8, device=self.rp.rt.device)
            out = self.frames
        if self.kind == "lists":
            return self.exchange.render_and_assemble(r, self.res, self.w2c_dev, self.lo, out)
        return self.exchange.render_and_assemble(r, self.res, self.w2c_dev, self.lo, self.n_frames, out, mode=self.mode, render_stream=self.render_stream)
