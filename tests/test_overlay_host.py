"""Host side of the sparse overlay transfer (cama_overlay_apply_host): byte movement only, CPU-testable."""
import ctypes

import numpy as np
import pytest

from cama_b200 import _native as N
from cama_b200.tools import MOSAIC_ROWS, concate_image

REC = np.dtype([("chunk", "<u4"), ("mask", "<u4"), ("bgr", "u1", (24,))])
REC_PAL = np.dtype([("chunk", "<u4"), ("index", "u1", (8,))])


def apply(records, frames, op=0, threads=0, fmt=N.OVERLAY_BGR, palette=None, tiles=None, shape=None):
    assert REC.itemsize == N.OVERLAY_RECORD_BYTES[N.OVERLAY_BGR] and REC_PAL.itemsize == N.OVERLAY_RECORD_BYTES[N.OVERLAY_PALETTE]
    F, C, H, W = shape if shape is not None else frames.shape[:4]
    target = N.OverlayTarget(frames.ctypes.data, F, C, H, W, 0 if tiles is None else 3, None if tiles is None else tiles.ctypes.data)
    return N.lib().cama_overlay_apply_host(records.ctypes.data if records is not None else None, len(records) if records is not None else 5,
                                           fmt, None if palette is None else palette.ctypes.data, ctypes.byref(target), op, threads)


def reference(records, frames, op=0, palette=None):
    flat = frames.reshape(-1, 8, 3)
    for r in records:
        if r["chunk"] >= len(flat):
            continue
        if palette is None:
            mask, bgr = int(r["mask"]), r["bgr"].reshape(8, 3)
        else:
            mask = sum(1 << k for k in range(8) if r["index"][k])
            bgr = np.where(r["index"][:, None] > 0, palette[r["index"]], 0)
        for k in range(8):
            whole = op in (N.OVERLAY_DRAW_CHUNKS, N.OVERLAY_BLANK_CHUNKS)
            if whole or (mask >> k) & 1:
                flat[r["chunk"], k] = 0 if op in (N.OVERLAY_BLANK, N.OVERLAY_BLANK_CHUNKS) else bgr[k]


def make_records(rng, n_chunks, n, fmt):
    recs = np.zeros(n, REC if fmt == N.OVERLAY_BGR else REC_PAL)
    recs["chunk"] = rng.permutation(n_chunks)[:n]                  # every lit chunk appears once, as the GPU emits them
    if fmt == N.OVERLAY_BGR:
        recs["mask"] = rng.integers(0, 256, n)
        recs["mask"][:10] = 0xFF
        recs["bgr"] = rng.integers(0, 256, (n, 24))
        recs["bgr"] *= ((recs["mask"][:, None] >> (np.arange(24) // 3)) & 1).astype(np.uint8)     # unpainted pixels carry zeros
    else:
        recs["index"] = rng.integers(0, 4, (n, 8)) * (rng.random((n, 8)) < 0.6)
        recs["index"][:10] = 2
    return recs


@pytest.mark.parametrize("fmt", [N.OVERLAY_BGR, N.OVERLAY_PALETTE])
def test_apply_matches_numpy(fmt):
    rng = np.random.default_rng(0)
    frames = rng.integers(0, 256, size=(4, 6, 64, 128, 3), dtype=np.uint8)
    n_chunks = frames.size // 24
    palette = np.zeros((256, 3), np.uint8)
    palette[1:4] = [[211, 211, 211], [0, 215, 255], [9, 8, 7]]
    pal = palette if fmt == N.OVERLAY_PALETTE else None
    recs = make_records(rng, n_chunks, 9000, fmt)                   # > 4096: the threaded path
    extra = np.zeros(1, recs.dtype)
    extra["chunk"] = n_chunks + 5
    recs = np.concatenate([recs, extra])                            # out of range: ignored
    for op in (N.OVERLAY_DRAW, N.OVERLAY_BLANK, N.OVERLAY_DRAW_CHUNKS, N.OVERLAY_BLANK_CHUNKS):
        want, got = frames.copy(), frames.copy()
        reference(recs, want, op, pal)
        N.check(apply(recs, got, op, fmt=fmt, palette=pal))
        assert np.array_equal(got, want), op
        single = frames.copy()
        N.check(apply(recs[:100], single, op, threads=1, fmt=fmt, palette=pal))
        want1 = frames.copy()
        reference(recs[:100], want1, op, pal)
        assert np.array_equal(single, want1), op
    assert apply(recs, frames, 7, fmt=fmt, palette=pal) == N.CAMA_E_INVALID
    assert apply(None, frames, 0, fmt=fmt, palette=pal) == N.CAMA_E_INVALID
    if fmt == N.OVERLAY_PALETTE:
        assert apply(recs, frames, 0, fmt=fmt, palette=None) == N.CAMA_E_INVALID
    N.check(apply(recs[:0], frames, 0, fmt=fmt, palette=pal))


def test_mosaic_layout_matches_concate_image():
    """Drawing into the 2x3 mosaic == drawing into [F,C,H,W,3] frames and np.concatenate (reference cama/tools.py:22-25)."""
    rng = np.random.default_rng(4)
    F, C, H, W = 3, 6, 16, 32
    names = ["camera_rear", "camera_front_left", "camera_front", "camera_front_right", "camera_rear_left", "camera_rear_right"]   # any camera_list order
    order = [n for row in MOSAIC_ROWS for n in row]
    tiles = np.array([order.index(n) for n in names], np.int32)
    recs = make_records(rng, F * C * H * W // 8, 600, N.OVERLAY_BGR)
    frames = rng.integers(0, 256, size=(F, C, H, W, 3), dtype=np.uint8)
    mosaic = np.stack([concate_image({n: frames[f, c] for c, n in enumerate(names)}) for f in range(F)])
    assert mosaic.shape == (F, 2 * H, 3 * W, 3)
    for op in (N.OVERLAY_DRAW, N.OVERLAY_BLANK, N.OVERLAY_DRAW_CHUNKS, N.OVERLAY_BLANK_CHUNKS):
        N.check(apply(recs, frames, op))
        N.check(apply(recs, mosaic, op, tiles=tiles, shape=(F, C, H, W)))
        want = np.stack([concate_image({n: frames[f, c] for c, n in enumerate(names)}) for f in range(F)])
        assert np.array_equal(mosaic, want), op
    bad = tiles.copy()
    bad[0] = 9
    assert apply(recs, mosaic, 0, tiles=bad, shape=(F, C, H, W)) == N.CAMA_E_INVALID


@pytest.mark.parametrize("n_colours", [1, 2, 15, 16, 40, 255])
def test_palette_draw_chunks_both_paths(n_colours):
    """DRAW_CHUNKS with palette records: up to 15 colours take the byte-shuffle path (SSSE3), more take the
    two-pixels-per-lookup table; both must write exactly palette[index] for all 8 pixels of every record, leave the
    other chunks alone, ignore chunks outside the target and entries above the palette size in use."""
    rng = np.random.default_rng(n_colours)
    frames = rng.integers(0, 256, size=(2, 3, 64, 256, 3), dtype=np.uint8)
    n_chunks = frames.size // 24
    palette = np.zeros((256, 3), np.uint8)
    palette[1:n_colours + 1] = rng.integers(1, 256, (n_colours, 3))
    recs = np.zeros(6000, REC_PAL)
    recs["chunk"] = rng.permutation(n_chunks)[:6000]
    recs["chunk"][-3:] = [n_chunks, n_chunks + 7, 0xFFFFFFFF]           # outside the target: ignored
    recs["index"] = rng.integers(0, n_colours + 1, (6000, 8))
    want = frames.copy()
    flat = want.reshape(-1, 8, 3)
    inside = recs["chunk"] < n_chunks
    flat[recs["chunk"][inside]] = palette[recs["index"][inside]]
    masked = frames.copy()                                              # DRAW: only painted pixels (index != 0) are written
    mflat = masked.reshape(-1, 8, 3)
    painted = recs["index"][inside] > 0
    mflat[recs["chunk"][inside]] = np.where(painted[:, :, None], palette[recs["index"][inside]], mflat[recs["chunk"][inside]])
    for threads in (1, 0):
        got = frames.copy()
        assert apply(recs, got, op=N.OVERLAY_DRAW_CHUNKS, threads=threads, fmt=N.OVERLAY_PALETTE, palette=palette) == 0
        assert np.array_equal(got, want), (n_colours, threads)
        got = frames.copy()
        assert apply(recs, got, op=N.OVERLAY_DRAW, threads=threads, fmt=N.OVERLAY_PALETTE, palette=palette) == 0
        assert np.array_equal(got, masked), (n_colours, threads)



def test_worker_pool_concurrent_callers_and_thread_counts():
    """The host loops run on the library's own worker pool (csrc/host_pool.h): callers on different Python threads
    (Reproject's helper thread blanks while the main thread may draw) are serialised region by region, the pool grows
    with the thread count asked for, and the result never depends on either."""
    import threading
    rng = np.random.default_rng(11)
    shape = (2, 3, 64, 256, 3)
    n_chunks = int(np.prod(shape)) // 24
    palette = np.zeros((256, 3), np.uint8)
    palette[1:4] = [[211, 211, 211], [0, 215, 255], [9, 8, 7]]
    recs = make_records(rng, n_chunks, 12000, N.OVERLAY_PALETTE)
    halves = [np.ascontiguousarray(recs[:6000]), np.ascontiguousarray(recs[6000:])]      # disjoint chunks
    want = np.zeros(shape, np.uint8)
    reference(recs, want, op=N.OVERLAY_DRAW_CHUNKS, palette=palette)
    for threads in (2, 3, 8, 0, 5):
        frames = np.zeros(shape, np.uint8)
        errors = []

        def work(part):
            for _ in range(20):
                if apply(part, frames, op=N.OVERLAY_DRAW_CHUNKS, threads=threads, fmt=N.OVERLAY_PALETTE, palette=palette) != 0:
                    errors.append(1)

        workers = [threading.Thread(target=work, args=(h,)) for h in halves]
        for w in workers:
            w.start()
        for w in workers:
            w.join(60)
        assert not errors and not any(w.is_alive() for w in workers)
        assert np.array_equal(frames, want), threads
        assert apply(recs, frames, op=N.OVERLAY_BLANK_CHUNKS, threads=threads, fmt=N.OVERLAY_PALETTE, palette=palette) == 0
        assert not frames.any()
