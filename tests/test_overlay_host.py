"""Host side of the sparse overlay transfer (cama_overlay_apply_host): byte movement only, CPU-testable."""
import ctypes

import numpy as np

from cama_b200 import _native as N

REC = np.dtype([("chunk", "<u4"), ("mask", "<u4"), ("bgr", "u1", (24,))])


def apply(records, frames, erase=0, threads=0):
    assert REC.itemsize == N.OVERLAY_RECORD_BYTES
    N.check(N.lib().cama_overlay_apply_host(records.ctypes.data, len(records), frames.ctypes.data, frames.size // 24, erase, threads))


def reference(records, frames, erase=0):
    flat = frames.reshape(-1, 8, 3)
    for r in records:  # noqa
        if r["chunk"] >= len(flat):
            continue
        for k in range(8):
            if (r["mask"] >> k) & 1:
                flat[r["chunk"], k] = 0 if erase else r["bgr"].reshape(8, 3)[k]


def test_apply_matches_numpy_and_erases():
    rng = np.random.default_rng(0)
    frames = rng.integers(0, 256, size=(2, 3, 16, 64, 3), dtype=np.uint8)
    n_chunks = frames.size // 24
    recs = np.zeros(9000, REC)                                     # > 4096: the threaded path
    frames = rng.integers(0, 256, size=(4, 6, 64, 128, 3), dtype=np.uint8)
    n_chunks = frames.size // 24
    recs["chunk"] = rng.permutation(n_chunks)[:9000]               # every lit chunk appears once, as the GPU emits them
    recs["mask"] = rng.integers(0, 256, len(recs))
    recs["mask"][:10] = 0xFF
    recs["bgr"] = rng.integers(0, 256, (len(recs), 24))
    extra = np.zeros(1, REC)
    extra["chunk"], extra["mask"], extra["bgr"] = n_chunks + 5, 0xFF, 7
    recs = np.concatenate([recs, extra])                           # out of range: ignored
    want = frames.copy()
    reference(recs, want)
    got = frames.copy()
    apply(recs, got)
    assert np.array_equal(got, want)
    single = frames.copy()
    apply(recs[:100], single, threads=1)
    want1 = frames.copy()
    reference(recs[:100], want1)
    assert np.array_equal(single, want1)
    # erase: painted pixels -> 0, everything else untouched
    reference(recs, want, erase=1)
    apply(recs, got, erase=1)
    assert np.array_equal(got, want)
    # whole-chunk variants: 24 bytes per record regardless of the mask
    blank = np.zeros_like(frames)
    apply(recs[:-1], blank, erase=N.OVERLAY_DRAW_CHUNKS)
    assert np.array_equal(blank.reshape(-1, 24)[recs["chunk"][:-1]], recs["bgr"][:-1])
    apply(recs[:-1], blank, erase=N.OVERLAY_BLANK_CHUNKS)
    assert not blank.any()
    assert N.lib().cama_overlay_apply_host(recs.ctypes.data, 5, got.ctypes.data, n_chunks, 7, 0) == N.CAMA_E_INVALID
    assert N.lib().cama_overlay_apply_host(None, 5, got.ctypes.data, n_chunks, 0, 0) == N.CAMA_E_INVALID
    N.check(N.lib().cama_overlay_apply_host(None, 0, None, 0, 0, 0))


def test_mosaic_layout_matches_concate_image():
    """Drawing into the 2x3 mosaic == drawing into [F,C,H,W,3] frames and np.concatenate (reference cama/tools.py:22-25)."""
    from cama_b200.tools import MOSAIC_ROWS, concate_image
    rng = np.random.default_rng(4)
    F, C, H, W = 3, 6, 16, 32
    names = ["camera_rear", "camera_front_left", "camera_front", "camera_front_right", "camera_rear_left", "camera_rear_right"]   # any camera_list order
    order = [n for row in MOSAIC_ROWS for n in row]
    tiles = np.array([order.index(n) for n in names], np.int32)
    n_chunks = F * C * H * W // 8
    recs = np.zeros(600, REC)
    recs["chunk"] = rng.permutation(n_chunks)[:600]
    recs["mask"] = rng.integers(1, 256, 600)
    recs["bgr"] = rng.integers(0, 256, (600, 24))
    frames = rng.integers(0, 256, size=(F, C, H, W, 3), dtype=np.uint8)
    mosaic = np.stack([concate_image({n: frames[f, c] for c, n in enumerate(names)}) for f in range(F)])
    assert mosaic.shape == (F, 2 * H, 3 * W, 3)
    for op in (N.OVERLAY_DRAW, N.OVERLAY_BLANK, N.OVERLAY_DRAW_CHUNKS, N.OVERLAY_BLANK_CHUNKS):
        apply(recs, frames, erase=op)
        N.check(N.lib().cama_overlay_apply_host_mosaic(recs.ctypes.data, len(recs), mosaic.ctypes.data, F, C, H, W, 3, tiles.ctypes.data, op, 0))
        want = np.stack([concate_image({n: frames[f, c] for c, n in enumerate(names)}) for f in range(F)])
        assert np.array_equal(mosaic, want), op
    bad = tiles.copy(); bad[0] = 9
    assert N.lib().cama_overlay_apply_host_mosaic(recs.ctypes.data, len(recs), mosaic.ctypes.data, F, C, H, W, 3, bad.ctypes.data, 0, 0) == N.CAMA_E_INVALID
