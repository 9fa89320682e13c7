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
