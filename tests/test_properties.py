"""Property tests (hypothesis) of the host-side logic: the pieces every device call leans on.

The reference has no tests; these pin the host helpers of this package against the oracle's
restatement of the reference on RANDOM inputs — densify (float32 NEP-50 arithmetic, end points dropped, short segments
dropped), instance packing, culling bounds, the frame partition of a sharded clip and the mailbox layout.  CPU-only.
"""
import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from cama_b200 import shard
from cama_b200.batched import pack_vertices, tile_bounds
from cama_b200.reproject import densify_polyline, pack_instances, unpack_instances
from oracle import cama_oracle as orc
from oracle import oracle_c

coord = st.floats(min_value=-300.0, max_value=300.0, allow_nan=False, allow_infinity=False, width=32)
polyline = st.lists(st.tuples(coord, coord), min_size=2, max_size=12)


@settings(max_examples=120, deadline=None)
@given(polyline)
def test_densify_matches_the_reference_restatement(points):
    """cama_b200.reproject.densify_polyline (vectorised) == the oracle's scalar loop (reference :49-63) == oracle.c, bit for bit."""
    pts = np.array(points, dtype=np.float64)
    seg = np.linalg.norm(np.diff(pts.astype(np.float32), axis=0), axis=-1)
    if int((seg / orc.RESOLUTION).astype(np.int64).sum()) == 0:
        with pytest.raises(IndexError):                    # the reference indexes an empty array there
            densify_polyline(points, orc.RESOLUTION)
        return
    got = densify_polyline(points, orc.RESOLUTION)
    want = orc._densify(points)
    assert got.dtype == np.float32 and np.array_equal(got, want)
    assert np.array_equal(oracle_c.densify(pts), want)
    # the last vertex of the polyline is never emitted unless another segment starts there
    assert len(got) == int((seg / orc.RESOLUTION).astype(np.int64).sum())


@settings(max_examples=60, deadline=None)
@given(st.lists(st.integers(min_value=0, max_value=40), min_size=0, max_size=12), st.integers(min_value=0, max_value=2**31 - 1))
def test_pack_unpack_round_trip(counts, seed):
    rng = np.random.default_rng(seed)
    instances = [{"class": f"c{i % 3}", "points": rng.standard_normal((n, 3)).astype(np.float32)} for i, n in enumerate(counts)]
    flat, offsets, classes = pack_instances(instances)
    assert offsets[0] == 0 and offsets[-1] == sum(counts) and len(offsets) == len(counts) + 1
    back = unpack_instances(flat, offsets, classes, drop_empty=False)
    assert len(back) == len(instances)
    for a, b in zip(instances, back):
        assert a["class"] == b["class"] and np.array_equal(a["points"], b["points"])
    kept = unpack_instances(flat, offsets, classes)            # the reference drops instances left empty
    assert [len(i["points"]) for i in kept] == [n for n in counts if n > 0]
    if instances:
        layout, verts, ordinal, bgr = pack_vertices(instances)
        assert verts.shape == (sum(counts), 4) and bgr.shape == (len(counts), 3)
        assert np.array_equal(verts[:, 3].view(np.int32), np.repeat(np.arange(len(counts), dtype=np.int32), counts))


@settings(max_examples=60, deadline=None)
@given(st.integers(min_value=1, max_value=700), st.sampled_from([32, 256]), st.integers(min_value=0, max_value=2**31 - 1))
def test_culling_bounds_contain_their_vertices(n, tile, seed):
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-500, 500, size=(n, 3))
    b = tile_bounds(xyz, tile)
    assert b.shape == ((n + tile - 1) // tile, 6)
    for t in range(b.shape[0]):
        pts = xyz[t * tile:(t + 1) * tile]
        assert np.all(np.abs(pts - b[t, :3]) <= b[t, 3:] + 1e-9)


@given(st.integers(min_value=0, max_value=5000), st.integers(min_value=1, max_value=8))
def test_frame_blocks_partition_the_clip(n_frames, world):
    blocks = [shard.frame_block(n_frames, r, world) for r in range(world)]
    covered = [f for lo, hi in blocks for f in range(lo, hi)]
    assert covered == list(range(n_frames))                     # contiguous, ordered, nothing twice, nothing missing
    assert all(hi - lo <= shard.block_size(n_frames, world) for lo, hi in blocks)


@given(st.integers(min_value=1, max_value=8), st.integers(min_value=0, max_value=10**7), st.sampled_from([12, 32]))
def test_mailbox_slots_are_disjoint_and_aligned(world, capacity, record_bytes):
    slot, total, offset = shard.slot_layout(world, capacity, record_bytes)
    assert slot % 256 == 0 and slot >= 256 + capacity * record_bytes
    offs = sorted(offset(p, s) for p in range(shard.PARITIES) for s in range(world))
    assert offs[0] == 0 and all(b - a == slot for a, b in zip(offs, offs[1:])) and offs[-1] + slot == total
