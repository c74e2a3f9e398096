"""Host-side partition logic of the Gaussian-sharded render (diff_gaussian_rasterization.sharded): property tests, CPU only."""
import os
import sys

from hypothesis import given, settings
from hypothesis import strategies as st

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "4dgs-slam_b200"))
from diff_gaussian_rasterization import sharded  # noqa: E402


@settings(max_examples=200, deadline=None)
@given(P=st.integers(0, 5_000_000), world=st.integers(1, 16))
def test_shard_bounds_partition_the_cloud(P, world):
    """Shards are contiguous, ordered, cover [0, P) exactly once and differ in size by at most one -- the order-preserving
    global id rank * Pmax + local of the sharded render relies on exactly that."""
    prev = 0
    sizes = []
    for r in range(world):
        lo, hi = sharded.shard_bounds(P, world, r)
        assert lo == prev and hi >= lo
        sizes.append(hi - lo)
        prev = hi
    assert prev == P
    assert max(sizes) - min(sizes) <= 1


@settings(max_examples=200, deadline=None)
@given(tiles_y=st.integers(1, 400), world=st.integers(1, 16))
def test_strip_bounds_partition_the_tile_rows(tiles_y, world):
    prev = 0
    for r in range(world):
        b, e = sharded.strip_bounds(tiles_y, world, r)
        assert b == prev and e >= b
        prev = e
    assert prev == tiles_y


@settings(max_examples=100, deadline=None)
@given(H=st.integers(1, 2400), world=st.integers(1, 16), cap=st.integers(1, 10_000_000), W=st.integers(1, 4000))
def test_strips_cover_every_pixel_row_once_and_payload_sizes_are_consistent(H, world, cap, W):
    """The pixel-row strips [16 b, min(H, 16 e)) derived from the tile-row strips tile the image, the padded strip height bounds
    every strip, and collective_bytes reports what the slabs / strips actually hold."""
    tiles_y, maxh = sharded._strip_rows(H, world)
    covered = 0
    for r in range(world):
        b, e = sharded.strip_bounds(tiles_y, world, r)
        y0, y1 = 16 * b, min(H, 16 * e)
        assert y0 == min(covered, 16 * b) or y1 <= y0
        if y1 > y0:
            assert y0 == covered and y1 - y0 <= maxh
            covered = y1
    assert covered == H
    cb = sharded.collective_bytes(W, H, world, cap)
    assert cb["all_to_all_records"] == world * (cap + 1) * 48 == cb["all_to_all_accumulators"]
    assert cb["all_gather_strip_ntouched_counts"] >= 5 * maxh * W * 4
