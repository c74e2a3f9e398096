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


@settings(max_examples=60, deadline=None)
@given(W=st.integers(1, 2000), H=st.integers(1, 1200), world=st.integers(1, 8))
def test_interleaved_tiles_are_owned_exactly_once(W, H, world):
    masks = [sharded.owned_tiles(W, H, world, r) for r in range(world)]
    total = sum(m.to(int) for m in masks)
    assert int(total.min()) == 1 and int(total.max()) == 1
    assert masks[0].numel() == ((W + 15) // 16) * ((H + 15) // 16)
