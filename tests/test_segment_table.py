"""CPU-only: the placement bookkeeping shared by ShardedCorpus and LocalShardedCorpus."""
import numpy as np
import pytest

from vlite_b200.sharded import SegmentTable


def test_least_full_placement_locate_and_merging():
    t = SegmentTable(3)
    assert t.dirty == {0, 1, 2}
    t.dirty.clear()
    placed = []
    for n in (5, 5, 5, 2, 9, 1, 1, 1):
        owner = t.least_full()
        placed.append((owner, t.add(owner, n)))
    assert [p[0] for p in placed] == [0, 1, 2, 0, 1, 2, 2, 0]          # least full, ties to the lowest shard
    assert [p[1] for p in placed] == [0, 5, 10, 15, 17, 26, 27, 28]    # global rows = insertion order
    # the two 1-row appends to shard 2 continue each other: ONE segment, one table entry, one upload
    assert t.segments[-2] == (26, 2, 5, 2) and t.segments[-1] == (28, 0, 7, 1) and len(t.segments) == 7
    assert t.table_of(2) == ([0, 5], [10, 26]) and t.dirty == {0, 1, 2}
    assert t.counts == [8, 14, 7] and t.total == 29
    for g in range(t.total):
        owner, local = t.locate(g)
        base, o, lstart, n = next(s for s in t.segments if s[0] <= g < s[0] + s[3])
        assert (owner, local) == (o, lstart + g - base)
    with pytest.raises(IndexError):
        t.locate(29)
    with pytest.raises(IndexError):
        t.locate(-1)
    assert list(t.overlapping(4, 8)) == [(0, 0, 4, 1), (1, 1, 0, 5), (6, 2, 0, 2)]


def test_single_shard_appends_stay_one_segment():
    t = SegmentTable(1)
    t.dirty.clear()
    for _ in range(1000):
        t.add(t.least_full(), 3)
    assert t.segments == [(0, 0, 0, 3000)] and t.table_of(0) == ([0], [0]) and t.dirty == {0}


def test_local_masks_follow_segments():
    rng = np.random.default_rng(0)
    t = SegmentTable(2)
    for n in (40, 70, 33, 5):
        t.add(t.least_full(), n)
    bits = rng.random(t.total) < 0.5
    words = np.packbits(np.pad(bits, (0, -t.total % 32)).reshape(-1, 32), axis=1, bitorder="little").view(np.uint32).reshape(-1)
    masks = t.local_masks(words)
    for owner in range(2):
        got = np.unpackbits(masks[owner].view(np.uint8), bitorder="little")[: t.counts[owner]].astype(bool)
        want = np.concatenate([bits[b:b + n] for b, o, ls, n in t.segments if o == owner])
        np.testing.assert_array_equal(got, want)
    only0 = t.local_masks(words, owners=[0])
    assert only0[1] is None and np.array_equal(only0[0], masks[0])
    t.clear()
    assert t.total == 0 and t.segments == [] and t.dirty == {0, 1}
