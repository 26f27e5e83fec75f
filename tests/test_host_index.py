"""Host-side bookkeeping of the VLite mirror that needs no GPU: the item-id prefix index must
select exactly what the reference's ``chunk_id.startswith(f"{id}_")`` scans select
(vlite/main.py:173, :195, :215, :235), in index order, through registrations, overwrites and
deletions."""
import threading

import numpy as np

from vlite_b200.main import VLite
from vlite_b200.metaindex import MetaIndex


class _NoDevice:
    rows = 0
    live_rows = 0
    width = 64

    def tombstone(self, rows):
        pass


def _bare():
    v = VLite.__new__(VLite)
    v._lock = threading.RLock()
    v._corpus, v._width = _NoDevice(), 64
    v._row_of, v._ids, v._texts, v._metas, v._extra = {}, [], {}, {}, {}
    v._meta_index, v._by_prefix = MetaIndex(), {}
    v.save = lambda *a, **k: None
    return v


def test_prefix_index_equals_startswith_scan():
    rng = np.random.default_rng(4)
    v = _bare()
    stems = ["a", "a_b", "a_b_c", "ab", "", "x_", "doc1", "doc10", "doc1_0"]
    row = 0

    def scan(id):
        return [c for c in v._row_of if c.startswith(f"{id}_")]

    for round_ in range(6):
        batch = []
        for _ in range(40):
            cid = f"{stems[int(rng.integers(len(stems)))]}_{int(rng.integers(12))}"
            batch.append(cid)
        if round_ % 2:
            v._register_many(batch, ["t"] * len(batch), [{} for _ in batch], row)     # may fall back on duplicates
        else:
            for i, cid in enumerate(batch):
                v._register(cid, "t", {}, row + i)
        row += len(batch)
        for probe in stems + ["a_b_c_1", "nope", "doc"]:
            assert v._chunks_of(probe) == scan(probe), probe
        victim = stems[int(rng.integers(len(stems)))]
        want = len(scan(victim))
        assert v.delete(victim) == want
        assert v._chunks_of(victim) == scan(victim) == []
        for probe in stems:
            assert v._chunks_of(probe) == scan(probe), probe
    assert set(v._by_prefix) == {c[:p] for c in v._row_of for p in range(len(c)) if c[p] == "_"}
