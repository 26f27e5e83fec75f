"""Smallest end-to-end case for compute-sanitizer: every kernel family once, tiny shapes."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from oracle import cscan, synth
from vlite_b200 import _capi
n, w = 3000, 128
corpus = synth.corpus_rows(0, 0, n, w)
with _capi.Corpus(w) as c:
    c.append(corpus[:1000]); c.fill_synthetic(0, 1000, 2000)
    c.tombstone([5, 6, 7])
    valid = np.ones(n, bool); valid[[5, 6, 7]] = False
    for nq, k in ((1, 5), (9, 10), (70, 3), (3, 300)):
        q = synth.uniform_queries(1, nq, w)
        s, r = c.search(q, k, 0)
        es, er = cscan.search(q, corpus, k, 0, synth.pack_mask(valid))
        assert np.array_equal(s, es) and np.array_equal(r, er)
    s, r = c.search(synth.uniform_queries(2, 4, w), 5, 1, synth.pack_mask(synth.filter_valid(2, 0, n, 0.3)))
    c.append_f32(np.full((4, w), -1.0, np.float32), _capi.SRC_F32_OFFSET)
    assert c.scores(corpus[0], 0).shape[0] == n + 4
print(_capi.quantize_binary(np.random.default_rng(0).standard_normal((3, 512)).astype(np.float32)).shape, "sanitize case ok")
