"""Step-by-step check of the tensor-core scan against the C oracle, smallest case first.
usage: mma_debug.py PATH   (2 = one CTA per tile, 3 = CTA pairs).  Prints mismatch diagnostics."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cscan, synth  # noqa: E402
from vlite_b200 import _capi  # noqa: E402

path = int(sys.argv[1])
ok = True
for (n, w, nq, k) in [(128, 128, 1, 1), (256, 128, 128, 4), (1000, 128, 5, 10), (20011, 128, 129, 10),
                      (20011, 64, 300, 10), (20011, 32, 100, 16), (500_000, 128, 1024, 10)]:
    corpus = synth.corpus_rows(5, 0, n, w)
    queries = synth.uniform_queries(6, nq, w)
    queries[0] = corpus[n - 1]
    with _capi.Corpus(w) as c:
        c.append(corpus)
        c.set_scan_path(path)
        s, r = c.search(queries, k, 0)
    es, er = cscan.search(queries, corpus, k, 0)
    good = np.array_equal(s, es) and np.array_equal(r, er)
    print(f"path {path} n={n} w={w} nq={nq} k={k}: {'OK' if good else 'MISMATCH'}", flush=True)
    if not good:
        ok = False
        bad = np.argwhere((s != es) | (r != er))
        print("  mismatching (query, slot) pairs:", len(bad), "first:", bad[:5].tolist())
        q = bad[0][0]
        print("  got   ", s[q][:8].tolist(), r[q][:8].tolist())
        print("  expect", es[q][:8].tolist(), er[q][:8].tolist())
        true = cscan.scores(queries[q], corpus, 0)
        rr = r[q][:8].astype(np.int64)
        rr = rr[rr < n]
        print("  true scores of returned rows", true[rr].tolist())
        break
sys.exit(0 if ok else 1)
