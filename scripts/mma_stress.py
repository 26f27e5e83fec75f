"""Repeat-launch stress of the tensor-core scan at 8+ tiles per CTA on every path (2 / 3: HAMMING as e2m1,
5 / 6: as int8) against the oracle: counts launches whose top-k differs.  This shape exposed the packed-box
hand-off race described in mmascan.cuh (expanders)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from oracle import cscan, synth
from vlite_b200 import _capi
HAMMING = 0
n, w = 150_003, 128
corpus = synth.corpus_rows(11, 0, n, w)
reps = int(os.environ.get("REPS", "20"))
for nq, path, k, metric in ((128, 2, 10, 0), (1024, 2, 10, 0), (1024, 3, 10, 0), (200, 3, 16, 0), (128, 5, 10, 0), (1024, 6, 10, 0),
                            (128, 2, 10, 1), (1024, 3, 10, 1), (1024, 6, 10, 1)):
    queries, src = synth.clustered_queries(12, 0, n, nq, w)
    es, er = cscan.search(queries, corpus, k, metric, None)
    with _capi.Corpus(w) as c:
        c.append(corpus)
        c.set_scan_path(path)
        bad = 0; tot = 0
        for rep in range(reps):
            s, r = c.search(queries, k, metric, None)
            b = int(((s != es) | (r != er)).sum()); bad += b > 0; tot += b
        print(f"nq={nq} path={path} k={k} metric={metric}: {bad}/{reps} launches wrong ({tot} entries)", flush=True)
