"""Where does a single VLite.retrieve call (config 1) spend its time?"""
import cProfile, os, pstats, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from oracle import synth
from vlite_b200 import VLite
n, w = 100_000, 64
corpus = synth.corpus_rows(0, 0, n, w)
queries, _ = synth.clustered_queries(1, 0, n, 8, w)
emb = lambda texts: np.unpackbits(queries[[int(t) for t in texts]], axis=1).astype(np.float32) * 2 - 1
with tempfile.TemporaryDirectory() as td:
    v = VLite("lat", ctx_dir=td, embedder=emb, width=w)
    first = v._append_vectors(corpus)
    for i in range(n):
        v._register(f"id{i}_0", "", {}, first + i)
    for i in range(20):
        v.retrieve(str(i % 8), top_k=5)
    t0 = time.perf_counter()
    for i in range(200):
        v.retrieve(str(i % 8), top_k=5)
    print(f"retrieve: {(time.perf_counter()-t0)/200*1e6:.1f} us per call")
    q = queries[:1].copy()
    t0 = time.perf_counter()
    for i in range(200):
        v._corpus.search(q, 5, 1)
    print(f"Corpus.search (host buffers): {(time.perf_counter()-t0)/200*1e6:.1f} us per call")
    pr = cProfile.Profile(); pr.enable()
    for i in range(200):
        v.retrieve(str(i % 8), top_k=5)
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
    v.close()
