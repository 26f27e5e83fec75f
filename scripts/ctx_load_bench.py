"""Time vl_corpus_load_file (CTX EMBEDDINGS payload -> HBM) on a float32 CTX file of N rows.
usage: ctx_load_bench.py [rows] [path]   (env VLITE_B200_IO_THREADS picks the reader thread count)"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import ctxfile as octx, synth  # noqa: E402  (data generation + check only)
from vlite_b200 import _capi  # noqa: E402
from vlite_b200.ctx import CtxFile  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
path = sys.argv[2] if len(sys.argv) > 2 else "/tmp/ctx_load_bench.ctx"
rows64 = synth.corpus_rows(9, 0, n, 64)
if not os.path.exists(path):
    octx.write_ctx(path, {"embedding_model": "m", "embedding_size": 64, "embedding_dtype": "float32",
                          "context_length": 512}, rows64, [], {})
cf = CtxFile(path)
cf.load(materialise=False)
off, ln = cf.embeddings_section
best = None
for rep in range(4):
    with _capi.Corpus(64, capacity_rows=n) as c:
        t0 = time.perf_counter()
        first, got = c.load_file(path, off, ln, _capi.SRC_F32_OFFSET)
        dt = time.perf_counter() - t0
        assert got == n
        if rep == 0:
            assert np.array_equal(c.read(0, 4096), rows64[:4096]) and np.array_equal(c.read(n - 4096, 4096), rows64[-4096:])
    best = dt if best is None else min(best, dt)
print(json.dumps({"rows": n, "file_GB": ln / 1e9, "threads": os.environ.get("VLITE_B200_IO_THREADS", "default"),
                  "best_s": best, "file_GBps": ln / 1e9 / best, "rows_per_s": n / best}))
