"""-m gpu: rescore kernel vs the oracle (float64 dot) and vs the outputs of the reference's cos_sim
(vlite/utils.py:205-208, the only dense similarity it holds: tests/golden/rescore_golden.npz), and the
single-GPU emulation of the sharded merge."""
import numpy as np
import pytest

from oracle import cscan, search as osearch, synth

pytestmark = pytest.mark.gpu
RTOL = 1e-5          # north_star: rescored float scores within 1e-5 relative


def _docs(gpu, torch, n, dims, seed=7):
    d = torch.empty((n, dims), dtype=torch.int8, device="cuda")
    gpu.synth_docs_i8(seed, 0, n, dims, d)
    torch.cuda.synchronize()
    host = synth.doc_rows_i8(seed, 0, n, dims)
    np.testing.assert_array_equal(d.cpu().numpy(), host)                # device generator == numpy generator
    return d, host


@pytest.mark.parametrize("qtype", ["f32", "i8"])
@pytest.mark.parametrize("n_cand,k", [(100, 10), (1000, 10), (37, 37), (2048, 50)])
def test_rescore_matches_oracle(gpu, qtype, n_cand, k):
    torch = pytest.importorskip("torch")
    n, dims, nq = 20000, 1024, 9
    docs, host_docs = _docs(gpu, torch, n, dims)
    rng = np.random.default_rng(5)
    if qtype == "f32":
        queries = rng.standard_normal((nq, dims)).astype(np.float32)
    else:
        queries = rng.integers(-128, 128, size=(nq, dims), dtype=np.int8)
    cand = np.stack([rng.choice(n, size=n_cand, replace=False) for _ in range(nq)]).astype(np.uint64)
    cand[0, -3:] = osearch.SENTINEL_ROW                                   # padded candidate lists
    with gpu.Corpus(128) as c:
        s, r = c.rescore(docs, n, dims, queries, cand, k)
    es, er = osearch.rescore(queries, host_docs, cand, k)
    if qtype == "i8":
        np.testing.assert_array_equal(r, er)                             # int32 accumulate: exact
        np.testing.assert_array_equal(s.astype(np.float64), es)
    else:
        # int8 x fp32 products and their fp64 sum are exact on the device: the only rounding is the
        # final one to float32.  north_star: within 1e-5 RELATIVE -- asserted per element, no absolute
        # slack, cancelling sums included (k = n_cand = 37 keeps every candidate, near-zero scores too)
        ok = er != osearch.SENTINEL_ROW
        rel = np.abs(s.astype(np.float64)[ok] - es[ok]) / np.maximum(np.abs(es[ok]), 1e-300)
        print(f"rescore fp32 queries: max relative error {rel.max():.3e} over {ok.sum()} scores")
        assert rel.max() <= RTOL, rel.max()
        assert rel.max() <= 1.2e-7                                       # = float32 rounding of an exact sum
        # ordering is decided on float32 scores: rows may differ from the float64 order only where two
        # float64 scores round to the same float32
        for qi, pos in zip(*np.nonzero(r != er)):
            assert np.float32(es[qi, pos]) == s[qi, pos]


def test_binary_then_rescore_pipeline(gpu):
    """config 3 in miniature: binary top-C on sign bits of the int8 docs, then int8 rescore to k."""
    torch = pytest.importorskip("torch")
    n, dims, nq, C, k = 30000, 1024, 6, 200, 10
    docs, host_docs = _docs(gpu, torch, n, dims, seed=11)
    ubin = osearch.quantize_binary(host_docs.astype(np.float32))
    qf = host_docs[[5, 77, 4000, 12345, 29999, 0]].astype(np.float32) + np.random.default_rng(1).standard_normal((nq, dims)).astype(np.float32) * 20
    qb = osearch.quantize_binary(qf)
    with gpu.Corpus(128) as c:
        # sign-quantise the int8 docs on the device and append (no host round trip of the corpus)
        packed = gpu.quantize_binary(docs.float().contiguous())
        np.testing.assert_array_equal(packed, ubin)
        c.append(packed)
        bs, br = c.search(qb, C, gpu.HAMMING)
        es, er = cscan.search(qb, ubin, C, gpu.HAMMING)
        np.testing.assert_array_equal(br, er)
        s, r = c.rescore(docs, n, dims, qf, br, k)
    os_, or_ = osearch.rescore(qf, host_docs, er, k)
    np.testing.assert_array_equal(r, or_)
    np.testing.assert_allclose(s, os_, rtol=RTOL)
    np.testing.assert_array_equal(r[:, 0], [5, 77, 4000, 12345, 29999, 0])


@pytest.mark.parametrize("shards", [2, 4, 8])
def test_sharded_merge_single_gpu_emulation(gpu, shards):
    """All shards on one GPU (the profiling guide's advice for fewer GPUs than ranks): per-shard
    top-k into the all-gather layout [scores | rows] per rank, then the strided CUDA merge."""
    torch = pytest.importorskip("torch")
    from vlite_b200.sharded import ShardPlan
    n, w, nq, k = 40001, 128, 33, 10
    nq_even = nq + (nq * k) % 2
    full = synth.corpus_rows(9, 0, n, w, period=700)
    queries = synth.uniform_queries(10, nq_even, w)
    plan = ShardPlan(n, shards)
    nb = nq_even * k
    block = torch.empty((shards, nb * 12), dtype=torch.uint8, device="cuda")
    dq = torch.from_numpy(queries).cuda()
    handles = []
    for g in range(shards):
        c = gpu.Corpus(w, base_row=plan.base_row(g))
        c.fill_synthetic(9, 0, plan.shard_rows(g), 700)
        ls = block[g, : nb * 4].view(torch.int32).view(nq_even, k)
        lr = block[g, nb * 4:].view(torch.int64).view(nq_even, k)
        c.search(dq, k, gpu.HAMMING, None, ls, lr)
        handles.append(c)
    out_s = torch.empty((nq_even, k), dtype=torch.int32, device="cuda")
    out_r = torch.empty((nq_even, k), dtype=torch.int64, device="cuda")
    base = block.data_ptr()
    gpu.merge_topk(base, base + nb * 4, shards, nq_even, k, k, out_s, out_r,
                   score_list_stride=nb * 3, row_list_stride=nb * 12 // 8)
    torch.cuda.synchronize()
    for c in handles:
        c.close()
    es, er = cscan.search(queries, full, k, gpu.HAMMING)
    np.testing.assert_array_equal(out_s.cpu().numpy().view(np.uint32), es)
    np.testing.assert_array_equal(out_r.cpu().numpy().view(np.uint64), er)


@pytest.mark.parametrize("qname", ["f32", "i8"])
def test_rescore_matches_reference_cos_sim_golden(gpu, qname):
    """rescore_golden.npz holds what the reference's cos_sim (vlite/utils.py:205-208) returned for seeded
    queries x int8 documents.  The device scores divided by the same norms must reproduce them within
    1e-5 relative (north_star), with identical documents ranked by row."""
    torch = pytest.importorskip("torch")
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "rescore_golden.npz"))
    docs = g["docs_i8"]
    n, dims = docs.shape
    qs = g[f"queries_{qname}"]
    d = torch.from_numpy(docs).cuda()
    cand = np.tile(np.arange(n, dtype=np.uint64), (len(qs), 1))
    with gpu.Corpus(128) as c:
        s, r = c.rescore(d, n, dims, qs, cand, n)
    doc_norm = np.linalg.norm(docs.astype(np.float64), axis=1)
    for qi, q in enumerate(qs):
        rows = r[qi].astype(np.int64)
        assert sorted(rows.tolist()) == list(range(n))
        cos = s[qi].astype(np.float64) / (np.linalg.norm(q.astype(np.float64)) * doc_norm[rows])
        ref = g[f"{qname}_q{qi}_cos_f64"][rows]
        rel = np.abs(cos - ref) / np.maximum(np.abs(ref), 1e-300)
        big = np.abs(ref) > 1e-6 * np.abs(ref).max()          # (a cosine that cancels to ~0 has no relative scale)
        assert rel[big].max() <= RTOL, rel[big].max()
        assert np.abs(cos - ref).max() <= 1e-7
        pos5, pos17 = int(np.nonzero(rows == 5)[0][0]), int(np.nonzero(rows == 17)[0][0])
        assert pos17 == pos5 + 1
