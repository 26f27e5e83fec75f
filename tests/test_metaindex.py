"""Host-side metadata bitmap index against the reference's row predicate
``all(md.get(k) == v for k, v in metadata.items())`` (vlite/main.py:162)."""
import numpy as np

from vlite_b200.metaindex import MetaIndex, indexable


def brute(metas, query, n):
    bits = np.zeros((n + 31) // 32 * 32, dtype=bool)
    for r, md in enumerate(metas):
        if md is not None and all(md.get(k) == v for k, v in query.items()):
            bits[r] = True
    return np.packbits(bits.reshape(-1, 32), axis=1, bitorder="little").view(np.uint32).reshape(-1)


def test_random_updates_match_row_predicate():
    rng = np.random.default_rng(5)
    idx, metas = MetaIndex(), []
    langs, years = ["en", "fr", "de", None], [2020, 2021, 2022.0, True]
    for r in range(1500):
        md = {}
        if rng.random() < 0.9:
            md["lang"] = langs[rng.integers(len(langs))]
        if rng.random() < 0.7:
            md["year"] = years[rng.integers(len(years))]
        if rng.random() < 0.1:
            md["tags"] = ["a", "b"]                       # not indexable
        metas.append(md)
        idx.set_row(md, r)
    for _ in range(300):                                   # in-place updates, as VLite.update does
        r = int(rng.integers(len(metas)))
        new = {"lang": langs[rng.integers(len(langs))]}
        idx.clear_row({k: metas[r][k] for k in new if k in metas[r]}, r)
        idx.set_row(new, r)
        metas[r].update(new)
    for n in (1500, 1499, 1473, 64, 1):
        for q in ({"lang": "en"}, {"lang": "fr", "year": 2021}, {"year": 1}, {"year": 2022},
                  {"lang": None}, {"lang": "xx"}, {"missing": 3}, {}):
            got = idx.mask(q, n)
            assert got.dtype == np.uint32 and got.shape == ((n + 31) // 32,)
            np.testing.assert_array_equal(got, brute(metas[:n], q, n), err_msg=f"{q} n={n}")
    assert idx.mask({"tags": ["a", "b"]}, 1500) is None      # caller falls back to a row walk
    assert idx.mask({"x": float("nan")}, 1500) is None
    assert not indexable([1]) and not indexable(float("nan")) and indexable(("a", 1))


def test_ranges_rows_and_cache():
    idx = MetaIndex()
    idx.set_range({"src": "bulk"}, 5, 200)                 # spans several words, ragged both ends
    idx.set_range({"src": "bulk"}, 300, 3)                 # inside one word
    idx.set_rows({"src": "picked"}, [0, 63, 64, 1000])
    want = np.r_[np.arange(5, 205), np.arange(300, 303)]
    np.testing.assert_array_equal(idx.rows({"src": "bulk"}, 2000), want)
    np.testing.assert_array_equal(idx.rows({"src": "bulk"}, 100), np.arange(5, 100))
    np.testing.assert_array_equal(idx.rows({"src": "picked"}, 1001), [0, 63, 64, 1000])
    np.testing.assert_array_equal(idx.rows({"src": "picked"}, 1000), [0, 63, 64])
    a = idx.mask({"src": "bulk"}, 2000)
    assert idx.mask({"src": "bulk"}, 2000) is a             # cached until the next mutation
    idx.set_row({"src": "bulk"}, 1999)
    b = idx.mask({"src": "bulk"}, 2000)
    assert b is not a and b[1999 // 32] >> (1999 % 32) & 1
    idx.clear()
    assert len(idx) == 0 and not idx.mask({"src": "bulk"}, 2000).any()


def test_bulk_registration_matches_row_by_row():
    rng = np.random.default_rng(9)
    metas = []
    for i in range(3000):
        md = {"lang": ["en", "fr", None][int(rng.integers(3))]} if rng.random() < 0.8 else {}
        if rng.random() < 0.5:
            md["year"] = int(rng.integers(2020, 2023))
        if i % 97 == 0:
            md["tags"] = [i]
        metas.append(md)
    one, bulk = MetaIndex(), MetaIndex()
    for r, md in enumerate(metas):
        one.set_row(md, r + 11)
    bulk.set_many(metas[:1000], 11)
    bulk.set_many(metas[1000:], 1011)
    shared = {"src": "same"}
    one.set_range(shared, 4000, 500)
    bulk.set_many([shared] * 500, 4000)                    # one dict shared by every row: range fill
    for q in ({"lang": "en"}, {"lang": None}, {"year": 2021, "lang": "fr"}, {"src": "same"}, {"nope": 1}, {}):
        np.testing.assert_array_equal(bulk.mask(q, 4600), one.mask(q, 4600), err_msg=str(q))
