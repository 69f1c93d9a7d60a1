"""gpu: the low-ceiling head filter on collections built to defeat it — results must stay bit-exact against
the C oracle, and the two counter-measures must engage: background-aware heads (deep shared ancestry) and the
FILTER -> STREAM steering (an outbreak: every candidate is a neighbour)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def hits_to_dict(hits):
    return {int(h["row"]): (int(h["dist"]), int(h["n1"]), int(h["n2"]), int(h["nboth"])) for h in hits}


def _pass_rate(eng, rows, ceiling):
    pr = eng.profile_queries(rows, ceiling, flush_l2=False, count_bytes=True)
    return float(pr["n_passed"].mean()) / max(eng.stats()["n_live"] - 1, 1)


def test_lineage_collection_background_heads(monkeypatch):
    """4 lineages x 4 sub-lineages: unrelated samples of one lineage share ~800 of their ~1 450 variant positions, so heads made
    of the 32 lowest coordinates let a quarter of the collection through; heads made of private variants do not."""
    from findneighbour3_b200 import synth
    from findneighbour3_b200.engine import Engine
    from oracle import c_oracle
    L = 1_500_000
    coll = synth.make_lineages(120, 50, genome_length=L, tb_mask=False, workers=4, k1=400, n_mean=1500, n_blocks=8)
    assert coll.n == 6000
    monkeypatch.setenv("FN3_BG_AUTO", "0")
    plain = Engine(L, 0)
    monkeypatch.setenv("FN3_BG_AUTO", "1")
    eng = Engine(L, 0)
    for e in (plain, eng):
        e.append_bulk(coll.pos, coll.off, coll.invalid)
    rng = np.random.default_rng(3)
    queries = [int(q) for q in rng.choice(np.flatnonzero(coll.invalid == 0), size=10, replace=False)]
    for ceiling in (12, 20):
        for q in queries:
            want = c_oracle.one_vs_all(coll.pos, coll.off, coll.invalid, q, ceiling, nthreads=4)
            assert hits_to_dict(plain.one_vs_all(q, ceiling)) == want
            assert hits_to_dict(eng.one_vs_all(q, ceiling)) == want, (ceiling, q)
    assert plain.stats()["head_refreshes"] == 0 and eng.stats()["head_refreshes"] == 1      # 6 000 rows >= 4 096: one refresh
    r_plain, r_bg = _pass_rate(plain, queries, 20), _pass_rate(eng, queries, 20)
    assert r_plain > 0.10 and r_bg < 0.25 * r_plain, (r_plain, r_bg)
    # rows inserted AFTER the refresh get background-aware heads on the host side (fill_head), removed rows stay out
    extra = [synth.new_sample_like(coll, q, 2, seed=900 + i) for i, q in enumerate(queries)]
    flat = coll
    for i, s in enumerate(extra):
        row, hits = eng.insert_query(s[0], s[1], s[2], 20)
        pos = np.concatenate([flat.pos, s[0]])
        off = np.concatenate([flat.off, flat.off[-1] + s[1][1:].astype(np.int64)])
        inv = np.concatenate([flat.invalid, np.zeros(1, np.uint8)])
        flat = synth.Collection(L, pos, off, inv, np.concatenate([flat.family, [-1]]).astype(np.int32))
        assert row == flat.n - 1
        assert hits_to_dict(hits) == c_oracle.one_vs_all(flat.pos, flat.off, flat.invalid, row, 20, nthreads=4)
    eng.remove(queries[0])
    want = c_oracle.one_vs_all(flat.pos, flat.off, flat.invalid, flat.n - 1, 20, nthreads=4)
    want.pop(queries[0], None)
    assert hits_to_dict(eng.one_vs_all(flat.n - 1, 20)) == want
    # an explicit refresh with another threshold, and clearing the background, change nothing in the answers
    for ppm in (200000, 0):
        eng.refresh_heads(ppm)
        assert hits_to_dict(eng.one_vs_all(flat.n - 1, 20)) == want
    assert eng.stats()["head_refreshes"] == 3
    plain.close()
    eng.close()


def test_outbreak_collection_is_steered_to_the_streaming_kernel():
    """every sample within a few SNVs of every other: nothing can be ruled out from a head; after such a query full scans at
    snpCeiling 20 go to k_scan (stats: scan_mode) and come back to the filter when the hit rate drops"""
    from findneighbour3_b200 import synth
    from findneighbour3_b200.engine import Engine
    from oracle import c_oracle
    L = 1_000_000
    ob = synth.make_outbreak(3000, piece=300, genome_length=L, tb_mask=False, workers=4, k1=400, n_mean=1000, n_blocks=6)
    eng = Engine(L, 0)
    eng.append_bulk(ob.pos, ob.off, ob.invalid)
    rng = np.random.default_rng(5)
    queries = [int(q) for q in rng.choice(np.flatnonzero(ob.invalid == 0), size=8, replace=False)]
    modes = []
    for q in queries:
        want = c_oracle.one_vs_all(ob.pos, ob.off, ob.invalid, q, 20, nthreads=4)
        assert len(want) > 0.9 * (ob.n - int(ob.invalid.sum()))
        assert hits_to_dict(eng.one_vs_all(q, 20)) == want, q
        modes.append(eng.stats()["scan_mode"])
    assert modes[-1] == 1, modes
    # unrelated samples appended and queried: few hits -> after a streak of such queries the filter is tried again
    star = synth.make_collection(60, 50, genome_length=L, seed=77, k1=400, n_mean=1000, n_blocks=6)
    first = eng.append_bulk(star.pos, star.off, star.invalid)
    flat = synth.concat(ob, star)
    for i in range(0, 12):
        q = first + 50 * i + 3
        if flat.invalid[q]:
            continue
        want = c_oracle.one_vs_all(flat.pos, flat.off, flat.invalid, q, 20, nthreads=4)
        assert hits_to_dict(eng.one_vs_all(q, 20)) == want
    assert eng.stats()["scan_mode"] == 0
    eng.close()
