"""GPU tests of the engine GROUP (include/fn3.h fn3_group_*, engine.GroupEngine, seqComparer(device=[...])): one process, the
store sharded over several GPUs — every neighbour list against the C oracle on the FLAT collection.

The devices are whatever the box shows: all of them when there are several, otherwise the same GPU listed three times
(three shards, three host threads, "peer" reads inside one device) — the group logic is identical."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _devices():
    import torch
    n = torch.cuda.device_count()
    return list(range(min(n, 8))) if n >= 2 else [0, 0, 0]


def _synthetic(n_anc, max_c, L, seed, **kw):
    from findneighbour3_b200 import synth
    excluded = synth.synthetic_mask(L, L // 15, 40, seed=seed + 100)
    return synth.make_collection(n_anc, max_c, genome_length=L, seed=seed, excluded=excluded, **kw)


def hits_to_dict(hits):
    return {int(h["row"]): (int(h["dist"]), int(h["n1"]), int(h["n2"]), int(h["nboth"])) for h in hits}


@pytest.mark.parametrize("shape", ["tb_like", "high_nm"])
def test_group_queries_vs_c_oracle(shape):
    from findneighbour3_b200.engine import GroupEngine
    from oracle import c_oracle
    if shape == "tb_like":
        coll = _synthetic(14, 25, 500000, 1, k1=500, s=3.0, n_blocks=20, n_mean=4000, rule="lss")
        ceilings = (12, 20, 250, 2 ** 31 - 1)
    else:
        coll = _synthetic(6, 30, 400000, 4, k1=3000, s=5.0, n_blocks=40, n_mean=30000, m_mean=300, rule="any",
                          n_lognormal_sigma=0.8, max_ns=60000, p_invalid=0.02)
        ceilings = (20, 5000)
    devs = _devices()
    eng = GroupEngine(coll.genome_length, devs)
    assert eng.n_devices == len(devs)
    bulk = coll.n - 40
    sub = coll.subset(np.arange(bulk))
    assert eng.append_bulk(sub.pos, sub.off, sub.invalid) == 0
    # streaming inserts, each queried against everything stored before (server insert(): persist + mcompare)
    for i in range(bulk, coll.n):
        p, o, inv = coll.lists(i)
        row, hits = eng.insert_query(p, o, inv, 20)
        assert row == i
        upto = coll.subset(np.arange(i + 1))
        assert hits_to_dict(hits) == c_oracle.one_vs_all(upto.pos, upto.off, upto.invalid, i, 20), (shape, i)
    st = eng.stats()
    assert st["n_rows"] == coll.n and st["n_invalid"] == int(coll.invalid.sum())
    rng = np.random.default_rng(0)
    queries = rng.choice(coll.n, size=14, replace=False)
    for ceiling in ceilings:
        for q in queries:
            want = c_oracle.one_vs_all(coll.pos, coll.off, coll.invalid, int(q), ceiling, nthreads=4)
            assert hits_to_dict(eng.one_vs_all(int(q), ceiling)) == want, (shape, ceiling, int(q))
            # the same sample as a query that is not stored: its own stored copy answers too
            p, o, inv = coll.lists(int(q))
            got = hits_to_dict(eng.lists_vs_all(p, o, inv, ceiling))
            got.pop(int(q), None)
            assert got == want
    # explicit candidate lists, stored rows read back, pairs across shards
    q = int(queries[0])
    cand = rng.choice(coll.n, size=60, replace=False)
    want = c_oracle.one_vs_all(coll.pos, coll.off, coll.invalid, q, 250, nthreads=4)
    got = hits_to_dict(eng.one_vs_all(q, 250, rows=cand))
    assert got == {r: v for r, v in want.items() if r in set(int(c) for c in cand)}
    for r in (0, 1, 2, coll.n - 1):
        p, o, inv = eng.row_get(r)
        wp, wo, winv = coll.lists(r)
        assert inv == winv and (inv or (np.array_equal(p, wp) and np.array_equal(o, wo)))
    for a, b in [(q, int(queries[1])), (q, q), (int(queries[2]), q), (3, 4), (4, 7)]:
        pa, oa, ia = coll.lists(a)
        pb, ob, ib = coll.lists(b)
        assert eng.pair(a, b, 10 ** 6) == c_oracle.pair(pa, oa, ia, pb, ob, ib, 10 ** 6), (a, b)
    # remove: the row never matches again, on whichever shard it lives
    victims = [r for r in want if r != q][:3]
    for v in victims:
        eng.remove(v)
    got = hits_to_dict(eng.one_vs_all(q, 250))
    assert got == {r: v for r, v in want.items() if r not in victims}
    if victims:
        with pytest.raises(KeyError):
            eng.one_vs_all(victims[0], 20)
    eng.close()


def test_group_all_vs_all_and_cluster_statistics_equal_one_engine():
    """distmat over the sharded store (queries read across shards), consensus and MSA over rows of several shards: the same
    answers as ONE engine holding the same rows (which the other tests pin to the oracle)."""
    from findneighbour3_b200 import synth
    from findneighbour3_b200.engine import Engine, GroupEngine
    coll = _synthetic(8, 20, 300000, 7, k1=400, s=3.0, n_blocks=10, n_mean=2500, m_mean=6, rule="any", p_invalid=0.04)
    one = Engine(coll.genome_length, 0)
    grp = GroupEngine(coll.genome_length, _devices())
    one.append_bulk(coll.pos, coll.off, coll.invalid)
    grp.append_bulk(coll.pos, coll.off, coll.invalid)
    for ceiling in (12, 250):
        for upper in (True, False):
            a = one.all_vs_all(ceiling, upper_only=upper)
            b = grp.all_vs_all(ceiling, upper_only=upper)
            key = lambda e: sorted(zip(e["row1"].tolist(), e["row2"].tolist(), e["dist"].tolist(), e["n1"].tolist(),
                                       e["n2"].tolist(), e["nboth"].tolist()))
            assert key(a) == key(b) and len(a) > 0, (ceiling, upper)
    ref = synth.reference_string(synth.random_reference(coll.genome_length))
    one.set_reference(ref)
    grp.set_reference(ref)
    members = [int(i) for i in np.flatnonzero((coll.family == 2) & (coll.invalid == 0))]
    members = members + members[:2]                                   # a row given twice votes twice
    for mv in (1, 3, len(members)):
        pa, oa = one.consensus(members, mv)
        pb, ob = grp.consensus(members, mv)
        assert np.array_equal(pa, pb) and np.array_equal(oa, ob)
    for n_is_call in (False, True):
        sa, sb = one.msa_sites(members, n_is_call), grp.msa_sites(members, n_is_call)
        assert np.array_equal(sa, sb) and sa.size > 0
        ca, na = one.msa_align(members, sa)
        cb, nb = grp.msa_align(members, sb)
        assert np.array_equal(ca, cb) and np.array_equal(na, nb)
    one.close()
    grp.close()


def test_class_on_a_group_runs_the_class_level_tests(monkeypatch):
    """seqComparer with FN3_DEVICES set (what a deployment of the unmodified server uses to hand the class several GPUs):
    the class-level golden tests — known-answer distances, byKey, random pairs, recompression, MSA, cluster statistics."""
    import test_gpu_parity as T
    monkeypatch.setenv("FN3_DEVICES", ",".join(str(d) for d in _devices()))
    T.test_kat_count_differences()
    T.test_kat_mcompare_bykey_hash()
    T.test_random_pairs_through_class(20)
    T.test_consensus_and_recompression_golden()
    T.test_msa_golden()
    T.test_cluster_golden_through_class()
    from findneighbour3_b200.seqComparer import seqComparer
    sc = seqComparer(reference="ACTG", maxNs=1e8, snpCeiling=20)
    assert sc.engine.n_devices == len(_devices())
    for s in ("AAAA", "CCCC", "RRCC", "NNTG"):
        sc.persist(sc.compress(s), s)
    got = {r[1]: r[2:6] for r in sc.mcompare("AAAA")}
    assert got == {"CCCC": [4, 0, 0, 0], "NNTG": [2, 0, 2, 2], "RRCC": [2, 0, 0, 0]}
    assert len(list(sc.distmat(half=False))) == 12


def test_class_insert_sequence_on_a_group():
    """the server's insert route from the sequence string (fn3_group_ingest_query): compress on the owner device, persist,
    one-vs-all over every shard — against compress + persist + mcompare of the Python port"""
    from findneighbour3_b200.seqComparer import seqComparer
    from oracle import seqcomparer_port as port
    sc = seqComparer(reference="ACTGACTGAC", maxNs=4, snpCeiling=20, excludePositions={9}, device=_devices())
    store = {}
    for s in ("ACTGACTGAC", "NCTGACTAAC", "A-TGRCTGAC", "CCCCCCCCCC", "NNNNNCTGAC", "ACTGACTGAA", "ACTGACTTAC"):
        found, hist = sc.insert_sequence(s, s)
        store[s] = port.compress(sc.reference, sorted(sc.included), 4, s)
        assert sc.seqProfile[s] == store[s]
        want = {r[1]: tuple(r[2:6]) for r in port.mcompare(store, {}, s, 20) if r[2] is not None}
        assert found == want, (s, found, want)
        assert int(hist[ord("A")]) == s.count("A")


def test_group_rejects_a_bad_sample_before_anything_is_stored():
    """a bulk append with ONE malformed sample (out of range, unsorted, overlapping lists) fails as a whole — no shard keeps
    its share — and the group stays in step: the next append and every query work, rows are numbered as if nothing happened"""
    from findneighbour3_b200.engine import GroupEngine
    from oracle import c_oracle
    coll = _synthetic(4, 12, 100000, 3, k1=100, s=2.0, n_blocks=4, n_mean=500, m_mean=2, rule="any")
    eng = GroupEngine(coll.genome_length, _devices())
    first = coll.subset(np.arange(20))
    assert eng.append_bulk(first.pos, first.off, first.invalid) == 0
    rest = coll.subset(np.arange(20, coll.n))
    for kind in ("range", "order", "overlap"):
        pos = rest.pos.copy()
        i = 11                                                    # a sample whose shard is not the first one
        o = rest.off[6 * i: 6 * i + 7]
        assert o[1] - o[0] >= 2
        if kind == "range":
            pos[o[0]] = coll.genome_length + 5
        elif kind == "order":
            pos[o[0]], pos[o[0] + 1] = pos[o[0] + 1], pos[o[0]]
        else:
            pos[o[4]] = pos[o[0]] if o[5] > o[4] else pos[o[0]]
            if o[5] == o[4]:
                continue
        with pytest.raises(ValueError):
            eng.append_bulk(pos, rest.off, rest.invalid)
        assert eng.stats()["n_rows"] == 20
    assert eng.append_bulk(rest.pos, rest.off, rest.invalid) == 20
    for q in (0, 5, 21, coll.n - 1):
        want = c_oracle.one_vs_all(coll.pos, coll.off, coll.invalid, q, 20)
        assert hits_to_dict(eng.one_vs_all(q, 20)) == want
    eng.close()
