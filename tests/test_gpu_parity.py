"""gpu: the CUDA path (through the C-ABI and through the drop-in seqComparer class) against
 (a) the golden vectors produced by the unmodified reference, and
 (b) the C oracle on seeded synthetic collections, bit-exact (integer work: no tolerance)."""
import os

import numpy as np
import pytest

from util import dict_from_json, golden, lists_from_dict

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def libs_loaded():
    from findneighbour3_b200 import _native
    _native.load()


def _sc(ref, maxNs, ceiling, excluded=(), **kw):
    from findneighbour3_b200.seqComparer import seqComparer
    return seqComparer(reference=ref, maxNs=maxNs, snpCeiling=ceiling, excludePositions=set(excluded), **kw)


def hits_to_dict(hits):
    return {int(h["row"]): (int(h["dist"]), int(h["n1"]), int(h["n2"]), int(h["nboth"])) for h in hits}


# ---------------------------------------------------------------------------------------------
# (a) golden vectors
# ---------------------------------------------------------------------------------------------
def test_compress_golden():
    for case in golden("compress.json"):
        sc = _sc(case["ref"], case["maxNs"], 20, case["excluded"])
        got = sc.compress(case["seq"])
        assert got == dict_from_json(case["got"]), case
        if "uncompress" in case:
            assert sc.uncompress(got) == case["uncompress"]


def test_kat_count_differences():
    n = 0
    for case in golden("kat.json"):
        if case["kind"] != "countDifferences":
            continue
        sc = _sc(case["ref"], case["maxNs"], case["snpCeiling"])
        sc.setComparator1(case["seq1"])
        sc.setComparator2(case["seq2"])
        assert sc.countDifferences() == case["literal"], case
        # and with _seq1/_seq2 assigned directly, as the reference's tests 16/17 do
        sc._seq1 = sc.compress(case["seq1"])
        sc._seq2 = sc.compress(case["seq2"])
        assert sc.countDifferences() == case["literal"], case
        n += 1
    assert n >= 20


def test_kat_mcompare_bykey_hash():
    for case in golden("kat.json"):
        if case["kind"] in ("mcompare", "mcompare_dist"):
            sc = _sc(case["ref"], case["maxNs"], case["snpCeiling"])
            for s in case["samples"]:
                sc.persist(sc.compress(s), s)
            res = sc.mcompare(case["query"])
            assert len(res) == len(case["samples"]) - 1
            if case["kind"] == "mcompare":
                assert {r[1]: r[2:6] for r in res} == case["literal"]
                by = {r[1]: r for r in res}
                assert by["NNTG"][6:9] == [set(), {0, 1}, {0, 1}]
            else:
                assert {r[1]: r[2] for r in res} == case["literal"]
            sub = sc.mcompare(case["query"], case["samples"][1:3])
            assert len(sub) == 2
        elif case["kind"] == "byKey":
            sc = _sc(case["ref"], case["maxNs"], case["snpCeiling"])
            with pytest.raises(KeyError):
                sc.countDifferences_byKey(("k1", "k2"))
            sc.persist(sc.compress(case["seq1"]), "k1")
            sc.persist(sc.compress(case["seq2"]), "k2")
            got = sc.countDifferences_byKey(("k1", "k2"), cutoff=None)
            assert list(got[2:6]) == case["got"]
            with pytest.raises(TypeError):
                sc.countDifferences_byKey(["k1", "k2"])
            with pytest.raises(TypeError):
                sc.countDifferences_byKey(("k1", "k2", "k3"))
        elif case["kind"] == "excluded_hash":
            sc = _sc(case["ref"], 1e8, 1, case["excluded"])
            assert sc.excluded_hash() == case["got"]
        elif case["kind"] == "setStats":
            sc = _sc(case["ref"], 1e8, 20)
            c1, c2 = sc.compress(case["seq1"]), sc.compress(case["seq2"])
            n1, n2, nall, rv1, rv2, rv = sc._setStats(c1[case["key"]], c2[case["key"]])
            assert [n1, n2, nall, sorted(rv)] == case["got"]


@pytest.mark.parametrize("ceiling", [20, 1000])
def test_random_pairs_through_class(ceiling):
    g = golden("random_pairs.json")
    sc = _sc(g["ref"], g["maxNs"], ceiling, g["excluded"])
    for i, s in enumerate(g["seqs"]):
        sc.persist(sc.compress(s), i)
    rows = []
    for i in range(len(g["seqs"])):
        for r in sc.mcompare(i):
            if r[2] is not None:
                rows.append([r[0], r[1], r[2], r[3], r[4], r[5]])
                # the position sets must agree with the counts (reference: _setStats)
                assert (len(r[6]), len(r[7]), len(r[8])) == (r[3], r[4], r[5])
    assert sorted(rows) == g["results"][str(ceiling)]


def _collection_engine(chunk_env=None):
    from findneighbour3_b200.engine import Engine
    g = golden("collection.json")
    pos, off, inv = np.array(g["pos"], np.int32), np.array(g["off"], np.int64), np.array(g["invalid"], np.uint8)
    eng = Engine(g["genome_length"], 0)
    first = eng.append_bulk(pos, off, inv)
    assert first == 0
    return g, eng, pos, off, inv


@pytest.mark.parametrize("ceiling", [12, 20, 250])
def test_collection_one_vs_all_golden(ceiling):
    g, eng, pos, off, inv = _collection_engine()
    n = inv.size
    for q in range(n):
        got = hits_to_dict(eng.one_vs_all(q, ceiling))
        want = {r[0]: tuple(r[1:]) for r in g["neighbours"][str(ceiling)][str(q)]}
        assert got == want, (q, ceiling)
    eng.close()


@pytest.mark.parametrize("ceiling", [12, 250])
def test_collection_all_vs_all_golden(ceiling):
    g, eng, pos, off, inv = _collection_engine()
    edges = eng.all_vs_all(ceiling, upper_only=False)
    got = {(int(e["row1"]), int(e["row2"])): (int(e["dist"]), int(e["n1"]), int(e["n2"]), int(e["nboth"])) for e in edges}
    want = {}
    for q, rows in g["neighbours"][str(ceiling)].items():
        for r in rows:
            want[(int(q), r[0])] = tuple(r[1:])
    assert got == want
    upper = eng.all_vs_all(ceiling, upper_only=True)
    got_u = {(int(e["row1"]), int(e["row2"])): int(e["dist"]) for e in upper}
    assert got_u == {k: v[0] for k, v in want.items() if k[0] < k[1]}
    # block-row ownership (multi-GPU sharding): the union over phases equals the whole
    parts = {}
    for phase in range(3):
        for e in eng.all_vs_all(ceiling, upper_only=True, stride=3, phase=phase):
            assert int(e["row1"]) % 3 == phase
            parts[(int(e["row1"]), int(e["row2"]))] = int(e["dist"])
    assert parts == got_u
    eng.close()


def test_collection_subset_remove_pair_rowget():
    g, eng, pos, off, inv = _collection_engine()
    want20 = {int(q): {r[0]: tuple(r[1:]) for r in rows} for q, rows in g["neighbours"]["20"].items()}
    q = max(want20, key=lambda k: len(want20[k]))
    rows = np.array(sorted(want20[q].keys())[:5] + [q, (q + 1) % inv.size], dtype=np.int64)
    got = hits_to_dict(eng.one_vs_all(q, 20, rows))
    assert got == {r: v for r, v in want20[q].items() if r in set(rows.tolist())}
    # pair == one entry of one-vs-all; self pair is 0 for valid rows
    r2 = sorted(want20[q])[0]
    assert eng.pair(q, r2, 20) == want20[q][r2]
    assert eng.pair(q, q, 20)[0] == 0
    bad = int(np.flatnonzero(inv)[0])
    assert eng.pair(q, bad, 20) is None and eng.pair(bad, q, 20) is None
    assert len(eng.one_vs_all(bad, 20)) == 0
    # row_get round trip
    o = off[6 * q: 6 * q + 7]
    p, f, i = eng.row_get(q)
    assert not i and np.array_equal(p, pos[o[0]:o[6]]) and np.array_equal(f, (o - o[0]).astype(np.int32))
    # lists_vs_all with the same content == one_vs_all (+ the stored twin at distance 0)
    lv = hits_to_dict(eng.lists_vs_all(p, f, False, 20))
    assert lv.pop(q)[0] == 0
    assert lv == want20[q]
    # remove a neighbour: it disappears; removing twice is silent; unknown row raises
    eng.remove(r2)
    eng.remove(r2)
    got = hits_to_dict(eng.one_vs_all(q, 20))
    assert got == {r: v for r, v in want20[q].items() if r != r2}
    with pytest.raises(KeyError):
        eng.remove(10 ** 6)
    with pytest.raises(KeyError):
        eng.one_vs_all(r2, 20)
    st = eng.stats()
    assert st["n_rows"] == inv.size and st["n_live"] == inv.size - 1
    eng.close()


def test_bad_input_is_rejected():
    from findneighbour3_b200.engine import Engine
    eng = Engine(100, 0)
    off = np.array([0, 2, 2, 2, 2, 2, 2], np.int32)
    with pytest.raises(ValueError):
        eng.append(np.array([5, 5], np.int32), off)          # not strictly increasing
    with pytest.raises(ValueError):
        eng.append(np.array([5, 100], np.int32), off)        # out of range
    # a position in two lists (A and N; C and an N run; N and M): every entry point that takes lists refuses it
    for pos_, off_ in (([5, 9, 9], [0, 2, 2, 2, 2, 3, 3]), ([7, 20, 4, 5, 6, 7, 8], [0, 0, 2, 2, 2, 7, 7]),
                       ([1, 2, 3, 3], [0, 0, 0, 0, 0, 3, 4])):
        p_, o_ = np.array(pos_, np.int32), np.array(off_, np.int32)
        with pytest.raises(ValueError):
            eng.append(p_, o_)
        with pytest.raises(ValueError):
            eng.insert_query(p_, o_, False, 20)
        with pytest.raises(ValueError):
            eng.lists_vs_all(p_, o_, False, 20)
        with pytest.raises(ValueError):
            eng.append_bulk(p_, np.array(off_, np.int64))
    assert eng.stats()["n_rows"] == 0
    eng.append(np.array([5, 9, 10], np.int32), np.array([0, 2, 2, 2, 2, 3, 3], np.int32))    # disjoint: accepted
    assert eng.stats()["n_rows"] == 1
    eng.close()


def test_consensus_and_recompression_golden():
    g = golden("consensus.json")
    sc = _sc("ACTG", 1e8, 10)
    assert sc.compressed_sequence_hash(sc.compress(g["hash"]["seq"])) == g["hash"]["literal"]
    for case in g["consensus"]:
        comp = [sc.compress(s) for s in case["seqs"]]
        cons = sc.consensus(comp, case["cutoff"])
        assert {k: sorted(v) for k, v in cons.items()} == case["got"]
        assert sc.compressed_sequence_hash(cons) == case["hash"]
    comp = [sc.compress(s) for s in g["consensus"][0]["seqs"]]
    cons = sc.consensus(comp, 0.6)
    for c, case in zip(comp, g["patch"]):
        patch = sc.generate_patch(c, cons)
        assert {k: sorted(v) for k, v in patch["+"].items()} == case["patch"]["+"]
        assert {k: sorted(v) for k, v in patch["-"].items()} == case["patch"]["-"]
        assert sc.apply_patch(patch, cons) == c
    r = g["recompress"]
    sc = _sc(r["ref"], 1e8, 20, snpCompressionCeiling=r["snpCompressionCeiling"])
    for i, s in enumerate(r["seqs"]):
        sc.persist(sc.compress(s), "g%d" % i)
    before = {x[1]: x[2:6] for x in sc.mcompare("g0")}
    assert before == r["mcompare_g0"]
    launches = sc.engine.stats()["n_rows"]
    visited = sc.compress_relative_to_consensus("g0", 0.8)
    assert sorted(visited) == r["visited"]
    assert sc.summarise_stored_items() == r["summary"]
    assert sc.engine.stats()["n_rows"] == launches          # re-encoding as patches did not touch the device store
    assert {x[1]: x[2:6] for x in sc.mcompare("g0")} == r["mcompare_g0"]
    for i, s in enumerate(r["seqs"]):
        assert sc.uncompress(sc.seqProfile["g%d" % i]) == s


# ---------------------------------------------------------------------------------------------
# (b) seeded synthetic collections vs the C oracle
# ---------------------------------------------------------------------------------------------
def _synthetic(n_anc, max_c, L, seed, **kw):
    from findneighbour3_b200 import synth
    excluded = synth.synthetic_mask(L, L // 15, 40, seed=seed + 100)
    return synth.make_collection(n_anc, max_c, genome_length=L, seed=seed, excluded=excluded, **kw)


@pytest.mark.parametrize("staged", [True, False])
@pytest.mark.parametrize("shape", ["tb_like", "high_nm", "tiny_genome", "long_runs", "scattered_n"])
def test_synthetic_vs_c_oracle(shape, staged, monkeypatch):
    """staged=False forces the kernels' global-memory query variants (what a query touching more blocks than
    fit in shared memory, or a genome whose block bitmap exceeds it, would use)."""
    from findneighbour3_b200.engine import Engine
    from oracle import c_oracle
    if not staged:
        monkeypatch.setenv("FN3_FORCE_GLOBAL_QUERY", "1")
    if shape == "tb_like":        # config 1 shape at reduced genome length
        coll = _synthetic(12, 25, 500000, 1, k1=500, s=3.0, n_blocks=20, n_mean=4000, rule="lss")
        ceilings = (20, 250, 2 ** 31 - 1)
    elif shape == "high_nm":      # config 4 shape: many N (lognormal) and IUPAC M, some invalid through maxNs
        coll = _synthetic(6, 30, 400000, 4, k1=3000, s=5.0, n_blocks=40, n_mean=30000, m_mean=300, rule="any",
                          n_lognormal_sigma=0.8, max_ns=60000, p_invalid=0.02)
        ceilings = (20, 5000)
    elif shape == "tiny_genome":  # genome shorter than one rank block / one run word
        coll = _synthetic(5, 8, 31, 9, k1=3, s=1.0, n_blocks=2, n_mean=4, m_mean=1, rule="any", p_invalid=0.1)
        ceilings = (0, 1, 20)
    elif shape == "scattered_n":  # i.i.d. N (the reference generator's model): tens of thousands of touched blocks,
        # more than shared memory can stage -> the engine itself falls back to the global-memory variant
        coll = _synthetic(3, 8, 4000000, 15, k1=300, s=2.0, n_blocks=60000, n_mean=60000, m_mean=50, rule="any")
        ceilings = (20, 400)
    else:                         # runs much longer than the run-word limit and > maxNs-like sizes
        coll = _synthetic(4, 10, 3000000, 12, k1=200, s=2.0, n_blocks=3, n_mean=120000, m_mean=20, rule="any")
        ceilings = (20, 400)
    eng = Engine(coll.genome_length, 0)
    eng.append_bulk(coll.pos, coll.off, coll.invalid)
    rng = np.random.default_rng(0)
    queries = rng.choice(coll.n, size=min(coll.n, 12), replace=False)
    for ceiling in ceilings:
        for q in queries:
            got = hits_to_dict(eng.one_vs_all(int(q), ceiling))
            want = c_oracle.one_vs_all(coll.pos, coll.off, coll.invalid, int(q), ceiling, nthreads=4)
            assert got == want, (shape, ceiling, int(q))
    eng.close()


@pytest.mark.parametrize("split", [None, 0, 1, 2, 3, 5])
def test_scan_every_split_of_the_head_tiles(split, monkeypatch):
    """k_scan hands its warps sub-units of a 32-row head tile; the split (fn3_engine.cu pick_scan_split, or FN3_SCAN_SS) must
    not show in the result: one full superblock of 8 192 rows plus a partly filled one, full scans and explicit candidate
    lists, at the recompression ceiling and with none."""
    from findneighbour3_b200.engine import Engine
    from oracle import c_oracle
    if split is not None:
        monkeypatch.setenv("FN3_SCAN_SS", str(split))
    coll = _synthetic(190, 50, 300000, 21, k1=150, s=3.0, n_blocks=6, n_mean=600, m_mean=2, rule="any", p_invalid=0.01)
    assert coll.n > 8192 + 300
    eng = Engine(coll.genome_length, 0)
    eng.append_bulk(coll.pos, coll.off, coll.invalid)
    rng = np.random.default_rng(split or 0)
    queries = [int(q) for q in rng.choice(coll.n, size=6, replace=False)]
    subset = np.sort(rng.choice(coll.n, size=3001, replace=False)).astype(np.int64)
    in_subset = set(int(r) for r in subset)
    for ceiling in (250, 2 ** 31 - 1):
        for q in queries:
            want = c_oracle.one_vs_all(coll.pos, coll.off, coll.invalid, q, ceiling, nthreads=4)
            assert hits_to_dict(eng.one_vs_all(q, ceiling)) == want, (split, ceiling, q)
            got = hits_to_dict(eng.one_vs_all(q, ceiling, rows=subset))
            assert got == {r: v for r, v in want.items() if r in in_subset}, (split, ceiling, q, "list")
    eng.close()


def test_incremental_insert_then_query_matches_bulk(monkeypatch):
    """config 1 flow: persist one by one with a query after each insert; small store chunks force the arena
    and the header table to roll over."""
    from findneighbour3_b200.engine import Engine
    from oracle import c_oracle
    monkeypatch.setenv("FN3_CHUNK_MB", "1")
    monkeypatch.setenv("FN3_HDR_SHIFT", "4")
    coll = _synthetic(6, 20, 300000, 21, k1=400, s=3.0, n_blocks=10, n_mean=3000)
    eng = Engine(coll.genome_length, 0)
    for i in range(coll.n):
        p, o, inv = coll.lists(i)
        assert eng.append(p, o, inv) == i
        if i % 7 == 3 or i == coll.n - 1:
            sub = coll.subset(np.arange(i + 1))
            want = c_oracle.one_vs_all(sub.pos, sub.off, sub.invalid, i, 20)
            assert hits_to_dict(eng.one_vs_all(i, 20)) == want
    st = eng.stats()
    assert st["n_rows"] == coll.n and st["store_bytes"] > (1 << 20)
    eng.close()


def test_insert_query_fused_call_matches_two_calls():
    """fn3_insert_query (persist + mcompare in one ABI call) == fn3_append followed by fn3_one_vs_all."""
    from findneighbour3_b200.engine import Engine
    from oracle import c_oracle
    coll = _synthetic(5, 16, 250000, 44, k1=300, s=3.0, n_blocks=8, n_mean=2500, m_mean=4, rule="any", p_invalid=0.05)
    a, b = Engine(coll.genome_length, 0), Engine(coll.genome_length, 0)
    for i in range(coll.n):
        p, o, inv = coll.lists(i)
        row, hits = a.insert_query(p, o, inv, 20)
        assert row == i == b.append(p, o, inv)
        two = hits_to_dict(b.one_vs_all(i, 20))
        assert hits_to_dict(hits) == two
        if i % 9 == 0:
            sub = coll.subset(np.arange(i + 1))
            assert two == c_oracle.one_vs_all(sub.pos, sub.off, sub.invalid, i, 20)
    assert a.stats()["n_rows"] == b.stats()["n_rows"] == coll.n
    a.close()
    b.close()


def test_properties_at_scale():
    """size-independent properties on a TB-sized genome: symmetry of the distance, zero self distance,
    hits at a low ceiling are a subset of hits at a higher one with identical values, monotone counts."""
    from findneighbour3_b200 import synth
    from findneighbour3_b200.engine import Engine
    coll = synth.make_collection(40, 50, seed=5, excluded=synth.synthetic_mask(synth.TB_LENGTH, synth.TB_EXCLUDED,
                                                                               synth.TB_EXCLUDED_RUNS))
    eng = Engine(coll.genome_length, 0)
    eng.append_bulk(coll.pos, coll.off, coll.invalid)
    rng = np.random.default_rng(1)
    valid = np.flatnonzero(coll.invalid == 0)
    for q in rng.choice(valid, size=10, replace=False):
        q = int(q)
        lo = hits_to_dict(eng.one_vs_all(q, 12))
        hi = hits_to_dict(eng.one_vs_all(q, 250))
        assert set(lo) <= set(hi)
        for r, v in lo.items():
            assert hi[r] == v and v[0] <= 12
        assert eng.pair(q, q, 0)[0] == 0
        for r in list(hi)[:20]:
            back = eng.pair(r, q, 250)
            assert back is not None and back[0] == hi[r][0]
        same_family = np.flatnonzero((coll.family == coll.family[q]) & (coll.invalid == 0))
        assert len(hi) >= 1 and set(hi) <= set(same_family.tolist())
    eng.close()


def test_threads_append_while_querying():
    """persist runs outside the server's write semaphore (findNeighbour3-server.py:513-516): appends from one
    thread while another thread queries must neither crash nor corrupt earlier rows."""
    import threading
    from findneighbour3_b200.engine import Engine
    from oracle import c_oracle
    coll = _synthetic(4, 25, 200000, 33, k1=300, s=3.0, n_blocks=8, n_mean=2000)
    eng = Engine(coll.genome_length, 0)
    half = coll.n // 2
    first = coll.subset(np.arange(half))
    eng.append_bulk(first.pos, first.off, first.invalid)
    want = {q: c_oracle.one_vs_all(first.pos, first.off, first.invalid, q, 20) for q in range(0, half, 5)}
    errors = []

    def writer():
        try:
            for i in range(half, coll.n):
                p, o, inv = coll.lists(i)
                eng.append(p, o, inv)
        except Exception as ex:      # pragma: no cover
            errors.append(ex)

    t = threading.Thread(target=writer)
    t.start()
    for _ in range(3):
        for q in want:
            got = {r: v for r, v in hits_to_dict(eng.one_vs_all(q, 20)).items() if r < half}
            assert got == want[q]
    t.join()
    assert not errors
    full = c_oracle.one_vs_all(coll.pos, coll.off, coll.invalid, 0, 20)
    assert hits_to_dict(eng.one_vs_all(0, 20)) == full
    eng.close()


@pytest.mark.parametrize("kind", ["engine", "group"])
def test_stress_inserts_removals_and_queries_racing(kind):
    """Stands in for the racecheck the pool does not offer: four threads hammer one store — fused inserts (the kernel places the
    head by value), appends, removals (k_put_head from another stream) and queries at both kernel variants — then, at
    quiescence, EVERY live row is queried and compared with the C oracle on exactly the rows that are live."""
    import threading
    from findneighbour3_b200.engine import Engine, GroupEngine
    from oracle import c_oracle
    coll = _synthetic(10, 40, 300000, 61, k1=300, s=3.0, n_blocks=8, n_mean=2000, m_mean=3, rule="any", p_invalid=0.03)
    eng = Engine(coll.genome_length, 0) if kind == "engine" else GroupEngine(coll.genome_length, [0, 0, 0])
    base = 200
    first = coll.subset(np.arange(base))
    eng.append_bulk(first.pos, first.off, first.invalid)
    errors, removed, row_of = [], set(), {i: i for i in range(base)}
    lock = threading.Lock()
    stop = threading.Event()

    def inserter(lo, hi, fused):
        try:
            for i in range(lo, hi):
                p, o, inv = coll.lists(i)
                with lock:                       # (the row number of a sample must be known to check it later; the engine calls
                    pass                         #  themselves run unlocked and race with the other threads)
                row = eng.insert_query(p, o, inv, 20)[0] if fused else eng.append(p, o, inv)
                with lock:
                    row_of[i] = row
        except Exception as ex:      # pragma: no cover
            errors.append(ex)

    def remover():
        rng = np.random.default_rng(1)
        try:
            for r in rng.choice(base, size=60, replace=False):
                eng.remove(int(r))
                with lock:
                    removed.add(int(r))
        except Exception as ex:      # pragma: no cover
            errors.append(ex)

    def querier():
        rng = np.random.default_rng(2)
        try:
            while not stop.is_set():
                q = int(rng.integers(0, base))
                try:
                    hits = eng.one_vs_all(q, 20 if rng.random() < 0.5 else 250)
                except KeyError:                 # the query row was removed meanwhile: the reference raises KeyError too
                    continue
                for h in hits:
                    assert int(h["row"]) != q and int(h["dist"]) <= 250
        except Exception as ex:      # pragma: no cover
            errors.append(ex)

    threads = [threading.Thread(target=inserter, args=(base, base + 100, True)),
               threading.Thread(target=inserter, args=(base + 100, coll.n, False)),
               threading.Thread(target=remover), threading.Thread(target=querier)]
    for t in threads:
        t.start()
    for t in threads[:3]:
        t.join()
    stop.set()
    threads[3].join()
    assert not errors, errors
    # quiescence: the live rows, in engine row order, against the oracle
    sample_of = {row: i for i, row in row_of.items()}
    live_rows = sorted(r for r in sample_of if r not in removed)
    live = coll.subset(np.array([sample_of[r] for r in live_rows]))
    index_of = {r: k for k, r in enumerate(live_rows)}
    assert eng.stats()["n_live"] == len(live_rows) and eng.stats()["n_rows"] == coll.n
    for ceiling in (12, 20, 250):
        for r in live_rows[::3]:
            want = c_oracle.one_vs_all(live.pos, live.off, live.invalid, index_of[r], ceiling)
            got = {index_of[int(h["row"])]: (int(h["dist"]), int(h["n1"]), int(h["n2"]), int(h["nboth"]))
                   for h in eng.one_vs_all(r, ceiling)}
            assert got == want, (kind, ceiling, r)
    eng.close()


def test_msa_golden():
    """multi_sequence_alignment / mixPORE statistics (seqComparer.py:691-1128) against the reference's outputs."""
    import math
    for case in golden("msa.json"):
        sc = _sc(case["ref"], case["maxNs"], 10)
        n = 0
        for _ in range(case["reps"]):
            for o in case["originals"]:
                n += 1
                sc.persist(sc.compress(o), guid="{0}-{1}".format(o, n))
        res = sc.multi_sequence_alignment(case["select"], output="dict", **case["kwargs"])
        want = case["got"]
        if want is None:
            assert res is None
            continue
        assert set(res.keys()) == set(want.keys()), case["label"]
        for k, v in want.items():
            got = res[k]
            if isinstance(v, dict):
                assert set(got.keys()) == set(v.keys()), (case["label"], k)
                for g, x in v.items():
                    y = got[g]
                    if isinstance(x, float) and y is not None:
                        assert math.isclose(float(y), x, rel_tol=1e-12, abs_tol=1e-15) or (math.isnan(x) and math.isnan(float(y))), (case["label"], k, g, x, y)
                    elif isinstance(x, list):
                        assert list(y) == x, (case["label"], k, g)
                    else:
                        assert (y is None and x is None) or y == x, (case["label"], k, g, x, y)
            else:
                assert list(got) == v, (case["label"], k)
        df = sc.multi_sequence_alignment(case["select"], output="df", **case["kwargs"])
        expected_cols = set(['what_tested', 'aligned_seq', 'aligned_seq_len', 'allN', 'alignN', 'allM', 'alignM', 'allN_or_M',
                             'alignN_or_M', 'p_value1', 'p_value2', 'p_value3', 'p_value4', 'observed_proportion',
                             'expected_proportion1', 'expected_proportion2', 'expected_proportion3', 'expected_proportion4'])
        assert set(df.columns.values) == expected_cols
        assert len(df.index) == len(want["guid2msa_seq"])


@pytest.mark.parametrize("slots", [8, None])
def test_all_vs_all_batched_tiles_vs_one_vs_all(slots, monkeypatch):
    """all-vs-all queues many query tiles between host synchronisations and grows its edge buffer on overflow:
    on a collection with more than FN3_SLOT_RING tiles (8 queries per tile -> 60 tiles) and with the default tile width
    its edges must equal the union of one-vs-all results."""
    from findneighbour3_b200.engine import Engine
    if slots:
        monkeypatch.setenv("FN3_SLOTS", str(slots))
    coll = _synthetic(12, 40, 300000, 71, k1=200, s=2.0, n_blocks=6, n_mean=1500, m_mean=2, rule="any", p_invalid=0.02)
    eng = Engine(coll.genome_length, 0)
    eng.append_bulk(coll.pos, coll.off, coll.invalid)
    assert coll.n > 32 * 8                                   # several batches of tiles
    for ceiling in (12, 400):
        edges = eng.all_vs_all(ceiling, upper_only=True)
        got = {(int(e["row1"]), int(e["row2"])): (int(e["dist"]), int(e["n1"]), int(e["n2"]), int(e["nboth"])) for e in edges}
        assert len(got) == len(edges)                        # no duplicates
        want = {}
        for q in range(coll.n):
            for r, v in hits_to_dict(eng.one_vs_all(q, ceiling)).items():
                if q < r:
                    want[(q, r)] = v
        assert got == want
    eng.close()


def test_config1_incremental_through_the_class_full_genome():
    """BASELINE.json configs[0] flow through the drop-in class on a full-length (4.4 Mb) genome: sequence STRINGS ->
    device compress -> persist -> mcompare after every insert, every neighbour list checked against the C oracle
    (reduced to 120 samples so the test stays within seconds; bench.py / test_properties_at_scale cover the sizes)."""
    from findneighbour3_b200 import synth
    from findneighbour3_b200.seqComparer import seqComparer
    from oracle import c_oracle
    L = synth.TB_LENGTH
    mask = synth.synthetic_mask(L, synth.TB_EXCLUDED, synth.TB_EXCLUDED_RUNS)
    coll = synth.make_collection(4, 30, seed=1, excluded=mask, m_mean=3, rule="any")
    ref_codes = synth.random_reference(L)
    ref = synth.reference_string(ref_codes)
    excluded = set(np.flatnonzero(mask).tolist())
    sc = seqComparer(reference=ref, maxNs=130000, snpCeiling=20, excludePositions=excluded)
    for i in range(coll.n):
        p, o, inv = coll.lists(i)
        if inv:
            seq = "N" * L                                  # an all-N sample: invalid under maxNs
        else:
            seq = synth.to_string(ref_codes, p, o)
        obj = sc.compress(seq)
        assert obj["invalid"] == (1 if inv else 0)
        if not inv:                                        # compress must give back exactly the generated lists
            got = np.concatenate([np.sort(np.fromiter(obj[k], np.int64, len(obj[k]))) for k in "ACGTN"] +
                                 [np.sort(np.fromiter(obj["M"].keys(), np.int64, len(obj["M"])))])
            assert np.array_equal(got, p.astype(np.int64))
        sc.persist(obj, "s%d" % i)
        if i % 4 == 3 or i == coll.n - 1:
            res = sc.mcompare("s%d" % i)
            assert len(res) == i
            got = {int(r[1][1:]): tuple(r[2:6]) for r in res if r[2] is not None}
            sub = coll.subset(np.arange(i + 1))
            assert got == c_oracle.one_vs_all(sub.pos, sub.off, sub.invalid, i, 20)
    assert sc.summarise_stored_items()["server|scstat|nSeqs"] == coll.n


# ---------------------------------------------------------------------------------------------
# (c) cluster statistics on the device: consensus vote, MSA sites / alignment / N-M tallies (csrc/fn3_cluster.cuh)
# ---------------------------------------------------------------------------------------------
def _cluster_class():
    """the collection of collection.json persisted into the drop-in class under the guids cluster.json uses"""
    from findneighbour3_b200 import synth
    c = golden("collection.json")
    pos, off, invalid = np.array(c["pos"], np.int32), np.array(c["off"], np.int64), np.array(c["invalid"], np.uint8)
    L = c["genome_length"]
    ref = synth.reference_string(synth.random_reference(L, c["ref_seed"]))
    coll = synth.Collection(L, pos, off, invalid, np.zeros(invalid.size, np.int32))
    excluded = set(np.flatnonzero(synth.synthetic_mask(L, 2000, 12, seed=3)).tolist())
    sc = _sc(ref, 700, 20, excluded)
    for i in range(coll.n):
        sc.persist(coll.as_dict(i), "s%d" % i)
    return sc, coll


def test_cluster_golden_through_class():
    """consensus(), multi_sequence_alignment(), estimate_expected_unk[_sites]() of the class — whose votes, variant
    sites, aligned characters and tallies come from the device — against the unmodified reference's outputs."""
    g = golden("cluster.json")
    sc, coll = _cluster_class()
    launches0 = sc.engine.launch_count()
    for case in g["consensus"]:
        objs = [sc.seqProfile["s%d" % i] for i in case["members"]]
        got = sc.consensus(objs, case["cutoff"])
        assert {k: sorted(v) for k, v in got.items()} == case["got"], (case["members"], case["cutoff"])
        stored = sc._consensus_of_stored(["s%d" % i for i in case["members"]], case["cutoff"])
        assert stored == got
    for case in g["msa"]:
        want = case["got"]
        res = sc.multi_sequence_alignment(case["select"], output="dict", sample_size=case["sample_size"],
                                          uncertain_base_type=case["unk"])
        assert list(res["variant_positions"]) == want["variant_positions"]
        assert sorted(res["invalid_guids"]) == sorted(want["invalid_guids"])
        for key in ("guid2msa_seq", "guid2allN", "guid2allM", "guid2allN_or_M", "guid2alignN", "guid2alignM",
                    "guid2alignN_or_M", "guid2expected_p1", "guid2expected_p2"):
            assert res[key] == want[key], (case["select"][:3], key)
        for s, text in res["guid2msa_seq"].items():
            assert res["guid2sequence"][s] == list(text)
    for case in g["unk_sites"]:
        got = sc.estimate_expected_unk_sites(sample_size=case["sample_size"], sites=set(case["sites"]), unk_type=case["unk"])
        assert got == case["got"]
    for unk, want in g["unk"].items():
        assert sc.estimate_expected_unk(sample_size=g["n_valid"], unk_type=unk) == want
    assert sc.estimate_expected_unk(sample_size=g["n_valid"] + 1, unk_type="N") is None
    assert sc.engine.launch_count() > launches0          # the statistics above were computed by kernels


@pytest.mark.parametrize("shape", ["high_nm", "tiny_genome", "tb_like", "tb_full_genome"])
def test_cluster_engine_vs_port(shape):
    """fn3_consensus / fn3_consensus_lists / fn3_msa_sites / fn3_msa_align through the C-ABI on seeded synthetic
    collections against the Python-set restatement (oracle/cluster_port.py, pinned by cluster.json)."""
    from findneighbour3_b200 import synth
    from findneighbour3_b200.engine import Engine
    from oracle import cluster_port as cp
    if shape == "high_nm":
        coll = _synthetic(5, 14, 120000, 31, k1=800, s=5.0, n_blocks=30, n_mean=9000, m_mean=400, rule="any",
                          max_ns=40000, p_invalid=0.05)
    elif shape == "tiny_genome":
        coll = _synthetic(5, 8, 31, 9, k1=3, s=1.0, n_blocks=2, n_mean=4, m_mean=2, rule="any", p_invalid=0.1)
    elif shape == "tb_full_genome":                          # L = 4 411 532: 88 MB of vote planes, 1 078 selection CTAs
        coll = synth.make_collection(3, 20, seed=12, m_mean=4, rule="any",
                                     excluded=synth.synthetic_mask(synth.TB_LENGTH, synth.TB_EXCLUDED, synth.TB_EXCLUDED_RUNS))
    else:
        coll = _synthetic(6, 20, 600000, 8, k1=500, s=3.0, n_blocks=20, n_mean=4000, m_mean=3, rule="any")
    L = coll.genome_length
    ref = synth.reference_string(synth.random_reference(L, 77))
    dicts = [coll.as_dict(i) for i in range(coll.n)]
    valid = [i for i in range(coll.n) if not coll.invalid[i]]
    eng = Engine(L, 0)
    eng.append_bulk(coll.pos, coll.off, coll.invalid)
    with pytest.raises(ValueError):
        eng.msa_sites(valid[:3])                             # the site rule needs the reference bases
    # NB a reference other than the one the collection was generated against: stored bases may EQUAL the reference
    # base here, which the reference's rule (a set of bases per position, :916-931) treats as one base
    eng.set_reference(ref)
    rng = np.random.default_rng(5)
    groups = [valid[:9], list(rng.choice(valid, size=min(len(valid), 17), replace=False)), valid[:2] + valid[:2], [valid[-1]],
              valid]
    # --- consensus: stored rows and shipped lists give the reference's vote
    for gsel in groups:
        for min_votes in sorted({1, 2, max(1, len(gsel) // 2), max(1, (4 * len(gsel) + 4) // 5), len(gsel), len(gsel) + 1}):
            want = {}
            for item in ("A", "C", "G", "T", "N"):
                votes = {}
                for i in gsel:
                    for p in dicts[i][item]:
                        votes[p] = votes.get(p, 0) + 1
                want[item] = sorted(p for p, v in votes.items() if v >= min_votes)
            pos, off = eng.consensus(gsel, min_votes)
            got = {k: pos[off[b]:off[b + 1]].tolist() for b, k in enumerate("ACGTN")}
            assert got == want, (shape, len(gsel), min_votes)
            sub = coll.subset(gsel)
            pos2, off2 = eng.consensus_lists(sub.pos, sub.off, min_votes)
            assert np.array_equal(pos, pos2) and np.array_equal(off, off2)
        for cutoff in (0.5, 0.7, 0.8):                       # the float threshold of the reference, via the port
            from findneighbour3_b200.seqComparer import seqComparer
            want = cp.consensus([dicts[i] for i in gsel], cutoff)
            pos, off = eng.consensus(gsel, seqComparer._min_votes(len(gsel), cutoff))
            assert seqComparer._consensus_dict(pos, off) == want
    # an invalid row votes for nothing
    bad = [i for i in range(coll.n) if coll.invalid[i]]
    if bad:
        pos, off = eng.consensus(valid[:4] + bad[:1], 1)
        pos2, off2 = eng.consensus(valid[:4], 1)
        assert np.array_equal(pos, pos2) and np.array_equal(off, off2)
        with pytest.raises(ValueError):
            eng.msa_sites(valid[:3] + bad[:1])
    # --- MSA: sites, aligned characters, tallies
    for gsel in groups:
        samples = [dicts[i] for i in gsel]
        ordered = cp.variant_positions(ref, samples)
        sites = eng.msa_sites(gsel)
        assert sites.tolist() == ordered, (shape, len(gsel))
        # the older rule of seqComparer_mt._msa (an N also accounts for a position): fn3_msa_sites_rule
        assert eng.msa_sites(gsel, n_is_call=True).tolist() == cp.variant_positions(ref, samples, n_is_call=True)
        codes, counts = eng.msa_align(gsel, sites, want_codes=True)
        lut = np.frombuffer(b"?ACGTNM", dtype=np.uint8)
        ref_at = np.frombuffer(ref.encode(), dtype=np.uint8)[sites]
        chars = np.where(codes == 0, ref_at[None, :], lut[codes])
        for j, i in enumerate(gsel):
            aligned = cp.aligned_string(ref, dicts[i], ordered)
            assert chars[j].tobytes().decode() == aligned
            assert tuple(int(x) for x in counts[j]) == cp.tallies(dicts[i], aligned)
        # counts-only form at foreign sites (estimate_expected_unk_sites) and with no sites at all
        foreign = np.unique(rng.integers(0, L, size=min(L, 300))).astype(np.int32)
        for at in (foreign, np.empty(0, np.int32)):
            none, c2 = eng.msa_align(gsel, at, want_codes=False)
            assert none is None
            sset = set(at.tolist())
            for j, i in enumerate(gsel):
                assert tuple(int(x) for x in c2[j]) == (len(dicts[i]["N"]), len(dicts[i]["M"]),
                                                        cp.unknowns_at_sites(dicts[i], sset, "N"),
                                                        cp.unknowns_at_sites(dicts[i], sset, "M"))
    # removed rows are refused
    eng.remove(valid[0])
    with pytest.raises(KeyError):
        eng.msa_sites(valid[:3])
    with pytest.raises(KeyError):
        eng.consensus(valid[:3], 1)
    with pytest.raises(ValueError):
        eng.msa_align(valid[1:3], np.array([5, 5], np.int32))        # sites must be strictly increasing
    eng.close()


# ---------------------------------------------------------------------------------------------
# (d) BASELINE.json configurations at their full sizes
# ---------------------------------------------------------------------------------------------
def test_config2_full_size_50k_sampled_queries_vs_oracle():
    """BASELINE.json configs[1] at full size: 50 000 TB-shaped samples (L = 4 411 532) resident on one GPU.
    Sampled stored queries, fresh inserts (fn3_insert_query) and un-stored queries (fn3_lists_vs_all) are checked
    bit-exact against the C oracle over ALL 50 000 candidates; the all-vs-all edges of sampled block rows at 12 SNV
    must equal the one-vs-all results (checksum over the whole matrix is compared between two tile widths in bench)."""
    from findneighbour3_b200 import synth
    from findneighbour3_b200.engine import Engine
    from oracle import c_oracle
    mask = synth.synthetic_mask(synth.TB_LENGTH, synth.TB_EXCLUDED, synth.TB_EXCLUDED_RUNS)
    coll = synth.make_collection(1000, 50, seed=2, excluded=mask)
    assert coll.n == 50000 and coll.genome_length == 4411532
    eng = Engine(coll.genome_length, 0)
    eng.append_bulk(coll.pos, coll.off, coll.invalid)
    rng = np.random.default_rng(3)
    nthreads = max(1, min(16, os.cpu_count() or 1))
    queries = [int(q) for q in rng.choice(coll.n, size=12, replace=False)]
    for ceiling in (20, 12, 250):
        for q in queries[: (12 if ceiling == 20 else 4)]:
            want = c_oracle.one_vs_all(coll.pos, coll.off, coll.invalid, q, ceiling, nthreads=nthreads)
            assert hits_to_dict(eng.one_vs_all(q, ceiling)) == want, (ceiling, q)
            if not coll.invalid[q]:
                p, o, inv = coll.lists(q)
                got = hits_to_dict(eng.lists_vs_all(p, o, inv, ceiling))
                assert got.pop(q)[0] == 0                    # the un-stored copy of q finds q itself at distance 0
                assert got == want
    # block rows of the all-vs-all matrix: rows r with r % 4000 == 7
    edges = eng.all_vs_all(12, upper_only=False, stride=4000, phase=7)
    got = {}
    for e in edges:
        got.setdefault(int(e["row1"]), {})[int(e["row2"])] = (int(e["dist"]), int(e["n1"]), int(e["n2"]), int(e["nboth"]))
    for r in range(7, coll.n, 4000):
        want = c_oracle.one_vs_all(coll.pos, coll.off, coll.invalid, r, 12, nthreads=nthreads)
        assert got.get(r, {}) == want, r
    # streaming inserts on top of the full store (configs[4] flow on one GPU): relatives of stored samples
    fresh = synth.make_collection(3, 4, seed=9001, excluded=mask)
    both = synth.concat(coll, fresh)                         # the oracle sees the first row+1 samples of this
    for i in range(fresh.n):
        p, o, inv = fresh.lists(i)
        row, hits = eng.insert_query(p, o, inv, 20)
        assert row == coll.n + i
        want = c_oracle.one_vs_all(both.pos, both.off[: 6 * (row + 1) + 1], both.invalid[: row + 1], row, 20,
                                   nthreads=nthreads)
        assert hits_to_dict(hits) == want
    eng.close()


@pytest.mark.parametrize("organism", ["salm", "ngon"])
def test_config4_high_n_and_mixed_collections(organism):
    """BASELINE.json configs[3] at the size SURVEY 8(d) names: 5 000 samples of an S. enterica-shaped (L = 4 857 432,
    SALM-exclude-sized mask) and of an N. gonorrhoeae-shaped (L = 2 232 025, no mask) collection with tens of thousands of N
    and hundreds of IUPAC positions per sample, some invalid through maxNs = 130 000, plus 2 % MIXED samples built as
    src/make_simulation.py:147-164 (mix_sequences) builds them — mixture-aware mcompare counts (dist, n1, n2, nboth) vs the
    C oracle, with mixtures among the queries."""
    from findneighbour3_b200 import synth
    from findneighbour3_b200.engine import Engine
    from oracle import c_oracle
    workers = max(1, min(16, os.cpu_count() or 1))
    if organism == "salm":
        L = 4857432
        coll = synth.make_families([4_000_000 + a for a in range(200)], 25, workers=workers, tb_mask=False,
                                   mask_spec=(L, 51897, 58, 44), genome_length=L, k1=5000, s=5.0, n_blocks=60, n_mean=30000,
                                   m_mean=300, rule="any", n_lognormal_sigma=0.9, max_ns=130000, ref_seed=4)
    else:
        L = 2232025
        coll = synth.make_families([5_000_000 + a for a in range(200)], 25, workers=workers, tb_mask=False, genome_length=L,
                                   k1=3000, s=5.0, n_blocks=50, n_mean=20000, m_mean=1000, rule="any", n_lognormal_sigma=0.9,
                                   max_ns=130000, p_invalid=0.02, ref_seed=5)
    assert coll.n == 5000
    coll = synth.add_mixtures(coll, 0.02, seed=5, max_ns=130000)
    assert coll.n == 5100
    eng = Engine(L, 0)
    eng.append_bulk(coll.pos, coll.off, coll.invalid)
    rng = np.random.default_rng(8)
    n_checked = 0
    mixed_valid = [q for q in range(5000, 5100) if not coll.invalid[q]]
    queries = [int(q) for q in rng.choice(5000, size=10, replace=False)] + mixed_valid[:6]
    for q in queries:
        for ceiling in (20, 100000):             # 100 000: every valid pair reports its counts (mixtures are far from everything)
            want = c_oracle.one_vs_all(coll.pos, coll.off, coll.invalid, q, ceiling, nthreads=workers)
            assert hits_to_dict(eng.one_vs_all(q, ceiling)) == want, (organism, q, ceiling)
            n_checked += len(want)
    assert n_checked > 0 and len(mixed_valid) > 0
    eng.close()


def test_config0_full_thousand_incremental_inserts_every_insert_checked():
    """BASELINE.json configs[0] as SURVEY 8(d) states it: 1 000 samples (20 families x 50, TB-length genome, TB-exclude-shaped
    mask, 1 % invalid) inserted one by one through the drop-in class, mcompare after EVERY insert, every row of every result —
    neighbours and the None rows — checked against the C oracle: 499 500 pairs."""
    from findneighbour3_b200 import synth
    from findneighbour3_b200.seqComparer import seqComparer
    from oracle import c_oracle
    L = synth.TB_LENGTH
    mask = synth.synthetic_mask(L, synth.TB_EXCLUDED, synth.TB_EXCLUDED_RUNS)
    coll = synth.make_families([1_000 + a for a in range(20)], 50, workers=max(1, min(16, os.cpu_count() or 1)), m_mean=2,
                               rule="any")
    assert coll.n == 1000
    ref = synth.reference_string(synth.random_reference(L))
    sc = seqComparer(reference=ref, maxNs=130000, snpCeiling=20, excludePositions=set(np.flatnonzero(mask).tolist()))
    pairs = neighbours = 0
    for i in range(coll.n):
        sc.persist(coll.as_dict(i), "s%d" % i)
        res = sc.mcompare("s%d" % i)
        assert len(res) == i
        got, none_rows = {}, 0
        for (g1, g2, d, n1, n2, nboth, a, b, c) in res:
            assert g1 == "s%d" % i
            if d is None:
                assert (n1, n2, nboth, a, b, c) == (None,) * 6
                none_rows += 1
            else:
                got[int(g2[1:])] = (d, n1, n2, nboth)
                assert (len(a), len(b), len(c)) == (n1, n2, nboth)         # the position sets of a surviving pair (:416-418)
        want = c_oracle.one_vs_all(coll.pos, coll.off, coll.invalid, i, 20, nthreads=4)
        want = {r: v for r, v in want.items() if r < i}
        assert got == want and none_rows == i - len(want), i
        pairs += i
        neighbours += len(want)
    assert pairs == 499500 and neighbours > 10000
    assert sc.summarise_stored_items()["server|scstat|nSeqs"] == 1000


def test_low_ceiling_filter_with_removed_and_invalid_rows():
    """ceilings below 16 take the two-candidates-per-thread filter that reads fingerprints only: removed and invalid
    rows (heads carrying the reserved always-mismatching id) never come back, in one-vs-all, all-vs-all and after a
    fused insert of an invalid sample; every result checked against the C oracle on the surviving store."""
    from findneighbour3_b200.engine import Engine
    from oracle import c_oracle
    coll = _synthetic(10, 30, 400000, 91, k1=300, s=2.0, n_blocks=8, n_mean=2000, m_mean=3, rule="any", p_invalid=0.05)
    eng = Engine(coll.genome_length, 0)
    eng.append_bulk(coll.pos, coll.off, coll.invalid)
    rng = np.random.default_rng(4)
    valid = np.flatnonzero(coll.invalid == 0)
    removed = set(int(r) for r in rng.choice(valid, size=25, replace=False))
    for r in removed:
        eng.remove(r)
    inv2 = coll.invalid.copy()
    inv2[list(removed)] = 1                                  # for the oracle a removed row is a row without distances
    row, hits = eng.insert_query(None, None, True, 12)       # an invalid sample: stored, no neighbours, never a candidate
    assert row == coll.n and len(hits) == 0
    for ceiling in (0, 5, 12, 15):
        for q in rng.choice(valid, size=8, replace=False):
            q = int(q)
            if q in removed:
                with pytest.raises(KeyError):
                    eng.one_vs_all(q, ceiling)
                continue
            want = c_oracle.one_vs_all(coll.pos, coll.off, inv2, q, ceiling, nthreads=4)
            assert hits_to_dict(eng.one_vs_all(q, ceiling)) == want, (ceiling, q)
    edges = eng.all_vs_all(12, upper_only=False)
    got = {}
    for e in edges:
        got.setdefault(int(e["row1"]), {})[int(e["row2"])] = (int(e["dist"]), int(e["n1"]), int(e["n2"]), int(e["nboth"]))
    assert not (set(got) & removed) and coll.n not in got
    for q in range(coll.n):
        if inv2[q]:
            assert q not in got
        else:
            assert got.get(q, {}) == c_oracle.one_vs_all(coll.pos, coll.off, inv2, q, 12, nthreads=4), q
    eng.close()


def test_persist_pickles_start_up_reload():
    """start-up reload from the reference's stored form (pickle protocol 2 per sample, mongoStore.py:344-363): natively
    parsed samples (valid, invalid), one in patch form (unpickled the ordinary way), lazily materialised host dicts;
    neighbours must equal the reference's (collection.json) and a plain persist() reload."""
    import pickle
    g = golden("collection.json")
    sc, coll = _cluster_class()                           # ordinary persist(), used as the twin
    want20 = {int(q): {r[0]: tuple(r[1:]) for r in rows} for q, rows in g["neighbours"]["20"].items()}
    # re-encode one family member as a patch in the twin, so that one stored object is NOT a plain compressed sequence
    q0 = max(want20, key=lambda k: len(want20[k]))
    sc.snpCompressionCeiling = 20
    sc.compress_relative_to_consensus("s%d" % q0, 0.8)
    kinds = [sc._stored_kind("s%d" % i) for i in range(coll.n)]
    assert "patch" in kinds and "invalid" in kinds and "compressed" in kinds
    guids = ["s%d" % i for i in range(coll.n)]
    blobs = [pickle.dumps(sc.seqProfile[gu], protocol=2) for gu in guids]
    sc2 = _sc(sc.reference, 700, 20, sc.excluded)
    sc2.consensi = dict(sc.consensi)
    sc2.persist_pickles(blobs, guids)
    assert sc2.engine.stats()["n_rows"] == coll.n
    assert sc2.summarise_stored_items() == sc.summarise_stored_items()
    lazy = sum(1 for v in sc2.seqProfile._d.values() if type(v).__name__ == "_LazyPickle")
    assert lazy == kinds.count("compressed") + kinds.count("invalid")
    for q in list(want20)[::7] + [q0]:
        got = {int(r[1][1:]): tuple(r[2:6]) for r in sc2.mcompare("s%d" % q) if r[2] is not None}
        assert got == want20[q], q
    assert sc2.seqProfile["s%d" % q0] == sc.seqProfile["s%d" % q0]
    some = next(i for i in range(coll.n) if kinds[i] == "compressed")
    assert sc2.seqProfile["s%d" % some] == coll.as_dict(some)          # materialised on demand, identical content
    n_before = sc2.engine.stats()["n_rows"]
    sc2.persist(coll.as_dict(some), "s%d" % some)                      # identical content: no new device row
    assert sc2.engine.stats()["n_rows"] == n_before
    res = sc2.multi_sequence_alignment(guids[:10], output="dict", sample_size=5)
    ref_res = sc.multi_sequence_alignment(guids[:10], output="dict", sample_size=5)
    assert res["variant_positions"] == ref_res["variant_positions"] and res["guid2msa_seq"] == ref_res["guid2msa_seq"]
