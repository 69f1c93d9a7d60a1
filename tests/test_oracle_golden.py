"""not-gpu: the oracle (Python-set port AND C restatement) against the golden vectors produced by the
UNMODIFIED reference (tests/golden/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

from oracle import c_oracle
from oracle import seqcomparer_port as port
from util import dict_from_json, golden, lists_from_dict


def _comp(ref, maxNs, seq, excluded=()):
    included = set(range(len(ref))) - set(excluded)
    return port.compress(ref, included, maxNs, seq)


def test_compress_matches_reference():
    for case in golden("compress.json"):
        got = _comp(case["ref"], case["maxNs"], case["seq"], case["excluded"])
        assert got == dict_from_json(case["got"]), case


def test_kat_count_differences_port_and_c():
    n = 0
    for case in golden("kat.json"):
        if case["kind"] != "countDifferences":
            continue
        a = _comp(case["ref"], case["maxNs"], case["seq1"])
        b = _comp(case["ref"], case["maxNs"], case["seq2"])
        want = case["got"]
        assert want == case["literal"]
        assert port.count_differences(a, b, case["snpCeiling"]) == want, case
        p1, o1, i1 = lists_from_dict(a)
        p2, o2, i2 = lists_from_dict(b)
        res = c_oracle.pair(p1, o1, i1, p2, o2, i2, case["snpCeiling"])
        assert (None if res is None else res[0]) == want, case
        n += 1
    assert n >= 20


def test_kat_mcompare_and_bykey():
    for case in golden("kat.json"):
        if case["kind"] == "mcompare":
            store = {s: _comp(case["ref"], case["maxNs"], s) for s in case["samples"]}
            got = {r[1]: r[2:6] for r in port.mcompare(store, {}, case["query"], case["snpCeiling"])}
            assert got == case["got"] == case["literal"]
        elif case["kind"] == "mcompare_dist":
            store = {s: _comp(case["ref"], case["maxNs"], s) for s in case["samples"]}
            got = {r[1]: r[2] for r in port.mcompare(store, {}, case["query"], case["snpCeiling"])}
            assert got == case["got"] == case["literal"]
        elif case["kind"] == "byKey":
            store = {"k1": _comp(case["ref"], case["maxNs"], case["seq1"]),
                     "k2": _comp(case["ref"], case["maxNs"], case["seq2"])}
            got = port.by_key(store, {}, "k1", "k2", case["snpCeiling"])
            assert list(got[2:6]) == case["got"]
            assert got[2] == case["literal"]


def _random_pairs_store():
    g = golden("random_pairs.json")
    included = set(range(len(g["ref"]))) - set(g["excluded"])
    store = {i: port.compress(g["ref"], included, g["maxNs"], s) for i, s in enumerate(g["seqs"])}
    return g, store


@pytest.mark.parametrize("ceiling", [20, 1000])
def test_random_pairs_port(ceiling):
    g, store = _random_pairs_store()
    rows = []
    for i in store:
        for r in port.mcompare(store, {}, i, ceiling):
            if r[2] is not None:
                rows.append([r[0], r[1], r[2], r[3], r[4], r[5]])
    assert sorted(rows) == g["results"][str(ceiling)]


@pytest.mark.parametrize("ceiling", [20, 1000])
def test_random_pairs_c_oracle(ceiling):
    g, store = _random_pairs_store()
    packed = [lists_from_dict(store[i]) for i in range(len(store))]
    rows = []
    for i, (p1, o1, i1) in enumerate(packed):
        for j, (p2, o2, i2) in enumerate(packed):
            if i == j:
                continue
            res = c_oracle.pair(p1, o1, i1, p2, o2, i2, ceiling)
            if res is not None:
                rows.append([i, j] + list(res))
    assert sorted(rows) == g["results"][str(ceiling)]


@pytest.mark.parametrize("ceiling", [12, 20, 250])
@pytest.mark.parametrize("threads", [1, 3])
def test_collection_c_oracle_one_vs_all(ceiling, threads):
    g = golden("collection.json")
    pos, off, inv = np.array(g["pos"], np.int32), np.array(g["off"], np.int64), np.array(g["invalid"], np.uint8)
    n = inv.size
    for q in range(n):
        got = c_oracle.one_vs_all(pos, off, inv, q, ceiling, nthreads=threads)
        want = {r[0]: tuple(r[1:]) for r in g["neighbours"][str(ceiling)][str(q)]}
        assert got == want, (q, ceiling)


def test_collection_port_sampled():
    from findneighbour3_b200.synth import Collection
    g = golden("collection.json")
    coll = Collection(g["genome_length"], np.array(g["pos"], np.int32), np.array(g["off"], np.int64),
                      np.array(g["invalid"], np.uint8), np.zeros(len(g["invalid"]), np.int32))
    store = {i: coll.as_dict(i) for i in range(coll.n)}
    for q in (0, 5, 17, 40, 95):
        got = port.neighbours(store, {}, q, 20)
        want = {r[0]: tuple(r[1:]) for r in g["neighbours"]["20"][str(q)]}
        assert got == want


# ---------------------------------------------------------------------------------------------
# consensus vote / MSA set algebra (oracle/cluster_port.py) against the reference's outputs (cluster.json)
# ---------------------------------------------------------------------------------------------
def _cluster_store():
    from findneighbour3_b200 import synth
    c = golden("collection.json")
    pos, off, invalid = np.array(c["pos"], np.int32), np.array(c["off"], np.int64), np.array(c["invalid"], np.uint8)
    ref = synth.reference_string(synth.random_reference(c["genome_length"], c["ref_seed"]))
    store = {}
    for i in range(invalid.size):
        if invalid[i]:
            store["s%d" % i] = {"invalid": 1}
            continue
        o = off[6 * i: 6 * i + 7]
        d = {"invalid": 0}
        for b, k in enumerate("ACGTN"):
            d[k] = set(int(x) for x in pos[o[b]:o[b + 1]])
        d["M"] = {int(x): "R" for x in pos[o[5]:o[6]]}
        store["s%d" % i] = d
    return ref, store


def test_cluster_port_matches_reference():
    from oracle import cluster_port as cp
    g = golden("cluster.json")
    ref, store = _cluster_store()
    valid = [k for k, v in store.items() if v["invalid"] == 0]
    assert len(valid) == g["n_valid"]
    for case in g["consensus"]:
        got = cp.consensus([store["s%d" % i] for i in case["members"]], case["cutoff"])
        assert {k: sorted(v) for k, v in got.items()} == case["got"], (case["members"], case["cutoff"])
    for case in g["msa"]:
        want = case["got"]
        sel = [s for s in case["select"] if store[s]["invalid"] == 0]
        assert sorted(s for s in case["select"] if store[s]["invalid"] == 1) == sorted(want["invalid_guids"])
        samples = [store[s] for s in sel]
        ordered = cp.variant_positions(ref, samples)
        assert ordered == want["variant_positions"]
        for s in sel:
            aligned = cp.aligned_string(ref, store[s], ordered)
            assert aligned == want["guid2msa_seq"][s]
            assert cp.tallies(store[s], aligned) == (want["guid2allN"][s], want["guid2allM"][s], want["guid2alignN"][s],
                                                     want["guid2alignM"][s])
        e2 = cp.expected_unknowns([store[s] for s in valid if s not in want["invalid_guids"]], set(ordered), case["unk"])
        for s in sel:
            if len(ordered) == 0:
                assert want["guid2expected_p2"][s] is None
            else:
                assert want["guid2expected_p2"][s] == e2 / len(ordered)
    for case in g["unk_sites"]:
        assert cp.expected_unknowns([store[s] for s in valid], set(case["sites"]), case["unk"]) == case["got"]
    for unk, want in g["unk"].items():
        assert cp.expected_unknowns([store[s] for s in valid], None, unk) == want
