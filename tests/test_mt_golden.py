"""The drop-in for the reference's older threaded class (findneighbour3_b200.seqComparer_mt)
against tests/golden/mt.json — the UNMODIFIED src/seqComparer_mt.py run on the family-structured collection
(tests/golden/make_golden.py --mt).

Two runs of the same checks: on the CUDA engine (-m gpu, the parity test proper) and, on the CPU-only build box, with the
oracle-backed stand-in engine of tests/fake_engine.py, which pins the host layer only."""
import numpy as np
import pytest

from findneighbour3_b200 import synth
from util import golden


def _collection():
    c = golden("collection.json")
    L = c["genome_length"]
    ref = synth.reference_string(synth.random_reference(L, c["ref_seed"]))
    pos, off, invalid = np.array(c["pos"], np.int32), np.array(c["off"], np.int64), np.array(c["invalid"], np.uint8)
    excluded = set(np.flatnonzero(synth.synthetic_mask(L, 2000, 12, seed=3)).tolist())
    dicts = []
    for i in range(invalid.size):
        if invalid[i]:
            dicts.append({"invalid": 1})
            continue
        o = off[6 * i: 6 * i + 7]
        d = {"invalid": 0}
        for b, k in enumerate("ACGTN"):
            d[k] = set(int(x) for x in pos[o[b]:o[b + 1]])
        dicts.append(d)
    return ref, excluded, dicts


def _rows_of(res):
    return sorted(([r[1], r[2], r[3], r[4], r[5], None if r[6] is None else len(r[6]), None if r[7] is None else len(r[7]),
                    None if r[8] is None else sorted(int(x) for x in r[8])[:5]] for r in res), key=repr)


def _setlist(d):
    return {k: sorted(int(x) for x in v) for k, v in d.items()}


def _same(a, b):
    if a is None or b is None:
        return a is None and b is None
    if a != a or b != b:                       # binomtest on an empty alignment: nan in the reference as well
        return a != a and b != b
    return a == pytest.approx(b, rel=1e-12, abs=1e-300)


def _check_everything():
    from findneighbour3_b200.seqComparer_mt import seqComparer
    g = golden("mt.json")
    ref, excluded, dicts = _collection()
    stores = {}

    def store(cpu, ceiling, comp):
        key = (cpu, ceiling, comp)
        if key not in stores:
            sc = seqComparer(reference=ref, maxNs=700, snpCeiling=ceiling, snpCompressionCeiling=comp, cpuCount=cpu,
                             excludePositions=excluded)
            for i, d in enumerate(dicts):
                sc.persist(d, "s%d" % i)
            stores[key] = sc
        return stores[key]

    # both mcompare conventions (lists without the query at snpCompressionCeiling / tuples with the query at snpCeiling)
    for case in g["mcompare"]:
        sc = store(case["cpuCount"], case["snpCeiling"], case["snpCompressionCeiling"])
        res = sc.mcompare(case["query"], case.get("guids"))
        assert sorted(set(type(r).__name__ for r in res)) == case["row_type"]
        assert _rows_of(res) == sorted(case["rows"], key=repr), (case["cpuCount"], case["snpCeiling"], case["query"])

    sc = store(1, 20, 250)
    assert sc.summarise_stored_items() == g["summary"]
    n_valid = sum(1 for d in dicts if d["invalid"] == 0)

    # the three-test MSA: sites under the "N accounts for a position" rule, characters, tallies, p values
    for case in g["msa"]:
        res = sc.multi_sequence_alignment(case["select"], output="dict", sample_size=case["sample_size"])
        want = case["got"]
        assert set(res.keys()) - {"guid2sequence"} == set(want.keys())
        for k in ("variant_positions", "invalid_guids"):
            assert list(res[k]) == want[k], k
        for k in ("guid2msa_seq", "guid2allN", "guid2alignN"):
            assert res[k] == want[k], k
        for guid, text in res["guid2msa_seq"].items():
            assert res["guid2sequence"][guid] == list(text)
        for k in ("guid2observed_proportion", "guid2expected_p1", "guid2expected_p2", "guid2expected_p3", "guid2pvalue1",
                  "guid2pvalue2", "guid2pvalue3"):
            assert set(res[k]) == set(want[k])
            for guid in want[k]:
                assert _same(res[k][guid], want[k][guid]), (k, guid, res[k][guid], want[k][guid])
        df = sc.multi_sequence_alignment(case["select"], output="df", sample_size=case["sample_size"])
        assert list(df.columns) == case["columns"]
        for guid, v in case["p_value3_column"].items():
            got = df.loc[guid, "p_value3"]
            assert _same(None if got is None or got != got else float(got), v)
        as_dict = sc.multi_sequence_alignment(case["select"], output="df_dict", sample_size=case["sample_size"])
        assert set(as_dict.keys()) == set(want["guid2msa_seq"].keys())
    assert float(sc.estimate_expected_N(sample_size=n_valid)) == g["expected_N"]
    assert sc.estimate_expected_N(sample_size=n_valid + 1) is g["expected_N_too_few"]
    assert float(sc.estimate_expected_N_sites(sample_size=n_valid, sites=set(g["msa"][0]["got"]["variant_positions"]))) == \
        g["expected_N_sites"]

    # consensus vote, {'add','subtract'} patches, hash
    valid = [i for i, d in enumerate(dicts) if d["invalid"] == 0]
    for case in g["consensus"]:
        members = [sc.seqProfile["s%d" % i] for i in case["members"]]
        cons = sc.consensus(members, case["cutoff"])
        assert _setlist(cons) == case["got"]
        assert sc.compressed_sequence_hash(cons) == case["hash"]
        patch = sc.generate_patch(members[0], cons)
        assert {k: _setlist(v) for k, v in patch.items()} == case["patch_of_first"]
        assert sc.apply_patch(patch, cons) == members[0]

    # recompression against the neighbours' consensus; distances unchanged afterwards
    rc = g["recompress"]
    visited = sc.compress_relative_to_consensus(rc["seed"])
    assert sorted(visited) == rc["visited"]
    assert len(sc.consensi) == rc["n_consensi"]
    assert sc.summarise_stored_items() == rc["summary"]
    for guid, patch in rc["patches"].items():
        assert {k: _setlist(v) for k, v in sc.seqProfile[guid]["patch"].items()} == patch
    assert _rows_of(sc.mcompare(rc["seed"])) == sorted(rc["neighbours_after"], key=repr)
    assert sc.uncompress(sc.seqProfile[rc["seed"]]) == sc.uncompress(dicts[valid[0]])

    # compress: '-' as N, invalid above maxNs, KeyError on anything outside ACGTN-
    small = seqComparer(reference="ACTGACTG", maxNs=3, snpCeiling=10, excludePositions={7})
    for case in g["compress"]:
        if "raises" in case:
            with pytest.raises(KeyError) as ei:
                small.compress(case["seq"])
            assert ei.value.args[0] == case["arg"]
            continue
        got = small.compress(case["seq"])
        if case["got"]["invalid"] == 1:
            assert got == {"invalid": 1}
        else:
            assert set(got.keys()) == {"A", "C", "G", "T", "N", "invalid"}
            assert dict(_setlist({k: got[k] for k in "ACGTN"}), invalid=0) == case["got"]
            assert small.uncompress(got) == case["roundtrip"]

    # assess_mixed: one row of this_guid per pair of related samples
    sc2 = store(1, 12, 40)
    related = ["s%d" % i for i in valid[1:5]]
    res = sc2.assess_mixed("s%d" % valid[0], related, max_sample_size=10)
    assert len(res.index) == 6 and list(res["pairid"]) == [1, 2, 3, 4, 5, 6]
    assert list(res.columns) == g["msa"][0]["columns"] + ["pairid"]
    assert sc2.assess_mixed("s%d" % valid[0], related[:1]) is None
    for sc in stores.values():
        sc.engine.close()


@pytest.mark.gpu
def test_mt_class_matches_the_reference_on_the_gpu():
    _check_everything()


def test_mt_class_host_layer_with_the_stand_in_engine(monkeypatch):
    import findneighbour3_b200.seqComparer as base
    from fake_engine import OracleEngine
    monkeypatch.setattr(base, "make_engine", lambda genome_length, device=0: OracleEngine(genome_length))
    _check_everything()
