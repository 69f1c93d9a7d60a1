#!/usr/bin/env python3
"""Generates tests/golden/*.json by running the UNMODIFIED reference (/root/reference/src/seqComparer.py,
imported through oracle/ref_shim.py).  Runs only in the build container; the fixtures are committed.

    python tests/golden/make_golden.py

Files
  kat.json            the reference's own known-answer tests for this path, each case
                      carrying BOTH the literal expectation transcribed from the reference's test source
                      (`literal`, with its line) and the value the reference produced here (`got`)
  compress.json       compress() outputs (incl. '-', IUPAC, arbitrary chars, mask, invalid)
  random_pairs.json   all ordered pairs of 60 random samples (L=300, 10 % masked, N/-/IUPAC, some invalid):
                      reference mcompare rows (dist,n1,n2,nboth) at ceilings 20 and 1000
  collection.json     a family-structured synthetic collection in list form (findneighbour3_b200.synth) and
                      the reference's neighbour lists for every query at ceilings 12, 20 and 250
  consensus.json      consensus / patch / hash vectors (double delta, seqComparer.py:460-565)
  msa.json            multi_sequence_alignment outputs on the reference's own test families (tests 45a-47c)
  cluster.json        consensus votes, _msa sites / aligned strings / N-M tallies and estimate_expected_unk_sites on
                      the collection of collection.json (`--cluster` regenerates only this file)
  nucleic.json        NucleicAcid.examine (src/NucleicAcid.py, unmodified): composition dicts / exception types
                      (`--nucleic` regenerates only this file)
  mt.json             the older threaded class (src/seqComparer_mt.py, unmodified) on the same collection without its
                      M lists: compress, both mcompare conventions, consensus / patches / recompression, the
                      three-test MSA and estimate_expected_N[_sites] (`--mt` regenerates only this file)
"""
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from findneighbour3_b200 import synth  # noqa: E402

REF = ref_shim.load_reference_module("seqComparer")


def jsonable(obj):
    """compressed_sequence dict -> JSON-friendly (sets as sorted lists, M keys as strings)."""
    if obj.get("invalid", 0) == 1 and "A" not in obj:
        return {"invalid": 1}
    out = {"invalid": obj["invalid"]}
    for k in "ACGTN":
        out[k] = sorted(int(x) for x in obj[k])
    out["M"] = {str(int(k)): v for k, v in sorted(obj["M"].items())}
    return out


def dump(name, payload):
    path = os.path.join(HERE, name)
    with open(path, "w") as fh:
        json.dump(payload, fh, separators=(",", ":"), sort_keys=True)
    print("wrote %s (%d bytes)" % (path, os.path.getsize(path)))


# ---------------------------------------------------------------------------------------------
def make_kat():
    cases = []

    def cd(line, ref, maxNs, ceiling, s1, s2, literal, direct=False):
        """countDifferences known answers.  direct=True: the test assigns _seq1/_seq2 = compress(..)"""
        sc = REF.seqComparer(reference=ref, maxNs=maxNs, snpCeiling=ceiling)
        if direct:
            sc._seq1 = sc.compress(s1)
            sc._seq2 = sc.compress(s2)
        else:
            sc.setComparator1(s1)
            sc.setComparator2(s2)
        got = sc.countDifferences()
        assert got == literal, (line, got, literal)
        cases.append({"kind": "countDifferences", "src_line": line, "ref": ref, "maxNs": maxNs, "snpCeiling": ceiling,
                      "seq1": s1, "seq2": s2, "literal": literal, "got": got})

    R = "ACTG"
    # seqComparer.py tests 9..17 (snpCeiling 20 / 10)
    cd(1629, R, 1e8, 20, "ACTG", "ACTG", 0)
    cd(1636, R, 1e8, 20, "TTTG", "ACTG", 2)
    cd(1643, R, 1e8, 20, "TTTG", "NNTG", 0)
    cd(1650, R, 1e8, 20, "TTTG", "MMTG", 0)
    cd(1657, R, 1e8, 20, "NNTG", "TTTG", 0)
    cd(1665, R, 1e8, 20, "--TG", "TTTG", 0)
    cd(1672, R, 1e8, 20, "RRRR", "TTTG", 0)
    cd(1679, R, 1e8, 20, "TTTG", "RRRR", 0)
    cd(1686, R, 1e8, 20, "--AG", "TTAA", 1)
    cd(1693, R, 1e8, 20, "TTAA", "--AG", 1)
    cd(1701, R, 1e8, 10, "AAAA", "CCCC", 4, direct=True)
    cd(1713, R, 1e8, 10, "AAAA", "RRCC", 2, direct=True)
    cd(1725, R, 1e8, 10, "AAAA", "RRNN", 0, direct=True)
    cd(1737, R, 3, 10, "AAAA", "NNNN", None, direct=True)
    # ceiling semantics, snpCeiling = 1 (tests 31..36)
    cd(1825, R, 1e8, 1, "ACTG", "ACTG", 0)
    cd(1832, R, 1e8, 1, "TTTG", "ACTG", None)
    cd(1839, R, 1e8, 1, "TTTG", "NNTG", 0)
    cd(1846, R, 1e8, 1, "NNTG", "TTTG", 0)
    cd(1860, R, 1e8, 1, "--TG", "TTTG", 0)
    cd(1867, R, 1e8, 1, "--AG", "TTAA", 1)
    cd(1874, R, 1e8, 1, "TTAA", "--AG", 1)

    # test 44 (:2054-2087): countDifferences_byKey with maxNs 2 -> None, maxNs 1e8 -> 0
    for maxNs, literal in ((2, None), (1e8, 0)):
        sc = REF.seqComparer(reference=R, maxNs=maxNs, snpCeiling=10)
        sc.persist(sc.compress("AAAA"), "k1")
        sc.persist(sc.compress("NNNN"), "k2")
        got = sc.countDifferences_byKey(("k1", "k2"), cutoff=None)
        assert got[2] == literal
        cases.append({"kind": "byKey", "src_line": 2054, "ref": R, "maxNs": maxNs, "snpCeiling": 10,
                      "seq1": "AAAA", "seq2": "NNNN", "literal": literal, "got": list(got[2:6])})

    # SURVEY appendix A worked vector + mc1 (seqComparer_mt.py:2254-2292; N-only, valid for both classes)
    sc = REF.seqComparer(reference=R, maxNs=1e8, snpCeiling=20)
    for s in ("AAAA", "CCCC", "RRCC", "NNTG"):
        sc.persist(sc.compress(s), s)
    rows = {r[1]: r[2:6] for r in sc.mcompare("AAAA")}
    assert rows == {"CCCC": [4, 0, 0, 0], "NNTG": [2, 0, 2, 2], "RRCC": [2, 0, 0, 0]}, rows
    cases.append({"kind": "mcompare", "src_line": 149, "ref": R, "maxNs": 1e8, "snpCeiling": 20,
                  "samples": ["AAAA", "CCCC", "RRCC", "NNTG"], "query": "AAAA",
                  "literal": {"CCCC": [4, 0, 0, 0], "NNTG": [2, 0, 2, 2], "RRCC": [2, 0, 0, 0]}, "got": rows})

    ref12 = "GGGGGGGGGGGG"
    originals = ["AAACACTGACTG", "CCCCACTGACTG", "TTTCACTGACTG", "GGGCACTGACTG", "NNNCACTGACTG", "ACTCACTGACTG",
                 "TCTNACTGACTG"]
    sc = REF.seqComparer(reference=ref12, maxNs=1e8, snpCeiling=10)
    for s in originals:
        sc.persist(sc.compress(s), s)
    rows = {r[1]: r[2] for r in sc.mcompare("AAACACTGACTG")}
    literal = {"CCCCACTGACTG": 3, "TTTCACTGACTG": 3, "GGGCACTGACTG": 3, "NNNCACTGACTG": 0, "ACTCACTGACTG": 2,
               "TCTNACTGACTG": 3}
    assert rows == literal, rows
    cases.append({"kind": "mcompare_dist", "src_line": "seqComparer_mt.py:2254", "ref": ref12, "maxNs": 1e8,
                  "snpCeiling": 10, "samples": originals, "query": "AAACACTGACTG", "literal": literal, "got": rows})

    # excluded_hash known answer (tests 37/38 :1881-1897)
    sc = REF.seqComparer(reference=R, maxNs=1e8, snpCeiling=1)
    assert sc.excluded_hash() == "Excl 0 nt [d751713988987e9331980363e24189ce]"
    cases.append({"kind": "excluded_hash", "src_line": 1888, "ref": R, "excluded": [],
                  "literal": "Excl 0 nt [d751713988987e9331980363e24189ce]", "got": sc.excluded_hash()})
    sc = REF.seqComparer(reference="ACTGACTG", maxNs=1e8, snpCeiling=1, excludePositions=set([1, 5, 6]))
    cases.append({"kind": "excluded_hash", "src_line": 240, "ref": "ACTGACTG", "excluded": [1, 5, 6],
                  "literal": None, "got": sc.excluded_hash()})

    # _setStats (test 29 :1780-1817)
    ref31 = "ACTGTTAATTTTTTTTTGGGGGGGGGGGGAA"
    sc = REF.seqComparer(reference=ref31, maxNs=1e8, snpCeiling=20)
    for s1, s2, key in (("GGGGTTAANNNNNNNNNGGGGGAAAAGGGAA", "ACTGTTAATTTTTTTTTNNNNNNNNNNNNNN", "N"),
                        ("NNNGTTAANNNNNNNNTGGGGGAAAAGGGAA", "ACTNNNNNTTTTTTTTTQQQQQQQQQQQQQQ", "N"),
                        ("NNNGTTAANNNNNNNNTGGGGGAAAAGGGAA", "ACTNNNNNTTTTTTTTTQQQQQQQQQQQQQQ", "M"),
                        ("qqqGTTAAqqqqqqqqTGGGGGAAAAGGGAA", "ACTqqqqqTTTTTTTTTqqqqqqqqqqqqqq", "M")):
        c1, c2 = sc.compress(s1), sc.compress(s2)
        n1, n2, nall, rv1, rv2, rv = sc._setStats(c1[key], c2[key])
        cases.append({"kind": "setStats", "src_line": 1780, "ref": ref31, "seq1": s1, "seq2": s2, "key": key,
                      "got": [n1, n2, nall, sorted(rv)]})
    dump("kat.json", cases)


# ---------------------------------------------------------------------------------------------
def make_compress():
    cases = []

    def one(ref, seq, maxNs=1e8, excluded=()):
        sc = REF.seqComparer(reference=ref, maxNs=maxNs, snpCeiling=20, excludePositions=set(excluded))
        out = sc.compress(seq)
        case = {"ref": ref, "seq": seq, "maxNs": maxNs, "excluded": sorted(excluded), "got": jsonable(out)}
        if out.get("invalid") == 0:
            case["uncompress"] = sc.uncompress(out)
        cases.append(case)

    for s in ("ACTG", "ACTQ", "NYTQ", "ACTN", "ACT-", "TCT-", "ATT-", "AAAA", "CCCC", "TTTT", "GGGG", "NNNN", "ACTC",
              "TCTN", "QRST", "acgt", "----", "RYSW"):
        one("ACTG", s)
    one("ACTG", "NNNN", maxNs=3)
    one("ACTG", "NNNQ", maxNs=3)
    one("ACTG", "NNNG", maxNs=3)
    one("ACTGTTAATTTTTTTTTGGGGGGGGGGGGAA", "ACTGTTAANNNNNNNNTGGGGGGGGGGGGAA")
    one("ACTGTTAATTTTTTTTTGGGGGGGGGGGGAA", "NNTGTTAANNNNNNNNTGGGGGGGGGGGGAA")
    rnd = random.Random(11)
    for _ in range(12):
        L = rnd.randint(20, 80)
        ref = "".join(rnd.choice("ACGT") for _ in range(L))
        seq = "".join(rnd.choice("ACGT" * 3 + "N-RYKq") if rnd.random() < 0.4 else ref[i] for i in range(L))
        excl = [i for i in range(L) if rnd.random() < 0.15]
        one(ref, seq, maxNs=rnd.choice([5, 1e8]), excluded=excl)
    dump("compress.json", cases)


# ---------------------------------------------------------------------------------------------
def make_random_pairs():
    rnd = random.Random(20181101)
    L, n = 300, 60
    ref = "".join(rnd.choice("ACGT") for _ in range(L))
    excluded = sorted(rnd.sample(range(L), 30))
    seqs = []
    base = None
    for i in range(n):
        if i % 6 == 0 or base is None:
            base = list(ref)
            for _ in range(rnd.randint(0, 25)):
                base[rnd.randrange(L)] = rnd.choice("ACGT")
        s = list(base)
        for _ in range(rnd.randint(0, 6)):
            s[rnd.randrange(L)] = rnd.choice("ACGT")
        # uncertain calls: runs of N / '-' and scattered IUPAC
        for _ in range(rnd.randint(0, 4)):
            st, ln = rnd.randrange(L), rnd.randint(1, 25 if rnd.random() < 0.8 else 60)
            ch = rnd.choice("NN-")
            for p in range(st, min(L, st + ln)):
                s[p] = ch
        for _ in range(rnd.randint(0, 8)):
            s[rnd.randrange(L)] = rnd.choice("RYSWKMq")
        seqs.append("".join(s))
    out = {"ref": ref, "excluded": excluded, "maxNs": 60, "seqs": seqs, "results": {}}
    for ceiling in (20, 1000):
        sc = REF.seqComparer(reference=ref, maxNs=60, snpCeiling=ceiling, excludePositions=set(excluded))
        for i, s in enumerate(seqs):
            sc.persist(sc.compress(s), i)
        rows = []
        for i in range(n):
            for r in sc.mcompare(i):
                if r[2] is not None:
                    rows.append([r[0], r[1], r[2], r[3], r[4], r[5]])
        out["results"][str(ceiling)] = sorted(rows)
        out["n_invalid"] = sum(1 for i in range(n) if sc.seqProfile[i].get("invalid") == 1)
    print("random_pairs: invalid samples", out["n_invalid"], "rows@20", len(out["results"]["20"]), "rows@1000",
          len(out["results"]["1000"]))
    dump("random_pairs.json", out)


# ---------------------------------------------------------------------------------------------
def make_collection():
    L = 30000
    excluded = synth.synthetic_mask(L, 2000, 12, seed=3)
    coll = synth.make_collection(n_anc=8, max_c=12, genome_length=L, k1=60, s=3.0, n_blocks=5, n_mean=400, m_mean=6,
                                 p_invalid=0.04, seed=17, rule="any", ref_seed=5, excluded=excluded, max_ns=700)
    ref = synth.reference_string(synth.random_reference(L, 5))
    out = {"genome_length": L, "ref_seed": 5, "pos": coll.pos.tolist(), "off": coll.off.tolist(),
           "invalid": coll.invalid.tolist(), "neighbours": {}}
    for ceiling in (12, 20, 250):
        sc = REF.seqComparer(reference=ref, maxNs=700, snpCeiling=ceiling, excludePositions=set(np.flatnonzero(excluded).tolist()))
        for i in range(coll.n):
            sc.persist(coll.as_dict(i), i)
        res = {}
        for i in range(coll.n):
            rows = [[r[1], r[2], r[3], r[4], r[5]] for r in sc.mcompare(i) if r[2] is not None]
            res[str(i)] = sorted(rows)
        out["neighbours"][str(ceiling)] = res
    npairs = sum(len(v) for v in out["neighbours"]["20"].values())
    print("collection: n", coll.n, "invalid", int(coll.invalid.sum()), "pairs<=20:", npairs,
          "pairs<=250:", sum(len(v) for v in out["neighbours"]["250"].values()))
    dump("collection.json", out)


# ---------------------------------------------------------------------------------------------
def make_consensus():
    out = {}
    sc = REF.seqComparer(reference="ACTG", maxNs=1e8, snpCeiling=10)
    seqs = ["TTAA", "TTTA", "TTGA", "TTAA", "NNAA", "ACTG"]
    comp = [sc.compress(s) for s in seqs]
    out["consensus"] = []
    for cutoff in (0.5, 0.6, 0.8, 1.0):
        cons = sc.consensus(comp, cutoff)
        out["consensus"].append({"ref": "ACTG", "seqs": seqs, "cutoff": cutoff,
                                 "got": {k: sorted(v) for k, v in cons.items()},
                                 "hash": sc.compressed_sequence_hash(cons)})
    # test 40 (:1975-1986): hash of a compressed sequence
    c = sc.compress("TTAA")
    out["hash"] = {"ref": "ACTG", "seq": "TTAA", "literal": "6ce0e55c4ab092f560e03c5d2de53098",
                   "got": sc.compressed_sequence_hash(c)}
    assert out["hash"]["got"] == out["hash"]["literal"]
    # patch round trip (tests 41-43)
    cons = sc.consensus(comp, 0.6)
    out["patch"] = []
    for s, c in zip(seqs, comp):
        patch = sc.generate_patch(c, cons)
        back = sc.apply_patch(patch, cons)
        assert back == c
        out["patch"].append({"seq": s, "patch": {"+": {k: sorted(v) for k, v in patch["+"].items()},
                                                  "-": {k: sorted(v) for k, v in patch["-"].items()},
                                                  "M": {str(k): v for k, v in patch["M"].items()}}})
    # compress_relative_to_consensus on a small store: distances must be unchanged, guids reported
    ref = "ACTGTTAATTTTTTTTTGGGGGGGGGGGGAA"
    rnd = random.Random(3)
    store = []
    for i in range(14):
        s = list(ref)
        for _ in range(rnd.randint(0, 3)):
            s[rnd.randrange(len(ref))] = rnd.choice("ACGTN")
        store.append("".join(s))
    sc = REF.seqComparer(reference=ref, maxNs=1e8, snpCeiling=20, snpCompressionCeiling=3)
    for i, s in enumerate(store):
        sc.persist(sc.compress(s), "g%d" % i)
    before = {r[1]: r[2:6] for r in sc.mcompare("g0")}
    visited = sc.compress_relative_to_consensus("g0", 0.8)
    after = {r[1]: r[2:6] for r in sc.mcompare("g0")}
    assert before == after
    out["recompress"] = {"ref": ref, "seqs": store, "snpCompressionCeiling": 3, "visited": sorted(visited),
                         "mcompare_g0": before, "summary": sc.summarise_stored_items()}
    dump("consensus.json", out)


def make_msa():
    """multi_sequence_alignment outputs (seqComparer.py:691-1128; tests 45a-47c) on cases whose result does not depend
    on set iteration order (the reference samples `sample_size` guids from an unordered set)."""
    def norm(v):
        if isinstance(v, (np.floating, float)):
            return float(v)
        if isinstance(v, (np.integer,)):
            return int(v)
        return v

    out = []
    ref = "GGGGGG"
    fam_n = ["AAACGN", "CCCCGN", "TTTCGN", "GGGGGN", "NNNCGN", "ACTCGN", "TCTNGN"]
    fam_m = ["AAACGY", "CCCCGY", "TTTCGY", "GGGGGY", "NNNCGY", "ACTCGY", "TCTQGY"]
    for label, originals, maxNs, reps, kwargs in (
            ("45a", fam_n, 1e8, 1, dict()),
            ("45b", fam_n[:6] + ["TCTGGN"], 3, 1, dict()),
            ("47c", fam_n, 3, 6, dict(expected_p1=0.995, sample_size=42)),
            ("47b2", fam_m, 3, 6, dict(uncertain_base_type="M", sample_size=42)),
            ("47a_nm", fam_n, 1e8, 6, dict(uncertain_base_type="N_or_M", sample_size=42))):
        sc = REF.seqComparer(maxNs=maxNs, reference=ref, snpCeiling=10)
        names = []
        n = 0
        for _ in range(reps):
            for o in originals:
                n += 1
                g = "{0}-{1}".format(o, n)
                sc.persist(sc.compress(o), guid=g)
                names.append(g)
        sel = names[0:8] if reps > 1 else names
        res = sc.multi_sequence_alignment(sel, output="dict", **kwargs)
        clean = {}
        for k, v in res.items():
            if isinstance(v, dict):
                clean[k] = {kk: ([norm(x) for x in vv] if isinstance(vv, list) else norm(vv)) for kk, vv in v.items()}
            else:
                clean[k] = [norm(x) for x in v]
        out.append({"label": label, "ref": ref, "maxNs": maxNs, "originals": originals, "reps": reps,
                    "select": sel, "kwargs": kwargs, "got": clean})
    sc = REF.seqComparer(maxNs=3, reference=ref, snpCeiling=10)
    sc.persist(sc.compress("NNNCGN"), guid="NNNCGN-1")
    assert sc.multi_sequence_alignment(["NNNCGN-1"]) is None
    out.append({"label": "45c", "ref": ref, "maxNs": 3, "originals": ["NNNCGN"], "reps": 1, "select": ["NNNCGN-1"],
                "kwargs": {}, "got": None})
    dump("msa.json", out)


def make_cluster():
    """consensus() votes, _msa steps 1-4 and estimate_expected_unk_sites of the reference on the family-structured
    collection of collection.json (N blocks, IUPAC M, invalid samples): the fixtures behind the device kernels of
    csrc/fn3_cluster.cuh.  Everything here is independent of set iteration order: sample_size is set to the
    number of valid stored samples, so the reference's "first sample_size guids of an unordered set" is all of them."""
    with open(os.path.join(HERE, "collection.json")) as fh:
        c = json.load(fh)
    L = c["genome_length"]
    ref = synth.reference_string(synth.random_reference(L, c["ref_seed"]))
    pos, off, invalid = np.array(c["pos"], np.int32), np.array(c["off"], np.int64), np.array(c["invalid"], np.uint8)
    n = invalid.size
    excluded = synth.synthetic_mask(L, 2000, 12, seed=3)

    def as_dict(i):
        if invalid[i]:
            return {"invalid": 1}
        o = off[6 * i: 6 * i + 7]
        d = {"invalid": 0}
        for b, k in enumerate("ACGTN"):
            d[k] = set(int(x) for x in pos[o[b]:o[b + 1]])
        d["M"] = {int(x): "R" for x in pos[o[5]:o[6]]}
        return d

    sc = REF.seqComparer(reference=ref, maxNs=700, snpCeiling=20, excludePositions=set(np.flatnonzero(excluded).tolist()))
    for i in range(n):
        sc.persist(as_dict(i), "s%d" % i)
    valid = [i for i in range(n) if not invalid[i]]
    rnd = random.Random(11)
    out = {"n_valid": len(valid), "consensus": [], "msa": [], "unk_sites": []}

    # consensus: whole families, random mixtures, a member listed twice; thresholds incl. float-rounding traps
    groups = [valid[:10], valid[10:22], rnd.sample(valid, 7), rnd.sample(valid, 15), valid[:3] + valid[:3],
              [valid[4]], valid[30:40] + [valid[30]]]
    for g in groups:
        for cutoff in (0.0, 0.3, 0.5, 0.7, 0.8, 1.0):
            cons = sc.consensus([sc.seqProfile["s%d" % i] for i in g], cutoff)
            out["consensus"].append({"members": g, "cutoff": cutoff,
                                     "got": {k: sorted(int(x) for x in v) for k, v in cons.items()}})

    # _msa: selections with and without invalid members; each of the three uncertain_base_type settings
    sels = [valid[:12], valid[12:20] + [i for i in range(n) if invalid[i]][:2], rnd.sample(valid, 9), valid[40:43],
            [valid[7]], rnd.sample(range(n), 25)]
    keep = ("variant_positions", "invalid_guids", "guid2msa_seq", "guid2allN", "guid2allM", "guid2allN_or_M",
            "guid2alignN", "guid2alignM", "guid2alignN_or_M", "guid2expected_p2", "guid2expected_p1")
    for sel, unk in zip(sels, ("N", "M", "N_or_M", "N", "M", "N_or_M")):
        guids = ["s%d" % i for i in sel]
        res = sc.multi_sequence_alignment(guids, output="dict", sample_size=len(valid), uncertain_base_type=unk)
        clean = {}
        for k in keep:
            v = res[k]
            if isinstance(v, dict):
                clean[k] = {kk: (None if vv is None else (float(vv) if isinstance(vv, (float, np.floating)) else
                                 (int(vv) if isinstance(vv, (int, np.integer)) else vv))) for kk, vv in v.items()}
            else:
                clean[k] = [int(x) if isinstance(x, (int, np.integer)) else x for x in v]
        out["msa"].append({"select": guids, "unk": unk, "sample_size": len(valid), "got": clean})

    # estimate_expected_unk / estimate_expected_unk_sites over everything stored
    site_sets = [sorted(rnd.sample(range(L), 200)), out["msa"][0]["got"]["variant_positions"], []]
    for sites in site_sets:
        for unk in ("N", "M", "N_or_M"):
            got = sc.estimate_expected_unk_sites(sample_size=len(valid), sites=set(sites), unk_type=unk)
            out["unk_sites"].append({"sites": sites, "unk": unk, "sample_size": len(valid),
                                     "got": None if got is None else float(got)})
    out["unk"] = {unk: float(sc.estimate_expected_unk(sample_size=len(valid), unk_type=unk)) for unk in ("N", "M", "N_or_M")}
    out["unk_too_few"] = sc.estimate_expected_unk(sample_size=len(valid) + 1, unk_type="N")
    dump("cluster.json", out)


def make_mt():
    """The older class `seqComparer_mt.seqComparer`: N-only samples and distances, the two
    mcompare conventions (single thread: skips the query, cut-off snpCompressionCeiling, lists; threads: includes the
    query, cut-off snpCeiling, tuples), {'add','subtract'} patches, the MSA whose site rule lets N account for a
    position.  Order-independent by construction (results keyed by guid, sample_size = all valid samples)."""
    MT = ref_shim.load_reference_module("seqComparer_mt")
    with open(os.path.join(HERE, "collection.json")) as fh:
        c = json.load(fh)
    L = c["genome_length"]
    ref = synth.reference_string(synth.random_reference(L, c["ref_seed"]))
    pos, off, invalid = np.array(c["pos"], np.int32), np.array(c["off"], np.int64), np.array(c["invalid"], np.uint8)
    n = invalid.size
    excluded = set(np.flatnonzero(synth.synthetic_mask(L, 2000, 12, seed=3)).tolist())

    def as_dict(i):
        if invalid[i]:
            return {"invalid": 1}
        o = off[6 * i: 6 * i + 7]
        d = {"invalid": 0}
        for b, k in enumerate("ACGTN"):
            d[k] = set(int(x) for x in pos[o[b]:o[b + 1]])
        return d

    def setlist(d):
        return {k: sorted(int(x) for x in v) for k, v in d.items()}

    def rows_of(res):
        return sorted([r[1], r[2], r[3], r[4], r[5], None if r[6] is None else len(r[6]), None if r[7] is None else len(r[7]),
                       None if r[8] is None else sorted(int(x) for x in r[8])[:5]] for r in res)

    valid = [i for i in range(n) if not invalid[i]]
    rnd = random.Random(23)
    out = {"mcompare": [], "msa": [], "consensus": [], "compress": []}

    def store(cpu, ceiling, comp):
        sc = MT.seqComparer(reference=ref, maxNs=700, snpCeiling=ceiling, snpCompressionCeiling=comp, cpuCount=cpu,
                            excludePositions=excluded)
        for i in range(n):
            sc.persist(as_dict(i), "s%d" % i)
        return sc

    queries = valid[:3] + [i for i in range(n) if invalid[i]][:1] + rnd.sample(valid, 4)
    for cpu, ceiling, comp in ((1, 20, 250), (1, 12, 40), (1, 20, None), (3, 20, 250), (2, 12, 250)):
        sc = store(cpu, ceiling, comp)
        for q in queries:
            res = sc.mcompare("s%d" % q)
            kind = sorted(set(type(r).__name__ for r in res))
            out["mcompare"].append({"cpuCount": cpu, "snpCeiling": ceiling, "snpCompressionCeiling": comp, "query": "s%d" % q,
                                    "row_type": kind, "rows": rows_of(res)})
        subset = ["s%d" % i for i in rnd.sample(range(n), 20)]
        res = sc.mcompare("s%d" % valid[5], subset)
        out["mcompare"].append({"cpuCount": cpu, "snpCeiling": ceiling, "snpCompressionCeiling": comp, "query": "s%d" % valid[5],
                                "guids": subset, "row_type": sorted(set(type(r).__name__ for r in res)), "rows": rows_of(res)})

    sc = store(1, 20, 250)
    out["summary"] = sc.summarise_stored_items()
    sels = [valid[:12], valid[12:20] + [i for i in range(n) if invalid[i]][:2], rnd.sample(valid, 9), valid[40:43], [valid[7]],
            rnd.sample(range(n), 25)]
    for sel in sels:
        guids = ["s%d" % i for i in sel]
        res = sc.multi_sequence_alignment(guids, output="dict", sample_size=len(valid))
        clean = {}
        for k, v in res.items():
            if k == "guid2sequence":
                continue
            if isinstance(v, dict):
                clean[k] = {kk: (None if vv is None else (float(vv) if isinstance(vv, (float, np.floating)) else
                                 (int(vv) if isinstance(vv, (int, np.integer)) else vv))) for kk, vv in v.items()}
            else:
                clean[k] = [int(x) if isinstance(x, (int, np.integer)) else x for x in v]
        df = sc.multi_sequence_alignment(guids, output="df", sample_size=len(valid))
        out["msa"].append({"select": guids, "sample_size": len(valid), "got": clean, "columns": list(df.columns),
                           "p_value3_column": {g: (None if v is None or v != v else float(v)) for g, v in df["p_value3"].items()}})
    out["expected_N"] = float(sc.estimate_expected_N(sample_size=len(valid)))
    out["expected_N_too_few"] = sc.estimate_expected_N(sample_size=len(valid) + 1)
    out["expected_N_sites"] = float(sc.estimate_expected_N_sites(sample_size=len(valid), sites=set(out["msa"][0]["got"]["variant_positions"])))

    groups = [valid[:10], valid[10:22], rnd.sample(valid, 7), valid[:3] + valid[:3]]
    for g in groups:
        for cutoff in (0.3, 0.5, 0.8, 1.0):
            cons = sc.consensus([sc.seqProfile["s%d" % i] for i in g], cutoff)
            member = sc.seqProfile["s%d" % g[0]]
            patch = sc.generate_patch(member, cons)
            back = sc.apply_patch(patch, cons)
            assert back == member
            out["consensus"].append({"members": g, "cutoff": cutoff, "got": setlist(cons), "hash": sc.compressed_sequence_hash(cons),
                                     "patch_of_first": {k: setlist(v) for k, v in patch.items()}})
    seed = "s%d" % valid[0]
    visited = sc.compress_relative_to_consensus(seed)
    out["recompress"] = {"seed": seed, "visited": sorted(visited), "n_consensi": len(sc.consensi),
                         "summary": sc.summarise_stored_items(),
                         "patches": {g: {k: setlist(v) for k, v in sc.seqProfile[g]["patch"].items()} for g in sorted(visited)[:6]},
                         "neighbours_after": rows_of(sc.mcompare(seed))}

    small = MT.seqComparer(reference="ACTGACTG", maxNs=3, snpCeiling=10, excludePositions={7})
    for seq in ("ACTGACTG", "NCTGACTA", "A-TGACTG", "CCCCCCCC", "NNNNACTG", "ACTGNNNN", "ACTRACTG", "ACTGACTR", "ACTQACTG"):
        try:
            got = small.compress(seq)
            got = {"invalid": 1} if got.get("invalid") == 1 else dict(setlist({k: got[k] for k in "ACGTN"}), invalid=0)
            out["compress"].append({"seq": seq, "got": got, "roundtrip": None if got["invalid"] else small.uncompress(small.compress(seq))})
        except KeyError as ex:
            out["compress"].append({"seq": seq, "raises": "KeyError", "arg": ex.args[0]})
    dump("mt.json", out)


def make_nucleic():
    """NucleicAcid.examine of the unmodified reference: composition (key order kept, examinationDate dropped), isValid,
    the cleaned string's length — or the exception type it raises."""
    NA = ref_shim.load_reference_module("NucleicAcid")
    rnd = random.Random(5)
    cases = ["ACGT", "ACGTN-", "ACGTN-X", "", "AACCCGT", "AAAN", "acgtn-rykmsw", "RYSWKMryswkm", "ACGT" * 1000 + "NN--",
             "A", "-", "nnnn", "ACGU", "AC GT", "ACGT\n", "ÀCGT", "ACGß", 35, None,
             "".join(rnd.choice("ACGT") for _ in range(4099)) + "".join(rnd.choice("ACGTN-RYKMSWacgtn") for _ in range(333)),
             "".join(rnd.choice("ACGTNX*") for _ in range(257))]
    out = []
    for c in cases:
        na = NA.NucleicAcid()
        rec = {"input": c}
        try:
            na.examine(c)
        except Exception as ex:
            rec["raises"] = type(ex).__name__
        comp = dict(na.composition)
        comp.pop("examinationDate")
        rec.update({"composition": comp, "keys": list(na.composition.keys()), "isValid": na.isValid, "nchars": na.nchars,
                    "cleaned": None if na.nucleicAcidString is None else na.nucleicAcidString.decode("utf-8")})
        out.append(rec)
    dump("nucleic.json", out)


if __name__ == "__main__":
    if "--nucleic" in sys.argv:
        make_nucleic()
        sys.exit(0)
    if "--mt" in sys.argv:
        make_mt()
        sys.exit(0)
    if "--cluster" in sys.argv:
        make_cluster()
        sys.exit(0)
    make_msa()
    make_kat()
    make_compress()
    make_random_pairs()
    make_collection()
    make_consensus()
    make_cluster()
    make_mt()
    make_nucleic()
