"""SURVEY.md §8(f)-1, the ingest side of findNeighbour3.insert (findNeighbour3-server.py:511-530): NucleicAcid.examine's
character counts on the device (fn3_byte_histogram), and the fused examine + compress + persist + mcompare call
(fn3_ingest_query, seqComparer.insert_sequence).

tests/golden/nucleic.json holds what the UNMODIFIED src/NucleicAcid.py produced (make_golden.py --nucleic).  The -m gpu
tests are the parity tests proper; the CPU tests pin the host layer with a numpy stand-in for the device histogram and the
oracle-backed stand-in engine."""
import unittest

import numpy as np
import pytest

from oracle import ref_shim
from util import golden


def _numpy_histogram(data, device=0):
    return np.bincount(np.frombuffer(data, dtype=np.uint8), minlength=256).astype(np.int64)


def _check_nucleic_golden():
    from findneighbour3_b200.NucleicAcid import NucleicAcid
    for case in golden("nucleic.json"):
        na = NucleicAcid()
        raised = None
        try:
            na.examine(case["input"])
        except Exception as ex:          # noqa: BLE001 — the exception TYPE is what is compared
            raised = type(ex).__name__
        assert raised == case.get("raises"), case["input"]
        comp = dict(na.composition)
        assert isinstance(comp.pop("examinationDate"), type(na.composition["examinationDate"]))
        assert comp == case["composition"], case["input"]
        assert list(na.composition.keys()) == case["keys"]
        assert na.isValid == case["isValid"] and na.nchars == case["nchars"]
        assert (None if na.nucleicAcidString is None else na.nucleicAcidString.decode("utf-8")) == case["cleaned"]


def _run_reference_nucleic_tests():
    """the reference's own unittest classes for NucleicAcid (src/NucleicAcid.py:67-), class swapped for the drop-in"""
    import findneighbour3_b200.NucleicAcid as dropin
    ref = ref_shim.load_reference_module("NucleicAcid")
    original = ref.NucleicAcid
    ran, bad = 0, {}
    try:
        ref.NucleicAcid = dropin.NucleicAcid
        for name in sorted(dir(ref)):
            obj = getattr(ref, name)
            if isinstance(obj, type) and issubclass(obj, unittest.TestCase) and name.startswith("Test_NucleicAcid_") \
                    and "Base" not in name:
                res = unittest.TestResult()
                unittest.defaultTestLoader.loadTestsFromTestCase(obj).run(res)
                ran += res.testsRun
                if res.errors or res.failures:
                    bad[name] = (res.errors + res.failures)[0][1].strip().splitlines()[-1]
    finally:
        ref.NucleicAcid = original
    return ran, bad


def test_nucleic_acid_host_layer(monkeypatch):
    import findneighbour3_b200.engine as engine
    monkeypatch.setattr(engine, "byte_histogram", _numpy_histogram)
    _check_nucleic_golden()
    if ref_shim.reference_available():
        ran, bad = _run_reference_nucleic_tests()
        assert ran >= 8 and not bad, bad


def test_nucleic_acid_fails_loudly_without_a_device():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("checks the no-GPU failure mode")
    except ImportError:
        pass
    from findneighbour3_b200.NucleicAcid import NucleicAcid
    from findneighbour3_b200.engine import EngineError
    with pytest.raises(EngineError):
        NucleicAcid().examine("ACGT")


def _check_insert_sequence(genome_length, n, seed, max_ns, mt=False):
    """insert_sequence == examine + compress + persist + mcompare, sample by sample"""
    if mt:
        from findneighbour3_b200.seqComparer_mt import seqComparer
    else:
        from findneighbour3_b200.seqComparer import seqComparer
    from findneighbour3_b200.NucleicAcid import NucleicAcid
    rng = np.random.default_rng(seed)
    ref = "".join(rng.choice(list("ACGT"), size=genome_length))
    excluded = set(rng.choice(genome_length, size=genome_length // 10, replace=False).tolist())
    fused = seqComparer(reference=ref, maxNs=max_ns, snpCeiling=12, excludePositions=excluded, cpuCount=1)
    plain = seqComparer(reference=ref, maxNs=max_ns, snpCeiling=12, excludePositions=excluded, cpuCount=1)
    base = np.frombuffer(ref.encode(), dtype=np.uint8)
    ancestors = []
    for i in range(n):
        if i % 5 == 0 or not ancestors:
            seq = base.copy()
            at = rng.choice(genome_length, size=min(genome_length, 40), replace=False)
            seq[at] = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=at.size)
            ancestors.append(seq)
        seq = ancestors[-1].copy()
        at = rng.choice(genome_length, size=min(genome_length, 6), replace=False)
        seq[at] = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=at.size)
        s0 = int(rng.integers(0, genome_length))
        seq[s0: s0 + int(rng.integers(0, max(2, genome_length // 20)))] = ord("N")
        if not mt:
            at = rng.choice(genome_length, size=min(genome_length, int(rng.integers(0, 8))), replace=False)
            seq[at] = rng.choice(np.frombuffer(b"RYKMSW-", np.uint8), size=at.size)
        text = seq.tobytes().decode()
        guid = "g%d" % i
        found, hist = fused.insert_sequence(guid, text)
        plain.persist(plain.compress(text), guid)
        want = {r[1]: tuple(r[2:6]) for r in plain.mcompare(guid) if r[2] is not None and r[2] <= 12}
        assert found == want, (i, found, want)
        assert fused.seqProfile[guid] == plain.seqProfile[guid]
        na1, na2 = NucleicAcid(), NucleicAcid()
        na1.examine(text, histogram=hist)
        na2.examine(text)
        c1, c2 = dict(na1.composition), dict(na2.composition)
        c1.pop("examinationDate"), c2.pop("examinationDate")
        assert c1 == c2 and c1["A"] == text.count("A") and c1["N"] == text.count("N") and c1["mixed"] == sum(
            text.count(ch) for ch in "RYKMSW")
    assert fused.summarise_stored_items() == plain.summarise_stored_items()
    # stored afterwards exactly as if persisted: every query answers the same
    for i in range(0, n, 7):
        a = {r[1]: tuple(r[2:6]) for r in fused.mcompare("g%d" % i)}
        b = {r[1]: tuple(r[2:6]) for r in plain.mcompare("g%d" % i)}
        assert a == b
    # re-inserting a guid replaces it; a wrong length raises like compress()
    fused.insert_sequence("g0", ref)
    assert fused.seqProfile["g0"] == plain.compress(ref)
    with pytest.raises(TypeError):
        fused.insert_sequence("bad", ref[:-1])
    if mt:
        p = next(i for i in range(genome_length) if i not in excluded)
        with pytest.raises(KeyError):
            fused.insert_sequence("iupac", ref[:p] + "R" + ref[p + 1:])
        assert not fused.iscachedinram("iupac")
    fused.engine.close()
    plain.engine.close()


@pytest.mark.parametrize("mt", [False, True])
def test_insert_sequence_host_layer(monkeypatch, mt):
    import findneighbour3_b200.engine as engine
    import findneighbour3_b200.seqComparer as base
    from fake_engine import OracleEngine
    monkeypatch.setattr(engine, "byte_histogram", _numpy_histogram)
    monkeypatch.setattr(base, "make_engine", lambda genome_length, device=0: OracleEngine(genome_length))
    _check_insert_sequence(600, 24, 3, max_ns=40, mt=mt)


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_byte_histogram_on_the_device():
    from findneighbour3_b200.engine import byte_histogram
    rng = np.random.default_rng(11)
    for n in (0, 1, 15, 16, 17, 255, 4096, 4097, 65536 + 3, 4411532, 9000001):
        data = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
        assert np.array_equal(byte_histogram(data), _numpy_histogram(data)), n
    # sequence-shaped input: four hot values (the register-counted path) and a few rare ones
    seq = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=4411532)
    seq[rng.choice(seq.size, size=5000, replace=False)] = rng.choice(np.frombuffer(b"N-RYKMSWacgtnX", np.uint8), size=5000)
    assert np.array_equal(byte_histogram(seq.tobytes()), _numpy_histogram(seq.tobytes()))
    for ch in (b"A", b"-", b"N", b"\x00", b"\xff"):
        assert byte_histogram(ch * 100003)[ch[0]] == 100003


def _check_concurrent_compress():
    """compress() and insert_sequence() from several request threads at once (the server compresses outside its write
    semaphore, findNeighbour3-server.py:513): every thread gets its own sample back, not another thread's buffer"""
    import threading
    from findneighbour3_b200.seqComparer import seqComparer
    rng = np.random.default_rng(21)
    L = 20000
    ref = "".join(rng.choice(list("ACGT"), size=L))
    sc = seqComparer(reference=ref, maxNs=L, snpCeiling=20)
    seqs = []
    for t in range(6):
        s = np.frombuffer(ref.encode(), dtype=np.uint8).copy()
        at = rng.choice(L, size=50 + 400 * t, replace=False)
        s[at] = rng.choice(np.frombuffer(b"ACGTNR", np.uint8), size=at.size)
        seqs.append(s.tobytes().decode())
    want = [sc.compress(s) for s in seqs]
    errors = []

    def work(t):
        try:
            for rep in range(8):
                assert sc.compress(seqs[t]) == want[t]
                sc.insert_sequence("t%d_%d" % (t, rep), seqs[t])
                assert sc.seqProfile["t%d_%d" % (t, rep)] == want[t]
        except Exception as ex:          # noqa: BLE001
            errors.append(repr(ex))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(6)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors[:3]
    assert len(sc.guidscachedinram()) == 48
    for t in range(6):                                       # the eight copies of a sample find each other at distance 0
        got = sc.neighbours("t%d_0" % t)
        assert sum(1 for g, v in got.items() if g.startswith("t%d_" % t) and v[0] == 0) == 7
    sc.engine.close()


def test_concurrent_compress_host_layer(monkeypatch):
    import findneighbour3_b200.seqComparer as base
    from fake_engine import OracleEngine
    monkeypatch.setattr(base, "make_engine", lambda genome_length, device=0: OracleEngine(genome_length))
    _check_concurrent_compress()


@pytest.mark.gpu
def test_concurrent_compress_on_the_device():
    _check_concurrent_compress()


@pytest.mark.gpu
def test_nucleic_acid_golden_on_the_device():
    _check_nucleic_golden()


@pytest.mark.gpu
@pytest.mark.parametrize("shape", ["small", "tiny", "mt", "tb_full_genome"])
def test_insert_sequence_matches_separate_calls(shape):
    if shape == "small":
        _check_insert_sequence(30000, 40, 5, max_ns=2500)
    elif shape == "tiny":
        _check_insert_sequence(31, 30, 6, max_ns=6)
    elif shape == "mt":
        _check_insert_sequence(5000, 30, 7, max_ns=400, mt=True)
    else:
        _check_insert_sequence(4411532, 12, 8, max_ns=130000)
