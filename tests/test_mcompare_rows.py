"""not-gpu: the lazily materialised result of seqComparer.mcompare (_McompareRows) behaves like the reference's list of
9-element lists (src/seqComparer.py:149-172) — checked against the LIVE unmodified reference where it is present, and
against the Python port otherwise.  Engine = the oracle-backed stand-in (this box has no GPU); `-m gpu` repeats it on the
CUDA engine."""
import random

import pytest

from oracle import ref_shim
from oracle import seqcomparer_port as port


def _scenario(sc_new, sc_ref=None):
    rnd = random.Random(5)
    L = 60
    ref = "".join(rnd.choice("ACGT") for _ in range(L))
    seqs = {}
    for i in range(25):
        base = list(ref)
        for _ in range(rnd.randint(0, 6)):
            base[rnd.randrange(L)] = rnd.choice("ACGTN-RY")
        seqs["g%d" % i] = "".join(base)
    seqs["bad"] = "N" * L
    a = sc_new(ref)
    b = sc_ref(ref) if sc_ref else None
    for g, s in seqs.items():
        a.persist(a.compress(s), g)
        if b:
            b.persist(b.compress(s), g)
    return a, b, seqs


def _check(a, b, seqs):
    for guid in ("g3", "g11", "bad"):
        res = a.mcompare(guid)
        if b is not None:
            want = b.mcompare(guid)
        else:
            want = port.mcompare({g: a.seqProfile[g] for g in seqs}, {}, guid, a.snpCeiling)
        key = lambda r: r[1]
        assert len(res) == len(want) == len(seqs) - 1
        assert sorted(res, key=key) == sorted(want, key=key)               # iteration, rows equal incl. the position sets
        assert sorted(list(res), key=key) == sorted(want, key=key)
        assert res == list(res) and list(res) == res[:] and res[0] == list(res)[0] and res[-1] == list(res)[-1]
        assert res[2:5] == list(res)[2:5] and [r[1] for r in res] == [r[1] for r in list(res)]
        assert all(isinstance(r, list) and len(r) == 9 for r in res)
        assert (res[1] in res) and (["x"] not in res)
        n_hits = sum(1 for r in res if r[2] is not None)
        assert n_hits == sum(1 for r in want if r[2] is not None)
        # the unmodified server's loop (findNeighbour3-server.py:532-538)
        links = {}
        for (guid1, guid2, dist, n1, n2, nboth, N1pos, N2pos, Nbothpos) in res:
            if not guid1 == guid2 and dist is not None and dist <= a.snpCeiling:
                links[guid2] = {"dist": dist, "n1": n1, "n2": n2, "nboth": nboth}
        assert links == {g: dict(zip(("dist", "n1", "n2", "nboth"), v)) for g, v in a.neighbours(guid).items()}
    sub = ["g1", "g2", "g3", "g3", "g9"]
    res = a.mcompare("g3", sub)
    assert sorted(r[1] for r in res) == ["g1", "g2", "g9"]
    with pytest.raises(KeyError):
        a.mcompare("nope")


def _new(ref, **kw):
    from findneighbour3_b200.seqComparer import seqComparer
    return seqComparer(reference=ref, maxNs=20, snpCeiling=8, **kw)


def test_mcompare_rows_host_layer(monkeypatch):
    import findneighbour3_b200.seqComparer as base
    from fake_engine import OracleEngine
    monkeypatch.setattr(base, "make_engine", lambda genome_length, device=0: OracleEngine(genome_length))
    sc_ref = None
    if ref_shim.reference_available():
        mod = ref_shim.load_reference_module("seqComparer")
        sc_ref = lambda ref: mod.seqComparer(reference=ref, maxNs=20, snpCeiling=8)
    _check(*_scenario(_new, sc_ref))


@pytest.mark.gpu
@pytest.mark.parametrize("device", [0, [0, 0]])
def test_mcompare_rows_on_the_device(device):
    _check(*_scenario(lambda ref: _new(ref, device=device)))
