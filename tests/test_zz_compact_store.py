"""seqComparer.compact_store(): the device store rebuilt without tombstones answers every query as before.
(CPU: oracle-backed stand-in engine; -m gpu: the CUDA engine: device-side fn3_compact for one engine, the host-side rebuild
for an engine group.)"""
import numpy as np
import pytest


def _check(device=None):
    from findneighbour3_b200 import synth
    from findneighbour3_b200.seqComparer import seqComparer
    L = 40000
    coll = synth.make_collection(4, 10, genome_length=L, seed=5, k1=60, s=2.0, n_blocks=3, n_mean=200, m_mean=3, rule="any",
                                 p_invalid=0.08)
    ref = synth.reference_string(synth.random_reference(L, 962))
    sc = seqComparer(reference=ref, maxNs=1e8, snpCeiling=20, device=device)
    for i in range(coll.n):
        sc.persist(coll.as_dict(i), "g%d" % i)
    for i in range(0, coll.n, 3):
        sc.remove("g%d" % i)
    sc.persist(coll.as_dict(1), "g2")                        # re-persist with other content: tombstone + append
    want = {g: sc.neighbours(g) for g in sc.guidscachedinram()}
    stats_before = sc.summarise_stored_items()
    before, after = sc.compact_store()
    assert before == coll.n + 1 and after == len(want) < before
    assert sc.compact_store() == (after, after)              # nothing left to do
    assert {g: sc.neighbours(g) for g in sc.guidscachedinram()} == want
    assert sc.summarise_stored_items() == stats_before
    assert sc.neighbours("g2")["g1"][0] == 0
    # the store keeps working: inserts, compress (the reference sequence is uploaded again on demand), removal
    p = ref[:100] + ("A" if ref[100] != "A" else "C") + ref[101:]
    sc.persist(sc.compress(p), "fresh")
    assert sc.countDifferences_byKey(("fresh", "fresh"))[2] == 0
    sc.remove("fresh")
    assert {g: sc.neighbours(g) for g in sc.guidscachedinram()} == want
    sc.engine.close()


def test_compact_store_host_layer(monkeypatch):
    import findneighbour3_b200.seqComparer as base
    from fake_engine import OracleEngine
    monkeypatch.setattr(base, "make_engine", lambda genome_length, device=0: OracleEngine(genome_length))
    _check()


@pytest.mark.gpu
@pytest.mark.parametrize("device", [0, [0, 0, 0]])
def test_compact_store_on_the_device(device):
    """one engine: fn3_compact (rows copied and heads rebuilt on the device); a group: the rebuild through the host"""
    _check(device)
