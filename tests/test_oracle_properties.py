"""not-gpu: randomised cross-checks of the test infrastructure itself, beyond the fixed golden vectors.

 * the two restatements (Python sets, oracle/seqcomparer_port.py; sorted arrays in C, oracle/fn3_oracle.c) agree on
   random samples with N, '-', IUPAC, arbitrary characters, exclusions and invalid samples;
 * the STREAMING IDENTITY the CUDA kernel is built on (SURVEY.md §8 "Spec") — a comparison that reads
   only the candidate's lists and looks positions up in a map of the query — restated here in numpy, equals the
   reference's set algebra for dist, n1, n2 and nboth (incl. quirk Q3), and S1 is a lower bound of dist;
 * in the build container, where /root/reference exists, the UNMODIFIED reference class is run on the same random
   inputs (this is a live check on top of the committed fixtures).
"""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import c_oracle, ref_shim
from oracle import seqcomparer_port as port
from util import lists_from_dict

ALPHABET = "ACGTACGTACGTN-RYKq"


def _random_case(rng, L, n):
    ref = "".join(rng.choice(list("ACGT"), size=L))
    excluded = set(rng.choice(L, size=int(rng.integers(0, max(1, L // 4))), replace=False).tolist())
    included = [i for i in range(L) if i not in excluded]
    max_ns = int(rng.integers(1, L))
    seqs = []
    for _ in range(n):
        s = list(ref)
        for p in rng.choice(L, size=int(rng.integers(0, L)), replace=False):
            s[int(p)] = ALPHABET[int(rng.integers(0, len(ALPHABET)))]
        if rng.random() < 0.3:                                   # an N block, as in mapped data
            a = int(rng.integers(0, L))
            for p in range(a, min(L, a + int(rng.integers(1, max(2, L // 3))))):
                s[p] = "N"
        seqs.append("".join(s))
    return ref, excluded, included, max_ns, seqs


def _streaming(q, c):
    """dist, n1, n2, nboth of (query q, candidate c) reading only c's lists + point lookups into q (the kernel's math)"""
    code_q = {}
    for b in "ACGT":
        for p in q[b]:
            code_q[p] = b
    Uq = q["N"] | set(q["M"].keys())
    Vq = set(code_q)
    S1 = qdefV = 0
    for b in "ACGT":
        for p in c[b]:
            if p in Uq:
                continue
            if p in code_q:
                qdefV += 1
                S1 += code_q[p] != b
            else:
                S1 += 1
    Uc = c["N"] | set(c["M"].keys())
    qdefU = len(Uc & Vq)
    dist = S1 + len(Vq) - qdefV - qdefU
    n1 = len(Uq)
    n2 = len(c["N"]) + len(q["M"]) - len(c["N"] & set(q["M"].keys()))
    nboth = n1 + len(c["N"]) - len(c["N"] & Uq)
    return S1, dist, n1, n2, nboth


@settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck))
@given(seed=st.integers(0, 2 ** 31 - 1), L=st.integers(4, 90), n=st.integers(2, 9), ceiling=st.sampled_from([0, 1, 3, 12, 20, 10 ** 6]))
def test_restatements_and_streaming_identity_agree(seed, L, n, ceiling):
    rng = np.random.default_rng(seed)
    ref, excluded, included, max_ns, seqs = _random_case(rng, L, n)
    store = {i: port.compress(ref, included, max_ns, s) for i, s in enumerate(seqs)}
    # C oracle over the packed lists
    packed = [lists_from_dict(store[i]) for i in range(n)]
    pos = np.concatenate([p[0] for p in packed]) if packed else np.empty(0, np.int32)
    off = np.zeros(6 * n + 1, np.int64)
    total = 0
    for i, (p, o, bad) in enumerate(packed):
        off[6 * i: 6 * i + 7] = total + o.astype(np.int64)
        total += int(o[6])
    inv = np.array([1 if p[2] else 0 for p in packed], np.uint8)
    for q in range(n):
        want = port.neighbours(store, {}, q, ceiling)
        assert c_oracle.one_vs_all(pos, off, inv, q, ceiling) == want
        if store[q].get("invalid") == 1:
            assert want == {}
            continue
        for c in range(n):
            if c == q or store[c].get("invalid") == 1:
                continue
            S1, dist, n1, n2, nboth = _streaming(store[q], store[c])
            full = port.by_key(store, {}, q, c, 10 ** 9)
            assert (dist, n1, n2, nboth) == tuple(full[2:6])
            assert S1 <= dist                                    # what makes the early exit sound
            assert (c in want) == (dist <= ceiling)
        # symmetry of the distance, zero self distance
        assert port.by_key(store, {}, q, q, 10 ** 9)[2] == 0
    for a in range(n):
        for b in range(a):
            assert port.by_key(store, {}, a, b, 10 ** 9)[2] == port.by_key(store, {}, b, a, 10 ** 9)[2]


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference checkout not present")
@settings(max_examples=15, deadline=None, suppress_health_check=list(HealthCheck))
@given(seed=st.integers(0, 2 ** 31 - 1), L=st.integers(4, 60), n=st.integers(2, 7), ceiling=st.sampled_from([0, 2, 20, 10 ** 6]))
def test_port_equals_the_live_reference(seed, L, n, ceiling):
    REF = ref_shim.load_reference_module("seqComparer")
    rng = np.random.default_rng(seed)
    ref, excluded, included, max_ns, seqs = _random_case(rng, L, n)
    sc = REF.seqComparer(reference=ref, maxNs=max_ns, snpCeiling=ceiling, excludePositions=set(excluded))
    store = {}
    for i, s in enumerate(seqs):
        obj = sc.compress(s)
        store[i] = port.compress(ref, included, max_ns, s)
        assert obj == store[i]
        sc.persist(obj, i)
    for q in range(n):
        got = {r[1]: tuple(r[2:6]) for r in sc.mcompare(q)}
        want = {r[1]: tuple(r[2:6]) for r in port.mcompare(store, {}, q, ceiling)}
        assert got == want
