"""not-gpu: the native reader of the reference's stored sample format (pickle protocol 2 of the compressed_sequence
dict, src/mongoStore.py:344-363) against Python's own unpickler.  fn3_unpickle_samples is a pure host function."""
import pickle
import random

import numpy as np

from findneighbour3_b200 import packing


def _sample(rnd, L, n_v, n_n, n_m):
    taken = set()

    def draw(k):
        out = set()
        while len(out) < k:
            p = rnd.randrange(L)
            if p not in taken:
                taken.add(p)
                out.add(p)
        return out
    d = {b: draw(rnd.randrange(n_v + 1)) for b in "ACGT"}
    d["N"] = draw(n_n)
    d["M"] = {p: rnd.choice("RYSWKMq?") for p in draw(n_m)}
    d["invalid"] = 0
    return d


def _lists(d):
    return [sorted(d[k]) for k in "ACGTN"] + [sorted(d["M"].keys())]


def test_native_unpickle_matches_pickle_loads():
    rnd = random.Random(5)
    objs = []
    for L, n_v, n_n, n_m in ((200, 20, 30, 5), (70000, 300, 2500, 40), (4411532, 400, 4000, 300), (4411532, 0, 0, 0)):
        for _ in range(6):
            objs.append(_sample(rnd, L, n_v, n_n, n_m))
    objs.append({"invalid": 1})
    objs.append({"A": set(), "C": set(), "G": set(), "T": set(), "N": set(), "M": {}, "invalid": 0})
    blobs = [pickle.dumps(o, protocol=2) for o in objs]
    pos, off, status = packing.unpickle_many(blobs)
    assert off.size == 6 * len(objs) + 1 and off[-1] == pos.size
    for i, o in enumerate(objs):
        if o.get("invalid") == 1:
            assert status[i] == 1 and off[6 * i] == off[6 * i + 6]
            continue
        assert status[i] == 0, i
        for b, want in enumerate(_lists(o)):
            assert pos[off[6 * i + b]:off[6 * i + b + 1]].tolist() == want, (i, b)
    # the same through the packing module's own dict route
    p2, o2, inv2 = packing.pack_many(objs)
    assert np.array_equal(p2, pos) and np.array_equal(o2, off) and np.array_equal(inv2, (status == 1).astype(np.uint8))


def test_native_unpickle_reports_what_it_does_not_read():
    plain = {"A": {1}, "C": set(), "G": set(), "T": set(), "N": {2, 3}, "M": {7: "R"}, "invalid": 0}
    patch = {"consensus_md5": "abc", "patch": {"+": {"A": {1}}, "-": {}, "M": {}}}
    numpy_ints = dict(plain, A={np.int64(5)})
    odd_invalid = dict(plain, invalid=1)
    missing_key = {k: v for k, v in plain.items() if k != "T"}
    blobs = [pickle.dumps(plain, protocol=2), pickle.dumps(patch, protocol=2), pickle.dumps(numpy_ints, protocol=2),
             pickle.dumps(plain, protocol=4), pickle.dumps(odd_invalid, protocol=2), pickle.dumps(missing_key, protocol=2),
             pickle.dumps([1, 2, 3], protocol=2), b"", b"\x80\x02}q\x00(", pickle.dumps(plain, protocol=2)[:-3],
             pickle.dumps(plain, protocol=0), pickle.dumps(plain, protocol=1)]
    pos, off, status = packing.unpickle_many(blobs)
    assert status[0] == 0
    assert status[1:10].tolist() == [2] * 9
    for i in (10, 11):                       # older protocols: either read correctly or declined, never misread
        if status[i] == 0:
            assert pos[off[6 * i]:off[6 * i + 6]].tolist() == [1, 2, 3, 7]
        else:
            assert status[i] == 2
    assert pos[off[0]:off[6]].tolist() == [1, 2, 3, 7]
    assert packing.unpickle_many([])[2].size == 0


def test_native_unpickle_python2_style_pickle():
    """a protocol-2 pickle as Python 2 wrote it: str keys as SHORT_BINSTRING, set via __builtin__"""
    blob = (b"\x80\x02}q\x00(U\x01Aq\x01c__builtin__\nset\nq\x02]q\x03(K\x05M\x10\x27e\x85q\x04Rq\x05"
            b"U\x01Cq\x06h\x02]\x85Rq\x07U\x01Gq\x08h\x02]\x85Rq\tU\x01Tq\nh\x02]\x85Rq\x0b"
            b"U\x01Nq\x0ch\x02]q\r(J\xa0\x86\x01\x00K\x02e\x85Rq\x0eU\x01Mq\x0f}q\x10K\tU\x01Rq\x11sU\x07invalidq\x12K\x00u.")
    want = pickle.loads(blob)
    assert want == {"A": {5, 10000}, "C": set(), "G": set(), "T": set(), "N": {100000, 2}, "M": {9: "R"}, "invalid": 0}
    pos, off, status = packing.unpickle_many([blob])
    assert status.tolist() == [0]
    assert [pos[off[b]:off[b + 1]].tolist() for b in range(6)] == [[5, 10000], [], [], [], [2, 100000], [9]]


def test_native_unpickle_survives_damaged_input():
    """mutated / truncated pickles: never a crash; whatever is still read as a valid sample equals what Python reads"""
    rnd = random.Random(1)
    base = {"A": {1, 5, 300, 70000}, "C": set(), "G": {7}, "T": {9, 10}, "N": set(range(1000, 2500)),
            "M": {9000: "R", 12000: "R", 40000: "y"}, "invalid": 0}
    b0 = pickle.dumps(base, protocol=2)
    blobs = []
    for _ in range(3000):
        b = bytearray(b0)
        for _ in range(rnd.choice((1, 1, 2, 3, 8))):
            op = rnd.randrange(4)
            if op == 0:
                b[rnd.randrange(len(b))] = rnd.randrange(256)
            elif op == 1 and len(b) > 1:
                del b[rnd.randrange(len(b))]
            elif op == 2:
                b.insert(rnd.randrange(len(b) + 1), rnd.randrange(256))
            else:
                b = b[: max(1, rnd.randrange(len(b) + 1))]
        blobs.append(bytes(b))
    pos, off, status = packing.unpickle_many(blobs)
    assert set(status.tolist()) <= {0, 1, 2} and (status == 2).sum() > 2000
    for i in np.flatnonzero(status == 0):
        try:
            o = pickle.loads(blobs[i])
            want = [sorted(o[k]) for k in "ACGTN"] + [sorted(o["M"].keys())]
        except Exception:
            continue
        assert [pos[off[6 * i + b]:off[6 * i + b + 1]].tolist() for b in range(6)] == want
