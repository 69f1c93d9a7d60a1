"""Host-side conversion between the reference's sample dicts and the engine's list format.

Reference formats (src/seqComparer.py:81-84):
  compressed_sequence  {'A','C','G','T','N': set[int], 'M': {pos: char}, 'invalid': 0} | {'invalid': 1}   (:290-305)
  patch_and_consensus  {'consensus_md5': md5, 'patch': {'+': {b:set}, '-': {b:set}, 'M': dict}}          (:503-548)
Engine format (include/fn3.h): six sorted int32 lists A,C,G,T,N,M concatenated + 7 offsets.
"""
import ctypes as C
import hashlib

import numpy as np

LIST_ORDER = ("A", "C", "G", "T", "N", "M")
COMPRESSED_KEYS = frozenset(("invalid", "A", "C", "G", "T", "N", "M"))
PATCHED_KEYS = frozenset(("consensus_md5", "patch"))


def is_invalid(obj):
    return isinstance(obj, dict) and obj.get("invalid", 0) == 1


def _sorted_i32(collection):
    n = len(collection)
    if n == 0:
        return np.empty(0, dtype=np.int32)
    a = np.fromiter(collection, dtype=np.int64, count=n)
    a.sort()
    return a.astype(np.int32)


def pack_compressed(obj):
    """compressed_sequence dict -> (pos int32[], off int32[7], invalid bool).

    The M characters are not needed by the compare path (seqComparer.py:449-450 use only
    the keys) and stay in the host copy of the dict."""
    if is_invalid(obj):
        return np.empty(0, dtype=np.int32), np.zeros(7, dtype=np.int32), True
    # (the older seqComparer_mt sample type has no 'M' key, src/seqComparer_mt.py:474: an empty M list)
    parts = [_sorted_i32(obj[k]) if k != "M" else _sorted_i32(obj.get("M", {}).keys()) for k in LIST_ORDER]
    off = np.zeros(7, dtype=np.int32)
    off[1:] = np.cumsum([p.size for p in parts])
    pos = np.concatenate(parts) if off[6] else np.empty(0, dtype=np.int32)
    return pos, off, False


def unpack_lists(pos, off, m_chars=None):
    """(pos, off) -> compressed_sequence dict (M characters default to 'M' when unknown)."""
    out = {"invalid": 0}
    for i, k in enumerate(LIST_ORDER[:5]):
        out[k] = set(int(x) for x in pos[off[i]:off[i + 1]])
    mpos = pos[off[5]:off[6]]
    if m_chars is None:
        out["M"] = {int(p): "M" for p in mpos}
    else:
        out["M"] = {int(p): c for p, c in zip(mpos, m_chars)}
    return out


def pack_many(objs):
    """list of compressed_sequence dicts -> (pos int32[], off int64[6n+1], invalid uint8[n]) for append_bulk."""
    n = len(objs)
    off = np.zeros(6 * n + 1, dtype=np.int64)
    inv = np.zeros(n, dtype=np.uint8)
    chunks = []
    total = 0
    for i, o in enumerate(objs):
        p, f, bad = pack_compressed(o)
        inv[i] = 1 if bad else 0
        off[6 * i: 6 * i + 7] = total + f.astype(np.int64)
        total += int(f[6])
        chunks.append(p)
    pos = np.concatenate(chunks) if total else np.empty(0, dtype=np.int32)
    return pos, off, inv


def content_key(pos, off, invalid):
    """cheap identity of a sample's expanded content (used to make re-persist of identical content a no-op)."""
    if invalid:
        return b"invalid"
    h = hashlib.blake2b(np.ascontiguousarray(off, dtype=np.int32).tobytes(), digest_size=16)
    h.update(np.ascontiguousarray(pos, dtype=np.int32).tobytes())
    return h.digest()


def unpickle_many(blobs):
    """The reference's stored samples — pickle.dumps(obj, protocol=2), src/mongoStore.py:344-363 — parsed natively
    (libfn3 fn3_unpickle_samples, host threads, no Python objects) into the engine's bulk format.

    -> (pos int32[], off int64[6n+1], status uint8[n]): status 0 = valid sample, 1 = {'invalid': 1},
    2 = not a plain compressed_sequence pickle (its lists are empty; unpickle it with `pickle.loads`)."""
    from . import _native as N
    lib = N.load()
    n = len(blobs)
    if n == 0:
        return np.empty(0, dtype=np.int32), np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.uint8)
    boff = np.zeros(n + 1, dtype=np.int64)
    if n:
        np.cumsum(np.fromiter((len(b) for b in blobs), dtype=np.int64, count=n), out=boff[1:])
    joined = np.frombuffer(b"".join(blobs), dtype=np.uint8)
    cap = max(1, int(boff[-1]) // 2)                    # an integer takes at least two bytes of pickle
    pos = np.empty(cap, dtype=np.int32)
    off = np.zeros(6 * n + 1, dtype=np.int64)
    status = np.zeros(n, dtype=np.uint8)
    n_pos = C.c_int64(0)
    rc = lib.fn3_unpickle_samples(n, joined.ctypes.data if joined.size else None, boff.ctypes.data, pos.ctypes.data, cap,
                                  off.ctypes.data, status.ctypes.data if n else None, C.byref(n_pos))
    if rc != N.FN3_OK:
        raise RuntimeError("fn3_unpickle_samples failed (%d)" % rc)
    return pos[: n_pos.value], off, status
