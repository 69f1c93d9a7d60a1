"""TEST INFRASTRUCTURE — ctypes face of oracle/fn3_oracle.c (the C restatement of the reference)."""
import ctypes as C

import numpy as np

from .build_oracle import build_oracle

_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build_oracle())
        i32p, i64p, u8p = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_uint8)
        lib.fn3o_pair.restype = C.c_int
        lib.fn3o_pair.argtypes = [i32p, i32p, C.c_int, i32p, i32p, C.c_int, C.c_int64, i32p]
        lib.fn3o_one_vs_all.restype = C.c_int64
        lib.fn3o_one_vs_all.argtypes = [i32p, i64p, u8p, C.c_int64, C.c_int64, C.c_int64, C.c_int, i64p, i32p]
        _lib = lib
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def pair(pos1, off1, inv1, pos2, off2, inv2, cutoff):
    """-> (dist,n1,n2,nboth) or None."""
    lib = _load()
    p1 = np.ascontiguousarray(pos1 if not inv1 else [], dtype=np.int32)
    p2 = np.ascontiguousarray(pos2 if not inv2 else [], dtype=np.int32)
    o1 = np.ascontiguousarray(off1 if not inv1 else np.zeros(7), dtype=np.int32)
    o2 = np.ascontiguousarray(off2 if not inv2 else np.zeros(7), dtype=np.int32)
    out = np.zeros(4, dtype=np.int32)
    r = lib.fn3o_pair(_p(p1, C.c_int32), _p(o1, C.c_int32), int(bool(inv1)), _p(p2, C.c_int32), _p(o2, C.c_int32),
                      int(bool(inv2)), int(cutoff), _p(out, C.c_int32))
    if r < 0:
        raise MemoryError("oracle out of memory")
    return tuple(int(v) for v in out) if r == 1 else None


def one_vs_all(pos, off, invalid, query, cutoff, nthreads=1):
    """store = (pos int32[], off int64[6n+1], invalid uint8[n] or None) -> dict {row: (dist,n1,n2,nboth)}."""
    lib = _load()
    pos = np.ascontiguousarray(pos, dtype=np.int32)
    off = np.ascontiguousarray(off, dtype=np.int64)
    n = (off.size - 1) // 6
    inv = None if invalid is None else np.ascontiguousarray(invalid, dtype=np.uint8)
    rows = np.zeros(max(n, 1), dtype=np.int64)
    vals = np.zeros((max(n, 1), 4), dtype=np.int32)
    cnt = lib.fn3o_one_vs_all(_p(pos, C.c_int32), _p(off, C.c_int64), _p(inv, C.c_uint8), n, int(query), int(cutoff),
                              int(nthreads), _p(rows, C.c_int64), _p(vals, C.c_int32))
    if cnt < 0:
        raise MemoryError("oracle out of memory")
    return {int(rows[i]): tuple(int(v) for v in vals[i]) for i in range(cnt)}
