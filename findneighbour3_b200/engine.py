"""Thin Python face of the C-ABI (include/fn3.h): numpy arrays in, numpy structured arrays out.

`Engine` owns one fn3_engine (one CUDA device).  It knows nothing about guids or the
reference's dict format — that is seqComparer.py / packing.py.  Every method maps to
exactly one C entry point; errors become the exception classes the reference's API uses
(SURVEY.md §8(b)): FN3_EARG -> ValueError, FN3_ENOTFOUND -> KeyError, the rest -> RuntimeError.
"""
import ctypes as C
import threading

import numpy as np

from . import _native as N

HIT_DTYPE = np.dtype([("row", "<i8"), ("dist", "<i4"), ("n1", "<i4"), ("n2", "<i4"), ("nboth", "<i4")])
EDGE_DTYPE = np.dtype([("row1", "<i8"), ("row2", "<i8"), ("dist", "<i4"), ("n1", "<i4"), ("n2", "<i4"),
                       ("nboth", "<i4")])
assert HIT_DTYPE.itemsize == C.sizeof(N.Hit) and EDGE_DTYPE.itemsize == C.sizeof(N.Edge)

_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_u8p = C.POINTER(C.c_uint8)
_f32p = C.POINTER(C.c_float)


class EngineError(RuntimeError):
    pass


def _as_i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _ptr(a, typ=None):
    """address of a numpy buffer for a void* argument (None -> NULL).  `ndarray.ctypes` costs 2-4 us per use; exporting
    the buffer to a one-byte ctypes view costs 0.5 us (it needs a writable, non-empty buffer: else the slow way)."""
    if a is None:
        return None
    try:
        return C.addressof(C.c_char.from_buffer(a))
    except (TypeError, ValueError):
        return a.ctypes.data


def _take(buf, n):
    """copy of the first n records of a structured buffer (a byte copy: slicing + .copy() of a structured dtype is 4x slower)"""
    isz = buf.dtype.itemsize
    return buf.view(np.uint8)[: n * isz].copy().view(buf.dtype)


def byte_histogram(data, device=0):
    """NucleicAcid.examine's character counts on the device (fn3_byte_histogram): bytes -> int64[256]."""
    lib = N.load()
    hist = np.zeros(256, dtype=np.int64)
    buf = np.frombuffer(data, dtype=np.uint8)
    rc = lib.fn3_byte_histogram(int(device), buf.ctypes.data if buf.size else None, buf.size, hist.ctypes.data)
    if rc != N.FN3_OK:
        raise EngineError("fn3_byte_histogram failed (%d): no usable CUDA device? There is no CPU path." % rc)
    return hist


# calls that exist both for one engine (fn3_<name>) and for an engine group (fn3_group_<name>)
_SHARED_CALLS = ("append", "append_bulk", "remove", "stats_get", "row_get", "set_reference", "one_vs_all", "insert_query",
                 "ingest_query", "lists_vs_all", "pair", "consensus", "consensus_lists", "msa_sites_rule", "msa_align",
                 "launch_count", "last_error", "destroy")
# calls that always run on ONE engine (a group uses its first member)
_LOCAL_CALLS = ("compress", "pair_lists", "profile_queries")


class Engine:
    def __init__(self, genome_length, device=0):
        self._lib = N.load()
        self._h = self._lib.fn3_create(int(genome_length), int(device))
        if not self._h:
            msg = self._lib.fn3_last_error(None)
            raise EngineError("fn3_create failed: %s" % (msg.decode() if msg else "unknown"))
        self._hl = self._h
        self._bind("fn3_")
        self.genome_length = int(genome_length)
        self.device = int(device)
        self.n_devices = 1
        self._init_buffers()

    def _bind(self, prefix):
        for name in _SHARED_CALLS:
            setattr(self, "_f_" + name, getattr(self._lib, prefix + name))
        for name in _LOCAL_CALLS:
            setattr(self, "_f_" + name, getattr(self._lib, "fn3_" + name))

    def _init_buffers(self):
        self._hits = np.empty(4096, dtype=HIT_DTYPE)
        self._edges = np.empty(1 << 16, dtype=EDGE_DTYPE)
        self._cpos = np.empty(1 << 16, dtype=np.int32)
        self._cm = C.create_string_buffer(1 << 12)
        self._n_rows = 0          # rows appended through this object (sizes the hit buffer without an ABI call)
        # the result buffers below are reused from call to call: calls that fill them are serialised per Engine object
        # (the library serialises queries on its own mutex anyway, so no concurrency is lost).  Appends and removals
        # take no Python lock and run concurrently with queries, as the server needs (SURVEY.md §8(b) "Threading").
        self._q_lock = threading.Lock()       # _hits, _edges
        self._c_lock = threading.Lock()       # _cpos, _cm (compress outputs); taken before _q_lock

    # -- plumbing -------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._f_destroy(self._h)
            self._h = None
            self._hl = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc == N.FN3_OK:
            return
        msg = self._f_last_error(self._h)
        msg = msg.decode() if msg else "fn3 error %d" % rc
        if rc == N.FN3_EARG:
            raise ValueError(msg)
        if rc == N.FN3_ENOTFOUND:
            raise KeyError(msg)
        if rc == N.FN3_ENOMEM:
            raise MemoryError(msg)
        raise EngineError("fn3 error %d: %s" % (rc, msg))

    # -- store ----------------------------------------------------------------------------
    def append(self, pos, off, invalid=False):
        """persist one sample (six sorted lists A,C,G,T,N,M as pos + 7 offsets) -> row index."""
        row = C.c_int64(-1)
        if invalid:
            rc = self._f_append(self._h, None, None, 1, C.byref(row))
        else:
            pos = _as_i32(pos)
            off = _as_i32(off)
            if off.shape != (7,):
                raise ValueError("off must have 7 entries")
            rc = self._f_append(self._h, _ptr(pos, _i32p), _ptr(off, _i32p), 0, C.byref(row))
        self._check(rc)
        self._n_rows = max(self._n_rows, row.value + 1)
        return row.value

    def append_bulk(self, pos, off, invalid=None):
        """persist n samples: pos concatenated, off int64[6n+1], invalid uint8[n] or None -> first row."""
        pos = _as_i32(pos)
        off = np.ascontiguousarray(off, dtype=np.int64)
        n = (off.size - 1) // 6
        if off.size != 6 * n + 1:
            raise ValueError("off must have 6*n+1 entries")
        inv = None if invalid is None else np.ascontiguousarray(invalid, dtype=np.uint8)
        if inv is not None and inv.size != n:
            raise ValueError("invalid must have n entries")
        first = C.c_int64(-1)
        rc = self._f_append_bulk(self._h, n, _ptr(pos, _i32p), _ptr(off, _i64p), _ptr(inv, _u8p), C.byref(first))
        self._check(rc)
        self._n_rows = max(self._n_rows, first.value + n)
        return first.value

    def remove(self, row):
        self._check(self._f_remove(self._h, int(row)))

    def stats(self):
        st = N.Stats()
        self._check(self._f_stats_get(self._h, C.byref(st)))
        return {k: getattr(st, k) for k, _ in N.Stats._fields_}

    def refresh_heads(self, min_ppm=10000):
        """rebuild the heads of the low-ceiling filter from a fresh background bitmap (fn3_refresh_heads; single engine: a
        group refreshes each member automatically)"""
        self._check(self._lib.fn3_refresh_heads(self._hl, int(min_ppm)))

    def compact(self):
        """reclaim tombstoned rows on the device (fn3_compact) -> int64 array: new row of every old row, -1 = was removed.
        Single engine only (a group keeps global row g on device g % G: its shards cannot be renumbered independently)."""
        if self.n_devices != 1:
            raise NotImplementedError("device-side compaction is per engine")
        n = int(self.stats()["n_rows"])
        m = np.empty(max(n, 1), dtype=np.int64)
        nb, na = C.c_int64(0), C.c_int64(0)
        self._check(self._lib.fn3_compact(self._hl, _ptr(m), m.size, C.byref(nb), C.byref(na)))
        self._n_rows = int(na.value)
        return m[: nb.value]

    def row_get(self, row):
        """-> (pos int32[], off int32[7], invalid bool) of a stored row."""
        off = np.zeros(7, dtype=np.int32)
        inv = C.c_int32(0)
        npos = C.c_int64(0)
        cap = 1024
        while True:
            pos = np.empty(cap, dtype=np.int32)
            rc = self._f_row_get(self._h, int(row), _ptr(pos, _i32p), cap, _ptr(off, _i32p),
                                       C.byref(inv), C.byref(npos))
            if rc == N.FN3_ECAPACITY:
                cap = int(npos.value)
                continue
            self._check(rc)
            return pos[: npos.value].copy(), off, bool(inv.value)

    # -- ingest ---------------------------------------------------------------------------
    def set_reference(self, reference, excluded=None):
        """upload the reference string and the excluded positions (needed by compress only)."""
        ref = reference.encode("ascii") if isinstance(reference, str) else bytes(reference)
        if len(ref) != self.genome_length:
            raise ValueError("reference length differs from the engine's genome length")
        ex = None
        if excluded is not None and len(excluded) > 0:
            ex = np.sort(np.fromiter(excluded, dtype=np.int64, count=len(excluded))).astype(np.int32)
        self._check(self._f_set_reference(self._h, ref, _ptr(ex, _i32p), 0 if ex is None else ex.size))

    def compress(self, seq_bytes):
        """device compress of one sequence (bytes, len = genome length) -> (pos int32[], off int32[7], m_chars bytes)."""
        with self._c_lock:
            if len(seq_bytes) != self.genome_length:
                raise ValueError("sequence length differs from the engine's genome length")
            off = np.zeros(7, dtype=np.int32)
            npos = C.c_int64(0)
            while True:
                rc = self._f_compress(self._hl, seq_bytes, _ptr(self._cpos, _i32p), self._cpos.size,
                                            _ptr(off, _i32p), self._cm, len(self._cm), C.byref(npos))
                if rc == N.FN3_ECAPACITY:
                    n_m = int(off[6] - off[5])
                    if npos.value > self._cpos.size:
                        self._cpos = np.empty(int(npos.value) * 2, dtype=np.int32)
                    if n_m > len(self._cm):
                        self._cm = C.create_string_buffer(n_m * 2)
                    continue
                self._check(rc)
                n_m = int(off[6] - off[5])
                return self._cpos[: npos.value].copy(), off, self._cm.raw[:n_m]

    # -- hot path -------------------------------------------------------------------------
    def _hit_buf(self, n):
        if self._hits.size < n:
            self._hits = np.empty(max(n, 2 * self._hits.size), dtype=HIT_DTYPE)
            self._hits_ptr = None
        if getattr(self, "_hits_ptr", None) is None:
            self._hits_ptr = self._hits.ctypes.data                          # cached: microseconds per call otherwise
        return self._hits

    def one_vs_all(self, query_row, ceiling, rows=None):
        """mcompare: stored row vs all (or `rows`); -> structured array of survivors (copy)."""
        with self._q_lock:
            rows_a = None if rows is None else np.ascontiguousarray(rows, dtype=np.int64)
            n_cand = self._n_rows if rows_a is None else rows_a.size
            buf = self._hit_buf(max(int(n_cand), 1))
            n_out = C.c_int64(0)
            rc = self._f_one_vs_all(self._h, int(query_row), int(ceiling), _ptr(rows_a, _i64p),
                                          0 if rows_a is None else rows_a.size,
                                          self._hits_ptr, buf.size, C.byref(n_out))
            self._check(rc)
            return _take(buf, n_out.value)

    def insert_query(self, pos, off, invalid, ceiling):
        """persist + mcompare in one ABI call -> (row, structured array of survivors among the earlier rows)."""
        with self._q_lock:
            buf = self._hit_buf(max(self._n_rows + 1, 1))
            row = C.c_int64(-1)
            n_out = C.c_int64(0)
            if invalid:
                pos_a, off_a = None, None
            else:
                pos_a, off_a = _as_i32(pos), _as_i32(off)
            rc = self._f_insert_query(self._h, _ptr(pos_a, _i32p), _ptr(off_a, _i32p), 1 if invalid else 0,
                                            int(ceiling), C.byref(row), self._hits_ptr, buf.size, C.byref(n_out))
            if row.value >= 0:
                self._n_rows = max(self._n_rows, row.value + 1)
            if rc == N.FN3_ECAPACITY and row.value >= 0:
                # the row IS stored (include/fn3.h); only the neighbour buffer was too small: ask again with room
                buf = self._hit_buf(int(n_out.value))
                rc = self._f_one_vs_all(self._h, row.value, int(ceiling), None, 0, self._hits_ptr, buf.size, C.byref(n_out))
            self._check(rc)
            return row.value, _take(buf, n_out.value)

    def ingest_query(self, seq_bytes, max_ns, ceiling, want_hist=True):
        """findNeighbour3.insert()'s data path in one ABI call (fn3_ingest_query): composition histogram + compress +
        maxNs decision + persist + mcompare from one host->device copy of the sequence.
        -> (row, invalid, pos int32[], off int32[7], m_chars bytes, survivors, hist int64[256] or None)."""
        if len(seq_bytes) != self.genome_length:
            raise ValueError("sequence length differs from the engine's genome length")
        with self._c_lock, self._q_lock:
            off = np.zeros(7, dtype=np.int32)
            hist = np.zeros(256, dtype=np.int64) if want_hist else None
            npos, row, n_out, invalid = C.c_int64(0), C.c_int64(-1), C.c_int64(0), C.c_int32(0)
            buf = self._hit_buf(max(self._n_rows + 1, 1))
            cap = int(max_ns) if max_ns < (1 << 62) else (1 << 62)
            while True:
                rc = self._f_ingest_query(self._h, seq_bytes, cap, int(ceiling), _ptr(hist), _ptr(self._cpos), self._cpos.size,
                                                _ptr(off), self._cm, len(self._cm), C.byref(npos), C.byref(invalid),
                                                C.byref(row), self._hits_ptr, buf.size, C.byref(n_out))
                if rc == N.FN3_ECAPACITY and row.value < 0:                 # the compress step ran out of room: nothing stored
                    n_m = int(off[6] - off[5])
                    if npos.value > self._cpos.size:
                        self._cpos = np.empty(int(npos.value) * 2, dtype=np.int32)
                    if n_m > len(self._cm):
                        self._cm = C.create_string_buffer(n_m * 2)
                    continue
                break
            if row.value >= 0:
                self._n_rows = max(self._n_rows, row.value + 1)
            if rc == N.FN3_ECAPACITY and row.value >= 0:           # stored; the neighbour buffer was too small
                buf = self._hit_buf(int(n_out.value))
                rc = self._f_one_vs_all(self._h, row.value, int(ceiling), None, 0, self._hits_ptr, buf.size, C.byref(n_out))
            self._check(rc)
            n_m = int(off[6] - off[5])
            return (row.value, bool(invalid.value), self._cpos[: npos.value].copy(), off, self._cm.raw[:n_m],
                    _take(buf, n_out.value), hist)

    def lists_vs_all(self, pos, off, invalid, ceiling, rows=None):
        with self._q_lock:
            rows_a = None if rows is None else np.ascontiguousarray(rows, dtype=np.int64)
            n_cand = self._n_rows if rows_a is None else rows_a.size
            buf = self._hit_buf(max(int(n_cand), 1))
            n_out = C.c_int64(0)
            if invalid:
                pos_a, off_a = None, None
            else:
                pos_a, off_a = _as_i32(pos), _as_i32(off)
            rc = self._f_lists_vs_all(self._h, _ptr(pos_a, _i32p), _ptr(off_a, _i32p), 1 if invalid else 0,
                                            int(ceiling), _ptr(rows_a, _i64p), 0 if rows_a is None else rows_a.size,
                                            self._hits_ptr, buf.size, C.byref(n_out))
            self._check(rc)
            return _take(buf, n_out.value)

    def all_vs_all(self, ceiling, upper_only=True, row_begin=0, row_end=None, stride=1, phase=0):
        """distmat over the block rows this caller owns -> structured array of edges (copy)."""
        with self._q_lock:
            if row_end is None:
                row_end = self.stats()["n_rows"]
            n_out = C.c_int64(0)
            buf = self._edges
            rc = self._lib.fn3_all_vs_all(self._h, int(ceiling), 1 if upper_only else 0, int(row_begin), int(row_end),
                                          int(stride), int(phase), buf.ctypes.data_as(C.POINTER(N.Edge)), buf.size,
                                          C.byref(n_out))
            if rc == N.FN3_ECAPACITY:                 # the edges are still on the device: fetch, do not recompute
                buf = self._edges = np.empty(int(n_out.value), dtype=EDGE_DTYPE)
                rc = self._lib.fn3_all_vs_all_fetch(self._h, buf.ctypes.data_as(C.POINTER(N.Edge)), buf.size, C.byref(n_out))
            self._check(rc)
            return buf[: n_out.value].copy()

    def pair(self, row1, row2, ceiling):
        """countDifferences_byKey on stored rows -> (dist,n1,n2,nboth) or None."""
        hit = N.Hit()
        has = C.c_int32(0)
        self._check(self._f_pair(self._h, int(row1), int(row2), int(ceiling), C.byref(hit), C.byref(has)))
        if not has.value:
            return None
        return (hit.dist, hit.n1, hit.n2, hit.nboth)

    def pair_lists(self, pos1, off1, invalid1, pos2, off2, invalid2, ceiling):
        """countDifferences on two un-stored samples -> (dist,n1,n2,nboth) or None."""
        hit = N.Hit()
        has = C.c_int32(0)
        p1 = None if invalid1 else _as_i32(pos1)
        o1 = None if invalid1 else _as_i32(off1)
        p2 = None if invalid2 else _as_i32(pos2)
        o2 = None if invalid2 else _as_i32(off2)
        rc = self._f_pair_lists(self._hl, _ptr(p1, _i32p), _ptr(o1, _i32p), 1 if invalid1 else 0,
                                      _ptr(p2, _i32p), _ptr(o2, _i32p), 1 if invalid2 else 0, int(ceiling),
                                      C.byref(hit), C.byref(has))
        self._check(rc)
        if not has.value:
            return None
        return (hit.dist, hit.n1, hit.n2, hit.nboth)

    # -- cluster statistics ---------------------------------------------------------------
    def _consensus_call(self, call):
        """shared output protocol of fn3_consensus / fn3_consensus_lists -> (pos int32[], off int32[6])."""
        off = np.zeros(6, dtype=np.int32)
        npos = C.c_int64(0)
        cap = 1 << 14
        while True:
            pos = np.empty(cap, dtype=np.int32)
            rc = call(_ptr(pos, _i32p), cap, _ptr(off, _i32p), C.byref(npos))
            if rc == N.FN3_ECAPACITY:
                cap = int(npos.value)
                continue
            self._check(rc)
            return pos[: npos.value].copy(), off

    def consensus(self, rows, min_votes):
        """per-list positions carried by >= min_votes of the stored rows -> (pos, off[6]) in list order A,C,G,T,N."""
        r = np.ascontiguousarray(rows, dtype=np.int64)
        return self._consensus_call(lambda p, cap, off, npos: self._f_consensus(
            self._h, _ptr(r, _i64p), r.size, int(min_votes), p, cap, off, npos))

    def consensus_lists(self, pos_in, off_in, min_votes):
        """the same vote over un-stored samples (pos concatenated, off int64[6n+1] as for append_bulk)."""
        pos_in = _as_i32(pos_in)
        off_in = np.ascontiguousarray(off_in, dtype=np.int64)
        n = (off_in.size - 1) // 6
        if off_in.size != 6 * n + 1:
            raise ValueError("off must have 6*n+1 entries")
        return self._consensus_call(lambda p, cap, off, npos: self._f_consensus_lists(
            self._h, n, _ptr(pos_in, _i32p), _ptr(off_in, _i64p), int(min_votes), p, cap, off, npos))

    def msa_sites(self, rows, n_is_call=False):
        """ordered variant positions of the (valid) stored rows -> int32[].  n_is_call: the older rule of
        seqComparer_mt._msa, where an N also accounts for a position (FN3_SITES_CALLS_OR_N)."""
        r = np.ascontiguousarray(rows, dtype=np.int64)
        n_sites = C.c_int64(0)
        cap = 1 << 14
        while True:
            sites = np.empty(cap, dtype=np.int32)
            rc = self._f_msa_sites_rule(self._h, _ptr(r, _i64p), r.size, 1 if n_is_call else 0, _ptr(sites, _i32p), cap,
                                              C.byref(n_sites))
            if rc == N.FN3_ECAPACITY:
                cap = int(n_sites.value)
                continue
            self._check(rc)
            return sites[: n_sites.value].copy()

    def msa_align(self, rows, sites, want_codes=True):
        """-> (codes uint8[n, n_sites] or None, counts int32[n, 4] = |N|, |M|, N at sites, M at sites)."""
        r = np.ascontiguousarray(rows, dtype=np.int64)
        s = _as_i32(sites)
        counts = np.zeros((r.size, 4), dtype=np.int32)
        codes = np.zeros((r.size, s.size), dtype=np.uint8) if want_codes else None
        rc = self._f_msa_align(self._h, _ptr(r, _i64p), r.size, _ptr(s, _i32p), s.size,
                                     _ptr(codes, _u8p) if codes is not None and codes.size else None,
                                     _ptr(counts, _i32p))
        self._check(rc)
        return codes, counts

    # -- measurement ----------------------------------------------------------------------
    def profile_queries(self, query_rows, ceiling, flush_l2=False, count_bytes=False, member=None):
        """device-only timing of one-vs-all for each stored row in query_rows (bench.py's `value` leg); rows are those of
        ONE engine (`member`: the handle of a group member, default the first)."""
        q = np.ascontiguousarray(query_rows, dtype=np.int64)
        n = q.size
        ms_build = np.zeros(n, dtype=np.float32)
        ms_cmp = np.zeros(n, dtype=np.float32)
        n_hits = np.zeros(n, dtype=np.int64)
        touched = np.zeros(n, dtype=np.int64) if count_bytes else None
        passed = np.zeros(n, dtype=np.int64) if count_bytes else None
        rc = self._f_profile_queries(self._hl if member is None else member, _ptr(q, _i64p), n, int(ceiling),
                                     1 if flush_l2 else 0, _ptr(ms_build, _f32p), _ptr(ms_cmp, _f32p), _ptr(n_hits, _i64p),
                                     _ptr(touched, _i64p), _ptr(passed, _i64p))
        self._check(rc)
        return {"ms_build": ms_build, "ms_compare": ms_cmp, "n_hits": n_hits, "bytes_touched": touched, "n_passed": passed}

    def launch_count(self):
        return int(self._f_launch_count(self._h))


class GroupEngine(Engine):
    """An engine GROUP (include/fn3.h, fn3_group_*): the store sharded over several GPUs of one box, driven from THIS
    process — global row g lives on devices[g % G].  Same methods, same row numbering (global rows) and same results as
    `Engine`; every query fans out to one host thread per device inside libfn3 and the neighbour records are merged there.
    compress / pair_lists / profile_queries run on the first member engine."""

    def __init__(self, genome_length, devices):
        self._lib = N.load()
        devs = np.ascontiguousarray(list(devices), dtype=np.int32)
        if devs.size < 1:
            raise ValueError("an engine group needs at least one device")
        self._h = self._lib.fn3_group_create(int(genome_length), int(devs.size), _ptr(devs))
        if not self._h:
            msg = self._lib.fn3_last_error(None)
            raise EngineError("fn3_group_create failed: %s" % (msg.decode() if msg else "unknown"))
        self._hl = self._lib.fn3_group_engine(self._h, 0)
        self._bind("fn3_group_")
        self.genome_length = int(genome_length)
        self.devices = [int(d) for d in devs]
        self.device = self.devices[0]
        self.n_devices = len(self.devices)
        self.peer_access = bool(self._lib.fn3_group_peer_access(self._h))
        self._init_buffers()

    def member(self, i):
        """handle of the i-th member engine (bench.py: per-device profiling)"""
        return self._lib.fn3_group_engine(self._h, int(i))

    def all_vs_all(self, ceiling, upper_only=True, row_begin=0, row_end=None, stride=1, phase=0):
        """distmat over the sharded store: every device takes every live row as a query against its own shard (rows of
        other shards are read over NVLink).  Block-row arguments do not apply to a group (it already uses every device)."""
        if row_begin != 0 or row_end is not None or stride != 1 or phase != 0:
            raise ValueError("an engine group computes the whole matrix: no block-row arguments")
        with self._q_lock:
            n_out = C.c_int64(0)
            buf = self._edges
            rc = self._lib.fn3_group_all_vs_all(self._h, int(ceiling), 1 if upper_only else 0, _ptr(buf), buf.size, C.byref(n_out))
            if rc == N.FN3_ECAPACITY:
                buf = self._edges = np.empty(int(n_out.value), dtype=EDGE_DTYPE)
                rc = self._lib.fn3_group_all_vs_all(self._h, int(ceiling), 1 if upper_only else 0, _ptr(buf), buf.size,
                                                    C.byref(n_out))
            self._check(rc)
            return buf[: n_out.value].copy()


def make_engine(genome_length, device=0):
    """`device`: one CUDA ordinal -> Engine; a list / tuple of ordinals -> GroupEngine over those GPUs."""
    if isinstance(device, (list, tuple)):
        return GroupEngine(genome_length, device)
    return Engine(genome_length, device)
