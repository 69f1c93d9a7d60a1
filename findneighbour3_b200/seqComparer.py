"""Drop-in replacement for findNeighbour3's `seqComparer` class with the distance path on a B200.

    from findneighbour3_b200.seqComparer import seqComparer      # instead of: from seqComparer import seqComparer

Same constructor, methods, return shapes and exception types as the reference class
(src/seqComparer.py:31-1128; the callers are listed in SURVEY.md §8(b)), but:
  * every distance (countDifferences, countDifferences_byKey, mcompare, distmat,
    compress_relative_to_consensus' neighbour scan) is computed by the CUDA engine behind
    include/fn3.h — there is no Python/CPU distance code in this module;
  * the engine's HBM store always holds EXPANDED rows; the double-delta representation
    (patch + consensus, :503-548) remains a host-RAM concern of `seqProfile`, exactly as in the
    reference, and re-persisting identical content is a no-op for the device store;
  * the host keeps the reference-format dict of every sample (needed by load(), uncompress(),
    the MSA code and MongoDB pickling — SURVEY.md §8(b) "Ownership").

Deliberate deviation (documented in INTEGRATION.md): a malformed object is rejected at
persist() time instead of at the first comparison.
"""
import hashlib
import json
import math
import os
import pickle
import threading
import uuid
from collections import Counter
from collections.abc import MutableMapping, Sequence

import numpy as np

from . import packing
from .engine import Engine, make_engine

_BASES5 = ("A", "C", "T", "G", "N")


class _LazyPickle:
    """a sample still in its stored form (the reference's GridFS pickle); becomes a dict on first access"""
    __slots__ = ("blob", "invalid")

    def __init__(self, blob, invalid):
        self.blob = blob
        self.invalid = invalid


class _LazyLists:
    """a sample inserted through insert_sequence(): still the packed lists the device returned; becomes a dict on first access"""
    __slots__ = ("pos", "off", "m_chars", "invalid")

    def __init__(self, pos, off, m_chars, invalid):
        self.pos, self.off, self.m_chars, self.invalid = pos, off, m_chars, invalid


class _McompareRows(Sequence):
    """What mcompare() returns: one 9-element list per other guid (src/seqComparer.py:165-169), built when somebody looks.

    The reference materialises N lists per call — at 50 000 stored samples that alone is milliseconds, for rows that are
    almost all `[guid, guid2, None x 7]`.  This sequence keeps the snapshot of the other guids taken at call time and the
    (few) device hits; rows are created on indexing / iteration, survivors' position sets only for survivors.
    len(), indexing, slicing, iteration, `in`, == with a list, sorted(), list() behave as for the reference's list."""
    __slots__ = ("_sc", "_guid", "_keys", "_found", "_s1")

    def __init__(self, sc, guid, keys, found):
        self._sc, self._guid, self._keys, self._found, self._s1 = sc, guid, keys, found, None

    def __len__(self):
        return len(self._keys)

    def _hit_row(self, key2, hit):
        sc = self._sc
        if self._s1 is None:
            self._s1 = sc._computeComparator(sc.seqProfile[self._guid])
        s2 = sc._computeComparator(sc.seqProfile[key2])
        N1pos, N2pos, Nbothpos = sc._uncertain_sets(self._s1, s2)
        return [self._guid, key2, hit[0], hit[1], hit[2], hit[3], N1pos, N2pos, Nbothpos]

    def _row(self, key2):
        hit = self._found.get(key2)
        if hit is None:
            return [self._guid, key2, None, None, None, None, None, None, None]
        return self._hit_row(key2, hit)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._row(k) for k in self._keys[i]]
        return self._row(self._keys[i])

    def __iter__(self):
        guid, found = self._guid, self._found
        if not found:
            for k in self._keys:
                yield [guid, k, None, None, None, None, None, None, None]
            return
        get = found.get
        for k in self._keys:
            hit = get(k)
            yield [guid, k, None, None, None, None, None, None, None] if hit is None else self._hit_row(k, hit)

    def __eq__(self, other):
        if isinstance(other, (list, _McompareRows)):
            return len(self) == len(other) and all(a == b for a, b in zip(self, other))
        return NotImplemented

    def __repr__(self):
        return repr(list(self))


class _Profile(MutableMapping):
    """`seqProfile`: guid -> sample dict.  Assignments and deletions keep the device store in step.
    Samples loaded by persist_pickles() stay pickled until somebody asks for their dict."""

    def __init__(self, owner):
        self._owner = owner
        self._d = {}

    def __getitem__(self, k):
        v = self._d[k]
        if type(v) is _LazyPickle:
            v = pickle.loads(v.blob)
            self._d[k] = v
        elif type(v) is _LazyLists:
            v = self._owner._dict_from_lists(v.pos, v.off, v.m_chars, v.invalid)
            self._d[k] = v
        return v

    def __setitem__(self, k, v):
        self._owner._store_put(k, v)

    def __delitem__(self, k):
        if k not in self._d:
            raise KeyError(k)
        self._owner._store_del(k)

    def __iter__(self):
        return iter(self._d)

    def __len__(self):
        return len(self._d)

    def __contains__(self, k):
        return k in self._d

    def keys(self):
        return self._d.keys()

    def __repr__(self):
        return "_Profile(%d samples)" % len(self._d)


class seqComparer():
    def __init__(self, reference, maxNs, snpCeiling, debugMode=False, excludePositions=set(),
                 snpCompressionCeiling=250, cpuCount=1, device=None):
        """Same arguments as the reference (src/seqComparer.py:32-40); `device` is new: one CUDA ordinal, or a list of
        ordinals — the store is then sharded over those GPUs and this ONE object (the server constructs one,
        findNeighbour3-server.py:340-345) drives all of them (engine.GroupEngine).  Left out, it is read from the
        environment variable FN3_DEVICES ("0" by default; "0,1,2,3,4,5,6,7" hands the unmodified server eight GPUs).
        cpuCount is accepted and ignored, as in the reference (:65, :112)."""
        self.compressed_sequence_keys = set(['invalid', 'A', 'C', 'G', 'T', 'N', 'M'])
        self.patch_and_consensus_keys = set(['consensus_md5', 'patch'])
        self.patch_keys = set(['+', '-', 'M'])
        self.consensus_keys = set(['A', 'C', 'G', 'T', 'N'])
        self.snpCeiling = snpCeiling
        self.snpCompressionCeiling = snpCompressionCeiling
        self.maxNs = maxNs
        self.debugMode = debugMode

        self.reference = str(reference)
        letters = Counter(self.reference)
        if len(set(letters.keys()) - set(['A', 'C', 'G', 'T'])) > 0:
            raise TypeError("Reference sequence supplied contains characters other than ACTG: {0}".format(letters))

        self.excluded = excludePositions
        self._included = None
        self._ref_uploaded = False
        self._refresh()

        self.consensi = {}
        self.cpuCount = 1

        # device store: created on first use, so that the host-only methods (compress, hashes, patches) work
        # anywhere; anything that needs a distance needs the CUDA engine and fails loudly without it.
        if device is None:
            ids = [int(x) for x in os.environ.get("FN3_DEVICES", "0").split(",") if x.strip() != ""]
            device = ids[0] if len(ids) == 1 else ids
        self._device = device
        self._engine = None
        # persist() may run on another request thread while a query is in flight (server.py:513-516 is outside the
        # write semaphore): the guid<->row maps and the device store change together under this lock
        self._lock = threading.RLock()
        self._row_of = {}            # guid -> live row
        self._guid_of = {}           # row -> guid
        self._key_of = {}            # guid -> content key of what is in the device store
        self.seqProfile = _Profile(self)

    # ------------------------------------------------------------------------------------------
    # plumbing between the host dicts and the device store
    # ------------------------------------------------------------------------------------------
    @property
    def included(self):
        """positions taking part in comparisons (src/seqComparer.py:102); built on first use (4.4 M ints for TB)."""
        if self._included is None:
            self._included = set(range(len(self.reference))) - set(self.excluded)
        return self._included

    @property
    def engine(self):
        if self._engine is None:
            if len(self.reference) == 0:
                raise TypeError("reference cannot be of zero length")
            self._engine = make_engine(len(self.reference), self._device)     # raises if libfn3.so / a GPU is missing
        return self._engine

    def _engine_with_reference(self):
        """the engine, with the reference sequence and the mask uploaded (device compress and the MSA site rule need them)"""
        eng = self.engine
        with self._lock:
            if not self._ref_uploaded:
                eng.set_reference(self.reference, self.excluded)
                self._ref_uploaded = True
        return eng

    def _lists_of(self, obj):
        """any stored representation -> (pos, off, invalid) for the engine; raises like _computeComparator (:309-335)."""
        if not isinstance(obj, dict):
            raise TypeError("Cannot use object of class {0} as a sequence".format(type(obj)))
        if packing.is_invalid(obj):
            return packing.pack_compressed({'invalid': 1})
        keys = set(obj.keys())
        if keys == self.compressed_sequence_keys:
            return packing.pack_compressed(obj)
        if keys == self.patch_and_consensus_keys:
            return packing.pack_compressed(self.apply_patch(obj['patch'], self.consensi[obj['consensus_md5']]))
        raise KeyError("Was passed a dictionary with keys {0} but cannot handle this".format(obj.keys()))

    def _store_put(self, guid, obj):
        pos, off, invalid = self._lists_of(obj)
        key = packing.content_key(pos, off, invalid)
        with self._lock:
            if guid in self._row_of and self._key_of.get(guid) is None:       # loaded by persist_pickles: key not computed yet
                self._key_of[guid] = packing.content_key(*self._lists_of(self.seqProfile[guid]))
            if guid in self._row_of and self._key_of.get(guid) == key:
                self.seqProfile._d[guid] = obj   # same expanded content (e.g. re-encoded as a patch): host-only change
                return
            if guid in self._row_of:
                self._drop_row(guid)
            row = self.engine.append(pos, off, invalid)
            self._row_of[guid] = row
            self._guid_of[row] = guid
            self._key_of[guid] = key
            self.seqProfile._d[guid] = obj

    def _drop_row(self, guid):
        row = self._row_of.pop(guid)
        self._guid_of.pop(row, None)
        self._key_of.pop(guid, None)
        self.engine.remove(row)

    def _store_del(self, guid):
        with self._lock:
            self._drop_row(guid)
            del self.seqProfile._d[guid]

    def persist_many(self, objects, guids):
        """Bulk persist (server start-up reload, findNeighbour3-server.py:358-362): one device copy for all."""
        objects, guids = list(objects), list(guids)
        fresh = [(o, g) for o, g in zip(objects, guids)]
        packed = [self._lists_of(o) for o, _ in fresh]
        n = len(packed)
        off = np.zeros(6 * n + 1, dtype=np.int64)
        inv = np.zeros(n, dtype=np.uint8)
        total = 0
        for i, (p, f, bad) in enumerate(packed):
            inv[i] = 1 if bad else 0
            off[6 * i: 6 * i + 7] = total + f.astype(np.int64)
            total += int(f[6])
        pos = np.concatenate([p for p, _, _ in packed]) if total else np.empty(0, dtype=np.int32)
        with self._lock:                                 # the maps and the device store change together
            for _, g in fresh:
                if g in self._row_of:
                    self._drop_row(g)
            first = self.engine.append_bulk(pos, off, inv)
            for i, ((o, g), (p, f, bad)) in enumerate(zip(fresh, packed)):
                self._row_of[g] = first + i
                self._guid_of[first + i] = g
                self._key_of[g] = packing.content_key(p, f, bad)
                self.seqProfile._d[g] = o

    def persist_packed(self, guids, pos, off, invalid=None, m_chars=None):
        """Bulk persist of samples already in the engine's list format (what device compress returns and what
        fn3_unpickle_samples produces): n samples, `pos` concatenated, `off` int64[6n+1], `invalid` uint8[n] or None,
        `m_chars` the mixed-base characters in M-list order (bytes/str, or None: 'M').  One device copy; a sample's dict of
        sets is built only when `seqProfile[guid]` / `load(guid)` is asked for."""
        guids = list(guids)
        n = len(guids)
        off = np.ascontiguousarray(off, dtype=np.int64)
        if off.size != 6 * n + 1:
            raise ValueError("off must have 6*n+1 entries")
        inv = np.zeros(n, dtype=np.uint8) if invalid is None else np.ascontiguousarray(invalid, dtype=np.uint8)
        m_at = 0
        with self._lock:
            for g in guids:
                if g in self._row_of:
                    self._drop_row(g)
            first = self.engine.append_bulk(pos, off, inv)
            for i, g in enumerate(guids):
                o = off[6 * i: 6 * i + 7]
                o7 = (o - o[0]).astype(np.int32)
                n_m = int(o7[6] - o7[5])
                mc = ['M'] * n_m if m_chars is None else [chr(c) if isinstance(c, int) else c for c in m_chars[m_at:m_at + n_m]]
                m_at += n_m
                self._row_of[g] = first + i
                self._guid_of[first + i] = g
                self._key_of[g] = None
                self.seqProfile._d[g] = _LazyLists(pos[o[0]:o[6]], o7, mc, bool(inv[i]))

    def persist_pickles(self, blobs, guids):
        """Start-up reload straight from the stored form (findNeighbour3-server.py:358-362 reads GridFS pickles written by
        mongoStore.py:344-363, `pickle.dumps(obj, protocol=2)`): the pickles are parsed natively into the packed lists
        (fn3_unpickle_samples, host threads) and appended in one bulk copy — no dict of Python sets is built per
        sample; `seqProfile[guid]` unpickles a sample the first time it is asked for.  Pickles that are not plain
        compressed sequences (patch form) take the ordinary route."""
        blobs, guids = list(blobs), list(guids)
        n = len(blobs)
        pos, off, status = packing.unpickle_many(blobs)
        objs = [None] * n
        other = np.flatnonzero(status == 2)
        invalid = (status == 1).astype(np.uint8)
        if other.size:
            lens = np.diff(off).reshape(n, 6)
            pieces, at = [], 0
            for i in other:
                i = int(i)
                objs[i] = pickle.loads(blobs[i])
                p, f, bad = self._lists_of(objs[i])
                pieces.append(pos[at:off[6 * i]])
                pieces.append(p)
                at = off[6 * i]
                lens[i] = np.diff(f)
                invalid[i] = 1 if bad else 0
            pieces.append(pos[at:])
            pos = np.concatenate(pieces)
            off = np.concatenate([[0], np.cumsum(lens.ravel())]).astype(np.int64)
        with self._lock:
            for g in guids:
                if g in self._row_of:
                    self._drop_row(g)
            first = self.engine.append_bulk(pos, off, invalid)
            for i, g in enumerate(guids):
                self._row_of[g] = first + i
                self._guid_of[first + i] = g
                self._key_of[g] = None                   # computed if the guid is ever persisted again
                self.seqProfile._d[g] = objs[i] if objs[i] is not None else _LazyPickle(blobs[i], bool(invalid[i]))

    def compact_store(self):
        """Rebuild the device store without the rows remove() and re-persist have tombstoned (they keep their words in
        HBM and their slot in the head tiles: rows are never reused, include/fn3.h).  Maintenance call for a long-running
        server.  One engine: fn3_compact does it ON THE DEVICE (live rows copied to fresh chunks, heads rebuilt, rows
        renumbered) and only the guid<->row maps change here.  An engine group (or a stand-in engine without compact()):
        every live row is read back and bulk-appended to a fresh engine.  Returns (rows before, rows after)."""
        with self._lock:
            old = self.engine
            live = sorted(self._guid_of.keys())
            before = int(old.stats()["n_rows"])
            if len(live) == before:
                return before, before
            if getattr(old, "n_devices", 0) == 1 and hasattr(old, "compact"):
                new_row = old.compact()
                self._row_of = {g: int(new_row[r]) for g, r in self._row_of.items()}
                self._guid_of = {row: guid for guid, row in self._row_of.items()}
                return before, len(live)
            packed = [old.row_get(r) for r in live]
            n = len(packed)
            off = np.zeros(6 * n + 1, dtype=np.int64)
            inv = np.zeros(n, dtype=np.uint8)
            total = 0
            for i, (p, f, bad) in enumerate(packed):
                inv[i] = 1 if bad else 0
                off[6 * i: 6 * i + 7] = total + (0 if bad else f.astype(np.int64))
                total += 0 if bad else int(f[6])
            pos = np.concatenate([p for p, _, bad in packed if not bad]) if total else np.empty(0, dtype=np.int32)
            fresh = make_engine(len(self.reference), self._device)
            first = fresh.append_bulk(pos, off, inv) if n else 0
            self._engine = fresh
            self._ref_uploaded = False                       # uploaded again by the next compress / MSA call
            self._row_of = {self._guid_of[r]: first + i for i, r in enumerate(live)}
            self._guid_of = {row: guid for guid, row in self._row_of.items()}
            old.close()
            return before, n

    def _stored_kind(self, guid):
        """'invalid' | 'compressed' | 'patch' for a stored guid, without unpickling a lazily loaded sample"""
        v = self.seqProfile._d[guid]
        if type(v) is _LazyPickle or type(v) is _LazyLists:
            return 'invalid' if v.invalid else 'compressed'
        if 'invalid' in v and v['invalid'] == 1:
            return 'invalid'
        return 'patch' if set(v.keys()) == self.patch_and_consensus_keys else 'compressed'

    # ------------------------------------------------------------------------------------------
    # reference API: store
    # ------------------------------------------------------------------------------------------
    def raise_error(self, token):
        """raises a ZeroDivisionError, with token as the message (src/seqComparer.py:114-117)."""
        raise ZeroDivisionError(token)

    def persist(self, object, guid):
        """keeps a reference compressed object in RAM and appends it to the HBM store (:119-124)."""
        self.seqProfile[guid] = object

    def remove(self, guid):
        """removes a sample; silent if absent (:125-134)."""
        try:
            del self.seqProfile[guid]
        except KeyError:
            pass

    def load(self, guid):
        return self.seqProfile[guid]

    def _refresh(self):
        self._seq1 = None
        self._seq2 = None
        self.seq1md5 = None
        self.seq2md5 = None

    def iscachedinram(self, guid):
        return guid in self.seqProfile

    def guidscachedinram(self):
        return set(self.seqProfile.keys())

    def _guid(self):
        return str(uuid.uuid1())

    def _delta(self, x):
        return x[1] - x[0]

    def summarise_stored_items(self):
        """counts stored sequences by kind (:199-219)."""
        out = {'server|scstat|nSeqs': len(self.seqProfile), 'server|scstat|nConsensi': len(self.consensi),
               'server|scstat|nInvalid': 0, 'server|scstat|nCompressed': 0, 'server|scstat|nRecompressed': 0}
        for guid in list(self.seqProfile.keys()):
            kind = self._stored_kind(guid)               # (an invalid sample also counts as "compressed", :213-217)
            if kind == 'invalid':
                out['server|scstat|nInvalid'] += 1
            if kind == 'patch':
                out['server|scstat|nRecompressed'] += 1
            else:
                out['server|scstat|nCompressed'] += 1
        return out

    def excluded_hash(self):
        """'Excl {n} nt [{md5 of the sorted json list}]' (:240-248)."""
        ordered = sorted(list(self.excluded))
        h = hashlib.md5()
        h.update(json.dumps(ordered).encode('utf-8'))
        return "Excl {0} nt [{1}]".format(len(ordered), h.hexdigest())

    # ------------------------------------------------------------------------------------------
    # reference API: compress / uncompress
    # ------------------------------------------------------------------------------------------
    def compress(self, sequence):
        """string -> compressed_sequence dict (:274-307).  The byte compare against the reference under the
        exclusion mask and the compaction into six sorted lists run on the device (fn3_compress); only the
        final list -> Python set conversion (the reference's return type) happens here."""
        if not len(sequence) == len(self.reference):
            raise TypeError("sequence must of the same length as reference; seq is {0} and ref is {1}".format(
                len(sequence), len(self.reference)))
        if len(self.reference) == 0:
            raise TypeError("reference cannot be of zero length")
        eng = self._engine_with_reference()
        # one byte per character; anything outside latin-1 becomes '?', which (like every non-ACGTN-
        # character) is a mixed base; its real character is read back from `sequence` below (:296-298)
        pos, off, _ = eng.compress(sequence.encode('latin-1', 'replace'))
        invalid = int(off[6] - off[4]) > self.maxNs                # len(N) + len(M) > maxNs (:301)
        return self._dict_from_lists(pos, off, [sequence[p] for p in pos[off[5]:off[6]].tolist()], invalid)

    def _dict_from_lists(self, pos, off, m_chars, invalid):
        """the device's six sorted lists -> the reference's compressed_sequence dict (:290-305)"""
        if invalid:
            return {'invalid': 1}
        out = {'invalid': 0}
        for i, b in enumerate('ACGT'):
            out[b] = set(pos[off[i]:off[i + 1]].tolist())
        out['N'] = set(pos[off[4]:off[5]].tolist())
        out['M'] = dict(zip(pos[off[5]:off[6]].tolist(), m_chars))
        return out

    def insert_sequence(self, guid, sequence, cutoff=None):
        """findNeighbour3.insert()'s data path (findNeighbour3-server.py:511-538) in ONE device call: the character
        counts NucleicAcid.examine needs, compress (:274-307), the maxNs decision (:301), persist (:119-124) and the
        one-vs-all of mcompare (:149-172) from a single host->device copy of `sequence` (fn3_ingest_query).

        -> ({guid2: (dist, n1, n2, nboth)} for the neighbours within `cutoff` (default snpCeiling) among the samples
        stored before, 256 byte counts of the sequence for `NucleicAcid.examine(sequence, histogram=...)`).
        `seqProfile[guid]` / `load(guid)` build the sample's dict of sets only when somebody asks for it."""
        if not len(sequence) == len(self.reference):
            raise TypeError("sequence must of the same length as reference; seq is {0} and ref is {1}".format(
                len(sequence), len(self.reference)))
        if len(self.reference) == 0:
            raise TypeError("reference cannot be of zero length")
        if cutoff is None:
            cutoff = self.snpCeiling
        eng = self._engine_with_reference()
        with self._lock:
            if guid in self._row_of:
                self._drop_row(guid)
            row, invalid, pos, off, _, hits, hist = eng.ingest_query(sequence.encode('latin-1', 'replace'), self.maxNs,
                                                                     self._ceil(cutoff))
            self._row_of[guid] = row
            self._guid_of[row] = guid
            self._key_of[guid] = None                        # computed if the guid is ever persisted again
            self.seqProfile._d[guid] = _LazyLists(pos, off, [sequence[p] for p in pos[off[5]:off[6]].tolist()], invalid)
            gof = self._guid_of
            found = {gof[int(h['row'])]: (int(h['dist']), int(h['n1']), int(h['n2']), int(h['nboth'])) for h in hits}
        return found, hist

    def uncompress(self, compressed_sequence):
        """compressed (or patched) sample -> sequence string with N at excluded positions (:250-272)."""
        if 'invalid' in compressed_sequence.keys():
            if compressed_sequence['invalid'] == 1:
                raise ValueError("Cannot uncompress an invalid sequence, because the sequence it is not stored {0}".format(
                    compressed_sequence.keys()))
        cs = self._computeComparator(compressed_sequence)
        seq = list(self.reference)
        for x in self.excluded:
            seq[x] = 'N'
        for item in _BASES5:
            for x in cs[item]:
                seq[x] = item
        for x, ch in cs['M'].items():
            seq[x] = ch
        return ''.join(seq)

    def _computeComparator(self, sequence):
        """str | compressed_sequence | patch_and_consensus -> compressed_sequence (:309-335)."""
        if isinstance(sequence, str):
            return self.compress(sequence)
        elif isinstance(sequence, dict):
            try:
                if sequence['invalid'] == 1:
                    raise ValueError("Cannot uncompress an invalid sequence, as it is not stored. {0}".format(sequence.keys()))
            except KeyError:
                pass
            if set(sequence.keys()) == self.compressed_sequence_keys:
                return sequence
            elif set(sequence.keys()) == self.patch_and_consensus_keys:
                return self.apply_patch(sequence['patch'], self.consensi[sequence['consensus_md5']])
            else:
                raise KeyError("Was passed a dictionary with keys {0} but cannot handle this".format(sequence.keys()))
        else:
            raise TypeError("Cannot use object of class {0} as a sequence".format(type(sequence)))

    def setComparator1(self, sequence):
        try:
            self._seq1 = self._computeComparator(sequence)
        except ValueError:
            self._seq1 = None

    def setComparator2(self, sequence):
        try:
            self._seq2 = self._computeComparator(sequence)
        except ValueError:
            self._seq2 = None

    def _setStats(self, i1, i2):
        """sizes and union of two position sets / dict key sets (:354-380)."""
        r1 = i1 if isinstance(i1, set) else set(i1.keys())
        r2 = i2 if isinstance(i2, set) else set(i2.keys())
        both = r2 | r1
        return (len(r1), len(r2), len(both), r1, r2, both)

    # ------------------------------------------------------------------------------------------
    # reference API: the hot path — all arithmetic happens on the device
    # ------------------------------------------------------------------------------------------
    def countDifferences(self, cutoff=None):
        """distance between self._seq1 and self._seq2 (:424-458): None if either is invalid or > cutoff."""
        if cutoff is None:
            cutoff = self.snpCeiling
        if self._seq1 is None or self._seq2 is None:
            return None
        if self._seq1['invalid'] == 1 or self._seq2['invalid'] == 1:
            return None
        p1, o1, i1 = packing.pack_compressed(self._seq1)
        p2, o2, i2 = packing.pack_compressed(self._seq2)
        res = self.engine.pair_lists(p1, o1, i1, p2, o2, i2, self._ceil(cutoff))
        return None if res is None else res[0]

    @staticmethod
    def _ceil(cutoff):
        """ceilings arrive as ints or floats (1e8 in the reference's tests); the ABI takes int32."""
        c = int(np.floor(cutoff))
        return max(-1, min(c, 2 ** 31 - 1))

    def _uncertain_sets(self, s1, s2):
        """N1pos, N2pos, Nbothpos for a surviving pair (:416-418; NB the second uses seq1's M keys)."""
        u1 = s1['N'] | set(s1['M'].keys())
        u2 = s2['N'] | set(s1['M'].keys())
        return u1, u2, u2 | u1

    def countDifferences_byKey(self, keyPair, cutoff=None):
        """9-tuple (key1,key2,dist,n1,n2,nboth,N1pos,N2pos,Nbothpos) for two stored samples (:382-421)."""
        if not type(keyPair) is tuple:
            raise TypeError("Wanted tuple keyPair, but got keyPair={0} with type {1}".format(keyPair, type(keyPair)))
        if not len(keyPair) == 2:
            raise TypeError("Wanted a keyPair with two elements, but got {0}".format(keyPair))
        (key1, key2) = keyPair
        if key1 not in self.seqProfile:
            raise KeyError("Key1={0} does not exist in the in-memory store.".format(key1))
        if key2 not in self.seqProfile:
            raise KeyError("Key1={0} does not exist in the in-memory store.".format(key1))
        if cutoff is None:
            cutoff = self.snpCeiling
        self.setComparator1(self.seqProfile[key1])
        self.setComparator2(self.seqProfile[key2])
        res = self.engine.pair(self._row_of[key1], self._row_of[key2], self._ceil(cutoff))
        if res is None:
            return (key1, key2, None, None, None, None, None, None, None)
        dist, n1, n2, nboth = res
        N1pos, N2pos, Nbothpos = self._uncertain_sets(self._seq1, self._seq2)
        return (key1, key2, dist, n1, n2, nboth, N1pos, N2pos, Nbothpos)

    def neighbours(self, guid, guids=None, cutoff=None):
        """Fast form of mcompare: only the survivors, {guid2: (dist, n1, n2, nboth)} — what
        findNeighbour3-server.py:532-538 keeps of mcompare's output."""
        if guid not in self.seqProfile:
            raise KeyError("Asked to compare {0}  but guid requested has not been stored.  call .persist() on the "
                           "sample to be added before using mcompare.".format(guid))
        if cutoff is None:
            cutoff = self.snpCeiling
        rows = None
        if guids is not None:
            rows = []
            for g in set(guids):
                if g == guid:
                    continue
                if g not in self._row_of:
                    raise KeyError("Key1={0} does not exist in the in-memory store.".format(guid))
                rows.append(self._row_of[g])
            rows = np.asarray(rows, dtype=np.int64)
        with self._lock:
            hits = self.engine.one_vs_all(self._row_of[guid], self._ceil(cutoff), rows)
            gof = self._guid_of
            return {gof[int(h['row'])]: (int(h['dist']), int(h['n1']), int(h['n2']), int(h['nboth'])) for h in hits}

    def mcompare(self, guid, guids=None):
        """one guid against all (or `guids`), one 9-element list per other guid (:149-172).  The rows come back as a
        lazily materialised sequence (_McompareRows): the device call and a snapshot of the guids happen here, the
        per-row Python lists are built when the caller iterates or indexes."""
        found = self.neighbours(guid, guids, self.snpCeiling)
        keys = list(self.seqProfile._d) if guids is None else list(set(guids))
        try:
            keys.remove(guid)                    # the query itself is skipped (:166)
        except ValueError:
            pass
        return _McompareRows(self, guid, keys, found)

    def distmat(self, half=True, diagonal=False):
        """generator over the all-vs-all matrix at snpCompressionCeiling (:174-197)."""
        ceiling = self._ceil(self.snpCompressionCeiling)
        edges = self.engine.all_vs_all(ceiling, upper_only=False)
        table = {(int(e['row1']), int(e['row2'])): (int(e['dist']), int(e['n1']), int(e['n2']), int(e['nboth']))
                 for e in edges}
        keys = list(self.seqProfile.keys())
        for guid1 in keys:
            for guid2 in keys:
                include = False
                if not half and not guid1 == guid2:
                    include = True
                if guid1 == guid2 and diagonal:
                    include = True
                if guid1 < guid2 and half:
                    include = True
                if not include:
                    continue
                if guid1 == guid2:
                    yield self.countDifferences_byKey((guid1, guid2), cutoff=self.snpCompressionCeiling)
                    continue
                hit = table.get((self._row_of[guid1], self._row_of[guid2]))
                if hit is None:
                    yield (guid1, guid2, None, None, None, None, None, None, None)
                else:
                    s1 = self._computeComparator(self.seqProfile[guid1])
                    s2 = self._computeComparator(self.seqProfile[guid2])
                    N1pos, N2pos, Nbothpos = self._uncertain_sets(s1, s2)
                    yield (guid1, guid2, hit[0], hit[1], hit[2], hit[3], N1pos, N2pos, Nbothpos)

    # ------------------------------------------------------------------------------------------
    # reference API: double delta (host-RAM representation only; distances never see it)
    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _min_votes(n_objects, cutoff_proportion):
        """the reference keeps a position iff count >= len(list)*cutoff_proportion (:493-499); counts are
        integers >= 1, so that is count >= max(1, smallest integer >= the float product)."""
        return max(1, int(math.ceil(n_objects * cutoff_proportion)))

    @staticmethod
    def _consensus_dict(pos, off):
        """engine output (lists A,C,G,T,N) -> the reference's consensus object"""
        return {'A': set(pos[off[0]:off[1]].tolist()), 'C': set(pos[off[1]:off[2]].tolist()),
                'T': set(pos[off[3]:off[4]].tolist()), 'G': set(pos[off[2]:off[3]].tolist()),
                'N': set(pos[off[4]:off[5]].tolist())}

    def consensus(self, compressed_sequences, cutoff_proportion):
        """majority (>= cutoff_proportion) variant calls over a list of samples (:460-501).

        The per-position vote and the threshold run on the device (fn3_consensus_lists).  As in the
        reference, the vote counts the *given* objects' own A/C/G/T/N sets (objects in patch form
        contribute nothing, :485-486) and the threshold uses the full list length (:493)."""
        any_valid = False
        for cs in compressed_sequences:
            if self._computeComparator(cs)['invalid'] == 0:
                any_valid = True
        if not any_valid:
            return {'A': set(), 'C': set(), 'T': set(), 'G': set(), 'N': set()}
        voters = [cs for cs in compressed_sequences if set(cs.keys()) == self.compressed_sequence_keys]
        if len(voters) == 0:
            return {'A': set(), 'C': set(), 'T': set(), 'G': set(), 'N': set()}
        pos_in, off_in, _ = packing.pack_many(voters)
        pos, off = self.engine.consensus_lists(pos_in, off_in, self._min_votes(len(compressed_sequences), cutoff_proportion))
        return self._consensus_dict(pos, off)

    def _consensus_of_stored(self, guids, cutoff_proportion):
        """consensus() over stored samples without shipping their lists: the device votes on its own rows
        (fn3_consensus).  Same quirks: a guid listed twice votes twice, samples currently held in patch
        form do not vote, the threshold uses len(guids)."""
        with self._lock:
            rows = [self._row_of[g] for g in guids if self._stored_kind(g) == 'compressed']
        if len(rows) == 0:
            return {'A': set(), 'C': set(), 'T': set(), 'G': set(), 'N': set()}
        pos, off = self.engine.consensus(np.asarray(rows, dtype=np.int64), self._min_votes(len(guids), cutoff_proportion))
        return self._consensus_dict(pos, off)

    def generate_patch(self, compressed_sequence, consensus):
        """difference between a sample and a consensus; empty sets are not stored (:503-524)."""
        plus, minus = {}, {}
        for item in _BASES5:
            a = compressed_sequence[item] - consensus[item]
            s = consensus[item] - compressed_sequence[item]
            if len(a) > 0:
                plus[item] = a
            if len(s) > 0:
                minus[item] = s
        return {'+': plus, '-': minus, 'M': compressed_sequence['M']}

    def apply_patch(self, patch, consensus):
        """patch + consensus -> compressed_sequence (:526-548)."""
        if not patch.keys() == self.patch_keys:
            raise TypeError("Patch passed has wrong keys {0}".format(patch.keys))
        if not consensus.keys() == self.consensus_keys:
            raise TypeError("Consensus passed has wrong keys {0}".format(consensus.keys))
        out = {'invalid': 0}
        for item in _BASES5:
            out[item] = (consensus[item] | patch['+'].get(item, set())) - patch['-'].get(item, set())
        out['M'] = patch['M']
        return out

    def compressed_sequence_hash(self, compressed_sequence):
        """md5 over 'key:sorted-list;' for the sorted keys (:550-565)."""
        text = ""
        for key in sorted(compressed_sequence.keys()):
            v = compressed_sequence[key]
            text = text + key + ":" + str(sorted(list(v)) if isinstance(v, set) else v) + ';'
        h = hashlib.md5()
        h.update(text.encode('utf-8'))
        return h.hexdigest()

    def remove_unused_consensi(self):
        used = set()
        for guid in self.seqProfile.keys():
            if 'consensus_md5' in self.seqProfile[guid].keys():
                used.add(self.seqProfile[guid]['consensus_md5'])
        for md5 in set(self.consensi.keys()):
            if md5 not in used:
                del self.consensi[md5]

    def compress_relative_to_consensus(self, guid, cutoff_proportion=0.8):
        """re-encode guid and its neighbours within snpCompressionCeiling against their consensus (:579-615).

        The neighbour scan is ONE device one-vs-all at ceiling snpCompressionCeiling (the reference does N
        countDifferences_byKey calls, :584-592); a neighbour qualifies if dist < ceiling (:589).  The
        consensus vote runs on the device over the visited rows (fn3_consensus)."""
        visited_guids = [guid]
        visited_sequences = [self.seqProfile[guid]]
        found = self.neighbours(guid, None, self.snpCompressionCeiling)
        for other in self.seqProfile.keys():
            if other == guid:
                # the reference skips the seed by IDENTITY (`result[1] is not guid`, :588): an equal but distinct
                # string object passes the test and the seed is visited twice (its self distance is 0)
                if other is not guid and not packing.is_invalid(self.seqProfile[other]) \
                        and 0 < self.snpCompressionCeiling:
                    visited_sequences.append(self.seqProfile[other])
                    visited_guids.append(other)
            elif other in found and found[other][0] < self.snpCompressionCeiling:
                visited_sequences.append(self.seqProfile[other])
                visited_guids.append(other)
        if len(visited_sequences) > 1:
            consensus = self._consensus_of_stored(visited_guids, cutoff_proportion)
            consensus_md5 = self.compressed_sequence_hash(consensus)
            self.consensi[consensus_md5] = consensus
            for g in visited_guids:
                current = self.seqProfile[g]
                self.seqProfile[g] = {'patch': self.generate_patch(self._computeComparator(current), consensus),
                                      'consensus_md5': consensus_md5}
            self.remove_unused_consensi()
        return visited_guids

    # ------------------------------------------------------------------------------------------
    # reference API: mixPORE / MSA: sites, alignment and N/M tallies on the device, tests on the host
    # ------------------------------------------------------------------------------------------
    def estimate_expected_proportion(self, seqs):
        from .msa import estimate_expected_proportion
        return estimate_expected_proportion(seqs)

    def estimate_expected_unk(self, sample_size=30, exclude_guids=set(), unk_type='N'):
        from .msa import estimate_expected_unk
        return estimate_expected_unk(self, sample_size, exclude_guids, unk_type)

    def estimate_expected_unk_sites(self, sample_size=30, sites=set(), exclude_guids=set(), unk_type='N'):
        from .msa import estimate_expected_unk_sites
        return estimate_expected_unk_sites(self, sample_size, sites, exclude_guids, unk_type)

    def multi_sequence_alignment(self, guids, output='dict', sample_size=30, expected_p1=None, uncertain_base_type='N'):
        from .msa import multi_sequence_alignment
        return multi_sequence_alignment(self, guids, output, sample_size, expected_p1, uncertain_base_type)

    def _msa(self, valid_guids, invalid_guids, expected_p1, output, sample_size, uncertain_base_type='N'):
        from .msa import msa
        return msa(self, valid_guids, invalid_guids, expected_p1, output, sample_size, uncertain_base_type)
