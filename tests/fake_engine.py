"""TEST-ONLY stand-in for findneighbour3_b200.engine.Engine built on the C oracle, so that the multi-rank host
logic (findneighbour3_b200/sharded.py) can be exercised with gloo on a CPU-only box."""
import numpy as np

from oracle import c_oracle

HIT_DTYPE = np.dtype([("row", "<i8"), ("dist", "<i4"), ("n1", "<i4"), ("n2", "<i4"), ("nboth", "<i4")])


class OracleEngine:
    def __init__(self, genome_length):
        self.genome_length = genome_length
        self.samples = []          # (pos, off, invalid)

    def _arrays(self, extra=None):
        items = self.samples + ([extra] if extra is not None else [])
        n = len(items)
        off = np.zeros(6 * n + 1, np.int64)
        inv = np.zeros(n, np.uint8)
        chunks, total = [], 0
        for i, (p, o, bad) in enumerate(items):
            inv[i] = 1 if bad else 0
            o = np.zeros(7, np.int64) if bad else np.asarray(o, np.int64)
            off[6 * i: 6 * i + 7] = total + o
            total += int(o[6])
            chunks.append(np.asarray(p, np.int32)[: int(o[6])])
        pos = np.concatenate(chunks) if chunks else np.empty(0, np.int32)
        return pos, off, inv

    @staticmethod
    def _hits(d):
        out = np.zeros(len(d), HIT_DTYPE)
        for i, (row, v) in enumerate(sorted(d.items())):
            out[i] = (row, v[0], v[1], v[2], v[3])
        return out

    def append(self, pos, off, invalid=False):
        self.samples.append((np.array(pos, np.int32), np.array(off, np.int32), bool(invalid)))
        return len(self.samples) - 1

    def append_bulk(self, pos, off, invalid=None):
        first = len(self.samples)
        n = (len(off) - 1) // 6
        for i in range(n):
            o = np.asarray(off[6 * i: 6 * i + 7], np.int64)
            self.append(pos[o[0]:o[6]], (o - o[0]).astype(np.int32), bool(invalid[i]) if invalid is not None else False)
        return first

    def row_get(self, row):
        return self.samples[row]

    def one_vs_all(self, row, ceiling, rows=None):
        pos, off, inv = self._arrays()
        return self._hits(c_oracle.one_vs_all(pos, off, inv, row, ceiling))

    def lists_vs_all(self, pos, off, invalid, ceiling, rows=None):
        p, o, inv = self._arrays((np.array(pos, np.int32), np.array(off, np.int32), bool(invalid)))
        return self._hits(c_oracle.one_vs_all(p, o, inv, len(self.samples), ceiling))

    def insert_query(self, pos, off, invalid, ceiling):
        row = self.append(pos, off, invalid)
        return row, self.one_vs_all(row, ceiling)

    # ---- the rest of the Engine face, so that the drop-in seqComparer CLASS can be driven on a CPU-only box
    # (tests/test_reference_unittests.py: the reference's own unit tests against the class; distances from the C oracle,
    # compress / consensus / MSA pieces from the set-based ports).  Nothing in the product imports this. ----
    def remove(self, row):
        if row < 0 or row >= len(self.samples):
            raise KeyError(row)
        self.samples[row] = (np.empty(0, np.int32), np.zeros(7, np.int32), True)
        self._removed = getattr(self, "_removed", set()) | {row}

    def _one_vs(self, row, ceiling, rows=None, keep_self=False):
        pos, off, inv = self._arrays()
        d = c_oracle.one_vs_all(pos, off, inv, row, ceiling)
        if rows is not None:
            keep = set(int(r) for r in rows)
            d = {r: v for r, v in d.items() if r in keep}
        return d

    def one_vs_all(self, row, ceiling, rows=None):  # noqa: F811  (supersedes the minimal version above)
        if row in getattr(self, "_removed", set()):
            raise KeyError(row)
        return self._hits(self._one_vs(row, ceiling, rows))

    def pair(self, row1, row2, ceiling):
        p1, o1, i1 = self.samples[row1]
        p2, o2, i2 = self.samples[row2]
        return c_oracle.pair(p1, o1, i1, p2, o2, i2, ceiling)

    def pair_lists(self, pos1, off1, invalid1, pos2, off2, invalid2, ceiling):
        return c_oracle.pair(pos1, off1, invalid1, pos2, off2, invalid2, ceiling)

    def all_vs_all(self, ceiling, upper_only=True, row_begin=0, row_end=None, stride=1, phase=0):
        edge = np.dtype([("row1", "<i8"), ("row2", "<i8"), ("dist", "<i4"), ("n1", "<i4"), ("n2", "<i4"), ("nboth", "<i4")])
        out = []
        n = len(self.samples)
        for q in range(row_begin, n if row_end is None else row_end):
            if q % stride != phase or self.samples[q][2]:
                continue
            for r, v in self._one_vs(q, ceiling).items():
                if not upper_only or q < r:
                    out.append((q, r) + tuple(v))
        return np.array(out, dtype=edge) if out else np.zeros(0, edge)

    def stats(self):
        return {"n_rows": len(self.samples)}

    def launch_count(self):
        return 0

    def close(self):
        pass

    def set_reference(self, reference, excluded=None):
        self._ref = reference if isinstance(reference, str) else bytes(reference).decode("latin-1")
        self._excluded = set(excluded or ())

    def compress(self, seq_bytes):
        from oracle import seqcomparer_port as port
        seq = seq_bytes.decode("latin-1")
        included = [i for i in range(len(self._ref)) if i not in self._excluded]
        d = port.compress(self._ref, included, float("inf"), seq)
        lists = [sorted(d[k]) for k in "ACGTN"] + [sorted(d["M"].keys())]
        off = np.zeros(7, np.int32)
        off[1:] = np.cumsum([len(x) for x in lists])
        pos = np.array([p for lst in lists for p in lst], dtype=np.int32)
        return pos, off, "".join(d["M"][p] for p in lists[5]).encode("latin-1")

    def ingest_query(self, seq_bytes, max_ns, ceiling, want_hist=True):
        pos, off, m_chars = self.compress(seq_bytes)
        invalid = int(off[6] - off[4]) > max_ns
        row, hits = self.insert_query(pos, off, invalid, ceiling)
        hist = np.bincount(np.frombuffer(seq_bytes, np.uint8), minlength=256).astype(np.int64) if want_hist else None
        return row, invalid, pos, off, m_chars, hits, hist

    def _dict_of(self, row):
        p, o, bad = self.samples[row]
        if bad:
            return None
        d = {k: set(int(x) for x in p[o[b]:o[b + 1]]) for b, k in enumerate("ACGTN")}
        d["M"] = {int(x): "M" for x in p[o[5]:o[6]]}
        d["invalid"] = 0
        return d

    @staticmethod
    def _vote(dicts, min_votes):
        lists = []
        for k in "ACGTN":
            votes = {}
            for d in dicts:
                for p in d[k]:
                    votes[p] = votes.get(p, 0) + 1
            lists.append(sorted(p for p, v in votes.items() if v >= min_votes))
        off = np.zeros(6, np.int32)
        off[1:] = np.cumsum([len(x) for x in lists])
        return np.array([p for lst in lists for p in lst], dtype=np.int32), off

    def consensus(self, rows, min_votes):
        return self._vote([d for d in (self._dict_of(int(r)) for r in rows) if d is not None], min_votes)

    def consensus_lists(self, pos_in, off_in, min_votes):
        n = (len(off_in) - 1) // 6
        dicts = []
        for i in range(n):
            o = off_in[6 * i: 6 * i + 7]
            dicts.append({k: set(int(x) for x in pos_in[o[b]:o[b + 1]]) for b, k in enumerate("ACGTN")})
        return self._vote(dicts, min_votes)

    def msa_sites(self, rows, n_is_call=False):
        from oracle import cluster_port as cp
        dicts = [self._dict_of(int(r)) for r in rows]
        if any(d is None for d in dicts):
            raise ValueError("invalid sample")
        return np.array(cp.variant_positions(self._ref, dicts, n_is_call), dtype=np.int32)

    def msa_align(self, rows, sites, want_codes=True):
        from oracle import cluster_port as cp
        code = {"A": 1, "C": 2, "G": 3, "T": 4, "N": 5, "M": 6}
        sites = [int(s) for s in sites]
        counts = np.zeros((len(rows), 4), np.int32)
        codes = np.zeros((len(rows), len(sites)), np.uint8) if want_codes else None
        for j, r in enumerate(rows):
            d = self._dict_of(int(r))
            if d is None:
                raise ValueError("invalid sample")
            if want_codes:
                aligned = cp.aligned_string("." * (max(sites) + 1 if sites else 0), d, sites)
                codes[j] = [code.get(ch, 0) for ch in aligned]
                counts[j] = cp.tallies(d, aligned)
            else:
                counts[j] = (len(d["N"]), len(d["M"]), cp.unknowns_at_sites(d, set(sites), "N"),
                             cp.unknowns_at_sites(d, set(sites), "M"))
        return codes, counts
