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
