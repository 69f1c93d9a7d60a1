"""Candidate set sharded over the GPUs of one box: one process per GPU (torch.distributed), one engine per process.

SURVEY.md §8(e): every (query, candidate) pair is independent, so the only exchange is
  (1) the new sample's packed lists going from the rank that received it (rank 0, where the server's insert()
      runs) to every rank  — a broadcast of a few KB;
  (2) the thresholded neighbour records coming back — an all-gather of fixed-size blocks (count + records).
Both go through torch.distributed on the group's backend: NCCL over NVLink with device tensors on the GPU box,
gloo with host tensors in the CPU tests.  Sample i (global insertion order) lives on rank i % G as local row
i // G, which keeps the shards balanced under streaming inserts.

The per-rank engine is duck-typed (append / insert_query / lists_vs_all / one_vs_all / row_get): the product
passes findneighbour3_b200.engine.Engine; the CPU-only gloo tests pass a stand-in built on the oracle.
"""
import numpy as np
import torch
import torch.distributed as dist

HIT_FIELDS = 5          # row, dist, n1, n2, nboth


class ShardedStore:
    def __init__(self, engine, group=None, max_hits=1024, local_rows=0):
        self.engine = engine
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        backend = dist.get_backend(group) if dist.is_initialized() else "gloo"
        self.dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        self.max_hits = int(max_hits)
        # `local_rows`: rows the engine already holds (every rank the same number, placed as by load_shard)
        self._local_rows = int(local_rows)
        self.n_global = self._local_rows * self.world       # samples inserted so far (same value on every rank)

    # -- id mapping -------------------------------------------------------------------------
    def owner(self, gid):
        return gid % self.world

    def local_row(self, gid):
        return gid // self.world

    def global_id(self, rank, row):
        return row * self.world + rank

    # -- bulk load (server start-up): every rank appends its own shard, no exchange ---------
    def load_shard(self, pos, off, invalid):
        """append this rank's shard; all ranks must hold the same number of samples."""
        first = self.engine.append_bulk(pos, off, invalid)
        n = (len(off) - 1) // 6
        assert first == self._local_rows
        self._local_rows += n
        counts = self._all_gather_i64(torch.tensor([n], dtype=torch.int64))
        if len(set(int(c) for c in counts.flatten())) != 1:
            raise ValueError("load_shard: shards must be of equal size (got %s)" % counts.flatten().tolist())
        self.n_global += n * self.world
        return first

    # -- collectives ------------------------------------------------------------------------
    def _bcast_sample(self, sample):
        """rank 0's (pos, off, invalid) -> every rank.  Two broadcasts: an 9-int header, then the positions."""
        if self.world == 1:
            return sample
        head = torch.zeros(9, dtype=torch.int32)
        if self.rank == 0:
            pos, off, invalid = sample
            head[0] = 0 if invalid else len(pos)
            head[1:8] = torch.from_numpy(np.asarray(off, dtype=np.int32)) if not invalid else 0
            head[8] = 1 if invalid else 0
        head = head.to(self.dev)
        dist.broadcast(head, src=0, group=self.group)
        head = head.cpu()
        n = int(head[0])
        payload = torch.zeros(max(n, 1), dtype=torch.int32)
        if self.rank == 0 and n:
            payload[:n] = torch.from_numpy(np.ascontiguousarray(sample[0], dtype=np.int32))
        payload = payload.to(self.dev)
        dist.broadcast(payload, src=0, group=self.group)
        pos = payload.cpu().numpy()[:n].copy()
        return pos, head[1:8].numpy().astype(np.int32), bool(head[8])

    def _all_gather_i64(self, t):
        if self.world == 1:
            return t.reshape(1, -1)
        t = t.to(self.dev)
        out = torch.empty(self.world * t.numel(), dtype=t.dtype, device=self.dev)
        dist.all_gather_into_tensor(out, t, group=self.group)
        return out.cpu().reshape(self.world, -1)

    def _gather_hits(self, hits):
        """each rank's structured hits (local rows) -> dict {global id: (dist, n1, n2, nboth)} on every rank."""
        cap = self.max_hits
        while True:
            block = torch.zeros(1 + HIT_FIELDS * cap, dtype=torch.int64)
            n = len(hits)
            block[0] = n
            m = min(n, cap)
            if m:
                rec = np.stack([hits["row"][:m].astype(np.int64), hits["dist"][:m].astype(np.int64),
                                hits["n1"][:m].astype(np.int64), hits["n2"][:m].astype(np.int64),
                                hits["nboth"][:m].astype(np.int64)], axis=1)
                block[1:1 + HIT_FIELDS * m] = torch.from_numpy(rec.reshape(-1))
            allb = self._all_gather_i64(block)
            counts = allb[:, 0]
            if int(counts.max()) <= cap:
                break
            cap = int(counts.max())             # a rank overflowed its block: every rank repeats with room for all
        out = {}
        for r in range(self.world):
            c = int(counts[r])
            rec = allb[r, 1:1 + HIT_FIELDS * c].reshape(c, HIT_FIELDS).numpy()
            for row, d, n1, n2, nb in rec:
                out[self.global_id(r, int(row))] = (int(d), int(n1), int(n2), int(nb))
        return out

    # -- the hot path -----------------------------------------------------------------------
    def insert_query(self, sample, ceiling):
        """server insert(): rank 0 passes the new sample (pos, off, invalid), the others pass None.
        The sample is appended on its owner rank and compared against every earlier sample on all ranks.
        Returns (global id, {global id: (dist, n1, n2, nboth)}) on every rank."""
        pos, off, invalid = self._bcast_sample(sample)
        gid = self.n_global
        if self.owner(gid) == self.rank:
            row, hits = self.engine.insert_query(pos, off, invalid, ceiling)
            assert row == self.local_row(gid) == self._local_rows, (row, gid, self._local_rows)
            self._local_rows += 1
        else:
            hits = self.engine.lists_vs_all(pos, off, invalid, ceiling)
        self.n_global += 1
        return gid, self._gather_hits(hits)

    def one_vs_all(self, gid, ceiling):
        """mcompare of an already stored sample: its owner ships the row to everybody, then as above."""
        own = self.owner(gid) == self.rank
        if self.world == 1:
            return self._gather_hits(self.engine.one_vs_all(self.local_row(gid), ceiling))
        # route through rank 0's broadcast: the owner first hands the lists to rank 0 (or is rank 0)
        sample = None
        src = self.owner(gid)
        if src == 0:
            if own:
                sample = self.engine.row_get(self.local_row(gid))
        else:
            head = torch.zeros(9, dtype=torch.int32)
            if own:
                p, o, inv = self.engine.row_get(self.local_row(gid))
                head[0] = len(p)
                head[1:8] = torch.from_numpy(np.asarray(o, dtype=np.int32))
                head[8] = 1 if inv else 0
            head = head.to(self.dev)
            dist.broadcast(head, src=src, group=self.group)
            head = head.cpu()
            n = int(head[0])
            payload = torch.zeros(max(n, 1), dtype=torch.int32)
            if own and n:
                payload[:n] = torch.from_numpy(np.ascontiguousarray(p, dtype=np.int32))
            payload = payload.to(self.dev)
            dist.broadcast(payload, src=src, group=self.group)
            if self.rank == 0:
                sample = (payload.cpu().numpy()[:n].copy(), head[1:8].numpy().astype(np.int32), bool(head[8]))
        pos, off, invalid = self._bcast_sample(sample)
        if own:
            hits = self.engine.one_vs_all(self.local_row(gid), ceiling)
        else:
            hits = self.engine.lists_vs_all(pos, off, invalid, ceiling)
        return self._gather_hits(hits)

    # -- all-vs-all: replicated store, block rows by rank (no exchange until the final gather) ---
    @staticmethod
    def block_rows(n_rows, rank, world):
        """row ownership for the upper triangle: row r belongs to rank r % world (rows are interleaved, so long
        and short rows — row r has n-r-1 pairs — are spread evenly)."""
        return dict(row_begin=0, row_end=n_rows, stride=world, phase=rank)
