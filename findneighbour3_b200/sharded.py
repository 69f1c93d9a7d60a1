"""Candidate set sharded over the GPUs of one box: one process per GPU (torch.distributed), one engine per process.

SURVEY.md §8(e): every (query, candidate) pair is independent, so the only exchange is
  (1) the new sample's packed lists going from the rank that received it (rank 0, where the server's insert()
      runs) to every rank  — a broadcast of a few KB;
  (2) the thresholded neighbour records coming back — an all-gather of fixed-size blocks (count + records).
Both go through torch.distributed on the group's backend: NCCL over NVLink with device tensors on the GPU box,
gloo with host tensors in the CPU tests.  Sample i (global insertion order) lives on rank i % G as local row
i // G, which keeps the shards balanced under streaming inserts.

On ONE box the product path is the engine GROUP (include/fn3.h fn3_group_*, engine.GroupEngine): one process, one host
thread per device inside libfn3, no collective at all.  This module is the multi-PROCESS form — one rank per GPU, as
`torchrun` launches them, and the form that extends over several boxes.  (A second, shared-memory transport that lived
here in round 1 was removed: the group does the same job without leaving the process.)

The per-rank engine is duck-typed (append / insert_query / lists_vs_all / one_vs_all / row_get): the product
passes findneighbour3_b200.engine.Engine; the CPU-only gloo tests pass a stand-in built on the oracle.
"""
import numpy as np
import torch
import torch.distributed as dist

HIT_FIELDS = 5          # row, dist, n1, n2, nboth
HDR = 16                # header words of the sample message: n, off[0..6], invalid, ctl[0..2] (op, ceiling, argument), spare
OP_INSERT, OP_QUERY, OP_REMOVE, OP_STOP = 1, 2, 3, 4     # ctl[0] of a served round (ShardedSeqComparer); 0 = caller knows


class ShardedStore:
    def __init__(self, engine, group=None, max_hits=512, local_rows=0):
        self.engine = engine
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        backend = dist.get_backend(group) if dist.is_initialized() else "gloo"
        self.dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        self.max_hits = int(max_hits)
        self.bcast_cap = 1 << 13                 # positions carried by the single fixed-size broadcast (32 KB)
        self._bbuf = torch.zeros(HDR + self.bcast_cap, dtype=torch.int32)
        self._hblk = torch.zeros(1 + HIT_FIELDS * self.max_hits, dtype=torch.int32)
        self._dbuf = self._dhblk = self._dhall = self._hhall = None
        if self.dev.type == "cuda":
            self._bbuf = self._bbuf.pin_memory()
            self._hblk = self._hblk.pin_memory()
            self._dbuf = torch.zeros(HDR + self.bcast_cap, dtype=torch.int32, device=self.dev)
            self._dhblk = torch.zeros(1 + HIT_FIELDS * self.max_hits, dtype=torch.int32, device=self.dev)
            self._dhall = torch.zeros(self.world * (1 + HIT_FIELDS * self.max_hits), dtype=torch.int32, device=self.dev)
            self._hhall = torch.zeros(self.world * (1 + HIT_FIELDS * self.max_hits), dtype=torch.int32).pin_memory()
        self._bbuf_np = self._bbuf.numpy()
        # `local_rows`: rows the engine already holds (every rank the same number, placed as by load_shard)
        self._local_rows = int(local_rows)
        self.n_global = self._local_rows * self.world       # samples inserted so far (same value on every rank)

    def close(self):
        pass

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
    def _bcast_sample(self, sample, src=0, ctl=(0, 0, 0)):
        """`src`'s (pos, off, invalid) and three control words -> every rank, in ONE broadcast of a fixed-size buffer
        [n, off[0..6], invalid, ctl[0..2], spare | positions ...] (a second one only for samples beyond `self.bcast_cap`).
        -> (pos, off, invalid, ctl)"""
        if self.world == 1:
            pos, off, invalid = sample if sample is not None else (np.empty(0, np.int32), np.zeros(7, np.int32), True)
            return pos, off, invalid, tuple(int(x) for x in ctl)
        cap = self.bcast_cap
        buf = self._bbuf_np                       # numpy view of the (pinned) staging tensor
        if self.rank == src:
            pos, off, invalid = sample if sample is not None else (None, None, True)
            n = 0 if invalid else len(pos)
            buf[0] = n
            buf[1:8] = 0 if invalid else off
            buf[8] = 1 if invalid else 0
            buf[9:12] = ctl
            m = min(n, cap)
            if m:
                buf[HDR:HDR + m] = pos[:m]
        if self.dev.type == "cuda":
            self._dbuf.copy_(self._bbuf, non_blocking=True)
            dist.broadcast(self._dbuf, src=src, group=self.group)
            self._bbuf.copy_(self._dbuf, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        else:
            dist.broadcast(self._bbuf, src=src, group=self.group)
        n = int(buf[0])
        off = buf[1:8].astype(np.int32)
        invalid = bool(buf[8])
        ctl = (int(buf[9]), int(buf[10]), int(buf[11]))
        pos = buf[HDR:HDR + min(n, cap)].copy()
        if n > cap:                               # rare: a sample larger than the fixed buffer
            rest = torch.zeros(n - cap, dtype=torch.int32)
            if self.rank == src:
                rest[:] = torch.from_numpy(np.ascontiguousarray(sample[0][cap:], dtype=np.int32))
            rest = rest.to(self.dev)
            dist.broadcast(rest, src=src, group=self.group)
            pos = np.concatenate([pos, rest.cpu().numpy()])
        return pos, off, invalid, ctl

    def _exchange(self, sample, src, compute, ctl=(0, 0, 0)):
        """one round: the sample and `ctl` go from `src` to every rank, every rank runs
        `compute(pos, off, invalid, ctl) -> hits`, everybody receives all hits -> (dict, ctl).
        A rank whose compute() raises still takes part in the gather (with an error marker), so that every rank raises
        together instead of leaving the others waiting in the collective."""
        pos, off, invalid, ctl = self._bcast_sample(sample, src, ctl)
        err = None
        try:
            hits = compute(pos, off, invalid, ctl)
        except Exception as ex:                      # noqa: BLE001 - re-raised on every rank below
            err = ex
            hits = np.zeros(0, dtype=[("row", "<i8"), ("dist", "<i4"), ("n1", "<i4"), ("n2", "<i4"), ("nboth", "<i4")])
        out = self._gather_hits(hits, failed=err is not None)
        if err is not None:
            raise err
        return out, ctl

    def _all_gather_i64(self, t):
        if self.world == 1:
            return t.reshape(1, -1)
        t = t.to(self.dev)
        out = torch.empty(self.world * t.numel(), dtype=t.dtype, device=self.dev)
        dist.all_gather_into_tensor(out, t, group=self.group)
        return out.cpu().reshape(self.world, -1)

    def _gather_hits(self, hits, failed=False):
        """each rank's structured hits (local rows) -> dict {global id: (dist, n1, n2, nboth)} on every rank.
        One all-gather of fixed-size int32 blocks [count | row, dist, n1, n2, nboth ...] through pre-allocated
        (pinned / device) buffers; a second, larger one only if some rank overflowed its block.  A count of -1 marks
        a rank whose engine call failed: every rank then raises."""
        cap = self.max_hits
        n = -1 if failed else len(hits)
        while True:
            if cap == self.max_hits:
                blk, dblk, dall, hall = self._hblk, self._dhblk, self._dhall, self._hhall
            else:                                   # overflow round: temporary buffers
                blk = torch.zeros(1 + HIT_FIELDS * cap, dtype=torch.int32)
                dblk = dall = hall = None
            b = blk.numpy()
            b[0] = n
            m = min(max(n, 0), cap)
            if m:
                rec = b[1:1 + HIT_FIELDS * m].reshape(m, HIT_FIELDS)
                rec[:, 0] = hits["row"][:m]
                rec[:, 1] = hits["dist"][:m]
                rec[:, 2] = hits["n1"][:m]
                rec[:, 3] = hits["n2"][:m]
                rec[:, 4] = hits["nboth"][:m]
            if self.world == 1:
                allb = b.reshape(1, -1)
            elif self.dev.type == "cuda":
                if dblk is None:
                    dblk = torch.empty_like(blk, device=self.dev)
                    dall = torch.empty(self.world * blk.numel(), dtype=torch.int32, device=self.dev)
                    hall = torch.empty(self.world * blk.numel(), dtype=torch.int32).pin_memory()
                dblk.copy_(blk, non_blocking=True)
                dist.all_gather_into_tensor(dall, dblk, group=self.group)
                hall.copy_(dall, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                allb = hall.numpy().reshape(self.world, -1)
            else:
                hall = torch.empty(self.world * blk.numel(), dtype=torch.int32)
                dist.all_gather_into_tensor(hall, blk, group=self.group)
                allb = hall.numpy().reshape(self.world, -1)
            counts = allb[:, 0]
            if int(counts.min()) < 0:
                if failed:
                    return {}
                raise RuntimeError("the engine call of this round failed on rank(s) %s" % np.flatnonzero(counts < 0).tolist())
            if int(counts.max()) <= cap:
                break
            cap = int(counts.max())             # a rank overflowed its block: every rank repeats with room for all
        out = {}
        for r in range(self.world):
            c = int(counts[r])
            if c == 0:
                continue
            rec = allb[r, 1:1 + HIT_FIELDS * c].reshape(c, HIT_FIELDS)
            gids = (rec[:, 0].astype(np.int64) * self.world + r).tolist()
            out.update(zip(gids, map(tuple, rec[:, 1:].tolist())))
        return out

    # -- the hot path -----------------------------------------------------------------------
    def insert_query(self, sample, ceiling):
        """server insert(): rank 0 passes the new sample (pos, off, invalid), the others pass None.
        The sample is appended on its owner rank and compared against every earlier sample on all ranks.
        Returns (global id, {global id: (dist, n1, n2, nboth)}) on every rank."""
        gid = self.n_global
        own = self.owner(gid) == self.rank

        def compute(pos, off, invalid, ctl):
            if own:
                row, hits = self.engine.insert_query(pos, off, invalid, ceiling)
                assert row == self.local_row(gid) == self._local_rows, (row, gid, self._local_rows)
                self._local_rows += 1
                return hits
            return self.engine.lists_vs_all(pos, off, invalid, ceiling)

        out, _ = self._exchange(sample, 0, compute)
        self.n_global += 1
        return gid, out

    def one_vs_all(self, gid, ceiling):
        """mcompare of an already stored sample: its owner ships the row to everybody, then as above."""
        own = self.owner(gid) == self.rank
        if self.world == 1:
            return self._gather_hits(self.engine.one_vs_all(self.local_row(gid), ceiling))
        src = self.owner(gid)
        sample = self.engine.row_get(self.local_row(gid)) if own else None

        def compute(pos, off, invalid, ctl):
            if own:
                return self.engine.one_vs_all(self.local_row(gid), ceiling)
            return self.engine.lists_vs_all(pos, off, invalid, ceiling)

        return self._exchange(sample, src, compute)[0]

    # -- all-vs-all: replicated store, block rows by rank (no exchange until the final gather) ---
    @staticmethod
    def block_rows(n_rows, rank, world):
        """row ownership for the upper triangle: row r belongs to rank r % world (rows are interleaved, so long
        and short rows — row r has n-r-1 pairs — are spread evenly)."""
        return dict(row_begin=0, row_end=n_rows, stride=world, phase=rank)


class ShardedSeqComparer:
    """Guid-level face of a ShardedStore for BASELINE configs[4] (a collection resident across the GPUs of one box,
    streaming inserts each queried one-vs-all): the server process is rank 0 and calls insert / neighbours / mcompare /
    remove with guids and the reference's sample dicts; every other rank runs serve().

    Every call is ONE round of the store's exchange: rank 0 ships (operation, cut-off, sample lists) — it keeps the host
    dict of every sample, as the reference's seqProfile does, so a query on a stored sample needs no hop through its
    owner rank — every rank compares against its own shard, the thresholded neighbour records come back.
    Samples are plain compressed_sequence dicts ({'invalid': 1} included); the double-delta form stays a host-RAM matter
    of the single-GPU class."""

    def __init__(self, store, snpCeiling):
        self.store = store
        self.snpCeiling = snpCeiling
        self.seqProfile = {}                 # guid -> sample dict (rank 0)
        self._gid_of = {}                    # guid -> global id (rank 0)
        self._guid_of = {}                   # global id -> guid (rank 0)

    # -- one round, all ranks ---------------------------------------------------------------
    def _round(self, ctl=(0, 0, 0), sample=None):
        st = self.store
        gid_new = st.n_global

        def compute(pos, off, invalid, c):
            op, ceiling, arg = c
            if op == OP_INSERT:
                if st.owner(gid_new) == st.rank:
                    row, hits = st.engine.insert_query(pos, off, invalid, ceiling)
                    assert row == st.local_row(gid_new) == st._local_rows, (row, gid_new, st._local_rows)
                    st._local_rows += 1
                    return hits
                return st.engine.lists_vs_all(pos, off, invalid, ceiling)
            if op == OP_QUERY:
                return st.engine.lists_vs_all(pos, off, invalid, ceiling)
            if op == OP_REMOVE and st.owner(arg) == st.rank:
                st.engine.remove(st.local_row(arg))
            return np.zeros(0, dtype=[("row", "<i8"), ("dist", "<i4"), ("n1", "<i4"), ("n2", "<i4"), ("nboth", "<i4")])

        hits, c = st._exchange(sample, 0, compute, ctl)
        if c[0] == OP_INSERT:
            st.n_global += 1
        return hits, c

    def serve(self):
        """ranks >= 1: take part in every round until rank 0 calls shutdown()"""
        while True:
            _, c = self._round()
            if c[0] == OP_STOP:
                return

    # -- rank 0 -----------------------------------------------------------------------------
    @staticmethod
    def _ceil(cutoff):
        return max(-1, min(int(np.floor(cutoff)), 2 ** 31 - 1))

    def _named(self, hits, skip=None):
        return {self._guid_of[g]: v for g, v in hits.items() if g != skip and g in self._guid_of}

    def insert(self, guid, obj, cutoff=None):
        """persist + the one-vs-all of mcompare (findNeighbour3-server.py:516, :530) across all shards
        -> {guid2: (dist, n1, n2, nboth)} for the samples stored before."""
        from . import packing
        if guid in self._gid_of:
            self.remove(guid)
        sample = packing.pack_compressed(obj)
        gid = self.store.n_global
        hits, _ = self._round((OP_INSERT, self._ceil(self.snpCeiling if cutoff is None else cutoff), 0), sample)
        self._gid_of[guid] = gid
        self._guid_of[gid] = guid
        self.seqProfile[guid] = obj
        return self._named(hits, skip=gid)

    def neighbours(self, guid, cutoff=None):
        """one stored sample against all shards -> {guid2: (dist, n1, n2, nboth)}"""
        from . import packing
        if guid not in self._gid_of:
            raise KeyError("Asked to compare {0}  but guid requested has not been stored.".format(guid))
        gid = self._gid_of[guid]
        sample = packing.pack_compressed(self.seqProfile[guid])
        hits, _ = self._round((OP_QUERY, self._ceil(self.snpCeiling if cutoff is None else cutoff), gid), sample)
        return self._named(hits, skip=gid)                   # its own stored copy answers with distance 0: dropped

    def mcompare(self, guid, guids=None):
        """the reference's return shape (seqComparer.py:149-172): one 9-element list per other guid"""
        found = self.neighbours(guid)
        s1 = self.seqProfile[guid]
        out = []
        for key2 in (set(self.seqProfile.keys()) if guids is None else set(guids)):
            if key2 == guid:
                continue
            if key2 not in self.seqProfile:
                raise KeyError("Key1={0} does not exist in the in-memory store.".format(guid))
            hit = found.get(key2)
            if hit is None:
                out.append([guid, key2, None, None, None, None, None, None, None])
            else:
                s2 = self.seqProfile[key2]
                u1 = s1['N'] | set(s1['M'].keys())
                u2 = s2['N'] | set(s1['M'].keys())           # sic (seqComparer.py:417)
                out.append([guid, key2, hit[0], hit[1], hit[2], hit[3], u1, u2, u2 | u1])
        return out

    def remove(self, guid):
        """silent if absent (seqComparer.py:131-134); the owner rank tombstones the row"""
        gid = self._gid_of.pop(guid, None)
        if gid is None:
            return
        self._guid_of.pop(gid, None)
        self.seqProfile.pop(guid, None)
        self._round((OP_REMOVE, 0, gid), None)

    def iscachedinram(self, guid):
        return guid in self._gid_of

    def guidscachedinram(self):
        return set(self._gid_of.keys())

    def shutdown(self):
        """rank 0: release the serving ranks"""
        self._round((OP_STOP, 0, 0), None)
