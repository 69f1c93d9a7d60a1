"""not-gpu: the N>1 host path (findneighbour3_b200/sharded.py) with world_size 2 over gloo on the CPU.
The per-rank engine is the oracle-backed stand-in of tests/fake_engine.py; what is under test is the sharding:
ownership i % G, broadcast of the new sample, all-gather of the neighbour blocks (incl. block overflow),
global id mapping, shard bulk load and queries on stored samples owned by either rank."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fake_engine import OracleEngine
        from findneighbour3_b200 import synth
        from findneighbour3_b200.sharded import ShardedStore
        from oracle import c_oracle

        L = 60000
        coll = synth.make_collection(4, 12, genome_length=L, seed=8, k1=80, s=2.0, n_blocks=4, n_mean=300, m_mean=3,
                                     rule="any", p_invalid=0.05)
        # (1) bulk load: rank r takes samples r, r+G, ... of the first 24
        base = 24
        mine = np.arange(rank, base, world)
        shard = coll.subset(mine)
        # a tiny block forces the overflow round of the gather
        store = ShardedStore(OracleEngine(L), max_hits=2)
        store.load_shard(shard.pos, shard.off, shard.invalid)
        assert store.n_global == base
        # (2) streaming inserts, sample known on rank 0 only
        results = {}
        for i in range(base, coll.n):
            sample = coll.lists(i) if rank == 0 else None
            gid, hits = store.insert_query(sample, 20)
            assert gid == i
            results[gid] = hits
        # (3) queries on stored samples owned by rank 0 and by rank 1
        stored = {g: store.one_vs_all(g, 20) for g in (0, 1, 6, 31)}
        # reference: one flat store in global order
        ok = True
        for gid, hits in results.items():
            sub = coll.subset(np.arange(gid + 1))
            want = c_oracle.one_vs_all(sub.pos, sub.off, sub.invalid, gid, 20)
            ok &= hits == want
        for g, hits in stored.items():
            want = c_oracle.one_vs_all(coll.pos, coll.off, coll.invalid, g, 20)
            ok &= hits == want
        q.put((rank, bool(ok), len(results), sum(len(h) for h in results.values())))
        store.close()
    finally:
        dist.destroy_process_group()


def test_sharded_store_world2_gloo():
    """the exchange over torch.distributed collectives (gloo here, NCCL on the GPU box)"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=90) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in out), out
    assert out[0][2] == out[1][2] == 24 and out[0][3] == out[1][3] > 0


def test_sharded_store_single_rank():
    sys.path.insert(0, HERE)
    from fake_engine import OracleEngine
    from findneighbour3_b200 import synth
    from findneighbour3_b200.sharded import ShardedStore
    from oracle import c_oracle
    coll = synth.make_collection(3, 8, genome_length=30000, seed=2, k1=50, s=2.0, n_blocks=3, n_mean=200, rule="any")
    store = ShardedStore(OracleEngine(30000))
    for i in range(coll.n):
        gid, hits = store.insert_query(coll.lists(i), 20)
        sub = coll.subset(np.arange(i + 1))
        assert gid == i and hits == c_oracle.one_vs_all(sub.pos, sub.off, sub.invalid, i, 20)
    assert ShardedStore.block_rows(10, 1, 4) == dict(row_begin=0, row_end=10, stride=4, phase=1)


def _class_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fake_engine import OracleEngine
        from findneighbour3_b200 import synth
        from findneighbour3_b200.sharded import ShardedSeqComparer, ShardedStore
        from oracle import seqcomparer_port as port_oracle

        L = 60000
        store = ShardedStore(OracleEngine(L), max_hits=256)
        sc = ShardedSeqComparer(store, snpCeiling=20)
        if rank != 0:
            sc.serve()                                   # until rank 0 shuts the service down
            q.put((rank, True, store.n_global))
            store.close()
            return
        coll = synth.make_collection(4, 10, genome_length=L, seed=18, k1=80, s=2.0, n_blocks=4, n_mean=300, m_mean=3,
                                     rule="any", p_invalid=0.05)
        flat = {}                                        # what one reference seqComparer would hold
        ok = True
        for i in range(coll.n):
            obj = coll.as_dict(i)
            got = sc.insert("g%d" % i, obj)
            flat["g%d" % i] = obj
            ok &= got == port_oracle.neighbours(flat, {}, "g%d" % i, 20)
        for g in ("g0", "g1", "g7", "g38"):              # stored samples owned by either rank
            ok &= sc.neighbours(g) == port_oracle.neighbours(flat, {}, g, 20)
            ok &= sc.neighbours(g, cutoff=3) == port_oracle.neighbours(flat, {}, g, 3)
        # the reference's mcompare shape, position sets included
        want = {r[1]: r[2:] for r in port_oracle.mcompare(flat, {}, "g5", 20)}
        got = {r[1]: r[2:] for r in sc.mcompare("g5")}
        ok &= got == want and len(got) == coll.n - 1
        # removal on both ranks; re-insert under the same guid
        for g in ("g2", "g3"):
            sc.remove(g)
            del flat[g]
        sc.remove("never-there")
        ok &= sc.neighbours("g4") == port_oracle.neighbours(flat, {}, "g4", 20)
        ok &= not sc.iscachedinram("g2") and sc.guidscachedinram() == set(flat)
        obj = coll.as_dict(2)
        got = sc.insert("g2", obj)
        flat["g2"] = obj
        ok &= got == port_oracle.neighbours(flat, {}, "g2", 20)
        try:
            sc.neighbours("nobody")
            ok = False
        except KeyError:
            pass
        sc.shutdown()
        q.put((rank, bool(ok), store.n_global))
        store.close()
    finally:
        dist.destroy_process_group()


def test_sharded_class_world2_gloo():
    """ShardedSeqComparer: rank 0 drives inserts / queries / removals by guid, rank 1 serves."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_class_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in out), out
    assert out[0][2] == out[1][2] == 41                  # 40 inserts + one re-insert, counted alike on both ranks


def test_sharded_class_single_rank():
    sys.path.insert(0, HERE)
    from fake_engine import OracleEngine
    from findneighbour3_b200 import synth
    from findneighbour3_b200.sharded import ShardedSeqComparer, ShardedStore
    from oracle import seqcomparer_port as port_oracle
    coll = synth.make_collection(3, 6, genome_length=30000, seed=4, k1=50, s=2.0, n_blocks=3, n_mean=200, m_mean=2, rule="any")
    sc = ShardedSeqComparer(ShardedStore(OracleEngine(30000)), snpCeiling=12)
    flat = {}
    for i in range(coll.n):
        flat[i] = coll.as_dict(i)
        assert sc.insert(i, flat[i]) == port_oracle.neighbours(flat, {}, i, 12)
    sc.remove(3)
    del flat[3]
    assert sc.neighbours(4) == port_oracle.neighbours(flat, {}, 4, 12)
    sc.shutdown()


def _big_and_bad_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fake_engine import OracleEngine
        from findneighbour3_b200 import synth
        from findneighbour3_b200.sharded import ShardedStore
        from oracle import c_oracle
        L = 400000
        # ~20 000 N positions per sample: far beyond the 8 192 positions of the fixed-size broadcast (MAXN_STORAGE is 130 000)
        coll = synth.make_collection(2, 5, genome_length=L, seed=3, k1=60, s=2.0, n_blocks=6, n_mean=20000, rule="any",
                                     p_invalid=0.0)
        store = ShardedStore(OracleEngine(L), max_hits=64)
        assert int(coll.positions_per_sample().max()) > store.bcast_cap
        ok = True
        for i in range(coll.n):
            gid, hits = store.insert_query(coll.lists(i) if rank == 0 else None, 20)
            sub = coll.subset(np.arange(i + 1))
            ok &= gid == i and hits == c_oracle.one_vs_all(sub.pos, sub.off, sub.invalid, i, 20)
        # an engine failure on ONE rank (the owner of the next row rejects the sample) must surface on EVERY rank, and the
        # ranks must stay in step afterwards
        owner = store.owner(store.n_global)
        real = store.engine.insert_query
        if rank == owner:
            def boom(*a, **k):
                raise ValueError("rejected by the engine")
            store.engine.insert_query = boom
        raised = False
        try:
            store.insert_query(coll.lists(0) if rank == 0 else None, 20)
        except (ValueError, RuntimeError):
            raised = True
        store.engine.insert_query = real
        gid, hits = store.insert_query(coll.lists(1) if rank == 0 else None, 20)      # same global id again: nothing was stored
        ok &= raised and gid == coll.n
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_samples_beyond_the_fixed_broadcast_and_failures_on_one_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_big_and_bad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        out = [q.get(timeout=120) for _ in procs]
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.kill()
    assert sorted(out) == [(0, True), (1, True)]
