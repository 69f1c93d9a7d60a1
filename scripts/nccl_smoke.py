"""torchrun --nproc-per-node 2 scripts/nccl_smoke.py : the sharded host layer on real engines over NCCL and over the
shared-memory exchange — ShardedStore.insert_query / one_vs_all and ShardedSeqComparer — against the Python port."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from findneighbour3_b200 import synth  # noqa: E402
from findneighbour3_b200.engine import Engine  # noqa: E402
from findneighbour3_b200.sharded import ShardedSeqComparer, ShardedStore  # noqa: E402
from oracle import seqcomparer_port as port  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
L = 60000
coll = synth.make_collection(3, 10, genome_length=L, seed=18, k1=80, s=2.0, n_blocks=4, n_mean=300, m_mean=3, rule="any",
                             p_invalid=0.05)
ok = True
for exchange in ("nccl", "shm"):
    store = ShardedStore(Engine(L, local), max_hits=256, exchange=exchange)
    flat = {}
    for i in range(coll.n):
        gid, hits = store.insert_query(coll.lists(i) if rank == 0 else None, 20)
        flat[i] = coll.as_dict(i)
        ok &= gid == i and hits == port.neighbours(flat, {}, i, 20)
    for g in (0, 1, 7):
        ok &= store.one_vs_all(g, 20) == port.neighbours(flat, {}, g, 20)
    store.close()
    store = ShardedStore(Engine(L, local), max_hits=256, exchange=exchange)
    sc = ShardedSeqComparer(store, snpCeiling=20)
    if rank != 0:
        sc.serve()
    else:
        flat = {}
        for i in range(coll.n):
            flat["g%d" % i] = coll.as_dict(i)
            ok &= sc.insert("g%d" % i, flat["g%d" % i]) == port.neighbours(flat, {}, "g%d" % i, 20)
        sc.remove("g3")
        del flat["g3"]
        ok &= sc.neighbours("g4") == port.neighbours(flat, {}, "g4", 20)
        sc.shutdown()
    store.close()
    print("rank", rank, exchange, "ok" if ok else "MISMATCH", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
