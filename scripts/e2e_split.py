"""where does the end-to-end insert+query time go?  (host pack / enqueue / wait, per call, and the Python overhead)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = ["bench.py"]
import numpy as np
import bench
from findneighbour3_b200.engine import Engine
common = bench.make_common()
coll = bench.make_shard(50000, 0, common)
qs = bench.make_queries(common, 700)
for zc in ("1", "0"):
    os.environ["FN3_ZERO_COPY"] = zc
    eng = Engine(coll.genome_length, 0)
    eng.append_bulk(coll.pos, coll.off, coll.invalid)
    for q in qs[:50]:
        eng.insert_query(q[0], q[1], q[2], 20)
    s0 = eng.stats()
    t0 = time.perf_counter()
    for q in qs[50:650]:
        eng.insert_query(q[0], q[1], q[2], 20)
    t = time.perf_counter() - t0
    s1 = eng.stats()
    n = s1["n_timed"] - s0["n_timed"]
    print("zero_copy", zc, "total %.1f us/call | pack %.1f enqueue %.1f wait %.1f | python+rest %.1f" % (
        1e6 * t / n, (s1["t_pack_ns"] - s0["t_pack_ns"]) / n / 1e3, (s1["t_enqueue_ns"] - s0["t_enqueue_ns"]) / n / 1e3,
        (s1["t_wait_ns"] - s0["t_wait_ns"]) / n / 1e3,
        1e6 * t / n - (s1["t_pack_ns"] - s0["t_pack_ns"] + s1["t_enqueue_ns"] - s0["t_enqueue_ns"] + s1["t_wait_ns"] - s0["t_wait_ns"]) / n / 1e3))
    eng.close()
