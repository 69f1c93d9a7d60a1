"""profiling aid: k_query time with FN3_DEBUG_STOP = 0 / 1 (build only) / 2 (build + phase A)."""
import os, sys, json, subprocess
code = r'''
import os, sys, numpy as np
sys.path.insert(0, ".")
sys.argv = ["bench.py"]
import bench
from findneighbour3_b200.engine import Engine
common = bench.make_common()
coll = bench.make_shard(50000, 0, common)
qs = bench.make_queries(common, 120)
for stop in ("3", "4", "1", "2", "0"):
    os.environ["FN3_DEBUG_STOP"] = stop
    eng = Engine(coll.genome_length, 0)
    eng.append_bulk(coll.pos, coll.off, coll.invalid)
    rows = [eng.append(*q) for q in qs]
    eng.profile_queries(rows[:10], 20, flush_l2=True)
    for flush in (True, False):
        pr = eng.profile_queries(rows[10:], 20, flush_l2=flush)
        print("stop", stop, "flush", flush, "k_query us: mean %.2f  p50 %.2f  min %.2f" % (1e3 * pr["ms_compare"].mean(), 1e3 * np.median(pr["ms_compare"]), 1e3 * pr["ms_compare"].min()))
    if stop == "0":
        pr = eng.profile_queries(rows[10:40], 2**31 - 1, flush_l2=True)
        print("fullscan us: mean %.2f" % (1e3 * pr["ms_compare"].mean()))
        pr = eng.profile_queries(rows[10:40], 250, flush_l2=True)
        print("ceiling 250 us: mean %.2f" % (1e3 * pr["ms_compare"].mean()))
    eng.close()
'''
subprocess.run([sys.executable, "-c", code], check=True)
