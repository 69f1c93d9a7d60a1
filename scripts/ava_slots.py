"""all-vs-all wall time for several tile widths (FN3_SLOTS)"""
import os, sys, subprocess
code = r'''
import os, sys, time, numpy as np
sys.path.insert(0, ".")
sys.argv = ["bench.py"]
import bench
from findneighbour3_b200.engine import Engine
common = bench.make_common()
coll = bench.make_shard(50000, 0, common)
eng = Engine(coll.genome_length, 0)
eng.append_bulk(coll.pos, coll.off, coll.invalid)
eng.all_vs_all(12, row_end=4096)
t0 = time.perf_counter(); e = eng.all_vs_all(12); t = time.perf_counter() - t0
n = coll.n
print("FN3_SLOTS", os.environ.get("FN3_SLOTS"), "all-vs-all %.3f s  %.2e pairs/s  edges %d" % (t, n*(n-1)/2/t, len(e)))
'''
for slots in ("64", "128"):
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, FN3_SLOTS=slots))
