"""kernel experiment harness: times k_query for several builds of libfn3 (FN3_LIB_PATH) in subprocesses."""
import os, sys, subprocess, glob
code = r'''
import os, sys, numpy as np
sys.path.insert(0, ".")
sys.argv = ["bench.py"]
import bench
from findneighbour3_b200.engine import Engine
common = bench.make_common()
coll = bench.make_shard(50000, 0, common)
qs = bench.make_queries(common, 60)
eng = Engine(coll.genome_length, 0)
eng.append_bulk(coll.pos, coll.off, coll.invalid)
rows = [eng.append(*q) for q in qs]
eng.profile_queries(rows[:10], 20, flush_l2=True)
out = []
for ceil in (20, 250, 2**31 - 1):
    pr = eng.profile_queries(rows[10:], ceil, flush_l2=True)
    out.append("ceil %s: %.2f us" % (ceil if ceil < 10**6 else "inf", 1e3 * pr["ms_compare"].mean()))
print(os.environ.get("FN3_LIB_PATH", "default"), " | ".join(out))
'''
for lib in sorted(glob.glob("variants/libfn3_*.so")):
    env = dict(os.environ, FN3_LIB_PATH=os.path.abspath(lib))
    subprocess.run([sys.executable, "-c", code], env=env)
