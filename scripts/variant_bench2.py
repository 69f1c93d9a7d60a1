"""kernel experiment harness: one-vs-all k_query time (ceilings 20 / 250 / none) and all-vs-all wall time for every build of
libfn3 under variants/ (FN3_LIB_PATH), each in its own subprocess; the 50 000-sample collection is generated once."""
import glob
import os
import subprocess
import sys

import numpy as np

sys.path.insert(0, ".")
sys.argv = ["bench.py"]
import bench  # noqa: E402

NPZ = "/tmp/fn3_variant_store.npz"
common = bench.make_common()
coll = bench.make_shard(50000, 0, common)
qs = bench.make_queries(common, 70)
np.savez(NPZ, pos=coll.pos, off=coll.off, invalid=coll.invalid, L=coll.genome_length,
         qpos=np.concatenate([q[0] for q in qs]), qoff=np.stack([q[1] for q in qs]),
         qinv=np.array([q[2] for q in qs], np.uint8))

code = r'''
import os, sys, time, numpy as np
sys.path.insert(0, ".")
from findneighbour3_b200.engine import Engine
z = np.load("%s")
eng = Engine(int(z["L"]), 0)
eng.append_bulk(z["pos"], z["off"], z["invalid"])
rows, at = [], 0
for o, inv in zip(z["qoff"], z["qinv"]):
    n = int(o[6]); rows.append(eng.append(z["qpos"][at:at + n], o, bool(inv))); at += n
eng.profile_queries(rows[:10], 20, flush_l2=True)
out = []
for ceil in (20, 250, 2**31 - 1):
    pr = eng.profile_queries(rows[10:], ceil, flush_l2=True)
    out.append("ceil %%s: %%.2f us (hits %%d)" %% (ceil if ceil < 10**6 else "inf", 1e3 * pr["ms_compare"].mean(), int(pr["n_hits"].sum())))
eng.all_vs_all(12, row_end=4096)
best = 1e9
for _ in range(3):
    t0 = time.perf_counter(); e = eng.all_vs_all(12); best = min(best, time.perf_counter() - t0)
out.append("all-vs-all %%.4f s edges %%d" %% (best, len(e)))
print(os.path.basename(os.environ.get("FN3_LIB_PATH", "default")), " | ".join(out), flush=True)
''' % NPZ
for lib in sorted(glob.glob("variants/libfn3_*.so")):
    env = dict(os.environ, FN3_LIB_PATH=os.path.abspath(lib))
    subprocess.run([sys.executable, "-c", code], env=env)
