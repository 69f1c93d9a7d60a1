"""all-vs-all wall time on the 200 000-sample configs[2] collection (or --n), FN3_CARVEOUT etc. from the environment"""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
from findneighbour3_b200 import synth
from findneighbour3_b200.engine import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
NPZ = "/tmp/fn3_ava_%d.npz" % n
if not os.path.exists(NPZ):
    c = synth.make_families([300000 + a for a in range(n // 100)], 100, workers=os.cpu_count())
    np.savez(NPZ, pos=c.pos, off=c.off, invalid=c.invalid)
z = np.load(NPZ)
eng = Engine(synth.TB_LENGTH, 0)
eng.append_bulk(z["pos"], z["off"], z["invalid"])
eng.all_vs_all(12, row_end=4096)
best = 1e9
for _ in range(3):
    t0 = time.perf_counter(); e = eng.all_vs_all(12); best = min(best, time.perf_counter() - t0)
print("all-vs-all n=%d carveout=%s: best %.4f s, %d edges" % (n, os.environ.get("FN3_CARVEOUT", "default"), best, len(e)), flush=True)
