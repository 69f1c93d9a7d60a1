"""profiling aid for k_scan (the STREAM variant): the 50 000-sample bench store, one-vs-all at ceilings 250 and none,
CUDA-event time per launch; FN3_LIB_PATH selects a build of libfn3 (variants/), FN3_DEBUG_STOP=1 stops after the query build."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
sys.argv = ["bench.py"]
import bench
from findneighbour3_b200.engine import Engine



def main():
    NPZ = "/tmp/fn3_scan_store.npz"
    if not os.path.exists(NPZ):
        common = bench.make_common(os.cpu_count())
        coll = bench.make_shard(50000, 0, common, os.cpu_count())
        qs = bench.make_queries(common, 40)
        np.savez(NPZ, pos=coll.pos, off=coll.off, invalid=coll.invalid, L=coll.genome_length,
                 qpos=np.concatenate([q[0] for q in qs]), qoff=np.stack([q[1] for q in qs]),
                 qinv=np.array([q[2] for q in qs], np.uint8))
    z = np.load(NPZ)
    eng = Engine(int(z["L"]), 0)
    eng.append_bulk(z["pos"], z["off"], z["invalid"])
    rows, at = [], 0
    for o, inv in zip(z["qoff"], z["qinv"]):
        n = int(o[6]); rows.append(eng.append(z["qpos"][at:at + n], o, bool(inv))); at += n
    eng.profile_queries(rows[:8], 250, flush_l2=True)
    out = []
    for ceil in (250, 2 ** 31 - 1):
        pr = eng.profile_queries(rows[8:], ceil, flush_l2=True, count_bytes=True)
        ms = np.sort(pr["ms_compare"])
        out.append("ceil %s: median %.2f us min %.2f max %.2f, %.1f MB moved -> %.2f TB/s (hits %d)" % (
            ceil if ceil < 10 ** 6 else "inf", 1e3 * np.median(ms), 1e3 * ms[0], 1e3 * ms[-1], pr["bytes_touched"].mean() / 1e6,
            pr["bytes_touched"].mean() / np.median(ms) / 1e9, int(pr["n_hits"].sum())))
    print(os.path.basename(os.environ.get("FN3_LIB_PATH", "default")), "stop=" + os.environ.get("FN3_DEBUG_STOP", "0"), " | ".join(out), flush=True)



if __name__ == "__main__":
    main()
