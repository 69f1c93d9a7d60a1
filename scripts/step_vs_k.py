"""what does the device step depend on: which queries, how many in the queue, head refresh on/off"""
import os, sys
import numpy as np
sys.path.insert(0, ".")
sys.argv = ["bench.py"]
import bench
from findneighbour3_b200.engine import Engine


def main():
    cpus = os.cpu_count()
    common = bench.make_common(cpus)
    coll = bench.make_shard(50000, 0, common, cpus)
    qs = bench.make_queries(common, 420)
    for auto in ("0", "1"):
        os.environ["FN3_BG_AUTO"] = auto
        for n_res in (25, 210, 420):
            eng = Engine(coll.genome_length, 0)
            eng.append_bulk(coll.pos, coll.off, coll.invalid)
            rows = [eng.append(*q) for q in qs[:n_res]]
            eng.profile_queries(rows[:5], 20, flush_l2=True)
            for lo, hi in ((5, 25), (10, 30), (10, 210), (210, 410)):
                if hi > n_res:
                    continue
                pr = eng.profile_queries(rows[lo:hi], 20, flush_l2=True)
                ms = (pr["ms_build"] + pr["ms_compare"]) * 1e3
                print("bg_auto=%s resident=%3d queries[%3d:%3d]: step median %.2f us compare median %.2f min %.2f hits %.1f refreshes %d" % (
                    auto, n_res, lo, hi, np.median(ms), 1e3 * np.median(pr["ms_compare"]), 1e3 * pr["ms_compare"].min(),
                    pr["n_hits"].mean(), eng.stats()["head_refreshes"]), flush=True)
            eng.close()


if __name__ == "__main__":
    main()
