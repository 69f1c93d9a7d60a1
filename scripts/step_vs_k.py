"""does the device step depend on how many steps are queued back to back, and on how the L2 is flushed between them?
(write sweep: the L2 is left full of dirty lines whose write-back runs under the next kernel; FN3_FLUSH_READ=1: read sweep)"""
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
    for flush_read in ("0", "1"):
        os.environ["FN3_FLUSH_READ"] = flush_read
        eng = Engine(coll.genome_length, 0)
        eng.append_bulk(coll.pos, coll.off, coll.invalid)
        rows = [eng.append(*q) for q in qs]
        eng.profile_queries(rows[:10], 20, flush_l2=True)
        for k in (20, 200, 400):
            for rep in range(2):
                pr = eng.profile_queries(rows[10:10 + k], 20, flush_l2=True)
                ms = (pr["ms_build"] + pr["ms_compare"]) * 1e3
                print("flush_read=%s K=%3d: step median %.2f us, first 20 median %.2f, last 20 median %.2f, compare median %.2f" % (
                    flush_read, k, np.median(ms), np.median(ms[:20]), np.median(ms[-20:]), 1e3 * np.median(pr["ms_compare"])), flush=True)
        pr = eng.profile_queries(rows[10:210], 20, flush_l2=False)
        print("flush_read=%s warm L2, K=200: compare median %.2f us" % (flush_read, 1e3 * np.median(pr["ms_compare"])), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
