"""A/B timing of libfn3 builds (FN3_LIB_PATH selects one of variants/): on the 50 000-sample bench store
   FILTER step (ceiling 20, L2 flushed and warm), STREAM at ceilings 250 / none, all-vs-all at 12 SNV, e2e insert+query;
   the hit counts printed with every figure must agree between builds."""
import os, sys, time, statistics
import numpy as np
sys.path.insert(0, ".")
sys.argv = ["bench.py"]
import bench
from findneighbour3_b200.engine import Engine


def main():
    NPZ = "/tmp/fn3_variant_store.npz"
    if not os.path.exists(NPZ):
        common = bench.make_common(os.cpu_count())
        coll = bench.make_shard(50000, 0, common, os.cpu_count())
        qs = bench.make_queries(common, 200)
        np.savez(NPZ, pos=coll.pos, off=coll.off, invalid=coll.invalid, L=coll.genome_length,
                 qpos=np.concatenate([q[0] for q in qs]), qoff=np.stack([q[1] for q in qs]),
                 qinv=np.array([q[2] for q in qs], np.uint8))
    z = np.load(NPZ)
    eng = Engine(int(z["L"]), 0)
    eng.append_bulk(z["pos"], z["off"], z["invalid"])
    samples, at = [], 0
    for o, inv in zip(z["qoff"], z["qinv"]):
        n = int(o[6]); samples.append((z["qpos"][at:at + n], o, bool(inv))); at += n
    rows = [eng.append(*s) for s in samples[:100]]
    name = os.path.basename(os.environ.get("FN3_LIB_PATH", "default"))
    out = []
    eng.profile_queries(rows[:8], 20, flush_l2=True)
    for flush in (True, False):
        pr = eng.profile_queries(rows[8:], 20, flush_l2=flush)
        ms = 1e3 * (pr["ms_build"] + pr["ms_compare"])
        out.append("filter %s: median %.2f us p10 %.2f p90 %.2f (hits %d)" % (
            "cold" if flush else "warm", np.median(ms), np.percentile(ms, 10), np.percentile(ms, 90), int(pr["n_hits"].sum())))
    eng.profile_queries(rows[:8], 250, flush_l2=True)
    for ceil in (250, 2 ** 31 - 1):
        pr = eng.profile_queries(rows[8:40], ceil, flush_l2=True)
        ms = 1e3 * pr["ms_compare"]
        out.append("scan %s: median %.2f us min %.2f (hits %d)" % (
            ceil if ceil < 10 ** 6 else "inf", np.median(ms), ms.min(), int(pr["n_hits"].sum())))
    eng.all_vs_all(12, row_end=4096)
    ts = []
    for rep in range(3):
        t0 = time.perf_counter(); e = eng.all_vs_all(12); ts.append(time.perf_counter() - t0)
    out.append("all-vs-all: best %.4f s (%d edges)" % (min(ts), len(e)))
    lat, nb = [], 0
    for s in samples[100:]:
        t0 = time.perf_counter(); _, hits = eng.insert_query(s[0], s[1], s[2], 20); lat.append(time.perf_counter() - t0)
        nb += len(hits)
    out.append("e2e insert+query: p50 %.1f us p95 %.1f (hits %d)" % (
        1e6 * statistics.median(lat[10:]), 1e6 * np.percentile(lat[10:], 95), nb))
    print(name, "\n   " + "\n   ".join(out), flush=True)


if __name__ == "__main__":
    main()
