"""In-kernel timeline of the compare kernels (variants/libfn3_trace.so, built with -DFN3_TRACE=1): thread 0 of every CTA leaves
   its SM clock at numbered points; printed as microseconds since the CTA's own entry, median / min / max over the CTAs.
       scripts/build_variants.sh trace "-DFN3_TRACE=1"; FN3_LIB_PATH=$PWD/variants/libfn3_trace.so python scripts/trace_prof.py"""
import ctypes as C
import os, sys
import numpy as np
sys.path.insert(0, ".")
sys.argv = ["bench.py"]
import bench
from findneighbour3_b200 import _native
from findneighbour3_b200.engine import Engine

PTS, CTAS, MHZ = 40, 512, 1965.0
NAMES = {0: "k_query entry", 1: "build: row loads issued", 2: "build: zeroed + row words arrived", 3: "build: touched blocks set",
         4: "build: prefix + entries cleared", 5: "build: masks", 6: "build: counts (done)", 7: "after build",
         8: "phase A: heads tested (warp 0)", 9: "phase A: survivors queued", 10: "phase A loop left", 11: "phase B barrier passed",
         12: "phase B drained", 13: "k_scan entry", 14: "mbarriers ready", 15: "first headers arrived", 16: "first copies issued",
         17: "query built", 19: "rows done", 20: "records flushed", 24: "build/prefix: popc of own words", 25: "build/prefix: block scan", 26: "build/prefix: written",
         27: "build/counts: warp sums", 28: "build/counts: block scan", 29: "build/counts: written",
         30: "build/counts: ranges computed", 31: "build/counts: own entries summed", 32: "build/prefix: ranges computed",
         33: "phase A: candidate located", 34: "phase A: head arrived, first 8 fingerprints"}


def main():
    common = bench.make_common(os.cpu_count())
    coll = bench.make_shard(50000, 0, common, os.cpu_count())
    qs = bench.make_queries(common, 40)
    eng = Engine(coll.genome_length, 0)
    eng.append_bulk(coll.pos, coll.off, coll.invalid)
    rows = [eng.append(*q) for q in qs]
    lib = C.CDLL(_native.LIB_PATH)
    buf = np.zeros(CTAS * PTS, dtype=np.uint64)
    fetch = lambda: lib.fn3_debug_trace(buf.ctypes.data_as(C.c_void_p), C.c_int(buf.size))
    for ceil, base, label in ((20, 0, "FILTER ceiling 20"), (250, 13, "STREAM ceiling 250"), (2 ** 31 - 1, 13, "STREAM no ceiling")):
        eng.profile_queries(rows[:6], ceil, flush_l2=True)
        fetch()
        acc, spans = [], []
        for q in rows[6:30]:
            pr = eng.profile_queries([q], ceil, flush_l2=True)
            assert fetch() == 0
            t = buf.reshape(CTAS, PTS).astype(np.int64)
            live = t[:, base] > 0
            g0, g1 = t[live][:, 22], t[live][:, 23]
            if not live.any() or not (g1 > 0).any():
                continue                                  # an invalid query: the kernel returns at once
            spans.append(((g0.max() - g0.min()) / 1e3, (g1[g1 > 0].max() - g0.min()) / 1e3))
            t[:, 22:24] = 0
            acc.append(((t[live] - t[live][:, base:base + 1]) / MHZ, t[live] > 0, float(pr["ms_compare"][0]) * 1e3, int(live.sum())))
        print("== %s: event-to-event %.2f us median, %d CTAs" % (label, np.median([a[2] for a in acc]), acc[0][3]))
        print("   %%globaltimer: CTA entries spread over %.2f us; first entry -> last CTA done %.2f us (medians)" % (
            np.median([s_[0] for s_ in spans]), np.median([s_[1] for s_ in spans])))
        for i in list(range(22)) + list(range(24, PTS)):
            med, lo, hi, n = [], [], [], 0
            for us, ok, _, _ in acc:
                v = us[:, i][ok[:, i]]
                if v.size:
                    med.append(np.median(v)); lo.append(v.min()); hi.append(v.max()); n += v.size
            if med:
                print("   %2d %-36s median %6.2f  earliest %6.2f  latest %6.2f us   (%d CTA samples)" % (
                    i, NAMES.get(i, ""), np.median(med), np.median(lo), np.median(hi), n))
    eng.close()


if __name__ == "__main__":
    main()
