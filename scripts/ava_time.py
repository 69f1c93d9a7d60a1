"""all-vs-all wall time on the 200 000-sample configs[2] collection (or --n), FN3_CARVEOUT etc. from the environment"""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
from findneighbour3_b200 import synth
from findneighbour3_b200.engine import Engine


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    NPZ = "/tmp/fn3_ava_%d.npz" % n
    if not os.path.exists(NPZ):
        c = synth.make_families([300000 + a for a in range(n // 100)], 100, workers=os.cpu_count())
        np.savez(NPZ, pos=c.pos, off=c.off, invalid=c.invalid)
    z = np.load(NPZ)
    eng = Engine(synth.TB_LENGTH, 0)
    eng.append_bulk(z["pos"], z["off"], z["invalid"])
    eng.all_vs_all(12, row_end=4096)
    times = []
    for rep in range(4):
        if rep == 2:
            time.sleep(8)                      # an idle GPU (clocks drop) before the call, as after a CPU-bound set-up phase
        t0 = time.perf_counter(); e = eng.all_vs_all(12); times.append(time.perf_counter() - t0)
    st = eng.stats()
    print("all-vs-all n=%d: %s s (3rd after 8 s idle), %d edges, head refreshes %d" % (
        n, " ".join("%.4f" % t for t in times), len(e), st["head_refreshes"]), flush=True)



if __name__ == "__main__":
    main()
