"""what the launch floor of the compare kernels is made of: the kernel returns at entry (FN3_DEBUG_STOP=3); grid size and dynamic
shared memory overridden (FN3_DBG_CTAS / FN3_DBG_SMEM); CUDA events around each launch, L2 flush kernel in between or not"""
import os, sys
import numpy as np
sys.path.insert(0, ".")
from findneighbour3_b200 import synth
from findneighbour3_b200.engine import Engine


def main():
    coll = synth.make_families([2_000_000 + a for a in range(100)], 50, workers=os.cpu_count())
    os.environ["FN3_DEBUG_STOP"] = "3"
    for ceiling, name in ((20, "k_query"), (250, "k_scan")):
        for ctas, smem in ((None, None), (None, 0), (1, None), (1, 0), (296, 0), (2368, 0)):
            if ctas is not None: os.environ["FN3_DBG_CTAS"] = str(ctas)
            else: os.environ.pop("FN3_DBG_CTAS", None)
            if smem is not None: os.environ["FN3_DBG_SMEM"] = str(smem)
            else: os.environ.pop("FN3_DBG_SMEM", None)
            eng = Engine(coll.genome_length, 0)
            eng.append_bulk(coll.pos, coll.off, coll.invalid)
            rows = list(range(10, 60))
            eng.profile_queries(rows[:5], ceiling, flush_l2=True)
            a = eng.profile_queries(rows, ceiling, flush_l2=True)
            b = eng.profile_queries(rows, ceiling, flush_l2=False)
            print("%s ctas=%s smem=%s: floor %.2f us behind the L2-flush kernel, %.2f us back to back" % (
                name, ctas or "default", "default" if smem is None else smem, 1e3 * np.median(a["ms_compare"]), 1e3 * np.median(b["ms_compare"])), flush=True)
            eng.close()


if __name__ == "__main__":
    main()
