import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from findneighbour3_b200 import synth
from findneighbour3_b200.engine import Engine
from oracle import c_oracle

def hd(hits):
    return {int(h["row"]): (int(h["dist"]), int(h["n1"]), int(h["n2"]), int(h["nboth"])) for h in hits}

L = 400000
excluded = synth.synthetic_mask(L, L // 15, 40, seed=104)
coll = synth.make_collection(6, 30, genome_length=L, seed=4, excluded=excluded, k1=3000, s=5.0, n_blocks=40, n_mean=30000,
                             m_mean=300, rule="any", n_lognormal_sigma=0.8, max_ns=60000, p_invalid=0.02)
for force in ("0", "1"):
    os.environ["FN3_FORCE_GLOBAL_QUERY"] = force
    eng = Engine(L, 0)
    eng.append_bulk(coll.pos, coll.off, coll.invalid)
    bad = 0
    for ceiling in (20, 5000, 2**31 - 1):
        for q in range(coll.n):
            got = hd(eng.one_vs_all(q, ceiling))
            want = c_oracle.one_vs_all(coll.pos, coll.off, coll.invalid, q, ceiling, nthreads=4)
            if got != want:
                bad += 1
                miss = sorted(set(want) - set(got)); extra = sorted(set(got) - set(want))
                diff = {r: (got[r], want[r]) for r in got if r in want and got[r] != want[r]}
                p, o, inv = coll.lists(q)
                print("force", force, "ceil", ceiling, "q", q, "off", o.tolist(), "nwant", len(want), "ngot", len(got),
                      "missing", miss[:5], "extra", extra[:5], "diff", list(diff.items())[:3])
    print("force_global", force, "bad", bad)
    eng.close()
