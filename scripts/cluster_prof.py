"""profiling target: the cluster-statistics kernels (consensus vote over 250 rows, MSA sites + alignment of a 50-sample
family) on a TB-sized genome."""
import sys

import numpy as np

sys.path.insert(0, ".")
from findneighbour3_b200 import synth  # noqa: E402
from findneighbour3_b200.engine import Engine  # noqa: E402

mask = synth.synthetic_mask(synth.TB_LENGTH, synth.TB_EXCLUDED, synth.TB_EXCLUDED_RUNS)
coll = synth.make_collection(100, 50, seed=2, excluded=mask)
eng = Engine(coll.genome_length, 0)
eng.append_bulk(coll.pos, coll.off, coll.invalid)
eng.set_reference(synth.reference_string(synth.random_reference(coll.genome_length)), np.flatnonzero(mask).tolist())
fam = np.flatnonzero((coll.family == coll.family[0]) & (coll.invalid == 0))[:50]
rows250 = np.flatnonzero(coll.invalid == 0)[:250]
for _ in range(2):
    sites = eng.msa_sites(fam)
    codes, counts = eng.msa_align(fam, sites)
    pos, off = eng.consensus(rows250, 200)
print("sites", sites.size, "consensus positions", pos.size, "launches", eng.launch_count())
