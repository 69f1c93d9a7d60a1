"""profiling target: the ingest kernels on a TB-sized genome — fn3_compress (k_compress_count/_scan/_write), the fused
insert route fn3_ingest_query (+ k_byte_hist + k_query) and the stand-alone fn3_byte_histogram.  Prints wall-clock per call."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from findneighbour3_b200 import synth  # noqa: E402
from findneighbour3_b200.engine import Engine, byte_histogram  # noqa: E402

mask = synth.synthetic_mask(synth.TB_LENGTH, synth.TB_EXCLUDED, synth.TB_EXCLUDED_RUNS)
coll = synth.make_collection(40, 50, seed=2, excluded=mask)
ref = synth.random_reference(coll.genome_length)
eng = Engine(coll.genome_length, 0)
eng.append_bulk(coll.pos, coll.off, coll.invalid)
eng.set_reference(synth.reference_string(ref), np.flatnonzero(mask).tolist())
seqs = []
for i in np.flatnonzero(coll.invalid == 0)[:4]:
    p, o, _ = coll.lists(int(i))
    seqs.append(synth.to_string(ref, p, o).encode("ascii"))
out = []
for name, fn in (("compress", lambda s: eng.compress(s)), ("ingest_query", lambda s: eng.ingest_query(s, 130000, 20)),
                 ("byte_histogram", lambda s: byte_histogram(s, 0))):
    fn(seqs[0])
    t0 = time.perf_counter()
    for s in seqs:
        fn(s)
    out.append("%s %.3f ms" % (name, 1e3 * (time.perf_counter() - t0) / len(seqs)))
print(" | ".join(out), "| launches", eng.launch_count())
