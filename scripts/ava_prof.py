"""profiling target: a few all-vs-all tiles (128 queries per k_query launch) on the 50 000-sample bench store."""
import sys

sys.path.insert(0, ".")
sys.argv = ["bench.py"]
import bench  # noqa: E402
from findneighbour3_b200.engine import Engine  # noqa: E402

common = bench.make_common()
coll = bench.make_shard(50000, 0, common)
eng = Engine(coll.genome_length, 0)
eng.append_bulk(coll.pos, coll.off, coll.invalid)
edges = eng.all_vs_all(12, row_begin=20000, row_end=20000 + 128 * 12)
print("edges", len(edges), "launches", eng.launch_count())
