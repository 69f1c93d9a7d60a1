#!/usr/bin/env python3
"""bench.py — headline benchmark of the masked pairwise SNV distance path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--samples S] [--ceiling C]

Workload of the headline line (config.workload): BASELINE.json configs[1] — a 50 000-sample synthetic M. tuberculosis-shaped
collection PER GPU (1 000 families x 50, L = 4 411 532, TB-exclude-shaped mask, ~480 variant + ~4 000 blocky N positions per
sample, 1 % invalid), one new-sample one-vs-all neighbour query per step at snpCeiling 20 (weak scaling: the collection
grows with N and every query is compared against all shards).

  value      pairwise comparisons/s, device-resident: the K query samples are already rows of the HBM store; a step = ONE
             k_query launch (query-structure build fused in), CUDA events on the engine's stream, L2 flushed (256 MiB
             write, untimed) before every step.  Built from the MEDIAN per-step time of the slowest rank (one host hiccup
             in an event-bracketed sample must not halve the figure); mean, p95 and the per-rank figures are beside it.
  e2e        the same metric through the C-ABI with HOST buffers: N = 1 `fn3_insert_query` (persist + mcompare of the
             server's insert(), one call); N > 1 `fn3_group_insert_query` — ONE process (rank 0) drives all N GPUs through
             an engine group, the other ranks wait on a host barrier.  Every step copies the new sample host->device and
             the neighbour records device->host.  A sample of the timed steps is checked against the C oracle.
  e2e_class  latency through the drop-in class at the same store size: seqComparer.mcompare / neighbours / insert_sequence.
  roofline   the timed kernel (k_query, FILTER variant) moves ~5 MB per launch: it is LATENCY bound, and says so; `frac`
             = bytes the kernel actually moved (instrumented, untimed pass of the same launches) / median launch time /
             measured HBM peak.  `frac_canonical` keeps SURVEY 8(d)'s canonical-bytes figure (> 1 by construction: the
             kernel proves "not a neighbour" from a 96-byte head).  `fullscan` = k_scan (the STREAM variant: rows by bulk
             async copy into shared-memory rings) at ceilings 250 (the recompression scan) and 2^31-1 (no pair may exit
             early, every row is streamed): the HBM-bound figure of this path.
  workloads  the same one-vs-all on collections that stress the head filter: `lineage` (deep shared ancestry) and
             `outbreak` (every candidate IS a neighbour), with the filter's pass rate, device and e2e latency, parity.
  all_vs_all BASELINE configs[2] at every N: 200 000 samples at 12 SNV, store replicated, block rows r % N == rank, edges
             gathered to rank 0 over NCCL inside the timed region, order-independent checksum (equal at every N), sampled rows
             against the C oracle.
  config4    BASELINE configs[4]: 125 000 samples per GPU (1 M at N = 8) resident in an engine group, streamed inserts each
             queried one-vs-all, p50 / p95, a sample of them against the C oracle.
"""
import argparse
import gc
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pairwise SNV comparisons/sec (one-vs-all, 50k-sample TB collection per GPU, snpCeiling 20)"
UNIT = "comparisons/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--samples", type=int, default=50000, help="samples per GPU shard (headline workload)")
    ap.add_argument("--ceiling", type=int, default=20)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--quick", action="store_true", help="headline legs only (value, roofline, e2e)")
    ap.add_argument("--legs", default="phases,class,workloads,ava,c4,extras", help="extra legs to run (comma separated)")
    ap.add_argument("--ava-samples", type=int, default=200000)
    ap.add_argument("--ava-ceiling", type=int, default=12)
    ap.add_argument("--c4-samples", type=int, default=125000, help="configs[4]: samples per GPU")
    ap.add_argument("--c4-inserts", type=int, default=200)
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """per-launch DRAM bytes from the committed ncu captures (profiles/traffic.json): a cross-check, not the figure"""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(path))
    except Exception:
        return None


def pct(a, q):
    a = sorted(a)
    return a[min(len(a) - 1, int(q * (len(a) - 1) + 0.5))] if a else None


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed regions (NVML, every 2 ms, own thread)."""

    def __init__(self, index, period=0.002):
        self.period = period
        self.samples, self.reasons = [], set()
        self.stop_flag = threading.Event()
        self.thread = self.h = self.max_mhz = None
        try:
            import pynvml
            import torch
            self.nv = pynvml
            pynvml.nvmlInit()
            pr = torch.cuda.get_device_properties(index)
            busid = "%08X:%02X:%02X.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            try:
                self.h = pynvml.nvmlDeviceGetHandleByPciBusId(busid.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def _loop(self):
        nv = self.nv
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
        while not self.stop_flag.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self.h is not None:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        if self.thread is None:
            return out
        self.stop_flag.set()
        self.thread.join(timeout=2)
        if self.samples:
            out.update(sm_mhz=statistics.median(self.samples), samples=len(self.samples),
                       sm_mhz_min=min(self.samples), sm_mhz_max_seen=max(self.samples))
        out["reasons"] = sorted(self.reasons)
        return out


# ----------------------------------------------------------------------------------------------------
# synthetic collections (seeded; the same families whatever N is)
# ----------------------------------------------------------------------------------------------------
COMMON_ANC = 20      # families every rank regenerates: the new samples are their relatives


def make_common(workers):
    from findneighbour3_b200 import synth
    return synth.make_families([777000 + a for a in range(COMMON_ANC)], 50, workers=workers)


def make_shard(samples, shard, common, workers):
    """shard `shard` of the headline collection (`samples` samples: families of 50); shard 0 ends with the `common` families"""
    from findneighbour3_b200 import synth
    n_anc = max(1, samples // 50)
    if shard == 0:
        own = max(0, n_anc - COMMON_ANC)
        if own == 0:
            return common
        return synth.concat(synth.make_families([2_000_000 + a for a in range(own)], 50, workers=workers), common)
    return synth.make_families([2_000_000 + 100_000 * shard + a for a in range(n_anc)], 50, workers=workers)


def make_queries(common, count, seed=12345):
    """`count` newly sequenced samples: relatives (0-3 extra SNVs) of random valid samples of the common families"""
    from findneighbour3_b200 import synth
    rng = np.random.default_rng(seed)
    valid = np.flatnonzero(common.invalid == 0)
    return [synth.new_sample_like(common, int(rng.choice(valid)), int(rng.integers(0, 4)), seed=seed + k) for k in range(count)]


def hits_to_dict(hits):
    return {int(h["row"]): (int(h["dist"]), int(h["n1"]), int(h["n2"]), int(h["nboth"])) for h in hits}


class OracleStore:
    """the flat collection as the C oracle sees it (TEST INFRASTRUCTURE: bench.py uses it only as the checker of the
    neighbour lists it times): a list of shards whose rows are numbered in append order, plus the samples inserted since."""

    def __init__(self, threads):
        self.shards, self.first, self.n, self.extra, self.threads = [], [], 0, [], max(1, threads)

    def add(self, coll):
        self.shards.append(coll)
        self.first.append(self.n)
        self.n += coll.n

    def inserted(self, sample):
        self.extra.append((self.n, sample))
        self.n += 1

    def neighbours(self, sample, ceiling, below=None):
        """{global row: (dist, n1, n2, nboth)} of `sample` against every row numbered < below (default: all)"""
        from oracle import c_oracle
        p, o, inv = sample
        out = {}
        if inv:
            return out
        below = self.n if below is None else below
        for coll, first in zip(self.shards, self.first):
            if first >= below:
                continue
            pos = np.concatenate([coll.pos, p])
            off = np.concatenate([coll.off, coll.off[-1] + o[1:].astype(np.int64)])
            invalid = np.concatenate([coll.invalid, np.zeros(1, np.uint8)])
            got = c_oracle.one_vs_all(pos, off, invalid, coll.n, ceiling, nthreads=self.threads)
            out.update({first + r: v for r, v in got.items() if first + r < below})
        for gid, (ep, eo, einv) in self.extra:
            if gid >= below or einv:
                continue
            v = c_oracle.pair(p, o, False, ep, eo, False, ceiling)
            if v is not None:
                out[gid] = v
        return out


# ----------------------------------------------------------------------------------------------------
# CPU legs (oracle = test/bench infrastructure; never on the product path)
# ----------------------------------------------------------------------------------------------------
def cpu_python_port(coll, queries, ceiling, budget_s):
    """the reference's algorithm (Python set algebra, single thread — the class the server imports is single-threaded
    and GIL-bound) on a bounded sample: the query against the last `m` stored samples."""
    from oracle import seqcomparer_port as port
    from findneighbour3_b200 import packing
    m = min(coll.n, 1500)
    store = {i: coll.as_dict(i) for i in range(coll.n - m, coll.n)}     # includes the queries' own families
    pairs, t_used, nq = 0, 0.0, 0
    for (p, o, inv) in queries:
        store["q"] = {"invalid": 1} if inv else packing.unpack_lists(p, o, None)
        t0 = time.perf_counter()
        port.mcompare(store, {}, "q", ceiling)
        t_used += time.perf_counter() - t0
        pairs += m
        nq += 1
        if t_used > budget_s:
            break
    return {"value": pairs / t_used, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "%d queries x last %d candidates of the shard, oracle/seqcomparer_port.py "
                      "(Python-set restatement of seqComparer.mcompare), %.1f s" % (nq, m, t_used)}


def cpu_c_oracle(coll, queries, ceiling, budget_s, threads):
    from oracle import c_oracle
    t_used, pairs, nq = 0.0, 0, 0
    for (p, o, inv) in queries:
        pos = np.concatenate([coll.pos, p])
        off = np.concatenate([coll.off, coll.off[-1] + o[1:].astype(np.int64)])
        invalid = np.concatenate([coll.invalid, np.array([1 if inv else 0], np.uint8)])
        t0 = time.perf_counter()
        c_oracle.one_vs_all(pos, off, invalid, coll.n, ceiling, nthreads=threads)
        t_used += time.perf_counter() - t0
        pairs += coll.n
        nq += 1
        if t_used > budget_s:
            break
    return {"value": pairs / t_used, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d queries x all %d candidates, oracle/fn3_oracle.c (C restatement, %d pthreads), %.1f s"
                      % (nq, coll.n, threads, t_used)}


def _reference_worker(conn, coll, lo, hi, ceiling):
    """one host process of the reference arm: the Python-set algorithm over candidates [lo, hi) of the shard"""
    from oracle import seqcomparer_port as port
    from findneighbour3_b200 import packing
    store = {i: coll.as_dict(i) for i in range(lo, hi)}
    conn.send(len(store))
    while True:
        msg = conn.recv()
        if msg is None:
            return
        p, o, inv = msg
        store["q"] = {"invalid": 1} if inv else packing.unpack_lists(p, o, None)
        conn.send(len(port.mcompare(store, {}, "q", ceiling)))


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path, timed on this host.
    The reference is pure Python and cannot travel to the GPU box, so its restatement (oracle port) runs.  The class the
    server imports is single-threaded and seqComparer_mt's threads share one GIL (src/seqComparer_mt.py:363-385), so to give
    the algorithm every host core the candidates of a step are split over one PROCESS per core, each holding its slice of
    the store as the reference holds it (dicts of Python sets)."""
    if rank != 0:
        return
    import multiprocessing as mp
    workers = os.cpu_count() or 1
    per = 1500                                                  # candidates per process and step: ~0.5-0.7 s of set algebra
    try:
        import psutil
        procs = max(1, min(workers, int(0.25 * psutil.virtual_memory().available / (per * 0.35e6))))
    except Exception:       # pragma: no cover
        procs = workers
    common = make_common(workers)
    m = min(int(args.samples), per * procs)
    procs = max(1, min(procs, m // 200))
    coll = make_shard(m, 0, common, workers)
    m = coll.n
    queries = make_queries(common, max(args.steps + args.warmup, 4))
    budget = max(5.0, min(90.0, args.cpu_seconds * 3))
    ctx = mp.get_context("fork")                                # the shard is inherited, not pickled; no CUDA in this arm
    bounds = [m * i // procs for i in range(procs + 1)]
    conns, kids = [], []
    for i in range(procs):
        a, b = ctx.Pipe()
        k = ctx.Process(target=_reference_worker, args=(b, coll, bounds[i], bounds[i + 1], args.ceiling), daemon=True)
        k.start()
        conns.append(a)
        kids.append(k)
    assert sum(c.recv() for c in conns) == m

    def step(q):
        t0 = time.perf_counter()
        for c in conns:
            c.send(q)
        n = sum(c.recv() for c in conns)
        dt = time.perf_counter() - t0
        assert n == m, (n, m)                                   # one row per candidate, as mcompare returns them
        return dt

    n_warm = min(args.warmup, 1)
    for q in queries[:n_warm]:
        step(q)
    t_used, steps = 0.0, 0
    for q in queries[args.warmup:args.warmup + args.steps]:
        t_used += step(q)
        steps += 1
        if t_used > budget:
            break
    for c in conns:
        c.send(None)
    for k in kids:
        k.join(10)
    value = steps * m / t_used
    sample = ("%d steps, each one query x %d candidates (bounded sample of the 50k workload) split over %d processes: "
              "Python-set restatement of seqComparer.mcompare (oracle/seqcomparer_port.py), the query sent to and the row "
              "counts read from each process inside the timed region" % (steps, m, procs))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": n_warm, "ms_per_step": 1e3 * t_used / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": "configs[1]: 50,000 synthetic TB sequences per GPU, one new-sample one-vs-all query "
                                   "per step, snpCeiling %d" % args.ceiling,
                       "samples_per_gpu": int(args.samples), "genome_length": int(coll.genome_length),
                       "candidates_per_step": m,
                       "sample": "each step compares the query with a bounded sample of %d of the shard's candidates "
                                 "(the reference's cost is linear in the candidates)" % m},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "host_cpus": os.cpu_count()}
    try:
        line["cpu_baseline_native"] = cpu_c_oracle(coll, queries[:3], args.ceiling, 10.0, os.cpu_count() or 1)
    except Exception as ex:       # pragma: no cover
        line["cpu_baseline_native"] = {"error": str(ex)}
    emit(line)


# ----------------------------------------------------------------------------------------------------
def emit(line):
    """the ONE JSON line goes to the real stdout; everything else that writes to fd 1 (NCCL's version banner,
    library chatter) was redirected to stderr at start-up."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def log(*a):
    sys.stderr.write("[bench %.1f] " % (time.time() - _T0) + " ".join(str(x) for x in a) + "\n")
    sys.stderr.flush()


_T0 = time.time()


def edge_checksum(edges):
    """order-independent 64-bit checksum of an edge set (sum and xor of a per-edge mix; wrap-around arithmetic)"""
    if len(edges) == 0:
        return {"sum": 0, "xor": 0}
    with np.errstate(over="ignore"):
        h = (edges["row1"].astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)) ^ (edges["row2"].astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F))
        h = h + edges["dist"].astype(np.uint64) * np.uint64(0x165667B19E3779F9) + edges["n1"].astype(np.uint64) * np.uint64(0x27D4EB2F165667C5)
        h = h ^ (edges["n2"].astype(np.uint64) * np.uint64(0x85EBCA77C2B2AE63)) ^ (edges["nboth"].astype(np.uint64) * np.uint64(0xFF51AFD7ED558CCD))
        h = (h ^ (h >> np.uint64(33))) * np.uint64(0xC4CEB9FE1A85EC53)
        return {"sum": int(np.add.reduce(h, dtype=np.uint64)), "xor": int(np.bitwise_xor.reduce(h))}


def main():
    global _REAL_STDOUT
    args = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(local)
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # a HOST barrier for the legs in which rank 0 alone drives every GPU: a rank parked in an NCCL collective would
        # keep a spinning kernel on the very GPU rank 0's engine group wants to use
        host_group = dist.new_group(backend="gloo")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def host_barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=host_group)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def gather_floats(xs):
        """every rank's list of floats -> list of lists (on every rank)"""
        if world == 1:
            return [list(xs)]
        t = torch.tensor(list(xs), dtype=torch.float64, device="cuda")
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [o.cpu().tolist() for o in out]

    from findneighbour3_b200 import synth
    from findneighbour3_b200.engine import Engine, GroupEngine

    cpus = os.cpu_count() or 1
    workers = max(1, cpus // world)
    K, W = args.steps, args.warmup
    ceiling = args.ceiling
    legs = set() if args.quick else set(x for x in args.legs.split(",") if x)
    peak, peak_src = peaks()
    errors = {}

    # ================================================================================================
    # leg 1: device-resident `value` (every rank: its own shard on its own GPU)
    # ================================================================================================
    t_gen = time.time()
    common = make_common(workers)
    coll = make_shard(args.samples, rank, common, workers)
    log("rank", rank, "shard of", coll.n, "samples generated in %.1f s" % (time.time() - t_gen))
    eng = Engine(coll.genome_length, local)
    t_bulk = time.perf_counter()
    eng.append_bulk(coll.pos, coll.off, coll.invalid)       # server start-up reload path: validate, pack, copy, place heads
    t_bulk = time.perf_counter() - t_bulk
    n_e2e = max(K, 200)                              # the latency legs use at least 200 steps, whatever K is
    queries = make_queries(common, (K + W) + 2 * (n_e2e + max(W, 2 * world)) + 64)      # the same new samples on every rank
    t_gen = time.time() - t_gen

    resident = [eng.append(p, o, inv) for (p, o, inv) in queries[:K + W]]
    n_rows = eng.stats()["n_rows"]
    bytes_canon_all = int(coll.bytes_canonical()[coll.invalid == 0].sum())
    for (p, o, inv) in queries[:K + W]:
        if not inv:
            bytes_canon_all += 4 * int(o[6]) + 32
    sampler = ClockSampler(local)
    barrier()
    eng.profile_queries(resident[:W], ceiling, flush_l2=True)
    barrier()
    sampler.start()
    launches0 = eng.launch_count()
    prof = eng.profile_queries(resident[W:], ceiling, flush_l2=True)          # EXACTLY K timed steps
    launches = eng.launch_count() - launches0 - K      # minus the untimed L2-flush launches
    barrier()
    step_ms = (prof["ms_build"] + prof["ms_compare"]).astype(np.float64)
    warm = eng.profile_queries(resident[W:], ceiling, flush_l2=False)
    touched = eng.profile_queries(resident[W:W + min(K, 50)], ceiling, flush_l2=False, count_bytes=True)
    my_stats = [float(np.median(step_ms)), float(step_ms.mean()), float(pct(step_ms.tolist(), 0.95)), float(step_ms.max()),
                float(n_rows - 1)]
    all_stats = gather_floats(my_stats)
    slow = int(np.argmax([s[0] for s in all_stats]))
    med_ms = all_stats[slow][0]                                  # median step of the slowest rank
    comps_per_step = sum(s[4] for s in all_stats)                # every other row of every shard
    value = comps_per_step / (med_ms * 1e-3)
    mean_ms = max(s[1] for s in all_stats)
    t_warm = max_over_ranks(float(np.median(warm["ms_build"] + warm["ms_compare"])))

    cmp_ms = float(np.median(prof["ms_compare"]))
    st = eng.stats()
    store_bytes_rows = 4 * st["store_words"] + 32 * st["n_rows"]
    moved = float(touched["bytes_touched"].mean())
    traffic = ncu_traffic() or {}
    roofline = {
        "bound": "hbm", "regime": "latency", "kernel": "k_query (FILTER variant)", "unit": "GB/s", "peak": peak, "peak_source": peak_src,
        "achieved": moved / (cmp_ms * 1e-3) / 1e9, "frac": moved / (cmp_ms * 1e-3) / 1e9 / peak,
        "traffic": moved, "traffic_source": "instrumented pass of the same launches in this run (heads + row words loaded)",
        "traffic_ncu": traffic.get("k_query_bytes_per_launch"), "traffic_ncu_source": traffic.get("source"),
        "launch_ms": cmp_ms, "launch_ms_mean": float(prof["ms_compare"].mean()),
        "algorithmic_bytes_per_launch": bytes_canon_all,
        "achieved_canonical": bytes_canon_all / (cmp_ms * 1e-3) / 1e9,
        "frac_canonical": bytes_canon_all / (cmp_ms * 1e-3) / 1e9 / peak,
        "store_bytes": store_bytes_rows,
        "filter_pass_rate": float(touched["n_passed"].mean()) / max(n_rows - 1, 1),
        "note": "the timed kernel decides ~99.9 % of the pairs from a 96-byte head: it moves ~5 MB in ~18 us and is bound by "
                "launch + query build + two dependent memory round trips, not by bandwidth; frac_canonical (SURVEY 8(d): "
                "canonical B_pair bytes / time) exceeds 1 by construction and is not a bandwidth figure",
    }

    # ---- fullscan: k_scan, the kernel that DOES stream the store (first-class sub-block) --------------------------------
    nq_scan = min(K, 24)
    scan = {}
    for name, c in (("ceiling_250", 250), ("no_ceiling", 2 ** 31 - 1)):
        eng.profile_queries(resident[W:W + 3], c, flush_l2=True)
        pr = eng.profile_queries(resident[W:W + nq_scan], c, flush_l2=True, count_bytes=True)
        ms = float(np.median(pr["ms_compare"]))
        mv = float(pr["bytes_touched"].mean())
        scan[name] = {"ceiling": c, "launch_ms": ms, "launch_ms_p95": float(pct(pr["ms_compare"].tolist(), 0.95)),
                      "moved_bytes_per_launch": mv, "achieved": mv / (ms * 1e-3) / 1e9, "frac": mv / (ms * 1e-3) / 1e9 / peak,
                      "frac_canonical": bytes_canon_all / (ms * 1e-3) / 1e9 / peak,
                      "comparisons_per_s": (n_rows - 1) / (ms * 1e-3), "neighbours_per_query": float(pr["n_hits"].mean())}
    roofline["fullscan"] = dict(scan["no_ceiling"], bound="hbm", kernel="k_scan (STREAM variant: bulk async copies into "
                                "per-warp shared-memory rings)", unit="GB/s", peak=peak,
                                frac_touched=scan["no_ceiling"]["frac"], traffic_ncu=traffic.get("k_scan_bytes_per_launch"),
                                recompression_scan=scan["ceiling_250"])

    # ---- phase split of the timed kernel (separate small engines with the kernel's debug stops; untimed for `value`) ----
    phases = None
    if rank == 0 and "phases" in legs:
        try:
            phases = {}
            for stop, name in ((3, "launch_floor"), (1, "query_build"), (2, "phase_a_heads")):
                os.environ["FN3_DEBUG_STOP"] = str(stop)
                e2 = Engine(coll.genome_length, local)
                os.environ.pop("FN3_DEBUG_STOP")
                e2.append_bulk(coll.pos, coll.off, coll.invalid)
                rows2 = [e2.append(p, o, inv) for (p, o, inv) in queries[:W + 30]]
                e2.profile_queries(rows2[:W], ceiling, flush_l2=True)
                pr = e2.profile_queries(rows2[W:], ceiling, flush_l2=True)
                phases[name + "_ms"] = float(np.median(pr["ms_compare"]))
                e2.close()
            phases["whole_kernel_ms"] = cmp_ms
            phases["note"] = "cumulative: kernel stopped at entry / after the query build / after the head filter (FN3_DEBUG_STOP)"
        except Exception as ex:   # pragma: no cover
            errors["phases"] = repr(ex)
        finally:
            os.environ.pop("FN3_DEBUG_STOP", None)
    roofline["phases"] = phases

    # ================================================================================================
    # leg 2: end to end through the C-ABI with host buffers
    # ================================================================================================
    fresh = queries[K + W:]
    oracle_threads = cpus

    WG = max(W, 2 * world)                           # group legs: every device has had its turn as the owner before the clock starts

    def timed_inserts(insert, samples, n_rows_now, check=None, check_every=0, warm=None):
        """`warm` untimed + the remaining timed calls of insert(sample) -> hits; optional parity check of a sample of the steps"""
        warm = W if warm is None else warm
        lat, n_nb, comps, checked = [], 0, 0, 0
        for s in samples[:warm]:
            hits = insert(s)
            if check:
                check.inserted(s)
        n_rows_now += warm
        timed = samples[warm:]
        pending = []
        gc.collect()
        gc.disable()                                 # a generation-2 pass over this process's big host-side collections
        t0 = time.perf_counter()                     # (the oracle's copy of every shard) is tens of ms: not the engine's
        for i, s in enumerate(timed):
            t1 = time.perf_counter()
            hits = insert(s)
            lat.append(time.perf_counter() - t1)
            n_nb += len(hits)
            comps += n_rows_now + i
            if check:
                if check_every and i % check_every == 0:
                    pending.append((s, check.n, hits_to_dict(hits)))
                check.inserted(s)
        t = time.perf_counter() - t0
        gc.enable()
        for s, below, got in pending:                # the oracle runs after the timed loop
            want = check.neighbours(s, ceiling, below=below)
            assert got == want, "e2e parity: neighbours of a timed insert differ from the C oracle"
            checked += 1
        order = sorted(range(len(lat)), key=lambda i_: -lat[i_])[:3]
        return {"seconds": t, "lat": lat, "neighbours": n_nb, "comparisons": comps, "parity_checked": checked,
                "slowest": [[i_, 1e3 * lat[i_]] for i_ in order]}

    e2e = None
    grp = None
    ostore = None
    # (NVML queries from eight processes at 500 Hz stalled the group's launches for milliseconds: during the end-to-end legs
    #  only rank 0 samples, every 25 ms)
    clocks = sampler.stop()                          # the device-timed legs are over: every rank stops polling NVML at 500 Hz
    sampler2 = ClockSampler(local, period=0.025) if rank == 0 else None
    if sampler2:
        sampler2.start()
    if world == 1:
        ostore = OracleStore(oracle_threads)
        ostore.add(coll)
        for s in queries[:K + W]:
            ostore.inserted(s)
        st0 = eng.stats()
        two = timed_inserts(lambda s: eng.one_vs_all(eng.append(*s), ceiling), fresh[:W + n_e2e], st0["n_rows"], ostore, 0)
        st1 = eng.stats()
        r = timed_inserts(lambda s: eng.insert_query(s[0], s[1], s[2], ceiling)[1], fresh[W + n_e2e:2 * (W + n_e2e)],
                          st1["n_rows"], ostore, max(1, n_e2e // 8))
        st2 = eng.stats()
        e2e = {"value": r["comparisons"] / r["seconds"], "unit": UNIT, "steps": n_e2e,
               "h2d_bytes_per_step": (st2["h2d_bytes"] - st1["h2d_bytes"]) / (n_e2e + W),
               "d2h_bytes_per_step": (st2["d2h_bytes"] - st1["d2h_bytes"]) / (n_e2e + W),
               "ms_per_step": 1e3 * r["seconds"] / n_e2e, "p50_ms": 1e3 * statistics.median(r["lat"]),
               "p95_ms": 1e3 * pct(r["lat"], 0.95), "neighbours_per_query": r["neighbours"] / n_e2e,
               "parity_steps_checked": r["parity_checked"], "slowest_steps_ms": r["slowest"],
               "call": "fn3_insert_query via ctypes (server insert(): persist + mcompare, findNeighbour3-server.py:516+530)",
               "two_calls": {"ms_per_step": 1e3 * two["seconds"] / n_e2e, "p50_ms": 1e3 * statistics.median(two["lat"]),
                             "call": "fn3_append + fn3_one_vs_all (seqComparer.persist, then seqComparer.mcompare)"}}
    else:
        # ONE process drives every GPU: rank 0 builds an engine group over all N devices holding the N shards
        host_barrier()
        if rank == 0:
            try:
                t0 = time.time()
                grp = GroupEngine(coll.genome_length, list(range(world)))
                ostore = OracleStore(oracle_threads)
                for r_ in range(world):
                    sh = coll if r_ == 0 else make_shard(args.samples, r_, common, cpus)
                    grp.append_bulk(sh.pos, sh.off, sh.invalid)
                    ostore.add(sh)
                log("group of", world, "devices loaded with", grp.stats()["n_rows"], "rows in %.1f s" % (time.time() - t0),
                    "peer access:", grp.peer_access)
                st1 = grp.stats()
                r = timed_inserts(lambda s: grp.insert_query(s[0], s[1], s[2], ceiling)[1], fresh[:WG + n_e2e], st1["n_rows"],
                                  ostore, max(1, n_e2e // 8), warm=WG)
                st2 = grp.stats()
                e2e = {"value": r["comparisons"] / r["seconds"], "unit": UNIT, "steps": n_e2e,
                       "h2d_bytes_per_step": (st2["h2d_bytes"] - st1["h2d_bytes"]) / (n_e2e + WG),
                       "d2h_bytes_per_step": (st2["d2h_bytes"] - st1["d2h_bytes"]) / (n_e2e + WG),
                       "ms_per_step": 1e3 * r["seconds"] / n_e2e, "p50_ms": 1e3 * statistics.median(r["lat"]),
                       "p95_ms": 1e3 * pct(r["lat"], 0.95), "neighbours_per_query": r["neighbours"] / n_e2e,
                       "parity_steps_checked": r["parity_checked"], "slowest_steps_ms": r["slowest"], "candidates": int(st1["n_rows"]),
                       "warmup_steps": WG,
                       "peer_access": grp.peer_access,
                       "call": "fn3_group_insert_query via ctypes: ONE process (rank 0) drives all %d GPUs — one host thread per "
                               "device inside libfn3, the sample host->every device, neighbour records through each device's "
                               "mapped pinned block; no collective on the data path" % world}
                # stored queries: the owner's row is read by the other devices over NVLink (peer access)
                qrows = list(range(st1["n_rows"] - 40, st1["n_rows"]))
                lat = []
                for q in qrows:
                    t1 = time.perf_counter()
                    grp.one_vs_all(q, ceiling)
                    lat.append(time.perf_counter() - t1)
                e2e["stored_query_p50_ms"] = 1e3 * statistics.median(lat[8:])
            except Exception as ex:   # pragma: no cover
                errors["e2e_group"] = repr(ex)
                log("e2e group leg failed:", repr(ex))
        host_barrier()
    if sampler2:
        c2 = sampler2.stop()
        clocks["e2e_region"] = {k: c2.get(k) for k in ("sm_mhz", "sm_mhz_min", "reasons", "samples")}

    # ================================================================================================
    # leg 3: latency through the drop-in class (rank 0)
    # ================================================================================================
    e2e_class = None
    if "class" in legs:
        host_barrier()                               # rank 0's class drives every GPU: nobody may sit in a device collective
    if rank == 0 and "class" in legs:
        try:
            from findneighbour3_b200.seqComparer import seqComparer
            dev = local if world == 1 else list(range(world))
            ref_codes = synth.random_reference(coll.genome_length)
            excl = np.flatnonzero(synth.synthetic_mask(synth.TB_LENGTH, synth.TB_EXCLUDED, synth.TB_EXCLUDED_RUNS)).tolist()
            sc = seqComparer(reference=synth.reference_string(ref_codes), maxNs=130000, snpCeiling=ceiling,
                             excludePositions=set(excl), device=dev)
            shards_c = [coll] if world == 1 else ostore.shards
            at = 0
            for sh in shards_c:
                sc.persist_packed(["s%d" % (at + i) for i in range(sh.n)], sh.pos, sh.off, sh.invalid)
                at += sh.n
            n_store = at
            cq = fresh[-64:]
            guids = []
            for i, (p, o, inv) in enumerate(cq):
                g = "q%d" % i
                sc.persist_packed([g], p, np.asarray(o, dtype=np.int64), np.array([1 if inv else 0], np.uint8))
                guids.append(g)
            lat_call, lat_iter, lat_nb = [], [], []
            for g in guids[8:40]:
                t1 = time.perf_counter()
                res = sc.mcompare(g)
                t2 = time.perf_counter()
                links = {}
                for (guid1, guid2, d, n1, n2, nboth, a, b, c_) in res:      # the server's own loop (server.py:532-538)
                    if not guid1 == guid2:
                        if d is not None:
                            if d <= ceiling:
                                links[guid2] = {"dist": d, "n1": n1, "n2": n2, "nboth": nboth}
                t3 = time.perf_counter()
                found = sc.neighbours(g)
                t4 = time.perf_counter()
                assert set(links) == set(found) and len(res) == n_store + len(guids) - 1
                lat_call.append(t2 - t1); lat_iter.append(t3 - t1); lat_nb.append(t4 - t3)
            e2e_class = {"stored_samples": n_store, "devices": world,
                         "mcompare_call_p50_ms": 1e3 * statistics.median(lat_call), "mcompare_call_p95_ms": 1e3 * pct(lat_call, 0.95),
                         "mcompare_and_server_loop_p50_ms": 1e3 * statistics.median(lat_iter),
                         "mcompare_and_server_loop_p95_ms": 1e3 * pct(lat_iter, 0.95),
                         "neighbours_p50_ms": 1e3 * statistics.median(lat_nb), "neighbours_p95_ms": 1e3 * pct(lat_nb, 0.95),
                         "note": "seqComparer.mcompare(guid) returns its rows lazily (call = device one-vs-all + a snapshot of "
                                 "the guids); `_and_server_loop` adds the unmodified server's loop over all rows "
                                 "(findNeighbour3-server.py:532-538: one tuple unpack + tests per stored sample, in Python); "
                                 "neighbours() is the survivors-only form"}
            # the whole insert route from the 4.4 MB sequence string: examine counts + compress + persist + mcompare
            seqs = []
            for (p, o, inv) in cq[40:52]:
                if not inv:
                    seqs.append(synth.to_string(ref_codes, p, o))
            lat_ins = []
            for i, sq in enumerate(seqs):
                t1 = time.perf_counter()
                sc.insert_sequence("n%d" % i, sq)
                lat_ins.append(time.perf_counter() - t1)
            if len(lat_ins) > 3:
                e2e_class["insert_sequence_p50_ms"] = 1e3 * statistics.median(lat_ins[2:])
                e2e_class["insert_sequence_p95_ms"] = 1e3 * pct(lat_ins[2:], 0.95)
            sc.engine.close()
            del sc
        except Exception as ex:   # pragma: no cover
            errors["e2e_class"] = repr(ex)
            log("e2e_class leg failed:", repr(ex))
    if "class" in legs:
        host_barrier()

    # ================================================================================================
    # leg 4 (N = 1): collections that stress the head filter
    # ================================================================================================
    workloads = None
    if world == 1 and "workloads" in legs:
        workloads = {"star": {"samples": int(coll.n), "filter_pass_rate": roofline["filter_pass_rate"],
                              "device_p50_ms": float(np.median(step_ms)), "device_p95_ms": float(pct(step_ms.tolist(), 0.95)),
                              "e2e_p50_ms": e2e["p50_ms"] if e2e else None,
                              "neighbours_per_query": float(prof["n_hits"].mean()),
                              "parity_queries_checked": e2e["parity_steps_checked"] if e2e else 0}}
        for name in ("lineage", "outbreak"):
            try:
                t0 = time.time()
                if name == "lineage":
                    wc = synth.make_lineages(args.samples // 50, 50, workers=cpus)
                else:
                    wc = synth.make_outbreak(args.samples, workers=cpus)
                we = Engine(wc.genome_length, local)
                we.append_bulk(wc.pos, wc.off, wc.invalid)
                rng = np.random.default_rng(99)
                valid = np.flatnonzero(wc.invalid == 0)
                wq = [synth.new_sample_like(wc, int(rng.choice(valid)), int(rng.integers(0, 4)), seed=500 + k) for k in range(48)]
                rows_w = [we.append(p, o, inv) for (p, o, inv) in wq[:24]]
                # the head filter's pass rate (heads refreshed on entry: the store holds >= 4 096 rows), then real queries: the
                # engine settles on FILTER or STREAM for this collection, and the timed launches take that variant
                tb0 = we.profile_queries(rows_w[4:], ceiling, flush_l2=False, count_bytes=True)
                for r_ in rows_w[:4]:
                    we.one_vs_all(r_, ceiling)
                we.profile_queries(rows_w[:4], ceiling, flush_l2=True)
                pr = we.profile_queries(rows_w[4:], ceiling, flush_l2=True)
                tb = we.profile_queries(rows_w[4:], ceiling, flush_l2=False, count_bytes=True)
                fs = we.profile_queries(rows_w[4:12], 2 ** 31 - 1, flush_l2=True)
                ms = (pr["ms_build"] + pr["ms_compare"]).astype(np.float64)
                wo = OracleStore(oracle_threads)
                wo.add(wc)
                for s in wq[:24]:
                    wo.inserted(s)
                lat, checked = [], 0
                for i, s in enumerate(wq[24:]):
                    t1 = time.perf_counter()
                    row, hits = we.insert_query(s[0], s[1], s[2], ceiling)
                    lat.append(time.perf_counter() - t1)
                    if i % 8 == 0:
                        assert hits_to_dict(hits) == wo.neighbours(s, ceiling), "workload %s: parity" % name
                        checked += 1
                    wo.inserted(s)
                st_w = we.stats()
                n_w = st_w["n_rows"]
                workloads[name] = {"samples": int(wc.n), "filter_pass_rate": float(tb0["n_passed"].mean()) / max(n_w - 1, 1),
                                   "head_refreshes": int(st_w["head_refreshes"]), "steered_to_k_scan": bool(st_w["scan_mode"]),
                                   "device_p50_ms": float(np.median(ms)), "device_p95_ms": float(pct(ms.tolist(), 0.95)),
                                   "e2e_p50_ms": 1e3 * statistics.median(lat[2:]), "e2e_p95_ms": 1e3 * pct(lat[2:], 0.95),
                                   "fullscan_ms": float(np.median(fs["ms_compare"])),
                                   "device_p50_over_fullscan": float(np.median(ms)) / float(np.median(fs["ms_compare"])),
                                   "moved_bytes_per_launch": float(tb["bytes_touched"].mean()),
                                   "neighbours_per_query": float(pr["n_hits"].mean()), "parity_queries_checked": checked,
                                   "setup_s": time.time() - t0}
                we.close()
                del wc, wo
            except Exception as ex:   # pragma: no cover
                errors["workload_" + name] = repr(ex)
                log("workload", name, "failed:", repr(ex))

    # ================================================================================================
    # leg 5 (every N): BASELINE configs[2] — all-vs-all sparse matrix at 12 SNV, block rows over the ranks
    # ================================================================================================
    ava = None
    if "ava" in legs:
        try:
            ava = run_allvsall(args, rank, world, local, workers, barrier, max_over_ranks, sum_over_ranks, peak, peak_src)
        except Exception as ex:   # pragma: no cover
            errors["all_vs_all"] = repr(ex)
            log("all_vs_all leg failed:", repr(ex))
            if world > 1:                        # the other ranks are inside this leg's collectives: fail fast, do not strand them
                raise

    # ================================================================================================
    # leg 6 (every N): BASELINE configs[4] — collection resident across the GPUs, streamed inserts each queried
    # ================================================================================================
    c4 = None
    if "c4" in legs:
        host_barrier()
        if rank == 0:
            try:
                t0 = time.time()
                per = args.c4_samples
                if grp is None and world > 1:
                    grp = GroupEngine(coll.genome_length, list(range(world)))
                    ostore = OracleStore(oracle_threads)
                target = eng if world == 1 else grp
                have = target.stats()["n_rows"]
                want_total = per * world
                fam_at = 0
                while have < want_total:                 # top the store up, 125 000 samples (2 500 families) at a time
                    n_f = min(2500, (want_total - have + 49) // 50)
                    extra = synth.make_families([5_000_000 + fam_at + a for a in range(n_f)], 50, workers=cpus)
                    fam_at += n_f
                    target.append_bulk(extra.pos, extra.off, extra.invalid)
                    ostore.add(extra)
                    have += extra.n
                t_setup = time.time() - t0
                c4q = make_queries(common, args.c4_inserts + WG, seed=777)
                st1 = target.stats()
                r = timed_inserts(lambda s: target.insert_query(s[0], s[1], s[2], ceiling)[1], c4q, st1["n_rows"], ostore,
                                  max(1, args.c4_inserts // 6), warm=WG)
                c4 = {"workload": "configs[4]: %d samples resident across %d GPU(s), %d streamed inserts each queried one-vs-all at "
                                  "snpCeiling %d" % (st1["n_rows"], world, args.c4_inserts, ceiling),
                      "samples": int(st1["n_rows"]), "samples_per_gpu": per, "inserts": args.c4_inserts,
                      "p50_ms": 1e3 * statistics.median(r["lat"]), "p95_ms": 1e3 * pct(r["lat"], 0.95),
                      "inserts_per_s": args.c4_inserts / r["seconds"], "comparisons_per_s": r["comparisons"] / r["seconds"],
                      "neighbours_per_query": r["neighbours"] / args.c4_inserts, "parity_rows_checked": r["parity_checked"],
                      "slowest_steps_ms": r["slowest"],
                      "store_bytes_per_gpu": int(target.stats()["store_bytes"] // world), "setup_s": t_setup}
                # the STREAM kernel over ONE device's share of this store (2.5 x the headline shard): launch + query build,
                # ~13 us of every k_scan launch whatever the store holds, weigh less the more rows a GPU holds
                try:
                    big = {}
                    qrows = list(range(1000, 1000 + min(K, 24) + 3))
                    for name, c in (("ceiling_250", 250), ("no_ceiling", 2 ** 31 - 1)):
                        target.profile_queries(qrows[:3], c, flush_l2=True)
                        pr = target.profile_queries(qrows[3:], c, flush_l2=True, count_bytes=True)
                        ms = float(np.median(pr["ms_compare"]))
                        mv = float(pr["bytes_touched"].mean())
                        big[name] = {"ceiling": c, "launch_ms": ms, "launch_ms_p95": float(pct(pr["ms_compare"].tolist(), 0.95)),
                                     "moved_bytes_per_launch": mv, "achieved": mv / (ms * 1e-3) / 1e9,
                                     "frac": mv / (ms * 1e-3) / 1e9 / peak, "neighbours_per_query": float(pr["n_hits"].mean())}
                    c4["scan_one_device"] = dict(big, bound="hbm", unit="GB/s", peak=peak, rows=int(per),
                                                 kernel="k_scan over the %d rows one device holds of this store" % per)
                except Exception as ex:   # pragma: no cover
                    errors["config4_scan"] = repr(ex)
                    log("config4 scan block failed:", repr(ex))
            except Exception as ex:   # pragma: no cover
                errors["config4"] = repr(ex)
                log("config4 leg failed:", repr(ex))
        host_barrier()
    if grp is not None:
        grp.close()

    # ---- extras at N = 1: device compress, cluster statistics, start-up reload ---------------------------------
    ingest = cluster = reload_leg = None
    if world == 1 and "extras" in legs:
        try:
            ingest, cluster, reload_leg = extras(eng, coll, local, ceiling)
        except Exception as ex:   # pragma: no cover
            errors["extras"] = repr(ex)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": med_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": {"workload": "configs[1]: 50,000 synthetic TB sequences per GPU, one new-sample one-vs-all query "
                               "per step, snpCeiling %d" % ceiling,
                   "samples_per_gpu": int(coll.n), "genome_length": int(coll.genome_length),
                   "positions_per_sample_mean": float(coll.positions_per_sample()[coll.invalid == 0].mean()),
                   "l2": "flushed (256 MiB write) before every timed step; store %.0f MB > L2" % (store_bytes_rows / 1e6),
                   "parallelism": "candidate shards x%d, no data-path collective in `value`" % world,
                   "statistic": "value = comparisons per step / MEDIAN per-step device time of the slowest rank (K steps)"},
        "value_from_mean": comps_per_step / (mean_ms * 1e-3),
        "step_ms": {"median": med_ms, "mean": mean_ms, "p95": max(s[2] for s in all_stats), "max": max(s[3] for s in all_stats),
                    "slowest_rank": slow,
                    "per_rank": [{"rank": i, "median": s[0], "mean": s[1], "p95": s[2], "max": s[3]} for i, s in enumerate(all_stats)]},
        "value_warm_l2": comps_per_step / (t_warm * 1e-3),
        "one_vs_all_p50_ms_device": med_ms,
        "one_vs_all_p50_ms_e2e": e2e["p50_ms"] if e2e else None,
        "build_ms": float(prof["ms_build"].mean()), "compare_ms": cmp_ms,
        "neighbours_per_query": float(prof["n_hits"].mean()),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(sum_over_ranks(float(launches))),
        "roofline": roofline, "host_cpus": cpus, "setup_s": t_gen,
        "bulk_append": {"rows": int(coll.n), "seconds": t_bulk, "rows_per_s": coll.n / t_bulk,
                        "positions_per_s": float(coll.pos.size) / t_bulk},
    }
    for key, val in (("e2e_class", e2e_class), ("workloads", workloads), ("all_vs_all", ava), ("config4", c4), ("ingest", ingest),
                     ("cluster_statistics", cluster), ("reload_from_pickles", reload_leg)):
        if val is not None:
            line[key] = val
    if errors:
        line["leg_errors"] = errors
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            line["cpu_baseline"] = cpu_python_port(coll, fresh[-40:], ceiling, args.cpu_seconds)
            line["cpu_baseline_native"] = cpu_c_oracle(coll, fresh[-4:], ceiling, args.cpu_seconds, cpus)
        except Exception as ex:   # pragma: no cover
            line["cpu_baseline"] = {"error": str(ex)}
    if rank == 0:
        emit(line)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def run_allvsall(args, rank, world, local, workers, barrier, max_over_ranks, sum_over_ranks, peak, peak_src):
    """BASELINE configs[2]: all-vs-all sparse matrix at 12 SNV (distmat, src/seqComparer.py:174-197).  Every rank holds the
    WHOLE collection (replicated store: 200 000 x ~3 KB = 0.6 GB) and owns the query rows r with r % world == rank
    (SURVEY 8(e)); nothing is exchanged until the edges are gathered on rank 0 (NCCL all-gather of the padded edge blocks,
    inside the timed region).  Each rank generates 1/world of the families; the parts are all-gathered and laid out in
    family order, so the row numbering — and the checksum of the edge set — is the same at every N."""
    import torch
    import torch.distributed as dist
    from findneighbour3_b200 import synth
    from findneighbour3_b200.engine import EDGE_DTYPE, Engine
    max_c = 100
    n_anc = max(world, args.ava_samples // max_c)
    mine = list(range(rank, n_anc, world))
    t0 = time.time()
    part = synth.make_families([300000 + a for a in mine], max_c, workers=workers)
    if world == 1:
        full = part
    else:
        sizes = torch.tensor([part.pos.size, part.invalid.size], dtype=torch.int64, device="cuda")
        allsz = [torch.zeros_like(sizes) for _ in range(world)]
        dist.all_gather(allsz, sizes)
        maxp, maxn = int(max(int(x[0]) for x in allsz)), int(max(int(x[1]) for x in allsz))
        tp = torch.zeros(maxp, dtype=torch.int32, device="cuda"); tp[:part.pos.size] = torch.from_numpy(part.pos).cuda()
        to = torch.zeros(6 * maxn + 1, dtype=torch.int64, device="cuda"); to[:part.off.size] = torch.from_numpy(part.off).cuda()
        ti = torch.zeros(maxn, dtype=torch.uint8, device="cuda"); ti[:part.invalid.size] = torch.from_numpy(part.invalid).cuda()
        gp = [torch.zeros_like(tp) for _ in range(world)]
        go = [torch.zeros_like(to) for _ in range(world)]
        gi = [torch.zeros_like(ti) for _ in range(world)]
        dist.all_gather(gp, tp); dist.all_gather(go, to); dist.all_gather(gi, ti)
        parts = []
        for r in range(world):
            npos, nn = int(allsz[r][0]), int(allsz[r][1])
            parts.append((gp[r][:npos].cpu().numpy(), go[r][:6 * nn + 1].cpu().numpy(), gi[r][:nn].cpu().numpy()))
        del gp, go, gi, tp, to, ti
        torch.cuda.empty_cache()
        # family a is block a // world of rank a % world's part: lay the families out in the order a = 0, 1, 2, ...
        pos_chunks, len_chunks, inv_chunks = [], [], []
        for a in range(n_anc):
            ppos, poff, pinv = parts[a % world]
            k = a // world
            lo, hi = k * max_c, (k + 1) * max_c
            pos_chunks.append(ppos[poff[6 * lo]:poff[6 * hi]])
            len_chunks.append(np.diff(poff[6 * lo:6 * hi + 1]))
            inv_chunks.append(pinv[lo:hi])
        lens = np.concatenate(len_chunks)
        off = np.zeros(lens.size + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        full = synth.Collection(synth.TB_LENGTH, np.concatenate(pos_chunks), off, np.concatenate(inv_chunks),
                                np.repeat(np.arange(n_anc, dtype=np.int32), max_c))
        del parts, pos_chunks
    eng = Engine(synth.TB_LENGTH, local)
    eng.append_bulk(full.pos, full.off, full.invalid)
    st = eng.stats()
    n = st["n_rows"]
    t_setup = time.time() - t0
    log("all-vs-all: rank", rank, "holds", n, "rows after %.1f s" % t_setup)
    eng.all_vs_all(args.ava_ceiling, upper_only=True, row_begin=0, row_end=min(n, 4096), stride=world, phase=rank)     # warm-up

    pinned = {}

    def gather_edges(edges):
        """every rank's edges -> list of per-rank arrays on rank 0 (None elsewhere): ONE padded NCCL gather into a single
        device buffer, one copy into pinned host memory (allocated by the first, untimed pass), no concatenation"""
        if world == 1:
            return [edges]
        isz = EDGE_DTYPE.itemsize
        cnt = torch.tensor([len(edges)], dtype=torch.int64, device="cuda")
        cnts = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(cnts, cnt)
        cnts = [int(c) for c in cnts]
        mx = max(max(cnts), 1)
        buf = torch.empty(mx * isz, dtype=torch.uint8, device="cuda")
        raw = torch.from_numpy(edges.view(np.uint8).reshape(-1))
        buf[:raw.numel()].copy_(raw, non_blocking=True)
        recv = torch.empty(world * mx * isz, dtype=torch.uint8, device="cuda") if rank == 0 else None
        dist.gather(buf, list(recv.view(world, -1).unbind(0)) if rank == 0 else None, dst=0)
        if rank != 0:
            return None
        if pinned.get("n", 0) < recv.numel():
            pinned["t"] = torch.empty(recv.numel(), dtype=torch.uint8).pin_memory()
            pinned["n"] = recv.numel()
        host = pinned["t"][:recv.numel()]
        host.copy_(recv, non_blocking=True)
        torch.cuda.synchronize()
        h = host.numpy()
        return [h[r * mx * isz: r * mx * isz + cnts[r] * isz].view(EDGE_DTYPE) for r in range(world)]

    runs = []
    for rep in range(2):                              # the first pass also sizes the edge buffers (device, pinned, numpy)
        barrier()
        launches0 = eng.launch_count()
        t1 = time.perf_counter()
        edges = eng.all_vs_all(args.ava_ceiling, upper_only=True, stride=world, phase=rank)
        t_mine = time.perf_counter() - t1
        all_edges = gather_edges(edges)
        torch.cuda.synchronize()
        t = time.perf_counter() - t1
        barrier()
        runs.append((max_over_ranks(t), max_over_ranks(t_mine)))
        launches = sum_over_ranks(float(eng.launch_count() - launches0))
        log("all-vs-all pass", rep, "rank", rank, "%.4f s (compute %.4f s)" % (t, t_mine))
    t, t_compute = runs[-1]
    # parity: sampled rows against the C oracle over all candidates (each rank checks the sampled rows it owns)
    from oracle import c_oracle
    rng = np.random.default_rng(2024)
    sample_rows = sorted(int(x) for x in rng.choice(n, size=min(50, n), replace=False))
    by_row = {}
    mine_rows = [r for r in sample_rows if r % world == rank]
    if mine_rows:
        sel = np.isin(edges["row1"], np.array(mine_rows, dtype=np.int64))
        for e in edges[sel]:
            by_row.setdefault(int(e["row1"]), {})[int(e["row2"])] = (int(e["dist"]), int(e["n1"]), int(e["n2"]), int(e["nboth"]))
    checked = bad = 0
    for r in mine_rows:
        want = c_oracle.one_vs_all(full.pos, full.off, full.invalid, r, args.ava_ceiling, nthreads=max(1, workers))
        want = {k: v for k, v in want.items() if k > r}
        if by_row.get(r, {}) != want:            # (no assert here: a rank that left the leg early would strand the others
            bad += 1                             #  in the collectives below)
            log("all-vs-all parity: row", r, "differs from the C oracle on rank", rank)
        checked += 1
    checked = int(sum_over_ranks(float(checked)))
    bad = int(sum_over_ranks(float(bad)))
    pairs = n * (n - 1) / 2
    b_pair = (4.0 * st["positions"] + 32.0 * st["n_live"]) / max(st["n_live"], 1)
    out = None
    if rank == 0:
        parts_cs = [edge_checksum(e_) for e_ in all_edges]
        cs = {"sum": sum(c["sum"] for c in parts_cs) % (1 << 64), "xor": 0}
        for c in parts_cs:
            cs["xor"] ^= c["xor"]
        n_edges_all = int(sum(len(e_) for e_ in all_edges))
        out = {"workload": "configs[2]: %d synthetic TB sequences all-vs-all at %d SNV, store replicated, query rows r %% %d == rank"
                           % (n, args.ava_ceiling, world),
               "samples": int(n), "ceiling": args.ava_ceiling, "n_gpus": world, "scaling": "strong",
               "seconds": t, "seconds_compute_max_rank": t_compute, "seconds_first_pass": runs[0][0],
               "edges": n_edges_all, "pairs": int(pairs),
               "comparisons_per_s": pairs / t, "edge_checksum": cs, "parity_rows_checked": checked, "parity_rows_wrong": bad,
               "gpu_launches": int(launches), "setup_s": t_setup,
               "canonical_GBps": pairs * b_pair / t / 1e9, "frac_canonical": pairs * b_pair / t / 1e9 / (peak * world),
               "note": "seconds = max over ranks of (fn3_all_vs_all of the rank's block rows + gather of all edges on rank 0); the "
                       "tile kernel works on 96-byte heads resident in L2 (ncu: ~2 MB of DRAM per 148-query tile), so the canonical "
                       "figure is not a bandwidth claim"}
    eng.close()
    del full
    return out


def extras(eng, coll, local, ceiling):
    """N = 1 extras: device compress / ingest, cluster statistics, start-up reload from the stored pickles"""
    import pickle
    from findneighbour3_b200 import packing, synth
    from findneighbour3_b200.engine import byte_histogram
    ref_codes = synth.random_reference(coll.genome_length)
    mask = synth.synthetic_mask(synth.TB_LENGTH, synth.TB_EXCLUDED, synth.TB_EXCLUDED_RUNS)
    eng.set_reference(synth.reference_string(ref_codes), np.flatnonzero(mask).tolist())
    seqs = []
    for i in np.flatnonzero(coll.invalid == 0)[:8]:
        p_, o_, _ = coll.lists(int(i))
        seqs.append(synth.to_string(ref_codes, p_, o_).encode("ascii"))
    for sq in seqs[:2]:
        eng.compress(sq)
    t0 = time.perf_counter()
    for rep in range(5):
        for sq in seqs:
            eng.compress(sq)
    t_c = (time.perf_counter() - t0) / (5 * len(seqs))
    eng.ingest_query(seqs[0], 130000, ceiling)
    byte_histogram(seqs[0], local)
    t0 = time.perf_counter()
    for sq in seqs:
        eng.ingest_query(sq, 130000, ceiling)
    t_i = (time.perf_counter() - t0) / len(seqs)
    t0 = time.perf_counter()
    for sq in seqs:
        byte_histogram(sq, local)
    t_h = (time.perf_counter() - t0) / len(seqs)
    ingest = {"compress_ms_per_sequence": 1e3 * t_c, "sequence_bytes": int(coll.genome_length),
              "GBps_host_to_lists": coll.genome_length / t_c / 1e9,
              "ingest_query_ms_per_sequence": 1e3 * t_i, "byte_histogram_ms_per_sequence": 1e3 * t_h,
              "note": "fn3_compress through ctypes: 4.4 MB sequence host->device, 3 kernels, lists device->host; "
                      "ingest_query = composition histogram + compress + persist + one-vs-all from the same copy "
                      "(fn3_ingest_query); byte_histogram = NucleicAcid.examine's counts (fn3_byte_histogram)"}
    fam = np.flatnonzero((coll.family == coll.family[0]) & (coll.invalid == 0))[:50]
    rows250 = np.flatnonzero(coll.invalid == 0)[:250]
    sites = eng.msa_sites(fam)
    eng.msa_align(fam, sites)
    eng.consensus(rows250, 200)
    reps = 10
    t0 = time.perf_counter()
    for _ in range(reps):
        sites = eng.msa_sites(fam)
    t_sites = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        eng.msa_align(fam, sites)
    t_align = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        cpos, coff = eng.consensus(rows250, 200)
    t_cons = (time.perf_counter() - t0) / reps
    L = int(coll.genome_length)
    cluster = {"msa_rows": int(fam.size), "msa_sites": int(sites.size), "msa_sites_ms": 1e3 * t_sites,
               "msa_align_ms": 1e3 * t_align, "consensus_rows": int(rows250.size), "consensus_min_votes": 200,
               "consensus_positions": int(cpos.size), "consensus_ms": 1e3 * t_cons,
               "msa_sites_GBps": 4 * 4 * L * 3 / t_sites / 1e9, "consensus_GBps": 5 * 4 * L * 3 / t_cons / 1e9,
               "note": "fn3_msa_sites / fn3_msa_align / fn3_consensus through ctypes, wall clock per call incl. host "
                       "synchronisations; plane bytes = vote planes cleared once and scanned twice (count, write)"}
    idx = np.flatnonzero(coll.invalid == 0)[:2000]
    blobs = [pickle.dumps(coll.as_dict(int(i)), protocol=2) for i in idx]
    packing.unpickle_many(blobs[:64])
    t0 = time.perf_counter()
    p_n, o_n, st_n = packing.unpickle_many(blobs)
    t_native = time.perf_counter() - t0
    t0 = time.perf_counter()
    p_p, o_p, inv_p = packing.pack_many([pickle.loads(b) for b in blobs])
    t_python = time.perf_counter() - t0
    assert np.array_equal(p_n, p_p) and np.array_equal(o_n, o_p) and not st_n.any()
    reload_leg = {"samples": int(idx.size), "pickle_bytes_mean": float(np.mean([len(b) for b in blobs])),
                  "native_samples_per_s": idx.size / t_native, "python_samples_per_s": idx.size / t_python,
                  "speedup": t_python / t_native, "host_threads": min(16, os.cpu_count() or 1),
                  "note": "packing.unpickle_many (fn3_unpickle_samples) vs pickle.loads + packing.pack_many; identical output"}
    return ingest, cluster, reload_leg


if __name__ == "__main__":
    main()
