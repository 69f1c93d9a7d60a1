#!/usr/bin/env python3
"""bench.py — headline benchmark of the masked pairwise SNV distance path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--samples S] [--ceiling C]

Workload (config.workload): BASELINE.json configs[1] — a 50 000-sample synthetic M. tuberculosis-shaped
collection (1 000 ancestors x 50, L = 4 411 532, TB-exclude-shaped mask, ~480 variant + ~4 000 blocky N
positions per sample, 1 % invalid), one new-sample one-vs-all neighbour query per step at snpCeiling 20.
With N GPUs every rank holds its OWN 50 000-sample shard (weak scaling: the collection grows with N) and
every query is compared against all shards.

  value     pairwise comparisons/s, device-resident: the K query samples are already rows of the HBM store;
            a step = query-map build + compare kernel, timed with CUDA events on the engine's stream, L2
            flushed (256 MiB write, untimed) before every step; max over ranks.
  e2e       the same metric through the C-ABI a user of the reference class would sit on
            (persist -> fn3_append, mcompare -> fn3_one_vs_all) with HOST buffers: every step copies the new
            sample's packed lists host->device and the neighbour records device->host; wall clock, max over ranks.
  roofline  dominant kernel = k_query.  achieved = algorithmic bytes / mean CUDA-event duration of the
            compare launch, algorithmic bytes = sum over compared candidates of B_pair = 4*(|A|+|C|+|G|+|T|+|N|+|M|)+32
            (SURVEY.md §8(d)).  Early exit and N-run coding make the kernel touch far fewer bytes than that,
            so `frac` can exceed 1; `touched_*` and `fullscan` (ceiling = 2^31-1: no pair may exit early,
            every stored byte is streamed) are reported beside it.  See DESIGN.md §6.
"""
import argparse
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
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--samples", type=int, default=50000, help="samples per GPU shard")
    ap.add_argument("--ceiling", type=int, default=20)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-allvsall", dest="allvsall", action="store_false", help="skip the extra all-vs-all leg")
    ap.add_argument("--workload", default="onevsall", choices=["onevsall", "allvsall"],
                    help="allvsall = BASELINE configs[2]: replicated store, block rows per rank (not the headline line)")
    ap.add_argument("--ava-ceiling", type=int, default=12)
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
    """per-launch DRAM bytes of k_query from the committed ncu capture, if any (profiles/traffic.json)."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path))
        except Exception:
            return None
    return None


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed regions (NVML, every 2 ms, own thread);
    the equivalent of the profiling recipe's nvidia-smi clocks line without its start-up delay."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = threading.Event()
        self.thread = None
        self.h = None
        self.max_mhz = None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES may remap ordinals: match by PCI bus id of the torch device
            import torch
            bus = torch.cuda.get_device_properties(index).pci_bus_id
            dom = torch.cuda.get_device_properties(index).pci_domain_id
            dev = torch.cuda.get_device_properties(index).pci_device_id
            busid = "%08X:%02X:%02X.0" % (dom, bus, dev)
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
            time.sleep(0.002)

    def start(self):
        if self.h is None:
            return
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


COMMON_ANC = 20      # families every rank can regenerate quickly: the new samples are their relatives


def _mask():
    from findneighbour3_b200 import synth
    return synth.synthetic_mask(synth.TB_LENGTH, synth.TB_EXCLUDED, synth.TB_EXCLUDED_RUNS)


def make_common():
    from findneighbour3_b200 import synth
    return synth.make_collection(COMMON_ANC, 50, seed=777, excluded=_mask())


def make_shard(samples, rank, common):
    """rank's shard of `samples` samples; rank 0's shard ends with the `common` families."""
    from findneighbour3_b200 import synth
    max_c = 50
    n_anc = max(1, samples // max_c)
    if rank == 0:
        own = max(0, n_anc - COMMON_ANC)
        if own == 0:
            return common
        return synth.concat(synth.make_collection(own, max_c, seed=2, excluded=_mask()), common)
    return synth.make_collection(n_anc, max_c, seed=2 + 1000 * rank, excluded=_mask())


def make_queries(common, count, seed=12345):
    """`count` newly sequenced samples: relatives (0-3 extra SNVs) of random valid samples of the common
    families — identical on every rank."""
    from findneighbour3_b200 import synth
    rng = np.random.default_rng(seed)
    valid = np.flatnonzero(common.invalid == 0)
    out = []
    for k in range(count):
        i = int(rng.choice(valid))
        out.append(synth.new_sample_like(common, i, int(rng.integers(0, 4)), seed=seed + k))
    return out


# ----------------------------------------------------------------------------------------------------
# CPU legs (oracle = test/bench infrastructure; never on the product path)
# ----------------------------------------------------------------------------------------------------
def cpu_python_port(coll, queries, ceiling, budget_s):
    """the reference's algorithm (Python set algebra, single thread — the class the server imports is
    single-threaded and GIL-bound) on a bounded sample: the query against the first `m` stored samples."""
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
    sub = coll
    t_used, pairs, nq = 0.0, 0, 0
    # the query is appended as the last row of a private copy of the store arrays
    for (p, o, inv) in queries:
        pos = np.concatenate([sub.pos, p])
        off = np.concatenate([sub.off, sub.off[-1] + o[1:].astype(np.int64)])
        invalid = np.concatenate([sub.invalid, np.array([1 if inv else 0], np.uint8)])
        t0 = time.perf_counter()
        c_oracle.one_vs_all(pos, off, invalid, sub.n, ceiling, nthreads=threads)
        t_used += time.perf_counter() - t0
        pairs += sub.n
        nq += 1
        if t_used > budget_s:
            break
    return {"value": pairs / t_used, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d queries x all %d candidates, oracle/fn3_oracle.c (C restatement, %d pthreads), %.1f s"
                      % (nq, sub.n, threads, t_used)}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path, timed on this host.
    The reference is pure Python and cannot travel to the GPU box, so its restatement (oracle port) runs."""
    if rank != 0:
        return
    common = make_common()
    coll = make_shard(min(args.samples, 5000), 0, common)
    queries = make_queries(common, max(args.steps + args.warmup, 4))
    budget = max(5.0, min(90.0, args.cpu_seconds * 3))
    # warm-up steps then timed steps; a step = one query against a bounded candidate sample
    from oracle import seqcomparer_port as port
    from findneighbour3_b200 import packing
    m = min(coll.n, 1500)
    store = {i: coll.as_dict(i) for i in range(coll.n - m, coll.n)}     # includes the queries' own families
    for (p, o, inv) in queries[:min(args.warmup, 1)]:
        store["q"] = {"invalid": 1} if inv else packing.unpack_lists(p, o, None)
        port.mcompare(store, {}, "q", args.ceiling)
    t_used, steps = 0.0, 0
    for (p, o, inv) in queries[args.warmup:args.warmup + args.steps]:
        store["q"] = {"invalid": 1} if inv else packing.unpack_lists(p, o, None)
        t0 = time.perf_counter()
        port.mcompare(store, {}, "q", args.ceiling)
        t_used += time.perf_counter() - t0
        steps += 1
        if t_used > budget:
            break
    value = steps * m / t_used
    sample = ("%d steps, each one query x %d candidates (bounded sample of the 50k workload), single thread: "
              "Python-set restatement of seqComparer.mcompare (oracle/seqcomparer_port.py)" % (steps, m))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * t_used / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": "configs[1]: 50,000 synthetic TB sequences per GPU, one new-sample one-vs-all query "
                                   "per step, snpCeiling %d" % args.ceiling,
                       "samples_per_gpu": int(args.samples), "genome_length": int(coll.genome_length),
                       "candidates_per_step": m,
                       "sample": "each step compares the query with a bounded sample of %d of the shard's candidates "
                                 "(the reference's cost is linear in the candidates)" % m},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "host_cpus": os.cpu_count()}
    try:
        line["cpu_baseline_native"] = cpu_c_oracle(coll, queries[:3], args.ceiling, 10.0, os.cpu_count() or 1)
    except Exception as ex:       # pragma: no cover
        line["cpu_baseline_native"] = {"error": str(ex)}
    emit(line)


def run_allvsall(args, rank, world, local, barrier, max_over_ranks, sum_over_ranks):
    """BASELINE configs[2]: all-vs-all sparse matrix at 12 SNV.  Every rank holds the WHOLE collection (replicated
    store) and owns the query rows r with r % world == rank; no exchange until the edge counts are
    summed.  Each rank generates 1/world of the families; the shards are all-gathered over NCCL."""
    import torch
    import torch.distributed as dist
    from findneighbour3_b200 import synth
    from findneighbour3_b200.engine import Engine
    max_c = 100
    n_anc = max(world, args.samples // max_c)
    mine = list(range(rank, n_anc, world))
    t0 = time.time()
    part = synth.make_families([300000 + a for a in mine], max_c, workers=max(1, (os.cpu_count() or 1) // world))
    pos, off, inv = part.pos, part.off, part.invalid
    eng = Engine(synth.TB_LENGTH, local)
    if world == 1:
        eng.append_bulk(pos, off, inv)
    else:
        sizes = torch.tensor([pos.size, inv.size], dtype=torch.int64, device="cuda")
        allsz = [torch.zeros_like(sizes) for _ in range(world)]
        dist.all_gather(allsz, sizes)
        maxp, maxn = int(max(int(x[0]) for x in allsz)), int(max(int(x[1]) for x in allsz))
        tp = torch.zeros(maxp, dtype=torch.int32, device="cuda"); tp[:pos.size] = torch.from_numpy(pos).cuda()
        to = torch.zeros(6 * maxn + 1, dtype=torch.int64, device="cuda"); to[:off.size] = torch.from_numpy(off).cuda()
        ti = torch.zeros(maxn, dtype=torch.uint8, device="cuda"); ti[:inv.size] = torch.from_numpy(inv).cuda()
        gp = [torch.zeros_like(tp) for _ in range(world)]
        go = [torch.zeros_like(to) for _ in range(world)]
        gi = [torch.zeros_like(ti) for _ in range(world)]
        dist.all_gather(gp, tp); dist.all_gather(go, to); dist.all_gather(gi, ti)
        for r in range(world):
            npos, nn = int(allsz[r][0]), int(allsz[r][1])
            eng.append_bulk(gp[r][:npos].cpu().numpy(), go[r][:6 * nn + 1].cpu().numpy(), gi[r][:nn].cpu().numpy())
        del gp, go, gi, tp, to, ti
        torch.cuda.empty_cache()
    st = eng.stats()
    n = st["n_rows"]
    t_setup = time.time() - t0
    # warm-up on a slice of the rows, then the timed full pass over this rank's block rows
    eng.all_vs_all(args.ava_ceiling, upper_only=True, row_begin=0, row_end=min(n, 2048), stride=world, phase=rank)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    launches0 = eng.launch_count()
    t1 = time.perf_counter()
    edges = eng.all_vs_all(args.ava_ceiling, upper_only=True, stride=world, phase=rank)
    torch.cuda.synchronize()
    t = time.perf_counter() - t1
    barrier()
    clocks = sampler.stop()
    t = max_over_ranks(t)
    n_edges = sum_over_ranks(float(len(edges)))
    launches = sum_over_ranks(float(eng.launch_count() - launches0))
    pairs = n * (n - 1) / 2
    b_pair = (4.0 * st["positions"] + 32.0 * st["n_live"]) / max(st["n_live"], 1)      # mean canonical bytes per comparison
    peak, peak_src = peaks()
    line = {"metric": "pairwise SNV comparisons/sec (all-vs-all sparse matrix at %d SNV, block rows over %d GPU)" % (args.ava_ceiling, world),
            "value": pairs / t, "unit": UNIT, "n_gpus": world, "steps": 1, "warmup": 1, "ms_per_step": 1e3 * t,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": "configs[2]: %d synthetic TB sequences all-vs-all at %d SNV, store replicated, query rows r %% %d == rank"
                                   % (n, args.ava_ceiling, world), "samples": int(n), "genome_length": int(synth.TB_LENGTH),
                       "l2": "head tiles %.0f MB + rows %.0f MB per GPU" % (n * 96 / 1e6, 4 * st["store_words"] / 1e6)},
            "edges": int(n_edges), "pairs": int(pairs), "seconds": t, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_query", "achieved": pairs * b_pair / t / 1e9, "peak": peak * world,
                         "peak_source": peak_src + " x n_gpus", "unit": "GB/s", "frac": pairs * b_pair / t / 1e9 / (peak * world),
                         "traffic": None,
                         "note": "canonical B_pair bytes (SURVEY 8(d)); the FILTER variant decides almost every pair from its 96-byte head"},
            "setup_s": t_setup, "host_cpus": os.cpu_count()}
    if rank == 0:
        emit(line)
    eng.close()


# ----------------------------------------------------------------------------------------------------
def emit(line):
    """the ONE JSON line goes to the real stdout; everything else that writes to fd 1 (NCCL's version banner,
    library chatter) was redirected to stderr at start-up."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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

    from findneighbour3_b200.engine import Engine

    if args.workload == "allvsall":
        run_allvsall(args, rank, world, local, barrier, max_over_ranks, sum_over_ranks)
        if world > 1:
            dist.destroy_process_group()
        return

    t_gen = time.time()
    common = make_common()
    coll = make_shard(args.samples, rank, common)
    eng = Engine(coll.genome_length, local)
    t_bulk = time.perf_counter()
    eng.append_bulk(coll.pos, coll.off, coll.invalid)       # server start-up reload path: validate, pack, copy, place heads
    t_bulk = time.perf_counter() - t_bulk
    K, W = args.steps, args.warmup
    queries = make_queries(common, 4 * (K + W))      # the same new samples on every rank
    t_gen = time.time() - t_gen

    # ---- leg 1: device-resident value ------------------------------------------------------------
    resident = [eng.append(p, o, inv) for (p, o, inv) in queries[:K + W]]
    n_rows = eng.stats()["n_rows"]
    bytes_canon_all = int(coll.bytes_canonical()[coll.invalid == 0].sum())
    for (p, o, inv) in queries[:K + W]:
        if not inv:
            bytes_canon_all += 4 * int(o[6]) + 32
    sampler = ClockSampler(local)
    barrier()
    eng.profile_queries(resident[:W], args.ceiling, flush_l2=True)
    barrier()
    sampler.start()
    launches0 = eng.launch_count()
    prof = eng.profile_queries(resident[W:], args.ceiling, flush_l2=True)
    launches = eng.launch_count() - launches0 - K      # minus the untimed L2-flush launches
    barrier()
    warm = eng.profile_queries(resident[W:], args.ceiling, flush_l2=False)
    full = eng.profile_queries(resident[W:W + min(K, 20)], 2 ** 31 - 1, flush_l2=True, count_bytes=True)
    touched = eng.profile_queries(resident[W:W + min(K, 20)], args.ceiling, flush_l2=False, count_bytes=True)
    barrier()
    t_dev = float(prof["ms_build"].sum() + prof["ms_compare"].sum()) / 1e3
    t_dev = max_over_ranks(t_dev)
    comps_per_step = sum_over_ranks(float(n_rows - 1))          # every other row of every shard
    value = comps_per_step * K / t_dev
    t_warm = max_over_ranks(float(warm["ms_build"].sum() + warm["ms_compare"].sum()) / 1e3)

    peak, peak_src = peaks()
    cmp_ms = float(prof["ms_compare"].mean())
    st = eng.stats()
    store_bytes_rows = 4 * st["store_words"] + 32 * st["n_rows"]
    achieved = bytes_canon_all / (cmp_ms * 1e-3) / 1e9
    traffic = ncu_traffic()
    roofline = {
        "bound": "hbm", "kernel": "k_query", "achieved": achieved, "peak": peak, "peak_source": peak_src,
        "unit": "GB/s", "frac": achieved / peak,
        "traffic": None if not traffic else traffic.get("k_query_bytes_per_launch"),
        "launch_ms": cmp_ms,
        "algorithmic_bytes_per_launch": bytes_canon_all,
        "store_bytes_per_launch": store_bytes_rows,
        "touched_bytes_per_launch": float(touched["bytes_touched"].mean()),
        "touched_GBps": float(touched["bytes_touched"].mean()) / (cmp_ms * 1e-3) / 1e9,
        "note": "frac uses SURVEY §8(d) canonical B_pair bytes; early exit + N-run coding let the kernel skip most "
                "of them (touched_bytes_per_launch), so frac > 1 is expected here; `fullscan` is the same kernel "
                "with early exit impossible",
    }
    full_ms = float(full["ms_compare"].mean())
    roofline["fullscan"] = {
        "ceiling": 2 ** 31 - 1, "launch_ms": full_ms,
        "achieved_canonical": bytes_canon_all / (full_ms * 1e-3) / 1e9,
        "frac_canonical": bytes_canon_all / (full_ms * 1e-3) / 1e9 / peak,
        "touched_bytes_per_launch": float(full["bytes_touched"].mean()),
        "achieved_touched": float(full["bytes_touched"].mean()) / (full_ms * 1e-3) / 1e9,
        "frac_touched": float(full["bytes_touched"].mean()) / (full_ms * 1e-3) / 1e9 / peak,
        "comparisons_per_s": (n_rows - 1) / (full_ms * 1e-3),
    }

    # ---- leg 2: end to end through the C-ABI with host buffers -----------------------------------
    fresh = queries[K + W:]
    ceiling = args.ceiling

    sharded = {"store": None}

    def make_sharded(exchange):
        from findneighbour3_b200.sharded import ShardedStore
        if sharded["store"] is not None:
            sharded["store"].close()
        rows = eng.stats()["n_rows"]
        sharded["store"] = ShardedStore(eng, local_rows=rows, exchange=exchange)

    def step_fused(sample):
        if sharded["store"] is not None:            # N ranks: sample out, local kernels, neighbour blocks back
            return sharded["store"].insert_query(sample if rank == 0 else None, ceiling)[1]
        p, o, inv = sample                          # server insert(): persist + mcompare, one ABI call
        return eng.insert_query(p, o, inv, ceiling)[1]

    def step_two_calls(sample):
        p, o, inv = sample
        row = eng.append(p, o, inv)                 # persist: pack -> pinned staging -> H2D
        return eng.one_vs_all(row, ceiling)         # mcompare: build + compare -> D2H of the neighbour records

    def run_e2e(step, samples):
        lat = []
        for s in samples[:W]:
            step(s)
        st0 = eng.stats()
        barrier()
        t0 = time.perf_counter()
        n_neighbours = 0
        comps = 0
        rows_now = st0["n_rows"]
        for s in samples[W:W + K]:
            t1 = time.perf_counter()
            hits = step(s)
            lat.append(time.perf_counter() - t1)
            n_neighbours += len(hits)
            comps += rows_now                        # candidates this rank compared in this step
            rows_now = eng._n_rows
        torch.cuda.synchronize()
        t = time.perf_counter() - t0
        barrier()
        st1 = eng.stats()
        t = max_over_ranks(t)
        comps = sum_over_ranks(float(comps))
        return {"value": comps / t, "unit": UNIT,
                "h2d_bytes_per_step": (st1["h2d_bytes"] - st0["h2d_bytes"]) / K,
                "d2h_bytes_per_step": (st1["d2h_bytes"] - st0["d2h_bytes"]) / K,
                "ms_per_step": 1e3 * t / K,
                "p50_ms": 1e3 * statistics.median(lat), "p95_ms": 1e3 * sorted(lat)[int(0.95 * (len(lat) - 1))],
                "neighbours_per_query": n_neighbours / K}

    third = len(fresh) // 3
    e2e_two = None
    if world == 1:
        e2e_two = run_e2e(step_two_calls, fresh[:third])
        e2e_two["call"] = "fn3_append + fn3_one_vs_all via ctypes (seqComparer.persist, then seqComparer.mcompare)"
        e2e = run_e2e(step_fused, fresh[third:2 * third])
        e2e["call"] = "fn3_insert_query via ctypes (server insert(): persist + mcompare, findNeighbour3-server.py:516+530)"
        e2e["two_calls"] = e2e_two
    else:
        def pad_to_world(samples):                  # keep the shards the same size on every rank
            extra = (-(W + K)) % world
            for s_ in samples[:extra]:
                step_fused(s_)
        make_sharded("nccl")
        e2e_nccl = run_e2e(step_fused, fresh[:third])
        pad_to_world(fresh[2 * third:2 * third + 16])
        e2e_nccl["call"] = ("ShardedStore(exchange='nccl').insert_query: NCCL broadcast of the sample, fn3_insert_query on the "
                            "owner rank / fn3_lists_vs_all on the others, NCCL all-gather of the neighbour blocks")
        make_sharded("shm")
        e2e = run_e2e(step_fused, fresh[third:2 * third])
        e2e["call"] = ("ShardedStore(exchange='shm').insert_query: the sample and the neighbour blocks travel through a shared "
                       "host-memory segment (one box); fn3_insert_query on the owner rank / fn3_lists_vs_all on the others")
        e2e["nccl"] = e2e_nccl
        sharded["store"].close()
    clocks = sampler.stop()

    # ---- leg 3 (extra, not the headline): all-vs-all sparse matrix of the shard at 12 SNV (configs[2] shape) ----
    ava = None
    if args.allvsall and world == 1:
        torch.cuda.synchronize()
        n_now = eng.stats()["n_rows"]
        eng.all_vs_all(12, upper_only=True, row_end=min(n_now, 8192))     # warm-up: edge buffers, first launches
        t_ava = 0.0
        for _ in range(3):
            t0 = time.perf_counter()
            edges = eng.all_vs_all(12, upper_only=True)
            t_ava += (time.perf_counter() - t0) / 3
        live = n_now - int(coll.invalid.sum())
        ava = {"samples": int(n_now), "ceiling": 12, "seconds": t_ava, "edges": int(len(edges)),
               "pairs": n_now * (n_now - 1) // 2, "comparisons_per_s": n_now * (n_now - 1) / 2 / t_ava,
               "canonical_GBps": (bytes_canon_all / max(live, 1)) * (n_now * (n_now - 1) / 2) / t_ava / 1e9,
               "note": "fn3_all_vs_all, upper triangle, tiles of one query per SM queued 32 at a time; wall clock incl. the copy of the edges to the host, mean of 3 runs after one partial warm-up run"}

    # ---- leg 4 (extra): device compress() of full-length sequence strings --------------------
    ingest = None
    if args.allvsall and world == 1:
        from findneighbour3_b200 import synth
        ref_codes = synth.random_reference(coll.genome_length)
        eng.set_reference(synth.reference_string(ref_codes), np.flatnonzero(_mask()).tolist())
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
        # the server's whole insert data path from the sequence string (findNeighbour3-server.py:511-530) in one call:
        # composition histogram + compress + maxNs decision + persist + one-vs-all (fn3_ingest_query); and the
        # stand-alone device histogram behind the NucleicAcid drop-in (fn3_byte_histogram)
        from findneighbour3_b200.engine import byte_histogram
        eng.ingest_query(seqs[0], 130000, args.ceiling)
        byte_histogram(seqs[0], local)
        t0 = time.perf_counter()
        for sq in seqs:
            eng.ingest_query(sq, 130000, args.ceiling)
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

    # ---- leg 5 (extra): cluster statistics on the device: consensus vote over 250 stored rows,
    # MSA variant sites + alignment + N/M tallies of a 50-sample cluster -------------------------------------------
    cluster = None
    if args.allvsall and world == 1:
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
                   "plane_bytes_msa": 4 * 4 * L * 3, "plane_bytes_consensus": 5 * 4 * L * 3,
                   "msa_sites_GBps": 4 * 4 * L * 3 / t_sites / 1e9, "consensus_GBps": 5 * 4 * L * 3 / t_cons / 1e9,
                   "note": "fn3_msa_sites / fn3_msa_align / fn3_consensus through ctypes, wall clock per call incl. host "
                           "synchronisations; plane bytes = vote planes cleared once and scanned twice (count, write)"}

    # ---- leg 6 (extra): start-up reload from the reference's stored form (pickle protocol 2 per sample):
    # native parse (fn3_unpickle_samples, host threads) vs pickle.loads + packing of the same 2 000 samples ----------
    reload_leg = None
    if args.allvsall and world == 1:
        import pickle
        from findneighbour3_b200 import packing
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

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": 1e3 * t_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": {"workload": "configs[1]: 50,000 synthetic TB sequences per GPU, one new-sample one-vs-all query "
                               "per step, snpCeiling %d" % ceiling,
                   "samples_per_gpu": int(coll.n), "genome_length": int(coll.genome_length),
                   "positions_per_sample_mean": float(coll.positions_per_sample()[coll.invalid == 0].mean()),
                   "l2": "flushed (256 MiB write) before every timed step; store %.0f MB > L2" % (store_bytes_rows / 1e6),
                   "parallelism": "candidate shards x%d, no data-path collective in `value`" % world},
        "value_warm_l2": comps_per_step * K / t_warm,
        "one_vs_all_p50_ms_device": float(np.median(prof["ms_build"] + prof["ms_compare"])),
        "one_vs_all_p50_ms_e2e": e2e["p50_ms"],
        "build_ms": float(prof["ms_build"].mean()), "compare_ms": cmp_ms,
        "neighbours_per_query": float(prof["n_hits"].mean()),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(sum_over_ranks(float(launches))),
        "roofline": roofline, "host_cpus": os.cpu_count(), "setup_s": t_gen,
        "bulk_append": {"rows": int(coll.n), "seconds": t_bulk, "rows_per_s": coll.n / t_bulk,
                        "positions_per_s": float(coll.pos.size) / t_bulk},
    }
    if ava is not None:
        line["all_vs_all"] = ava
    if ingest is not None:
        line["ingest"] = ingest
    if cluster is not None:
        line["cluster_statistics"] = cluster
    if reload_leg is not None:
        line["reload_from_pickles"] = reload_leg
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            line["cpu_baseline"] = cpu_python_port(coll, fresh[W:], ceiling, args.cpu_seconds)
            line["cpu_baseline_native"] = cpu_c_oracle(coll, fresh[W:W + 4], ceiling, args.cpu_seconds,
                                                       os.cpu_count() or 1)
        except Exception as ex:   # pragma: no cover
            line["cpu_baseline"] = {"error": str(ex)}
    if rank == 0:
        emit(line)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
