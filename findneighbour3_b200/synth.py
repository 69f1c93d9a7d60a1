"""Seeded synthetic "M. tuberculosis-shaped" collections, generated directly in list form.

Semantics follow the reference's generator (src/LargeSequenceSetGenerator.py:64-141 and its
documented TB invocation `1000 500 3 50 1e-8`, doc/demos_technical.md:91):
  * n_anc ancestors = reference + k1 random substitutions (:92-93);
  * every ancestor roots a random tree of max_c nodes; a child = its parent + Poisson(s)
    substitutions (:95-98).  ete3 is not available, so the topology is "node i>0 picks a
    uniform random earlier node of the same family as parent" (SURVEY.md §8(d));
  * substitution rule 'lss': G->T, anything else ->G (:136-139); rule 'any': a uniformly
    random different base (exercises all four base lists);
  * the reference's i.i.d. N model with p=1e-8 (:112-127) yields ~0 N, unlike mapped reads,
    so N is added as `n_blocks` contiguous blocks totalling ~Poisson(n_mean) positions
    (SURVEY.md §8(d)); mixed bases (M) as Poisson(m_mean) scattered positions;
  * a fraction `p_invalid` of samples is all-N, i.e. invalid under maxNs.
Excluded positions are dropped exactly as compress() does (src/seqComparer.py:292).

Everything is numpy; a 50 000-sample TB collection takes a few tens of seconds.
"""
from dataclasses import dataclass
from functools import lru_cache

import numpy as np

TB_LENGTH = 4411532          # src/seqComparer.py:2140
TB_EXCLUDED = 288069         # reference/TB-exclude.txt line count
TB_EXCLUDED_RUNS = 481

_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


@dataclass
class Collection:
    genome_length: int
    pos: np.ndarray          # int32, all samples' lists concatenated
    off: np.ndarray          # int64 [6n+1]
    invalid: np.ndarray      # uint8 [n]
    family: np.ndarray       # int32 [n] ancestor index of each sample

    @property
    def n(self):
        return int(self.invalid.size)

    def lists(self, i):
        """-> (pos int32[], off int32[7], invalid bool) of sample i (engine.append format)."""
        o = self.off[6 * i: 6 * i + 7]
        return self.pos[o[0]:o[6]], (o - o[0]).astype(np.int32), bool(self.invalid[i])

    def as_dict(self, i, m_char="R"):
        """sample i as the reference's compressed_sequence dict (for the oracle port)."""
        if self.invalid[i]:
            return {"invalid": 1}
        p, o, _ = self.lists(i)
        d = {"invalid": 0}
        for b, k in enumerate("ACGTN"):
            d[k] = set(int(x) for x in p[o[b]:o[b + 1]])
        d["M"] = {int(x): m_char for x in p[o[5]:o[6]]}
        return d

    def positions_per_sample(self):
        return (self.off[6::6] - self.off[0:-1:6]).astype(np.int64)

    def bytes_canonical(self):
        """Sum over samples of B_pair = 4*(|A|+|C|+|G|+|T|+|N|+|M|) + 32."""
        return 4 * self.positions_per_sample() + 32

    def subset(self, idx):
        idx = np.asarray(idx, dtype=np.int64)
        chunks, off, total = [], np.zeros(6 * idx.size + 1, dtype=np.int64), 0
        for j, i in enumerate(idx):
            o = self.off[6 * i: 6 * i + 7]
            chunks.append(self.pos[o[0]:o[6]])
            off[6 * j: 6 * j + 7] = total + (o - o[0])
            total += int(o[6] - o[0])
        pos = np.concatenate(chunks) if chunks else np.empty(0, dtype=np.int32)
        return Collection(self.genome_length, pos, off, self.invalid[idx].copy(), self.family[idx].copy())


@lru_cache(maxsize=4)
def random_reference(length, seed=962):
    """uint8 codes 0..3 (A,C,G,T): seeded stand-in for the missing NC_000962 FASTA (read-only, cached)."""
    ref = np.random.default_rng(seed).integers(0, 4, size=length, dtype=np.uint8)
    ref.setflags(write=False)
    return ref


def concat(a, b):
    """two Collections over the same genome -> one (b's samples after a's)."""
    assert a.genome_length == b.genome_length
    off = np.concatenate([a.off, a.off[-1] + b.off[1:]])
    return Collection(a.genome_length, np.concatenate([a.pos, b.pos]), off, np.concatenate([a.invalid, b.invalid]),
                      np.concatenate([a.family, b.family + (int(a.family.max()) + 1 if a.n else 0)]))


def reference_string(ref_codes):
    return _BASES[ref_codes].tobytes().decode("ascii")


def synthetic_mask(length, n_excluded, n_runs, seed=7):
    """boolean[length], True = excluded; n_runs runs totalling ~n_excluded (TB-exclude-shaped)."""
    rng = np.random.default_rng(seed)
    ex = np.zeros(length, dtype=bool)
    if n_excluded <= 0 or n_runs <= 0:
        return ex
    lens = rng.multinomial(n_excluded, np.ones(n_runs) / n_runs)
    starts = np.sort(rng.integers(0, max(1, length - int(lens.max()) - 1), size=n_runs))
    for s, l in zip(starts, lens):
        ex[s:s + l] = True
    return ex


def _mutate(vpos, vbase, ref, count, rng, rule):
    """apply `count` substitutions at uniformly random positions to the variant map (vpos sorted, vbase codes)."""
    if count <= 0:
        return vpos, vbase
    new_pos = np.unique(rng.integers(0, ref.size, size=count))
    # current base at those positions: variant base if present else reference
    idx = np.searchsorted(vpos, new_pos)
    present = (idx < vpos.size)
    present[present] = vpos[idx[present]] == new_pos[present]
    cur = ref[new_pos].copy()
    cur[present] = vbase[idx[present]]
    if rule == "lss":
        nb = np.where(cur == 2, 3, 2).astype(np.uint8)          # G->T else ->G
    else:
        nb = ((cur + rng.integers(1, 4, size=new_pos.size)) % 4).astype(np.uint8)
    # merge: drop old entries at these positions, add the new ones unless they revert to reference
    keep = np.ones(vpos.size, dtype=bool)
    keep[idx[present]] = False
    add = nb != ref[new_pos]
    pos = np.concatenate([vpos[keep], new_pos[add]])
    base = np.concatenate([vbase[keep], nb[add]])
    order = np.argsort(pos, kind="stable")
    return pos[order], base[order]


def _blocks(length, n_blocks, total, rng):
    """sorted unique positions of `n_blocks` contiguous blocks with `total` positions overall."""
    if total <= 0 or n_blocks <= 0:
        return np.empty(0, dtype=np.int64)
    lens = rng.multinomial(total, np.ones(n_blocks) / n_blocks)
    starts = rng.integers(0, max(1, length - int(lens.max()) - 1), size=n_blocks)
    parts = [np.arange(s, s + l, dtype=np.int64) for s, l in zip(starts, lens) if l > 0]
    return np.unique(np.concatenate(parts)) if parts else np.empty(0, dtype=np.int64)


def make_collection(n_anc, max_c, genome_length=TB_LENGTH, k1=500, s=3.0, n_blocks=20, n_mean=4000, m_mean=0,
                    p_invalid=0.01, seed=1, rule="lss", ref_seed=962, excluded=None, n_lognormal_sigma=None,
                    max_ns=130000, base=None, star=False, anc_seed=None):
    """Generate n_anc*max_c samples -> Collection.  `excluded`: boolean[L] or None.
    `base` = (positions, base codes): variants every ancestor starts from (shared ancestry: lineage-defining SNPs) instead
    of the bare reference.  `star`: every sample descends from the family's ancestor directly (an outbreak: all members
    within a few SNVs of each other) instead of from a random earlier node.  `anc_seed`: the ancestor's own substitutions are
    drawn from this seed (several calls then share ONE ancestor: an outbreak generated in parallel pieces)."""
    rng = np.random.default_rng(seed)
    ref = random_reference(genome_length, ref_seed)
    base_vp = np.empty(0, dtype=np.int64) if base is None else np.asarray(base[0], dtype=np.int64)
    base_vb = np.empty(0, dtype=np.uint8) if base is None else np.asarray(base[1], dtype=np.uint8)
    n = n_anc * max_c
    off = np.zeros(6 * n + 1, dtype=np.int64)
    invalid = np.zeros(n, dtype=np.uint8)
    family = np.repeat(np.arange(n_anc, dtype=np.int32), max_c)
    chunks = []
    total = 0
    i = 0
    for a in range(n_anc):
        nodes = []
        for c in range(max_c):
            if c == 0:
                arng = rng if anc_seed is None else np.random.default_rng(anc_seed)
                vp, vb = _mutate(base_vp, base_vb, ref, k1, arng, rule)
                if anc_seed is not None:                 # a piece of a larger family: its first member is a descendant too
                    nodes.append((vp, vb))
                    vp, vb = _mutate(vp, vb, ref, int(rng.poisson(s)), rng, rule)
            else:
                pp, pb = nodes[0] if star else nodes[int(rng.integers(0, len(nodes)))]
                vp, vb = _mutate(pp, pb, ref, int(rng.poisson(s)), rng, rule)
            nodes.append((vp, vb))
            if rng.random() < p_invalid:
                invalid[i] = 1
                off[6 * i + 1: 6 * i + 7] = total
                i += 1
                continue
            if n_lognormal_sigma is None:
                n_total = int(rng.poisson(n_mean)) if n_mean > 0 else 0
            else:
                n_total = int(rng.lognormal(np.log(max(n_mean, 1)), n_lognormal_sigma))
            npos = _blocks(genome_length, n_blocks, n_total, rng)
            mpos = np.unique(rng.integers(0, genome_length, size=int(rng.poisson(m_mean)))) if m_mean > 0 \
                else np.empty(0, dtype=np.int64)
            if excluded is not None:
                npos = npos[~excluded[npos]]
                mpos = mpos[~excluded[mpos]]
            mpos = np.setdiff1d(mpos, npos, assume_unique=True)
            if npos.size + mpos.size > max_ns:
                invalid[i] = 1
                off[6 * i + 1: 6 * i + 7] = total
                i += 1
                continue
            keep = np.ones(vp.size, dtype=bool)
            if excluded is not None:
                keep &= ~excluded[vp]
            if npos.size:
                j = np.searchsorted(npos, vp)
                hit = j < npos.size
                hit[hit] = npos[j[hit]] == vp[hit]
                keep &= ~hit
            if mpos.size:
                j = np.searchsorted(mpos, vp)
                hit = j < mpos.size
                hit[hit] = mpos[j[hit]] == vp[hit]
                keep &= ~hit
            p, b = vp[keep], vb[keep]
            lists = [p[b == code] for code in range(4)] + [npos, mpos]
            sizes = [x.size for x in lists]
            off[6 * i + 1: 6 * i + 7] = total + np.cumsum(sizes)
            total += int(sum(sizes))
            chunks.append(np.concatenate(lists).astype(np.int32))
            i += 1
    pos = np.concatenate(chunks) if chunks else np.empty(0, dtype=np.int32)
    return Collection(genome_length, pos, off, invalid, family)


def _families_task(task):
    seeds, max_c, tb_mask, kw, bases = task
    kw = dict(kw)
    spec = kw.pop("mask_spec", None)             # (length, excluded positions, runs, seed) of another organism's mask
    mask = synthetic_mask(*spec) if spec else (synthetic_mask(TB_LENGTH, TB_EXCLUDED, TB_EXCLUDED_RUNS) if tb_mask else None)
    parts = []
    for k, sd in enumerate(seeds):
        base = None
        if bases is not None:
            base = lineage_base(kw.get("genome_length", TB_LENGTH), bases[k][0], bases[k][1], kw.get("ref_seed", 962),
                                kw.get("rule", "lss"))
        parts.append(make_collection(1, max_c, seed=int(sd), excluded=mask, base=base, **kw))
    return (np.concatenate([c.pos for c in parts]), np.concatenate([np.diff(c.off) for c in parts]),
            np.concatenate([c.invalid for c in parts]))


def make_families(seeds, max_c, workers=None, tb_mask=True, bases=None, **kw):
    """One TB-shaped family per seed — make_collection(1, max_c, seed=seed, **kw), the collection BASELINE configs[1..4] are
    built from — concatenated in seed order, generated by a pool of processes (the generator is single-threaded numpy:
    200 000 samples take minutes on one core).  `bases[f]` = (lineage, sub-lineage) of family f (see lineage_base) or None.
    -> Collection; family f holds the samples of seeds[f]."""
    import multiprocessing
    import os
    from concurrent.futures import ProcessPoolExecutor
    seeds = [int(x) for x in seeds]
    workers = max(1, min(workers or (os.cpu_count() or 1), len(seeds)))
    per = max(1, min(25, (len(seeds) + workers - 1) // workers))
    tasks = [(seeds[i:i + per], max_c, tb_mask, kw, None if bases is None else bases[i:i + per])
             for i in range(0, len(seeds), per)]
    if workers == 1:
        results = [_families_task(t) for t in tasks]
    else:                                # spawn: the parent may hold a CUDA context, which must not be forked
        try:
            with ProcessPoolExecutor(max_workers=workers, mp_context=multiprocessing.get_context("spawn")) as pool:
                results = list(pool.map(_families_task, tasks))
        except Exception:                # e.g. a __main__ that cannot be re-imported (stdin, REPL): same result, one core
            results = [_families_task(t) for t in tasks]
    pos = np.concatenate([r[0] for r in results]) if results else np.empty(0, dtype=np.int32)
    lens = np.concatenate([r[1] for r in results]) if results else np.empty(0, dtype=np.int64)
    off = np.zeros(lens.size + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    invalid = np.concatenate([r[2] for r in results]) if results else np.empty(0, dtype=np.uint8)
    family = np.repeat(np.arange(len(seeds), dtype=np.int32), max_c)
    return Collection(kw.get("genome_length", TB_LENGTH), pos, off, invalid, family)


# ---- collections that stress the early-exit filter -------------------------------------------------

@lru_cache(maxsize=64)
def lineage_base(genome_length, lineage, sublineage, ref_seed=962, rule="lss", k_lineage=800, k_sub=150):
    """variants shared by every member of (lineage, sub-lineage): reference + k_lineage lineage-defining substitutions +
    k_sub sub-lineage-defining ones — real M. tuberculosis collections look like this relative to H37Rv: two unrelated
    lineage-4 isolates agree on hundreds of non-reference positions."""
    ref = random_reference(genome_length, ref_seed)
    vp, vb = _mutate(np.empty(0, dtype=np.int64), np.empty(0, dtype=np.uint8), ref, k_lineage,
                     np.random.default_rng(1_000_003 * (int(lineage) + 1)), rule)
    vp, vb = _mutate(vp, vb, ref, k_sub, np.random.default_rng(7_000_003 * (int(lineage) + 1) + int(sublineage) + 1), rule)
    vp.setflags(write=False); vb.setflags(write=False)
    return vp, vb


def make_lineages(n_anc, max_c, seed=11, n_lineages=4, n_sub=4, workers=None, k1=500, **kw):
    """Deep shared ancestry: every family's ancestor = its (lineage, sub-lineage) base (800 + 150 shared SNPs) + k1 own
    substitutions, then the usual random tree.  Family f belongs to lineage f % n_lineages, sub-lineage (f // n_lineages) % n_sub.
    -> Collection"""
    bases = [(f % n_lineages, (f // n_lineages) % n_sub) for f in range(n_anc)]
    return make_families([seed * 1_000_000 + f for f in range(n_anc)], max_c, workers=workers, bases=bases, k1=k1, **kw)


def make_outbreak(n, seed=13, piece=500, workers=None, s=2.0, **kw):
    """An outbreak: n samples that ALL descend directly from one ancestor (reference + k1 substitutions) with Poisson(s)
    private substitutions each — every pair is within a few SNVs, so no pair can be ruled out early and every candidate is
    a neighbour at snpCeiling 12..20.  Generated in parallel pieces that share the ancestor."""
    n_pieces = (n + piece - 1) // piece
    coll = make_families([seed * 1_000_000 + 17 * k for k in range(n_pieces)], piece, workers=workers, star=True, s=s,
                         anc_seed=seed * 31 + 5, **kw)
    if coll.n > n:
        coll = Collection(coll.genome_length, coll.pos[: coll.off[6 * n]], coll.off[: 6 * n + 1], coll.invalid[:n], coll.family[:n])
    return coll


def mix_samples(a, b, max_ns=None):
    """consensus call of a MIXTURE of two samples, as src/make_simulation.py:147-164 (mix_sequences) builds it from strings:
    a position keeps its character where the two sequences agree and becomes a mixed base where they differ.
    a, b = (pos, off, invalid) valid samples -> (pos, off, invalid)."""
    (pa, oa, _), (pb, ob, _) = a, b
    lists = []
    for l in range(4):
        lists.append(np.intersect1d(pa[oa[l]:oa[l + 1]], pb[ob[l]:ob[l + 1]], assume_unique=True))
    n_both = np.intersect1d(pa[oa[4]:oa[5]], pb[ob[4]:ob[5]], assume_unique=True)
    m_both = np.intersect1d(pa[oa[5]:oa[6]], pb[ob[5]:ob[6]], assume_unique=True)     # the same IUPAC character in both
    touched = np.union1d(pa[:oa[6]], pb[:ob[6]])
    same = np.concatenate(lists + [n_both])
    mixed = np.union1d(np.setdiff1d(touched, same, assume_unique=False), m_both)
    lists += [n_both, mixed]
    off = np.zeros(7, dtype=np.int32)
    off[1:] = np.cumsum([x.size for x in lists])
    invalid = max_ns is not None and (n_both.size + mixed.size) > max_ns
    return np.concatenate(lists).astype(np.int32), off, bool(invalid)


def add_mixtures(coll, fraction=0.02, seed=5, max_ns=None):
    """append round(fraction * n) mixed samples, each the mixture of two random valid samples"""
    rng = np.random.default_rng(seed)
    valid = np.flatnonzero(coll.invalid == 0)
    k = int(round(fraction * coll.n))
    chunks, offs, inv = [coll.pos], [coll.off], [coll.invalid]
    total = int(coll.off[-1])
    for _ in range(k):
        i, j = rng.choice(valid, size=2, replace=False)
        p, o, bad = mix_samples(coll.lists(int(i)), coll.lists(int(j)), max_ns)
        if bad:
            p, o = np.empty(0, np.int32), np.zeros(7, np.int32)
        chunks.append(p)
        offs.append(total + o[1:].astype(np.int64))
        total += int(o[6])
        inv.append(np.array([1 if bad else 0], np.uint8))
    return Collection(coll.genome_length, np.concatenate(chunks), np.concatenate(offs), np.concatenate(inv),
                      np.concatenate([coll.family, np.full(k, -1, np.int32)]))


def new_sample_like(coll, i, extra_snps, seed, rule="lss", ref_seed=962):
    """a 'newly sequenced' relative of sample i: its lists + `extra_snps` fresh substitutions (same N/M)."""
    rng = np.random.default_rng(seed)
    ref = random_reference(coll.genome_length, ref_seed)
    p, o, inv = coll.lists(i)
    if inv:
        return p, o, inv
    vp = np.concatenate([p[o[b]:o[b + 1]] for b in range(4)]).astype(np.int64)
    vb = np.concatenate([np.full(o[b + 1] - o[b], b, dtype=np.uint8) for b in range(4)])
    order = np.argsort(vp, kind="stable")
    vp, vb = _mutate(vp[order], vb[order], ref, extra_snps, rng, rule)
    unc = np.concatenate([p[o[4]:o[5]], p[o[5]:o[6]]]).astype(np.int64)
    if unc.size:
        keep = ~np.isin(vp, unc)
        vp, vb = vp[keep], vb[keep]
    lists = [vp[vb == code] for code in range(4)] + [p[o[4]:o[5]].astype(np.int64), p[o[5]:o[6]].astype(np.int64)]
    off = np.zeros(7, dtype=np.int32)
    off[1:] = np.cumsum([x.size for x in lists])
    return np.concatenate(lists).astype(np.int32), off, False


def to_string(ref_codes, pos, off, excluded=None, m_char="R"):
    """render one sample as a sequence string (small genomes only; used to exercise compress())."""
    seq = _BASES[ref_codes].copy()
    for b in range(4):
        seq[pos[off[b]:off[b + 1]]] = _BASES[b]
    seq[pos[off[4]:off[5]]] = ord("N")
    seq[pos[off[5]:off[6]]] = ord(m_char)
    return seq.tobytes().decode("ascii")
