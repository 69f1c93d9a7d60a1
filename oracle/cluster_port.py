"""TEST INFRASTRUCTURE — CPU restatement (pure Python sets) of the reference's consensus vote and of the
set algebra inside its multi-sequence alignment.

Restates, against the reference's line numbers (src/seqComparer.py):
  consensus                    :460-501
  _msa steps 1-4               :895-950   (variant positions, aligned characters)
  N/M tallies                  :905-908 (whole sample), :975-977 (in the alignment)
  estimate_expected_unk[_sites] :632-689

Pinned against the unmodified reference by tests/golden/cluster.json (tests/golden/make_golden.py --cluster).
Only tests/ may import this module; the product (findneighbour3_b200/) never does.

Samples are plain compressed_sequence dicts ({'A','C','G','T','N': set, 'M': dict, 'invalid': 0}).
"""
import statistics


def consensus(samples, cutoff_proportion):
    """positions carried by at least len(samples)*cutoff_proportion of the samples, per key (:480-500).
    Samples without the key (patch form, invalid) do not vote but do count in the threshold (:485, :493)."""
    cutoff_number = len(samples) * cutoff_proportion
    out = {}
    for item in ("A", "C", "T", "G", "N"):
        votes = {}
        for s in samples:
            if item in s:
                for p in s[item]:
                    votes[p] = votes.get(p, 0) + 1
        out[item] = set(p for p, v in votes.items() if v >= cutoff_number)
    return out


def variant_positions(reference, samples, n_is_call=False):
    """_msa steps 1-3 (:910-931): positions where the samples show more than one base, a sample with no
    A/C/T/G call at the position standing for the reference base.  n_is_call: the older rule of
    src/seqComparer_mt.py:1127-1137, where a sample's N at the position also accounts for it."""
    seen = {}
    for s in samples:
        for base in ("A", "C", "T", "G"):
            for p in s[base]:
                seen.setdefault(p, set()).add(base)
    for s in samples:
        for p in seen:
            if not (p in s["A"] or p in s["C"] or p in s["T"] or p in s["G"] or (n_is_call and p in s["N"])):
                seen[p].add(reference[p])
    return sorted(p for p, bases in seen.items() if len(bases) > 1)


def aligned_string(reference, sample, ordered_positions):
    """_msa step 4 (:936-949): the later key of the scan order A,C,T,G,N,M wins."""
    row = []
    for p in ordered_positions:
        ch = reference[p]
        for base in ("A", "C", "T", "G", "N"):
            if p in sample[base]:
                ch = base
        if p in sample["M"]:
            ch = "M"
        row.append(ch)
    return "".join(row)


def tallies(sample, aligned):
    """(allN, allM, alignN, alignM): :905-908 and :975-977"""
    return (len(sample["N"]), len(sample["M"]), aligned.count("N"), aligned.count("M"))


def unknowns_at_sites(sample, sites, unk_type):
    """one sample's term of estimate_expected_unk_sites (:676-681); sites=None: estimate_expected_unk (:648-652)"""
    n = 0
    if unk_type in ("N", "N_or_M"):
        n += len(sample["N"]) if sites is None else len(sample["N"].intersection(sites))
    if unk_type in ("M", "N_or_M"):
        n += len(sample["M"]) if sites is None else len(set(sample["M"].keys()).intersection(sites))
    return n


def expected_unknowns(samples, sites, unk_type):
    """median over the given samples (the reference takes the first sample_size valid ones of an unordered set;
    callers pass exactly the samples they want in)"""
    return float(statistics.median(unknowns_at_sites(s, sites, unk_type) for s in samples))
