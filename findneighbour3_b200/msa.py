"""mixPORE / multi-sequence-alignment statistics of the reference's seqComparer (host side).

Restates src/seqComparer.py:617-1128 (estimate_expected_*, multi_sequence_alignment, _msa) over the
host copies of the sample dicts.  This is SURVEY.md §8(f)-3 ("next" row): the variant-site extraction and
per-sample N/M counts are set operations over the same lists the device store holds; the binomial tests
stay on the host (scipy).  Output keys, column names and value conventions follow the reference exactly,
including its quirk that column 'expected_proportion2' carries expected_p3 (:1071-1072).
"""
import numpy as np
import pandas as pd
from scipy.stats import binomtest


def _binom_greater(x, n, p):
    """scipy.stats.binom_test(x, n, p, alternative='greater') — removed upstream; same exact test."""
    return binomtest(int(x), int(n), p, alternative='greater').pvalue


def estimate_expected_proportion(seqs):
    """median N count of the aligned strings over the alignment length; None for < 3 strings or empty (:617-630)."""
    if len(seqs) < 3:
        return None
    if len(seqs[0]) == 0:
        return None
    return np.median([s.count('N') for s in seqs]) / len(seqs[0])


def _sampled_unknowns(sc, sample_size, exclude_guids, counter):
    """shared loop of estimate_expected_unk / _unk_sites (:632-689): median over the first sample_size valid guids."""
    guids = list(set(sc.seqProfile.keys()) - set(exclude_guids))
    counts = []
    for guid in guids:
        try:
            seq = sc._computeComparator(sc.seqProfile[guid])
            counts.append(counter(seq))
        except ValueError:
            pass
        if len(counts) >= sample_size:
            break
    if len(counts) >= sample_size:
        return np.median(counts)
    return None


def estimate_expected_unk(sc, sample_size=30, exclude_guids=set(), unk_type='N'):
    def counter(seq):
        n = 0
        if unk_type in ('N', 'N_or_M'):
            n += len(seq['N'])
        if unk_type in ('M', 'N_or_M'):
            n += len(seq['M'])
        return n
    return _sampled_unknowns(sc, sample_size, exclude_guids, counter)


def estimate_expected_unk_sites(sc, sample_size=30, sites=set(), exclude_guids=set(), unk_type='N'):
    sites = set(sites)

    def counter(seq):
        n = 0
        if unk_type in ('N', 'N_or_M'):
            n += len(seq['N'].intersection(sites))
        if unk_type in ('M', 'N_or_M'):
            n += len(set(seq['M'].keys()).intersection(sites))
        return n
    return _sampled_unknowns(sc, sample_size, exclude_guids, counter)


def multi_sequence_alignment(sc, guids, output='dict', sample_size=30, expected_p1=None, uncertain_base_type='N'):
    """(:691-802) split guids into valid / invalid, estimate expected_p1 if not given, delegate to msa()."""
    if expected_p1 is not None:
        if expected_p1 < 0 or expected_p1 > 1:
            raise ValueError("Expected_p1 must lie between 0 and 1")
    if sample_size is None:
        sample_size = 30
    valid, invalid = [], []
    for guid in guids:
        try:
            sc._computeComparator(sc.seqProfile[guid])
            valid.append(guid)
        except ValueError:
            invalid.append(guid)
    if expected_p1 is None:
        expected_n1 = estimate_expected_unk(sc, sample_size=sample_size, exclude_guids=invalid)
        expected_p1 = None if expected_n1 is None else expected_n1 / len(sc.reference)
    return msa(sc, valid, invalid, expected_p1, output, sample_size, uncertain_base_type)


def msa(sc, valid_guids, invalid_guids, expected_p1, output, sample_size, uncertain_base_type='N'):
    """variant-site alignment + four one-sided binomial tests per sample (:804-1128)."""
    ref = sc.reference
    seqs = {g: sc._computeComparator(sc.seqProfile[g]) for g in valid_guids}
    all_unk = {'N': {}, 'M': {}, 'N_or_M': {}}
    align_unk = {'N': {}, 'M': {}, 'N_or_M': {}}

    # steps 1-3: positions where the valid samples (reference base where a sample has no call) disagree
    seen = {}
    for g in valid_guids:
        s = seqs[g]
        all_unk['N'][g] = len(s['N'])
        all_unk['M'][g] = len(s['M'])
        all_unk['N_or_M'][g] = all_unk['N'][g] + all_unk['M'][g]
        for base in ('A', 'C', 'T', 'G'):
            for p in s[base]:
                seen.setdefault(p, set()).add(base)
    for g in valid_guids:
        s = seqs[g]
        for p in seen:
            if not (p in s['A'] or p in s['C'] or p in s['T'] or p in s['G']):
                seen[p].add(ref[p])
    ordered = sorted(p for p, bases in seen.items() if len(bases) > 1)

    # step 4: aligned strings (later keys win, as in the reference's A,C,T,G,N,M scan :941-948)
    guid2seq, guid2msa = {}, {}
    for g in valid_guids:
        s = seqs[g]
        row = []
        for p in ordered:
            ch = ref[p]
            for base in ('A', 'C', 'T', 'G', 'N'):
                if p in s[base]:
                    ch = base
            if p in s['M']:
                ch = 'M'
            row.append(ch)
        guid2seq[g] = row
        guid2msa[g] = ''.join(row)

    # step 5: expectation at the aligned sites
    expected_n2 = estimate_expected_unk_sites(sc, sample_size=sample_size, exclude_guids=invalid_guids,
                                              sites=set(ordered), unk_type=uncertain_base_type)
    if expected_n2 is None or len(ordered) == 0:
        expected_p2 = None
    else:
        expected_p2 = expected_n2 / len(ordered)

    if len(valid_guids) == 0:
        return None

    p1, p2, p3, p4 = {}, {}, {}, {}
    observed, e1, e2, e3, e4 = {}, {}, {}, {}, {}
    for g in valid_guids:
        aligned = guid2msa[g]
        for t in ('N', 'M'):
            align_unk[t][g] = aligned.count(t)
        align_unk['N_or_M'][g] = align_unk['N'][g] + align_unk['M'][g]
        x = align_unk[uncertain_base_type][g]
        n = len(aligned)

        if expected_p1 is None or n == 0:
            p1[g], observed[g] = None, None
        else:
            observed[g] = x / n
            p1[g] = _binom_greater(x, n, expected_p1)
        e1[g] = expected_p1

        if expected_p2 is None or n == 0:
            p2[g] = None
        else:
            p2[g] = _binom_greater(x, n, expected_p2)
        e2[g] = expected_p2

        e3[g] = (all_unk[uncertain_base_type][g] - x) / (len(ref) - n)
        p3[g] = _binom_greater(x, n, e3[g])

        others = [guid2msa[o] for o in guid2msa.keys() if not o == g]
        e4[g] = estimate_expected_proportion(others)
        p4[g] = None if e4[g] is None else _binom_greater(x, n, e4[g])

    def col(d, name):
        f = pd.DataFrame.from_dict(d, orient='index')
        f.columns = [name]
        return f

    df = col(guid2msa, 'aligned_seq')
    df['aligned_seq_len'] = len(ordered)
    frames = [col(all_unk['N'], 'allN'), col(align_unk['N'], 'alignN'), col(all_unk['M'], 'allM'),
              col(align_unk['M'], 'alignM'), col(all_unk['N_or_M'], 'allN_or_M'),
              col(align_unk['N_or_M'], 'alignN_or_M'), col(p1, 'p_value1'), col(p2, 'p_value2'),
              col(p3, 'p_value3'), col(p4, 'p_value4'), col(observed, 'observed_proportion'),
              col(e1, 'expected_proportion1'), col(e3, 'expected_proportion2'), col(e3, 'expected_proportion3'),
              col(e4, 'expected_proportion4')]
    frames[-1]['what_tested'] = uncertain_base_type
    for f in frames:
        df = df.merge(f, left_index=True, right_index=True)

    result = {
        'variant_positions': ordered,
        'invalid_guids': invalid_guids,
        'what_tested': [uncertain_base_type] * len(df.index),
        'guid2sequence': guid2seq,
        'guid2allN': all_unk['N'], 'guid2allM': all_unk['M'], 'guid2allN_or_M': all_unk['N_or_M'],
        'guid2msa_seq': guid2msa,
        'guid2observed_proportion': observed,
        'guid2expected_p1': e1, 'guid2expected_p2': e2, 'guid2expected_p3': e3, 'guid2expected_p4': e4,
        'guid2pvalue1': p1, 'guid2pvalue2': p2, 'guid2pvalue3': p3, 'guid2pvalue4': p4,
        'guid2alignN': align_unk['N'], 'guid2alignM': align_unk['M'], 'guid2alignN_or_M': align_unk['N_or_M'],
    }
    if output == 'dict':
        return result
    elif output == 'df':
        return df
    elif output == 'df_dict':
        return df.to_dict(orient='index')
    raise ValueError("Don't know how to format {0}.  Valid options are 'df','df_dict', 'dict'".format(output))
