"""mixPORE / multi-sequence-alignment statistics of the reference's seqComparer.

Restates src/seqComparer.py:617-1128 (estimate_expected_*, multi_sequence_alignment, _msa).  This is SURVEY.md
§8(f)-3: the variant-site extraction, the aligned characters and every N/M tally are computed by the device
engine on its own rows (fn3_msa_sites / fn3_msa_align, csrc/fn3_cluster.cuh) — this module holds no set algebra
over sample lists; what stays here is the reference's control flow, its float arithmetic (medians, proportions)
and the four binomial tests (scipy).  Output keys, column names and value conventions follow the reference
exactly, including its quirk that column 'expected_proportion2' carries expected_p3 (:1071-1072).
"""
import numpy as np
import pandas as pd
from scipy.stats import binomtest

_CODE_CHARS = np.frombuffer(b"?ACGTNM", dtype=np.uint8)      # fn3_msa_align codes 1..6; 0 = reference base


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


def _sampled_unknowns(sc, sample_size, exclude_guids, sites, unk_type):
    """shared loop of estimate_expected_unk / _unk_sites (:632-689): median, over the first sample_size valid
    guids, of the N and/or M count (sites=None: whole sample, :648-652; else at the given sites, :678-681)."""
    guids = list(set(sc.seqProfile.keys()) - set(exclude_guids))
    with sc._lock:
        rows = []
        for guid in guids:
            if sc._stored_kind(guid) != 'invalid':            # (no unpickling of lazily loaded samples)
                rows.append(sc._row_of[guid])
            if len(rows) >= sample_size:
                break
    if len(rows) < sample_size or len(rows) == 0:
        return None
    at = np.empty(0, dtype=np.int32) if sites is None else np.asarray(sorted(sites), dtype=np.int32)
    _, counts = sc.engine.msa_align(rows, at, want_codes=False)
    col_n, col_m = (0, 1) if sites is None else (2, 3)
    per_sample = np.zeros(len(rows), dtype=np.int64)
    if unk_type in ('N', 'N_or_M'):
        per_sample += counts[:, col_n]
    if unk_type in ('M', 'N_or_M'):
        per_sample += counts[:, col_m]
    return np.median(per_sample.tolist())


def estimate_expected_unk(sc, sample_size=30, exclude_guids=set(), unk_type='N'):
    return _sampled_unknowns(sc, sample_size, exclude_guids, None, unk_type)


def estimate_expected_unk_sites(sc, sample_size=30, sites=set(), exclude_guids=set(), unk_type='N'):
    return _sampled_unknowns(sc, sample_size, exclude_guids, set(sites), unk_type)


def multi_sequence_alignment(sc, guids, output='dict', sample_size=30, expected_p1=None, uncertain_base_type='N'):
    """(:691-802) split guids into valid / invalid, estimate expected_p1 if not given, delegate to msa()."""
    if expected_p1 is not None:
        if expected_p1 < 0 or expected_p1 > 1:
            raise ValueError("Expected_p1 must lie between 0 and 1")
    if sample_size is None:
        sample_size = 30
    valid, invalid = [], []
    for guid in guids:
        if guid not in sc.seqProfile:
            raise KeyError(guid)
        if sc._stored_kind(guid) != 'invalid':
            valid.append(guid)
        else:
            invalid.append(guid)
    if expected_p1 is None:
        expected_n1 = estimate_expected_unk(sc, sample_size=sample_size, exclude_guids=invalid)
        expected_p1 = None if expected_n1 is None else expected_n1 / len(sc.reference)
    return msa(sc, valid, invalid, expected_p1, output, sample_size, uncertain_base_type)


def msa(sc, valid_guids, invalid_guids, expected_p1, output, sample_size, uncertain_base_type='N'):
    """variant-site alignment + four one-sided binomial tests per sample (:804-1128)."""
    ref = sc.reference
    if len(valid_guids) == 0:
        return None
    all_unk = {'N': {}, 'M': {}, 'N_or_M': {}}
    align_unk = {'N': {}, 'M': {}, 'N_or_M': {}}

    # steps 1-4 on the device: variant sites of these rows, then every row's character and N/M tallies there
    with sc._lock:
        rows = [sc._row_of[g] for g in valid_guids]
    eng = sc._engine_with_reference()
    sites = eng.msa_sites(rows)
    codes, counts = eng.msa_align(rows, sites, want_codes=True)
    ordered = sites.tolist()
    chars = _CODE_CHARS[codes]
    if len(ordered):
        ref_at = np.frombuffer(''.join(ref[p] for p in ordered).encode('ascii'), dtype=np.uint8)
        chars = np.where(codes == 0, ref_at[None, :], chars)
    guid2seq, guid2msa = {}, {}
    for i, g in enumerate(valid_guids):
        text = chars[i].tobytes().decode('ascii')
        guid2seq[g] = list(text)
        guid2msa[g] = text
        all_unk['N'][g] = int(counts[i, 0])
        all_unk['M'][g] = int(counts[i, 1])
        all_unk['N_or_M'][g] = all_unk['N'][g] + all_unk['M'][g]
        align_unk['N'][g] = int(counts[i, 2])
        align_unk['M'][g] = int(counts[i, 3])
        align_unk['N_or_M'][g] = align_unk['N'][g] + align_unk['M'][g]

    # step 5: expectation at the aligned sites
    expected_n2 = estimate_expected_unk_sites(sc, sample_size=sample_size, exclude_guids=invalid_guids,
                                              sites=set(ordered), unk_type=uncertain_base_type)
    if expected_n2 is None or len(ordered) == 0:
        expected_p2 = None
    else:
        expected_p2 = expected_n2 / len(ordered)

    p1, p2, p3, p4 = {}, {}, {}, {}
    observed, e1, e2, e3, e4 = {}, {}, {}, {}, {}
    for g in valid_guids:
        aligned = guid2msa[g]
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
