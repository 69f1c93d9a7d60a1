"""Drop-in replacement for findNeighbour3's older, threaded `seqComparer_mt.seqComparer` with the distance path on a B200.

    from findneighbour3_b200.seqComparer_mt import seqComparer     # instead of: from seqComparer_mt import seqComparer

`src/seqComparer_mt.py` is the reference's multi-threaded generation of the class: the same store
and the same distance, but an earlier API — samples carry no mixed-base dict (keys invalid,A,C,G,T,N, :260), `compress`
raises KeyError on any character outside ACGTN- (:477-478, quirk Q5), distances mask by N alone (:611-614; identical to the
current formula when M is empty, which it always is here), patches are {'add','subtract'} (:262, :670-713), the statistics
keys lack the 'server|' prefix (:389-408), the MSA lets an N account for a position (:1127-1137) and reports three tests
instead of four (:1044-1274), and `mcompare` (:337-387)
  * with cpuCount == 1 skips the query itself, uses snpCompressionCeiling as the cut-off and returns lists;
  * with cpuCount > 1 splits the candidates over `ComparingThread`s (:41-209), which use snpCeiling, INCLUDE the
    query itself (quirk Q6; expected in test mc1, :2268-2274) and return tuples.
Here both forms are ONE kernel launch over the HBM store (there is nothing left to thread); the two return conventions
are kept.  Everything else is inherited from the drop-in `seqComparer` class; no distance is computed on the CPU.
"""
import multiprocessing

import numpy as np
import pandas as pd

from . import msa as _msa_mod
from .seqComparer import seqComparer as _seqComparer

_BASES5 = ("A", "C", "T", "G", "N")


class seqComparer(_seqComparer):
    def __init__(self, reference, maxNs, snpCeiling, debugMode=False, excludePositions=set(),
                 snpCompressionCeiling=250, cpuCount=None, device=None):
        """Same arguments as src/seqComparer_mt.py:212-220; `device` (CUDA ordinal) is new.  cpuCount only selects
        which of the reference's two mcompare conventions is returned (:294-300 clamp it to 1..cpu_count)."""
        super().__init__(reference, maxNs, snpCeiling, debugMode, excludePositions, snpCompressionCeiling, 1, device)
        self.compressed_sequence_keys = set(['invalid', 'A', 'C', 'G', 'T', 'N'])
        self.patch_keys = set(['add', 'subtract'])
        n_cpus = multiprocessing.cpu_count()
        self.cpuCount = cpuCount
        if self.cpuCount is None:
            self.cpuCount = n_cpus
        if self.cpuCount > n_cpus:
            self.cpuCount = n_cpus
        if self.cpuCount < 1:
            self.cpuCount = 1

    # ------------------------------------------------------------------------------------------
    # sample type without M
    # ------------------------------------------------------------------------------------------
    def compress(self, sequence):
        """string -> {'A','C','G','T','N': set, 'invalid': 0} | {'invalid': 1} (:458-491); device kernel fn3_compress.
        A character outside ACGTN- raises KeyError, as `diffDict[sequence[i]]` does in the reference (:478)."""
        if not len(sequence) == len(self.reference):
            raise TypeError("sequence must of the same length as reference; seq is {0} and ref is {1}".format(
                len(sequence), len(self.reference)))
        if len(self.reference) == 0:
            raise TypeError("reference cannot be of zero length")
        eng = self._engine_with_reference()
        pos, off, _ = eng.compress(sequence.encode('latin-1', 'replace'))
        if off[6] > off[5]:
            raise KeyError(sequence[int(pos[off[5]])])
        return self._dict_from_lists(pos, off, [], int(off[5] - off[4]) > self.maxNs)

    def _dict_from_lists(self, pos, off, m_chars, invalid):
        if invalid:
            return {'invalid': 1}
        out = {}
        for i, b in enumerate('ACGT'):
            out[b] = set(pos[off[i]:off[i + 1]].tolist())
        out['N'] = set(pos[off[4]:off[5]].tolist())
        out['invalid'] = 0
        return out

    def insert_sequence(self, guid, sequence, cutoff=None):
        """as in the current class; a character outside ACGTN- raises KeyError like compress() and stores nothing"""
        found, hist = super().insert_sequence(guid, sequence, self.snpCeiling if cutoff is None else cutoff)
        v = self.seqProfile._d[guid]
        if v.off[6] > v.off[5]:
            self.remove(guid)
            raise KeyError(sequence[int(v.pos[v.off[5]])])
        return found, hist

    def uncompress(self, compressed_sequence):
        """(:440-456)"""
        if 'invalid' in compressed_sequence.keys():
            if compressed_sequence['invalid'] == 1:
                raise ValueError("Cannot uncompress an invalid sequence, because the sequence it is not stored {0}".format(
                    compressed_sequence.keys()))
        cs = self._computeComparator(compressed_sequence)
        seq = list(self.reference)
        for x in self.excluded:
            seq[x] = 'N'
        for item in _BASES5:
            for x in cs[item]:
                seq[x] = item
        return ''.join(seq)

    def _uncertain_sets(self, s1, s2):
        """N1pos, N2pos, Nbothpos = _setStats(seq1['N'], seq2['N']) (:538-551, :587)"""
        return s1['N'], s2['N'], s1['N'] | s2['N']

    # ------------------------------------------------------------------------------------------
    # the hot path
    # ------------------------------------------------------------------------------------------
    def mcompare(self, guid, guids=None):
        """one guid against all (or `guids`) (:337-387) — one device one-vs-all in either convention."""
        if guids is None:
            guids = set(self.seqProfile.keys())
        if guid not in self.seqProfile:
            raise KeyError("Asked to compare {0}  but guid requested has not been stored.  call .persist() on the sample "
                           "to be added before using mcompare.".format(guid))
        guids = list(set(guids))
        threaded = self.cpuCount != 1
        # single thread: countDifferences_byKey(cutoff=self.snpCompressionCeiling) (:359-360; None falls back to
        # snpCeiling, :567-568); threads: ComparingThread uses snpCeiling (:66, :375)
        cutoff = self.snpCeiling if threaded or self.snpCompressionCeiling is None else self.snpCompressionCeiling
        found = self.neighbours(guid, guids, cutoff)
        s1 = None
        out = []
        for key2 in guids:
            if key2 == guid:
                if not threaded:
                    continue                                   # the single-thread loop skips the query (:358)
                hit = None                                     # the threads compare it with itself: distance 0 (:64-67)
                if self._stored_kind(guid) != 'invalid' and 0 <= cutoff:
                    n_self = len(self._computeComparator(self.seqProfile[guid])['N'])
                    hit = (0, n_self, n_self, n_self)
            else:
                hit = found.get(key2)
            if hit is None:
                row = [guid, key2, None, None, None, None, None, None, None]
            else:
                if s1 is None:
                    s1 = self._computeComparator(self.seqProfile[guid])
                s2 = self._computeComparator(self.seqProfile[key2])
                N1pos, N2pos, Nbothpos = self._uncertain_sets(s1, s2)
                row = [guid, key2, hit[0], hit[1], hit[2], hit[3], N1pos, N2pos, Nbothpos]
            out.append(tuple(row) if threaded else row)
        return out

    def summarise_stored_items(self):
        """(:389-408): the counts of the current class under the older key names"""
        return {k.replace('server|', ''): v for k, v in super().summarise_stored_items().items()}

    # ------------------------------------------------------------------------------------------
    # double delta with {'add','subtract'} patches
    # ------------------------------------------------------------------------------------------
    def generate_patch(self, compressed_sequence, consensus):
        """(:670-691); empty sets are not stored"""
        add, subtract = {}, {}
        for item in _BASES5:
            a = compressed_sequence[item] - consensus[item]
            s = consensus[item] - compressed_sequence[item]
            if len(a) > 0:
                add[item] = a
            if len(s) > 0:
                subtract[item] = s
        return {'add': add, 'subtract': subtract}

    def apply_patch(self, patch, consensus):
        """(:692-713)"""
        if not patch.keys() == self.patch_keys:
            raise TypeError("Patch passed has wrong keys {0}".format(patch.keys))
        if not consensus.keys() == self.consensus_keys:
            raise TypeError("Consensus passed has wrong keys {0}".format(consensus.keys))
        out = {'invalid': 0}
        for item in _BASES5:
            out[item] = (consensus[item] | patch['add'].get(item, set())) - patch['subtract'].get(item, set())
        return out

    # ------------------------------------------------------------------------------------------
    # mixture statistics of the older class: device sites / alignment / tallies, three binomial tests on the host
    # ------------------------------------------------------------------------------------------
    def estimate_expected_N(self, sample_size=30, exclude_guids=set()):
        """(:779-800)"""
        return _msa_mod.estimate_expected_unk(self, sample_size, exclude_guids, 'N')

    def estimate_expected_N_sites(self, sample_size=30, sites=set(), exclude_guids=set()):
        """(:801-823)"""
        return _msa_mod.estimate_expected_unk_sites(self, sample_size, sites, exclude_guids, 'N')

    def multi_sequence_alignment(self, guids, output='dict', sample_size=30, expected_p1=None):
        """(:939-1042)"""
        if expected_p1 is not None:
            if expected_p1 < 0 or expected_p1 > 1:
                raise ValueError("Expected_p1 must lie between 0 and 1")
        if sample_size is None:
            sample_size = 30
        valid, invalid = [], []
        for guid in guids:
            if guid not in self.seqProfile:
                raise KeyError(guid)
            (invalid if self._stored_kind(guid) == 'invalid' else valid).append(guid)
        if expected_p1 is None:
            expected_n1 = self.estimate_expected_N(sample_size=sample_size, exclude_guids=invalid)
            expected_p1 = None if expected_n1 is None else expected_n1 / len(self.reference)
        return self._msa(valid, invalid, expected_p1, output, sample_size)

    def _msa(self, valid_guids, invalid_guids, expected_p1, output, sample_size):
        """(:1044-1274).  Column quirks kept: 'p_value3' carries test 2's p value (:1232-1233), 'expected_proportion2'
        carries expected_p3 (:1238-1239)."""
        if len(valid_guids) == 0:
            return None
        ref = self.reference
        with self._lock:
            rows = [self._row_of[g] for g in valid_guids]
        eng = self._engine_with_reference()
        sites = eng.msa_sites(rows, n_is_call=True)
        codes, counts = eng.msa_align(rows, sites, want_codes=True)
        ordered = sites.tolist()
        chars = _msa_mod._CODE_CHARS[codes]
        if len(ordered):
            ref_at = np.frombuffer(''.join(ref[p] for p in ordered).encode('ascii'), dtype=np.uint8)
            chars = np.where(codes == 0, ref_at[None, :], chars)
        guid2seq, guid2msa, all_n, align_n = {}, {}, {}, {}
        for i, g in enumerate(valid_guids):
            text = chars[i].tobytes().decode('ascii')
            guid2seq[g] = list(text)
            guid2msa[g] = text
            all_n[g] = int(counts[i, 0])
            align_n[g] = int(counts[i, 2])

        expected_n2 = self.estimate_expected_N_sites(sample_size=sample_size, exclude_guids=invalid_guids, sites=set(ordered))
        expected_p2 = None if expected_n2 is None or len(ordered) == 0 else expected_n2 / len(ordered)

        p1, p2, p3, observed, e1, e2, e3 = {}, {}, {}, {}, {}, {}, {}
        for g in valid_guids:
            x, n = align_n[g], len(guid2msa[g])
            if expected_p1 is None or n == 0:
                p1[g], observed[g] = None, None
            else:
                observed[g] = x / n
                p1[g] = _msa_mod._binom_greater(x, n, expected_p1)
            e1[g] = expected_p1
            p2[g] = None if expected_p2 is None or n == 0 else _msa_mod._binom_greater(x, n, expected_p2)
            e2[g] = expected_p2
            e3[g] = (all_n[g] - x) / (len(ref) - n)
            p3[g] = _msa_mod._binom_greater(x, n, e3[g])

        def col(d, name):
            f = pd.DataFrame.from_dict(d, orient='index')
            f.columns = [name]
            return f

        df = col(guid2msa, 'aligned_seq')
        for f in (col(all_n, 'allN'), col(align_n, 'alignN'), col(p1, 'p_value1'), col(p2, 'p_value2'), col(p2, 'p_value3'),
                  col(observed, 'observed_proportion'), col(e1, 'expected_proportion1'), col(e3, 'expected_proportion2'),
                  col(e3, 'expected_proportion3')):
            df = df.merge(f, left_index=True, right_index=True)
        result = {'variant_positions': ordered, 'invalid_guids': invalid_guids, 'guid2sequence': guid2seq,
                  'guid2allN': all_n, 'guid2msa_seq': guid2msa, 'guid2observed_proportion': observed,
                  'guid2expected_p1': e1, 'guid2expected_p2': e2, 'guid2expected_p3': e3,
                  'guid2pvalue1': p1, 'guid2pvalue2': p2, 'guid2pvalue3': p3, 'guid2alignN': align_n}
        if output == 'dict':
            return result
        elif output == 'df':
            return df
        elif output == 'df_dict':
            return df.to_dict(orient='index')
        raise ValueError("Don't know how to format {0}.  Valid options are 'df','dict'".format(output))

    def assess_mixed(self, this_guid, related_guids, max_sample_size=30):
        """mixture estimate for one sample from every pair of (sampled) related samples (:825-938): one three-sample
        alignment per pair, the row of this_guid kept.  The reference grows the frame with DataFrame.append (:934),
        which pandas 2 removed; pd.concat(ignore_index=True) yields the frame that call produced."""
        sample_size = 30
        if self._stored_kind(this_guid) == 'invalid':
            raise ValueError("{0} is invalid".format(this_guid))
        expected_n1 = self.estimate_expected_N(sample_size=sample_size, exclude_guids=related_guids)
        expected_p1 = None if expected_n1 is None else expected_n1 / len(self.reference)
        valid_related = [g for g in related_guids if self._stored_kind(g) != 'invalid']
        if len(valid_related) > max_sample_size:
            np.random.shuffle(valid_related)
            valid_related = valid_related[0:max_sample_size]
        if len(valid_related) < 2:
            return None
        frames = []
        for i in range(len(valid_related)):
            for j in range(i):
                df = self._msa(valid_guids=[this_guid, valid_related[i], valid_related[j]], invalid_guids=[],
                               expected_p1=expected_p1, output='df', sample_size=30)
                df = df[df.index == this_guid].copy()
                df['pairid'] = len(frames) + 1
                frames.append(df)
        return frames[0] if len(frames) == 1 else pd.concat(frames, ignore_index=True)
