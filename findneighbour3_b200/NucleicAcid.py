"""Drop-in replacement for findNeighbour3's `NucleicAcid` class (src/NucleicAcid.py:10-66) with the character counts
on the device.

    from findneighbour3_b200.NucleicAcid import NucleicAcid        # instead of: from NucleicAcid import NucleicAcid

`findNeighbour3.insert` examines every sequence before compressing it (findNeighbour3-server.py:511-513); the
reference counts the 18 permitted characters with one `bytes.count()` pass each over the 4.4 MB string
(NucleicAcid.py:50-58).  Here the counts come from ONE pass on the GPU (fn3_byte_histogram, or — when the caller
inserts through `seqComparer.insert_sequence` — from the histogram the fused ingest call computed on its own copy of the
sequence, passed in as `histogram`).  The upper-casing and the MD5 of the cleaned string (:35, :38) stay on the host:
a hash is a sequential chain, not a data-parallel pass.  Same attributes, same composition keys in the same order,
same exceptions.
"""
import datetime
import hashlib

_IUPAC = ['R', 'r', 'Y', 'y', 'S', 's', 'W', 'w', 'K', 'k', 'M', 'm']
_VALID = ['A', 'C', 'G', 'T', 'N', '-'] + _IUPAC


class NucleicAcid():
    def __init__(self, device=0):
        self._device = device
        self._reset()

    def _reset(self):
        self.isValid = None
        self.composition = {'examinationDate': datetime.datetime.now(), 'MD5': None, 'length': 0}
        self.nchars = 0
        self.nucleicAcidString = None

    def examine(self, nucleicAcidString, histogram=None):
        """evaluates the string passed (:29-66).  `histogram` (256 byte counts of the upper-cased string, as returned by
        seqComparer.insert_sequence / Engine.ingest_query) spares the device pass."""
        self._reset()
        if not isinstance(nucleicAcidString, str):
            self.isValid = False
            raise TypeError("nucleicAcidString must be of type str, but it is %s " % type(nucleicAcidString))
        self.nucleicAcidString = bytes(nucleicAcidString.upper().encode(encoding='utf-8'))
        self.composition['MD5'] = hashlib.md5(self.nucleicAcidString).hexdigest()
        self.nchars = len(self.nucleicAcidString)
        if self.nchars == 0:
            self.isValid = False
            raise ValueError("nucleicAcidString must be of type str, and of non-zero length")
        if histogram is None:
            from .engine import byte_histogram
            histogram = byte_histogram(self.nucleicAcidString, self._device)
        elif int(sum(histogram)) != self.nchars:
            raise ValueError("histogram does not belong to this string")
        # the 18 permitted characters in the reference's order (upper-case codes first within each pair, :46-48); after the
        # upper-casing above the lower-case keys are always 0, as in the reference
        counts = {ch: int(histogram[ord(ch)]) for ch in _VALID}
        comp = self.composition
        comp.update(counts)
        comp['invalid'] = self.nchars - sum(counts.values())
        comp['length'] = self.nchars
        comp['ACTG'] = counts['A'] + counts['C'] + counts['G'] + counts['T']
        comp['mixed'] = sum(counts[ch] for ch in _IUPAC)
        comp['propACTG'] = float(comp['ACTG']) / float(comp['length'])
        if self.composition['invalid'] == 0:
            self.isValid = True
        else:
            self.isValid = False
            raise ValueError("nucleicAcidString must be of type str, and consist only of ACGTN-.  There are %i invalid "
                             "characters" % self.composition['invalid'])
