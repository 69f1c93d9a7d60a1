"""TEST INFRASTRUCTURE — loads the UNMODIFIED reference modules from /root/reference/src.

Only usable in the build container (the GPU box has no /root/reference).  Used by
tests/golden/make_golden.py to generate golden vectors and by the container-only tests
that run the reference's own unittest classes against the drop-in class.

The reference imports Biopython "only for unit testing" (src/seqComparer.py:25-29) and
`scipy.stats.binom_test` (removed upstream); neither is installed here, so empty stand-in
modules and a binomtest adapter are registered before import.  Nothing in the reference
source is altered or copied.
"""
import importlib.util
import os
import sys
import types

REFERENCE_SRC = os.environ.get("FN3_REFERENCE_SRC", "/root/reference/src")


def reference_available():
    return os.path.exists(os.path.join(REFERENCE_SRC, "seqComparer.py"))


def _install_standins():
    if "Bio" not in sys.modules:
        bio = types.ModuleType("Bio")
        seqio = types.ModuleType("Bio.SeqIO")
        seq = types.ModuleType("Bio.Seq")
        seqrec = types.ModuleType("Bio.SeqRecord")
        alpha = types.ModuleType("Bio.Alphabet")
        seq.Seq = str
        seqrec.SeqRecord = object
        alpha.generic_nucleotide = None
        bio.SeqIO, bio.Seq, bio.SeqRecord, bio.Alphabet = seqio, seq, seqrec, alpha
        sys.modules.update({"Bio": bio, "Bio.SeqIO": seqio, "Bio.Seq": seq, "Bio.SeqRecord": seqrec,
                            "Bio.Alphabet": alpha})
    import scipy.stats as st
    if not hasattr(st, "binom_test"):
        def binom_test(x, n, p, alternative="two-sided"):
            return st.binomtest(int(x), int(n), p, alternative=alternative).pvalue
        st.binom_test = binom_test


def load_reference_module(name="seqComparer"):
    """Import /root/reference/src/<name>.py under the module name `_fn3ref_<name>`."""
    if not reference_available():
        raise RuntimeError("reference checkout not present at %s" % REFERENCE_SRC)
    _install_standins()
    modname = "_fn3ref_" + name
    if modname in sys.modules:
        return sys.modules[modname]
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REFERENCE_SRC, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod
