"""not-gpu: libfn3.so loads on a CPU-only host, exports every symbol include/fn3.h declares, and refuses to
create an engine without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fn3.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fn3_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for must in ("fn3_create", "fn3_destroy", "fn3_append", "fn3_append_bulk", "fn3_remove", "fn3_one_vs_all",
                 "fn3_all_vs_all", "fn3_pair", "fn3_pair_lists", "fn3_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from findneighbour3_b200 import _native
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), "libfn3.so does not export %s" % name
    assert set(_native.SIGNATURES) == set(declared_symbols())
    assert lib.fn3_abi_version() == 4


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_device():
    from findneighbour3_b200.engine import Engine, EngineError
    with pytest.raises(EngineError) as ei:
        Engine(1000, 0)
    assert "no CUDA device" in str(ei.value) or "CPU" in str(ei.value)
    from findneighbour3_b200.seqComparer import seqComparer
    sc = seqComparer(reference="ACTG", maxNs=1e8, snpCeiling=20)
    with pytest.raises(EngineError):
        sc.persist(sc.compress("AAAA"), "x")


def test_product_does_not_import_oracle():
    """the shipped package must never import, load or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "findneighbour3_b200")
    bad = re.compile(r"^\s*(from\s+\.*oracle|import\s+oracle)|libfn3_oracle|oracle/_build|fn3o_", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not bad.search(text), "%s reaches into oracle/" % f


def test_dropin_directory_shadows_the_reference_module_names():
    """PYTHONPATH=<repo>/findneighbour3_b200/dropin:<repo> makes the server's own import lines
    (findNeighbour3-server.py:72, :74) pick up the B200 classes without touching the server."""
    import subprocess
    import sys
    code = ("from seqComparer import seqComparer as a; from seqComparer_mt import seqComparer as b; "
            "from NucleicAcid import NucleicAcid as c; "
            "print(a.__module__, b.__module__, c.__module__)")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "findneighbour3_b200", "dropin"), ROOT]))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd="/")
    assert out.returncode == 0, out.stderr
    assert out.stdout.split() == ["findneighbour3_b200.seqComparer", "findneighbour3_b200.seqComparer_mt",
                                  "findneighbour3_b200.NucleicAcid"]


def test_scan_split_rule_is_host_logic_with_the_measured_choices():
    """k_scan's split of the 32-row head tiles (fn3_engine.cu pick_scan_split; no device involved): the choices the A/B
    runs of profiles/README.md found best, and the invariants of the rule."""
    from findneighbour3_b200 import _native
    lib = ctypes.CDLL(_native.LIB_PATH)
    f = lib.fn3_debug_scan_split
    f.argtypes = [ctypes.c_int64, ctypes.c_int32, ctypes.c_int64, ctypes.c_int64]
    f.restype = ctypes.c_int
    warps = 148 * 24                                            # one CTA of 24 warps per SM of a B200
    assert f(50040, 13, -1, warps) == 1                         # the bench store: ONE 16-row sub-unit per warp
    assert f(125016, 13, -1, warps) == 3                        # configs[4]'s share of a GPU: eight 4-row sub-units per tile
    assert f(1000000, 13, -1, warps) == 0                       # nine whole tiles per warp: nothing to gain from drawing more often
    for n_rows in (1, 31, 255, 256, 4097, 8192, 8193, 20000, 60000, 300000):
        for w in (24, warps // 128, warps):
            ss = f(n_rows, 13, -1, w)
            assert 0 <= ss <= 3
            # never whole tiles while that leaves warps of the launch without a row
            full, rem = divmod(n_rows, 8192)
            tiles_used = full * 256 + min(256, rem)
            if tiles_used < w and n_rows >= 2 * w:
                assert ss > 0, (n_rows, w)
    assert f(0, 13, 3001, warps) in (0, 1, 2, 3)                # an explicit candidate list
    assert f(0, 13, 32 * warps * 40, warps) == 0                # forty whole tiles per warp
