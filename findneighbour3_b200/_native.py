"""ctypes binding of libfn3.so (the C-ABI declared in include/fn3.h).

There is deliberately no fallback: if the shared object is missing or a symbol is
absent, importing/using the engine raises.  The product path never touches oracle/.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FN3_LIB_PATH") or os.path.join(HERE, "libfn3.so")   # override: kernel experiments only

FN3_OK = 0
FN3_EARG = -1
FN3_ENOTFOUND = -2
FN3_ECAPACITY = -3
FN3_ECUDA = -4
FN3_ENOMEM = -5


class Hit(C.Structure):
    _fields_ = [("row", C.c_int64), ("dist", C.c_int32), ("n1", C.c_int32),
                ("n2", C.c_int32), ("nboth", C.c_int32)]


class Edge(C.Structure):
    _fields_ = [("row1", C.c_int64), ("row2", C.c_int64), ("dist", C.c_int32),
                ("n1", C.c_int32), ("n2", C.c_int32), ("nboth", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("n_rows", C.c_int64), ("n_live", C.c_int64), ("n_invalid", C.c_int64),
                ("store_words", C.c_int64), ("store_bytes", C.c_int64), ("positions", C.c_int64),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("t_pack_ns", C.c_int64),
                ("t_enqueue_ns", C.c_int64), ("t_wait_ns", C.c_int64), ("n_timed", C.c_int64),
                ("head_refreshes", C.c_int64), ("scan_mode", C.c_int64)]


# array arguments (int32_t*, int64_t*, uint8_t*, float*) are declared void*: the callers pass plain integer addresses of
# numpy buffers (engine._ptr) — building a typed ctypes pointer per argument costs ~3.6 us, a third of an insert+query call
_vp = C.c_void_p
_i32p = _i64p = _u8p = _f32p = _vp

# name -> (restype, argtypes); mirrors include/fn3.h one to one
SIGNATURES = {
    "fn3_abi_version": (C.c_int, []),
    "fn3_create": (_vp, [C.c_int64, C.c_int32]),
    "fn3_destroy": (C.c_int, [_vp]),
    "fn3_last_error": (C.c_char_p, [_vp]),
    "fn3_append": (C.c_int, [_vp, _i32p, _i32p, C.c_int32, _i64p]),
    "fn3_append_bulk": (C.c_int, [_vp, C.c_int64, _i32p, _i64p, _u8p, _i64p]),
    "fn3_remove": (C.c_int, [_vp, C.c_int64]),
    "fn3_stats_get": (C.c_int, [_vp, _vp]),
    "fn3_refresh_heads": (C.c_int, [_vp, C.c_int64]),
    "fn3_compact": (C.c_int, [_vp, _i64p, C.c_int64, _i64p, _i64p]),
    "fn3_row_get": (C.c_int, [_vp, C.c_int64, _i32p, C.c_int64, _i32p, _i32p, _i64p]),
    "fn3_one_vs_all": (C.c_int, [_vp, C.c_int64, C.c_int32, _i64p, C.c_int64, _vp, C.c_int64, _i64p]),
    "fn3_insert_query": (C.c_int, [_vp, _i32p, _i32p, C.c_int32, C.c_int32, _i64p, _vp, C.c_int64, _i64p]),
    "fn3_lists_vs_all": (C.c_int, [_vp, _i32p, _i32p, C.c_int32, C.c_int32, _i64p, C.c_int64,
                                   _vp, C.c_int64, _i64p]),
    "fn3_all_vs_all": (C.c_int, [_vp, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                 _vp, C.c_int64, _i64p]),
    "fn3_all_vs_all_fetch": (C.c_int, [_vp, _vp, C.c_int64, _i64p]),
    "fn3_pair": (C.c_int, [_vp, C.c_int64, C.c_int64, C.c_int32, _vp, _i32p]),
    "fn3_pair_lists": (C.c_int, [_vp, _i32p, _i32p, C.c_int32, _i32p, _i32p, C.c_int32, C.c_int32,
                                 _vp, _i32p]),
    "fn3_set_reference": (C.c_int, [_vp, C.c_char_p, _i32p, C.c_int64]),
    "fn3_compress": (C.c_int, [_vp, C.c_char_p, _i32p, C.c_int64, _i32p, C.c_char_p, C.c_int64, _i64p]),
    "fn3_ingest_query": (C.c_int, [_vp, C.c_char_p, C.c_int64, C.c_int32, _i64p, _i32p, C.c_int64, _i32p, C.c_char_p, C.c_int64,
                                   _i64p, _i32p, _i64p, _vp, C.c_int64, _i64p]),
    "fn3_byte_histogram": (C.c_int, [C.c_int32, _u8p, C.c_int64, _i64p]),
    "fn3_consensus": (C.c_int, [_vp, _i64p, C.c_int64, C.c_int64, _i32p, C.c_int64, _i32p, _i64p]),
    "fn3_consensus_lists": (C.c_int, [_vp, C.c_int64, _i32p, _i64p, C.c_int64, _i32p, C.c_int64, _i32p, _i64p]),
    "fn3_msa_sites": (C.c_int, [_vp, _i64p, C.c_int64, _i32p, C.c_int64, _i64p]),
    "fn3_msa_sites_rule": (C.c_int, [_vp, _i64p, C.c_int64, C.c_int32, _i32p, C.c_int64, _i64p]),
    "fn3_msa_align": (C.c_int, [_vp, _i64p, C.c_int64, _i32p, C.c_int64, _u8p, _i32p]),
    "fn3_unpickle_samples": (C.c_int, [C.c_int64, _u8p, _i64p, _i32p, C.c_int64, _i64p, _u8p, _i64p]),
    "fn3_profile_queries": (C.c_int, [_vp, _i64p, C.c_int32, C.c_int32, C.c_int32, _f32p, _f32p, _i64p, _i64p, _i64p]),
    "fn3_launch_count": (C.c_int64, [_vp]),
    # engine groups (one process, several GPUs): same argument lists as the single-engine calls of the same name
    "fn3_group_create": (_vp, [C.c_int64, C.c_int32, _i32p]),
    "fn3_group_destroy": (C.c_int, [_vp]),
    "fn3_group_last_error": (C.c_char_p, [_vp]),
    "fn3_group_n_devices": (C.c_int32, [_vp]),
    "fn3_group_peer_access": (C.c_int32, [_vp]),
    "fn3_group_engine": (_vp, [_vp, C.c_int32]),
    "fn3_group_launch_count": (C.c_int64, [_vp]),
    "fn3_group_append": (C.c_int, [_vp, _i32p, _i32p, C.c_int32, _i64p]),
    "fn3_group_append_bulk": (C.c_int, [_vp, C.c_int64, _i32p, _i64p, _u8p, _i64p]),
    "fn3_group_remove": (C.c_int, [_vp, C.c_int64]),
    "fn3_group_row_get": (C.c_int, [_vp, C.c_int64, _i32p, C.c_int64, _i32p, _i32p, _i64p]),
    "fn3_group_stats_get": (C.c_int, [_vp, _vp]),
    "fn3_group_set_reference": (C.c_int, [_vp, C.c_char_p, _i32p, C.c_int64]),
    "fn3_group_insert_query": (C.c_int, [_vp, _i32p, _i32p, C.c_int32, C.c_int32, _i64p, _vp, C.c_int64, _i64p]),
    "fn3_group_ingest_query": (C.c_int, [_vp, C.c_char_p, C.c_int64, C.c_int32, _i64p, _i32p, C.c_int64, _i32p, C.c_char_p,
                                         C.c_int64, _i64p, _i32p, _i64p, _vp, C.c_int64, _i64p]),
    "fn3_group_one_vs_all": (C.c_int, [_vp, C.c_int64, C.c_int32, _i64p, C.c_int64, _vp, C.c_int64, _i64p]),
    "fn3_group_lists_vs_all": (C.c_int, [_vp, _i32p, _i32p, C.c_int32, C.c_int32, _i64p, C.c_int64, _vp, C.c_int64, _i64p]),
    "fn3_group_pair": (C.c_int, [_vp, C.c_int64, C.c_int64, C.c_int32, _vp, _i32p]),
    "fn3_group_all_vs_all": (C.c_int, [_vp, C.c_int32, C.c_int32, _vp, C.c_int64, _i64p]),
    "fn3_group_consensus": (C.c_int, [_vp, _i64p, C.c_int64, C.c_int64, _i32p, C.c_int64, _i32p, _i64p]),
    "fn3_group_consensus_lists": (C.c_int, [_vp, C.c_int64, _i32p, _i64p, C.c_int64, _i32p, C.c_int64, _i32p, _i64p]),
    "fn3_group_msa_sites_rule": (C.c_int, [_vp, _i64p, C.c_int64, C.c_int32, _i32p, C.c_int64, _i64p]),
    "fn3_group_msa_align": (C.c_int, [_vp, _i64p, C.c_int64, _i32p, C.c_int64, _u8p, _i32p]),
}

_lib = None


def load():
    """Load libfn3.so and bind every symbol of include/fn3.h.  Raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libfn3.so is not built (%s). Run `python -m findneighbour3_b200.build` "
            "(or __graft_entry__.build()). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
