"""ctypes binding of include/concepthash_b200.h.  Fails loudly: there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# (CONCEPTHASH_B200_LIB: another build of the same library -- kernel experiments under dev/; never a fallback)
LIB_PATH = os.environ.get("CONCEPTHASH_B200_LIB") or os.path.join(HERE, "libconcepthash_b200.so")

CH_F32, CH_F16, CH_BF16, CH_F64, CH_I64, CH_I32, CH_U8, CH_I16, CH_I8 = range(9)
CH_MEM_DEVICE, CH_MEM_HOST = 0, 1
CH_LAB_NONE, CH_LAB_ID, CH_LAB_MASK = 0, 1, 2
CH_EMIT_NONE, CH_EMIT_RELEVANT, CH_EMIT_CANDIDATES = 0, 1, 2
CH_QUERY_NOLABEL, CH_GALLERY_NOLABEL = 0xFFFFFFFF, 0xFFFFFFFE
CH_MAX_NBIT, CH_MAX_R, CH_MAX_PR = 256, 8, 32

P = C.c_void_p


class HistArgs(C.Structure):
    _fields_ = [(n, P) for n in ("q_bits", "q_nz", "g_bits", "g_nz", "q_lab", "g_lab", "slab_all", "slab_rel",
                                 "thresh", "rec_off", "rec_cap", "rec_cnt", "recs", "err_flag")] + \
               [(n, C.c_int64) for n in ("nq", "nq_pad", "ndb")] + \
               [(n, C.c_int32) for n in ("nbit", "ternary", "label_mode", "mask_words", "emit_mode",
                                         "nstripes", "threads", "rows_per_stripe")] + \
               [("row_base", C.c_int64), ("key_limit", C.c_int32)]


class FinalArgs(C.Structure):
    _fields_ = [(n, P) for n in ("recs", "rec_off", "rec_cnt", "base0_all", "base0_rel", "sbase_all", "sbase_rel",
                                 "first_rel", "partial", "cols")] + \
               [(n, C.c_int64) for n in ("nq", "nq_pad")] + \
               [(n, C.c_int32) for n in ("nstripes", "nbins", "remove_first", "nR", "nPR")] + \
               [("r_eff", C.c_int64 * CH_MAX_R), ("pr_k", C.c_int64 * CH_MAX_PR)]


class SelectArgs(C.Structure):
    _fields_ = [(n, P) for n in ("q_i8", "g_i8", "cand_off", "cand_cap", "cand_cnt", "cand_rows", "err_flag", "thresh")] + \
               [(n, C.c_int64) for n in ("nq", "nq_pad", "ndb", "row_base")] + \
               [(n, C.c_int32) for n in ("nbit", "nstripes", "rows_per_stripe", "dense", "ternary")] + \
               [("pair", C.c_int32), ("bad", P), ("q_stripe_bytes", C.c_int64)]


class CandArgs(C.Structure):
    _fields_ = [(n, P) for n in ("cand_off", "cand_cnt", "cand_rows", "cand_key", "q_bits", "g_bits", "q_nz", "g_nz",
                                 "g_plane", "q_lab", "g_lab",
                                 "tot_all", "tot_rel", "base0_all", "base0_rel", "first_rel", "first_rel_out", "key_max", "cols",
                                 "ids", "keys", "err_flag")] + \
               [(n, C.c_int64) for n in ("nq", "nq_pad", "R", "row_offset")] + \
               [(n, C.c_int32) for n in ("nstripes", "nbins", "nbit", "label_mode", "mask_words", "remove_first",
                                         "nR", "nPR", "mode")] + \
               [("r_eff", C.c_int64 * CH_MAX_R), ("pr_k", C.c_int64 * CH_MAX_PR)]


CH_LOADER_MAX_JOBS = 6
CH_LOADER_PACK, CH_LOADER_COPY = 0, 1


class LoaderJob(C.Structure):
    _fields_ = [("codes_host", P), ("n", C.c_int64), ("nbit", C.c_int32), ("kind", C.c_int32),
                ("row_stride", C.c_int64), ("out_bits_dev", P), ("flags_dev", P)]


# name -> (restype, argtypes); the authoritative list of exported symbols (tests check it against the header)
SIGNATURES = {
    "ch_abi_version": (C.c_int, []),
    "ch_last_error": (C.c_char_p, []),
    "ch_workspace_create": (C.c_int, [C.c_int, C.POINTER(P)]),
    "ch_workspace_destroy": (C.c_int, [P]),
    "ch_device_info": (C.c_int, [P] + [C.POINTER(C.c_int)] * 4),
    "ch_padded_rows": (C.c_int64, [C.c_int64]),
    "ch_code_words": (C.c_int, [C.c_int]),
    "ch_pack_sign": (C.c_int, [P, P, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int64, C.c_int64, C.c_double,
                               P, P, P, P, P]),
    "ch_host_pack_sign": (C.c_int, [P, C.c_int64, C.c_int, C.c_int64, P, P, C.c_int]),
    "ch_host_pack_threads": (C.c_int, [P]),
    "ch_host_loader_start": (C.c_int, [P, C.POINTER(LoaderJob), C.c_int, P, C.POINTER(C.c_void_p)]),
    "ch_host_loader_wait": (C.c_int, [P, C.c_int, C.c_int64, P, C.c_int]),
    "ch_host_loader_join": (C.c_int, [P, C.POINTER(C.c_uint32)]),
    "ch_column_sums": (C.c_int, [P, P, C.c_int, C.c_int64, C.c_int, C.c_int64, C.c_int64, P, P]),
    "ch_pack_labels": (C.c_int, [P, P, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int64, C.c_int64, C.c_uint32,
                                 P, P, P, P]),
    "ch_hist_geometry": (C.c_int, [P, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                                   C.POINTER(C.c_int32)]),
    "ch_hamming_hist": (C.c_int, [P, C.POINTER(HistArgs), P]),
    "ch_tc_code_bytes": (C.c_int, [C.c_int]),
    "ch_tc_queries_per_cta": (C.c_int, []),
    "ch_tc_code_bytes_bare": (C.c_int, [C.c_int]),
    "ch_tc_code_bytes_pair": (C.c_int, [C.c_int, C.c_int]),
    "ch_tc_tile_rows": (C.c_int, [C.c_int]),
    "ch_expand_i8": (C.c_int, [P, P, P, C.c_int64, C.c_int, C.c_int, C.c_int, P, C.c_int64, P, C.c_int64, P]),
    "ch_hamming_select_tc": (C.c_int, [P, C.POINTER(SelectArgs), P]),
    "ch_expand_i8_query_stripes": (C.c_int, [P, P, P, C.c_int64, C.c_int, C.c_int, P, C.c_int64, P, P, C.c_int, C.c_int,
                                             C.c_int64, P]),
    "ch_cand_hist": (C.c_int, [P, C.POINTER(CandArgs), P]),
    "ch_gather_plane_words": (C.c_int, [C.c_int]),
    "ch_gather_plane": (C.c_int, [P, P, P, C.c_int64, C.c_int, P, P]),
    "ch_cand_finalize": (C.c_int, [P, C.POINTER(CandArgs), P]),
    "ch_cand_rank": (C.c_int, [P, C.POINTER(CandArgs), C.c_int64, C.c_int64, P, P, P]),
    "ch_cand_caps": (C.c_int, [P, P, P, P, P, P, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int, P, C.c_int, P,
                               P]),
    "ch_slab_totals": (C.c_int, [P, P, C.c_int, C.c_int, C.c_int64, P, P]),
    "ch_slab_exscan": (C.c_int, [P, P, C.c_int, C.c_int, C.c_int64, P]),
    "ch_slab_scan": (C.c_int, [P, P, C.c_int, C.c_int, C.c_int, C.c_int64, P, P]),
    "ch_scan_bases": (C.c_int, [P, P, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, P, P, P, P]),
    "ch_record_caps": (C.c_int, [P, C.c_int, P, P, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int,
                                 P, P]),
    "ch_record_offsets_async": (C.c_int, [P, P, C.c_int, C.c_int64, C.c_int64, P, P, C.c_uint64, C.c_uint32, P, P, P]),
    "ch_scan_bases_pair": (C.c_int, [P, P, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                     P, P, P, P, P, P, P]),
    "ch_gather_rows": (C.c_int, [P, P, C.c_int64, C.c_int, C.c_int64, P, C.c_int64, P]),
    "ch_record_offsets": (C.c_int, [P, P, C.c_int, C.c_int64, C.c_int64, P, C.POINTER(C.c_uint64), P,
                                    C.POINTER(C.c_uint32), P]),
    "ch_check_counts": (C.c_int, [P, P, C.c_int64, C.c_int64, P, P]),
    "ch_class_counts": (C.c_int, [P, P, C.c_int64, C.c_int, C.c_int, P, P]),
    "ch_finalize_records": (C.c_int, [P, C.POINTER(FinalArgs), P]),
    "ch_first_relevant": (C.c_int, [P, C.POINTER(FinalArgs), P, P]),
    "ch_reduce_means": (C.c_int, [P, P, P, P, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64), P,
                                  C.POINTER(C.c_double), P, C.POINTER(C.c_uint32), C.c_int, P]),
    "ch_reduce_means_enqueue": (C.c_int, [P, P, P, P, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64), P, P, C.c_int, P]),
    "ch_reduce_means_fetch": (C.c_int, [P, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_uint32), C.c_int, P]),
    "ch_scatter_ranked": (C.c_int, [P, C.POINTER(FinalArgs), C.c_int64, C.c_int64, P, P, P]),
    "ch_ap_from_ranked": (C.c_int, [P, P, C.c_int64, C.c_int64, P, P, C.c_int, C.c_int, C.c_int,
                                    C.POINTER(C.c_int64), P, P]),
    "ch_hamming_matrix": (C.c_int, [P, P, P, P, P, C.c_int64, C.c_int64, C.c_int, C.c_int, P, P]),
    "ch_popc_peak": (C.c_int, [P, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "ch_launch_count": (C.c_int64, [P]),
}

_lib = None


class NativeLibraryError(RuntimeError):
    pass


def load():
    """Loads libconcepthash_b200.so (built in-tree by ``python -m concepthash_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} is missing: build it with `python -m concepthash_b200.build` "
            "(needs nvcc; sm_100a only).  There is no CPU fallback.")
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as e:  # e.g. libcudart not found
        raise NativeLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise NativeLibraryError(f"{LIB_PATH} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    if lib.ch_abi_version() != 5:
        raise NativeLibraryError("ABI version mismatch: rebuild the library")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().ch_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"concepthash_b200 {what} failed: {msg}")
