"""
ctypes binding of libgk_typing.so (the C ABI declared in include/gk_typing.h).

Importing this module never touches the GPU; :func:`load` does, and it raises
if the shared object is missing -- there is no CPU fallback for the product
path.  The numpy dtypes below mirror the C structs field for field and are
checked against ``gk_sizeof`` at load time.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

GK_MAX_CN = 8
GK_KB = 64
GK_RT = 32
GK_LIK_READS = 128

LIB_PATH = os.environ.get("GK_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib",
                                                   "libgk_typing.so")

MATRIX_DTYPE = np.dtype([
    ("mem_off", "<i8"), ("entoff_off", "<i8"), ("L_off", "<i8"), ("LT_off", "<i8"), ("col_off", "<i8"),
    ("n_reads", "<i4"), ("n_alleles", "<i4"), ("n_words", "<i4"), ("r_pad", "<i4"),
    ("a_tile", "<i4"), ("n_ablk", "<i4"), ("n_reads_total", "<i4"), ("m_max", "<i4"),
], align=True)

SEARCH_DTYPE = np.dtype([
    ("P_off", "<i8"), ("S_off", "<i8"), ("cand_off", "<i8"), ("flag_off", "<i8"),
    ("alive_off", "<i8"), ("cnt_off", "<i8"),
    ("matrix", "<i4"), ("n_cand", "<i4"), ("s_stride", "<i4"), ("alive_cap", "<i4"),
    ("n_kblk", "<i4"), ("pad", "<i4"),
], align=True)

LIK_ITEM_DTYPE = np.dtype([("matrix", "<i4"), ("a_blk", "<i4"), ("r0", "<i4"), ("flags", "<i4")])
GK_LIK_COLSUM_ONLY = 1
EXPAND_ITEM_DTYPE = np.dtype([("matrix", "<i4"), ("r0", "<i4"), ("hdr_base", "<i4"), ("keep_off", "<i4"),
                              ("stream_off", "<u4"), ("ent_off", "<u4")])
SCORE_ITEM_DTYPE = np.dtype([("search", "<i4"), ("k_blk", "<i4"), ("a_blk", "<i4"),
                             ("r0", "<i4"), ("r1", "<i4"), ("shape", "<i4")])
COUNT_ITEM_DTYPE = np.dtype([("search", "<i4"), ("f0", "<i4"), ("r0", "<i4"), ("r1", "<i4")])
P_ITEM_DTYPE = np.dtype([("search", "<i4"), ("k_blk", "<i4"), ("r0", "<i4"), ("r1", "<i4")])
STEP_INFO_DTYPE = np.dtype([("n_kept", "<i4"), ("n_unique", "<i4"), ("n_alive", "<i4"), ("cut", "<i4"),
                            ("bar", "<u4"), ("tie_flags", "<i4"), ("best_rank", "<i4"), ("pad1", "<i4")])

EM_PROBLEM_DTYPE = np.dtype([("row_off", "<i8"), ("wgt_off", "<i8"), ("len_off", "<i8"), ("out_off", "<i8"),
                             ("n_rows", "<i4"), ("n_alleles", "<i4"), ("n_awords", "<i4"), ("pad", "<i4")])

_STRUCTS = {
    "GkMatrix": MATRIX_DTYPE, "GkSearch": SEARCH_DTYPE, "GkLikItem": LIK_ITEM_DTYPE,
    "GkExpandItem": EXPAND_ITEM_DTYPE,
    "GkScoreItem": SCORE_ITEM_DTYPE, "GkCountItem": COUNT_ITEM_DTYPE, "GkPItem": P_ITEM_DTYPE,
    "GkStepInfo": STEP_INFO_DTYPE, "GkEmProblem": EM_PROBLEM_DTYPE,
}

# every symbol include/gk_typing.h declares
EXPORTS = (
    "gk_last_error", "gk_abi_version", "gk_sizeof", "gk_wire_encode", "gk_expand_reads", "gk_likelihood", "gk_first_step", "gk_score",
    "gk_select", "gk_rescore_count", "gk_rank", "gk_write_p", "gk_em_compat", "gk_em_squarem",
    "gk_group_reads", "gk_cn_fit", "gk_json_scan", "gk_json_fill", "gk_json_free", "gk_pack_entries", "gk_sam_walk",
    "gk_sam_extract", "gk_sam_extract_fill", "gk_sam_extract_free", "gk_sam_extract_json",
    "gk_plan_score_tiles", "gk_plan_score_items", "gk_plan_grid_items",
)

_lib = None


class GkError(RuntimeError):
    """A launcher of libgk_typing.so returned an error."""


def load(path: str | None = None) -> ctypes.CDLL:
    """dlopen the CUDA library; raise (never fall back) when it is absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise GkError(
            f"{path} not found: build it with `python -m kir_graph_b200.build` "
            "(kir_graph_b200 has no CPU fallback)")
    lib = ctypes.CDLL(path)
    for name in EXPORTS:
        if not hasattr(lib, name):
            raise GkError(f"{path} does not export {name}")
    lib.gk_last_error.restype = ctypes.c_char_p
    lib.gk_sizeof.argtypes = [ctypes.c_char_p]
    _vp, _i64, _i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32
    lib.gk_plan_score_tiles.restype = _i64
    lib.gk_plan_score_tiles.argtypes = [ctypes.c_int, _vp, _vp, _vp, _vp]
    lib.gk_plan_score_items.restype = _i64
    lib.gk_plan_score_items.argtypes = [ctypes.c_int, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64]
    lib.gk_plan_grid_items.restype = _i64
    lib.gk_plan_grid_items.argtypes = [ctypes.c_int, _vp, _vp, _i32, _vp, _i64, _vp, _i64]
    for name, dtype in _STRUCTS.items():
        size = lib.gk_sizeof(name.encode())
        if size != dtype.itemsize:
            raise GkError(f"struct {name}: C sizeof {size} != numpy itemsize {dtype.itemsize}")
    _lib = lib
    return lib


def call(fn_name: str, *args) -> None:
    """Invoke a launcher with raw pointers / ints; translate its status into an exception."""
    lib = load()
    fn = getattr(lib, fn_name)
    conv = []
    for a in args:
        if isinstance(a, float):
            conv.append(ctypes.c_double(a))
        elif isinstance(a, (int, np.integer)):
            conv.append(ctypes.c_longlong(int(a)) if abs(int(a)) >= 2 ** 31 else ctypes.c_int(int(a)))
        elif a is None:
            conv.append(ctypes.c_void_p(0))
        else:
            conv.append(a)
    status = fn(*conv)
    if status != 0:
        raise GkError(f"{fn_name} failed ({status}): {lib.gk_last_error().decode()}")


def ptr(value: int) -> ctypes.c_void_p:
    return ctypes.c_void_p(int(value))
