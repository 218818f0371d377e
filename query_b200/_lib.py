"""ctypes binding of libn1gpu.so (include/n1gpu.h).  There is no fallback: a missing library is an error."""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libn1gpu.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "n1gpu.h")

OK, E_INVALID, E_PARSE, E_INELIGIBLE, E_CUDA, E_IO, E_CANCELLED, E_NOMEM = 0, -1, -2, -3, -4, -5, -6, -7
C_MISSING, C_NULL, C_FALSE, C_TRUE, C_INT, C_FLOAT, C_STRING, C_OTHER = range(8)


class N1GpuError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("n1gpu error %d: %s" % (code, message))
        self.code = code
        self.message = message


class Ineligible(N1GpuError):
    """The plan is outside the substituted subset: the caller keeps its own operators."""


_lib = None

_P = C.c_void_p
_I64P = C.POINTER(C.c_int64)
_U8P = C.POINTER(C.c_uint8)
_SIGS = {
    "n1gpu_init": (C.c_int, [C.c_int]),
    "n1gpu_shutdown": (C.c_int, []),
    "n1gpu_last_error": (C.c_char_p, []),
    "n1gpu_version": (C.c_char_p, []),
    "n1gpu_launch_count": (C.c_uint64, []),
    "n1gpu_table_create": (C.c_int, [C.POINTER(_P)]),
    "n1gpu_table_add_column": (C.c_int, [_P, C.c_char_p]),
    "n1gpu_table_find_column": (C.c_int, [_P, C.c_char_p]),
    "n1gpu_table_append_json": (C.c_int, [_P, C.c_char_p, _I64P, C.c_int64, C.c_int]),
    "n1gpu_table_load_dir": (C.c_int, [_P, C.c_char_p, C.c_int]),
    "n1gpu_table_set_column": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, C.c_int64, C.c_char_p, _I64P, C.c_int64]),
    "n1gpu_table_set_column_device": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, C.c_int64, C.c_char_p, _I64P, C.c_int64]),
    "n1gpu_table_seal": (C.c_int, [_P]),
    "n1gpu_table_num_rows": (C.c_int64, [_P]),
    "n1gpu_table_num_columns": (C.c_int, [_P]),
    "n1gpu_table_column_scan_bytes": (C.c_int, [_P, C.c_int]),
    "n1gpu_table_dict_export": (C.c_int, [_P, C.c_int, C.c_char_p, C.c_int64, _I64P, C.c_int64, _I64P, _I64P]),
    "n1gpu_table_dict_import": (C.c_int, [_P, C.c_int, C.c_char_p, _I64P, C.c_int64]),
    "n1gpu_table_dict_merge": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_char_p), C.POINTER(_I64P), _I64P]),
    "n1gpu_table_stats_get": (C.c_int, [_P, C.c_int, _I64P]),
    "n1gpu_table_stats_set": (C.c_int, [_P, C.c_int, _I64P]),
    "n1gpu_table_set_global_rows": (C.c_int, [_P, C.c_int64]),
    "n1gpu_table_load_ndjson": (C.c_int, [_P, C.c_char_p, C.c_int]),
    "n1gpu_table_set_segment_output": (C.c_int, [_P, C.c_char_p, C.c_char_p]),
    "n1gpu_table_load_segment": (C.c_int, [_P, C.c_char_p, C.c_char_p, C.POINTER(C.c_int)]),
    "n1gpu_set_segment_dir": (C.c_int, [C.c_char_p]),
    "n1gpu_table_column_peek": (C.c_int, [_P, C.c_int, _I64P, _U8P, C.c_int64]),
    "n1gpu_table_free": (C.c_int, [_P]),
    "n1gpu_query_compile": (C.c_int, [_P, C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_int,
                                      C.POINTER(C.c_char_p), C.c_int, C.POINTER(_P)]),
    "n1gpu_query_compile_params": (C.c_int, [_P, C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.POINTER(C.c_char_p), C.c_int,
                                             C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_int, C.POINTER(_P)]),
    "n1gpu_jit_stats": (C.c_int, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "n1gpu_query_execute": (C.c_int, [_P, C.POINTER(_P)]),
    "n1gpu_query_launch": (C.c_int, [_P]),
    "n1gpu_query_collect": (C.c_int, [_P, C.POINTER(_P)]),
    "n1gpu_query_cancel": (C.c_int, [_P]),
    "n1gpu_query_kernel_source": (C.c_char_p, [_P]),
    "n1gpu_query_part_source": (C.c_char_p, [_P]),
    "n1gpu_query_info": (C.c_int, [_P, _I64P]),
    "n1gpu_query_last_scan_ns": (C.c_int64, [_P]),
    "n1gpu_query_rebind": (C.c_int, [_P, _P]),
    "n1gpu_query_set_stream": (C.c_int, [_P, _P]),
    "n1gpu_query_set_timing": (C.c_int, [_P, C.c_int]),
    "n1gpu_query_free": (C.c_int, [_P]),
    "n1gpu_query_scan_partial": (C.c_int, [_P]),
    "n1gpu_query_partial_counts": (C.c_int, [_P, _I64P, _I64P, C.POINTER(C.c_int)]),
    "n1gpu_query_partial_export": (C.c_int, [_P, C.c_int, _P, C.c_int64, _I64P, _P, C.c_int64, _I64P]),
    "n1gpu_query_partial_reset": (C.c_int, [_P]),
    "n1gpu_query_partial_import": (C.c_int, [_P, _P, C.c_int64, _P, C.c_int64]),
    "n1gpu_query_finalize": (C.c_int, [_P, C.POINTER(_P)]),
    "n1gpu_query_state_words": (C.c_int, [_P, C.POINTER(_P), _I64P]),
    "n1gpu_query_word_ops": (C.c_int, [_P, C.POINTER(C.c_int), C.c_int]),
    "n1gpu_query_merge_words": (C.c_int, [_P, _P, C.c_int]),
    "n1gpu_mailbox_create": (C.c_int, [C.c_int, C.c_int, C.c_int64, C.POINTER(_P)]),
    "n1gpu_mailbox_create_arena": (C.c_int, [C.c_int, C.c_int, C.c_int64, C.c_int64, C.POINTER(_P)]),
    "n1gpu_mailbox_set_peer": (C.c_int, [_P, C.c_int, _P]),
    "n1gpu_mailbox_base": (_P, [_P]),
    "n1gpu_mailbox_ipc_handle": (C.c_int, [_P, C.c_char_p]),
    "n1gpu_mailbox_open_peers": (C.c_int, [_P, C.c_char_p]),
    "n1gpu_mailbox_free": (C.c_int, [_P]),
    "n1gpu_query_set_mailbox": (C.c_int, [_P, _P]),
    "n1gpu_query_peer_mode": (C.c_int, [_P]),
    "n1gpu_result_num_groups": (C.c_int64, [_P]),
    "n1gpu_result_num_keys": (C.c_int, [_P]),
    "n1gpu_result_num_aggregates": (C.c_int, [_P]),
    "n1gpu_result_fetch": (C.c_int, [_P, _U8P, _I64P, _U8P, _I64P]),
    "n1gpu_result_string": (C.c_int, [_P, C.c_int64, C.POINTER(C.c_char_p), _I64P]),
    "n1gpu_result_stats": (C.c_int, [_P, _I64P]),
    "n1gpu_result_free": (C.c_int, [_P]),
    "n1gpu_plan_build": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(_P), C.POINTER(C.c_int)]),
    "n1gpu_operator_run_once": (C.c_int, [_P, C.POINTER(_P)]),
    "n1gpu_operator_send_stop": (C.c_int, [_P]),
    "n1gpu_operator_marshal_json": (C.c_int, [_P, C.c_char_p, C.c_int64, _I64P]),
    "n1gpu_result_to_json": (C.c_int, [_P, C.c_char_p, C.c_int64, _I64P]),
    "n1gpu_operator_free": (C.c_int, [_P]),
    "n1gpu_plan_build_tail": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "n1gpu_operator_tail_operators": (C.c_int, [_P, C.c_char_p, C.c_int64, C.POINTER(C.c_int64)]),
    "n1gpu_operator_import_result": (C.c_int, [_P, C.c_int64, _U8P, _I64P, _U8P, _I64P, C.c_char_p, _I64P, C.c_int64, C.POINTER(C.c_void_p)]),
    "n1gpu_operator_num_keys": (C.c_int, [_P]),
    "n1gpu_operator_num_aggregates": (C.c_int, [_P]),
    "n1gpu_operator_run_tail": (C.c_int, [_P, _P, C.c_char_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
}


def declared_symbols():
    """Every function name include/n1gpu.h declares."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(n1gpu_[a-z0-9_]+)\s*\(", text)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise N1GpuError(E_CUDA, "libn1gpu.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                                     "query_b200 has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc >= 0:
        return rc
    msg = lib().n1gpu_last_error().decode("utf-8", "replace")
    if rc == E_INELIGIBLE:
        raise Ineligible(rc, msg)
    raise N1GpuError(rc, msg)
