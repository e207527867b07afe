"""ctypes binding of libchiron_b200.so (the C ABI declared in include/chiron_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, an exception is raised."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int8, c_int32, c_longlong, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CHIRON_B200_LIB") or os.path.join(_HERE, "lib", "libchiron_b200.so")   # env override: A/B builds

CB_OK = 0
CB_ERR_ARG = -1
PREC_FP32, PREC_TC_SPLIT = 0, 1
ASM_SIMPLE, ASM_GLUE, ASM_STICK = 0, 1, 2
BN_POPULATION, BN_BATCH = 0, 1
BN_MODES = {"population": BN_POPULATION, "batch": BN_BATCH}
PRECISIONS = {"fp32": PREC_FP32, "tc": PREC_TC_SPLIT, "tc_split": PREC_TC_SPLIT}
ASM_KERNELS = {"simple": ASM_SIMPLE, "glue": ASM_GLUE, "stick": ASM_STICK}

# name -> (restype, argtypes); mirrors include/chiron_b200.h one to one
SIGNATURES = {
    "cb_create": (c_int, [c_void_p, c_size_t, c_int, c_int, POINTER(c_void_p)]),
    "cb_destroy": (c_int, [c_void_p]),
    "cb_last_error": (c_char_p, []),
    "cb_version": (c_char_p, []),
    "cb_set_bn_mode": (c_int, [c_void_p, c_int]),
    "cb_bn_mode": (c_int, [c_void_p]),
    "cb_out_len": (c_int, [c_void_p, c_int]),
    "cb_n_class": (c_int, [c_void_p]),
    "cb_precision": (c_int, [c_void_p]),
    "cb_workspace_bytes": (c_size_t, [c_void_p]),
    "cb_seq_len_out": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "cb_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "cb_check_status": (c_int, [c_void_p, c_void_p]),
    "cb_decode_greedy": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "cb_decode_beam": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "cb_decode_beam_scored": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "cb_assemble": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                            c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "cb_basecall_host": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                 c_void_p]),
    "cb_basecall_submit": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int]),
    "cb_basecall_collect": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "cb_assemble_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    "cb_host_parse_signal": (c_longlong, [c_char_p, c_size_t, c_void_p, c_size_t]),
    "cb_host_normalize": (c_int, [c_void_p, c_size_t, c_int, c_void_p]),
    "cb_host_windows": (c_longlong, [c_void_p, c_size_t, c_int, c_int, c_void_p, c_void_p, c_size_t]),
    "cb_host_format_segments": (c_longlong, [c_char_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_size_t]),
    "cb_launch_count": (c_longlong, [c_void_p]),
    "cb_last_forward_ms": (c_int, [c_void_p, POINTER(c_float), c_int]),
    "cb_enable_timing": (None, [c_void_p, c_int]),
    "cb_reserve_sms": (c_int, [c_void_p, c_int]),
    "cb_last_forward_profile": (c_int, [c_void_p, POINTER(c_float), POINTER(c_int), c_int]),
    "cb_debug_fetch": (c_longlong, [c_void_p, c_int, c_void_p, c_size_t]),
    "cb_host_alloc": (c_void_p, [c_size_t]),
    "cb_host_free": (None, [c_void_p]),
}

# include/chiron_b200_selftest.h (test-only hooks; never used by the product path)
SELFTEST_SIGNATURES = {
    "cb_selftest_beam": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "cb_selftest_beam16": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "cb_selftest_disp": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int]),
}

_LIB = None


class ChironB200Error(RuntimeError):
    pass


def load():
    """Load the library (once).  Raises ChironB200Error when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ChironB200Error(
            "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C chiron_b200/csrc` (there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in list(SIGNATURES.items()) + list(SELFTEST_SIGNATURES.items()):
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch, which must be loud
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != CB_OK:
        msg = load().cb_last_error().decode("utf-8", "replace")
        raise ChironB200Error("%s failed (%d): %s" % (what or "chiron_b200 call", rc, msg))
