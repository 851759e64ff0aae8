"""ctypes binding of libvqb200.so (the C ABI in include/vqb200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, a
``RuntimeError`` is raised.  Nothing here imports ``oracle/``.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VQB200_LIB", os.path.join(_HERE, "lib", "libvqb200.so"))   # override only for A/B experiments

ALGO_AUTO, ALGO_SIMT, ALGO_TC = 0, 1, 2
ALGO_PREPARED = 256
ALGOS = {"auto": ALGO_AUTO, "simt": ALGO_SIMT, "tc": ALGO_TC}

# scalar slots / result slots (mirror of the enums in include/vqb200.h)
S_SUM_MIN_D, S_COMMIT_SQ, S_MASK_SUM, S_UNSAFE_ROWS, S_COUNT_TOTAL = 0, 1, 2, 3, 4
NUM_SCALARS = 16
R_FIT, R_COMMIT, R_ENTROPY, R_USAGE, R_DK, R_USED_CURR = 0, 1, 2, 3, 4, 5
NUM_RESULTS = 8

_c_void_p, _c_i64, _c_int, _c_size_t, _c_double = (ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                                                   ctypes.c_size_t, ctypes.c_double)

# name -> (restype, argtypes); every symbol include/vqb200.h declares
SIGNATURES = {
    "vq_version": (_c_int, []),
    "vq_last_error": (ctypes.c_char_p, []),
    "vq_device_supported": (_c_int, []),
    "vq_workspace_bytes": (_c_size_t, [_c_i64, _c_i64, _c_int, _c_int]),
    "vq_assign": (_c_int, [_c_void_p, _c_i64, _c_i64, _c_i64, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p,
                           _c_void_p, _c_size_t, _c_int, _c_void_p]),
    "vq_assign_bf16": (_c_int, [_c_void_p, _c_i64, _c_i64, _c_i64, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p,
                                _c_void_p, _c_size_t, _c_int, _c_void_p]),
    "vq_assign_grouped": (_c_int, [_c_void_p, _c_i64, _c_i64, _c_i64, _c_void_p, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p,
                                   _c_void_p, _c_void_p, _c_void_p, _c_size_t, _c_void_p]),
    "vq_assign_debug": (_c_int, [_c_void_p, _c_i64, _c_i64, _c_i64, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p,
                                 _c_void_p, _c_size_t, _c_void_p, _c_void_p, _c_int]),
    "vq_gather_st_fwd": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_i64, _c_i64, _c_i64, _c_int,
                                  _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
    "vq_gather_st_fwd_ema_supported": (_c_int, [_c_i64, _c_i64, _c_int]),
    "vq_gather_st_fwd_ema": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_i64, _c_i64, _c_i64, _c_int,
                                      _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
    "vq_gather_st_bwd": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                  _c_i64, _c_i64, _c_i64, _c_int, _c_void_p, _c_void_p]),
    "vq_decode": (_c_int, [_c_void_p, _c_void_p, _c_i64, _c_i64, _c_i64, _c_int, _c_void_p, _c_void_p]),
    "vq_ema_accumulate": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_i64, _c_i64, _c_i64, _c_int, _c_void_p, _c_void_p, _c_void_p]),
    "vq_ema_finalize": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int,
                                 _c_double, _c_double, _c_double, _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
    "vq_p2p_region_bytes": (_c_size_t, [_c_int, _c_int]),
    "vq_p2p_alloc": (_c_int, [_c_size_t, _c_void_p, _c_void_p]),
    "vq_p2p_open": (_c_int, [_c_void_p, _c_void_p]),
    "vq_p2p_close": (_c_int, [_c_void_p]),
    "vq_p2p_free": (_c_int, [_c_void_p]),
    "vq_p2p_stats_slot": (_c_void_p, [_c_void_p, ctypes.c_uint, _c_int, _c_int]),
    "vq_p2p_krand_slot": (_c_void_p, [_c_void_p, ctypes.c_uint, _c_int, _c_int]),
    "vq_p2p_exchange": (_c_int, [_c_void_p, _c_int, _c_int, ctypes.c_uint, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p]),
    "vq_p2p_publish": (_c_int, [_c_void_p, _c_int, _c_int, ctypes.c_uint, _c_void_p]),
    "vq_p2p_collect": (_c_int, [_c_void_p, _c_int, _c_int, ctypes.c_uint, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p]),
    "vq_gather_rows": (_c_int, [_c_void_p, _c_void_p, _c_i64, _c_i64, _c_i64, _c_i64, _c_void_p, _c_void_p]),
    "vq_restart_rows_device": (_c_int, [_c_void_p, _c_void_p, _c_i64, _c_i64, _c_i64, _c_int, ctypes.c_uint64, _c_void_p, _c_void_p,
                                        _c_void_p, _c_void_p]),
    "vq_profile_enable": (_c_int, [_c_int]),
    "vq_profile_read": (_c_int, [_c_void_p]),
    "vq_host_ctx_create": (_c_void_p, [_c_int, _c_i64, _c_int, _c_int]),
    "vq_host_ctx_destroy": (None, [_c_void_p]),
    "vq_host_ctx_x_staging": (_c_void_p, [_c_void_p]),
    "vq_host_ctx_idx_staging": (_c_void_p, [_c_void_p]),
    "vq_host_ctx_set_codebook": (_c_int, [_c_void_p, _c_void_p]),
    "vq_host_ctx_codes_staging": (_c_void_p, [_c_void_p]),
    "vq_encode_host_u16": (_c_int, [_c_void_p, _c_void_p, _c_i64, _c_i64, _c_void_p, _c_void_p, _c_void_p]),
    "vq_encode_host": (_c_int, [_c_void_p, _c_void_p, _c_i64, _c_i64, _c_void_p, _c_void_p]),
}

_lock = threading.Lock()
_lib = None


def load():
    """Load libvqb200.so (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"vqb200: {LIB_PATH} is missing.  Build it with `python __graft_entry__.py` "
                "(or speech-masters-thesis_b200/build.py); there is no CPU or PyTorch fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here means header and library disagree
            fn.restype = res
            fn.argtypes = args
        if lib.vq_version() != 100:
            raise RuntimeError("vqb200: library/header version mismatch")
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().vq_last_error()
        raise RuntimeError(f"vqb200: {what} failed: {msg.decode() if msg else 'unknown error'}")


def ptr(t):
    """Device (or host) address of a tensor, None -> NULL."""
    return None if t is None else t.data_ptr()
