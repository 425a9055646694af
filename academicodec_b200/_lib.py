"""ctypes binding of libacq_b200.so (the C ABI declared in include/acq_b200.h).

There is no CPU fallback and no other backend: if the shared library is missing the import of
any op fails loudly with instructions to build it.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# ACQ_B200_LIB points at another build of the same library (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("ACQ_B200_LIB") or os.path.join(_HERE, "lib", "libacq_b200.so")

# symbols include/acq_b200.h declares (checked by tests/test_cpu_host.py::test_cabi_exports_every_declared_symbol)
SYMBOLS = (
    "acq_version", "acq_last_error", "acq_codebook_half_norms", "acq_rvq_search", "acq_vq_decode",
    "acq_ema_stats", "acq_ema_apply", "acq_pipeline_create", "acq_pipeline_destroy",
    "acq_rvq_encode_host", "acq_vq_decode_host", "acq_pipeline_last_launches",
    "acq_tc_pack_bytes", "acq_tc_workspace_bytes", "acq_tc_pack_codebooks", "acq_debug_tc_scores",
    "acq_rvq_codec_host", "acq_rvq_replay", "acq_packed_bytes", "acq_pack_codes", "acq_unpack_codes",
    "acq_tc_configure", "acq_tc_query", "acq_pipeline_wait_stream", "acq_peer_allreduce", "acq_grvq_backward",
)

ACQ_STE = 1
ACQ_LOSS_RAW = 2
ACQ_IMPL_AUTO, ACQ_IMPL_SIMT, ACQ_IMPL_TC = 0, 1, 2
ACQ_MAX_TABLE = 64

_lib = None


class AcqError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing. Build it with `python -m academicodec_b200.build` "
            "(nvcc, sm_100a). academicodec_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    pp = POINTER(c_void_p)
    lib.acq_version.restype = c_int
    lib.acq_last_error.restype = c_char_p
    lib.acq_codebook_half_norms.argtypes = [pp, c_int, c_int, c_int, c_void_p, c_void_p]
    lib.acq_rvq_search.argtypes = [c_void_p, pp, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                   c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p]
    lib.acq_peer_allreduce.argtypes = [c_void_p, pp, c_int, c_int, c_size_t, c_void_p]
    lib.acq_grvq_backward.argtypes = [c_void_p, c_void_p, pp, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                      c_void_p, c_double, c_double, c_void_p, pp, c_void_p]
    lib.acq_tc_configure.argtypes = [c_int, c_int, c_int]
    lib.acq_tc_query.argtypes = [c_int]
    lib.acq_tc_pack_bytes.argtypes = [c_int, c_int, c_int]
    lib.acq_tc_pack_bytes.restype = c_size_t
    lib.acq_tc_workspace_bytes.argtypes = [c_int]
    lib.acq_tc_workspace_bytes.restype = c_size_t
    lib.acq_tc_pack_codebooks.argtypes = [pp, c_int, c_int, c_int, c_void_p, c_void_p]
    lib.acq_debug_tc_scores.argtypes = [c_void_p, pp, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                        c_void_p, c_void_p, c_void_p]
    lib.acq_vq_decode.argtypes = [c_void_p, c_int64, c_int64, pp, c_int, c_int, c_int, c_int, c_int,
                                  c_int, c_void_p, c_void_p, c_void_p]
    lib.acq_ema_stats.argtypes = [c_void_p, c_void_p, pp, c_int, c_int, c_int, c_int, c_int, c_int,
                                  c_void_p, c_void_p]
    lib.acq_rvq_replay.argtypes = [c_void_p, c_void_p, pp, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.acq_packed_bytes.argtypes = [c_int64, c_int]
    lib.acq_packed_bytes.restype = c_int64
    lib.acq_pack_codes.argtypes = [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]
    lib.acq_unpack_codes.argtypes = [c_void_p, c_int64, c_int, c_void_p, c_void_p]
    lib.acq_ema_apply.argtypes = [c_void_p, pp, pp, pp, c_int, c_int, c_int, c_double, c_double,
                                  c_void_p]
    lib.acq_pipeline_create.argtypes = [POINTER(c_void_p), c_int, c_size_t]
    lib.acq_pipeline_destroy.argtypes = [c_void_p]
    lib.acq_pipeline_destroy.restype = None
    lib.acq_pipeline_last_launches.argtypes = [c_void_p]
    lib.acq_pipeline_wait_stream.argtypes = [c_void_p, c_void_p]
    lib.acq_rvq_encode_host.argtypes = [c_void_p, c_void_p, pp, c_void_p, c_void_p, c_int, c_int, c_int,
                                        c_int, c_int, c_int, c_int, c_int, c_void_p]
    lib.acq_rvq_codec_host.argtypes = [c_void_p, c_void_p, pp, c_void_p, c_void_p, c_int, c_int, c_int,
                                       c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]
    lib.acq_vq_decode_host.argtypes = [c_void_p, c_void_p, c_int64, c_int64, pp, c_int, c_int, c_int,
                                       c_int, c_int, c_int, c_void_p]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if name not in ("acq_last_error", "acq_pipeline_destroy", "acq_tc_pack_bytes",
                        "acq_tc_workspace_bytes", "acq_packed_bytes"):
            fn.restype = c_int
    _lib = lib
    global _tc_defaults
    _tc_defaults = tuple(int(lib.acq_tc_query(i)) for i in range(3))
    return lib


_tc_defaults = None


def tc_config_defaults():
    """(variant, cluster, split) the library started with (environment / build defaults)."""
    load()
    return _tc_defaults


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().acq_last_error().decode("utf-8", "replace")
        raise AcqError(f"{what} failed (rc={rc}): {msg}")


def ptr_table(tensors):
    """HOST array of device pointers, as the C ABI expects for codebook tables."""
    arr = (c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    return ctypes.cast(arr, POINTER(c_void_p)), arr   # keep `arr` alive while the call runs
