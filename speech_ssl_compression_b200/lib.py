"""ctypes binding of libmh_b200.so (the C ABI declared in include/mh_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, a
RuntimeError is raised (north_star: "no CPU fallback").
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_longlong, c_uint32, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MH_B200_LIB") or os.path.join(_HERE, "libmh_b200.so")  # override: A/B builds of the same ABI

EPI_BF16, EPI_GELU, EPI_RES, EPI_F32, EPI_DGELU, EPI_ADD, EPI_DELTA = range(7)


class GemmArgs(ctypes.Structure):
    _fields_ = [
        ("M", c_int), ("N", c_int), ("K", c_int),
        ("A", c_void_p), ("lda", c_longlong), ("a_mn", c_int),
        ("B", c_void_p), ("ldb", c_longlong), ("b_mn", c_int),
        ("D", c_void_p), ("ldd", c_longlong),
        ("epilogue", c_int),
        ("bias", c_void_p),
        ("aux_in", c_void_p),
        ("aux_out", c_void_p),
        ("ld_aux", c_longlong),
        ("mask", c_void_p),
        ("p_drop", c_float), ("seed", c_uint64), ("site", c_uint32),
        ("block_n", c_int),
        ("split_k", c_int),
        ("delta", c_void_p),
        ("delta_T", c_int),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C speech_ssl_compression_b200/csrc).  There is no CPU fallback.")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.mh_last_error.restype = c_char_p
        _lib.mh_launch_count.restype = c_longlong
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().mh_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def launch_count():
    return int(lib().mh_launch_count())


def ptr(t):
    """Raw device pointer of a tensor (or None)."""
    return None if t is None else c_void_p(t.data_ptr())


def stream_ptr():
    import torch

    return c_void_p(torch.cuda.current_stream().cuda_stream)
