"""ctypes binding of libneucam_b200.so (the C ABI declared in include/neucam_b200.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libneucam_b200.so")
HEADER_PATH = os.path.join(_HERE, "..", "include", "neucam_b200.h")

_lib = None


class NeucamNativeError(RuntimeError):
    pass


def declared_symbols():
    """Names of every function include/neucam_b200.h declares."""
    src = open(HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\bint\s+(neucam_\w+)\s*\(", src)))


def lib():
    """Loads the library (building is a separate, explicit step: `python -m neucam_b200.build`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NeucamNativeError(
                f"{LIB_PATH} not found: the CUDA extension is required (no CPU fallback). "
                "Run `python -m neucam_b200.build`.")
        _lib = ctypes.CDLL(LIB_PATH)
        for name in declared_symbols():
            getattr(_lib, name).restype = ctypes.c_int
    return _lib


def check(rc, what):
    if rc != 0:
        raise NeucamNativeError(f"{what} failed with status {rc}")


def ptr(t):
    """Device pointer of a torch tensor (or None -> NULL) as c_void_p."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def cur_stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
