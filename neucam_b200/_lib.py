"""ctypes binding of libneucam_b200.so (the C ABI declared in include/neucam_b200.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libneucam_b200.so")
DEBUG_LIB_PATH = os.path.join(_HERE, "libneucam_b200_debug.so")
HEADER_PATH = os.path.join(_HERE, "..", "include", "neucam_b200.h")
DEBUG_HEADER_PATH = os.path.join(_HERE, "..", "include", "neucam_b200_debug.h")

_lib = None
_debug_lib = None


class NeucamNativeError(RuntimeError):
    pass


def declared_symbols(header=None):
    """Names of every function include/neucam_b200.h (or the given header) declares."""
    src = open(header or HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\bint\s+(neucam_\w+)\s*\(", src)))


def lib():
    """Loads the library (building is a separate, explicit step: `python -m neucam_b200.build`)."""
    global _lib
    if _lib is None:
        variant = os.environ.get("NEUCAM_B200_LIB_VARIANT")
        if variant:
            # kernel A/B experiments (tools/ only): an alternative build of the SAME ABI placed next to the product
            # library as libneucam_b200_<variant>.so; never set by the package, the tests or bench.py
            path = os.path.join(_HERE, f"libneucam_b200_{variant}.so")
            if not os.path.exists(path):
                raise NeucamNativeError(f"NEUCAM_B200_LIB_VARIANT={variant}: {path} not found")
            _lib = ctypes.CDLL(path)
            for name in declared_symbols():
                getattr(_lib, name).restype = ctypes.c_int
            return _lib
        if not os.path.exists(LIB_PATH):
            raise NeucamNativeError(
                f"{LIB_PATH} not found: the CUDA extension is required (no CPU fallback). "
                "Run `python -m neucam_b200.build`.")
        _lib = ctypes.CDLL(LIB_PATH)
        for name in declared_symbols():
            getattr(_lib, name).restype = ctypes.c_int
    return _lib


def use_debug_library():
    """Diagnostic tools only (tools/*.py): makes lib() return the DEBUG build (same sources, -DNEUCAM_DEBUG: in-kernel
    timelines, experiment switches, building-block probes).  Must be called before the first lib() call; tests and
    the product path never call it."""
    global _lib
    if _lib is not None and _lib is not _debug_lib:
        raise NeucamNativeError("use_debug_library() must come before the product library is loaded")
    _lib = debug_lib()
    return _lib


def debug_lib():
    """The debug flavour of the library (include/neucam_b200_debug.h), loaded on its own handle."""
    global _debug_lib
    if _debug_lib is None:
        if not os.path.exists(DEBUG_LIB_PATH):
            raise NeucamNativeError(f"{DEBUG_LIB_PATH} not found. Run `python -m neucam_b200.build`.")
        _debug_lib = ctypes.CDLL(DEBUG_LIB_PATH)
        for name in declared_symbols() + declared_symbols(DEBUG_HEADER_PATH):
            getattr(_debug_lib, name).restype = ctypes.c_int
    return _debug_lib


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
