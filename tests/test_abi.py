"""CPU-side checks of the drop-in boundary: the shared library loads and exports every symbol that
include/neucam_b200.h declares (no compute calls -- there is no GPU in the build container)."""
import ctypes
import os

from neucam_b200 import _lib


def test_library_is_built():
    assert os.path.exists(_lib.LIB_PATH), "run `python -m neucam_b200.build`"


def test_exports_every_declared_symbol():
    names = _lib.declared_symbols()
    assert "neucam_abi_version" in names and len(names) >= 3
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_abi_version_matches_header():
    src = open(_lib.HEADER_PATH).read()
    import re
    ver = int(re.search(r"#define NEUCAM_ABI_VERSION (\d+)", src).group(1))
    assert _lib.lib().neucam_abi_version() == ver


def test_debug_entry_points_are_not_in_the_product_library():
    """Probes, microbenchmarks and in-kernel timelines live in libneucam_b200_debug.so only."""
    prod = ctypes.CDLL(_lib.LIB_PATH)
    debug_only = [n for n in _lib.declared_symbols(_lib.DEBUG_HEADER_PATH) if n not in _lib.declared_symbols()]
    assert len(debug_only) >= 6
    assert not [n for n in debug_only if hasattr(prod, n)]
    dbg = ctypes.CDLL(_lib.DEBUG_LIB_PATH)
    assert not [n for n in debug_only + _lib.declared_symbols() if not hasattr(dbg, n)]


def test_step_args_struct_matches_the_header():
    """ctypes mirror of `struct neucam_step_args`: same field names in the same order as include/neucam_b200.h."""
    import re
    from neucam_b200 import ops
    src = open(_lib.HEADER_PATH).read()
    start = src.index("typedef struct neucam_step_args {") + len("typedef struct neucam_step_args {")
    body = src[start:src.index("} neucam_step_args;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl or decl.startswith("typedef"):
            continue
        for part in decl.split(","):
            m = re.search(r"\*?\s*\*?(\w+)\s*(\[\d+\])?$", part.strip())
            names.append(m.group(1))
    assert names == [f[0] for f in ops.StepArgs._fields_], (names, [f[0] for f in ops.StepArgs._fields_])
