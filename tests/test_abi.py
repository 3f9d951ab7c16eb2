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
