"""TEST INFRASTRUCTURE ONLY -- import shim that lets the UNMODIFIED reference (/root/reference) run in
the build container (CPU-only, several optional dependencies absent).

Used by oracle/make_golden.py and tests that validate oracle/neucam_oracle.py against the real
reference when /root/reference is present.  Nothing under neucam_b200/ may import this.

What it does (no reference source is edited or copied):
  * registers an empty `torchmeta` package whose __path__ points into the reference so that
    `torchmeta.modules` imports without executing torchmeta/__init__.py (which needs h5py);
  * stubs modules the hot path never touches (matplotlib, skimage, skvideo, imageio, cmapy,
    tensorboardX, configargparse);
  * on a CUDA-less host maps `.cuda()` / `.to('cuda')` to no-ops, because the reference hard-codes
    CUDA placement (modules.py:223,230; training.py:92-93,181,244,250; utils.py:443,610).
"""
import importlib
import os
import sys
import types
from unittest import mock

REFERENCE_ROOT = os.environ.get("NEUCAM_REFERENCE_ROOT", "/root/reference")

_installed = False


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "modules.py"))


def install():
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import torch

    pkg = types.ModuleType("torchmeta")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "torchmeta")]
    sys.modules.setdefault("torchmeta", pkg)

    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "skimage", "skimage.measure",
                 "skimage.filters", "skimage.metrics", "skvideo", "skvideo.io", "imageio", "cmapy",
                 "tensorboardX", "configargparse", "skimage.transform"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = mock.MagicMock(name=name)

    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        _orig_to = torch.Tensor.to

        def _to(self, *args, **kwargs):
            args = tuple("cpu" if (isinstance(a, str) and a.startswith("cuda")) else a for a in args)
            if isinstance(kwargs.get("device"), str) and kwargs["device"].startswith("cuda"):
                kwargs["device"] = "cpu"
            return _orig_to(self, *args, **kwargs)

        torch.Tensor.to = _to

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def load(name):
    """Imports a top-level reference module (modules, training, utils, loss_functions, dataio)."""
    install()
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return importlib.import_module(name)
