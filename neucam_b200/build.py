"""Builds the C-ABI shared libraries in-tree with nvcc for sm_100a (no torch headers, pure C ABI).

    python -m neucam_b200.build [--force] [--debug]

  libneucam_b200.so        the product library (include/neucam_b200.h)
  libneucam_b200_debug.so  the same sources with -DNEUCAM_DEBUG: adds the building-block probes, the tensor-pipe
                           microbenchmark and the in-kernel timelines (include/neucam_b200_debug.h); used by
                           tests/test_umma_probe.py and tools/, never by the product path

Sources compile to objects in parallel (one nvcc per .cu), then link.  The built .so files are git-ignored but
travel to the GPU box with the gpurun snapshot.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libneucam_b200.so")
OUT_DEBUG = os.path.join(HERE, "libneucam_b200_debug.so")
OBJ_DIR = os.path.join(HERE, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    # no --use_fast_math: it would turn tanhf into tanh.approx (2^-11), too coarse for the CRF / exposure
    # paths; the hot epilogues call __sinf/__cosf/__expf explicitly instead.
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(extra):
    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC))
    files.append(os.path.join(HERE, "..", "include", "neucam_b200.h"))
    files.append(os.path.join(HERE, "..", "include", "neucam_b200_debug.h"))
    files.append(os.path.abspath(__file__))
    for f in files:
        with open(f, "rb") as fh:
            h.update(f.encode())
            h.update(fh.read())
    h.update(repr(extra).encode())
    return h.hexdigest()


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _build_one(out, defines, force, verbose, extra_flags=()):
    tag = os.path.basename(out).replace(".so", "")
    stamp = os.path.join(HERE, f".build_stamp_{tag}")
    dig = _digest((defines, tuple(extra_flags)))
    if not force and os.path.exists(out) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return out
    obj_dir = os.path.join(OBJ_DIR, tag)
    os.makedirs(obj_dir, exist_ok=True)
    inc = ["-I", os.path.join(HERE, "..", "include")]
    defs = [f"-D{d}" for d in defines]

    def compile_one(src):
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc_path()] + NVCC_FLAGS + list(extra_flags) + defs + inc + ["-c", src, "-o", obj]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        return obj, " ".join(cmd) + "\n" + proc.stdout + proc.stderr, proc.returncode

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, _sources()))
    log = "".join(r[1] for r in results)
    rc = max(r[2] for r in results)
    if rc == 0:
        cmd = [nvcc_path(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + [r[0] for r in results]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        log += " ".join(cmd) + "\n" + proc.stdout + proc.stderr
        rc = proc.returncode
    with open(os.path.join(HERE, f"build_{tag}.log"), "w") as fh:
        fh.write(log)
    if rc != 0:
        sys.stderr.write(log)
        raise RuntimeError(f"nvcc failed building {os.path.basename(out)}")
    if verbose:
        print(log)
    with open(stamp, "w") as fh:
        fh.write(dig)
    return out


def build(force=False, verbose=False, debug=True):
    """Builds the product library and (debug=True) the debug flavour; returns the product library's path."""
    out = _build_one(OUT, (), force, verbose)
    if debug:
        _build_one(OUT_DEBUG, ("NEUCAM_DEBUG",), force, verbose)
    return out


def build_variant(name, defines, force=False, verbose=False):
    """Kernel A/B experiments: the same sources with extra -D flags -> libneucam_b200_<name>.so, selected at run time by
    NEUCAM_B200_LIB_VARIANT=<name> (tools/ only; see _lib.lib())."""
    return _build_one(os.path.join(HERE, f"libneucam_b200_{name}.so"), tuple(defines), force, verbose)


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:], force="--force" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose="--quiet" not in sys.argv, debug=True))
