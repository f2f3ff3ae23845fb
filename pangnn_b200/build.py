"""In-tree nvcc build of the C-ABI shared library (sm_100a only).

``python -m pangnn_b200.build`` or ``pangnn_b200.build.build()``; ``__graft_entry__.build()`` calls
the latter.  The ``.so`` is git-ignored but travels to the GPU box with the repo snapshot.
"""
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libpangnn_b200.so")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _digest():
    """Hash of every source, header and flag that goes into the library (paths relative to the repository, so
    the stamp is the same on any checkout or box)."""
    h = hashlib.sha256()
    root = os.path.dirname(HERE)
    files = _sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + \
        [os.path.join(root, "include", "pangnn_b200.h")]
    for f in files:
        with open(f, "rb") as fh:
            h.update(os.path.relpath(f, root).encode()); h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def have_sources():
    return bool(_sources())


def is_current():
    """True when the built library carries the digest of the sources beside it (the digest is compiled into
    the library, ``pangnn_source_digest``; ``_abi.load`` refuses a stale one)."""
    if not os.path.exists(LIB_PATH):
        return False
    import ctypes
    try:
        fn = ctypes.CDLL(LIB_PATH).pangnn_source_digest
    except (OSError, AttributeError):
        return False
    fn.restype = ctypes.c_char_p
    return fn().decode() == _digest()


def build(force=False, verbose=False, prof=False):
    """Compile ``csrc/*.cu`` into ``_lib/libpangnn_b200.so`` unless it is up to date.  ``prof`` builds the
    development variant ``libpangnn_b200_prof.so`` (per-phase cycle counters in the scorer, ``tools/``)."""
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if prof:
        out = os.path.join(LIB_DIR, "libpangnn_b200_prof.so")
        r = subprocess.run([nvcc] + NVCC_FLAGS + ["-DPANGNN_SCORER_PROF", f'-DPANGNN_SRC_DIGEST="{_digest()}"', "-o", out]
                           + _sources(), capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed ({r.returncode}):\n{r.stdout}\n{r.stderr}")
        return out
    digest = _digest()
    if not force and is_current():
        return LIB_PATH
    cmd = [nvcc] + NVCC_FLAGS + [f'-DPANGNN_SRC_DIGEST="{digest}"'] + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", LIB_PATH] + _sources()
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed ({r.returncode}):\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, prof="--prof" in sys.argv))
