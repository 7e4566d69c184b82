"""Import the UNMODIFIED reference (staged by oracle/build_ref.py) as the package `lnasr`  --  TEST / BENCH INFRASTRUCTURE ONLY.

Used by `bench.py --impl reference` / the `cpu_baseline` leg (worker processes) and by tests that cross-check the numpy port
(oracle/lnasr_oracle.py) against the real thing on the GPU box.  The reference imports absolutely (`from lnasr.utils import
lse`), i.e. it must BE `lnasr` in the importing process: call `load()` only in a process (a forked / spawned worker) that
does not use this repository's own `lnasr` package afterwards -- it evicts it from sys.modules.

Import shims (the reference's code is untouched): stub modules for pyaudio / librosa / h5py (absent from the
image, unused on the hot path) and the numpy-1 aliases np.mat / np.float the reference still uses (gmm.py:63,85,97).
"""
import os
import sys
import types
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ARCHIVE = os.path.join(HERE, "_ref", "lnasr_reference.zip")


def available() -> bool:
    return os.path.exists(ARCHIVE)


def load():
    """-> dict(utils, mfcc, gmm, hmm, gmmhmm) of reference modules."""
    import numpy as np
    if not available():
        raise RuntimeError("oracle/_ref/lnasr_reference.zip is missing: run `python oracle/build_ref.py` where /root/reference exists")
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    for name in ("pyaudio", "librosa", "h5py"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if not hasattr(np, "mat"):
        np.mat = np.asmatrix
    if not hasattr(np, "float"):
        np.float = float
    for name in [k for k in sys.modules if k == "lnasr" or k.startswith("lnasr.")]:
        mod = sys.modules[name]
        if ARCHIVE not in (getattr(mod, "__file__", "") or ""):
            del sys.modules[name]
    if ARCHIVE not in sys.path[:1]:
        sys.path.insert(0, ARCHIVE)
    import importlib
    importlib.invalidate_caches()
    out = {}
    for m in ("utils", "mfcc", "gmm", "hmm", "gmmhmm"):
        out[m] = importlib.import_module("lnasr." + m)
        assert ARCHIVE in out[m].__file__, f"lnasr.{m} was not imported from the staged reference archive"
    return out
