"""Model files either side of the hot path: HMM.save / load (hmm.py:252-266), GMMHMM.save / load
(gmmhmm.py:190-208) and recognizer.AcousticModel (recognizer.py:20-26) through the h5py API.  h5py is absent from the image;
tests/shims/h5py_shim.py stands in for it (same calls, its own container), so these code paths are executed.  CPU only."""
import importlib
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture()
def h5py_shim(monkeypatch):
    spec = importlib.util.spec_from_file_location("h5py", os.path.join(HERE, "shims", "h5py_shim.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    monkeypatch.setitem(sys.modules, "h5py", mod)
    return mod


def test_gmmhmm_h5_round_trip(tmp_path, h5py_shim):
    from lnasr.gmmhmm import GMMHMM
    np.random.seed(11)
    g = GMMHMM(n=6, m=3, d=39)
    g.reset("random")
    g.si = g.si * 2.5
    path = str(tmp_path / "am.h5")
    g.save(path)                                                       # h5py.File(path, 'w') + create_dataset x5 (gmmhmm.py:190-198)
    with h5py_shim.File(path, "r") as f:
        assert sorted(f.keys()) == ["A", "mu", "pi", "si", "w"] and f["si"].shape == (6, 3, 39, 39)
    h = GMMHMM(n=2, m=2, d=39).load(path)                              # shapes come from the file (gmmhmm.py:205-207)
    assert (h.n, h.m, h.d) == (6, 3, 39)
    for k in ("A", "pi", "w", "mu", "si"):
        assert np.array_equal(getattr(h, k), getattr(g, k)) and isinstance(getattr(h, k), np.ndarray)


def test_hmm_h5_round_trip(tmp_path, h5py_shim):
    from lnasr.hmm import HMM
    np.random.seed(12)
    a = HMM(n=4, m=7)
    a.reset("random")
    path = str(tmp_path / "hmm.h5")
    a.save(path)                                                       # datasets A, B, pi (hmm.py:252-258)
    b = HMM()
    b.load(path)
    assert (b.n, b.m) == (4, 7)
    assert np.array_equal(b.A, a.A) and np.array_equal(b.B, a.B) and np.array_equal(b.pi, a.pi)


def test_recognizer_loads_an_h5_acoustic_model(tmp_path, h5py_shim):
    """recognizer.AcousticModel: MFCC() + GMMHMM(n=2, m=2, d=MFCC.D).load(file) (recognizer.py:23-26) from an .h5 name."""
    from lnasr.gmmhmm import GMMHMM
    from lnasr.recognizer import AcousticModel
    np.random.seed(13)
    g = GMMHMM(n=5, m=4, d=39)
    g.reset("random")
    path = str(tmp_path / "word.h5")
    g.save(path)
    am = AcousticModel(path)
    assert (am.gmmhmm.n, am.gmmhmm.m, am.gmmhmm.d) == (5, 4, 39) and np.array_equal(am.gmmhmm.mu, g.mu)


def test_without_h5py_the_h5_names_fail_loudly(tmp_path, monkeypatch):
    from lnasr.gmmhmm import GMMHMM
    monkeypatch.setitem(sys.modules, "h5py", None)                      # import h5py -> ImportError
    g = GMMHMM(n=2, m=2, d=3)
    g.reset("uniform")
    with pytest.raises(ImportError):
        g.save(str(tmp_path / "x.h5"))
    g.save(str(tmp_path / "x.npz"))                                     # the numpy side door needs nothing
    assert np.array_equal(GMMHMM().load(str(tmp_path / "x.npz")).mu, g.mu)
