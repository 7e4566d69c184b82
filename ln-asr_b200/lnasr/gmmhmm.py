"""GMM-HMM -- API mirror of the reference's lnasr/gmmhmm.py (gmmhmm.py:17-208) on a B200.

An HMM whose state emissions are M-component Gaussian mixtures over d-dimensional features.
A, pi and w are natural-log probabilities; mu [n,m,d] and si [n,m,d,d] are linear (gmmhmm.py:31).
Emissions come from csrc/gmm_exact.cu (float64) or, for large models, the tensor-core kernel
csrc/gmm_tc.cu; the trellis and the Baum-Welch statistics from csrc/trellis.cu and csrc/em.cu.

The build target is DIAGONAL covariance (BASELINE.json north star): `si` keeps the
reference's [n,m,d,d] shape, a full `si` is honoured when scoring (full-covariance kernel), and
re-estimation returns diag(Sigma') embedded in [n,m,d,d] -- the diagonal of what the reference computes
(its off-diagonals are dropped).  `var` is the [n,m,d] diagonal view.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np

from . import _capi, _dev
from .gmm import PackedGmm
from .hmm import HMM

__all__ = ["GMMHMM"]


class GMMHMM(HMM):
    """n states x m mixtures x d dims (gmmhmm.py:17-44)."""

    _obs_np_dtype = np.float64

    def __init__(self, n: int = 1, m: int = 1, d: int = 1, A=None, pi=None, w=None, mu=None, si=None,
                 precision=np.double):
        super().__init__(n, m, A, None, pi, precision)
        self.d = d
        self.w = w
        self.mu = mu
        self.si = si
        self.min_std = 0.01        # added to the VARIANCE diagonal in the M-step (gmmhmm.py:44,147)
        self._bm_np = None
        self._bm_dev = None
        self._gmm_key = None
        self._gmm = None

    # ------------------------------------------------------------------ views
    @property
    def var(self) -> np.ndarray:
        """Diagonal of `si`, [n,m,d]."""
        return np.diagonal(np.asarray(self.si), axis1=2, axis2=3).copy()

    @property
    def bm(self):
        """Per-mixture log densities [n, m, T] of the last `_map_b` (gmmhmm.py:59), downloaded on first access."""
        if self._bm_np is None and self._bm_dev is not None:
            T = self._bm_dev.shape[0]
            a = self._bm_dev.cpu().numpy().reshape(T, self.n, self.m)
            self._bm_np = np.ascontiguousarray(a.transpose(1, 2, 0)).astype(self.precision, copy=False)
        return self._bm_np

    @bm.setter
    def bm(self, value):
        self._bm_np = None if value is None else np.asarray(value)
        self._bm_dev = None

    def _bm_device(self):
        if self._bm_dev is None:
            if self._bm_np is None:
                raise AttributeError("no cached mixture densities: call _map_b(obs) first")
            t = _dev.torch()
            a = np.asarray(self._bm_np, dtype=np.float64)                  # [n, m, T]
            self._bm_dev = _dev.to_dev(np.ascontiguousarray(a.transpose(2, 0, 1)).reshape(a.shape[2], -1), t.float64)
        return self._bm_dev

    def _packed(self) -> PackedGmm:
        w, mu, si = np.asarray(self.w), np.asarray(self.mu), np.asarray(self.si)
        key = self._content_key(w, mu, si)                 # content hash: in-place edits of w / mu / si are seen
        if self._gmm is None or self._gmm_key != key:
            self._gmm = PackedGmm(w, mu, si)
            self._gmm_key = key
        return self._gmm

    # ------------------------------------------------------------------ emissions
    def _obs_device(self, obs):
        t = _dev.torch()
        if isinstance(obs, t.Tensor):
            return _dev.to_dev(obs)                       # float32 / float64 features already on the device
        return _dev.to_dev(np.asarray(obs, dtype=np.float64), t.float64)

    def _emissions_dev(self, x_d, want_bm=False):
        """x_d [T, d] device -> (b [T,n] float64, bm [T, n*m] float64 or None)   (gmmhmm.py:46-67)."""
        if x_d.dim() != 2 or x_d.shape[1] < self.d:
            raise ValueError(f"observations must be [T, {self.d}], got {tuple(x_d.shape)}")
        return self._packed().loglik(x_d, want_bm=want_bm)

    def _map_b(self, obs: np.ndarray):
        """Fill the cached `b` [n,T] and `bm` [n,m,T] (gmmhmm.py:46-67)."""
        self._b_np = self._bm_np = None
        self._b_dev, self._bm_dev = self._emissions_dev(self._obs_device(obs), want_bm=True)

    # ------------------------------------------------------------------ EM
    def _acc_size(self) -> int:
        return int(_capi.lib().lnasr_em_acc_doubles(self.n, self.m, self.d))

    _fb_split = None      # None: choose by batch size; True / False force the two-kernel / the fused Baum-Welch path

    def _accumulate(self, x_d, b, bm, alpha, f_off_d, n_utt, total, acc, split=None):
        """Baum-Welch statistics of a batch into `acc`.  split=True: two-kernel form (lane groups per utterance + frame-parallel
        statistics), False: one fused kernel with a CTA per utterance; default: split when the batch has >= 8 utterances."""
        t = _dev.torch()
        g = self._packed()
        lib = _capi.lib()
        if g.full and not getattr(self, "_warned_full", False):
            import warnings
            warnings.warn("GMMHMM re-estimation is built for diagonal covariance: the scatter statistics keep diag(Sigma') only "
                          "(the reference's train() returns the full matrix, gmmhmm.py:138-147); the off-diagonals of `si` are "
                          "used for scoring and dropped by the update", RuntimeWarning, stacklevel=3)
            self._warned_full = True
        if self.n > 32:
            # backward trellis pass (any n, ELL transition lists) + the frame-parallel general statistics kernels
            beta = self._trans().backward(b, f_off_d, n_utt, total)
            nbytes = int(lib.lnasr_em_statistics_workspace_bytes(total, self.n))
            ws = _dev.empty((nbytes,), t.uint8)
            _capi.check(lib.lnasr_em_statistics_ws(
                _capi.ptr(x_d), _dev.np_dtype_code(x_d.dtype), x_d.stride(0), _capi.ptr(b), _capi.ptr(bm), _capi.ptr(alpha),
                _capi.ptr(beta), _capi.ptr(f_off_d), n_utt, total, _capi.ptr(self._trans().logA), _capi.ptr(g.logw_d),
                _capi.ptr(g.mu_d), self.n, self.m, self.d, _capi.ptr(acc), None, _capi.ptr(ws), nbytes, _capi.stream_ptr()),
                "em_statistics_ws")
            return
        if split is None:
            split = self._fb_split if self._fb_split is not None else n_utt >= 8
        ws, ws_bytes = None, 0
        if split:
            ws_bytes = int(lib.lnasr_fb_workspace_bytes(total, self.n))
            ws = _dev.empty((ws_bytes,), t.uint8)
        _capi.check(lib.lnasr_fb_accumulate(
            _capi.ptr(x_d), _dev.np_dtype_code(x_d.dtype), x_d.stride(0), _capi.ptr(b), _capi.ptr(bm), _capi.ptr(alpha),
            _capi.ptr(f_off_d), n_utt, total, _capi.ptr(self._trans().logA), _capi.ptr(g.logw_d), _capi.ptr(g.mu_d),
            self.n, self.m, self.d, _capi.ptr(acc), None, None, _capi.ptr(ws), ws_bytes, _capi.stream_ptr()), "fb_accumulate")

    def _finalize(self, acc) -> dict:
        t = _dev.torch()
        n, m, d = self.n, self.m, self.d
        A = _dev.empty((n, n), t.float64)
        pi = _dev.empty((n,), t.float64)
        w = _dev.empty((n, m), t.float64)
        mu = _dev.empty((n, m, d), t.float64)
        var = _dev.empty((n, m, d), t.float64)
        _capi.check(_capi.lib().lnasr_mstep_finalize(_capi.ptr(acc), n, m, d, float(self.min_std), _capi.ptr(A),
                                                    _capi.ptr(pi), _capi.ptr(w), _capi.ptr(mu), _capi.ptr(var),
                                                    _capi.stream_ptr()), "mstep_finalize")
        var_np = self._out(var)
        si = np.zeros((n, m, d, d), dtype=var_np.dtype)
        idx = np.arange(d)
        si[:, :, idx, idx] = var_np
        return {"A": self._out(A), "pi": self._out(pi), "w": self._out(w), "mu": self._out(mu), "si": si}

    def _estimate(self, obs: np.ndarray, alpha: np.ndarray, beta: np.ndarray, xi: np.ndarray, gamma: np.ndarray) -> dict:
        """M-step -> {'A','pi','w','mu','si'} (gmmhmm.py:69-156), si = diag(Sigma') embedded in [n,m,d,d].
        Statistics are re-derived on the device from `alpha`, the cached `b`/`bm` and `obs` by the fused
        backward/accumulate kernel; `beta`, `xi`, `gamma` are accepted for signature parity."""
        t = _dev.torch()
        b, bm = self._b_device(), self._bm_device()
        T = b.shape[0]
        al = _dev.to_dev(np.asarray(alpha, dtype=np.float64), t.float64)
        acc = _dev.zeros((self._acc_size(),), t.float64)
        self._accumulate(self._obs_device(obs), b, bm, al, self._one(T), 1, T, acc)
        return self._finalize(acc)

    def _updatemodel(self, model):
        """Adopt re-estimated parameters (gmmhmm.py:158-164)."""
        self.A = model["A"]
        self.pi = model["pi"]
        self.mu = model["mu"]
        self.si = model["si"]
        self.w = model["w"]
        self._drop_emission_cache()

    def _drop_emission_cache(self):
        self._b_np = self._b_dev = self._bm_np = self._bm_dev = None

    def _remember_emissions(self, obs_list, b, bm=None):
        if len(obs_list) == 1:
            self._b_np, self._b_dev = None, b
            self._bm_np, self._bm_dev = None, bm

    def _empty_obs(self):
        return np.zeros((0, self.d), dtype=self._obs_np_dtype)

    def reset(self, init_type):
        """'uniform' or 'random' initial parameters, same draw order as the reference (gmmhmm.py:166-188)."""
        n, m, d, p = self.n, self.m, self.d, self.precision
        if init_type == "uniform":
            self.A = np.log(np.full((n, n), 1.0 / n, dtype=p))
            self.pi = np.log(np.full(n, 1.0 / n, dtype=p))
            self.w = np.log(np.full((n, m), 1.0 / m, dtype=p))
            self.mu = np.zeros((n, m, d), dtype=p)
        elif init_type == "random":
            a = np.random.random_sample((n, n))
            pi = np.random.random_sample(n)
            w = np.random.random_sample((n, m))
            with np.errstate(divide="ignore"):
                self.A = np.log(a / a.sum(axis=1, keepdims=True))
                self.pi = np.log(pi / pi.sum())
                self.w = np.log(w / w.sum(axis=1, keepdims=True))
            self.mu = np.array(0.6 * np.random.random_sample((n, m, d)) - 0.3, dtype=p)
        else:
            return
        self.si = np.zeros((n, m, d, d), dtype=p)
        self.si[:, :] = np.eye(d, dtype=p)

    def save(self, filename):
        """HDF5 with datasets 'A','pi','w','mu','si' at the root (gmmhmm.py:190-198); '.npz' names use numpy."""
        arrs = {k: np.asarray(getattr(self, k)) for k in ("A", "pi", "w", "mu", "si")}
        if str(filename).endswith(".npz"):
            np.savez(filename, **arrs)
            return
        import h5py
        with h5py.File(filename, "w") as f:
            for k, v in arrs.items():
                f.create_dataset(k, data=v)

    def load(self, filename):
        """Read 'A','pi','w','mu','si'; n, m, d come from mu's shape (gmmhmm.py:200-208)."""
        if str(filename).endswith(".npz"):
            f = np.load(filename)
            vals = {k: np.array(f[k]) for k in ("A", "pi", "w", "mu", "si")}
        else:
            import h5py
            with h5py.File(filename, "r") as f:
                vals = {k: np.array(f[k]) for k in ("A", "pi", "w", "mu", "si")}
        self.A, self.pi, self.w, self.mu, self.si = (vals[k] for k in ("A", "pi", "w", "mu", "si"))
        self.n, self.m, self.d = self.mu.shape
        return self
