"""Discrete-emission HMM in the log domain -- API mirror of the reference's lnasr/hmm.py (hmm.py:16-313),
with the trellis (forward, backward, Viterbi + backtrace) and the Baum-Welch statistics computed by
csrc/trellis.cu and csrc/em.cu on a B200.

A, B and pi are natural-log probabilities (hmm.py:27).  Every method keeps the reference's name,
arguments and return shapes; results are numpy arrays of dtype `precision`.  Additive batch methods
(`*_batch`) take lists of observation sequences and keep the work on the device.
"""
from __future__ import annotations

import zlib
from typing import List, Sequence

import numpy as np

from . import _capi, _dev
from ._trellis import TransitionBank

__all__ = ["HMM"]


class HMM:
    """n hidden states, m observation symbols; A [n,n], B [n,m], pi [n] in log domain (hmm.py:17-45)."""

    def __init__(self, n: int = 1, m: int = 1, A=None, B=None, pi=None, precision=np.double):
        self.n = n
        self.m = m
        self.A = A
        self.B = B
        self.pi = pi
        self.precision = precision
        self._b_np = None          # cached emissions, reference layout [n, T]
        self._b_dev = None         # same on the device, frame-major [T, n]
        self._bank_key = None
        self._bank = None

    # ------------------------------------------------------------------ cached emissions (hmm.py:59)
    @property
    def b(self):
        """Emission log-probabilities [n, T] of the last `_map_b` (downloaded on first access)."""
        if self._b_np is None and self._b_dev is not None:
            self._b_np = np.ascontiguousarray(self._b_dev.cpu().numpy().T).astype(self.precision, copy=False)
        return self._b_np

    @b.setter
    def b(self, value):
        self._b_np = None if value is None else np.asarray(value)
        self._b_dev = None

    def _b_device(self):
        """[T, n] float64 device copy of the cached emissions."""
        if self._b_dev is None:
            if self._b_np is None:
                raise AttributeError("no cached emissions: call _map_b(obs) first or pass remap=True")
            t = _dev.torch()
            self._b_dev = _dev.to_dev(np.ascontiguousarray(np.asarray(self._b_np, dtype=np.float64).T), t.float64)
        return self._b_dev

    # ------------------------------------------------------------------ device views of the parameters
    @staticmethod
    def _content_key(*arrays):
        """Cache key from the CONTENT of the parameter arrays (shape + CRC-32 + Adler-32 of the bytes): an in-place edit such as
        h.A[i, j] = -inf, or a new array that happens to reuse a freed object's id, must not leave stale device copies behind."""
        key = []
        for a in arrays:
            a = np.ascontiguousarray(a)
            if a.nbytes > (64 << 20):
                # very large parameter arrays (a 2000-state full `si` is 390 MB): hash ~1M evenly spaced elements plus the
                # object's identity; an in-place edit that misses every sampled element needs an explicit invalidate()
                flat = a.reshape(-1)
                sample = np.ascontiguousarray(flat[::max(1, flat.size >> 20)])
                buf = sample.view(np.uint8)
                key.append((a.shape, a.dtype.str, id(a), zlib.crc32(buf), zlib.adler32(buf)))
                continue
            buf = a.view(np.uint8).reshape(-1) if a.size else b""
            key.append((a.shape, a.dtype.str, zlib.crc32(buf), zlib.adler32(buf)))
        return tuple(key)

    def invalidate(self):
        """Drop every cached device copy of the parameters and emissions (they are re-derived from the numpy attributes on
        the next call).  Needed only after an in-place edit of a parameter array larger than 64 MB."""
        self._bank = self._bank_key = None
        self._drop_emission_cache()
        if hasattr(self, "_gmm"):
            self._gmm = self._gmm_key = None

    def _trans(self) -> TransitionBank:
        A = np.asarray(self.A, dtype=np.float64)
        pi = np.asarray(self.pi, dtype=np.float64).reshape(-1)
        key = self._content_key(A, pi)
        if self._bank is None or self._bank_key != key:
            self._bank = TransitionBank(A, pi)
            self._bank_key = key
        return self._bank

    @staticmethod
    def _one(T):
        t = _dev.torch()
        off = np.array([0, T], dtype=np.int64)
        return _dev.to_dev(off, t.int64)

    def _out(self, x):
        return x.cpu().numpy().astype(self.precision, copy=False)

    # ------------------------------------------------------------------ reference methods
    def _map_b(self, obs: np.ndarray):
        """b[:, t] = B[:, obs[t]] (hmm.py:47-61)."""
        self._b_np = None
        self._b_dev = self._emissions_dev(self._obs_device(np.asarray(obs)))[0]

    def _obs_device(self, obs):
        """Observations of one or many concatenated utterances on the device (symbol indices, int32)."""
        t = _dev.torch()
        return _dev.to_dev(np.asarray(obs).astype(np.int32), t.int32)

    def _emissions_dev(self, obs_d, want_bm=False):
        """obs_d [T] int32 device -> (b [T,n] float64 device, None)."""
        t = _dev.torch()
        logB = _dev.to_dev(np.asarray(self.B, dtype=np.float64), t.float64)
        n, m = logB.shape
        b = _dev.empty((obs_d.numel(), n), t.float64)
        _capi.check(_capi.lib().lnasr_emit_gather(_capi.ptr(logB), _capi.ptr(obs_d), obs_d.numel(), n, m, _capi.ptr(b),
                                                 _capi.stream_ptr()), "emit_gather")
        return b, None

    def _forward(self, obs: np.ndarray, remap: bool = False) -> np.ndarray:
        """alpha [T, n] (hmm.py:63-100)."""
        if remap:
            self._map_b(obs)
        b = self._b_device()
        T = b.shape[0]
        alpha, _ = self._trans().forward(b, self._one(T), 1, T, want_alpha=True)
        return self._out(alpha)

    def _backward(self, obs: np.ndarray, remap: bool = False) -> np.ndarray:
        """beta [T, n] (hmm.py:102-131)."""
        if remap:
            self._map_b(obs)
        b = self._b_device()
        T = b.shape[0]
        return self._out(self._trans().backward(b, self._one(T), 1, T))

    def _viterbi(self, obs: np.ndarray, remap: bool = False):
        """(v [T,n], bt [T,n] uint32, path [T] uint32) (hmm.py:133-172)."""
        if remap:
            self._map_b(obs)
        b = self._b_device()
        T = b.shape[0]
        r = self._trans().viterbi(b, self._one(T), 1, T, want_path=True, want_score=False, want_v=True, want_bt=True)
        return (self._out(r["v"]), r["bt"].cpu().numpy().astype(np.uint32), r["path"][0].cpu().numpy().astype(np.uint32))

    def _baumwelch(self, obs: np.ndarray, alpha: np.ndarray, beta: np.ndarray, remap: bool = False):
        """(xi [T,n,n], gamma [T,n]), row T-1 = -inf (hmm.py:174-200)."""
        t = _dev.torch()
        if remap:
            self._map_b(obs)
        b = self._b_device()
        T, n = b.shape
        al = _dev.to_dev(np.asarray(alpha, dtype=np.float64), t.float64)
        be = _dev.to_dev(np.asarray(beta, dtype=np.float64), t.float64)
        xi = _dev.empty((T, n, n), t.float64)
        gamma = _dev.empty((T, n), t.float64)
        _capi.check(_capi.lib().lnasr_baumwelch_xi(_capi.ptr(al), _capi.ptr(be), _capi.ptr(b), _capi.ptr(self._trans().logA),
                                                  _capi.ptr(self._one(T)), 1, T, n, _capi.ptr(xi), _capi.ptr(gamma),
                                                  _capi.stream_ptr()), "baumwelch_xi")
        return self._out(xi), self._out(gamma)

    def _acc_size(self) -> int:
        return int(_capi.lib().lnasr_hmm_acc_doubles(self.n, self.m))

    def _accumulate(self, obs_d, b, bm, alpha, f_off_d, n_utt, total, acc):
        """Baum-Welch statistics of a batch into `acc` (csrc/em.cu): the backward-fused kernel for n <= 32, else the backward
        trellis pass (any n, ELL transition lists) + the frame-parallel general statistics kernels."""
        t = _dev.torch()
        lib = _capi.lib()
        if self.n > 32:
            beta = self._trans().backward(b, f_off_d, n_utt, total)
            nbytes = int(lib.lnasr_em_statistics_workspace_bytes(total, self.n))
            ws = _dev.empty((nbytes,), t.uint8)
            _capi.check(lib.lnasr_hmm_em_statistics_ws(
                _capi.ptr(b), _capi.ptr(alpha), _capi.ptr(beta), _capi.ptr(obs_d), _capi.ptr(f_off_d), n_utt, total,
                _capi.ptr(self._trans().logA), self.n, self.m, _capi.ptr(acc), None, _capi.ptr(ws), nbytes, _capi.stream_ptr()),
                "hmm_em_statistics_ws")
            return
        _capi.check(lib.lnasr_hmm_fb_accumulate(
            _capi.ptr(b), _capi.ptr(alpha), _capi.ptr(obs_d), _capi.ptr(f_off_d), n_utt, total,
            _capi.ptr(self._trans().logA), self.n, self.m, _capi.ptr(acc), None, None, _capi.stream_ptr()),
            "hmm_fb_accumulate")

    def _finalize(self, acc) -> dict:
        t = _dev.torch()
        A = _dev.empty((self.n, self.n), t.float64)
        B = _dev.empty((self.n, self.m), t.float64)
        pi = _dev.empty((self.n,), t.float64)
        _capi.check(_capi.lib().lnasr_hmm_mstep_finalize(_capi.ptr(acc), self.n, self.m, _capi.ptr(A), _capi.ptr(B),
                                                        _capi.ptr(pi), _capi.stream_ptr()), "hmm_mstep_finalize")
        return {"A": self._out(A), "B": self._out(B), "pi": self._out(pi)}

    def _estimate(self, obs: np.ndarray, alpha: np.ndarray, beta: np.ndarray, xi: np.ndarray, gamma: np.ndarray) -> dict:
        """M-step -> {'A','B','pi'} (hmm.py:202-225).  The statistics are re-derived on the device from `alpha`
        and the cached emissions by the fused backward/accumulate kernel; `beta`, `xi`, `gamma` are accepted
        for signature parity (they are functions of alpha, A and b)."""
        t = _dev.torch()
        b = self._b_device()
        T = b.shape[0]
        al = _dev.to_dev(np.asarray(alpha, dtype=np.float64), t.float64)
        acc = _dev.zeros((self._acc_size(),), t.float64)
        self._accumulate(self._obs_device(obs), b, None, al, self._one(T), 1, T, acc)
        return self._finalize(acc)

    def _updatemodel(self, model: dict):
        """Adopt re-estimated parameters (hmm.py:227-231)."""
        self.A = model["A"]
        self.B = model["B"]
        self.pi = model["pi"]
        self._drop_emission_cache()

    def _drop_emission_cache(self):
        """Emissions cached under the OLD parameters must not feed a later _forward(obs) / _backward(obs) without remap."""
        self._b_np = self._b_dev = None

    def reset(self, init_type):
        """'uniform' or 'random' initial parameters (hmm.py:233-250); random draws lie in (0, 1] so no log(0)."""
        n, m = self.n, self.m
        if init_type == "uniform":
            self.A = np.log(np.full((n, n), 1.0 / n, dtype=self.precision))
            self.B = np.log(np.full((n, m), 1.0 / m, dtype=self.precision))
            self.pi = np.log(np.full(n, 1.0 / n, dtype=self.precision))
        elif init_type == "random":
            a = 1.0 - np.random.random_sample((n, n))
            b = 1.0 - np.random.random_sample((n, m))
            p = 1.0 - np.random.random_sample(n)
            self.A = np.log(a / a.sum(axis=1, keepdims=True))
            self.B = np.log(b / b.sum(axis=1, keepdims=True))
            self.pi = np.log(p / p.sum())

    def save(self, filename):
        """HDF5 with datasets 'A', 'B', 'pi' at the root (hmm.py:252-258)."""
        import h5py
        with h5py.File(filename, "w") as f:
            f.create_dataset("A", data=np.asarray(self.A))
            f.create_dataset("B", data=np.asarray(self.B))
            f.create_dataset("pi", data=np.asarray(self.pi))

    def load(self, filename):
        """Read 'A', 'B', 'pi' (hmm.py:260-266); n, m come from B's shape.  Arrays are materialised."""
        import h5py
        with h5py.File(filename, "r") as f:
            self.A = np.array(f["A"])
            self.B = np.array(f["B"])
            self.pi = np.array(f["pi"])
        self.n, self.m = self.B.shape

    def calc_prob(self, obs: np.ndarray) -> float:
        """log P(obs | model) = lse(alpha[T-1]) (hmm.py:268-274)."""
        return float(self.calc_prob_batch([obs])[0])

    def decode(self, obs: np.ndarray) -> np.ndarray:
        """Most likely state path, uint32 [T] (hmm.py:276-283)."""
        return self.decode_batch([obs])[0]

    def train(self, obs, iters=1, eps=0.0001, verbose: bool = True):
        """Baum-Welch on one observation sequence with the reference's stopping rule and log line (hmm.py:285-313)."""
        return self.train_batch([obs], iters=iters, eps=eps, verbose=verbose)

    # ------------------------------------------------------------------ batch API (additive)
    _obs_np_dtype = np.int32

    def _pack(self, obs_list: Sequence):
        """-> (observations concatenated on the device, host offsets, device offsets)."""
        t = _dev.torch()
        cat, off = _dev.pack_ragged(obs_list, self._obs_np_dtype)
        if len(obs_list) == 0:
            cat = self._empty_obs()
        return self._obs_device(cat), off, _dev.to_dev(off, t.int64)

    def _empty_obs(self):
        return np.zeros((0,), dtype=self._obs_np_dtype)

    def _remember_emissions(self, obs_list, b, bm=None):
        """The reference's calc_prob / decode / train run _forward(obs, True) / _viterbi(obs, True), which leave self.b (and
        self.bm) filled as a side effect (hmm.py:82-83, 154-155): keep that for the single-sequence calls."""
        if len(obs_list) == 1:
            self._b_np, self._b_dev = None, b

    def calc_prob_batch(self, obs_list: Sequence) -> np.ndarray:
        x_d, off, off_d = self._pack(obs_list)
        if int(off[-1]) == 0:
            return np.full(len(off) - 1, -np.inf)
        b, bm = self._emissions_dev(x_d, want_bm=len(obs_list) == 1)
        self._remember_emissions(obs_list, b, bm)
        _, lp = self._trans().forward(b, off_d, len(off) - 1, int(off[-1]))
        return lp[:, 0].cpu().numpy()

    def decode_batch(self, obs_list: Sequence) -> List[np.ndarray]:
        x_d, off, off_d = self._pack(obs_list)
        if int(off[-1]) == 0:
            return [np.zeros(0, dtype=np.uint32) for _ in range(len(off) - 1)]
        b, bm = self._emissions_dev(x_d, want_bm=len(obs_list) == 1)
        self._remember_emissions(obs_list, b, bm)
        r = self._trans().viterbi(b, off_d, len(off) - 1, int(off[-1]), want_path=True, want_score=False)
        return _dev.split_ragged(r["path"][0].cpu().numpy().astype(np.uint32), off)

    def _em_iteration(self, x_d, off, off_d, reduce_fn=None):
        """One E+M step over a batch; returns (model dict, total log-likelihood before the update)."""
        t = _dev.torch()
        n_utt, total = len(off) - 1, int(off[-1])
        acc = _dev.zeros((self._acc_size(),), t.float64)
        if total > 0:                                     # a rank whose shard is empty still enters the collective below
            b, bm = self._emissions_dev(x_d, want_bm=True)
            alpha, _ = self._trans().forward(b, off_d, n_utt, total, want_alpha=True)
            self._accumulate(x_d, b, bm, alpha, off_d, n_utt, total, acc)
        if reduce_fn is not None:
            reduce_fn(acc)
        loglik = float(acc[-2].item())
        return self._finalize(acc), loglik

    def train_batch(self, obs_list: Sequence, iters=1, eps=0.0001, verbose=True, reduce_fn=None):
        """Multi-utterance Baum-Welch: statistics summed over utterances (and over ranks by `reduce_fn`, an
        in-place all-reduce of the accumulator tensor).  With one utterance this is the reference's `train`."""
        x_d, off, off_d = self._pack(obs_list)
        history = []
        for k in range(iters):
            model, prob_old = self._em_iteration(x_d, off, off_d, reduce_fn)
            self._updatemodel(model)
            prob_new = self._total_loglik(x_d, off, off_d, reduce_fn)
            d = abs(prob_old - prob_new)
            history.append((prob_old, prob_new))
            if verbose:
                print("Iter: {:3}, L(lambda|O) = {:.6e}, L(lambda_new|O) = {:.6e}, eps = {:.6f}".format(k, prob_old, prob_new, d))
            if d < eps:
                break
        return history

    def _total_loglik(self, x_d, off, off_d, reduce_fn=None) -> float:
        t = _dev.torch()
        if int(off[-1]) > 0:
            b, bm = self._emissions_dev(x_d, want_bm=len(off) == 2)
            if len(off) == 2:
                self._remember_emissions([None], b, bm)      # train() ends on calc_prob(obs) of the new model (hmm.py:303)
            _, lp = self._trans().forward(b, off_d, len(off) - 1, int(off[-1]))
            s = lp.sum().reshape(1)
        else:
            s = _dev.zeros((1,), t.float64)
        if reduce_fn is not None:
            reduce_fn(s)
        return float(s.item())
