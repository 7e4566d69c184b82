"""Device-resident Baum-Welch for a GMM-HMM over a ragged batch of utterances (additive API; BASELINE config 3).

`GMMHMM.train` / `train_batch` mirror the reference's loop (hmm.py:285-313): the model lives in numpy attributes and every
iteration makes a host round trip.  `EmTrainer` keeps the model ON THE DEVICE for the whole run; one iteration is

    lnasr_gmm_pack_exact -> lnasr_gmm_loglik_exact (b, bm) -> lnasr_forward_batch (alpha) || lnasr_backward_batch (beta, on a
    side stream) -> lnasr_em_statistics_ws (frame-parallel: gamma, xi, occupancies, S0/S1/S2c; any number of states)
    -> [all-reduce of the packed float64 accumulator over ranks] -> lnasr_mstep_finalize (in place)

with no host synchronisation, so iterations can be queued back to back and the only multi-GPU traffic is the
accumulator (lnasr_em_acc_doubles values, 15 KB for 3 states x 8 mixtures x 39 dims).  Every rank finalises the same
numbers, so replicas stay bit-identical.  Semantics = GMMHMM._estimate (gmmhmm.py:69-156) in
sufficient-statistic form with a diagonal covariance.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _capi, _dev

__all__ = ["EmTrainer"]


class EmTrainer:
    def __init__(self, model, min_std: Optional[float] = None):
        """model: lnasr.gmmhmm.GMMHMM (or a dict with A, pi, w, mu and var [n,m,d] / si [n,m,d,d])."""
        t = _dev.torch()
        get = (lambda k: model[k]) if isinstance(model, dict) else (lambda k: getattr(model, k))
        if isinstance(model, dict) and "var" in model:
            var = np.asarray(model["var"], dtype=np.float64)
        else:
            si = np.asarray(get("si"), dtype=np.float64)
            var = np.diagonal(si, axis1=2, axis2=3).copy()
            d_ = si.shape[-1]
            if np.any(si.reshape(-1, d_ * d_)[:, np.flatnonzero(~np.eye(d_, dtype=bool).reshape(-1))] != 0.0):
                raise _capi.LnasrError("EmTrainer trains diagonal-covariance models: `si` has non-zero off-diagonal entries "
                                       "(pass var = diag(si) explicitly to project, or score with ModelBank(emission='exact'))")
        self.n, self.m, self.d = var.shape
        self.min_std = float(min_std if min_std is not None else (0.01 if isinstance(model, dict) else model.min_std))
        self.logA = _dev.to_dev(np.asarray(get("A"), dtype=np.float64), t.float64)
        self.logpi = _dev.to_dev(np.asarray(get("pi"), dtype=np.float64).reshape(-1), t.float64)
        self.logw = _dev.to_dev(np.asarray(get("w"), dtype=np.float64), t.float64)
        self.mu = _dev.to_dev(np.asarray(get("mu"), dtype=np.float64), t.float64)
        self.var = _dev.to_dev(var, t.float64)
        lib = _capi.lib()
        self.packed = _dev.empty((int(lib.lnasr_gmm_exact_packed_doubles(self.n, self.m, self.d)),), t.float64)
        self.acc = _dev.zeros((int(lib.lnasr_em_acc_doubles(self.n, self.m, self.d)),), t.float64)
        # n > 32: the trellis passes read ELL lists of the finite transitions, re-packed from the device-resident log A at every
        # iteration.  EM never turns an impossible transition into a possible one, so the initial in / out degrees bound them.
        self._ell = None
        max_in = max_out = 0
        if self.n > 32:
            deg = _dev.zeros((4,), t.int32)
            _capi.check(lib.lnasr_trans_degree(_capi.ptr(self.logA), 1, self.n, _capi.ptr(deg), _capi.stream_ptr()), "trans_degree")
            d_in, d_out, _, _ = (int(v) for v in deg.cpu().tolist())
            max_in, max_out = max(1, d_in), max(1, d_out)
            self._ell = (_dev.empty((1, max_in, self.n), t.int32), _dev.empty((1, max_in, self.n), t.float64),
                         _dev.empty((1, max_out, self.n), t.int32), _dev.empty((1, max_out, self.n), t.float64))
        e = self._ell
        self._bank = _capi.HmmBank(n_models=1, N=self.n, logA=self.logA.data_ptr(), logpi=self.logpi.data_ptr(),
                                   max_in=max_in, max_out=max_out, pred_idx=e[0].data_ptr() if e else 0,
                                   pred_val=e[1].data_ptr() if e else 0, succ_idx=e[2].data_ptr() if e else 0,
                                   succ_val=e[3].data_ptr() if e else 0, band=0)
        self._x = None

    def load_batch(self, feats, frame_offsets):
        """feats [T_total, d] CUDA tensor (float32/float64); frame_offsets [n_utt+1] host.  Buffers are allocated once."""
        t = _dev.torch()
        off = np.asarray(frame_offsets, dtype=np.int64)
        self._x = _dev.to_dev(feats)
        self._off_d = _dev.to_dev(off, t.int64)
        self._n_utt, self._total = len(off) - 1, int(off[-1])
        T = self._total
        self._b = _dev.empty((T, self.n), t.float64)
        self._bm = _dev.empty((T, self.n * self.m), t.float64)
        self._alpha = _dev.empty((T, self.n), t.float64)
        self._lp = _dev.empty((self._n_utt, 1), t.float64)
        self._beta = _dev.empty((T, self.n), t.float64)
        self._ws_bytes = int(_capi.lib().lnasr_em_statistics_workspace_bytes(T, self.n))
        self._ws = _dev.empty((self._ws_bytes,), t.uint8)
        self._side = t.cuda.Stream()
        return self

    def iterate(self, reduce_fn=None):
        """One EM iteration, fully asynchronous.  Returns a 1-element device tensor: the total log-likelihood of the
        batch (all ranks) under the model BEFORE the update (hmm.py:301)."""
        if self._x is None:
            raise _capi.LnasrError("EmTrainer.iterate: call load_batch first")
        lib, ptr, st = _capi.lib(), _capi.ptr, _capi.stream_ptr()
        x, n, m, d, T = self._x, self.n, self.m, self.d, self._total
        _capi.check(lib.lnasr_gmm_pack_exact(ptr(self.logw), ptr(self.mu), ptr(self.var), n, m, d, ptr(self.packed), st), "gmm_pack_exact")
        _capi.check(lib.lnasr_gmm_loglik_exact(ptr(x), _dev.np_dtype_code(x.dtype), x.stride(0), T, ptr(self.packed), n, m, d,
                                               ptr(self._b), _capi.DT_F64, n, ptr(self._bm), st), "gmm_loglik_exact")
        if self._ell is not None:
            e = self._ell
            _capi.check(lib.lnasr_trans_pack(ptr(self.logA), 1, n, self._bank.max_in, self._bank.max_out, ptr(e[0]), ptr(e[1]),
                                             ptr(e[2]), ptr(e[3]), st), "trans_pack")
        # the two recursions are independent: backward on a side stream, forward here; everything else is frame-parallel
        t = _dev.torch()
        main = t.cuda.current_stream()
        self._side.wait_stream(main)
        _capi.check(lib.lnasr_backward_batch(ptr(self._b), _capi.DT_F64, n, ptr(self._off_d), self._n_utt, T, C.byref(self._bank),
                                             ptr(self._beta), C.c_void_p(self._side.cuda_stream)), "backward_batch")
        _capi.check(lib.lnasr_forward_batch(ptr(self._b), _capi.DT_F64, n, ptr(self._off_d), self._n_utt, T,
                                            C.byref(self._bank), ptr(self._alpha), ptr(self._lp), st), "forward_batch")
        self.acc.zero_()
        main.wait_stream(self._side)
        _capi.check(lib.lnasr_em_statistics_ws(ptr(x), _dev.np_dtype_code(x.dtype), x.stride(0), ptr(self._b), ptr(self._bm),
                                               ptr(self._alpha), ptr(self._beta), ptr(self._off_d), self._n_utt, T, ptr(self.logA),
                                               ptr(self.logw), ptr(self.mu), n, m, d, ptr(self.acc), None, ptr(self._ws),
                                               self._ws_bytes, st), "em_statistics_ws")
        if reduce_fn is not None:
            reduce_fn(self.acc)                       # SUM over ranks of S0 | S1 | S2c | xi_sum | pi_occ | loglik | n_utt
        loglik = self.acc[-2:-1].clone()
        _capi.check(lib.lnasr_mstep_finalize(ptr(self.acc), n, m, d, self.min_std, ptr(self.logA), ptr(self.logpi),
                                             ptr(self.logw), ptr(self.mu), ptr(self.var), st), "mstep_finalize")
        return loglik

    def export(self) -> dict:
        """Current model as numpy arrays: A, pi, w (log), mu, var [n,m,d] and si = diag(var) [n,m,d,d]."""
        var = self.var.cpu().numpy()
        si = np.zeros((self.n, self.m, self.d, self.d))
        idx = np.arange(self.d)
        si[:, :, idx, idx] = var
        return {"A": self.logA.cpu().numpy(), "pi": self.logpi.cpu().numpy(), "w": self.logw.cpu().numpy(),
                "mu": self.mu.cpu().numpy(), "var": var, "si": si}
