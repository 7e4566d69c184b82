"""Tensor-core GMM emissions (csrc/gmm_tc.cu): packed 3xTF32 image of a bank of diagonal GMMs and its scorer.

Used by lnasr.bank.ModelBank for large models (BASELINE config 4: 2000 states x 16 mixtures x 39 dims)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi, _dev

__all__ = ["PackedGmmTC"]


class PackedGmmTC:
    def __init__(self, logw, mu, var, shift=None):
        """logw [n_states, n_mix] (log), mu / var [n_states, n_mix, dim] (diagonal variances); shift [dim] optional
        feature centre (scores are mathematically unchanged; the fp32 cancellation error shrinks)."""
        t = _dev.torch()
        lib = _capi.lib()
        logw = np.asarray(logw, dtype=np.float64)
        mu = np.asarray(mu, dtype=np.float64)
        var = np.asarray(var, dtype=np.float64)
        self.n_states, self.n_mix = logw.shape
        self.dim = mu.shape[-1]
        if mu.shape != (self.n_states, self.n_mix, self.dim) or var.shape != mu.shape:
            raise ValueError("PackedGmmTC: mu / var must be [n_states, n_mix, dim] (diagonal covariance)")
        nbytes = C.c_size_t(0)
        _capi.check(lib.lnasr_gmm_tc_packed_bytes(self.n_states, self.n_mix, self.dim, C.byref(nbytes)), "gmm_tc_packed_bytes")
        self.packed = _dev.zeros((nbytes.value,), t.uint8)
        lw, m, v = (_dev.to_dev(x, t.float64) for x in (logw, mu, var))
        sh = None if shift is None else _dev.to_dev(np.asarray(shift, dtype=np.float64), t.float64)
        _capi.check(lib.lnasr_gmm_pack_tc(_capi.ptr(lw), _capi.ptr(m), _capi.ptr(v), self.n_states, self.n_mix, self.dim,
                                          _capi.ptr(sh), _capi.ptr(self.packed), _capi.stream_ptr()), "gmm_pack_tc")

    def loglik(self, feats, b_out=None):
        """feats [T, >=dim] float32 CUDA tensor -> b [T, n_states] float32."""
        t = _dev.torch()
        if feats.dtype != t.float32:
            raise _capi.LnasrError("the tensor-core GMM path takes float32 features")
        T = feats.shape[0]
        b = b_out if b_out is not None else _dev.empty((T, self.n_states), t.float32)
        need = C.c_size_t(0)
        _capi.check(_capi.lib().lnasr_gmm_tc_workspace_bytes(T, C.byref(need)), "gmm_tc_workspace_bytes")
        ws = _dev.empty((need.value,), t.uint8)
        _capi.check(_capi.lib().lnasr_gmm_loglik_tc(_capi.ptr(feats), feats.stride(0), T, _capi.ptr(self.packed), self.n_states,
                                                   self.n_mix, self.dim, _capi.ptr(b), b.stride(0), _capi.ptr(ws), need.value,
                                                   _capi.stream_ptr()), "gmm_loglik_tc")
        return b
