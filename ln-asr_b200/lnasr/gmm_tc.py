"""Tensor-core GMM emissions (csrc/gmm_tc.cu): packed 3xTF32 image of a bank of diagonal GMMs and its scorer.

Used by lnasr.bank.ModelBank for large models (BASELINE config 4: 2000 states x 16 mixtures x 39 dims)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi, _dev

__all__ = ["PackedGmmTC"]


class PackedGmmTC:
    def __init__(self, logw, mu, var, shift=None, fmt="tf32"):
        """logw [n_states, n_mix] (log), mu / var [n_states, n_mix, dim] (diagonal variances); shift [dim] optional
        feature centre (scores are mathematically unchanged; the fp32 cancellation error shrinks).
        fmt: 'tf32' (3xTF32, csrc/gmm_tc.cu) or 'f16' (two FP16 slices per operand, csrc/gmm_tc16.cu: half the MMAs; the
        features are centred on the model and scaled by powers of two so that they fit FP16's exponent range)."""
        t = _dev.torch()
        lib = _capi.lib()
        logw = np.asarray(logw, dtype=np.float64)
        mu = np.asarray(mu, dtype=np.float64)
        var = np.asarray(var, dtype=np.float64)
        self.n_states, self.n_mix = logw.shape
        self.dim = mu.shape[-1]
        if mu.shape != (self.n_states, self.n_mix, self.dim) or var.shape != mu.shape:
            raise ValueError("PackedGmmTC: mu / var must be [n_states, n_mix, dim] (diagonal covariance)")
        if fmt not in ("tf32", "f16"):
            raise ValueError("PackedGmmTC: fmt must be 'tf32' or 'f16'")
        self.fmt = fmt
        nbytes = C.c_size_t(0)
        lw, m, v = (_dev.to_dev(x, t.float64) for x in (logw, mu, var))
        if fmt == "f16":
            # centre on the model (weighted by nothing: the mid-range of the means) unless the caller gave a centre, then
            # scale every dimension by a power of two so that mean +- 4 sigma of every Gaussian stays within +-32
            finite = np.isfinite(logw)
            mm, vv = mu[finite], var[finite]
            if shift is None:
                shift = 0.5 * (mm.min(axis=0) + mm.max(axis=0)) if len(mm) else np.zeros(self.dim)
            # the staging kernel subtracts the shift in float32: use a float32-representable centre so that the weights
            # (float64 on the host) are built about exactly the same origin
            shift = np.asarray(shift, dtype=np.float32).astype(np.float64)
            reach = (np.abs(mm - shift) + 4.0 * np.sqrt(vv)).max(axis=0) if len(mm) else np.ones(self.dim)
            scale = np.exp2(np.floor(np.log2(32.0 / np.maximum(reach, 1e-30))))
            self.shift, self.scale = shift, scale
            _capi.check(lib.lnasr_gmm_tc16_packed_bytes(self.n_states, self.n_mix, self.dim, C.byref(nbytes)), "gmm_tc16_packed_bytes")
            self.packed = _dev.zeros((nbytes.value,), t.uint8)
            sh, sc = _dev.to_dev(shift, t.float64), _dev.to_dev(scale, t.float64)
            _capi.check(lib.lnasr_gmm_pack_tc16(_capi.ptr(lw), _capi.ptr(m), _capi.ptr(v), self.n_states, self.n_mix, self.dim,
                                                _capi.ptr(sh), _capi.ptr(sc), _capi.ptr(self.packed), _capi.stream_ptr()), "gmm_pack_tc16")
            return
        _capi.check(lib.lnasr_gmm_tc_packed_bytes(self.n_states, self.n_mix, self.dim, C.byref(nbytes)), "gmm_tc_packed_bytes")
        self.packed = _dev.zeros((nbytes.value,), t.uint8)
        sh = None if shift is None else _dev.to_dev(np.asarray(shift, dtype=np.float32).astype(np.float64), t.float64)
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
        fn = _capi.lib().lnasr_gmm_loglik_tc16 if self.fmt == "f16" else _capi.lib().lnasr_gmm_loglik_tc
        _capi.check(fn(_capi.ptr(feats), feats.stride(0), T, _capi.ptr(self.packed), self.n_states, self.n_mix, self.dim,
                       _capi.ptr(b), b.stride(0), _capi.ptr(ws), need.value, _capi.stream_ptr()), "gmm_loglik_" + self.fmt)
        return b
