"""Gaussian / Gaussian-mixture densities -- API mirror of the reference's lnasr/gmm.py (gmm.py:16-100),
evaluated by csrc/gmm_exact.cu (float64) on a B200.

Log densities are natural logs; means and (co)variances are linear, as in the reference (gmm.py:7).
Covariance matrices may be diagonal (the build target) or full: a full Sigma is
whitened once on the host (Cholesky of the D x D parameter, a model-pack step) and the data term is
evaluated by the full-covariance kernel.
"""
from __future__ import annotations

import numpy as np

from . import _capi, _dev

__all__ = [
    "gaussian_distribution", "gaussian_distribution_log",
    "gaussian_multi_distribution", "gaussian_multi_distribution_log",
    "gaussian_mixture_distribution", "gaussian_mixture_distribution_log",
]


def _is_diag(S: np.ndarray) -> bool:
    S = np.asarray(S)
    d = S.shape[-1]
    off = S.reshape(-1, d, d) * (1.0 - np.eye(d))
    return not np.any(off)


class PackedGmm:
    """Device image of a bank of GMMs ([n_states, n_mix] Gaussians of dimension dim) for the float64 kernels."""

    def __init__(self, logw, mu, cov):
        t = _dev.torch()
        lib = _capi.lib()
        logw = np.asarray(logw, dtype=np.float64)
        mu = np.asarray(mu, dtype=np.float64)
        cov = np.asarray(cov, dtype=np.float64)
        self.n_states, self.n_mix = logw.shape
        self.dim = mu.shape[-1]
        if mu.shape != (self.n_states, self.n_mix, self.dim):
            raise ValueError(f"mu has shape {mu.shape}, expected {(self.n_states, self.n_mix, self.dim)}")
        if cov.ndim == 4:
            if cov.shape != (self.n_states, self.n_mix, self.dim, self.dim):
                raise ValueError(f"si has shape {cov.shape}")
            self.full = not _is_diag(cov)
            var = np.diagonal(cov, axis1=2, axis2=3) if not self.full else None
        else:
            if cov.shape != mu.shape:
                raise ValueError(f"var has shape {cov.shape}, expected {mu.shape}")
            self.full, var = False, cov
        self.logw_d = _dev.to_dev(logw, t.float64)
        self.mu_d = _dev.to_dev(mu, t.float64)
        ng = self.n_states * self.n_mix
        if not self.full:
            self.var_d = _dev.to_dev(np.ascontiguousarray(var), t.float64)
            self.packed = _dev.empty((int(lib.lnasr_gmm_exact_packed_doubles(self.n_states, self.n_mix, self.dim)),), t.float64)
            _capi.check(lib.lnasr_gmm_pack_exact(_capi.ptr(self.logw_d), _capi.ptr(self.mu_d), _capi.ptr(self.var_d),
                                                 self.n_states, self.n_mix, self.dim, _capi.ptr(self.packed),
                                                 _capi.stream_ptr()), "gmm_pack_exact")
        else:
            # model-pack on the host: Sigma = L L^T ; quad = |L^-1 (x-mu)|^2 ; ln det = 2 sum ln L_ii
            D = self.dim
            Lc = np.linalg.cholesky(cov.reshape(ng, D, D))
            Linv = np.linalg.inv(Lc)
            c = -0.5 * (D * np.log(2.0 * np.pi) + 2.0 * np.log(np.diagonal(Lc, axis1=1, axis2=2)).sum(axis=1))
            self.var_d = None
            self.linv_d = _dev.to_dev(np.ascontiguousarray(Linv), t.float64)
            self.nu_d = _dev.to_dev(np.einsum("gde,ge->gd", Linv, mu.reshape(ng, D)), t.float64)
            self.c_d = _dev.to_dev(c, t.float64)

    def loglik(self, feats, want_bm=False, out_dtype=None, b_out=None):
        """feats [T, >=dim] CUDA tensor (float32/float64) -> b [T, n_states] (+ bm [T, n_states*n_mix] float64)."""
        t = _dev.torch()
        lib = _capi.lib()
        T = feats.shape[0]
        out_dtype = out_dtype or t.float64
        b = b_out if b_out is not None else _dev.empty((T, self.n_states), out_dtype)
        bm = _dev.empty((T, self.n_states * self.n_mix), t.float64) if want_bm else None
        if T == 0:
            return b, bm
        if not self.full:
            _capi.check(lib.lnasr_gmm_loglik_exact(
                _capi.ptr(feats), _dev.np_dtype_code(feats.dtype), feats.stride(0), T, _capi.ptr(self.packed),
                self.n_states, self.n_mix, self.dim, _capi.ptr(b), _dev.np_dtype_code(b.dtype), b.stride(0),
                _capi.ptr(bm), _capi.stream_ptr()), "gmm_loglik_exact")
        else:
            _capi.check(lib.lnasr_gmm_loglik_full(
                _capi.ptr(feats), _dev.np_dtype_code(feats.dtype), feats.stride(0), T, _capi.ptr(self.linv_d),
                _capi.ptr(self.nu_d), _capi.ptr(self.c_d), _capi.ptr(self.logw_d), self.n_states, self.n_mix, self.dim,
                _capi.ptr(b), _dev.np_dtype_code(b.dtype), b.stride(0), _capi.ptr(bm), _capi.stream_ptr()),
                "gmm_loglik_full")
        return b, bm


def _single_logpdf(x, mu, Sigma) -> np.ndarray:
    t = _dev.torch()
    x = np.asarray(x, dtype=np.float64)
    if x.ndim != 2:
        raise ValueError("x must be [L, D]")
    D = x.shape[1]
    mu = np.asarray(mu, dtype=np.float64).reshape(1, 1, D)
    Sigma = np.asarray(Sigma, dtype=np.float64).reshape(1, 1, D, D)
    g = PackedGmm(np.zeros((1, 1)), mu, Sigma)
    b, _ = g.loglik(_dev.to_dev(x, t.float64))
    return b[:, 0].cpu().numpy()


def gaussian_distribution_log(x: np.ndarray, mu, sigma2) -> np.ndarray:
    """Univariate Gaussian log-density, elementwise over x (gmm.py:28-30)."""
    x = np.asarray(x, dtype=np.float64)
    out = _single_logpdf(x.reshape(-1, 1), [mu], [[sigma2]])
    return out.reshape(x.shape)


def gaussian_distribution(x: np.ndarray, mu, sigma2) -> np.ndarray:
    """Univariate Gaussian density (gmm.py:16-26)."""
    return np.exp(gaussian_distribution_log(x, mu, sigma2))


def gaussian_multi_distribution_log(x: np.ndarray, mu: np.ndarray, Sigma: np.ndarray) -> np.ndarray:
    """Multivariate Gaussian log-density of the rows of x [L,D] -> [L] (gmm.py:59-68)."""
    return _single_logpdf(x, mu, Sigma)


def gaussian_multi_distribution(x: np.ndarray, mu: np.ndarray, Sigma: np.ndarray) -> np.ndarray:
    """Multivariate Gaussian density -> [L] (gmm.py:32-57)."""
    return np.exp(_single_logpdf(x, mu, Sigma))


def gaussian_mixture_distribution_log(c: np.ndarray, x: np.ndarray, mu: np.ndarray, Sigma: np.ndarray):
    """log sum_k exp(c_k + log N(x; mu_k, Sigma_k)); c holds LOG mixture weights -> [L] (gmm.py:90-100)."""
    t = _dev.torch()
    x = np.asarray(x, dtype=np.float64)
    c = np.asarray(c, dtype=np.float64).reshape(1, -1)
    M, D = c.shape[1], x.shape[1]
    g = PackedGmm(c, np.asarray(mu, dtype=np.float64).reshape(1, M, D),
                  np.asarray(Sigma, dtype=np.float64).reshape(1, M, D, D))
    b, _ = g.loglik(_dev.to_dev(x, t.float64))
    return b[:, 0].cpu().numpy()


def gaussian_mixture_distribution(c: np.ndarray, x: np.ndarray, mu: np.ndarray, Sigma: np.ndarray):
    """sum_k c_k N(x; mu_k, Sigma_k) with LINEAR weights c -> [1, L] as the reference's np.dot (gmm.py:70-88)."""
    c = np.asarray(c, dtype=np.float64).ravel()
    with np.errstate(divide="ignore"):
        out = np.exp(gaussian_mixture_distribution_log(np.log(c), x, mu, Sigma))
    return out.reshape(1, -1)
