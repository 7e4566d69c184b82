"""Device-side trellis engine shared by lnasr.hmm.HMM, lnasr.gmmhmm.GMMHMM and lnasr.bank.ModelBank:
thin wrappers over lnasr_viterbi_batch / lnasr_forward_batch / lnasr_backward_batch (csrc/trellis.cu)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi, _dev

SMALL_N = 32     # csrc/trellis.cu: N <= 32 runs the warp-shuffle kernel on dense log A


class TransitionBank:
    """log A [K,N,N] and log pi [K,N] of K models with the same N on the device, plus -- for N > 32 -- the ELL
    predecessor / successor lists of the finite transitions (lnasr_trans_pack)."""

    def __init__(self, logA, logpi):
        t = _dev.torch()
        lib = _capi.lib()
        logA = np.asarray(logA, dtype=np.float64)
        logpi = np.asarray(logpi, dtype=np.float64)
        if logA.ndim == 2:
            logA, logpi = logA[None], logpi[None]
        self.K, self.N = logA.shape[0], logA.shape[1]
        if logA.shape != (self.K, self.N, self.N) or logpi.shape != (self.K, self.N):
            raise ValueError(f"transition bank: A {logA.shape} / pi {logpi.shape} do not agree")
        self.logA = _dev.to_dev(logA, t.float64)
        self.logpi = _dev.to_dev(logpi, t.float64)
        self.max_in = self.max_out = 0
        self.band = 0
        self.pred_idx = self.pred_val = self.succ_idx = self.succ_val = None
        if self.N > SMALL_N:
            deg = _dev.zeros((4,), t.int32)
            _capi.check(lib.lnasr_trans_degree(_capi.ptr(self.logA), self.K, self.N, _capi.ptr(deg), _capi.stream_ptr()),
                        "trans_degree")
            d_in, d_out, fwd, back = (int(v) for v in deg.cpu().tolist())
            self.max_in, self.max_out = max(1, d_in), max(1, d_out)
            # left-to-right band (self loop, next, skip): the decode path keeps the trellis column in registers
            self.band = 1 + fwd if (back == 0 and fwd <= 2) else 0
            self.pred_idx = _dev.empty((self.K, self.max_in, self.N), t.int32)
            self.pred_val = _dev.empty((self.K, self.max_in, self.N), t.float64)
            self.succ_idx = _dev.empty((self.K, self.max_out, self.N), t.int32)
            self.succ_val = _dev.empty((self.K, self.max_out, self.N), t.float64)
            _capi.check(lib.lnasr_trans_pack(_capi.ptr(self.logA), self.K, self.N, self.max_in, self.max_out,
                                             _capi.ptr(self.pred_idx), _capi.ptr(self.pred_val),
                                             _capi.ptr(self.succ_idx), _capi.ptr(self.succ_val), _capi.stream_ptr()),
                        "trans_pack")
        self.struct = _capi.HmmBank(
            n_models=self.K, N=self.N, logA=self.logA.data_ptr(), logpi=self.logpi.data_ptr(),
            max_in=self.max_in, max_out=self.max_out,
            pred_idx=0 if self.pred_idx is None else self.pred_idx.data_ptr(),
            pred_val=0 if self.pred_val is None else self.pred_val.data_ptr(),
            succ_idx=0 if self.succ_idx is None else self.succ_idx.data_ptr(),
            succ_val=0 if self.succ_val is None else self.succ_val.data_ptr(), band=self.band)

    @property
    def ncol(self):
        return self.K * self.N

    # ------------------------------------------------------------------
    def viterbi(self, b, f_off_d, n_utt, total, want_path=True, want_score=True, want_v=False, want_bt=False,
                workspace=None):
        """b [total, K*N] (float32/float64).  Returns dict(path [K,total] int32 view of uint32, score [n_utt,K], v, bt)."""
        t = _dev.torch()
        lib = _capi.lib()
        path = _dev.empty((self.K, total), t.int32) if want_path else None       # uint32 values (hmm.py:168), < 2^31
        score = _dev.empty((n_utt, self.K), t.float64) if want_score else None
        v = _dev.empty((total, self.ncol), t.float64) if want_v else None
        bt = _dev.empty((total, self.ncol), t.int32) if want_bt else None
        ws_bytes = 0
        ws = None
        if want_path:
            ws_bytes = int(lib.lnasr_viterbi_workspace_bytes(total, n_utt, self.K, self.N))
            ws = workspace if (workspace is not None and workspace.numel() >= ws_bytes) else _dev.empty((ws_bytes,), t.uint8)
        if total and n_utt:
            _capi.check(lib.lnasr_viterbi_batch(
                _capi.ptr(b), _dev.np_dtype_code(b.dtype), b.stride(0), _capi.ptr(f_off_d), n_utt, total,
                C.byref(self.struct), _capi.ptr(path), _capi.ptr(score), _capi.ptr(v), _capi.ptr(bt),
                _capi.ptr(ws), ws_bytes, _capi.stream_ptr()), "viterbi_batch")
        return {"path": path, "score": score, "v": v, "bt": bt}

    def forward(self, b, f_off_d, n_utt, total, want_alpha=False):
        t = _dev.torch()
        alpha = _dev.empty((total, self.ncol), t.float64) if want_alpha else None
        logprob = _dev.empty((n_utt, self.K), t.float64)
        if total and n_utt:
            _capi.check(_capi.lib().lnasr_forward_batch(
                _capi.ptr(b), _dev.np_dtype_code(b.dtype), b.stride(0), _capi.ptr(f_off_d), n_utt, total,
                C.byref(self.struct), _capi.ptr(alpha), _capi.ptr(logprob), _capi.stream_ptr()), "forward_batch")
        return alpha, logprob

    def backward(self, b, f_off_d, n_utt, total):
        t = _dev.torch()
        beta = _dev.empty((total, self.ncol), t.float64)
        if total and n_utt:
            _capi.check(_capi.lib().lnasr_backward_batch(
                _capi.ptr(b), _dev.np_dtype_code(b.dtype), b.stride(0), _capi.ptr(f_off_d), n_utt, total,
                C.byref(self.struct), _capi.ptr(beta), _capi.stream_ptr()), "backward_batch")
        return beta
