"""Banks of GMM-HMMs scored together, and the PCM -> word pipeline (additive API, SURVEY.md 8(b) "new-build
additions").

The reference scores one utterance against one model per call.  Isolated-word recognition (BASELINE config 2)
and large-vocabulary acoustic scoring (config 4) are batches of utterances against K models that share
(n, m, d): here all K*n*m Gaussians are one packed image, emissions are one [frames, K*n] matrix, and the
trellis kernels run every (utterance, model) pair in one launch.  Word decision = argmax over models of the
Viterbi score max_j v[T-1, j], lowest index on ties.
"""
from __future__ import annotations

from typing import NamedTuple, Optional, Sequence

import numpy as np

from . import _capi, _dev
from ._trellis import TransitionBank
from .gmm import PackedGmm

__all__ = ["ModelBank", "AcousticPipeline", "BankResult"]


class BankResult(NamedTuple):
    words: "object"                 # [n_utt] int32 CUDA tensor: index of the best model
    best: "object"                  # [n_utt] float64: its Viterbi score
    scores: "object"                # [n_utt, K] float64 Viterbi scores max_j v[T-1,j]
    paths: Optional["object"]       # [K, total_frames] int32 state paths per model, or None
    logprob: Optional["object"]     # [n_utt, K] float64 forward log-probabilities, or None


class ModelBank:
    """K GMM-HMMs with common (n, m, d)."""

    def __init__(self, models: Sequence, emission: str = "auto"):
        """models: GMMHMM objects or dicts with keys A, pi, w, mu and si ([n,m,d,d]) or var ([n,m,d]).
        emission: 'exact' (float64 SIMT), 'tc' (3xTF32 tensor cores) or 'auto' (tc when K*n*m >= 2048)."""
        def get(mdl, k):
            return mdl[k] if isinstance(mdl, dict) else getattr(mdl, k)

        def cov(mdl):
            if isinstance(mdl, dict) and "var" in mdl:
                return np.asarray(mdl["var"], dtype=np.float64)
            si = np.asarray(get(mdl, "si"), dtype=np.float64)
            return si
        A = np.stack([np.asarray(get(m, "A"), dtype=np.float64) for m in models])
        pi = np.stack([np.asarray(get(m, "pi"), dtype=np.float64).reshape(-1) for m in models])
        w = np.concatenate([np.asarray(get(m, "w"), dtype=np.float64) for m in models], axis=0)
        mu = np.concatenate([np.asarray(get(m, "mu"), dtype=np.float64) for m in models], axis=0)
        cv = np.concatenate([cov(m) for m in models], axis=0)
        self.K, self.n = A.shape[0], A.shape[1]
        self.m, self.d = w.shape[1], mu.shape[2]
        self.trans = TransitionBank(A, pi)
        self.gmm = PackedGmm(w, mu, cv)
        if emission == "auto":
            emission = "tc" if (self.K * self.n * self.m >= 2048 and not self.gmm.full) else "exact"
        if emission not in ("exact", "tc"):
            raise ValueError("emission must be 'exact', 'tc' or 'auto'")
        self.emission = emission
        self._tc = None
        if emission == "tc":
            from .gmm_tc import PackedGmmTC
            self._tc = PackedGmmTC(w, mu, cv if cv.ndim == 3 else np.diagonal(cv, axis1=2, axis2=3))

    def emissions(self, feats, b_out=None):
        """feats [T, d] float32/float64 device -> b [T, K*n] (float64 exact / float32 tc)."""
        if self._tc is not None:
            return self._tc.loglik(feats, b_out=b_out)
        return self.gmm.loglik(feats, b_out=b_out)[0]

    def score(self, feats, f_off_d, n_utt: int, total: int, want_paths=False, want_logprob=False,
              workspace=None) -> BankResult:
        t = _dev.torch()
        b = self.emissions(feats)
        r = self.trans.viterbi(b, f_off_d, n_utt, total, want_path=want_paths, want_score=True, workspace=workspace)
        words = _dev.empty((n_utt,), t.int32)
        best = _dev.empty((n_utt,), t.float64)
        _capi.check(_capi.lib().lnasr_argmax_rows(_capi.ptr(r["score"]), n_utt, self.K, _capi.ptr(words), _capi.ptr(best),
                                                 _capi.stream_ptr()), "argmax_rows")
        lp = self.trans.forward(b, f_off_d, n_utt, total)[1] if want_logprob else None
        return BankResult(words, best, r["score"], r["path"], lp)


class AcousticPipeline:
    """PCM -> MFCC -> GMM emissions -> Viterbi -> word, for a ragged batch (BASELINE north star (1)-(3))."""

    def __init__(self, mfcc, bank: ModelBank):
        self.mfcc = mfcc
        self.bank = bank

    def recognize_batch(self, pcm, sample_offsets, want_paths=False, want_logprob=False):
        """pcm: concatenated int16/float32 samples (numpy, pinned CPU tensor or CUDA tensor); sample_offsets [n+1] host.
        Returns (BankResult, MfccBatch); everything stays on the device."""
        mb = self.mfcc.calc_mfcc_batch(pcm, sample_offsets=sample_offsets)
        n_utt = len(mb.frame_offsets_host) - 1
        total = int(mb.frame_offsets_host[-1])
        res = self.bank.score(mb.feats, mb.frame_offsets, n_utt, total, want_paths=want_paths, want_logprob=want_logprob)
        return res, mb

    def recognize(self, signals: Sequence[np.ndarray]):
        """Convenience: list of utterances -> numpy array of word indices."""
        arrs = [np.asarray(s) for s in signals]
        i16 = all(a.dtype == np.int16 for a in arrs)
        pcm, off = _dev.pack_ragged(arrs, np.int16 if i16 else np.float32)
        res, _ = self.recognize_batch(pcm, off)
        return res.words.cpu().numpy()
