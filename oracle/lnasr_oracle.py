"""CPU oracle for the ln-asr acoustic hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a float64 numpy restatement of the reference algorithms
(yehuohan/ln-asr, `lnasr/{utils,mfcc,gmm,hmm,gmmhmm}.py`).  It is the checker the
CUDA path is compared with; it is never on the product path.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it.  The product package (`ln-asr_b200/lnasr`) must not.

Parity pinning: every function below is checked (tests/test_oracle_golden.py)
against fixtures under tests/golden/ that were produced by running the
UNMODIFIED reference source in the authoring container
(tests/golden/make_golden.py, which imports /root/reference with import shims).
The reference's own tests hold no numeric assertions, so those
generated fixtures plus the ice-cream HMM of test/hmm-test.py:10-19 are the pin.

Each function cites the reference file:line it restates.  All citations are
relative to /root/reference/.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "lse", "split_frames", "create_hamming", "num_frames",
    "MfccParams", "high_pass", "mel", "mel_inv", "mel_filter_bank", "mfcc",
    "gauss_logpdf_full", "gauss_logpdf_diag", "map_b",
    "hmm_map_b", "forward", "backward", "viterbi", "baumwelch",
    "estimate_hmm", "estimate_gmmhmm", "estimate_gmmhmm_diag",
    "calc_prob", "decode", "train_gmmhmm", "em_accumulate_diag", "em_finalize_diag",
    "em_step_multi_diag", "random_gmmhmm", "score_words", "score_words_viterbi", "vad_ltsd",
]

NEG_INF = -np.inf


# --------------------------------------------------------------------------- utils
def lse(seq: np.ndarray, axis=None) -> np.ndarray:
    """log-sum-exp with the reference's exact fold order (lnasr/utils.py:140-179).

    The reference reduces an axis by folding `np.logaddexp` left to right
    (utils.py:176-179); a tuple of (ascending) axes is reduced last axis first
    (utils.py:166-167); a zero-length axis yields -inf over the non-zero dims
    (utils.py:169-170).  We keep that order so results are bit-identical.
    """
    seq = np.asarray(seq)
    if axis is None:
        return lse(seq.ravel(), 0)
    if isinstance(axis, tuple):
        out = seq
        for ax in reversed(axis):
            out = lse(out, ax)
        return out
    n = seq.shape[axis]
    if n == 0:
        return np.full(tuple(d for d in seq.shape if d != 0), NEG_INF)
    moved = np.moveaxis(seq, axis, 0)
    acc = moved[0]
    for k in range(1, n):
        acc = np.logaddexp(acc, moved[k])
    return acc


def num_frames(n_samples: int, L: int = 400, S: int = 160) -> int:
    """Frame count rule of lnasr/utils.py:118 -- note the abs()."""
    return int(np.ceil(np.abs(n_samples - (L - S)) / S))


def split_frames(signal: np.ndarray, L: int, S: int) -> np.ndarray:
    """Framing with zero padding of the tail (lnasr/utils.py:106-130)."""
    signal = np.asarray(signal)
    n = num_frames(len(signal), L, S)
    need = n * S + (L - S)
    if need > len(signal):
        signal = np.concatenate([signal, np.zeros(need - len(signal))])
    starts = S * np.arange(n)
    idx = starts[:, None] + np.arange(L)[None, :]
    return signal[idx]


def create_hamming(N: int) -> np.ndarray:
    """Hamming window 0.54-0.46cos(2 pi n/(N-1)) (lnasr/utils.py:132-138)."""
    return 0.54 - 0.46 * np.cos(2 * np.pi * np.arange(N) / (N - 1))


# --------------------------------------------------------------------------- MFCC
class MfccParams:
    """Derived constants of MFCC.__init__ (lnasr/mfcc.py:23-48)."""

    def __init__(self, fs=16000, frame_T=25e-3, frame_stride=10e-3, hpf_alpha=0.97,
                 fft_N=512, mfbank_num=40, cepstrum_num=12):
        self.fs = fs
        self.frame_L = int(fs * frame_T)
        self.frame_step = int(fs * frame_stride)
        self.hpf_alpha = hpf_alpha
        self.fft_N = fft_N
        self.fft_size = fft_N // 2 + 1
        self.mfbank_num = mfbank_num
        self.cepstrum_num = cepstrum_num


def high_pass(signal: np.ndarray, alpha: float = 0.97) -> np.ndarray:
    """y[0]=x[0]; y[t]=x[t]-alpha*x[t-1] (lnasr/mfcc.py:50-64)."""
    signal = np.asarray(signal)
    out = np.empty(len(signal), dtype=np.float64)
    out[0] = signal[0]
    out[1:] = signal[1:] - alpha * signal[:-1]
    return out


def mel(hz):
    """2595*ln(1+hz/700) -- natural log (lnasr/mfcc.py:66-72)."""
    return 2595 * np.log(1 + np.asarray(hz, dtype=np.float64) / 700)


def mel_inv(m):
    """Inverse of `mel` (lnasr/mfcc.py:74-76)."""
    return 700 * (np.exp(np.asarray(m, dtype=np.float64) / 2595) - 1)


def mel_filter_bank(p: MfccParams = None) -> np.ndarray:
    """Triangular mel filters on FFT bins (lnasr/mfcc.py:78-106).  -> [M, fft_N/2+1]."""
    p = p or MfccParams()
    M, N = p.mfbank_num, p.fft_N
    edges_hz = mel_inv(np.linspace(mel(0), mel(p.fs / 2), M + 2))
    bins = np.floor((N / p.fs) * edges_hz)
    fb = np.zeros((M, N // 2 + 1))
    for m in range(M):
        lo, mi, hi = int(bins[m]), int(bins[m + 1]), int(bins[m + 2])
        fb[m, lo:mi] = (np.arange(lo, mi) - lo) / (mi - lo)
        fb[m, mi:hi] = (hi - np.arange(mi, hi)) / (hi - mi)
    return fb


def _dct2_ortho(x: np.ndarray) -> np.ndarray:
    """DCT-II, norm='ortho', along axis 1 (scipy.fftpack.dct as used at lnasr/mfcc.py:152).

    y_k = f_k * 2 * sum_n x_n cos(pi k (2n+1) / (2N)),  f_0 = sqrt(1/(4N)), f_k = sqrt(1/(2N)).
    """
    from scipy.fftpack import dct  # same third-party routine the reference calls
    return dct(x, type=2, axis=1, norm="ortho")


def dct2_ortho_matrix(N: int) -> np.ndarray:
    """Explicit DCT-II ortho matrix [k, n]; used to cross-check `_dct2_ortho`."""
    n = np.arange(N)
    k = np.arange(N)[:, None]
    mat = 2.0 * np.cos(np.pi * k * (2 * n + 1) / (2 * N))
    scale = np.full(N, np.sqrt(1.0 / (2 * N)))
    scale[0] = np.sqrt(1.0 / (4 * N))
    return mat * scale[:, None]


def mfcc(signal: np.ndarray, p: MfccParams = None):
    """MFCC.calc_mfcc (lnasr/mfcc.py:108-175) -> (power [F,257], cepstrum [F,40], features [F,39]).

    Quirks kept on purpose:
      * zero mel energy is replaced by float eps before the log (mfcc.py:149);
      * dB scale 20*log10, then column mean over ALL frames + 1e-8 removed (mfcc.py:150-151);
      * columns 1..12 of the DCT output are a view and are mean-normalised in place,
        so the returned cepstrum has those columns altered too (mfcc.py:155-156);
      * 13th coefficient is ln(sum_k power) (mfcc.py:158-160);
      * delta row 0 is a copy of static row 1, delta-delta row 0 a copy of delta row 1
        (mfcc.py:162-172).
    """
    p = p or MfccParams()
    frames = split_frames(high_pass(signal, p.hpf_alpha), p.frame_L, p.frame_step)
    frames = frames * create_hamming(p.frame_L)
    mag = np.abs(np.fft.rfft(frames, p.fft_N))
    power = (1.0 / p.fft_N) * mag ** 2
    logmel = np.dot(power, mel_filter_bank(p).T)
    logmel = np.where(logmel == 0, np.finfo(float).eps, logmel)
    logmel = 20 * np.log10(logmel)
    logmel = logmel - (np.mean(logmel, axis=0) + 1e-8)
    ceps = _dct2_ortho(logmel)
    static = ceps[:, 1:1 + p.cepstrum_num]          # view into ceps, as in the reference
    static -= (np.mean(static, axis=0) + 1e-8)
    with np.errstate(divide="ignore"):
        energy = np.log(np.sum(power, axis=1))[:, None]
    s13 = np.concatenate([static, energy], axis=1)
    d13 = np.concatenate([s13[1:2], s13[1:] - s13[:-1]], axis=0)
    dd13 = np.concatenate([d13[1:2], d13[1:] - d13[:-1]], axis=0)
    return power, ceps, np.concatenate([s13, d13, dd13], axis=1)


def vad_ltsd(data, winsize=2048, stepsize=1024, order=6, threshold=-6, alpha=None):
    """VadLtsd.detect (third/vadlstd/VadLstd.py:44-74): LTSD in dB per frame; frames within `order` of either end stay 0."""
    pad = np.append(np.zeros(stepsize), np.asarray(data, dtype=np.float64))
    frames = split_frames(pad, winsize, stepsize) * create_hamming(winsize)
    F = frames.shape[0]
    amp = np.abs(np.fft.rfft(frames, winsize))
    noise = np.average(amp[0:2], axis=0) ** 2
    ltsd = np.zeros(F)
    for k in range(order, F - order):
        ltse = amp[k - order:k + order].max(axis=0)
        ltsd[k] = 10 * np.log10(np.sum(ltse ** 2 / noise) / winsize)
        if alpha is not None and ltsd[k] < threshold:
            noise = alpha * noise + (1 - alpha) * (np.sum(ltse) / winsize)
    return ltsd


# --------------------------------------------------------------------------- GMM
def gauss_logpdf_full(x: np.ndarray, mu: np.ndarray, Sigma: np.ndarray) -> np.ndarray:
    """gaussian_multi_distribution_log (lnasr/gmm.py:59-68): full covariance via det + inverse."""
    x = np.asarray(x, dtype=np.float64)
    L, D = x.shape
    det = np.linalg.det(Sigma)
    inv = np.linalg.inv(np.asarray(Sigma, dtype=np.float64))
    diff = x - mu
    quad = np.einsum("ld,de,le->l", diff, inv, diff)
    return -0.5 * (D * np.log(2.0 * np.pi) + np.log(det) + quad)


def gauss_logpdf_diag(x: np.ndarray, mu: np.ndarray, var: np.ndarray) -> np.ndarray:
    """Diagonal restatement of gmm.py:59-68."""
    x = np.asarray(x, dtype=np.float64)
    D = x.shape[1]
    diff = x - mu
    return -0.5 * (D * np.log(2.0 * np.pi) + np.sum(np.log(var)) + np.sum(diff * diff / var, axis=1))


def map_b(obs, w, mu, si):
    """GMMHMM._map_b (lnasr/gmmhmm.py:46-67) -> b [N,T], bm [N,M,T].

    `si` is [N,M,D,D] (full, as in the reference) or [N,M,D] (diagonal variances).
    """
    obs = np.asarray(obs, dtype=np.float64)
    N, M = w.shape
    T = obs.shape[0]
    bm = np.empty((N, M, T))
    b = np.empty((N, T))
    diag = (np.ndim(si) == 3)
    for j in range(N):
        for m in range(M):
            bm[j, m] = (gauss_logpdf_diag if diag else gauss_logpdf_full)(obs, mu[j, m], si[j, m])
        b[j] = lse(w[j][:, None] + bm[j], axis=0)
    return b, bm


# --------------------------------------------------------------------------- HMM
def hmm_map_b(B: np.ndarray, obs: np.ndarray) -> np.ndarray:
    """HMM._map_b for discrete symbols (lnasr/hmm.py:47-61): b[:, t] = B[:, obs[t]]."""
    return np.asarray(B)[:, np.asarray(obs)]


def forward(pi, A, b):
    """HMM._forward (lnasr/hmm.py:63-100).  b is state-major [N,T]; returns alpha [T,N]."""
    N, T = b.shape
    alpha = np.empty((T, N))
    alpha[0] = pi + b[:, 0]
    for t in range(1, T):
        # lse over predecessor i for every j at once; fold order over i as in utils.py:176-179
        alpha[t] = lse(alpha[t - 1][:, None] + A, axis=0) + b[:, t]
    return alpha


def backward(A, b):
    """HMM._backward (lnasr/hmm.py:102-131).  Returns beta [T,N]."""
    N, T = b.shape
    beta = np.empty((T, N))
    beta[T - 1] = 0.0
    for t in range(T - 2, -1, -1):
        # same association as the reference: (beta[t+1] + A[i, :]) + b[:, t+1]  (hmm.py:130)
        beta[t] = lse((beta[t + 1][None, :] + A) + b[:, t + 1][None, :], axis=1)
    return beta


def viterbi(pi, A, b):
    """HMM._viterbi (lnasr/hmm.py:133-172) -> (v [T,N] f64, bt [T,N] u32, path [T] u32).

    argmax takes the first maximum (np.argmax, hmm.py:166,169); bt[0] stays 0.
    """
    N, T = b.shape
    v = np.empty((T, N))
    bt = np.zeros((T, N), dtype=np.uint32)
    v[0] = pi + b[:, 0]
    for t in range(1, T):
        cand = v[t - 1][:, None] + A            # [i, j]
        bt[t] = np.argmax(cand, axis=0)
        v[t] = np.max(cand, axis=0) + b[:, t]
    path = np.empty(T, dtype=np.uint32)
    path[T - 1] = np.argmax(v[T - 1])
    for t in range(T - 2, -1, -1):
        path[t] = bt[t + 1, path[t + 1]]
    return v, bt, path


def baumwelch(A, b, alpha, beta):
    """HMM._baumwelch (lnasr/hmm.py:174-200) -> xi [T,N,N], gamma [T,N]; row T-1 is -inf."""
    N, T = b.shape
    xi = np.full((T, N, N), NEG_INF)
    for t in range(T - 1):
        # reference adds the three terms in this order: (alpha + A) + b + beta (hmm.py:193-196)
        numer = (alpha[t][:, None] + A) + b[:, t + 1][None, :] + beta[t + 1][None, :]
        xi[t] = numer - lse(numer)
    with np.errstate(invalid="ignore"):
        gamma = lse(xi, axis=2)
    return xi, gamma


def estimate_hmm(obs, m, xi, gamma):
    """HMM._estimate, discrete M-step (lnasr/hmm.py:202-225)."""
    obs = np.asarray(obs)
    N = gamma.shape[1]
    denom = lse(gamma, axis=0)
    A = lse(xi, axis=0) - denom[:, None]
    B = np.empty((N, m))
    for k in range(m):
        B[:, k] = lse(gamma[obs == k, :], axis=0) - denom
    return {"A": A, "B": B, "pi": gamma[0]}


def _mixture_posteriors(alpha, beta, w, bm):
    """xi_mix [T,N,M] in log domain (lnasr/gmmhmm.py:89-100)."""
    T, N = alpha.shape
    M = w.shape[1]
    out = np.empty((T, N, M))
    for t in range(T):
        ab = alpha[t] + beta[t]
        wb = w + bm[:, :, t]
        out[t] = ab[:, None] + w + bm[:, :, t]
        out[t] -= lse(ab)
        out[t] -= lse(wb, axis=1)[:, None]
    return out


def estimate_gmmhmm(obs, alpha, beta, xi, gamma, w, mu, bm, min_std=0.01):
    """GMMHMM._estimate (lnasr/gmmhmm.py:69-156), full covariance exactly as the reference.

    Sigma' is the occupancy-weighted scatter about the OLD mean plus min_std on the
    diagonal (gmmhmm.py:138-147).  Written with einsum instead of materialising
    [N,M,T,D,D]; same sums.
    """
    obs = np.asarray(obs, dtype=np.float64)
    N, M, D = mu.shape
    A = lse(xi, axis=0) - lse(gamma, axis=0)[:, None]
    pi = gamma[0]
    lmix = _mixture_posteriors(alpha, beta, w, bm)
    w_new = lse(lmix, axis=0) - lse(lmix, axis=(0, 2))[:, None]
    occ = np.exp(lmix)                                  # [T,N,M]
    s0 = occ.sum(axis=0)                                # [N,M]
    mu_new = np.einsum("tnm,td->nmd", occ, obs) / s0[:, :, None]
    dt = obs[None, None, :, :] - mu[:, :, None, :]      # [N,M,T,D] about the old mean
    si_new = np.einsum("tnm,nmtd,nmte->nmde", occ, dt, dt) / s0[:, :, None, None]
    si_new = si_new + min_std * np.eye(D)
    return {"A": A, "pi": pi, "w": w_new, "mu": mu_new, "si": si_new}


def estimate_gmmhmm_diag(obs, alpha, beta, xi, gamma, w, mu, bm, min_std=0.01):
    """Same M-step keeping only diag(Sigma'): 'var' is [N,M,D]."""
    full = estimate_gmmhmm(obs, alpha, beta, xi, gamma, w, mu, bm, min_std)
    full["var"] = np.diagonal(full["si"], axis1=2, axis2=3).copy()
    return full


# --------------------------------------------------------------------------- drivers
def calc_prob(obs, pi, A, w, mu, si):
    """HMM.calc_prob on a GMMHMM (lnasr/hmm.py:268-274)."""
    b, _ = map_b(obs, w, mu, si)
    return lse(forward(pi, A, b)[-1])


def decode(obs, pi, A, w, mu, si):
    """HMM.decode on a GMMHMM (lnasr/hmm.py:276-283)."""
    b, _ = map_b(obs, w, mu, si)
    return viterbi(pi, A, b)[2]


def train_gmmhmm(obs, model, iters=1, eps=1e-4, project_diag=False, min_std=0.01):
    """HMM.train loop on a GMMHMM (lnasr/hmm.py:285-313).  `model` = dict(A,pi,w,mu,si).

    project_diag=True zeroes the off-diagonals of Sigma after every update, which is the
    oracle definition for multi-iteration parity of the diagonal build.
    Returns (model, history of (prob_old, prob_new)).
    """
    m = {k: np.array(v, dtype=np.float64) for k, v in model.items()}
    hist = []
    for _ in range(iters):
        b, bm = map_b(obs, m["w"], m["mu"], m["si"])
        alpha = forward(m["pi"], m["A"], b)
        beta = backward(m["A"], b)
        xi, gamma = baumwelch(m["A"], b, alpha, beta)
        new = estimate_gmmhmm(obs, alpha, beta, xi, gamma, m["w"], m["mu"], bm, min_std)
        prob_old = lse(alpha[-1])
        if project_diag:
            d = np.diagonal(new["si"], axis1=2, axis2=3)
            new["si"] = d[..., None] * np.eye(d.shape[-1])
        m = new
        prob_new = calc_prob(obs, m["pi"], m["A"], m["w"], m["mu"], m["si"])
        hist.append((float(prob_old), float(prob_new)))
        if abs(prob_old - prob_new) < eps:
            break
    return m, hist


# ------------------------------------------------- multi-utterance EM (diagonal), SURVEY 8(c)
def em_accumulate_diag(obs, pi, A, w, mu, var):
    """Linear-domain sufficient statistics of ONE utterance, built from the reference pieces.

    Returns dict: S0 [N,M], S1 [N,M,D], S2c [N,M,D] (scatter about the OLD mean), xi_sum [N,N]
    (sum_t exp xi[t], t<T-1), pi_occ [N] (exp gamma[0]), loglik (lse alpha[T-1]).
    With one utterance `em_finalize_diag` of these equals the reference `_estimate`
    (diagonal of Sigma') -- tests/test_oracle_golden.py checks that.
    """
    obs = np.asarray(obs, dtype=np.float64)
    si = var[..., None] * np.eye(var.shape[-1])
    b, bm = map_b(obs, w, mu, si)
    alpha = forward(pi, A, b)
    beta = backward(A, b)
    xi, gamma = baumwelch(A, b, alpha, beta)
    occ = np.exp(_mixture_posteriors(alpha, beta, w, bm))
    dt2 = (obs[None, None, :, :] - mu[:, :, None, :]) ** 2
    with np.errstate(under="ignore"):
        return {
            "S0": occ.sum(axis=0),
            "S1": np.einsum("tnm,td->nmd", occ, obs),
            "S2c": np.einsum("tnm,nmtd->nmd", occ, dt2),
            "xi_sum": np.exp(xi[:-1]).sum(axis=0),
            "pi_occ": np.exp(gamma[0]),
            "loglik": float(lse(alpha[-1])),
            "n_utt": 1.0,
        }


def em_finalize_diag(acc, min_std=0.01):
    """M-step from summed statistics: A, pi, w in log domain; mu, var linear."""
    with np.errstate(divide="ignore"):
        A = np.log(acc["xi_sum"]) - np.log(acc["xi_sum"].sum(axis=1))[:, None]
        pi = np.log(acc["pi_occ"] / acc["n_utt"])
        w = np.log(acc["S0"]) - np.log(acc["S0"].sum(axis=1))[:, None]
    mu = acc["S1"] / acc["S0"][:, :, None]
    var = acc["S2c"] / acc["S0"][:, :, None] + min_std
    return {"A": A, "pi": pi, "w": w, "mu": mu, "var": var}


def em_step_multi_diag(utts, model, min_std=0.01):
    """One EM iteration over a list of utterances: sum of per-utterance statistics."""
    tot = None
    for obs in utts:
        acc = em_accumulate_diag(obs, model["pi"], model["A"], model["w"], model["mu"], model["var"])
        tot = acc if tot is None else {k: tot[k] + acc[k] for k in tot}
    return em_finalize_diag(tot, min_std), tot


# --------------------------------------------------------------------------- helpers
def random_gmmhmm(n, m, d, rng: np.random.Generator, var_lo=0.5, var_hi=1.5, left_to_right=False):
    """Random-init model with GMMHMM.reset('random') semantics (lnasr/gmmhmm.py:179-188):
    row-normalised uniform pi/A/w in log domain, mu ~ U(-0.3,0.3); variances U(var_lo,var_hi)
    on the diagonal.  left_to_right keeps self/next/skip transitions only."""
    a = rng.random((n, n))
    if left_to_right:
        mask = np.zeros((n, n), dtype=bool)
        for i in range(n):
            for j in (i, i + 1, i + 2):
                if j < n:
                    mask[i, j] = True
        a = np.where(mask, a, 0.0)
    p = rng.random(n)
    ww = rng.random((n, m))
    with np.errstate(divide="ignore"):
        A = np.log(a / a.sum(axis=1, keepdims=True))
        pi = np.log(p / p.sum())
        w = np.log(ww / ww.sum(axis=1, keepdims=True))
    mu = 0.6 * rng.random((n, m, d)) - 0.3
    var = var_lo + (var_hi - var_lo) * rng.random((n, m, d))
    return {"A": A, "pi": pi, "w": w, "mu": mu, "var": var}


def score_words(obs, models):
    """Isolated-word decode: per word model the Viterbi score max(v[T-1]),
    the path and calc_prob; word = argmax (lowest index on ties)."""
    scores, probs, paths = [], [], []
    for mdl in models:
        b, _ = map_b(obs, mdl["w"], mdl["mu"], mdl["var"])
        v, _, path = viterbi(mdl["pi"], mdl["A"], b)
        scores.append(np.max(v[-1]))
        paths.append(path)
        probs.append(lse(forward(mdl["pi"], mdl["A"], b)[-1]))
    scores = np.array(scores)
    probs = np.array(probs)
    return int(np.argmax(scores)), scores, int(np.argmax(probs)), probs, paths


def score_words_viterbi(obs, models):
    """Viterbi-only isolated-word decision (what the batched GPU decode computes): word, scores, paths."""
    scores, paths = [], []
    for mdl in models:
        b, _ = map_b(obs, mdl["w"], mdl["mu"], mdl["var"])
        v, _, path = viterbi(mdl["pi"], mdl["A"], b)
        scores.append(np.max(v[-1]))
        paths.append(path)
    scores = np.array(scores)
    return int(np.argmax(scores)), scores, paths
