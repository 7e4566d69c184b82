#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (yehuohan/ln-asr).

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference is imported as-is under four import shims: stub modules for
pyaudio / librosa / h5py (absent here, not used on the hot path) and the numpy-1 aliases
`np.mat`, `np.float`.  No reference source is copied; only input/output vectors are stored.
The GPU box has no /root/reference, so tests read only the committed .npz files.
"""
import os
import sys
import types
import warnings

import numpy as np

warnings.filterwarnings("ignore")
REF = os.environ.get("LNASR_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    for name in ("pyaudio", "librosa", "h5py"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if not hasattr(np, "mat"):
        np.mat = np.asmatrix
    if not hasattr(np, "float"):
        np.float = float
    sys.path.insert(0, REF)
    import lnasr.utils as ru
    import lnasr.mfcc as rm
    import lnasr.gmm as rg
    import lnasr.hmm as rh
    import lnasr.gmmhmm as rgh
    return ru, rm, rg, rh, rgh


def synth_pcm(rng, n, f0):
    """SURVEY.md section 8(d) synthetic audio: noise + tone, int16."""
    t = np.arange(n) / 16000.0
    x = np.round(3000.0 * rng.standard_normal(n)) + 2000.0 * np.sin(2 * np.pi * f0 * t)
    return np.clip(x, -32767, 32767).astype(np.int16)


def main():
    ru, rm, rg, rh, rgh = import_reference()
    np.seterr(all="ignore")

    # ---------------------------------------------------------------- lse / framing
    rng = np.random.default_rng(7)
    a3 = rng.standard_normal((4, 5, 6)) * 3
    a3[1, 2, :] = -np.inf
    np.savez_compressed(
        os.path.join(OUT, "utils.npz"),
        a3=a3,
        lse_all=ru.lse(a3), lse_ax0=ru.lse(a3, 0), lse_ax2=ru.lse(a3, 2), lse_ax02=ru.lse(a3, (0, 2)),
        hamming400=ru.create_hamming(400),
        frames_1000=ru.split_frames(np.arange(1000.0), 400, 160),
        frames_100=ru.split_frames(np.arange(100.0), 400, 160),
        nframes_S=np.array([100, 240, 241, 400, 401, 560, 561, 16000, 21751, 960000]),
        nframes_F=np.array([ru.split_frames(np.zeros(s), 400, 160).shape[0]
                            for s in [100, 240, 241, 400, 401, 560, 561, 16000, 21751, 960000]]),
    )

    # ---------------------------------------------------------------- MFCC
    pcm = ru.read_pcm(os.path.join(REF, "test", "data.raw"))
    m = rm.MFCC()
    P, C, F = m.calc_mfcc(pcm)
    out = dict(pcm_data_raw=pcm, power=P, ceps=C, feats=F, fbank=m.create_filter_bank())
    rng = np.random.default_rng(1234)
    for tag, n in (("s401", 401), ("s1000", 1000), ("s16000", 16000), ("s48017", 48017)):
        x = synth_pcm(rng, n, rng.uniform(80, 400))
        p_, c_, f_ = m.calc_mfcc(x)
        out["pcm_" + tag], out["power_" + tag], out["ceps_" + tag], out["feats_" + tag] = x, p_, c_, f_
    # float input, and an utterance with an all-zero frame in the middle (-> -inf energy, eps mel)
    xf = rng.standard_normal(4000) * 0.1
    p_, c_, f_ = m.calc_mfcc(xf)
    out["pcm_float"], out["power_float"], out["ceps_float"], out["feats_float"] = xf, p_, c_, f_
    xz = synth_pcm(rng, 3200, 200.0)
    xz[800:1700] = 0
    p_, c_, f_ = m.calc_mfcc(xz)
    out["pcm_zero"], out["power_zero"], out["ceps_zero"], out["feats_zero"] = xz, p_, c_, f_
    np.savez_compressed(os.path.join(OUT, "mfcc.npz"), **out)

    # ---------------------------------------------------------------- ice-cream HMM (test/hmm-test.py:10-39)
    A = np.array([[0.6, 0.4], [0.5, 0.5]])
    B = np.array([[0.2, 0.4, 0.4], [0.5, 0.4, 0.1]])
    pi = np.array([0.8, 0.2])
    obs = np.array([2, 0, 2], dtype=np.int32)
    h = rh.HMM(2, 3, np.log(A), np.log(B), np.log(pi))
    alpha = h._forward(obs, True)
    beta = h._backward(obs)
    v, bt, path = h._viterbi(obs)
    xi, gamma = h._baumwelch(obs, alpha, beta)
    mdl = h._estimate(obs, alpha, beta, xi, gamma)
    ice = dict(A=np.log(A), B=np.log(B), pi=np.log(pi), obs=obs, alpha=alpha, beta=beta, v=v, bt=bt, path=path,
               xi=xi, gamma=gamma, A_new=mdl["A"], B_new=mdl["B"], pi_new=mdl["pi"], prob=h.calc_prob(obs),
               decode=h.decode(obs))
    # seeded discrete HMM training, shapes of test/hmm-test.py:42-55
    np.random.seed(11)
    obs2 = np.array((0, 1, 1, 2, 0, 4, 2, 3, 4, 2, 0, 2, 3, 4, 4, 2, 1, 0, 1) * 3)
    h2 = rh.HMM(4, 5)
    h2.reset("random")
    ice.update(d_A0=h2.A.copy(), d_B0=h2.B.copy(), d_pi0=h2.pi.copy(), d_obs=obs2)
    h2.train(obs2, 5, verbose=False)
    ice.update(d_A5=h2.A, d_B5=h2.B, d_pi5=h2.pi, d_prob5=h2.calc_prob(obs2), d_path5=h2.decode(obs2))
    np.savez_compressed(os.path.join(OUT, "hmm_discrete.npz"), **ice)

    # ---------------------------------------------------------------- GMMHMM (5,4,39) on standardised data.raw MFCC
    np.random.seed(0)
    g = rgh.GMMHMM(5, 4, 39)
    g.reset("random")
    Fs = (F - F.mean(0)) / F.std(0) * 0.3
    gm = dict(obs=Fs, A=g.A.copy(), pi=g.pi.copy(), w=g.w.copy(), mu=g.mu.copy(), si=g.si.copy())
    gm["prob"] = g.calc_prob(Fs)
    gm["b"], gm["bm"] = g.b.copy(), g.bm.copy()
    alpha = g._forward(Fs)
    beta = g._backward(Fs)
    v, bt, path = g._viterbi(Fs)
    xi, gamma = g._baumwelch(Fs, alpha, beta)
    est = g._estimate(Fs, alpha, beta, xi, gamma)
    gm.update(alpha=alpha, beta=beta, v=v, bt=bt, path=path, gamma=gamma,
              xi_lse_t=ru.lse(xi, axis=0),
              est_A=est["A"], est_pi=est["pi"], est_w=est["w"], est_mu=est["mu"], est_si=est["si"])
    # diagonal, non-identity variances (what the build targets): emissions + one estimate
    rng = np.random.default_rng(5)
    var = 0.5 + rng.random((5, 4, 39))
    g2 = rgh.GMMHMM(5, 4, 39, A=g.A, pi=g.pi, w=g.w, mu=g.mu, si=var[..., None] * np.eye(39))
    g2._map_b(Fs)
    a2 = g2._forward(Fs)
    b2 = g2._backward(Fs)
    xi2, gam2 = g2._baumwelch(Fs, a2, b2)
    e2 = g2._estimate(Fs, a2, b2, xi2, gam2)
    v2, bt2, p2 = g2._viterbi(Fs)
    gm.update(dg_var=var, dg_b=g2.b, dg_bm=g2.bm, dg_alpha=a2, dg_beta=b2, dg_path=p2, dg_vlast=v2[-1],
              dg_prob=ru.lse(a2[-1]), dg_est_A=e2["A"], dg_est_pi=e2["pi"], dg_est_w=e2["w"], dg_est_mu=e2["mu"],
              dg_est_var=np.diagonal(e2["si"], axis1=2, axis2=3))
    np.savez_compressed(os.path.join(OUT, "gmmhmm_5_4_39.npz"), **gm)

    # ---------------------------------------------------------------- config 1: test/gmmhmm-test.py shapes, seeded
    np.random.seed(3)
    n, mm, d, T = 5, 3, 2, 40
    obs = np.array(0.6 * np.random.random_sample((T, d)) - 0.3, dtype=np.double)
    gh = rgh.GMMHMM(n, mm, d)
    gh.reset("random")
    c1 = dict(obs=obs, A0=gh.A.copy(), pi0=gh.pi.copy(), w0=gh.w.copy(), mu0=gh.mu.copy(), si0=gh.si.copy())
    gh.train(obs, 1, verbose=False)
    c1.update(A1=gh.A.copy(), pi1=gh.pi.copy(), w1=gh.w.copy(), mu1=gh.mu.copy(), si1=gh.si.copy(),
              prob1=gh.calc_prob(obs), path1=gh.decode(obs))

    # 10 iterations with off-diagonals zeroed after every update (oracle definition for the diagonal
    # build): a harness subclass, reference files untouched.
    class DiagGMMHMM(rgh.GMMHMM):
        def _updatemodel(self, model):
            super()._updatemodel(model)
            dg = np.diagonal(self.si, axis1=2, axis2=3)
            self.si = dg[..., None] * np.eye(self.d)

    gd = DiagGMMHMM(n, mm, d, A=c1["A0"].copy(), pi=c1["pi0"].copy(), w=c1["w0"].copy(), mu=c1["mu0"].copy(),
                    si=c1["si0"].copy())
    gd.train(obs, 10, eps=0.0, verbose=False)
    c1.update(dA10=gd.A, dpi10=gd.pi, dw10=gd.w, dmu10=gd.mu, dvar10=np.diagonal(gd.si, axis1=2, axis2=3),
              dprob10=gd.calc_prob(obs), dpath10=gd.decode(obs))
    np.savez_compressed(os.path.join(OUT, "gmmhmm_config1.npz"), **c1)

    # ---------------------------------------------------------------- gmm.py density functions
    rng = np.random.default_rng(9)
    x = rng.standard_normal((17, 3))
    mu = rng.standard_normal(3)
    Lm = rng.standard_normal((3, 3))
    Sig = Lm @ Lm.T + 0.5 * np.eye(3)
    cw = np.array([0.2, 0.5, 0.3])
    mus = rng.standard_normal((3, 3))
    Sigs = np.stack([Sig, np.eye(3), 2 * np.eye(3)])
    np.savez_compressed(
        os.path.join(OUT, "gmm.npz"), x=x, mu=mu, Sigma=Sig, c=cw, mus=mus, Sigmas=Sigs,
        g1=rg.gaussian_distribution(x[:, 0], 0.3, 1.7), g1log=rg.gaussian_distribution_log(x[:, 0], 0.3, 1.7),
        gm=rg.gaussian_multi_distribution(x, mu, Sig), gmlog=rg.gaussian_multi_distribution_log(x, mu, Sig),
        gmix=rg.gaussian_mixture_distribution(cw, x, mus, Sigs),
        gmixlog=rg.gaussian_mixture_distribution_log(np.log(cw), x, mus, Sigs),
    )
    # ---------------------------------------------------------------- VadLtsd (third/vadlstd/VadLstd.py), the step before the path
    import importlib.util
    sys.modules.setdefault("snoop", types.ModuleType("snoop"))           # imported at VadLstd.py:19, used only in a comment
    spec = importlib.util.spec_from_file_location("ref_vadlstd", os.path.join(REF, "third", "vadlstd", "VadLstd.py"))
    rv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rv)
    rng = np.random.default_rng(21)
    vad = {}
    for tag, n, kw in (("a", 40000, {}), ("b", 30001, {"alpha": 0.9, "threshold": 30.0}), ("c", 9000, {"winsize": 512, "stepsize": 256, "order": 3}),
                       ("d", 2500, {})):
        t = np.arange(n) / 16000.0
        x = 300.0 * rng.standard_normal(n)
        x[n // 3:2 * n // 3] += 4000.0 * np.sin(2 * np.pi * 220.0 * t[n // 3:2 * n // 3])      # a loud stretch in the middle
        x = np.clip(np.round(x), -32767, 32767).astype(np.int16)
        vad["pcm_" + tag] = x
        vad["ltsd_" + tag] = rv.VadLtsd(**kw).detect(x.astype(np.float64))
    np.savez_compressed(os.path.join(OUT, "vad.npz"), **vad)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
