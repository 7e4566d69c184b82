"""Baum-Welch: fused backward + sufficient-statistic accumulation and the M-step (csrc/em.cu) against the
reference golden vectors and the oracle.  Bar: re-estimated means and variances within 1e-4 relative."""
import numpy as np
import pytest

from helpers import assert_close

pytestmark = pytest.mark.gpu


def test_discrete_estimate_icecream(golden):
    from lnasr.hmm import HMM
    g = golden("hmm_discrete")
    h = HMM(2, 3, g["A"], g["B"], g["pi"])
    alpha = h._forward(g["obs"], True)
    beta = h._backward(g["obs"])
    xi, gamma = h._baumwelch(g["obs"], alpha, beta)
    est = h._estimate(g["obs"], alpha, beta, xi, gamma)
    with np.errstate(all="ignore"):
        assert_close(np.exp(est["A"]), np.exp(g["A_new"]), 1e-10, "A")
        assert_close(np.exp(est["B"]), np.exp(g["B_new"]), 1e-10, "B")
        assert_close(np.exp(est["pi"]), np.exp(g["pi_new"]), 1e-10, "pi")
    np.testing.assert_allclose(np.exp(est["A"]), [[.5372005, .4627995], [.75341028, .24658972]], atol=1e-7)


def test_discrete_training_matches_reference(golden, capsys):
    from lnasr.hmm import HMM
    g = golden("hmm_discrete")
    h = HMM(4, 5, g["d_A0"].copy(), g["d_B0"].copy(), g["d_pi0"].copy())
    h.train(g["d_obs"], 5, verbose=True)
    out = capsys.readouterr().out
    assert out.count("Iter:") == 5 and "L(lambda|O)" in out and "eps =" in out          # hmm.py:307-311 format
    assert_close(np.exp(h.A), np.exp(g["d_A5"]), 1e-8, "A5")
    assert_close(np.exp(h.B), np.exp(g["d_B5"]), 1e-8, "B5")
    assert_close(np.exp(h.pi), np.exp(g["d_pi5"]), 1e-8, "pi5")
    assert abs(h.calc_prob(g["d_obs"]) - float(g["d_prob5"])) < 1e-8
    assert np.array_equal(h.decode(g["d_obs"]), g["d_path5"])


def test_gmmhmm_estimate_matches_reference_diagonal(golden):
    """One reference `_estimate` on a diagonal model: A, pi, w, mu and diag(Sigma')."""
    from lnasr.gmmhmm import GMMHMM
    g = golden("gmmhmm_5_4_39")
    si = g["dg_var"][..., None] * np.eye(39)
    m = GMMHMM(5, 4, 39, A=g["A"], pi=g["pi"], w=g["w"], mu=g["mu"], si=si)
    obs = g["obs"]
    m._map_b(obs)
    alpha = m._forward(obs)
    beta = m._backward(obs)
    np.testing.assert_allclose(alpha, g["dg_alpha"], rtol=1e-11, atol=1e-9)
    np.testing.assert_allclose(beta, g["dg_beta"], rtol=1e-11, atol=1e-9)
    est = m._estimate(obs, alpha, beta, None, None)
    assert est["si"].shape == (5, 4, 39, 39)
    assert_close(np.exp(est["A"]), np.exp(g["dg_est_A"]), 1e-9, "A")
    assert_close(np.exp(est["pi"]), np.exp(g["dg_est_pi"]), 1e-9, "pi")
    assert_close(np.exp(est["w"]), np.exp(g["dg_est_w"]), 1e-9, "w")
    assert_close(est["mu"], g["dg_est_mu"], 1e-9, "mu")                    # bar: 1e-4 relative
    assert_close(np.diagonal(est["si"], axis1=2, axis2=3), g["dg_est_var"], 1e-9, "var")


def test_gmmhmm_estimate_from_identity_si_matches_reference_full_run(golden):
    """The reference's own run (si = I): our diag(Sigma') must equal the diagonal of its full Sigma'."""
    from lnasr.gmmhmm import GMMHMM
    g = golden("gmmhmm_5_4_39")
    m = GMMHMM(5, 4, 39, A=g["A"], pi=g["pi"], w=g["w"], mu=g["mu"], si=g["si"])
    obs = g["obs"]
    m._map_b(obs)
    alpha = m._forward(obs)
    est = m._estimate(obs, alpha, None, None, None)
    assert_close(est["mu"], g["est_mu"], 1e-9, "mu")
    assert_close(np.diagonal(est["si"], axis1=2, axis2=3), np.diagonal(g["est_si"], axis1=2, axis2=3), 1e-9, "var")
    assert_close(np.exp(est["w"]), np.exp(g["est_w"]), 1e-9, "w")
    assert_close(np.exp(est["A"]), np.exp(g["est_A"]), 1e-9, "A")


def test_config1_ten_iterations_match_projected_reference(golden):
    """BASELINE config 1 (test/gmmhmm-test.py shapes, seeded): 10 EM iterations against the reference with
    off-diagonals zeroed after every update (the oracle definition for the diagonal build)."""
    from lnasr.gmmhmm import GMMHMM
    g = golden("gmmhmm_config1")
    m = GMMHMM(5, 3, 2, A=g["A0"].copy(), pi=g["pi0"].copy(), w=g["w0"].copy(), mu=g["mu0"].copy(), si=g["si0"].copy())
    m.train(g["obs"], 10, eps=0.0, verbose=False)
    assert_close(m.mu, g["dmu10"], 1e-7, "mu10")
    assert_close(m.var, g["dvar10"], 1e-7, "var10")
    assert_close(np.exp(m.w), np.exp(g["dw10"]), 1e-7, "w10")
    assert_close(np.exp(m.A), np.exp(g["dA10"]), 1e-7, "A10")
    assert abs(m.calc_prob(g["obs"]) - float(g["dprob10"])) < 1e-6
    assert np.array_equal(m.decode(g["obs"]), g["dpath10"])


def test_config1_first_iteration_matches_reference_exactly(golden):
    """After ONE reference train step the means / weights / transitions and diag(Sigma) agree with the unprojected reference."""
    from lnasr.gmmhmm import GMMHMM
    g = golden("gmmhmm_config1")
    m = GMMHMM(5, 3, 2, A=g["A0"].copy(), pi=g["pi0"].copy(), w=g["w0"].copy(), mu=g["mu0"].copy(), si=g["si0"].copy())
    m.train(g["obs"], 1, verbose=False)
    assert_close(m.mu, g["mu1"], 1e-9, "mu1")
    assert_close(m.var, np.diagonal(g["si1"], axis1=2, axis2=3), 1e-9, "var1")
    assert_close(np.exp(m.A), np.exp(g["A1"]), 1e-9, "A1")


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("n,mix,d,n_utt", [(3, 8, 39, 12), (5, 4, 39, 7), (1, 2, 3, 4), (8, 2, 5, 9), (2, 1, 7, 33), (17, 3, 11, 5)])
def test_multi_utterance_em_against_oracle(n, mix, d, n_utt, split):
    """BASELINE config 3 shapes (3 states x 8 mixtures x 39 dims): statistics summed over utterances."""
    from lnasr.gmmhmm import GMMHMM
    from oracle import lnasr_oracle as O
    rng = np.random.default_rng(n * 31 + mix)
    mdl = O.random_gmmhmm(n, mix, d, rng)
    utts = [0.6 * rng.standard_normal((int(T), d)) for T in rng.integers(2, 60, size=n_utt)]
    utts[0] = utts[0][:1]                                  # a one-frame utterance: no transition statistics
    new_ref, tot = O.em_step_multi_diag(utts, mdl)
    m = GMMHMM(n, mix, d, A=mdl["A"], pi=mdl["pi"], w=mdl["w"], mu=mdl["mu"], si=mdl["var"][..., None] * np.eye(d))
    m._fb_split = split                                    # fused single kernel / lane-group recursion + frame-parallel statistics
    hist = m.train_batch(utts, iters=1, verbose=False)
    assert abs(hist[0][0] - tot["loglik"]) < 1e-3 * n_utt                   # 1e-3 absolute per utterance
    assert abs(hist[0][0] - tot["loglik"]) < 1e-8 * abs(tot["loglik"])
    assert_close(m.mu, new_ref["mu"], 1e-8, "mu")
    assert_close(m.var, new_ref["var"], 1e-8, "var")
    with np.errstate(all="ignore"):
        assert_close(np.exp(m.w), np.exp(new_ref["w"]), 1e-8, "w")
        assert_close(np.exp(m.A), np.exp(new_ref["A"]), 1e-8, "A")
        assert_close(np.exp(m.pi), np.exp(new_ref["pi"]), 1e-8, "pi")


@pytest.mark.parametrize("n,mix,d,n_utt", [(3, 8, 39, 9), (1, 2, 3, 5), (17, 3, 11, 4), (32, 2, 5, 3), (5, 4, 39, 40)])
def test_device_resident_em_trainer_matches_oracle(n, mix, d, n_utt):
    """lnasr.em.EmTrainer (model stays on the device; forward || backward on two streams, then the frame-parallel
    lnasr_em_statistics): three iterations against the oracle's multi-utterance EM, and against GMMHMM.train_batch."""
    import torch
    from lnasr.em import EmTrainer
    from lnasr.gmmhmm import GMMHMM
    from oracle import lnasr_oracle as O
    rng = np.random.default_rng(3 + n)
    mdl = O.random_gmmhmm(n, mix, d, rng)
    utts = [0.6 * rng.standard_normal((int(T), d)) for T in rng.integers(20, 60, size=n_utt)]
    utts[-1] = utts[-1][:1]                                   # a one-frame utterance: no transition / pi statistics
    utts[0] = utts[0][:2]
    off = np.concatenate([[0], np.cumsum([len(u) for u in utts])])
    tr = EmTrainer(mdl).load_batch(torch.from_numpy(np.concatenate(utts)).cuda(), off)
    ref = dict(mdl)
    m = GMMHMM(n, mix, d, A=mdl["A"], pi=mdl["pi"], w=mdl["w"], mu=mdl["mu"], si=mdl["var"][..., None] * np.eye(d))
    for it in range(3):
        ll = float(tr.iterate().item())
        ref, tot = O.em_step_multi_diag(utts, ref)
        assert abs(ll - tot["loglik"]) < 1e-8 * abs(tot["loglik"]), it
    hist = m.train_batch(utts, iters=3, eps=0.0, verbose=False)
    got = tr.export()
    assert_close(got["mu"], ref["mu"], 1e-7, "mu")                          # bar: 1e-4 relative
    assert_close(got["var"], ref["var"], 1e-7, "var")
    assert_close(got["mu"], m.mu, 1e-9, "mu vs train_batch")
    assert_close(got["var"], m.var, 1e-9, "var vs train_batch")
    with np.errstate(all="ignore"):
        assert_close(np.exp(got["w"]), np.exp(ref["w"]), 1e-7, "w")
        assert_close(np.exp(got["A"]), np.exp(ref["A"]), 1e-7, "A")
        assert_close(np.exp(got["pi"]), np.exp(ref["pi"]), 1e-7, "pi")
    assert got["si"].shape == (n, mix, d, d) and len(hist) == 3
