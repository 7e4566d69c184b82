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


@pytest.mark.parametrize("n,mix,d,ltr", [(200, 4, 39, True), (40, 3, 39, False), (33, 1, 5, False), (70, 2, 13, True)])
def test_baum_welch_beyond_thirty_two_states(n, mix, d, ltr):
    """HMM.train (hmm.py:285-313) / GMMHMM._estimate (gmmhmm.py:69-156) take any n: above 32 states the backward pass is a plain
    trellis pass (ELL transition lists) and the statistics come from the general frame-parallel kernels (compact transition
    list, accumulators in global memory).  GMMHMM.train_batch and the device-resident EmTrainer against the oracle's
    multi-utterance EM: left-to-right (sparse) and dense transition matrices, one-frame and two-frame utterances."""
    import torch
    from lnasr.em import EmTrainer
    from lnasr.gmmhmm import GMMHMM
    from oracle import lnasr_oracle as O
    rng = np.random.default_rng(5 * n + mix)
    mdl = O.random_gmmhmm(n, mix, d, rng, left_to_right=ltr)
    lens = [47, 1, 33, 2, 60] if n <= 70 else [41, 1, 30]
    utts = [0.6 * rng.standard_normal((T, d)) for T in lens]
    with np.errstate(all="ignore"):
        new_ref, tot = O.em_step_multi_diag(utts, mdl)
    m = GMMHMM(n, mix, d, A=mdl["A"], pi=mdl["pi"], w=mdl["w"], mu=mdl["mu"], si=mdl["var"][..., None] * np.eye(d))
    hist = m.train_batch(utts, iters=1, verbose=False)
    assert abs(hist[0][0] - tot["loglik"]) < 1e-8 * abs(tot["loglik"])
    off = np.concatenate([[0], np.cumsum(lens)])
    tr = EmTrainer(mdl).load_batch(torch.from_numpy(np.concatenate(utts)).cuda(), off)
    ll = float(tr.iterate().item())
    assert abs(ll - tot["loglik"]) < 1e-8 * abs(tot["loglik"])
    got = tr.export()
    for name, mu_, var_, w_, A_, pi_ in (("train_batch", m.mu, m.var, m.w, m.A, m.pi),
                                         ("EmTrainer", got["mu"], got["var"], got["w"], got["A"], got["pi"])):
        assert_close(mu_, new_ref["mu"], 1e-8, name + " mu")                 # bar: 1e-4 relative
        assert_close(var_, new_ref["var"], 1e-8, name + " var")
        with np.errstate(all="ignore"):
            assert_close(np.exp(w_), np.exp(new_ref["w"]), 1e-8, name + " w")
            assert_close(np.exp(A_), np.exp(new_ref["A"]), 1e-8, name + " A", per_column=False)
            assert_close(np.exp(pi_), np.exp(new_ref["pi"]), 1e-8, name + " pi")
    # a second iteration runs on the re-packed transition lists of the updated model
    ll2 = float(tr.iterate().item())
    with np.errstate(all="ignore"):
        _, tot2 = O.em_step_multi_diag(utts, new_ref)
    assert abs(ll2 - tot2["loglik"]) < 1e-7 * abs(tot2["loglik"]) and ll2 >= ll - 1e-6


def test_moment_kernels_agree():
    """Above 32 states the Gaussian moments run on the FP64 tensor cores (raw moments about a point of the tile, dim <= 39) or
    on the SIMT kernel (centred on mu_old; dim <= 64): the two give the same re-estimates, far from the origin too."""
    from lnasr import _capi
    from lnasr.gmmhmm import GMMHMM
    from oracle import lnasr_oracle as O
    rng = np.random.default_rng(77)
    n, mix, d = 40, 3, 39
    mdl = O.random_gmmhmm(n, mix, d, rng, left_to_right=True)
    mdl["mu"] = mdl["mu"] + 250.0                                   # |mu| / sigma ~ 250: the raw-moment form must not cancel
    utts = [250.0 + 0.7 * rng.standard_normal((T, d)) for T in (50, 3, 41)]
    res = []
    for simt in (0, 1):
        _capi.lib().lnasr_debug_set_em_moments_simt(simt)
        try:
            m = GMMHMM(n, mix, d, A=mdl["A"], pi=mdl["pi"], w=mdl["w"], mu=mdl["mu"], si=mdl["var"][..., None] * np.eye(d))
            m.train_batch(utts, iters=1, verbose=False)
        finally:
            _capi.lib().lnasr_debug_set_em_moments_simt(0)
        res.append((m.mu.copy(), m.var.copy(), m.w.copy()))
    with np.errstate(all="ignore"):
        new_ref, _ = O.em_step_multi_diag(utts, mdl)
    for mu_, var_, w_ in res:
        assert_close(mu_, new_ref["mu"], 1e-9, "mu")
        assert_close(var_, new_ref["var"], 1e-7, "var")
    assert_close(res[0][1], res[1][1], 1e-8, "var: tensor-core vs SIMT moments")


def test_large_model_training_from_float32_features():
    """The features a pipeline hands to the trainer are float32 (MFCC output): the moments kernel fetches them with 4-byte
    asynchronous copies.  40 states x 3 mixtures, a Gaussian tile that is not full (120 = 3 x 32 + 24), utterances of 1..57 frames;
    against the oracle on the same float32 values."""
    import torch
    from lnasr.em import EmTrainer
    from oracle import lnasr_oracle as O
    rng = np.random.default_rng(404)
    n, mix, d = 40, 3, 39
    mdl = O.random_gmmhmm(n, mix, d, rng, left_to_right=True)
    lens = [57, 1, 32, 31, 33, 2]
    utts32 = [(0.6 * rng.standard_normal((T, d))).astype(np.float32) for T in lens]
    with np.errstate(all="ignore"):
        new_ref, tot = O.em_step_multi_diag([u.astype(np.float64) for u in utts32], mdl)
    off = np.concatenate([[0], np.cumsum(lens)])
    tr = EmTrainer(mdl).load_batch(torch.from_numpy(np.concatenate(utts32)).cuda(), off)
    ll = float(tr.iterate().item())
    assert abs(ll - tot["loglik"]) < 1e-8 * abs(tot["loglik"])
    got = tr.export()
    assert_close(got["mu"], new_ref["mu"], 1e-8, "mu")
    assert_close(got["var"], new_ref["var"], 1e-8, "var")
    with np.errstate(all="ignore"):
        assert_close(np.exp(got["w"]), np.exp(new_ref["w"]), 1e-8, "w")


def test_discrete_hmm_training_beyond_thirty_two_states():
    """Plain HMM._estimate (hmm.py:202-225) with 48 states and 10 symbols: one Baum-Welch step against the oracle."""
    from lnasr.hmm import HMM
    from oracle import lnasr_oracle as O
    rng = np.random.default_rng(48)
    n, m, T = 48, 10, 90
    np.random.seed(4)
    h = HMM(n, m)
    h.reset("random")
    A, B, pi = h.A.copy(), h.B.copy(), h.pi.copy()
    obs = rng.integers(0, m, size=T)
    b = O.hmm_map_b(B, obs)
    with np.errstate(all="ignore"):
        alpha, beta = O.forward(pi, A, b), O.backward(A, b)
        xi, gamma = O.baumwelch(A, b, alpha, beta)
        ref = O.estimate_hmm(obs, m, xi, gamma)
    hist = h.train(obs, iters=1, verbose=False)
    assert abs(hist[0][0] - O.lse(alpha[-1])) < 1e-9 * abs(O.lse(alpha[-1]))
    with np.errstate(all="ignore"):
        assert_close(np.exp(h.A), np.exp(ref["A"]), 1e-9, "A", per_column=False)
        assert_close(np.exp(h.B), np.exp(ref["B"]), 1e-9, "B", per_column=False)
        assert_close(np.exp(h.pi), np.exp(ref["pi"]), 1e-9, "pi")


def test_in_place_parameter_edits_are_seen():
    """The device copies of the parameters are keyed on the CONTENT of the numpy arrays: an in-place edit of A or mu changes
    the next result, and single-sequence calc_prob / decode leave self.b / self.bm filled like the reference's remap=True."""
    from lnasr.gmmhmm import GMMHMM
    from oracle import lnasr_oracle as O
    rng = np.random.default_rng(2)
    n, mix, d = 60, 2, 5
    mdl = O.random_gmmhmm(n, mix, d, rng, left_to_right=True)
    x = 0.6 * rng.standard_normal((40, d))
    g = GMMHMM(n, mix, d, A=mdl["A"].copy(), pi=mdl["pi"].copy(), w=mdl["w"].copy(), mu=mdl["mu"].copy(),
               si=mdl["var"][..., None] * np.eye(d))
    p0 = g.calc_prob(x)
    assert g.b.shape == (n, 40) and g.bm.shape == (n, mix, 40)
    np.testing.assert_allclose(g.b, O.map_b(x, mdl["w"], mdl["mu"], mdl["var"])[0], rtol=1e-10)
    g.A[0, 1] = -np.inf                                                   # in place: same object, same shape
    g.mu[3, 1, 2] += 0.5
    mdl["A"][0, 1] = -np.inf
    mdl["mu"][3, 1, 2] += 0.5
    with np.errstate(all="ignore"):
        p_ref = O.calc_prob(x, mdl["pi"], mdl["A"], mdl["w"], mdl["mu"], mdl["var"])
    p1 = g.calc_prob(x)
    assert p1 != p0 and abs(p1 - p_ref) < 1e-9 * abs(p_ref)
    with np.errstate(all="ignore"):
        path_ref = O.decode(x, mdl["pi"], mdl["A"], mdl["w"], mdl["mu"], mdl["var"])
    assert np.array_equal(g.decode(x), path_ref)
    alpha = g._forward(x)                                                 # no remap: uses the emissions decode() left behind
    assert abs(O.lse(alpha[-1]) - p1) < 1e-9 * abs(p1)


def test_empty_shard_still_enters_the_collective():
    """A rank whose utterance shard is empty (world_size > number of utterances) must call the all-reduce with a zero
    accumulator instead of failing before it -- the other ranks would hang.  The reduce hook stands in for the collective."""
    import torch
    from lnasr.gmmhmm import GMMHMM
    from oracle import lnasr_oracle as O
    rng = np.random.default_rng(8)
    n, mix, d = 3, 2, 5
    mdl = O.random_gmmhmm(n, mix, d, rng)
    utts = [0.6 * rng.standard_normal((T, d)) for T in (20, 31)]

    def model():
        return GMMHMM(n, mix, d, A=mdl["A"], pi=mdl["pi"], w=mdl["w"], mu=mdl["mu"], si=mdl["var"][..., None] * np.eye(d))

    seen = []
    full = model()
    full.train_batch(utts, iters=1, verbose=False, reduce_fn=lambda t: seen.append(t.clone()) or t)
    acc_full, ll_full = seen[0], seen[1]                                  # the accumulator, then the log-likelihood sum of the new model
    calls = []

    def fake_allreduce(t):                                                # what NCCL would deliver: the sum over ranks
        calls.append(tuple(t.shape))
        assert float(t.abs().sum()) == 0.0                                # this rank contributes nothing
        t.copy_(acc_full if t.numel() == acc_full.numel() else ll_full)
        return t

    empty = model()
    hist = empty.train_batch([], iters=1, verbose=False, reduce_fn=fake_allreduce)
    assert len(calls) == 2 and len(hist) == 1
    np.testing.assert_allclose(empty.mu, full.mu, rtol=1e-12)
    np.testing.assert_allclose(empty.A, full.A, rtol=1e-12)
    assert empty.calc_prob_batch([]).shape == (0,) and empty.decode_batch([]) == []
