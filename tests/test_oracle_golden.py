"""Pin the numpy oracle (oracle/lnasr_oracle.py) to fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from oracle import lnasr_oracle as O

TIGHT = dict(rtol=1e-12, atol=1e-12)


def test_lse_matches_reference_bitwise(golden):
    g = golden("utils")
    a3 = g["a3"]
    assert np.array_equal(O.lse(a3), g["lse_all"])
    assert np.array_equal(O.lse(a3, 0), g["lse_ax0"])
    assert np.array_equal(O.lse(a3, 2), g["lse_ax2"])
    assert np.array_equal(O.lse(a3, (0, 2)), g["lse_ax02"])
    assert np.all(O.lse(np.zeros((0, 3)), 0) == -np.inf)


def test_framing_and_window(golden):
    g = golden("utils")
    assert np.array_equal(O.create_hamming(400), g["hamming400"])
    assert np.array_equal(O.split_frames(np.arange(1000.0), 400, 160), g["frames_1000"])
    assert np.array_equal(O.split_frames(np.arange(100.0), 400, 160), g["frames_100"])
    for s, f in zip(g["nframes_S"], g["nframes_F"]):
        assert O.num_frames(int(s)) == int(f)


@pytest.mark.parametrize("tag", ["", "_s401", "_s1000", "_s16000", "_s48017", "_float", "_zero"])
def test_mfcc_matches_reference(golden, tag):
    g = golden("mfcc")
    pcm = g["pcm_data_raw"] if tag == "" else g["pcm" + tag]
    P, C, F = O.mfcc(pcm)
    assert np.array_equal(P, g["power" + tag])
    np.testing.assert_allclose(C, g["ceps" + tag], **TIGHT)
    np.testing.assert_allclose(F, g["feats" + tag], **TIGHT)


def test_mfcc_fingerprints_from_survey(golden):
    """SURVEY.md 8(c) fingerprints recorded from the reference on test/data.raw."""
    g = golden("mfcc")
    P, C, F = O.mfcc(g["pcm_data_raw"])
    assert P.shape == (135, 257) and C.shape == (135, 40) and F.shape == (135, 39)
    np.testing.assert_allclose(F[0, :5], [48.82825422, -13.27565024, 11.73653465, 28.20323954, 27.03861366], atol=1e-7)
    np.testing.assert_allclose(F[134, 12], 14.1284645622, atol=1e-9)
    np.testing.assert_allclose(F.sum(), 2389.7636666679, atol=1e-5)
    assert np.array_equal(O.mel_filter_bank(), g["fbank"])
    assert np.count_nonzero(g["fbank"]) == 453
    np.testing.assert_allclose(O.dct2_ortho_matrix(40) @ np.arange(40.0), O._dct2_ortho(np.arange(40.0)[None])[0], atol=1e-10)


def test_icecream_hmm(golden):
    """test/hmm-test.py:10-39 known-answer values."""
    g = golden("hmm_discrete")
    b = O.hmm_map_b(g["B"], g["obs"])
    alpha = O.forward(g["pi"], g["A"], b)
    beta = O.backward(g["A"], b)
    v, bt, path = O.viterbi(g["pi"], g["A"], b)
    xi, gamma = O.baumwelch(g["A"], b, alpha, beta)
    est = O.estimate_hmm(g["obs"], 3, xi, gamma)
    assert np.array_equal(alpha, g["alpha"]) and np.array_equal(beta, g["beta"])
    assert np.array_equal(v, g["v"]) and np.array_equal(bt, g["bt"]) and np.array_equal(path, g["path"])
    assert np.array_equal(xi, g["xi"]) and np.array_equal(gamma, g["gamma"])
    np.testing.assert_allclose(est["A"], g["A_new"], **TIGHT)
    np.testing.assert_allclose(est["B"], g["B_new"], **TIGHT)
    assert np.array_equal(est["pi"], g["pi_new"])
    np.testing.assert_allclose(np.exp(alpha), [[.32, .02], [.0404, .069], [.023496, .005066]], atol=1e-9)
    np.testing.assert_allclose(np.exp(O.lse(alpha[-1])), 0.028562, atol=1e-9)
    assert list(path) == [0, 1, 0]
    assert np.all(gamma[2] == -np.inf)


def test_discrete_hmm_training(golden):
    g = golden("hmm_discrete")
    A, B, pi, obs = g["d_A0"], g["d_B0"], g["d_pi0"], g["d_obs"]
    for _ in range(5):
        b = O.hmm_map_b(B, obs)
        alpha, beta = O.forward(pi, A, b), O.backward(A, b)
        xi, gamma = O.baumwelch(A, b, alpha, beta)
        m = O.estimate_hmm(obs, 5, xi, gamma)
        old = O.lse(alpha[-1])
        A, B, pi = m["A"], m["B"], m["pi"]
        new = O.lse(O.forward(pi, A, O.hmm_map_b(B, obs))[-1])
        if abs(old - new) < 1e-4:
            break
    np.testing.assert_allclose(A, g["d_A5"], **TIGHT)
    np.testing.assert_allclose(B, g["d_B5"], **TIGHT)
    np.testing.assert_allclose(new, g["d_prob5"], **TIGHT)
    assert np.array_equal(O.viterbi(pi, A, O.hmm_map_b(B, obs))[2], g["d_path5"])


def test_gmmhmm_5_4_39(golden):
    g = golden("gmmhmm_5_4_39")
    obs = g["obs"]
    b, bm = O.map_b(obs, g["w"], g["mu"], g["si"])
    np.testing.assert_allclose(bm, g["bm"], rtol=1e-12, atol=1e-10)
    np.testing.assert_allclose(b, g["b"], rtol=1e-12, atol=1e-10)
    alpha, beta = O.forward(g["pi"], g["A"], b), O.backward(g["A"], b)
    np.testing.assert_allclose(alpha, g["alpha"], rtol=1e-12)
    np.testing.assert_allclose(beta, g["beta"], rtol=1e-12)
    np.testing.assert_allclose(O.lse(alpha[-1]), -5147.0613398313, atol=1e-8)
    v, bt, path = O.viterbi(g["pi"], g["A"], b)
    assert np.array_equal(path, g["path"]) and np.array_equal(bt, g["bt"])
    assert list(path[:20]) == [2, 3, 3, 3, 3, 3, 4, 1, 3, 3, 4, 1, 3, 2, 3, 3, 2, 3, 3, 2]
    np.testing.assert_allclose(np.max(v[-1]), -5291.9737418548, atol=1e-8)
    xi, gamma = O.baumwelch(g["A"], b, alpha, beta)
    est = O.estimate_gmmhmm(obs, alpha, beta, xi, gamma, g["w"], g["mu"], bm)
    for k in ("A", "pi", "w", "mu", "si"):
        np.testing.assert_allclose(est[k], g["est_" + k], rtol=1e-9, atol=1e-12)


def test_gmmhmm_diag_and_sufficient_statistics(golden):
    """Diagonal restatement == reference full-covariance code on diagonal Sigma; the linear-domain
    sufficient statistics (what the CUDA path accumulates) reproduce the reference M-step."""
    g = golden("gmmhmm_5_4_39")
    obs, var = g["obs"], g["dg_var"]
    b, bm = O.map_b(obs, g["w"], g["mu"], var)
    np.testing.assert_allclose(bm, g["dg_bm"], rtol=1e-12, atol=1e-10)
    np.testing.assert_allclose(b, g["dg_b"], rtol=1e-12, atol=1e-10)
    acc = O.em_accumulate_diag(obs, g["pi"], g["A"], g["w"], g["mu"], var)
    np.testing.assert_allclose(acc["loglik"], g["dg_prob"], rtol=1e-13)
    new = O.em_finalize_diag(acc)
    np.testing.assert_allclose(new["A"], g["dg_est_A"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(new["pi"], g["dg_est_pi"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(new["w"], g["dg_est_w"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(new["mu"], g["dg_est_mu"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(new["var"], g["dg_est_var"], rtol=1e-9, atol=1e-12)
    assert np.array_equal(O.viterbi(g["pi"], g["A"], b)[2], g["dg_path"])


def test_config1_training(golden):
    """test/gmmhmm-test.py shapes (n=5,m=3,d=2,T=40), seeded: one reference iteration, and ten
    iterations of the diagonal-projected harness."""
    g = golden("gmmhmm_config1")
    obs = g["obs"]
    m0 = dict(A=g["A0"], pi=g["pi0"], w=g["w0"], mu=g["mu0"], si=g["si0"])
    m1, _ = O.train_gmmhmm(obs, m0, iters=1)
    for k in ("A", "pi", "w", "mu", "si"):
        np.testing.assert_allclose(m1[k], g[k + "1"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(O.calc_prob(obs, m1["pi"], m1["A"], m1["w"], m1["mu"], m1["si"]), g["prob1"], rtol=1e-12)
    md, _ = O.train_gmmhmm(obs, m0, iters=10, eps=0.0, project_diag=True)
    np.testing.assert_allclose(md["mu"], g["dmu10"], rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(np.diagonal(md["si"], axis1=2, axis2=3), g["dvar10"], rtol=1e-8)
    np.testing.assert_allclose(md["A"], g["dA10"], rtol=1e-8, atol=1e-10)
    # the sufficient-statistic path must walk the same 10 iterations
    mdl = dict(A=g["A0"], pi=g["pi0"], w=g["w0"], mu=g["mu0"], var=np.diagonal(g["si0"], axis1=2, axis2=3).copy())
    for _ in range(10):
        mdl, _ = O.em_step_multi_diag([obs], mdl)
    np.testing.assert_allclose(mdl["mu"], g["dmu10"], rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(mdl["var"], g["dvar10"], rtol=1e-7)
    np.testing.assert_allclose(mdl["w"], g["dw10"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(mdl["pi"], g["dpi10"], rtol=1e-7, atol=1e-9)


def test_gauss_density(golden):
    g = golden("gmm")
    np.testing.assert_allclose(O.gauss_logpdf_full(g["x"], g["mu"], g["Sigma"]), g["gmlog"], **TIGHT)
    np.testing.assert_allclose(np.exp(O.gauss_logpdf_full(g["x"], g["mu"], g["Sigma"])), g["gm"], **TIGHT)


def test_port_agrees_with_the_staged_reference_in_a_worker():
    """oracle/_ref/lnasr_reference.zip (the unmodified reference staged by oracle/build_ref.py, what `bench.py --impl reference`
    times) against the numpy port, from PCM, in a separate process (the reference must be imported as `lnasr` there)."""
    import subprocess
    import sys
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference archive not staged (run python oracle/build_ref.py where /root/reference exists)")
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r)
np.seterr(all="ignore")
from oracle import ref_loader, lnasr_oracle as O
R = ref_loader.load()
rng = np.random.default_rng(5)
x = np.clip(np.round(3000 * rng.standard_normal(9000)), -32767, 32767).astype(np.int16)
F = R["mfcc"].MFCC().calc_mfcc(x)[2]
assert np.array_equal(F, O.mfcc(x)[2])
m = O.random_gmmhmm(6, 3, 39, rng)
m["mu"] *= 100.0; m["var"] *= 400.0
si = m["var"][..., None] * np.eye(39)
g = R["gmmhmm"].GMMHMM(6, 3, 39, A=m["A"], pi=m["pi"], w=m["w"], mu=m["mu"], si=si)
v, bt, path = g._viterbi(F, True)
b, _ = O.map_b(F, m["w"], m["mu"], m["var"])
v2, bt2, path2 = O.viterbi(m["pi"], m["A"], b)
assert np.array_equal(path, path2) and np.allclose(v, v2, rtol=1e-10, atol=1e-8)
assert abs(g.calc_prob(F) - O.lse(O.forward(m["pi"], m["A"], b)[-1])) < 1e-8
print("ok")
''' % ROOT
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "ok" in p.stdout, p.stderr[-2000:]


@pytest.mark.parametrize("tag,kw", [("a", {}), ("b", {"alpha": 0.9, "threshold": 30.0}),
                                    ("c", {"winsize": 512, "stepsize": 256, "order": 3}), ("d", {})])
def test_vad_ltsd_oracle_matches_reference_golden(golden, tag, kw):
    """O.vad_ltsd against VadLtsd.detect of the unmodified reference (third/vadlstd/VadLstd.py:44-74), incl. noise adaptation."""
    g = golden("vad")
    got = O.vad_ltsd(g["pcm_" + tag], **kw)
    assert got.shape == g["ltsd_" + tag].shape
    np.testing.assert_allclose(got, g["ltsd_" + tag], rtol=1e-12, atol=1e-12)
