"""Trellis kernels (csrc/trellis.cu) through the reference API: forward, backward, Viterbi + backtrace.

Bars (BASELINE.json): state paths bit-exact, ties aside; total log-probability within 1e-3 absolute per utterance
(met here at ~1e-9: the trellis state is float64)."""
import numpy as np
import pytest

from helpers import assert_close, assert_paths_equal_ties_aside

pytestmark = pytest.mark.gpu


def test_icecream_hmm_known_answers(golden):
    """test/hmm-test.py:10-39 values recorded from the reference."""
    from lnasr.hmm import HMM
    g = golden("hmm_discrete")
    h = HMM(2, 3, g["A"], g["B"], g["pi"])
    alpha = h._forward(g["obs"], True)
    beta = h._backward(g["obs"])
    v, bt, path = h._viterbi(g["obs"])
    assert np.array_equal(h.b, g["B"][:, g["obs"]])
    np.testing.assert_allclose(alpha, g["alpha"], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(beta, g["beta"], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(v, g["v"], rtol=1e-13, atol=1e-13)
    assert bt.dtype == np.uint32 and path.dtype == np.uint32
    assert np.array_equal(bt, g["bt"]) and np.array_equal(path, g["path"]) and list(path) == [0, 1, 0]
    np.testing.assert_allclose(np.exp(alpha), [[.32, .02], [.0404, .069], [.023496, .005066]], atol=1e-9)
    np.testing.assert_allclose(np.exp(h.calc_prob(g["obs"])), 0.028562, atol=1e-9)
    assert np.array_equal(h.decode(g["obs"]), g["decode"])
    xi, gamma = h._baumwelch(g["obs"], alpha, beta)
    assert_close(xi, g["xi"], 1e-12, "xi")
    assert_close(gamma, g["gamma"], 1e-12, "gamma")
    assert np.all(gamma[2] == -np.inf)


def test_gmmhmm_golden_trellis(golden):
    from lnasr.gmmhmm import GMMHMM
    g = golden("gmmhmm_5_4_39")
    m = GMMHMM(5, 4, 39, A=g["A"], pi=g["pi"], w=g["w"], mu=g["mu"], si=g["si"])
    obs = g["obs"]
    assert abs(m.calc_prob(obs) - float(g["prob"])) < 1e-8
    assert abs(m.calc_prob(obs) - (-5147.0613398313)) < 1e-6                     # SURVEY.md 8(c) fingerprint
    alpha = m._forward(obs, True)
    beta = m._backward(obs)
    v, bt, path = m._viterbi(obs)
    np.testing.assert_allclose(alpha, g["alpha"], rtol=1e-11, atol=1e-9)
    np.testing.assert_allclose(beta, g["beta"], rtol=1e-11, atol=1e-9)
    np.testing.assert_allclose(v, g["v"], rtol=1e-11, atol=1e-9)
    assert np.array_equal(bt, g["bt"])
    assert np.array_equal(path, g["path"])
    assert np.array_equal(m.decode(obs), g["path"])
    assert list(m.decode(obs)[:20]) == [2, 3, 3, 3, 3, 3, 4, 1, 3, 3, 4, 1, 3, 2, 3, 3, 2, 3, 3, 2]
    xi, gamma = m._baumwelch(obs, alpha, beta)
    np.testing.assert_allclose(gamma[:-1], g["gamma"][:-1], rtol=1e-9, atol=1e-9)
    assert np.all(gamma[-1] == -np.inf) and xi.shape == (135, 5, 5)


def _random_case(rng, n, T, left_to_right=False, impossible=False):
    from oracle import lnasr_oracle as O
    mdl = O.random_gmmhmm(n, 1, 1, rng, left_to_right=left_to_right)
    b = -5.0 * rng.random((n, T)) - 0.5
    if impossible:
        b[rng.integers(0, n), rng.integers(0, T)] = -np.inf
    return mdl["pi"], mdl["A"], b


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 13, 32, 33, 64, 200])
@pytest.mark.parametrize("ltr", [False, True])
def test_trellis_batch_against_oracle(n, ltr):
    """Ragged batch (T from 1 to 70) against the numpy oracle, small-N (warp shuffle) and general (ELL) kernels."""
    import torch
    from lnasr._trellis import TransitionBank
    from oracle import lnasr_oracle as O
    rng = np.random.default_rng(100 + n + (1000 if ltr else 0))
    pi, A, _ = _random_case(rng, n, 1, ltr)
    if ltr:
        pi = np.full(n, -np.inf)
        pi[0] = 0.0
    lens = [1, 2, 70, 17, 33, 5]
    bs = [-5.0 * rng.random((n, T)) - 0.5 for T in lens]
    bs[3][rng.integers(0, n), 4] = -np.inf
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    b_cat = np.ascontiguousarray(np.concatenate([b.T for b in bs], axis=0))
    bank = TransitionBank(A, pi)
    b_d = torch.from_numpy(np.ascontiguousarray(b_cat)).cuda()
    off_d = torch.from_numpy(off).cuda()
    total = int(off[-1])
    r = bank.viterbi(b_d, off_d, len(lens), total, want_path=True, want_score=True, want_v=True, want_bt=True)
    alpha, lp = bank.forward(b_d, off_d, len(lens), total, want_alpha=True)
    beta = bank.backward(b_d, off_d, len(lens), total)
    path = r["path"][0].cpu().numpy()
    with np.errstate(all="ignore"):
        for u, b in enumerate(bs):
            sl = slice(off[u], off[u + 1])
            v_ref, bt_ref, p_ref = O.viterbi(pi, A, b)
            a_ref = O.forward(pi, A, b)
            be_ref = O.backward(A, b)
            assert_close(r["v"].cpu().numpy()[sl], v_ref, 1e-12, f"v[{u}]")
            assert_close(alpha.cpu().numpy()[sl], a_ref, 1e-11, f"alpha[{u}]")
            assert_close(beta.cpu().numpy()[sl], be_ref, 1e-11, f"beta[{u}]")
            assert_paths_equal_ties_aside(path[sl], p_ref, pi, A, b, what=f"path[{u}]")
            if np.array_equal(path[sl], p_ref):
                assert np.array_equal(r["bt"].cpu().numpy()[sl][np.isfinite(v_ref)], bt_ref[np.isfinite(v_ref)])
            assert_close(r["score"].cpu().numpy()[u, 0], np.max(v_ref[-1]), 1e-12, f"score[{u}]")
            assert_close(lp.cpu().numpy()[u, 0], O.lse(a_ref[-1]), 1e-11, f"logprob[{u}]")
    # float32 emissions (the layout the tensor-core path produces) give the same paths on float32-rounded inputs
    b32 = b_cat.astype(np.float32)
    r32 = bank.viterbi(torch.from_numpy(b32).cuda(), off_d, len(lens), total, want_path=True, want_score=True)
    p32 = r32["path"][0].cpu().numpy()
    with np.errstate(all="ignore"):
        for u in range(len(lens)):
            sl = slice(off[u], off[u + 1])
            bb = b32[sl].astype(np.float64).T
            assert_paths_equal_ties_aside(p32[sl], O.viterbi(pi, A, bb)[2], pi, A, bb, what=f"path32[{u}]")


@pytest.mark.parametrize("n,skip", [(40, True), (250, False), (1030, True), (1100, True), (2000, True)])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_banded_viterbi_large_n(n, skip, dtype):
    """Left-to-right banks (self / next / skip transitions) take the register-resident banded kernel on the decode path
    (paths + scores, no v / bt outputs): 4 and 8 states per thread, state counts that do not fill the last thread, both
    emission dtypes, ragged lengths from 1 frame, an impossible emission, and an entry state other than 0."""
    import torch
    from lnasr._trellis import TransitionBank
    from oracle import lnasr_oracle as O
    rng = np.random.default_rng(7 * n + (1 if skip else 0))
    mdl = O.random_gmmhmm(n, 1, 1, rng, left_to_right=True)
    A = mdl["A"].copy()
    if not skip:
        for i in range(n - 2):
            A[i, i + 2] = -np.inf
    pi = np.full(n, -np.inf)
    pi[0] = np.log(0.7)
    pi[3] = np.log(0.3)
    lens = [1, 2, 45, 9, 30]
    bs = [(-5.0 * rng.random((n, T)) - 0.5).astype(dtype).astype(np.float64) for T in lens]
    bs[2][rng.integers(0, n), 7] = -np.inf
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    b_cat = np.ascontiguousarray(np.concatenate([b.T for b in bs], axis=0)).astype(dtype)
    bank = TransitionBank(A, pi)
    assert bank.band == (3 if skip else 2)
    off_d = torch.from_numpy(off).cuda()
    total = int(off[-1])
    r = bank.viterbi(torch.from_numpy(b_cat).cuda(), off_d, len(lens), total, want_path=True, want_score=True)
    path = r["path"][0].cpu().numpy()
    score = r["score"].cpu().numpy()
    # the general kernel (API-parity outputs requested) must agree bit for bit on the scores
    r_gen = bank.viterbi(torch.from_numpy(b_cat).cuda(), off_d, len(lens), total, want_path=True, want_score=True, want_v=True)
    assert np.array_equal(score, r_gen["score"].cpu().numpy())
    with np.errstate(all="ignore"):
        for u, b in enumerate(bs):
            sl = slice(off[u], off[u + 1])
            v_ref, _, p_ref = O.viterbi(pi, A, b)
            assert_paths_equal_ties_aside(path[sl], p_ref, pi, A, b, what=f"path[{u}]")
            assert_close(score[u, 0], np.max(v_ref[-1]), 1e-12, f"score[{u}]")


def test_exact_ties_take_lowest_index():
    """np.argmax semantics (hmm.py:166,169): uniform model, constant emissions -> every argmax is index 0."""
    from lnasr.hmm import HMM
    n, T = 6, 12
    h = HMM(n, 2)
    h.reset("uniform")
    v, bt, path = h._viterbi(np.zeros(T, dtype=np.int64), True)
    assert np.all(bt == 0) and np.all(path == 0)


def test_model_bank_scores_and_words():
    """Isolated-word decode (BASELINE config 2 shapes): 10 word models x (5 states, 4 mixtures, 39 dims)."""
    import torch
    from lnasr.bank import ModelBank
    from oracle import lnasr_oracle as O
    rng = np.random.default_rng(77)
    models = [O.random_gmmhmm(5, 4, 39, rng) for _ in range(10)]
    lens = [50, 149, 2, 80, 101]
    utts = [0.5 * rng.standard_normal((T, 39)) for T in lens]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    bank = ModelBank(models, emission="exact")
    feats = torch.from_numpy(np.concatenate(utts)).cuda()
    res = bank.score(feats, torch.from_numpy(off).cuda(), len(lens), int(off[-1]), want_paths=True, want_logprob=True)
    words = res.words.cpu().numpy()
    scores = res.scores.cpu().numpy()
    lps = res.logprob.cpu().numpy()
    paths = res.paths.cpu().numpy()
    for u, x in enumerate(utts):
        w_ref, s_ref, wp_ref, p_ref, paths_ref = O.score_words(x, models)
        np.testing.assert_allclose(scores[u], s_ref, rtol=1e-11)
        np.testing.assert_allclose(lps[u], p_ref, rtol=1e-11)           # well inside 1e-3 absolute
        assert words[u] == w_ref
        for k in range(10):
            b, _ = O.map_b(x, models[k]["w"], models[k]["mu"], models[k]["var"])
            assert_paths_equal_ties_aside(paths[k, off[u]:off[u + 1]], paths_ref[k], models[k]["pi"], models[k]["A"], b)


@pytest.mark.parametrize("n", [64, 1024, 2000, 2048])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_banded_decode_kernel_against_first_form_and_oracle(n, dtype):
    """trellis_band2_kernel (barrier-free warps, 2-bit backpointers, backtrace fused, 64 rows per chunk) against
    trellis_band_kernel + backtrace_kernel bit for bit, and against the oracle: utterance lengths around the backtrace chunk
    size (1, 2, 63, 64, 65, 129, 300 frames), 4 and 8 states per thread, both emission dtypes, a bank of two models, an
    impossible emission and an unreachable stretch of states (backpointer code 3 = np.argmax of an all -inf column)."""
    import torch
    from lnasr import _capi
    from lnasr._trellis import TransitionBank
    from oracle import lnasr_oracle as O
    rng = np.random.default_rng(11 * n)
    K = 2
    As, pis = [], []
    for k in range(K):
        mdl = O.random_gmmhmm(n, 1, 1, rng, left_to_right=True)
        A = mdl["A"].copy()
        if k == 1:
            A[n // 2 - 1, n // 2] = A[n // 2 - 1, n // 2 + 1] = A[n // 2 - 2, n // 2] = -np.inf     # states >= n/2 unreachable
        pi = np.full(n, -np.inf)
        pi[0], pi[2] = np.log(0.6), np.log(0.4)
        As.append(A)
        pis.append(pi)
    lens = [1, 2, 63, 64, 65, 129, 300]
    total = sum(lens)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    b = (-6.0 * rng.random((total, K * n)) - 0.5).astype(dtype)
    b[off[4] + 7, rng.integers(0, n)] = -np.inf
    bank = TransitionBank(np.stack(As), np.stack(pis))
    assert bank.band == 3
    bd, off_d = torch.from_numpy(b).cuda(), torch.from_numpy(off).cuda()
    lib = _capi.lib()
    r2 = bank.viterbi(bd, off_d, len(lens), total, want_path=True, want_score=True)
    try:
        lib.lnasr_debug_set_band_kernel(1)
        r1 = bank.viterbi(bd, off_d, len(lens), total, want_path=True, want_score=True)
    finally:
        lib.lnasr_debug_set_band_kernel(2)
    assert torch.equal(r1["score"], r2["score"]) and torch.equal(r1["path"], r2["path"])
    s_only = bank.viterbi(bd, off_d, len(lens), total, want_path=False, want_score=True)       # no workspace, no backtrace
    assert torch.equal(s_only["score"], r2["score"])
    path, score = r2["path"].cpu().numpy(), r2["score"].cpu().numpy()
    b64 = b.astype(np.float64)
    check = range(len(lens)) if n <= 1024 else (0, 1, 4)                                        # the oracle is slow at N = 2000
    with np.errstate(all="ignore"):
        for k in range(K):
            for u in check:
                bb = b64[off[u]:off[u + 1], k * n:(k + 1) * n].T
                v_ref, _, p_ref = O.viterbi(pis[k], As[k], bb)
                assert_paths_equal_ties_aside(path[k, off[u]:off[u + 1]], p_ref, pis[k], As[k], bb, what=f"model {k} utt {u}")
                assert_close(score[u, k], np.max(v_ref[-1]), 1e-12, f"score[{u},{k}]")
