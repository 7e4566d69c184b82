"""The other consumer of lnasr.hmm.HMM in the reference: scripts/seg/seg.py:102-157, a 4-state (B/M/E/S) character tagger over
m = 65 536 symbols (ord(char)), trained by COUNTING tagged data and decoded with HMM.decode.  SURVEY.md 8(f) N3: the same
trellis kernels with an emission gather B[:, obs[t]] instead of the GMM, at the `Seg` shape.

`Seg.train` as written in the reference calls self.reset() without the argument HMM.reset requires (seg.py:138 against
hmm.py:233) and leaves A, B, pi as linear probabilities although the HMM works in the log domain; the class below keeps
its structure (same state map, same counting, same add-one smoothing of B) with explicit zero arrays and a final log."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

STATES = {"B": 0, "M": 1, "E": 2, "S": 3}


def _seg_class():
    from lnasr.hmm import HMM

    class Seg(HMM):
        def __init__(self, A=None, B=None, pi=None):
            super().__init__(len(STATES), 65536, A, B, pi, np.double)          # seg.py:113-114

        def _map_data(self, data):
            return np.array([ord(c) for c in data])                           # seg.py:116-118

        def train(self, datas):                                               # seg.py:133-151, counting
            n, m = self.n, self.m
            A, B, pi = np.zeros((n, n)), np.zeros((n, m)), np.zeros(n)
            for ds in datas:
                data = self._map_data(ds["data"])
                state = np.array([STATES[s] for s in ds["state"]])
                for k in range(len(data) - 1):
                    A[state[k], state[k + 1]] += 1
                for k in range(len(data)):
                    B[state[k], data[k]] += 1
                pi[state[0]] += 1
            A /= A.sum(axis=1, keepdims=True)
            B += 1
            B /= B.sum(axis=1, keepdims=True)
            pi /= pi.sum()
            with np.errstate(divide="ignore"):
                self.A, self.B, self.pi = np.log(A), np.log(B), np.log(pi)

        def test(self, text):                                                 # seg.py:153-158
            return self.decode(self._map_data(text))
    return Seg


def _tag(words):
    data, state = "", ""
    for w in words:
        data += w
        state += "S" if len(w) == 1 else "B" + "M" * (len(w) - 2) + "E"
    return {"data": data, "state": state}


def _corpus(rng, n_sent):
    # synthetic "words" over CJK code points (0x4e00..0x9fff): a small lexicon with Zipf-like use
    lex = ["".join(chr(int(c)) for c in rng.integers(0x4e00, 0x9fff, size=int(L))) for L in rng.choice([1, 2, 2, 3, 4], size=300)]
    p = 1.0 / np.arange(1, len(lex) + 1)
    p /= p.sum()
    return [[lex[i] for i in rng.choice(len(lex), size=int(rng.integers(3, 12)), p=p)] for _ in range(n_sent)]


def test_seg_shape_counting_and_decode_against_oracle():
    from oracle import lnasr_oracle as O
    from helpers import assert_paths_equal_ties_aside
    rng = np.random.default_rng(65536)
    sents = _corpus(rng, 400)
    seg = _seg_class()()
    seg.train(_tag(s) for s in sents[:300])
    assert seg.A.shape == (4, 4) and seg.B.shape == (4, 65536)
    assert np.isneginf(seg.A[0, 0]) and np.isneginf(seg.A[2, 1]) and np.isneginf(seg.pi[1])     # B->B, E->M, start in M never seen
    texts = ["".join(s) for s in sents[300:]]
    paths = seg.decode_batch([seg._map_data(t) for t in texts])                # 100 sentences, one launch
    n_exact = 0
    with np.errstate(all="ignore"):
        for t, p in zip(texts, paths):
            obs = seg._map_data(t)
            b = O.hmm_map_b(seg.B, obs)
            v, _, p_ref = O.viterbi(seg.pi, seg.A, b)
            assert_paths_equal_ties_aside(p, p_ref, seg.pi, seg.A, b, what=t)
            n_exact += int(np.array_equal(p, p_ref))
    assert n_exact >= 95
    assert np.array_equal(seg.test(texts[0]), paths[0])                        # the reference's single-sentence entry point
    # words of the training lexicon are recovered: the tags cut sentence 0 where its words end
    tags = "".join("BMES"[k] for k in seg.test("".join(sents[0])))
    assert tags == _tag(sents[0])["state"]


def test_seg_shape_baum_welch_step_against_oracle():
    """One Baum-Welch step (HMM.train, hmm.py:285-313) at n = 4, m = 65 536: the B statistics are a [4, 65 536] scatter."""
    from lnasr.hmm import HMM
    from oracle import lnasr_oracle as O
    from helpers import assert_close
    rng = np.random.default_rng(7)
    np.random.seed(9)
    h = HMM(4, 65536)
    h.reset("random")
    A, B, pi = h.A.copy(), h.B.copy(), h.pi.copy()
    obs = rng.integers(0x4e00, 0x4e40, size=120)                                # 64 distinct symbols, each seen about twice
    b = O.hmm_map_b(B, obs)
    with np.errstate(all="ignore"):
        alpha, beta = O.forward(pi, A, b), O.backward(A, b)
        xi, gamma = O.baumwelch(A, b, alpha, beta)
        denom = O.lse(gamma, axis=0)
        A_ref = O.lse(xi, axis=0) - denom[:, None]
        seen = np.unique(obs)
        B_ref_seen = np.stack([O.lse(gamma[obs == k, :], axis=0) - denom for k in seen], axis=1)
    h.train(obs, iters=1, verbose=False)
    with np.errstate(all="ignore"):
        assert_close(np.exp(h.A), np.exp(A_ref), 1e-9, "A", per_column=False)
        assert_close(np.exp(h.B[:, seen]), np.exp(B_ref_seen), 1e-9, "B seen", per_column=False)
        assert_close(np.exp(h.pi), np.exp(gamma[0]), 1e-9, "pi")
    unseen = np.setdiff1d(np.arange(65536), seen)
    assert np.isneginf(h.B[:, unseen]).all()                                   # lse over an empty selection = -inf (utils.py:169-170)
