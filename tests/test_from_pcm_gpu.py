"""End-to-end parity FROM THE PCM: the oracle runs its own chain  O.mfcc -> O.map_b -> O.forward / O.viterbi  on the raw
samples (float64 all the way, mfcc.py:108-175 -> gmmhmm.py:46-67 -> hmm.py:63-100, 133-172, 268-283) and the GPU runs
AcousticPipeline on the same samples.  Nothing the GPU produced is fed to the oracle, so every stage's error (float32
feature storage, tensor-core emissions, float32 emission storage) lands in the compared numbers.

Bars (BASELINE.json): words and Viterbi state paths bit-exact, ties aside; total log-probability within 1e-3 absolute per
utterance.  The measured errors are printed (run with -s)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LOGPROB_TOL = 1e-3          # BASELINE.json: total log-probability within 1e-3 absolute per utterance
NEAR_TIE_TOL = 1e-4         # a path that differs from the oracle's must score within this of it under the ORACLE's emissions


def _utterances(seed, lens):
    rng = np.random.default_rng(seed)
    sigs = []
    for n in lens:
        t = np.arange(n) / 16000.0
        x = np.round(3000.0 * rng.standard_normal(n)) + 2000.0 * np.sin(2 * np.pi * rng.uniform(80, 400) * t)
        sigs.append(np.clip(x, -32767, 32767).astype(np.int16))
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    return np.concatenate(sigs), off, sigs


def _oracle_chain(sig, mdl):
    """PCM -> (features, b [N,T], log-prob, Viterbi score, path) in float64 with the oracle only."""
    from oracle import lnasr_oracle as O
    with np.errstate(all="ignore"):
        _, _, F = O.mfcc(sig)
        b, _ = O.map_b(F, mdl["w"], mdl["mu"], mdl["var"])
        lp = float(O.lse(O.forward(mdl["pi"], mdl["A"], b)[-1]))
        v, _, path = O.viterbi(mdl["pi"], mdl["A"], b)
    return F, b, lp, float(np.max(v[-1])), path


def _check_path(got, ref, mdl, b, what):
    """bit-exact, ties aside: a different path is accepted only if it scores like the oracle's under the oracle's emissions."""
    from helpers import path_score
    got, ref = np.asarray(got, dtype=np.int64), np.asarray(ref, dtype=np.int64)
    if np.array_equal(got, ref):
        return 0
    with np.errstate(all="ignore"):
        s_got, s_ref = path_score(got, mdl["pi"], mdl["A"], b), path_score(ref, mdl["pi"], mdl["A"], b)
    assert np.isfinite(s_got) and abs(s_got - s_ref) <= NEAR_TIE_TOL, (
        f"{what}: paths differ in {int((got != ref).sum())} frames and are not near-ties: {s_got!r} vs {s_ref!r}")
    return int((got != ref).sum())


@pytest.mark.parametrize("emission", ["tc16", "tc"])
def test_config4_model_scale_from_pcm(emission):
    """BASELINE configs[3] model (bench.make_models('c4'): 2000 states x 16 mixtures x 39 dims, left-to-right, means and
    variances at MFCC scale -- |b| ~ 170 per frame) on three utterances of 50-100 frames."""
    import bench
    from lnasr.bank import AcousticPipeline, ModelBank
    from lnasr.mfcc import MFCC
    mdl = bench.make_models("c4")[0]
    lens = [240 + 160 * 60, 240 + 160 * 81, 240 + 160 * 100]            # 60, 81 and 100 frames
    pcm, off, sigs = _utterances(404, lens)
    pipe = AcousticPipeline(MFCC(), ModelBank([mdl], emission=emission))
    res, mb = pipe.recognize_batch(pcm, off, want_paths=True, want_logprob=True)
    foff = mb.frame_offsets_host
    feats = mb.feats.double().cpu().numpy()
    b_gpu = pipe.bank.emissions(mb.feats).double().cpu().numpy()
    paths = res.paths.cpu().numpy()[0]
    worst_lp = worst_v = 0.0
    for u, sig in enumerate(sigs):
        F, b, lp, vs, path = _oracle_chain(sig, mdl)
        T = F.shape[0]
        assert foff[u + 1] - foff[u] == T == [60, 81, 100][u]
        got_f = feats[foff[u]:foff[u + 1]]
        scale = np.sqrt(np.mean(F ** 2, axis=0))
        assert np.all(np.abs(got_f - F) <= 1e-4 * (np.abs(F) + scale)), "MFCC features from PCM (per-column scale)"
        bg = b_gpu[foff[u]:foff[u + 1]].T
        e_b = np.abs(bg - b)
        assert np.max(e_b / np.abs(b)) < 1e-4, "emissions from PCM: 1e-4 relative"
        d_lp = float(res.logprob[u, 0]) - lp
        d_v = float(res.scores[u, 0]) - vs
        n_diff = _check_path(paths[foff[u]:foff[u + 1]], path, mdl, b, f"{emission} utt {u}")
        print(f"c4/{emission} utt {u} T={T}: |b| mean {np.abs(b).mean():.1f}, emission err max {e_b.max():.2e} mean {e_b.mean():.2e}; "
              f"log-prob err {d_lp:+.3e}, Viterbi score err {d_v:+.3e}, path frames differing {n_diff}")
        worst_lp, worst_v = max(worst_lp, abs(d_lp)), max(worst_v, abs(d_v))
        assert abs(d_lp) <= LOGPROB_TOL and abs(d_v) <= LOGPROB_TOL
    # per-frame error rate -> what a 300-frame utterance of the bench would accumulate
    print(f"c4/{emission}: worst |dlogprob| {worst_lp:.3e}, worst |dscore| {worst_v:.3e} (bar {LOGPROB_TOL})")


def test_config4_three_hundred_frames_from_pcm():
    """The bench's own utterance length (3 s = 300 frames) at the bench's model scale, tensor-core (FP16-slice) emissions:
    |dlogprob| <= 1e-3 on the whole utterance."""
    import bench
    from lnasr.bank import AcousticPipeline, ModelBank
    from lnasr.mfcc import MFCC
    mdl = bench.make_models("c4")[0]
    pcm, off, sigs = _utterances(405, [48000])
    pipe = AcousticPipeline(MFCC(), ModelBank([mdl], emission="tc16"))
    res, mb = pipe.recognize_batch(pcm, off, want_paths=True, want_logprob=True)
    F, b, lp, vs, path = _oracle_chain(sigs[0], mdl)
    assert F.shape[0] == 299 or F.shape[0] == 300
    d_lp, d_v = float(res.logprob[0, 0]) - lp, float(res.scores[0, 0]) - vs
    n_diff = _check_path(res.paths.cpu().numpy()[0], path, mdl, b, "c4 300 frames")
    print(f"c4/tc16 T={F.shape[0]}: log-prob {lp:.4f} err {d_lp:+.3e}, Viterbi score err {d_v:+.3e}, path frames differing {n_diff}")
    assert abs(d_lp) <= LOGPROB_TOL and abs(d_v) <= LOGPROB_TOL


def test_config2_from_pcm_sixty_four_utterances():
    """BASELINE configs[1] (10 words x 5 states x 4 mixtures, float64 emissions): 64 utterances of 0.5-1.5 s; words, paths of
    the winning model, Viterbi scores and forward log-probabilities of ALL ten models against the float64 oracle chain."""
    import bench
    from lnasr.bank import AcousticPipeline, ModelBank
    from lnasr.mfcc import MFCC
    from oracle import lnasr_oracle as O
    models = bench.make_models("c2")
    rng = np.random.default_rng(64)
    lens = [int(x) for x in rng.integers(8000, 24001, size=64)]
    pcm, off, sigs = _utterances(202, lens)
    pipe = AcousticPipeline(MFCC(), ModelBank(models, emission="exact"))
    res, mb = pipe.recognize_batch(pcm, off, want_paths=True, want_logprob=True)
    foff = mb.frame_offsets_host
    words, scores, logprob = res.words.cpu().numpy(), res.scores.cpu().numpy(), res.logprob.cpu().numpy()
    paths = res.paths.cpu().numpy()
    worst_lp = worst_v = 0.0
    n_path_diff = 0
    for u, sig in enumerate(sigs):
        with np.errstate(all="ignore"):
            _, _, F = O.mfcc(sig)
            w_ref, s_ref, _, p_ref, paths_ref = O.score_words(F, models)
        assert words[u] == w_ref, f"utt {u}: word {words[u]} != {w_ref}"
        worst_v = max(worst_v, float(np.max(np.abs(scores[u] - s_ref))))
        worst_lp = max(worst_lp, float(np.max(np.abs(logprob[u] - p_ref))))
        for k in (w_ref, (w_ref + 3) % 10):
            b, _ = O.map_b(F, models[k]["w"], models[k]["mu"], models[k]["var"])
            n_path_diff += _check_path(paths[k, foff[u]:foff[u + 1]], paths_ref[k], models[k], b, f"c2 utt {u} model {k}")
    print(f"c2 from PCM, 64 utterances x 10 models: worst |dlogprob| {worst_lp:.3e}, worst |dViterbi score| {worst_v:.3e}, "
          f"path frames differing (near-ties) {n_path_diff}")
    assert worst_lp <= LOGPROB_TOL and worst_v <= LOGPROB_TOL


def test_float32_feature_storage_cost():
    """The features cross from the MFCC kernels to the emission kernels as float32.  Score the SAME utterances from the oracle's
    float64 features rounded to float32 and from the unrounded float64 features (float64 emission kernel takes both): the
    difference in total log-probability is the cost of the float32 hand-over."""
    import torch
    import bench
    from lnasr.bank import ModelBank
    from oracle import lnasr_oracle as O
    models = bench.make_models("c2")
    _, _, sigs = _utterances(203, [16000, 24000, 9000])
    bank = ModelBank(models, emission="exact")
    worst = 0.0
    for sig in sigs:
        _, _, F = O.mfcc(sig)
        T = F.shape[0]
        f_off = torch.tensor([0, T], dtype=torch.int64, device="cuda")
        r64 = bank.score(torch.from_numpy(F).cuda(), f_off, 1, T, want_logprob=True)
        r32 = bank.score(torch.from_numpy(F.astype(np.float32)).cuda(), f_off, 1, T, want_logprob=True)
        worst = max(worst, float((r64.logprob - r32.logprob).abs().max()))
    print(f"float32 feature storage: worst |dlogprob| {worst:.3e} per utterance")
    assert worst < 1e-4
