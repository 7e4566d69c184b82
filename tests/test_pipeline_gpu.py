"""PCM -> word pipeline (lnasr.bank.AcousticPipeline): the staged path, the single-C-call path (lnasr_recognize_batch) and
the host-audio path (chunked, overlapped copies) must agree with each other and with the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _batch(rng, lens):
    sigs = []
    for n in lens:
        t = np.arange(n) / 16000.0
        x = np.round(3000.0 * rng.standard_normal(n)) + 2000.0 * np.sin(2 * np.pi * rng.uniform(80, 400) * t)
        sigs.append(np.clip(x, -32767, 32767).astype(np.int16))
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    return np.concatenate(sigs), off, sigs


def _models(rng, K, n, m):
    from oracle import lnasr_oracle as O
    models = [O.random_gmmhmm(n, m, 39, rng) for _ in range(K)]
    for mdl in models:
        mdl["mu"] = mdl["mu"] * 100.0
        mdl["var"] = mdl["var"] * 400.0
    return models


@pytest.mark.parametrize("K,n,m", [(10, 5, 4), (3, 3, 8), (4, 7, 2)])
def test_three_entry_points_agree_and_match_oracle(K, n, m):
    import torch
    from lnasr.bank import AcousticPipeline, ModelBank
    from lnasr.mfcc import MFCC
    from oracle import lnasr_oracle as O
    rng = np.random.default_rng(K * 100 + n)
    lens = [int(x) for x in rng.integers(4000, 20000, size=37)]
    pcm, off, sigs = _batch(rng, lens)
    models = _models(rng, K, n, m)
    pipe = AcousticPipeline(MFCC(), ModelBank(models, emission="exact"))
    res, mb = pipe.recognize_batch(pcm, off, want_paths=True)
    w_f, b_f, p_f = pipe.recognize_batch_fused(torch.from_numpy(pcm).cuda(), off, want_paths=True)
    assert torch.equal(res.words, w_f) and torch.equal(res.best, b_f) and torch.equal(res.paths, p_f)
    for chunks, streams in ((1, 1), (3, 2), (8, 2), (64, 2)):
        w_h, b_h = pipe.recognize_batch_host(torch.from_numpy(pcm).pin_memory(), off, n_chunks=chunks, want_paths=True,
                                             n_compute_streams=streams)
        assert np.array_equal(w_h.numpy(), res.words.cpu().numpy()), (chunks, streams)
        assert np.array_equal(b_h.numpy(), res.best.cpu().numpy()), (chunks, streams)
    # oracle on a few utterances: same features -> same word and score
    feats = mb.feats.double().cpu().numpy()
    foff = mb.frame_offsets_host
    for u in (0, 5, 36):
        w_ref, s_ref, _, _, _ = O.score_words(feats[foff[u]:foff[u + 1]], models)
        assert int(res.words[u]) == w_ref
        np.testing.assert_allclose(res.scores[u].cpu().numpy(), s_ref, rtol=1e-9)


def test_fused_entry_point_with_tensor_core_emissions():
    import torch
    from lnasr.bank import AcousticPipeline, ModelBank
    from lnasr.mfcc import MFCC
    from oracle import lnasr_oracle as O
    rng = np.random.default_rng(9)
    pcm, off, _ = _batch(rng, [8000, 12000, 5000, 16000])
    mdl = O.random_gmmhmm(64, 16, 39, rng, left_to_right=True)
    mdl["mu"] = mdl["mu"] * 100.0
    mdl["var"] = mdl["var"] * 400.0
    pipe = AcousticPipeline(MFCC(), ModelBank([mdl], emission="tc"))
    res, _ = pipe.recognize_batch(pcm, off, want_paths=True)
    w_f, b_f, p_f = pipe.recognize_batch_fused(torch.from_numpy(pcm).cuda(), off, want_paths=True)
    assert torch.equal(res.words, w_f) and torch.equal(res.best, b_f) and torch.equal(res.paths, p_f)


def test_sixty_second_utterance_and_batch_of_one():
    """BASELINE config 5 extremes through the whole path: one 60 s utterance (5 999 frames) alone in a batch, next to 0.1 s
    ones in another; paths against the oracle's Viterbi on the same emissions."""
    import torch
    from lnasr.bank import AcousticPipeline, ModelBank
    from lnasr.mfcc import MFCC
    from oracle import lnasr_oracle as O
    from helpers import assert_paths_equal_ties_aside
    rng = np.random.default_rng(60)
    models = _models(rng, 3, 5, 4)
    pipe = AcousticPipeline(MFCC(), ModelBank(models, emission="exact"))
    pcm1, off1, _ = _batch(rng, [960000])
    res1, mb1 = pipe.recognize_batch(pcm1, off1, want_paths=True)
    assert int(mb1.frame_offsets_host[-1]) == 5999 and res1.paths.shape == (3, 5999)
    pcm2, off2, _ = _batch(rng, [1600, 960000, 2000, 1700])
    pcm2[off2[1]:off2[2]] = pcm1                                    # the same long utterance inside a ragged batch
    res2, mb2 = pipe.recognize_batch(pcm2, off2, want_paths=True)
    f = mb2.frame_offsets_host
    assert torch.equal(res2.scores[1], res1.scores[0]) and torch.equal(res2.paths[:, f[1]:f[2]], res1.paths)
    w_f, b_f, p_f = pipe.recognize_batch_fused(torch.from_numpy(pcm2).cuda(), off2, want_paths=True)
    assert torch.equal(w_f, res2.words) and torch.equal(p_f, res2.paths)
    # oracle Viterbi of model 1 on the long utterance, from the emissions the GPU produced
    b = pipe.bank.emissions(mb1.feats).cpu().numpy()
    k = 1
    v, bt, path = O.viterbi(models[k]["pi"], models[k]["A"], b[:, 5 * k:5 * k + 5].T.copy())
    assert abs(v[-1].max() - float(res1.scores[0, k])) < 1e-6
    assert_paths_equal_ties_aside(res1.paths[k].cpu().numpy(), path, models[k]["pi"], models[k]["A"], b[:, 5 * k:5 * k + 5].T, what="60 s path")


def test_empty_batch():
    import torch
    from lnasr.bank import AcousticPipeline, ModelBank
    from lnasr.mfcc import MFCC
    rng = np.random.default_rng(1)
    pipe = AcousticPipeline(MFCC(), ModelBank(_models(rng, 2, 3, 2), emission="exact"))
    w, b = pipe.recognize_batch_host(torch.empty((0,), dtype=torch.int16).pin_memory(), np.zeros(1, dtype=np.int64))
    assert w.numel() == 0 and b.numel() == 0
    w, b, p = pipe.recognize_batch_fused(torch.empty((0,), dtype=torch.int16, device="cuda"), np.zeros(1, dtype=np.int64), want_paths=True)
    assert w.numel() == 0 and p.shape == (2, 0)


def test_recognizer_isolated_words(tmp_path):
    """lnasr.recognizer.Recognizer with several acoustic models: the word is the model with the best Viterbi score."""
    from lnasr.bank import AcousticPipeline, ModelBank
    from lnasr.gmmhmm import GMMHMM
    from lnasr.mfcc import MFCC
    from lnasr.recognizer import Recognizer
    rng = np.random.default_rng(12)
    models = _models(rng, 4, 5, 4)
    files = {}
    for name, mdl in zip(["yes", "no", "stop", "go"], models):
        g = GMMHMM(5, 4, 39, A=mdl["A"], pi=mdl["pi"], w=mdl["w"], mu=mdl["mu"], si=mdl["var"][..., None] * np.eye(39))
        files[name] = str(tmp_path / f"{name}.npz")
        g.save(files[name])
    r = Recognizer({"am": files, "lm": None})
    _, _, sigs = _batch(rng, [9000, 12000, 7000, 15000, 8000])
    words = r.recognize_batch(sigs)
    ref = AcousticPipeline(MFCC(), ModelBank(models, emission="exact")).recognize(sigs)
    assert words == [list(files)[i] for i in ref]
    assert r.recognize(sigs[2]) == words[2]


def test_submit_host_keeps_batches_in_flight():
    """AcousticPipeline.submit_host: several batches queued before any result is collected; every handle returns what the
    synchronous call returns for its batch."""
    import torch
    from lnasr.bank import AcousticPipeline, ModelBank
    from lnasr.mfcc import MFCC
    rng = np.random.default_rng(33)
    pipe = AcousticPipeline(MFCC(), ModelBank(_models(rng, 5, 5, 4), emission="exact"))
    batches = []
    for _ in range(4):
        pcm, off, _ = _batch(rng, [int(x) for x in rng.integers(3000, 15000, size=29)])
        batches.append((torch.from_numpy(pcm).pin_memory(), off))
    ref = [pipe.recognize_batch_host(p, o, n_chunks=3) for p, o in batches]
    handles = [pipe.submit_host(p, o, n_chunks=5, want_paths=True) for p, o in batches]       # all four in flight
    for h, (w_ref, b_ref) in zip(reversed(handles), reversed(ref)):                         # collected out of order
        w, b = h.result()
        assert h.done() and np.array_equal(w.numpy(), w_ref.numpy()) and np.array_equal(b.numpy(), b_ref.numpy())


def test_first_use_on_side_streams():
    """Device-side state that is built lazily (MFCC tables and the mel plan, packed models, transition lists) is complete before
    another stream can use it: fresh objects whose FIRST call runs inside submit_host (copy + two compute streams) and inside a
    user's side stream give the results of the plain call."""
    import torch
    from lnasr.bank import AcousticPipeline, ModelBank
    from lnasr.mfcc import MFCC
    rng = np.random.default_rng(77)
    models = _models(rng, 3, 5, 4)
    pcm, off, _ = _batch(rng, [int(x) for x in rng.integers(3000, 12000, size=17)])
    ref_res, ref_mb = AcousticPipeline(MFCC(), ModelBank(models, emission="exact")).recognize_batch(pcm, off)
    ref_w = ref_res.words.cpu().numpy()
    ref_f = ref_mb.feats.cpu()
    for _ in range(3):
        pipe = AcousticPipeline(MFCC(), ModelBank(models, emission="exact"))      # nothing of it exists on the device yet
        h = pipe.submit_host(torch.from_numpy(pcm).pin_memory(), off)
        assert np.array_equal(h.result()[0].cpu().numpy(), ref_w)
        side = torch.cuda.Stream()
        mf = MFCC()
        with torch.cuda.stream(side):
            feats = mf.calc_mfcc_batch(torch.from_numpy(pcm).cuda(), sample_offsets=off).feats
        side.synchronize()
        assert torch.equal(feats.cpu().nan_to_num(), ref_f.nan_to_num())
