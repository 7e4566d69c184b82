"""BASELINE.json's FULL sizes, checked through size-independent properties (no oracle run: the numpy port needs minutes
at these sizes).  Config 2: 4096 utterances x 10 word models (5 states x 4 mixtures x 39 dims).  Config 4 shape: 2000 states
x 16 mixtures x 39 dims, left-to-right, tensor-core emissions.

Properties: a Viterbi path's score re-summed from the emissions equals the reported maximum (float64, exact to rounding);
max-product <= sum-product (Viterbi score <= forward log-probability); every step of a path is an allowed transition;
results do not depend on the batch an utterance is scored in (permutation of the batch permutes the results, bit for bit);
the three exact-emission kernels agree; emissions obey max_m(w+bm) <= b <= max_m(w+bm) + ln M."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _path_scores(torch, b, paths, f_off, logA, logpi, K, N):
    """Score of every (utterance, model) path re-summed on the device in float64: pi[s0] + b[0,s0] + sum A[s_{t-1},s_t] + b[t,s_t]."""
    T = b.shape[0]
    n_utt = len(f_off) - 1
    out = torch.empty((n_utt, K), dtype=torch.float64, device=b.device)
    off = torch.as_tensor(f_off, device=b.device)
    first = torch.zeros(T, dtype=torch.bool, device=b.device)
    first[off[:-1]] = True
    utt = torch.repeat_interleave(torch.arange(n_utt, device=b.device), off[1:] - off[:-1])
    for k in range(K):
        s = paths[k].long()
        e = b[torch.arange(T, device=b.device), k * N + s].double()
        prev = torch.roll(s, 1)
        tr = torch.where(first, logpi[k][s], logA[k][prev, s])
        out[:, k] = torch.zeros(n_utt, dtype=torch.float64, device=b.device).index_add_(0, utt, e + tr)
    return out


def test_config2_full_size_properties():
    import torch
    import bench
    from lnasr import _capi
    from lnasr.bank import AcousticPipeline, ModelBank
    from lnasr.mfcc import MFCC
    models = bench.make_models("c2")
    pipe = AcousticPipeline(MFCC(), ModelBank(models, emission="exact"))
    pcm, off = bench.synth_batch(77, 4096, 8000, 24000)
    res, mb = pipe.recognize_batch(pcm, off, want_paths=True, want_logprob=True)
    K, N = pipe.bank.K, pipe.bank.n
    T = int(mb.frame_offsets_host[-1])
    assert T > 390000 and res.paths.shape == (K, T)
    feats = mb.feats
    assert torch.isfinite(feats).all()
    b = pipe.bank.emissions(feats)
    # emissions: max_m(w + bm) <= b <= max_m(w + bm) + ln M
    b2, bm = pipe.bank.gmm.loglik(feats[:50000], want_bm=True)
    wbm = bm.view(50000, K * N, 4) + pipe.bank.gmm.logw_d.view(1, K * N, 4)
    mx = wbm.max(dim=2).values
    assert torch.all(b2 >= mx - 1e-9) and torch.all(b2 <= mx + np.log(4.0) + 1e-9)
    # the three exact kernels agree at full size
    lib = _capi.lib()
    try:
        lib.lnasr_debug_set_gmm_simt(1)
        b_simt = pipe.bank.emissions(feats)
        lib.lnasr_debug_set_gmm_simt(2)
        b_stream = pipe.bank.emissions(feats)
    finally:
        lib.lnasr_debug_set_gmm_simt(0)
    assert torch.allclose(b, b_simt, rtol=1e-11, atol=1e-9) and torch.allclose(b, b_stream, rtol=1e-13, atol=1e-11)
    # Viterbi: path score == reported score; score <= forward log-probability; word = argmax with the lowest index on ties
    ps = _path_scores(torch, b, res.paths, mb.frame_offsets_host, pipe.bank.trans.logA, pipe.bank.trans.logpi, K, N)
    assert torch.allclose(ps, res.scores, rtol=1e-12, atol=1e-8)
    assert torch.all(res.scores <= res.logprob + 1e-9)
    assert torch.equal(res.words.long(), res.scores.argmax(dim=1)) and torch.equal(res.best, res.scores.max(dim=1).values)
    # batch independence: score the utterances in reverse order
    n_utt = len(off) - 1
    order = np.arange(n_utt)[::-1]
    lens = np.diff(off)
    off_r = np.concatenate([[0], np.cumsum(lens[order])]).astype(np.int64)
    pcm_r = np.concatenate([pcm[off[u]:off[u + 1]] for u in order])
    res_r, _ = pipe.recognize_batch(pcm_r, off_r)
    assert torch.equal(res_r.scores.flip(0), res.scores) and torch.equal(res_r.words.flip(0), res.words)
    # and the host-audio entry point decodes the same words
    w_h, b_h = pipe.recognize_batch_host(torch.from_numpy(pcm).pin_memory(), off, want_paths=True)
    assert np.array_equal(w_h.numpy(), res.words.cpu().numpy()) and np.array_equal(b_h.numpy(), res.best.cpu().numpy())


@pytest.mark.parametrize("emission", ["tc", "tc16"])
def test_config4_shape_properties(emission):
    import torch
    import bench
    from lnasr.bank import AcousticPipeline, ModelBank
    from lnasr.mfcc import MFCC
    models = bench.make_models("c4")
    pipe = AcousticPipeline(MFCC(), ModelBank(models, emission=emission))
    pcm, off = bench.synth_batch(78, 48, 40000, 48000)
    res, mb = pipe.recognize_batch(pcm, off, want_paths=True, want_logprob=True)
    N = pipe.bank.n
    T = int(mb.frame_offsets_host[-1])
    b = pipe.bank.emissions(mb.feats)
    assert b.dtype == torch.float32 and b.shape == (T, N) and torch.isfinite(b).all()
    ps = _path_scores(torch, b, res.paths, mb.frame_offsets_host, pipe.bank.trans.logA, pipe.bank.trans.logpi, 1, N)
    assert torch.isfinite(ps).all()                                   # every step is an allowed (finite) transition
    assert torch.allclose(ps, res.scores, rtol=1e-12, atol=1e-7)
    assert torch.all(res.scores <= res.logprob + 1e-9)
    # left-to-right topology, start in state 0: paths are non-decreasing and start at 0
    p = res.paths[0].long()
    foff = torch.as_tensor(mb.frame_offsets_host, device=p.device)
    assert torch.all(p[foff[:-1]] == 0)
    d = p[1:] - p[:-1]
    inner = torch.ones(T - 1, dtype=torch.bool, device=p.device)
    inner[foff[1:-1] - 1] = False                                     # differences across utterance borders do not count
    assert torch.all(d[inner] >= 0) and torch.all(d[inner] <= 2)
    # tensor-core emissions against the float64 kernel on the same features: 1e-4 relative (BASELINE.json)
    from lnasr.gmm import PackedGmm
    g = PackedGmm(models[0]["w"], models[0]["mu"], models[0]["var"])
    b64, _ = g.loglik(mb.feats[:4096])
    rel = ((b[:4096].double() - b64).abs() / b64.abs()).max().item()
    assert rel < 1e-4, rel
