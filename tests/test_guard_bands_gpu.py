"""Out-of-bounds WRITES, found without compute-sanitizer (the tool is closed on this GPU pool): every output and workspace
buffer of the main C-ABI entry points is carved out of a larger allocation whose margins hold a sentinel pattern; the call
gets exactly the number of bytes its *_workspace_bytes query asks for, and the margins must come back untouched.  The
results are compared with the ordinary (wrapper-allocated) call, so a kernel that stayed inside its buffer but wrote the
wrong place is caught as well."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BAND = 1 << 16            # 64 KB of sentinel on either side
PATTERN = 0xA5


class Guarded:
    """nbytes of device memory between two sentinel bands; .ptr is 256-byte aligned."""

    def __init__(self, nbytes):
        import torch
        self.nbytes = int(nbytes)
        self.raw = torch.full((2 * BAND + self.nbytes + 512,), PATTERN, dtype=torch.uint8, device="cuda")
        base = self.raw.data_ptr() + BAND
        self.off = BAND + ((-base) % 256)
        self.ptr = C.c_void_p(self.raw.data_ptr() + self.off)

    def view(self, dtype, shape):
        import torch
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        assert n <= self.nbytes
        return self.raw[self.off:self.off + n].view(dtype).view(shape)

    def check(self, what):
        lo, hi = self.raw[:self.off], self.raw[self.off + self.nbytes:]
        assert bool((lo == PATTERN).all()), f"{what}: bytes BEFORE the buffer were overwritten"
        assert bool((hi == PATTERN).all()), f"{what}: bytes AFTER the buffer were overwritten"


def _batch(rng, lens):
    sigs = [np.clip(np.round(3000.0 * rng.standard_normal(n)), -32767, 32767).astype(np.int16) for n in lens]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    return np.concatenate(sigs), off


@pytest.mark.parametrize("emission,n,m", [("exact", 5, 4), ("tc16", 2000, 16), ("tc", 88, 16)])
def test_decode_chain_stays_inside_its_buffers(emission, n, m):
    import torch
    from lnasr import _capi, _dev
    from lnasr.bank import AcousticPipeline, ModelBank
    from lnasr.mfcc import MFCC
    from oracle import lnasr_oracle as O
    rng = np.random.default_rng(n)
    K = 3 if n == 5 else 1
    models = [O.random_gmmhmm(n, m, 39, rng, left_to_right=(n > 32)) for _ in range(K)]
    for mdl in models:
        mdl["mu"] = mdl["mu"] * 100.0
        mdl["var"] = mdl["var"] * 400.0
    pcm, off = _batch(rng, [16000, 240 + 160 * 37, 9999, 4000, 31111])
    pipe = AcousticPipeline(MFCC(), ModelBank(models, emission=emission))
    ref, mb = pipe.recognize_batch(pcm, off, want_paths=True)
    lib = _capi.lib()
    t = torch
    n_utt, total = len(off) - 1, int(mb.frame_offsets_host[-1])
    pcm_d = torch.from_numpy(pcm).cuda()
    meta = _dev.to_dev(np.concatenate([off, mb.frame_offsets_host]), t.int64)
    s_off, f_off = C.c_void_p(meta.data_ptr()), C.c_void_p(meta.data_ptr() + 8 * (n_utt + 1))
    st = _capi.stream_ptr()

    # ---- MFCC alone
    need = C.c_size_t(0)
    _capi.check(lib.lnasr_mfcc_workspace_bytes(total, n_utt, C.byref(need)), "mfcc_workspace_bytes")
    g_feats, g_ws = Guarded(total * 39 * 4), Guarded(need.value)
    tb = pipe.mfcc._device_tables()
    _capi.check(lib.lnasr_mfcc_batch(_capi.ptr(pcm_d), _capi.DT_I16, s_off, f_off, n_utt, total, C.byref(tb), g_feats.ptr, 39, None, None,
                                     g_ws.ptr, need.value, st), "mfcc_batch")
    torch.cuda.synchronize()
    g_feats.check("mfcc features")
    g_ws.check("mfcc workspace")
    feats = g_feats.view(t.float32, (total, 39))
    assert torch.equal(feats.nan_to_num(), mb.feats.nan_to_num())
    # the same call without a prepared mel plan (lnasr_mfcc_tables.mel_plan = NULL): planned inside the workspace, same bits
    tb2 = _capi.MfccTables.from_buffer_copy(tb)
    tb2.mel_plan = None
    g_feats2, g_ws2 = Guarded(total * 39 * 4), Guarded(need.value)
    _capi.check(lib.lnasr_mfcc_batch(_capi.ptr(pcm_d), _capi.DT_I16, s_off, f_off, n_utt, total, C.byref(tb2), g_feats2.ptr, 39, None, None,
                                     g_ws2.ptr, need.value, st), "mfcc_batch")
    torch.cuda.synchronize()
    g_feats2.check("mfcc features (own plan)")
    g_ws2.check("mfcc workspace (own plan)")
    assert torch.equal(g_feats2.view(t.float32, (total, 39)).nan_to_num(), feats.nan_to_num())

    # ---- emissions alone (tensor-core kernels: workspace = operand images + flags)
    ncol = K * n
    if emission != "exact":
        _capi.check(lib.lnasr_gmm_tc_workspace_bytes(total, C.byref(need)), "gmm_tc_workspace_bytes")
        g_b, g_gws = Guarded(total * ncol * 4), Guarded(need.value)
        fn = lib.lnasr_gmm_loglik_tc16 if emission == "tc16" else lib.lnasr_gmm_loglik_tc
        _capi.check(fn(_capi.ptr(mb.feats), 39, total, _capi.ptr(pipe.bank._tc.packed), ncol, m, 39, g_b.ptr, ncol, g_gws.ptr, need.value, st),
                    "gmm_loglik_tc")
        torch.cuda.synchronize()
        g_b.check("tensor-core emissions")
        g_gws.check("tensor-core workspace")
        b = g_b.view(t.float32, (total, ncol))
        code = _capi.DT_F32
    else:
        g_b = Guarded(total * ncol * 8)
        _capi.check(lib.lnasr_gmm_loglik_exact(_capi.ptr(mb.feats), _capi.DT_F32, 39, total, _capi.ptr(pipe.bank.gmm.packed), ncol, m, 39,
                                               g_b.ptr, _capi.DT_F64, ncol, None, st), "gmm_loglik_exact")
        torch.cuda.synchronize()
        g_b.check("float64 emissions")
        b = g_b.view(t.float64, (total, ncol))
        code = _capi.DT_F64
    assert torch.equal(b, pipe.bank.emissions(mb.feats))

    # ---- Viterbi + backtrace alone
    vbytes = int(lib.lnasr_viterbi_workspace_bytes(total, n_utt, K, n))
    g_path, g_score, g_vws = Guarded(K * total * 4), Guarded(n_utt * K * 8), Guarded(vbytes)
    _capi.check(lib.lnasr_viterbi_batch(C.c_void_p(b.data_ptr()), code, ncol, f_off, n_utt, total, C.byref(pipe.bank.trans.struct),
                                        g_path.ptr, g_score.ptr, None, None, g_vws.ptr, vbytes, st), "viterbi_batch")
    torch.cuda.synchronize()
    for g, what in ((g_path, "paths"), (g_score, "scores"), (g_vws, "viterbi workspace")):
        g.check(what)
    assert torch.equal(g_path.view(t.int32, (K, total)), ref.paths) and torch.equal(g_score.view(t.float64, (n_utt, K)), ref.scores)

    # ---- the whole chain as one call on one workspace
    rec = pipe._recognizer_struct()
    _capi.check(lib.lnasr_recognize_workspace_bytes(total, n_utt, C.byref(rec), 1, C.byref(need)), "recognize_workspace_bytes")
    g_words, g_best, g_paths, g_rws = Guarded(n_utt * 4), Guarded(n_utt * 8), Guarded(K * total * 4), Guarded(need.value)
    _capi.check(lib.lnasr_recognize_batch(_capi.ptr(pcm_d), _capi.DT_I16, s_off, f_off, n_utt, total, C.byref(rec), g_words.ptr, g_best.ptr,
                                          g_paths.ptr, g_rws.ptr, need.value, st), "recognize_batch")
    torch.cuda.synchronize()
    for g, what in ((g_words, "words"), (g_best, "best"), (g_paths, "paths"), (g_rws, "recognize workspace")):
        g.check(what)
    assert torch.equal(g_words.view(t.int32, (n_utt,)), ref.words) and torch.equal(g_paths.view(t.int32, (K, total)), ref.paths)


@pytest.mark.parametrize("n,mix", [(3, 8), (70, 2)])
def test_baum_welch_statistics_stay_inside_their_buffers(n, mix):
    import torch
    from lnasr import _capi, _dev
    from lnasr._trellis import TransitionBank
    from lnasr.gmm import PackedGmm
    from oracle import lnasr_oracle as O
    rng = np.random.default_rng(n)
    d = 39
    mdl = O.random_gmmhmm(n, mix, d, rng, left_to_right=(n > 32))
    lens = [33, 1, 57, 2]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    total, n_utt = int(off[-1]), len(lens)
    x = torch.from_numpy(0.6 * rng.standard_normal((total, d))).cuda()
    g = PackedGmm(mdl["w"], mdl["mu"], mdl["var"])
    b, bm = g.loglik(x, want_bm=True)
    bank = TransitionBank(mdl["A"], mdl["pi"])
    off_d = _dev.to_dev(off, torch.int64)
    alpha, _ = bank.forward(b, off_d, n_utt, total, want_alpha=True)
    beta = bank.backward(b, off_d, n_utt, total)
    lib = _capi.lib()
    n_acc = int(lib.lnasr_em_acc_doubles(n, mix, d))
    wbytes = int(lib.lnasr_em_statistics_workspace_bytes(total, n))
    g_acc, g_ws = Guarded(n_acc * 8), Guarded(wbytes)
    acc = g_acc.view(torch.float64, (n_acc,))
    acc.zero_()
    _capi.check(lib.lnasr_em_statistics_ws(_capi.ptr(x), _capi.DT_F64, d, _capi.ptr(b), _capi.ptr(bm), _capi.ptr(alpha), _capi.ptr(beta),
                                           _capi.ptr(off_d), n_utt, total, _capi.ptr(bank.logA), _capi.ptr(g.logw_d), _capi.ptr(g.mu_d),
                                           n, mix, d, g_acc.ptr, None, g_ws.ptr, wbytes, _capi.stream_ptr()), "em_statistics_ws")
    torch.cuda.synchronize()
    g_acc.check("EM accumulator")
    g_ws.check("EM workspace")
    with np.errstate(all="ignore"):
        _, tot = O.em_step_multi_diag([x[off[u]:off[u + 1]].cpu().numpy() for u in range(n_utt)], mdl)
    NM = n * mix
    np.testing.assert_allclose(acc[:NM].cpu().numpy().reshape(n, mix), tot["S0"], rtol=1e-9, atol=1e-12)
    assert abs(float(acc[-2]) - tot["loglik"]) < 1e-8 * abs(tot["loglik"]) and float(acc[-1]) == n_utt
