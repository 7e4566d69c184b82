"""Tensor-core GMM path (csrc/gmm_tc.cu): descriptor bring-up probe and parity of the 3xTF32 contraction + fused
log-sum-exp against the float64 oracle.  Bar: emission log-likelihoods within 1e-4 relative (BASELINE.json)."""
import ctypes as C

import numpy as np
import pytest

from helpers import assert_close

pytestmark = pytest.mark.gpu


def _tf32(x):
    """Round float32 to TF32 precision (10-bit mantissa) so that products are exact in fp32."""
    u = np.asarray(x, dtype=np.float32).view(np.uint32)
    u = (u + 0x1000) & 0xFFFFE000
    return u.view(np.float32)


def _kmajor_image(mat):
    """[rows, K] float32 -> UMMA canonical K-major no-swizzle image [rows/8][K/4][8][4] (csrc/gmm_tc.cu kmajor_off)."""
    rows, K = mat.shape
    return np.ascontiguousarray(mat.reshape(rows // 8, 8, K // 4, 4).transpose(0, 2, 1, 3)).ravel()


def _probe(a, b, a_lbo, a_sbo, b_lbo, b_sbo):
    import torch
    from lnasr import _capi
    lib = _capi.lib()
    a_img = torch.from_numpy(_kmajor_image(a)).cuda()
    b_img = torch.from_numpy(_kmajor_image(b)).cuda()
    d = torch.zeros((128, 128), dtype=torch.float32, device="cuda")
    idesc = (1 << 4) | (2 << 7) | (2 << 10) | ((128 >> 3) << 17) | ((128 >> 4) << 24)
    _capi.check(lib.lnasr_debug_umma_probe(_capi.ptr(a_img), a_img.numel() * 4, _capi.ptr(b_img), b_img.numel() * 4,
                                           a_lbo, a_sbo, b_lbo, b_sbo, 0, idesc, _capi.ptr(d), _capi.stream_ptr()), "probe")
    torch.cuda.synchronize()
    return d.cpu().numpy()


def test_umma_descriptor_convention():
    """One tcgen05.mma M=128 N=128 K=8 kind::tf32: D = A . B^T with LBO = K-chunk stride, SBO = 8-row-group stride."""
    rng = np.random.default_rng(0)
    a = _tf32(rng.standard_normal((128, 8)))
    b = _tf32(rng.standard_normal((128, 8)))
    ref = a.astype(np.float64) @ b.astype(np.float64).T
    d = _probe(a, b, 128, 256, 128, 256)
    err_doc = np.abs(d - ref).max()
    d_sw = _probe(a, b, 256, 128, 256, 128)
    err_sw = np.abs(d_sw - ref).max()
    print(f"umma probe: max err documented convention {err_doc:.3e}, swapped {err_sw:.3e}")
    assert err_doc < 1e-4, f"documented LBO/SBO convention is wrong on this hardware (swapped err {err_sw:.3e})"


@pytest.mark.parametrize("fmt", ["tf32", "f16"])
@pytest.mark.parametrize("n,mix,d,T", [(14, 16, 39, 128), (30, 16, 39, 1000), (200, 16, 39, 777), (64, 4, 13, 300),
                                       (100, 8, 39, 129), (33, 3, 20, 260)])
def test_gmm_tc_matches_oracle(n, mix, d, T, fmt):
    import torch
    from lnasr.gmm_tc import PackedGmmTC
    from oracle import lnasr_oracle as O
    rng = np.random.default_rng(n + 7 * mix)
    mdl = O.random_gmmhmm(n, mix, d, rng)
    x = (0.5 * rng.standard_normal((T, d))).astype(np.float32)
    b_ref, _ = O.map_b(x.astype(np.float64), mdl["w"], mdl["mu"], mdl["var"])
    g = PackedGmmTC(mdl["w"], mdl["mu"], mdl["var"], fmt=fmt)
    b = g.loglik(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    got = b.cpu().numpy().T
    err = np.abs(got - b_ref)
    print(f"tc {fmt} n={n} m={mix} d={d} T={T}: max abs err {err.max():.3e}, max rel {np.max(err / np.abs(b_ref)):.3e}, "
          f"mean signed err {np.mean(got - b_ref):+.3e}")
    assert np.max(err / np.abs(b_ref)) < 1e-4
    assert err.max() < 2e-4         # fp32-level: what the per-utterance 1e-3 log-probability budget needs
    # padded feature rows (ld = 40) take the 16-byte load path and must give identical numbers
    xp = torch.zeros((T, 40), dtype=torch.float32, device="cuda")
    xp[:, :d] = torch.from_numpy(x).cuda()
    b2 = g.loglik(xp)
    assert torch.equal(b, b2)


@pytest.mark.parametrize("n,mix,d,T", [(200, 16, 39, 7777), (2000, 16, 39, 2100), (9, 1, 5, 1000), (70, 2, 39, 4000)])
def test_gmm_tc16_many_tiles(n, mix, d, T):
    """Several W tiles x several frame ranges per CTA (ring slots, TMEM stages and output staging wrap many times), at the
    config-4 model size among others; also: two runs give identical bits."""
    import torch
    from lnasr.gmm_tc import PackedGmmTC
    from oracle import lnasr_oracle as O
    rng = np.random.default_rng(n + mix + T)
    mdl = O.random_gmmhmm(n, mix, d, rng)
    x = (0.7 * rng.standard_normal((T, d))).astype(np.float32)
    g = PackedGmmTC(mdl["w"], mdl["mu"], mdl["var"], fmt="f16")
    xd = torch.from_numpy(x).cuda()
    b1 = g.loglik(xd).clone()
    b2 = g.loglik(xd).clone()
    torch.cuda.synchronize()
    assert torch.equal(b1, b2)
    rows = rng.choice(T, size=min(T, 300), replace=False)          # the oracle on a sample of frames (it is slow at n = 2000)
    b_ref, _ = O.map_b(x[rows].astype(np.float64), mdl["w"], mdl["mu"], mdl["var"])
    got = b1[torch.from_numpy(rows).cuda()].cpu().numpy().T
    err = np.abs(got - b_ref)
    assert np.max(err / np.abs(b_ref)) < 1e-4 and err.max() < 2e-4


@pytest.mark.parametrize("fmt", ["tf32", "f16"])
def test_gmm_tc_large_feature_magnitudes_with_shift(fmt):
    """Real MFCC magnitudes (|x| ~ 100) need the centring shift to stay inside 1e-4 relative comfortably."""
    import torch
    from lnasr.gmm_tc import PackedGmmTC
    from oracle import lnasr_oracle as O
    rng = np.random.default_rng(2)
    n, mix, d, T = 28, 16, 39, 256
    mdl = O.random_gmmhmm(n, mix, d, rng)
    centre = rng.uniform(-50, 50, size=d)
    mdl["mu"] = mdl["mu"] * 10 + centre
    x = (centre + 3.0 * rng.standard_normal((T, d))).astype(np.float32)
    b_ref, _ = O.map_b(x.astype(np.float64), mdl["w"], mdl["mu"], mdl["var"])
    g = PackedGmmTC(mdl["w"], mdl["mu"], mdl["var"], shift=centre, fmt=fmt)
    b = g.loglik(torch.from_numpy(x).cuda()).cpu().numpy().T
    print(f"tc {fmt} large magnitudes: max abs err {np.abs(b - b_ref).max():.3e}, max rel {np.max(np.abs(b - b_ref) / np.abs(b_ref)):.3e}")
    assert np.max(np.abs(b - b_ref) / np.abs(b_ref)) < 1e-4


@pytest.mark.parametrize("emission", ["tc", "tc16"])
def test_bank_with_tc_emissions_decodes_like_oracle(emission):
    """Viterbi on tensor-core emissions: paths bit-exact, ties / fp32 near-ties aside."""
    import torch
    from lnasr.bank import ModelBank
    from oracle import lnasr_oracle as O
    from helpers import path_score
    rng = np.random.default_rng(21)
    mdl = O.random_gmmhmm(140, 16, 39, rng, left_to_right=True)
    mdl["pi"] = np.full(140, -np.inf)
    mdl["pi"][0] = 0.0
    lens = [300, 150, 299]
    utts = [(0.5 * rng.standard_normal((T, 39))).astype(np.float32) for T in lens]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    bank = ModelBank([mdl], emission=emission)
    res = bank.score(torch.from_numpy(np.concatenate(utts)).cuda(), torch.from_numpy(off).cuda(), 3, int(off[-1]),
                     want_paths=True, want_logprob=True)
    paths = res.paths.cpu().numpy()[0]
    with np.errstate(all="ignore"):
        for u, x in enumerate(utts):
            b, _ = O.map_b(x.astype(np.float64), mdl["w"], mdl["mu"], mdl["var"])
            v, _, p_ref = O.viterbi(mdl["pi"], mdl["A"], b)
            got = paths[off[u]:off[u + 1]]
            s_got, s_ref = path_score(got, mdl["pi"], mdl["A"], b), path_score(p_ref, mdl["pi"], mdl["A"], b)
            assert abs(s_got - s_ref) < 1e-3, (s_got, s_ref)               # same path up to fp32 near-ties
            lp_ref = O.lse(O.forward(mdl["pi"], mdl["A"], b)[-1])
            d_v = res.scores.cpu().numpy()[u, 0] - np.max(v[-1])
            d_lp = res.logprob.cpu().numpy()[u, 0] - lp_ref
            print(f"{emission} utt T={lens[u]}: viterbi score err {d_v:+.3e}, log-prob err {d_lp:+.3e}")
            assert abs(d_v) < 1e-3 and abs(d_lp) < 1e-3              # BASELINE.json: 1e-3 absolute per utterance


@pytest.mark.parametrize("fmt", ["tf32", "f16"])
def test_out_of_range_and_non_finite_frames_take_the_exact_path(fmt):
    """No silent clamp: a frame whose features leave the operand format's range, or hold inf / NaN (the log energy of an
    all-zero frame is -inf, mfcc.py:158-160, and the deltas next to it are NaN), is flagged by the staging kernel and re-scored
    by the float64 fix-up kernel: finite far-away features exactly; inf / NaN features as the reference does -- NaN for every
    state (its full-matrix quadratic form, gmm.py:59-68, multiplies the zero off-diagonals of Sigma^-1 by inf)."""
    import torch
    from lnasr.gmm_tc import PackedGmmTC
    from oracle import lnasr_oracle as O
    rng = np.random.default_rng(17)
    n, mix, d, T = 40, 16, 39, 300
    mdl = O.random_gmmhmm(n, mix, d, rng)
    mdl["mu"] = mdl["mu"] * 100.0
    mdl["var"] = mdl["var"] * 400.0
    x = (20.0 * rng.standard_normal((T, d))).astype(np.float32)
    far, minf, nan, both = 5, 130, 131, 299
    x[far, 3] = 2.0e4                      # 1000 sigma away: |x'| > 240 (f16) and |x - shift| > 4096 (tf32)
    x[17, 0] = 9000.0
    x[minf, 12] = -np.inf
    x[nan, 25] = np.nan
    x[both, 12], x[both, 38] = -np.inf, np.nan
    si = mdl["var"][..., None] * np.eye(d)                       # the reference's [n,m,d,d] layout -> its 0 * inf = NaN
    with np.errstate(all="ignore"):
        b_ref, _ = O.map_b(x.astype(np.float64), mdl["w"], mdl["mu"], si)
    assert np.isnan(b_ref[:, [minf, nan, both]]).all() and np.isfinite(np.delete(b_ref, [minf, nan, both], axis=1)).all()
    b = PackedGmmTC(mdl["w"], mdl["mu"], mdl["var"], fmt=fmt).loglik(torch.from_numpy(x).cuda()).cpu().numpy().T.astype(np.float64)
    assert np.array_equal(np.isnan(b), np.isnan(b_ref)) and not np.isinf(b).any()
    fin = np.isfinite(b_ref)
    rel = np.abs(b[fin] - b_ref[fin]) / np.abs(b_ref[fin])
    assert rel.max() < 1e-4
    for t in (far, 17):                    # the exact path: only the float32 rounding of the result is left
        assert np.max(np.abs(b[:, t] - b_ref[:, t]) / np.abs(b_ref[:, t])) < 2e-7
    assert np.all(b[:, far] < -1e5)        # and nothing like the finite value a clamp to +-240 would have produced


@pytest.mark.parametrize("emission", ["exact", "tc", "tc16"])
def test_all_zero_frames_through_the_pipeline(emission):
    """PCM with a silent (all-zero) stretch: energy -inf, NaN deltas (the `_zero` case of the MFCC goldens).  Every emission
    kernel gives what the reference gives from the same features: NaN for every state of a frame with a non-finite feature,
    the usual values elsewhere."""
    import torch
    from lnasr.bank import ModelBank
    from lnasr.mfcc import MFCC
    from oracle import lnasr_oracle as O
    rng = np.random.default_rng(3)
    pcm = np.clip(np.round(3000.0 * rng.standard_normal(16000)), -32767, 32767).astype(np.int16)
    pcm[6000:8000] = 0
    mb = MFCC().calc_mfcc_batch([pcm])
    feats = mb.feats
    assert torch.isinf(feats[:, 12]).any() and torch.isnan(feats).any()
    mdl = O.random_gmmhmm(48, 8, 39, rng)
    mdl["mu"] = mdl["mu"] * 100.0
    mdl["var"] = mdl["var"] * 400.0
    with np.errstate(all="ignore"):
        b_ref, _ = O.map_b(feats.double().cpu().numpy(), mdl["w"], mdl["mu"], mdl["var"][..., None] * np.eye(39))
    b_ref = b_ref.T
    bad = ~np.isfinite(feats.cpu().numpy()).all(axis=1)
    assert bad.any() and not bad.all() and np.isnan(b_ref[bad]).all() and np.isfinite(b_ref[~bad]).all()
    b = ModelBank([mdl], emission=emission).emissions(feats).double().cpu().numpy()
    assert np.array_equal(np.isnan(b), np.isnan(b_ref)) and not np.isinf(b).any()
    assert np.max(np.abs(b[~bad] - b_ref[~bad]) / np.abs(b_ref[~bad])) < (1e-10 if emission == "exact" else 1e-4)


def test_tc16_epilogue_variants_agree():
    """The FP16-slice kernel's epilogue variants (lnasr_debug_set_tc16_groups / _poly): one or two groups of epilogue warps do
    the same arithmetic in the same order -> identical bits; the polynomial exp2 (max relative error 1.7e-7, inside the MUFU's
    2 ulp) moves an emission by less than 1e-6 and stays inside the parity bar against the oracle."""
    import torch
    from lnasr import _capi
    from lnasr.gmm_tc import PackedGmmTC
    from oracle import lnasr_oracle as O
    rng = np.random.default_rng(99)
    n, mix, d, T = 200, 16, 39, 3000
    mdl = O.random_gmmhmm(n, mix, d, rng)
    mdl["mu"] = mdl["mu"] * 100.0
    mdl["var"] = mdl["var"] * 400.0
    x = torch.from_numpy((25.0 * rng.standard_normal((T, d))).astype(np.float32)).cuda()
    g = PackedGmmTC(mdl["w"], mdl["mu"], mdl["var"], fmt="f16")
    lib = _capi.lib()
    out = {}
    try:
        for groups in (1, 2):
            for poly in (0, 1, 3):
                lib.lnasr_debug_set_tc16_groups(groups)
                lib.lnasr_debug_set_tc16_poly(poly)
                out[groups, poly] = g.loglik(x).clone()
    finally:
        lib.lnasr_debug_set_tc16_groups(2)
        lib.lnasr_debug_set_tc16_poly(0)
    for poly in (0, 1, 3):
        assert torch.equal(out[1, poly], out[2, poly]), poly
    rows = rng.choice(T, size=200, replace=False)
    b_ref, _ = O.map_b(x[torch.from_numpy(rows).cuda()].double().cpu().numpy(), mdl["w"], mdl["mu"], mdl["var"])
    for poly in (1, 3):
        assert (out[2, poly] - out[2, 0]).abs().max().item() < 5e-5            # float32 ulp at |b| in [128, 256) is 1.5e-5: at most 3 ulp
        got = out[2, poly][torch.from_numpy(rows).cuda()].cpu().numpy().T
        assert np.max(np.abs(got - b_ref) / np.abs(b_ref)) < 1e-6
