"""MFCC kernel (csrc/mfcc.cu through lnasr.mfcc.MFCC) against the golden vectors of the unmodified reference
and against the numpy oracle.  Tolerance: 1e-4 relative (BASELINE.json), see helpers.assert_close."""
import numpy as np
import pytest

from helpers import assert_close

pytestmark = pytest.mark.gpu

CASES = ["", "_s401", "_s1000", "_s16000", "_s48017", "_float", "_zero"]


@pytest.fixture(scope="module")
def mfcc():
    from lnasr.mfcc import MFCC
    return MFCC()


@pytest.mark.parametrize("tag", CASES)
def test_calc_mfcc_matches_reference_golden(golden, mfcc, tag):
    g = golden("mfcc")
    pcm = g["pcm_data_raw"] if tag == "" else g["pcm" + tag]
    P, C, F = mfcc.calc_mfcc(pcm)
    assert P.dtype == np.float64 and F.shape == g["feats" + tag].shape
    assert_close(P, g["power" + tag], 1e-4, "power" + tag)
    assert_close(C, g["ceps" + tag], 1e-4, "cepstrum" + tag)
    assert_close(F, g["feats" + tag], 1e-4, "features" + tag)


def test_attributes_and_tables(golden, mfcc):
    g = golden("mfcc")
    assert (mfcc.frame_L, mfcc.frame_step, mfcc.fft_size, mfcc.D) == (400, 160, 257, 39)
    assert np.array_equal(mfcc.create_filter_bank(), g["fbank"])
    x = g["pcm_data_raw"]
    y = mfcc.calc_high_pass_filter(x)
    ref = np.append(x[0], x[1:] - 0.97 * x[:-1])
    np.testing.assert_allclose(y, ref, rtol=1e-14, atol=1e-9)


def test_ragged_batch_equals_per_utterance_oracle(mfcc):
    from oracle import lnasr_oracle as O
    rng = np.random.default_rng(5)
    lens = [401, 16000, 800, 5000, 48017, 720, 2400, 33333]
    sigs = []
    for n in lens:
        t = np.arange(n) / 16000.0
        x = np.round(3000.0 * rng.standard_normal(n)) + 2000.0 * np.sin(2 * np.pi * rng.uniform(80, 400) * t)
        sigs.append(np.clip(x, -32767, 32767).astype(np.int16))
    res = mfcc.calc_mfcc_batch(sigs)
    feats = res.feats.double().cpu().numpy()
    off = res.frame_offsets_host
    assert list(np.diff(off)) == [O.num_frames(n) for n in lens]
    for u, s in enumerate(sigs):
        _, _, F = O.mfcc(s)
        assert_close(feats[off[u]:off[u + 1]], F, 1e-4, f"utt{u}")


def test_short_utterance_raises_like_reference(mfcc):
    # S=300 -> 1 frame: the reference indexes features[1] and raises IndexError (mfcc.py:164)
    with pytest.raises(IndexError):
        mfcc.calc_mfcc(np.ones(300, dtype=np.int16))


def test_large_batch_properties(mfcc):
    """Full-size property checks (no oracle): shape rule, per-utterance zero mean of c1..c12, delta identities."""
    rng = np.random.default_rng(11)
    n_utt, n = 512, 16000
    pcm = np.clip(np.round(3000.0 * rng.standard_normal(n_utt * n)), -32767, 32767).astype(np.int16)
    off = np.arange(n_utt + 1, dtype=np.int64) * n
    res = mfcc.calc_mfcc_batch(pcm, sample_offsets=off)
    F = res.feats.double().cpu().numpy().reshape(n_utt, 99, 39)
    assert np.all(np.isfinite(F))
    assert np.abs(F[:, :, :12].mean(axis=1) + 1e-8).max() < 1e-3                      # CMN (mfcc.py:156)
    np.testing.assert_allclose(F[:, 1:, 13:26], F[:, 1:, :13] - F[:, :-1, :13], atol=2e-4)    # delta (mfcc.py:162-166)
    np.testing.assert_allclose(F[:, 0, 13:26], F[:, 1, :13], atol=1e-6)                       # delta row 0 = static row 1
    np.testing.assert_allclose(F[:, 1:, 26:], F[:, 1:, 13:26] - F[:, :-1, 13:26], atol=4e-4)  # delta-delta
    np.testing.assert_allclose(F[:, 0, 26:], F[:, 1, 13:26], atol=1e-6)


@pytest.mark.parametrize("kw", [
    dict(frame_T=20e-3, frame_stride=10e-3, mfbank_num=26),                  # 320-sample frames, 26 filters
    dict(fs=8000, frame_T=25e-3, frame_stride=10e-3, mfbank_num=20, hpf_alpha=0.95),
    dict(hpf_alpha=0.9375, mfbank_num=64, cepstrum_num=12),                  # alpha*100 not an integer -> float64 pre-emphasis; 64 filters
    dict(frame_T=32e-3, frame_stride=8e-3, mfbank_num=33, cepstrum_num=10),  # 512-sample frames (= fft_N), 11 static coefficients
])
def test_non_default_constructor_arguments_match_oracle(kw):
    """MFCC(fs, frame_T, frame_stride, hpf_alpha, fft_N, mfbank_num, cepstrum_num) (mfcc.py:23-27): the kernel is generic in
    frame length (<= fft_N), hop, filter count (<= 64), coefficient count and pre-emphasis; int16 and float input."""
    from lnasr.mfcc import MFCC
    from oracle import lnasr_oracle as O
    rng = np.random.default_rng(17)
    m = MFCC(**kw)
    p = O.MfccParams(**kw)
    for n, as_float in ((9000, False), (2 * p.frame_L + 3, False), (7001, True)):
        t = np.arange(n) / float(p.fs)
        x = np.round(3000.0 * rng.standard_normal(n)) + 2000.0 * np.sin(2 * np.pi * 210.0 * t)
        x = np.clip(x, -32767, 32767).astype(np.int16)
        sig = x.astype(np.float32) * 0.37 if as_float else x
        P, C, F = m.calc_mfcc(sig)
        Pr, Cr, Fr = O.mfcc(sig.astype(np.float64), p)
        assert F.shape == Fr.shape == (O.num_frames(n, p.frame_L, p.frame_step), 3 * (p.cepstrum_num + 1))
        assert_close(P, Pr, 1e-4, "power")
        assert_close(C, Cr, 1e-4, "cepstrum")
        assert_close(F, Fr, 1e-4, "features")


def test_fuzz_lengths_and_alignments(mfcc):
    """Many short utterances of arbitrary (odd and even) lengths packed back to back: every frame start alignment, the
    zero-padded last frames and the row-0 delta rule, int16 and float32, against the oracle."""
    from oracle import lnasr_oracle as O
    rng = np.random.default_rng(2024)
    lens = [401, 402, 559, 560, 561, 562, 720, 721] + [int(x) for x in rng.integers(401, 4000, size=56)]
    sigs = [np.clip(np.round(4000.0 * rng.standard_normal(n)), -32767, 32767).astype(np.int16) for n in lens]
    for as_float in (False, True):
        batch = [s.astype(np.float32) * 0.5 for s in sigs] if as_float else sigs
        res = mfcc.calc_mfcc_batch(batch)
        feats = res.feats.double().cpu().numpy()
        off = res.frame_offsets_host
        assert list(np.diff(off)) == [O.num_frames(n) for n in lens]
        for u, s in enumerate(batch):
            _, _, F = O.mfcc(s.astype(np.float64))
            assert_close(feats[off[u]:off[u + 1]], F, 1e-4, f"utt{u} len {lens[u]} float={as_float}")


def test_filter_bank_without_triangle_structure(monkeypatch):
    """The frames kernel sums mel energies per interval between filter peaks (a bin feeds at most two neighbouring triangles,
    mfcc.py:66-106).  A bank that is NOT of that shape -- every bin inside three wide filters -- must take the kernel's general
    plan (one interval per filter support) and still equal power . bank^T."""
    from lnasr.mfcc import MFCC
    from oracle import lnasr_oracle as O
    rng = np.random.default_rng(5)
    n_mel = 20
    bank = np.zeros((n_mel, 257))
    for m in range(n_mel):
        lo = 11 * m
        bank[m, lo:lo + 33] = rng.uniform(0.1, 1.0, size=33)
    bank[:, 257:] = 0.0

    class WideBank(MFCC):
        def create_filter_bank(self):
            return bank.copy()

    monkeypatch.setattr(O, "mel_filter_bank", lambda p=None: bank.copy())
    m = WideBank(mfbank_num=n_mel)
    p = O.MfccParams(mfbank_num=n_mel)
    x = np.clip(np.round(3000.0 * rng.standard_normal(12000)), -32767, 32767).astype(np.int16)
    P, C, F = m.calc_mfcc(x)
    Pr, Cr, Fr = O.mfcc(x.astype(np.float64), p)
    assert_close(P, Pr, 1e-4, "power")
    assert_close(C, Cr, 1e-4, "cepstrum")
    assert_close(F, Fr, 1e-4, "features")


def test_float64_samples_are_not_narrowed(mfcc):
    """The reference takes float64 audio as it is (mfcc.py:108, e.g. samples normalised to [-1, 1]).  Such samples go through the
    C ABI as float64 (LNASR_DT_F64): same results as the oracle, and not the results of the same audio rounded to float32."""
    from oracle import lnasr_oracle as O
    rng = np.random.default_rng(99)
    n = 16000
    t = np.arange(n) / 16000.0
    x = (0.3 * np.sin(2 * np.pi * 173.0 * t) + 0.05 * rng.standard_normal(n)) / 3.0          # float64, not float32-representable
    assert x.dtype == np.float64 and not np.array_equal(x, x.astype(np.float32).astype(np.float64))
    P, C, F = mfcc.calc_mfcc(x)
    Pr, Cr, Fr = O.mfcc(x)
    assert_close(P, Pr, 1e-4, "power")
    assert_close(C, Cr, 1e-4, "cepstrum")
    assert_close(F, Fr, 1e-4, "features")
    # power is returned as float32: compare where it is well above the float32 rounding of the narrowed input
    P32, _, _ = mfcc.calc_mfcc(x.astype(np.float32))
    rel64 = np.abs(P - Pr) / Pr
    rel32 = np.abs(P32 - Pr) / Pr
    assert np.median(rel64) < 2e-7                       # float32 storage of the result
    assert rel32.max() > 10 * rel64.max()                # the narrowed input is visibly different in the weakest bins
    # a ragged batch mixing float64 utterances of different lengths
    sigs = [x[:4000], x[100:9000], x[5:805]]
    res = mfcc.calc_mfcc_batch(sigs)
    off = res.frame_offsets_host
    feats = res.feats.double().cpu().numpy()
    for u, s in enumerate(sigs):
        assert_close(feats[off[u]:off[u + 1]], O.mfcc(s)[2], 1e-4, f"float64 utt {u}")


def test_every_base_alignment_of_the_sample_buffer(mfcc):
    """The frames kernel copies samples in 16-byte chunks from the boundary at or below x[-1]; the buffer a caller passes may start
    anywhere inside such a chunk (AcousticPipeline hands out sub-ranges of one device buffer).  Views starting 0..9 samples into
    an allocation, int16 and float32, two utterances each: same features as the oracle, and nothing outside the view is used."""
    import torch
    from oracle import lnasr_oracle as O
    rng = np.random.default_rng(31)
    lens = [1203, 2500]
    sigs = [np.clip(np.round(4000.0 * rng.standard_normal(n)), -32767, 32767).astype(np.int16) for n in lens]
    ref = [O.mfcc(s.astype(np.float64))[2] for s in sigs]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    for dt, scale in ((torch.int16, 1), (torch.float32, 1)):
        for o in range(10):
            # poison on both sides of the view: a read outside it that reached the arithmetic would show up in the features
            buf = torch.full((o + int(off[-1]) + 64,), 32000, dtype=dt, device="cuda")
            buf[o:o + int(off[-1])] = torch.from_numpy(np.concatenate(sigs)).to(dt).cuda()
            res = mfcc.calc_mfcc_batch(buf[o:o + int(off[-1])], sample_offsets=off)
            feats = res.feats.double().cpu().numpy()
            fo = res.frame_offsets_host
            for u in range(2):
                assert_close(feats[fo[u]:fo[u + 1]], ref[u], 1e-4, f"{dt} view at sample {o}, utterance {u}")
