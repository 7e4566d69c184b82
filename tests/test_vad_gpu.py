"""lnasr.vad.VadLtsd (csrc/vad.cu) against the reference's VadLtsd.detect (third/vadlstd/VadLstd.py:44-74): golden vectors of
the unmodified reference and the oracle on further inputs (int16 / float32 / float64 PCM, non-default window, noise adaptation,
utterances shorter than the envelope).  The VAD shares the framing rule and window of the MFCC front end."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag,kw", [("a", {}), ("b", {"alpha": 0.9, "threshold": 30.0}),
                                    ("c", {"winsize": 512, "stepsize": 256, "order": 3}), ("d", {})])
def test_vad_matches_reference_golden(golden, tag, kw):
    from lnasr.vad import VadLtsd
    g = golden("vad")
    got = VadLtsd(**kw).detect(g["pcm_" + tag])
    ref = g["ltsd_" + tag]
    assert got.shape == ref.shape and got.dtype == np.float64
    np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-9)


def test_vad_constructor_and_dtypes():
    from lnasr.vad import VadLtsd
    from oracle import lnasr_oracle as O
    v = VadLtsd()
    assert (v.freq, v.winsize, v.stepsize, v.order, v.threshold, v.alpha, v.fft_size) == (16000, 2048, 1024, 6, -6, None, 1025)
    rng = np.random.default_rng(5)
    x = 1000.0 * rng.standard_normal(33333)
    x[12000:20000] *= 20.0
    ref = O.vad_ltsd(x)
    np.testing.assert_allclose(v.detect(x), ref, rtol=1e-9, atol=1e-9)                          # float64 PCM, as the reference takes it
    x32 = x.astype(np.float32)
    np.testing.assert_allclose(v.detect(x32), O.vad_ltsd(x32.astype(np.float64)), rtol=1e-9, atol=1e-9)
    xi = np.round(x).clip(-32767, 32767).astype(np.int16)
    np.testing.assert_allclose(v.detect(xi), O.vad_ltsd(xi), rtol=1e-9, atol=1e-9)
    # loud stretch -> larger divergence than the quiet lead-in
    d = v.detect(xi)
    assert d[14:18].mean() > d[6:9].mean() + 10.0
    a = VadLtsd(winsize=1024, stepsize=512, order=4, threshold=25.0, alpha=0.8)
    np.testing.assert_allclose(a.detect(xi), O.vad_ltsd(xi, 1024, 512, 4, 25.0, 0.8), rtol=1e-9, atol=1e-9)
    assert VadLtsd().detect(np.zeros(0, dtype=np.int16)).shape == (O.vad_ltsd(np.zeros(0)).shape)
