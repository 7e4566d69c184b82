"""CPU checks: the C-ABI library loads and exports every symbol include/lnasr_b200.h declares; the Python mirror has
the reference's classes / methods / signatures; the product path refuses to run without CUDA."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "lnasr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lnasr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from lnasr import _capi
    assert os.path.exists(_capi.LIB_PATH), "build liblnasr_b200.so first: python ln-asr_b200/build.py"
    lib = ctypes.CDLL(_capi.LIB_PATH)
    names = _declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/lnasr_b200.h but not exported"
    # and the binding table covers the header
    assert set(names) == set(_capi.SIGNATURES), set(names) ^ set(_capi.SIGNATURES)
    assert _capi.lib().lnasr_version() >= 100


def test_host_helpers_without_gpu():
    from lnasr import _capi
    lib = _capi.lib()
    # frame-count rule incl. the abs() (utils.py:118)
    for s, f in [(100, 1), (240, 0), (241, 1), (400, 1), (401, 2), (560, 2), (561, 3), (16000, 99), (21751, 135), (960000, 5999)]:
        assert lib.lnasr_mfcc_num_frames(s, 400, 160) == f
    assert lib.lnasr_em_acc_doubles(3, 8, 39) == 24 * 79 + 9 + 3 + 2            # 1 913 doubles
    n = ctypes.c_size_t(0)
    assert lib.lnasr_gmm_tc_packed_bytes(2000, 16, 39, ctypes.byref(n)) == 0 and n.value > 2000 * 16 * 80 * 8
    assert lib.lnasr_gmm_tc_packed_bytes(10, 17, 39, ctypes.byref(n)) != 0
    assert b"n_mix" in lib.lnasr_last_error()


def test_reference_signatures():
    from lnasr.gmmhmm import GMMHMM
    from lnasr.hmm import HMM
    from lnasr.mfcc import MFCC
    from lnasr import gmm, utils

    def params(f):
        return list(inspect.signature(f).parameters)

    assert params(MFCC.__init__) == ["self", "fs", "frame_T", "frame_stride", "hpf_alpha", "fft_N", "mfbank_num", "cepstrum_num"]
    assert MFCC.D == 39
    m = MFCC()
    assert (m.fs, m.frame_L, m.frame_step, m.fft_N, m.fft_size, m.mfbank_num, m.cepstrum_num, m.hpf_alpha) == \
        (16000, 400, 160, 512, 257, 40, 12, 0.97)
    for name in ["calc_high_pass_filter", "calc_mel", "calc_mel_inv", "create_filter_bank", "calc_mfcc"]:
        assert callable(getattr(MFCC, name))
    assert params(HMM.__init__) == ["self", "n", "m", "A", "B", "pi", "precision"]
    assert params(GMMHMM.__init__) == ["self", "n", "m", "d", "A", "pi", "w", "mu", "si", "precision"]
    expect = {"_map_b": ["self", "obs"], "_forward": ["self", "obs", "remap"], "_backward": ["self", "obs", "remap"],
              "_viterbi": ["self", "obs", "remap"], "_baumwelch": ["self", "obs", "alpha", "beta", "remap"],
              "_estimate": ["self", "obs", "alpha", "beta", "xi", "gamma"], "_updatemodel": ["self", "model"],
              "reset": ["self", "init_type"], "save": ["self", "filename"], "load": ["self", "filename"],
              "calc_prob": ["self", "obs"], "decode": ["self", "obs"], "train": ["self", "obs", "iters", "eps", "verbose"]}
    for cls in (HMM, GMMHMM):
        for name, p in expect.items():
            assert params(getattr(cls, name)) == p, (cls.__name__, name)
    assert inspect.signature(HMM.train).parameters["eps"].default == 0.0001
    assert GMMHMM(2, 2, 39).min_std == 0.01
    for name, p in {"gaussian_distribution": ["x", "mu", "sigma2"], "gaussian_distribution_log": ["x", "mu", "sigma2"],
                    "gaussian_multi_distribution": ["x", "mu", "Sigma"], "gaussian_multi_distribution_log": ["x", "mu", "Sigma"],
                    "gaussian_mixture_distribution": ["c", "x", "mu", "Sigma"],
                    "gaussian_mixture_distribution_log": ["c", "x", "mu", "Sigma"]}.items():
        assert params(getattr(gmm, name)) == p
    for name in ["recording", "read_wave", "read_pcm", "split_frames", "create_hamming", "lse", "lse2",
                 "Punctuation_Ascii_Set", "Punctuation_Unicode_Set"]:
        assert hasattr(utils, name)


def test_utils_match_reference_golden(golden):
    from lnasr import utils
    g = golden("utils")
    a3 = g["a3"]
    assert np.array_equal(utils.lse(a3), g["lse_all"]) and np.array_equal(utils.lse(a3, (0, 2)), g["lse_ax02"])
    assert np.array_equal(utils.lse(a3, 0), g["lse_ax0"]) and np.array_equal(utils.lse(a3, 2), g["lse_ax2"])
    assert np.array_equal(utils.create_hamming(400), g["hamming400"])
    assert np.array_equal(utils.split_frames(np.arange(1000.0), 400, 160), g["frames_1000"])
    assert np.array_equal(utils.split_frames(np.arange(100.0), 400, 160), g["frames_100"])


def test_reset_reproduces_reference_draw_order(golden):
    """Same np.random.seed -> same initial model as the reference's reset('random') (gmmhmm.py:179-188, hmm.py:243-250)."""
    from lnasr.gmmhmm import GMMHMM
    from lnasr.hmm import HMM
    g = golden("gmmhmm_5_4_39")
    np.random.seed(0)
    m = GMMHMM(5, 4, 39)
    m.reset("random")
    for k in ("A", "pi", "w", "mu", "si"):
        assert np.array_equal(getattr(m, k), g[k]), k
    gd = golden("hmm_discrete")
    np.random.seed(11)
    h = HMM(4, 5)
    h.reset("random")
    assert np.array_equal(h.A, gd["d_A0"]) and np.array_equal(h.B, gd["d_B0"]) and np.array_equal(h.pi, gd["d_pi0"])
    u = GMMHMM(3, 2, 4)
    u.reset("uniform")
    assert np.allclose(np.exp(u.A).sum(1), 1) and np.all(u.mu == 0) and np.array_equal(u.si[1, 1], np.eye(4))


def test_filter_bank_table_matches_reference(golden):
    from lnasr.mfcc import MFCC
    assert np.array_equal(MFCC().create_filter_bank(), golden("mfcc")["fbank"])


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from lnasr import _capi
    from lnasr.gmmhmm import GMMHMM
    from lnasr.mfcc import MFCC
    with pytest.raises(_capi.LnasrError):
        MFCC().calc_mfcc(np.zeros(16000, dtype=np.int16))
    m = GMMHMM(2, 2, 3)
    m.reset("uniform")
    with pytest.raises(_capi.LnasrError):
        m.decode(np.zeros((5, 3)))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ln-asr_b200", "lnasr")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("oracle definition", ""), f"{fn} mentions the oracle"


def test_shard_utterances_balances_frames():
    from lnasr.dist import shard_utterances
    rng = np.random.default_rng(0)
    lens = rng.integers(50, 400, size=1000)
    parts = [shard_utterances(lens, 8, r) for r in range(8)]
    assert sorted(i for p in parts for i in p) == list(range(1000))
    loads = [int(lens[p].sum()) for p in parts]
    assert max(loads) - min(loads) <= lens.max()
