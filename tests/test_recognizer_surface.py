"""The reference's recognizer.py must run UNCHANGED against this package (BASELINE north star): its imports
(recognizer.py:11-17) and its constructors (recognizer.py:23-32).  Also the n-gram import surface (lnasr/ngram.py), which
is out of scope of the hot path but needed by that file.  CPU only: constructing MFCC / GMMHMM and loading a model needs no GPU."""
import importlib
import os
import sys
import types

import numpy as np
import pytest

REF = "/root/reference/lnasr"

CORPUS = ["the cat sat on the mat", "the dog sat on the log", "a cat and a dog", "the cat saw the dog on the mat"]


def _tokens(mod):
    return tuple(mod.Tokenizer.get_tokens(t) for t in CORPUS)


def test_ngram_backoff_model_and_arpa_round_trip(tmp_path):
    from lnasr import ngram
    toks = _tokens(ngram)
    assert toks[0][0] == "<s>" and toks[0][-1] == "</s>"
    assert ngram.Tokenizer.get_tokens("你好，世界", s=False) == ("你好", "世界")
    cnt = ngram.NGramCounter(3, toks)
    assert cnt.order == 3 and cnt.backoff.order == 2 and cnt.backoff.backoff.backoff is None
    assert cnt[("the", "cat")]["sat"] == 1 and cnt.backoff[("the",)]["cat"] == 2
    lm = ngram.NGramModel(cnt)
    seen = lm._logprob("sat", ("the", "cat"))
    assert seen == pytest.approx(np.log10(0.7 * 1 / 2))
    # unseen trigram backs off: alpha(the,cat) + log P(on | cat) -> alpha(cat) + log P(on)
    assert lm._logprob("on", ("the", "cat")) < seen
    p = lm.calc_prob(toks[0])
    assert np.isfinite(p) and p < 0 and lm.calc_ppl(toks[0]) > 1.0
    path = tmp_path / "tri.lm"
    ngram.NGramModelARPA().save(lm, str(path))
    lm2 = ngram.NGramModel(ngram.NGramModelARPA().load(str(path)))
    assert lm2.order == 3
    for t in toks + (ngram.Tokenizer.get_tokens("the dog saw a cat on the log"),):
        assert lm2.calc_prob(t) == pytest.approx(lm.calc_prob(t), rel=1e-12)


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only present in the authoring container")
def test_ngram_matches_reference_implementation():
    """Same probabilities as the unmodified reference ngram.py on a small corpus (loaded by path, shims as in make_golden.py)."""
    from lnasr import ngram
    for name in ("pyaudio", "librosa", "h5py"):
        sys.modules.setdefault(name, types.ModuleType(name))
    spec_u = importlib.util.spec_from_file_location("ref_lnasr_utils", os.path.join(REF, "utils.py"))
    ref_utils = importlib.util.module_from_spec(spec_u)
    spec_u.loader.exec_module(ref_utils)
    src = open(os.path.join(REF, "ngram.py"), encoding="utf-8").read().replace("from lnasr.utils import", "from ref_lnasr_utils import")
    sys.modules["ref_lnasr_utils"] = ref_utils
    ref = types.ModuleType("ref_ngram")
    exec(compile(src, "ref_ngram.py", "exec"), ref.__dict__)
    toks = _tokens(ngram)
    assert toks == tuple(ref.Tokenizer.get_tokens(t) for t in CORPUS)
    ours, theirs = ngram.NGramModel(ngram.NGramCounter(3, toks)), ref.NGramModel(ref.NGramCounter(3, toks))
    for t in toks + (ngram.Tokenizer.get_tokens("the dog saw a cat on the log"),):
        assert ours.calc_prob(t) == pytest.approx(theirs.calc_prob(t), rel=1e-12)
        assert ours.calc_ppl(t) == pytest.approx(theirs.calc_ppl(t), rel=1e-12)


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only present in the authoring container")
def test_reference_recognizer_runs_unchanged(tmp_path):
    """exec the reference's recognizer.py text, byte for byte, with `lnasr` = this package (and a stub h5py: absent here)."""
    import lnasr
    from lnasr import ngram
    from lnasr.gmmhmm import GMMHMM
    from lnasr.mfcc import MFCC
    assert os.path.dirname(lnasr.__file__).endswith(os.path.join("ln-asr_b200", "lnasr"))
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    src = open(os.path.join(REF, "recognizer.py"), encoding="utf-8").read()
    rec = types.ModuleType("recognizer_unchanged")
    exec(compile(src, "recognizer.py", "exec"), rec.__dict__)
    # model files: an acoustic model written by this package, an ARPA LM written by this package
    np.random.seed(3)
    am = GMMHMM(n=4, m=3, d=39)
    am.reset("random")
    am_path = str(tmp_path / "am.npz")
    am.save(am_path)
    lm_path = str(tmp_path / "tri.lm")
    ngram.NGramModelARPA().save(ngram.NGramModel(ngram.NGramCounter(3, _tokens(ngram))), lm_path)
    r = rec.Recognizer({"am": am_path, "lm": lm_path})
    assert isinstance(r.am.mfcc, MFCC) and isinstance(r.am.gmmhmm, GMMHMM)
    assert (r.am.gmmhmm.n, r.am.gmmhmm.m, r.am.gmmhmm.d) == (4, 3, 39)      # shapes recovered from the file (gmmhmm.py:205-207)
    assert np.array_equal(r.am.gmmhmm.mu, am.mu)
    assert r.lm.ngram.order == 3 and np.isfinite(r.lm.ngram.calc_prob(ngram.Tokenizer.get_tokens(CORPUS[0])))
    assert r.recognize(np.zeros(16000, dtype=np.int16)) is None             # unimplemented in the reference (recognizer.py:46-48)


def test_package_recognizer_module(tmp_path):
    """`from lnasr.recognizer import Recognizer` (test/recognizer-test.py:13) resolves in this package too; with one acoustic
    model `recognize` returns None like the reference's stub (recognizer.py:46-48)."""
    from lnasr import ngram
    from lnasr.gmmhmm import GMMHMM
    from lnasr.recognizer import AcousticModel, LanguageModel, Recognizer
    np.random.seed(5)
    am = GMMHMM(n=3, m=2, d=39)
    am.reset("random")
    am_path, lm_path = str(tmp_path / "am.npz"), str(tmp_path / "tri.lm")
    am.save(am_path)
    ngram.NGramModelARPA().save(ngram.NGramModel(ngram.NGramCounter(2, _tokens(ngram))), lm_path)
    r = Recognizer({"am": am_path, "lm": lm_path})
    assert isinstance(r.am, AcousticModel) and isinstance(r.lm, LanguageModel)
    assert (r.am.gmmhmm.n, r.am.gmmhmm.m, r.am.gmmhmm.d) == (3, 2, 39) and r.lm.ngram.order == 2
    assert r.recognize(np.zeros(16000, dtype=np.int16)) is None
    with pytest.raises(ValueError):
        r.recognize_batch([np.zeros(16000, dtype=np.int16)])
