"""Recognizer shell -- the classes of the reference's lnasr/recognizer.py (recognizer.py:20-48) on top of the B200 path.

The reference file itself runs unchanged against this package (tests/test_recognizer_surface.py executes its text); this
module exists so that `from lnasr.recognizer import Recognizer` (test/recognizer-test.py:13) also resolves when the whole
package is swapped.  Constructor arguments and attributes are the reference's: `AcousticModel(modelfile)` builds `MFCC()`
and `GMMHMM(n=2, m=2, d=MFCC.D).load(modelfile)`; `LanguageModel(modelfile)` builds `NGramModel(NGramModelARPA().load(..))`;
`Recognizer(model)` takes `{'am': ..., 'lm': ...}`.

`recognize` is `pass` in the reference (recognizer.py:46-48).  Here it is the isolated-word decision of SURVEY.md 8(c) when
`model['am']` names SEVERAL acoustic models (a list of files, or a dict word -> file): PCM -> MFCC -> emissions of every
word model -> Viterbi -> the word whose model scores best, all in one call of the C library (lnasr_recognize_batch).  With a
single acoustic-model file there is nothing to choose between and `recognize` returns None like the reference.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Union

import numpy as np

from .gmmhmm import GMMHMM
from .lexicon import Lexicon
from .mfcc import MFCC
from .ngram import NGramModel, NGramModelARPA

__all__ = ["AcousticModel", "LanguageModel", "Recognizer"]


class AcousticModel:
    """MFCC front end + GMM-HMM loaded from `modelfile` (recognizer.py:20-26)."""

    def __init__(self, modelfile):
        self.mfcc = MFCC()
        self.gmmhmm = GMMHMM(n=2, m=2, d=self.mfcc.D)
        self.gmmhmm.load(modelfile)


class LanguageModel:
    """N-gram model loaded from an ARPA file (recognizer.py:28-32)."""

    def __init__(self, modelfile):
        self.ngram = NGramModel(NGramModelARPA().load(modelfile))


class Recognizer:
    def __init__(self, model: dict):
        """model: {'am': acoustic model file | list of files | {word: file}, 'lm': ARPA file (optional here)}."""
        am = model["am"]
        self.words: Optional[List[str]] = None
        if isinstance(am, dict):
            self.words = [str(w) for w in am]
            files = list(am.values())
        elif isinstance(am, (list, tuple)):
            files = list(am)
            self.words = [str(i) for i in range(len(files))]
        else:
            files = [am]
        self.ams = [AcousticModel(f) for f in files]
        self.am = self.ams[0]
        self.lm = LanguageModel(model["lm"]) if model.get("lm") is not None else None
        self.lexicon = Lexicon(self.words) if self.words is not None else Lexicon()
        self._pipe = None

    def _pipeline(self):
        if self._pipe is None:
            from .bank import AcousticPipeline, ModelBank
            self._pipe = AcousticPipeline(self.am.mfcc, ModelBank([a.gmmhmm for a in self.ams], emission="auto"))
        return self._pipe

    def recognize_batch(self, signals: Sequence[np.ndarray]) -> List[str]:
        """Isolated-word recognition of several utterances at once (additive API)."""
        if self.words is None:
            raise ValueError("recognize_batch needs several acoustic models: pass model['am'] as a list or a {word: file} dict")
        idx = self._pipeline().recognize(signals)
        return list(self.lexicon.map(idx))

    def recognize(self, data: np.ndarray) -> Optional[str]:
        """One utterance of PCM -> word (see the module docstring); None with a single acoustic model, as in the reference."""
        if self.words is None:
            return None
        return self.recognize_batch([np.asarray(data)])[0]
