"""Pronunciation lexicon -- mirror of the reference's lnasr/lexicon.py (lexicon.py:11-18), which is an empty
stub there (`map` is `pass`).  Here `map` is usable for isolated-word models: a dict {word index -> word}."""
from __future__ import annotations

from typing import Tuple

import numpy as np


class Lexicon(dict):
    """Connects acoustic-model outputs to words."""

    def __init__(self, words=None):
        super().__init__()
        if words is not None:
            for k, w in (words.items() if isinstance(words, dict) else enumerate(words)):
                self[int(k)] = w

    def map(self, state: np.ndarray) -> Tuple[str]:
        """Word indices (e.g. the argmax over a ModelBank) -> words; unknown indices map to '<unk>'."""
        return tuple(self.get(int(s), "<unk>") for s in np.atleast_1d(np.asarray(state)))
