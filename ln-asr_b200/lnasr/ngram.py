"""N-gram language model -- import surface of the reference's lnasr/ngram.py (ngram.py:20-378).

OUT OF SCOPE of the acoustic hot path: dictionary-based text processing with no numeric hot loop.
It exists because the reference's recognizer.py, which must run unchanged against this package, does
`from lnasr.ngram import NGramCounter, NGramModel, NGramModelARPA` and builds `NGramModel(NGramModelARPA().load(file))`
(recognizer.py:14,31-32).  A compact host-side implementation with the reference's class names, constructor arguments and
semantics: fixed-discount (0.7) Katz back-off, log10 probabilities, ARPA text files.
"""
from __future__ import annotations

import math
import re
from collections import Counter, defaultdict
from typing import Dict, Iterable, Optional, Tuple

from .utils import Punctuation_Unicode_Set

__all__ = ["Tokenizer", "NGramCounter", "NGramModel", "NGramModelARPA"]


class Tokenizer:
    """Text -> tuple of tokens; punctuation acts as white space (ngram.py:20-44)."""
    Pnctt_Set = Punctuation_Unicode_Set

    @staticmethod
    def get_tokens(text: str, s: bool = True) -> Tuple[str, ...]:
        cleaned = "".join(" " if ch in Tokenizer.Pnctt_Set else ch for ch in text)
        words = cleaned.split()
        return tuple(["<s>"] + words + ["</s>"]) if s else tuple(words)


class NGramCounter(defaultdict):
    """counts[context (n-1 tokens)][word] for order n, with the lower orders in `.backoff` (ngram.py:46-114)."""

    def __init__(self, order: int, tokens: Iterable[Tuple[str, ...]]):
        super().__init__(Counter)
        self.order = order
        self.ngrams = set()
        tokens = tuple(tokens)
        for sent in tokens:
            for k in range(order - 1, len(sent)):
                ctx = tuple(sent[k - order + 1:k])
                self[ctx][sent[k]] += 1
                self.ngrams.add(ctx + (sent[k],))
        self.backoff = NGramCounter(order - 1, tokens) if order > 1 else None

    def __str__(self):
        lines = [f"{self.order}-grams"] + [f"{k} : {v}" for k, v in self.items()]
        return "\n".join(lines + ([str(self.backoff)] if self.backoff is not None else []))


class NGramModel(dict):
    """Katz back-off model over log10 probabilities, from counts or from a loaded ARPA model (ngram.py:116-273).

    self[context][word] = discounted probability; `.prob[ngram]` = log10 p; `.prob_bo[context]` = log10 alpha(context)."""
    DisCount = 0.7
    K = 5
    NInf = -1e300

    def __init__(self, ng):
        super().__init__()
        self.order = ng.order
        self.backoff: Optional["NGramModel"] = None
        self.prob: Dict[Tuple[str, ...], float] = {}
        self.prob_bo: Dict[Tuple[str, ...], float] = {}
        if isinstance(ng, NGramCounter):
            for ctx, words in ng.items():
                total = float(sum(words.values()))
                self[ctx] = {w: self.DisCount * c / total for w, c in words.items()}
                for w, p in self[ctx].items():
                    self.prob[ctx + (w,)] = math.log10(p) if p > 0 else self.NInf
            if self.order > 1:
                self.backoff = NGramModel(ng.backoff)
                for ctx, words in self.items():
                    left = 1.0 - sum(words.values())                          # mass freed by the discount
                    lower = sum(self.backoff[ctx[1:]][w] for w in words)      # lower-order mass of the seen words
                    alpha = left / (1.0 - lower)
                    self.prob_bo[ctx] = math.log10(alpha) if alpha > 0 else self.NInf
        elif isinstance(ng, NGramModelARPA):
            self.prob = ng.prob
            if self.order > 1:
                self.backoff = NGramModel(ng.backoff)
                self.prob_bo = ng.backoff.prob_bo          # ARPA keeps alpha(context) on the (n-1)-gram line
        else:
            raise TypeError("NGramModel needs an NGramCounter or an NGramModelARPA")

    def _logalpha(self, context: Tuple[str, ...]) -> float:
        return self.prob_bo.get(tuple(context), 0.0)

    def _logprob(self, word: str, context: Tuple[str, ...]) -> float:
        """log10 P(word | context) with back-off to shorter contexts."""
        key = tuple(context) + (word,)
        if key in self.prob or self.order == 1:
            return self.prob[key]
        return self._logalpha(context) + self.backoff._logprob(word, tuple(context)[1:])

    def calc_prob(self, sentence: Tuple[str, ...]) -> float:
        return sum(self._logprob(sentence[k], tuple(sentence[k - self.order + 1:k]))
                   for k in range(self.order - 1, len(sentence)))

    def calc_ppl(self, sentence: Tuple[str, ...]) -> float:
        return math.pow(10.0, -self.calc_prob(sentence) / len(sentence))


class NGramModelARPA:
    """ARPA text format: \\data\\, 'ngram n=count', '\\n-grams:' sections of 'log10p<TAB>w1 .. wn[<TAB>log10 alpha]', \\end\\
    (ngram.py:275-378)."""

    def __init__(self):
        self.order = 0
        self.prob: Dict[Tuple[str, ...], float] = {}
        self.prob_bo: Dict[Tuple[str, ...], float] = {}
        self.backoff: Optional["NGramModelARPA"] = None

    # chain of orders, highest first
    def _levels(self):
        lv, out = self, []
        while lv is not None:
            out.append(lv)
            lv = lv.backoff
        return out

    def _init(self, model: NGramModel):
        self.order, self.prob, self.prob_bo = model.order, model.prob, {}
        if model.order > 1:
            self.backoff = NGramModelARPA()
            self.backoff._init(model.backoff)
            self.backoff.prob_bo = model.prob_bo
        else:
            self.backoff = None

    def save(self, model: NGramModel, filename):
        self._init(model)
        levels = self._levels()[::-1]                                          # 1-grams first
        out = ["", "\\data\\"] + [f"ngram {lv.order}={len(lv.prob)}" for lv in levels]
        for lv in levels:
            out += ["", f"\\{lv.order}-grams:"]
            for gram, p in lv.prob.items():
                line = f"{p}\t{' '.join(gram)}"
                if gram in lv.prob_bo:
                    line += f"\t{lv.prob_bo[gram]}"
                out.append(line)
        out += ["", "\\end\\", ""]
        with open(filename, "w", encoding="utf-8") as fp:
            fp.write("\n".join(out))

    def load(self, filename):
        sections: Dict[int, list] = {}
        cur = None
        with open(filename, "r", encoding="utf-8") as fp:
            for raw in fp:
                line = raw.strip()
                if not line or line == "\\data\\" or re.match(r"ngram \d+=\d+", line):
                    continue
                if line == "\\end\\":
                    break
                m = re.match(r"\\(\d+)-grams:", line)
                if m:
                    cur = int(m.group(1))
                    sections[cur] = []
                elif cur is not None:
                    sections[cur].append(line.split())
        self.order = max(sections) if sections else 0
        lv = self
        for n in range(self.order, 0, -1):
            lv.order, lv.prob, lv.prob_bo = n, {}, {}
            for f in sections.get(n, []):
                gram = tuple(f[1:n + 1])
                lv.prob[gram] = float(f[0])
                if len(f) >= n + 2:
                    lv.prob_bo[gram] = float(f[n + 1])
            lv.backoff = NGramModelARPA() if n > 1 else None
            lv = lv.backoff
        return self
