"""Text side of the evaluation loop: token ids -> string, and the word normalisation applied before the WER.

Reference behaviour restated here (checked against outputs of the unmodified reference, tests/golden/norm_text.json):
* ``ids_to_text``: ``TextTransform.post_process`` (/root/reference/src/tokenizer/spm_tokenizer.py:46-54): drop ids of -1,
  concatenate the token strings, ``<space>`` -> " ", the sentencepiece word marker U+2581 -> " ", strip.
* ``norm_string`` (/root/reference/src/tokenizer/norm_text.py:3-134): every whitespace-separated word that contains one of a
  fixed set of punctuation characters is classified (word + trailing punctuation, contraction, hyphenated compound,
  percentage, dollar / pound amount, whisper tag, decimal number, abbreviation, domain name) and rewritten by its class;
  everything else only has non-alphanumeric characters blanked.  The result is upper-cased with single spaces.

Written from the behaviour, as a classification table + one rewrite rule per class.
"""
from __future__ import annotations

import re
from typing import Callable, List, Optional, Sequence, Tuple

_EDGE = ".,!?;:'\"-][~+"                      # characters stripped from both ends of a word before it is classified
_TRIGGER = frozenset("%$!\"&*+:£|<>/])~[_(-.,';?=@#^\\`{}’")

_AMOUNT = r"\d{1,10}[\.,]*(?:,\d{3})*\d*"

# (class name, pattern, matched against the edge-stripped word?) in the reference's order of precedence
_CLASSES: List[Tuple[str, "re.Pattern[str]", bool]] = [
    ("trailing_punct", re.compile(r"^\w+[.,!?;:]+$"), False),
    ("contraction", re.compile(r"^[A-Za-z]?[a-z]+(?:['’](?:[a-z]{1,2}|m|re|ve|ll|s|t))?$"), True),
    ("hyphenated", re.compile(r"^[a-zA-Z]+(?:-[a-zA-Z]+)+$"), True),
    ("percent", re.compile(r"^[0-9]+(?:\.[0-9]+)?%$"), True),
    ("dollar", re.compile(rf"(?:{_AMOUNT}\$$)|(?:\${_AMOUNT}$)"), True),
    ("pound", re.compile(rf"(?:{_AMOUNT}£$)|(?:£{_AMOUNT}$)"), True),
    ("whisper_tag", re.compile(r"^[a-zA-Z]+[.,?!']*<\|\w+\|><\|(translate|transcribe)\|>$"), True),
    ("decimal", re.compile(r"^[0-9]+[\.,]+[0-9]+$"), True),
    ("abbreviation", re.compile(r"[a-z]{1}(\.[a-z]{1})+$"), True),
    ("domain", re.compile(r"^[a-zA-Z0-9]+(?:\.[a-zA-Z0-9]+)+$"), True),
]


def _classify(word: str) -> str:
    low = word.lower()
    core = low.strip(_EDGE)
    for name, pat, on_core in _CLASSES:
        if pat.match(core if on_core else low):
            return name
    return "other"


def _spoken_number(w: str) -> str:
    return w.replace(",", "").replace(".", " point ")


_REWRITE = {
    "trailing_punct": lambda w: w,
    "contraction": lambda w: w,
    "hyphenated": lambda w: w.replace("-", " "),
    "percent": lambda w: _spoken_number(w).replace("%", " percent"),
    "dollar": lambda w: _spoken_number(w.replace("$", "")) + " dollar",
    "pound": lambda w: _spoken_number(w.replace("£", "")) + " pound",
    "decimal": lambda w: w.replace(".", " point ").replace(",", ""),
    "domain": lambda w: w.replace(".", " dot "),
    "abbreviation": lambda w: w.replace(".", ""),
    "other": lambda w: re.sub(r"[^a-zA-Z0-9' ]", " ", w),
}


def _rewrite(word: str, cls: str) -> str:
    up = word.upper()
    if cls == "whisper_tag":
        out = up.split("<")[0].strip(_EDGE)
    else:
        out = _REWRITE[cls](up.strip(_EDGE))
    return re.sub(r"\s+", " ", out).upper()


def norm_string(text: str) -> str:
    words = []
    for w in text.strip().split():
        cls = _classify(w) if (_TRIGGER & set(w)) else "other"
        words.append(_rewrite(w, cls))
    return " ".join(words)


def ids_to_text(token_ids: Sequence[int], token_list: Sequence[str]) -> str:
    """``TextTransform.post_process``: ids (ignore_id -1 dropped) -> text."""
    pieces = [token_list[int(t)] for t in token_ids if int(t) != -1]
    return "".join(pieces).replace("<space>", " ").replace("▁", " ").strip()


def make_text_functions(token_list: Sequence[str]) -> Tuple[Callable[[Sequence[int]], str], Callable[[str], str]]:
    """(ids_to_text, normalize) for ``evaluation.evaluate_sharded``, i.e. what eval_lrs2 applies to hypotheses and labels
    (script/evaluation.py:105-107, 392-399): post_process, "<eos>" / "<unk>" removed by the driver, then norm_string."""
    return (lambda ids: ids_to_text(ids, token_list)), norm_string
