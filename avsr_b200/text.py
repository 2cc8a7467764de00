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


# ------------------------------------------------------------------------------------------------- segments, VTT, stitching
def fixed_chunks(duration: float, max_length: float = 15) -> List[Tuple[float, float]]:
    """``InferenceEngine.chunk_video`` without an ASD file (/root/reference/script/evaluation.py:247-270): the clip is cut
    into ceil(duration / max_length) chunks of a whole number of seconds, on a centisecond grid; returns (start, end) in
    seconds.  The reference divides by zero for an empty clip; here that is a ValueError."""
    import math
    if not duration > 0:
        raise ValueError("fixed_chunks needs a positive duration")
    num_chunks = math.ceil(duration / max_length)
    chunk_size = math.ceil(duration / num_chunks)
    steps, step_size = int(duration * 100), int(chunk_size * 100)
    return [(i / 100, min((i + step_size) / 100, duration)) for i in range(0, steps, step_size)]


def format_vtt_timestamp(timestamp: float) -> str:
    """``InferenceEngine.format_vtt_timestamp`` (script/evaluation.py:272-278): HH:MM:SS.mmm, everything truncated."""
    hours = int(timestamp // 3600)
    minutes = int((timestamp % 3600) // 60)
    seconds = int(timestamp % 60)
    milliseconds = int((timestamp - int(timestamp)) * 1000)
    return f"{hours:02d}:{minutes:02d}:{seconds:02d}.{milliseconds:03d}"


def segment_hypotheses(segments: Sequence[Tuple[float, float]], outputs: Sequence[str], offset: float = 0.0) -> List[dict]:
    """The list ``infer_video`` returns (script/evaluation.py:327-333): one dict per segment, times shifted by the track's
    start offset."""
    if len(segments) != len(outputs):
        raise ValueError("segments and outputs differ in length")
    return [{"start_time": s[0] + offset, "end_time": s[1] + offset, "text": o} for s, o in zip(segments, outputs)]


def write_vtt(hypotheses: Sequence[dict]) -> str:
    """The per-speaker .vtt body of ``mcorec_session_infer`` (script/evaluation.py:376-385): ``<unk>`` removed, empty cues
    dropped, cues in the order given."""
    parts = ["WEBVTT\n\n"]
    for hyp in hypotheses:
        text = hyp["text"].strip().replace("<unk>", "").strip()
        if len(text) == 0:
            continue
        parts.append(f"{format_vtt_timestamp(hyp['start_time'])} --> {format_vtt_timestamp(hyp['end_time'])}\n{text}\n\n")
    return "".join(parts)


def stitch_outputs(start_times: Sequence[float], outputs: Sequence[str], normalize: Optional[Callable[[str], str]] = None) -> str:
    """Per-video concatenation of the AVCocktail evaluation (script/evaluation.py:449-451): chunk outputs sorted by
    (start time, text), joined with spaces, ``<unk>`` removed, normalised."""
    if len(start_times) != len(outputs):
        raise ValueError("start_times and outputs differ in length")
    ordered = [o for _, o in sorted(zip(start_times, outputs))]
    return (normalize or norm_string)(" ".join(ordered).replace("<unk>", ""))
