"""Sharded, batched evaluation driver for the B200 path.

What the reference does (/root/reference/script/evaluation.py:387-404, ``eval_lrs2``): loop over the samples, one
``AVSRCocktailModel.inference(videos, audios)`` each, normalise hypothesis and label, ``jiwer.wer`` over the two lists, print
``WER: ...`` (``:553``).  Here the same result comes from: shard the utterances over the ranks (``sharding.shard_utterances``),
decode each rank's share in length-bucketed batches through ``AVSRCocktailB200.infer_batch`` (every utterance evolves as its
own B=1 run), ``all_gather`` the 1-best token ids, ``all_reduce`` [edits, reference words].  No collective on the hot path.

The model only needs ``infer_batch(videos[B,1,T,88,88], audios[B,104,T], lengths) -> n-best per utterance`` and ``eos``;
the CPU tests drive this module with a stand-in model under ``gloo``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import sharding as S


@dataclass
class EvalResult:
    wer: float
    edits: int
    ref_words: int
    hyp_tokens: Dict[int, List[int]] = field(default_factory=dict)      # utterance id -> 1-best token ids (no sos / eos)
    hyp_text: Dict[int, str] = field(default_factory=dict)
    audio_seconds: float = 0.0
    n_batches: int = 0


def strip_sos_eos(yseq: Sequence[int], eos: int) -> List[int]:
    """``nbest[0]["yseq"][1:]`` with the trailing ``<eos>`` dropped: what script/evaluation.py:105-107 turns into text
    (it strips the literal "<eos>" from the string instead)."""
    toks = [int(t) for t in yseq[1:]]
    while toks and toks[-1] == eos:
        toks.pop()
    return toks


_pinned: dict = {}          # reused pinned staging buffers of pad_batch (host samples), grown on demand


def _staging(name: str, numel: int) -> torch.Tensor:
    buf = _pinned.get(name)
    if buf is None or buf.numel() < numel:
        buf = torch.empty(int(numel * 1.25) + 1, dtype=torch.float32, pin_memory=torch.cuda.is_available())
        _pinned[name] = buf
    return buf[:numel]


def pad_batch(samples: Sequence[Tuple[torch.Tensor, torch.Tensor]]) -> Tuple[torch.Tensor, torch.Tensor, List[int]]:
    """samples: (video [1,T,88,88] or [T,88,88], audio [104,T]) per utterance -> videos [B,1,Tmax,88,88], audios
    [B,104,Tmax] zero-padded, and the frame counts (collate_pad of avhubert_dataset.py:277-312 for ready features).
    Device samples are padded on their device; host samples go into a reused PINNED staging buffer, which the encoder
    uploads in chunks under the video frontend.  The staging buffer is overwritten by the next call: consume the batch
    (``infer_batch``) before padding the next one."""
    lengths = [int(a.shape[-1]) for _, a in samples]
    tmax = max(lengths)
    B = len(samples)
    dev = samples[0][0].device
    if dev.type == "cuda":
        videos = torch.empty(B, 1, tmax, 88, 88, dtype=torch.float32, device=dev)
        audios = torch.empty(B, 104, tmax, dtype=torch.float32, device=dev)
    else:
        videos = _staging("video", B * tmax * 88 * 88).view(B, 1, tmax, 88, 88)
        audios = _staging("audio", B * 104 * tmax).view(B, 104, tmax)
    for b, (v, a) in enumerate(samples):
        v = v.reshape(-1, 88, 88)
        if v.shape[0] != lengths[b] or a.shape[0] != 104:
            raise RuntimeError(f"utterance {b}: video {tuple(v.shape)} and audio {tuple(a.shape)} disagree")
        videos[b, 0, :lengths[b]] = v
        audios[b, :, :lengths[b]] = a
        if lengths[b] < tmax:                          # only the tails need zeroing
            videos[b, 0, lengths[b]:] = 0
            audios[b, :, lengths[b]:] = 0
    return videos, audios, lengths


def evaluate_sharded(model, lengths: Sequence[int], load_sample: Callable[[int], Tuple[torch.Tensor, torch.Tensor]],
                     references: Optional[Sequence[str]] = None, ids_to_text: Optional[Callable[[Sequence[int]], str]] = None,
                     normalize: Optional[Callable[[str], str]] = None, max_utts: Optional[int] = None, max_frames: int = 12288,
                     device="cpu", fps: float = 25.0, group=None, collate: Optional[Callable] = None) -> EvalResult:
    """Decode utterances 0..N-1 (``lengths[i]`` frames each, inputs from ``load_sample(i)``) on all ranks of the default
    process group and return the corpus result on every rank.

    references / ids_to_text / normalize: label strings, the tokenizer's ``post_process`` and ``norm_string``; when any is
    missing the WER is computed on token ids written as decimal words (references then are token-id strings too).
    ``device``: where the gather / reduce tensors live ("cuda" under nccl, "cpu" under gloo).
    ``collate``: optional replacement for ``pad_batch``: ``collate(samples) -> (videos [B,1,T,88,88], audios [B,104,T], lengths)``,
    e.g. the GPU input pipeline on raw (uint8 frames, waveform) samples (avsr_b200.input_pipeline.DataCollator)."""
    rank, world = S._world(group)
    mine = S.shard_utterances(lengths, world)[rank]
    # max_utts=None: cost-optimal cuts (few utterances per rank -> small batches, many -> 100+); an int: fixed-size buckets
    batches = (S.plan_batches(mine, lengths, max_frames=max_frames) if max_utts is None
               else S.bucket_batches(mine, lengths, max_utts=max_utts, max_frames=max_frames))
    eos = int(model.eos)
    ids, toks = [], []
    for batch in batches:
        samples = [load_sample(i) for i in batch]
        videos, audios, lens = pad_batch(samples) if collate is None else collate(samples)
        if lens != [int(lengths[i]) for i in batch]:
            raise RuntimeError("load_sample returned utterances whose lengths differ from `lengths`")
        nbest = model.infer_batch(videos, audios, lens)
        for i, hyps in zip(batch, nbest):
            ids.append(i)
            toks.append(strip_sos_eos(hyps[0].yseq.tolist() if hasattr(hyps[0].yseq, "tolist") else hyps[0].yseq, eos))
    all_toks = S.gather_hypotheses(ids, toks, device=device, group=group)
    if sorted(all_toks) != list(range(len(lengths))):
        raise RuntimeError("sharding lost or duplicated utterances")
    to_text = ids_to_text if ids_to_text is not None else (lambda t: " ".join(str(int(x)) for x in t))
    norm = normalize if normalize is not None else (lambda s: s)
    text = {i: norm(to_text(all_toks[i]).replace("<eos>", "").replace("<unk>", "")) for i in all_toks}
    edits = nref = 0
    if references is not None:
        # each rank scores its own utterances; the sums meet in one all_reduce
        for i in mine:
            e, n = S.corpus_wer([norm(references[i].replace("<unk>", ""))], [text[i]])
            edits += e
            nref += n
    wer, e, n = S.reduce_wer(edits, nref, device=device, group=group)
    return EvalResult(wer=wer, edits=e, ref_words=n, hyp_tokens=all_toks, hyp_text=text,
                      audio_seconds=sum(int(t) for t in lengths) / fps, n_batches=len(batches))


# ---------------------------------------------------------------------------------------------------- AVCocktail loop
def parse_vtt(vtt_text: str) -> List[Tuple[float, float, str]]:
    """Cues of a WebVTT document as (start seconds, end seconds, text): what ``webvtt.read`` yields to ``eval_avcocktail``
    (script/evaluation.py:411-433; the ``webvtt`` package is not a dependency here).  Cue text lines are joined with "\\n"."""
    import re
    ts = re.compile(r"(?:(\d+):)?(\d{2}):(\d{2})[.,](\d{3})\s*-->\s*(?:(\d+):)?(\d{2}):(\d{2})[.,](\d{3})")
    cues, cur = [], None
    for line in vtt_text.replace("\r\n", "\n").split("\n"):
        m = ts.search(line)
        if m:
            g = m.groups()
            start = int(g[0] or 0) * 3600 + int(g[1]) * 60 + int(g[2]) + int(g[3]) / 1000
            end = int(g[4] or 0) * 3600 + int(g[5]) * 60 + int(g[6]) + int(g[7]) / 1000
            cur = [start, end, []]
            cues.append(cur)
        elif line.strip() == "":
            cur = None
        elif cur is not None:
            cur[2].append(line.strip())
    return [(s, e, "\n".join(t)) for s, e, t in cues]


def avcocktail_label(vtt_text: str, normalize: Callable[[str], str]) -> Tuple[str, float, float]:
    """The label side of ``eval_avcocktail`` (script/evaluation.py:411-433): non-empty cues sorted by start time, joined with
    spaces and normalised; plus the earliest start / latest end of the cues (the window chunks are filtered against)."""
    cues = [c for c in parse_vtt(vtt_text) if c[2] != ""]
    if not cues:
        raise ValueError("label VTT holds no cue")
    start, end = min(c[0] for c in cues), max(c[1] for c in cues)
    ordered = [t for _, t in sorted((c[0], c[2]) for c in cues)]
    return normalize(" ".join(ordered)), start, end


CHUNK_TYPES = ("asd_chunk", "fixed_chunk", "gold_chunk")


def evaluate_avcocktail(model, videos: Dict[str, dict], ids_to_text: Callable[[Sequence[int]], str],
                        normalize: Optional[Callable[[str], str]] = None, max_utts: Optional[int] = 32, max_frames: int = 12288,
                        device="cpu", group=None, collate: Optional[Callable] = None):
    """``eval_avcocktail`` over a set of videos (script/evaluation.py:406-453 and the ``*`` branch of ``main``, :556-570), sharded.

    videos: ``{set_id: {"label": <VTT text>, "asd_chunk" | "fixed_chunk" | "gold_chunk": [{"start_time", "end_time",
    "frames": T, "load": callable -> sample}]}}``.  As in the reference, a chunk is skipped when it starts more than 1 s before
    the first label cue or ends more than 1 s after the last one; the outputs of a (video, chunk type) are concatenated in
    start-time order, normalised and scored against the video's label with a word error rate; per chunk type the scores are
    averaged with the labels' word counts as weights.  All kept chunks of all videos and chunk types form ONE utterance list
    that ``evaluate_sharded`` deals over the ranks (chunks are independent; no collective on the hot path), every rank then
    stitches and scores locally from the gathered token ids.

    Returns ``(per_video {set_id: {chunk_type: wer}}, num_words {set_id: n}, average {chunk_type: wer})``."""
    from .text import norm_string, stitch_outputs
    norm = normalize or norm_string
    labels, flat = {}, []
    for set_id in sorted(videos):
        v = videos[set_id]
        label_text, t0, t1 = avcocktail_label(v["label"], norm)
        labels[set_id] = label_text
        for ct in CHUNK_TYPES:
            for ch in v.get(ct, []):
                s, e = float(ch["start_time"]), float(ch["end_time"])
                if s + 1 < t0 or e - 1 > t1:                   # :441-442
                    continue
                flat.append((set_id, ct, s, ch))
    lengths = [int(ch["frames"]) for _, _, _, ch in flat]
    res = evaluate_sharded(model, lengths, lambda i: flat[i][3]["load"](), ids_to_text=ids_to_text, normalize=lambda s: s,
                           max_utts=max_utts, max_frames=max_frames, device=device, group=group, collate=collate) if flat else None
    per_video: Dict[str, Dict[str, float]] = {}
    num_words = {k: len(t.split()) for k, t in labels.items()}
    for set_id in labels:
        per_video[set_id] = {}
        for ct in CHUNK_TYPES:
            idx = [i for i, f in enumerate(flat) if f[0] == set_id and f[1] == ct]
            if ct not in videos[set_id]:
                continue
            out_text = stitch_outputs([flat[i][2] for i in idx], [res.hyp_text[i] for i in idx], normalize=norm) if idx else ""
            e, n = S.corpus_wer([labels[set_id]], [out_text])
            per_video[set_id][ct] = e / n if n else float("nan")
    average = {}
    for ct in CHUNK_TYPES:
        pairs = [(per_video[k][ct], num_words[k]) for k in per_video if ct in per_video[k]]
        if pairs:
            average[ct] = sum(w * n for w, n in pairs) / max(1, sum(n for _, n in pairs))       # [wer] * num_words, then the mean (:564-570)
    return per_video, num_words, average
