"""Utterance sharding across the GPUs of one box, and the end-of-run gather / WER reduce.

The reference decodes one utterance at a time in a Python loop (/root/reference/script/evaluation.py:387-404, ``eval_lrs2``)
and computes a corpus-level word error rate with jiwer (``:402``).  Utterances are independent, so the B200 path shards
them: one process per GPU, every rank holds the full weights, NO collective on the hot path.  This module holds the host
logic of that scheme, all of it backend-agnostic (``nccl`` on the GPUs, ``gloo`` in the CPU tests):

* ``utterance_cost`` / ``shard_utterances``: length-sorted greedy dealing to the least-loaded rank;
* ``bucket_batches`` / ``plan_batches``: length-bucketed batches inside a rank (bounded frames per batch); the second cuts
  the sorted list where a measured cost model of a decode batch is smallest;
* ``gather_hypotheses``: ``all_gather`` of a padded int32 token matrix + lengths + utterance ids (KBs);
* ``word_edit_distance`` / ``reduce_wer``: word-level Levenshtein and ``all_reduce(SUM)`` of [edits, reference words], i.e.
  corpus WER = sum(edits) / sum(ref words), which is what jiwer's ``wer(list, list)`` returns.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

# per-frame encoder FLOPs and the attention term: T * (1 268 871 168 + 98 304 T)
ENC_FLOPS_PER_FRAME = 1_268_871_168
ENC_ATTN_FLOPS_PER_FRAME2 = 98_304
# one decode position of one utterance costs about as much wall time as this many encoder FLOPs (measured on B200 at
# B=32, beam 3: ~0.79 ms per position for 32 utterances vs ~38 ms for 32 x 489.65 GFLOP of encoder)
DEC_FLOPS_EQUIV_PER_POSITION = 1.0e10


def utterance_cost(T: int, decode_positions: Optional[int] = None) -> float:
    """Relative cost of one utterance of T frames: encoder FLOPs + decode positions (random-init models decode T positions,
    trained ones ~ #tokens + 3; the caller may pass an estimate)."""
    T = int(T)
    pos = T if decode_positions is None else int(decode_positions)
    return T * (ENC_FLOPS_PER_FRAME + ENC_ATTN_FLOPS_PER_FRAME2 * T) + DEC_FLOPS_EQUIV_PER_POSITION * pos


def shard_utterances(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Longest-first greedy bin packing: returns, per rank, the indices of its utterances (each sorted by length, longest
    first).  Deterministic: ties go to the lowest rank / lowest index, so every rank computes the same plan locally."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    load = [0.0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        shards[r].append(i)
        load[r] += utterance_cost(lengths[i])
    return shards


def bucket_batches(indices: Sequence[int], lengths: Sequence[int], max_utts: int = 32, max_frames: int = 12288) -> List[List[int]]:
    """Length-bucketed batches of a rank's utterances: walk them longest first and cut a batch when it would exceed
    ``max_utts`` utterances or ``max_frames`` packed frames.  A decode batch runs max(T) positions, so similar lengths
    together waste the fewest no-op positions."""
    if max_utts < 1 or max_frames < 1:
        raise ValueError("max_utts and max_frames must be positive")
    order = sorted(indices, key=lambda i: (-int(lengths[i]), i))
    batches, cur, frames = [], [], 0
    for i in order:
        t = int(lengths[i])
        if cur and (len(cur) >= max_utts or frames + t > max_frames):
            batches.append(cur)
            cur, frames = [], 0
        cur.append(i)
        frames += t
    if cur:
        batches.append(cur)
    return batches


# Cost of one decode batch on a B200, fitted on the configs[2] sweep in profiles/cfg2_sharded_r01.jsonl (1243 utterances,
# 32 .. 384 utterances per batch: the model reproduces the six single-GPU wall times within 6 %):
#   t(batch) = BATCH_FIXED_MS + T_max * (POSITION_FIXED_MS + POSITION_PER_UTT_MS * n)        (+ the encoder, which is per frame)
# i.e. a fixed cost per batch (session switch, graph replays, result collection), a per-position floor set by the chain of
# dependent launches, and a per-utterance term (the decode kernels work on all n * beam rows at every position until the
# longest utterance of the batch has ended).
BATCH_FIXED_MS = 14.0
POSITION_FIXED_MS = 0.2
POSITION_PER_UTT_MS = 0.0125


def batch_cost_ms(t_max: int, n: int) -> float:
    return BATCH_FIXED_MS + t_max * (POSITION_FIXED_MS + POSITION_PER_UTT_MS * n)


def plan_batches(indices: Sequence[int], lengths: Sequence[int], max_utts: int = 256, max_frames: int = 12288) -> List[List[int]]:
    """Cost-optimal length-bucketed batches of a rank's utterances: sort longest first and cut the sorted list where the
    sum of ``batch_cost_ms`` is smallest (dynamic programme over the cut points, O(N * max_utts)).  Few long utterances per
    rank -> small batches (no positions wasted on the short ones); many utterances -> batches of 100+ (the per-position
    floor is shared).  Every batch respects ``max_utts`` and ``max_frames``; deterministic."""
    if max_utts < 1 or max_frames < 1:
        raise ValueError("max_utts and max_frames must be positive")
    order = sorted(indices, key=lambda i: (-int(lengths[i]), i))
    n = len(order)
    if n == 0:
        return []
    L = np.array([int(lengths[i]) for i in order], dtype=np.int64)
    if int(L.max()) > max_frames:
        raise ValueError(f"an utterance of {int(L.max())} frames exceeds max_frames={max_frames}")
    csum = np.concatenate(([0], np.cumsum(L)))
    best = np.full(n + 1, np.inf)
    best[0] = 0.0
    cut = np.zeros(n + 1, dtype=np.int64)
    for j in range(1, n + 1):
        lo = max(0, j - max_utts)
        i = np.arange(lo, j)
        ok = (csum[j] - csum[i]) <= max_frames
        cost = best[i] + BATCH_FIXED_MS + L[i] * (POSITION_FIXED_MS + POSITION_PER_UTT_MS * (j - i))
        cost = np.where(ok, cost, np.inf)
        k = int(np.argmin(cost))                      # first minimum: ties go to the larger batch
        best[j], cut[j] = cost[k], lo + k
    batches, j = [], n
    while j > 0:
        batches.append(order[int(cut[j]):j])
        j = int(cut[j])
    return batches[::-1]


# ------------------------------------------------------------------------------------------------- gather of hypotheses
def _world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def gather_hypotheses(utt_ids: Sequence[int], token_seqs: Sequence[Sequence[int]], device="cpu", group=None) -> Dict[int, List[int]]:
    """Every rank contributes its (utterance id, token ids) pairs; every rank gets the complete {id: tokens} map.
    Three fixed-shape ``all_gather`` calls (counts, then a padded int32 matrix with a length and an id column); a few KB."""
    if len(utt_ids) != len(token_seqs):
        raise ValueError("utt_ids and token_seqs differ in length")
    _, world = _world(group)
    if world == 1:
        return {int(u): [int(t) for t in s] for u, s in zip(utt_ids, token_seqs)}
    dev = torch.device(device)
    n_local = len(utt_ids)
    l_local = max([len(s) for s in token_seqs], default=0)
    meta = torch.tensor([n_local, l_local], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    n_max = max(int(m[0]) for m in metas)
    l_max = max(int(m[1]) for m in metas)
    mat = torch.full((max(n_max, 1), l_max + 2), -1, dtype=torch.int32)
    for r, (u, s) in enumerate(zip(utt_ids, token_seqs)):
        mat[r, 0] = int(u)
        mat[r, 1] = len(s)
        if len(s):
            mat[r, 2:2 + len(s)] = torch.as_tensor(list(s), dtype=torch.int32)
    mat = mat.to(dev)
    mats = [torch.empty_like(mat) for _ in range(world)]
    dist.all_gather(mats, mat, group=group)
    out: Dict[int, List[int]] = {}
    for m, meta_r in zip(mats, metas):
        m = m.cpu().numpy()
        for r in range(int(meta_r[0])):
            u, n = int(m[r, 0]), int(m[r, 1])
            if u in out:
                raise RuntimeError(f"utterance {u} was decoded by more than one rank")
            out[u] = m[r, 2:2 + n].tolist()
    return out


# ------------------------------------------------------------------------------------------------- word error rate
def word_edit_distance(ref: Sequence[str], hyp: Sequence[str]) -> int:
    """Levenshtein distance between two word sequences (substitution = insertion = deletion = 1): the numerator of jiwer's
    ``wer``.  Row-vectorised dynamic programme, O(len(ref) * len(hyp))."""
    n, m = len(ref), len(hyp)
    if n == 0:
        return m
    if m == 0:
        return n
    ids: Dict[str, int] = {}
    a = np.array([ids.setdefault(w, len(ids)) for w in ref], dtype=np.int64)
    b = np.array([ids.setdefault(w, len(ids)) for w in hyp], dtype=np.int64)
    prev = np.arange(m + 1, dtype=np.int64)
    offs = np.arange(m + 1, dtype=np.int64)
    for i in range(1, n + 1):
        sub = prev[:-1] + (b != a[i - 1])
        dele = prev[1:] + 1
        best = np.minimum(sub, dele)
        # insertions chain along the row: cur[j] = min_k<=j (cand[k] + (j - k)) -> running minimum of cand[k] - k
        cand = np.concatenate(([i], best))
        cur = np.minimum.accumulate(cand - offs) + offs
        prev = cur
    return int(prev[m])


def reduce_wer(edits: int, ref_words: int, device="cpu", group=None) -> Tuple[float, int, int]:
    """Corpus WER over all ranks: all_reduce(SUM) of [edit distance, reference words] (int64)."""
    _, world = _world(group)
    t = torch.tensor([int(edits), int(ref_words)], dtype=torch.int64, device=torch.device(device))
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    e, n = int(t[0]), int(t[1])
    return (e / n if n > 0 else float("nan")), e, n


def corpus_wer(refs: Sequence[str], hyps: Sequence[str]) -> Tuple[int, int]:
    """(edits, reference words) of paired sentences, whitespace-tokenised like jiwer's default transform."""
    if len(refs) != len(hyps):
        raise ValueError("refs and hyps differ in length")
    e = n = 0
    for r, h in zip(refs, hyps):
        rw, hw = r.split(), h.split()
        e += word_edit_distance(rw, hw)
        n += len(rw)
    return e, n
