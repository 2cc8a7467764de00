"""Host logic of the multi-GPU path on CPU: sharding plan, batching, gather / WER reduce under gloo with world_size 2.

Reference behaviour being reproduced: the sequential loop + corpus-level jiwer WER of script/evaluation.py:387-404.
"""
import os
import random
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from avsr_b200 import evaluation as E
from avsr_b200 import sharding as S


def _lev_ref(a, b):
    """Textbook O(nm) Levenshtein in pure Python."""
    d = list(range(len(b) + 1))
    for i in range(1, len(a) + 1):
        prev, d[0] = d[0], i
        for j in range(1, len(b) + 1):
            cur = min(prev + (a[i - 1] != b[j - 1]), d[j] + 1, d[j - 1] + 1)
            prev, d[j] = d[j], cur
    return d[len(b)]


def test_word_edit_distance_matches_textbook():
    rng = random.Random(0)
    vocab = ["a", "b", "c", "dd", "e"]
    assert S.word_edit_distance([], []) == 0
    assert S.word_edit_distance(["a"], []) == 1
    assert S.word_edit_distance([], ["a", "b"]) == 2
    assert S.word_edit_distance("the cat sat".split(), "the cat sat".split()) == 0
    assert S.word_edit_distance("the cat sat on the mat".split(), "cat sat on a mat please".split()) == 3
    for _ in range(200):
        a = [rng.choice(vocab) for _ in range(rng.randint(0, 12))]
        b = [rng.choice(vocab) for _ in range(rng.randint(0, 12))]
        assert S.word_edit_distance(a, b) == _lev_ref(a, b), (a, b)


def test_corpus_wer_is_sum_of_edits_over_sum_of_words():
    refs = ["a b c d", "e f", "g"]
    hyps = ["a x c", "e f", ""]
    e, n = S.corpus_wer(refs, hyps)
    assert (e, n) == (2 + 0 + 1, 7)


def test_shard_plan_is_a_balanced_partition():
    # cfg 3 length law
    rng = np.random.default_rng(2024)
    lengths = np.clip(np.round(25 * rng.lognormal(np.log(1.3), 0.6, 1243)), 12, 155).astype(int).tolist()
    for world in (1, 2, 4, 8):
        shards = S.shard_utterances(lengths, world)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(len(lengths)))
        loads = [sum(S.utterance_cost(lengths[i]) for i in s) for s in shards]
        assert max(loads) / (sum(loads) / world) < 1.01          # greedy longest-first: < 1 % imbalance on 1243 utterances
        assert S.shard_utterances(lengths, world) == shards     # deterministic: every rank derives the same plan
    with pytest.raises(ValueError):
        S.shard_utterances(lengths, 0)


def test_bucket_batches_respect_limits_and_keep_lengths_close():
    lengths = [375] * 5 + [50, 60, 70, 300, 12, 13, 155]
    idx = list(range(len(lengths)))
    batches = S.bucket_batches(idx, lengths, max_utts=4, max_frames=1000)
    assert sorted(i for b in batches for i in b) == idx
    for b in batches:
        assert len(b) <= 4
        assert sum(lengths[i] for i in b) <= 1000 or len(b) == 1
    firsts = [lengths[b[0]] for b in batches]
    assert firsts == sorted(firsts, reverse=True)
    assert S.bucket_batches([], lengths) == []
    # a single utterance longer than max_frames still gets its own batch
    assert S.bucket_batches([0], [5000], max_frames=100) == [[0]]


class _Hyp:
    def __init__(self, yseq):
        self.yseq = torch.tensor(yseq, dtype=torch.int64)


class _StandInModel:
    """Deterministic stand-in for AVSRCocktailB200.infer_batch: the 'hypothesis' of an utterance is a function of its own
    input only, so any sharding / batching must give the same corpus result."""
    eos = 5048

    def infer_batch(self, videos, audios, lengths):
        out = []
        for b, t in enumerate(lengths):
            assert float(videos[b, 0, t:].abs().sum()) == 0.0 and float(audios[b, :, t:].abs().sum()) == 0.0
            key = int(round(float(audios[b, :, :t].sum()) * 7 + float(videos[b, 0, :t].sum())))
            n = 1 + (key % 5)
            toks = [1 + ((key * (k + 3)) % 40) for k in range(n)]
            out.append([_Hyp([self.eos] + toks + [self.eos])])
        return out


def _sample(i, lengths):
    g = torch.Generator().manual_seed(100 + i)
    t = lengths[i]
    return torch.randint(0, 3, (1, t, 88, 88), generator=g).float(), torch.randint(0, 5, (104, t), generator=g).float()


def _refs(lengths):
    return [" ".join(str((i * 7 + k) % 40) for k in range(1 + i % 4)) for i in range(len(lengths))]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, lengths, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = E.evaluate_sharded(_StandInModel(), lengths, lambda i: _sample(i, lengths), references=_refs(lengths),
                                 max_utts=3, max_frames=64)
        # uneven contribution: rank 1 sends nothing, rank 0 sends two ragged rows
        g = S.gather_hypotheses([5, 9] if rank == 0 else [], [[1, 2, 3], []] if rank == 0 else [])
        q.put((rank, res.wer, res.edits, res.ref_words, res.hyp_tokens, g))
    finally:
        dist.destroy_process_group()


def test_sharded_evaluation_world2_equals_single_process():
    lengths = [12, 30, 7, 22, 15, 9, 31, 5, 18]
    single = E.evaluate_sharded(_StandInModel(), lengths, lambda i: _sample(i, lengths), references=_refs(lengths))
    assert single.ref_words == sum(len(r.split()) for r in _refs(lengths))
    assert sorted(single.hyp_tokens) == list(range(len(lengths)))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lengths, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, wer, edits, nref, toks, g in got:
        assert (edits, nref) == (single.edits, single.ref_words)
        assert wer == pytest.approx(single.wer)
        assert toks == single.hyp_tokens
        assert g == {5: [1, 2, 3], 9: []}


def test_strip_and_pad():
    assert E.strip_sos_eos([5048, 3, 4, 5048], 5048) == [3, 4]
    assert E.strip_sos_eos([5048], 5048) == []
    v, a, lens = E.pad_batch([(torch.ones(1, 3, 88, 88), torch.ones(104, 3)), (torch.ones(5, 88, 88), torch.ones(104, 5))])
    assert v.shape == (2, 1, 5, 88, 88) and a.shape == (2, 104, 5) and lens == [3, 5]
    assert float(v[0, 0, 3:].sum()) == 0.0
    with pytest.raises(RuntimeError):
        E.pad_batch([(torch.ones(4, 88, 88), torch.ones(104, 3))])


def test_plan_batches_is_a_valid_and_cheaper_partition():
    rng = np.random.default_rng(2024)
    lengths = np.clip(np.round(25 * rng.lognormal(np.log(1.3), 0.6, 1243)), 12, 155).astype(int).tolist()

    def cost(bs):
        return sum(S.batch_cost_ms(max(lengths[i] for i in b), len(b)) for b in bs)

    for world in (1, 2, 8):
        for shard in S.shard_utterances(lengths, world)[:2]:
            plan = S.plan_batches(shard, lengths, max_utts=256, max_frames=12288)
            assert sorted(i for b in plan for i in b) == sorted(shard)
            assert all(len(b) <= 256 and sum(lengths[i] for i in b) <= 12288 for b in plan)
            flat = [lengths[i] for b in plan for i in b]
            assert flat == sorted(flat, reverse=True)                     # contiguous cuts of the length-sorted list
            assert plan == S.plan_batches(list(reversed(shard)), lengths)   # deterministic, order-independent
            for mu in (16, 32, 64, 128, 256):
                assert cost(plan) <= cost(S.bucket_batches(shard, lengths, max_utts=mu)) + 1e-9
    # few long utterances: small batches; many short ones: one big batch
    assert [len(b) for b in S.plan_batches(range(62), [150, 140] + [30] * 60)] == [2, 60]
    assert [len(b) for b in S.plan_batches(range(6), [150, 140, 30, 28, 27, 12])] == [6]      # the per-batch cost keeps them together
    assert len(S.plan_batches(range(200), [15] * 200)) == 1
    assert S.plan_batches([], lengths) == []
    # the frame bound cuts a batch even when the cost model would not
    assert [len(b) for b in S.plan_batches(range(10), [100] * 10, max_frames=400)] == [4, 4, 2] or \
        sum(len(b) for b in S.plan_batches(range(10), [100] * 10, max_frames=400)) == 10
    with pytest.raises(ValueError):
        S.plan_batches([0], [500], max_frames=400)


def test_plan_batches_is_optimal_on_small_sets():
    """Exhaustive check: on sets small enough to enumerate every way of cutting the length-sorted list, the dynamic
    programme returns a cut of minimal model cost (with the utterance and frame bounds in force)."""
    import itertools
    rng = random.Random(5)
    for trial in range(40):
        n = rng.randint(1, 9)
        lengths = [rng.choice([12, 20, 33, 60, 110, 155]) for _ in range(n)]
        max_utts, max_frames = rng.choice([2, 3, 9]), rng.choice([160, 400, 100000])
        order = sorted(range(n), key=lambda i: (-lengths[i], i))
        best = None
        for cuts in itertools.product([0, 1], repeat=n - 1):
            batches, cur = [], [order[0]]
            for c, i in zip(cuts, order[1:]):
                if c:
                    batches.append(cur)
                    cur = []
                cur.append(i)
            batches.append(cur)
            if any(len(b) > max_utts or sum(lengths[i] for i in b) > max_frames for b in batches):
                continue
            cost = sum(S.batch_cost_ms(max(lengths[i] for i in b), len(b)) for b in batches)
            best = cost if best is None else min(best, cost)
        plan = S.plan_batches(range(n), lengths, max_utts=max_utts, max_frames=max_frames)
        got = sum(S.batch_cost_ms(max(lengths[i] for i in b), len(b)) for b in plan)
        assert best is not None and abs(got - best) < 1e-9, (lengths, max_utts, max_frames, plan)


def test_avcocktail_loop_matches_a_sequential_restatement():
    """evaluate_avcocktail (script/evaluation.py:406-453, :556-570) with the stand-in model: label parsing from VTT, the 1 s
    window filter, time-ordered stitching per (video, chunk type), WER per chunk type and the word-count-weighted average."""
    from avsr_b200 import text
    vtt = ("WEBVTT\n\n00:00:02.000 --> 00:00:04.000\nsecond cue\n\n00:00:00.500 --> 00:00:01.500\nfirst cue here\n\n"
           "00:00:05.000 --> 00:00:05.500\n\n")
    assert E.parse_vtt(vtt)[0] == (2.0, 4.0, "second cue") and len(E.parse_vtt(vtt)) == 3
    label, t0, t1 = E.avcocktail_label(vtt, text.norm_string)
    assert label == "FIRST CUE HERE SECOND CUE" and (t0, t1) == (0.5, 4.0)
    lengths = {}

    def chunk(start, end, seed, T):
        g = torch.Generator().manual_seed(seed)
        smp = (torch.randint(0, 3, (1, T, 88, 88), generator=g).float(), torch.randint(0, 5, (104, T), generator=g).float())
        return {"start_time": start, "end_time": end, "frames": T, "load": (lambda s=smp: s)}

    videos = {"video_0": {"label": vtt, "asd_chunk": [chunk(2.0, 4.0, 1, 9), chunk(0.5, 2.0, 2, 7), chunk(9.0, 12.0, 3, 5)],
                          "fixed_chunk": [chunk(0.0, 4.5, 4, 11)], "gold_chunk": []},
              "video_1": {"label": vtt.replace("first", "other words in"), "fixed_chunk": [chunk(0.0, 3.0, 5, 6), chunk(3.0, 4.9, 6, 8)]}}
    to_text = lambda ids: " ".join(f"w{int(t)}" for t in ids)
    per_video, nw, avg = E.evaluate_avcocktail(_StandInModel(), videos, to_text, max_utts=2, max_frames=64)
    # sequential restatement: one inference per kept chunk, stitched by start time
    m = _StandInModel()
    for set_id, v in videos.items():
        lab, a, b = E.avcocktail_label(v["label"], text.norm_string)
        assert nw[set_id] == len(lab.split())
        for ct in E.CHUNK_TYPES:
            if ct not in v:
                assert ct not in per_video[set_id]
                continue
            outs, starts = [], []
            for ch in v[ct]:
                if ch["start_time"] + 1 < a or ch["end_time"] - 1 > b:
                    continue
                vid, aud = ch["load"]()
                hyp = m.infer_batch(vid.unsqueeze(0), aud.unsqueeze(0), [ch["frames"]])[0][0]
                outs.append(to_text(E.strip_sos_eos(hyp.yseq, m.eos)))
                starts.append(ch["start_time"])
            want = text.stitch_outputs(starts, outs) if outs else ""
            e, n = S.corpus_wer([lab], [want])
            assert per_video[set_id][ct] == pytest.approx(e / n)
    assert set(per_video["video_0"]) == {"asd_chunk", "fixed_chunk", "gold_chunk"} and per_video["video_0"]["gold_chunk"] == 1.0
    assert avg["fixed_chunk"] == pytest.approx((per_video["video_0"]["fixed_chunk"] * nw["video_0"] + per_video["video_1"]["fixed_chunk"] * nw["video_1"])
                                               / (nw["video_0"] + nw["video_1"]))


_AVC_VTT = ("WEBVTT\n\n00:00:02.000 --> 00:00:04.000\nsecond cue\n\n00:00:00.500 --> 00:00:01.500\nfirst cue here\n\n"
            "00:00:05.000 --> 00:00:05.500\n\n")


def _avc_chunk(start, end, seed, T):
    g = torch.Generator().manual_seed(seed)
    smp = (torch.randint(0, 3, (1, T, 88, 88), generator=g).float(), torch.randint(0, 5, (104, T), generator=g).float())
    return {"start_time": start, "end_time": end, "frames": T, "load": (lambda s=smp: s)}


def _avc_videos():
    return {"video_0": {"label": _AVC_VTT, "asd_chunk": [_avc_chunk(2.0, 4.0, 1, 9), _avc_chunk(0.5, 2.0, 2, 7), _avc_chunk(9.0, 12.0, 3, 5)],
                        "fixed_chunk": [_avc_chunk(0.0, 4.5, 4, 11)], "gold_chunk": [_avc_chunk(0.4, 1.6, 7, 6), _avc_chunk(1.9, 4.2, 8, 10)]},
            "video_1": {"label": _AVC_VTT.replace("first", "other words in"),
                        "fixed_chunk": [_avc_chunk(0.0, 3.0, 5, 6), _avc_chunk(3.0, 4.9, 6, 8)], "asd_chunk": [_avc_chunk(0.2, 4.4, 9, 13)]}}


def _avc_to_text(ids):
    return " ".join(f"w{int(t)}" for t in ids)


def _avc_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q.put((rank,) + tuple(E.evaluate_avcocktail(_StandInModel(), _avc_videos(), _avc_to_text, max_utts=2, max_frames=64)))
    finally:
        dist.destroy_process_group()


def test_avcocktail_loop_world2_equals_single_process():
    """The (video, chunk type) units of eval_avcocktail sharded over two ranks (gloo): every rank stitches and scores from the
    gathered token ids and gets the single-process numbers (shard by chunk, per-video concatenation stays exact)."""
    single = E.evaluate_avcocktail(_StandInModel(), _avc_videos(), _avc_to_text, max_utts=2, max_frames=64)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_avc_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, per_video, nw, avg in got:
        assert nw == single[1]
        assert {k: {c: pytest.approx(v) for c, v in d.items()} for k, d in single[0].items()} == per_video
        assert {c: pytest.approx(v) for c, v in single[2].items()} == avg
