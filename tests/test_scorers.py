"""The scorer plug-in API (avsr_b200/scorers.py) under the REFERENCE's own beam search.

SURVEY.md 8(b): the reference's ``BatchBeamSearch`` (src/nets/batch_beam_search.py:138-285) drives its scorers through
``batch_init_state / batch_score / batch_score_partial / select_state`` (src/nets/scorer_interface.py:9-186).  Here that
search - the unmodified reference from ``baseline/_ref`` (tools/install_ref.sh) when present, and always a 40-line restatement
of its ``search`` loop over the same API - runs with the B200 scorers swapped in and must produce the golden n-best.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def _check_nbest(nbest, golden, T, beam, get):
    yseq, score = golden[f"nbest_T{T}_b{beam}_yseq"], golden[f"nbest_T{T}_b{beam}_score"]
    n = int((score > -1e8).sum())
    assert len(nbest) >= n >= 1
    for k in range(n):
        y, sc, parts = get(nbest[k])
        assert y == yseq[k].tolist(), (T, beam, k)
        assert abs(sc - score[k]) <= 1e-3 * len(y)
        assert abs(parts["decoder"] - golden[f"nbest_T{T}_b{beam}_dec"][k]) <= 1e-3 * len(y)
        assert abs(parts["ctc"] - golden[f"nbest_T{T}_b{beam}_ctc"][k]) <= 1e-2 * len(y)


def mini_batch_beam_search(scorers, weights, x, beam, V, sos, eos, pre_beam_ratio=1.5):
    """BatchBeamSearch.forward / search / post_process restated over the scorer API only (no end detection: the goldens
    used here are maxlen runs).  Hypothesis = dict(yseq, score, scores, states)."""
    full = {k: v for k, v in scorers.items() if not hasattr(v, "batch_score_partial")}
    part = {k: v for k, v in scorers.items() if hasattr(v, "batch_score_partial")}
    S = int(pre_beam_ratio * beam)
    run = [dict(yseq=[sos], score=0.0, scores={k: 0.0 for k in scorers}, states={k: v.batch_init_state(x) for k, v in scorers.items()})]
    ended, T = [], x.shape[0]
    for i in range(T):
        n = len(run)
        ys = torch.tensor([h["yseq"] for h in run], dtype=torch.int64, device=x.device)
        w = torch.zeros(n, V, dtype=x.dtype, device=x.device)
        sc, st = {}, {}
        for k, d in full.items():
            sc[k], st[k] = d.batch_score(ys, [h["states"][k] for h in run], x.expand(n, *x.shape))
            w += weights[k] * sc[k]
        ids = torch.topk(sc["decoder"], S, dim=-1)[1]
        for k, d in part.items():
            sc[k], st[k] = d.batch_score_partial(ys, ids, [h["states"][k] for h in run], x)
            w += weights[k] * sc[k]
        w += torch.tensor([h["score"] for h in run], dtype=x.dtype, device=x.device).unsqueeze(1)
        top = w.view(-1).topk(beam)[1]
        new = []
        for p, t in zip(torch.div(top, V, rounding_mode="trunc").tolist(), (top % V).tolist()):
            states = {k: full[k].select_state(st[k], p) for k in full}
            states.update({k: part[k].select_state(st[k], p, t) for k in part})
            new.append(dict(yseq=run[p]["yseq"] + [t], score=float(w[p, t]),
                            scores={k: run[p]["scores"][k] + float(sc[k][p, t]) for k in scorers}, states=states))
        if i == T - 1:
            for h in new:
                h["yseq"] = h["yseq"] + [eos]
        ended += [h for h in new if h["yseq"][-1] == eos]
        run = [h for h in new if h["yseq"][-1] != eos]
        if not run:
            break
    return sorted(ended, key=lambda h: h["score"], reverse=True)


def test_scorer_classes_follow_the_interface():
    """CPU: the classes exist, expose the reference's method names and derive from a ScorerInterface."""
    from avsr_b200 import scorers as SC
    dec_cls, ctc_cls = SC.scorer_classes()
    for name in ("init_state", "batch_init_state", "batch_score", "select_state", "final_score"):
        assert callable(getattr(dec_cls, name))
    for name in ("batch_init_state", "batch_score_partial", "select_state", "final_score"):
        assert callable(getattr(ctc_cls, name))
    full_base, part_base = SC.reference_interfaces()
    assert issubclass(dec_cls, full_base) and issubclass(ctc_cls, part_base)
    assert not hasattr(dec_cls, "batch_score_partial")          # the decoder is a FULL scorer (beam_search.py:83-86)


@pytest.mark.gpu
@pytest.mark.parametrize("T,beam", [(12, 3), (12, 5), (30, 3)])
def test_mini_driver_with_b200_scorers(gpu_model, golden, T, beam):
    x = torch.from_numpy(golden[f"enc_T{T}"]).cuda()
    scorers = {"decoder": gpu_model.decoder, "ctc": gpu_model.ctc_prefix_scorer()}
    nbest = mini_batch_beam_search(scorers, {"decoder": 0.9, "ctc": 0.1}, x, beam, 5049, 5048, 5048)
    _check_nbest(nbest, golden, T, beam, lambda h: (h["yseq"], h["score"], h["scores"]))


@pytest.mark.gpu
def test_ctc_head_matches_reference_posteriors(gpu_model, golden):
    x = torch.from_numpy(golden["enc_T12"]).cuda()
    lp = gpu_model.ctc.log_softmax(x.unsqueeze(0))
    assert tuple(lp.shape) == (1, 12, 5049)
    assert np.abs(lp[0, :, ::37].cpu().numpy() - golden["ctc_logp_T12"]).max() < 1e-4


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "nets")), reason="baseline/_ref missing (tools/install_ref.sh)")
@pytest.mark.parametrize("which", ["both", "decoder_only", "ctc_only", "reference_factory"])
def test_reference_batch_beam_search_drives_b200_scorers(state_dict, gpu_model, golden, which):
    """The UNMODIFIED reference BatchBeamSearch with B200 scorers: both swapped in, each one alone next to the reference's
    own torch scorer on the GPU (the bisect configurations), and through the reference's get_beam_search_decoder(model, ...)
    with the B200 model object (model.decoder + the reference's CTCPrefixScorer over model.ctc)."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from src.avhubert_avsr.avhubert_avsr_model import get_beam_search_decoder
    from src.nets.batch_beam_search import BatchBeamSearch
    from src.nets.scorers.ctc import CTCPrefixScorer
    from src.nets.scorers.length_bonus import LengthBonus
    from avsr_b200 import scorers as SC
    from avsr_b200.scorers import B200CTCPrefixScorer, B200DecoderScorer
    T, beam = 12, 3
    token_list = ["<blank>"] + [f"u{i}" for i in range(5047)] + ["<eos>"]
    x = torch.from_numpy(golden[f"enc_T{T}"]).cuda()
    w = gpu_model.decoder_weights
    get = lambda h: (h.yseq.tolist(), float(h.score), {k: float(v) for k, v in h.scores.items()})
    if which == "reference_factory":
        bs = get_beam_search_decoder(gpu_model, token_list, ctc_weight=0.1, beam_size=beam)
        assert isinstance(bs.scorers["decoder"], SC.reference_interfaces()[0])
    else:
        if which in ("both", "decoder_only"):
            dec = B200DecoderScorer(w, "cuda:0")
        else:
            sys.path.insert(0, ROOT)
            from baseline import reference_arm as RA
            dec = RA.build_model(state_dict, device="cuda").decoder
        if which in ("both", "ctc_only"):
            ctc = B200CTCPrefixScorer(w, "cuda:0", eos=5048)
        else:
            ctc = CTCPrefixScorer(gpu_model.ctc, 5048)
        bs = BatchBeamSearch(beam_size=beam, vocab_size=5049, weights={"decoder": 0.9, "ctc": 0.1, "lm": 0.0, "length_bonus": 0.0},
                             scorers={"decoder": dec, "ctc": ctc, "length_bonus": LengthBonus(5049), "lm": None}, sos=5048, eos=5048,
                             token_list=token_list, pre_beam_score_key="decoder")
    with torch.no_grad():
        nbest = bs(x)
    _check_nbest(nbest, golden, T, beam, get)
