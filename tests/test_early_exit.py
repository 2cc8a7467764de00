"""The early-exit regime of the beam search and the search options the default evaluation does not use.

Goldens: tests/golden/eos_early.npz (oracle/gen_golden_eos.py) and tests/golden/search_options.npz
(oracle/gen_golden_options.py), both n-bests of the UNMODIFIED reference BatchBeamSearch.  With the eos bias of the decoder's
output layer raised, hypotheses end mid-sequence, ``ended_hyps`` accumulates, the running set shrinks below the beam and the
loop stops through ``end_detect`` or because no hypothesis is left (src/nets/batch_beam_search.py:287-349,
src/nets/beam_search.py:363-376, src/nets/e2e_asr_common.py:18-48) - the regime every trained checkpoint runs in.

CPU tests pin the oracle on the goldens; GPU tests compare the CUDA path (graph and eager) with the goldens, check that a
mixed batch whose utterances stop at different positions equals the B=1 runs, and that the host stops queueing positions
within one graph replay of the last utterance's stop.
"""
import copy
import os

import numpy as np
import pytest
import torch

from avsr_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [(shift, T, beam) for shift in (9, 6) for T in (30, 125) for beam in (3, 5)]


@pytest.fixture(scope="module")
def eos_golden():
    g = np.load(os.path.join(HERE, "golden", "eos_early.npz"))
    return {k: g[k] for k in g.files}


@pytest.fixture(scope="module")
def opt_golden():
    g = np.load(os.path.join(HERE, "golden", "search_options.npz"))
    return {k: g[k] for k in g.files}


def _enc(eos_golden, golden, T):
    return torch.from_numpy(eos_golden["enc_T125"] if T == 125 else golden[f"enc_T{T}"])


def _shifted_sd(state_dict, shift):
    sd = dict(state_dict)
    b = sd["decoder.output_layer.bias"].clone()
    b[synth.EOS] += float(shift)
    sd["decoder.output_layer.bias"] = b
    return sd


def _check(hyps, g, key, names=("decoder", "ctc"), get=None):
    """hyps: objects with yseq / score (+ per-scorer parts through `get`) vs the golden n-best `key`, incl. the order."""
    ys, ln, sc = g[key + "_yseq"], g[key + "_len"], g[key + "_score"]
    n = int((sc > -1e8).sum())
    assert len(hyps) == len(sc) and n >= 1, (key, len(hyps), len(sc))
    for k in range(n):
        y = hyps[k].yseq.tolist() if hasattr(hyps[k].yseq, "tolist") else list(hyps[k].yseq)
        assert y == ys[k, :ln[k]].tolist(), (key, k)
        assert abs(float(hyps[k].score) - sc[k]) <= 1e-3 * ln[k], (key, k)
        if get is not None:
            for nm in names:
                tol = 1e-2 if nm == "ctc" else 1e-3
                want = g[f"{key}_{nm}"] if f"{key}_{nm}" in g else g[f"{key}_dec"]
                assert abs(get(hyps[k], nm) - want[k]) <= tol * ln[k], (key, k, nm)


# ----------------------------------------------------------------------------------------------------------- CPU: oracle
# the CPU suite re-checks the T = 30 cases and one T = 125 case on every run (the others take a minute each on the CPU; the
# generator compared all eight with the oracle when it wrote the golden, and AVSR_SLOW_TESTS=1 re-runs them here)
CPU_CASES = [c for c in CASES if c[1] == 30 or c == (6, 125, 3) or os.environ.get("AVSR_SLOW_TESTS") == "1"]


@pytest.mark.parametrize("shift,T,beam", CPU_CASES)
def test_oracle_early_exit_matches_reference(state_dict, golden, eos_golden, shift, T, beam):
    from oracle import avsr_oracle as O
    trace = []
    hyps = O.beam_search(_shifted_sd(state_dict, shift), _enc(eos_golden, golden, T), beam, kv_cache=True, trace=trace)
    key = f"s{shift}_T{T}_b{beam}"
    _check(hyps, eos_golden, key, get=lambda h, nm: h.dec_score if nm == "decoder" else h.ctc_score)
    assert len(trace) - 1 == int(eos_golden[key + "_stop"]) < T - 1          # stopped early, at the reference's position
    assert len(set(eos_golden[key + "_len"].tolist())) >= 1


def test_oracle_search_options_match_reference(state_dict, golden, opt_golden):
    from oracle import avsr_oracle as O
    for T in (12, 30):
        for beam in (3, 5):
            hyps = O.beam_search(state_dict, torch.from_numpy(golden[f"enc_T{T}"]), beam, ctc_weight=0.0)
            _check(hyps, opt_golden, f"dec_only_T{T}_b{beam}", names=("decoder",), get=lambda h, nm: h.dec_score)
    x = torch.from_numpy(golden["enc_T30"])
    for tag in ("half", "const7", "minlen", "ratio1"):
        mlr = float(opt_golden[f"{tag}_ratios"][0])
        for beam in (3, 5):
            _check(O.beam_search(state_dict, x, beam, maxlenratio=mlr), opt_golden, f"{tag}_b{beam}",
                   get=lambda h, nm: h.dec_score if nm == "decoder" else h.ctc_score)


# ----------------------------------------------------------------------------------------------------------- GPU
def _shifted_weights(gpu_model, shift):
    w = copy.copy(gpu_model.decoder_weights)           # shares every tensor but the output bias
    w.out_b = gpu_model.decoder_weights.out_b.clone()
    w.out_b[synth.EOS] += float(shift)
    return w


def _part(h, nm):
    return float(h.scores[nm])


@pytest.mark.gpu
@pytest.mark.parametrize("graph", [True, False])
@pytest.mark.parametrize("shift,T,beam", CASES)
def test_gpu_early_exit_matches_reference(gpu_model, golden, eos_golden, shift, T, beam, graph):
    from avsr_b200.beam_search import BatchedBeamSearch
    bs = BatchedBeamSearch(_shifted_weights(gpu_model, shift), beam_size=beam, use_graph=graph)
    nbest = bs(_enc(eos_golden, golden, T).cuda())
    key = f"s{shift}_T{T}_b{beam}"
    _check(nbest, eos_golden, key, get=_part)
    stop = int(eos_golden[key + "_stop"])
    assert bs.last_stop_positions == [stop]
    # the host reads the "still running" flag one replay late: at most two replays of slack behind the stop position
    assert bs.last_positions_queued <= stop + 1 + 2 * bs.POLL_EVERY
    assert bs.last_positions_queued < T or stop + 1 + 2 * bs.POLL_EVERY >= T


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["bf16x3", "fp32"])
def test_gpu_mixed_batch_stops_per_utterance(gpu_model, golden, eos_golden, precision):
    """Utterances of one batch end at different positions (and one of them through the maxlen branch): each equals its B=1
    golden, and the stop positions are the reference's."""
    from avsr_b200.beam_search import BatchedBeamSearch
    bs = BatchedBeamSearch(_shifted_weights(gpu_model, 9), beam_size=3, precision=precision)
    x30, x125, x12 = _enc(eos_golden, golden, 30).cuda(), _enc(eos_golden, golden, 125).cuda(), _enc(eos_golden, golden, 12).cuda()
    out = bs.decode_batch(torch.cat([x30, x125, x12, x30], 0).contiguous(), [30, 125, 12, 30])
    _check(out[0], eos_golden, "s9_T30_b3", get=_part)
    _check(out[1], eos_golden, "s9_T125_b3", get=_part)
    _check(out[3], eos_golden, "s9_T30_b3", get=_part)
    stops = list(bs.last_stop_positions)
    single = bs(x12)
    assert [h.yseq.tolist() for h in out[2]] == [h.yseq.tolist() for h in single]
    assert [float(h.score) for h in out[2]] == [float(h.score) for h in single]
    assert stops[2] == bs.last_stop_positions[0]
    assert stops[0] == stops[3] == int(eos_golden["s9_T30_b3_stop"]) and stops[1] == int(eos_golden["s9_T125_b3_stop"])


@pytest.mark.gpu
@pytest.mark.parametrize("beam", [3, 5])
def test_gpu_attention_only_search_matches_reference(gpu_model, golden, opt_golden, beam):
    from avsr_b200.beam_search import BatchedBeamSearch
    bs = BatchedBeamSearch(gpu_model.decoder_weights, beam_size=beam, ctc_weight=0.0)
    for T in (12, 30):
        nbest = bs(torch.from_numpy(golden[f"enc_T{T}"]).cuda())
        _check(nbest, opt_golden, f"dec_only_T{T}_b{beam}", names=("decoder",), get=_part)
        assert set(nbest[0].scores) == {"decoder"}


@pytest.mark.gpu
@pytest.mark.parametrize("beam", [3, 5])
def test_gpu_maxlenratio_minlenratio_match_reference(state_dict, gpu_model, golden, opt_golden, beam):
    from avsr_b200.beam_search import BatchedBeamSearch
    from oracle import avsr_oracle as O
    bs = BatchedBeamSearch(gpu_model.decoder_weights, beam_size=beam)
    x = torch.from_numpy(golden["enc_T30"]).cuda()
    for tag in ("half", "const7", "minlen", "ratio1"):
        mlr, mnr = [float(v) for v in opt_golden[f"{tag}_ratios"]]
        _check(bs(x, maxlenratio=mlr, minlenratio=mnr), opt_golden, f"{tag}_b{beam}", get=_part)
    with pytest.raises(RuntimeError):
        bs(x, maxlenratio=2.0)
    # maxlenratio != 0 switches end detection off (beam_search.py:369): with the eos-biased weights the search then runs until
    # no hypothesis is left; compared with the oracle (pinned on both pieces separately above)
    bs9 = BatchedBeamSearch(_shifted_weights(gpu_model, 9), beam_size=beam)
    got = bs9(x, maxlenratio=1.0)
    want = O.beam_search(_shifted_sd(state_dict, 9), x.cpu(), beam, maxlenratio=1.0)
    assert len(got) == len(want)
    for a, b in zip(got, want):
        if b.score > -1e8:
            assert a.yseq.tolist() == b.yseq
            assert abs(float(a.score) - b.score) <= 1e-3 * len(b.yseq)
