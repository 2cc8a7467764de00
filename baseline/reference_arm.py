"""The reference arm of bench.py: the UNMODIFIED reference (quanpn90/avsr) run through its own public API.

``baseline/_ref/src`` is a verbatim copy of ``/root/reference/src`` made by ``tools/install_ref.sh`` (git-ignored; it travels
to the GPU box with the snapshot).  Nothing from ``avsr_b200`` is on this path except ``synth`` (the shared deterministic
state_dict and inputs, loaded with ``load_state_dict(strict=True)``): the model is the reference's ``AVHubertAVSR``, the
search is the reference's ``get_beam_search_decoder`` -> ``BatchBeamSearch`` (src/avhubert_avsr/avhubert_avsr_model.py:12-36),
called the way ``AVSRCocktailModel.inference`` calls them (script/evaluation.py:96-108).

One full utterance of the workload costs the reference minutes on a CPU (134 s on 8 vCPU), so a bench "step"
is a bounded SAMPLE of one utterance:
  * the full encoder forward (T frames),
  * ``init_hyp`` (CTC head + CTCPrefixScoreTH construction),
  * the stock ``BatchBeamSearch.search`` + ``post_process`` for ONE position at each of ``positions`` prefix lengths spread
    over 0 .. T-1.  A position in the middle of a search needs running hypotheses of that length; they are fabricated with
    the shapes the search itself would hold there (``beam`` hyps, per-layer decoder output caches [L-1, 1024], CTC forward
    variables [T, 2]) and random contents - cost does not depend on the values.  The reference re-embeds and re-projects the
    whole prefix at every position (transformer/decoder.py:153-183), so its cost per position grows with the prefix while
    the CTC time loop shrinks (ctc_prefix_score.py:150-161); sampling the whole range captures both.
The decode time of the utterance is the sum over all T positions of the piecewise-linear interpolation of the samples.
``validate()`` (run in the build container) compares this estimate with a real full search.
"""
from __future__ import annotations

import os
import sys
import time
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_DIR = os.path.join(HERE, "_ref")


def available() -> Optional[str]:
    """None if baseline/_ref holds the reference, else the reason it does not."""
    if not os.path.isdir(os.path.join(REF_DIR, "src", "nets")):
        return "baseline/_ref/src missing: run tools/install_ref.sh in the build container (needs /root/reference)"
    return None


def load_modules():
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    from src.avhubert_avsr.avhubert_avsr_model import AVHubertAVSR, get_beam_search_decoder
    from src.avhubert_avsr.configuration_avhubert_avsr import AVHubertAVSRConfig
    from src.nets.batch_beam_search import BatchHypothesis
    return AVHubertAVSR, AVHubertAVSRConfig, get_beam_search_decoder, BatchHypothesis


def build_model(sd: Dict[str, torch.Tensor], device="cpu", dtype=torch.float32):
    """AVHubertAVSR(AVHubertAVSRConfig()) with the shared deterministic weights; returns the E2E module (`.avsr`)."""
    AVHubertAVSR, AVHubertAVSRConfig, _, _ = load_modules()
    m = AVHubertAVSR(AVHubertAVSRConfig()).eval()
    m.avsr.load_state_dict(sd, strict=True)
    return m.avsr.to(device=device, dtype=dtype)


def token_list(V: int = 5049) -> List[str]:
    return ["<blank>"] + [f"u{i}" for i in range(V - 2)] + ["<eos>"]


def fabricate_running(bs, BatchHypothesis, x: torch.Tensor, pos: int, beam: int, seed: int = 0):
    """Running hypotheses as BatchBeamSearch holds them when it enters position `pos` (prefix length pos + 1 incl. sos)."""
    if pos == 0:
        return bs.init_hyp(x)
    g = torch.Generator(device="cpu").manual_seed(seed + pos)
    T, D = x.shape
    L = pos + 1
    V = bs.n_vocab
    dev, dt = x.device, x.dtype
    yseq = torch.randint(1, V - 1, (beam, L), generator=g)
    yseq[:, 0] = bs.sos
    n_layers = len(bs.scorers["decoder"].decoders)
    dec_states = [[torch.randn(L - 1, D, generator=g).to(dev, dt) for _ in range(n_layers)] for _ in range(beam)]
    ctc_states = []
    for _ in range(beam):
        r = (-(torch.rand(T, 2, generator=g) * 40.0 + 1.0) * torch.arange(1, T + 1).unsqueeze(1) / T * 8.0).to(dev, dt)
        s = (-(torch.rand((), generator=g) * 50.0 + 5.0)).to(dev, dt).expand(V)
        ctc_states.append((r, s, 0, 0))
    scores = -torch.rand(beam, generator=g) * 100.0
    return BatchHypothesis(yseq=yseq.to(dev), score=scores.to(dev), length=torch.full((beam,), L, dtype=torch.int64, device=dev),
                           scores={"decoder": scores.clone().to(dev), "ctc": scores.clone().to(dev)},
                           states={"decoder": dec_states, "ctc": ctc_states})


def _sync(dev):
    if torch.device(dev).type == "cuda":
        torch.cuda.synchronize()


def sample_utterance(model, bs, BatchHypothesis, video: torch.Tensor, audio: torch.Tensor, beam: int,
                     positions: Sequence[int]) -> dict:
    """One bounded sample: encoder + init_hyp + one stock search/post_process at each of `positions`.  Returns the timings
    and the extrapolated seconds for the whole utterance."""
    dev = video.device
    T = audio.shape[-1]
    with torch.no_grad():
        _sync(dev)
        t0 = time.perf_counter()
        x = model.encoder(input_features=audio, video=video).last_hidden_state.squeeze(0)
        _sync(dev)
        t_enc = time.perf_counter() - t0
        t0 = time.perf_counter()
        bs.init_hyp(x)
        _sync(dev)
        t_init = time.perf_counter() - t0
        per_pos = []
        for p in positions:
            running = fabricate_running(bs, BatchHypothesis, x, p, beam)      # also (re)builds the CTC scorer for x at p == 0
            if p != 0:
                bs.init_hyp(x)
            _sync(dev)
            t0 = time.perf_counter()
            best = bs.search(running, x)
            bs.post_process(p, T, 0.0, best, [])
            _sync(dev)
            per_pos.append(time.perf_counter() - t0)
    t_dec = float(np.interp(np.arange(T), np.asarray(positions, dtype=np.float64), np.asarray(per_pos)).sum())
    return dict(t_enc=t_enc, t_init=t_init, per_pos=per_pos, positions=list(positions), t_dec=t_dec, t_utt=t_enc + t_init + t_dec,
                t_sample=t_enc + t_init + float(sum(per_pos)))


def default_positions(T: int, n: int = 5) -> List[int]:
    return sorted(set(int(round(v)) for v in np.linspace(0, T - 1, n)))


def full_utterance(model, bs, video, audio) -> dict:
    """One WHOLE utterance through the stock path (what AVSRCocktailModel.inference runs); minutes on a CPU."""
    dev = video.device
    with torch.no_grad():
        _sync(dev)
        t0 = time.perf_counter()
        x = model.encoder(input_features=audio, video=video).last_hidden_state.squeeze(0)
        _sync(dev)
        t_enc = time.perf_counter() - t0
        t0 = time.perf_counter()
        nbest = bs(x)
        _sync(dev)
        t_dec = time.perf_counter() - t0
    return dict(t_enc=t_enc, t_dec=t_dec, t_utt=t_enc + t_dec, yseq=nbest[0].yseq.tolist())


def validate(T: int = 100, beam: int = 3):
    """Build-container check of the sampling estimate against a real search of T positions (prints both)."""
    sys.path.insert(0, ROOT)
    from avsr_b200 import synth
    _, _, get_bs, BH = load_modules()
    torch.set_num_threads(os.cpu_count() or 1)
    sd = synth.make_state_dict(0)
    model = build_model(sd)
    bs = get_bs(model, token_list(), beam_size=beam)
    video, audio = synth.make_inputs(1234, T)
    full = full_utterance(model, bs, video, audio)
    est = sample_utterance(model, bs, BH, video, audio, beam, default_positions(T))
    print(f"T={T} beam={beam}: real search {full['t_dec']:.2f}s (+ encoder {full['t_enc']:.2f}s); sampled estimate "
          f"{est['t_dec']:.2f}s + init {est['t_init']:.2f}s (encoder {est['t_enc']:.2f}s); per-position samples "
          f"{[round(v * 1e3) for v in est['per_pos']]} ms at {est['positions']}")


if __name__ == "__main__":
    validate(int(sys.argv[1]) if len(sys.argv) > 1 else 100)
