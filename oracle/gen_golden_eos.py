"""Golden vectors for the EARLY-EXIT regime of the beam search, from the UNMODIFIED reference (build container only).

    python oracle/gen_golden_eos.py        # needs /root/reference; writes tests/golden/eos_early.npz

Random-init weights never let ``<eos>`` win, so every other golden of this repo is a maxlen run.  Trained checkpoints stop
through ``end_detect`` instead: hypotheses end mid-sequence, ``ended_hyps`` accumulates, the number of running hyps drops
below the beam, and the loop breaks early (src/nets/batch_beam_search.py:287-349, src/nets/beam_search.py:363-376,
src/nets/e2e_asr_common.py:18-48).  To reach that regime with deterministic weights, the bias of ``<eos>`` in the decoder's
output layer is raised (``eos_bias_shift``): ``sd = make_state_dict(0); sd["decoder.output_layer.bias"][5048] += shift``.
The reference ``BatchBeamSearch`` is then run unmodified on reference encoder outputs (T = 30 from model_seed0.npz, T = 125
computed here) at beam 3 and 5; the number of ``search`` calls (= stop position + 1) is counted by wrapping the bound method.
The oracle is compared on the spot and the script aborts on a mismatch.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from avsr_b200 import synth  # noqa: E402
from oracle import avsr_oracle as O  # noqa: E402
from oracle.gen_golden import build_reference  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SHIFTS = (9.0, 6.0)
EOS = synth.EOS


def shifted_state_dict(shift: float):
    sd = synth.make_state_dict(0)
    sd["decoder.output_layer.bias"] = sd["decoder.output_layer.bias"].clone()
    sd["decoder.output_layer.bias"][EOS] += shift
    return sd


def main():
    from src.avhubert_avsr.avhubert_avsr_model import get_beam_search_decoder
    torch.set_num_threads(os.cpu_count())
    g = np.load(os.path.join(GOLD, "model_seed0.npz"))
    token_list = ["<blank>"] + [f"u{i}" for i in range(5047)] + ["<eos>"]
    out = {"shifts": np.array(SHIFTS)}
    with torch.no_grad():
        sd0 = synth.make_state_dict(0)
        ref0 = build_reference(sd0)
        video, audio = synth.make_inputs(1236, 125)
        t0 = time.time()
        x125 = ref0.encoder(input_features=audio, video=video).last_hidden_state[0]
        print(f"reference encoder T=125: {time.time() - t0:.1f}s", flush=True)
        out["enc_T125"] = x125.numpy()
        out["enc_T125_seed"] = np.array(1236)
        xs = {30: torch.from_numpy(g["enc_T30"]), 125: x125}
        del ref0
        for shift in SHIFTS:
            sd = shifted_state_dict(shift)
            ref = build_reference(sd)
            for T in (30, 125):
                for beam in (3, 5):
                    bs = get_beam_search_decoder(ref, token_list, beam_size=beam)
                    calls = [0]
                    inner = bs.search

                    def counted(running, x, _inner=inner, _c=calls):
                        _c[0] += 1
                        return _inner(running, x)

                    bs.search = counted                      # counts positions; the reference code itself is untouched
                    t0 = time.time()
                    nbest = bs(xs[T])
                    t_ref = time.time() - t0
                    stop = calls[0] - 1                       # index of the last position searched
                    hyps = O.beam_search(sd, xs[T], beam, kv_cache=True)
                    n_cmp = sum(1 for h in nbest if float(h.score) > -1e8)
                    assert n_cmp > 0 and len(hyps) == len(nbest), (len(hyps), len(nbest))
                    for a, b in list(zip(nbest, hyps))[:n_cmp]:
                        assert a.yseq.tolist() == b.yseq, (shift, T, beam, a.yseq.tolist(), b.yseq)
                        assert abs(float(a.score) - b.score) < 1e-3 * len(b.yseq)
                    lens = [len(h.yseq) for h in nbest]
                    print(f"shift {shift} T={T} beam={beam}: {len(nbest)} ended hyps (lengths {min(lens)}..{max(lens)}), stop position "
                          f"{stop} of {T}, ref {t_ref:.1f}s, oracle matches on {n_cmp}", flush=True)
                    key = f"s{int(shift)}_T{T}_b{beam}"
                    ml = max(lens)
                    out[key + "_yseq"] = np.array([h.yseq.tolist() + [-1] * (ml - len(h.yseq)) for h in nbest], dtype=np.int64)
                    out[key + "_len"] = np.array(lens)
                    out[key + "_score"] = np.array([float(h.score) for h in nbest], dtype=np.float64)
                    out[key + "_dec"] = np.array([float(h.scores["decoder"]) for h in nbest])
                    out[key + "_ctc"] = np.array([float(h.scores["ctc"]) for h in nbest])
                    out[key + "_stop"] = np.array(stop)
            del ref
    np.savez_compressed(os.path.join(GOLD, "eos_early.npz"), **out)
    print("eos_early golden written", len(out))


if __name__ == "__main__":
    main()
