"""Golden vectors for the search options of the reference that the default evaluation does not use (build container only).

    python oracle/gen_golden_options.py    # needs /root/reference; writes tests/golden/search_options.npz

* ``ctc_weight = 0.0``: the factory keeps only the decoder scorer (weight-0 scorers are dropped, src/nets/beam_search.py:72-75),
  so there is no partial scorer and no pre-beam (:100-104): attention-only search over the full vocabulary.
* ``maxlenratio`` > 0 and < 0 (src/nets/beam_search.py:349-354): shorter maxlen, eos appended there, and NO end detection (:369).
* ``minlenratio`` > 0: accepted, no effect on the result (:355-358).
All on the reference encoder outputs of tests/golden/model_seed0.npz, unmodified reference BatchBeamSearch; the oracle is
compared on the spot.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from avsr_b200 import synth  # noqa: E402
from oracle import avsr_oracle as O  # noqa: E402
from oracle.gen_golden import build_reference  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def save(out, key, nbest, names):
    lens = [len(h.yseq) for h in nbest]
    ml = max(lens)
    out[key + "_yseq"] = np.array([h.yseq.tolist() + [-1] * (ml - len(h.yseq)) for h in nbest], dtype=np.int64)
    out[key + "_len"] = np.array(lens)
    out[key + "_score"] = np.array([float(h.score) for h in nbest], dtype=np.float64)
    for n in names:
        out[key + "_" + n] = np.array([float(h.scores[n]) for h in nbest])


def main():
    from src.avhubert_avsr.avhubert_avsr_model import get_beam_search_decoder
    torch.set_num_threads(os.cpu_count())
    g = np.load(os.path.join(GOLD, "model_seed0.npz"))
    sd = synth.make_state_dict(0)
    ref = build_reference(sd)
    token_list = ["<blank>"] + [f"u{i}" for i in range(5047)] + ["<eos>"]
    out = {}
    with torch.no_grad():
        for T in (12, 30):
            x = torch.from_numpy(g[f"enc_T{T}"])
            for beam in (3, 5):
                bs = get_beam_search_decoder(ref, token_list, ctc_weight=0.0, beam_size=beam)
                assert list(bs.scorers) == ["decoder"] and not bs.do_pre_beam
                nbest = bs(x)
                hyps = O.beam_search(sd, x, beam, ctc_weight=0.0)
                assert [h.yseq.tolist() for h in nbest] == [h.yseq for h in hyps], (T, beam)
                assert all(abs(float(a.score) - b.score) < 1e-3 * len(b.yseq) for a, b in zip(nbest, hyps))
                save(out, f"dec_only_T{T}_b{beam}", nbest, ("decoder",))
                print(f"ctc_weight 0: T={T} beam={beam}: {len(nbest)} hyps, best score {float(nbest[0].score):.4f}, oracle matches")
        x = torch.from_numpy(g["enc_T30"])
        for tag, mlr, mnr in (("half", 0.5, 0.0), ("const7", -7.0, 0.0), ("minlen", 0.0, 0.3), ("ratio1", 1.0, 0.2)):
            for beam in (3, 5):
                bs = get_beam_search_decoder(ref, token_list, beam_size=beam)
                nbest = bs(x, maxlenratio=mlr, minlenratio=mnr)
                hyps = O.beam_search(sd, x, beam, maxlenratio=mlr)
                n_cmp = sum(1 for h in nbest if float(h.score) > -1e8)
                for a, b in list(zip(nbest, hyps))[:n_cmp]:
                    assert a.yseq.tolist() == b.yseq, (tag, beam)
                    assert abs(float(a.score) - b.score) < 1e-3 * len(b.yseq)
                save(out, f"{tag}_b{beam}", nbest, ("decoder", "ctc"))
                print(f"maxlenratio {mlr} minlenratio {mnr} beam={beam}: {len(nbest)} hyps of lengths {sorted(set(len(h.yseq) for h in nbest))}, "
                      f"oracle matches on {n_cmp}")
            out[f"{tag}_ratios"] = np.array([mlr, mnr])
    np.savez_compressed(os.path.join(GOLD, "search_options.npz"), **out)
    print("search_options golden written", len(out))


if __name__ == "__main__":
    main()
