"""Golden vectors for BASELINE.json configs[0] from the UNMODIFIED reference (run in the build container only).

    python oracle/gen_golden_cfg0.py       # needs /root/reference; writes tests/golden/cfg0_T375.npz  (~10 min on 8 cores)

One synthetic 15 s utterance (T=375, seed 1234, SURVEY.md 8d "Cfg 1"): the reference encoder output and the reference
``BatchBeamSearch`` n-best at beam 3 and beam 5 on that output.  The oracle (oracle/avsr_oracle.py, KV-cache form) is run on
the same encoder output and its n-best is stored next to the reference's, together with whether they agree: at this length
the fp32 log-domain CTC forward variables sit near -3000 (one ulp = 2.4e-4), so agreement is an empirical fact about rounding
noise that the tests report rather than assume.
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


def main():
    from src.avhubert_avsr.avhubert_avsr_model import get_beam_search_decoder
    torch.set_num_threads(os.cpu_count())
    sd = synth.make_state_dict(0)
    ref = build_reference(sd)
    token_list = ["<blank>"] + [f"u{i}" for i in range(5047)] + ["<eos>"]
    T, seed = 375, 1234
    out = {"fingerprint": np.array(synth.fingerprint(sd)), "T": np.array(T), "seed": np.array(seed)}
    with torch.no_grad():
        video, audio = synth.make_inputs(seed, T)
        t0 = time.time()
        x_ref = ref.encoder(input_features=audio, video=video).last_hidden_state[0]
        print(f"reference encoder T={T}: {time.time() - t0:.1f}s", flush=True)
        x_or = O.encoder_forward(sd, audio, video)[0]
        out["enc_ref_vs_oracle_maxabs"] = np.array((x_ref - x_or).abs().max().item())
        out["enc"] = x_ref.numpy()
        for beam in (3, 5):
            bs = get_beam_search_decoder(ref, token_list, beam_size=beam)
            t0 = time.time()
            nbest = bs(x_ref)
            t_ref = time.time() - t0
            t0 = time.time()
            hyps = O.beam_search(sd, x_ref, beam, kv_cache=True)
            t_or = time.time() - t0
            n_cmp = sum(1 for h in nbest if float(h.score) > -1e8)
            same = [a.yseq.tolist() == b.yseq for a, b in list(zip(nbest, hyps))[:n_cmp]]
            print(f"beam {beam}: reference {t_ref:.1f}s, oracle(kv) {t_or:.1f}s, n-best identical: {same}, "
                  f"scores ref {[round(float(h.score), 4) for h in nbest[:n_cmp]]} oracle {[round(h.score, 4) for h in hyps[:n_cmp]]}",
                  flush=True)
            out[f"ref_seconds_b{beam}"] = np.array(t_ref)
            out[f"nbest_b{beam}_yseq"] = np.array([h.yseq.tolist() for h in nbest])
            out[f"nbest_b{beam}_score"] = np.array([float(h.score) for h in nbest], dtype=np.float64)
            out[f"nbest_b{beam}_dec"] = np.array([float(h.scores["decoder"]) for h in nbest])
            out[f"nbest_b{beam}_ctc"] = np.array([float(h.scores["ctc"]) for h in nbest])
            out[f"oracle_b{beam}_yseq"] = np.array([h.yseq for h in hyps[:n_cmp]])
            out[f"oracle_b{beam}_score"] = np.array([h.score for h in hyps[:n_cmp]], dtype=np.float64)
            out[f"oracle_matches_b{beam}"] = np.array(same)
    np.savez_compressed(os.path.join(GOLD, "cfg0_T375.npz"), **out)
    print("cfg0 golden written")


if __name__ == "__main__":
    main()
