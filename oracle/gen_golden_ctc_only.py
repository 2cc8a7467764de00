"""Generates tests/golden/ctc_only.npz: n-best of the UNMODIFIED reference beam search at ctc_weight = 1.0
(src/avhubert_avsr/avhubert_avsr_model.py:12-36: the decoder scorer is dropped, pre_beam_score_key = None, so every step
scores the full vocabulary with CTCPrefixScoreTH, ctc_prefix_score.py:115-119), from the reference encoder outputs already
stored in tests/golden/model_seed0.npz.  Run in the build container only (needs /root/reference)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from avsr_b200 import synth                                              # noqa: E402
from oracle.gen_golden import build_reference                            # noqa: E402

from src.avhubert_avsr.avhubert_avsr_model import get_beam_search_decoder   # noqa: E402

torch.set_num_threads(os.cpu_count())
g = np.load(os.path.join(ROOT, "tests", "golden", "model_seed0.npz"))
ref = build_reference(synth.make_state_dict(0))
token_list = ["<blank>"] + [f"u{i}" for i in range(5047)] + ["<eos>"]
out = {}
with torch.no_grad():
    for T in (12, 30):
        x = torch.from_numpy(g[f"enc_T{T}"])
        for beam in (3, 5):
            bs = get_beam_search_decoder(ref, token_list, ctc_weight=1.0, beam_size=beam)
            assert list(bs.scorers) == ["ctc"] and not bs.do_pre_beam
            t0 = time.time()
            nbest = bs(x)
            print(f"T={T} beam={beam}: {len(nbest)} hyps in {time.time() - t0:.1f}s; best {nbest[0].yseq.tolist()[:8]}... score {float(nbest[0].score):.4f}")
            ml = max(len(h.yseq) for h in nbest)
            out[f"nbest_T{T}_b{beam}_yseq"] = np.array([h.yseq.tolist() + [-1] * (ml - len(h.yseq)) for h in nbest], dtype=np.int64)
            out[f"nbest_T{T}_b{beam}_score"] = np.array([float(h.score) for h in nbest], dtype=np.float64)
            out[f"nbest_T{T}_b{beam}_ctc"] = np.array([float(h.scores["ctc"]) for h in nbest])
            out[f"nbest_T{T}_b{beam}_len"] = np.array([len(h.yseq) for h in nbest])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ctc_only.npz"), **out)
print("written")
