"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    python oracle/gen_golden.py            # needs /root/reference; writes tests/golden/*.npz

The reference has no tests or golden vectors for this path, so parity is pinned on
outputs of the reference itself: its modules are imported from /root/reference, loaded with the
deterministic weights of ``avsr_b200.synth.make_state_dict`` (``load_state_dict(strict=True)``), and run
on seeded inputs.  While generating, every vector is also compared with ``oracle/avsr_oracle.py`` and
the script aborts on a mismatch.  /root/reference does not exist on the GPU box, so nothing at test
time imports this file.
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

GOLD = os.path.join(ROOT, "tests", "golden")


def build_reference(sd):
    from src.avhubert_avsr.avhubert_avsr_model import AVHubertAVSR
    from src.avhubert_avsr.configuration_avhubert_avsr import AVHubertAVSRConfig
    m = AVHubertAVSR(AVHubertAVSRConfig()).eval()
    missing, unexpected = m.avsr.load_state_dict(sd, strict=True), None
    return m.avsr


def gold_ctc_prefix():
    """CTCPrefixScoreTH known answers: random log-posteriors, 4 steps, pre-beam and full-vocab modes."""
    from src.nets.ctc_prefix_score import CTCPrefixScoreTH
    out = {}
    for tag, T, V, n_h, S in (("small", 23, 40, 3, 4), ("vocab", 31, 5049, 5, 7)):
        g = torch.Generator().manual_seed(77)
        logp = torch.log_softmax(torch.randn(1, T, V, generator=g) * 2.0, dim=-1)
        for mode in ("prebeam", "full"):
            impl = CTCPrefixScoreTH(logp.clone(), torch.tensor([T]), 0, V - 1)
            y = [[V - 1]] * 1
            state = None
            o_rn, o_rb, o_s = O.ctc_initial_state(logp[0])
            o_rn, o_rb = o_rn.unsqueeze(1), o_rb.unsqueeze(1)
            for step in range(4):
                n = len(y)
                cand = None
                if mode == "prebeam":
                    cand = torch.stack([torch.randperm(V - 1, generator=g)[:S] + (0 if step else 1) for _ in range(n)])
                    if step == 2:
                        cand[0, 1] = y[0][-1]          # repeated-token candidate
                        cand[-1, 0] = 0                # blank inside the pre-beam
                ys = [torch.tensor(v) for v in y]
                scores, new_state = impl(ys, state, cand)
                r, log_psi = new_state[0], new_state[1]
                o_scores, o_psi, o_rnn, o_rbb = O.ctc_prefix_scores(
                    logp[0], o_rn, o_rb, o_s, [v[-1] for v in y], len(y[0]) - 1, cand, 0, V - 1)
                err = (scores - o_scores).abs().max().item()
                assert err < 2e-4 * max(1.0, scores[scores > -1e9].abs().max().item()) or err < 1e-3, (tag, mode, step, err)
                assert torch.allclose(r[:, 0], o_rnn, atol=1e-3, rtol=1e-5) and torch.allclose(r[:, 1], o_rbb, atol=1e-3, rtol=1e-5)
                out[f"{tag}_{mode}_scores{step}"] = scores.numpy().copy()
                if cand is not None:
                    out[f"{tag}_{mode}_cand{step}"] = cand.numpy().copy()
                out[f"{tag}_{mode}_last{step}"] = np.array([v[-1] for v in y])
                # next hyps: pick n_h (hyp, token) pairs deterministically among scored candidates
                picks = []
                for j in range(n_h):
                    h = j % n
                    if cand is not None:
                        tok = int(cand[h, (j + step) % S])
                        if tok == 0:
                            tok = int(cand[h, (j + step + 1) % S])
                    else:
                        tok = 1 + (7 * j + 13 * step) % (V - 2)
                        if step == 1 and j == 0:
                            tok = y[h][-1]
                    picks.append((h, tok))
                out[f"{tag}_{mode}_picks{step}"] = np.array(picks)
                # reference state selection (scorers/ctc.py:40-63) vs oracle selection
                idmap = new_state[4]
                sel_r, sel_s = [], []
                for h, tok in picks:
                    col = idmap[h, tok] if idmap is not None else tok
                    sel_r.append(r[:, :, h, col])
                    sel_s.append(log_psi[h, tok].expand(V))
                state = (torch.stack(sel_r, dim=2), torch.stack(sel_s), 0, 0)
                hs = torch.tensor([p[0] for p in picks])
                if cand is not None:
                    cols = torch.tensor([int((cand[h] == tok).nonzero()[-1]) for h, tok in picks])
                else:
                    cols = torch.tensor([p[1] for p in picks])
                o_rn, o_rb = o_rnn[:, hs, cols], o_rbb[:, hs, cols]
                o_s = torch.stack([o_psi[h, tok] for h, tok in picks])
                y = [y[h] + [tok] for h, tok in picks]
        out[f"{tag}_logp"] = logp[0].numpy() if V < 100 else None
        out[f"{tag}_shape"] = np.array([T, V, n_h, S])
    out = {k: v for k, v in out.items() if v is not None}
    np.savez_compressed(os.path.join(GOLD, "ctc_prefix.npz"), **out)
    print("ctc_prefix golden written", len(out))


def gold_model():
    from src.avhubert_avsr.avhubert_avsr_model import get_beam_search_decoder
    sd = synth.make_state_dict(0)
    ref = build_reference(sd)
    token_list = ["<blank>"] + [f"u{i}" for i in range(5047)] + ["<eos>"]
    out = {"fingerprint": np.array(synth.fingerprint(sd))}
    with torch.no_grad():
        for T, seed in ((12, 1234), (30, 1235)):
            video, audio = synth.make_inputs(seed, T)
            t0 = time.time()
            x_ref = ref.encoder(input_features=audio, video=video).last_hidden_state[0]
            taps = {}
            x_or = O.encoder_forward(sd, audio, video, taps)[0]
            err = (x_ref - x_or).abs().max().item()
            print(f"T={T} encoder ref-vs-oracle max-abs {err:.3e}  ({time.time() - t0:.1f}s)")
            assert err < 2e-4, err
            out[f"enc_T{T}"] = x_ref.numpy()
            out[f"trunk_T{T}"] = taps["trunk"][0].numpy()
            out[f"fused_T{T}"] = taps["fused"][0].numpy()
            out[f"posconv_T{T}"] = taps["posconv"][0].numpy()
            out[f"enc_layer0_T{T}"] = taps["enc_layer0"][0].numpy()
            out[f"frontend3d_T{T}"] = taps["frontend3d"][:, ::8, ::3, ::3].numpy()   # subsampled [T,8,8,8]
            lp_ref = ref.ctc.log_softmax(x_ref.unsqueeze(0))[0]
            lp_or = O.ctc_log_softmax(sd, x_ref.unsqueeze(0))[0]
            assert (lp_ref - lp_or).abs().max().item() < 1e-4
            out[f"ctc_logp_T{T}"] = lp_ref[:, ::37].numpy()
            for beam in (3, 5):
                bs = get_beam_search_decoder(ref, token_list, beam_size=beam)
                t0 = time.time()
                nbest = bs(x_ref)
                t_ref = time.time() - t0
                for kv in (False, True):
                    hyps = O.beam_search(sd, x_ref, beam, kv_cache=kv)
                    n_cmp = sum(1 for h in nbest if float(h.score) > -1e8)
                    assert n_cmp > 0
                    for a, b in list(zip(nbest, hyps))[:n_cmp]:
                        assert a.yseq.tolist() == b.yseq, (T, beam, kv, a.yseq.tolist(), b.yseq)
                        assert abs(float(a.score) - b.score) < 1e-3 * len(b.yseq)
                print(f"T={T} beam={beam}: {len(nbest)} hyps, ref {t_ref:.1f}s, oracle matches (faithful + kv)")
                out[f"nbest_T{T}_b{beam}_yseq"] = np.array([h.yseq.tolist() for h in nbest])
                out[f"nbest_T{T}_b{beam}_score"] = np.array([float(h.score) for h in nbest], dtype=np.float64)
                out[f"nbest_T{T}_b{beam}_dec"] = np.array([float(h.scores["decoder"]) for h in nbest])
                out[f"nbest_T{T}_b{beam}_ctc"] = np.array([float(h.scores["ctc"]) for h in nbest])
        # one decoder scoring step, for the decoder kernels: prefix of 3 tokens, 2 hyps
        T = 12
        x_ref = torch.from_numpy(out["enc_T12"])
        ys = torch.tensor([[5048, 17, 99], [5048, 4000, 3]])
        lp, _ = O.decoder_batch_score(sd, ys[:, :1], None, x_ref.unsqueeze(0).expand(2, T, 1024))
        st = None
        for L in (1, 2, 3):
            lp_ref, st = ref.decoder.batch_score(ys[:, :L], [None, None] if st is None else st,
                                                 x_ref.unsqueeze(0).expand(2, T, 1024))
        caches = None
        for L in (1, 2, 3):
            lp_or, caches = O.decoder_batch_score(sd, ys[:, :L], caches, x_ref.unsqueeze(0).expand(2, T, 1024))
        assert (lp_ref - lp_or).abs().max().item() < 1e-4
        out["dec_step_ys"] = ys.numpy()
        out["dec_step_logp"] = lp_ref.numpy()
    np.savez_compressed(os.path.join(GOLD, "model_seed0.npz"), **out)
    print("model golden written")


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    gold_ctc_prefix()
    gold_model()
