"""CPU: the oracle restatement is re-checked against the committed golden vectors, which were produced by running the
UNMODIFIED reference (oracle/gen_golden.py, build container only).  This is the parity pin of the oracle."""
import numpy as np
import torch

from avsr_b200 import synth
from oracle import avsr_oracle as O


def test_weight_generator_is_stable(state_dict, golden):
    # golden vectors are only meaningful if the deterministic weights are bit-identical to the ones they were made with
    assert abs(synth.fingerprint(state_dict) - float(golden["fingerprint"])) < 1e-6
    assert len(state_dict) == 699
    assert sum(v.numel() for v in state_dict.values()) == 428859374


def test_encoder_matches_reference_T12(state_dict, golden):
    video, audio = synth.make_inputs(1234, 12)
    taps = {}
    x = O.encoder_forward(state_dict, audio, video, taps)[0]
    assert np.abs(x.numpy() - golden["enc_T12"]).max() < 2e-4
    assert np.abs(taps["trunk"][0].numpy() - golden["trunk_T12"]).max() < 2e-4
    assert np.abs(taps["fused"][0].numpy() - golden["fused_T12"]).max() < 2e-4
    assert np.abs(taps["posconv"][0].numpy() - golden["posconv_T12"]).max() < 2e-4
    assert np.abs(taps["enc_layer0"][0].numpy() - golden["enc_layer0_T12"]).max() < 2e-4
    lp = O.ctc_log_softmax(state_dict, torch.from_numpy(golden["enc_T12"]).unsqueeze(0))[0]
    assert np.abs(lp[:, ::37].numpy() - golden["ctc_logp_T12"]).max() < 1e-4


def test_decoder_step_matches_reference(state_dict, golden):
    x = torch.from_numpy(golden["enc_T12"])
    ys = torch.from_numpy(golden["dec_step_ys"])
    caches = None
    for L in (1, 2, 3):
        lp, caches = O.decoder_batch_score(state_dict, ys[:, :L], caches, x.unsqueeze(0).expand(2, 12, 1024))
    assert np.abs(lp.numpy() - golden["dec_step_logp"]).max() < 1e-4
    # the KV-cache form equals the reference's output-cache form
    kv = O.KVDecoder(state_dict, x)
    st = None
    for p in range(3):
        lp2, st = kv.step(ys[:, p], p, st)
    assert np.abs(lp2.numpy() - golden["dec_step_logp"]).max() < 1e-4


def _check_nbest(hyps, golden, T, beam):
    yseq, score = golden[f"nbest_T{T}_b{beam}_yseq"], golden[f"nbest_T{T}_b{beam}_score"]
    n = int((score > -1e8).sum())
    assert n >= 1
    for k in range(n):
        assert hyps[k].yseq == yseq[k].tolist(), (T, beam, k)
        assert abs(hyps[k].score - score[k]) < 1e-3 * len(hyps[k].yseq)
        assert abs(hyps[k].dec_score - golden[f"nbest_T{T}_b{beam}_dec"][k]) < 1e-3 * len(hyps[k].yseq)
        assert abs(hyps[k].ctc_score - golden[f"nbest_T{T}_b{beam}_ctc"][k]) < 1e-2 * len(hyps[k].yseq)


def test_beam_search_matches_reference_T12(state_dict, golden):
    x = torch.from_numpy(golden["enc_T12"])
    for beam in (3, 5):
        _check_nbest(O.beam_search(state_dict, x, beam, kv_cache=True), golden, 12, beam)
    _check_nbest(O.beam_search(state_dict, x, 3, kv_cache=False), golden, 12, 3)


def test_beam_search_matches_reference_T30(state_dict, golden):
    x = torch.from_numpy(golden["enc_T30"])
    for beam in (3, 5):
        _check_nbest(O.beam_search(state_dict, x, beam, kv_cache=True), golden, 30, beam)


def test_beam_search_matches_reference_cfg0_T375(state_dict, golden_cfg0):
    """configs[0] at full size (T=375, beam 3): the oracle reproduces the reference n-best token for token."""
    g = golden_cfg0
    assert np.array_equal(g["fingerprint"], np.array(__import__("avsr_b200.synth", fromlist=["x"]).fingerprint(state_dict)))
    hyps = O.beam_search(state_dict, torch.from_numpy(g["enc"]), 3, kv_cache=True)
    score = g["nbest_b3_score"]
    n = int((score > -1e8).sum())
    assert n == 3
    for k in range(n):
        assert hyps[k].yseq == g["nbest_b3_yseq"][k].tolist(), k
        assert abs(hyps[k].score - score[k]) < 1e-3 * len(hyps[k].yseq)


def test_ctc_prefix_matches_reference(golden_ctc):
    """Replays the golden CTCPrefixScoreTH call sequences (pre-beam + full-vocabulary, repeated-token, blank-in-pre-beam
    and eos cases) through the oracle."""
    g = golden_ctc
    for tag in ("small", "vocab"):
        T, V, n_h, S = [int(v) for v in g[f"{tag}_shape"]]
        gen = torch.Generator().manual_seed(77)
        logp = torch.log_softmax(torch.randn(1, T, V, generator=gen) * 2.0, dim=-1)[0]
        if f"{tag}_logp" in g:
            assert np.abs(logp.numpy() - g[f"{tag}_logp"]).max() < 1e-6
        for mode in ("prebeam", "full"):
            rn, rb, sp = O.ctc_initial_state(logp)
            rn, rb = rn.unsqueeze(1), rb.unsqueeze(1)
            y = [[V - 1]]
            for step in range(4):
                cand = torch.from_numpy(g[f"{tag}_{mode}_cand{step}"]) if mode == "prebeam" else None
                assert [v[-1] for v in y] == g[f"{tag}_{mode}_last{step}"].tolist()
                sc, psi, rnn, rbb = O.ctc_prefix_scores(logp, rn, rb, sp, [v[-1] for v in y], len(y[0]) - 1, cand, 0, V - 1)
                ref = g[f"{tag}_{mode}_scores{step}"]
                live = ref > -1e9
                assert np.abs(sc.numpy()[live] - ref[live]).max() < 1e-3
                assert np.allclose(sc.numpy()[~live], ref[~live], rtol=1e-6)
                picks = g[f"{tag}_{mode}_picks{step}"]
                hs = torch.tensor(picks[:, 0])
                if cand is not None:
                    cols = torch.tensor([int((cand[h] == t).nonzero()[-1]) for h, t in picks])
                else:
                    cols = torch.tensor(picks[:, 1])
                rn, rb = rnn[:, hs, cols], rbb[:, hs, cols]
                sp = torch.stack([psi[h, t] for h, t in picks])
                y = [y[h] + [int(t)] for h, t in picks]


def test_end_detect():
    H = O.Hyp
    ended = [H([1] * 5, -1.0), H([1] * 6, -20.0), H([1] * 7, -25.0), H([1] * 8, -30.0)]
    assert O.end_detect(ended, 8)            # lengths 8,7,6 all > 10 below the best
    assert not O.end_detect(ended, 7)        # length 5 is the best itself (diff 0)
    assert not O.end_detect(ended[:3], 8)    # length 8 missing
    assert not O.end_detect([], 3)


def test_ctc_only_beam_search_matches_reference(state_dict, golden):
    """ctc_weight = 1.0 (decoder scorer dropped, full-vocabulary CTC prefix scoring every step): oracle vs the n-best of the
    unmodified reference (tests/golden/ctc_only.npz, oracle/gen_golden_ctc_only.py)."""
    import os
    import numpy as np
    import torch
    from oracle import avsr_oracle as O
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ctc_only.npz"))
    for T, beam in ((12, 3), (12, 5), (30, 3)):
        x = torch.from_numpy(golden[f"enc_T{T}"])
        hyps = O.beam_search(state_dict, x, beam, ctc_weight=1.0)
        ys, sc, ln = g[f"nbest_T{T}_b{beam}_yseq"], g[f"nbest_T{T}_b{beam}_score"], g[f"nbest_T{T}_b{beam}_len"]
        assert len(hyps) == len(sc)
        for k, h in enumerate(hyps):
            assert h.yseq == ys[k, :ln[k]].tolist(), (T, beam, k)
            assert abs(h.score - sc[k]) < 1e-3 * ln[k]
