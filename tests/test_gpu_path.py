"""GPU: the hot path through the public API (avsr_b200.model / encoder / beam_search -> C ABI) against (i) the golden
vectors produced by the unmodified reference and (ii) the CPU oracle on the same seeded inputs.

Tolerances: bf16 encoder vs fp32 reference: rel-RMSE <= 1.5e-2, max-abs <= 0.10, cosine >= 0.9999;
decode: identical token sequences for every n-best entry with score > -1e8, |d score| <= 1e-3 * len.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from avsr_b200 import synth

pytestmark = pytest.mark.gpu


def _enc_metrics(x, ref):
    d = (x - ref).double()
    return dict(max_abs=d.abs().max().item(), rel_rmse=(d.pow(2).mean().sqrt() / ref.double().pow(2).mean().sqrt()).item(),
                cos=torch.nn.functional.cosine_similarity(x.double().flatten(), ref.double().flatten(), dim=0).item())


def _check_enc(x, ref, what):
    m = _enc_metrics(x, ref)
    assert m["max_abs"] <= 0.10 and m["rel_rmse"] <= 1.5e-2 and m["cos"] >= 0.9999, (what, m)


@pytest.mark.parametrize("T,seed", [(12, 1234), (30, 1235)])
def test_encoder_vs_reference_golden(gpu_model, golden, T, seed):
    video, audio = synth.make_inputs(seed, T)
    taps = {}
    enc = gpu_model.encoder
    x = enc.forward_packed(video[0, 0].cuda().contiguous(), audio.cuda(), [T], taps).cpu()
    # stage by stage first, so that a failure names the kernel that broke
    tr, tr_ref = taps["trunk"].cpu(), torch.from_numpy(golden[f"trunk_T{T}"])
    assert (tr - tr_ref).abs().max().item() <= 0.02 * tr_ref.abs().max().item() + 0.02, ("trunk", _enc_metrics(tr, tr_ref))
    for name in ("fused", "posconv", "enc_layer0"):
        got, ref = taps[name].cpu(), torch.from_numpy(golden[f"{name}_T{T}"])
        m = _enc_metrics(got, ref)
        assert m["rel_rmse"] <= 1.5e-2, (name, m)
    _check_enc(x, torch.from_numpy(golden[f"enc_T{T}"]), f"encoder T={T}")


def test_encoder_vs_reference_cfg0_T375(gpu_model, golden_cfg0):
    """BASELINE.json configs[0] at full size: the bf16 encoder on the 375-frame utterance against the reference's fp32 encoder
    output (tests/golden/cfg0_T375.npz), alone and as utterance 0 of a 4-utterance batch."""
    video, audio = synth.make_inputs(1234, 375)
    ref = torch.from_numpy(golden_cfg0["enc"])
    x = gpu_model.encoder(input_features=audio.cuda(), video=video.cuda()).last_hidden_state[0]
    _check_enc(x.cpu(), ref, "encoder T=375")
    vids, auds = zip(*[synth.make_inputs(1234 + i, 375) for i in range(4)])
    xb = gpu_model.encoder(input_features=torch.cat(auds, 0).cuda(), video=torch.cat(vids, 0).cuda()).last_hidden_state
    _check_enc(xb[0].cpu(), ref, "encoder T=375, batch of 4")


def test_implicit_posconv_and_frontend_equal_explicit_paths(gpu_model):
    """The implicit-GEMM forms of the positional conv (one launch, per-utterance tensor maps) and of the 3D-conv frontend give
    the results of the explicit patch-matrix paths on a mixed-length batch that spans several 128-frame work items."""
    enc = gpu_model.encoder
    lengths = [200, 3, 131]
    vids, auds = zip(*[synth.make_inputs(60 + i, t) for i, t in enumerate(lengths)])
    video = torch.cat([v[0, 0] for v in vids], 0).cuda().contiguous()
    audio = torch.zeros(3, 104, max(lengths))
    for b, a in enumerate(auds):
        audio[b, :, :lengths[b]] = a[0]
    audio = audio.cuda()
    outs = {}
    for imp in (True, False):
        enc.implicit_posconv = enc.implicit_frontend = imp
        taps = {}
        x = enc.forward_packed(video, audio, lengths, taps)
        outs[imp] = (taps["trunk"].clone(), taps["fused"].clone(), taps["posconv"].clone(), x.clone())
    enc.implicit_posconv = enc.implicit_frontend = True
    for a, b, name in zip(outs[True], outs[False], ("trunk", "fused", "posconv", "output")):
        d = (a - b).abs().max().item()
        assert d <= 2e-2 * max(1.0, b.abs().max().item()), (name, d)
    # the two forms accumulate in a different order but from the same bf16 operands: the positional conv itself agrees tightly
    pc_t, pc_f = outs[True][2] - outs[True][1], outs[False][2] - outs[False][1]
    assert (pc_t - pc_f).abs().max().item() < 5e-3 * max(1.0, pc_f.abs().max().item())


def test_halo_layer1_equals_dense_layer1(gpu_model):
    """ResNet layer1 on the padded layout (halo convolution, pitched max-pool, pitched entry of layer2) gives bit-identical trunk
    features to the dense implicit-GEMM path, on a batch that is not a multiple of anything."""
    enc = gpu_model.encoder
    lengths = [77, 5, 40]
    vids, auds = zip(*[synth.make_inputs(80 + i, t) for i, t in enumerate(lengths)])
    video = torch.cat([v[0, 0] for v in vids], 0).cuda().contiguous()
    audio = torch.zeros(3, 104, max(lengths))
    for b, a in enumerate(auds):
        audio[b, :, :lengths[b]] = a[0]
    audio = audio.cuda()
    outs = {}
    try:
        for halo in (True, False):
            enc.halo_conv = halo
            taps = {}
            x = enc.forward_packed(video, audio, lengths, taps)
            outs[halo] = (torch.cat(taps["frontend3d"], 0).clone(), taps["trunk"].clone(), x.clone())
    finally:
        enc.halo_conv = True
    for a, b, name in zip(outs[True], outs[False], ("pooled frontend", "trunk", "output")):
        assert torch.equal(a, b), name


def _check_nbest(nbest, golden, T, beam):
    yseq, score = golden[f"nbest_T{T}_b{beam}_yseq"], golden[f"nbest_T{T}_b{beam}_score"]
    n = int((score > -1e8).sum())
    assert len(nbest) >= n >= 1
    for k in range(n):
        assert nbest[k].yseq.tolist() == yseq[k].tolist(), (T, beam, k, nbest[k].yseq.tolist(), yseq[k].tolist())
        ln = len(yseq[k])
        assert abs(float(nbest[k].score) - score[k]) <= 1e-3 * ln
        assert abs(float(nbest[k].scores["decoder"]) - golden[f"nbest_T{T}_b{beam}_dec"][k]) <= 1e-3 * ln
        assert abs(float(nbest[k].scores["ctc"]) - golden[f"nbest_T{T}_b{beam}_ctc"][k]) <= 1e-2 * ln


@pytest.mark.parametrize("T", [12, 30])
@pytest.mark.parametrize("beam", [3, 5])
@pytest.mark.parametrize("graph,precision", [(False, "fp32"), (True, "fp32"), (False, "bf16x3"), (True, "bf16x3")])
def test_beam_search_vs_reference_golden(state_dict, gpu_model, golden, T, beam, graph, precision):
    """Decode from the reference's own fp32 encoder output: token-identical n-best at beam 3 and 5, on both numerics paths
    (fp32 CUDA-core FMA and the three-term bf16 split on the tensor cores)."""
    from avsr_b200.beam_search import BatchedBeamSearch
    bs = BatchedBeamSearch(gpu_model.decoder_weights, beam_size=beam, use_graph=graph, precision=precision)
    x = torch.from_numpy(golden[f"enc_T{T}"]).cuda()
    nbest = bs(x)
    _check_nbest(nbest, golden, T, beam)
    h = nbest[0]
    assert h.yseq.dtype == torch.int64 and h.yseq[0].item() == 5048 and h.yseq[-1].item() == 5048
    d = h.asdict()
    assert isinstance(d["yseq"], list) and isinstance(d["score"], float) and set(d["scores"]) == {"decoder", "ctc"}


@pytest.mark.parametrize("beam,graph", [(3, True), (5, False)])
def test_chained_projections_match_reference_golden(gpu_model, golden, beam, graph):
    """Opt-in variant of the bf16x3 position (avsr_gemm_x3_chain: a projection launch first finishes the rows of the previous
    projection): same token-identical n-best as the default sequence of launches."""
    from avsr_b200.beam_search import BatchedBeamSearch
    bs = BatchedBeamSearch(gpu_model.decoder_weights, beam_size=beam, use_graph=graph)
    bs.chain = True
    for T in (30, 12):
        _check_nbest(bs(torch.from_numpy(golden[f"enc_T{T}"]).cuda()), golden, T, beam)


@pytest.mark.parametrize("beam", [3, 5])
def test_beam_search_vs_reference_cfg0_T375(gpu_model, golden_cfg0, beam):
    """BASELINE.json configs[0] at full size: T=375 frames, decode from the reference's encoder output, n-best
    token-identical to the reference's BatchBeamSearch at beam 3 and 5 (377-token hypotheses)."""
    from avsr_b200.beam_search import BatchedBeamSearch
    g = golden_cfg0
    bs = BatchedBeamSearch(gpu_model.decoder_weights, beam_size=beam)
    nbest = bs(torch.from_numpy(g["enc"]).cuda())
    score = g[f"nbest_b{beam}_score"]
    n = int((score > -1e8).sum())
    assert len(nbest) >= n >= beam
    for k in range(n):
        assert nbest[k].yseq.tolist() == g[f"nbest_b{beam}_yseq"][k].tolist(), (beam, k)
        ln = len(g[f"nbest_b{beam}_yseq"][k])
        assert abs(float(nbest[k].score) - score[k]) <= 1e-3 * ln
        assert abs(float(nbest[k].scores["decoder"]) - g[f"nbest_b{beam}_dec"][k]) <= 1e-3 * ln
        assert abs(float(nbest[k].scores["ctc"]) - g[f"nbest_b{beam}_ctc"][k]) <= 1e-2 * ln


def test_batched_decode_equals_single_runs(gpu_model, golden):
    """Mixed-length batch: every utterance must evolve exactly as its own B=1 run."""
    from avsr_b200.beam_search import BatchedBeamSearch
    bs = BatchedBeamSearch(gpu_model.decoder_weights, beam_size=3)
    x12, x30 = torch.from_numpy(golden["enc_T12"]).cuda(), torch.from_numpy(golden["enc_T30"]).cuda()
    out = bs.decode_batch(torch.cat([x30, x12, x30[:20].contiguous()], 0), [30, 12, 20])
    _check_nbest(out[0], golden, 30, 3)
    _check_nbest(out[1], golden, 12, 3)
    single = bs(x30[:20].contiguous())
    assert [h.yseq.tolist() for h in out[2]] == [h.yseq.tolist() for h in single]
    assert all(abs(float(a.score) - float(b.score)) < 1e-4 for a, b in zip(out[2], single))


def test_concurrent_decode_groups_equal_single_runs(gpu_model, golden):
    """decode_batch with the utterances split into groups on separate streams (n_groups > 1): same n-best as one chain."""
    from avsr_b200.beam_search import BatchedBeamSearch
    bs = BatchedBeamSearch(gpu_model.decoder_weights, beam_size=3)
    bs.n_groups = 2
    x12, x30 = torch.from_numpy(golden["enc_T12"]).cuda(), torch.from_numpy(golden["enc_T30"]).cuda()
    out = bs.decode_batch(torch.cat([x30, x12, x12, x30], 0), [30, 12, 12, 30])
    assert len(bs.last_sessions) == 2
    for o, T in zip(out, (30, 12, 12, 30)):
        _check_nbest(o, golden, T, 3)


def test_batched_encoder_equals_single_runs(gpu_model):
    """Padded mixed-length batch == per-utterance runs (temporal conv / pos-conv / attention stop at utterance ends)."""
    enc = gpu_model.encoder
    v1, a1 = synth.make_inputs(7, 20)
    v2, a2 = synth.make_inputs(8, 9)
    video = torch.zeros(2, 1, 20, 88, 88)
    audio = torch.zeros(2, 104, 20)
    video[0], audio[0] = v1[0], a1[0]
    video[1, :, :9], audio[1, :, :9] = v2[0], a2[0]
    video[1, :, 9:] = 5.0                     # garbage in the padding must be invisible
    audio[1, :, 9:] = -3.0
    both = enc(input_features=audio.cuda(), video=video.cuda(), lengths=[20, 9]).last_hidden_state.cpu()
    s1 = enc(input_features=a1.cuda(), video=v1.cuda()).last_hidden_state.cpu()[0]
    s2 = enc(input_features=a2.cuda(), video=v2.cuda()).last_hidden_state.cpu()[0]
    assert (both[0] - s1).abs().max().item() < 2e-2
    assert (both[1, :9] - s2).abs().max().item() < 2e-2
    assert both[1, 9:].abs().max().item() == 0.0


def test_full_path_vs_oracle(state_dict, gpu_model):
    """Encoder (bf16) -> beam search end to end through the evaluation-style call, against the CPU oracle."""
    from oracle import avsr_oracle as O
    T = 25
    video, audio = synth.make_inputs(4321, T)
    x_ref = O.encoder_forward(state_dict, audio, video)[0]
    x = gpu_model.encoder(input_features=audio.cuda(), video=video.cuda()).last_hidden_state[0]
    _check_enc(x.cpu(), x_ref, "encoder T=25")
    ref = O.beam_search(state_dict, x_ref, 3, kv_cache=True)
    nbest = gpu_model.beam_search(x_ref.cuda())
    for a, b in zip(nbest, ref):
        if b.score > -1e8:
            assert a.yseq.tolist() == b.yseq
            assert abs(float(a.score) - b.score) < 1e-3 * len(b.yseq)
    # evaluation-style call (script/evaluation.py:96-108): token ids without sos
    ids = gpu_model.inference(video.cuda(), audio.cuda())
    assert isinstance(ids, list) and ids[-1] == 5048 and len(ids) == T + 1


def test_ctc_prefix_kernels_vs_reference_golden(golden_ctc):
    """avsr_ctc_prefix_prebeam / avsr_ctc_prefix_full driven with the golden CTCPrefixScoreTH call sequences."""
    from avsr_b200 import _lib as L
    lib = L.load()
    g = golden_ctc
    LOGZERO = -1e10
    for tag in ("small", "vocab"):
        T, V, n_h, S = [int(v) for v in g[f"{tag}_shape"]]
        gen = torch.Generator().manual_seed(77)
        ldp = (V + 3) // 4 * 4
        logp = torch.zeros(T, ldp)
        logp[:, :V] = torch.log_softmax(torch.randn(1, T, V, generator=gen) * 2.0, dim=-1)[0]
        logp = logp.cuda()
        beam, R = n_h, n_h
        i32 = lambda v: torch.tensor(v, dtype=torch.int32, device="cuda")
        utt_off, utt_T = i32([0]), i32([T])
        for mode in ("prebeam", "full"):
            Sk = S if mode == "prebeam" else 1
            r_buf = torch.full((2, R * Sk, T, 2), LOGZERO, device="cuda")
            rprev = i32([0] * R)
            s_prev = torch.zeros(R, device="cuda")
            n_hyp = 1
            for step in range(4):
                last = g[f"{tag}_{mode}_last{step}"].tolist()
                n_run, last_tok, step_t = i32([n_hyp]), i32(last + [0] * (R - n_hyp)), i32([step])
                ref = g[f"{tag}_{mode}_scores{step}"]
                picks = g[f"{tag}_{mode}_picks{step}"]
                psi = torch.zeros(R, Sk, device="cuda")
                rsum = torch.zeros(R, device="cuda")
                if mode == "prebeam":
                    cand = g[f"{tag}_{mode}_cand{step}"]
                    part = torch.zeros(R, S, dtype=torch.int32, device="cuda")
                    part[:n_hyp] = torch.from_numpy(cand).int().cuda()
                    L.check(lib.avsr_ctc_prefix_prebeam(L.ptr(logp), V, ldp, 0, L.ptr(utt_off), L.ptr(utt_T), L.ptr(n_run), beam, R, S,
                                                        L.ptr(last_tok), L.ptr(part), L.ptr(rprev), L.ptr(r_buf), T, L.ptr(step_t),
                                                        L.ptr(psi), L.ptr(rsum), L.stream()), "prebeam")
                    psi_c, rsum_c, sp = psi.cpu().numpy(), rsum.cpu().numpy(), s_prev.cpu().numpy()
                    for h in range(n_hyp):
                        for s_, c in enumerate(cand[h]):
                            want = ref[h, c]
                            got = (LOGZERO if c == 0 else (rsum_c[h] if c == V - 1 else psi_c[h, s_])) - sp[h]
                            assert abs(got - want) < 1e-3 or (want < -1e9 and got < -1e9), (tag, mode, step, h, c, got, want)
                        assert abs((rsum_c[h] - sp[h]) - ref[h, V - 1]) < 1e-3
                    new_psi = []
                    for h, t in picks:
                        col = int(np.nonzero(cand[h] == t)[0][-1])
                        new_psi.append((int(h) * S + col, float(psi_c[h, col])))
                    rprev = i32([p[0] for p in new_psi] + [0] * (R - len(new_psi)))
                    s_prev = torch.tensor([p[1] for p in new_psi] + [0.0] * (R - len(new_psi)), device="cuda")
                else:
                    scores = torch.zeros(R, V, device="cuda")
                    ncg, ts = C.c_int(0), C.c_int(0)
                    L.check(lib.avsr_ctc_prefix_full_plan(1, V, C.byref(ncg), C.byref(ts)), "plan")
                    assert ts.value == 16 and ncg.value == -(-V // 512)            # one utterance: the time axis is split 16 ways
                    fpart = torch.empty(1, ts.value, beam, V, device="cuda")
                    ftick = torch.zeros(1, ncg.value, dtype=torch.int32, device="cuda")
                    for _ in range(2):                                             # twice: the tickets must re-arm themselves
                        L.check(lib.avsr_ctc_prefix_full(L.ptr(logp), V, ldp, 0, V - 1, L.ptr(utt_off), L.ptr(utt_T), L.ptr(n_run), beam, 1, 1,
                                                         L.ptr(last_tok), L.ptr(rprev), L.ptr(r_buf), T, L.ptr(step_t), L.ptr(s_prev),
                                                         L.ptr(scores), L.ptr(fpart), L.ptr(ftick), L.stream()), "full")
                    assert int(ftick.abs().sum().item()) == 0
                    got = scores[:n_hyp].cpu().numpy()
                    live = ref > -1e9
                    assert np.abs(got[live] - ref[live]).max() < 1e-3, (tag, mode, step)
                    assert (got[~live] < -1e9).all()
                    # survivor chains are recomputed by the pre-beam kernel on the picked tokens (S = 1 candidate per new hyp)
                    part = i32([[int(t)] for _, t in picks])
                    rows_last = i32([last[int(h)] for h, _ in picks])
                    rp = i32([int(rprev[int(h)].item()) for h, _ in picks])
                    n_new = i32([len(picks)])
                    psi2 = torch.zeros(R, 1, device="cuda")
                    L.check(lib.avsr_ctc_prefix_prebeam(L.ptr(logp), V, ldp, 0, L.ptr(utt_off), L.ptr(utt_T), L.ptr(n_new), beam, R, 1,
                                                        L.ptr(rows_last), L.ptr(part), L.ptr(rp), L.ptr(r_buf), T, L.ptr(step_t),
                                                        L.ptr(psi2), L.ptr(rsum), L.stream()), "prebeam(recompute)")
                    sp_old = s_prev.cpu().numpy()
                    for j, (h, t) in enumerate(picks):
                        assert abs((psi2[j, 0].item() - sp_old[h]) - ref[h, t]) < 1e-3
                    rprev = i32(list(range(len(picks))) + [0] * (R - len(picks)))
                    s_prev = torch.cat([psi2[:len(picks), 0], torch.zeros(R - len(picks), device="cuda")])
                n_hyp = len(picks)


@pytest.mark.gpu
def test_sharded_evaluation_driver_equals_single_runs(gpu_model):
    """evaluate_sharded (length-bucketed batches, the eval_lrs2 loop of script/evaluation.py:387-404) on a mixed-length
    set: every utterance must get the token ids of its own `inference()` call, and the corpus WER must follow from them."""
    from avsr_b200 import evaluation as E
    from avsr_b200 import sharding as S
    lengths = [14, 9, 22, 5, 17]
    samples = [synth.make_inputs(900 + i, t) for i, t in enumerate(lengths)]
    load = lambda i: (samples[i][0][0], samples[i][1][0])
    want = []
    for v, a in samples:
        ids = gpu_model.inference(v.cuda(), a.cuda())
        want.append(E.strip_sos_eos([gpu_model.sos] + ids, gpu_model.eos))
    refs = [" ".join(str(t) for t in w) for w in want]
    refs[2] = refs[2] + " 7"                               # one deletion in utterance 2
    res = E.evaluate_sharded(gpu_model, lengths, load, references=refs, max_utts=2, max_frames=40, device="cuda")
    assert res.n_batches >= 3
    assert [res.hyp_tokens[i] for i in range(len(lengths))] == want
    assert res.edits == 1 and res.ref_words == sum(len(r.split()) for r in refs)
    assert res.wer == pytest.approx(1 / res.ref_words)
    assert S.shard_utterances(lengths, 1) == [[2, 4, 0, 1, 3]]


def test_large_decode_batches_equal_small_ones(gpu_model):
    """evaluate_sharded's default plan packs 100+ short utterances into one decode batch (450 hypothesis rows per position
    here): the token ids must not depend on how the set was cut into batches."""
    from avsr_b200 import evaluation as E
    rng = np.random.default_rng(77)
    lengths = rng.integers(6, 21, size=150).tolist()
    samples = {}

    def load(i):
        if i not in samples:
            v, a = synth.make_inputs(5000 + i, lengths[i])
            samples[i] = (v[0], a[0])
        return samples[i]

    big = E.evaluate_sharded(gpu_model, lengths, load, device="cuda")
    small = E.evaluate_sharded(gpu_model, lengths, load, max_utts=16, device="cuda")
    assert big.n_batches <= 2 and small.n_batches >= 10
    assert len(big.hyp_tokens) == 150
    diff = [i for i in range(150) if big.hyp_tokens[i] != small.hyp_tokens[i]]
    # The packed offset of an utterance shifts the key-block boundaries of the encoder attention (last-bit differences in the
    # bf16 encoder output), and batches above 128 hypothesis rows use the split-K projections: a 1-best may only differ where
    # the search itself has a TIE - the two best hypotheses of that utterance, decoded alone, score within 1e-4 of each other
    # and are exactly the two answers (seen: utterance 62, both at -73.969513).
    assert len(diff) <= 2, diff
    for i in diff:
        v, a = load(i)
        x = gpu_model.encoder(input_features=a[None].cuda(), video=v[None].cuda()).last_hidden_state[0]
        nbest = gpu_model.beam_search(x)
        assert len(nbest) >= 2 and abs(float(nbest[0].score) - float(nbest[1].score)) < 1e-4, (i, [float(h.score) for h in nbest[:3]])
        top2 = {tuple(h.yseq.tolist()[1:-1]) for h in nbest[:2]}                 # hyp_tokens carry no sos / eos
        assert {tuple(big.hyp_tokens[i]), tuple(small.hyp_tokens[i])} == top2, (i, big.hyp_tokens[i], small.hyp_tokens[i], top2)


def test_pinned_host_inputs_are_uploaded_in_chunks(gpu_model):
    """infer_batch / encoder with pinned HOST video (chunked upload on a copy stream under the video frontend) gives the
    bit-identical encoder output of the same call with device-resident inputs."""
    enc = gpu_model.encoder
    old = enc.CHUNK_FRAMES
    enc.CHUNK_FRAMES = 8                         # several chunks + halo frames across chunk borders
    try:
        vids, auds = zip(*[synth.make_inputs(40 + i, 13) for i in range(3)])
        video, audio = torch.cat(vids, 0), torch.cat(auds, 0)
        ref = enc(input_features=audio.cuda(), video=video.cuda()).last_hidden_state
        for _ in range(2):                       # twice: the upload buffer is reused
            got = enc(input_features=audio.pin_memory(), video=video.pin_memory()).last_hidden_state
            assert torch.equal(got, ref)
    finally:
        enc.CHUNK_FRAMES = old


def test_full_size_batch_properties(gpu_model):
    """BASELINE.json configs[1] at full size (32 utterances x 15 s, beam 3) through size-independent properties: every
    utterance of the batch equals its own B=1 run (checked on three of them), hypotheses are complete and sorted, the fused
    score is the weighted sum of its scorer parts, and a second run is bit-identical."""
    B, T = 32, 375
    vids, auds = zip(*[synth.make_inputs(1234 + i, T) for i in range(B)])
    video, audio = torch.cat(vids, 0).cuda(), torch.cat(auds, 0).cuda()
    nb = gpu_model.infer_batch(video, audio)
    assert len(nb) == B
    for hyps in nb:
        assert 1 <= len(hyps) <= 3
        scores = [float(h.score) for h in hyps]
        assert scores == sorted(scores, reverse=True) and all(np.isfinite(scores))
        for h in hyps:
            y = h.yseq.tolist()
            assert len(y) == T + 2 and y[0] == 5048 and y[-1] == 5048           # random init: maxlen = T positions + sos + eos
            assert all(0 < t < 5049 for t in y[1:-1])                            # never the CTC blank
            fused = 0.9 * float(h.scores["decoder"]) + 0.1 * float(h.scores["ctc"])
            assert abs(fused - float(h.score)) < 2e-3 * len(y)
    again = gpu_model.infer_batch(video, audio)
    for a, b in zip(nb, again):
        assert [h.yseq.tolist() for h in a] == [h.yseq.tolist() for h in b]
        assert [float(h.score) for h in a] == [float(h.score) for h in b]
    for i in (0, 13, 31):
        single = gpu_model.infer_batch(video[i:i + 1], audio[i:i + 1])[0]
        assert [h.yseq.tolist() for h in single] == [h.yseq.tolist() for h in nb[i]], i
        # the decode is bit-identical for a given encoder output; the bf16 encoder itself is only tolerance-identical between a
        # batch and a B=1 run (max-abs ~1e-2, see test_batched_encoder_equals_single_runs), which moves the scores slightly
        assert all(abs(float(x.score) - float(y.score)) < 0.5 for x, y in zip(single, nb[i]))
    x = gpu_model.encoder(input_features=audio, video=video).packed
    full = gpu_model.beam_search.decode_batch(x, [T] * B)
    for i in (5, 20):
        one = gpu_model.beam_search.decode_batch(x[i * T:(i + 1) * T].contiguous(), [T])[0]
        assert [h.yseq.tolist() for h in one] == [h.yseq.tolist() for h in full[i]]
        assert [float(h.score) for h in one] == [float(h.score) for h in full[i]]          # bit-identical


@pytest.mark.parametrize("T,beam,graph", [(12, 3, False), (12, 5, True), (30, 3, True)])
def test_ctc_only_beam_search_vs_reference_golden(gpu_model, golden, T, beam, graph):
    """ctc_weight = 1.0: the decoder scorer is dropped and every position scores the full vocabulary with the
    HBM-bound CTC kernel; n-best token-identical to the unmodified reference (tests/golden/ctc_only.npz)."""
    import os
    from avsr_b200.beam_search import BatchedBeamSearch
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ctc_only.npz"))
    bs = BatchedBeamSearch(gpu_model.decoder_weights, beam_size=beam, ctc_weight=1.0, use_graph=graph)
    x = torch.from_numpy(golden[f"enc_T{T}"]).cuda()
    nbest = bs(x)
    ys, sc, ln = g[f"nbest_T{T}_b{beam}_yseq"], g[f"nbest_T{T}_b{beam}_score"], g[f"nbest_T{T}_b{beam}_len"]
    assert len(nbest) == len(sc)
    for k, h in enumerate(nbest):
        assert h.yseq.tolist() == ys[k, :ln[k]].tolist(), (T, beam, k)
        assert abs(float(h.score) - sc[k]) < 1e-3 * ln[k]
        assert set(h.scores) == {"ctc"} and abs(float(h.scores["ctc"]) - g[f"nbest_T{T}_b{beam}_ctc"][k]) < 1e-2 * ln[k]
    # batched: two utterances of different lengths evolve as their own runs
    x12, x30 = torch.from_numpy(golden["enc_T12"]).cuda(), torch.from_numpy(golden["enc_T30"]).cuda()
    if T == 12 and beam == 3:
        out = bs.decode_batch(torch.cat([x30, x12], 0), [30, 12])
        assert [h.yseq.tolist() for h in out[1]] == [h.yseq.tolist() for h in nbest]
        assert out[0][0].yseq.tolist() == g["nbest_T30_b3_yseq"][0, :g["nbest_T30_b3_len"][0]].tolist()


@pytest.mark.parametrize("T,beam", [(1, 3), (2, 3), (5, 1), (9, 8), (3, 5)])
def test_edge_shapes_match_the_oracle(state_dict, gpu_model, T, beam):
    """Shortest utterances and the extreme beam sizes (1 and 8) against the CPU oracle, from the oracle's fp32 encoder output."""
    from avsr_b200.beam_search import BatchedBeamSearch
    from oracle import avsr_oracle as O
    video, audio = synth.make_inputs(700 + T, T)
    x = O.encoder_forward(state_dict, audio, video)[0]
    ref = O.beam_search(state_dict, x, beam, kv_cache=True)
    bs = BatchedBeamSearch(gpu_model.decoder_weights, beam_size=beam)
    nbest = bs(x.cuda())
    assert len(nbest) == len(ref) >= 1
    for a, b in zip(nbest, ref):
        if b.score > -1e8:
            assert a.yseq.tolist() == b.yseq, (T, beam)
            assert abs(float(a.score) - b.score) < 1e-3 * len(b.yseq)


def test_long_utterances_batch_equals_single(gpu_model):
    """30 s utterances (T = 750, BASELINE.json configs[4] upper end) next to a very short one: batch == single runs, complete
    hypotheses."""
    lengths = [750, 3, 401]
    g = torch.Generator().manual_seed(5)
    xs = [torch.nn.functional.layer_norm(torch.randn(t, 1024, generator=g), (1024,)).cuda() for t in lengths]
    bs = gpu_model.beam_search
    out = bs.decode_batch(torch.cat(xs, 0).contiguous(), lengths)
    for x, t, hyps in zip(xs, lengths, out):
        single = bs.decode_batch(x.contiguous(), [t])[0]
        assert [h.yseq.tolist() for h in hyps] == [h.yseq.tolist() for h in single]
        assert [float(h.score) for h in hyps] == [float(h.score) for h in single]
        assert all(len(h.yseq) == t + 2 for h in hyps)
