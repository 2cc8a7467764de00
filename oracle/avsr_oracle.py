"""CPU oracle for the avsr_cocktail inference hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, with plain torch-CPU fp32 ops, the algorithm of the reference's PyTorch path
(quanpn90/avsr): the AV-HuBERT-large encoder forward, the CTC head, the 6-layer transformer decoder
scorer, ``CTCPrefixScoreTH`` and the joint CTC/attention ``BatchBeamSearch``.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
it; the product (``avsr_b200/``) never does.

Parity pin: the reference holds no golden vectors or tests for this path, so the
oracle is pinned against outputs of the reference itself, run in the build container by
``oracle/gen_golden.py`` (which imports /root/reference) and committed under ``tests/golden/``.
``tests/test_oracle_golden.py`` re-checks the oracle against those vectors on every run.

Every function cites the reference file:line it follows (paths relative to /root/reference).
All functions take the reference ``E2E.state_dict()`` (keys ``encoder.* decoder.* ctc.*``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

LOGZERO = -10000000000.0  # src/nets/ctc_prefix_score.py:33
SD = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------------------------
# Encoder
# --------------------------------------------------------------------------------------------
def _bn(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"],
                        sd[p + ".bias"], training=False, eps=1e-5)


def frontend3d(sd: SD, video: torch.Tensor) -> torch.Tensor:
    """[B,1,T,88,88] -> [B*T,64,22,22].  src/nets/backend/backbones/resnet.py:132-136,151-164."""
    p = "encoder.feature_extractor_video.resnet.frontend3D."
    x = F.conv3d(video, sd[p + "0.weight"], None, stride=(1, 2, 2), padding=(2, 3, 3))
    x = _bn(sd, p + "1", x)
    x = F.prelu(x, sd[p + "2.weight"])
    x = F.max_pool3d(x, kernel_size=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1))
    B, C, T, H, W = x.shape
    return x.transpose(1, 2).reshape(B * T, C, H, W)


def basic_block(sd: SD, p: str, x: torch.Tensor, stride: int) -> torch.Tensor:
    """src/nets/backend/backbones/resnet.py:56-69 (PReLU variant)."""
    out = F.conv2d(x, sd[p + "conv1.weight"], None, stride=stride, padding=1)
    out = F.prelu(_bn(sd, p + "bn1", out), sd[p + "relu1.weight"])
    out = _bn(sd, p + "bn2", F.conv2d(out, sd[p + "conv2.weight"], None, stride=1, padding=1))
    if (p + "downsample.0.weight") in sd:
        res = _bn(sd, p + "downsample.1", F.conv2d(x, sd[p + "downsample.0.weight"], None, stride=stride))
    else:
        res = x
    return F.prelu(out + res, sd[p + "relu2.weight"])


def resnet_trunk(sd: SD, x: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
    """[N,64,22,22] -> [N,512].  resnet.py:72-124."""
    r = "encoder.feature_extractor_video.resnet.trunk."
    for li in (1, 2, 3, 4):
        for bi in (0, 1):
            x = basic_block(sd, f"{r}layer{li}.{bi}.", x, 2 if (li > 1 and bi == 0) else 1)
        if taps is not None:
            taps[f"layer{li}"] = x
    return x.mean(dim=(2, 3))


def pos_conv_weight(sd: SD) -> torch.Tensor:
    """Effective weight of the weight-normed (dim=2) positional conv: W = g * v / ||v||_(0,1).
    HF Wav2Vec2PositionalConvEmbedding (transformers/models/wav2vec2/modeling_wav2vec2.py:326-368),
    called at src/nets/backend/backbones/avhubert.py:698."""
    p = "encoder.encoder.pos_conv_embed.conv.parametrizations.weight."
    g, v = sd[p + "original0"], sd[p + "original1"]
    return v * (g / v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt())


def encoder_layer(sd: SD, p: str, h: torch.Tensor, heads: int = 16) -> torch.Tensor:
    """Pre-LN layer.  avhubert.py:747-768 + HF Wav2Vec2Attention/FeedForward (eager attention)."""
    B, T, Dm = h.shape
    a = F.layer_norm(h, (Dm,), sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"], 1e-5)
    q = F.linear(a, sd[p + "attention.q_proj.weight"], sd[p + "attention.q_proj.bias"])
    k = F.linear(a, sd[p + "attention.k_proj.weight"], sd[p + "attention.k_proj.bias"])
    v = F.linear(a, sd[p + "attention.v_proj.weight"], sd[p + "attention.v_proj.bias"])
    dh = Dm // heads
    q = q.view(B, T, heads, dh).transpose(1, 2)
    k = k.view(B, T, heads, dh).transpose(1, 2)
    v = v.view(B, T, heads, dh).transpose(1, 2)
    w = torch.softmax(torch.matmul(q, k.transpose(2, 3)) * (dh ** -0.5), dim=-1)
    o = torch.matmul(w, v).transpose(1, 2).reshape(B, T, Dm)
    h = h + F.linear(o, sd[p + "attention.out_proj.weight"], sd[p + "attention.out_proj.bias"])
    a = F.layer_norm(h, (Dm,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], 1e-5)
    f = F.gelu(F.linear(a, sd[p + "feed_forward.intermediate_dense.weight"],
                        sd[p + "feed_forward.intermediate_dense.bias"]))
    return h + F.linear(f, sd[p + "feed_forward.output_dense.weight"], sd[p + "feed_forward.output_dense.bias"])


def encoder_forward(sd: SD, audio: torch.Tensor, video: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
    """AVHubertModel.forward inference branch (mask=False, features_only=True, padding_mask=None).
    avhubert.py:546-561 -> forward_gen :448-524 -> AVHubertEncoder.forward :672-745.
    audio [B,104,T], video [B,1,T,88,88] -> [B,T,1024]."""
    e = "encoder."
    B, _, T = audio.shape
    fa = F.linear(audio.transpose(1, 2), sd[e + "feature_extractor_audio.proj.weight"],
                  sd[e + "feature_extractor_audio.proj.bias"])                      # avhubert.py:193-198
    f3 = frontend3d(sd, video)
    if taps is not None:
        taps["frontend3d"] = f3
    fv = resnet_trunk(sd, f3, taps).view(B, T, 512)
    if taps is not None:
        taps["trunk"] = fv
    fv = F.linear(fv, sd[e + "feature_extractor_video.proj.weight"], sd[e + "feature_extractor_video.proj.bias"])
    feats = torch.cat([fa, fv], dim=-1)                                             # audio first, :486-487
    feats = F.layer_norm(feats, (feats.shape[-1],), sd[e + "layer_norm.weight"], sd[e + "layer_norm.bias"], 1e-5)
    h = F.linear(feats, sd[e + "post_extract_proj.weight"], sd[e + "post_extract_proj.bias"])
    if taps is not None:
        taps["fused"] = h
    pc = F.conv1d(h.transpose(1, 2), pos_conv_weight(sd), sd[e + "encoder.pos_conv_embed.conv.bias"],
                  padding=64, groups=16)[:, :, :-1]                                 # SamePad drops the last frame
    h = h + F.gelu(pc).transpose(1, 2)                                              # no LN here (:700 commented)
    if taps is not None:
        taps["posconv"] = h
    n_layers = 0
    while f"{e}encoder.layers.{n_layers}.layer_norm.weight" in sd:
        n_layers += 1
    for l in range(n_layers):
        h = encoder_layer(sd, f"{e}encoder.layers.{l}.", h)
        if taps is not None:
            taps[f"enc_layer{l}"] = h
    return F.layer_norm(h, (h.shape[-1],), sd[e + "encoder.layer_norm.weight"], sd[e + "encoder.layer_norm.bias"], 1e-5)


def ctc_log_softmax(sd: SD, x: torch.Tensor) -> torch.Tensor:
    """src/nets/backend/ctc.py:163-170."""
    return torch.log_softmax(F.linear(x, sd["ctc.ctc_lo.weight"], sd["ctc.ctc_lo.bias"]), dim=-1)


# --------------------------------------------------------------------------------------------
# Decoder scorer
# --------------------------------------------------------------------------------------------
_PE_CACHE: Dict[int, torch.Tensor] = {}


def positional_table(n: int, d: int = 1024) -> torch.Tensor:
    """fp32-built sinusoid table.  src/nets/backend/transformer/embedding.py:62-76."""
    if n not in _PE_CACHE:
        pos = torch.arange(0, n, dtype=torch.float32).unsqueeze(1)
        div = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * -(math.log(10000.0) / d))
        pe = torch.zeros(n, d)
        pe[:, 0::2] = torch.sin(pos * div)
        pe[:, 1::2] = torch.cos(pos * div)
        _PE_CACHE[n] = pe
    return _PE_CACHE[n]


def _mha(sd: SD, p: str, q_in, kv_in, heads=16, mask=None):
    """src/nets/backend/transformer/attention.py:38-106."""
    n, tq, Dm = q_in.shape
    dh = Dm // heads
    q = F.linear(q_in, sd[p + "linear_q.weight"], sd[p + "linear_q.bias"]).view(n, -1, heads, dh).transpose(1, 2)
    k = F.linear(kv_in, sd[p + "linear_k.weight"], sd[p + "linear_k.bias"]).view(n, -1, heads, dh).transpose(1, 2)
    v = F.linear(kv_in, sd[p + "linear_v.weight"], sd[p + "linear_v.bias"]).view(n, -1, heads, dh).transpose(1, 2)
    s = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(dh)
    if mask is not None:
        s = s.masked_fill(~mask, torch.finfo(s.dtype).min)
        w = torch.softmax(s, dim=-1).masked_fill(~mask, 0.0)
    else:
        w = torch.softmax(s, dim=-1)
    o = torch.matmul(w, v).transpose(1, 2).reshape(n, tq, Dm)
    return F.linear(o, sd[p + "linear_out.weight"], sd[p + "linear_out.bias"])


def _dec_ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-12)   # layer_norm.py:19


def decoder_batch_score(sd: SD, ys: torch.Tensor, caches: Optional[List[torch.Tensor]], memory: torch.Tensor
                        ) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """Faithful restatement of ``Decoder.batch_score``/``forward_one_step`` with the ESPnet output cache:
    every call re-projects K/V of all L prefix tokens and of all T memory frames for every hyp
    (src/nets/backend/transformer/decoder.py:153-227, decoder_layer.py:58-121).
    ys [n,L] int64; caches: per layer [n,L-1,D] or None; memory [n,T,D] -> (logp [n,V], new caches)."""
    n, L = ys.shape
    Dm = memory.shape[-1]
    x = sd["decoder.embed.0.weight"][ys] * math.sqrt(Dm) + positional_table(5000, Dm)[:L]   # embedding.py:86
    new_caches = []
    n_layers = 0
    while f"decoder.decoders.{n_layers}.norm1.weight" in sd:
        n_layers += 1
    causal = torch.ones(L, L, dtype=torch.bool).tril()
    for l in range(n_layers):
        p = f"decoder.decoders.{l}."
        t = _dec_ln(sd, p + "norm1", x)
        if caches is None:
            q, res, m = t, x, causal.view(1, 1, L, L)
        else:
            q, res, m = t[:, -1:], x[:, -1:], causal[-1:].view(1, 1, 1, L)
        h = res + _mha(sd, p + "self_attn.", q, t, mask=m)
        h = h + _mha(sd, p + "src_attn.", _dec_ln(sd, p + "norm2", h), memory)
        c = _dec_ln(sd, p + "norm3", h)
        h = h + F.linear(torch.relu(F.linear(c, sd[p + "feed_forward.w_1.weight"], sd[p + "feed_forward.w_1.bias"])),
                         sd[p + "feed_forward.w_2.weight"], sd[p + "feed_forward.w_2.bias"])
        if caches is not None:
            h = torch.cat([caches[l], h], dim=1)
        new_caches.append(h)
        x = h
    y = _dec_ln(sd, "decoder.after_norm", x[:, -1])
    logp = torch.log_softmax(F.linear(y, sd["decoder.output_layer.weight"], sd["decoder.output_layer.bias"]), dim=-1)
    return logp, new_caches


class KVDecoder:
    """Mathematically equal KV-cache form of the decoder step, used to keep the oracle
    fast at large T.  Cross-attention K/V are projected once per utterance; self-attention K/V are cached
    per hyp.  Checked against ``decoder_batch_score`` and the reference's ``Decoder.batch_score`` output in
    tests/test_oracle_golden.py::test_decoder_step_matches_reference."""

    def __init__(self, sd: SD, memory: torch.Tensor, heads: int = 16):
        self.sd, self.heads = sd, heads
        self.Dm = memory.shape[-1]
        self.n_layers = 0
        while f"decoder.decoders.{self.n_layers}.norm1.weight" in sd:
            self.n_layers += 1
        dh = self.Dm // heads
        self.ck, self.cv = [], []
        for l in range(self.n_layers):
            p = f"decoder.decoders.{l}.src_attn."
            k = F.linear(memory, sd[p + "linear_k.weight"], sd[p + "linear_k.bias"])
            v = F.linear(memory, sd[p + "linear_v.weight"], sd[p + "linear_v.bias"])
            self.ck.append(k.view(-1, heads, dh).transpose(0, 1))       # [H,T,dh]
            self.cv.append(v.view(-1, heads, dh).transpose(0, 1))

    def step(self, tokens: torch.Tensor, pos: int, kv: Optional[list]):
        """tokens [n] (last token of each hyp), kv: per layer (K [n,H,pos,dh], V) or None."""
        sd, H = self.sd, self.heads
        n = tokens.shape[0]
        dh = self.Dm // H
        h = sd["decoder.embed.0.weight"][tokens] * math.sqrt(self.Dm) + positional_table(5000, self.Dm)[pos]
        new_kv = []
        for l in range(self.n_layers):
            p = f"decoder.decoders.{l}."
            a = _dec_ln(sd, p + "norm1", h)
            q = F.linear(a, sd[p + "self_attn.linear_q.weight"], sd[p + "self_attn.linear_q.bias"]).view(n, H, 1, dh)
            k = F.linear(a, sd[p + "self_attn.linear_k.weight"], sd[p + "self_attn.linear_k.bias"]).view(n, H, 1, dh)
            v = F.linear(a, sd[p + "self_attn.linear_v.weight"], sd[p + "self_attn.linear_v.bias"]).view(n, H, 1, dh)
            if kv is not None:
                k = torch.cat([kv[l][0], k], dim=2)
                v = torch.cat([kv[l][1], v], dim=2)
            new_kv.append((k, v))
            w = torch.softmax(torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(dh), dim=-1)
            o = torch.matmul(w, v).reshape(n, self.Dm)
            h = h + F.linear(o, sd[p + "self_attn.linear_out.weight"], sd[p + "self_attn.linear_out.bias"])
            b = _dec_ln(sd, p + "norm2", h)
            q = F.linear(b, sd[p + "src_attn.linear_q.weight"], sd[p + "src_attn.linear_q.bias"]).view(n, H, 1, dh)
            w = torch.softmax(torch.matmul(q, self.ck[l].transpose(-2, -1)) / math.sqrt(dh), dim=-1)
            o = torch.matmul(w, self.cv[l]).reshape(n, self.Dm)
            h = h + F.linear(o, sd[p + "src_attn.linear_out.weight"], sd[p + "src_attn.linear_out.bias"])
            c = _dec_ln(sd, p + "norm3", h)
            h = h + F.linear(torch.relu(F.linear(c, sd[p + "feed_forward.w_1.weight"], sd[p + "feed_forward.w_1.bias"])),
                             sd[p + "feed_forward.w_2.weight"], sd[p + "feed_forward.w_2.bias"])
        y = _dec_ln(sd, "decoder.after_norm", h)
        logp = torch.log_softmax(F.linear(y, sd["decoder.output_layer.weight"], sd["decoder.output_layer.bias"]), dim=-1)
        return logp, new_kv


# --------------------------------------------------------------------------------------------
# CTC prefix scoring
# --------------------------------------------------------------------------------------------
def ctc_initial_state(logp: torch.Tensor, blank: int = 0):
    """First-call state of CTCPrefixScoreTH: r^n = logzero, r^b = cumsum_t logp[t, blank], s_prev = 0.
    src/nets/ctc_prefix_score.py:83-93.  Returns (rn [T], rb [T], s_prev)."""
    T = logp.shape[0]
    return torch.full((T,), LOGZERO), torch.cumsum(logp[:, blank], 0), 0.0


def ctc_prefix_scores(logp: torch.Tensor, rn_prev: torch.Tensor, rb_prev: torch.Tensor, s_prev,
                      last: Sequence[int], out_len: int, cand: Optional[torch.Tensor],
                      blank: int = 0, eos: Optional[int] = None):
    """One ``CTCPrefixScoreTH.__call__`` (src/nets/ctc_prefix_score.py:68-187) for the n hyps of one utterance.

    logp [T,V]; rn_prev, rb_prev [T,n]; s_prev [n] (or 0.0); last [n] last token of each hyp;
    out_len = len(y)-1; cand [n,S] candidate ids (pre-beam) or None (full vocabulary).
    Returns (scores [n,V] = log_psi - s_prev, log_psi [n,V], rn [T,n,S'], rb [T,n,S']) where S' = S or V.
    """
    T, Vv = logp.shape
    n = rn_prev.shape[1]
    eos = Vv - 1 if eos is None else eos
    if cand is None:
        x = logp.unsqueeze(1).expand(T, n, Vv)                              # :115-119
    else:
        x = logp[:, cand.reshape(-1)].view(T, n, -1)                        # :99-114
    S = x.shape[2]
    xb = logp[:, blank].view(T, 1, 1)
    r_sum = torch.logsumexp(torch.stack([rn_prev, rb_prev]), 0)             # :132  [T,n]
    phi = r_sum.unsqueeze(2).repeat(1, 1, S)                                # :133
    for i in range(n):                                                      # :134-141
        if cand is None:
            phi[:, i, last[i]] = rb_prev[:, i]
        else:
            hit = (cand[i] == last[i]).nonzero()
            if hit.numel():
                phi[:, i, int(hit[-1])] = rb_prev[:, i]
    rn = torch.full((T, n, S), LOGZERO)
    rb = torch.full((T, n, S), LOGZERO)
    if out_len == 0:
        rn[0] = x[0]                                                        # :129-130
    start = max(out_len, 1)                                                 # :150-153
    for t in range(start, T):                                               # :156-161
        rn[t] = torch.logsumexp(torch.stack([rn[t - 1], phi[t - 1]]), 0) + x[t]
        rb[t] = torch.logsumexp(torch.stack([rn[t - 1], rb[t - 1]]), 0) + xb[t]
    phi_x = torch.cat([phi[:1], phi[:-1]], 0) + x                           # :164
    psi_c = torch.logsumexp(torch.cat([phi_x[start:T], rn[start - 1].unsqueeze(0)], 0), 0)   # :169-179  [n,S]
    if cand is None:
        log_psi = psi_c.clone()
    else:
        log_psi = torch.full((n, Vv), LOGZERO)
        for i in range(n):
            log_psi[i, cand[i]] = psi_c[i]
    log_psi[:, eos] = r_sum[T - 1]                                          # :181-182
    log_psi[:, blank] = LOGZERO                                             # :185
    sp = s_prev if isinstance(s_prev, float) else s_prev.view(n, 1)
    return log_psi - sp, log_psi, rn, rb


# --------------------------------------------------------------------------------------------
# Joint CTC/attention beam search
# --------------------------------------------------------------------------------------------
@dataclass
class Hyp:
    yseq: List[int]
    score: float
    dec_score: float = 0.0
    ctc_score: float = 0.0
    # tensors kept in fp32 exactly like the reference keeps 0-dim tensors
    _score_t: torch.Tensor = field(default=None, repr=False)


def end_detect(ended: List[Hyp], i: int, M: int = 3, d_end: float = math.log(math.exp(-10))) -> bool:
    """src/nets/e2e_asr_common.py:18-48."""
    if not ended:
        return False
    best = max(h.score for h in ended)
    count = 0
    for m in range(M):
        same = [h.score for h in ended if len(h.yseq) == i - m]
        if same and max(same) - best < d_end:
            count += 1
    return count == M


def beam_search(sd: SD, x: torch.Tensor, beam_size: int = 3, ctc_weight: float = 0.1,
                pre_beam_ratio: float = 1.5, maxlenratio: float = 0.0, kv_cache: bool = True,
                max_steps: Optional[int] = None, trace: Optional[list] = None) -> List[Hyp]:
    """``BatchBeamSearch.forward`` as wired by ``get_beam_search_decoder``
    (src/avhubert_avsr/avhubert_avsr_model.py:12-36; src/nets/beam_search.py:330-406;
    src/nets/batch_beam_search.py:86-110,208-349).  x [T,D] -> ended hyps sorted by score (desc).

    kv_cache=False runs the decoder exactly in the reference's compute pattern (the CPU baseline);
    kv_cache=True uses the equal KV-cache form.  ``max_steps`` truncates the loop (bench sampling only).
    """
    T, Dm = x.shape
    Vv = sd["decoder.output_layer.weight"].shape[0]
    sos = eos = Vv - 1
    blank = 0
    w_dec = torch.tensor(1.0 - ctc_weight, dtype=torch.float32)
    w_ctc = torch.tensor(ctc_weight, dtype=torch.float32)
    S = int(pre_beam_ratio * beam_size)                                     # beam_search.py:91
    maxlen = T if maxlenratio == 0 else (-int(maxlenratio) if maxlenratio < 0 else max(1, int(maxlenratio * T)))
    logp_ctc = ctc_log_softmax(sd, x.unsqueeze(0))[0]                       # scorers/ctc.py:96
    ctc_only = ctc_weight == 1.0       # decoder dropped (weight 0, beam_search.py:72-75), no pre-beam (avhubert_avsr_model.py:35)
    kvdec = KVDecoder(sd, x) if (kv_cache and not ctc_only) else None

    # running hyps (all the same length)
    yseqs = [[sos]]
    score = torch.zeros(1)
    dsc = torch.zeros(1)
    csc = torch.zeros(1)
    dec_state = None
    rn0, rb0, _ = ctc_initial_state(logp_ctc, blank)
    rn_prev, rb_prev = rn0.unsqueeze(1), rb0.unsqueeze(1)
    s_prev = 0.0
    ended: List[Hyp] = []
    n_steps = maxlen if max_steps is None else min(maxlen, max_steps)
    for i in range(n_steps):
        n = len(yseqs)
        ys = torch.tensor(yseqs, dtype=torch.int64)
        if ctc_only:
            dec, new_state = torch.zeros(n, Vv), None
            weighted = torch.zeros(n, Vv)
            part = None                                                     # batch_beam_search.py:219: no pre-beam
        else:
            if kv_cache:
                dec, new_state = kvdec.step(ys[:, -1], i, dec_state)
            else:
                dec, new_state = decoder_batch_score(sd, ys, dec_state, x.unsqueeze(0).expand(n, T, Dm))
            weighted = torch.zeros(n, Vv) + w_dec * dec                     # batch_beam_search.py:222-227
            part = torch.topk(dec, S, dim=-1)[1]                            # :235
        ctc, log_psi, rn, rb = ctc_prefix_scores(logp_ctc, rn_prev, rb_prev, s_prev,
                                                 [y[-1] for y in yseqs], len(yseqs[0]) - 1, part, blank, eos)
        weighted = weighted + w_ctc * ctc                                   # :240-241
        weighted = weighted + score.unsqueeze(1)                            # :243-245
        flat = weighted.view(-1)
        # torch.topk ties are unspecified; the oracle breaks them by lowest flat index (stable sort)
        top = torch.sort(flat, descending=True, stable=True)[1][:beam_size]  # :104
        prev = torch.div(top, Vv, rounding_mode="trunc")
        tok = top % Vv
        if trace is not None:
            trace.append(dict(dec=dec.clone(), part=None if part is None else part.clone(), ctc=ctc.clone(), weighted=weighted.clone(),
                              prev=prev.clone(), tok=tok.clone()))
        n_yseqs, n_score, n_dsc, n_csc, keep_prev, keep_tok = [], [], [], [], [], []
        for pj, tj in zip(prev.tolist(), tok.tolist()):
            n_yseqs.append(yseqs[pj] + [tj])
            n_score.append(weighted[pj, tj])
            n_dsc.append(dsc[pj] + dec[pj, tj])
            n_csc.append(csc[pj] + ctc[pj, tj])
            keep_prev.append(pj)
            keep_tok.append(tj)
        if i == maxlen - 1:                                                 # :321-337 eos appended, score unchanged
            n_yseqs = [y + [eos] for y in n_yseqs]
        run = []
        for j, y in enumerate(n_yseqs):                                     # :341-349
            if y[-1] == eos:
                ended.append(Hyp(y, float(n_score[j]), float(n_dsc[j]), float(n_csc[j])))
            else:
                run.append(j)
        if maxlenratio == 0.0 and end_detect(ended, i):                     # beam_search.py:369
            break
        if not run:
            break
        # select states of the survivors (scorers/ctc.py:40-63; decoder caches by prev hyp)
        yseqs = [n_yseqs[j] for j in run]
        score = torch.stack([n_score[j] for j in run])
        dsc = torch.stack([n_dsc[j] for j in run])
        csc = torch.stack([n_csc[j] for j in run])
        pidx = torch.tensor([keep_prev[j] for j in run])
        if ctc_only:
            dec_state = None
        elif kv_cache:
            dec_state = [(k[pidx], v[pidx]) for k, v in new_state]
        else:
            dec_state = [c[pidx] for c in new_state]
        cols = []
        for j in run:
            pj, tj = keep_prev[j], keep_tok[j]
            if ctc_only:
                cols.append(tj)                                             # full vocabulary: the state column is the token
                continue
            hit = (part[pj] == tj).nonzero()
            cols.append(int(hit[-1]) if hit.numel() else S - 1)              # idmap -1 -> last column (eos quirk)
        cols_t = torch.tensor(cols)
        rn_prev = rn[:, pidx, cols_t]
        rb_prev = rb[:, pidx, cols_t]
        s_prev = torch.stack([log_psi[keep_prev[j], keep_tok[j]] for j in run])
    ended.sort(key=lambda h: h.score, reverse=True)
    return ended


# --------------------------------------------------------------------------------------------
# Whole path (what AVSRCocktailModel.inference does, script/evaluation.py:96-108)
# --------------------------------------------------------------------------------------------
def infer(sd: SD, video: torch.Tensor, audio: torch.Tensor, beam_size: int = 3, kv_cache: bool = False,
          max_steps: Optional[int] = None) -> List[List[Hyp]]:
    """Sequential B=1 loop like the reference's eval loop (script/evaluation.py:390-400)."""
    out = []
    for b in range(video.shape[0]):
        x = encoder_forward(sd, audio[b:b + 1], video[b:b + 1])[0]
        out.append(beam_search(sd, x, beam_size, kv_cache=kv_cache, max_steps=max_steps))
    return out
