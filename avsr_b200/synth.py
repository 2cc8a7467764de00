"""Deterministic random-init weights and synthetic inputs for the avsr_cocktail hot path.

The reference builds its model with ``AVHubertAVSR(AVHubertAVSRConfig())``
(/root/reference/src/avhubert_avsr/avhubert_avsr_model.py:45-50) and the GPU box has no copy of the
reference, so parity tests and the bench need a weight source that travels: this module produces a
``state_dict`` with exactly the reference's key names and shapes (SURVEY.md App. A;
/root/reference/src/nets/backend/e2e_asr_avhubert.py:24-117) from per-tensor seeded CPU generators.
BatchNorm statistics, PReLU slopes and LayerNorm affines are randomised (the reference's defaults are
the identity) so that every folding step of the weight repacker is exercised.
"""
from __future__ import annotations

import hashlib
import math
from typing import Dict, Tuple

import torch

V = 5049          # src/tokenizer/spm_tokenizer.py:38
D = 1024
ENC_LAYERS = 24
ENC_FFN = 4096
DEC_LAYERS = 6
DEC_FFN = 3072
HEADS = 16
AUDIO_DIM = 104
POS_K = 128
POS_G = 16
SOS = EOS = V - 1  # e2e_asr_avhubert.py:96-98
BLANK = 0


def _gen(name: str, seed: int) -> torch.Generator:
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:7], "little"))
    return g


def _uniform(name, seed, shape, lo, hi):
    return torch.rand(shape, generator=_gen(name, seed), dtype=torch.float32) * (hi - lo) + lo


def _normal(name, seed, shape, std, mean=0.0):
    return torch.randn(shape, generator=_gen(name, seed), dtype=torch.float32) * std + mean


def make_state_dict(seed: int = 0, enc_layers: int = ENC_LAYERS, dec_layers: int = DEC_LAYERS,
                    randomize_norms: bool = True) -> Dict[str, torch.Tensor]:
    """state_dict of the reference ``E2E`` module (keys without the ``avsr.`` prefix)."""
    sd: Dict[str, torch.Tensor] = {}

    def linear(prefix, out_f, in_f):
        b = 1.0 / math.sqrt(in_f)
        sd[prefix + ".weight"] = _uniform(prefix + ".weight", seed, (out_f, in_f), -b, b)
        sd[prefix + ".bias"] = _uniform(prefix + ".bias", seed, (out_f,), -b, b)

    def norm(prefix, n):
        if randomize_norms:
            sd[prefix + ".weight"] = _uniform(prefix + ".weight", seed, (n,), 0.8, 1.2)
            sd[prefix + ".bias"] = _normal(prefix + ".bias", seed, (n,), 0.05)
        else:
            sd[prefix + ".weight"] = torch.ones(n)
            sd[prefix + ".bias"] = torch.zeros(n)

    def bn(prefix, c):
        norm(prefix, c)
        if randomize_norms:
            sd[prefix + ".running_mean"] = _normal(prefix + ".running_mean", seed, (c,), 0.1)
            sd[prefix + ".running_var"] = _uniform(prefix + ".running_var", seed, (c,), 0.5, 1.5)
        else:
            sd[prefix + ".running_mean"] = torch.zeros(c)
            sd[prefix + ".running_var"] = torch.ones(c)
        sd[prefix + ".num_batches_tracked"] = torch.zeros((), dtype=torch.int64)

    def prelu(name, c):
        sd[name] = _uniform(name, seed, (c,), 0.1, 0.4) if randomize_norms else torch.full((c,), 0.25)

    def conv(name, shape):
        # resnet.py:83-86: normal(0, sqrt(2 / (k*k*out)))
        n = shape[0]
        for k in shape[2:]:
            n *= k
        sd[name] = _normal(name, seed, shape, math.sqrt(2.0 / n))

    e = "encoder."
    sd[e + "mask_emb"] = _uniform(e + "mask_emb", seed, (AUDIO_DIM,), 0, 1)
    sd[e + "label_embs_concat"] = _uniform(e + "label_embs_concat", seed, (2004, 256), 0, 1)
    linear(e + "feature_extractor_audio.proj", D, AUDIO_DIM)
    r = e + "feature_extractor_video.resnet."
    sd[r + "frontend3D.0.weight"] = _normal(r + "frontend3D.0.weight", seed, (64, 1, 5, 7, 7), 0.037)
    bn(r + "frontend3D.1", 64)
    prelu(r + "frontend3D.2.weight", 64)
    inpl = 64
    for li, planes in enumerate((64, 128, 256, 512), start=1):
        for bi in range(2):
            p = f"{r}trunk.layer{li}.{bi}."
            cin = inpl if bi == 0 else planes
            conv(p + "conv1.weight", (planes, cin, 3, 3))
            bn(p + "bn1", planes)
            prelu(p + "relu1.weight", planes)
            prelu(p + "relu2.weight", planes)
            conv(p + "conv2.weight", (planes, planes, 3, 3))
            bn(p + "bn2", planes)
            if bi == 0 and li > 1:
                conv(p + "downsample.0.weight", (planes, cin, 1, 1))
                bn(p + "downsample.1", planes)
        inpl = planes
    linear(e + "feature_extractor_video.proj", D, 512)
    linear(e + "post_extract_proj", D, 2 * D)
    pc = e + "encoder.pos_conv_embed.conv."
    sd[pc + "bias"] = _uniform(pc + "bias", seed, (D,), -0.011, 0.011)
    v = _normal(pc + "v", seed, (D, D // POS_G, POS_K), 2.0 * math.sqrt(1.0 / (POS_K * D)))
    sd[pc + "parametrizations.weight.original1"] = v
    g = v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt()       # weight_norm(dim=2) initial g = ||v||
    if randomize_norms:
        g = g * _uniform(pc + "g", seed, (1, 1, POS_K), 0.9, 1.1)
    sd[pc + "parametrizations.weight.original0"] = g
    norm(e + "encoder.layer_norm", D)
    for l in range(enc_layers):
        p = f"{e}encoder.layers.{l}."
        for nm in ("k_proj", "v_proj", "q_proj", "out_proj"):
            linear(p + "attention." + nm, D, D)
        norm(p + "layer_norm", D)
        linear(p + "feed_forward.intermediate_dense", ENC_FFN, D)
        linear(p + "feed_forward.output_dense", D, ENC_FFN)
        norm(p + "final_layer_norm", D)
    norm(e + "layer_norm", 2 * D)

    d = "decoder."
    sd[d + "embed.0.weight"] = _normal(d + "embed.0.weight", seed, (V, D), 1.0)
    for l in range(dec_layers):
        p = f"{d}decoders.{l}."
        for att in ("self_attn", "src_attn"):
            for nm in ("linear_q", "linear_k", "linear_v", "linear_out"):
                linear(f"{p}{att}.{nm}", D, D)
        linear(p + "feed_forward.w_1", DEC_FFN, D)
        linear(p + "feed_forward.w_2", D, DEC_FFN)
        for nm in ("norm1", "norm2", "norm3"):
            norm(p + nm, D)
    norm(d + "after_norm", D)
    linear(d + "output_layer", V, D)
    linear("ctc.ctc_lo", V, D)
    return sd


def make_inputs(seed: int, T: int, B: int = 1) -> Tuple[torch.Tensor, torch.Tensor]:
    """(video [B,1,T,88,88], audio [B,104,T]) fp32 N(0,1), the cfg-1 recipe of SURVEY.md 8(d)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    video = torch.randn(B, 1, T, 88, 88, generator=g)
    audio = torch.randn(B, AUDIO_DIM, T, generator=g)
    return video, audio


def fingerprint(sd: Dict[str, torch.Tensor]) -> float:
    """Cheap checksum over a few tensors, stored beside golden vectors to detect RNG drift."""
    keys = ["ctc.ctc_lo.weight", "decoder.embed.0.weight",
            "encoder.feature_extractor_video.resnet.frontend3D.0.weight",
            "encoder.encoder.layers.0.attention.q_proj.weight"]
    return float(sum(sd[k].double().abs().sum() for k in keys if k in sd))
