"""Repack the reference ``E2E.state_dict()`` into the device layouts the sm_100a kernels consume.

Key names are the reference's (/root/reference/src/nets/backend/e2e_asr_avhubert.py:24-117); an
``avsr.`` prefix (``AVHubertAVSR.state_dict()``, src/avhubert_avsr/avhubert_avsr_model.py:45-50) is accepted.
Folding done here, once, on the host in fp64/fp32:
  * eval-mode BatchNorm into the preceding conv (weight scale + bias)             resnet.py:56-69,132-136
  * weight-norm (dim=2) of the positional conv into a plain weight               modeling_wav2vec2.py:326-368
  * the attention scale 1/sqrt(64) into Wq, bq (exact: a power of two)           modeling_wav2vec2.py:438-463
  * q|k concatenation (encoder), q|k|v (decoder self-attn), all 6 layers' cross k|v (decoder memory projection)
Encoder GEMM operands are bf16 (tcgen05), everything on the decode side stays fp32 (token parity).
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from . import synth


def _strip(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    if any(k.startswith("avsr.") for k in sd):
        return {k[len("avsr."):]: v for k, v in sd.items() if k.startswith("avsr.")}
    return sd


def _fold_bn(sd, conv_key: str, bn_prefix: str, eps: float = 1e-5):
    w = sd[conv_key].double()
    g, b = sd[bn_prefix + ".weight"].double(), sd[bn_prefix + ".bias"].double()
    mean, var = sd[bn_prefix + ".running_mean"].double(), sd[bn_prefix + ".running_var"].double()
    scale = g / torch.sqrt(var + eps)
    w = w * scale.view(-1, *([1] * (w.dim() - 1)))
    bias = b - mean * scale
    return w, bias


class EncoderWeights:
    """bf16 GEMM operands + fp32 vectors of the AV-HuBERT-large encoder."""

    def __init__(self, sd: Dict[str, torch.Tensor], device: torch.device):
        sd = _strip(sd)
        e = "encoder."
        bf = lambda t: t.to(torch.float32).to(device=device, dtype=torch.bfloat16).contiguous()
        f32 = lambda t: t.to(device=device, dtype=torch.float32).contiguous()
        self._check(sd)
        r = e + "feature_extractor_video.resnet."
        w, b = _fold_bn(sd, r + "frontend3D.0.weight", r + "frontend3D.1")
        w = w.reshape(64, 245)
        wpad = torch.zeros(64, 256, dtype=torch.float64)
        wpad[:, :245] = w                                   # k = (dt*7 + dy)*7 + dx, matches im2col_frontend
        self.front_w, self.front_b = bf(wpad), f32(b)
        # implicit-GEMM layout (csrc/frontend_conv.cu): k = (dt*7 + dy)*8 + dx, the eighth dx and k >= 280 are zero
        w8 = torch.zeros(64, 40, 8, dtype=torch.float64)
        w8[:, :35, :7] = w.reshape(64, 35, 7)
        self.front_w8 = bf(w8.reshape(64, 320))
        self.front_prelu = f32(sd[r + "frontend3D.2.weight"])
        self.blocks = []
        for li in (1, 2, 3, 4):
            for bi in (0, 1):
                p = f"{r}trunk.layer{li}.{bi}."
                blk = {}
                for ci, bn in (("conv1", "bn1"), ("conv2", "bn2")):
                    w, b = _fold_bn(sd, p + ci + ".weight", p + bn)
                    blk[ci + "_w"] = bf(w.permute(0, 2, 3, 1).reshape(w.shape[0], -1))   # [Cout, (ky,kx,cin)]
                    blk[ci + "_b"] = f32(b)
                blk["prelu1"], blk["prelu2"] = f32(sd[p + "relu1.weight"]), f32(sd[p + "relu2.weight"])
                blk["stride"] = 2 if (li > 1 and bi == 0) else 1
                blk["cout"] = sd[p + "conv1.weight"].shape[0]
                if (p + "downsample.0.weight") in sd:
                    w, b = _fold_bn(sd, p + "downsample.0.weight", p + "downsample.1")
                    blk["down_w"], blk["down_b"] = bf(w.reshape(w.shape[0], -1)), f32(b)
                self.blocks.append(blk)
        self.vproj_w, self.vproj_b = bf(sd[e + "feature_extractor_video.proj.weight"]), f32(sd[e + "feature_extractor_video.proj.bias"])
        self.aproj_w, self.aproj_b = bf(sd[e + "feature_extractor_audio.proj.weight"]), f32(sd[e + "feature_extractor_audio.proj.bias"])
        self.fuse_ln_g, self.fuse_ln_b = f32(sd[e + "layer_norm.weight"]), f32(sd[e + "layer_norm.bias"])
        self.post_w, self.post_b = bf(sd[e + "post_extract_proj.weight"]), f32(sd[e + "post_extract_proj.bias"])
        pc = e + "encoder.pos_conv_embed.conv."
        g = sd[pc + "parametrizations.weight.original0"].double()
        v = sd[pc + "parametrizations.weight.original1"].double()
        weff = v * (g / v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt())            # [1024 out, 64 in, 128 tap]
        # per group: Wg[o][tap*64 + c]
        self.pos_w = bf(weff.view(16, 64, 64, 128).permute(0, 1, 3, 2).reshape(16, 64, 128 * 64))
        self.pos_b = f32(sd[pc + "bias"])
        self.layers = []
        l = 0
        while f"{e}encoder.layers.{l}.layer_norm.weight" in sd:
            p = f"{e}encoder.layers.{l}."
            a = p + "attention."
            lay = dict(
                ln1_g=f32(sd[p + "layer_norm.weight"]), ln1_b=f32(sd[p + "layer_norm.bias"]),
                wqk=bf(torch.cat([sd[a + "q_proj.weight"] * 0.125, sd[a + "k_proj.weight"]], 0)),
                bqk=f32(torch.cat([sd[a + "q_proj.bias"] * 0.125, sd[a + "k_proj.bias"]], 0)),
                wv=bf(sd[a + "v_proj.weight"]), bv=f32(sd[a + "v_proj.bias"]),
                wo=bf(sd[a + "out_proj.weight"]), bo=f32(sd[a + "out_proj.bias"]),
                ln2_g=f32(sd[p + "final_layer_norm.weight"]), ln2_b=f32(sd[p + "final_layer_norm.bias"]),
                w1=bf(sd[p + "feed_forward.intermediate_dense.weight"]), b1=f32(sd[p + "feed_forward.intermediate_dense.bias"]),
                w2=bf(sd[p + "feed_forward.output_dense.weight"]), b2=f32(sd[p + "feed_forward.output_dense.bias"]),
            )
            self.layers.append(lay)
            l += 1
        self.final_ln_g, self.final_ln_b = f32(sd[e + "encoder.layer_norm.weight"]), f32(sd[e + "encoder.layer_norm.bias"])

    @staticmethod
    def _check(sd):
        """Fail loudly on a checkpoint whose architecture differs from the one the kernels are written for."""
        e = "encoder."
        exp = {
            e + "feature_extractor_audio.proj.weight": (1024, 104),
            e + "feature_extractor_video.proj.weight": (1024, 512),
            e + "post_extract_proj.weight": (1024, 2048),
            e + "encoder.pos_conv_embed.conv.parametrizations.weight.original1": (1024, 64, 128),
            e + "encoder.layers.0.feed_forward.intermediate_dense.weight": (4096, 1024),
            e + "feature_extractor_video.resnet.frontend3D.0.weight": (64, 1, 5, 7, 7),
        }
        for k, shp in exp.items():
            if k not in sd or tuple(sd[k].shape) != shp:
                raise RuntimeError(f"unsupported checkpoint: {k} is {tuple(sd[k].shape) if k in sd else 'missing'}, expected {shp}")


def split3_weight(w: torch.Tensor) -> torch.Tensor:
    """fp32 [N,K] -> bf16 [N,6K] = [w1|w2|w1|w3|w2|w1] with w = w1+w2+w3 (three bf16 terms, ~24 mantissa bits).
    Paired with activations laid out [a1|a1|a2|a1|a2|a3] (csrc/common.cuh avsr_split3_store) one bf16 tensor-core GEMM
    over 6K accumulates the six largest cross terms of (a1+a2+a3)(w1+w2+w3): fp32-level accuracy at tensor-core speed."""
    w = w.float()
    w1 = w.bfloat16()
    r = w - w1.float()
    w2 = r.bfloat16()
    w3 = (r - w2.float()).bfloat16()
    return torch.cat([w1, w2, w1, w3, w2, w1], dim=1).contiguous()


def split3_weight_compact(w: torch.Tensor) -> torch.Tensor:
    """fp32 [N,K] -> bf16 [N,3K] = [w1|w2|w3] (w = w1+w2+w3): operand of the decoder step's own GEMM (csrc/gemm_x3.cu), which
    forms the six cross terms with six MMAs per k step, so a weight is streamed from HBM as 6 bytes instead of 12."""
    w = w.float()
    w1 = w.bfloat16()
    r = w - w1.float()
    w2 = r.bfloat16()
    w3 = (r - w2.float()).bfloat16()
    return torch.cat([w1, w2, w3], dim=1).contiguous()


def fold_layernorm(w: torch.Tensor, bias, gamma: torch.Tensor, beta: torch.Tensor):
    """LayerNorm folded into the nn.Linear that consumes it (csrc/gemm_x3c.cu, avsr_dec_proj_folded):
    LayerNorm(x) W^T + bias = rstd * (x (gamma . W)^T - mean * u) + c with u = W gamma, c = W beta + bias.
    Returns (compact bf16x3 of gamma . W, u fp32 [N], c fp32 [N]); products and sums in float64."""
    wd = w.double().cpu()
    g, b = gamma.double().cpu(), beta.double().cpu()
    wg = wd * g.unsqueeze(0)
    u = wg.sum(1)
    c = wd @ b
    if bias is not None:
        c = c + bias.double().cpu()
    dev = w.device
    return split3_weight_compact(wg.float().to(dev)), u.float().to(dev).contiguous(), c.float().to(dev).contiguous()


class DecoderWeights:
    """Operands of the 6-layer transformer decoder + CTC head: fp32 (CUDA-core path) and bf16x3 (tensor-core path)."""

    def __init__(self, sd: Dict[str, torch.Tensor], device: torch.device, keep_fp32: bool = True):
        sd = _strip(sd)
        f32 = lambda t: t.to(device=device, dtype=torch.float32).contiguous()
        d = "decoder."
        self.V, self.D = sd[d + "output_layer.weight"].shape
        if self.D != 1024 or sd[d + "decoders.0.feed_forward.w_1.weight"].shape[0] != 3072:
            raise RuntimeError("unsupported decoder geometry (expected adim 1024, dunits 3072)")
        self.embed = f32(sd[d + "embed.0.weight"])
        # fp32-built sinusoid table, exactly as src/nets/backend/transformer/embedding.py:62-76 builds it
        n = 5000
        pos = torch.arange(0, n, dtype=torch.float32).unsqueeze(1)
        div = torch.exp(torch.arange(0, self.D, 2, dtype=torch.float32) * -(math.log(10000.0) / self.D))
        pe = torch.zeros(n, self.D)
        pe[:, 0::2] = torch.sin(pos * div)
        pe[:, 1::2] = torch.cos(pos * div)
        self.pe = f32(pe)
        self.layers = []
        ckv_w, ckv_b = [], []
        l = 0
        while f"{d}decoders.{l}.norm1.weight" in sd:
            p = f"{d}decoders.{l}."
            sa, ca = p + "self_attn.", p + "src_attn."
            self.layers.append(dict(
                n1_g=f32(sd[p + "norm1.weight"]), n1_b=f32(sd[p + "norm1.bias"]),
                n2_g=f32(sd[p + "norm2.weight"]), n2_b=f32(sd[p + "norm2.bias"]),
                n3_g=f32(sd[p + "norm3.weight"]), n3_b=f32(sd[p + "norm3.bias"]),
                wqkv=f32(torch.cat([sd[sa + "linear_q.weight"], sd[sa + "linear_k.weight"], sd[sa + "linear_v.weight"]], 0)),
                bqkv=f32(torch.cat([sd[sa + "linear_q.bias"], sd[sa + "linear_k.bias"], sd[sa + "linear_v.bias"]], 0)),
                wo=f32(sd[sa + "linear_out.weight"]), bo=f32(sd[sa + "linear_out.bias"]),
                wq2=f32(sd[ca + "linear_q.weight"]), bq2=f32(sd[ca + "linear_q.bias"]),
                wo2=f32(sd[ca + "linear_out.weight"]), bo2=f32(sd[ca + "linear_out.bias"]),
                w1=f32(sd[p + "feed_forward.w_1.weight"]), b1=f32(sd[p + "feed_forward.w_1.bias"]),
                w2=f32(sd[p + "feed_forward.w_2.weight"]), b2=f32(sd[p + "feed_forward.w_2.bias"]),
            ))
            ckv_w += [sd[ca + "linear_k.weight"], sd[ca + "linear_v.weight"]]
            ckv_b += [sd[ca + "linear_k.bias"], sd[ca + "linear_v.bias"]]
            l += 1
        self.n_layers = l
        self.ckv_w, self.ckv_b = f32(torch.cat(ckv_w, 0)), f32(torch.cat(ckv_b, 0))     # [L*2*1024, 1024]
        self.after_g, self.after_b = f32(sd[d + "after_norm.weight"]), f32(sd[d + "after_norm.bias"])
        self.out_w, self.out_b = f32(sd[d + "output_layer.weight"]), f32(sd[d + "output_layer.bias"])
        self.ctc_w, self.ctc_b = f32(sd["ctc.ctc_lo.weight"]), f32(sd["ctc.ctc_lo.bias"])
        if self.ctc_w.shape[0] != self.V:
            raise RuntimeError("CTC head and decoder vocabulary sizes differ")
        self.sos = self.eos = self.V - 1          # e2e_asr_avhubert.py:96-98
        self.blank = 0
        # bf16x3 operands for the tensor-core decode path
        for lay in self.layers:
            for k in ("wqkv", "wo", "wq2", "wo2", "w1", "w2"):
                lay[k + "3"] = split3_weight_compact(lay[k])            # per-position projections (gemm_x3)
        self.out_w3 = split3_weight_compact(self.out_w)
        # LayerNorm-folded operands of the projections that consume a LayerNorm (norm1 -> q|k|v, norm2 -> src q, norm3 -> w_1,
        # after_norm -> output layer; decoder_layer.py:82-116, decoder.py:176-181)
        for lay in self.layers:
            lay["wqkv3g"], lay["uqkv"], lay["cqkv"] = fold_layernorm(lay["wqkv"], lay["bqkv"], lay["n1_g"], lay["n1_b"])
            lay["wq23g"], lay["uq2"], lay["cq2"] = fold_layernorm(lay["wq2"], lay["bq2"], lay["n2_g"], lay["n2_b"])
            lay["w13g"], lay["u1"], lay["c1"] = fold_layernorm(lay["w1"], lay["b1"], lay["n3_g"], lay["n3_b"])
        self.out_w3g, self.out_u, self.out_c = fold_layernorm(self.out_w, None, self.after_g, self.after_b)     # out_b is added by the softmax kernel
        # "Query merge" (csrc/gemm_x3c.cu avsr_dec_proj_dual, csrc/dec_attn.cu avsr_dec_attn_fold_query): the source-attention
        # query LayerNorm2(x1) Wq^T + bq, x1 = x + att Wo^T + bo, needs x1 (g2 . Wq)^T = x (g2 . Wq)^T + att ((g2 . Wq) Wo)^T +
        # (g2 . Wq) bo: the first product rides along with q | k | v (operand: the raw row x), the second with the attention-output
        # projection (operand: att), and the attention kernel applies the LayerNorm's rstd / mean itself (uq2, cq2 above).
        for lay in self.layers:
            wg = lay["wq2"].double().cpu() * lay["n2_g"].double().cpu().unsqueeze(0)              # g2 . Wq   [1024, 1024]
            wprime = (wg @ lay["wo"].double().cpu()).float().to(device)                             # (g2 . Wq) Wo
            dvec = (wg @ lay["bo"].double().cpu()).float().to(device)                               # (g2 . Wq) bo
            zeros = torch.zeros(1024, dtype=torch.float32, device=device)
            lay["wcat1_3"] = torch.cat([lay["wqkv3g"], lay["wq23g"]], 0).contiguous()               # [3072 + 1024, 3 K]
            lay["ucat1"] = torch.cat([lay["uqkv"], zeros]).contiguous()
            lay["ccat1"] = torch.cat([lay["cqkv"], zeros]).contiguous()
            lay["wcat2_3"] = torch.cat([lay["wo3"], split3_weight_compact(wprime)], 0).contiguous()  # [1024 + 1024, 3 K]
            lay["bcat2"] = torch.cat([lay["bo"], dvec]).contiguous()
        self.ctc_w6 = split3_weight(self.ctc_w)                         # once-per-utterance projections (generic GEMM, K' = 6K)
        self.ckv_w6 = split3_weight(self.ckv_w)
