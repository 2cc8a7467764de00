"""CPU: the host-side weight algebra of the decoder step (avsr_b200/weights.py) against plain float64 formulas.

* three-term bf16 split: w1 + w2 + w3 reproduces an fp32 weight to ~2^-24 relative;
* LayerNorm folded into the consuming nn.Linear (decoder_layer.py:82-116): rstd (x (g . W)^T - mean u) + c;
* "query merge": the source-attention query LayerNorm2(x + att Wo^T + bo) Wq^T + bq (decoder_layer.py:97-107) from the two
  stacked projections x (g2 . Wq)^T and att ((g2 . Wq) Wo)^T + (g2 . Wq) bo and the row statistics, exactly as
  DecoderWeights builds the stacked operands (wcat1_3 / wcat2_3 / bcat2) and as csrc/dec_attn.cu finishes the query.
"""
import torch

from avsr_b200.weights import fold_layernorm, split3_weight_compact


def _unsplit(w3: torch.Tensor) -> torch.Tensor:
    k = w3.shape[1] // 3
    return w3[:, :k].double() + w3[:, k:2 * k].double() + w3[:, 2 * k:].double()


def test_three_term_split_is_fp32_accurate():
    g = torch.Generator().manual_seed(0)
    w = torch.randn(64, 96, generator=g) * torch.logspace(-3, 2, 96)
    err = (_unsplit(split3_weight_compact(w)) - w.double()).abs() / w.double().abs().clamp_min(1e-30)
    assert err.max().item() < 2 ** -22


def test_folded_layernorm_equals_layernorm_then_linear():
    g = torch.Generator().manual_seed(1)
    D, N, R = 256, 48, 7
    x = torch.randn(R, D, generator=g).double() * 2 + 0.5
    w, b = torch.randn(N, D, generator=g) * 0.1, torch.randn(N, generator=g)
    gamma, beta = 1 + 0.3 * torch.randn(D, generator=g), 0.2 * torch.randn(D, generator=g)
    w3g, u, c = fold_layernorm(w, b, gamma, beta)
    mean, var = x.mean(1, keepdim=True), x.var(1, unbiased=False, keepdim=True)
    rstd = 1.0 / torch.sqrt(var + 1e-12)
    got = rstd * (x @ _unsplit(w3g).t() - mean * u.double()) + c.double()
    want = torch.nn.functional.layer_norm(x, (D,), gamma.double(), beta.double(), 1e-12) @ w.double().t() + b.double()
    assert (got - want).abs().max().item() < 1e-5 * max(1.0, want.abs().max().item())


def test_query_merge_algebra():
    g = torch.Generator().manual_seed(2)
    D, R = 128, 5
    x, att = torch.randn(R, D, generator=g).double() + 0.3, torch.randn(R, D, generator=g).double()
    wo, bo = torch.randn(D, D, generator=g) * 0.1, torch.randn(D, generator=g) * 0.1
    wq, bq = torch.randn(D, D, generator=g) * 0.1, torch.randn(D, generator=g) * 0.1
    g2, b2 = 1 + 0.3 * torch.randn(D, generator=g), 0.2 * torch.randn(D, generator=g)
    # what DecoderWeights.__init__ builds
    wq3g, uq, cq = fold_layernorm(wq, bq, g2, b2)
    wg = wq.double() * g2.double().unsqueeze(0)
    wprime3 = split3_weight_compact((wg @ wo.double()).float())
    dvec = (wg @ bo.double()).float()
    # what the three kernels compute: tq rides with q|k|v, q2raw with the attention-output projection, the attention kernel
    # finishes it from the statistics of x1
    tq = x @ _unsplit(wq3g).t()
    x1 = x + att @ wo.double().t() + bo.double()
    q2raw = tq + att @ _unsplit(wprime3).t() + dvec.double()
    mean, var = x1.mean(1, keepdim=True), x1.var(1, unbiased=False, keepdim=True)
    got = (1.0 / torch.sqrt(var + 1e-12)) * (q2raw - mean * uq.double()) + cq.double()
    want = torch.nn.functional.layer_norm(x1, (D,), g2.double(), b2.double(), 1e-12) @ wq.double().t() + bq.double()
    assert (got - want).abs().max().item() < 2e-5 * max(1.0, want.abs().max().item())
