"""GPU: every kernel of libavsr_b200.so, called through the C ABI, against a plain PyTorch fp32 restatement of the op."""
import ctypes as C
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from avsr_b200 import _lib
    _lib.load()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return _lib


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


@pytest.mark.parametrize("M,N,K,bn", [(128, 64, 64, 64), (128, 128, 64, 128), (300, 256, 128, 256), (1000, 1024, 1024, 0),
                                      (777, 200, 104, 0), (5000, 64, 576, 0), (12000, 4096, 1024, 0), (1024, 3001, 1024, 0)])
def test_gemm_tc_plain(L, M, N, K, bn):
    a, b = _rand(M, K, seed=1).bfloat16(), _rand(N, K, seed=2).bfloat16()
    ld = (N + 7) // 8 * 8
    out16 = torch.zeros(M, ld, dtype=torch.bfloat16, device="cuda")
    out32 = torch.zeros(M, ld, dtype=torch.float32, device="cuda")
    L.gemm_bf16(a, b, M, N, K, L.make_epilogue(out_bf16=out16, ld_bf16=ld, out_f32=out32, ld_f32=ld), bn_hint=bn)
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t()
    err = (out32[:, :N] - ref).abs().max().item()
    assert err < 2e-3 * math.sqrt(K), f"fp32 out max-abs {err}"
    assert (out16[:, :N].float() - ref).abs().max().item() < 0.02 * ref.abs().max().item() + 1e-2
    if ld > N:
        assert out32[:, N:].abs().max().item() == 0.0      # nothing written past column N


def test_gemm_tc_epilogues(L):
    M, N, K = 900, 256, 192
    a, b = _rand(M, K, seed=3).bfloat16(), _rand(N, K, seed=4, scale=0.1).bfloat16()
    bias, slope = _rand(N, seed=5), torch.rand(N, device="cuda") * 0.3 + 0.1
    res32, res16 = _rand(M, N, seed=6), _rand(M, N, seed=7).bfloat16()
    acc = a.float() @ b.float().t()
    out = torch.empty(M, N, device="cuda")
    # bias + GELU, then residual (fp32)
    L.gemm_bf16(a, b, M, N, K, L.make_epilogue(bias=bias, act=L.ACT_GELU, residual=res32, ldr=N, out_f32=out, ld_f32=N))
    assert (out - (F.gelu(acc + bias) + res32)).abs().max().item() < 2e-3
    # bias, bf16 residual, PReLU after the residual (ResNet block tail)
    L.gemm_bf16(a, b, M, N, K, L.make_epilogue(bias=bias, act=L.ACT_PRELU, prelu=slope, residual=res16, ldr=N,
                                               act_after_residual=True, out_f32=out, ld_f32=N))
    ref = acc + bias + res16.float()
    ref = torch.where(ref >= 0, ref, ref * slope)
    assert (out - ref).abs().max().item() < 2e-3
    # per-row bias (V^T projection), ReLU
    rb = _rand(M, seed=8)
    L.gemm_bf16(a, b, M, N, K, L.make_epilogue(bias=rb, bias_mode=2, act=L.ACT_RELU, out_f32=out, ld_f32=N))
    assert (out - torch.relu(acc + rb[:, None])).abs().max().item() < 2e-3
    # in-place residual stream update
    h = res32.clone()
    L.gemm_bf16(a, b, M, N, K, L.make_epilogue(bias=bias, residual=h, ldr=N, out_f32=h, ld_f32=N))
    assert (h - (acc + bias + res32)).abs().max().item() < 2e-3


@pytest.mark.parametrize("M,N,K", [(500, 333, 1024), (96, 3072, 1024), (3000, 5049, 1024)])
def test_sgemm(L, M, N, K):
    a, w, bias = _rand(M, K, seed=1), _rand(N, K, seed=2, scale=0.03), _rand(N, seed=3)
    out = torch.empty(M, N, device="cuda")
    L.sgemm(a, w, M, N, K, L.make_epilogue(bias=bias, out_f32=out, ld_f32=N))
    ref = (a.double() @ w.double().t() + bias.double()).float()
    assert (out - ref).abs().max().item() < 2e-5 * math.sqrt(K)


@pytest.mark.parametrize("R,N,K", [(96, 1024, 1024), (160, 3072, 1024), (96, 1024, 3072), (15, 5049, 1024), (3, 1024, 1024)])
def test_sgemm_skinny_and_epilogue(L, R, N, K):
    lib = L.load()
    a, w, bias = _rand(R, K, seed=1), _rand(N, K, seed=2, scale=0.03), _rand(N, seed=3)
    ns = lib.avsr_sgemm_skinny_splits(R, N, K)
    part = torch.empty(ns, R, N, device="cuda")
    L.check(lib.avsr_sgemm_skinny(L.ptr(a), L.ll(K), L.ptr(w), L.ll(K), R, N, K, L.ptr(part), ns, L.stream()), "skinny")
    ref = (a.double() @ w.double().t()).float()
    assert (part.sum(0) - ref).abs().max().item() < 2e-5 * math.sqrt(K)
    if N == 1024:
        res, g, b = _rand(R, N, seed=4), _rand(N, seed=5), _rand(N, seed=6)
        act = torch.ones(R, dtype=torch.int32, device="cuda")
        act[R // 2] = 0
        out, ln = torch.zeros(R, N, device="cuda"), torch.zeros(R, N, device="cuda")
        L.check(lib.avsr_splitk_epilogue(L.ptr(part), ns, R, N, L.ptr(bias), 0, L.ptr(res), L.ll(N), L.ptr(out), L.ll(N), L.ptr(g),
                                         L.ptr(b), C.c_float(1e-12), L.ptr(ln), L.ll(N), L.ptr(act), None, L.stream()), "epilogue")
        want = ref + bias + res
        keep = act.bool()
        assert (out[keep] - want[keep]).abs().max().item() < 1e-4
        assert (ln[keep] - F.layer_norm(want, (N,), g, b, 1e-12)[keep]).abs().max().item() < 1e-4
        assert out[~keep].abs().max().item() == 0.0
        # compact bf16x3 output of the same LayerNorm ([a1|a2|a3]): the three terms must add back to the fp32 value (~2^-24 relative)
        sp = torch.zeros(R, 3 * N, dtype=torch.bfloat16, device="cuda")
        L.check(lib.avsr_splitk_epilogue(L.ptr(part), ns, R, N, L.ptr(bias), 0, L.ptr(res), L.ll(N), None, L.ll(N), L.ptr(g),
                                         L.ptr(b), C.c_float(1e-12), None, L.ll(N), L.ptr(act), L.ptr(sp), L.stream()), "epilogue split")
        blocks = sp.float().view(R, 3, N)
        back = blocks[:, 0].double() + blocks[:, 1].double() + blocks[:, 2].double()
        assert (back[keep] - ln[keep].double()).abs().max().item() < 1e-6
        assert (blocks[:, 1][keep].abs() <= blocks[:, 0][keep].abs() * 2.0 ** -7 + 1e-30).all()


def _split3c(a):
    a1 = a.bfloat16()
    r = a - a1.float()
    a2 = r.bfloat16()
    a3 = (r - a2.float()).bfloat16()
    return torch.cat([a1, a2, a3], 1).contiguous()


@pytest.mark.parametrize("R,N,K", [(96, 1024, 1024), (96, 3072, 1024), (160, 1024, 3072), (96, 5049, 1024), (3, 1024, 1024), (300, 1024, 64)])
def test_gemm_x3_is_fp32_accurate(L, R, N, K):
    """The decoder step's own GEMM: compact three-term operands, six MMAs per k step, split-K partials (csrc/gemm_x3.cu)."""
    from avsr_b200.weights import split3_weight_compact
    lib = L.load()
    a, w = _rand(R, K, seed=1), _rand(N, K, seed=2, scale=0.03)
    a3, w3 = _split3c(a), split3_weight_compact(w)
    ns = lib.avsr_gemm_x3_splits(R, N, K)
    assert 1 <= ns <= K // 64
    part = torch.full((ns, R, N), float("nan"), device="cuda")
    for _ in range(2):
        L.check(lib.avsr_gemm_x3_splitk(L.ptr(a3), L.ll(3 * K), L.ptr(w3), L.ll(3 * K), R, N, K, L.ptr(part), L.stream()), "gemm_x3")
    ref = a.double() @ w.double().t()
    got = part.double().sum(0)
    assert not torch.isnan(got).any()
    err = (got - ref).abs().max().item()
    fp32_err = ((a @ w.t()).double() - ref).abs().max().item()
    assert err < 2e-5 * (K ** 0.5) and err < 4 * fp32_err + 1e-6, (err, fp32_err)


@pytest.mark.parametrize("R,N,K", [(96, 1024, 1024), (96, 3072, 1024), (160, 1024, 3072), (96, 5049, 1024), (3, 1024, 1024)])
def test_bf16x3_tensor_core_gemm_is_fp32_accurate(L, R, N, K):
    """Decode-side projections on the tensor cores: activations / weights split into three bf16 terms, split-K partials."""
    from avsr_b200.weights import split3_weight
    lib = L.load()
    a, w = _rand(R, K, seed=1), _rand(N, K, seed=2, scale=0.03)
    a6 = torch.zeros(R, 6 * K, dtype=torch.bfloat16, device="cuda")
    L.check(lib.avsr_split3(L.ptr(a), L.ll(K), L.ptr(a6), L.ll(R), K, L.stream()), "split3")
    w6 = split3_weight(w)
    bn, ns = (64 if N <= 1024 else 128), 6
    part = torch.full((ns, R, N), float("nan"), device="cuda")
    L.check(lib.avsr_gemm_bf16_tc_splitk(L.ptr(a6), L.ll(6 * K), L.ptr(w6), L.ll(6 * K), R, N, 6 * K, L.ptr(part), ns, bn, L.stream()), "tc splitk")
    ref = a.double() @ w.double().t()
    got = part.double().sum(0)
    err = (got - ref).abs().max().item()
    fp32_err = ((a @ w.t()).double() - ref).abs().max().item()
    assert err < 2e-5 * (K ** 0.5) and err < 8 * fp32_err + 1e-6, (err, fp32_err)


@pytest.mark.parametrize("rows,N", [(777, 2048), (12003, 1024), (5, 1024), (64, 1792)])
def test_layernorm(L, rows, N):
    """Block-per-row kernel (any N) and the warp-per-row kernel of the transformer width (N = 1024), ragged row counts, in place."""
    x, g, b = _rand(rows, N, seed=1, scale=3.0) + 2.0, _rand(N, seed=2), _rand(N, seed=3)
    o16 = torch.empty(rows, N, dtype=torch.bfloat16, device="cuda")
    o32 = torch.empty(rows, N, device="cuda")
    L.layernorm(x, g, b, 1e-5, out_bf16=o16, out_f32=o32)
    ref = F.layer_norm(x, (N,), g, b, 1e-5)
    assert (o32 - ref).abs().max().item() < 1e-4
    assert (o16.float() - ref).abs().max().item() < 0.05
    xi = x.clone()
    L.layernorm(xi, g, b, 1e-5, out_f32=xi)
    assert torch.equal(xi, o32)


def test_im2col_pool_kernels(L):
    lib = L.load()
    # 2D im2col against F.unfold (channels-last k ordering (ky, kx, c))
    for (H, Cc, ks, s) in ((22, 64, 3, 1), (22, 64, 3, 2), (11, 128, 1, 2), (6, 256, 3, 2), (3, 512, 3, 1)):
        Fr = 5
        x = _rand(Fr, H, H, Cc, seed=H).bfloat16()
        pad = ks // 2
        Ho = (H + 2 * pad - ks) // s + 1
        out = torch.empty(Fr * Ho * Ho, ks * ks * Cc, dtype=torch.bfloat16, device="cuda")
        L.check(lib.avsr_im2col2d(L.ptr(x), L.ptr(out), L.ll(Fr), H, H, Cc, ks, s, L.stream()), "im2col2d")
        u = F.unfold(x.float().permute(0, 3, 1, 2), ks, padding=pad, stride=s)          # [F, C*ks*ks, L]
        u = u.view(Fr, Cc, ks * ks, Ho * Ho).permute(0, 3, 2, 1).reshape(Fr * Ho * Ho, ks * ks * Cc)
        assert torch.equal(out.float(), u), (H, Cc, ks, s)
    # max pool
    x = _rand(4, 44, 44, 64, seed=9).bfloat16()
    out = torch.empty(4, 22, 22, 64, dtype=torch.bfloat16, device="cuda")
    L.check(lib.avsr_maxpool3x3s2(L.ptr(x), L.ptr(out), L.ll(4), 44, 44, 64, L.stream()), "maxpool")
    ref = F.max_pool2d(x.float().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    assert torch.equal(out.float(), ref)
    # avg pool
    x = _rand(7, 9, 512, seed=10).bfloat16()
    out = torch.empty(7, 512, dtype=torch.bfloat16, device="cuda")
    L.check(lib.avsr_avgpool(L.ptr(x), L.ptr(out), L.ll(7), 9, 512, L.stream()), "avgpool")
    assert (out.float() - x.float().mean(1)).abs().max().item() < 0.02


@pytest.mark.parametrize("Fr,H,Cin,Cout,ks,stride", [(3, 22, 64, 64, 3, 1), (5, 11, 128, 128, 3, 1), (7, 6, 256, 256, 3, 1),
                                                     (9, 3, 512, 512, 3, 1), (40, 22, 64, 64, 3, 1), (1, 3, 64, 128, 3, 1),
                                                     (5, 22, 64, 128, 3, 2), (5, 22, 64, 128, 1, 2), (6, 11, 128, 256, 3, 2),
                                                     (6, 11, 128, 256, 1, 2), (7, 6, 256, 512, 3, 2), (7, 6, 256, 512, 1, 2)])
def test_implicit_gemm_conv_equals_im2col_gemm(L, Fr, H, Cin, Cout, ks, stride):
    """avsr_conv2d_bf16_tc (im2col-mode TMA, nothing materialised) against the explicit im2col + GEMM path (same k order, same
    MMAs: bit-identical) and against F.conv2d in fp32 (BasicBlock / downsample convolutions, resnet.py:30-69)."""
    lib = L.load()
    pad = ks // 2
    Ho = (H + 2 * pad - ks) // stride + 1
    x = _rand(Fr, H, H, Cin, seed=H + Cin).bfloat16()
    w = (_rand(Cout, ks * ks * Cin, seed=3) * 0.05).bfloat16()
    bias = _rand(Cout, seed=4)
    slope = torch.full((Cout,), 0.25, device="cuda")
    M = Fr * Ho * Ho
    res = _rand(M, Cout, seed=5).bfloat16()
    out_i = torch.zeros(M, Cout, dtype=torch.bfloat16, device="cuda")
    out_e = torch.zeros(M, Cout, dtype=torch.bfloat16, device="cuda")
    kw = dict(bias=bias, act=L.ACT_PRELU, prelu=slope, residual=res, ldr=Cout, act_after_residual=True, ld_bf16=Cout)
    L.conv2d_bf16(x, w, Fr, H, H, Cin, Cout, ks, stride, L.make_epilogue(out_bf16=out_i, **kw))
    col = torch.empty(M, ks * ks * Cin, dtype=torch.bfloat16, device="cuda")
    L.check(lib.avsr_im2col2d(L.ptr(x), L.ptr(col), L.ll(Fr), H, H, Cin, ks, stride, L.stream()), "im2col2d")
    L.gemm_bf16(col, w, M, Cout, ks * ks * Cin, L.make_epilogue(out_bf16=out_e, **kw))
    torch.cuda.synchronize()
    assert torch.equal(out_i, out_e), (out_i.float() - out_e.float()).abs().max().item()
    wt = w.float().view(Cout, ks, ks, Cin).permute(0, 3, 1, 2)                      # [Cout, Cin, ky, kx]
    y = F.conv2d(x.float().permute(0, 3, 1, 2), wt, bias, padding=pad, stride=stride).permute(0, 2, 3, 1).reshape(M, Cout) + res.float()
    y = torch.where(y >= 0, y, 0.25 * y)
    assert (out_i.float() - y).abs().max().item() < 0.05 * max(1.0, y.abs().max().item() / 8)


def test_frontend_and_posconv_im2col(L):
    lib = L.load()
    lengths = [5, 3]
    Fr = sum(lengths)
    video = _rand(Fr, 88, 88, seed=11)
    ft = torch.tensor([0, 1, 2, 3, 4, 0, 1, 2], dtype=torch.int32, device="cuda")
    fT = torch.tensor([5] * 5 + [3] * 3, dtype=torch.int32, device="cuda")
    out = torch.empty(Fr * 1936, 256, dtype=torch.bfloat16, device="cuda")
    L.check(lib.avsr_im2col_frontend(L.ptr(video), L.ptr(ft), L.ptr(fT), 0, Fr, L.ptr(out), L.stream()), "im2col_frontend")
    # restatement: per utterance, pad time by 2 and space by 3, unfold 5x7x7 with stride (1,2,2)
    o = 0
    for T in lengths:
        v = F.pad(video[o:o + T].bfloat16().float(), (3, 3, 3, 3, 2, 2))              # [T+4, 94, 94]
        p = v.unfold(0, 5, 1).unfold(1, 7, 2).unfold(2, 7, 2)                          # [T,44,44,5,7,7]
        ref = p.reshape(T * 1936, 245)
        got = out[o * 1936:(o + T) * 1936].float()
        assert torch.equal(got[:, :245], ref)
        assert got[:, 245:].abs().max().item() == 0.0
        o += T
    # positional-conv patches
    x = _rand(Fr, 1024, seed=12).bfloat16()
    pc = torch.empty(16, Fr, 8192, dtype=torch.bfloat16, device="cuda")
    L.check(lib.avsr_posconv_im2col(L.ptr(x), L.ptr(pc), L.ptr(ft), L.ptr(fT), L.ll(Fr), 0, 16, L.stream()), "posconv_im2col")
    o = 0
    for T in lengths:
        xp = F.pad(x[o:o + T].float(), (0, 0, 64, 64))                                  # [T+128, 1024]
        for g in (0, 7, 15):
            for t in range(T):
                ref = xp[t:t + 128, g * 64:(g + 1) * 64].reshape(-1)
                assert torch.equal(pc[g, o + t].float(), ref)
        o += T


@pytest.mark.parametrize("lengths", [[12], [30, 7], [375], [130, 257, 64], [700]])
def test_attention_varlen(L, lengths):
    lib = L.load()
    Fr = sum(lengths)
    Fld = (Fr + 7) // 8 * 8
    qk = _rand(Fr, 2048, seed=21, scale=0.6).bfloat16()
    v = _rand(Fr, 1024, seed=22).bfloat16()
    vt = torch.zeros(1024, Fld, dtype=torch.bfloat16, device="cuda")
    vt[:, :Fr] = v.t()
    out = torch.zeros(Fr, 1024, dtype=torch.bfloat16, device="cuda")
    offs = np.concatenate([[0], np.cumsum(lengths)[:-1]])
    work = [(int(offs[b]), t, q0) for b, t in enumerate(lengths) for q0 in range(0, t, 128)]
    wo, wT, wq = (torch.tensor([w[i] for w in work], dtype=torch.int32, device="cuda") for i in range(3))
    L.check(lib.avsr_attention_varlen(L.ptr(qk), L.ptr(vt), L.ll(Fld), L.ptr(out), L.ll(Fr), L.ptr(wo), L.ptr(wT), L.ptr(wq),
                                      len(work), max(lengths), L.stream()), "attention")
    torch.cuda.synchronize()
    o = 0
    for T in lengths:
        q = qk[o:o + T, :1024].float().view(T, 16, 64).transpose(0, 1)
        k = qk[o:o + T, 1024:].float().view(T, 16, 64).transpose(0, 1)
        vv = v[o:o + T].float().view(T, 16, 64).transpose(0, 1)
        ref = (torch.softmax(q @ k.transpose(1, 2), -1) @ vv).transpose(0, 1).reshape(T, 1024)     # scale pre-folded into q
        err = (out[o:o + T].float() - ref).abs().max().item()
        assert err < 0.03, (T, err)
        o += T


def test_log_softmax_and_logits_topk(L):
    lib = L.load()
    V, R, beam, S = 5049, 6, 3, 4
    x = _rand(50, V, seed=31, scale=3.0)
    ref = torch.log_softmax(x, -1)
    L.check(lib.avsr_log_softmax_rows(L.ptr(x), L.ll(V), L.ll(50), V, L.stream()), "log_softmax")
    assert (x - ref).abs().max().item() < 1e-5
    part = _rand(2, R, V, seed=32)
    bias = _rand(V, seed=33)
    n_run = torch.tensor([3, 2], dtype=torch.int32, device="cuda")
    logp = torch.zeros(R, V, device="cuda")
    ids = torch.full((R, S), -1, dtype=torch.int32, device="cuda")
    L.check(lib.avsr_dec_logits_lsm_topk(L.ptr(part), 2, R, V, L.ptr(bias), L.ptr(n_run), beam, L.ptr(logp), L.ptr(ids), S, L.stream()), "lsm_topk")
    want = torch.log_softmax(part.sum(0) + bias, -1)
    for r in (0, 1, 2, 3, 4):
        assert (logp[r] - want[r]).abs().max().item() < 1e-5
        assert ids[r].tolist() == torch.topk(want[r], S)[1].tolist()
    assert ids[5].tolist() == [-1] * S                       # dead row untouched


def _as_partials(x, nsplit, seed):
    """[R, N] -> (part [nsplit, R, N], bias [N]) with sum_z part[z] + bias == x up to fp32 rounding (the attention kernels can
    take their query straight from the split-K partial sums of the projection)."""
    if nsplit == 0:
        return x, None
    g = torch.Generator().manual_seed(seed)
    bias = torch.randn(x.shape[1], generator=g).cuda()
    part = torch.randn(nsplit, *x.shape, generator=g).cuda()
    part[nsplit - 1] = x - bias - part[:nsplit - 1].sum(0)
    return part.contiguous(), bias


@pytest.mark.parametrize("beam,step,nsplit,dense", [(3, 0, 0, False), (3, 5, 6, True), (3, 127, 0, False), (3, 128, 3, True),
                                                    (5, 300, 0, True), (3, 375, 6, True), (8, 140, 2, False), (4, 260, 0, True),
                                                    (3, 200, 0, False)])
def test_dec_attn_step_self(L, beam, step, nsplit, dense):
    """Self-attention of one decode position over a random cache and a random ancestry table (decoder_layer.py:82-93,
    attention.py:38-106 restated in torch fp32): streaming kernel incl. the distinct-row list, the multi-chunk merge, the
    k/v append and the query taken from split-K partial sums."""
    lib = L.load()
    B, lmax = 4, 376
    R = B * beam
    g = torch.Generator().manual_seed(100 + step)
    kc_log = torch.randn(B, 16, lmax, beam, 64, generator=g).cuda()    # logical [utt][head][pos][slot][64]
    vc = torch.randn(B, 16, lmax, beam, 64, generator=g).cuda()
    # physical key cache: transposed in 32-byte groups, [utt][head][8][pos*beam+slot][8]
    to_phys = lambda k: k.reshape(B, 16, lmax * beam, 8, 8).permute(0, 1, 3, 2, 4).contiguous()
    to_log = lambda k: k.permute(0, 1, 3, 2, 4).reshape(B, 16, lmax, beam, 64)
    kc = to_phys(kc_log)
    qkv = torch.randn(R, 3072, generator=g).cuda()
    q_in, q_bias = _as_partials(qkv, nsplit, 5)
    n_run = torch.tensor([beam, max(1, beam - 1), 0, 1][:B], dtype=torch.int32, device="cuda")
    anc = torch.randint(0, beam, (2, R, lmax), generator=g, dtype=torch.uint8)
    shared = max(0, step - 6)                           # old positions: all hyps of an utterance share one ancestor slot
    anc[:, :, :shared] = anc[:, ::beam, :shared].repeat_interleave(beam, 1)
    anc = anc.cuda()
    step_t = torch.tensor([step], dtype=torch.int32, device="cuda")
    out = torch.full((R, 1024), 7.0, device="cuda")
    out6 = torch.zeros(R, 3 * 1024, dtype=torch.bfloat16, device="cuda")
    kc0, vc0 = kc_log.clone(), vc.clone()
    kd = vd = conv = None
    if dense:
        # dense caches of the converged prefix, filled by the promotion kernel in a few calls (at most 32 positions each); the
        # per-slot rows of promoted positions are then poisoned: the attention must read the dense copies
        kd = torch.zeros(1, B, 16, 8, lmax, 8, device="cuda")
        vd = torch.zeros(1, B, 16, lmax, 64, device="cuda")
        conv = torch.zeros(2, B, dtype=torch.int32, device="cuda")
        for _ in range(step // 32 + 2):
            L.check(lib.avsr_dec_cache_promote(L.ptr(kc), L.ptr(vc), L.ptr(kd), L.ptr(vd), 1, L.ptr(anc), lmax, L.ptr(n_run), beam, R,
                                               L.ptr(step_t), L.ptr(conv), L.stream()), "promote")
            conv[step & 1] = conv[(step + 1) & 1]            # next call continues where this one stopped
        torch.cuda.synchronize()
        cl = conv[(step + 1) & 1].cpu().tolist()
        a_now = anc[step & 1].cpu().long()
        for b in range(B):
            nh = int(n_run[b])
            want_c = 0
            while nh > 0 and want_c < step and len({int(a_now[b * beam + h, want_c]) for h in range(nh)}) == 1:
                want_c += 1
            assert cl[b] == want_c, (b, cl[b], want_c)
            if cl[b] > 0:
                slots = a_now[b * beam, :cl[b]]
                kd_log = kd[0, b].permute(0, 2, 1, 3).reshape(16, lmax, 64)
                assert torch.equal(kd_log[:, :cl[b]], kc_log[b][:, torch.arange(cl[b]), slots])
                assert torch.equal(vd[0, b][:, :cl[b]], vc[b][:, torch.arange(cl[b]), slots])
                kcl = to_log(kc)
                kcl[b, :, :cl[b]] = float("nan")
                kc = to_phys(kcl)
                vc[b, :, :cl[b]] = float("nan")
        vc = vc.contiguous().view(B, 16, lmax * beam, 64)
    for _ in range(2):                                   # twice: the second call re-reads the k / v the first one appended
        L.check(lib.avsr_dec_attn_step(0, L.ptr(q_in), L.ll(3072), nsplit, L.ptr(q_bias), L.ptr(kc), L.ptr(vc), L.ptr(anc), lmax,
                                       L.ptr(n_run), None, None, beam, R, L.ptr(step_t), L.ptr(out), L.ll(0), L.ptr(out6),
                                       L.ptr(kd), L.ptr(vd), L.ptr(conv), L.stream()), "dec_attn_step(self)")
    torch.cuda.synchronize()
    vc = vc.view(B, 16, lmax, beam, 64)
    kc = to_log(kc)
    a = anc[step & 1].cpu().long()
    tol = 2e-5 if nsplit == 0 else 1e-4                  # the partial-sum form adds the rounding of the split sums
    for b in range(B):
        for h in range(beam):
            row = b * beam + h
            if h >= int(n_run[b]):
                assert (out[row] == 7.0).all()           # dead rows untouched
                assert (kc[b, :, step, h] == kc0[b, :, step, h]).all()
                continue
            slots = torch.tensor([int(a[row, p]) for p in range(step)], dtype=torch.long)
            q = qkv[row, :1024].view(16, 64)
            kcur, vcur = qkv[row, 1024:2048].view(16, 64), qkv[row, 2048:].view(16, 64)
            assert (kc[b, :, step, h] - kcur).abs().max().item() < tol and (vc[b, :, step, h] - vcur).abs().max().item() < tol
            K = torch.cat([kc0[b][:, torch.arange(step), slots], kc[b, :, step, h][:, None]], 1)          # [16, step+1, 64]
            Vv = torch.cat([vc0[b][:, torch.arange(step), slots], vc[b, :, step, h][:, None]], 1)
            att = torch.softmax(torch.einsum("hd,hpd->hp", q, K) / 8.0, -1)
            want = torch.einsum("hp,hpd->hd", att, Vv).reshape(1024)
            assert (out[row] - want).abs().max().item() < tol, (b, h)
            o3 = out6[row].float().view(3, 1024)
            assert (o3[0] + o3[1] + o3[2] - want).abs().max().item() < tol
    # positions other than the current one are never written (poisoned rows of the dense variant stay poisoned)
    keep = torch.ones(lmax, dtype=torch.bool)
    keep[step] = False
    same = lambda x, y: bool(((x == y) | (x != x)).all())
    assert same(kc[:, :, keep], kc0[:, :, keep]) and same(vc[:, :, keep], vc0[:, :, keep])


@pytest.mark.parametrize("beam,lengths,nsplit", [(3, [375, 12, 130, 257], 0), (5, [128, 129, 1, 375], 16), (3, [400, 384, 385, 900], 4)])
def test_dec_attn_step_cross(L, beam, lengths, nsplit):
    """Source attention of one decode position: every live hyp of an utterance attends over that utterance's frames."""
    lib = L.load()
    B, tmax, Fr = len(lengths), max(lengths), sum(lengths)
    R = B * beam
    g = torch.Generator().manual_seed(7)
    kc = torch.randn(16, Fr, 64, generator=g).cuda()
    vc = torch.randn(16, Fr, 64, generator=g).cuda()
    kc_phys = kc.reshape(16, Fr, 8, 8).permute(0, 2, 1, 3).contiguous()        # keys transposed in 32-byte groups: [head][8][F][8]
    q = torch.randn(R, 1024, generator=g).cuda()
    q_in, q_bias = _as_partials(q, nsplit, 6)
    n_run = torch.tensor([beam, 1, 0, beam - 1][:B], dtype=torch.int32, device="cuda")
    offs = np.concatenate([[0], np.cumsum(lengths)[:-1]]).astype(np.int32)
    utt_off = torch.from_numpy(offs).cuda()
    utt_T = torch.tensor(lengths, dtype=torch.int32, device="cuda")
    step_t = torch.tensor([3], dtype=torch.int32, device="cuda")
    out = torch.full((R, 1024), 7.0, device="cuda")
    for _ in range(2):
        L.check(lib.avsr_dec_attn_step(1, L.ptr(q_in), L.ll(1024), nsplit, L.ptr(q_bias), L.ptr(kc_phys), L.ptr(vc), None, tmax + 1, L.ptr(n_run),
                                       L.ptr(utt_off), L.ptr(utt_T), beam, R, L.ptr(step_t), L.ptr(out), L.ll(Fr), None, None, None, None,
                                       L.stream()), "dec_attn_step(src)")
    torch.cuda.synchronize()
    tol = 2e-5 if nsplit == 0 else 1e-4
    for b in range(B):
        K, Vv = kc[:, offs[b]:offs[b] + lengths[b]], vc[:, offs[b]:offs[b] + lengths[b]]
        for h in range(beam):
            row = b * beam + h
            if h >= int(n_run[b]):
                assert (out[row] == 7.0).all()
                continue
            att = torch.softmax(torch.einsum("hd,hpd->hp", q[row].view(16, 64), K) / 8.0, -1)
            want = torch.einsum("hp,hpd->hd", att, Vv).reshape(1024)
            assert (out[row] - want).abs().max().item() < tol, (b, h)


@pytest.mark.parametrize("R", [3, 96, 128])
def test_query_merge_equals_separate_projection(L, R):
    """avsr_dec_proj_dual + avsr_dec_attn_fold_query: the source-attention query computed as extra output columns of the
    q|k|v projection (operand x) and of the attention-output projection (operand att), finished by the attention kernel from
    the row statistics, against LayerNorm2(x + att Wo^T + bo) Wq^T + bq in float64; and the first column ranges (q|k|v with the
    folded LayerNorm1, x1 with its statistics) against the single-range launches they replace."""
    from avsr_b200.weights import fold_layernorm, split3_weight_compact
    lib = L.load()
    g = torch.Generator().manual_seed(R)
    D = 1024
    rnd = lambda *sh, sc=1.0: (torch.randn(*sh, generator=g) * sc).cuda()
    x, att = rnd(R, D) + 0.7, rnd(R, D)
    wqkv, bqkv, wo, bo, wq2, bq2 = rnd(3 * D, D, sc=0.04), rnd(3 * D, sc=0.1), rnd(D, D, sc=0.04), rnd(D, sc=0.1), rnd(D, D, sc=0.04), rnd(D, sc=0.1)
    g1, b1, g2, b2 = 1 + rnd(D, sc=0.2), rnd(D, sc=0.2), 1 + rnd(D, sc=0.2), rnd(D, sc=0.2)
    wqkv3g, uqkv, cqkv = fold_layernorm(wqkv, bqkv, g1, b1)
    wq23g, uq2, cq2 = fold_layernorm(wq2, bq2, g2, b2)
    wg = wq2.double().cpu() * g2.double().cpu().unsqueeze(0)
    wcat1 = torch.cat([wqkv3g, wq23g], 0).contiguous()
    ucat1, ccat1 = torch.cat([uqkv, torch.zeros(D, device="cuda")]).contiguous(), torch.cat([cqkv, torch.zeros(D, device="cuda")]).contiguous()
    wo3 = split3_weight_compact(wo)
    wcat2 = torch.cat([wo3, split3_weight_compact((wg @ wo.double().cpu()).float().cuda())], 0).contiguous()
    bcat2 = torch.cat([bo, (wg @ bo.double().cpu()).float().cuda()]).contiguous()
    # operands as the chain has them: x3 / att3 = compact bf16x3 rows, stats of x in the tile format
    x3, att3 = split3_weight_compact(x), split3_weight_compact(att)
    xt = x.view(R, 8, 128)
    stats = torch.stack([xt.mean(-1), ((xt - xt.mean(-1, keepdim=True)) ** 2).sum(-1)], -1).permute(1, 0, 2).contiguous()      # [8][R][2]
    qkv_d, tq = torch.empty(R, 3 * D, device="cuda"), torch.empty(R, D, device="cuda")
    L.check(lib.avsr_dec_proj_dual(L.ptr(x3), L.ll(3 * D), L.ptr(stats), C.c_float(1e-12), L.ptr(ucat1), L.ptr(ccat1), L.ptr(wcat1), L.ll(3 * D),
                                   R, 4 * D, D, 3 * D, L.ACT_NONE, None, L.ll(D), L.ptr(qkv_d), L.ll(3 * D), None, L.ll(D), L.ptr(tq), L.ll(D),
                                   None, None, L.ll(0), L.stream()), "dual qkv")
    qkv_s = torch.empty(R, 3 * D, device="cuda")
    L.check(lib.avsr_dec_proj_folded(L.ptr(x3), L.ll(3 * D), L.ptr(stats), C.c_float(1e-12), L.ptr(uqkv), L.ptr(cqkv), L.ptr(wqkv3g), L.ll(3 * D),
                                     R, 3 * D, D, L.ACT_NONE, None, L.ll(D), L.ptr(qkv_s), L.ll(3 * D), None, None, None, L.ll(0), L.stream()), "folded qkv")
    x1_d, q2raw, stats2 = x.clone(), torch.empty(R, D, device="cuda"), torch.zeros(8, R, 2, device="cuda")
    L.check(lib.avsr_dec_proj_dual(L.ptr(att3), L.ll(3 * D), None, C.c_float(1e-12), None, L.ptr(bcat2), L.ptr(wcat2), L.ll(3 * D), R, 2 * D, D, D,
                                   L.ACT_NONE, L.ptr(x1_d), L.ll(D), L.ptr(x1_d), L.ll(D), L.ptr(tq), L.ll(D), L.ptr(q2raw), L.ll(D),
                                   L.ptr(stats2), None, L.ll(0), L.stream()), "dual out")
    x1_s, stats2_s = x.clone(), torch.zeros(8, R, 2, device="cuda")
    L.check(lib.avsr_dec_proj(L.ptr(att3), L.ll(3 * D), None, L.ll(0), None, None, None, C.c_float(1e-12), L.ptr(wo3), L.ll(3 * D), R, D, D, L.ptr(bo), 0,
                              L.ptr(x1_s), L.ll(D), L.ptr(x1_s), L.ll(D), None, L.ptr(stats2_s), None, L.ll(0), L.stream()), "single out")
    torch.cuda.synchronize()
    # the stacked launches plan a different cluster size (more output tiles): the same sums in a different split order
    assert (x1_d - x1_s).abs().max().item() < 2e-5 * max(1.0, x1_s.abs().max().item())
    assert (stats2 - stats2_s).abs().max().item() < 1e-3
    assert (qkv_d - qkv_s).abs().max().item() < 2e-5 * max(1.0, qkv_s.abs().max().item())
    # finished query: what the attention kernel computes from q2raw and the statistics of x1
    x1 = x.double().cpu() + att.double().cpu() @ wo.double().cpu().t() + bo.double().cpu()
    ln = torch.nn.functional.layer_norm(x1, (D,), g2.double().cpu(), b2.double().cpu(), 1e-12)
    want = ln @ wq2.double().cpu().t() + bq2.double().cpu()
    st = stats2.double().cpu()
    mean = st[:, :, 0].mean(0)
    m2 = (st[:, :, 1] + 128.0 * (st[:, :, 0] - mean) ** 2).sum(0)
    rstd = 1.0 / torch.sqrt(m2 / D + 1e-12)
    got = rstd[:, None] * (q2raw.double().cpu() - mean[:, None] * uq2.double().cpu()) + cq2.double().cpu()
    err = (got - want).abs().max().item()
    assert err < 5e-5 * max(1.0, want.abs().max().item()), err


@pytest.mark.parametrize("nf,with_res", [(3, False), (37, True), (300, True)])
def test_conv3x3_halo_equals_implicit_gemm(L, nf, with_res):
    """ResNet layer1 convolution on the padded layout (halo staged once, nine taps = nine shifted descriptors of the same tile)
    against the generic implicit-GEMM convolution on the dense layout: bit-identical on the valid pixels (same tap / k order),
    pad cells of the output untouched; the pitched max-pool and the pitched convolution read / write that layout in place."""
    lib = L.load()
    H = W = 22
    Hp, Wp = H + 1, W + 2
    g = torch.Generator().manual_seed(nf)
    x = torch.randn(nf, H, W, 64, generator=g).bfloat16().cuda()
    w = (torch.randn(64, 576, generator=g) * 0.05).bfloat16().cuda()
    bias, slope = torch.randn(64, generator=g).cuda(), torch.rand(64, generator=g).cuda()
    res = torch.randn(nf, H, W, 64, generator=g).bfloat16().cuda()
    xp = torch.zeros(nf, Hp, Wp, 64, dtype=torch.bfloat16, device="cuda")
    xp[:, :H, :W] = x
    rp = torch.zeros_like(xp)
    rp[:, :H, :W] = res
    want = torch.empty(nf * H * W, 64, dtype=torch.bfloat16, device="cuda")
    kw = dict(bias=bias, act=L.ACT_PRELU, prelu=slope)
    kw_d = dict(kw, residual=res.view(-1, 64), ldr=64, act_after_residual=True) if with_res else kw
    kw_p = dict(kw, residual=rp.view(-1, 64), ldr=64, act_after_residual=True) if with_res else kw
    L.conv2d_bf16(x, w, nf, H, W, 64, 64, 3, 1, L.make_epilogue(out_bf16=want, ld_bf16=64, **kw_d))
    got = torch.full((nf, Hp, Wp, 64), 3.0, dtype=torch.bfloat16, device="cuda")
    epp = L.make_epilogue(out_bf16=got.view(-1, 64), ld_bf16=64, **kw_p)
    L.check(lib.avsr_conv3x3_halo_bf16(L.ptr(xp), L.ptr(w), L.ll(nf), H, W, C.byref(epp), L.stream()), "halo conv")
    torch.cuda.synchronize()
    assert torch.equal(got[:, :H, :W].reshape(-1, 64), want)
    assert (got[:, H] == 3.0).all() and (got[:, :, W:] == 3.0).all()
    # pitched max-pool writes the padded layout in place; pitched convolution (stride 2, the entry of layer2) reads it in place
    c0 = torch.randn(nf, 44, 44, 64, generator=g).bfloat16().cuda()
    dense = torch.empty(nf, H, W, 64, dtype=torch.bfloat16, device="cuda")
    padded = torch.zeros(nf, Hp, Wp, 64, dtype=torch.bfloat16, device="cuda")
    L.check(lib.avsr_maxpool3x3s2(L.ptr(c0), L.ptr(dense), L.ll(nf), 44, 44, 64, L.stream()), "maxpool")
    L.check(lib.avsr_maxpool3x3s2_pitched(L.ptr(c0), L.ptr(padded), L.ll(nf), 44, 44, 64, L.ll(Wp), L.ll(Hp * Wp), L.stream()), "maxpool pitched")
    torch.cuda.synchronize()
    assert torch.equal(padded[:, :H, :W], dense) and (padded[:, H] == 0).all() and (padded[:, :, W:] == 0).all()
    w2 = (torch.randn(128, 576, generator=g) * 0.05).bfloat16().cuda()
    o_d = torch.empty(nf * 121, 128, dtype=torch.bfloat16, device="cuda")
    o_p = torch.empty_like(o_d)
    L.conv2d_bf16(dense, w2, nf, H, W, 64, 128, 3, 2, L.make_epilogue(out_bf16=o_d, ld_bf16=128))
    ep2 = L.make_epilogue(out_bf16=o_p, ld_bf16=128)
    L.check(lib.avsr_conv2d_bf16_tc_pitched(L.ptr(padded), L.ptr(w2), L.ll(nf), H, W, 64, 128, 3, 2, L.ll(Wp), L.ll(Hp * Wp), C.byref(ep2), L.stream()),
            "conv pitched")
    torch.cuda.synchronize()
    assert torch.equal(o_d, o_p)


@pytest.mark.parametrize("M,N,K,mode", [(12000, 1024, 1024, "res"), (9000, 4096, 512, "gelu"), (5000, 5049, 384, "f32"),
                                        (37889, 1024, 128, "prelu")])
def test_gemm_cta_pair_kernel(L, M, N, K, mode):
    """The CTA-pair (tcgen05 cta_group::2) GEMM the big encoder projections dispatch to (>= 148 tiles of 256 x 256): ragged M
    (last pair half / fully out of range), N not a multiple of 256, every epilogue flavour, against torch fp32 on the same
    bf16 operands; and bit-identical to the single-CTA kernel (same k order, same fp32 accumulation)."""
    g = torch.Generator().manual_seed(M + N)
    a = (torch.randn(M, K, generator=g) * 0.5).bfloat16().cuda()
    b = (torch.randn(N, K, generator=g) * 0.5).bfloat16().cuda()
    bias = torch.randn(N, generator=g).cuda()
    ref = a.float() @ b.float().t() + bias
    kw = dict(bias=bias)
    if mode == "res":
        res = torch.randn(M, N, generator=g).cuda()
        kw.update(residual=res, ldr=N)
        ref = ref + res
    elif mode == "gelu":
        kw.update(act=L.ACT_GELU)
        ref = torch.nn.functional.gelu(ref)
    elif mode == "prelu":
        slope = torch.rand(N, generator=g).cuda()
        kw.update(act=L.ACT_PRELU, prelu=slope)
        ref = torch.where(ref > 0, ref, ref * slope)
    outs = []
    for bn in (0, 256):                           # 0: automatic (the pair kernel at these sizes), 256: the single-CTA kernel
        o32 = torch.full((M, N), 7.0, device="cuda")
        o16 = torch.full((M, N), 7.0, device="cuda", dtype=torch.bfloat16)
        L.gemm_bf16(a, b, M, N, K, L.make_epilogue(out_f32=o32, ld_f32=N, out_bf16=o16, ld_bf16=N, **kw), bn_hint=bn)
        torch.cuda.synchronize()
        outs.append((o32, o16))
    err = (outs[0][0] - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), err
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("B,beam,V,step", [(1, 3, 5049, 2), (5, 5, 5049, 7), (80, 3, 700, 3), (3, 8, 2600, 1), (2, 3, 5049, 0)])
def test_ctc_prefix_full_vs_torch(L, B, beam, V, step):
    """Full-vocabulary CTC prefix scores (ctc_prefix_score.py:68-187 with scoring_ids=None) against the log-domain formula in
    torch, for every split plan: B=1 (16 time splits), B=5 (14), B=80 (no split: scores straight from registers), beam 8 (two
    passes over the block), step 0 (empty prefix)."""
    lib = L.load()
    g = torch.Generator().manual_seed(B * 100 + beam)
    lengths = [int(v) for v in torch.randint(20, 60, (B,), generator=g)]
    tmax, Fr = max(lengths), sum(lengths)
    ldp = (V + 3) // 4 * 4                                              # 16-byte aligned rows, padding filled with garbage
    logp_p = torch.full((Fr, ldp), float("nan"))
    logp_p[:, :V] = torch.log_softmax(torch.randn(Fr, V, generator=g) * 3.0, -1)
    logp_p = logp_p.cuda()
    logp = logp_p[:, :V]
    offs = np.concatenate([[0], np.cumsum(lengths)[:-1]]).astype(np.int32)
    R = B * beam
    n_run_l = [1 if step == 0 else 1 + (b % beam) for b in range(B)]
    if B > 2:
        n_run_l[1] = 0
    blank, eos = 0, V - 1
    r_buf = torch.full((2, R, tmax, 2), -1e10)
    r_buf[step & 1] = -torch.rand(R, tmax, 2, generator=g) * 40.0 - torch.arange(tmax).view(1, tmax, 1) * 3.0
    r_buf[step & 1, :, : max(0, step - 1)] = -1e10                      # a prefix of length L needs at least L frames
    r_buf = r_buf.cuda()
    last = torch.randint(1, V - 1, (R,), generator=g, dtype=torch.int32)
    if step == 0:
        last[:] = eos
    s_prev = (-torch.rand(R, generator=g) * 50).cuda() if step else torch.zeros(R).cuda()
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device="cuda")
    rprev = i32(list(range(R)))
    scores = torch.full((R, V), 123.0, device="cuda")
    ncg, ts = C.c_int(0), C.c_int(0)
    L.check(lib.avsr_ctc_prefix_full_plan(B, V, C.byref(ncg), C.byref(ts)), "plan")
    part = torch.empty(B, ts.value, beam, V, device="cuda")
    tick = torch.zeros(B, ncg.value, dtype=torch.int32, device="cuda")
    d_off, d_T, d_run, d_last, d_step = i32(offs), i32(lengths), i32(n_run_l), last.cuda(), i32([step])   # keep the buffers alive
    L.check(lib.avsr_ctc_prefix_full(L.ptr(logp_p), V, ldp, blank, eos, L.ptr(d_off), L.ptr(d_T), L.ptr(d_run), beam, B, 1,
                                     L.ptr(d_last), L.ptr(rprev), L.ptr(r_buf), tmax, L.ptr(d_step), L.ptr(s_prev),
                                     L.ptr(scores), L.ptr(part), L.ptr(tick), L.stream()), "ctc_full")
    torch.cuda.synchronize()
    assert int(tick.abs().sum().item()) == 0
    # the same with the posteriors precomputed once (what the CTC-only search does at every position): bit-identical
    probs = torch.empty_like(logp_p)
    scores_p = torch.full((R, V), 123.0, device="cuda")
    L.check(lib.avsr_ctc_exp_posteriors(L.ptr(logp_p), L.ll(logp_p.numel()), L.ptr(probs), L.stream()), "ctc_exp")
    L.check(lib.avsr_ctc_prefix_full_probs(L.ptr(logp_p), L.ptr(probs), V, ldp, blank, eos, L.ptr(d_off), L.ptr(d_T), L.ptr(d_run), beam, B, 1,
                                           L.ptr(d_last), L.ptr(rprev), L.ptr(r_buf), tmax, L.ptr(d_step), L.ptr(s_prev),
                                           L.ptr(scores_p), L.ptr(part), L.ptr(tick), L.stream()), "ctc_full_probs")
    torch.cuda.synchronize()
    assert int(tick.abs().sum().item()) == 0
    assert torch.equal(scores, scores_p)
    start = max(step, 1)
    for b in range(B):
        T = lengths[b]
        x = logp[offs[b]:offs[b] + T].double().cpu()
        for h in range(beam):
            row = b * beam + h
            if h >= n_run_l[b]:
                assert (scores[row] == 123.0).all()
                continue
            if step == 0:
                pn = torch.full((T,), -1e10, dtype=torch.float64)
                pb = torch.cumsum(x[:, blank], 0)
            else:
                pn, pb = r_buf[step & 1, row, :T, 0].double().cpu(), r_buf[step & 1, row, :T, 1].double().cpu()
            rs = torch.logaddexp(pn, pb)
            phi = rs.unsqueeze(1).repeat(1, V)
            phi[:, int(last[row])] = pb
            terms = phi[start - 1:T - 1] + x[start:T]
            x0 = x[0] if step == 0 else torch.full((V,), -1e10, dtype=torch.float64)
            want = torch.logsumexp(torch.cat([terms, x0.unsqueeze(0)], 0), 0)
            want[eos] = rs[T - 1]
            want[blank] = -1e10
            want = want - float(s_prev[row])
            got = scores[row].double().cpu()
            live = want > -1e9
            err = (got[live] - want[live]).abs().max().item()
            assert err < 2e-4, (b, h, err)
            assert (got[~live] < -1e9).all()


@pytest.mark.parametrize("R,N,K,mode", [(96, 1024, 1024, "ln"), (96, 3072, 1024, "relu"), (96, 1024, 3072, "ln"), (3, 3072, 1024, "plain"),
                                        (128, 1024, 1024, "ln")])
def test_gemm_x3_fused_epilogue(L, R, N, K, mode):
    """Projection + grid barrier + row-wise epilogue in one launch (decoder_layer.py:58-121 glue): bias / ReLU / residual /
    LayerNorm(eps 1e-12) / compact bf16x3 output, dead rows untouched; launched repeatedly (the barrier state is reused)."""
    from avsr_b200.weights import split3_weight_compact
    lib = L.load()
    a, w, bias = _rand(R, K, seed=11), _rand(N, K, seed=12, scale=0.03), _rand(N, seed=13)
    res, g, b = _rand(R, 1024, seed=14), _rand(N, seed=15), _rand(N, seed=16)
    a3, w3 = _split3c(a), split3_weight_compact(w)
    ns = lib.avsr_gemm_x3_splits(R, N, K)
    part = torch.empty(ns, R, N, device="cuda")
    gbar = torch.zeros(2, dtype=torch.int32, device="cuda")
    act = torch.ones(R, dtype=torch.int32, device="cuda")
    act[R // 2] = 0
    keep = act.bool()
    ref = (a.double() @ w.double().t()).float() + bias
    for it in range(3):
        out = torch.full((R, N), 5.0, device="cuda")
        sp = torch.zeros(R, 3 * N, dtype=torch.bfloat16, device="cuda")
        if mode == "ln":
            x = res.clone()
            L.check(lib.avsr_gemm_x3_fused(L.ptr(a3), L.ll(3 * K), L.ptr(w3), L.ll(3 * K), R, N, K, L.ptr(part), L.ptr(bias), 0, L.ptr(x),
                                           L.ll(1024), L.ptr(x), L.ll(N), L.ptr(g), L.ptr(b), C.c_float(1e-12), None, L.ll(1024), L.ptr(act),
                                           L.ptr(sp), L.ptr(gbar), L.stream()), "fused")
            want = ref + res
            assert (x[keep] - want[keep]).abs().max().item() < 1e-4
            assert torch.equal(x[~keep], res[~keep])
            lnw = F.layer_norm(want, (N,), g, b, 1e-12)
            back = sp.float().view(R, 3, N).sum(1)
            assert (back[keep] - lnw[keep]).abs().max().item() < 2e-4
        else:
            a_code = 2 if mode == "relu" else 0
            L.check(lib.avsr_gemm_x3_fused(L.ptr(a3), L.ll(3 * K), L.ptr(w3), L.ll(3 * K), R, N, K, L.ptr(part), L.ptr(bias), a_code, None,
                                           L.ll(1024), L.ptr(out), L.ll(N), None, None, C.c_float(1e-12), None, L.ll(1024), L.ptr(act),
                                           L.ptr(sp), L.ptr(gbar), L.stream()), "fused")
            want = torch.relu(ref) if mode == "relu" else ref
            assert (out[keep] - want[keep]).abs().max().item() < 1e-4
            assert (out[~keep] == 5.0).all()
            back = sp.float().view(R, 3, N).sum(1)
            assert (back[keep] - want[keep]).abs().max().item() < 1e-4
    assert int(gbar[0].item()) == 0 and int(gbar[1].item()) == 3


def _tile_stats(x):
    """(mean, M2) of every 128-column tile of the rows of x -> [K/128][R][2] (what avsr_dec_proj leaves in stats_out)."""
    R, K = x.shape
    t = x.double().view(R, K // 128, 128)
    mean = t.mean(-1)
    m2 = ((t - mean.unsqueeze(-1)) ** 2).sum(-1)
    return torch.stack([mean, m2], -1).permute(1, 0, 2).contiguous().float()


@pytest.mark.parametrize("R,N,K,mode", [(96, 1024, 1024, "tma_res_stats"), (96, 3072, 1024, "ln_bias"), (96, 3072, 1024, "ln_relu_split"),
                                        (96, 1024, 3072, "tma_res_stats"), (96, 5049, 1024, "ln_plain"), (96, 3072, 1024, "tma_bias"),
                                        (160, 1024, 1024, "tma_res_stats"), (160, 3072, 1024, "ln_bias"), (8, 1024, 1024, "ln_bias"),
                                        (3, 3072, 1024, "tma_bias"), (300, 1024, 3072, "tma_res_stats"), (96, 1024, 1024, "ln_bias")])
def test_dec_proj_cluster(L, R, N, K, mode):
    """avsr_dec_proj (csrc/gemm_x3c.cu): cluster split-K reduced through DSMEM, every operand / epilogue mode of a decode
    position, against float64: fp32-level accuracy, LayerNorm applied while staging, tile statistics, bf16x3 output rows,
    run-to-run bit identity."""
    from avsr_b200.weights import split3_weight_compact
    lib = L.load()
    ns = lib.avsr_dec_proj_splits(R, N, K)
    assert 1 <= ns <= min(16, K // 64)
    w = _rand(N, K, seed=2, scale=0.03)
    w3 = split3_weight_compact(w)
    bias = _rand(N, seed=3, scale=0.1)
    ln = mode.startswith("ln")
    if ln:
        x = _rand(R, K, seed=1, scale=2.0) + 0.5
        g, b = torch.rand(K, device="cuda") * 0.4 + 0.8, _rand(K, seed=5, scale=0.05)
        stats_in = _tile_stats(x).cuda()
        a = F.layer_norm(x.double(), (K,), g.double(), b.double(), 1e-12)
        a3 = None
    else:
        a32 = _rand(R, K, seed=1)
        a3 = _split3c(a32)
        a = a32.double()
    ref = a @ w.double().t()
    fp32_err = ((a.float() @ w.t()).double() - ref).abs().max().item()
    use_bias = mode != "ln_plain"
    if use_bias:
        ref = ref + bias.double()
    if "relu" in mode:
        ref = torch.relu(ref)
    res = None
    if "res" in mode:
        res = _rand(R, N, seed=7)
        ref = ref + res.double()
    ldo = N
    out = torch.full((R, ldo), float("nan"), device="cuda") if "split" not in mode else None
    if res is not None:
        out = res.clone()                                   # in place, like the residual stream of the decoder step
    split = torch.zeros(R, 3 * N, dtype=torch.bfloat16, device="cuda") if "split" in mode else None
    stats = torch.full((N // 128, R, 2), float("nan"), device="cuda") if "stats" in mode else None

    def run(o):
        L.check(lib.avsr_dec_proj(L.ptr(a3), L.ll(3 * K), L.ptr(x) if ln else None, L.ll(K), L.ptr(stats_in) if ln else None,
                                  L.ptr(g) if ln else None, L.ptr(b) if ln else None, C.c_float(1e-12), L.ptr(w3), L.ll(3 * K), R, N, K,
                                  L.ptr(bias) if use_bias else None, L.ACT_RELU if "relu" in mode else L.ACT_NONE,
                                  L.ptr(o) if res is not None else None, L.ll(N), L.ptr(o), L.ll(ldo), L.ptr(split), L.ptr(stats), None, L.ll(0), L.stream()),
                "avsr_dec_proj")
    run(out)
    torch.cuda.synchronize()
    tol = max(2e-5 * (K ** 0.5), 4 * fp32_err + 1e-6) * (4.0 if ln else 1.0)
    if out is not None:
        assert not torch.isnan(out).any()
        err = (out.double() - ref).abs().max().item()
        assert err < tol, (err, fp32_err)
        if res is not None:
            again = res.clone()
            run(again)
            torch.cuda.synchronize()
            assert torch.equal(again, out)                  # deterministic reduction order
    if split is not None:
        got = split.view(R, 3, N).double().sum(1)
        assert (got - ref).abs().max().item() < tol
    if stats is not None:
        want = _tile_stats(out)
        assert (stats.cpu() - want.cpu()).abs().max().item() < 1e-3 * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("R,N,mode", [(96, 1024, "plain"), (96, 3072, "relu_split"), (96, 5049, "nobias"), (160, 3072, "plain"), (5, 1024, "plain")])
def test_dec_proj_folded_layernorm(L, R, N, mode):
    """avsr_dec_proj_folded: LayerNorm folded into the projection (raw rows through TMA, mean / rstd applied to the finished
    sums) against float64 LayerNorm -> Linear, with a row mean that is NOT small against the spread."""
    from avsr_b200.weights import fold_layernorm
    lib = L.load()
    K = 1024
    x = _rand(R, K, seed=1, scale=2.0) + 1.5                 # |mean| ~ 0.75 sigma: the cancellation the folded form has to survive
    w = _rand(N, K, seed=2, scale=0.03)
    bias = None if mode == "nobias" else _rand(N, seed=3, scale=0.1)
    g, b = torch.rand(K, device="cuda") * 0.4 + 0.8, _rand(K, seed=5, scale=0.05)
    w3g, u, c = fold_layernorm(w, bias, g, b)
    x3 = _split3c(x)
    stats_in = _tile_stats(x).cuda()
    ref = F.layer_norm(x.double(), (K,), g.double(), b.double(), 1e-12) @ w.double().t()
    if bias is not None:
        ref = ref + bias.double()
    if "relu" in mode:
        ref = torch.relu(ref)
    fp32 = F.layer_norm(x, (K,), g, b, 1e-12) @ w.t() + (bias if bias is not None else 0)
    if "relu" in mode:
        fp32 = torch.relu(fp32)
    fp32_err = (fp32.double() - ref).abs().max().item()
    out = torch.full((R, N), float("nan"), device="cuda") if "split" not in mode else None
    split = torch.zeros(R, 3 * N, dtype=torch.bfloat16, device="cuda") if "split" in mode else None
    L.check(lib.avsr_dec_proj_folded(L.ptr(x3), L.ll(3 * K), L.ptr(stats_in), C.c_float(1e-12), L.ptr(u), L.ptr(c), L.ptr(w3g), L.ll(3 * K), R, N, K,
                                     L.ACT_RELU if "relu" in mode else L.ACT_NONE, None, L.ll(N), L.ptr(out), L.ll(N), L.ptr(split), None, None,
                                     L.ll(0), L.stream()), "avsr_dec_proj_folded")
    torch.cuda.synchronize()
    got = out.double() if out is not None else split.view(R, 3, N).double().sum(1)
    assert not torch.isnan(got).any()
    err = (got - ref).abs().max().item()
    assert err < max(4e-5 * (K ** 0.5), 6 * fp32_err + 1e-6), (err, fp32_err)


@pytest.mark.parametrize("lengths", [[4, 3], [1], [9, 1, 2]])
def test_frontend_conv3d_implicit_gemm(L, lengths):
    """avsr_frontend_conv3d (csrc/frontend_conv.cu): Conv3d 5x7x7 / stride 1x2x2 / pad 2x3x3 + bias + PReLU as an implicit GEMM
    against F.conv3d per utterance (the temporal halo must not reach into the neighbouring utterance of the packed batch)."""
    lib = L.load()
    Fr = sum(lengths)
    video = _rand(Fr, 88, 88, seed=1)
    w = _rand(64, 1, 5, 7, 7, seed=2, scale=0.05)
    bias, slope = _rand(64, seed=3, scale=0.2), torch.rand(64, device="cuda") * 0.3 + 0.1
    w8 = torch.zeros(64, 40, 8, device="cuda")
    w8[:, :35, :7] = w.reshape(64, 35, 7)
    w8 = w8.reshape(64, 320).bfloat16().contiguous()
    ft = torch.cat([torch.arange(t) for t in lengths]).int().cuda()
    fT = torch.cat([torch.full((t,), t) for t in lengths]).int().cuda()
    out = torch.full((Fr, 44, 44, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
    L.check(lib.avsr_frontend_conv3d(L.ptr(video), L.ptr(ft), L.ptr(fT), 0, Fr, L.ptr(w8), L.ptr(bias), L.ptr(slope), L.ptr(out), L.stream()),
            "avsr_frontend_conv3d")
    torch.cuda.synchronize()
    vb, wb = video.bfloat16().float(), w.bfloat16().float()
    o = 0
    for t in lengths:
        ref = F.conv3d(vb[o:o + t].view(1, 1, t, 88, 88), wb, None, stride=(1, 2, 2), padding=(2, 3, 3))[0]      # [64, t, 44, 44]
        ref = ref + bias.view(64, 1, 1, 1)
        ref = torch.where(ref >= 0, ref, ref * slope.view(64, 1, 1, 1)).permute(1, 2, 3, 0)                      # [t, 44, 44, 64]
        got = out[o:o + t].float()
        assert not torch.isnan(got).any()
        assert (got - ref).abs().max().item() < 0.02 * ref.abs().max().item() + 1e-2, (lengths, o)
        o += t
    # a second call on a sub-range of the frames writes the same values (f0 / nf chunking of the encoder)
    if Fr >= 3:
        sub = torch.zeros(2, 44, 44, 64, dtype=torch.bfloat16, device="cuda")
        L.check(lib.avsr_frontend_conv3d(L.ptr(video), L.ptr(ft), L.ptr(fT), 1, 2, L.ptr(w8), L.ptr(bias), L.ptr(slope), L.ptr(sub), L.stream()),
                "avsr_frontend_conv3d")
        assert torch.equal(sub, out[1:3])
