"""Dev aid for ncu: a few launches of the kernels the roofline names, at the bench workload's shapes (no timing claims).

    python tools/profile_kernels.py [ctc] [attn] [gemm] [skinny] [proj] [front] [posconv] [halo] [encattn]
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from avsr_b200 import _lib as L
from avsr_b200.weights import split3_weight_compact

lib = L.load()
dev = "cuda"
which = set(sys.argv[1:]) or {"ctc", "attn", "gemm", "proj", "front", "posconv", "halo", "encattn", "dual", "ln"}
B, beam, T, V = 32, 3, 375, 5049
R = B * beam
i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)
flush = torch.empty(64 * 1024 * 1024, device=dev)
n_run, utt_off, utt_T = i32([beam] * B), i32([b * T for b in range(B)]), i32([T] * B)
step_t = i32([187])

if "ctc" in which:
    ldp = (V + 31) // 32 * 32
    logp = torch.zeros(B * T, ldp, device=dev)
    logp[:, :V] = torch.log_softmax(torch.randn(B * T, V, device=dev), -1)
    last = i32([7] * R)
    r_buf = torch.full((2, R, T, 2), -1e10, device=dev)
    r_buf[..., 1] = -5.0
    rprev, st2 = i32(list(range(R))), i32([2])
    s_prev, scores = torch.zeros(R, device=dev), torch.empty(R, V, device=dev)
    ncg, ts = C.c_int(0), C.c_int(0)
    L.check(lib.avsr_ctc_prefix_full_plan(B, V, C.byref(ncg), C.byref(ts)), "plan")
    fpart = torch.empty(B, ts.value, beam, V, device=dev)
    ftick = torch.zeros(B, ncg.value, dtype=torch.int32, device=dev)
    probs = torch.empty_like(logp)
    L.check(lib.avsr_ctc_exp_posteriors(L.ptr(logp), L.ll(logp.numel()), L.ptr(probs), L.stream()), "exp")
    for _ in range(3):
        flush.zero_()
        L.check(lib.avsr_ctc_prefix_full_probs(L.ptr(logp), L.ptr(probs), V, ldp, 0, V - 1, L.ptr(utt_off), L.ptr(utt_T), L.ptr(n_run), beam, B, 1, L.ptr(last),
                                         L.ptr(rprev), L.ptr(r_buf), T, L.ptr(st2), L.ptr(s_prev), L.ptr(scores), L.ptr(fpart),
                                         L.ptr(ftick), L.stream()), "ctc_full")
    torch.cuda.synchronize()

if "attn" in which:
    lmax, nl = T + 1, 2
    qkv, q2 = torch.randn(R, 3072, device=dev), torch.randn(R, 1024, device=dev)
    kc = torch.randn(nl, B, 16, 8, lmax * beam, 8, device=dev)
    vc = torch.randn(nl, B, 16, lmax * beam, 64, device=dev)
    anc = torch.zeros(2, R, lmax, dtype=torch.uint8, device=dev)
    ckv = torch.randn(nl, 2, 16, B * T, 64, device=dev)
    att6 = torch.empty(R, 3072, device=dev, dtype=torch.bfloat16)
    for it in range(3):
        l = it % nl
        flush.zero_()
        L.check(lib.avsr_dec_attn_step(0, L.ptr(qkv), L.ll(3072), 0, None, L.ptr(kc[l]), L.ptr(vc[l]), L.ptr(anc), lmax, L.ptr(n_run),
                                       L.ptr(utt_off), L.ptr(utt_T), beam, R, L.ptr(step_t), None, L.ll(0), L.ptr(att6), None, None, None, L.stream()), "self")
        L.check(lib.avsr_dec_attn_step(1, L.ptr(q2), L.ll(1024), 0, None, L.ptr(ckv[l, 0]), L.ptr(ckv[l, 1]), None, lmax, L.ptr(n_run),
                                       L.ptr(utt_off), L.ptr(utt_T), beam, R, L.ptr(step_t), None, L.ll(B * T), L.ptr(att6), None, None, None, L.stream()), "cross")
    torch.cuda.synchronize()

if "gemm" in which:
    M, N, K = B * T, 4096, 1024
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
    bias = torch.randn(N, device=dev)
    o16 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ep = L.make_epilogue(bias=bias, act=L.ACT_GELU, out_bf16=o16, ld_bf16=N)
    for _ in range(3):
        L.gemm_bf16(a, w, M, N, K, ep)                     # automatic: the CTA-pair kernel at this size
    for _ in range(2):
        L.gemm_bf16(a, w, M, N, K, ep, bn_hint=256)        # the single-CTA kernel, for comparison
    torch.cuda.synchronize()

if "skinny" in which:
    N, K = 3072, 1024
    ws = [split3_weight_compact(torch.randn(N, K, device=dev) * 0.02) for _ in range(3)]
    a3 = torch.randn(R, 3 * K, device=dev).bfloat16()
    ns = lib.avsr_gemm_x3_splits(R, N, K)
    part = torch.empty(ns, R, N, device=dev)
    for i in range(3):
        flush.zero_()
        L.check(lib.avsr_gemm_x3_splitk(L.ptr(a3), L.ll(3 * K), L.ptr(ws[i]), L.ll(3 * K), R, N, K, L.ptr(part), L.stream()), "g")
    torch.cuda.synchronize()
if "proj" in which:
    # the cluster projections of a decode position: q|k|v / w_1 shape, attention-output shape, w_2 shape (weights cycled: cold HBM)
    x = torch.randn(R, 1024, device=dev)
    stats = torch.zeros(8, R, 2, device=dev)
    bias = torch.zeros(3072, device=dev)
    a3 = {k: torch.randn(R, 3 * k, device=dev).bfloat16() for k in (1024, 3072)}
    out = torch.zeros(R, 3072, device=dev)
    for rep in range(2):
        for (N, K), res in (((1024, 1024), True), ((1024, 3072), True), ((3072, 1024), False)):
            w = split3_weight_compact(torch.randn(N, K, device=dev) * 0.03)
            flush.zero_()
            L.check(lib.avsr_dec_proj(L.ptr(a3[K]), L.ll(3 * K), None, L.ll(0), None, None, None, C.c_float(1e-12), L.ptr(w), L.ll(3 * K), R, N, K,
                                      L.ptr(bias), 0, L.ptr(x) if res else None, L.ll(1024), L.ptr(x) if res else L.ptr(out), L.ll(N), None,
                                      L.ptr(stats) if res else None, None, L.ll(0), L.stream()), "proj")
    torch.cuda.synchronize()

if "front" in which:
    nf = 2048
    video = torch.randn(nf, 88, 88, device=dev)
    w8 = (torch.randn(64, 320, device=dev) * 0.05).bfloat16()
    fb, fs = torch.randn(64, device=dev), torch.rand(64, device=dev)
    ft = (torch.arange(nf, device=dev) % 375).int()
    fT = torch.full((nf,), 375, dtype=torch.int32, device=dev)
    fo = torch.empty(nf, 44, 44, 64, dtype=torch.bfloat16, device=dev)
    for _ in range(2):
        L.check(lib.avsr_frontend_conv3d(L.ptr(video), L.ptr(ft), L.ptr(fT), 0, nf, L.ptr(w8), L.ptr(fb), L.ptr(fs), L.ptr(fo), L.stream()), "front")
    torch.cuda.synchronize()

if "posconv" in which:
    Fr = B * T
    hb = torch.randn(Fr, 1024, device=dev).bfloat16()
    h = torch.randn(Fr, 1024, device=dev)
    pw = (torch.randn(1024, 8192, device=dev) * 0.01).bfloat16()
    pb = torch.randn(1024, device=dev)
    offs = torch.arange(B, dtype=torch.int64) * T
    lens = torch.full((B,), T, dtype=torch.int32)
    host = torch.zeros(B, 128, dtype=torch.uint8)
    L.check(lib.avsr_posconv_encode_maps(L.ptr(hb), C.c_void_p(offs.data_ptr()), C.c_void_p(lens.data_ptr()), B, C.c_void_p(host.data_ptr())), "maps")
    maps = host.to(dev)
    work = [(b, b * T, T, q0) for b in range(B) for q0 in range(0, T, 128)]
    wu, wo_, wT, wq = (i32([x[j] for x in work]) for j in range(4))
    ep = L.make_epilogue(bias=pb, act=L.ACT_GELU, residual=h, ldr=1024, out_f32=h, ld_f32=1024)
    for _ in range(2):
        L.check(lib.avsr_posconv_bf16_tc(L.ptr(maps), L.ptr(pw), len(work), L.ptr(wu), L.ptr(wo_), L.ptr(wT), L.ptr(wq), C.byref(ep), L.stream()), "posconv")
    torch.cuda.synchronize()
if "halo" in which:
    nf, H = 2048, 22
    xp = torch.zeros(nf, H + 1, H + 2, 64, dtype=torch.bfloat16, device=dev)
    xp[:, :H, :H] = torch.randn(nf, H, H, 64, device=dev).bfloat16()
    w9 = (torch.randn(64, 576, device=dev) * 0.05).bfloat16()
    hb, hs = torch.randn(64, device=dev), torch.rand(64, device=dev)
    out = torch.zeros_like(xp)
    ep1 = L.make_epilogue(bias=hb, act=L.ACT_PRELU, prelu=hs, out_bf16=out.view(-1, 64), ld_bf16=64)
    ep2 = L.make_epilogue(bias=hb, act=L.ACT_PRELU, prelu=hs, residual=xp.view(-1, 64), ldr=64, act_after_residual=True, out_bf16=out.view(-1, 64), ld_bf16=64)
    for ep in (ep1, ep2, ep1, ep2):
        L.check(lib.avsr_conv3x3_halo_bf16(L.ptr(xp), L.ptr(w9), L.ll(nf), H, H, C.byref(ep), L.stream()), "halo")
    xd = torch.randn(nf, H, H, 64, device=dev).bfloat16()
    od = torch.empty(nf * H * H, 64, dtype=torch.bfloat16, device=dev)
    for _ in range(2):                                     # the generic implicit GEMM on the dense layout, for comparison
        L.conv2d_bf16(xd, w9, nf, H, H, 64, 64, 3, 1, L.make_epilogue(bias=hb, act=L.ACT_PRELU, prelu=hs, out_bf16=od, ld_bf16=64))
    torch.cuda.synchronize()

if "encattn" in which:
    Fr = B * T
    qk = torch.randn(Fr, 2048, device=dev).bfloat16()
    vt = torch.randn(1024, Fr + 8, device=dev).bfloat16()
    o = torch.empty(Fr, 1024, dtype=torch.bfloat16, device=dev)
    work = [(b * T, T, q0) for b in range(B) for q0 in range(0, T, 128)]
    wo2, wT2, wq2 = (i32([x[j] for x in work]) for j in range(3))
    for _ in range(2):
        L.check(lib.avsr_attention_varlen(L.ptr(qk), L.ptr(vt), L.ll(Fr + 8), L.ptr(o), L.ll(Fr), L.ptr(wo2), L.ptr(wT2), L.ptr(wq2), len(work), T,
                                          L.stream()), "attention")
    torch.cuda.synchronize()
if "dual" in which:
    # the two stacked projections of the query merge: q|k|v + tq (N = 4096, folded LayerNorm on the first 3072 columns) and
    # attention output + source query (N = 2048), weights cycled (cold HBM)
    x3 = torch.randn(R, 3 * 1024, device=dev).bfloat16()
    stats = torch.zeros(8, R, 2, device=dev)
    stats[..., 1] = 128.0
    xres, tq, q2, qkv = torch.randn(R, 1024, device=dev), torch.zeros(R, 1024, device=dev), torch.zeros(R, 1024, device=dev), torch.zeros(R, 3072, device=dev)
    u1, c1, b2 = torch.zeros(4096, device=dev), torch.zeros(4096, device=dev), torch.zeros(2048, device=dev)
    for rep in range(2):
        w1 = split3_weight_compact(torch.randn(4096, 1024, device=dev) * 0.03)
        w2 = split3_weight_compact(torch.randn(2048, 1024, device=dev) * 0.03)
        flush.zero_()
        L.check(lib.avsr_dec_proj_dual(L.ptr(x3), L.ll(3072), L.ptr(stats), C.c_float(1e-12), L.ptr(u1), L.ptr(c1), L.ptr(w1), L.ll(3072), R, 4096, 1024,
                                       3072, 0, None, L.ll(1024), L.ptr(qkv), L.ll(3072), None, L.ll(1024), L.ptr(tq), L.ll(1024), None, None, L.ll(0),
                                       L.stream()), "dual qkv")
        flush.zero_()
        L.check(lib.avsr_dec_proj_dual(L.ptr(x3), L.ll(3072), None, C.c_float(1e-12), None, L.ptr(b2), L.ptr(w2), L.ll(3072), R, 2048, 1024, 1024, 0,
                                       L.ptr(xres), L.ll(1024), L.ptr(xres), L.ll(1024), L.ptr(tq), L.ll(1024), L.ptr(q2), L.ll(1024), L.ptr(stats),
                                       None, L.ll(0), L.stream()), "dual out")
    torch.cuda.synchronize()

if "ln" in which:
    hx = torch.randn(B * T, 1024, device=dev)
    lg, lb = torch.ones(1024, device=dev), torch.zeros(1024, device=dev)
    lo = torch.empty(B * T, 1024, dtype=torch.bfloat16, device=dev)
    for _ in range(2):
        flush.zero_()
        L.layernorm(hx, lg, lb, 1e-5, out_bf16=lo)
    torch.cuda.synchronize()
print("ok")
