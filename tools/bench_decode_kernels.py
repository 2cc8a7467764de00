"""Dev aid: CUDA-event timings of the individual decode-step kernels at a chosen position (default: mid-utterance)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from avsr_b200 import _lib as L
from avsr_b200.beam_search import BatchedBeamSearch
from avsr_b200.weights import split3_weight_compact

lib = L.load()
dev = "cuda"
if os.environ.get("AVSR_L2_GRAN"):
    from cuda import cudart
    torch.zeros(1, device=dev)
    print("cudaLimitMaxL2FetchGranularity before:", cudart.cudaDeviceGetLimit(cudart.cudaLimit.cudaLimitMaxL2FetchGranularity))
    print("set:", cudart.cudaDeviceSetLimit(cudart.cudaLimit.cudaLimitMaxL2FetchGranularity, int(os.environ["AVSR_L2_GRAN"])))
    print("after:", cudart.cudaDeviceGetLimit(cudart.cudaLimit.cudaLimitMaxL2FetchGranularity))
B, beam, T, V = 32, 3, 375, 5049
step = int(sys.argv[1]) if len(sys.argv) > 1 else 187
R, S, lmax, nl = B * beam, 4, T + 1, 6
i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)


def timeit(fn, n=24):
    """n launches captured in one CUDA graph (as in the real decode loop: no host launch overhead), replayed 5 times."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * n) * 1e3


n_run, utt_off, utt_T = i32([beam] * B), i32([b * T for b in range(B)]), i32([T] * B)
step_t = i32([step])
qkv = torch.randn(R, 3072, device=dev)
q2 = torch.randn(R, 1024, device=dev)
kc = torch.randn(nl, B, 16, 8, lmax * beam, 8, device=dev)        # keys transposed in 32-byte groups
vc = torch.randn(nl, B, 16, lmax * beam, 64, device=dev)
anc = torch.randint(0, beam, (2, R, lmax), dtype=torch.uint8, device=dev)
if len(sys.argv) > 2 and sys.argv[2] == 'shared':
    anc.zero_()          # all hyps of an utterance descend from slot 0 (converged beam): physical rows are shared
ckv = torch.randn(nl, 2, 16, B * T, 64, device=dev)
att = torch.empty(R, 1024, device=dev)
att6 = torch.empty(R, 3072, device=dev, dtype=torch.bfloat16)
li = {"i": 0}
# the attention kernels take q (and the current k, v) from the split-K partial sums of their projection, as in the real step
NSQ, NSC = lib.avsr_gemm_x3_splits(R, 3072, 1024), lib.avsr_gemm_x3_splits(R, 1024, 1024)
qkv_p, qkv_b = torch.randn(NSQ, R, 3072, device=dev), torch.randn(3072, device=dev)
q2_p, q2_b = torch.randn(NSC, R, 1024, device=dev), torch.randn(1024, device=dev)


def self_attn():
    l = li["i"] % nl; li["i"] += 1
    L.check(lib.avsr_dec_attn_step(0, L.ptr(qkv_p), L.ll(3072), NSQ, L.ptr(qkv_b), L.ptr(kc[l]), L.ptr(vc[l]), L.ptr(anc), lmax, L.ptr(n_run),
                                   L.ptr(utt_off), L.ptr(utt_T), beam, R, L.ptr(step_t), None, L.ll(0), L.ptr(att6), None, None, None, L.stream()), "self")


def cross_attn():
    l = li["i"] % nl; li["i"] += 1
    L.check(lib.avsr_dec_attn_step(1, L.ptr(q2_p), L.ll(1024), NSC, L.ptr(q2_b), L.ptr(ckv[l, 0]), L.ptr(ckv[l, 1]), None, lmax,
                                   L.ptr(n_run), L.ptr(utt_off), L.ptr(utt_T), beam, R, L.ptr(step_t), None, L.ll(B * T),
                                   L.ptr(att6), None, None, None, L.stream()), "cross")


print(f"step={step}  R={R}")
print(f"self-attn  : {timeit(self_attn):8.1f} us")
print(f"cross-attn : {timeit(cross_attn):8.1f} us   (K/V bytes per launch {B*T*2048*4/1e6:.0f} MB -> {B*T*2048*4/1e3/timeit(cross_attn):.0f} GB/s)")

# split-K tensor-core projections (weights cycled so they stream from HBM)
for (N, K) in ((3072, 1024), (1024, 1024), (1024, 3072), (V, 1024)):
    ws = [split3_weight_compact(torch.randn(N, K, device=dev) * 0.02) for _ in range(8)]
    a3 = torch.randn(R, 3 * K, device=dev).bfloat16()
    ns = lib.avsr_gemm_x3_splits(R, N, K)
    part = torch.empty(ns, R, N, device=dev)
    def g():
        w = ws[li["i"] % 8]; li["i"] += 1
        L.check(lib.avsr_gemm_x3_splitk(L.ptr(a3), L.ll(3 * K), L.ptr(w), L.ll(3 * K), R, N, K, L.ptr(part), L.stream()), "g")
    t = timeit(g)
    print(f"proj N={N} K={K} splits={ns}: {t:8.1f} us   ({N*K*6/1e3/t:.0f} GB/s of bf16x3 weights)")
    bias, res, gam, bet = (torch.randn(N, device=dev) for _ in range(2)) if False else (torch.randn(N, device=dev), torch.randn(R, 1024, device=dev), torch.randn(N, device=dev), torch.randn(N, device=dev))
    act = torch.ones(R, dtype=torch.int32, device=dev)
    if N == 1024:
        x = torch.randn(R, N, device=dev)
        a6o = torch.empty(R, 3 * N, device=dev, dtype=torch.bfloat16)
        def e():
            L.check(lib.avsr_splitk_epilogue(L.ptr(part), ns, R, N, L.ptr(bias), 0, L.ptr(x), L.ll(N), L.ptr(x), L.ll(N), L.ptr(gam), L.ptr(bet),
                                             C.c_float(1e-12), None, L.ll(N), L.ptr(act), L.ptr(a6o), L.stream()), "e")
        print(f"   epilogue(res+LN+split) ns={ns}: {timeit(e):8.1f} us")
    del ws

# CTC prefix (pre-beam) and fusion
logp = torch.log_softmax(torch.randn(B * T, V, device=dev), -1)
last, part_ids = i32([7] * R), torch.randint(1, V - 1, (R, S), dtype=torch.int32, device=dev)
r_buf = torch.full((2, R * S, T, 2), -30.0, device=dev)
rprev = i32([r * S for r in range(R)])
psi, rsum = torch.empty(R, S, device=dev), torch.empty(R, device=dev)
def ctc():
    L.check(lib.avsr_ctc_prefix_prebeam(L.ptr(logp), V, V, 0, L.ptr(utt_off), L.ptr(utt_T), L.ptr(n_run), beam, R, S, L.ptr(last), L.ptr(part_ids),
                                        L.ptr(rprev), L.ptr(r_buf), T, L.ptr(step_t), L.ptr(psi), L.ptr(rsum), L.stream()), "ctc")
print(f"ctc prebeam: {timeit(ctc):8.1f} us")

# floor of a dependent launch in the chain: a trivial kernel (one thread bumps the step counter) and a small row kernel
step2, nrun2, anyr = i32([0]), i32([beam] * B), i32([0])
def adv():
    L.check(lib.avsr_beam_step_advance(L.ptr(step2), L.ptr(nrun2), B, L.ptr(anyr), L.stream()), "adv")
print(f"launch floor (1-thread kernel in the PDL chain): {timeit(adv):8.2f} us")
emb, pe = torch.randn(V, 1024, device=dev), torch.randn(5000, 1024, device=dev)
g1, b1 = torch.ones(1024, device=dev), torch.zeros(1024, device=dev)
xx, a3 = torch.empty(R, 1024, device=dev), torch.empty(R, 3 * 1024, device=dev, dtype=torch.bfloat16)
lt = i32([5] * R)
def emb_ln():
    L.check(lib.avsr_dec_embed_ln(L.ptr(emb), L.ptr(pe), L.ptr(lt), L.ptr(n_run), beam, R, L.ptr(step_t), L.ptr(g1), L.ptr(b1),
                                  C.c_float(1e-12), L.ptr(xx), None, L.ptr(a3), L.stream()), "embed")
print(f"embed + LayerNorm row kernel (96 CTAs x 256 threads): {timeit(emb_ln):8.2f} us")
