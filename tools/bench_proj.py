"""Dev aid: the decoder-step projections in isolation (CUDA graph of PDL-chained launches, weights cycling over the six layers
so that they stream from HBM): round 1's split-K + row-epilogue pair against the cluster kernel at forced cluster sizes."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from avsr_b200 import _lib as L
from avsr_b200.weights import split3_weight_compact

lib = L.load()
dev = "cuda"
R = int(sys.argv[1]) if len(sys.argv) > 1 else 96


def timeit(fn, n=24, reps=5):
    for _ in range(2):
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
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * n) * 1e3


NW = 16          # matrices per shape: 16 x 19 MB > the 126 MB L2, so every launch streams from HBM unless fetched ahead


def weights(N, K, n=NW):
    return [split3_weight_compact(torch.randn(N, K, device=dev) * 0.03) for _ in range(n)]


x = torch.randn(R, 1024, device=dev)
stats = torch.zeros(8, R, 2, device=dev)
stats[:, :, 1] = 128.0
g1, b1 = torch.ones(1024, device=dev), torch.zeros(1024, device=dev)
bias = {n: torch.randn(n, device=dev) * 0.1 for n in (1024, 3072, 5052)}
a3 = {k: torch.randn(R, 3 * k, device=dev).bfloat16() for k in (1024, 3072)}
out = {n: torch.zeros(R, n, device=dev) for n in (1024, 3072, 5049)}
split_o = torch.zeros(R, 3 * 3072, dtype=torch.bfloat16, device=dev)
W = {(3072, 1024): weights(3072, 1024), (1024, 1024): weights(1024, 1024), (1024, 3072): weights(1024, 3072)}
cnt = {"i": 0}


PF = {"on": False, "w": None}


def new(N, K, mode, nxt=None):
    w = W[(N, K)][cnt["i"] % NW]
    PF["w"] = W[(N, K)][(cnt["i"] + 1) % NW] if nxt is None else W[nxt][(cnt["i"] + 1) % NW]      # the weights of the next launch of the chain
    cnt["i"] += 1
    ln = mode.startswith("ln")
    L.check(lib.avsr_dec_proj(None if ln else L.ptr(a3[K]), L.ll(3 * K), L.ptr(x) if ln else None, L.ll(1024), L.ptr(stats) if ln else None,
                              L.ptr(g1) if ln else None, L.ptr(b1) if ln else None, C.c_float(1e-12), L.ptr(w), L.ll(3 * K), R, N, K,
                              L.ptr(bias[N]), L.ACT_RELU if "relu" in mode else 0, L.ptr(x) if "res" in mode else None, L.ll(1024),
                              L.ptr(x) if "res" in mode else (None if "split" in mode else L.ptr(out[N])), L.ll(N),
                              L.ptr(split_o) if "split" in mode else None, L.ptr(stats) if "res" in mode else None,
                              L.ptr(PF["w"]) if PF["on"] else None, L.ll(PF["w"].numel() * 2 if PF["on"] else 0), L.stream()), "proj")


part = torch.empty(16 * R * 3072, device=dev)
a3o = torch.zeros(R, 3 * 3072, dtype=torch.bfloat16, device=dev)


def old(N, K, epi=True):
    w = W[(N, K)][cnt["i"] % NW]
    cnt["i"] += 1
    ns = lib.avsr_gemm_x3_splits(R, N, K)
    L.check(lib.avsr_gemm_x3_splitk(L.ptr(a3[K]), L.ll(3 * K), L.ptr(w), L.ll(3 * K), R, N, K, L.ptr(part), L.stream()), "x3")
    if epi:
        ln = N == 1024
        L.check(lib.avsr_splitk_epilogue_pf(L.ptr(part), ns, R, N, L.ptr(bias[N]), 0, L.ptr(x) if ln else None, L.ll(1024), L.ptr(x) if ln else None,
                                            L.ll(N), L.ptr(g1) if ln else None, L.ptr(b1) if ln else None, C.c_float(1e-12), None, L.ll(1024), None,
                                            L.ptr(a3o), None, L.ll(0), L.stream()), "epi")


print(f"R={R}; max clusters by size:", {cs: lib.avsr_dec_proj_max_clusters(cs, 96) for cs in (2, 4, 5, 6, 8, 10, 12, 16)})
for N, K in ((1024, 1024), (3072, 1024), (1024, 3072)):
    print(f"--- N={N} K={K}: old split-K {lib.avsr_gemm_x3_splits(R, N, K)} splits: proj only {timeit(lambda: old(N, K, False)):6.2f} us, "
          f"proj + row epilogue {timeit(lambda: old(N, K, True)):6.2f} us per pair")
    for s in (4, 5, 8, 10):
        if s > K // 64 or (N // 128) * s > 148 * 2:
            continue
        lib.avsr_dec_proj_force_splits(s)
        modes = ("tma_plain", "tma_res", "ln_plain") if N == 1024 else ("tma_plain", "ln_plain", "ln_relu_split")
        if K == 3072:
            modes = ("tma_plain", "tma_res")
        PF["on"] = False
        line = f"    cluster {s:2d}: " + "  ".join(f"{m} {timeit(lambda: new(N, K, m)):6.2f} us" for m in modes)
        PF["on"] = True
        print(line + "   | with L2 fetch-ahead: " + "  ".join(f"{timeit(lambda: new(N, K, m)):6.2f}" for m in modes))
        PF["on"] = False
lib.avsr_dec_proj_force_splits(0)


def layer_new():
    seq = [(3072, 1024, "ln_plain"), (1024, 1024, "tma_res"), (1024, 1024, "ln_plain"), (1024, 1024, "tma_res"), (3072, 1024, "ln_relu_split"),
           (1024, 3072, "tma_res")]
    for i, (N, K, m) in enumerate(seq):
        nN, nK, _ = seq[(i + 1) % 6]
        new(N, K, m, nxt=(nN, nK))


def layer_old():
    old(3072, 1024, False); old(1024, 1024); old(1024, 1024, False); old(1024, 1024); old(3072, 1024); old(1024, 3072)


t_new = timeit(layer_new, n=6)
PF["on"] = True
t_pf = timeit(layer_new, n=6)
print(f"six projections of a layer back to back: cluster {t_new:7.2f} us, with L2 fetch-ahead {t_pf:7.2f} us, split-K + epilogues {timeit(layer_old, n=6):7.2f} us")
