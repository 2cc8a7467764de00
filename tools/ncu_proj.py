"""Dev aid: a short PDL chain of cluster projections (the shapes of one decoder layer) for an ncu capture of dec_proj_kernel."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from avsr_b200 import _lib as L
from avsr_b200.weights import split3_weight_compact

lib = L.load()
R = 96
dev = "cuda"
x = torch.randn(R, 1024, device=dev)
stats = torch.zeros(8, R, 2, device=dev)
bias = torch.zeros(3072, device=dev)
a3 = {k: torch.randn(R, 3 * k, device=dev).bfloat16() for k in (1024, 3072)}
out = torch.zeros(R, 3072, device=dev)
W = {(n, k): split3_weight_compact(torch.randn(n, k, device=dev) * 0.03) for n, k in ((3072, 1024), (1024, 1024), (1024, 3072))}
for rep in range(3):
    for (N, K), res in (((3072, 1024), False), ((1024, 1024), True), ((1024, 3072), True)):
        L.check(lib.avsr_dec_proj(L.ptr(a3[K]), L.ll(3 * K), None, L.ll(0), None, None, None, C.c_float(1e-12), L.ptr(W[(N, K)]), L.ll(3 * K), R, N, K,
                                  L.ptr(bias), 0, L.ptr(x) if res else None, L.ll(1024), L.ptr(x) if res else L.ptr(out), L.ll(N), None,
                                  L.ptr(stats) if res else None, None, L.ll(0), L.stream()), "proj")
torch.cuda.synchronize()
print("done")
