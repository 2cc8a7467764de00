"""Dev aid: tcgen05 GEMM throughput at the encoder's shapes, several epilogues / tile widths."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avsr_b200 import _lib as L
L.load()
dev = "cuda"
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3
which = sys.argv[1] if len(sys.argv) > 1 else "all"
M = 12000
for (N, K) in ((4096, 1024), (1024, 4096), (2048, 1024), (1024, 1024)):
    a = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
    bias = torch.randn(N, device=dev); res = torch.randn(M, N, device=dev)
    o16 = torch.empty(M, N, device=dev, dtype=torch.bfloat16); o32 = torch.empty(M, N, device=dev)
    variants = {
        "plain->bf16": L.make_epilogue(out_bf16=o16, ld_bf16=N),
        "bias+gelu->bf16": L.make_epilogue(bias=bias, act=L.ACT_GELU, out_bf16=o16, ld_bf16=N),
        "bias+res->f32": L.make_epilogue(bias=bias, residual=res, ldr=N, out_f32=o32, ld_f32=N),
    }
    for name, ep in variants.items():
        if which != "all" and which not in name: continue
        for bn in (128, 256):
            t = timeit(lambda: L.gemm_bf16(a, w, M, N, K, ep, bn_hint=bn))
            print(f"M={M} N={N} K={K} bn={bn} {name:18s} {t*1e6:8.1f} us  {2.0*M*N*K/t/1e12:7.1f} TFLOP/s", flush=True)
    if which != "all": break
