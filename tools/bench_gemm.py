"""Dev aid: the encoder's transformer GEMM shapes at the bench workload (M = 32 x 375 frames), CTA-pair kernel against the
single-CTA kernel (bn_hint 256 forces the latter).

    python tools/bench_gemm.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from avsr_b200 import _lib as L

L.load()
dev = "cuda"
M = 12000
for name, N, K, act in (("qkv", 3072, 1024, L.ACT_NONE), ("out", 1024, 1024, L.ACT_NONE), ("ffn1", 4096, 1024, L.ACT_GELU), ("ffn2", 1024, 4096, L.ACT_NONE)):
    a = torch.randn(M, K, device=dev).bfloat16()
    ws = [(torch.randn(N, K, device=dev) * 0.02).bfloat16() for _ in range(4)]
    bias = torch.randn(N, device=dev)
    o16 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ep = L.make_epilogue(bias=bias, act=act, out_bf16=o16, ld_bf16=N)
    for label, bn in (("pair", 0), ("single", 256)):
        i = [0]

        def run():
            L.gemm_bf16(a, ws[i[0] % 4], M, N, K, ep, bn_hint=bn)
            i[0] += 1
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                run()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 100 * 1e-3
        print(f"{name:5s} [{M},{K}]x[{N},{K}]^T {label:6s}: {t * 1e6:7.1f} us  {2.0 * M * N * K / t / 1e12:7.1f} TFLOP/s")
