"""Dev aid: the implicit-GEMM frontend conv (csrc/frontend_conv.cu) on one encoder chunk of 2048 frames, timed with CUDA events."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from avsr_b200 import _lib as L

lib = L.load()
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
video = torch.randn(nf, 88, 88, device="cuda")
w8 = (torch.randn(64, 320, device="cuda") * 0.05).bfloat16()
bias, slope = torch.randn(64, device="cuda"), torch.rand(64, device="cuda")
ft = (torch.arange(nf, device="cuda") % 375).int()
fT = torch.full((nf,), 375, dtype=torch.int32, device="cuda")
out = torch.empty(nf, 44, 44, 64, dtype=torch.bfloat16, device="cuda")


def run():
    L.check(lib.avsr_frontend_conv3d(L.ptr(video), L.ptr(ft), L.ptr(fT), 0, nf, L.ptr(w8), L.ptr(bias), L.ptr(slope), L.ptr(out), L.stream()), "front")


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
fl = 2.0 * nf * 1936 * 64 * 245
print(f"frontend conv, {nf} frames: {ms * 1e3:.1f} us per launch, {fl / ms / 1e9:.1f} TFLOP/s (245-tap flops), "
      f"{nf * 16 / 148 :.0f} tiles per SM, {ms * 1e3 / (nf * 16 / 148):.2f} us per tile")
