"""Dev aid: the encoder alone at the bench workload (B = 32 utterances of T = 375 frames): ms per pass, achieved TFLOP/s
against the algorithmic FLOPs of SURVEY.md 8(d) and the measured sustained bf16 peak.

    python tools/bench_encoder.py [B=32] [T=375] [passes=6]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from avsr_b200 import synth
from avsr_b200.encoder import Encoder

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 375
n = int(sys.argv[3]) if len(sys.argv) > 3 else 6
_pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
peak = json.load(open(_pk))["bf16_tflops_sustained"] if os.path.exists(_pk) else 1389.5
enc = Encoder(synth.make_state_dict(0), "cuda")
g = torch.Generator().manual_seed(B * 1000 + T)
video = torch.randn(B * T, 88, 88, generator=g).cuda()
audio = torch.randn(B, 104, T, generator=g).cuda()
for _ in range(2):
    enc.forward_packed(video, audio, [T] * B)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    enc.forward_packed(video, audio, [T] * B)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
flops = B * T * (1_268_871_168 + 98_304 * T)
tf = flops / (ms * 1e-3) / 1e12
print(f"encoder B={B} T={T}: {ms:.2f} ms per pass, {tf:.0f} TFLOP/s = {tf / peak:.3f} of the sustained bf16 peak ({peak:.0f})")
