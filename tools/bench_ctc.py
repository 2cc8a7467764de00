"""Dev aid: the full-vocabulary CTC prefix scoring kernel alone, at the bench workload's shape (same measurement as the
`ctc_prefix_full_vocab` roofline entry of bench.py: three posterior blocks used in turn, 12 launches per CUDA graph).

    python tools/bench_ctc.py [B=32] [T=375] [beam=3] [step=2]
"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from avsr_b200 import _lib as L

lib = L.load()
dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 375
nh = int(sys.argv[3]) if len(sys.argv) > 3 else 3
step = int(sys.argv[4]) if len(sys.argv) > 4 else 2
V = 5049
ldp = (V + 31) // 32 * 32
logps = []
for _ in range(3):
    lp = torch.zeros(B * T, ldp, device=dev)
    lp[:, :V] = torch.log_softmax(torch.randn(B * T, V, device=dev), -1)
    logps.append(lp)
i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)
utt_off, utt_T = i32([b * T for b in range(B)]), i32([T] * B)
n_run, last = i32([nh] * B), i32([7] * (B * nh))
r_buf = torch.full((2, B * nh, T, 2), -1e10, device=dev)
r_buf[..., 1] = -5.0
rprev, step_t = i32(list(range(B * nh))), i32([step])
s_prev, scores = torch.zeros(B * nh, device=dev), torch.empty(B * nh, V, device=dev)
ncg, ts = C.c_int(0), C.c_int(0)
L.check(lib.avsr_ctc_prefix_full_plan(B, V, C.byref(ncg), C.byref(ts)), "plan")
fpart = torch.empty(B, ts.value, nh, V, device=dev)
ftick = torch.zeros(B, ncg.value, dtype=torch.int32, device=dev)
use_probs = os.environ.get("AVSR_CTC_PROBS", "1") != "0"
probs = [torch.empty_like(lp) for lp in logps] if use_probs else [None] * 3
for lp, pr in zip(logps, probs):
    if pr is not None:
        L.check(lib.avsr_ctc_exp_posteriors(L.ptr(lp), L.ll(lp.numel()), L.ptr(pr), L.stream()), "exp")
cnt = {"i": 0}


def ctc_full():
    lp, pr = logps[cnt["i"] % 3], probs[cnt["i"] % 3]
    cnt["i"] += 1
    L.check(lib.avsr_ctc_prefix_full_probs(L.ptr(lp), L.ptr(pr), V, ldp, 0, V - 1, L.ptr(utt_off), L.ptr(utt_T), L.ptr(n_run), nh, B, 1, L.ptr(last),
                                     L.ptr(rprev), L.ptr(r_buf), T, L.ptr(step_t), L.ptr(s_prev), L.ptr(scores), L.ptr(fpart),
                                     L.ptr(ftick), L.stream()), "ctc_full")


for _ in range(3):
    ctc_full()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(12):
        ctc_full()
g.replay()
torch.cuda.synchronize()
best = None
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 48 * 1e-3
    best = t if best is None else min(best, t)
byts = B * (4.0 * T * V + 4.0 * nh * V + 16.0 * T * nh)
peak = 6551.0
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
print(f"ctc_prefix_full probs={int(use_probs)} B={B} T={T} hyps={nh} step={step}: {ncg.value} column groups x {ts.value} time splits, {best * 1e6:.1f} us per launch, "
      f"{byts / best / 1e9:.0f} GB/s = {byts / best / 1e9 / peak:.3f} of {peak:.0f}")
