"""Dev aid: in-situ cost of the kernel groups of ONE decode position (CUDA-graph replays, programmatic dependent launch
intact).  Decodes the bench batch up to a chosen position with the real path, then freezes the beam state (the
fusion / advance kernels are left out, so every replay recomputes the same position) and times the position with one kernel
group removed at a time.  The outputs of ablated runs are meaningless; only the times are used.

    python tools/ablate_step.py [position=187] [B=32] [T=375]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from avsr_b200 import synth
from avsr_b200.model import AVSRCocktailB200

pos = int(sys.argv[1]) if len(sys.argv) > 1 else 187
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
T = int(sys.argv[3]) if len(sys.argv) > 3 else 375
sd = synth.make_state_dict(0)
m = AVSRCocktailB200(sd, beam_size=3)
bs = m.beam_search
bs.n_groups = 1                       # one chain: this tool looks at a single session
x = torch.randn(B * T, 1024, device="cuda")
x = torch.nn.functional.layer_norm(x, (1024,))
bs.decode_batch(x, [T] * B, max_steps=pos)          # real decode up to `pos` (realistic caches / ancestry tables)
s = bs.last_session
torch.cuda.synchronize()
step_now = int(s['step'].item())
print(f"position {step_now}, live hyps per utterance {s['n_run'].tolist()[:8]}...")
anc = s["anc"][step_now & 1].view(B, 3, -1)[:, :, :step_now].cpu().numpy()
import numpy as np
one = np.array([[anc[b, 0, p] for p in range(step_now) if len(set(anc[b, :, p])) == 1] for b in range(B)], dtype=object)
allv = np.concatenate([np.asarray(o, dtype=np.int64) for o in one])
print("slot of the single surviving ancestor (share of positions):", [round(float((allv == k).mean()), 3) for k in range(3)])
runs = np.concatenate([np.diff(np.asarray(o, dtype=np.int64)) != 0 for o in one])
print(f"adjacent converged positions whose survivor sits in a different slot: {runs.mean():.3f}")
distinct = np.array([[len(set(anc[b, :, p])) for p in range(step_now)] for b in range(B)])
print(f"distinct history rows per utterance: mean {distinct.sum(1).mean() + 3:.0f} of {3 * (step_now + 1)} "
      f"(positions with 1 / 2 / 3 distinct ancestors: {[(distinct == k).mean().round(3) for k in (1, 2, 3)]})")


def time_step(skip, n=8, reps=6):
    bs._skip = frozenset(skip) | {"advance"}
    bs._step(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            bs._step(s)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (n * reps) * 1e3


full = time_step(set())
print(f"full position (without fusion/advance): {full:8.1f} us")
cluster = bs.proj == "cluster" and s["R"] <= 128
groups = [("self-attention x6", {"self"}), ("source attention x6", {"cross"}), ("both attentions", {"self", "cross"}),
          ("projections x37", {"gemm"}), ("softmax/top-S + CTC", {"tail"}), ("everything but projections", {"self", "cross", "tail"}),
          ("everything but attention", {"gemm", "tail"})]
if not cluster:
    groups += [("row epilogues x24", {"epi"}), ("projections + epilogues", {"gemm", "epi"})]
print("projection path:", "cluster (54 launches per position)" if cluster else "split-K + row epilogues (78 launches)")
for name, skip in groups:
    t = time_step(skip)
    print(f"  without {name:40s}: {t:8.1f} us   (group costs {full - t:7.1f} us in place)")
bs._skip = frozenset()
