"""Dev aid: times the two input-pipeline kernels at the bench workload's shape (32 utterances x 15 s) with CUDA events,
device-resident inputs, and the oracle (numpy, one core) on one utterance beside them.

    python tools/bench_input_pipeline.py [--batch 32] [--frames 375] [--iters 20]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from avsr_b200 import input_pipeline as P

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--frames", type=int, default=375)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--no-oracle", action="store_true")
a = ap.parse_args()
B, T = a.batch, a.frames
rng = np.random.default_rng(0)
waves = [torch.from_numpy((0.2 * rng.standard_normal(T * 640)).astype(np.float32)).cuda() for _ in range(B)]
vids = [torch.from_numpy(rng.integers(0, 256, size=(T, 96, 96), dtype=np.uint8)).cuda() for _ in range(B)]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timed(fn):
    for _ in range(3):
        fn()
    ms = []
    for _ in range(a.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


# the wrappers also concatenate the batch and upload three small index vectors; time the whole call as a user sees it
t_a = timed(lambda: P.fbank_stack_ln_batch(waves))
t_v = timed(lambda: P.video_transform_batch(vids))
vid_bytes = B * T * (88 * 88 + 88 * 88 * 4)

# kernels alone: inputs already concatenated on the device, index vectors resident
from avsr_b200 import _lib as L
lib = L.load()
wave, frames = torch.cat(waves), torch.cat(vids)
i64 = lambda v: torch.tensor(v, dtype=torch.int64, device="cuda")
i32 = lambda v: torch.tensor(v, dtype=torch.int32, device="cuda")
w_off, w_len, f_off, f_T = i64([b * T * 640 for b in range(B)]), i32([T * 640] * B), i64([b * T for b in range(B)]), i32([T] * B)
aud = torch.empty(B, 104, T, device="cuda")
vid = torch.empty(B, 1, T, 88, 88, device="cuda")
k_a = timed(lambda: L.check(lib.avsr_fbank_stack_ln(L.ptr(wave), L.ptr(w_off), L.ptr(w_len), L.ptr(w_len), B, T, L.ptr(aud), L.stream()), "fbank"))
k_v = timed(lambda: L.check(lib.avsr_video_u8_transform(L.ptr(frames), L.ptr(f_off), L.ptr(f_T), B, T, 96, 96, L.ptr(vid), L.stream()), "video"))
print(f"kernels alone: fbank {1e3 * k_a:.1f} us, video {1e3 * k_v:.1f} us ({vid_bytes / (k_v * 1e-3) / 1e9:.0f} GB/s)")
print(f"fbank+stack+layernorm  B={B} T={T}: {t_a:.3f} ms per batch ({B * T * 640 / 16000 / (t_a * 1e-3):.0f} audio-s/s)")
print(f"video u8 -> fp32       B={B} T={T}: {t_v:.3f} ms per batch ({vid_bytes / (t_v * 1e-3) / 1e9:.0f} GB/s of crop-in + fp32-out bytes)")
if not a.no_oracle:
    from oracle import input_oracle as O
    w, v = waves[0].cpu().numpy(), vids[0].cpu().numpy()
    t0 = time.perf_counter()
    O.fbanks_and_stack(w)
    t1 = time.perf_counter()
    O.video_transform(v)
    t2 = time.perf_counter()
    print(f"oracle (numpy, 1 core), ONE utterance: fbank {1e3 * (t1 - t0):.1f} ms, video {1e3 * (t2 - t1):.1f} ms")
