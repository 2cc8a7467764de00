"""BASELINE.json configs[4]: encoder-only sweep (sequence length 2-30 s x batch 1-64) and the full-vocabulary CTC prefix
scoring sweep.  For every point: encoder ms per pass, achieved TFLOP/s against the algorithmic FLOPs of SURVEY.md 8(d)
(T * (1 268 871 168 + 98 304 T) per utterance) and its fraction of the measured sustained bf16 peak; for the CTC kernel the
algorithmic bytes 4 T V + 4 n_h V + 16 T n_h per utterance-step against the measured HBM peak.  Prints JSON lines.

    python tools/sweep_cfg5.py [--quick]
"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from avsr_b200 import _lib as L
from avsr_b200 import synth
from avsr_b200.encoder import Encoder

quick = "--quick" in sys.argv
_pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
peaks = json.load(open(_pk)) if os.path.exists(_pk) else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
lib = L.load()
dev = "cuda"
enc = Encoder(synth.make_state_dict(0), dev)


def timeit(fn, n):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


Ts = (50, 375) if quick else (50, 125, 250, 375, 500, 750)
Bs = (1, 32) if quick else (1, 2, 4, 8, 16, 32, 64)
for T in Ts:
    for B in Bs:
        if B * T > 26000:
            continue
        g = torch.Generator().manual_seed(B * 1000 + T)
        video = torch.randn(B * T, 88, 88, generator=g).cuda()
        audio = torch.randn(B, 104, T, generator=g).cuda()
        ms = timeit(lambda: enc.forward_packed(video, audio, [T] * B), 3 if B * T > 4000 else 6)
        flops = B * T * (1_268_871_168 + 98_304 * T)
        tf = flops / (ms * 1e-3) / 1e12
        print(json.dumps({"sweep": "encoder", "T": T, "B": B, "ms": round(ms, 3), "tflops": round(tf, 1),
                          "frac_of_sustained_bf16_peak": round(tf / peaks["bf16_tflops_sustained"], 3),
                          "audio_s_per_s": round(B * T / 25.0 / (ms * 1e-3), 1)}), flush=True)
        del video, audio

V = 5049
ldp = (V + 31) // 32 * 32
i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)
for T in Ts:
    for nh in (3, 5):
        for B in ((1, 32) if quick else (1, 8, 32, 64)):
            logps = []
            for _ in range(3 if B * T * ldp * 4 < 150e6 else 2):
                lp = torch.zeros(B * T, ldp, device=dev)
                lp[:, :V] = torch.log_softmax(torch.randn(B * T, V, device=dev), -1)
                logps.append(lp)
            utt_off, utt_T = i32([b * T for b in range(B)]), i32([T] * B)
            n_run, last = i32([nh] * B), i32([7] * (B * nh))
            r_buf = torch.full((2, B * nh, T, 2), -1e10, device=dev)
            r_buf[..., 1] = -5.0
            rprev, step_t = i32(list(range(B * nh))), i32([2])
            s_prev, scores = torch.zeros(B * nh, device=dev), torch.empty(B * nh, V, device=dev)
            ncg, ts = C.c_int(0), C.c_int(0)
            L.check(lib.avsr_ctc_prefix_full_plan(B, V, C.byref(ncg), C.byref(ts)), "plan")
            fpart = torch.empty(B, ts.value, nh, V, device=dev)
            ftick = torch.zeros(B, ncg.value, dtype=torch.int32, device=dev)
            k = {"i": 0}

            def run():
                lp = logps[k["i"] % len(logps)]
                k["i"] += 1
                L.check(lib.avsr_ctc_prefix_full(L.ptr(lp), V, ldp, 0, V - 1, L.ptr(utt_off), L.ptr(utt_T), L.ptr(n_run), nh, B, 1, L.ptr(last),
                                                 L.ptr(rprev), L.ptr(r_buf), T, L.ptr(step_t), L.ptr(s_prev), L.ptr(scores), L.ptr(fpart),
                                                 L.ptr(ftick), L.stream()), "ctc_full")
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for _ in range(12):
                    run()
            ms = timeit(gr.replay, 4) / 12
            byts = B * (4.0 * T * V + 4.0 * nh * V + 16.0 * T * nh)
            gbs = byts / (ms * 1e-3) / 1e9
            print(json.dumps({"sweep": "ctc_prefix_full", "T": T, "n_h": nh, "B": B, "us": round(ms * 1e3, 2), "gbs": round(gbs, 1),
                              "frac_of_hbm_peak": round(gbs / peaks["hbm_gbs"], 3), "column_groups": ncg.value, "time_splits": ts.value}),
                  flush=True)
            del logps
