"""Generates tests/golden/stitch.json from the UNMODIFIED reference script (/root/reference/script/evaluation.py):
``InferenceEngine.chunk_video`` without ASD (:247-270, torchaudio.load replaced by a stub that returns a clip of the wanted
duration) and ``InferenceEngine.format_vtt_timestamp`` (:272-278).  Modules the script imports that are absent from this
image (jiwer, webvtt, torchcodec, python_speech_features) are stubbed; none of them is on the two code paths run here.
Build container only."""
import json
import os
import sys
import types

import datasets  # noqa: F401  (before the stubs: its config probes torchcodec with find_spec)
import torch

sys.path.insert(0, "/root/reference")
for name in ["jiwer", "webvtt", "torchcodec", "torchcodec.decoders", "python_speech_features"]:
    sys.modules[name] = types.ModuleType(name)
sys.modules["jiwer"].wer = lambda **k: 0.0
sys.modules["torchcodec.decoders"].VideoDecoder = sys.modules["torchcodec.decoders"].AudioDecoder = object
sys.modules["python_speech_features"].logfbank = None
from script import evaluation as EV                       # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
chunks = []
for n_samples, max_len in [(16000 * 15, 15), (16000 * 15 + 1, 15), (16000 * 31, 15), (123457, 15), (16000 * 60, 10), (8000, 15),
                           (16000 * 44 + 5000, 15), (16000 * 100, 7), (1600, 15), (16000 * 29 + 15999, 10)]:
    EV.torchaudio.load = lambda path, n=n_samples: (torch.zeros(1, n), 16000)
    segs = EV.InferenceEngine.chunk_video(None, "clip.mp4", None, max_length=max_len)
    chunks.append({"duration": n_samples / 16000, "max_length": max_len, "segments": [list(s) for s in segs]})
stamps = [0, 0.001, 0.9996, 1.0005, 12.34, 59.9999, 60, 61.5, 3599.999, 3600, 3725.5, 7322.042, 86399.25, 100000.125]
out = {"chunks": chunks, "timestamps": [[t, EV.InferenceEngine.format_vtt_timestamp(None, t)] for t in stamps]}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "stitch.json"), "w"), indent=0)
print(len(chunks), "chunkings,", len(stamps), "timestamps")
