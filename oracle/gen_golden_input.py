"""Generates tests/golden/input_pipeline.npz from the UNMODIFIED reference classes of
/root/reference/src/dataset/avhubert_dataset.py: ``FBanksAndStack`` (:86-116), ``VideoTransform('test')`` (:225-246),
``cut_or_pad`` (:22-33) and ``collate_pad`` (:277-312) + the permutes of ``DataCollator.__call__`` (:345-349).

Two modules that file imports are absent from this image and are injected before the import:
  * ``torchcodec`` (file decoding only; never called here) -> an empty stub;
  * ``python_speech_features`` (==0.6, requirements.txt:14) -> a module whose ``logfbank`` is oracle/input_oracle.py's
    restatement.  So the golden pins everything the reference does AROUND logfbank (stacking, zero rows, layer norm in
    torch, crop / normalise in torchvision, padding, layouts); the logfbank restatement itself stays unpinned.
Run in the build container only; the committed npz is what tests check the oracle and the CUDA path against."""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from oracle import input_oracle as O                     # noqa: E402

tc = types.ModuleType("torchcodec")
tcd = types.ModuleType("torchcodec.decoders")
tcd.VideoDecoder = tcd.AudioDecoder = object
tc.decoders = tcd
psf = types.ModuleType("python_speech_features")
psf.logfbank = lambda signal, samplerate=16000: O.logfbank_psf(signal, samplerate)
sys.modules.update({"torchcodec": tc, "torchcodec.decoders": tcd, "python_speech_features": psf})
from src.dataset import avhubert_dataset as R              # noqa: E402

rng = np.random.default_rng(20260118)
T_list = [9, 5, 12]                                       # video frames per utterance
H, W = 96, 96
videos = [rng.integers(0, 256, size=(t, 1, H, W), dtype=np.uint8) for t in T_list]
videos[1][0] = 0                                          # a black and a white frame: the ends of the value range
videos[1][1] = 255
waves = []
for i, t in enumerate(T_list):
    n = t * O.RATE_RATIO + (37, -211, 0)[i]               # longer, shorter and exactly rate_ratio * T
    tt = np.arange(n) / 16000.0
    w = 0.3 * np.sin(2 * np.pi * (180 + 90 * i) * tt) + 0.05 * rng.standard_normal(n)
    waves.append((w / np.abs(w).max()).astype(np.float32))
waves[1][300:1500] = 0.0                                  # a stretch of digital silence -> frames whose spectrum is all zero

vt, at = R.VideoTransform("test"), R.AudioTransform("test")
samples = []
for v, a in zip(videos, waves):
    video = torch.from_numpy(v)
    audio = R.cut_or_pad(torch.from_numpy(a)[:, None], len(video) * 640)
    samples.append({"video": vt(video), "audio": at(audio)})
batch = R.collate_pad(samples)
batch["videos"] = batch["videos"].permute(0, 2, 1, 3, 4)
batch["audios"] = batch["audios"].permute(0, 2, 1)

# odd-length waveforms straight through FBanksAndStack (frame counts 1, 2, not a multiple of 4, ...)
odd_lens = [1, 7, 400, 401, 560, 561, 1040, 3333]
odd = {}
fb = R.FBanksAndStack()
for n in odd_lens:
    w = (0.5 * rng.standard_normal(n)).astype(np.float32)
    odd[f"odd_wave_{n}"] = w
    odd[f"odd_feat_{n}"] = fb(torch.from_numpy(w)[:, None]).numpy() if n > 1 else np.zeros((0,), np.float32)
# (n == 1: x.squeeze() is 0-d and the reference itself fails inside psf; no golden for it)

# a non-96 frame size through the video transform
v2 = rng.integers(0, 256, size=(3, 1, 100, 120), dtype=np.uint8)
out = {
    "T_list": np.array(T_list), "videos_out": batch["videos"].numpy(), "audios_out": batch["audios"].numpy(),
    "video_lengths": batch["video_lengths"].numpy(), "audio_lengths": batch["audio_lengths"].numpy(),
    "video_100x120": v2, "video_100x120_out": vt(torch.from_numpy(v2)).numpy(), "odd_lens": np.array([n for n in odd_lens if n > 1]),
}
for i, (v, a) in enumerate(zip(videos, waves)):
    out[f"video_{i}"] = v
    out[f"wave_{i}"] = a
out.update({k: v for k, v in odd.items() if not k.endswith("_1")})
path = os.path.join(ROOT, "tests", "golden", "input_pipeline.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in out.items() if k.endswith("_out")})
