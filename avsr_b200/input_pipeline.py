"""GPU input pipeline: the step right before the hot path.

Mirrors the inference side of /root/reference/src/dataset/avhubert_dataset.py with the same names and argument meaning:

* ``cut_or_pad`` (:22-33), ``FBanksAndStack`` (:86-116), ``VideoTransform("test" | "val")`` (:225-246),
  ``AudioTransform("test" | "val")`` (:249-275; at test time the reference's ``AddNoise`` has no noise file and is the
  identity, :160-170), ``DataCollator`` (:314-349) for already decoded inputs (file decoding with torchcodec / cv2 stays
  on the host and is out of scope);
* the reference runs ``python_speech_features.logfbank`` in numpy on the CPU, one utterance at a time; here a batch is two
  kernel launches (``avsr_fbank_stack_ln``, ``avsr_video_u8_transform``, avsr_b200/csrc/input.cu) that write the collated,
  zero-padded ``audios [B,104,T]`` / ``videos [B,1,T,88,88]`` tensors the encoder takes, and the host uploads uint8 frames
  and raw samples (9.2 KB + 2.5 KB per video frame) instead of the 31 KB + 0.4 KB of float32 features.

The train-time augmentations (time masking, interferer / noise mixing, random crop) are not part of the inference path and
raise ``NotImplementedError``.  No CPU fallback: tensors are moved to the model's CUDA device and the library must load.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib as L

RATE_RATIO = 640
N_FEAT = 104
CROP = 88


def fbank_rows(n_samples: int) -> int:
    """Rows ``FBanksAndStack`` yields for a waveform of n samples: ceil(frames / 4), frames = 1 + ceil((n - 400) / 160)."""
    return int(L.load().avsr_fbank_rows(int(n_samples)))


def cut_or_pad(data: torch.Tensor, size: int, dim: int = 0) -> torch.Tensor:
    """Pads (zeros) or trims ``data`` along dim 0, like the reference (which only supports dim 0 as well)."""
    if dim != 0:
        raise ValueError("cut_or_pad works along dim 0")
    if data.size(0) < size:
        pad = data.new_zeros((size - data.size(0),) + tuple(data.shape[1:]))
        data = torch.cat([data, pad], 0)
    elif data.size(0) > size:
        data = data[:size]
    return data


def _i32(v, dev):
    return torch.tensor(list(v), dtype=torch.int32, device=dev)


def _i64(v, dev):
    return torch.tensor(list(v), dtype=torch.int64, device=dev)


def fbank_stack_ln_batch(waveforms: Sequence[torch.Tensor], n_samples: Optional[Sequence[int]] = None, device="cuda:0",
                         t_max: Optional[int] = None) -> Tuple[torch.Tensor, List[int]]:
    """Batch of waveforms ([n] or [n,1] float32, host or device) -> (audios [B,104,Tmax] fp32 on ``device``, rows per
    utterance).  ``n_samples[b]``: length utterance b is cut / zero-padded to first (default: its own length)."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("avsr_b200 input pipeline needs a CUDA device (no CPU path)")
    lib = L.load()
    flat = [w.reshape(-1) for w in waveforms]
    lens = [int(w.numel()) for w in flat]
    if not flat or min(lens) < 1:
        raise ValueError("every waveform needs at least one sample")
    n_samples = lens if n_samples is None else [int(n) for n in n_samples]
    if len(n_samples) != len(flat) or min(n_samples) < 1:
        raise ValueError("n_samples must hold one positive length per waveform")
    rows = [fbank_rows(n) for n in n_samples]
    tmax = max(rows) if t_max is None else int(t_max)
    if tmax < max(rows):
        raise ValueError(f"t_max={tmax} is smaller than the longest utterance ({max(rows)} rows)")
    offs, acc = [], 0
    for n in lens:
        offs.append(acc)
        acc += n
    # straight into one device buffer: pinned host tensors go by asynchronous DMA, pageable ones through torch's staging,
    # no host-side concatenation (device-resident inputs are concatenated on the device)
    if all(w.is_cuda for w in flat):
        wave = torch.cat([w.to(dev, torch.float32) for w in flat])
    else:
        wave = torch.empty(acc, dtype=torch.float32, device=dev)
        for w, o, n in zip(flat, offs, lens):
            wave[o:o + n].copy_(w, non_blocking=True)
    B = len(flat)
    out = torch.empty(B, N_FEAT, tmax, dtype=torch.float32, device=dev)
    d_off, d_len, d_n = _i64(offs, dev), _i32(lens, dev), _i32(n_samples, dev)      # named: they must outlive the launch call
    with torch.cuda.device(dev):
        L.check(lib.avsr_fbank_stack_ln(L.ptr(wave), L.ptr(d_off), L.ptr(d_len), L.ptr(d_n), B, tmax, L.ptr(out), L.stream()),
                "avsr_fbank_stack_ln")
    return out, rows


def video_transform_batch(videos: Sequence[torch.Tensor], device="cuda:0", t_max: Optional[int] = None) -> Tuple[torch.Tensor, List[int]]:
    """Batch of uint8 grey videos ([T,1,H,W] or [T,H,W], one frame size for the batch) -> (videos [B,1,Tmax,88,88] fp32 on
    ``device``, frames per utterance)."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("avsr_b200 input pipeline needs a CUDA device (no CPU path)")
    lib = L.load()
    vs = []
    for v in videos:
        if v.dtype != torch.uint8:
            raise ValueError("videos must be uint8 (the decoded grey frames, before x / 255)")
        if v.dim() == 4:
            if v.size(1) != 1:
                raise ValueError("videos must have one channel")
            v = v[:, 0]
        if v.dim() != 3 or v.size(0) < 1:
            raise ValueError("a video is [T,1,H,W] or [T,H,W] with T >= 1")
        vs.append(v)
    if not vs:
        raise ValueError("empty batch")
    H, W = int(vs[0].size(1)), int(vs[0].size(2))
    if any((int(v.size(1)), int(v.size(2))) != (H, W) for v in vs):
        raise ValueError("all videos of a batch must share one frame size")
    if H < CROP or W < CROP:
        raise ValueError(f"frames of {H}x{W} are smaller than the {CROP}x{CROP} crop")
    T = [int(v.size(0)) for v in vs]
    tmax = max(T) if t_max is None else int(t_max)
    if tmax < max(T):
        raise ValueError(f"t_max={tmax} is smaller than the longest utterance ({max(T)} frames)")
    offs, acc = [], 0
    for t in T:
        offs.append(acc)
        acc += t
    if all(v.is_cuda for v in vs):
        frames = torch.cat([v.to(dev).contiguous() for v in vs])
    else:
        frames = torch.empty(acc, H, W, dtype=torch.uint8, device=dev)
        for v, o, t in zip(vs, offs, T):
            frames[o:o + t].copy_(v, non_blocking=True)
    B = len(vs)
    out = torch.empty(B, 1, tmax, CROP, CROP, dtype=torch.float32, device=dev)
    d_off, d_T = _i64(offs, dev), _i32(T, dev)
    with torch.cuda.device(dev):
        L.check(lib.avsr_video_u8_transform(L.ptr(frames), L.ptr(d_off), L.ptr(d_T), B, tmax, H, W, L.ptr(out), L.stream()),
                "avsr_video_u8_transform")
    return out, T


def add_noise(waveform: torch.Tensor, noise: torch.Tensor, snr: torch.Tensor, lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``torchaudio.functional.add_noise`` for [B, L] (or [L]) CUDA tensors: the interferer / noise mixing of ``AddMultiSpk``
    and ``AddNoise`` (avhubert_dataset.py:160-222).  Same argument meaning; ``snr`` in dB per row."""
    squeeze = waveform.dim() == 1
    if squeeze:
        waveform, noise, snr = waveform[None], noise[None], snr.reshape(1)
        lengths = None if lengths is None else lengths.reshape(1)
    if waveform.dim() != 2 or noise.dim() != 2 or snr.dim() != 1 or (lengths is not None and lengths.dim() != 1):
        raise ValueError("Input leading dimensions don't match.")
    if waveform.size(-1) != noise.size(-1):
        raise ValueError(f"Length dimensions of waveform and noise don't match (got {waveform.size(-1)} and {noise.size(-1)}).")
    if waveform.shape != noise.shape or snr.numel() != waveform.size(0):
        raise ValueError("Input leading dimensions don't match.")
    if not waveform.is_cuda:
        raise RuntimeError("avsr_b200 input pipeline needs CUDA tensors (no CPU path)")
    dev = waveform.device
    w = waveform.to(torch.float32).contiguous()
    z = noise.to(dev, torch.float32).contiguous()
    s = snr.to(dev, torch.float32).contiguous()
    ln = None if lengths is None else lengths.to(dev, torch.int32).contiguous()
    B, n = w.shape
    out = torch.empty_like(w)
    energy = torch.empty(2 * B, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        L.check(L.load().avsr_add_noise(L.ptr(w), L.ptr(z), L.ptr(s), L.ptr(ln), B, L.ll(n), L.ptr(out), L.ptr(energy), L.stream()),
                "avsr_add_noise")
    return out[0] if squeeze else out


class FBanksAndStack(torch.nn.Module):
    """Per-utterance form, same call as the reference module: waveform [n,1] -> [rows,104] (on the GPU)."""

    def __init__(self, stack_order: int = 4, device="cuda:0"):
        super().__init__()
        if stack_order != 4:
            raise ValueError("the avsr_cocktail audio front end stacks 4 frames")
        self.device = device

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        out, rows = fbank_stack_ln_batch([x], device=self.device)
        return out[0].t()


class VideoTransform:
    def __init__(self, subset: str, device="cuda:0"):
        if subset not in ("val", "test"):
            raise NotImplementedError("only the inference transforms (val / test) are part of the B200 path")
        self.device = device

    def __call__(self, sample: torch.Tensor) -> torch.Tensor:
        """[T,1,H,W] uint8 -> [T,1,88,88] fp32."""
        out, _ = video_transform_batch([sample], device=self.device)
        return out[0].permute(1, 0, 2, 3)


class AudioTransform:
    def __init__(self, subset: str, speech_dataset=None, snr_target=None, device="cuda:0"):
        if subset not in ("val", "test"):
            raise NotImplementedError("only the inference transforms (val / test) are part of the B200 path")
        # the reference's test-time AddNoise(snr_target) loads no noise file and returns its input (:160-170)
        self.fbank = FBanksAndStack(device=device)

    def __call__(self, sample: torch.Tensor) -> torch.Tensor:
        return self.fbank(sample)


@dataclass
class DataCollator:
    """``DataCollator.__call__`` (:314-349) for decoded features: each feature holds ``"video"`` (uint8 [T,1,H,W]) and
    ``"audio"`` (float waveform [n,1]), optionally ``"label"`` (token ids).  Returns the reference's batch dict with the
    tensors on ``device``; lengths stay host tensors like the reference's."""
    text_transform: object = None
    video_transform: Optional[VideoTransform] = None
    audio_transform: Optional[AudioTransform] = None
    rate_ratio: int = RATE_RATIO
    device: str = "cuda:0"

    def __call__(self, features: List[Dict[str, torch.Tensor]]) -> Dict[str, torch.Tensor]:
        videos = [f["video"] for f in features]
        waves = [f["audio"] for f in features]
        vids, T = video_transform_batch(videos, device=self.device)
        auds, rows = fbank_stack_ln_batch(waves, n_samples=[t * self.rate_ratio for t in T], device=self.device)
        batch = {"videos": vids, "video_lengths": torch.tensor(T), "audios": auds, "audio_lengths": torch.tensor(rows)}
        if all("label" in f for f in features):
            labels = [torch.as_tensor(f["label"] if self.text_transform is None or not isinstance(f["label"], str)
                                      else self.text_transform.tokenize(f["label"])) for f in features]
            lmax = max(int(l.numel()) for l in labels)
            lab = torch.full((len(labels), lmax), -1, dtype=labels[0].dtype)
            for i, l in enumerate(labels):
                lab[i, :l.numel()] = l
            batch["labels"] = lab.unsqueeze(1)               # collate_pad: 1-D targets get a middle axis (:297-298)
            batch["label_lengths"] = torch.tensor([int(l.numel()) for l in labels])
        return batch
