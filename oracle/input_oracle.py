"""CPU oracle of the input pipeline that sits right before the hot path.  TEST INFRASTRUCTURE ONLY: imported
by tests/, oracle/gen_golden_input.py and nothing in the product path.

What it restates
  * /root/reference/src/dataset/avhubert_dataset.py:86-116  ``FBanksAndStack``: ``logfbank(x, samplerate=16000)`` ->
    zero-pad the frame count to a multiple of 4, stack 4 frames -> ``F.layer_norm`` over the 104 stacked features;
  * ``:225-246`` ``VideoTransform("test")``: ``x / 255.0`` -> ``CenterCrop(88)`` -> ``Normalize(0.421, 0.165)``;
  * ``:22-33`` ``cut_or_pad`` and ``:277-312`` ``pad`` / ``collate_pad`` + the two permutes of ``DataCollator.__call__``
    (``:345-349``): videos [B,1,Tmax,88,88], audios [B,104,Tmax], zero padded.

``logfbank`` lives in a third-party dependency that is NOT in /root/reference and not installed here:
python_speech_features==0.6 (requirements.txt:14).  ``logfbank_psf`` restates its published algorithm (base.py ``fbank`` /
``logfbank`` / ``get_filterbanks`` / ``hz2mel`` / ``mel2hz``; sigproc.py ``preemphasis`` / ``framesig`` / ``powspec``) with
the defaults the reference call uses: winlen 25 ms, winstep 10 ms, nfilt 26, nfft 512, lowfreq 0, highfreq 8000,
preemph 0.97, rectangular window.  PARITY UNPINNED for that function: the reference holds no golden vectors for it and
the package cannot be run here.  Everything around it IS pinned: oracle/gen_golden_input.py runs the reference's own
``FBanksAndStack`` / ``VideoTransform`` / ``collate_pad`` (unmodified, with this ``logfbank_psf`` injected as the missing
module) and tests/test_input_oracle.py checks this file against those outputs.
"""
import math

import numpy as np

SAMPLE_RATE = 16000
FRAME_LEN = 400          # round_half_up(0.025 * 16000)
FRAME_STEP = 160         # round_half_up(0.010 * 16000)
NFFT = 512
NFILT = 26
STACK = 4
PREEMPH = 0.97
CROP = 88
VIDEO_MEAN, VIDEO_STD = 0.421, 0.165
RATE_RATIO = 640         # audio samples per video frame (DataCollator.rate_ratio, avhubert_dataset.py:321)


def hz2mel(hz):
    return 2595 * np.log10(1 + hz / 700.0)


def mel2hz(mel):
    return 700 * (10 ** (mel / 2595.0) - 1)


def filterbank_bins(nfilt=NFILT, nfft=NFFT, samplerate=SAMPLE_RATE, lowfreq=0, highfreq=None):
    """FFT-bin edges of the triangular mel filters (psf base.get_filterbanks): nfilt + 2 integers (as float64)."""
    highfreq = highfreq or samplerate / 2
    melpoints = np.linspace(hz2mel(lowfreq), hz2mel(highfreq), nfilt + 2)
    return np.floor((nfft + 1) * mel2hz(melpoints) / samplerate)


def get_filterbanks(nfilt=NFILT, nfft=NFFT, samplerate=SAMPLE_RATE, lowfreq=0, highfreq=None):
    b = filterbank_bins(nfilt, nfft, samplerate, lowfreq, highfreq)
    fb = np.zeros([nfilt, nfft // 2 + 1])
    for j in range(nfilt):
        for i in range(int(b[j]), int(b[j + 1])):
            fb[j, i] = (i - b[j]) / (b[j + 1] - b[j])
        for i in range(int(b[j + 1]), int(b[j + 2])):
            fb[j, i] = (b[j + 2] - i) / (b[j + 2] - b[j + 1])
    return fb


def num_frames(n_samples):
    """psf sigproc.framesig: one frame if the signal fits in it, else 1 + ceil((n - frame_len) / frame_step)."""
    if n_samples <= FRAME_LEN:
        return 1
    return 1 + int(math.ceil((1.0 * n_samples - FRAME_LEN) / FRAME_STEP))


def logfbank_psf(signal, samplerate=SAMPLE_RATE):
    """python_speech_features 0.6 ``logfbank`` with its defaults.  ``signal`` 1-D; returns [frames, 26] float64.
    Pre-emphasis runs in the signal's own dtype (float32 in the reference call, avhubert_dataset.py:110); the zero padding
    of framesig promotes to float64 and everything after it is float64."""
    assert samplerate == SAMPLE_RATE
    signal = np.asarray(signal)
    if signal.ndim != 1 or signal.shape[0] < 1:
        raise ValueError("logfbank_psf wants a non-empty 1-D signal")
    pre = np.append(signal[0], signal[1:] - PREEMPH * signal[:-1])
    n = len(pre)
    nf = num_frames(n)
    padlen = (nf - 1) * FRAME_STEP + FRAME_LEN
    padded = np.concatenate((pre, np.zeros((padlen - n,))))
    idx = np.arange(FRAME_LEN)[None, :] + (np.arange(nf) * FRAME_STEP)[:, None]
    frames = padded[idx] * np.ones((FRAME_LEN,))
    pspec = 1.0 / NFFT * np.square(np.absolute(np.fft.rfft(frames, NFFT)))
    feat = np.dot(pspec, get_filterbanks().T)
    feat = np.where(feat == 0, np.finfo(float).eps, feat)
    return np.log(feat)


def stacker(feats, stack_order=STACK):
    """avhubert_dataset.py:91-105: zero rows up to a multiple of stack_order, then [T/4, 4 * F]."""
    dim = feats.shape[1]
    if len(feats) % stack_order != 0:
        res = stack_order - len(feats) % stack_order
        feats = np.concatenate([feats, np.zeros([res, dim]).astype(feats.dtype)], axis=0)
    return feats.reshape((-1, stack_order, dim)).reshape(-1, stack_order * dim)


def layer_norm_rows(x, eps=1e-5):
    """F.layer_norm(x, x.shape[1:]) without affine (avhubert_dataset.py:113-114); float64 accumulation, float32 result."""
    x64 = x.astype(np.float64)
    mean = x64.mean(axis=1, keepdims=True)
    var = ((x64 - mean) ** 2).mean(axis=1, keepdims=True)
    return ((x64 - mean) / np.sqrt(var + eps)).astype(np.float32)


def fbanks_and_stack(waveform):
    """``FBanksAndStack.forward``: waveform [n] or [n,1] float32 -> [ceil(frames / 4), 104] float32."""
    x = np.asarray(waveform, dtype=np.float32).reshape(-1)
    feats = logfbank_psf(x).astype(np.float32)
    return layer_norm_rows(stacker(feats))


def center_crop_offsets(h, w, crop=CROP):
    """torchvision.transforms.functional.center_crop: int(round((h - crop) / 2.0)) (Python's banker's round)."""
    if h < crop or w < crop:
        raise ValueError("frames smaller than the crop are not supported")
    return int(round((h - crop) / 2.0)), int(round((w - crop) / 2.0))


def video_transform(video_u8):
    """``VideoTransform('test')``: [T,1,H,W] (or [T,H,W]) uint8 -> [T,1,88,88] float32, all arithmetic in float32."""
    v = np.asarray(video_u8)
    if v.ndim == 3:
        v = v[:, None]
    top, left = center_crop_offsets(v.shape[2], v.shape[3])
    x = v.astype(np.float32) / np.float32(255.0)
    x = x[:, :, top:top + CROP, left:left + CROP]
    return ((x - np.float32(VIDEO_MEAN)) / np.float32(VIDEO_STD)).astype(np.float32)


def cut_or_pad(wave, size):
    """avhubert_dataset.py:22-33 on a 1-D waveform."""
    wave = np.asarray(wave, dtype=np.float32).reshape(-1)
    if len(wave) < size:
        return np.concatenate([wave, np.zeros(size - len(wave), dtype=np.float32)])
    return wave[:size]


def collate(videos_u8, waveforms):
    """``DataCollator.__call__`` for decoded inputs: per utterance cut_or_pad(audio, T * 640), the two transforms,
    zero padding to the longest, permutes.  Returns (videos [B,1,Tmax,88,88], audios [B,104,Tmax], video_lengths,
    audio_lengths)."""
    vids, auds = [], []
    for v, a in zip(videos_u8, waveforms):
        vids.append(video_transform(v))
        auds.append(fbanks_and_stack(cut_or_pad(a, len(v) * RATE_RATIO)))
    vl, al = [len(v) for v in vids], [len(a) for a in auds]
    B, tv, ta = len(vids), max(vl), max(al)
    videos = np.zeros((B, tv, 1, CROP, CROP), np.float32)
    audios = np.zeros((B, ta, NFILT * STACK), np.float32)
    for b in range(B):
        videos[b, :vl[b]] = vids[b]
        audios[b, :al[b]] = auds[b]
    return videos.transpose(0, 2, 1, 3, 4), audios.transpose(0, 2, 1), vl, al


def add_noise(waveform, noise, snr, lengths=None):
    """torchaudio.functional.add_noise (2.x), the mixing ``AddMultiSpk`` / ``AddNoise`` do (avhubert_dataset.py:160-222):
    waveform, noise [B, L]; snr [B] in dB; energies over the first lengths[b] samples.  float32 from the energies on."""
    w = np.asarray(waveform, dtype=np.float32)
    z = np.asarray(noise, dtype=np.float32)
    if w.shape != z.shape or w.ndim != 2:
        raise ValueError("waveform and noise must both be [B, L]")
    L = w.shape[1]
    mask = np.ones_like(w) if lengths is None else (np.arange(L)[None, :] < np.asarray(lengths)[:, None]).astype(np.float32)
    es = np.float32(1) * np.sqrt(((w * mask).astype(np.float64) ** 2).sum(1)).astype(np.float32) ** 2
    en = np.float32(1) * np.sqrt(((z * mask).astype(np.float64) ** 2).sum(1)).astype(np.float32) ** 2
    with np.errstate(divide="ignore", invalid="ignore"):
        snr0 = np.float32(10) * (np.log10(es) - np.log10(en))
        scale = np.float32(10) ** ((snr0 - np.asarray(snr, dtype=np.float32)) / np.float32(20))
        return (w + scale[:, None].astype(np.float32) * z).astype(np.float32)
