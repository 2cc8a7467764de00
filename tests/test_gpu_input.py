"""GPU: the input pipeline kernels (avsr_fbank_stack_ln, avsr_video_u8_transform) through avsr_b200/input_pipeline.py, against
outputs of the reference classes (tests/golden/input_pipeline.npz) and against the oracle on seeded inputs.

Tolerances: the video transform is float32 arithmetic on 256 possible inputs -> bit exact.  The filterbank is float64 up to
the float32 rounding the reference does too, then a layer norm (torch float32 in the reference, float64 here and in the
oracle): 2e-5 absolute on values of order 1."""
import os

import numpy as np
import pytest
import torch

from oracle import input_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AUDIO_TOL = 2e-5


@pytest.fixture(scope="module")
def P():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from avsr_b200 import input_pipeline
    return input_pipeline


@pytest.fixture(scope="module")
def G():
    g = np.load(os.path.join(ROOT, "tests", "golden", "input_pipeline.npz"))
    return {k: g[k] for k in g.files}


def test_collator_matches_reference_batch(P, G):
    n = len(G["T_list"])
    feats = [{"video": torch.from_numpy(G[f"video_{i}"]), "audio": torch.from_numpy(G[f"wave_{i}"])[:, None]} for i in range(n)]
    batch = P.DataCollator()(feats)
    assert batch["video_lengths"].tolist() == G["video_lengths"].tolist()
    assert batch["audio_lengths"].tolist() == G["audio_lengths"].tolist()
    assert batch["videos"].is_cuda and batch["audios"].is_cuda
    assert np.array_equal(batch["videos"].cpu().numpy(), G["videos_out"])
    got = batch["audios"].cpu().numpy()
    assert got.shape == G["audios_out"].shape
    assert np.abs(got - G["audios_out"]).max() < AUDIO_TOL
    for b, rows in enumerate(G["audio_lengths"].tolist()):                # collate_pad's zero padding, exactly zero
        assert np.all(got[b, :, rows:] == 0)


def test_per_sample_modules_match_reference(P, G):
    fb = P.FBanksAndStack()
    for n in G["odd_lens"].tolist():
        got = fb(torch.from_numpy(G[f"odd_wave_{n}"])[:, None]).cpu().numpy()
        want = G[f"odd_feat_{n}"]
        assert got.shape == want.shape, n
        assert np.abs(got - want).max() < AUDIO_TOL, n
    vt = P.VideoTransform("test")
    got = vt(torch.from_numpy(G["video_100x120"])).cpu().numpy()          # unaligned crop offsets: the byte-load variant
    assert np.array_equal(got, G["video_100x120_out"])
    with pytest.raises(NotImplementedError):
        P.VideoTransform("train")
    with pytest.raises(ValueError):
        vt(torch.zeros(2, 1, 80, 96, dtype=torch.uint8))
    with pytest.raises(ValueError):
        vt(torch.zeros(2, 1, 96, 96))


def test_against_oracle_at_output_rounding(P):
    """The oracle normalises in float64 like the CUDA path, so the two differ only by the float32 rounding of the log
    energies' last bit and of the result: a much tighter bound than against torch's float32 layer norm."""
    rng = np.random.default_rng(3)
    w = (0.4 * rng.standard_normal(640 * 40)).astype(np.float32)
    got, rows = P.fbank_stack_ln_batch([torch.from_numpy(w)])
    want = O.fbanks_and_stack(w)
    assert rows == [40]
    assert np.abs(got[0].t().cpu().numpy() - want).max() < 2e-6


@pytest.mark.parametrize("H,W", [(96, 96), (88, 88), (97, 101), (112, 112)])
def test_video_transform_frame_sizes(P, H, W):
    rng = np.random.default_rng(H * 1000 + W)
    vids = [rng.integers(0, 256, size=(t, 1, H, W), dtype=np.uint8) for t in (3, 1, 6)]
    out, T = P.video_transform_batch([torch.from_numpy(v) for v in vids])
    assert T == [3, 1, 6] and tuple(out.shape) == (3, 1, 6, 88, 88)
    out = out.cpu().numpy()
    for b, v in enumerate(vids):
        assert np.array_equal(out[b, 0, :T[b]], O.video_transform(v)[:, 0])
        assert np.all(out[b, 0, T[b]:] == 0)


def test_device_resident_inputs_and_t_max(P):
    rng = np.random.default_rng(5)
    w = [torch.from_numpy((0.1 * rng.standard_normal(n)).astype(np.float32)).cuda() for n in (3000, 12345)]
    a, rows = P.fbank_stack_ln_batch(w, t_max=32)
    assert tuple(a.shape) == (2, 104, 32) and rows == [P.fbank_rows(3000), P.fbank_rows(12345)]
    for b in range(2):
        want = O.fbanks_and_stack(w[b].cpu().numpy())
        assert np.abs(a[b, :, :rows[b]].t().cpu().numpy() - want).max() < AUDIO_TOL
        assert torch.all(a[b, :, rows[b]:] == 0)
    with pytest.raises(ValueError):
        P.fbank_stack_ln_batch(w, t_max=4)
    # cut: n_samples below the waveform length; pad: above it
    a2, rows2 = P.fbank_stack_ln_batch(w, n_samples=[1600, 16000])
    for b, n in enumerate((1600, 16000)):
        want = O.fbanks_and_stack(O.cut_or_pad(w[b].cpu().numpy(), n))
        assert np.abs(a2[b, :, :rows2[b]].t().cpu().numpy() - want).max() < AUDIO_TOL


def test_silence_and_extreme_amplitudes(P):
    waves = [np.zeros(6400, np.float32), np.full(6400, 1.0, np.float32), (1e-4 * np.sin(np.arange(6400) * 0.3)).astype(np.float32),
             np.concatenate([np.zeros(3200, np.float32), np.ones(3200, np.float32)])]
    a, rows = P.fbank_stack_ln_batch([torch.from_numpy(w) for w in waves])
    for b, w in enumerate(waves):
        want = O.fbanks_and_stack(w)
        got = a[b, :, :rows[b]].t().cpu().numpy()
        assert np.isfinite(got).all()
        # rows with a constant log energy (all-silent, var = 0) normalise to exactly 0 in both
        assert np.abs(got - want).max() < 1e-4, b


def test_full_size_batch_properties(P):
    """BASELINE configs[1] shape: 32 utterances of 15 s.  Batch results equal the single-utterance results bit for bit
    (every CTA owns one output row), every row is normalised, and a sample of utterances matches the oracle."""
    rng = np.random.default_rng(11)
    B, T = 32, 375
    waves = [(0.2 * rng.standard_normal(T * 640)).astype(np.float32) for _ in range(B)]
    vids = [torch.from_numpy(rng.integers(0, 256, size=(T, 1, 96, 96), dtype=np.uint8)) for _ in range(B)]
    batch = P.DataCollator()([{"video": v, "audio": torch.from_numpy(w)} for v, w in zip(vids, waves)])
    a, v = batch["audios"], batch["videos"]
    assert tuple(a.shape) == (B, 104, T) and tuple(v.shape) == (B, 1, T, 88, 88)
    assert torch.all(a.mean(1).abs() < 1e-5) and torch.all((a.var(1, unbiased=False) - 1).abs() < 1e-3)
    for b in (0, 17, 31):
        single, _ = P.fbank_stack_ln_batch([torch.from_numpy(waves[b])])
        assert torch.equal(single[0], a[b])
        assert np.abs(a[b].t().cpu().numpy() - O.fbanks_and_stack(waves[b])).max() < AUDIO_TOL
        assert np.array_equal(v[b, 0].cpu().numpy(), O.video_transform(vids[b].numpy())[:, 0])
    # the value table: only 256 distinct outputs exist
    assert torch.unique(v).numel() <= 256


def test_pipeline_feeds_the_encoder(P, gpu_model):
    """The collated tensors go straight into the encoder call of the reference API."""
    rng = np.random.default_rng(2)
    T = [20, 11]
    feats = [{"video": torch.from_numpy(rng.integers(0, 256, size=(t, 1, 96, 96), dtype=np.uint8)),
              "audio": torch.from_numpy((0.1 * rng.standard_normal(t * 640 + 13)).astype(np.float32))[:, None]} for t in T]
    batch = P.DataCollator()(feats)
    out = gpu_model.encoder(input_features=batch["audios"], video=batch["videos"]).last_hidden_state
    assert tuple(out.shape)[0] == 2 and tuple(out.shape)[2] == 1024 and torch.isfinite(out).all()


def test_add_noise_matches_torchaudio_and_oracle(P):
    import torchaudio
    rng = np.random.default_rng(21)
    B, n = 5, 160000                                                        # configs[3]: 10 s chunks
    w = torch.from_numpy((0.3 * rng.standard_normal((B, n))).astype(np.float32))
    z = torch.from_numpy((0.02 * rng.standard_normal((B, n))).astype(np.float32))
    snr = torch.tensor([-5.0, 0.0, 5.0, 10.0, 15.0])
    for lengths in (None, torch.tensor([n, n // 2, 77777, 1000, 1])):
        want = torchaudio.functional.add_noise(w, z, snr, lengths)
        got = P.add_noise(w.cuda(), z.cuda(), snr.cuda(), None if lengths is None else lengths.cuda()).cpu()
        tol = 1e-5 * want.abs().max().item()
        assert (got - want).abs().max().item() <= tol
        ora = O.add_noise(w.numpy(), z.numpy(), snr.numpy(), None if lengths is None else lengths.numpy())
        assert np.abs(got.numpy() - ora).max() <= tol
    one = P.add_noise(w[0].cuda(), z[0].cuda(), snr[0].cuda())
    assert one.shape == (n,) and torch.equal(one.cpu(), P.add_noise(w.cuda(), z.cuda(), snr.cuda()).cpu()[0])
    with pytest.raises(ValueError):
        P.add_noise(w.cuda(), z[:, :100].cuda(), snr.cuda())
    with pytest.raises(RuntimeError):
        P.add_noise(w, z, snr)


def test_interferer_mix_then_features(P):
    """configs[3] shape: a 10 s chunk mixed with two interferers the way AddMultiSpk chains add_noise (:196-222), then the
    feature kernel; against the oracle chain."""
    rng = np.random.default_rng(33)
    n = 160000
    sp, i1, i2 = [(a * rng.standard_normal(n)).astype(np.float32) for a in (0.2, 0.1, 0.3)]
    mix_o = O.add_noise(i1[None], i2[None], np.array([5.0], np.float32))
    out_o = O.add_noise(sp[None], mix_o, np.array([0.0], np.float32))[0]
    t = lambda a: torch.from_numpy(a).cuda()
    mix = P.add_noise(t(i1), t(i2), torch.tensor(5.0).cuda())
    out = P.add_noise(t(sp), mix, torch.tensor(0.0).cuda())
    assert np.abs(out.cpu().numpy() - out_o).max() < 1e-5
    feats, rows = P.fbank_stack_ln_batch([out])
    assert rows == [250]
    assert np.abs(feats[0].t().cpu().numpy() - O.fbanks_and_stack(out.cpu().numpy())).max() < AUDIO_TOL
