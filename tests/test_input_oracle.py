"""oracle/input_oracle.py against outputs of the unmodified reference classes (tests/golden/input_pipeline.npz, made by
oracle/gen_golden_input.py), plus independent checks of the restated python_speech_features pieces (that package is absent
here, see the oracle header: its logfbank is restated from the published algorithm, parity unpinned)."""
import os

import numpy as np
import pytest

from oracle import input_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def G():
    g = np.load(os.path.join(ROOT, "tests", "golden", "input_pipeline.npz"))
    return {k: g[k] for k in g.files}


def test_collate_matches_reference_outputs(G):
    n = len(G["T_list"])
    videos, audios, vl, al = O.collate([G[f"video_{i}"] for i in range(n)], [G[f"wave_{i}"] for i in range(n)])
    assert vl == G["video_lengths"].tolist() and al == G["audio_lengths"].tolist()
    assert videos.shape == G["videos_out"].shape and audios.shape == G["audios_out"].shape
    assert np.array_equal(videos, G["videos_out"])                       # uint8 -> float32 arithmetic: bit exact
    # torch's float32 layer norm vs the oracle's float64 one: rounding only
    assert np.abs(audios - G["audios_out"]).max() < 5e-6


def test_odd_waveform_lengths(G):
    for n in G["odd_lens"].tolist():
        got = O.fbanks_and_stack(G[f"odd_wave_{n}"])
        want = G[f"odd_feat_{n}"]
        assert got.shape == want.shape == ((O.num_frames(n) + 3) // 4, 104), n
        assert np.abs(got - want).max() < 5e-6, n


def test_video_transform_other_frame_size(G):
    assert np.array_equal(O.video_transform(G["video_100x120"]), G["video_100x120_out"])
    with pytest.raises(ValueError):
        O.video_transform(np.zeros((1, 1, 80, 96), np.uint8))


def test_frame_count_and_zero_rows():
    assert [O.num_frames(n) for n in (1, 400, 401, 560, 561, 640 * 375)] == [1, 1, 2, 2, 3, 4 * 375 - 1]
    x = np.random.default_rng(0).standard_normal(640 * 5).astype(np.float32)
    raw = O.stacker(O.logfbank_psf(x).astype(np.float32))
    assert raw.shape == (5, 104) and np.all(raw[-1, 78:] == 0) and np.all(raw[-1, :78] != 0)


def test_filterbank_matches_a_direct_construction():
    """The restated get_filterbanks: 26 triangles on FFT bins, peak 1 at the middle edge, each rising from its left edge."""
    b = O.filterbank_bins()
    assert b[0] == 0 and b[-1] == 256 and np.all(np.diff(b) >= 1) and len(b) == 28
    fb = O.get_filterbanks()
    assert fb.shape == (26, 257)
    for j in range(26):
        lo, mid, hi = int(b[j]), int(b[j + 1]), int(b[j + 2])
        assert fb[j, mid] == 1.0 and fb[j, lo] == 0.0 and np.all(fb[j, :lo] == 0) and np.all(fb[j, hi:] == 0)
        assert np.all(np.diff(fb[j, lo:mid + 1]) > 0) and np.all(np.diff(fb[j, mid:hi]) < 0)


def test_power_spectrum_against_a_direct_dft():
    """rfft-based frames of logfbank_psf vs an explicit O(N^2) DFT of the pre-emphasised, zero-padded frames."""
    rng = np.random.default_rng(1)
    x = rng.standard_normal(1000).astype(np.float32)
    pre = np.append(x[0], x[1:] - np.float32(0.97) * x[:-1]).astype(np.float64)
    nf = O.num_frames(len(x))
    pad = np.concatenate([pre, np.zeros((nf - 1) * 160 + 400 - len(x))])
    k = np.arange(257)[:, None] * np.arange(400)[None, :]
    E = np.exp(-2j * np.pi * k / 512)
    fb = O.get_filterbanks()
    want = np.stack([np.log(fb @ (np.abs(E @ pad[f * 160:f * 160 + 400]) ** 2 / 512)) for f in range(nf)])
    assert np.abs(O.logfbank_psf(x) - want).max() < 1e-9


def test_silent_frames_take_the_eps_floor():
    f = O.logfbank_psf(np.zeros(2000, np.float32))
    assert np.all(f == np.log(np.finfo(float).eps))


def test_add_noise_matches_torchaudio():
    """torchaudio is a library present in the image (not the reference tree); the oracle's mix against its add_noise."""
    torch = pytest.importorskip("torch")
    ta = pytest.importorskip("torchaudio")
    rng = np.random.default_rng(9)
    w = (0.3 * rng.standard_normal((4, 5000))).astype(np.float32)
    z = (0.05 * rng.standard_normal((4, 5000))).astype(np.float32)
    snr = np.array([-5, 0, 10, 15], np.float32)
    for lengths in (None, np.array([5000, 4000, 123, 1])):
        want = ta.functional.add_noise(torch.from_numpy(w), torch.from_numpy(z), torch.from_numpy(snr),
                                       None if lengths is None else torch.from_numpy(lengths)).numpy()
        got = O.add_noise(w, z, snr, lengths)
        assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
