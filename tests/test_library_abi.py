"""CPU: the C-ABI library loads and exports every symbol include/avsr_b200.h declares; host-side logic that needs no GPU."""
import ctypes
import os

import pytest
import torch

from avsr_b200 import _lib, synth


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = _lib.load()
    names = _lib.declared_symbols()
    assert len(names) >= 24
    for n in names:
        assert hasattr(lib, n), n
    assert lib.avsr_abi_version() == 1


def test_struct_sizes_match_header():
    # AvsrEpilogue: 13 fields with natural alignment; AvsrBeamState: 10 ints + 25 pointers + 1 double + 1 pointer (utt_maxlen)
    assert ctypes.sizeof(_lib.Epilogue) == 96
    assert ctypes.sizeof(_lib.BeamState) == 40 + 25 * 8 + 8 + 8


def test_argument_errors_are_reported_not_crashed():
    lib = _lib.load()
    rc = lib.avsr_layernorm(None, ctypes.c_longlong(0), ctypes.c_longlong(0), 0, None, None, ctypes.c_float(1e-5), None,
                            ctypes.c_longlong(0), None, ctypes.c_longlong(0), None)
    assert rc == -2
    assert b"avsr_layernorm" in lib.avsr_last_error()


def test_product_fails_loudly_without_gpu(state_dict):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from avsr_b200.model import AVSRCocktailB200
    with pytest.raises(RuntimeError):
        AVSRCocktailB200(state_dict, device="cuda:0")
    with pytest.raises(RuntimeError):
        AVSRCocktailB200(state_dict, device="cpu")


def test_product_never_imports_oracle():
    import re
    root = os.path.dirname(_lib._HERE)
    for fn in os.listdir(_lib._HERE):
        if fn.endswith(".py"):
            src = open(os.path.join(_lib._HERE, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), fn


def test_weight_repack_on_cpu_tensors(state_dict):
    """Folding math of the repacker (BN fold, weight-norm fold, q-scale fold) against the oracle's plain formulas."""
    from avsr_b200.weights import _fold_bn
    from oracle import avsr_oracle as O
    import torch.nn.functional as F
    p = "encoder.feature_extractor_video.resnet.trunk.layer2.0."
    w, b = _fold_bn(state_dict, p + "conv1.weight", p + "bn1")
    x = torch.randn(2, 64, 9, 9, dtype=torch.float64)
    y0 = F.conv2d(x, w, b, stride=2, padding=1)
    y1 = O._bn({k: v.double() if v.is_floating_point() else v for k, v in state_dict.items() if k.startswith(p)}, p + "bn1",
               F.conv2d(x, state_dict[p + "conv1.weight"].double(), None, stride=2, padding=1))
    assert (y0 - y1).abs().max() < 1e-10


def test_filterbank_host_helpers_match_oracle():
    """avsr_fbank_bins / avsr_fbank_rows are host-only: the mel bin edges and frame arithmetic of the CUDA path against the
    oracle's restatement of python_speech_features."""
    from oracle import input_oracle as IO
    lib = _lib.load()
    bins = (ctypes.c_int * 28)()
    assert lib.avsr_fbank_bins(bins) == 0
    assert list(bins) == [int(b) for b in IO.filterbank_bins()]
    for n in (1, 2, 399, 400, 401, 560, 561, 640, 1040, 16000, 640 * 375, 640 * 375 + 1):
        assert lib.avsr_fbank_rows(n) == (IO.num_frames(n) + 3) // 4, n
    assert lib.avsr_fbank_rows(0) == 0


def test_input_pipeline_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from avsr_b200 import input_pipeline as P
    with pytest.raises(RuntimeError):
        P.fbank_stack_ln_batch([torch.zeros(1000)], device="cpu")
    with pytest.raises(Exception):
        P.DataCollator()([{"video": torch.zeros(3, 1, 96, 96, dtype=torch.uint8), "audio": torch.zeros(1920, 1)}])
