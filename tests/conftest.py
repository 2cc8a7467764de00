import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def state_dict():
    from avsr_b200 import synth
    return synth.make_state_dict(0)


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    g = np.load(os.path.join(ROOT, "tests", "golden", "model_seed0.npz"))
    return {k: g[k] for k in g.files}


@pytest.fixture(scope="session")
def golden_cfg0():
    """BASELINE.json configs[0]: one 15 s utterance (T=375), reference encoder output + reference n-best at beam 3 / 5
    (oracle/gen_golden_cfg0.py)."""
    import numpy as np
    g = np.load(os.path.join(ROOT, "tests", "golden", "cfg0_T375.npz"))
    return {k: g[k] for k in g.files}


@pytest.fixture(scope="session")
def golden_ctc():
    import numpy as np
    g = np.load(os.path.join(ROOT, "tests", "golden", "ctc_prefix.npz"))
    return {k: g[k] for k in g.files}


@pytest.fixture(scope="session")
def gpu_model(state_dict):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from avsr_b200.model import AVSRCocktailB200
    return AVSRCocktailB200(state_dict, device="cuda:0", beam_size=3)
