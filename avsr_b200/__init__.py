"""avsr_b200: B200-native (sm_100a) inference hot path of AVSRCocktail (quanpn90/avsr, model type ``avsr_cocktail``).

Encoder forward (AV-HuBERT-large) + joint CTC/attention beam search, behind the reference's own model/beam-search API.
All compute runs in hand-written CUDA kernels exported by ``libavsr_b200.so`` (C ABI in ``include/avsr_b200.h``);
PyTorch is used for device memory, streams, CUDA graphs and ``torch.distributed`` only.
"""
from .synth import make_inputs, make_state_dict  # noqa: F401

__all__ = ["make_inputs", "make_state_dict", "AVSRCocktailB200", "Encoder", "BatchedBeamSearch", "Hypothesis",
           "get_beam_search_decoder"]


def __getattr__(name):
    # heavy modules are imported lazily so that `import avsr_b200` works on a box without the built library
    if name in ("AVSRCocktailB200",):
        from .model import AVSRCocktailB200
        return AVSRCocktailB200
    if name == "Encoder":
        from .encoder import Encoder
        return Encoder
    if name in ("BatchedBeamSearch", "Hypothesis", "get_beam_search_decoder"):
        from . import beam_search
        return getattr(beam_search, name)
    raise AttributeError(name)
