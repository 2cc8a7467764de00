"""``avsr_cocktail`` model object for the B200 path.

Mirrors what the reference's evaluation code touches (/root/reference/script/evaluation.py:63-108): an object with
``.encoder`` (callable like ``AVHubertModel.forward``), ``.sos`` / ``.eos`` / ``.odim`` (E2E attributes,
src/nets/backend/e2e_asr_avhubert.py:64-117) that ``get_beam_search_decoder(model, token_list, ...)`` accepts, and an
``inference(videos, audios)`` that returns the 1-best token ids exactly like ``AVSRCocktailModel.inference``.
Weights come in as the reference ``state_dict`` (same key names).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import _lib as L
from .beam_search import BatchedBeamSearch, Hypothesis, get_beam_search_decoder
from .encoder import Encoder
from .weights import DecoderWeights


class AVSRCocktailB200:
    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda:0", beam_size: int = 3, ctc_weight: float = 0.1,
                 token_list: Optional[Sequence[str]] = None):
        self.device = torch.device(device)
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise RuntimeError("AVSRCocktailB200 needs a CUDA device: the hot path has no CPU fallback")
        L.load()
        torch.cuda.set_device(self.device)
        self.encoder = Encoder(state_dict, self.device)
        self.decoder_weights = DecoderWeights(state_dict, self.device)
        self.odim = self.decoder_weights.V
        self.sos = self.eos = self.odim - 1
        self.blank = 0
        self.token_list = token_list
        self._decoder_scorer = None
        self._ctc_head = None
        self.beam_search = BatchedBeamSearch(self.decoder_weights, beam_size=beam_size, ctc_weight=ctc_weight,
                                             token_list=token_list, device=self.device)

    # E2E attributes the reference's own factory reads (src/avhubert_avsr/avhubert_avsr_model.py:12-17: ``model.decoder``,
    # ``CTCPrefixScorer(model.ctc, model.eos)``): scorer plug-in objects over the same kernels, created on first use so that
    # they derive from the reference's ScorerInterface when its package has been imported by then (avsr_b200/scorers.py)
    @property
    def decoder(self):
        from . import scorers
        if not isinstance(self._decoder_scorer, scorers.scorer_classes()[0]):       # also re-made once the reference's interfaces appear
            self._decoder_scorer = scorers.B200DecoderScorer(self.decoder_weights, self.device)
        return self._decoder_scorer

    @property
    def ctc(self):
        if self._ctc_head is None:
            from . import scorers
            self._ctc_head = scorers.B200CTCHead(self.decoder_weights, self.device)
        return self._ctc_head

    def ctc_prefix_scorer(self):
        """B200 stand-in for the reference's ``CTCPrefixScorer(model.ctc, model.eos)`` (src/nets/scorers/ctc.py:10-126)."""
        from . import scorers
        return scorers.B200CTCPrefixScorer(self.decoder_weights, self.device, eos=self.eos)

    def eval(self):
        return self

    def cuda(self):
        return self

    # script/evaluation.py:96-108 -------------------------------------------------------------------------------
    def inference(self, videos: torch.Tensor, audios: torch.Tensor) -> List[int]:
        """One utterance: videos [1,1,T,88,88], audios [1,104,T] -> token ids of the 1-best without sos
        (``nbest[0].asdict()["yseq"][1:]``)."""
        x = self.encoder(input_features=audios, video=videos).last_hidden_state.squeeze(0)
        nbest = self.beam_search(x)
        return nbest[0].asdict()["yseq"][1:]

    def infer_batch(self, videos: torch.Tensor, audios: torch.Tensor, lengths: Optional[Sequence[int]] = None,
                    max_steps: Optional[int] = None) -> List[List[Hypothesis]]:
        """Many utterances at once: videos [B,1,T,88,88], audios [B,104,T], optional frame counts.
        Returns the n-best list of every utterance, each equal to its own B=1 run."""
        out = self.encoder(input_features=audios, video=videos, lengths=lengths)
        return self.beam_search.decode_batch(out.packed, out.lengths, max_steps=max_steps)
