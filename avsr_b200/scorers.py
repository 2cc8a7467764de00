"""The reference's scorer plug-in API over the B200 kernels: the bisectable integration point.

The reference's beam search drives its scorers through ``ScorerInterface`` / ``BatchScorerInterface`` /
``BatchPartialScorerInterface`` (/root/reference/src/nets/scorer_interface.py:9-186).  The classes here implement that API
on the C ABI, so the reference's OWN ``BatchBeamSearch`` (src/nets/batch_beam_search.py:138-285) can run unmodified with
either or both of them swapped in:

* ``B200DecoderScorer``  - ``Decoder.batch_score`` / ``batch_init_state`` / ``select_state``
  (src/nets/backend/transformer/decoder.py:195-227, scorer_interface.py:27-50,86-95): next-token log-probabilities of every
  running hypothesis from the KV-cache decoder-step kernels.
* ``B200CTCPrefixScorer`` - ``CTCPrefixScorer.batch_init_state`` / ``batch_score_partial`` / ``select_state``
  (src/nets/scorers/ctc.py:40-63,87-126): CTC prefix scores of the pre-beam candidates from ``avsr_ctc_prefix_prebeam``.
* ``B200CTCHead`` - the ``model.ctc`` object of ``E2E`` (``log_softmax(hs_pad)``, src/nets/backend/ctc.py:163-170), enough for the
  reference's own ``CTCPrefixScorer(model.ctc, model.eos)`` (avhubert_avsr_model.py:15).

``AVSRCocktailB200.decoder`` / ``.ctc`` return these, so ``get_beam_search_decoder`` of the REFERENCE accepts the B200 model.
When the reference's ``src.nets.scorer_interface`` is importable the classes derive from ITS interfaces (its ``BeamSearch``
asserts ``isinstance(v, ScorerInterface)``, beam_search.py:79-81); otherwise from equivalent local ones.

Scorer states are what the reference threads through its hypotheses.  Decoder: the hypothesis' ancestry, one cache slot per
position (the kernels keep K/V of position p at (p, slot) and never copy them).  CTC: the index of the hypothesis' forward
variables in the kernel's ping-pong buffer plus ``log_psi`` of its prefix.  Both are only valid while the search that
produced them is running, like the reference's own states (its CTC scorer keeps ``self.impl``).

This is the integration surface, not the fast path: ``BatchedBeamSearch`` fuses all of it on the device.
"""
from __future__ import annotations

import ctypes as C
import sys
from typing import Any, List, Optional, Tuple

import numpy as np
import torch

from . import _lib as L
from .beam_search import BatchedBeamSearch
from .weights import DecoderWeights

LOGZERO = -10000000000.0          # ctc_prefix_score.py:33
MAX_HYPS = 8                      # rows one scorer call may carry (cache slots per position)


# ------------------------------------------------------------------------------------------------ interfaces
class _ScorerInterface:
    """Local mirror of src/nets/scorer_interface.py:9-81 (used when the reference is not importable)."""

    def init_state(self, x: torch.Tensor) -> Any:
        return None

    def select_state(self, state: Any, i: int, new_id: int = None) -> Any:
        return None if state is None else state[i]

    def score(self, y, state, x):
        raise NotImplementedError

    def final_score(self, state: Any) -> float:
        return 0.0


class _BatchScorerInterface(_ScorerInterface):
    """scorer_interface.py:83-126."""

    def batch_init_state(self, x: torch.Tensor) -> Any:
        return self.init_state(x)

    def batch_score(self, ys, states, xs):
        raise NotImplementedError


class _PartialScorerInterface(_ScorerInterface):
    """scorer_interface.py:129-159."""

    def score_partial(self, y, next_tokens, state, x):
        raise NotImplementedError


class _BatchPartialScorerInterface(_BatchScorerInterface, _PartialScorerInterface):
    """scorer_interface.py:162-186."""

    def batch_score_partial(self, ys, next_tokens, states, xs):
        raise NotImplementedError


def reference_interfaces():
    """(BatchScorerInterface, BatchPartialScorerInterface) of the reference if its package is importable, else the mirrors."""
    mod = sys.modules.get("src.nets.scorer_interface")
    if mod is None:
        try:
            import importlib
            mod = importlib.import_module("src.nets.scorer_interface")
        except Exception:
            mod = None
    if mod is not None and hasattr(mod, "BatchPartialScorerInterface"):
        return mod.BatchScorerInterface, mod.BatchPartialScorerInterface
    return _BatchScorerInterface, _BatchPartialScorerInterface


# ------------------------------------------------------------------------------------------------ implementations
class _DecoderScorerImpl:
    def __init__(self, weights: DecoderWeights, device, precision: str = "bf16x3"):
        self.device = torch.device(device)
        self._bs = BatchedBeamSearch(weights, beam_size=MAX_HYPS, ctc_weight=0.5, device=self.device, use_graph=False, precision=precision)
        self._s = None
        self._x_key = None

    # scorer_interface.py:27-50, 86-95
    def init_state(self, x):
        return None

    def batch_init_state(self, x):
        return None

    def select_state(self, state, i, new_id=None):
        return None if state is None else state[i]

    def final_score(self, state):
        return 0.0

    def _session_for(self, x: torch.Tensor):
        """Cross-attention K/V of the utterance, computed when a search starts (the reference re-projects them at every
        position, attention.py:50-52)."""
        T = x.shape[0]
        s = self._bs._session(1, T, T)
        self._bs.prepare(s, x, [T], ctc=False)
        self._s = s
        return s

    def batch_score(self, ys: torch.Tensor, states: List[Any], xs: torch.Tensor) -> Tuple[torch.Tensor, List[Any]]:
        """decoder.py:195-227: ys [n, L] int64 prefixes, states = per-hyp scorer states (None at the first position), xs
        [n, T, 1024] the encoder output expanded -> (log-probabilities [n, V], new per-hyp states)."""
        n, Lp = ys.shape
        if n > MAX_HYPS:
            raise RuntimeError(f"B200DecoderScorer scores at most {MAX_HYPS} hypotheses per call, got {n}")
        if not xs.is_cuda or xs.dtype != torch.float32:
            raise RuntimeError("encoder output must be a float32 CUDA tensor (avsr_b200 has no CPU path)")
        lib = L.load()
        bs = self._bs
        with torch.cuda.device(self.device):
            if states[0] is None:
                if Lp != 1:
                    raise RuntimeError("B200DecoderScorer: a hypothesis without state must be the empty prefix [sos]")
                s = self._session_for(xs[0].contiguous())
            else:
                s = self._s
            if s is None:
                raise RuntimeError("B200DecoderScorer: batch_score with states but no search in progress")
            step = Lp - 1
            if step >= s["lmax"]:
                raise RuntimeError("prefix longer than the utterance allows")
            s["step"].fill_(step)
            s["n_run"][0] = n
            s["last_tok"][:n] = ys[:, -1].to(torch.int32)
            if step > 0:
                paths = np.frombuffer(b"".join(states), dtype=np.uint8).reshape(n, step)
                s["anc"][step & 1, :n, :step] = torch.from_numpy(paths.copy()).to(self.device)
            part, ns = bs._decoder_layers(s, dense=False)
            L.check(lib.avsr_dec_logits_lsm_topk(L.ptr(part), ns, s["R"], bs.n_vocab, L.ptr(bs.w.out_b), L.ptr(s["n_run"]), MAX_HYPS,
                                                 L.ptr(s["dec_logp"]), L.ptr(s["part_ids"]), bs.pre_beam_size, L.stream()),
                    "avsr_dec_logits_lsm_topk")
            logp = s["dec_logp"][:n].clone()
        # K/V of this position sit at (position, slot = row); a hypothesis extends its parent's path by the parent's row
        new_states = [(states[h] if step > 0 else b"") + bytes([h]) for h in range(n)]
        return logp, new_states


class _CTCScorerImpl:
    def __init__(self, weights: DecoderWeights, device, eos: Optional[int] = None, precision: str = "bf16x3"):
        self.device = torch.device(device)
        self.eos = weights.eos if eos is None else eos
        self._bs = BatchedBeamSearch(weights, beam_size=MAX_HYPS, ctc_weight=0.5, device=self.device, use_graph=False, precision=precision)
        self._s = None
        self.impl = None          # the reference keeps its CTCPrefixScoreTH here (scorers/ctc.py:24,98)

    def init_state(self, x):
        return self.batch_init_state(x)

    def final_score(self, state):
        return 0.0

    def batch_init_state(self, x: torch.Tensor):
        """scorers/ctc.py:87-99: CTC posteriors of the utterance (assuming batch_size = 1, as the reference says)."""
        if not x.is_cuda or x.dtype != torch.float32:
            raise RuntimeError("encoder output must be a float32 CUDA tensor (avsr_b200 has no CPU path)")
        T = x.shape[0]
        with torch.cuda.device(self.device):
            s = self._bs._session(1, T, T)
            self._bs.prepare(s, x.contiguous(), [T], cross_kv=False)
        self._s = s
        self.impl = self
        return None

    def batch_score_partial(self, y: torch.Tensor, ids: torch.Tensor, state: List[Any], x: torch.Tensor):
        """scorers/ctc.py:101-126 -> CTCPrefixScoreTH.__call__ (ctc_prefix_score.py:68-187): y [n, L] prefixes, ids [n, S]
        pre-beam candidates, state = per-hyp states (None at the first position) -> (scores [n, V], batch state)."""
        s, bs = self._s, self._bs
        if s is None:
            raise RuntimeError("B200CTCPrefixScorer: batch_init_state(x) has not been called")
        if ids is None:
            raise RuntimeError("B200CTCPrefixScorer scores pre-beam candidates; full-vocabulary scoring is BatchedBeamSearch(ctc_weight=1.0)")
        n, Lp = y.shape
        S = ids.shape[-1]
        if n > MAX_HYPS or S > bs.pre_beam_size:
            raise RuntimeError(f"B200CTCPrefixScorer: at most {MAX_HYPS} hypotheses x {bs.pre_beam_size} candidates per call")
        lib = L.load()
        V = bs.n_vocab
        step = Lp - 1
        with torch.cuda.device(self.device):
            s["step"].fill_(step)
            s["n_run"][0] = n
            s["last_tok"][:n] = y[:, -1].to(torch.int32)
            if state[0] is None:
                s["rprev_idx"].zero_()
                s["s_prev"].zero_()
            else:
                s["rprev_idx"][:n] = torch.tensor([st[0] for st in state], dtype=torch.int32, device=self.device)
                s["s_prev"][:n] = torch.tensor([st[1] for st in state], dtype=torch.float32, device=self.device)
            part = torch.zeros(s["R"], S, dtype=torch.int32, device=self.device)
            part[:n] = ids.to(torch.int32)
            psi = torch.zeros(s["R"], S, dtype=torch.float32, device=self.device)
            L.check(lib.avsr_ctc_prefix_prebeam(L.ptr(s["logp"]), V, s["ldp"], bs.w.blank, L.ptr(s["utt_off"]), L.ptr(s["utt_T"]), L.ptr(s["n_run"]),
                                                MAX_HYPS, s["R"], S, L.ptr(s["last_tok"]), L.ptr(part), L.ptr(s["rprev_idx"]), L.ptr(s["r_buf"]),
                                                s["tmax"], L.ptr(s["step"]), L.ptr(psi), L.ptr(s["rsum_last"]), L.stream()),
                    "avsr_ctc_prefix_prebeam")
            scores = torch.empty(n, V, dtype=torch.float32, device=self.device)
            L.check(lib.avsr_ctc_scores_dense(L.ptr(psi), L.ptr(s["rsum_last"]), L.ptr(s["s_prev"]), L.ptr(part), n, S, V, bs.w.blank, self.eos,
                                              L.ptr(scores), L.stream()), "avsr_ctc_scores_dense")
            # the batch state select_state picks from: candidate ids, their log_psi, log_psi of eos, chain pitch S
            batch_state = (ids.cpu().numpy(), psi[:n].cpu().numpy(), s["rsum_last"][:n].cpu().numpy(), S)
        return scores, batch_state

    def select_state(self, state, i, new_id=None):
        """scorers/ctc.py:40-63: state of the hypothesis that extends row i with token new_id = the chain the kernel wrote for
        that candidate (row * S + column; a token outside the candidates maps to the last column, the reference's
        ``scoring_idmap == -1`` quirk) and log_psi[i, new_id]."""
        if state is None:
            return None
        if not (isinstance(state, tuple) and len(state) == 4):
            return state[i]                   # a list of per-hypothesis states (BatchBeamSearch._batch_select / unbatchfy)
        ids, psi, rsum, S = state
        i, new_id = int(i), int(new_id)
        hit = np.nonzero(ids[i] == new_id)[0]
        col = int(hit[-1]) if hit.size else S - 1
        if new_id == self._bs.w.blank:
            lp = LOGZERO
        elif new_id == self.eos:
            lp = float(rsum[i])
        else:
            lp = float(psi[i, col]) if hit.size else LOGZERO
        return (i * S + col, lp)


class B200CTCHead:
    """``model.ctc`` (E2E.ctc, e2e_asr_avhubert.py:104-111): ``log_softmax(hs_pad [B, T, 1024]) -> [B, T, V]``
    (src/nets/backend/ctc.py:163-170) on the tensor cores (bf16x3, fp32-level accuracy)."""

    def __init__(self, weights: DecoderWeights, device):
        self.w = weights
        self.device = torch.device(device)

    def log_softmax(self, hs_pad: torch.Tensor) -> torch.Tensor:
        if not hs_pad.is_cuda or hs_pad.dtype != torch.float32:
            raise RuntimeError("hs_pad must be a float32 CUDA tensor (avsr_b200 has no CPU path)")
        lib = L.load()
        w = self.w
        B, T, D = hs_pad.shape
        F = B * T
        ldp = (w.V + 31) // 32 * 32
        with torch.cuda.device(self.device):
            x = hs_pad.reshape(F, D).contiguous()
            x6 = torch.empty(F, 6 * D, dtype=torch.bfloat16, device=self.device)
            out = torch.zeros(F, ldp, dtype=torch.float32, device=self.device)
            L.check(lib.avsr_split3(L.ptr(x), L.ll(D), L.ptr(x6), L.ll(F), D, L.stream()), "avsr_split3")
            L.gemm_bf16(x6, w.ctc_w6, F, w.V, 6 * D, L.make_epilogue(bias=w.ctc_b, out_f32=out, ld_f32=ldp))
            L.check(lib.avsr_log_softmax_rows(L.ptr(out), L.ll(ldp), L.ll(F), w.V, L.stream()), "avsr_log_softmax_rows")
        return out.view(B, T, ldp)[:, :, :w.V]


_CLASS_CACHE = {}


def scorer_classes():
    """(B200DecoderScorer, B200CTCPrefixScorer) deriving from the scorer interfaces in force (see the module docstring)."""
    full_base, part_base = reference_interfaces()
    key = (id(full_base), id(part_base))
    if key not in _CLASS_CACHE:
        dec = type("B200DecoderScorer", (_DecoderScorerImpl, full_base), {"__doc__": _DecoderScorerImpl.batch_score.__doc__})
        ctc = type("B200CTCPrefixScorer", (_CTCScorerImpl, part_base), {"__doc__": _CTCScorerImpl.batch_score_partial.__doc__})
        _CLASS_CACHE[key] = (dec, ctc)
    return _CLASS_CACHE[key]


def B200DecoderScorer(weights: DecoderWeights, device="cuda:0", precision: str = "bf16x3"):
    return scorer_classes()[0](weights, device, precision=precision)


def B200CTCPrefixScorer(weights: DecoderWeights, device="cuda:0", eos: Optional[int] = None, precision: str = "bf16x3"):
    return scorer_classes()[1](weights, device, eos=eos, precision=precision)
