"""Joint CTC/attention beam search on B200 with the hypothesis state resident on the device.

Drop-in for the object returned by the reference's ``get_beam_search_decoder``
(/root/reference/src/avhubert_avsr/avhubert_avsr_model.py:12-36), i.e. ``BatchBeamSearch``
(src/nets/batch_beam_search.py:26-349, src/nets/beam_search.py:33-105,330-406): ``search(x[T,1024]) -> List[Hypothesis]``
sorted best first, each with ``.yseq`` (int64, starts with sos and ends with eos), ``.score``,
``.scores{"decoder","ctc"}`` and ``.asdict()``.  Semantics kept: weights {decoder: 1-ctc_weight, ctc: ctc_weight},
pre-beam on the decoder scores with ``int(1.5*beam)`` candidates, ``maxlen = T``, eos appended after the last step,
no length normalisation, ``end_detect`` with M=3 / D_end=-10.

Beyond the reference, ``decode_batch`` runs many utterances (mixed lengths) at once; every utterance evolves exactly as
its own B=1 run (per-utterance maxlen, top-k, early exit).  One decode step is a fixed sequence of
kernel launches that read the step index and liveness from device memory, so it is captured once into a CUDA graph and
replayed; the host only polls an "any utterance still running" flag every few steps.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, NamedTuple, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib as L
from .weights import DecoderWeights


class Hypothesis(NamedTuple):
    """Same fields as the reference's Hypothesis (src/nets/beam_search.py:13-27)."""

    yseq: torch.Tensor
    score: Union[float, torch.Tensor] = 0
    scores: Dict[str, Union[float, torch.Tensor]] = dict()
    states: Dict[str, object] = dict()

    def asdict(self) -> dict:
        return self._replace(
            yseq=self.yseq.tolist(),
            score=float(self.score),
            scores={k: float(v) for k, v in self.scores.items()},
        )._asdict()


D_END = math.log(1 * math.exp(-10))      # e2e_asr_common.py:18


class BatchedBeamSearch:
    POLL_EVERY = 16

    def __init__(self, weights: DecoderWeights, beam_size: int = 3, ctc_weight: float = 0.1, pre_beam_ratio: float = 1.5,
                 token_list: Optional[Sequence[str]] = None, device="cuda:0", use_graph: bool = True,
                 precision: str = "bf16x3"):
        """precision: numerics of the decoder / CTC-head projections.  "bf16x3" = three-term bf16 split of both operands on
        the tcgen05 tensor cores (fp32-level accuracy, default); "fp32" = plain fp32 FMA on the CUDA cores."""
        if precision not in ("bf16x3", "fp32"):
            raise RuntimeError(f"unknown precision {precision!r}")
        self.precision = precision
        self.w = weights
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("avsr_b200 beam search needs a CUDA device (no CPU fallback)")
        if not (1 <= beam_size <= 8):
            raise RuntimeError("beam_size must be in [1, 8]")
        if not (0.0 <= ctc_weight <= 1.0):
            raise RuntimeError("ctc_weight must be in [0, 1]")
        # ctc_weight == 1.0: the reference drops the decoder scorer (weight 0, beam_search.py:72-75) and the pre-beam
        # (avhubert_avsr_model.py:35): every step scores the FULL vocabulary with CTCPrefixScoreTH (ctc_prefix_score.py:115-119)
        self.ctc_only = ctc_weight == 1.0
        # ctc_weight == 0.0: the CTC scorer is dropped instead (no partial scorer left -> no pre-beam, beam_search.py:100-104):
        # attention-only search, top-k over the decoder's full-vocabulary log-probabilities
        self.dec_only = ctc_weight == 0.0
        if token_list is not None and len(token_list) != weights.V:
            raise RuntimeError(f"token_list has {len(token_list)} entries, model vocabulary is {weights.V}")
        self.beam_size = beam_size
        # beam_search.py:91; single-scorer searches have no pre-beam (ctc-only keeps one chain per survivor)
        self.pre_beam_size = 1 if (self.ctc_only or self.dec_only) else int(pre_beam_ratio * beam_size)
        self.n_vocab = weights.V
        self.sos, self.eos = weights.sos, weights.eos
        self.token_list = token_list
        self.weights = {"decoder": 1.0 - ctc_weight, "ctc": ctc_weight, "lm": 0.0, "length_bonus": 0.0}
        self.scorers = ("ctc",) if self.ctc_only else (("decoder",) if self.dec_only else ("decoder", "ctc"))
        self.w_dec = float(np.float32(1.0 - ctc_weight))
        self.w_ctc = float(np.float32(ctc_weight))
        self.use_graph = use_graph
        self.fuse_epilogue = False
        # stage the layer's cross-attention K/V in L2 ahead of the source attention (AVSR_L2_PREFETCH=0: dev A/B switch)
        # (measured SLOWER on B200, 477 vs 461 ms per 32-utterance pass, so it is off unless AVSR_L2_PREFETCH=1)
        self.l2_prefetch = os.environ.get("AVSR_L2_PREFETCH", "0") == "1"
        self._skip = frozenset()      # dev aid (tools/ablate_step.py): kernel groups left out of _step for timing ablations
        # bf16x3 path, opt-in (AVSR_CHAIN=1): a projection launch also finishes the rows of the previous projection (its own
        # operand), 54 launches per position instead of 78.  Measured SLOWER on B200 (873 vs 815 us per position): releasing
        # the operand rows through a device counter (threadfence + proxy fence + atomic + acquire spin) costs more than the
        # programmatic-dependent-launch boundary it replaces, so the default keeps the separate row-epilogue launches.
        self.chain = os.environ.get("AVSR_CHAIN", "0") == "1"
        # decoder-step projections of the bf16x3 path: "cluster" (default) = split-K inside a thread-block cluster reduced through
        # DSMEM with the LayerNorm applied by the consuming projection (csrc/gemm_x3c.cu, 54 launches per position);
        # "splitk" = round 1's partial sums through L2 + row-epilogue launches (csrc/gemm_x3.cu, 78 launches), kept for A/B runs
        self.proj = os.environ.get("AVSR_PROJ", "cluster")
        self.weight_prefetch = os.environ.get("AVSR_WEIGHT_PREFETCH", "1") != "0"      # dev A/B switch of the L2 fetch-ahead
        # LayerNorm folded into the consuming projection (avsr_dec_proj_folded: raw rows through TMA, mean / rstd applied to the
        # finished sums) instead of normalising while the operand is staged; AVSR_LN_FOLD=0 keeps the staged form (A/B switch)
        self.ln_fold = os.environ.get("AVSR_LN_FOLD", "1") != "0"
        # MB of the layer's cross-attention K/V that the self-attention output projection (two launches before the source
        # attention) asks the L2 to fetch (0 = off); the HBM stream then overlaps the latency-bound projections in between
        self.kv_prefetch_mb = float(os.environ.get("AVSR_KV_PREFETCH_MB", "0"))
        # the projection two launches before a self-attention asks the L2 for that layer's dense K/V history (AVSR_SELF_KV_PREFETCH)
        self.self_kv_prefetch = os.environ.get("AVSR_SELF_KV_PREFETCH", "0") == "1"
        if self.proj not in ("cluster", "splitk"):
            raise RuntimeError(f"AVSR_PROJ must be cluster or splitk, got {self.proj!r}")
        self.graph_launches = 0       # kernels launched through graph replays (bench.py adds them to gpu_launches)
        self.last_session = None
        self.last_sessions = []
        self.last_stop_positions = []
        self.last_positions_queued = 0
        self._sessions = {}
        self._streams = []
        # independent groups of utterances decoded concurrently on separate streams (AVSR_DECODE_GROUPS).  Default 1: on B200
        # two concurrent chains measured SLOWER (449 vs 373 ms per 32-utterance pass; 3 groups 524, 4 groups 587): the
        # projection launches of the chains each want one CTA with ~200 KB of shared memory on every SM and serialise.
        self.n_groups = max(1, int(os.environ.get("AVSR_DECODE_GROUPS", "1")))
        # AVSR_FUSE_STEP=1: the fusion kernel also closes the position (53 launches per position instead of 54).  Measured neutral
        # (295.0 vs 294.6 ms per pass): the one-thread step kernel already overlaps its neighbours through programmatic dependent
        # launch, so the default stays the separate launch.
        self.fuse_step = os.environ.get("AVSR_FUSE_STEP", "0") != "0"
        # query merge: the source-attention query projection of every layer rides along with q | k | v and the attention-output
        # projection (7 launches per layer instead of 8: 48 per position); AVSR_QUERY_MERGE=0 = its own launch
        self.query_merge = os.environ.get("AVSR_QUERY_MERGE", "1") != "0"
        L.load()
        if self.proj == "cluster" and (self.n_groups > 1 or "AVSR_SM_BUDGET" in os.environ):
            # concurrent chains: plan every projection for its share of the SMs so that the chains' clusters are co-resident
            # (AVSR_SM_BUDGET with one chain: consecutive launches of the chain co-resident, dev experiment)
            sms = torch.cuda.get_device_properties(self.device).multi_processor_count
            L.check(L.load().avsr_dec_proj_set_sm_budget(int(os.environ.get("AVSR_SM_BUDGET", sms // self.n_groups))), "avsr_dec_proj_set_sm_budget")
            L.launch_count -= 1

    # ------------------------------------------------------------------------------------------ session buffers
    T_BUCKET = 16        # sessions (buffers + the captured graphs) are shared by all batches whose longest utterance rounds up to the same multiple
    # utterance slots per session: fine enough that a batch wastes < 15 % of its rows on parked slots (every projection works on
    # all rows of the session), coarse enough that the batches of an evaluation plan share a handful of sessions
    B_BUCKETS = (1, 2, 4, 8, 12, 16, 24, 32, 40, 48, 64, 80, 96, 112, 128, 144, 160, 192, 224, 256, 288, 320, 384, 448, 512, 640, 768, 1024)
    SESSION_BYTES = int(os.environ.get("AVSR_SESSION_GB", "96")) << 30     # retained sessions are evicted least-recently-used beyond this

    @classmethod
    def bucket_B(cls, B: int) -> int:
        """Utterance slots of the session that serves a batch of B utterances: the next bucket size (unused slots are parked
        with n_run = 0, every kernel skips them), so the batches of an evaluation plan share a handful of sessions and graphs."""
        for b in cls.B_BUCKETS:
            if b >= B:
                return b
        return -(-B // 256) * 256

    def _session(self, B: int, tmax: int, F: int, slot: int = 0, no_end_detect: bool = False):
        """Buffers and CUDA graphs for batches of up to bucket_B(B) utterances of at most `tmax` frames (rounded up to a
        multiple of T_BUCKET): everything is sized for that capacity, so mixed batches of an evaluation run reuse a few
        sessions instead of allocating and capturing one per (B, lengths) combination.  Sessions are kept least-recently-used
        within SESSION_BYTES (AVSR_SESSION_GB)."""
        B = self.bucket_B(B)
        tmax = -(-tmax // self.T_BUCKET) * self.T_BUCKET
        F = B * tmax                       # capacity in frames; a batch uses the first sum(lengths) of them
        key = (B, tmax, slot, bool(no_end_detect))     # slot: concurrent groups of one decode_batch call never share buffers
        s = self._sessions.pop(key, None)
        if s is not None:
            self._sessions[key] = s        # most recently used last
            return s
        dev, beam, S, V = self.device, self.beam_size, self.pre_beam_size, self.n_vocab
        R = B * beam
        nl = self.w.n_layers
        lmax = tmax + 1
        i32 = lambda *shape: torch.zeros(*shape, dtype=torch.int32, device=dev)
        f32 = lambda *shape: torch.zeros(*shape, dtype=torch.float32, device=dev)
        s = dict(B=B, R=R, tmax=tmax, lmax=lmax, F=F)
        s["utt_T"], s["utt_off"], s["utt_maxlen"] = i32(B), i32(B), i32(B)
        s["step"], s["any_running"] = i32(1), i32(1)
        s["ticket"] = i32(1)
        s["tq"] = f32(R, 1024)                     # query merge: x (g2 . Wq)^T of the layer input, waiting for the attention-output projection                       # avsr_beam_fuse_topk_advance_step: CTAs that have finished the position
        # host copies of any_running (asynchronous poll, double-buffered: a flag is read one replay late)
        s["poll"] = [torch.zeros(1, dtype=torch.int32).pin_memory() for _ in range(2)]
        s["poll_event"] = [torch.cuda.Event(), torch.cuda.Event()]
        s["n_run"], s["row_active"], s["last_tok"], s["rprev_idx"] = i32(B), i32(R), i32(R), i32(R)
        for k in ("score", "dec_sc", "ctc_sc", "s_prev", "rsum_last"):
            s[k] = f32(R)
        s["anc"] = torch.zeros(2, R, lmax, dtype=torch.uint8, device=dev)
        s["hist_tok"], s["hist_prev"], s["run2j"] = i32(B, tmax, beam), i32(B, tmax, beam), i32(B, tmax, beam)
        cap = beam * (tmax + 1)
        s["cap"] = cap
        s["n_ended"], s["done"], s["overflow"] = i32(B), i32(B), i32(1)
        s["end_step"], s["end_j"], s["end_len"] = i32(B, cap), i32(B, cap), i32(B, cap)
        s["end_score"], s["end_dec"], s["end_ctc"] = f32(B, cap), f32(B, cap), f32(B, cap)
        s["best_len"], s["best_all"] = f32(B, tmax + 4), f32(B)
        # activations of one decoder step
        s["x"], s["a"], s["att"] = f32(R, 1024), f32(R, 1024), f32(R, 1024)
        s["ffn"] = f32(R, 3072)
        s["dec_logp"] = f32(R, V)
        s["part_ids"], s["psi"] = i32(R, S), f32(R, S)
        if self.dec_only:
            s["ctc_full"] = f32(R, V)                 # the dropped CTC scorer contributes 0 * 0
            s["rc_last"], s["rc_chain"], s["rc_tok"] = i32(R), i32(R), i32(R)
        if self.ctc_only:
            lib0 = L.load()
            s["ctc_full"] = f32(R, V)
            s["rc_last"], s["rc_chain"], s["rc_tok"] = i32(R), i32(R), i32(R)
            s["iota"] = torch.arange(R, dtype=torch.int32, device=dev)
            ncg, ts = C.c_int(0), C.c_int(0)
            L.check(lib0.avsr_ctc_prefix_full_plan(B, V, C.byref(ncg), C.byref(ts)), "avsr_ctc_prefix_full_plan")
            s["fpart"] = torch.empty(B, ts.value, beam, V, dtype=torch.float32, device=dev)
            s["ftick"] = i32(B, ncg.value)
            s["want_probs"] = True
        # self-attention caches, one contiguous span per (utterance, head): keys transposed in 32-byte groups
        # [layer][utt][head][8][pos*beam+slot][8], values [layer][utt][head][pos*beam+slot][64] (csrc/dec_attn.cu)
        s["kc"] = torch.empty(nl, B, 16, 8, lmax * beam, 8, dtype=torch.float32, device=dev)
        s["vc"] = torch.empty(nl, B, 16, lmax * beam, 64, dtype=torch.float32, device=dev)
        # dense copy of the converged history prefix (all live hyps share the ancestor): consecutive positions = consecutive rows
        s["kd"] = torch.empty(nl, B, 16, 8, lmax, 8, dtype=torch.float32, device=dev)
        s["vd"] = torch.empty(nl, B, 16, lmax, 64, dtype=torch.float32, device=dev)
        s["conv_len"] = i32(2, B)
        s["r_buf"] = torch.empty(2, R * S, tmax, 2, dtype=torch.float32, device=dev)
        lib = L.load()
        shapes = ((3072, 1024), (1024, 1024), (1024, 3072), (V, 1024))
        if self.precision == "bf16x3":
            # activations of the step in compact bf16x3 form [a1|a2|a3] (operands of csrc/gemm_x3.cu)
            s["a3"] = torch.zeros(R, 3 * 1024, dtype=torch.bfloat16, device=dev)
            s["att3"] = torch.zeros(R, 3 * 1024, dtype=torch.bfloat16, device=dev)
            s["ffn3"] = torch.zeros(R, 3 * 3072, dtype=torch.bfloat16, device=dev)
            s["x6"] = torch.empty(F, 6 * 1024, dtype=torch.bfloat16, device=dev)
            # cluster projections: finished q | k | v and cross-attention queries, the (mean, M2) of every 128-column tile of
            # the residual stream (consumed by the LayerNorm of the next projection), dense logits
            s["qkv"], s["q2"] = f32(R, 3072), f32(R, 1024)
            s["stats"] = f32(8, R, 2)
            s["x3"] = torch.zeros(R, 3 * 1024, dtype=torch.bfloat16, device=dev)      # the residual stream as compact bf16x3 (folded LayerNorm)
            s["logits"] = f32(R, V)
            n_part = max(lib.avsr_gemm_x3_splits(R, n, k) * R * n for n, k in shapes)
            s["gbar"] = torch.zeros(2, dtype=torch.int32, device=dev)          # grid-barrier state of the fused projections
            # the fused form needs one CTA per work item, all resident: tiles x splits <= SMs (true up to R = 128 rows)
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            s["fused_ok"] = {(n, k): (-(-n // 128)) * (-(-R // 128)) * lib.avsr_gemm_x3_splits(R, n, k) <= sms and R <= 128
                             for n, k in shapes}
        else:
            n_part = max(lib.avsr_sgemm_skinny_splits(R, n, k) * R * n for n, k in shapes)
        s["part"] = torch.empty(n_part, dtype=torch.float32, device=dev)
        # chained projections (csrc/gemm_x3.cu): consecutive projections alternate between two partial-sum buffers, and a
        # pair of device counters releases the operand rows that a launch finishes for itself
        s["part2"] = torch.empty(n_part, dtype=torch.float32, device=dev) if self.precision == "bf16x3" else None
        s["ready"] = torch.zeros(2, dtype=torch.int32, device=dev)
        # per-utterance precomputed tensors
        s["ldp"] = (V + 31) // 32 * 32              # posterior row pitch: 128-byte aligned rows (V = 5049 -> 5056)
        s["logp"] = torch.zeros(F, s["ldp"], dtype=torch.float32, device=dev)
        # CTC-only search: the posteriors themselves, exp(logp) once per batch, so that the full-vocabulary kernel (run at every
        # position over the whole block) streams them without an exponential per element
        s["probs"] = torch.zeros(F, s["ldp"], dtype=torch.float32, device=dev) if s.get("want_probs") else None
        s["ckv"] = torch.empty(F, nl * 2 * 1024, dtype=torch.float32, device=dev)
        # cross-attention K/V, head-major: [layer][k|v][head][frame][64] (one contiguous span per (utterance, head))
        s["ckv_t"] = torch.empty(nl, 2, 16, F, 64, dtype=torch.float32, device=dev)
        st = L.BeamState()
        st.B, st.beam, st.S, st.V, st.lmax, st.tmax = B, beam, S, V, lmax, tmax
        st.blank, st.eos, st.cap = self.w.blank, self.eos, cap
        st.no_end_detect = 1 if no_end_detect else 0        # baked into the captured graphs, hence part of the session key
        st.utt_maxlen = s["utt_maxlen"].data_ptr()
        for name in ("utt_T", "step", "n_run", "row_active", "last_tok", "score", "dec_sc", "ctc_sc", "s_prev", "rprev_idx", "anc",
                     "hist_tok", "hist_prev", "run2j", "n_ended", "end_step", "end_j", "end_score", "end_dec", "end_ctc", "end_len",
                     "best_len", "best_all", "done", "overflow"):
            setattr(st, name, s[name].data_ptr())
        st.d_end = D_END
        s["state"] = st
        s["graph"] = s["graph1"] = None
        s["bytes"] = sum(v.numel() * v.element_size() for v in s.values() if isinstance(v, torch.Tensor))
        self._sessions[key] = s
        total = sum(x["bytes"] for x in self._sessions.values())
        for k in list(self._sessions):
            if total <= self.SESSION_BYTES or k == key:
                continue
            total -= self._sessions.pop(k)["bytes"]           # oldest first; its tensors and graphs are freed with it
        return s

    # ------------------------------------------------------------------------------------------ one decode step
    def _proj(self, s, key_a, lay, name, N, K):
        """partial sums of act[R,K] @ W[N,K]^T into s['part']; returns the number of K splits."""
        lib = L.load()
        R = s["R"]
        if "gemm" in self._skip:
            return lib.avsr_gemm_x3_splits(R, N, K) if self.precision == "bf16x3" else lib.avsr_sgemm_skinny_splits(R, N, K)
        if self.precision == "bf16x3":
            ns = lib.avsr_gemm_x3_splits(R, N, K)
            L.check(lib.avsr_gemm_x3_splitk(L.ptr(s[key_a + "3"]), L.ll(3 * K), L.ptr(lay[name + "3"]), L.ll(3 * K), R, N, K,
                                            L.ptr(s["part"]), L.stream()), "avsr_gemm_x3_splitk")
        else:
            a = s[key_a]
            ns = lib.avsr_sgemm_skinny_splits(R, N, K)
            L.check(lib.avsr_sgemm_skinny(L.ptr(a), L.ll(K), L.ptr(lay[name]), L.ll(K), R, N, K, L.ptr(s["part"]), ns, L.stream()),
                    "avsr_sgemm_skinny")
        return ns

    def _linear(self, s, key_a, lay, name, N, K, bias, act=L.ACT_NONE, residual=None, out=None, ln=None, key_out=None, prefetch=None):
        """One nn.Linear of the step plus its glue: act[R,K] @ W[N,K]^T + bias ; act ; + residual -> out ; LayerNorm -> the
        operand of the next projection (fp32 s[key_out] on the CUDA-core path, compact bf16x3 s[key_out + '3'] on the
        tensor-core path).  Split-K projection + row-wise epilogue kernel.  ``fuse_epilogue=True`` runs both in ONE launch
        (projection, grid barrier, epilogue; avsr_gemm_x3_fused); it measured SLOWER on B200 (1.41 vs 1.19 ms per position):
        two projection CTAs cannot share an SM (197 KB of shared memory each), so back-to-back fused projections lose the
        weight prefetch that a programmatic dependent launch gets behind a small epilogue kernel."""
        lib = L.load()
        R = s["R"]
        g, b = (ln if ln is not None else (None, None))
        tc = self.precision == "bf16x3"
        ln_out = s[key_out] if (key_out is not None and ln is not None and not tc) else None
        if key_out is not None and ln is None and not tc:
            out = s[key_out]
        split = s[key_out + "3"] if (key_out is not None and tc) else None
        if tc and self.fuse_epilogue and s["fused_ok"][(N, K)]:
            L.check(lib.avsr_gemm_x3_fused(L.ptr(s[key_a + "3"]), L.ll(3 * K), L.ptr(lay[name + "3"]), L.ll(3 * K), R, N, K, L.ptr(s["part"]),
                                           L.ptr(bias), act, L.ptr(residual), L.ll(1024), L.ptr(out), L.ll(N), L.ptr(g), L.ptr(b),
                                           C.c_float(1e-12), L.ptr(ln_out), L.ll(1024), L.ptr(s["row_active"]), L.ptr(split),
                                           L.ptr(s["gbar"]), L.stream()), "avsr_gemm_x3_fused")
            return
        ns = self._proj(s, key_a, lay, name, N, K)
        if "epi" in self._skip:
            return
        pf_bytes = prefetch.numel() * prefetch.element_size() if (prefetch is not None and self.l2_prefetch) else 0
        L.check(lib.avsr_splitk_epilogue_pf(L.ptr(s["part"]), ns, R, N, L.ptr(bias), act, L.ptr(residual), L.ll(1024),
                                            L.ptr(out), L.ll(N), L.ptr(g), L.ptr(b), C.c_float(1e-12), L.ptr(ln_out), L.ll(1024),
                                            L.ptr(s["row_active"]), L.ptr(split), L.ptr(prefetch) if pf_bytes else None, L.ll(pf_bytes),
                                            L.stream()), "avsr_splitk_epilogue")

    def _step_chain(self, s):
        """The bf16x3 position with chained projections: every nn.Linear whose input is the row-wise glue of the previous one
        (bias / residual / LayerNorm / ReLU, decoder_layer.py:58-121) finishes those rows inside its own launch
        (avsr_gemm_x3_chain), so a position is 54 launches instead of 78."""
        lib = L.load()
        w = self.w
        R, beam, S, V, lmax = s["R"], self.beam_size, self.pre_beam_size, self.n_vocab, s["lmax"]
        st = L.stream
        nl = w.n_layers
        parts = (s["part"], s["part2"])
        state = {"cur": 0, "rq": 0}

        def proj(key_a, w3, N, K, pro=None):
            """act[R,K] @ W^T -> partial sums in the other buffer; pro = row glue of the previous projection to finish first:
            dict(N, bias, act, residual, out, ln, split)."""
            prev, cur = parts[state["cur"]], parts[state["cur"] ^ 1]
            state["cur"] ^= 1
            if pro is None:
                L.check(lib.avsr_gemm_x3_splitk(L.ptr(s[key_a]), L.ll(3 * K), L.ptr(w3), L.ll(3 * K), R, N, K, L.ptr(cur), st()),
                        "avsr_gemm_x3_splitk")
            else:
                g, b = pro.get("ln") or (None, None)
                ns_prev = lib.avsr_gemm_x3_splits(R, pro["N"], pro["K"])
                L.check(lib.avsr_gemm_x3_chain(L.ptr(s[key_a]), L.ll(3 * K), L.ptr(w3), L.ll(3 * K), R, N, K, L.ptr(cur),
                                               L.ptr(prev), ns_prev, pro["N"], L.ptr(pro["bias"]), pro.get("act", L.ACT_NONE),
                                               L.ptr(pro.get("residual")), L.ll(1024), L.ptr(pro.get("out")), L.ll(pro["N"]), L.ptr(g),
                                               L.ptr(b), C.c_float(1e-12), L.ptr(s["row_active"]), L.ptr(s[key_a]), L.ptr(s["ready"]),
                                               state["rq"], st()), "avsr_gemm_x3_chain")
                state["rq"] ^= 1
            return cur, lib.avsr_gemm_x3_splits(R, N, K)

        l0 = w.layers[0]
        L.check(lib.avsr_dec_cache_promote(L.ptr(s["kc"]), L.ptr(s["vc"]), L.ptr(s["kd"]), L.ptr(s["vd"]), nl, L.ptr(s["anc"]), lmax,
                                           L.ptr(s["n_run"]), beam, R, L.ptr(s["step"]), L.ptr(s["conv_len"]), st()), "avsr_dec_cache_promote")
        L.check(lib.avsr_dec_embed_ln(L.ptr(w.embed), L.ptr(w.pe), L.ptr(s["last_tok"]), L.ptr(s["n_run"]), beam, R, L.ptr(s["step"]),
                                      L.ptr(l0["n1_g"]), L.ptr(l0["n1_b"]), C.c_float(1e-12), L.ptr(s["x"]), None, L.ptr(s["a3"]), st()),
                "avsr_dec_embed_ln")
        att_split = L.ptr(s["att3"])
        pending = None           # row glue of the last projection that the next chained projection has to finish
        for li, lay in enumerate(w.layers):
            # self-attention (decoder_layer.py:82-93); q | k | v come straight from the split-K partial sums
            part, ns = proj("a3", lay["wqkv3"], 3072, 1024, pending)
            L.check(lib.avsr_dec_attn_step(0, L.ptr(part), L.ll(3072), ns, L.ptr(lay["bqkv"]), L.ptr(s["kc"][li]), L.ptr(s["vc"][li]),
                                           L.ptr(s["anc"]), lmax, L.ptr(s["n_run"]), L.ptr(s["utt_off"]), L.ptr(s["utt_T"]), beam, R,
                                           L.ptr(s["step"]), None, L.ll(0), att_split, L.ptr(s["kd"][li]), L.ptr(s["vd"][li]),
                                           L.ptr(s["conv_len"]), st()), "avsr_dec_attn_step(self)")
            proj("att3", lay["wo3"], 1024, 1024)
            # source attention (decoder_layer.py:97-107): its query projection first finishes x += wo(att) + bo ; LayerNorm n2
            part, ns = proj("a3", lay["wq23"], 1024, 1024, dict(N=1024, K=1024, bias=lay["bo"], residual=s["x"], out=s["x"],
                                                                ln=(lay["n2_g"], lay["n2_b"])))
            L.check(lib.avsr_dec_attn_step(1, L.ptr(part), L.ll(1024), ns, L.ptr(lay["bq2"]), L.ptr(s["ckv_t"][li, 0]),
                                           L.ptr(s["ckv_t"][li, 1]), None, lmax, L.ptr(s["n_run"]), L.ptr(s["utt_off"]), L.ptr(s["utt_T"]),
                                           beam, R, L.ptr(s["step"]), None, L.ll(s["F"]), att_split, None, None, None, st()),
                    "avsr_dec_attn_step(src)")
            proj("att3", lay["wo23"], 1024, 1024)
            # feed-forward (decoder_layer.py:112-116)
            proj("a3", lay["w13"], 3072, 1024, dict(N=1024, K=1024, bias=lay["bo2"], residual=s["x"], out=s["x"],
                                                    ln=(lay["n3_g"], lay["n3_b"])))
            proj("ffn3", lay["w23"], 1024, 3072, dict(N=3072, K=1024, bias=lay["b1"], act=L.ACT_RELU))
            nxt = (w.layers[li + 1]["n1_g"], w.layers[li + 1]["n1_b"]) if li + 1 < nl else (w.after_g, w.after_b)
            pending = dict(N=1024, K=3072, bias=lay["b2"], residual=s["x"], out=s["x"], ln=nxt)
        # output layer (decoder.py:176-181) finishes the last feed-forward + after_norm first
        part, ns = proj("a3", w.out_w3, V, 1024, pending)
        self._tail(s, part, ns)

    def _tail(self, s, part, ns):
        """log_softmax + pre-beam, CTC prefix scores, fusion / top-k / bookkeeping (batch_beam_search.py:222-349)."""
        lib = L.load()
        w = self.w
        R, beam, S, V = s["R"], self.beam_size, self.pre_beam_size, self.n_vocab
        st = L.stream
        if "tail" in self._skip:
            return
        # one launch: output-layer log_softmax + pre-beam top-S and the CTC prefix scores of those candidates
        L.check(lib.avsr_dec_tail(L.ptr(part), ns, L.ptr(w.out_b), L.ptr(s["dec_logp"]), L.ptr(s["part_ids"]), L.ptr(s["logp"]), V, s["ldp"],
                                  w.blank, L.ptr(s["utt_off"]), L.ptr(s["utt_T"]), L.ptr(s["n_run"]), beam, R, S, L.ptr(s["last_tok"]),
                                  L.ptr(s["rprev_idx"]), L.ptr(s["r_buf"]), s["tmax"], L.ptr(s["step"]), L.ptr(s["psi"]),
                                  L.ptr(s["rsum_last"]), st()), "avsr_dec_tail")
        if "advance" in self._skip:
            return
        if self.fuse_step:
            # fusion / top-k / bookkeeping AND the end of the position (liveness count, step + 1) in one launch
            L.check(lib.avsr_beam_fuse_topk_advance_step(C.byref(s["state"]), L.ptr(s["dec_logp"]), L.ptr(s["part_ids"]), L.ptr(s["psi"]),
                                                         L.ptr(s["rsum_last"]), C.c_float(self.w_dec), C.c_float(self.w_ctc),
                                                         L.ptr(s["any_running"]), L.ptr(s["ticket"]), st()), "avsr_beam_fuse_topk_advance_step")
            return
        L.check(lib.avsr_beam_fuse_topk_advance(C.byref(s["state"]), L.ptr(s["dec_logp"]), L.ptr(s["part_ids"]), L.ptr(s["psi"]),
                                                L.ptr(s["rsum_last"]), C.c_float(self.w_dec), C.c_float(self.w_ctc), st()),
                "avsr_beam_fuse_topk_advance")
        L.check(lib.avsr_beam_step_advance(L.ptr(s["step"]), L.ptr(s["n_run"]), s["B"], L.ptr(s["any_running"]), st()),
                "avsr_beam_step_advance")

    def _step_ctc_only(self, s):
        """One position of the single-scorer search (ctc_weight = 1.0): full-vocabulary CTC prefix scores of every live hyp
        (the HBM-bound kernel), fusion / top-k / bookkeeping over all n_h * V entries, then the forward variables of the
        survivors recomputed on their chosen token (CTCPrefixScorer.select_state, scorers/ctc.py:40-63, without ever
        materialising r[T, 2, n_h, V])."""
        lib = L.load()
        w = self.w
        R, beam, V = s["R"], self.beam_size, self.n_vocab
        st = L.stream
        L.check(lib.avsr_ctc_prefix_full_probs(L.ptr(s["logp"]), L.ptr(s["probs"]), V, s["ldp"], w.blank, self.eos, L.ptr(s["utt_off"]), L.ptr(s["utt_T"]),
                                         L.ptr(s["n_run"]), beam, s["B"], 1, L.ptr(s["last_tok"]), L.ptr(s["iota"]), L.ptr(s["r_buf"]),
                                         s["tmax"], L.ptr(s["step"]), L.ptr(s["psi"]), L.ptr(s["ctc_full"]), L.ptr(s["fpart"]),
                                         L.ptr(s["ftick"]), st()), "avsr_ctc_prefix_full_probs")
        L.check(lib.avsr_beam_fuse_topk_advance_full(C.byref(s["state"]), L.ptr(s["dec_logp"]), L.ptr(s["ctc_full"]), C.c_float(0.0),
                                                     C.c_float(1.0), L.ptr(s["rc_last"]), L.ptr(s["rc_chain"]), L.ptr(s["rc_tok"]), st()),
                "avsr_beam_fuse_topk_advance_full")
        L.check(lib.avsr_ctc_prefix_prebeam(L.ptr(s["logp"]), V, s["ldp"], w.blank, L.ptr(s["utt_off"]), L.ptr(s["utt_T"]), L.ptr(s["n_run"]),
                                            beam, R, 1, L.ptr(s["rc_last"]), L.ptr(s["rc_tok"]), L.ptr(s["rc_chain"]), L.ptr(s["r_buf"]),
                                            s["tmax"], L.ptr(s["step"]), L.ptr(s["psi"]), L.ptr(s["rsum_last"]), st()),
                "avsr_ctc_prefix_prebeam(recompute)")
        L.check(lib.avsr_beam_step_advance(L.ptr(s["step"]), L.ptr(s["n_run"]), s["B"], L.ptr(s["any_running"]), st()),
                "avsr_beam_step_advance")

    def _step(self, s):
        """Decoder.batch_score + CTC partial scoring + fusion/top-k/bookkeeping for position *step."""
        if self.ctc_only:
            return self._step_ctc_only(s)
        if self.precision == "bf16x3" and self.chain and not self.fuse_epilogue and not self.dec_only and not self._skip - {"advance", "tail"}:
            return self._step_chain(s)
        part, ns = self._decoder_layers(s)
        if self.dec_only:
            return self._tail_dec_only(s, part, ns)
        self._tail(s, part, ns)

    def _tail_dec_only(self, s, part, ns):
        """Attention-only search (ctc_weight = 0): log_softmax of the output layer, then the full-vocabulary form of the
        fusion / top-k / bookkeeping kernel with a zero CTC matrix and weight (batch_beam_search.py:208-285 without partial
        scorers)."""
        lib = L.load()
        w = self.w
        R, beam, V = s["R"], self.beam_size, self.n_vocab
        st = L.stream
        L.check(lib.avsr_dec_logits_lsm_topk(L.ptr(part), ns, R, V, L.ptr(w.out_b), L.ptr(s["n_run"]), beam, L.ptr(s["dec_logp"]),
                                             L.ptr(s["part_ids"]), 1, st()), "avsr_dec_logits_lsm_topk")
        L.check(lib.avsr_beam_fuse_topk_advance_full(C.byref(s["state"]), L.ptr(s["dec_logp"]), L.ptr(s["ctc_full"]), C.c_float(self.w_dec),
                                                     C.c_float(0.0), L.ptr(s["rc_last"]), L.ptr(s["rc_chain"]), L.ptr(s["rc_tok"]), st()),
                "avsr_beam_fuse_topk_advance_full")
        L.check(lib.avsr_beam_step_advance(L.ptr(s["step"]), L.ptr(s["n_run"]), s["B"], L.ptr(s["any_running"]), st()),
                "avsr_beam_step_advance")

    def _cproj(self, s, W3, N, K, a3=None, ln=None, bias=None, act=L.ACT_NONE, residual=None, out=None, ldo=None, split=None, stats_out=False,
               nxt=None):
        """One cluster projection (avsr_dec_proj): operand = compact bf16x3 rows `a3` or LayerNorm(x) with ln = (gamma, beta).
        nxt: the weights the NEXT projection of the chain streams; this launch asks the L2 to fetch them."""
        if "gemm" in self._skip:
            return
        lib = L.load()
        R = s["R"]
        g, b = ln if ln is not None else (None, None)
        pf = nxt if (nxt is not None and self.weight_prefetch) else None
        L.check(lib.avsr_dec_proj(L.ptr(a3), L.ll(3 * K), L.ptr(s["x"]) if ln is not None else None, L.ll(1024),
                                  L.ptr(s["stats"]) if ln is not None else None, L.ptr(g), L.ptr(b), C.c_float(1e-12), L.ptr(W3), L.ll(3 * K),
                                  R, N, K, L.ptr(bias), act, L.ptr(residual), L.ll(1024), L.ptr(out), L.ll(N if ldo is None else ldo),
                                  L.ptr(split), L.ptr(s["stats"]) if stats_out else None, L.ptr(pf),
                                  L.ll(pf.numel() * pf.element_size() if pf is not None else 0), L.stream()), "avsr_dec_proj")

    def _cattn(self, s, mode, q, ldq, kc, vc, dense, li, nxt):
        """Self (mode 0) / source (mode 1) attention of the position; also starts the L2 fetch of the next projection's weights."""
        if ("self" if mode == 0 else "cross") in self._skip:
            return
        lib = L.load()
        beam, R, lmax = self.beam_size, s["R"], s["lmax"]
        pf = nxt if self.weight_prefetch else None
        use_dense = mode == 0 and dense
        L.check(lib.avsr_dec_attn_step_pf(mode, L.ptr(q), L.ll(ldq), 0, None, L.ptr(kc), L.ptr(vc), L.ptr(s["anc"]) if mode == 0 else None, lmax,
                                          L.ptr(s["n_run"]), L.ptr(s["utt_off"]), L.ptr(s["utt_T"]), beam, R, L.ptr(s["step"]), None,
                                          L.ll(0 if mode == 0 else s["F"]), L.ptr(s["att3"]), L.ptr(s["kd"][li]) if use_dense else None,
                                          L.ptr(s["vd"][li]) if use_dense else None, L.ptr(s["conv_len"]) if use_dense else None, L.ptr(pf),
                                          L.ll(pf.numel() * pf.element_size() if pf is not None else 0), L.stream()),
                "avsr_dec_attn_step(self)" if mode == 0 else "avsr_dec_attn_step(src)")

    def _cfold(self, s, W3g, u, c, N, act=L.ACT_NONE, out=None, ldo=None, split=None, nxt=None):
        """Projection of LayerNorm(x) with the LayerNorm folded in (avsr_dec_proj_folded): operand = the raw rows s['x3']."""
        if "gemm" in self._skip:
            return
        lib = L.load()
        pf = nxt if (nxt is not None and self.weight_prefetch) else None
        L.check(lib.avsr_dec_proj_folded(L.ptr(s["x3"]), L.ll(3 * 1024), L.ptr(s["stats"]), C.c_float(1e-12), L.ptr(u), L.ptr(c), L.ptr(W3g),
                                         L.ll(3 * 1024), s["R"], N, 1024, act, None, L.ll(1024), L.ptr(out), L.ll(N if ldo is None else ldo),
                                         L.ptr(split), None, L.ptr(pf), L.ll(pf.numel() * pf.element_size() if pf is not None else 0),
                                         L.stream()), "avsr_dec_proj_folded")

    def _decoder_layers_cluster(self, s, dense: bool = True):
        """Decoder.forward_one_step up to the output layer with the cluster projections: 8 launches per layer (6 projections + 2
        attentions).  The residual stream x stays fp32; every projection that updates it also leaves the row as compact bf16x3
        (x3) and the per-tile LayerNorm statistics, and the projection that consumes LayerNorm(x) takes x3 through TMA with the
        LayerNorm folded into its weights and epilogue (ln_fold) or normalises while it stages its operand.  Every kernel asks
        the L2 for the weights of the projection that follows it."""
        lib = L.load()
        w = self.w
        R, beam, V, lmax = s["R"], self.beam_size, self.n_vocab, s["lmax"]
        st = L.stream
        l0 = w.layers[0]
        nl = w.n_layers
        fold = self.ln_fold
        x3 = s["x3"] if fold else None
        if dense:
            L.check(lib.avsr_dec_cache_promote(L.ptr(s["kc"]), L.ptr(s["vc"]), L.ptr(s["kd"]), L.ptr(s["vd"]), w.n_layers, L.ptr(s["anc"]), lmax,
                                               L.ptr(s["n_run"]), beam, R, L.ptr(s["step"]), L.ptr(s["conv_len"]), st()), "avsr_dec_cache_promote")
        merge_all = fold and self.query_merge and "gemm" not in self._skip
        if merge_all:
            # layer 0 takes its LayerNorm folded like the others: the embedding is written raw, with its tile statistics
            L.check(lib.avsr_dec_embed_raw(L.ptr(w.embed), L.ptr(w.pe), L.ptr(s["last_tok"]), L.ptr(s["n_run"]), beam, R, L.ptr(s["step"]),
                                           L.ptr(s["x"]), L.ptr(s["x3"]), L.ptr(s["stats"]), st()), "avsr_dec_embed_raw")
        else:
            L.check(lib.avsr_dec_embed_ln(L.ptr(w.embed), L.ptr(w.pe), L.ptr(s["last_tok"]), L.ptr(s["n_run"]), beam, R, L.ptr(s["step"]),
                                          L.ptr(l0["n1_g"]), L.ptr(l0["n1_b"]), C.c_float(1e-12), L.ptr(s["x"]), None, L.ptr(s["a3"]), st()),
                    "avsr_dec_embed_ln")
        for li, lay in enumerate(w.layers):
            # self-attention (decoder_layer.py:82-93): q | k | v finished by the projection, bias included
            merge = merge_all
            if merge:
                # q | k | v of LayerNorm1(x) and, riding along, tq = x (g2 . Wq)^T for the source-attention query
                pf = lay["wcat2_3"] if self.weight_prefetch else None
                L.check(lib.avsr_dec_proj_dual(L.ptr(s["x3"]), L.ll(3 * 1024), L.ptr(s["stats"]), C.c_float(1e-12), L.ptr(lay["ucat1"]),
                                               L.ptr(lay["ccat1"]), L.ptr(lay["wcat1_3"]), L.ll(3 * 1024), R, 4096, 1024, 3072, L.ACT_NONE,
                                               None, L.ll(1024), L.ptr(s["qkv"]), L.ll(3072), None, L.ll(1024), L.ptr(s["tq"]), L.ll(1024),
                                               None, L.ptr(pf), L.ll(pf.numel() * pf.element_size() if pf is not None else 0), st()),
                        "avsr_dec_proj_dual(qkv)")
            elif li == 0:
                self._cproj(s, lay["wqkv3"], 3072, 1024, a3=s["a3"], bias=lay["bqkv"], out=s["qkv"])
            elif fold:
                self._cfold(s, lay["wqkv3g"], lay["uqkv"], lay["cqkv"], 3072, out=s["qkv"])
            else:
                self._cproj(s, lay["wqkv3"], 3072, 1024, ln=(lay["n1_g"], lay["n1_b"]), bias=lay["bqkv"], out=s["qkv"])
            self._cattn(s, 0, s["qkv"], 3072, s["kc"][li], s["vc"][li], dense, li, lay["wcat2_3"] if merge else lay["wo3"])
            if self.kv_prefetch_mb > 0:
                ckv = s["ckv_t"][li]
                nbytes = min(int(self.kv_prefetch_mb * (1 << 20)), ckv.numel() * 4) // 16 * 16
                L.check(lib.avsr_dec_proj_also_prefetch(L.ptr(ckv), L.ll(nbytes)), "avsr_dec_proj_also_prefetch")
                L.launch_count -= 1
            if merge:
                # x += att Wo^T + bo (with the LayerNorm2 statistics) and q2raw = tq + att ((g2 . Wq) Wo)^T + (g2 . Wq) bo; the
                # source attention applies the LayerNorm's rstd / mean to q2raw itself
                L.check(lib.avsr_dec_proj_dual(L.ptr(s["att3"]), L.ll(3 * 1024), None, C.c_float(1e-12), None, L.ptr(lay["bcat2"]),
                                               L.ptr(lay["wcat2_3"]), L.ll(3 * 1024), R, 2048, 1024, 1024, L.ACT_NONE, L.ptr(s["x"]), L.ll(1024),
                                               L.ptr(s["x"]), L.ll(1024), L.ptr(s["tq"]), L.ll(1024), L.ptr(s["q2"]), L.ll(1024),
                                               L.ptr(s["stats"]), None, L.ll(0), st()), "avsr_dec_proj_dual(out)")
                if "cross" not in self._skip:
                    L.check(lib.avsr_dec_attn_fold_query(L.ptr(s["stats"]), L.ptr(lay["uq2"]), L.ptr(lay["cq2"]), C.c_float(1e-12)),
                            "avsr_dec_attn_fold_query")
                    L.launch_count -= 1
            else:
                self._cproj(s, lay["wo3"], 1024, 1024, a3=s["att3"], bias=lay["bo"], residual=s["x"], out=s["x"], split=x3, stats_out=True,
                            nxt=lay["wq23g"] if fold else lay["wq23"])
                # source attention (decoder_layer.py:97-107)
                if fold:
                    self._cfold(s, lay["wq23g"], lay["uq2"], lay["cq2"], 1024, out=s["q2"])
                else:
                    self._cproj(s, lay["wq23"], 1024, 1024, ln=(lay["n2_g"], lay["n2_b"]), bias=lay["bq2"], out=s["q2"])
            self._cattn(s, 1, s["q2"], 1024, s["ckv_t"][li, 0], s["ckv_t"][li, 1], dense, li, lay["wo23"])
            self._cproj(s, lay["wo23"], 1024, 1024, a3=s["att3"], bias=lay["bo2"], residual=s["x"], out=s["x"], split=x3, stats_out=True,
                        nxt=lay["w13g"] if fold else lay["w13"])
            # feed-forward (decoder_layer.py:112-116): ReLU(w_1 LN(x)) goes straight to the bf16x3 operand of w_2
            if fold:
                self._cfold(s, lay["w13g"], lay["u1"], lay["c1"], 3072, act=L.ACT_RELU, split=s["ffn3"], nxt=lay["w23"])
            else:
                self._cproj(s, lay["w13"], 3072, 1024, ln=(lay["n3_g"], lay["n3_b"]), bias=lay["b1"], act=L.ACT_RELU, split=s["ffn3"], nxt=lay["w23"])
            if li + 1 < nl:
                nxt = w.layers[li + 1]["wqkv3g"] if fold else w.layers[li + 1]["wqkv3"]
                if fold and self.query_merge:
                    nxt = w.layers[li + 1]["wcat1_3"]
            else:
                nxt = w.out_w3g if fold else w.out_w3
            if self.self_kv_prefetch and dense and li + 1 < nl:
                L.check(lib.avsr_dec_proj_prefetch_self_kv(L.ptr(s["kd"][li + 1]), L.ptr(s["vd"][li + 1]), lmax, s["B"] * 16, L.ptr(s["step"])),
                        "avsr_dec_proj_prefetch_self_kv")
                L.launch_count -= 1
            self._cproj(s, lay["w23"], 1024, 3072, a3=s["ffn3"], bias=lay["b2"], residual=s["x"], out=s["x"], split=x3, stats_out=True, nxt=nxt)
        # after_norm + output layer (decoder.py:176-181); the output bias is added by the softmax kernel that follows.  It
        # fetches the first projection of the NEXT position (only small kernels run in between)
        if fold:
            self._cfold(s, w.out_w3g, w.out_u, w.out_c, V, out=s["logits"], ldo=V, nxt=l0["wcat1_3"] if merge_all else l0["wqkv3"])
        else:
            self._cproj(s, w.out_w3, V, 1024, ln=(w.after_g, w.after_b), out=s["logits"], ldo=V, nxt=l0["wqkv3"])
        return s["logits"], 1

    def _decoder_layers(self, s, dense: bool = True):
        """Decoder.forward_one_step up to the output layer (decoder.py:153-181) for the rows of the session: embedding +
        positional encoding, six layers, after_norm, output projection.  Returns (partial logits [ns][R][V], ns).
        dense=False: no converged-prefix caches (the scorer plug-in API drives the kernels without the beam bookkeeping)."""
        # cluster projections serve up to 128 hypothesis rows (one row tile per cluster, everything co-resident in one wave);
        # larger batches (configs[2]-style batches of 100+ short utterances) amortise the launch chain over several row tiles
        # and run the persistent split-K kernels instead (measured on configs[2]: 792 vs 1090 ms per pass)
        if self.precision == "bf16x3" and self.proj == "cluster" and s["R"] <= 128 and not ("epi" in self._skip) and not self.fuse_epilogue:
            return self._decoder_layers_cluster(s, dense)
        lib = L.load()
        w = self.w
        R, beam, S, V, lmax = s["R"], self.beam_size, self.pre_beam_size, self.n_vocab, s["lmax"]
        st = L.stream
        tc = self.precision == "bf16x3"
        l0 = w.layers[0]
        # newly converged history positions -> dense self-attention caches (all layers)
        if dense:
            L.check(lib.avsr_dec_cache_promote(L.ptr(s["kc"]), L.ptr(s["vc"]), L.ptr(s["kd"]), L.ptr(s["vd"]), w.n_layers, L.ptr(s["anc"]), lmax,
                                               L.ptr(s["n_run"]), beam, R, L.ptr(s["step"]), L.ptr(s["conv_len"]), st()), "avsr_dec_cache_promote")
        L.check(lib.avsr_dec_embed_ln(L.ptr(w.embed), L.ptr(w.pe), L.ptr(s["last_tok"]), L.ptr(s["n_run"]), beam, R, L.ptr(s["step"]),
                                      L.ptr(l0["n1_g"]), L.ptr(l0["n1_b"]), C.c_float(1e-12), L.ptr(s["x"]),
                                      None if tc else L.ptr(s["a"]), L.ptr(s["a3"]) if tc else None, st()), "avsr_dec_embed_ln")
        nl = w.n_layers
        att_f32 = None if tc else L.ptr(s["att"])
        att_split = L.ptr(s["att3"]) if tc else None
        for li, lay in enumerate(w.layers):
            # self-attention (decoder_layer.py:82-93); the attention kernel sums the split-K partials of q | k | v itself
            ns = self._proj(s, "a", lay, "wqkv", 3072, 1024)
            if "self" not in self._skip:
                L.check(lib.avsr_dec_attn_step(0, L.ptr(s["part"]), L.ll(3072), ns, L.ptr(lay["bqkv"]), L.ptr(s["kc"][li]),
                                               L.ptr(s["vc"][li]), L.ptr(s["anc"]), lmax, L.ptr(s["n_run"]), L.ptr(s["utt_off"]),
                                               L.ptr(s["utt_T"]), beam, R, L.ptr(s["step"]), att_f32, L.ll(0), att_split,
                                               L.ptr(s["kd"][li]) if dense else None, L.ptr(s["vd"][li]) if dense else None,
                                               L.ptr(s["conv_len"]) if dense else None, st()),
                        "avsr_dec_attn_step(self)")
            # (its row epilogue also asks the L2 for this layer's cross K/V, which the source attention streams two kernels later)
            self._linear(s, "att", lay, "wo", 1024, 1024, lay["bo"], residual=s["x"], out=s["x"], ln=(lay["n2_g"], lay["n2_b"]), key_out="a",
                         prefetch=s["ckv_t"][li])
            # source attention over the precomputed K/V of the utterance's frames (decoder_layer.py:97-107)
            ns = self._proj(s, "a", lay, "wq2", 1024, 1024)
            ck, cv = s["ckv_t"][li, 0], s["ckv_t"][li, 1]
            if "cross" not in self._skip:
                L.check(lib.avsr_dec_attn_step(1, L.ptr(s["part"]), L.ll(1024), ns, L.ptr(lay["bq2"]), L.ptr(ck), L.ptr(cv), None, lmax,
                                               L.ptr(s["n_run"]), L.ptr(s["utt_off"]), L.ptr(s["utt_T"]), beam, R, L.ptr(s["step"]),
                                               att_f32, L.ll(s["F"]), att_split, None, None, None, st()), "avsr_dec_attn_step(src)")
            self._linear(s, "att", lay, "wo2", 1024, 1024, lay["bo2"], residual=s["x"], out=s["x"], ln=(lay["n3_g"], lay["n3_b"]), key_out="a")
            # feed-forward (decoder_layer.py:112-116)
            self._linear(s, "a", lay, "w1", 3072, 1024, lay["b1"], act=L.ACT_RELU, key_out="ffn")
            nxt = (w.layers[li + 1]["n1_g"], w.layers[li + 1]["n1_b"]) if li + 1 < nl else (w.after_g, w.after_b)
            self._linear(s, "ffn", lay, "w2", 1024, 3072, lay["b2"], residual=s["x"], out=s["x"], ln=nxt, key_out="a")
        # output layer (decoder.py:176-181)
        ns = self._proj(s, "a", {"out": w.out_w, "out3": getattr(w, "out_w3", None)}, "out", V, 1024)
        return s["part"], ns

    # ------------------------------------------------------------------------------------------ public API
    def prepare(self, s, x_packed: torch.Tensor, lengths: Sequence[int], maxlens: Optional[Sequence[int]] = None,
                ctc: bool = True, cross_kv: bool = True):
        """CTC posteriors (scorers/ctc.py:87-99) and the once-per-utterance cross-attention K/V projection; resets the beam
        state.  maxlens: positions per utterance after which eos is appended (default: its frame count, beam_search.py:349-350)."""
        w, V = self.w, self.n_vocab
        F = x_packed.shape[0]
        lib = L.load()
        n = w.ckv_w.shape[0]
        ctc = ctc and not self.dec_only               # attention-only search never reads the posteriors
        cross_kv = cross_kv and not self.ctc_only     # CTC-only search never runs the decoder
        if self.precision == "bf16x3":
            # fp32-accurate projections on the tensor cores: K' = 6 * 1024 (see weights.split3_weight)
            L.check(lib.avsr_split3(L.ptr(x_packed), L.ll(1024), L.ptr(s["x6"]), L.ll(F), 1024, L.stream()), "avsr_split3")
            if ctc:
                L.gemm_bf16(s["x6"], w.ctc_w6, F, V, 6144, L.make_epilogue(bias=w.ctc_b, out_f32=s["logp"], ld_f32=s["ldp"]))
            if cross_kv:
                L.gemm_bf16(s["x6"], w.ckv_w6, F, n, 6144, L.make_epilogue(bias=w.ckv_b, out_f32=s["ckv"], ld_f32=n))
        else:
            if ctc:
                L.sgemm(x_packed, w.ctc_w, F, V, 1024, L.make_epilogue(bias=w.ctc_b, out_f32=s["logp"], ld_f32=s["ldp"]))
            if cross_kv:
                L.sgemm(x_packed, w.ckv_w, F, n, 1024, L.make_epilogue(bias=w.ckv_b, out_f32=s["ckv"], ld_f32=n))
        if ctc:
            L.check(lib.avsr_log_softmax_rows(L.ptr(s["logp"]), L.ll(s["ldp"]), L.ll(F), V, L.stream()), "avsr_log_softmax_rows")
            if s.get("probs") is not None:
                L.check(lib.avsr_ctc_exp_posteriors(L.ptr(s["logp"]), L.ll(F * s["ldp"]), L.ptr(s["probs"]), L.stream()), "avsr_ctc_exp_posteriors")
        if cross_kv:
            L.check(lib.avsr_kv_head_major(L.ptr(s["ckv"]), L.ptr(s["ckv_t"]), L.ll(F), L.ll(s["F"]), n, 1, L.stream()), "avsr_kv_head_major")
        B, beam = s["B"], self.beam_size
        nb = len(lengths)                             # utterances of this batch; slots nb .. B-1 of the session stay parked
        offs = np.zeros(B, dtype=np.int32)
        offs[:nb] = np.concatenate([[0], np.cumsum(lengths)[:-1]])
        lens = np.ones(B, dtype=np.int32)
        lens[:nb] = lengths
        s["utt_T"].copy_(torch.from_numpy(lens))
        s["utt_off"].copy_(torch.from_numpy(offs))
        ml = lens.copy()
        if maxlens is not None:
            ml[:nb] = maxlens
        s["utt_maxlen"].copy_(torch.from_numpy(ml))
        s["n_utts"] = nb
        for k in ("step", "any_running", "ticket", "rprev_idx", "n_ended", "done", "overflow", "score", "dec_sc", "ctc_sc", "s_prev", "conv_len"):
            s[k].zero_()
        if self.ctc_only:
            s["psi"].zero_()                          # log_psi of the empty prefix
            s["dec_logp"].zero_()                     # the dropped decoder scorer contributes 0 * 0
        if self.dec_only:
            s["ctc_full"].zero_()
        s["n_run"].zero_()
        s["n_run"][:nb] = 1
        s["row_active"].zero_()
        s["row_active"].view(B, beam)[:nb, 0] = 1
        s["last_tok"].fill_(self.sos)
        s["best_len"].fill_(float("-inf"))
        s["best_all"].fill_(float("-inf"))

    @staticmethod
    def max_length(T: int, maxlenratio: float) -> int:
        """Positions searched for an utterance of T frames (beam_search.py:349-354)."""
        if maxlenratio == 0:
            return T
        if maxlenratio < 0:
            return -1 * int(maxlenratio)
        return max(1, int(maxlenratio * T))

    def decode_batch(self, x_packed: torch.Tensor, lengths: Sequence[int], max_steps: Optional[int] = None,
                     maxlenratio: float = 0.0) -> List[List[Hypothesis]]:
        """x_packed [sum(T),1024] fp32 encoder outputs (utterances back to back) -> n-best list per utterance.

        maxlenratio as in BeamSearch.forward (beam_search.py:330-355): 0 = up to T positions with end detection; > 0 = at
        most max(1, int(ratio * T)) positions, < 0 = at most -ratio positions, both WITHOUT end detection (:369).  The
        reference's CTCPrefixScoreTH indexes the frame axis with the prefix length (ctc_prefix_score.py:169-172), so more
        positions than frames raise there; here they raise a RuntimeError up front.

        With `self.n_groups` > 1 the utterances are decoded in independent groups on separate CUDA streams (every utterance
        evolves independently of its batch, so the split does not change any result; measured slower on B200, see __init__)."""
        L.require_cuda(x_packed, torch.float32, "encoder output")
        lengths = [int(t) for t in lengths]
        if x_packed.dim() != 2 or x_packed.shape[1] != 1024 or x_packed.shape[0] != sum(lengths) or min(lengths) < 1:
            raise RuntimeError(f"bad decode input: x {tuple(x_packed.shape)}, lengths {lengths}")
        B = len(lengths)
        self._maxlenratio = float(maxlenratio)
        if maxlenratio != 0.0 and any(self.max_length(t, maxlenratio) > t for t in lengths):
            raise RuntimeError(f"maxlenratio {maxlenratio} asks for more positions than an utterance has frames")
        G = max(1, min(self.n_groups, B))
        if G == 1:
            return self._decode_groups([(x_packed, lengths)], max_steps)[0]
        # contiguous groups of (almost) equal size; the packed rows of a group are contiguous
        bounds = [round(g * B / G) for g in range(G + 1)]
        offs = [0]
        for t in lengths:
            offs.append(offs[-1] + t)
        parts = [(x_packed[offs[bounds[g]]:offs[bounds[g + 1]]], lengths[bounds[g]:bounds[g + 1]]) for g in range(G)]
        out = []
        for r in self._decode_groups(parts, max_steps):
            out += r
        return out

    def _decode_groups(self, parts, max_steps):
        """parts: [(x_packed_g, lengths_g)]; each group gets its own session (buffers + graph) and stream."""
        G = len(parts)
        main = torch.cuda.current_stream()
        if len(self._streams) < G:
            self._streams += [torch.cuda.Stream(device=self.device) for _ in range(G - len(self._streams))]
        streams = [main] if G == 1 else self._streams[:G]
        ready = torch.cuda.Event()
        ready.record(main)
        chunk = self.POLL_EVERY
        sess, nsteps, done = [], [], []
        self.last_sessions = sess
        # ---- set-up per group: posteriors, cross K/V, position 0 eagerly (also warms the kernels up), graph capture
        mlr = getattr(self, "_maxlenratio", 0.0)
        for g, (xg, lg) in enumerate(parts):
            tmax = max(lg)
            s = self._session(len(lg), tmax, xg.shape[0], slot=g, no_end_detect=mlr != 0.0)
            self.last_session = s
            maxlens = [self.max_length(t, mlr) for t in lg]
            with torch.cuda.stream(streams[g]):
                streams[g].wait_event(ready)
                self.prepare(s, xg, lg, maxlens)
                n = max(maxlens) if max_steps is None else min(max(maxlens), max_steps)
                self._step(s)
                if n > 1 and self.use_graph and s["graph"] is None:
                    torch.cuda.synchronize()
                    # ONE graph holds POLL_EVERY positions (every kernel reads the position and liveness from device memory),
                    # so the kernels of consecutive positions chain through programmatic dependent launch inside the graph
                    # and the host only replays + polls.  Capture runs no kernels; state is untouched.
                    gr = torch.cuda.CUDAGraph()
                    n0 = L.launch_count
                    with torch.cuda.graph(gr):
                        for _ in range(chunk):
                            self._step(s)
                    s["launches_per_graph"] = L.launch_count - n0
                    # ... and a one-position graph for the last (tmax - 1) % POLL_EVERY positions of an utterance
                    gr1 = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gr1):
                        self._step(s)
                    L.launch_count = n0
                    s["graph"], s["graph1"] = gr, gr1
            sess.append(s)
            nsteps.append(n)
            done.append(1)
        # ---- position loop: one graph replay per group and round.  The "any utterance still running" flag of a replay is
        #      read one round LATE, after the next replay has been queued: the GPU never waits for the host between replays
        #      (a 16-position graph has ~1200 nodes; queueing it takes the host a few hundred microseconds).  If the flag
        #      says everything had already ended, the replay queued meanwhile only ran no-op kernels.
        active = [n > 1 for n in nsteps]
        self.last_positions_queued = 1           # positions the host queued (the search may have stopped earlier: one replay of slack)
        had_prev = [False] * G
        rnd = 0
        while any(active):
            for g in range(G):
                if not active[g] or done[g] >= nsteps[g]:
                    continue
                s = sess[g]
                n = min(chunk, nsteps[g] - done[g])
                with torch.cuda.stream(streams[g]):
                    if self.use_graph and n == chunk:
                        s["graph"].replay()
                        self.graph_launches += s["launches_per_graph"]
                    elif self.use_graph and max_steps is None:
                        for _ in range(n):
                            s["graph1"].replay()
                        self.graph_launches += n * (s["launches_per_graph"] // chunk)
                    else:
                        for _ in range(n):
                            self._step(s)
                    s["poll"][rnd & 1].copy_(s["any_running"], non_blocking=True)
                    s["poll_event"][rnd & 1].record(streams[g])
                done[g] += n
                self.last_positions_queued = max(self.last_positions_queued, done[g])
            for g in range(G):
                if not active[g]:
                    continue
                if had_prev[g]:
                    sess[g]["poll_event"][(rnd - 1) & 1].synchronize()
                    if int(sess[g]["poll"][(rnd - 1) & 1][0]) == 0:
                        active[g] = False
                        continue
                had_prev[g] = True
                if done[g] >= nsteps[g]:
                    active[g] = False                 # everything is queued; _collect synchronises
            rnd += 1
        out = []
        for g, (xg, lg) in enumerate(parts):
            with torch.cuda.stream(streams[g]):
                if int(sess[g]["overflow"].item()) != 0:
                    raise RuntimeError("ended-hypothesis table overflowed")
                out.append(self._collect(sess[g], lg, truncated=max_steps is not None))
            if G > 1:
                main.wait_stream(streams[g])
        return out

    def _collect(self, s, lengths, truncated=False) -> List[List[Hypothesis]]:
        """Backtrace the ended hypotheses on the host (one D2H at the end instead of the reference's per-step syncs), all
        hypotheses of the batch at once: one vectorised step per position instead of a Python loop per token."""
        tok = s["hist_tok"].cpu().numpy()
        prev = s["hist_prev"].cpu().numpy()
        r2j = s["run2j"].cpu().numpy()
        n_end = s["n_ended"].cpu().numpy()
        e_step, e_j, e_len = s["end_step"].cpu().numpy(), s["end_j"].cpu().numpy(), s["end_len"].cpu().numpy()
        e_sc, e_dec, e_ctc = s["end_score"].cpu(), s["end_dec"].cpu(), s["end_ctc"].cpu()
        B = s.get("n_utts", s["B"])
        self.last_stop_positions = (s["done"].cpu().numpy()[:B] - 1).tolist()     # position at which each utterance stopped
        hb = np.concatenate([np.full(int(n_end[b]), b, dtype=np.int64) for b in range(B)]) if B else np.zeros(0, np.int64)
        he = np.concatenate([np.arange(int(n_end[b]), dtype=np.int64) for b in range(B)]) if B else np.zeros(0, np.int64)
        n = len(hb)
        out = [[] for _ in range(B)]
        if n == 0:
            return out
        last = e_step[hb, he].astype(np.int64)            # position of the hypothesis' last token
        jj = e_j[hb, he].astype(np.int64)
        toks = np.zeros((n, int(last.max()) + 1), dtype=np.int64)
        for ii in range(int(last.max()), -1, -1):
            live = last >= ii                             # hypotheses that have a token at position ii
            if ii < int(last.max()):
                # step from position ii + 1 to its parent: candidate index at ii of the running slot it extended
                up = last >= ii + 1
                p = prev[hb[up], ii + 1, jj[up]]
                jj[up] = r2j[hb[up], ii, p]
            toks[live, ii] = tok[hb[live], ii, jj[live]]
        for k in range(n):
            b, e = int(hb[k]), int(he[k])
            yseq = [self.sos] + toks[k, :last[k] + 1].tolist()
            if int(e_len[b, e]) == len(yseq) + 1:        # eos appended at the last position (batch_beam_search.py:321-337)
                yseq.append(self.eos)
            scores = ({"ctc": e_ctc[b, e]} if self.ctc_only else
                      ({"decoder": e_dec[b, e]} if self.dec_only else {"decoder": e_dec[b, e], "ctc": e_ctc[b, e]}))
            out[b].append(Hypothesis(yseq=torch.tensor(yseq, dtype=torch.int64), score=e_sc[b, e], scores=scores, states={}))
        for b in range(B):
            out[b].sort(key=lambda h: float(h.score), reverse=True)     # stable, like sorted() in beam_search.py:378
        return out

    def forward(self, x: torch.Tensor, maxlenratio: float = 0.0, minlenratio: float = 0.0) -> List[Hypothesis]:
        """Reference entry point: x [T, 1024] -> n-best (beam_search.py:330-406)."""
        # minlenratio only feeds a debug log and the retry taken when NO hypothesis ended (:355-358, :380-389); eos is
        # appended at the last position, so that branch is unreachable and the value has no effect on the result
        x = x.to(self.device, torch.float32).contiguous()
        return self.decode_batch(x, [x.shape[0]], maxlenratio=maxlenratio)[0]

    __call__ = forward


def get_beam_search_decoder(model, token_list, ctc_weight=0.1, beam_size=3):
    """Same signature as the reference factory (src/avhubert_avsr/avhubert_avsr_model.py:12-36); ``model`` is an
    ``avsr_b200.model.AVSRCocktailB200`` (exposes ``.decoder_weights``, ``.sos``, ``.eos``)."""
    return BatchedBeamSearch(model.decoder_weights, beam_size=beam_size, ctc_weight=ctc_weight, token_list=token_list,
                             device=model.device)
