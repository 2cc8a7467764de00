// CTC prefix scoring, score fusion, top-k pruning and beam bookkeeping of the joint CTC/attention beam search.
// Reference: CTCPrefixScoreTH.__call__ (src/nets/ctc_prefix_score.py:68-187; scalar restatement in SURVEY.md App. B),
// CTCPrefixScorer.select_state (src/nets/scorers/ctc.py:40-63), BatchBeamSearch.search / batch_beam / post_process
// (src/nets/batch_beam_search.py:86-110,208-349), end_detect (src/nets/e2e_asr_common.py:18-48).
// The reference runs ~16.7k ATen ops + per-hyp host syncs per step here; this file does it in two launches per step
// for all utterances at once, with the hypothesis state resident on the device.
#include <stdlib.h>
#include "common.cuh"
#include "dec_tail.cuh"

namespace {

constexpr float LOGZERO = -10000000000.0f;   // ctc_prefix_score.py:33

// torch.logsumexp over two elements: m + log(exp(a-m) + exp(b-m)).
__device__ __forceinline__ float lse2(float a, float b) {
    const float m = fmaxf(a, b), n = fminf(a, b);
    return m + logf(1.f + expf(n - m));
}

// logaddexp of the serial forward chain: ex2.approx / lg2.approx instead of the IEEE-exact sequences.  The operands are log
// probabilities of magnitude 1e0..1e3 whose own fp32 spacing (1e-7..6e-5) is far above the ~2e-7 absolute error of the
// approximations, and the chain is the critical path of the kernel (T dependent steps).
__device__ __forceinline__ float lse2_chain(float a, float b) {
    const float m = fmaxf(a, b), n = fminf(a, b);
    return m + __logf(1.f + __expf(n - m));
}

// ------------------------------------------------------------------------------------------------------------------
// Pre-beam mode: one CTA per hypothesis row, two phases.
//   Phase A (all 128 threads, parallel over time): gather the candidates' log-posterior columns x[t][c_s] and the blank
//   column into shared memory, turn the parent's forward variables into log_phi (r_sum) and the blank-ending part.
//   Phase B: warp 0 runs the S serial forward chains (one lane each; per step only the two logaddexp of
//   ctc_prefix_score.py:156-161 are on the dependency chain, everything else was precomputed), while warps 1-3 reduce
//   log_psi = logsumexp_t(log_phi[t-1] + x[t]) in parallel (max, then sum of exponentials, as torch.logsumexp does).
// r_buf [2][R*S][tmax][2] ping-pongs on step parity; rprev_idx[row] is the chain (in the "current" half) that the
// surviving hypothesis inherited.  Outputs psi[row][s] (log prefix probability) and rsum_last[row] = r_sum[T-1].
constexpr int PB_THREADS = 128;
constexpr int PB_MAXS = 12;

struct PrebeamArgs {
    const float* logp; int V, ldp, blank;
    const int* utt_off; const int* utt_T; const int* n_run; int beam, R, S;
    const int* last_tok; const int* rprev_idx; float* r_buf; int tmax, pitch; const int* step_p;
    float* psi; float* rsum_last;
};

// One hypothesis row on a CTA of NT threads.  pb_sm: (S + 3) * pitch floats = xs [S][pitch], xb [pitch], phi [pitch],
// pbk [pitch]; ids: the row's S candidate tokens (global or shared memory).  The caller has already filled xb (blank column,
// which does not depend on the previous kernel) and made sure the row is live.
struct PrebeamRowState {        // beam state of the row, requested by the caller right after griddepcontrol.wait
    int step, last, rprev;
};
__device__ __forceinline__ PrebeamRowState prebeam_row_state(const PrebeamArgs& a, int row) {
    PrebeamRowState st;
    st.step = *a.step_p;
    st.last = a.last_tok[row];
    st.rprev = a.rprev_idx[row];
    return st;
}

template <int NT>
__device__ __forceinline__ void ctc_prebeam_row(const PrebeamArgs& a, int row, const PrebeamRowState& rs, float* pb_sm, const int* ids,
                                                float (*s_rmax)[PB_MAXS], float (*s_rsum)[PB_MAXS]) {
    constexpr int NR = NT - 32;                      // threads of the reduction warps
    constexpr int NRW = NR / 32;
    const int utt = row / a.beam;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = a.S, pitch = a.pitch, R = a.R, tmax = a.tmax;
    float* xs = pb_sm;
    float* xb = pb_sm + (size_t)S * pitch;
    float* phi = xb + pitch;
    float* pbk = phi + pitch;
    const int T = a.utt_T[utt];
    const float* lp = a.logp + (long long)a.utt_off[utt] * a.ldp;
    const int step = rs.step;
    const int cur = step & 1;
    const int start = step > 1 ? step : 1;
    const int last = rs.last;
    const float2* rp = reinterpret_cast<const float2*>(a.r_buf) + ((long long)cur * R * S + (step > 0 ? rs.rprev : 0)) * tmax;

    // ---- phase A
    for (int t = tid; t < T; t += NT) {
        for (int s = 0; s < S; ++s) xs[s * pitch + t] = __ldg(lp + (long long)t * a.ldp + ids[s]);
        if (step > 0) {
            const float2 p = rp[t];
            phi[t] = lse2(p.x, p.y);
            pbk[t] = p.y;
        }
    }
    __syncthreads();
    if (step == 0 && tid == 0) {
        // empty prefix: r_prev[:,0] = logzero, r_prev[:,1] = cumsum(x[:, blank]) (ctc_prefix_score.py:58-63)
        float cum = 0.f;
        for (int t = 0; t < T; ++t) {
            cum += xb[t];
            pbk[t] = cum;
            phi[t] = lse2(LOGZERO, cum);
        }
    }
    if (step == 0) __syncthreads();

    // ---- phase B
    if (warp == 0) {
        if (lane < S) {
            const int s = lane;
            const float* ph = (ids[s] == last) ? pbk : phi;
            const float* x = xs + s * pitch;
            float2* ro = reinterpret_cast<float2*>(a.r_buf) + ((long long)(cur ^ 1) * R * S + (long long)row * S + s) * tmax;
            float rn = (step == 0) ? x[0] : LOGZERO;
            float rb = LOGZERO;
            ro[start - 1] = make_float2(rn, rb);
#pragma unroll 4
            for (int t = start; t < T; ++t) {
                const float nrn = lse2_chain(rn, ph[t - 1]) + x[t];
                const float nrb = lse2_chain(rn, rb) + xb[t];
                rn = nrn;
                rb = nrb;
                ro[t] = make_float2(rn, rb);
            }
        }
        return;
    }
    const int tt = tid - 32, w3 = warp - 1;
    for (int s = 0; s < S; ++s) {
        const float* ph = (ids[s] == last) ? pbk : phi;
        const float* x = xs + s * pitch;
        float mx = -INFINITY;
        for (int t = start + tt; t < T; t += NR) mx = fmaxf(mx, ph[t - 1] + x[t]);
        mx = warp_max(mx);
        if (lane == 0) s_rmax[w3][s] = mx;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NR) : "memory");
    for (int s = 0; s < S; ++s) {
        const float* ph = (ids[s] == last) ? pbk : phi;
        const float* x = xs + s * pitch;
        const float r0 = (step == 0) ? x[0] : LOGZERO;                       // r[start-1, 0]
        float M = r0;
#pragma unroll
        for (int w = 0; w < NRW; ++w) M = fmaxf(M, s_rmax[w][s]);
        float sum = 0.f;
        for (int t = start + tt; t < T; t += NR) sum += expf(ph[t - 1] + x[t] - M);
        sum = warp_sum(sum);
        if (lane == 0) s_rsum[w3][s] = sum;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NR) : "memory");
    if (tt < S) {
        const int s = tt;
        const float r0 = (step == 0) ? xs[s * pitch] : LOGZERO;
        float M = r0;
#pragma unroll
        for (int w = 0; w < NRW; ++w) M = fmaxf(M, s_rmax[w][s]);
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < NRW; ++w) sum += s_rsum[w][s];
        sum += expf(r0 - M);
        a.psi[row * S + s] = M + logf(sum);
    }
    if (tt == 32) a.rsum_last[row] = phi[T - 1];
}

__global__ void __launch_bounds__(PB_THREADS)
ctc_prefix_prebeam_kernel(const PrebeamArgs a, const int* __restrict__ part_ids) {
    extern __shared__ float pb_sm[];
    __shared__ float s_rmax[PB_THREADS / 32 - 1][PB_MAXS], s_rsum[PB_THREADS / 32 - 1][PB_MAXS];
    const int row = blockIdx.x, utt = row / a.beam;
    pdl_trigger();
    // the posteriors were written before the chain of step kernels started: the blank column is fetched before the wait
    const int T = a.utt_T[utt];
    const float* lp = a.logp + (long long)a.utt_off[utt] * a.ldp;
    float* xb = pb_sm + (size_t)a.S * a.pitch;
    for (int t = threadIdx.x; t < T; t += PB_THREADS) xb[t] = __ldg(lp + (long long)t * a.ldp + a.blank);
    pdl_wait();
    const int nrun = a.n_run[utt];
    const PrebeamRowState rs = prebeam_row_state(a, row);
    if ((row % a.beam) >= nrun) return;
    ctc_prebeam_row<PB_THREADS>(a, row, rs, pb_sm, part_ids + row * a.S, s_rmax, s_rsum);
}

// Fused tail of a decode position, one CTA (512 threads) per hypothesis row: output-layer log_softmax + pre-beam top-S
// (dec_tail.cuh) and, for the S candidates just found, the CTC prefix scores - one launch instead of two on the critical
// chain (the candidates stay in shared memory).
template <int ITER>
__global__ void __launch_bounds__(LSM_THREADS)
dec_tail_kernel(const float* __restrict__ part, int nsplit, const float* __restrict__ bias, float* __restrict__ dec_logp,
                int* __restrict__ part_ids, const PrebeamArgs a) {
    extern __shared__ float pb_sm[];
    __shared__ LsmSmem lsm;
    __shared__ int s_ids[PB_MAXS];
    __shared__ float s_rmax[LSM_THREADS / 32 - 1][PB_MAXS], s_rsum[LSM_THREADS / 32 - 1][PB_MAXS];
    const int row = blockIdx.x, utt = row / a.beam;
    pdl_trigger();
    float v[ITER];
    lsm_load_bias<ITER>(v, bias, a.V);
    const int T = a.utt_T[utt];
    const float* lp = a.logp + (long long)a.utt_off[utt] * a.ldp;
    float* xb = pb_sm + (size_t)a.S * a.pitch;
    for (int t = threadIdx.x; t < T; t += LSM_THREADS) xb[t] = __ldg(lp + (long long)t * a.ldp + a.blank);
    pdl_wait();
    const int nrun = a.n_run[utt];
    const PrebeamRowState rs = prebeam_row_state(a, row);       // in flight while the logits are reduced
    if ((row % a.beam) >= nrun) return;
    lsm_topk_row<ITER>(v, lsm, part, nsplit, a.R, a.V, row, dec_logp, part_ids + row * a.S, s_ids, a.S);
    __syncthreads();
    ctc_prebeam_row<LSM_THREADS>(a, row, rs, pb_sm, s_ids, s_rmax, s_rsum);
}

// ------------------------------------------------------------------------------------------------------------------
// Full-vocabulary mode (scoring_ids=None, ctc_prefix_score.py:115-119).  For every live hyp h and token c
//     log_psi[h][c] = logsumexp( r[start-1,0][c], { log_phi[h][t-1](c) + x[t][c] : t = start .. T-1 } )
// where log_phi[h][t] = r_sum[h][t] of the parent prefix, except for c == last token of h where it is the blank-ending
// forward variable.  None of this depends on the NEW forward variables r[t][c] (they are only state for the next step
// and are recomputed for the survivors by the pre-beam kernel), so the sum over t is a skinny matrix product
//     Psi[h][c] = exp(M_h) * sum_t E[h][t-1] * P[t][c],   E[h][t] = exp(r_sum[h][t] - M_h),  P = exp(x),
// with ONE exp per log-posterior shared by all hyps (the log-domain recursion needs ~5 MUFU ops per (t, hyp, token) and
// is SFU-bound at ~11 % of HBM peak).  The log-posteriors are read from HBM exactly once, coalesced; E lives in shared
// memory.  Cells the linear form cannot represent (c == last token, or a sum that underflows fp32) are re-evaluated
// exactly in the log domain by a warp each, so every cell matches the reference to fp32 rounding.
//
// Work decomposition (HBM-bound streaming): the log-posteriors of an utterance are one dense [T][ldp] block whose rows are
// 16-byte aligned (ldp % 4 == 0; the producer pads V = 5049 to 5056).  CTA = (utterance, group of up to FV_CG columns, time
// split).  Warp 5's elected thread streams the CTA's row segments into a shared-memory ring with bulk async copies
// (cp.async.bulk + mbarrier transaction counts: ~70 KB per CTA in flight, requested before the preamble runs); the 160
// consumer threads own four consecutive columns each, read them back with one 16-byte shared-memory load per row and keep
// the sums for up to FV_NHP hyps in registers.  (History: 4-byte loads of 5 strided columns, too few bytes in flight, 45 %
// of the HBM peak; ~27 instructions per posterior, issue-bound, 37 %; 16-byte loads double-buffered in registers, 2 x 12
// rows per thread, 62 %.)  Warp 6 evaluates the c == last-token cells exactly while the stream runs.  With more than one
// time split the partial sums are published and the last CTA of a (utterance, column group) to finish (ticket) adds them in
// split order, takes the logarithm and writes the scores.
constexpr int FV_MAXH = 8;
constexpr int FV_THREADS = 160;           // consumer threads: a thread owns four consecutive columns
constexpr int FV_CG = FV_THREADS * 4;     // most columns a group can have (the plan sizes the groups to fill the SMs)
constexpr int FV_ALL = FV_THREADS + 64;   // + warp 5: row-stream producer, warp 6: exact evaluation of the c == last-token cells
constexpr int FV_FR = 8;                  // rows per ring slot (fewer rows per slot measured slower: 4 rows 70 us, 2 rows 111 us against 55 us)
constexpr int FV_MAXST = 4;               // ring slots (the host takes fewer when the tables of a long utterance need the room; 2 .. 4 measure the same)
constexpr int FV_NHP = 5;                 // hyps accumulated per pass over the block (beam <= 5: the block is read once)
constexpr int FV_ES = 8;                  // row stride of the E table in shared memory (floats): one or two 16-byte reads per row
constexpr int FV_NSPECIAL = 1024;
constexpr float FV_TINY = 1e-30f;

__device__ __forceinline__ float fv_exp(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
    return r;
}

// ---- the row stream: shared-memory ring filled by bulk async copies (one per row segment), consumed with 16-byte reads.
// A chunk = FV_FR consecutive rows of this CTA's column group; a ring slot holds one chunk.  The producer (one elected
// thread) keeps every slot of the ring requested, so a CTA has nst * FV_FR rows (~70 KB) in flight without holding them in
// registers - the register-staged form (2 x 12 rows per thread) stalled at 62 % of the HBM peak.
__device__ __forceinline__ void fv_bulk_row(uint32_t dst, const float* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fv_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    for (uint32_t i = 0; !ok; ++i) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (i > (1u << 28)) __trap();                // seconds, not a hung GPU, if a byte count was ever wrong
    }
}

// rows [t, t + nr) of the chunk at shared-memory address xs: acc[k][h] += E[h][t'-1] * exp(x[t'][c0 + k]).
// FULL = all FV_FR rows, no per-row branch: the rows of a chunk are independent until the final FMAs, and with ~2.5 warps per
// scheduler it is this instruction-level parallelism (8 rows of LDS -> ex2 -> FMA in flight) that hides the latencies; with
// a (uniform) branch per row every row paid its own LDS + MUFU latency chain, ~185 cycles, and the CONSUMERS, not HBM, paced
// the kernel (measured: time = 0.40 us per chunk + 0.097 us per row, whatever the ring depth).
template <int NH, bool FULL, bool PRE>
__device__ __forceinline__ void fv_consume_chunk(const float* __restrict__ s_E, int t, int nr, uint32_t xs, uint32_t row_bytes,
                                                 float (&acc)[4][FV_NHP]) {
    constexpr int G = FULL ? 4 : 1;                  // rows in flight together (registers: 2 CTAs of 7 warps leave 128 per thread)
#pragma unroll(FULL ? FV_FR / G : 1)
    for (int u0 = 0; u0 < (FULL ? FV_FR : nr); u0 += G) {
        float4 x[G];
#pragma unroll
        for (int i = 0; i < G; ++i)
            if (FULL || u0 + i < nr)
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(x[i].x), "=f"(x[i].y), "=f"(x[i].z), "=f"(x[i].w) : "r"(xs + (uint32_t)(u0 + i) * row_bytes));
#pragma unroll
        for (int i = 0; i < G; ++i) {
            if (FULL || u0 + i < nr) {               // uniform across the CTA
                const int u = u0 + i;
                float ev[8];
                const float4 e0 = *reinterpret_cast<const float4*>(s_E + (t + u - 1) * FV_ES);
                ev[0] = e0.x; ev[1] = e0.y; ev[2] = e0.z; ev[3] = e0.w;
                if (NH > 4) {
                    const float4 e1 = *reinterpret_cast<const float4*>(s_E + (t + u - 1) * FV_ES + 4);
                    ev[4] = e1.x; ev[5] = e1.y; ev[6] = e1.z; ev[7] = e1.w;
                }
                // ex2.approx.ftz(x * log2 e): one FMUL + one MUFU per posterior (__expf adds a denormal-range fix-up: a compare
                // and two more multiplies).  Relative error ~2^-21, far below the rounding of the fp32 sum itself; posteriors
                // below e^-87 flush to zero, where the IEEE result (< 1e-38) could not change a sum that is checked against 1e-30.
                // PRE: the stream already holds the posteriors (avsr_ctc_exp_posteriors: the same ex2 of the same inputs, done once
                // per utterance instead of once per decode position)
                const float pr[4] = {PRE ? x[i].x : fv_exp(x[i].x), PRE ? x[i].y : fv_exp(x[i].y), PRE ? x[i].z : fv_exp(x[i].z),
                                     PRE ? x[i].w : fv_exp(x[i].w)};
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int h = 0; h < NH; ++h) acc[k][h] = fmaf(ev[h], pr[k], acc[k][h]);
            }
        }
    }
}

// exact log-domain value of one cell (the reference's formula): logsumexp over t of phi[t-1] + x[t][c], plus r[start-1,0].
// The column is read ONCE: every lane requests all of its (strided, one sector each) posteriors before the first use and
// keeps the terms in registers for the second pass - one memory round trip instead of ~2 T / 32 partly serialised ones; this
// warp runs after the streaming pass, when the rest of its CTA is idle, so its latency is the tail of the kernel.
constexpr int FV_XR = 16;                 // terms per lane kept in registers (T <= 512); longer utterances loop in chunks
__device__ __forceinline__ float fv_exact_warp(const float* __restrict__ lp, int ldp, int c, int start, int T, const float* phi, int nh,
                                               int h, float x0, int lane) {
    if (T - start <= 32 * FV_XR) {
        float v[FV_XR];
#pragma unroll
        for (int u = 0; u < FV_XR; ++u) {
            const int t = start + lane + 32 * u;
            v[u] = (t < T) ? __ldg(lp + (long long)t * ldp + c) : 0.f;
        }
        float mx = x0;
#pragma unroll
        for (int u = 0; u < FV_XR; ++u) {
            const int t = start + lane + 32 * u;
            v[u] = (t < T) ? phi[(t - 1) * nh + h] + v[u] : -INFINITY;
            mx = fmaxf(mx, v[u]);
        }
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int u = 0; u < FV_XR; ++u) sum += expf(v[u] - mx);          // exp(-inf) = 0 for the slots past the end
        sum = warp_sum(sum);
        return mx + logf(sum + expf(x0 - mx));
    }
    float mx = x0;
    for (int t = start + lane; t < T; t += 32) mx = fmaxf(mx, phi[(t - 1) * nh + h] + lp[(long long)t * ldp + c]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int t = start + lane; t < T; t += 32) sum += expf(phi[(t - 1) * nh + h] + lp[(long long)t * ldp + c] - mx);
    sum = warp_sum(sum);
    return mx + logf(sum + expf(x0 - mx));
}

// One pass of the consumers over the CTA's rows: chunk g of the ring sequence (g runs on across passes), slot g % nst.
template <int NH, bool PRE>
__device__ __forceinline__ void fv_stream_ring(const float* __restrict__ s_E, int t_lo, int t_hi, int nch, int nst, int& g, uint32_t bar_s,
                                               uint32_t ring_s, uint32_t slot_bytes, uint32_t row_bytes, bool active, int tid, int lane,
                                               float (&acc)[4][FV_NHP]) {
    for (int ch = 0; ch < nch; ++ch, ++g) {
        const int slot = g % nst;
        fv_mbar_wait(bar_s + 8u * slot, (uint32_t)(g / nst) & 1u);                       // full[slot] (all lanes poll: one poller per warp + __syncwarp measured 45 % slower)
        const int t = t_lo + ch * FV_FR, nr = min(FV_FR, t_hi - t);
        if (active) {
            if (nr == FV_FR) fv_consume_chunk<NH, true, PRE>(s_E, t, nr, ring_s + slot * slot_bytes + 16u * tid, row_bytes, acc);
            else fv_consume_chunk<NH, false, PRE>(s_E, t, nr, ring_s + slot * slot_bytes + 16u * tid, row_bytes, acc);
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_s + 8u * (FV_MAXST + slot)) : "memory");   // empty[slot]
    }
}

template <bool PRE>
__global__ void __launch_bounds__(FV_ALL, 2)
ctc_prefix_full_kernel(const float* __restrict__ logp, const float* __restrict__ probs, int V, int ldp, int blank, int eos, const int* __restrict__ utt_off,
                       const int* __restrict__ utt_T, const int* __restrict__ n_run, int beam, int R, int S,
                       const int* __restrict__ last_tok, const int* __restrict__ rprev_idx, const float* __restrict__ r_buf,
                       int tmax, const int* __restrict__ step_p, const float* __restrict__ s_prev, float* __restrict__ scores,
                       int ncg, int cgw, int tsplit, float* __restrict__ part, int* __restrict__ tickets, int nst, int tab_bytes, int pdl) {
    extern __shared__ __align__(128) float sm[];    // s_E [T][FV_ES], s_rs [T][nh] (r_sum), s_pb [T][nh] (blank-ending); ring at tab_bytes
    __shared__ float s_M[FV_MAXH];
    __shared__ float s_red[FV_ALL / 32][FV_MAXH];
    __shared__ int s_nspecial, s_last;
    __shared__ int s_special[FV_NSPECIAL];
    __shared__ __align__(8) uint64_t s_bar[2 * FV_MAXST];        // full[FV_MAXST], empty[FV_MAXST]
    const int utt = blockIdx.y;
    const int cg = blockIdx.x % ncg, z = blockIdx.x / ncg;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(s_bar);
    if (tid == 0) {
        s_nspecial = 0;
        for (int i = 0; i < FV_MAXST; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_s + 8u * i), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_s + 8u * (FV_MAXST + i)), "r"(FV_THREADS / 32));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // programmatic dependent launch: the next kernel of the stream may place its CTAs as ours retire; nothing written by an
    // earlier kernel (posteriors, beam state, forward variables) is read before the wait
    if (pdl) pdl_trigger();
    pdl_wait();
    const int nh = n_run[utt];
    const int step = *step_p;
    const int T = utt_T[utt];
    const long long uoff = utt_off[utt];
    if (nh == 0) return;
    const float* lp = logp + uoff * ldp;
    const float* src = (PRE ? probs : logp) + uoff * ldp;        // what the ring streams
    float* s_E = sm;
    float* s_rs = sm + (size_t)T * FV_ES;
    float* s_pb = s_rs + (size_t)T * nh;
    const int cur = step & 1;
    const int start = step > 1 ? step : 1;
    const int per = (T - start + tsplit - 1) / tsplit;
    const int t_lo = start + z * per, t_hi = min(T, t_lo + per);
    // ---- the row stream of this CTA: columns [cbase, cbase + width) of rows [t_lo, t_hi), in chunks of FV_FR rows
    const int cbase = cg * cgw;
    const int width = min(cgw, ldp - cbase);         // > 0: the plan never makes an empty group; a multiple of 4 (ldp, cgw are)
    const uint32_t row_bytes = (uint32_t)cgw * 4u, seg_bytes = (uint32_t)width * 4u, slot_bytes = FV_FR * row_bytes;
    const int nch = t_hi > t_lo ? (t_hi - t_lo + FV_FR - 1) / FV_FR : 0;
    const int total = nch * ((nh + FV_NHP - 1) / FV_NHP);        // the block is streamed once per FV_NHP hyps
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(sm) + (uint32_t)tab_bytes;
    __syncthreads();
    auto issue = [&](int g) {                        // producer thread only (one copy per lane of a warp measured slower)
        const int slot = g % nst, t = t_lo + (g % nch) * FV_FR, nr = min(FV_FR, t_hi - t);
        const uint32_t fb = bar_s + 8u * slot;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"((uint32_t)nr * seg_bytes) : "memory");
        for (int u = 0; u < nr; ++u)
            fv_bulk_row(ring_s + slot * slot_bytes + (uint32_t)u * row_bytes, src + (long long)(t + u) * ldp + cbase, seg_bytes, fb);
    };
    // the whole ring is requested now: its latency is hidden by the preamble below
    if (tid == FV_THREADS)
        for (int g = 0; g < min(nst, total); ++g) issue(g);
    // this thread's columns c0 .. c0+3 (columns >= V lie in the row padding; they are dropped at the end)
    const bool active = tid < FV_THREADS && 4 * tid < width;
    const int c0 = active ? cbase + 4 * tid : ldp;
    // ---- parent forward variables -> r_sum, blank-ending part, per-hyp maximum over the time range that is used
    if (step == 0) {
        // r_prev[:,1] = cumsum(x[:, blank]) (ctc_prefix_score.py:58-63), r_prev[:,0] = logzero: the column is fetched by all
        // threads, summed in frame order by one (the order of the reference's CPU cumsum), finished by all
        for (int t = tid; t < T; t += FV_ALL) s_pb[t * nh] = lp[(long long)t * ldp + blank];
        __syncthreads();
        if (tid == 0) {
            float cs = 0.f;
            for (int t = 0; t < T; ++t) {
                cs += s_pb[t * nh];
                for (int h = 0; h < nh; ++h) s_pb[t * nh + h] = cs;
            }
        }
        __syncthreads();
        for (int i = tid; i < T * nh; i += FV_ALL) s_rs[i] = lse2(LOGZERO, s_pb[i]);
    } else {
        for (int i = tid; i < T * nh; i += FV_ALL) {
            const int t = i / nh, h = i % nh;
            const float2 v = (reinterpret_cast<const float2*>(r_buf) + ((long long)cur * R * S + rprev_idx[utt * beam + h]) * tmax)[t];
            s_pb[i] = v.y;
            s_rs[i] = lse2(v.x, v.y);
        }
    }
    __syncthreads();
    for (int h = 0; h < nh; ++h) {
        float mx = -INFINITY;
        for (int t = start - 1 + tid; t <= T - 2; t += FV_ALL) mx = fmaxf(mx, s_rs[t * nh + h]);
        mx = warp_max(mx);
        if (lane == 0) s_red[warp][h] = mx;
    }
    __syncthreads();
    if (tid < nh) {
        float mx = -INFINITY;
        for (int w = 0; w < FV_ALL / 32; ++w) mx = fmaxf(mx, s_red[w][tid]);
        s_M[tid] = (mx == -INFINITY) ? 0.f : mx;    // T == 1: no recursion term at all
    }
    __syncthreads();
    for (int i = tid; i < T * FV_ES; i += FV_ALL) {              // E table of the first pass: hyp j in column j
        const int t = i / FV_ES, j = i % FV_ES;
        s_E[i] = j < min(FV_NHP, nh) ? expf(s_rs[t * nh + j] - s_M[j]) : 0.f;
    }
    __syncthreads();

    // ---- finalisation of one cell from its linear-domain sum `a` (all time splits added)
    auto finish = [&](int h, int c, float a) {
        const int row = utt * beam + h;
        if (c == last_tok[row]) return;                          // evaluated exactly by warp 6 of the z == 0 CTA
        const float x0 = (step == 0) ? lp[c] : LOGZERO;          // r[start-1, 0]: x[0][c] for the empty prefix, else logzero
        a += expf(x0 - s_M[h]);
        if (!(a > FV_TINY) || !(a < 1e30f)) {
            const int slot = atomicAdd(&s_nspecial, 1);          // exact log-domain evaluation by a warp below
            if (slot < FV_NSPECIAL) { s_special[slot] = c * FV_MAXH + h; return; }
            const float* phi = s_rs;                             // list full (pathological input): this thread does it alone
            float mx = x0;
            for (int t = start; t < T; ++t) mx = fmaxf(mx, phi[(t - 1) * nh + h] + lp[(long long)t * ldp + c]);
            float sum = expf(x0 - mx);
            for (int t = start; t < T; ++t) sum += expf(phi[(t - 1) * nh + h] + lp[(long long)t * ldp + c] - mx);
            float lpsi = mx + logf(sum);
            if (c == eos) lpsi = s_rs[(T - 1) * nh + h];
            if (c == blank) lpsi = LOGZERO;
            scores[(long long)row * V + c] = lpsi - s_prev[row];
            return;
        }
        float lpsi = s_M[h] + logf(a);
        if (c == eos) lpsi = s_rs[(T - 1) * nh + h];
        if (c == blank) lpsi = LOGZERO;
        scores[(long long)row * V + c] = lpsi - s_prev[row];
    };

    if (warp == FV_THREADS / 32) {
        // ---- producer: refill a slot as soon as the five consumer warps have released it
        if (lane == 0) {
            for (int g = nst; g < total; ++g) {
                fv_mbar_wait(bar_s + 8u * (FV_MAXST + g % nst), (uint32_t)(g / nst - 1) & 1u);
                issue(g);
            }
        }
    } else if (warp == FV_THREADS / 32 + 1) {
        // ---- the c == last-token cell of every hyp (its phi is the blank-ending forward variable): always exact, and known
        //      up front, so it is evaluated while the stream runs instead of after it
        if (z == 0) {
            for (int h = 0; h < nh; ++h) {
                const int row = utt * beam + h;
                const int cx = last_tok[row];
                if (cx < cbase || cx >= cbase + cgw || cx >= V) continue;                // warp-uniform
                const float x0 = (step == 0) ? lp[cx] : LOGZERO;
                float lpsi = fv_exact_warp(lp, ldp, cx, start, T, s_pb, nh, h, x0, lane);
                if (lane == 0) {
                    if (cx == eos) lpsi = s_rs[(T - 1) * nh + h];
                    if (cx == blank) lpsi = LOGZERO;
                    scores[(long long)row * V + cx] = lpsi - s_prev[row];
                }
            }
        }
    } else {
        // ---- consumers: stream this CTA's rows, FV_NHP hyps at a time
        int g = 0;
        for (int h0 = 0; h0 < nh; h0 += FV_NHP) {
            const int ng = min(FV_NHP, nh - h0);
            float acc[4][FV_NHP];
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int h = 0; h < FV_NHP; ++h) acc[k][h] = 0.f;
            if (h0 > 0) {
                asm volatile("bar.sync 1, %0;" ::"n"(FV_THREADS) : "memory");              // the previous pass is done with the table
                for (int i = tid; i < T * FV_ES; i += FV_THREADS) {                        // E table of this pass: hyp h0 + j in column j
                    const int t = i / FV_ES, j = i % FV_ES;
                    s_E[i] = j < ng ? expf(s_rs[t * nh + h0 + j] - s_M[h0 + j]) : 0.f;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(FV_THREADS) : "memory");
            }
            switch (ng) {
                case 1: fv_stream_ring<1, PRE>(s_E, t_lo, t_hi, nch, nst, g, bar_s, ring_s, slot_bytes, row_bytes, active, tid, lane, acc); break;
                case 2: fv_stream_ring<2, PRE>(s_E, t_lo, t_hi, nch, nst, g, bar_s, ring_s, slot_bytes, row_bytes, active, tid, lane, acc); break;
                case 3: fv_stream_ring<3, PRE>(s_E, t_lo, t_hi, nch, nst, g, bar_s, ring_s, slot_bytes, row_bytes, active, tid, lane, acc); break;
                case 4: fv_stream_ring<4, PRE>(s_E, t_lo, t_hi, nch, nst, g, bar_s, ring_s, slot_bytes, row_bytes, active, tid, lane, acc); break;
                default: fv_stream_ring<5, PRE>(s_E, t_lo, t_hi, nch, nst, g, bar_s, ring_s, slot_bytes, row_bytes, active, tid, lane, acc); break;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = c0 + k;
                if (c < V) {
#pragma unroll
                    for (int h = 0; h < FV_NHP; ++h) {
                        if (h < ng) {
                            if (tsplit == 1) finish(h0 + h, c, acc[k][h]);
                            else part[(((long long)utt * tsplit + z) * beam + h0 + h) * V + c] = acc[k][h];
                        }
                    }
                }
            }
        }
    }
    if (tsplit > 1) {
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const int tk = atomicAdd(&tickets[utt * ncg + cg], 1);
            s_last = (tk == tsplit - 1) ? 1 : 0;
            if (s_last) tickets[utt * ncg + cg] = 0;             // re-armed for the next launch
        }
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        for (int h = 0; h < nh; ++h) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = c0 + k;
                if (c < V) {
                    float a = 0.f;
                    for (int zz = 0; zz < tsplit; ++zz) a += __ldcg(part + (((long long)utt * tsplit + zz) * beam + h) * V + c);
                    finish(h, c, a);
                }
            }
        }
    }
    __syncthreads();
    // ---- exact path: one warp per cell whose linear-domain sum left the fp32 range
    const int nsp = min(s_nspecial, FV_NSPECIAL);
    for (int i = warp; i < nsp; i += FV_ALL / 32) {
        const int cx = s_special[i] / FV_MAXH, h = s_special[i] % FV_MAXH;
        const int row = utt * beam + h;
        const float x0 = (step == 0) ? lp[cx] : LOGZERO;
        float lpsi = fv_exact_warp(lp, ldp, cx, start, T, s_rs, nh, h, x0, lane);
        if (lane == 0) {
            if (cx == eos) lpsi = s_rs[(T - 1) * nh + h];
            if (cx == blank) lpsi = LOGZERO;
            scores[(long long)row * V + cx] = lpsi - s_prev[row];
        }
    }
}

__global__ void __launch_bounds__(256)
ctc_exp_posteriors_kernel(const float4* __restrict__ logp, long long n4, float4* __restrict__ probs) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        const float4 x = __ldg(logp + i);
        probs[i] = make_float4(fv_exp(x.x), fv_exp(x.y), fv_exp(x.z), fv_exp(x.w));
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Fusion + pruning + bookkeeping: one CTA per utterance.
__device__ __forceinline__ float ctc_logpsi(const AvsrBeamState& st, const int* __restrict__ part_ids, const float* __restrict__ psi,
                                            const float* __restrict__ rsum_last, int row, int v, int* col) {
    int hit = -1;
    for (int s = 0; s < st.S; ++s)
        if (part_ids[row * st.S + s] == v) hit = s;
    *col = hit >= 0 ? hit : st.S - 1;                       // idmap == -1 selects the last column (reference quirk)
    if (v == st.blank) return LOGZERO;
    if (v == st.eos) return rsum_last[row];
    return hit >= 0 ? psi[row * st.S + hit] : LOGZERO;
}

constexpr int MAXB = 8;
constexpr int FAST_MAXS = 12;             // pre-beam candidates per hyp handled by the fast path of the fusion kernel

__global__ void __launch_bounds__(256)
beam_fuse_topk_advance_kernel(const AvsrBeamState st, const float* __restrict__ dec_logp, const int* __restrict__ part_ids,
                              const float* __restrict__ psi, const float* __restrict__ rsum_last, float w_dec, float w_ctc,
                              const float* __restrict__ ctc_full, int* __restrict__ rc_last, int* __restrict__ rc_chain,
                              int* __restrict__ rc_tok, int* __restrict__ any_running, int* __restrict__ ticket) {
    __shared__ float cval[256 * MAXB];
    __shared__ int cidx[256 * MAXB];
    __shared__ float redv[8];
    __shared__ int redi[8];
    __shared__ int redo[8];
    __shared__ float selv[MAXB];
    __shared__ int seli[MAXB];
    __shared__ int s_parent[MAXB];
    __shared__ int s_newcnt;
    const int b = blockIdx.x;
    pdl_trigger();
    pdl_wait();
    const int nrun = st.n_run[b];
    const int step = *st.step;
    // any_running != NULL: this launch also closes the position (what avsr_beam_step_advance does in a launch of its own): the
    // LAST CTA to get here counts the utterances that still run and advances *step; every CTA has read *step by then
    auto close_position = [&]() {
        if (any_running == nullptr || threadIdx.x != 0) return;
        __threadfence();                               // this CTA's n_run is visible before its ticket
        if (atomicAdd(ticket, 1) == (int)gridDim.x - 1) {
            *ticket = 0;                               // re-armed for the next position
            __threadfence();
            int live = 0;
            for (int u = 0; u < st.B; ++u) live += __ldcg(st.n_run + u) > 0;
            *any_running = live;
            *const_cast<int*>(st.step) = step + 1;      // the struct carries it read-only for everybody else
        }
    };
    if (nrun == 0) { close_position(); return; }
    const int beam = st.beam, V = st.V, S = st.S;
    const int base = b * beam;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- fast path.  Only the pre-beam candidates and eos have a CTC score other than logzero (batch_beam_search.py:229-247:
    // the partial scorer fills everything else with -1e10), so the top-`beam` of the flattened [nrun, V] fused scores are
    // found among those <= nrun * (S + 1) entries whenever `beam` of them beat the best score any other entry could have.
    __shared__ float fw[MAXB * (FAST_MAXS + 1)];
    __shared__ int fi[MAXB * (FAST_MAXS + 1)];
    __shared__ int s_fast;
    if (tid == 0) s_fast = 0;
    __syncthreads();
    if (ctc_full == nullptr && S <= FAST_MAXS && warp == 0) {
        const int ncand = nrun * (S + 1);
        float bound = -INFINITY;                     // upper bound of every entry that is NOT a candidate (dec_logp <= 0)
        for (int h = 0; h < nrun; ++h)
            bound = fmaxf(bound, __fadd_rn(__fmul_rn(w_ctc, __fsub_rn(LOGZERO, st.s_prev[base + h])), st.score[base + h]));
        for (int c = lane; c < ncand; c += 32) {
            const int h = c / (S + 1), s = c - h * (S + 1);
            const int row = base + h;
            const int v = (s < S) ? part_ids[row * S + s] : st.eos;
            float w = -INFINITY;
            int ix = 0x7fffffff;
            if (!(s < S && (v == st.eos || v == st.blank))) {        // eos is listed once; blank scores logzero like a non-candidate
                const float lpsi = (s < S) ? psi[row * S + s] : rsum_last[row];
                const float cs = __fsub_rn(lpsi, st.s_prev[row]);
                w = __fadd_rn(__fadd_rn(__fmul_rn(w_dec, dec_logp[(long long)row * V + v]), __fmul_rn(w_ctc, cs)), st.score[row]);
                ix = h * V + v;
            }
            fw[c] = w;
            fi[c] = ix;
        }
        __syncwarp();
        bool ok = true;
        for (int j = 0; j < beam; ++j) {
            float bv = -INFINITY;
            int bi = 0x7fffffff, bo = -1;
            for (int c = lane; c < ncand; c += 32) {
                const float v = fw[c];
                const int ix = fi[c];
                if (v > bv || (v == bv && ix < bi)) { bv = v; bi = ix; bo = c; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                const int oo = __shfl_xor_sync(0xffffffffu, bo, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; bo = oo; }
            }
            if (!(bv > bound)) ok = false;           // a non-candidate could tie or win: take the exhaustive path
            if (lane == 0) {
                selv[j] = bv; seli[j] = bi;
                if (bo >= 0) { fw[bo] = -INFINITY; fi[bo] = 0x7fffffff; }
            }
            __syncwarp();
        }
        if (lane == 0 && ok) s_fast = 1;
    }
    __syncthreads();
    if (!s_fast) {
    // ---- exhaustive path: per-thread top-`beam` over the flattened [nrun, V] fused scores
    float lv[MAXB];
    int li[MAXB];
#pragma unroll
    for (int k = 0; k < MAXB; ++k) { lv[k] = -INFINITY; li[k] = 0x7fffffff; }
    const int total = nrun * V;
    for (int i = tid; i < total; i += 256) {
        const int h = i / V, v = i - h * V;
        const int row = base + h;
        float cs;
        if (ctc_full != nullptr) {
            cs = ctc_full[(long long)row * V + v];               // full-vocabulary CTC scores (log_psi - s_prev) of this hyp
        } else {
            float lpsi = LOGZERO;
            if (v == st.eos) lpsi = rsum_last[row];
            else if (v != st.blank) {
                for (int s = 0; s < S; ++s)
                    if (part_ids[row * S + s] == v) lpsi = psi[row * S + s];
            }
            cs = __fsub_rn(lpsi, st.s_prev[row]);
        }
        const float w = __fadd_rn(__fadd_rn(__fmul_rn(w_dec, dec_logp[(long long)row * V + v]), __fmul_rn(w_ctc, cs)), st.score[row]);
        if (w > lv[beam - 1]) {                       // strict: on ties the lower flat index (seen first) stays
            int k = beam - 1;
            while (k > 0 && w > lv[k - 1]) { lv[k] = lv[k - 1]; li[k] = li[k - 1]; --k; }
            lv[k] = w; li[k] = i;
        }
    }
#pragma unroll
    for (int k = 0; k < MAXB; ++k) { cval[tid * MAXB + k] = lv[k]; cidx[tid * MAXB + k] = li[k]; }
    __syncthreads();
    // ---- `beam` rounds of block arg-max (value desc, flat index asc)
    for (int j = 0; j < beam; ++j) {
        float bv = -INFINITY;
        int bi = 0x7fffffff, bo = -1;
        for (int k = 0; k < beam; ++k) {
            const float v = cval[tid * MAXB + k];
            const int ix = cidx[tid * MAXB + k];
            if (v > bv || (v == bv && ix < bi)) { bv = v; bi = ix; bo = tid * MAXB + k; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const int oo = __shfl_xor_sync(0xffffffffu, bo, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; bo = oo; }
        }
        if (lane == 0) { redv[warp] = bv; redi[warp] = bi; redo[warp] = bo; }
        __syncthreads();
        if (tid == 0) {
            for (int w2 = 1; w2 < 8; ++w2)
                if (redv[w2] > bv || (redv[w2] == bv && redi[w2] < bi)) { bv = redv[w2]; bi = redi[w2]; bo = redo[w2]; }
            selv[j] = bv; seli[j] = bi;
            if (bo >= 0) { cval[bo] = -INFINITY; cidx[bo] = 0x7fffffff; }
        }
        __syncthreads();
    }

    }

    // ---- bookkeeping (serial over <= beam candidates)
    if (tid == 0) {
        // maxlen of the utterance: its frame count, or what maxlenratio asks for (beam_search.py:349-354)
        const int maxlen = st.utt_maxlen != nullptr ? st.utt_maxlen[b] : st.utt_T[b];
        const bool is_last = (step == maxlen - 1);
        float o_dec[MAXB], o_ctc[MAXB], o_sp[MAXB];
        int o_last[MAXB], o_chain[MAXB];
        for (int h = 0; h < nrun; ++h) {
            o_dec[h] = st.dec_sc[base + h]; o_ctc[h] = st.ctc_sc[base + h]; o_sp[h] = st.s_prev[base + h];
            o_last[h] = st.last_tok[base + h]; o_chain[h] = st.rprev_idx[base + h];
        }
        int n_tok[MAXB], n_ridx[MAXB];
        float n_score[MAXB], n_dec[MAXB], n_ctc[MAXB], n_sp[MAXB];
        int cnt = 0;
        const long long hb = ((long long)b * st.tmax + step) * beam;
        for (int j = 0; j < beam; ++j) {
            const int idx = seli[j];
            const int h = idx / V, v = idx - h * V;
            const int row = base + h;
            int col = 0;
            float lpsi = 0.f, cs;
            if (ctc_full != nullptr) cs = ctc_full[(long long)row * V + v];
            else {
                lpsi = ctc_logpsi(st, part_ids, psi, rsum_last, row, v, &col);
                cs = __fsub_rn(lpsi, o_sp[h]);
            }
            const float nd = __fadd_rn(o_dec[h], dec_logp[(long long)row * V + v]);
            const float nc = __fadd_rn(o_ctc[h], cs);
            st.hist_tok[hb + j] = v;
            st.hist_prev[hb + j] = h;
            if (v == st.eos || is_last) {
                const int e = st.n_ended[b]++;
                if (e < st.cap) {
                    const long long eb = (long long)b * st.cap + e;
                    st.end_step[eb] = step; st.end_j[eb] = j; st.end_score[eb] = selv[j]; st.end_dec[eb] = nd; st.end_ctc[eb] = nc;
                    const int len = step + 2 + (is_last ? 1 : 0);
                    st.end_len[eb] = len;
                    float* bl = st.best_len + (long long)b * (st.tmax + 4);
                    if (selv[j] > bl[len]) bl[len] = selv[j];
                    if (selv[j] > st.best_all[b]) st.best_all[b] = selv[j];
                } else {
                    st.overflow[0] = 1;
                }
            } else {
                n_tok[cnt] = v; n_score[cnt] = selv[j]; n_dec[cnt] = nd; n_ctc[cnt] = nc; n_sp[cnt] = lpsi;
                n_ridx[cnt] = row * S + col;
                if (ctc_full != nullptr) {
                    // full-vocabulary mode: the survivor's forward variables are recomputed by the pre-beam kernel on the
                    // chosen token (S = 1) from its parent's chain; the new chain then sits at the row's own index
                    rc_last[base + cnt] = o_last[h];
                    rc_chain[base + cnt] = o_chain[h];
                    rc_tok[base + cnt] = v;
                    n_ridx[cnt] = base + cnt;
                }
                s_parent[cnt] = h;
                st.run2j[hb + cnt] = j;
                ++cnt;
            }
        }
        for (int r = 0; r < cnt; ++r) {
            st.last_tok[base + r] = n_tok[r]; st.score[base + r] = n_score[r]; st.dec_sc[base + r] = n_dec[r];
            st.ctc_sc[base + r] = n_ctc[r]; st.rprev_idx[base + r] = n_ridx[r];
            if (ctc_full == nullptr) st.s_prev[base + r] = n_sp[r];      // full mode: s_prev = log_psi from the recompute kernel
        }
        // end detection (e2e_asr_common.py:18-48): M = 3 consecutive lengths, each > |D_end| below the best
        bool fin = false;
        if (!st.no_end_detect && st.n_ended[b] > 0) {       // only consulted when maxlenratio == 0 (beam_search.py:369)
            const float* bl = st.best_len + (long long)b * (st.tmax + 4);
            int count = 0;
            for (int m = 0; m < 3; ++m) {
                const int len = step - m;
                if (len >= 0 && bl[len] > -INFINITY && (double)bl[len] - (double)st.best_all[b] < st.d_end) ++count;
            }
            fin = (count == 3);
        }
        if (fin || cnt == 0) { cnt = 0; st.done[b] = step + 1; }   // 1 + the position the search stopped at (beam_search.py:369-374)
        st.n_run[b] = cnt;
        s_newcnt = cnt;
    }
    __syncthreads();
    // ---- ancestry of the new running rows: copy the parent's history, append the parent slot at `step`
    const int cnt = s_newcnt;
    const long long asz = (long long)st.B * beam * st.lmax;
    const unsigned char* a_old = st.anc + (long long)(step & 1) * asz;
    unsigned char* a_new = st.anc + (long long)((step + 1) & 1) * asz;
    for (int r = 0; r < cnt; ++r) {
        const int par = s_parent[r];
        for (int p = tid; p < step; p += 256) a_new[(long long)(base + r) * st.lmax + p] = a_old[(long long)(base + par) * st.lmax + p];
        if (tid == 0) a_new[(long long)(base + r) * st.lmax + step] = (unsigned char)par;
    }
    for (int r = tid; r < beam; r += 256) st.row_active[base + r] = r < cnt ? 1 : 0;
    close_position();                                  // thread 0 wrote n_run[b] above, in program order
}

// Dense score matrix of one CTCPrefixScoreTH.__call__ in pre-beam mode, as the scorer plug-in API has to return it
// (src/nets/scorers/ctc.py:101-126 -> ctc_prefix_score.py:164-187): every entry logzero, the S candidates of a row psi, then
// eos = r_sum[T-1] and blank = logzero (in that order, as the reference overwrites them), all minus the row's s_prev.
__global__ void __launch_bounds__(256)
ctc_scores_dense_kernel(const float* __restrict__ psi, const float* __restrict__ rsum_last, const float* __restrict__ s_prev,
                        const int* __restrict__ part_ids, int S, int V, int blank, int eos, float* __restrict__ out) {
    const int row = blockIdx.x;
    const float sp = s_prev[row];
    float* o = out + (long long)row * V;
    const float lz = __fsub_rn(LOGZERO, sp);
    for (int v = threadIdx.x; v < V; v += blockDim.x) o[v] = lz;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) o[part_ids[row * S + s]] = __fsub_rn(psi[row * S + s], sp);    // later duplicates win, like index_put
        o[eos] = __fsub_rn(rsum_last[row], sp);
        o[blank] = lz;
    }
}

__global__ void beam_step_advance_kernel(int* step, const int* n_run, int B, int* any_running) {
    pdl_trigger();
    pdl_wait();
    int live = 0;
    for (int b = 0; b < B; ++b) live += n_run[b] > 0;
    *any_running = live;
    *step += 1;
}

}  // namespace

static int prebeam_args(PrebeamArgs* a, size_t* smem, const float* logp, int V, int ldp, int blank, const int* utt_off, const int* utt_T,
                        const int* n_run, int beam, int R, int S, const int* last_tok, const int* rprev_idx, float* r_buf, int tmax,
                        const int* step, float* psi, float* rsum_last) {
    AVSR_REQUIRE(logp && utt_off && utt_T && n_run && last_tok && rprev_idx && r_buf && step && psi && rsum_last,
                 "avsr_ctc_prefix_prebeam: null argument");
    AVSR_REQUIRE(R > 0 && S > 0 && S <= PB_MAXS && beam > 0 && tmax > 0 && ldp >= V, "avsr_ctc_prefix_prebeam: bad sizes (S <= %d)", PB_MAXS);
    const int pitch = tmax | 1;                       // odd pitch: the S chains read their columns without bank conflicts
    *smem = (size_t)(S + 3) * pitch * sizeof(float);
    AVSR_REQUIRE(*smem <= 200 * 1024, "avsr_ctc_prefix_prebeam: %d frames x %d candidates do not fit shared memory", tmax, S);
    *a = PrebeamArgs{logp, V, ldp, blank, utt_off, utt_T, n_run, beam, R, S, last_tok, rprev_idx, r_buf, tmax, pitch, step, psi, rsum_last};
    return AVSR_OK;
}

extern "C" int avsr_ctc_prefix_prebeam(const float* logp, int V, int ldp, int blank, const int* utt_off, const int* utt_T, const int* n_run,
                                       int beam, int R, int S, const int* last_tok, const int* part_ids, const int* rprev_idx,
                                       float* r_buf, int tmax, const int* step, float* psi, float* rsum_last, cudaStream_t stream) {
    AVSR_REQUIRE(part_ids, "avsr_ctc_prefix_prebeam: null argument");
    PrebeamArgs a;
    size_t smem;
    int rc = prebeam_args(&a, &smem, logp, V, ldp, blank, utt_off, utt_T, n_run, beam, R, S, last_tok, rprev_idx, r_buf, tmax, step, psi,
                          rsum_last);
    if (rc != AVSR_OK) return rc;
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        AVSR_CHECK_CUDA(cudaFuncSetAttribute(ctc_prefix_prebeam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = 200 * 1024;
    }
    AVSR_CHECK_CUDA(avsr_launch_pdl(ctc_prefix_prebeam_kernel, dim3(R), dim3(PB_THREADS), smem, stream, a, part_ids));
    return AVSR_OK;
}

// avsr_dec_logits_lsm_topk followed by avsr_ctc_prefix_prebeam in ONE launch (one CTA per hypothesis row): part [nsplit][R][V]
// partial logits + bias -> dec_logp [R][V], part_ids [R][S] (pre-beam candidates) -> psi [R][S], rsum_last [R] and the new
// forward variables in r_buf; the remaining arguments as in the two stand-alone calls.
extern "C" int avsr_dec_tail(const float* part, int nsplit, const float* bias, float* dec_logp, int* part_ids, const float* logp, int V,
                             int ldp, int blank, const int* utt_off, const int* utt_T, const int* n_run, int beam, int R, int S,
                             const int* last_tok, const int* rprev_idx, float* r_buf, int tmax, const int* step, float* psi,
                             float* rsum_last, cudaStream_t stream) {
    AVSR_REQUIRE(part && bias && dec_logp && part_ids && nsplit >= 1, "avsr_dec_tail: null argument");
    AVSR_REQUIRE(V <= 16 * LSM_THREADS, "avsr_dec_tail: vocabulary %d too large (max %d)", V, 16 * LSM_THREADS);
    PrebeamArgs a;
    size_t smem;
    int rc = prebeam_args(&a, &smem, logp, V, ldp, blank, utt_off, utt_T, n_run, beam, R, S, last_tok, rprev_idx, r_buf, tmax, step, psi,
                          rsum_last);
    if (rc != AVSR_OK) return rc;
    const int iter = (V + LSM_THREADS - 1) / LSM_THREADS;
    static bool configured = false;
    if (!configured) {
        AVSR_CHECK_CUDA(cudaFuncSetAttribute(dec_tail_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        AVSR_CHECK_CUDA(cudaFuncSetAttribute(dec_tail_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        AVSR_CHECK_CUDA(cudaFuncSetAttribute(dec_tail_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
    }
    if (iter <= 4) AVSR_CHECK_CUDA(avsr_launch_pdl(dec_tail_kernel<4>, dim3(R), dim3(LSM_THREADS), smem, stream, part, nsplit, bias, dec_logp, part_ids, a));
    else if (iter <= 10) AVSR_CHECK_CUDA(avsr_launch_pdl(dec_tail_kernel<10>, dim3(R), dim3(LSM_THREADS), smem, stream, part, nsplit, bias, dec_logp, part_ids, a));
    else AVSR_CHECK_CUDA(avsr_launch_pdl(dec_tail_kernel<16>, dim3(R), dim3(LSM_THREADS), smem, stream, part, nsplit, bias, dec_logp, part_ids, a));
    return AVSR_OK;
}

// Work split of the full-vocabulary kernel: *ncg column groups x *tsplit time splits per utterance, sized so that one wave of
// CTAs (two per SM) covers the batch.  Scratch the caller provides: part [B][tsplit][beam][V] fp32 (unused if tsplit == 1),
// tickets [B][ncg] int32 zeroed once.
extern "C" int avsr_ctc_prefix_full_plan(int B, int V, int* ncg, int* tsplit) {
    AVSR_REQUIRE(B > 0 && V > 0 && ncg && tsplit, "avsr_ctc_prefix_full_plan: bad arguments");
    int sms = 0, dev = 0;
    AVSR_CHECK_CUDA(cudaGetDevice(&dev));
    AVSR_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // two CTAs per SM, all resident at once: column groups first (at least 128 busy threads each), then time splits
    const int v4 = cdiv(V, 4);
    int g = (2 * sms) / B;
    const int gmin = cdiv(v4, FV_THREADS), gmax = cdiv(v4, 128) > gmin ? cdiv(v4, 128) : gmin;
    g = g < gmin ? gmin : (g > gmax ? gmax : g);
    *ncg = g;
    int ts = (2 * sms) / (B * *ncg);
    static int force = -1;                            // dev knob AVSR_CTC_TSPLIT
    if (force < 0) { const char* e = getenv("AVSR_CTC_TSPLIT"); force = e ? atoi(e) : 0; }
    if (force > 0) ts = force;
    *tsplit = ts < 1 ? 1 : (ts > 16 ? 16 : ts);
    return AVSR_OK;
}

// probs (optional) = exp(logp) as avsr_ctc_exp_posteriors wrote it, same pitch: the kernel then streams the posteriors
// themselves and its inner loop is pure FMA (the ex2 per posterior, 60 M MUFU operations per launch at 16 per clock and SM,
// paced the kernel more than HBM did: 63 us with, 47 us without).  The log-posteriors are still needed (exact path, eos /
// empty-prefix terms).  Results are bit-identical with and without probs.
extern "C" int avsr_ctc_prefix_full_probs(const float* logp, const float* probs, int V, int ldp, int blank, int eos, const int* utt_off,
                                          const int* utt_T, const int* n_run, int beam, int B, int S, const int* last_tok, const int* rprev_idx,
                                          const float* r_buf, int tmax, const int* step, const float* s_prev, float* scores, float* part,
                                          int* tickets, cudaStream_t stream) {
    AVSR_REQUIRE(logp && utt_off && utt_T && n_run && last_tok && rprev_idx && r_buf && step && s_prev && scores,
                 "avsr_ctc_prefix_full: null argument");
    AVSR_REQUIRE(beam >= 1 && beam <= FV_MAXH, "avsr_ctc_prefix_full: beam %d exceeds %d", beam, FV_MAXH);
    AVSR_REQUIRE(B > 0 && B <= 65535 && V > 0 && tmax > 0, "avsr_ctc_prefix_full: bad sizes");
    AVSR_REQUIRE(ldp >= V && (ldp & 3) == 0 && (reinterpret_cast<uintptr_t>(logp) & 15) == 0 && (reinterpret_cast<uintptr_t>(probs) & 15) == 0,
                 "avsr_ctc_prefix_full: posterior rows must be 16-byte aligned (pitch %d floats, base %p)", ldp, (const void*)logp);
    int ncg = 0, tsplit = 0;
    int rc = avsr_ctc_prefix_full_plan(B, V, &ncg, &tsplit);
    if (rc != AVSR_OK) return rc;
    AVSR_REQUIRE(tsplit == 1 || (part && tickets), "avsr_ctc_prefix_full: %d time splits need the scratch buffers", tsplit);
    const int cgw = cdiv(cdiv(V, 4), ncg) * 4;       // columns per group
    // shared memory: the tables of the utterance, then the row ring (as many slots as leave room for two CTAs per SM)
    const size_t tab = (((size_t)tmax * FV_ES + (size_t)2 * tmax * beam) * sizeof(float) + 127) / 128 * 128;
    const size_t slot = (size_t)FV_FR * cgw * sizeof(float);
    const size_t lim = 200 * 1024, two = 108 * 1024;
    AVSR_REQUIRE(tab + 2 * slot <= lim, "avsr_ctc_prefix_full: T*beam too large for shared memory");
    int nst = tab + 2 * slot <= two ? (int)((two - tab) / slot) : (int)((lim - tab) / slot);
    nst = nst > FV_MAXST ? FV_MAXST : nst;
    const size_t smem = tab + nst * slot;
    static bool configured = false;
    static int k_pdl = 0;                             // dev knob AVSR_CTC_PDL: let the next kernel's CTAs in early (measured neutral)
    if (!configured) {
        AVSR_CHECK_CUDA(cudaFuncSetAttribute(ctc_prefix_full_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lim));
        AVSR_CHECK_CUDA(cudaFuncSetAttribute(ctc_prefix_full_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lim));
        const char* e = getenv("AVSR_CTC_PDL");
        k_pdl = e ? atoi(e) : 0;
        configured = true;
    }
    dim3 grid(ncg * tsplit, B);
    if (probs != nullptr)
        AVSR_CHECK_CUDA(avsr_launch_pdl(ctc_prefix_full_kernel<true>, grid, dim3(FV_ALL), smem, stream, logp, probs, V, ldp, blank, eos, utt_off,
                                        utt_T, n_run, beam, B * beam, S, last_tok, rprev_idx, r_buf, tmax, step, s_prev, scores, ncg, cgw,
                                        tsplit, part, tickets, nst, (int)tab, k_pdl));
    else
        AVSR_CHECK_CUDA(avsr_launch_pdl(ctc_prefix_full_kernel<false>, grid, dim3(FV_ALL), smem, stream, logp, probs, V, ldp, blank, eos, utt_off,
                                        utt_T, n_run, beam, B * beam, S, last_tok, rprev_idx, r_buf, tmax, step, s_prev, scores, ncg, cgw,
                                        tsplit, part, tickets, nst, (int)tab, k_pdl));
    return AVSR_OK;
}

extern "C" int avsr_ctc_prefix_full(const float* logp, int V, int ldp, int blank, int eos, const int* utt_off, const int* utt_T, const int* n_run,
                                    int beam, int B, int S, const int* last_tok, const int* rprev_idx, const float* r_buf, int tmax,
                                    const int* step, const float* s_prev, float* scores, float* part, int* tickets, cudaStream_t stream) {
    return avsr_ctc_prefix_full_probs(logp, nullptr, V, ldp, blank, eos, utt_off, utt_T, n_run, beam, B, S, last_tok, rprev_idx, r_buf, tmax, step,
                                      s_prev, scores, part, tickets, stream);
}

// probs[i] = exp(logp[i]) for the n floats of a posterior block (padding included), with the very ex2.approx.ftz the
// full-vocabulary kernel applies when it is given log-posteriors only.  Once per utterance batch (n % 4 == 0, 16-byte aligned).
extern "C" int avsr_ctc_exp_posteriors(const float* logp, long long n, float* probs, cudaStream_t stream) {
    AVSR_REQUIRE(logp && probs && n > 0 && (n & 3) == 0 && ((reinterpret_cast<uintptr_t>(logp) | reinterpret_cast<uintptr_t>(probs)) & 15) == 0,
                 "avsr_ctc_exp_posteriors: bad arguments");
    const long long n4 = n / 4;
    const int blocks = (int)((n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16);
    ctc_exp_posteriors_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const float4*>(logp), n4, reinterpret_cast<float4*>(probs));
    AVSR_LAUNCH_CHECK();
    return AVSR_OK;
}

extern "C" int avsr_beam_fuse_topk_advance(const AvsrBeamState* st, const float* dec_logp, const int* part_ids, const float* psi,
                                           const float* rsum_last, float w_dec, float w_ctc, cudaStream_t stream) {
    AVSR_REQUIRE(st && dec_logp && part_ids && psi && rsum_last, "avsr_beam_fuse_topk_advance: null argument");
    AVSR_REQUIRE(st->beam >= 1 && st->beam <= MAXB && st->B > 0, "avsr_beam_fuse_topk_advance: beam %d unsupported (max %d)", st->beam, MAXB);
    AVSR_REQUIRE(st->beam <= 255, "avsr_beam_fuse_topk_advance: ancestry slots are 8-bit");
    AVSR_CHECK_CUDA(avsr_launch_pdl(beam_fuse_topk_advance_kernel, dim3(st->B), dim3(256), 0, stream, *st, dec_logp, part_ids, psi, rsum_last,
                                    w_dec, w_ctc, (const float*)nullptr, (int*)nullptr, (int*)nullptr, (int*)nullptr, (int*)nullptr,
                                    (int*)nullptr));
    return AVSR_OK;
}

// The same, also closing the position in this launch (avsr_beam_step_advance folded in: one launch less per position):
// *any_running = utterances still running, *st->step += 1, done by the last CTA to finish; `ticket` = one int32, zero before the
// first call (the kernel re-arms it).
extern "C" int avsr_beam_fuse_topk_advance_step(const AvsrBeamState* st, const float* dec_logp, const int* part_ids, const float* psi,
                                                const float* rsum_last, float w_dec, float w_ctc, int* any_running, int* ticket,
                                                cudaStream_t stream) {
    AVSR_REQUIRE(st && dec_logp && part_ids && psi && rsum_last && any_running && ticket, "avsr_beam_fuse_topk_advance_step: null argument");
    AVSR_REQUIRE(st->beam >= 1 && st->beam <= MAXB && st->B > 0, "avsr_beam_fuse_topk_advance: beam %d unsupported (max %d)", st->beam, MAXB);
    AVSR_REQUIRE(st->beam <= 255, "avsr_beam_fuse_topk_advance: ancestry slots are 8-bit");
    AVSR_CHECK_CUDA(avsr_launch_pdl(beam_fuse_topk_advance_kernel, dim3(st->B), dim3(256), 0, stream, *st, dec_logp, part_ids, psi, rsum_last,
                                    w_dec, w_ctc, (const float*)nullptr, (int*)nullptr, (int*)nullptr, (int*)nullptr, any_running, ticket));
    return AVSR_OK;
}

// Full-vocabulary form (single-scorer search, ctc_weight = 1.0: src/avhubert_avsr/avhubert_avsr_model.py:35 drops the decoder
// and the pre-beam): fused score = w_dec * dec_logp + w_ctc * ctc_full + previous score over all n_h * V entries, with
// ctc_full [R][V] from avsr_ctc_prefix_full (dec_logp may be a zero matrix with w_dec = 0).  For every new running hyp the
// kernel leaves what avsr_ctc_prefix_prebeam (S = 1) needs to recompute its forward variables: rc_last [R] = the parent's
// last token, rc_chain [R] = the parent's chain index, rc_tok [R] = the chosen token; st->s_prev is not touched (the
// recompute's psi output is the new log_psi) and st->rprev_idx[row] = row.  The state must have S = 1.
extern "C" int avsr_beam_fuse_topk_advance_full(const AvsrBeamState* st, const float* dec_logp, const float* ctc_full, float w_dec,
                                                float w_ctc, int* rc_last, int* rc_chain, int* rc_tok, cudaStream_t stream) {
    AVSR_REQUIRE(st && dec_logp && ctc_full && rc_last && rc_chain && rc_tok, "avsr_beam_fuse_topk_advance_full: null argument");
    AVSR_REQUIRE(st->beam >= 1 && st->beam <= MAXB && st->B > 0 && st->S == 1, "avsr_beam_fuse_topk_advance_full: beam %d / S %d unsupported",
                 st->beam, st->S);
    AVSR_CHECK_CUDA(avsr_launch_pdl(beam_fuse_topk_advance_kernel, dim3(st->B), dim3(256), 0, stream, *st, dec_logp, (const int*)nullptr,
                                    (const float*)nullptr, (const float*)nullptr, w_dec, w_ctc, ctc_full, rc_last, rc_chain, rc_tok,
                                    (int*)nullptr, (int*)nullptr));
    return AVSR_OK;
}

extern "C" int avsr_beam_step_advance(int* step, const int* n_run, int B, int* any_running, cudaStream_t stream) {
    AVSR_REQUIRE(step && n_run && any_running && B > 0, "avsr_beam_step_advance: bad arguments");
    AVSR_CHECK_CUDA(avsr_launch_pdl(beam_step_advance_kernel, dim3(1), dim3(1), 0, stream, step, n_run, B, any_running));
    return AVSR_OK;
}

// scores [n][V] = CTCPrefixScoreTH.__call__'s first return value (log_psi - s_prev) from the outputs of
// avsr_ctc_prefix_prebeam: psi [n][S], rsum_last [n], part_ids [n][S], s_prev [n].
extern "C" int avsr_ctc_scores_dense(const float* psi, const float* rsum_last, const float* s_prev, const int* part_ids, int n, int S, int V,
                                     int blank, int eos, float* scores, cudaStream_t stream) {
    AVSR_REQUIRE(psi && rsum_last && s_prev && part_ids && scores && n > 0 && S > 0 && V > 0 && blank >= 0 && blank < V && eos >= 0 && eos < V,
                 "avsr_ctc_scores_dense: bad arguments");
    ctc_scores_dense_kernel<<<n, 256, 0, stream>>>(psi, rsum_last, s_prev, part_ids, S, V, blank, eos, scores);
    AVSR_LAUNCH_CHECK();
    return AVSR_OK;
}
