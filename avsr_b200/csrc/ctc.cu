// CTC prefix scoring, score fusion, top-k pruning and beam bookkeeping of the joint CTC/attention beam search.
// Reference: CTCPrefixScoreTH.__call__ (src/nets/ctc_prefix_score.py:68-187; scalar restatement in SURVEY.md App. B),
// CTCPrefixScorer.select_state (src/nets/scorers/ctc.py:40-63), BatchBeamSearch.search / batch_beam / post_process
// (src/nets/batch_beam_search.py:86-110,208-349), end_detect (src/nets/e2e_asr_common.py:18-48).
// The reference runs ~16.7k ATen ops + per-hyp host syncs per step here; this file does it in two launches per step
// for all utterances at once, with the hypothesis state resident on the device.
#include "common.cuh"

namespace {

constexpr float LOGZERO = -10000000000.0f;   // ctc_prefix_score.py:33

// torch.logsumexp over two elements: m + log(exp(a-m) + exp(b-m)).
__device__ __forceinline__ float lse2(float a, float b) {
    const float m = fmaxf(a, b), n = fminf(a, b);
    return m + logf(1.f + expf(n - m));
}

// ------------------------------------------------------------------------------------------------------------------
// Pre-beam mode: one thread per (row, candidate) owns one forward chain over time.
// r_buf [2][R*S][tmax][2] ping-pongs on step parity; rprev_idx[row] is the chain (in the "current" half) that the
// surviving hypothesis inherited.  Outputs psi[row][s] (log prefix probability) and rsum_last[row] = r_sum[T-1].
__global__ void __launch_bounds__(128)
ctc_prefix_prebeam_kernel(const float* __restrict__ logp, int V, int blank, const int* __restrict__ utt_off,
                          const int* __restrict__ utt_T, const int* __restrict__ n_run, int beam, int R, int S,
                          const int* __restrict__ last_tok, const int* __restrict__ part_ids, const int* __restrict__ rprev_idx,
                          float* __restrict__ r_buf, int tmax, const int* __restrict__ step_p, float* __restrict__ psi,
                          float* __restrict__ rsum_last) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= R * S) return;
    const int row = gid / S, s = gid % S;
    const int utt = row / beam;
    if ((row % beam) >= n_run[utt]) return;
    const int step = *step_p;
    const int T = utt_T[utt];
    const float* lp = logp + (long long)utt_off[utt] * V;
    const int c = part_ids[row * S + s];
    const bool same = (c == last_tok[row]);
    const int cur = step & 1;
    const float2* rp = reinterpret_cast<const float2*>(r_buf) + ((long long)cur * R * S + (step > 0 ? rprev_idx[row] : 0)) * tmax;
    float2* ro = reinterpret_cast<float2*>(r_buf) + ((long long)(cur ^ 1) * R * S + gid) * tmax;

    const int start = step > 1 ? step : 1;
    float rn = (step == 0) ? lp[c] : LOGZERO;
    float rb = LOGZERO;
    ro[start - 1] = make_float2(rn, rb);
    float M = rn, Ssum = 1.f;                                  // running logsumexp of {rn[start-1]} U {phi[t-1] + x[t]}
    float cum = 0.f;                                           // step 0: running sum of blank log-probs
    float pn = LOGZERO, pb = LOGZERO;                          // previous-label chains at t-1
    if (step == 0) {
        for (int u = 0; u < start; ++u) cum += lp[(long long)u * V + blank];
    }
    // The chain is serial in t, but its inputs are not: fetch them CH steps ahead (double-buffered registers) so that the
    // ~1 us global-load latency overlaps the logaddexp arithmetic instead of being paid once per time step.
    constexpr int CH = 8;
    float xa[CH], xba[CH], xn[CH], xbn[CH];
    float2 pa[CH], pnx[CH];
    auto fetch = [&](int t0, float (&xs)[CH], float (&xbs)[CH], float2 (&ps)[CH]) {
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            const int t = t0 + u;
            if (t < T) {
                xs[u] = __ldg(lp + (long long)t * V + c);
                xbs[u] = __ldg(lp + (long long)t * V + blank);
                if (step > 0) ps[u] = rp[t - 1];
            }
        }
    };
    fetch(start, xa, xba, pa);
    for (int t0 = start; t0 < T; t0 += CH) {
        fetch(t0 + CH, xn, xbn, pnx);
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            const int t = t0 + u;
            if (t < T) {
                if (step == 0) { pn = LOGZERO; pb = cum; }
                else { pn = pa[u].x; pb = pa[u].y; }
                const float x = xa[u], xb = xba[u];
                const float phi = same ? pb : lse2(pn, pb);
                const float term = phi + x;
                if (term > M) { Ssum = Ssum * expf(M - term) + 1.f; M = term; }
                else Ssum += expf(term - M);
                const float nrn = lse2(rn, phi) + x;
                const float nrb = lse2(rn, rb) + xb;
                rn = nrn; rb = nrb;
                ro[t] = make_float2(rn, rb);
                if (step == 0) cum += xb;
            }
        }
#pragma unroll
        for (int u = 0; u < CH; ++u) { xa[u] = xn[u]; xba[u] = xbn[u]; pa[u] = pnx[u]; }
    }
    psi[gid] = M + logf(Ssum);
    if (s == 0) {
        float en, eb;
        if (step == 0) {
            // cum currently holds sum_{u<T} x[u,blank] only when the loop ran to T; recompute for clarity
            float cs = 0.f;
            for (int u = 0; u < T; ++u) cs += lp[(long long)u * V + blank];
            en = LOGZERO; eb = cs;
        } else {
            const float2 v = rp[T - 1];
            en = v.x; eb = v.y;
        }
        rsum_last[row] = lse2(en, eb);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Full-vocabulary mode (scoring_ids=None, ctc_prefix_score.py:115-119): one thread per token column scores every live
// hypothesis of its utterance, so each log-posterior row is read from HBM once (coalesced) and shared by all hyps.
// The previous-label chains and r_sum are staged in shared memory per CTA.  Writes only scores[row][V]
// (= log_psi - s_prev); survivor chains are recomputed by the pre-beam kernel on the chosen tokens.
constexpr int FV_MAXH = 8;
__global__ void __launch_bounds__(128)
ctc_prefix_full_kernel(const float* __restrict__ logp, int V, int blank, int eos, const int* __restrict__ utt_off,
                       const int* __restrict__ utt_T, const int* __restrict__ n_run, int beam, int R, int S,
                       const int* __restrict__ last_tok, const int* __restrict__ rprev_idx, const float* __restrict__ r_buf,
                       int tmax, const int* __restrict__ step_p, const float* __restrict__ s_prev, float* __restrict__ scores) {
    extern __shared__ float sm[];                   // [T][nh] phi_same (= rb_prev), [T][nh] r_sum
    const int utt = blockIdx.y;
    const int nh = n_run[utt];
    if (nh == 0) return;
    const int step = *step_p;
    const int T = utt_T[utt];
    const float* lp = logp + (long long)utt_off[utt] * V;
    float* s_pb = sm;
    float* s_rs = sm + (size_t)T * nh;
    const int cur = step & 1;
    for (int i = threadIdx.x; i < T * nh; i += blockDim.x) {
        const int t = i / nh, h = i % nh;
        float pn, pb;
        if (step == 0) {
            pn = LOGZERO;
            float cs = 0.f;                         // cumulative blank log-prob (first call only; O(T^2/2) adds per CTA)
            for (int u = 0; u <= t; ++u) cs += lp[(long long)u * V + blank];
            pb = cs;
        } else {
            const float2 v = (reinterpret_cast<const float2*>(r_buf) + ((long long)cur * R * S + rprev_idx[utt * beam + h]) * tmax)[t];
            pn = v.x; pb = v.y;
        }
        s_pb[i] = pb;
        s_rs[i] = lse2(pn, pb);
    }
    __syncthreads();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= V) return;
    const int start = step > 1 ? step : 1;
    float rn[FV_MAXH], rb[FV_MAXH], M[FV_MAXH], Ss[FV_MAXH];
    int lastt[FV_MAXH];
#pragma unroll
    for (int h = 0; h < FV_MAXH; ++h) {
        rn[h] = (step == 0) ? lp[c] : LOGZERO;
        rb[h] = LOGZERO;
        M[h] = rn[h];
        Ss[h] = 1.f;
        lastt[h] = h < nh ? last_tok[utt * beam + h] : -1;
    }
    for (int t = start; t < T; ++t) {
        const float x = lp[(long long)t * V + c];
        const float xb = lp[(long long)t * V + blank];
#pragma unroll
        for (int h = 0; h < FV_MAXH; ++h) {
            if (h < nh) {
                const float phi = (c == lastt[h]) ? s_pb[(t - 1) * nh + h] : s_rs[(t - 1) * nh + h];
                const float term = phi + x;
                if (term > M[h]) { Ss[h] = Ss[h] * expf(M[h] - term) + 1.f; M[h] = term; }
                else Ss[h] += expf(term - M[h]);
                const float nrn = lse2(rn[h], phi) + x;
                rb[h] = lse2(rn[h], rb[h]) + xb;
                rn[h] = nrn;
            }
        }
    }
#pragma unroll
    for (int h = 0; h < FV_MAXH; ++h) {
        if (h < nh) {
            float lpsi = M[h] + logf(Ss[h]);
            if (c == eos) lpsi = s_rs[(T - 1) * nh + h];
            if (c == blank) lpsi = LOGZERO;
            scores[(long long)(utt * beam + h) * V + c] = lpsi - s_prev[utt * beam + h];
        }
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

__global__ void __launch_bounds__(256)
beam_fuse_topk_advance_kernel(const AvsrBeamState st, const float* __restrict__ dec_logp, const int* __restrict__ part_ids,
                              const float* __restrict__ psi, const float* __restrict__ rsum_last, float w_dec, float w_ctc) {
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
    const int nrun = st.n_run[b];
    if (nrun == 0) return;
    const int beam = st.beam, V = st.V, S = st.S;
    const int base = b * beam;
    const int step = *st.step;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- per-thread top-`beam` over the flattened [nrun, V] fused scores
    float lv[MAXB];
    int li[MAXB];
#pragma unroll
    for (int k = 0; k < MAXB; ++k) { lv[k] = -INFINITY; li[k] = 0x7fffffff; }
    const int total = nrun * V;
    for (int i = tid; i < total; i += 256) {
        const int h = i / V, v = i - h * V;
        const int row = base + h;
        float lpsi = LOGZERO;
        if (v == st.eos) lpsi = rsum_last[row];
        else if (v != st.blank) {
            for (int s = 0; s < S; ++s)
                if (part_ids[row * S + s] == v) lpsi = psi[row * S + s];
        }
        const float cs = __fsub_rn(lpsi, st.s_prev[row]);
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

    // ---- bookkeeping (serial over <= beam candidates)
    if (tid == 0) {
        const int T = st.utt_T[b];
        const bool is_last = (step == T - 1);
        float o_dec[MAXB], o_ctc[MAXB], o_sp[MAXB];
        for (int h = 0; h < nrun; ++h) { o_dec[h] = st.dec_sc[base + h]; o_ctc[h] = st.ctc_sc[base + h]; o_sp[h] = st.s_prev[base + h]; }
        int n_tok[MAXB], n_ridx[MAXB];
        float n_score[MAXB], n_dec[MAXB], n_ctc[MAXB], n_sp[MAXB];
        int cnt = 0;
        const long long hb = ((long long)b * st.tmax + step) * beam;
        for (int j = 0; j < beam; ++j) {
            const int idx = seli[j];
            const int h = idx / V, v = idx - h * V;
            const int row = base + h;
            int col;
            const float lpsi = ctc_logpsi(st, part_ids, psi, rsum_last, row, v, &col);
            const float cs = __fsub_rn(lpsi, o_sp[h]);
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
                s_parent[cnt] = h;
                st.run2j[hb + cnt] = j;
                ++cnt;
            }
        }
        for (int r = 0; r < cnt; ++r) {
            st.last_tok[base + r] = n_tok[r]; st.score[base + r] = n_score[r]; st.dec_sc[base + r] = n_dec[r];
            st.ctc_sc[base + r] = n_ctc[r]; st.s_prev[base + r] = n_sp[r]; st.rprev_idx[base + r] = n_ridx[r];
        }
        // end detection (e2e_asr_common.py:18-48): M = 3 consecutive lengths, each > |D_end| below the best
        bool fin = false;
        if (st.n_ended[b] > 0) {
            const float* bl = st.best_len + (long long)b * (st.tmax + 4);
            int count = 0;
            for (int m = 0; m < 3; ++m) {
                const int len = step - m;
                if (len >= 0 && bl[len] > -INFINITY && (double)bl[len] - (double)st.best_all[b] < st.d_end) ++count;
            }
            fin = (count == 3);
        }
        if (fin || cnt == 0) { cnt = 0; st.done[b] = 1; }
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
}

__global__ void beam_step_advance_kernel(int* step, const int* n_run, int B, int* any_running) {
    int live = 0;
    for (int b = 0; b < B; ++b) live += n_run[b] > 0;
    *any_running = live;
    *step += 1;
}

}  // namespace

extern "C" int avsr_ctc_prefix_prebeam(const float* logp, int V, int blank, const int* utt_off, const int* utt_T, const int* n_run,
                                       int beam, int R, int S, const int* last_tok, const int* part_ids, const int* rprev_idx,
                                       float* r_buf, int tmax, const int* step, float* psi, float* rsum_last, cudaStream_t stream) {
    AVSR_REQUIRE(logp && utt_off && utt_T && n_run && last_tok && part_ids && rprev_idx && r_buf && step && psi && rsum_last,
                 "avsr_ctc_prefix_prebeam: null argument");
    AVSR_REQUIRE(R > 0 && S > 0 && beam > 0 && tmax > 0, "avsr_ctc_prefix_prebeam: bad sizes");
    ctc_prefix_prebeam_kernel<<<cdiv((long long)R * S, 128), 128, 0, stream>>>(logp, V, blank, utt_off, utt_T, n_run, beam, R, S, last_tok,
                                                                             part_ids, rprev_idx, r_buf, tmax, step, psi, rsum_last);
    AVSR_LAUNCH_CHECK();
    return AVSR_OK;
}

extern "C" int avsr_ctc_prefix_full(const float* logp, int V, int blank, int eos, const int* utt_off, const int* utt_T, const int* n_run,
                                    int beam, int B, int S, const int* last_tok, const int* rprev_idx, const float* r_buf, int tmax,
                                    const int* step, const float* s_prev, float* scores, cudaStream_t stream) {
    AVSR_REQUIRE(logp && utt_off && utt_T && n_run && last_tok && rprev_idx && r_buf && step && s_prev && scores,
                 "avsr_ctc_prefix_full: null argument");
    AVSR_REQUIRE(beam <= FV_MAXH, "avsr_ctc_prefix_full: beam %d exceeds %d", beam, FV_MAXH);
    const size_t smem = (size_t)2 * tmax * beam * sizeof(float);
    AVSR_REQUIRE(smem <= 48 * 1024, "avsr_ctc_prefix_full: T*beam too large for shared memory");
    dim3 grid(cdiv(V, 128), B);
    ctc_prefix_full_kernel<<<grid, 128, smem, stream>>>(logp, V, blank, eos, utt_off, utt_T, n_run, beam, B * beam, S, last_tok, rprev_idx,
                                                       r_buf, tmax, step, s_prev, scores);
    AVSR_LAUNCH_CHECK();
    return AVSR_OK;
}

extern "C" int avsr_beam_fuse_topk_advance(const AvsrBeamState* st, const float* dec_logp, const int* part_ids, const float* psi,
                                           const float* rsum_last, float w_dec, float w_ctc, cudaStream_t stream) {
    AVSR_REQUIRE(st && dec_logp && part_ids && psi && rsum_last, "avsr_beam_fuse_topk_advance: null argument");
    AVSR_REQUIRE(st->beam >= 1 && st->beam <= MAXB && st->B > 0, "avsr_beam_fuse_topk_advance: beam %d unsupported (max %d)", st->beam, MAXB);
    AVSR_REQUIRE(st->beam <= 255, "avsr_beam_fuse_topk_advance: ancestry slots are 8-bit");
    beam_fuse_topk_advance_kernel<<<st->B, 256, 0, stream>>>(*st, dec_logp, part_ids, psi, rsum_last, w_dec, w_ctc);
    AVSR_LAUNCH_CHECK();
    return AVSR_OK;
}

extern "C" int avsr_beam_step_advance(int* step, const int* n_run, int B, int* any_running, cudaStream_t stream) {
    AVSR_REQUIRE(step && n_run && any_running && B > 0, "avsr_beam_step_advance: bad arguments");
    beam_step_advance_kernel<<<1, 1, 0, stream>>>(step, n_run, B, any_running);
    AVSR_LAUNCH_CHECK();
    return AVSR_OK;
}
