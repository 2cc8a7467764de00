// Per-step kernels of the transformer decoder scorer (fp32), KV-cache form of the reference's output-cache step:
// src/nets/backend/transformer/decoder.py:153-227 (forward_one_step / batch_score), decoder_layer.py:58-121,
// attention.py:38-106, embedding.py:78-87, layer_norm.py:12-33 (eps 1e-12).  SURVEY.md App. A restates the math.
//
// Rows: R = n_utt * beam, row = utt * beam + slot; slot < n_run[utt] is live.  All hyps of all utterances are at the same
// position `*step`.  Self-attention K/V of position p are stored at [layer][p][row][1024] by the row
// that computed them; a hyp finds its history through anc[row][p] = slot (within its utterance) that holds position p,
// so beam reordering never copies the cache (reference copies per-hyp caches in batch_beam_search.py:250-284).
#include "common.cuh"

namespace {

constexpr int D = 1024;
constexpr int HEADS = 16;
constexpr int DH = 64;

// x = sqrt(D) * E[tok] + PE[step]; a = LayerNorm(x) (first layer's norm1).  One CTA (256 threads) per row.
__global__ void __launch_bounds__(256)
dec_embed_ln_kernel(const float* __restrict__ emb, const float* __restrict__ pe, const int* __restrict__ last_tok,
                    const int* __restrict__ n_run, int beam, const int* __restrict__ step_p, const float* __restrict__ g,
                    const float* __restrict__ b, float eps, float* __restrict__ x, float* __restrict__ a,
                    __nv_bfloat16* __restrict__ a_split) {
    __shared__ float red[32];
    const int row = blockIdx.x;
    if ((row % beam) >= n_run[row / beam]) return;
    const int step = *step_p;
    const int c = threadIdx.x * 4;
    const float4 e = *reinterpret_cast<const float4*>(emb + (long long)last_tok[row] * D + c);
    const float4 p = *reinterpret_cast<const float4*>(pe + (long long)step * D + c);
    float4 v;
    v.x = e.x * 32.f + p.x; v.y = e.y * 32.f + p.y; v.z = e.z * 32.f + p.z; v.w = e.w * 32.f + p.w;
    *reinterpret_cast<float4*>(x + (long long)row * D + c) = v;
    const float mean = block_sum(v.x + v.y + v.z + v.w, red) / (float)D;
    const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
    const float rstd = rsqrtf(block_sum(d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3, red) / (float)D + eps);
    const float4 gg = *reinterpret_cast<const float4*>(g + c);
    const float4 bb = *reinterpret_cast<const float4*>(b + c);
    const float4 y = make_float4(d0 * rstd * gg.x + bb.x, d1 * rstd * gg.y + bb.y, d2 * rstd * gg.z + bb.z, d3 * rstd * gg.w + bb.w);
    if (a) *reinterpret_cast<float4*>(a + (long long)row * D + c) = y;
    if (a_split) {
        __nv_bfloat16* sr = a_split + (long long)row * 6 * D;
        avsr_split3_store(sr, D, c, y.x); avsr_split3_store(sr, D, c + 1, y.y);
        avsr_split3_store(sr, D, c + 2, y.z); avsr_split3_store(sr, D, c + 3, y.w);
    }
}

// Single-query attention: one CTA (4 warps) per (row, head); the keys are split across the warps and the partial
// (max, sum, weighted V) results are merged through shared memory.  MODE 0: self-attention over the cached positions
// 0..step-1 plus the current token (whose k, v are read from qkv and appended to the cache).  MODE 1: cross-attention
// over the T frames of the row's utterance.
template <int MODE>
__global__ void __launch_bounds__(128)
dec_attn_step_kernel(const float* __restrict__ q_in, long long ldq, float* __restrict__ kc, float* __restrict__ vc,
                     const unsigned char* __restrict__ anc, int lmax, const int* __restrict__ n_run, const int* __restrict__ utt_off,
                     const int* __restrict__ utt_T, int beam, int R, const int* __restrict__ step_p, float* __restrict__ out,
                     int smax, long long kv_ld, __nv_bfloat16* __restrict__ out_split) {
    extern __shared__ float sc_all[];          // [smax] scores / probabilities of all keys
    __shared__ float s_m[4], s_s[4];
    __shared__ float s_o[4][DH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x;
    const int head = blockIdx.y;
    const int utt = row / beam;
    if ((row % beam) >= n_run[utt]) return;
    const int step = *step_p;

    float q[DH];
    const float* qp = q_in + (long long)row * ldq + head * DH;
#pragma unroll
    for (int i = 0; i < DH; i += 4) {
        const float4 t = *reinterpret_cast<const float4*>(qp + i);
        q[i] = t.x; q[i + 1] = t.y; q[i + 2] = t.z; q[i + 3] = t.w;
    }
    int n;                       // number of keys
    const float* kbase;
    const float* vbase;
    long long kstride;
    if (MODE == 0) {
        n = step + 1;
        if (warp == 0) {         // append this position's k, v to the cache (slot = this row)
            const float* kp = qp + D;
            const float* vp = qp + 2 * D;
            float* kd = kc + ((long long)step * R + row) * D + head * DH;
            float* vd = vc + ((long long)step * R + row) * D + head * DH;
            kd[lane] = kp[lane]; kd[lane + 32] = kp[lane + 32];
            vd[lane] = vp[lane]; vd[lane + 32] = vp[lane + 32];
        }
        kbase = kc + head * DH;
        vbase = vc + head * DH;
        kstride = (long long)R * D;
    } else {
        n = utt_T[utt];
        kbase = kc + (long long)utt_off[utt] * kv_ld + head * DH;
        vbase = vc + (long long)utt_off[utt] * kv_ld + head * DH;
        kstride = kv_ld;
    }
    // anc is double-buffered on step parity: [2][R][lmax]; the half written at the end of step-1 is (step & 1)
    const unsigned char* arow = anc + ((long long)(step & 1) * R + row) * lmax;
    const int rbase = utt * beam;
    const int per = (n + 3) >> 2;
    const int p0 = warp * per, p1 = min(n, p0 + per);

    float mx = -INFINITY;
    for (int p = p0 + lane; p < p1; p += 32) {
        const float* kp;
        if (MODE == 0) {
            kp = (p == step) ? (qp + D) : (kbase + (long long)p * kstride + (long long)(rbase + arow[p]) * D);
        } else {
            kp = kbase + (long long)p * kstride;
        }
        float4 kk[DH / 4];
#pragma unroll
        for (int i = 0; i < DH / 4; ++i) kk[i] = *reinterpret_cast<const float4*>(kp + 4 * i);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < DH / 4; ++i) {
            s = fmaf(q[4 * i], kk[i].x, s); s = fmaf(q[4 * i + 1], kk[i].y, s);
            s = fmaf(q[4 * i + 2], kk[i].z, s); s = fmaf(q[4 * i + 3], kk[i].w, s);
        }
        s *= 0.125f;
        sc_all[p] = s;
        mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int p = p0 + lane; p < p1; p += 32) {
        const float e = expf(sc_all[p] - mx);
        sc_all[p] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
    int p = p0;
    for (; p + 1 < p1; p += 2) {
        const float* va;
        const float* vb;
        if (MODE == 0) {
            va = (p == step) ? (qp + 2 * D) : (vbase + (long long)p * kstride + (long long)(rbase + arow[p]) * D);
            vb = (p + 1 == step) ? (qp + 2 * D) : (vbase + (long long)(p + 1) * kstride + (long long)(rbase + arow[p + 1]) * D);
        } else {
            va = vbase + (long long)p * kstride;
            vb = va + kstride;
        }
        const float2 ta = *reinterpret_cast<const float2*>(va + lane * 2);
        const float2 tb = *reinterpret_cast<const float2*>(vb + lane * 2);
        const float wa = sc_all[p], wb = sc_all[p + 1];
        o0 = fmaf(wa, ta.x, o0); o1 = fmaf(wa, ta.y, o1);
        o2 = fmaf(wb, tb.x, o2); o3 = fmaf(wb, tb.y, o3);
    }
    if (p < p1) {
        const float* va;
        if (MODE == 0) va = (p == step) ? (qp + 2 * D) : (vbase + (long long)p * kstride + (long long)(rbase + arow[p]) * D);
        else va = vbase + (long long)p * kstride;
        const float2 ta = *reinterpret_cast<const float2*>(va + lane * 2);
        const float wa = sc_all[p];
        o0 = fmaf(wa, ta.x, o0); o1 = fmaf(wa, ta.y, o1);
    }
    o0 += o2; o1 += o3;
    if (lane == 0) { s_m[warp] = mx; s_s[warp] = sum; }
    s_o[warp][lane * 2] = o0;
    s_o[warp][lane * 2 + 1] = o1;
    __syncthreads();
    if (warp == 0) {
        const float M = fmaxf(fmaxf(s_m[0], s_m[1]), fmaxf(s_m[2], s_m[3]));
        float tot = 0.f, a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const float f = (s_m[w] == -INFINITY) ? 0.f : expf(s_m[w] - M);     // a warp with no keys contributes nothing
            tot += s_s[w] * f;
            a0 += s_o[w][lane * 2] * f;
            a1 += s_o[w][lane * 2 + 1] * f;
        }
        const float inv = 1.f / tot;
        if (out) *reinterpret_cast<float2*>(out + (long long)row * D + head * DH + lane * 2) = make_float2(a0 * inv, a1 * inv);
        if (out_split) {
            __nv_bfloat16* sr = out_split + (long long)row * 6 * D;
            avsr_split3_store(sr, D, head * DH + lane * 2, a0 * inv);
            avsr_split3_store(sr, D, head * DH + lane * 2 + 1, a1 * inv);
        }
    }
}


// logits = sum_z part[z][row] + bias ; logp = log_softmax(logits) -> dec_logp[row] ; part_ids[row] = top-S token ids.
__global__ void __launch_bounds__(256)
dec_logits_lsm_topk_kernel(const float* __restrict__ part, int nsplit, int R, int V, const float* __restrict__ bias,
                           const int* __restrict__ n_run, int beam, float* __restrict__ logp, int* __restrict__ part_ids, int S) {
    extern __shared__ float rowv[];
    __shared__ float red[32];
    __shared__ int redi[32];
    const int row = blockIdx.x;
    if ((row % beam) >= n_run[row / beam]) return;
    float mx = -INFINITY;
    for (int c = threadIdx.x; c < V; c += blockDim.x) {
        float v = 0.f;
        for (int z = 0; z < nsplit; ++z) v += part[((long long)z * R + row) * V + c];
        v += bias[c];
        rowv[c] = v;
        mx = fmaxf(mx, v);
    }
    mx = block_max(mx, red);
    float s = 0.f;
    for (int c = threadIdx.x; c < V; c += blockDim.x) s += expf(rowv[c] - mx);
    s = block_sum(s, red);
    const float lse = logf(s);
    for (int c = threadIdx.x; c < V; c += blockDim.x) {
        const float v = (rowv[c] - mx) - lse;
        rowv[c] = v;
        logp[(long long)row * V + c] = v;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int k = 0; k < S; ++k) {
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int c = threadIdx.x; c < V; c += blockDim.x) {
            const float v = rowv[c];
            if (v > bv) { bv = v; bi = c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { red[w] = bv; redi[w] = bi; }
        __syncthreads();
        if (w == 0) {
            bv = lane < 8 ? red[lane] : -INFINITY;
            bi = lane < 8 ? redi[lane] : 0x7fffffff;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) {
                part_ids[row * S + k] = bi;
                rowv[bi] = -INFINITY;
            }
        }
        __syncthreads();
    }
}

// log_softmax over rows of a [rows, V] fp32 matrix, in place (CTC head: src/nets/backend/ctc.py:163-170).
__global__ void __launch_bounds__(256) log_softmax_rows_kernel(float* __restrict__ x, long long ld, int V) {
    __shared__ float red[32];
    float* r = x + (long long)blockIdx.x * ld;
    float mx = -INFINITY;
    for (int c = threadIdx.x; c < V; c += blockDim.x) mx = fmaxf(mx, r[c]);
    mx = block_max(mx, red);
    float s = 0.f;
    for (int c = threadIdx.x; c < V; c += blockDim.x) s += expf(r[c] - mx);
    s = block_sum(s, red);
    const float lse = logf(s);
    for (int c = threadIdx.x; c < V; c += blockDim.x) r[c] = (r[c] - mx) - lse;
}

}  // namespace

extern "C" int avsr_dec_embed_ln(const float* emb, const float* pe, const int* last_tok, const int* n_run, int beam, int R,
                                 const int* step, const float* gamma, const float* beta, float eps, float* x, float* a,
                                 void* a_split, cudaStream_t stream) {
    AVSR_REQUIRE(emb && pe && last_tok && n_run && step && gamma && beta && x && (a || a_split) && R > 0 && beam > 0,
                 "avsr_dec_embed_ln: bad arguments");
    dec_embed_ln_kernel<<<R, 256, 0, stream>>>(emb, pe, last_tok, n_run, beam, step, gamma, beta, eps, x, a, (__nv_bfloat16*)a_split);
    AVSR_LAUNCH_CHECK();
    return AVSR_OK;
}

// mode 0: self-attention step. q_in = qkv [R, 3072] (q | k | v of the current position), kc/vc = this layer's caches
// [lmax][R][1024]; anc [R][lmax].  mode 1: cross-attention. q_in = q [R, 1024], kc/vc = this layer's cross K/V [F][1024].
extern "C" int avsr_dec_attn_step(int mode, const float* q_in, long long ldq, float* kc, float* vc, const unsigned char* anc, int lmax,
                                  const int* n_run, const int* utt_off, const int* utt_T, int beam, int R, const int* step,
                                  float* out, int max_keys, long long kv_ld, void* out_split, cudaStream_t stream) {
    AVSR_REQUIRE(q_in && kc && vc && n_run && step && (out || out_split) && R > 0 && beam > 0 && max_keys > 0, "avsr_dec_attn_step: bad arguments");
    AVSR_REQUIRE(mode == 0 ? (anc != nullptr) : (utt_off && utt_T), "avsr_dec_attn_step: missing index arrays for mode %d", mode);
    const int smax = (max_keys + 3) & ~3;
    const size_t smem = (size_t)smax * sizeof(float);
    AVSR_REQUIRE(smem <= 48 * 1024, "avsr_dec_attn_step: %d keys exceed the shared-memory strip", max_keys);
    dim3 grid(R, HEADS);
    if (mode == 0)
        dec_attn_step_kernel<0><<<grid, 128, smem, stream>>>(q_in, ldq, kc, vc, anc, lmax, n_run, utt_off, utt_T, beam, R, step, out, smax, kv_ld, (__nv_bfloat16*)out_split);
    else
        dec_attn_step_kernel<1><<<grid, 128, smem, stream>>>(q_in, ldq, kc, vc, anc, lmax, n_run, utt_off, utt_T, beam, R, step, out, smax, kv_ld, (__nv_bfloat16*)out_split);
    AVSR_LAUNCH_CHECK();
    return AVSR_OK;
}

extern "C" int avsr_dec_logits_lsm_topk(const float* part, int nsplit, int R, int V, const float* bias, const int* n_run, int beam,
                                        float* logp, int* part_ids, int S, cudaStream_t stream) {
    AVSR_REQUIRE(part && bias && n_run && logp && part_ids && R > 0 && V > 0 && S > 0 && S <= V, "avsr_dec_logits_lsm_topk: bad arguments");
    AVSR_REQUIRE((size_t)V * 4 <= 48 * 1024, "avsr_dec_logits_lsm_topk: vocabulary %d too large for the shared-memory row", V);
    dec_logits_lsm_topk_kernel<<<R, 256, (size_t)V * 4, stream>>>(part, nsplit, R, V, bias, n_run, beam, logp, part_ids, S);
    AVSR_LAUNCH_CHECK();
    return AVSR_OK;
}

extern "C" int avsr_log_softmax_rows(float* x, long long ld, long long rows, int V, cudaStream_t stream) {
    AVSR_REQUIRE(x && rows > 0 && V > 0, "avsr_log_softmax_rows: bad arguments");
    log_softmax_rows_kernel<<<(unsigned)rows, 256, 0, stream>>>(x, ld, V);
    AVSR_LAUNCH_CHECK();
    return AVSR_OK;
}
