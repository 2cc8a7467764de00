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
        avsr_split3_store4(a_split + (long long)row * 6 * D, D, c, y);
    }
}

// Single-query attention with the K / V rows staged through shared memory in 128-key tiles (coalesced 256-byte row loads).
// MODE 1 (cross-attention): one CTA per (utterance, head); the utterance's T frames of K and V are read from HBM ONCE and
//   shared by all its live hyps (<= MAXH), which is what bounds this kernel: 2 * T * 256 B per (utterance, head).
// MODE 0 (self-attention): one CTA per (row, head); keys are the row's own history, gathered through the ancestry table,
//   plus the current token whose k, v come from qkv (and are appended to the cache here).
// Pass 1 computes all scores (thread = key), a block-wide softmax follows, pass 2 accumulates V (warp = 32 keys of the
// tile, lane = 2 output dims) and the four warps' partial sums are merged through shared memory.
constexpr int MAXH = 8;
constexpr int KT = 128;                  // keys per staged tile
constexpr int KSTR = DH + 1;             // padded row stride (floats) of the staged K tile: conflict-free row-per-thread reads

template <int MODE>
__global__ void __launch_bounds__(128)
dec_attn_step_kernel(const float* __restrict__ q_in, long long ldq, float* __restrict__ kc, float* __restrict__ vc,
                     const unsigned char* __restrict__ anc, int lmax, const int* __restrict__ n_run, const int* __restrict__ utt_off,
                     const int* __restrict__ utt_T, int beam, int R, const int* __restrict__ step_p, float* __restrict__ out,
                     int smax, long long kv_ld, long long head_stride, __nv_bfloat16* __restrict__ out_split) {
    extern __shared__ float smem[];
    float* tile = smem;                              // [KT][KSTR] (K pass) / [KT][DH] (V pass)
    float* qs = tile + KT * KSTR;                    // [nh][DH]
    float* sc = qs + MAXH * DH;                      // [nh][smax]
    __shared__ float s_red[4][MAXH];
    __shared__ float s_o[4][MAXH][DH];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int head = blockIdx.y;
    const int step = *step_p;
    int nh, row0, utt, n;
    if (MODE == 1) {
        utt = blockIdx.x;
        nh = n_run[utt];
        row0 = utt * beam;
        n = utt_T[utt];
    } else {
        row0 = blockIdx.x;
        utt = row0 / beam;
        nh = ((row0 % beam) < n_run[utt]) ? 1 : 0;
        n = step + 1;
    }
    if (nh == 0) return;
    // ancestry row staged in shared memory first, so that a key row's address does not hang on a dependent global load
    unsigned char* arow = reinterpret_cast<unsigned char*>(sc + (size_t)(MODE == 1 ? MAXH : 1) * smax);
    if (MODE == 0) {
        const unsigned char* ag = anc + ((long long)(step & 1) * R + row0) * lmax;
        for (int i = tid; i < step; i += 128) arow[i] = ag[i];
    }
    const int rbase = utt * beam;
    const float* qrow0 = q_in + (long long)row0 * ldq + head * DH;

    for (int i = tid; i < nh * DH; i += 128) qs[i] = q_in[(long long)(row0 + i / DH) * ldq + head * DH + (i % DH)];
    if (MODE == 0 && warp == 0) {                    // append this position's k, v to the cache (slot = this row)
        float* kd = kc + head * head_stride + ((long long)step * R + row0) * kv_ld;
        float* vd = vc + head * head_stride + ((long long)step * R + row0) * kv_ld;
        kd[lane] = qrow0[D + lane]; kd[lane + 32] = qrow0[D + lane + 32];
        vd[lane] = qrow0[2 * D + lane]; vd[lane + 32] = qrow0[2 * D + lane + 32];
    }
    auto src_ptr = [&](const float* base, int p, int which) -> const float* {
        if (MODE == 1) return base + head * head_stride + ((long long)utt_off[utt] + p) * kv_ld;
        if (p == step) return qrow0 + (which + 1) * D;
        return base + head * head_stride + ((long long)p * R + rbase + arow[p]) * kv_ld;
    };

    // ---------------- pass 1: scores
    for (int t0 = 0; t0 < n; t0 += KT) {
        __syncthreads();
        {                                            // 16 consecutive threads fetch one key row (256 B); all 16 row loads
            const float* src[16];                    // of a thread are in flight before the first one is consumed
            float4 v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int p = t0 + (tid >> 4) + 8 * j;
                src[j] = (p < n) ? src_ptr(kc, p, 0) + 4 * (tid & 15) : nullptr;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = src[j] ? *reinterpret_cast<const float4*>(src[j]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float* d = tile + ((tid >> 4) + 8 * j) * KSTR + 4 * (tid & 15);
                d[0] = v[j].x; d[1] = v[j].y; d[2] = v[j].z; d[3] = v[j].w;
            }
        }
        __syncthreads();
        const int p = t0 + tid;
        if (p < n) {
            float kk[DH];
#pragma unroll
            for (int i = 0; i < DH; ++i) kk[i] = tile[tid * KSTR + i];
            for (int h = 0; h < nh; ++h) {
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < DH; ++i) s = fmaf(qs[h * DH + i], kk[i], s);
                sc[h * smax + p] = s * 0.125f;
            }
        }
    }
    __syncthreads();
    // ---------------- softmax statistics per hyp (block-wide)
    float inv[MAXH];
    for (int h = 0; h < nh; ++h) {
        float mx = -INFINITY;
        for (int p = tid; p < n; p += 128) mx = fmaxf(mx, sc[h * smax + p]);
        mx = warp_max(mx);
        if (lane == 0) s_red[warp][h] = mx;
    }
    __syncthreads();
    for (int h = 0; h < nh; ++h) {
        const float mx = fmaxf(fmaxf(s_red[0][h], s_red[1][h]), fmaxf(s_red[2][h], s_red[3][h]));
        float sum = 0.f;
        for (int p = tid; p < n; p += 128) {
            const float e = expf(sc[h * smax + p] - mx);
            sc[h * smax + p] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        __syncthreads();                             // all reads of s_red[.][h] (max) are done before it is reused
        if (lane == 0) s_red[warp][h] = sum;
        __syncthreads();
        inv[h] = 1.f / (s_red[0][h] + s_red[1][h] + s_red[2][h] + s_red[3][h]);
        __syncthreads();
    }
    // ---------------- pass 2: weighted sum of V
    float acc[MAXH][2];
#pragma unroll
    for (int h = 0; h < MAXH; ++h) acc[h][0] = acc[h][1] = 0.f;
    for (int t0 = 0; t0 < n; t0 += KT) {
        __syncthreads();
        {
            const float* src[16];
            float4 v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int p = t0 + (tid >> 4) + 8 * j;
                src[j] = (p < n) ? src_ptr(vc, p, 1) + 4 * (tid & 15) : nullptr;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = src[j] ? *reinterpret_cast<const float4*>(src[j]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 16; ++j) *reinterpret_cast<float4*>(tile + ((tid >> 4) + 8 * j) * DH + 4 * (tid & 15)) = v[j];
        }
        __syncthreads();
        const int kend = min(32, n - t0 - warp * 32);
        for (int k = 0; k < kend; ++k) {
            const int key = warp * 32 + k;
            const float2 v = *reinterpret_cast<const float2*>(tile + key * DH + lane * 2);
#pragma unroll
            for (int h = 0; h < MAXH; ++h) {
                if (h < nh) {
                    const float w = sc[h * smax + t0 + key];
                    acc[h][0] = fmaf(w, v.x, acc[h][0]);
                    acc[h][1] = fmaf(w, v.y, acc[h][1]);
                }
            }
        }
    }
#pragma unroll
    for (int h = 0; h < MAXH; ++h) {
        if (h < nh) { s_o[warp][h][lane * 2] = acc[h][0]; s_o[warp][h][lane * 2 + 1] = acc[h][1]; }
    }
    __syncthreads();
    for (int i = tid; i < nh * DH; i += 128) {
        const int h = i / DH, d = i % DH;
        const float v = (s_o[0][h][d] + s_o[1][h][d] + s_o[2][h][d] + s_o[3][h][d]) * inv[h];
        const long long row = row0 + h;
        if (out) out[row * D + head * DH + d] = v;
        if (out_split) avsr_split3_store(out_split + row * 6 * D, D, head * DH + d, v);
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

// mode 0: self-attention step. q_in = qkv [R, 3072] (q | k | v of the current position), kc/vc = this layer's caches,
// element (pos, row, head, d) at head*head_stride + (pos*R + row)*kv_ld + d; anc [2][R][lmax].
// mode 1: cross-attention. q_in = q [R, 1024], kc/vc = this layer's cross K / V, element (frame, head, d) at
// head*head_stride + frame*kv_ld + d.  out (fp32 [R,1024]) and/or out_split (bf16x3 [R, 6*1024]).
extern "C" int avsr_dec_attn_step(int mode, const float* q_in, long long ldq, float* kc, float* vc, const unsigned char* anc, int lmax,
                                  const int* n_run, const int* utt_off, const int* utt_T, int beam, int R, const int* step,
                                  float* out, int max_keys, long long kv_ld, long long head_stride, void* out_split, cudaStream_t stream) {
    AVSR_REQUIRE(q_in && kc && vc && n_run && step && (out || out_split) && R > 0 && beam > 0 && max_keys > 0, "avsr_dec_attn_step: bad arguments");
    AVSR_REQUIRE(mode == 0 ? (anc != nullptr) : (utt_off && utt_T), "avsr_dec_attn_step: missing index arrays for mode %d", mode);
    AVSR_REQUIRE(beam <= MAXH && R % beam == 0, "avsr_dec_attn_step: beam %d unsupported (max %d)", beam, MAXH);
    const int smax = (max_keys + 3) & ~3;
    const int nh = mode == 1 ? beam : 1;
    const size_t smem = ((size_t)KT * KSTR + MAXH * DH + (size_t)nh * smax) * sizeof(float) + (mode == 0 ? (size_t)smax : 0);
    AVSR_REQUIRE(smem <= 160 * 1024, "avsr_dec_attn_step: %d keys x %d hyps exceed shared memory", max_keys, nh);
    static size_t configured[2] = {0, 0};
    if (smem > 48 * 1024 && smem > configured[mode]) {
        if (mode == 0) AVSR_CHECK_CUDA(cudaFuncSetAttribute(dec_attn_step_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        else AVSR_CHECK_CUDA(cudaFuncSetAttribute(dec_attn_step_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        configured[mode] = 160 * 1024;
    }
    if (mode == 0)
        dec_attn_step_kernel<0><<<dim3(R, HEADS), 128, smem, stream>>>(q_in, ldq, kc, vc, anc, lmax, n_run, utt_off, utt_T, beam, R, step, out,
                                                                      smax, kv_ld, head_stride, (__nv_bfloat16*)out_split);
    else
        dec_attn_step_kernel<1><<<dim3(R / beam, HEADS), 128, smem, stream>>>(q_in, ldq, kc, vc, anc, lmax, n_run, utt_off, utt_T, beam, R, step,
                                                                             out, smax, kv_ld, head_stride, (__nv_bfloat16*)out_split);
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
