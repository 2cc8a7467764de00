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
    pdl_trigger();
    pdl_wait();
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
        avsr_split3c_store4(a_split + (long long)row * 3 * D, D, c, y);
    }
}

// Single-query attention of one decode position, split over key chunks (flash-decoding form).
// One CTA (128 threads) per (utterance, head, chunk of CK keys); all live hyps (<= MAXH) of the utterance are served by
// the same CTA so that a K / V row shared by several hyps is read from HBM once:
//   MODE 1 (cross-attention): keys = the utterance's T frames, identical for every hyp (2 * T * 256 B per (utt, head)).
//   MODE 0 (self-attention):  keys = each hyp's own history, gathered through the ancestry table (hyps of a beam mostly
//                             share their ancestors), plus the current token whose k, v come from qkv and are appended
//                             to the cache here.
// Pass 1: thread = key, the 256-byte K row goes straight to registers (16 float4 loads in flight per thread).
// Local softmax statistics (max, sum) per hyp.  Pass 2: half-warp = key, lane = 4 output dims, 8 keys in flight per lane.
// With more than one chunk the partial (max, sum, sum e*v) are written to scratch and the LAST CTA of the (utt, head)
// group (atomic ticket) merges them in chunk order, so the result does not depend on which CTA arrives last.
// (A variant that staged the rows with 256-byte bulk async copies into shared memory measured slower: 49 vs 41 us for the
// cross-attention of 32 x 375 frames; the kernel is bound by its chain of dependent round trips, not by load issue.)
constexpr int MAXH = 8;
constexpr int CK = 128;                  // keys per chunk

template <int MODE>
__global__ void __launch_bounds__(128)
dec_attn_step_kernel(const float* __restrict__ q_in, long long ldq, float* kc, float* vc, const unsigned char* __restrict__ anc,
                     int lmax, const int* __restrict__ n_run, const int* __restrict__ utt_off, const int* __restrict__ utt_T,
                     int beam, int R, const int* __restrict__ step_p, float* __restrict__ out, long long kv_ld,
                     long long head_stride, __nv_bfloat16* __restrict__ out_split, float* __restrict__ part_o,
                     float* __restrict__ part_ms, int* __restrict__ tickets) {
    __shared__ __align__(16) float qs[MAXH][DH];
    __shared__ float sc[MAXH][CK];
    __shared__ unsigned char aslot[MAXH][CK];
    __shared__ __align__(16) float s_o[4][MAXH][DH];
    __shared__ float s_red[4][MAXH];
    __shared__ float s_m[MAXH], s_s[MAXH];
    __shared__ int s_last;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int utt = blockIdx.x, head = blockIdx.y, chunk = blockIdx.z, nch = gridDim.z;
    const int row0 = utt * beam;
    const int p0 = chunk * CK;
    const int p = p0 + tid;
    const float* kbase = kc + head * head_stride;
    const float* vbase = vc + head * head_stride;
    pdl_trigger();
    // Cross-attention K / V, utt_T and utt_off were written before this chain of kernels started: the thread's K row is
    // requested BEFORE waiting for the predecessor kernel (which produces q), so its HBM latency overlaps that kernel.
    const int T_utt = (MODE == 1) ? utt_T[utt] : 0;
    const long long uoff = (MODE == 1) ? (long long)utt_off[utt] : 0;
    float kk[DH];
    if (MODE == 1 && p < T_utt) {
        const float4* src = reinterpret_cast<const float4*>(kbase + (uoff + p) * kv_ld);
#pragma unroll
        for (int j = 0; j < DH / 4; ++j) {
            const float4 v = __ldg(src + j);
            kk[4 * j] = v.x; kk[4 * j + 1] = v.y; kk[4 * j + 2] = v.z; kk[4 * j + 3] = v.w;
        }
    }
    pdl_wait();
    const int nh = n_run[utt];
    const int step = *step_p;
    if (nh == 0) return;
    const int n = (MODE == 1) ? T_utt : step + 1;
    const int nact = (n + CK - 1) / CK;
    if (chunk >= nact) return;
    const bool valid = p < n;

    for (int i = tid; i < nh * DH; i += 128) qs[i / DH][i % DH] = q_in[(long long)(row0 + i / DH) * ldq + head * DH + (i % DH)];
    if (MODE == 0) {
        if (step >= p0 && step < p0 + CK) {          // this chunk owns the current position: append k, v (slot = the row itself)
            for (int i = tid; i < nh * DH; i += 128) {
                const int h = i / DH, d = i % DH;
                const float* qr = q_in + (long long)(row0 + h) * ldq + head * DH + d;
                const long long o = head * head_stride + ((long long)step * R + row0 + h) * kv_ld + d;
                kc[o] = qr[D];
                vc[o] = qr[2 * D];
            }
        }
        for (int h = 0; h < nh; ++h) {
            unsigned char s = 0;
            if (p < step) s = anc[((long long)(step & 1) * R + row0 + h) * lmax + p];
            else if (p == step) s = (unsigned char)h;
            aslot[h][tid] = s;
        }
    }
    __syncthreads();                                 // qs / aslot ready, appended k / v visible to the whole CTA

    auto row_ptr = [&](const float* base, int pos, int slot) -> const float* {
        if (MODE == 1) return base + (uoff + pos) * kv_ld;
        return base + ((long long)pos * R + row0 + slot) * kv_ld;
    };

    // ---------------- pass 1: scores of this chunk's keys for every hyp
    {
        int cur = (MODE == 1) ? 0 : -1;
        for (int h = 0; h < nh; ++h) {
            const int slot = (MODE == 1) ? 0 : (int)aslot[h][tid];
            if (valid && slot != cur) {
                const float4* src = reinterpret_cast<const float4*>(row_ptr(kbase, p, slot));
#pragma unroll
                for (int j = 0; j < DH / 4; ++j) {
                    const float4 v = src[j];
                    kk[4 * j] = v.x; kk[4 * j + 1] = v.y; kk[4 * j + 2] = v.z; kk[4 * j + 3] = v.w;
                }
                cur = slot;
            }
            float s = 0.f;
            if (valid) {
#pragma unroll
                for (int i = 0; i < DH; ++i) s = fmaf(qs[h][i], kk[i], s);
            }
            sc[h][tid] = valid ? s * 0.125f : -INFINITY;
        }
    }
    // ---------------- local softmax statistics per hyp
    for (int h = 0; h < nh; ++h) {
        const float mx = warp_max(sc[h][tid]);
        if (lane == 0) s_red[warp][h] = mx;
    }
    __syncthreads();
    float e_own[MAXH];
#pragma unroll
    for (int h = 0; h < MAXH; ++h) {
        e_own[h] = 0.f;
        if (h < nh) {
            const float mx = fmaxf(fmaxf(s_red[0][h], s_red[1][h]), fmaxf(s_red[2][h], s_red[3][h]));
            e_own[h] = valid ? expf(sc[h][tid] - mx) : 0.f;
            if (tid == 0) s_m[h] = mx;
        }
    }
    __syncthreads();                                 // every read of s_red (max) done before it is reused for the sums
#pragma unroll
    for (int h = 0; h < MAXH; ++h) {
        if (h < nh) {
            sc[h][tid] = e_own[h];
            const float sm = warp_sum(e_own[h]);
            if (lane == 0) s_red[warp][h] = sm;
        }
    }
    __syncthreads();
    if (tid < nh) s_s[tid] = s_red[0][tid] + s_red[1][tid] + s_red[2][tid] + s_red[3][tid];

    // ---------------- pass 2: sum_p e[p] * V[p] ; half-warp = key, lane = 4 dims
    float acc[MAXH][4];
#pragma unroll
    for (int h = 0; h < MAXH; ++h) acc[h][0] = acc[h][1] = acc[h][2] = acc[h][3] = 0.f;
    const int half = lane >> 4, l4 = (lane & 15) * 4;
#pragma unroll
    for (int i0 = 0; i0 < 16; i0 += 8) {
        float4 v0[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int kl = warp * 32 + half + 2 * (i0 + i);
            const int pos = p0 + kl;
            v0[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pos < n) v0[i] = *reinterpret_cast<const float4*>(row_ptr(vbase, pos, (MODE == 1) ? 0 : (int)aslot[0][kl]) + l4);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int kl = warp * 32 + half + 2 * (i0 + i);
            const int pos = p0 + kl;
            if (pos < n) {
                const int slot0 = (MODE == 1) ? 0 : (int)aslot[0][kl];
#pragma unroll
                for (int h = 0; h < MAXH; ++h) {
                    if (h < nh) {
                        float4 v = v0[i];
                        if (MODE == 0 && h > 0) {
                            const int slot = (int)aslot[h][kl];
                            if (slot != slot0) v = *reinterpret_cast<const float4*>(row_ptr(vbase, pos, slot) + l4);
                        }
                        const float w = sc[h][kl];
                        acc[h][0] = fmaf(w, v.x, acc[h][0]); acc[h][1] = fmaf(w, v.y, acc[h][1]);
                        acc[h][2] = fmaf(w, v.z, acc[h][2]); acc[h][3] = fmaf(w, v.w, acc[h][3]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int h = 0; h < MAXH; ++h) {
        if (h < nh) {
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[h][j] += __shfl_xor_sync(0xffffffffu, acc[h][j], 16);
            if (half == 0) *reinterpret_cast<float4*>(&s_o[warp][h][l4]) = make_float4(acc[h][0], acc[h][1], acc[h][2], acc[h][3]);
        }
    }
    __syncthreads();

    auto store_out = [&](int h, int d, float v) {
        const long long row = row0 + h;
        if (out) out[row * D + head * DH + d] = v;
        if (out_split) avsr_split3c_store(out_split + row * 3 * D, D, head * DH + d, v);
    };
    if (nact == 1) {
        for (int i = tid; i < nh * DH; i += 128) {
            const int h = i / DH, d = i % DH;
            store_out(h, d, (s_o[0][h][d] + s_o[1][h][d] + s_o[2][h][d] + s_o[3][h][d]) / s_s[h]);
        }
        return;
    }
    // ---------------- several chunks: publish the partial, the last CTA of the group merges all of them
    const long long grp = (long long)utt * gridDim.y + head;
    float* po = part_o + ((grp * nch + chunk) * beam) * DH;
    float* pms = part_ms + ((grp * nch + chunk) * beam) * 2;
    for (int i = tid; i < nh * DH; i += 128) {
        const int h = i / DH, d = i % DH;
        po[i] = s_o[0][h][d] + s_o[1][h][d] + s_o[2][h][d] + s_o[3][h][d];
    }
    if (tid < nh) { pms[2 * tid] = s_m[tid]; pms[2 * tid + 1] = s_s[tid]; }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const int t = atomicAdd(&tickets[grp], 1);
        s_last = (t == nact - 1) ? 1 : 0;
        if (s_last) tickets[grp] = 0;                // re-armed for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int i = tid; i < nh * DH; i += 128) {
        const int h = i / DH, d = i % DH;
        float M = -INFINITY;
        for (int c = 0; c < nact; ++c) M = fmaxf(M, __ldcg(part_ms + ((grp * nch + c) * beam + h) * 2));
        float S = 0.f, o = 0.f;
        for (int c = 0; c < nact; ++c) {
            const float* q = part_ms + ((grp * nch + c) * beam + h) * 2;
            const float f = expf(__ldcg(q) - M);
            S = fmaf(__ldcg(q + 1), f, S);
            o = fmaf(__ldcg(part_o + ((grp * nch + c) * beam) * DH + i), f, o);
        }
        store_out(h, d, o / S);
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
    pdl_trigger();
    pdl_wait();
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
    AVSR_CHECK_CUDA(avsr_launch_pdl(dec_embed_ln_kernel, dim3(R), dim3(256), 0, stream, emb, pe, last_tok, n_run, beam, step, gamma, beta, eps, x,
                                    a, (__nv_bfloat16*)a_split));
    return AVSR_OK;
}

// mode 0: self-attention step. q_in = qkv [R, 3072] (q | k | v of the current position), kc/vc = this layer's caches,
// element (pos, row, head, d) at head*head_stride + (pos*R + row)*kv_ld + d; anc [2][R][lmax].
// mode 1: cross-attention. q_in = q [R, 1024], kc/vc = this layer's cross K / V, element (frame, head, d) at
// head*head_stride + frame*kv_ld + d.  out (fp32 [R,1024]) and/or out_split (bf16x3 [R, 6*1024]).
// Scratch for the split over key chunks (nch = avsr_dec_attn_chunks(max_keys)): part_o [R/beam][16][nch][beam][64],
// part_ms [R/beam][16][nch][beam][2] fp32, tickets [R/beam][16] int32 zeroed once by the caller (the kernel re-arms them).
extern "C" int avsr_dec_attn_chunks(int max_keys) { return (max_keys + CK - 1) / CK; }

extern "C" int avsr_dec_attn_step(int mode, const float* q_in, long long ldq, float* kc, float* vc, const unsigned char* anc, int lmax,
                                  const int* n_run, const int* utt_off, const int* utt_T, int beam, int R, const int* step,
                                  float* out, int max_keys, long long kv_ld, long long head_stride, void* out_split, float* part_o,
                                  float* part_ms, int* tickets, cudaStream_t stream) {
    AVSR_REQUIRE(q_in && kc && vc && n_run && step && (out || out_split) && R > 0 && beam > 0 && max_keys > 0, "avsr_dec_attn_step: bad arguments");
    AVSR_REQUIRE(mode == 0 ? (anc != nullptr) : (utt_off && utt_T), "avsr_dec_attn_step: missing index arrays for mode %d", mode);
    AVSR_REQUIRE(beam <= MAXH && R % beam == 0, "avsr_dec_attn_step: beam %d unsupported (max %d)", beam, MAXH);
    AVSR_REQUIRE((kv_ld & 3) == 0 && (head_stride & 3) == 0, "avsr_dec_attn_step: K/V rows must be 16-byte aligned");
    const int nch = (max_keys + CK - 1) / CK;
    AVSR_REQUIRE(nch == 1 || (part_o && part_ms && tickets), "avsr_dec_attn_step: %d keys need the chunk scratch buffers", max_keys);
    AVSR_REQUIRE(nch <= 65535, "avsr_dec_attn_step: too many keys");
    const dim3 grid(R / beam, HEADS, nch);
    if (mode == 0)
        AVSR_CHECK_CUDA(avsr_launch_pdl(dec_attn_step_kernel<0>, grid, dim3(128), 0, stream, q_in, ldq, kc, vc, anc, lmax, n_run, utt_off, utt_T,
                                        beam, R, step, out, kv_ld, head_stride, (__nv_bfloat16*)out_split, part_o, part_ms, tickets));
    else
        AVSR_CHECK_CUDA(avsr_launch_pdl(dec_attn_step_kernel<1>, grid, dim3(128), 0, stream, q_in, ldq, kc, vc, anc, lmax, n_run, utt_off, utt_T,
                                        beam, R, step, out, kv_ld, head_stride, (__nv_bfloat16*)out_split, part_o, part_ms, tickets));
    return AVSR_OK;
}

extern "C" int avsr_dec_logits_lsm_topk(const float* part, int nsplit, int R, int V, const float* bias, const int* n_run, int beam,
                                        float* logp, int* part_ids, int S, cudaStream_t stream) {
    AVSR_REQUIRE(part && bias && n_run && logp && part_ids && R > 0 && V > 0 && S > 0 && S <= V, "avsr_dec_logits_lsm_topk: bad arguments");
    AVSR_REQUIRE((size_t)V * 4 <= 48 * 1024, "avsr_dec_logits_lsm_topk: vocabulary %d too large for the shared-memory row", V);
    AVSR_CHECK_CUDA(avsr_launch_pdl(dec_logits_lsm_topk_kernel, dim3(R), dim3(256), (size_t)V * 4, stream, part, nsplit, R, V, bias, n_run, beam,
                                    logp, part_ids, S));
    return AVSR_OK;
}

extern "C" int avsr_log_softmax_rows(float* x, long long ld, long long rows, int V, cudaStream_t stream) {
    AVSR_REQUIRE(x && rows > 0 && V > 0, "avsr_log_softmax_rows: bad arguments");
    log_softmax_rows_kernel<<<(unsigned)rows, 256, 0, stream>>>(x, ld, V);
    AVSR_LAUNCH_CHECK();
    return AVSR_OK;
}
