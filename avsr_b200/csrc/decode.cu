// Per-step kernels of the transformer decoder scorer (fp32), KV-cache form of the reference's output-cache step:
// src/nets/backend/transformer/decoder.py:153-227 (forward_one_step / batch_score), decoder_layer.py:58-121,
// attention.py:38-106, embedding.py:78-87, layer_norm.py:12-33 (eps 1e-12).  SURVEY.md App. A restates the math.
//
// Rows: R = n_utt * beam, row = utt * beam + slot; slot < n_run[utt] is live.  All hyps of all utterances are at the same
// position `*step`.  Self-attention K/V of position p are stored at [layer][p][row][1024] by the row
// that computed them; a hyp finds its history through anc[row][p] = slot (within its utterance) that holds position p,
// so beam reordering never copies the cache (reference copies per-hyp caches in batch_beam_search.py:250-284).
#include "common.cuh"
#include "dec_tail.cuh"

namespace {

constexpr int D = 1024;

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

// Output-layer log_softmax + pre-beam top-S, one CTA per row (device code in dec_tail.cuh).
template <int ITER>
__global__ void __launch_bounds__(LSM_THREADS)
dec_logits_lsm_topk_kernel(const float* __restrict__ part, int nsplit, int R, int V, const float* __restrict__ bias,
                           const int* __restrict__ n_run, int beam, float* __restrict__ logp, int* __restrict__ part_ids, int S) {
    __shared__ LsmSmem sm;
    const int row = blockIdx.x;
    pdl_trigger();
    float v[ITER];
    lsm_load_bias<ITER>(v, bias, V);
    pdl_wait();
    if ((row % beam) >= n_run[row / beam]) return;
    lsm_topk_row<ITER>(v, sm, part, nsplit, R, V, row, logp, part_ids + row * S, nullptr, S);
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

// Embedding of the position WITHOUT the first LayerNorm (the q|k|v projection of layer 0 takes it folded, like every other
// layer): x = emb[tok] * sqrt(d) + pe[step] (decoder.py:153-175, embedding.py:78-87) as fp32, as compact bf16x3 and the row's
// (mean, M2) per 128-column tile in the layout avsr_dec_proj consumes (stats [8][R][2]).  Warp w owns columns [128 w, +128).
__global__ void __launch_bounds__(256)
dec_embed_raw_kernel(const float* __restrict__ emb, const float* __restrict__ pe, const int* __restrict__ last_tok,
                     const int* __restrict__ n_run, int beam, int R, const int* __restrict__ step_p, float* __restrict__ x,
                     __nv_bfloat16* __restrict__ x_split, float* __restrict__ stats) {
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
    avsr_split3c_store4(x_split + (long long)row * 3 * D, D, c, v);
    const float mean = warp_sum((v.x + v.y) + (v.z + v.w)) * (1.f / 128.f);
    const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
    const float m2 = warp_sum((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3));
    if ((threadIdx.x & 31) == 0) *reinterpret_cast<float2*>(stats + ((long long)(threadIdx.x >> 5) * R + row) * 2) = make_float2(mean, m2);
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

// x [R, 1024] fp32, x_split [R, 3 * 1024] compact bf16x3, stats [8][R][2]: the embedded position, not normalised.
extern "C" int avsr_dec_embed_raw(const float* emb, const float* pe, const int* last_tok, const int* n_run, int beam, int R,
                                  const int* step, float* x, void* x_split, float* stats, cudaStream_t stream) {
    AVSR_REQUIRE(emb && pe && last_tok && n_run && step && x && x_split && stats && R > 0 && beam > 0, "avsr_dec_embed_raw: bad arguments");
    AVSR_CHECK_CUDA(avsr_launch_pdl(dec_embed_raw_kernel, dim3(R), dim3(256), 0, stream, emb, pe, last_tok, n_run, beam, R, step, x,
                                    (__nv_bfloat16*)x_split, stats));
    return AVSR_OK;
}

extern "C" int avsr_dec_logits_lsm_topk(const float* part, int nsplit, int R, int V, const float* bias, const int* n_run, int beam,
                                        float* logp, int* part_ids, int S, cudaStream_t stream) {
    AVSR_REQUIRE(part && bias && n_run && logp && part_ids && R > 0 && V > 0 && S > 0 && S <= V, "avsr_dec_logits_lsm_topk: bad arguments");
    AVSR_REQUIRE(V <= 16 * LSM_THREADS, "avsr_dec_logits_lsm_topk: vocabulary %d too large (max %d)", V, 16 * LSM_THREADS);
    const int iter = (V + LSM_THREADS - 1) / LSM_THREADS;
#define AVSR_LSM(IT)                                                                                                              \
    AVSR_CHECK_CUDA(avsr_launch_pdl(dec_logits_lsm_topk_kernel<IT>, dim3(R), dim3(LSM_THREADS), 0, stream, part, nsplit, R, V, bias, n_run, \
                                    beam, logp, part_ids, S))
    if (iter <= 4) AVSR_LSM(4);
    else if (iter <= 10) AVSR_LSM(10);
    else AVSR_LSM(16);
#undef AVSR_LSM
    return AVSR_OK;
}

extern "C" int avsr_log_softmax_rows(float* x, long long ld, long long rows, int V, cudaStream_t stream) {
    AVSR_REQUIRE(x && rows > 0 && V > 0, "avsr_log_softmax_rows: bad arguments");
    log_softmax_rows_kernel<<<(unsigned)rows, 256, 0, stream>>>(x, ld, V);
    AVSR_LAUNCH_CHECK();
    return AVSR_OK;
}
