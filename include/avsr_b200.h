/* libavsr_b200.so - C ABI of the B200-native AVSRCocktail inference hot path.
 *
 * The reference (quanpn90/avsr) is pure Python/PyTorch and has no FFI; each entry point below replaces a span of the
 * reference's Python that today expands into ATen/cuBLAS/cuDNN library calls (file:line relative to /root/reference).
 * Host code stays Python (avsr_b200/*.py) and binds these symbols with ctypes (avsr_b200/_lib.py); INTEGRATION.md shows
 * the binding a maintainer of the reference would add.
 *
 * Conventions: every function returns 0 on success or a negative AVSR_ERR_* code and records a message retrievable with
 * avsr_last_error().  All pointers are DEVICE pointers unless stated otherwise; the library never allocates, frees or
 * synchronises: buffers (including workspaces) belong to the caller (torch tensors), work is enqueued on `stream`.
 * bf16 buffers are passed as void*.  Leading dimensions are in elements.
 */
#ifndef AVSR_B200_H
#define AVSR_B200_H

#ifdef __CUDACC__
#include <cuda_runtime.h>
typedef cudaStream_t avsr_stream_t;
#else
typedef void* avsr_stream_t;
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define AVSR_OK 0
#define AVSR_ERR_CUDA -1
#define AVSR_ERR_ARG -2
#define AVSR_ERR_UNSUPPORTED -3

enum { AVSR_ACT_NONE = 0, AVSR_ACT_GELU = 1, AVSR_ACT_RELU = 2, AVSR_ACT_PRELU = 3 };

/* Fused GEMM epilogue:  v = acc (+ bias[col] | bias[row]); [v = act(v)]; v += residual[row,col];
 * [v = act(v) if act_after_residual]; rows with row_mask == 0 become 0; store bf16 and/or fp32. */
typedef struct AvsrEpilogue {
    const float* bias;      /* NULL = none */
    int bias_mode;          /* 1 = per output column, 2 = per output row */
    int act;                /* AVSR_ACT_* */
    const float* prelu;     /* [N] slopes when act == AVSR_ACT_PRELU */
    const void* residual;   /* NULL = none */
    int res_dtype;          /* 0 fp32, 1 bf16 */
    long long ldr;
    void* out_bf16;         /* NULL = skip */
    long long ld_bf16;
    float* out_f32;         /* NULL = skip */
    long long ld_f32;
    const int* row_mask;    /* optional [M] */
    int act_after_residual;
} AvsrEpilogue;

/* Device-resident state of the batched joint CTC/attention beam search (one row = utterance * beam + slot).
 * Replaces the Python Hypothesis / BatchHypothesis lists of src/nets/beam_search.py:13-27 and
 * src/nets/batch_beam_search.py:12-84. */
typedef struct AvsrBeamState {
    int B, beam, S, V, lmax, tmax, blank, eos, cap, no_end_detect;   /* no_end_detect: maxlenratio != 0 (beam_search.py:369) */
    const int* utt_T;       /* [B] frames per utterance (= maxlen, beam_search.py:349-350) */
    const int* step;        /* current position i */
    int* n_run;             /* [B] running hyps; 0 = utterance finished */
    int* row_active;        /* [R] */
    int* last_tok;          /* [R] */
    float* score;           /* [R] fused score */
    float* dec_sc;          /* [R] accumulated decoder log-prob */
    float* ctc_sc;          /* [R] accumulated CTC score */
    float* s_prev;          /* [R] log_psi of the prefix (CTC state) */
    int* rprev_idx;         /* [R] chain index of the inherited CTC forward variables */
    unsigned char* anc;     /* [2][R][lmax] self-attention cache ancestry */
    int* hist_tok;          /* [B][tmax][beam] chosen token of candidate j at step i */
    int* hist_prev;         /* [B][tmax][beam] its parent running index */
    int* run2j;             /* [B][tmax][beam] running index (after step i) -> candidate j */
    int* n_ended;           /* [B] */
    int* end_step;          /* [B][cap] */
    int* end_j;             /* [B][cap] */
    float* end_score;       /* [B][cap] */
    float* end_dec;         /* [B][cap] */
    float* end_ctc;         /* [B][cap] */
    int* end_len;           /* [B][cap] len(yseq) incl. sos and eos */
    float* best_len;        /* [B][tmax+4] best ended score per yseq length, -inf = none */
    float* best_all;        /* [B] */
    int* done;              /* [B] 0 = still searching, else 1 + the position at which the utterance stopped (end_detect fired or no running hyp left) */
    int* overflow;          /* [1] set if an ended list overflowed */
    double d_end;           /* end_detect threshold log(exp(-10)), e2e_asr_common.py:18 */
    const int* utt_maxlen;  /* [B] positions after which eos is appended; NULL = utt_T (maxlenratio == 0, beam_search.py:349-354) */
} AvsrBeamState;

const char* avsr_last_error(void);
int avsr_abi_version(void);

/* ---- dense encoder ops (tcgen05 tensor cores) --------------------------------------------------------------------
 * C[M,N] = epilogue(A[M,K] * B[N,K]^T), A/B bf16 row-major.  Replaces every nn.Linear / Conv (as im2col GEMM) of the
 * encoder: src/nets/backend/backbones/avhubert.py:187-198,486-502,747-768, resnet.py:30-164, and HF
 * Wav2Vec2Attention/FeedForward/PositionalConvEmbedding (modeling_wav2vec2.py:326-573). */
int avsr_gemm_bf16_tc(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K, const AvsrEpilogue* ep,
                      int bn_hint, avsr_stream_t stream);
/* ks x ks (1 or 3) / stride 1 or 2 / pad ks/2 convolution of an NHWC bf16 tensor as an IMPLICIT GEMM (im2col-mode TMA loads of
 * the activation tensor; the patch matrix is never written): in [F,H,W,C], Wt [Cout, ks*ks*C] with k = (ky*ks + kx)*C + c,
 * outputs [F*Ho*Wo, Cout] through the epilogue.  C % 64 == 0.  BasicBlock / downsample convolutions,
 * src/nets/backend/backbones/resnet.py:30-69. */
int avsr_conv2d_bf16_tc(const void* in, const void* Wt, long long F, int H, int W, int C, int Cout, int ks, int stride,
                        const AvsrEpilogue* ep, avsr_stream_t stream);
/* The same reading the input image rows / frames at the given pixel pitches (0 = dense), e.g. in place from the padded layout
 * below. */
int avsr_conv2d_bf16_tc_pitched(const void* in, const void* Wt, long long F, int H, int W, int C, int Cout, int ks, int stride,
                                long long in_row_pitch_px, long long in_frame_pitch_px, const AvsrEpilogue* ep, avsr_stream_t stream);
/* 3x3 / stride 1 / pad 1 convolution 64 -> 64 channels (ResNet layer1, src/nets/backend/backbones/resnet.py:30-69) on the PADDED
 * layout [F][H + 1][W + 2][64] bf16 whose pad cells are zero: the halo of a 128-pixel tile is one contiguous run of the array,
 * staged in shared memory once and read by all nine filter taps.  Output / residual of `ep` use the same layout (ld = 64); pad
 * cells of the output are left untouched.  Wt as for avsr_conv2d_bf16_tc.  Bit-identical to it on the valid pixels. */
int avsr_conv3x3_halo_bf16(const void* in, const void* Wt, long long F, int H, int W, const AvsrEpilogue* ep, avsr_stream_t stream);
/* Split-K form for skinny operands: part[z][M][N] fp32 raw partial sums (reduced by avsr_splitk_epilogue). With the
 * "bf16x3" operand layout (avsr_split3 / *_split outputs: [a1|a1|a2|a1|a2|a3] x [w1|w2|w1|w3|w2|w1]) this gives
 * fp32-accurate decoder projections on the tensor cores (src/nets/backend/transformer/decoder_layer.py:58-121). */
int avsr_gemm_bf16_tc_splitk(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K, float* part,
                             int splits, int bn_hint, avsr_stream_t stream);
/* Decoder-step projections, compact bf16x3 form: part[z][R][N] (fp32, z < avsr_gemm_x3_splits) = A3 * W3^T restricted to
 * the z-th K range, A3 = [a1|a2|a3] ([R, 3K] bf16), W3 = [w1|w2|w3] ([N, 3K] bf16), six MMAs per k step (fp32-level
 * accuracy, 6 bytes per weight streamed).  Launched with programmatic dependent launch: the weight tiles are requested
 * while the kernel producing A3 is still running.  K % 64 == 0. */
int avsr_gemm_x3_splits(int R, int N, int K);
int avsr_gemm_x3_splitk(const void* A3, long long lda, const void* W3, long long ldw, int R, int N, int K, float* part,
                        avsr_stream_t stream);
/* The same projection followed IN THE SAME LAUNCH (grid-wide barrier) by its row-wise epilogue, arguments as
 * avsr_splitk_epilogue; gbar = two zero-initialised uint32 (barrier state, reusable by every launch on the stream).
 * Needs tiles * splits <= number of SMs (one CTA per work item, all resident). */
int avsr_gemm_x3_fused(const void* A3, long long lda, const void* W3, long long ldw, int R, int N, int K, float* part,
                       const float* bias, int act, const float* residual, long long ldr, float* out, long long ldo,
                       const float* ln_g, const float* ln_b, float ln_eps, float* ln_out, long long ld_ln,
                       const int* row_active, void* split_out, unsigned* gbar, avsr_stream_t stream);
/* Chained form: the launch first finishes the rows of the PREVIOUS projection that produce its operand (arguments p_* as
 * avsr_splitk_epilogue; p_split_out must be A3's buffer), one row per CTA, then runs the projection; no row-epilogue launch
 * in between.  ready = two zero-initialised uint32 (caller-owned); consecutive chained launches on a stream alternate
 * parity 0, 1, 0, ...; part must not alias p_part; needs tiles * splits <= number of SMs. */
int avsr_gemm_x3_chain(const void* A3, long long lda, const void* W3, long long ldw, int R, int N, int K, float* part,
                       const float* p_part, int p_nsplit, int p_N, const float* p_bias, int p_act, const float* p_residual,
                       long long p_ldr, float* p_out, long long p_ldo, const float* p_ln_g, const float* p_ln_b, float p_ln_eps,
                       const int* row_active, void* p_split_out, unsigned* ready, int parity, avsr_stream_t stream);
/* Decoder-step projection, second generation (csrc/gemm_x3c.cu): y = act(a W^T + bias) + residual for R rows with the K splits
 * of an output tile in ONE thread-block cluster, reduced through distributed shared memory in split order (deterministic; no
 * partial sums in global memory, no separate row-epilogue launch).  Replaces one nn.Linear of DecoderLayer.forward plus the
 * glue around it (src/nets/backend/transformer/decoder_layer.py:58-121; LayerNorm eps 1e-12, layer_norm.py:12-33).
 *   operand:  A3 != NULL: compact bf16x3 rows [R, 3K] (pitch lda); else a = LayerNorm(x): x [R, K] fp32 (pitch ldx), stats_in
 *             [K/128][R][2] = (mean, M2) per 128-column tile of x as an earlier call wrote them via stats_out, ln_g / ln_b [K].
 *   W3:       compact bf16x3 weights [N, 3K] (pitch ldw), K % 64 == 0.
 *   outputs:  out [R, N] fp32 (pitch ldo) and / or split_out [R, 3N] compact bf16x3 (N % 4 == 0); stats_out [N/128][R][2] (N % 128 == 0).
 *             residual [R, N] (pitch ldr) may alias out.  act: AVSR_ACT_NONE or AVSR_ACT_RELU (applied before the residual).
 *   l2_prefetch / l2_prefetch_bytes: optional span (the weights of the NEXT projection of the chain, 16-byte aligned) that the
 *             kernel asks the L2 to fetch while it runs, so the next launch does not pay the HBM latency after its wait.
 * avsr_dec_proj_splits: the cluster size (= K splits) chosen for a shape on the current device. */
int avsr_dec_proj_splits(int R, int N, int K);
int avsr_dec_proj_force_splits(int splits);                /* dev knob: 0 = automatic */
int avsr_dec_proj_prefetch_self_kv(const float* kd, const float* vd, int lmax, int n_utt_heads, const int* step);   /* one-shot: the next launch fetches [0, *step) of a layer's dense self-attention caches into L2 */
int avsr_dec_proj_also_prefetch(const void* span, long long bytes);   /* one-shot second L2 fetch-ahead span of the next avsr_dec_proj* launch */
int avsr_dec_proj_set_sm_budget(int sms);                  /* SMs one launch is planned for (0 = all): concurrent decode chains */
int avsr_dec_proj_max_clusters(int cluster_size, int nb);   /* resident clusters of that size (operand tiles of nb rows) */
int avsr_dec_proj(const void* A3, long long lda, const float* x, long long ldx, const float* stats_in, const float* ln_g,
                  const float* ln_b, float ln_eps, const void* W3, long long ldw, int R, int N, int K, const float* bias, int act,
                  const float* residual, long long ldr, float* out, long long ldo, void* split_out, float* stats_out,
                  const void* l2_prefetch, long long l2_prefetch_bytes, avsr_stream_t stream);
/* LayerNorm FOLDED into the projection (what the decode position uses): y = act(LayerNorm(x) W^T + bias) + residual computed as
 * y = act(rstd * (x (gamma . W)^T - mean * fold_u) + fold_c) + residual with fold_u[n] = sum_k gamma[k] W[n][k] and
 * fold_c[n] = sum_k beta[k] W[n][k] + bias[n] (host, float64).  X3 = the RAW rows of x as compact bf16x3 [R, 3K] (written by the
 * projection that produced x through split_out), W3g = compact bf16x3 of gamma . W; mean / rstd of a row come from stats_in
 * [K/128][R][2] in the epilogue, so nothing is normalised on the critical path.  K % 128 == 0. */
int avsr_dec_proj_folded(const void* X3, long long lda, const float* stats_in, float ln_eps, const float* fold_u, const float* fold_c,
                         const void* W3g, long long ldw, int R, int N, int K, int act, const float* residual, long long ldr, float* out,
                         long long ldo, void* split_out, float* stats_out, const void* l2_prefetch, long long l2_prefetch_bytes,
                         avsr_stream_t stream);
/* fp32 [rows, K] -> bf16 [rows, 6K] in the bf16x3 activation layout. */
int avsr_split3(const float* in, long long ldi, void* out, long long rows, int K, avsr_stream_t stream);
/* softmax(q k^T) v per head over packed variable-length utterances (modeling_wav2vec2.py:438-549 via avhubert.py:751). */
int avsr_attention_varlen(const void* qk, const void* vt, long long ld_vt, void* out, long long F, const int* work_off,
                          const int* work_T, const int* work_q0, int n_work, int max_T, avsr_stream_t stream);
/* torch.nn.LayerNorm over the last dim (avhubert.py:495, 705, 734; HF encoder layers). */
int avsr_layernorm(const float* x, long long ldx, long long rows, int N, const float* gamma, const float* beta, float eps,
                   void* out_bf16, long long ld_bf16, float* out_f32, long long ld_f32, avsr_stream_t stream);
/* Conv3d(1->64, 5x7x7, s 1x2x2, p 2x3x3) patches of packed frames (resnet.py:132). */
int avsr_im2col_frontend(const float* video, const int* frame_t, const int* frame_T, int f0, int nf, void* out, avsr_stream_t stream);
/* Conv3d(1 -> 64, 5x7x7, stride 1x2x2, pad 2x3x3, no bias) + BatchNorm3d (eval, folded into w / bias) + PReLU(64) of frames
 * [f0, f0 + nf) of the packed video as an IMPLICIT GEMM on the tensor cores (csrc/frontend_conv.cu): the patch matrix of
 * avsr_im2col_frontend is never written.  src/nets/backend/backbones/resnet.py:132-135.  w = [64][320] bf16 with
 * k = (dt*7 + dy)*8 + dx (dx = 7 and k >= 280 are zero), bias / prelu [64]; out = [nf][44][44][64] bf16 (NHWC); temporal zero
 * padding stops at utterance boundaries (frame_t / frame_T as in avsr_im2col_frontend). */
int avsr_frontend_conv3d(const float* video, const int* frame_t, const int* frame_T, int f0, int nf, const void* w, const float* bias,
                         const float* prelu, void* out, avsr_stream_t stream);
/* Conv2d 3x3 / 1x1 patches, channels-last bf16 (resnet.py:10-19). */
int avsr_im2col2d(const void* in, void* out, long long F, int H, int W, int C, int ks, int stride, avsr_stream_t stream);
/* MaxPool3d (1,3,3)/(1,2,2)/(0,1,1) (resnet.py:136) and AdaptiveAvgPool2d(1) (resnet.py:76). */
int avsr_maxpool3x3s2(const void* in, void* out, long long F, int H, int W, int C, avsr_stream_t stream);
/* ... writing the Ho x Wo pooled pixels of every frame at the given output pixel pitches (0 = dense). */
int avsr_maxpool3x3s2_pitched(const void* in, void* out, long long F, int H, int W, int C, long long out_row_pitch_px,
                              long long out_frame_pitch_px, avsr_stream_t stream);
int avsr_avgpool(const void* in, void* out, long long F, int HW, int C, avsr_stream_t stream);
/* audio [B,104,Tpad] fp32 -> packed [F,104] bf16 (the transpose of avhubert.py:196). */
int avsr_audio_pack(const float* audio, void* out, const int* frame_b, const int* frame_t, long long F, int Cin, int Tpad,
                    avsr_stream_t stream);
/* patches of the grouped positional Conv1d (k=128, pad=64, g=16), zero outside the utterance. */
int avsr_posconv_im2col(const void* x, void* out, const int* frame_t, const int* frame_T, long long F, int g0, int ng,
                        avsr_stream_t stream);
/* Positional convolution x + GELU(Conv1d_{k=128, pad=64, groups=16}(x)[..., :-1] + bias) (HF Wav2Vec2PositionalConvEmbedding,
 * modeling_wav2vec2.py:326-379, called from backbones/avhubert.py:698-699) as an IMPLICIT banded GEMM, one launch over all 16
 * groups: tap j of an output tile reads frames [q0 + j - 64, +128) of the tile's utterance through that utterance's own tensor
 * map (zero outside the utterance).  avsr_posconv_encode_maps (host only) writes the B tensor maps (128 bytes each) for the
 * packed bf16 activations x [F, 1024] into maps_host; the caller uploads them (64-byte aligned) and passes them as utt_maps.
 * W [1024][8192] bf16, k = tap * 64 + channel-in-group; work_* = 128-frame work items (utterance index, first packed frame of
 * the utterance, its length, first frame of the item); ep = per-column bias [1024], GELU, residual / outputs by packed frame. */
int avsr_posconv_encode_maps(const void* x, const long long* utt_off, const int* utt_T, int B, void* maps_host);
int avsr_posconv_bf16_tc(const void* utt_maps, const void* W, int n_work, const int* work_utt, const int* work_off, const int* work_T,
                         const int* work_q0, const AvsrEpilogue* ep, avsr_stream_t stream);
int avsr_cast_bf16(const float* in, long long ldi, void* out, long long ldo, long long rows, int cols, avsr_stream_t stream);

/* ---- fp32 decode-side ops ----------------------------------------------------------------------------------------- */
/* CTC head / cross-attention K,V projection (src/nets/backend/ctc.py:163-170; transformer/attention.py:50-52). */
int avsr_sgemm(const float* A, long long lda, const float* W, long long ldw, int M, int N, int K, const AvsrEpilogue* ep,
               avsr_stream_t stream);
int avsr_sgemm_skinny_splits(int M, int N, int K);
int avsr_sgemm_skinny(const float* A, long long lda, const float* W, long long ldw, int M, int N, int K, float* part, int nsplit,
                      avsr_stream_t stream);
int avsr_splitk_epilogue(const float* part, int nsplit, int M, int N, const float* bias, int act, const float* residual,
                         long long ldr, float* out, long long ldo, const float* ln_g, const float* ln_b, float ln_eps,
                         float* ln_out, long long ld_ln, const int* row_active, void* split_out, avsr_stream_t stream);
/* The same, and before it waits for its producer the kernel asks the L2 to fetch [l2_prefetch, +l2_prefetch_bytes): a span a
 * later kernel of the decode step streams (the layer's cross-attention K/V; must not be written by the chain itself). */
int avsr_splitk_epilogue_pf(const float* part, int nsplit, int M, int N, const float* bias, int act, const float* residual,
                            long long ldr, float* out, long long ldo, const float* ln_g, const float* ln_b, float ln_eps,
                            float* ln_out, long long ld_ln, const int* row_active, void* split_out, const void* l2_prefetch,
                            long long l2_prefetch_bytes, avsr_stream_t stream);
int avsr_log_softmax_rows(float* x, long long ld, long long rows, int V, avsr_stream_t stream);
/* Decoder.forward_one_step pieces (src/nets/backend/transformer/decoder.py:153-183, decoder_layer.py:58-121). */
int avsr_dec_embed_ln(const float* emb, const float* pe, const int* last_tok, const int* n_run, int beam, int R, const int* step,
                      const float* gamma, const float* beta, float eps, float* x, float* a, void* a_split, avsr_stream_t stream);
/* The embedded position without the LayerNorm: fp32 rows, compact bf16x3 rows and (mean, M2) per 128-column tile
 * (stats [8][R][2]) for a projection that takes the LayerNorm folded (avsr_dec_proj_folded / avsr_dec_proj_dual). */
int avsr_dec_embed_raw(const float* emb, const float* pe, const int* last_tok, const int* n_run, int beam, int R, const int* step,
                       float* x, void* x_split, float* stats, avsr_stream_t stream);
/* One decode position of MultiHeadedAttention (transformer/attention.py:38-106, called from decoder_layer.py:82-107) with
 * cached K/V, streamed once (csrc/dec_attn.cu).  Keys are stored transposed in 32-byte groups, values row-major.
 * mode 0 = self-attention over the hypothesis' own history: query / current k / current v = columns [0,1024) / [1024,2048) /
 *   [2048,3072) of q_in; kc / vc = this layer's caches: with row = pos*beam + slot and nr = lmax*beam, key element
 *   (utt, head, row, d) at ((utt*16 + head)*8 + d/8)*nr*8 + row*8 + d%8, value element at ((utt*16 + head)*nr + row)*64 + d;
 *   anc [2][R][lmax] = slot that holds position pos of a row's history (double-buffered on step parity); the current k / v
 *   are appended at (pos = *step, slot = own slot).
 * mode 1 = source attention over the utterance's precomputed K/V (n_frames packed frames of all utterances): key element
 *   (head, frame, d) at (head*8 + d/8)*n_frames*8 + frame*8 + d%8, value element at (head*n_frames + frame)*64 + d.
 * nsplit > 0: q_in = split-K partial sums part[z][R][ldq] of the projection, summed here in a fixed order + q_bias.
 * One CTA per (utterance, head) walks all keys in tiles of 128.  kd / vd / conv_len: optional dense caches of the converged
 * history prefix (mode 0 only, see avsr_dec_cache_promote); pass NULL to read everything through the ancestry table. */
int avsr_dec_attn_step(int mode, const float* q_in, long long ldq, int nsplit, const float* q_bias, float* kc, float* vc,
                       const unsigned char* anc, int lmax, const int* n_run, const int* utt_off, const int* utt_T, int beam, int R,
                       const int* step, float* out, long long n_frames, void* out_split, const float* kd, const float* vd,
                       const int* conv_len, avsr_stream_t stream);
/* The same, and every CTA first asks the L2 to fetch its share of [l2_prefetch, +l2_prefetch_bytes): the weights of the
 * projection that follows the attention in the chain (must not be written by the chain). */
/* "Query merge": the source-attention query of a decoder layer, LayerNorm(x1) Wq^T + bq with x1 = x + att Wo^T + bo, is linear in
 * (x, att) up to the LayerNorm's row statistics: x1 (gamma . Wq)^T = x (gamma . Wq)^T + att ((gamma . Wq) Wo)^T + const.  The two
 * products ride along in the projections that have x and att as their operand anyway (avsr_dec_proj_dual: extra output columns),
 * and the source-attention kernel applies rstd (. - mean u) + c itself.  One projection launch less per layer. */
int avsr_dec_proj_dual(const void* A3, long long lda, const float* stats_in, float ln_eps, const float* fold_u, const float* bias,
                       const void* W3, long long ldw, int R, int N, int K, int n1, int act, const float* residual, long long ldr,
                       float* out, long long ldo, const float* residual2, long long ldr2, float* out2, long long ldo2,
                       float* stats_out, const void* l2_prefetch, long long l2_prefetch_bytes, avsr_stream_t stream);
/* One-shot: the next mode-1 avsr_dec_attn_step* launch finishes its query q = rstd (q_in - mean u) + c from the tile statistics
 * stats [8][R][2] of the row the query was projected from (u, c [1024] as weights.fold_layernorm builds them). */
int avsr_dec_attn_fold_query(const float* stats, const float* u, const float* c, float eps);
int avsr_dec_attn_step_pf(int mode, const float* q_in, long long ldq, int nsplit, const float* q_bias, float* kc, float* vc,
                          const unsigned char* anc, int lmax, const int* n_run, const int* utt_off, const int* utt_T, int beam, int R,
                          const int* step, float* out, long long n_frames, void* out_split, const float* kd, const float* vd,
                          const int* conv_len, const void* l2_prefetch, long long l2_prefetch_bytes, avsr_stream_t stream);
int avsr_dec_attn_debug(unsigned long long* buf);   /* dev aid: phase timestamps of one self-attention CTA (8 uint64, NULL = off) */
/* Dense copy of the CONVERGED history prefix (positions where all live hyps of an utterance share their ancestor): copies the
 * newly converged rows of all layers from the per-slot caches to kd / vd (key element (layer, utt, head, pos, d) at
 * layer*B*16*lmax*64 + ((utt*16 + head)*8 + d/8)*lmax*8 + pos*8 + d%8, value element at layer*B*16*lmax*64 +
 * ((utt*16 + head)*lmax + pos)*64 + d) and advances conv_len[(*step + 1) & 1][utt].  avsr_dec_attn_step (mode 0) then
 * streams consecutive rows for that prefix.  Call once per position before the layers run. */
int avsr_dec_cache_promote(const float* kc, const float* vc, float* kd, float* vd, int n_layers, const unsigned char* anc, int lmax,
                           const int* n_run, int beam, int R, const int* step, int* conv_len, avsr_stream_t stream);
/* [F, ncol] fp32 -> [ncol/64][F_capacity][64] (head-major K/V: every (utterance, head) reads one contiguous span; the
 * destination has room for F_capacity >= F frames per block so one allocation serves batches of different sizes).  k_transposed:
 * columns are [k(1024) | v(1024)] pairs and the K blocks are written [block][8][F][8] (transposed in 32-byte groups), the
 * key layout avsr_dec_attn_step mode 1 reads. */
int avsr_kv_head_major(const float* in, float* out, long long F, long long F_capacity, int ncol, int k_transposed,
                       avsr_stream_t stream);
int avsr_dec_logits_lsm_topk(const float* part, int nsplit, int R, int V, const float* bias, const int* n_run, int beam, float* logp,
                             int* part_ids, int S, avsr_stream_t stream);
/* CTCPrefixScoreTH.__call__ (src/nets/ctc_prefix_score.py:68-187): pre-beam and full-vocabulary modes.  logp = CTC
 * log-posteriors [sum(T), V] with row pitch ldp floats (the full-vocabulary kernel needs ldp % 4 == 0 and a 16-byte
 * aligned base: its rows are read with 16-byte loads). */
int avsr_ctc_prefix_prebeam(const float* logp, int V, int ldp, int blank, const int* utt_off, const int* utt_T, const int* n_run, int beam,
                            int R, int S, const int* last_tok, const int* part_ids, const int* rprev_idx, float* r_buf, int tmax,
                            const int* step, float* psi, float* rsum_last, avsr_stream_t stream);
/* avsr_dec_logits_lsm_topk + avsr_ctc_prefix_prebeam in ONE launch (one CTA per hypothesis row; the pre-beam candidates stay
 * in shared memory): the tail of a decode position, decoder.py:176-181 + batch_beam_search.py:229-235 + ctc_prefix_score.py:68-187. */
int avsr_dec_tail(const float* part, int nsplit, const float* bias, float* dec_logp, int* part_ids, const float* logp, int V, int ldp,
                  int blank, const int* utt_off, const int* utt_T, const int* n_run, int beam, int R, int S, const int* last_tok,
                  const int* rprev_idx, float* r_buf, int tmax, const int* step, float* psi, float* rsum_last, avsr_stream_t stream);
/* Dense [n][V] score matrix CTCPrefixScoreTH.__call__ returns in pre-beam mode (log_psi - s_prev; logzero outside the
 * candidates, eos = r_sum[T-1], blank = logzero; ctc_prefix_score.py:164-187) from avsr_ctc_prefix_prebeam's compact outputs.
 * Only the scorer plug-in API (src/nets/scorer_interface.py:162-186) needs it; the fused search never materialises it. */
int avsr_ctc_scores_dense(const float* psi, const float* rsum_last, const float* s_prev, const int* part_ids, int n, int S, int V,
                          int blank, int eos, float* scores, avsr_stream_t stream);
/* Full-vocabulary mode.  avsr_ctc_prefix_full_plan gives the work split (*ncg column groups x *tsplit time splits per
 * utterance); caller-owned scratch: part [B][tsplit][beam][V] fp32 (may be NULL when tsplit == 1), tickets [B][ncg] int32
 * zeroed once. */
int avsr_ctc_prefix_full_plan(int B, int V, int* ncg, int* tsplit);
int avsr_ctc_prefix_full(const float* logp, int V, int ldp, int blank, int eos, const int* utt_off, const int* utt_T, const int* n_run,
                         int beam, int B, int S, const int* last_tok, const int* rprev_idx, const float* r_buf, int tmax,
                         const int* step, const float* s_prev, float* scores, float* part, int* tickets, avsr_stream_t stream);
/* The same with the posteriors themselves given as well: probs = exp(logp) as avsr_ctc_exp_posteriors writes it (same pitch);
 * the kernel streams probs and its inner loop needs no exponential.  The search calls avsr_ctc_exp_posteriors once per batch
 * of utterances and this entry point at every position.  probs == NULL: identical to avsr_ctc_prefix_full.  Results are
 * bit-identical either way. */
int avsr_ctc_prefix_full_probs(const float* logp, const float* probs, int V, int ldp, int blank, int eos, const int* utt_off,
                               const int* utt_T, const int* n_run, int beam, int B, int S, const int* last_tok, const int* rprev_idx,
                               const float* r_buf, int tmax, const int* step, const float* s_prev, float* scores, float* part,
                               int* tickets, cudaStream_t stream);
int avsr_ctc_exp_posteriors(const float* logp, long long n, float* probs, cudaStream_t stream);
/* BatchBeamSearch.search fusion + batch_beam top-k + post_process + end_detect
 * (src/nets/batch_beam_search.py:86-110,222-349; src/nets/e2e_asr_common.py:18-48). */
int avsr_beam_fuse_topk_advance(const AvsrBeamState* st, const float* dec_logp, const int* part_ids, const float* psi,
                                const float* rsum_last, float w_dec, float w_ctc, avsr_stream_t stream);
/* Full-vocabulary form for the single-scorer search (ctc_weight = 1.0, avhubert_avsr_model.py:35): ctc_full [R][V] from
 * avsr_ctc_prefix_full replaces the pre-beam candidates; rc_last / rc_chain / rc_tok [R] receive, per new running hyp, the
 * inputs of the avsr_ctc_prefix_prebeam (S = 1) call that recomputes its forward variables.  The state must have S = 1. */
int avsr_beam_fuse_topk_advance_full(const AvsrBeamState* st, const float* dec_logp, const float* ctc_full, float w_dec, float w_ctc,
                                     int* rc_last, int* rc_chain, int* rc_tok, avsr_stream_t stream);
/* avsr_beam_fuse_topk_advance + avsr_beam_step_advance in one launch: the last CTA to finish sets *any_running and advances
 * *st->step.  ticket: one int32 in device memory, zero before the first call. */
int avsr_beam_fuse_topk_advance_step(const AvsrBeamState* st, const float* dec_logp, const int* part_ids, const float* psi,
                                     const float* rsum_last, float w_dec, float w_ctc, int* any_running, int* ticket,
                                     avsr_stream_t stream);
int avsr_beam_step_advance(int* step, const int* n_run, int B, int* any_running, avsr_stream_t stream);

/* ---- input pipeline: the data format on the input side of the path (src/dataset/avhubert_dataset.py) -------------- */
/* FBanksAndStack.forward (:86-116) over a batch, with cut_or_pad (:22-33) and collate_pad + permute (:277-312, :347) folded
 * in.  wave = fp32 samples of all utterances back to back; utterance b starts at wave_off[b], has wave_len[b] samples and
 * is cut or zero-padded to n_samples[b] (= 640 * video frames in DataCollator.__call__, :335) before
 * python_speech_features.logfbank (25 ms / 10 ms frames, 512-point FFT, 26 mel filters, pre-emphasis 0.97), zero rows up
 * to a multiple of 4 frames, 4-frame stacking and F.layer_norm over the 104 features.  out = [B][104][Tmax] fp32, zero for
 * rows >= avsr_fbank_rows(n_samples[b]); Tmax must be >= every utterance's row count. */
int avsr_fbank_stack_ln(const float* wave, const long long* wave_off, const int* wave_len, const int* n_samples, int B, int Tmax,
                        float* out, avsr_stream_t stream);
/* Host helpers (no GPU work): rows FBanksAndStack yields for n samples; the 28 FFT-bin edges of the mel filters. */
int avsr_fbank_rows(int n_samples);
int avsr_fbank_bins(int* bins28);
/* torchaudio.functional.add_noise as AddMultiSpk / AddNoise call it (:160-222): wave, noise, out [B][L] fp32, snr_db [B],
 * lengths [B] or NULL (energies over the first lengths[b] samples), energy = caller-owned scratch of 2*B doubles. */
int avsr_add_noise(const float* wave, const float* noise, const float* snr_db, const int* lengths, int B, long long L, float* out,
                   double* energy, avsr_stream_t stream);
/* VideoTransform("test") (:225-246): uint8 grey frames [sum(T)][H][W] (utterance b = frames frame_off[b] .. + utt_T[b]) ->
 * x / 255, CenterCrop(88), Normalize(0.421, 0.165) -> out [B][1][Tmax][88][88] fp32, zero frames behind each utterance. */
int avsr_video_u8_transform(const unsigned char* frames, const long long* frame_off, const int* utt_T, int B, int Tmax, int H, int W,
                            float* out, avsr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* AVSR_B200_H */
