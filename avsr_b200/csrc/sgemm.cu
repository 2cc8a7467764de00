// fp32 CUDA-core GEMM for the decode side of the hot path, where token-identical beam search needs fp32 weights and
// fp32 accumulation (hard part 1):
//     C[M,N] = epilogue( A[M,K] * W[N,K]^T )     A, W fp32 row-major (K contiguous).
// Used for: decoder step projections (reference src/nets/backend/transformer/decoder_layer.py:58-121,
// attention.py:38-106, positionwise_feed_forward.py:11-30, decoder.py:176-181), the once-per-utterance
// cross-attention K/V projection of the encoder memory (attention.py:50-52 hoisted out of the step loop) and the
// CTC head (src/nets/backend/ctc.py:163-170).
//
// Two tilings: a 128x128x16 register-tiled kernel for tall operands (M = frames) and a 96x64x16 split-K kernel for the
// skinny per-step operands (M = utterances x beam <= ~192 rows), whose deterministic partial sums are combined by
// avsr_splitk_epilogue (fixed summation order -> run-to-run identical tokens).
#include "common.cuh"
#include "splitk_epilogue.cuh"

namespace {

template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
sgemm_tn_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ W, long long ldw, float* __restrict__ part,
                int M, int N, int K, int k_per_split, const AvsrEpilogue ep, int direct) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int PAD = 4;
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Ws[2][BK][BN + PAD];
    constexpr int A_F4 = BM * BK / 4, W_F4 = BN * BK / 4;
    constexpr int A_PER = (A_F4 + NT - 1) / NT, W_PER = (W_F4 + NT - 1) / NT;

    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * k_per_split;
    const int kend = min(K, kbeg + k_per_split);
    const int num_kt = (kend - kbeg + BK - 1) / BK;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float4 ra[A_PER], rw[W_PER];
    auto gload = [&](int kt) {
        const int k0 = kbeg + kt * BK;
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            const int f = tid + i * NT;
            const int r = f / (BK / 4), kq = (f % (BK / 4)) * 4;
            ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (f < A_F4 && m0 + r < M && k0 + kq < kend) ra[i] = *reinterpret_cast<const float4*>(A + (long long)(m0 + r) * lda + k0 + kq);
        }
#pragma unroll
        for (int i = 0; i < W_PER; ++i) {
            const int f = tid + i * NT;
            const int r = f / (BK / 4), kq = (f % (BK / 4)) * 4;
            rw[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (f < W_F4 && n0 + r < N && k0 + kq < kend) rw[i] = *reinterpret_cast<const float4*>(W + (long long)(n0 + r) * ldw + k0 + kq);
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            const int f = tid + i * NT;
            if (f < A_F4) {
                const int r = f / (BK / 4), kq = (f % (BK / 4)) * 4;
                As[buf][kq + 0][r] = ra[i].x; As[buf][kq + 1][r] = ra[i].y; As[buf][kq + 2][r] = ra[i].z; As[buf][kq + 3][r] = ra[i].w;
            }
        }
#pragma unroll
        for (int i = 0; i < W_PER; ++i) {
            const int f = tid + i * NT;
            if (f < W_F4) {
                const int r = f / (BK / 4), kq = (f % (BK / 4)) * 4;
                Ws[buf][kq + 0][r] = rw[i].x; Ws[buf][kq + 1][r] = rw[i].y; Ws[buf][kq + 2][r] = rw[i].z; Ws[buf][kq + 3][r] = rw[i].w;
            }
        }
    };

    if (num_kt > 0) {
        gload(0);
        sstore(0);
    }
    __syncthreads();
    for (int kt = 0; kt < num_kt; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < num_kt) gload(kt + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], w[TN];
#pragma unroll
            for (int i = 0; i < TM; i += 2) {
                const float2 t = *reinterpret_cast<const float2*>(&As[buf][k][ty * TM + i]);
                a[i] = t.x; a[i + 1] = t.y;
            }
#pragma unroll
            for (int j = 0; j < TN; j += 4) {
                const float4 t = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * TN + j]);
                w[j] = t.x; w[j + 1] = t.y; w[j + 2] = t.z; w[j + 3] = t.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        if (kt + 1 < num_kt) sstore(buf ^ 1);
        __syncthreads();
    }

    if (!direct) {
        float* p = part + (long long)blockIdx.z * M * N;
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int row = m0 + ty * TM + i;
            if (row >= M) continue;
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const int col = n0 + tx * TN + j;
                if (col < N) p[(long long)row * N + col] = acc[i][j];
            }
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int row = m0 + ty * TM + i;
        if (row >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int col = n0 + tx * TN + j;
            if (col >= N) continue;
            float v = acc[i][j];
            if (ep.bias) v += (ep.bias_mode == 2) ? ep.bias[row] : ep.bias[col];
            if (!ep.act_after_residual) v = avsr_apply_act(v, ep.act, ep.act == AVSR_ACT_PRELU ? ep.prelu[col] : 0.f);
            if (ep.residual) {
                v += ep.res_dtype == 0 ? reinterpret_cast<const float*>(ep.residual)[(long long)row * ep.ldr + col]
                                       : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(ep.residual)[(long long)row * ep.ldr + col]);
            }
            if (ep.act_after_residual) v = avsr_apply_act(v, ep.act, ep.act == AVSR_ACT_PRELU ? ep.prelu[col] : 0.f);
            if (ep.out_f32) ep.out_f32[(long long)row * ep.ld_f32 + col] = v;
            if (ep.out_bf16) reinterpret_cast<__nv_bfloat16*>(ep.out_bf16)[(long long)row * ep.ld_bf16 + col] = __float2bfloat16_rn(v);
        }
    }
}

// One CTA per output row (see splitk_epilogue.cuh).
// ---- fast row epilogue for the decode step: exactly one group of four columns per thread (N = 4 * blockDim).  The chain of
// dependent launches makes the kernel's own critical path matter, so every load is requested as early as it can be: bias /
// gamma / beta (static) before griddepcontrol.wait; row_active, the residual and ALL partial sums right after it, the
// partial sums through in-order loads followed by an "issue barrier" (an empty volatile asm that owns the loaded registers)
// because ptxas otherwise sinks the loads between the adds and keeps only ~4 in flight (4-5 L2 round trips instead of one).
__device__ __forceinline__ float4 ldcg_f4_inorder(const float* p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
template <int MAXT, int BATCH>
__global__ void __launch_bounds__(MAXT)
splitk_epilogue_fast_kernel(const SplitKEpi e) {
    extern __shared__ float rowbuf[];
    __shared__ float red[32];
    const int row = blockIdx.x, c = threadIdx.x * 4, N = e.N, nsplit = e.nsplit;
    pdl_trigger();
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 b4 = e.bias ? __ldg(reinterpret_cast<const float4*>(e.bias + c)) : zero4;
    float4 g4 = zero4, be4 = zero4;
    if (e.ln_g) {
        g4 = __ldg(reinterpret_cast<const float4*>(e.ln_g + c));
        be4 = __ldg(reinterpret_cast<const float4*>(e.ln_b + c));
    }
    pdl_wait();
    const int active = e.row_active ? e.row_active[row] : 1;
    const float4 r4 = e.residual ? *reinterpret_cast<const float4*>(e.residual + (long long)row * e.ldr + c) : zero4;
    const long long zstride = (long long)e.M * N;
    const float* p = e.part + (long long)row * N + c;
    const int rt_zero = e.act >> 16;                  // 0 at run time (act is a small enum), opaque at compile time
    float4 v = zero4;
    for (int z = 0; z < nsplit; z += BATCH) {
        float4 t[BATCH];
#pragma unroll
        for (int u = 0; u < BATCH; ++u) t[u] = ldcg_f4_inorder(p + (long long)min(z + u, nsplit - 1) * zstride);
        // the first add depends on EVERY load of the batch (an OR over one word of each, masked to zero by a value the
        // compiler cannot know), so all loads are in flight before anything waits: one L2 round trip per batch
        int dep = 0;
#pragma unroll
        for (int u = 0; u < BATCH; ++u) dep |= __float_as_int(t[u].x);
        v.x += __int_as_float(dep & rt_zero);
#pragma unroll
        for (int u = 0; u < BATCH; ++u)
            if (z + u < nsplit) { v.x += t[u].x; v.y += t[u].y; v.z += t[u].z; v.w += t[u].w; }
    }
    if (!active) return;                              // uniform for the CTA
    v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
    if (e.act == AVSR_ACT_RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    else if (e.act == AVSR_ACT_GELU) { v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w); }
    v.x += r4.x; v.y += r4.y; v.z += r4.z; v.w += r4.w;
    if (e.out) *reinterpret_cast<float4*>(e.out + (long long)row * e.ldo + c) = v;
    if (e.ln_g == nullptr) {
        if (e.split_out) avsr_split3c_store4(e.split_out + (long long)row * 3 * N, N, c, v);
        return;
    }
    const float mean = block_sum((v.x + v.y) + (v.z + v.w), red) / (float)N;
    const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
    const float rstd = rsqrtf(block_sum((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3), red) / (float)N + e.ln_eps);
    const float4 y = make_float4(d0 * rstd * g4.x + be4.x, d1 * rstd * g4.y + be4.y, d2 * rstd * g4.z + be4.z, d3 * rstd * g4.w + be4.w);
    if (e.ln_out) *reinterpret_cast<float4*>(e.ln_out + (long long)row * e.ld_ln + c) = y;
    if (e.split_out) avsr_split3c_store4(e.split_out + (long long)row * 3 * N, N, c, y);
}

// pf / pf_bytes: optional span that a LATER kernel of the step streams (the cross-attention K/V of the layer, written before
// the chain of step kernels started).  Before waiting for its producer, CTA b asks the L2 to fetch its 1/gridDim share in
// 32 KB pieces (cp.async.bulk.prefetch.L2, fire and forget): the decode step is a chain of latency-bound launches during
// which HBM is mostly idle, so the span is L2-resident by the time the attention kernel asks for it.
template <int P>
__global__ void __launch_bounds__(768)
splitk_epilogue_kernel(const SplitKEpi e, const char* pf, long long pf_bytes) {
    extern __shared__ float rowbuf[];
    __shared__ float red[32];
    pdl_trigger();
    if (pf != nullptr) {
        constexpr long long PIECE = 32 * 1024;
        const long long share = ((pf_bytes + gridDim.x - 1) / gridDim.x + PIECE - 1) / PIECE * PIECE;
        const long long lo = (long long)blockIdx.x * share, hi = min(pf_bytes, lo + share);
        for (long long o = lo + (long long)threadIdx.x * PIECE; o < hi; o += (long long)blockDim.x * PIECE) {
            const unsigned n = (unsigned)min(PIECE, hi - o) & ~15u;
            if (n) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pf + o), "r"(n) : "memory");
        }
    }
    pdl_wait();
    avsr_splitk_epilogue_row<P>(e, blockIdx.x, rowbuf, red);
}

int g_sms = 0;
int sm_count() {
    if (g_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return g_sms;
}

}  // namespace

// Tall fp32 GEMM with the fused epilogue (no split-K).
extern "C" int avsr_sgemm(const float* A, long long lda, const float* W, long long ldw, int M, int N, int K,
                          const AvsrEpilogue* ep, cudaStream_t stream) {
    AVSR_REQUIRE(A && W && ep, "avsr_sgemm: null operand");
    AVSR_REQUIRE(M > 0 && N > 0 && K > 0 && (K & 3) == 0 && (lda & 3) == 0 && (ldw & 3) == 0,
                 "avsr_sgemm: bad shape/stride M=%d N=%d K=%d lda=%lld ldw=%lld", M, N, K, lda, ldw);
    dim3 grid(cdiv(N, 128), cdiv(M, 128), 1);
    sgemm_tn_kernel<128, 128, 16, 8, 8><<<grid, 256, 0, stream>>>(A, lda, W, ldw, nullptr, M, N, K, K, *ep, 1);
    AVSR_LAUNCH_CHECK();
    return AVSR_OK;
}

// Number of K splits avsr_sgemm_skinny will use (the caller sizes the partial-sum workspace: nsplit*M*N floats).
extern "C" int avsr_sgemm_skinny_splits(int M, int N, int K) {
    const int tiles = cdiv(N, 64) * cdiv(M, 96);
    int ns = cdiv(2 * sm_count(), tiles);
    const int max_ns = K / 64 > 0 ? K / 64 : 1;
    if (ns > max_ns) ns = max_ns;
    if (ns < 1) ns = 1;
    return ns;
}

// Skinny fp32 GEMM: writes raw partial sums part[nsplit][M][N]; combine with avsr_splitk_epilogue.
extern "C" int avsr_sgemm_skinny(const float* A, long long lda, const float* W, long long ldw, int M, int N, int K, float* part,
                                 int nsplit, cudaStream_t stream) {
    AVSR_REQUIRE(A && W && part, "avsr_sgemm_skinny: null operand");
    AVSR_REQUIRE(M > 0 && N > 0 && K > 0 && (K & 3) == 0 && (lda & 3) == 0 && (ldw & 3) == 0 && nsplit >= 1,
                 "avsr_sgemm_skinny: bad shape/stride M=%d N=%d K=%d", M, N, K);
    int kps = cdiv(K, nsplit);
    kps = ((kps + 15) / 16) * 16;
    AvsrEpilogue ep = {};
    dim3 grid(cdiv(N, 64), cdiv(M, 96), nsplit);
    sgemm_tn_kernel<96, 64, 16, 6, 4><<<grid, 256, 0, stream>>>(A, lda, W, ldw, part, M, N, K, kps, ep, 0);
    AVSR_LAUNCH_CHECK();
    return AVSR_OK;
}

extern "C" int avsr_splitk_epilogue(const float* part, int nsplit, int M, int N, const float* bias, int act, const float* residual,
                                    long long ldr, float* out, long long ldo, const float* ln_g, const float* ln_b, float ln_eps,
                                    float* ln_out, long long ld_ln, const int* row_active, void* split_out, cudaStream_t stream) {
    return avsr_splitk_epilogue_pf(part, nsplit, M, N, bias, act, residual, ldr, out, ldo, ln_g, ln_b, ln_eps, ln_out, ld_ln, row_active,
                                   split_out, nullptr, 0, stream);
}

extern "C" int avsr_splitk_epilogue_pf(const float* part, int nsplit, int M, int N, const float* bias, int act, const float* residual,
                                       long long ldr, float* out, long long ldo, const float* ln_g, const float* ln_b, float ln_eps,
                                       float* ln_out, long long ld_ln, const int* row_active, void* split_out, const void* l2_prefetch,
                                       long long l2_prefetch_bytes, cudaStream_t stream) {
    AVSR_REQUIRE(part && M > 0 && N > 0 && nsplit >= 1, "avsr_splitk_epilogue: bad arguments");
    AVSR_REQUIRE(!l2_prefetch || (((uintptr_t)l2_prefetch & 15) == 0 && l2_prefetch_bytes > 0), "avsr_splitk_epilogue: prefetch span must be 16-byte aligned");
    AVSR_REQUIRE(out || ln_out || split_out, "avsr_splitk_epilogue: no output");
    AVSR_REQUIRE(!ln_out || ln_g, "avsr_splitk_epilogue: ln_out needs gamma/beta");
    AVSR_REQUIRE(!ln_g || (ln_b && N * 4 <= 48 * 1024), "avsr_splitk_epilogue: LayerNorm needs gamma/beta and N <= 12288");
    const size_t smem = ln_g ? (size_t)N * 4 : 0;
    SplitKEpi e = {part, nsplit, M, N, bias, act, residual, ldr, out, ldo, ln_g, ln_b, ln_eps, ln_out, ld_ln, row_active, (__nv_bfloat16*)split_out};
    // one thread per 4 columns up to 768 threads.  (Splitting a column group's partial sums over 2 or 4 threads, P > 1, cuts
    // the loads per thread but measured slower on B200 at these sizes: 4.9 (P = 2) / 5.8 (P = 4) vs 4.0 us for N = 1024, 16 splits - the row-wide
    // reductions of the LayerNorm then run over 1024 mostly idle threads.)
    int threads = ((N + 3) / 4 + 31) / 32 * 32;
    threads = threads < 256 ? 256 : (threads > 768 ? 768 : threads);
    const bool vec = (N & 3) == 0 && (ldr & 3) == 0 && (ldo & 3) == 0 && (ld_ln & 3) == 0;
    if (vec && l2_prefetch == nullptr && N / 4 == threads && N == 4 * threads) {
        // one column group per thread: the fast kernel with every load requested up front
        if (threads <= 256 && nsplit <= 16)
            AVSR_CHECK_CUDA(avsr_launch_pdl(splitk_epilogue_fast_kernel<256, 16>, dim3(M), dim3(threads), 0, stream, e));
        else if (threads <= 256)
            AVSR_CHECK_CUDA(avsr_launch_pdl(splitk_epilogue_fast_kernel<256, 20>, dim3(M), dim3(threads), 0, stream, e));
        else
            AVSR_CHECK_CUDA(avsr_launch_pdl(splitk_epilogue_fast_kernel<768, 8>, dim3(M), dim3(threads), 0, stream, e));
        return AVSR_OK;
    }
    AVSR_CHECK_CUDA(avsr_launch_pdl(splitk_epilogue_kernel<1>, dim3(M), dim3(threads), smem, stream, e, (const char*)l2_prefetch, l2_prefetch_bytes));
    return AVSR_OK;
}
