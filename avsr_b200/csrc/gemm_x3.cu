// Skinny fp32-accurate projections of the decoder step on the tcgen05 tensor cores ("bf16x3", compact form).
//
//     part[z][r][n] = sum_{k in split z} act[r][k] * W[n][k]        act [R, K] fp32 (R = utterances x beam <= a few hundred rows)
//
// Reference: every nn.Linear of Decoder.forward_one_step (src/nets/backend/transformer/decoder.py:153-183,
// decoder_layer.py:58-121, attention.py:38-106, positionwise_feed_forward.py:11-30) evaluated in fp32.  Both operands are
// stored as three bf16 terms (x = x1 + x2 + x3 to ~2^-24): activations [R, 3K] = [a1 | a2 | a3], weights [N, 3K] =
// [w1 | w2 | w3].  Per 16-wide k step SIX MMAs accumulate the six largest cross terms
//     a1 w1 + a1 w2 + a2 w1 + a1 w3 + a2 w2 + a3 w1
// into one fp32 accumulator in TMEM, so a weight is streamed from HBM as 6 bytes (the earlier layout repeated the terms
// along K, [w1|w2|w1|w3|w2|w1], and streamed 12).  The step is bound by that weight stream: 93 M parameters per position.
//
// Orientation: the WEIGHT tile is the M side of the MMA (128 output features = 128 TMEM lanes), the activations are the N
// side (NB = 32..128 rows), so no tensor-core rows are padding and the epilogue's stores are coalesced (32 lanes = 32
// consecutive features of one activation row).  Split-K over the SMs; the partial sums are reduced in a fixed order by
// avsr_splitk_epilogue / the attention kernels.
//
// Fused tail (avsr_gemm_x3_fused): after its partial sums are written every CTA joins a grid-wide barrier and then finishes
// rows of the output (sum of the K splits in split order, bias, activation, residual, LayerNorm, compact bf16x3 form of
// the result = the operand of the next projection), which used to be a separate launch per projection (42 per decoded
// position).  The barrier cannot deadlock: the grid is at most one CTA per SM, and a kernel only lets its dependents
// start (griddepcontrol.launch_dependents) once all of its own CTAs are resident.
//
// (Tried and measured slower on B200: a separate 3-stage weight ring + 2-stage activation ring with its own producer warp, so
// that all <= 3 k blocks of a CTA are requested before the wait: 10.0 / 6.8 us per projection instead of 9.3 / 6.4, the
// position 812 instead of 783 us - the activation tiles then queue behind 144 KB of weight requests per SM.)
//
// One CTA per work item (m-tile, n-tile, k-split): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2-5 =
// epilogue.  The kernel is launched with programmatic dependent launch: the weight tiles of the first pipeline stages are
// requested BEFORE griddepcontrol.wait, i.e. while the kernel that produces the activations is still running.
#include <stdlib.h>
#include "common.cuh"
#include "splitk_epilogue.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int BM = 128;                   // output features per tile (TMEM lanes)
constexpr int BK = 64;
constexpr int W_TILE = BM * BK * 2;       // 16 KB
constexpr int A_TILE_MAX = 128 * BK * 2;  // NB <= 128 activation rows; a stage holds 3 weight tiles + 3 tiles of NB rows
constexpr int STAGE_MAX = 3 * W_TILE + 3 * A_TILE_MAX;
constexpr int STAGES = 2;
constexpr int ROWBUF_BYTES = 3072 * 4;    // LayerNorm row of the fused tail (N <= 3072)
constexpr int SMEM_MAX = STAGES * STAGE_MAX + 1024 + 256 + ROWBUF_BYTES;
// Shared memory actually requested: stages sized by the real NB and the LayerNorm row only when a row tail / prologue runs
// in the launch.  At NB = 96 that is 170 KB instead of 205 KB, which leaves room for one CTA of the neighbouring attention
// kernel on the same SM: the projection's prologue (and the attention's K/V prefetch) can then overlap the other kernel.
static inline int smem_bytes(int nb, bool rowbuf, int nstages) {
    return nstages * (3 * W_TILE + 3 * nb * BK * 2) + 1024 + 256 + (rowbuf ? ROWBUF_BYTES : 0);
}
constexpr int NUM_THREADS = 64 + 4 * 32;
constexpr int TMEM_COLS = 128;

// mbarrier wait that traps instead of spinning forever if a transaction count was ever wrong (seconds, not a hung GPU)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t i = 0; !tc::mbar_try_wait(bar, parity); ++i)
        if (i > (1u << 28)) __trap();
}

// Sense-reversing barrier over all CTAs of the grid (thread 0 of each CTA): bar[0] = arrival count, bar[1] = generation.
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned nblocks) {
    unsigned my_gen;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(my_gen) : "l"(bar + 1) : "memory");
    __threadfence();
    const unsigned old = atomicAdd(bar, 1u);
    if (old == nblocks - 1) {
        atomicExch(bar, 0u);
        __threadfence();
        atomicAdd(bar + 1, 1u);
    } else {
        unsigned g;
        for (uint32_t i = 0;; ++i) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(g) : "l"(bar + 1) : "memory");
            if (g != my_gen) break;
            if (i > (1u << 26)) __trap();
            __nanosleep(32);
        }
    }
    __threadfence();
}

// Spin (one thread) until the rows-ready counter of this launch reaches `rows`; traps instead of hanging.
__device__ __forceinline__ void wait_rows_ready(const unsigned* ctr, unsigned rows) {
    unsigned v;
    for (uint32_t i = 0;; ++i) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        if (v >= rows) break;
        if (i > (1u << 26)) __trap();
        __nanosleep(20);
    }
    asm volatile("fence.proxy.async;" ::: "memory");
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_x3_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmA, int R, int N, int K, int NB,
               float* __restrict__ part, int splits, int tiles_m, int tiles_n, int fuse, const SplitKEpi epi, unsigned* gbar,
               int pro_on, const SplitKEpi pro, unsigned* ready, int rq, int nstages) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int A_TILE = NB * BK * 2, STAGE_BYTES = 3 * W_TILE + 3 * A_TILE;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + nstages * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 1);
    float* rowbuf = reinterpret_cast<float*>(smem + nstages * STAGE_BYTES + 256);
    __shared__ float red[32];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_items = tiles_m * tiles_n * splits;
    const int nkb = K / BK;
    const uint32_t stage_tx = 3 * W_TILE + 3 * (uint32_t)NB * BK * 2;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
        tc::mbar_init(tfull, 1);
        tc::mbar_init(tempty, 4);
        tc::fence_barrier_init();
        tc::tma_prefetch_desc(&tmW);
        tc::tma_prefetch_desc(&tmA);
    }
    if (warp == 1) {
        tc::tmem_alloc(tmem_slot, TMEM_COLS);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();                                    // dependents may start their own prologue

    auto item_coords = [&](int item, int& m0, int& n0, int& kb0, int& kb1) {
        const int z = item / (tiles_m * tiles_n), t2 = item - z * tiles_m * tiles_n;
        m0 = (t2 % tiles_m) * BM;
        n0 = (t2 / tiles_m) * NB;
        kb0 = (int)((long long)z * nkb / splits);
        kb1 = (int)((long long)(z + 1) * nkb / splits);
    };

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            bool waited = false;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                int m0, n0, kb0, kb1;
                item_coords(item, m0, n0, kb0, kb1);
                int kb = kb0;
                if (!waited) {
                    // weights do not depend on the previous kernel: fill the pipeline with them first, then wait for the
                    // grid that produces the activations, then complete the same stages with the activation tiles
                    const int npre = min(nstages, kb1 - kb0);
                    for (int i = 0; i < npre; ++i) {
                        uint8_t* sw = smem + i * STAGE_BYTES;
                        tc::mbar_arrive_expect_tx(&full[i], stage_tx);
#pragma unroll
                        for (int j = 0; j < 3; ++j) tc::tma_load_2d(sw + j * W_TILE, &tmW, &full[i], (kb0 + i) * BK + j * K, m0);
                    }
                    pdl_wait();
                    waited = true;
                    if (pro_on) wait_rows_ready(ready + rq, (unsigned)pro.M);      // the operand rows are finished by this launch
                    for (int i = 0; i < npre; ++i) {
                        uint8_t* sa = smem + i * STAGE_BYTES + 3 * W_TILE;
#pragma unroll
                        for (int j = 0; j < 3; ++j) tc::tma_load_2d(sa + j * A_TILE, &tmA, &full[i], (kb0 + i) * BK + j * K, n0);
                    }
                    kb = kb0 + npre;
                    stage = npre % nstages;
                    phase = (npre == nstages) ? 1u : 0u;
                }
                for (; kb < kb1; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sw = smem + stage * STAGE_BYTES;
                    tc::mbar_arrive_expect_tx(&full[stage], stage_tx);
#pragma unroll
                    for (int j = 0; j < 3; ++j) tc::tma_load_2d(sw + j * W_TILE, &tmW, &full[stage], kb * BK + j * K, m0);
#pragma unroll
                    for (int j = 0; j < 3; ++j) tc::tma_load_2d(sw + 3 * W_TILE + j * A_TILE, &tmA, &full[stage], kb * BK + j * K, n0);
                    if (++stage == nstages) { stage = 0; phase ^= 1; }
                }
            }
            if (!waited) pdl_wait();
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = tc::umma_idesc_bf16(BM, (uint32_t)NB);
            int stage = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                int m0, n0, kb0, kb1;
                item_coords(item, m0, n0, kb0, kb1);
                mbar_wait(tempty, acc_phase ^ 1);
                tc::tc_fence_after();
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc::tc_fence_after();
                    const uint32_t sw = tc::smem_u32(smem + stage * STAGE_BYTES);
                    const uint32_t sa = sw + 3 * W_TILE;
                    const uint64_t w1 = tc::umma_desc_sw128(sw), w2 = tc::umma_desc_sw128(sw + W_TILE), w3 = tc::umma_desc_sw128(sw + 2 * W_TILE);
                    const uint64_t a1 = tc::umma_desc_sw128(sa), a2 = tc::umma_desc_sw128(sa + A_TILE), a3 = tc::umma_desc_sw128(sa + 2 * A_TILE);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint32_t first = (kb > kb0 || k != 0) ? 1u : 0u;
                        tc::umma_bf16(tmem_base, w3 + 2 * k, a1 + 2 * k, idesc, first);      // smallest terms first
                        tc::umma_bf16(tmem_base, w1 + 2 * k, a3 + 2 * k, idesc, 1u);
                        tc::umma_bf16(tmem_base, w2 + 2 * k, a2 + 2 * k, idesc, 1u);
                        tc::umma_bf16(tmem_base, w2 + 2 * k, a1 + 2 * k, idesc, 1u);
                        tc::umma_bf16(tmem_base, w1 + 2 * k, a2 + 2 * k, idesc, 1u);
                        tc::umma_bf16(tmem_base, w1 + 2 * k, a1 + 2 * k, idesc, 1u);
                    }
                    tc::umma_commit(&empty[stage]);
                    if (++stage == nstages) { stage = 0; phase ^= 1; }
                }
                tc::umma_commit(tfull);
                acc_phase ^= 1;
            }
        }
    } else {
        const int quad = warp & 3;
        uint32_t acc_phase = 0;
        pdl_wait();                                   // the partial-sum buffer may still be read by the previous consumer
        if (pro_on) {
            // ---- prologue: finish rows of the PREVIOUS projection (split-K sum, bias, activation, residual, LayerNorm, bf16x3
            //      split) - they are this projection's operand.  One row per CTA; a device counter tells every CTA's TMA
            //      producer when all rows are there.  The counter of the next chained launch is re-armed here.
            const int gt = threadIdx.x - 64;
            if (blockIdx.x == 0 && gt == 0) ready[rq ^ 1] = 0u;
            unsigned done = 0;
            for (int row = blockIdx.x; row < pro.M; row += gridDim.x, ++done) avsr_splitk_epilogue_row_group<128>(pro, row, rowbuf, red, gt, 1);
            if (done) {
                __threadfence();
                asm volatile("fence.proxy.async;" ::: "memory");      // generic-proxy stores -> visible to the TMA loads of other CTAs
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (gt == 0) atomicAdd(ready + rq, done);
            }
        }
        for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
            int m0, n0, kb0, kb1;
            item_coords(item, m0, n0, kb0, kb1);
            const int z = item / (tiles_m * tiles_n);
            mbar_wait(tfull, acc_phase);
            tc::tc_fence_after();
            const int f = m0 + quad * 32 + lane;      // output feature of this thread
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
            float* pz = part + (long long)z * R * N;
#pragma unroll 1
            for (int c = 0; c < NB / 32; ++c) {
                uint32_t r[32];
                tc::tmem_ld_32x32(taddr + c * 32, r);
                tc::tmem_ld_wait();
                if (f < N) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int row = n0 + c * 32 + j;
                        if (row < R) pz[(long long)row * N + f] = __uint_as_float(r[j]);
                    }
                }
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(tempty);
            acc_phase ^= 1;
        }
        __threadfence();                              // partial sums visible device-wide before this CTA joins the barrier
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, TMEM_COLS);
    }
    if (fuse) {
        if (threadIdx.x == 0) grid_barrier(gbar, gridDim.x);
        __syncthreads();
        for (int row = blockIdx.x; row < R; row += gridDim.x) avsr_splitk_epilogue_row(epi, row, rowbuf, red);
    }
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

int plan(int R, int N, int K, int* nb, int* tiles_m, int* tiles_n, int* splits) {
    const int r32 = ((R + 31) / 32) * 32;
    *nb = r32 < 128 ? r32 : 128;
    *tiles_m = cdiv(N, BM);
    *tiles_n = cdiv(R, *nb);
    const int nkb = K / BK;
    int s = sm_count() / (*tiles_m * *tiles_n);
    static int min_kb = -1;                           // dev knob: at least this many k blocks per CTA (AVSR_X3_MIN_KB)
    if (min_kb < 0) { const char* e = getenv("AVSR_X3_MIN_KB"); min_kb = e ? atoi(e) : 1; if (min_kb < 1) min_kb = 1; }
    if (s > nkb / min_kb) s = nkb / min_kb;
    if (s > nkb) s = nkb;
    if (s < 1) s = 1;
    *splits = s;
    return AVSR_OK;
}

}  // namespace

// Number of K splits avsr_gemm_x3_splitk uses for this shape (the caller sizes part[splits][R][N]).
extern "C" int avsr_gemm_x3_splits(int R, int N, int K) {
    if (R <= 0 || N <= 0 || K <= 0 || (K % BK) != 0) return AVSR_ERR_ARG;
    int nb, tm, tn, s;
    plan(R, N, K, &nb, &tm, &tn, &s);
    return s;
}

static int x3_launch(const void* A3, long long lda, const void* W3, long long ldw, int R, int N, int K, float* part, int fuse,
                     const SplitKEpi& epi, unsigned* gbar, cudaStream_t stream, int pro_on = 0, const SplitKEpi* pro = nullptr,
                     unsigned* ready = nullptr, int rq = 0) {
    AVSR_REQUIRE(A3 && W3 && part, "avsr_gemm_x3: null operand");
    AVSR_REQUIRE(R > 0 && N > 0 && K > 0 && (K % BK) == 0, "avsr_gemm_x3: bad shape R=%d N=%d K=%d (K must be a multiple of 64)", R, N, K);
    int nb, tiles_m, tiles_n, splits;
    plan(R, N, K, &nb, &tiles_m, &tiles_n, &splits);
    CUtensorMap tw, ta;
    int rc = tc::make_tmap_2d_bf16(&tw, W3, (uint64_t)N, (uint64_t)3 * K, (uint64_t)ldw, BM, BK);
    if (rc != AVSR_OK) return rc;
    rc = tc::make_tmap_2d_bf16(&ta, A3, (uint64_t)R, (uint64_t)3 * K, (uint64_t)lda, (uint32_t)nb, BK);
    if (rc != AVSR_OK) return rc;
    static bool configured = false;
    if (!configured) {
        AVSR_CHECK_CUDA(cudaFuncSetAttribute(gemm_x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
        configured = true;
    }
    const int items = tiles_m * tiles_n * splits;
    const int grid = items < sm_count() ? items : sm_count();
    AVSR_REQUIRE(!pro_on || items <= sm_count(), "avsr_gemm_x3_chain: %d work items exceed one CTA per SM (R=%d N=%d)", items, R, N);
    SplitKEpi none = {};
    // (a single stage for the projections whose CTAs have one k block each - 86 KB, so that three attention CTAs fit beside
    // one on an SM - measured slower: 339 vs 330 ms per pass; two CTAs of the projection itself then share SMs)
    static int one_stage = -1;                        // dev knob AVSR_X3_ONE_STAGE=1
    if (one_stage < 0) { const char* e = getenv("AVSR_X3_ONE_STAGE"); one_stage = (e && atoi(e) == 1) ? 1 : 0; }
    const int nstages = (one_stage && items <= sm_count() && cdiv(K / BK, splits) <= 1) ? 1 : STAGES;
    AVSR_CHECK_CUDA(avsr_launch_pdl(gemm_x3_kernel, dim3(grid), dim3(NUM_THREADS), smem_bytes(nb, fuse || pro_on, nstages), stream, tw, ta, R, N, K, nb, part, splits,
                                    tiles_m, tiles_n, fuse, epi, gbar, pro_on, pro_on ? *pro : none, ready, rq, nstages));
    return AVSR_OK;
}

// part[z][R][N] (fp32) = A3[R, z-th K range] * W3[N, same]^T with A3 = [a1|a2|a3] ([R, 3K] bf16, pitch lda) and
// W3 = [w1|w2|w3] ([N, 3K] bf16, pitch ldw); K % 64 == 0.
extern "C" int avsr_gemm_x3_splitk(const void* A3, long long lda, const void* W3, long long ldw, int R, int N, int K, float* part,
                                   cudaStream_t stream) {
    SplitKEpi e = {};
    return x3_launch(A3, lda, W3, ldw, R, N, K, part, 0, e, nullptr, stream);
}

// The projection and its row-wise epilogue (arguments as avsr_splitk_epilogue) in ONE launch; gbar = two zero-initialised
// uint32 owned by the caller (grid barrier state, reusable by every launch on the same stream).
extern "C" int avsr_gemm_x3_fused(const void* A3, long long lda, const void* W3, long long ldw, int R, int N, int K, float* part,
                                  const float* bias, int act, const float* residual, long long ldr, float* out, long long ldo,
                                  const float* ln_g, const float* ln_b, float ln_eps, float* ln_out, long long ld_ln,
                                  const int* row_active, void* split_out, unsigned* gbar, cudaStream_t stream) {
    AVSR_REQUIRE(gbar, "avsr_gemm_x3_fused: missing barrier state");
    AVSR_REQUIRE(out || ln_out || split_out, "avsr_gemm_x3_fused: no output");
    AVSR_REQUIRE(!ln_out || ln_g, "avsr_gemm_x3_fused: ln_out needs gamma/beta");
    AVSR_REQUIRE(!ln_g || (ln_b && N * 4 <= ROWBUF_BYTES), "avsr_gemm_x3_fused: LayerNorm needs gamma/beta and N <= 3072");
    int nb, tiles_m, tiles_n, splits;
    plan(R, N, K, &nb, &tiles_m, &tiles_n, &splits);
    AVSR_REQUIRE(tiles_m * tiles_n * splits <= sm_count(), "avsr_gemm_x3_fused: %d work items exceed one CTA per SM (R=%d N=%d)",
                 tiles_m * tiles_n * splits, R, N);
    SplitKEpi e = {part, splits, R, N, bias, act, residual, ldr, out, ldo, ln_g, ln_b, ln_eps, ln_out, ld_ln, row_active, (__nv_bfloat16*)split_out};
    return x3_launch(A3, lda, W3, ldw, R, N, K, part, 1, e, gbar, stream);
}

// Chained form: before the projection loads its operand A3, the SAME launch finishes the rows of the previous projection
// that produce it (arguments as avsr_splitk_epilogue: p_part [p_nsplit][R][p_N] partial sums -> bias, activation, residual,
// out, LayerNorm -> split_out, which must be A3's buffer), one row per CTA, and a device counter releases the TMA loads
// once all R rows are there.  That removes the row-epilogue launch between two projections.  ready = two zero-initialised
// uint32 owned by the caller; consecutive chained launches on a stream alternate parity = 0, 1, 0, ... (each launch re-arms
// the other counter).  part must not alias p_part.  Needs tiles * splits <= number of SMs.
extern "C" int avsr_gemm_x3_chain(const void* A3, long long lda, const void* W3, long long ldw, int R, int N, int K, float* part,
                                  const float* p_part, int p_nsplit, int p_N, const float* p_bias, int p_act, const float* p_residual,
                                  long long p_ldr, float* p_out, long long p_ldo, const float* p_ln_g, const float* p_ln_b, float p_ln_eps,
                                  const int* row_active, void* p_split_out, unsigned* ready, int parity, cudaStream_t stream) {
    AVSR_REQUIRE(p_part && p_nsplit >= 1 && p_N > 0 && (p_N & 3) == 0 && (p_ldr & 3) == 0 && (p_ldo & 3) == 0 && ready && (parity == 0 || parity == 1),
                 "avsr_gemm_x3_chain: bad prologue arguments");
    AVSR_REQUIRE(p_part != part, "avsr_gemm_x3_chain: the partial-sum buffers of consecutive projections must differ");
    AVSR_REQUIRE(p_out || p_split_out, "avsr_gemm_x3_chain: the prologue has no output");
    AVSR_REQUIRE(!p_ln_g || (p_ln_b && p_N * 4 <= ROWBUF_BYTES), "avsr_gemm_x3_chain: LayerNorm needs gamma/beta and N <= 3072");
    SplitKEpi e = {};
    SplitKEpi pro = {p_part, p_nsplit, R, p_N, p_bias, p_act, p_residual, p_ldr, p_out, p_ldo, p_ln_g, p_ln_b, p_ln_eps, nullptr, 0, row_active,
                     (__nv_bfloat16*)p_split_out};
    return x3_launch(A3, lda, W3, ldw, R, N, K, part, 0, e, nullptr, stream, 1, &pro, ready, parity);
}
