// Decoder-step projections, second generation: split-K inside a THREAD-BLOCK CLUSTER, reduced through distributed shared
// memory, with the row-wise glue of the step folded into the two ends of the projection.
//
//     y[r][n] = act( sum_k a[r][k] * W[n][k] + bias[n] ) + residual[r][n]          R = utterances x beam <= a few hundred rows
//
// Reference: every nn.Linear of Decoder.forward_one_step and the LayerNorm / ReLU / residual glue between them
// (src/nets/backend/transformer/decoder.py:153-183, decoder_layer.py:58-121, attention.py:38-106,
// positionwise_feed_forward.py:11-30, layer_norm.py:12-33), evaluated with fp32-level accuracy: both operands as three bf16
// terms, six tcgen05 MMAs per 16-wide k step (csrc/gemm_x3.cu explains the arithmetic).
//
// What changed against gemm_x3.cu + splitk_epilogue (round 1: 37 projections + 24 row-epilogue launches per position, the K
// splits round-tripped through L2 as fp32 partial sums):
//   * grid = (m tiles of 128 output features) x (K splits) x (row tiles); the K splits of one tile form ONE CLUSTER
//     (cluster dims (1, splits, 1), up to 16 CTAs).  Every CTA leaves its fp32 accumulator tile in its own shared memory,
//     the cluster synchronises (barrier.cluster, hardware), and CTA c then sums rows [c * rows/splits, ...) of all peers in
//     split order through DSMEM (ld.shared::cluster): deterministic, no global partial sums, no second launch.
//   * the finishing CTA adds bias, activation and the residual, writes the fp32 result, and - for the results that feed a
//     LayerNorm - the per-row (mean, M2) of its 128 features (stats_out[tile][row]).  The LayerNorm itself is applied by
//     the NEXT projection while it stages its activation operand: its epilogue warps (idle until the MMAs finish) read the
//     fp32 rows of their k blocks, merge the eight tile statistics of a row in a fixed order (Chan), normalise, split into
//     three bf16 terms and store them in the 128-byte-swizzled layout the UMMA descriptors expect.  No counter handshake:
//     the dependency is the programmatic-dependent-launch boundary that exists anyway.
//   * results that feed another projection directly (ReLU(w_1 x)) are written as compact bf16x3 rows, which that projection
//     loads with TMA as before; results that feed an attention kernel (q | k | v) are written finished, bias included.
// A decode position is 54 launches instead of 78, and a projection no longer pays two L2 round trips of its partial sums.
//
// One CTA per work item, 192 threads: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2-5 = operand staging
// (LayerNorm mode) and epilogue.  The weight tiles of the first stages are requested BEFORE griddepcontrol.wait.
#include <stdlib.h>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int BM = 128;                   // output features per tile (TMEM lanes)
constexpr int BK = 64;
constexpr int W_TILE = BM * BK * 2;       // 16 KB
constexpr int STAGES = 2;
constexpr int NUM_THREADS = 64 + 4 * 32;
constexpr int TMEM_COLS = 128;
constexpr int MAX_NB = 128;
constexpr int MAX_CLUSTER = 16;

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t i = 0; !tc::mbar_try_wait(bar, parity); ++i)
        if (i > (1u << 28)) __trap();                 // seconds, not a hung GPU, if a transaction count was ever wrong
}

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// Arrival without memory ordering: the barrier before exit only keeps a CTA's shared memory alive while peers may read it.
// (arrive.release compiles to MEMBAR.ALL.GPU + ERRBAR: after the epilogue's global stores that is a ~1 us wait per launch.)
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_relaxed() { asm volatile("barrier.cluster.wait.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t mapa(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

struct ProjArgs {
    int R, N, K, NB;
    // activation operand: either TMA over compact bf16x3 rows (tmA), or LayerNorm of fp32 rows staged by the epilogue warps
    int ln_mode;
    const float* x; long long ldx;
    const float* stats_in;              // [K / 128][R][2] (mean, M2) of each 128-column tile of x
    const float* ln_g; const float* ln_b; float ln_eps;
    // folded LayerNorm (TMA operand = the RAW rows x as bf16x3, weights pre-multiplied by gamma): the epilogue applies
    // y = rstd * (acc - mean * fold_u[n]) + fold_c[n], with fold_u = W gamma, fold_c = W beta + bias (both from the host)
    const float* fold_u;                // fold_c travels as `bias`
    // epilogue
    const float* bias; int act;
    const float* residual; long long ldr;
    float* out; long long ldo;
    __nv_bfloat16* split_out;           // compact bf16x3 rows [R][3 N]
    float* stats_out;                   // [N / 128][R][2]
    const char* pf; long long pf_bytes;  // span the NEXT kernel of the chain streams (its weights): fetched into L2 from here
    const char* pf2; long long pf2_bytes;  // a second span (K/V a later attention kernel streams), same treatment
    // dense self-attention caches of the layer whose attention runs two launches from now (avsr_dec_proj_prefetch_self_kv):
    // positions [0, *step) of every (utterance, head) are fetched into L2 after this kernel's griddepcontrol.wait
    const float* skd; const float* svd; int s_lmax, s_nuh; const int* s_step;
    // second column range (avsr_dec_proj_dual): output features [n1, N) go to out2 / take residual2 (both indexed from column
    // n1 on) and get neither the folded LayerNorm nor statistics / a bf16x3 copy; n1 = 0: one range.  n1 % 128 == 0.
    int n1; const float* residual2; long long ldr2; float* out2; long long ldo2;
    int early_trigger;                  // dev knob AVSR_PROJ_EARLY_TRIGGER: griddepcontrol.launch_dependents before the prologue
};

// three bf16 terms of 8 consecutive fp32 values -> one 16-byte chunk per term
__device__ __forceinline__ void split3_chunk(const float (&v)[8], uint4& c1, uint4& c2, uint4& c3) {
    __align__(16) __nv_bfloat16 t1[8], t2[8], t3[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        t1[i] = __float2bfloat16_rn(v[i]);
        const float r1 = v[i] - __bfloat162float(t1[i]);
        t2[i] = __float2bfloat16_rn(r1);
        t3[i] = __float2bfloat16_rn(r1 - __bfloat162float(t2[i]));
    }
    c1 = *reinterpret_cast<const uint4*>(t1);
    c2 = *reinterpret_cast<const uint4*>(t2);
    c3 = *reinterpret_cast<const uint4*>(t3);
}

// LN = the operand is LayerNorm(x) staged by the epilogue warps (its own instantiation: the staging code is a third of the
// kernel's instructions and the decode position never runs it with the LayerNorm folded - a smaller kernel spends less of a
// ~8 us launch on instruction fetch, ncu: no_instruction 21-26 % of the stall samples)
template <bool LN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
dec_proj_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmA, const ProjArgs p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int NB = p.NB, K = p.K, N = p.N, R = p.R;
    const int A_TILE = NB * BK * 2, STAGE_BYTES = 3 * W_TILE + 3 * A_TILE;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);
    float* s_mean = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256);      // [MAX_NB]
    float* s_rstd = s_mean + MAX_NB;                                                   // [MAX_NB]
    float* red = reinterpret_cast<float*>(smem);       // [NB][128] fp32 accumulator tile of this CTA (aliases the pipeline stages)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (p.early_trigger) pdl_trigger();
    const uint32_t crank = cluster_ctarank(), csize = cluster_nctarank();
    const int m0 = blockIdx.x * BM;                    // first output feature of the tile
    const int n0 = blockIdx.z * NB;                    // first activation row of the tile
    const int nkb = K / BK;
    const int kb0 = (int)((long long)crank * nkb / csize), kb1 = (int)((long long)(crank + 1) * nkb / csize);
    const uint32_t w_tx = 3 * W_TILE, a_tx = 3 * (uint32_t)A_TILE;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            tc::mbar_init(&full[s], LN ? 1 + 4 : 1);      // producer's expect_tx arrival (+ one per staging warp)
            tc::mbar_init(&empty[s], 1);
        }
        tc::mbar_init(tfull, 1);
        tc::fence_barrier_init();
        tc::tma_prefetch_desc(&tmW);
        if (!LN) tc::tma_prefetch_desc(&tmA);
    }
    if (warp == 1) {
        tc::tmem_alloc(tmem_slot, TMEM_COLS);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (!p.early_trigger) pdl_trigger();               // dependents may start their own prologue

    if (warp == 0) {
        if (lane == 0) {
            // weights do not depend on the previous kernel: fill the pipeline with them first, then wait for the grid that
            // produces the activations, then complete the same stages with the activation tiles
            const int npre = min(STAGES, kb1 - kb0);
            for (int i = 0; i < npre; ++i) {
                uint8_t* sw = smem + i * STAGE_BYTES;
                tc::mbar_arrive_expect_tx(&full[i], LN ? w_tx : w_tx + a_tx);
#pragma unroll
                for (int j = 0; j < 3; ++j) tc::tma_load_2d(sw + j * W_TILE, &tmW, &full[i], (kb0 + i) * BK + j * K, m0);
            }
            pdl_wait();
            if (!LN) {
                for (int i = 0; i < npre; ++i) {
                    uint8_t* sa = smem + i * STAGE_BYTES + 3 * W_TILE;
#pragma unroll
                    for (int j = 0; j < 3; ++j) tc::tma_load_2d(sa + j * A_TILE, &tmA, &full[i], (kb0 + i) * BK + j * K, n0);
                }
            }
            if (p.skd != nullptr) {
                // (utterance, head) spans of the dense caches: K = 8 planes of [pos][8] floats, V = [pos][64] floats
                const int L = *p.s_step;                   // positions 0 .. L-1 exist (read after the wait: the chain wrote it)
                if (L > 0) {
                    const long long ncta = (long long)gridDim.x * gridDim.y * gridDim.z;
                    const long long cta = ((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
                    const unsigned kb_bytes = (unsigned)L * 32u, v_bytes = (unsigned)L * 256u;
                    for (long long uh = cta; uh < p.s_nuh; uh += ncta) {
                        const float* kbase = p.skd + uh * (long long)p.s_lmax * 64;
                        const float* vbase = p.svd + uh * (long long)p.s_lmax * 64;
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(kbase + (long long)j * p.s_lmax * 8), "r"(kb_bytes) : "memory");
                        for (unsigned o = 0; o < v_bytes; o += 32768u) {
                            const unsigned n = min(32768u, v_bytes - o);
                            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const char*>(vbase) + o), "r"(n) : "memory");
                        }
                    }
                }
            }
            int stage = npre % STAGES;
            uint32_t phase = (npre == STAGES) ? 1u : 0u;
            for (int kb = kb0 + npre; kb < kb1; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* sw = smem + stage * STAGE_BYTES;
                tc::mbar_arrive_expect_tx(&full[stage], LN ? w_tx : w_tx + a_tx);
#pragma unroll
                for (int j = 0; j < 3; ++j) tc::tma_load_2d(sw + j * W_TILE, &tmW, &full[stage], kb * BK + j * K, m0);
                if (!LN) {
#pragma unroll
                    for (int j = 0; j < 3; ++j) tc::tma_load_2d(sw + 3 * W_TILE + j * A_TILE, &tmA, &full[stage], kb * BK + j * K, n0);
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // the weights of the next projection (and K/V a later attention streams) do not depend on anything: ask the L2 for
            // this CTA's share now (fire and forget), so that their consumers find them on chip instead of paying HBM latency
#pragma unroll
            for (int which = 0; which < 2; ++which) {
                const char* pfp = which == 0 ? p.pf : p.pf2;
                const long long pfb = which == 0 ? p.pf_bytes : p.pf2_bytes;
                if (pfp == nullptr) continue;
                constexpr long long PIECE = 16 * 1024;
                const long long ncta = (long long)gridDim.x * gridDim.y * gridDim.z;
                const long long cta = ((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
                const long long share = ((pfb + ncta - 1) / ncta + PIECE - 1) / PIECE * PIECE;
                const long long lo = cta * share, hi = min(pfb, lo + share);
                for (long long o = lo; o < hi; o += PIECE) {
                    const unsigned n = (unsigned)min(PIECE, hi - o) & ~15u;
                    if (n) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pfp + o), "r"(n) : "memory");
                }
            }
            const uint32_t idesc = tc::umma_idesc_bf16(BM, (uint32_t)NB);
            int stage = 0;
            uint32_t phase = 0;
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
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            tc::umma_commit(tfull);
        }
    } else {
        const int et = threadIdx.x - 64;               // 0 .. 127
        const int quad = warp & 3;
        pdl_wait();                                    // x / stats / residual are written by the previous kernels of the chain
        if (LN) {
            // Everything this phase needs from global memory is requested in as few dependent round trips as possible: the
            // fp32 rows of the first k block and the row statistics go out together; the rows of k block i + 1 are requested
            // before k block i is converted and stored.
            const int ch = et & 7;                     // 16-byte chunk = 8 consecutive columns; fixed per thread
            const int rbase = et >> 3;                 // this thread's rows: rbase + 16 i
            constexpr int MAXI = MAX_NB / 16;
            const int NI = NB / 16;
            float4 xa[MAXI][2], xb[MAXI][2], ga[4], gb[4];
            auto load = [&](float4 (&xr)[MAXI][2], float4 (&gv)[4], int kb) {
                const int c0 = kb * BK + ch * 8;
                gv[0] = __ldg(reinterpret_cast<const float4*>(p.ln_g + c0)); gv[1] = __ldg(reinterpret_cast<const float4*>(p.ln_g + c0 + 4));
                gv[2] = __ldg(reinterpret_cast<const float4*>(p.ln_b + c0)); gv[3] = __ldg(reinterpret_cast<const float4*>(p.ln_b + c0 + 4));
#pragma unroll
                for (int i = 0; i < MAXI; ++i) {
                    const int row = n0 + rbase + 16 * i;
                    if (i < NI && row < R) {
                        const float* xp = p.x + (long long)row * p.ldx + c0;
                        xr[i][0] = __ldcg(reinterpret_cast<const float4*>(xp));
                        xr[i][1] = __ldcg(reinterpret_cast<const float4*>(xp + 4));
                    } else {
                        xr[i][0] = xr[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
            };
            load(xa, ga, kb0);
            // ---- row statistics: merge the (mean, M2) of the K / 128 column tiles of every row in tile order
            if (et < NB) {
                const int row = n0 + et;
                float mean = 0.f, rstd = 0.f;
                if (row < R) {
                    const int nt = K / 128;            // <= 8 (checked on the host): one batch of loads, one L2 round trip
                    float2 st[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        st[u] = (u < nt) ? __ldcg(reinterpret_cast<const float2*>(p.stats_in + ((long long)u * R + row) * 2)) : make_float2(0.f, 0.f);
                    float ms = 0.f;
#pragma unroll
                    for (int u = 0; u < 8; ++u) ms += st[u].x;
                    mean = ms / (float)nt;
                    // Chan's merge of equal-sized tiles in tile order: M2 = sum_t (M2_t + 128 (mean_t - mean)^2)
                    float m2 = 0.f;
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const float d = st[u].x - mean;
                        if (u < nt) m2 += st[u].y + 128.f * d * d;
                    }
                    rstd = rsqrtf(m2 / (float)K + p.ln_eps);
                }
                s_mean[et] = mean;
                s_rstd[et] = rstd;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            // ---- stage LayerNorm(x) of this CTA's k blocks as three bf16 terms in the UMMA layout: a tile is NB rows of
            //      128 bytes (64 bf16), 16-byte chunk c of row r at r * 128 + ((c ^ (r & 7)) << 4) (128-byte swizzle)
            int stage = 0;
            uint32_t phase = 0;
            auto process = [&](const float4 (&xr)[MAXI][2], const float4 (&gv)[4], int kb) {
                if (kb - kb0 >= STAGES) mbar_wait(&empty[stage], phase ^ 1);
                const float g[8] = {gv[0].x, gv[0].y, gv[0].z, gv[0].w, gv[1].x, gv[1].y, gv[1].z, gv[1].w};
                const float b[8] = {gv[2].x, gv[2].y, gv[2].z, gv[2].w, gv[3].x, gv[3].y, gv[3].z, gv[3].w};
                uint8_t* sa = smem + stage * STAGE_BYTES + 3 * W_TILE;
#pragma unroll
                for (int i = 0; i < MAXI; ++i) {
                    if (i < NI) {
                        const int r = rbase + 16 * i;
                        const float xin[8] = {xr[i][0].x, xr[i][0].y, xr[i][0].z, xr[i][0].w, xr[i][1].x, xr[i][1].y, xr[i][1].z, xr[i][1].w};
                        const float mean = s_mean[r], rstd = s_rstd[r];
                        float v[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = (n0 + r < R) ? (xin[j] - mean) * rstd * g[j] + b[j] : 0.f;
                        uint4 c1, c2, c3;
                        split3_chunk(v, c1, c2, c3);
                        const uint32_t off = (uint32_t)r * 128u + (uint32_t)((ch ^ (r & 7)) << 4);
                        *reinterpret_cast<uint4*>(sa + off) = c1;
                        *reinterpret_cast<uint4*>(sa + A_TILE + off) = c2;
                        *reinterpret_cast<uint4*>(sa + 2 * A_TILE + off) = c3;
                    }
                }
                tc::fence_proxy_async();               // generic-proxy stores -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&full[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            };
            for (int kb = kb0; kb < kb1; kb += 2) {
                if (kb + 1 < kb1) load(xb, gb, kb + 1);
                process(xa, ga, kb);
                if (kb + 1 < kb1) {
                    if (kb + 2 < kb1) load(xa, ga, kb + 2);
                    process(xb, gb, kb + 1);
                }
            }
        }
        // ---- accumulator tile -> shared memory as [row][feature] fp32 (conflict-free: a register = a row, lanes = features)
        mbar_wait(tfull, 0);
        tc::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
        const int f = quad * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < NB / 32; ++c) {
            uint32_t r[32];
            tc::tmem_ld_32x32(taddr + c * 32, r);
            tc::tmem_ld_wait();
            const uint32_t rs = tc::smem_u32(red) + (uint32_t)((c * 32) * BM + f) * 4u;
#pragma unroll
            for (int j = 0; j < 32; ++j)                  // explicit st.shared: the compiler only sees a generic pointer here
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(rs + (uint32_t)(j * BM * 4)), "r"(r[j]) : "memory");
        }
        tc::tc_fence_before();
    }
    // ---- all K splits of the tile are in the shared memories of the cluster
    __syncwarp();
    cluster_arrive();
    cluster_wait();
    if (warp >= 2) {
        const int ew = warp - 2;                       // 0 .. 3
        const int rows_here = min(NB, R - n0);         // valid rows of this row tile
        const int rpc = (rows_here + (int)csize - 1) / (int)csize;
        const int r_lo = (int)crank * rpc, r_hi = min(rows_here, r_lo + rpc);
        const int fcol = m0 + lane * 4;                // this lane's four output features
        // vector path: N and the pitches are multiples of 4, so a group of four features is in or out as a whole; otherwise
        // (the output layer, N = 5049 with dense rows) every element is guarded and stored on its own
        const bool second = p.n1 > 0 && m0 >= p.n1;    // tile-uniform: which column range this tile belongs to
        const float* const residual = second ? p.residual2 : p.residual;
        const long long ldr = second ? p.ldr2 : p.ldr;
        float* const out = second ? p.out2 : p.out;
        const long long ldo = second ? p.ldo2 : p.ldo;
        const int cb = second ? p.n1 : 0;              // first column of the range's own buffers
        const int nsw = p.n1 > 0 ? p.n1 : N;           // row width of the bf16x3 copy (first range only)
        const float* const fold_u = second ? nullptr : p.fold_u;
        __nv_bfloat16* const split_out = second ? nullptr : p.split_out;
        float* const stats_out = second ? nullptr : p.stats_out;
        const bool vec = (N & 3) == 0 && (ldo & 3) == 0 && (ldr & 3) == 0;
        const bool fok = fcol < N;
        const uint32_t red_s = tc::smem_u32(red);
        float4 u4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (fold_u != nullptr && fok) {
            if (vec) u4 = __ldg(reinterpret_cast<const float4*>(fold_u + fcol));
            else {
                u4.x = __ldg(fold_u + fcol);
                if (fcol + 1 < N) u4.y = __ldg(fold_u + fcol + 1);
                if (fcol + 2 < N) u4.z = __ldg(fold_u + fcol + 2);
                if (fcol + 3 < N) u4.w = __ldg(fold_u + fcol + 3);
            }
        }
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias != nullptr && fok) {
            if (vec) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + fcol));
            else {
                bias4.x = __ldg(p.bias + fcol);
                if (fcol + 1 < N) bias4.y = __ldg(p.bias + fcol + 1);
                if (fcol + 2 < N) bias4.z = __ldg(p.bias + fcol + 2);
                if (fcol + 3 < N) bias4.w = __ldg(p.bias + fcol + 3);
            }
        }
        for (int r = r_lo + ew; r < r_hi; r += 4) {
            const int row = n0 + r;
            float4 res4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (residual != nullptr && fok) {
                const float* rp = residual + (long long)row * ldr + (fcol - cb);
                if (vec) res4 = __ldcg(reinterpret_cast<const float4*>(rp));
                else {
                    res4.x = __ldcg(rp);
                    if (fcol + 1 < N) res4.y = __ldcg(rp + 1);
                    if (fcol + 2 < N) res4.z = __ldcg(rp + 2);
                    if (fcol + 3 < N) res4.w = __ldcg(rp + 3);
                }
            }
            const uint32_t a = red_s + (uint32_t)(r * BM + lane * 4) * 4u;
            float4 t[MAX_CLUSTER];
#pragma unroll
            for (int z = 0; z < MAX_CLUSTER; ++z)
                if (z < (int)csize) t[z] = ld_dsmem_f4(mapa(a, (uint32_t)z));
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int z = 0; z < MAX_CLUSTER; ++z)
                if (z < (int)csize) { v.x += t[z].x; v.y += t[z].y; v.z += t[z].z; v.w += t[z].w; }       // split order: deterministic
            if (fold_u != nullptr) {
                // LayerNorm of the operand row, applied to the finished sums (see ProjArgs): merge the row's tile statistics
                // in tile order (lanes 0 .. K/128-1 hold one tile each; xor-shuffle trees are order-independent per lane count)
                const int nt = K / 128;
                float2 st = make_float2(0.f, 0.f);
                if (lane < nt) st = __ldcg(reinterpret_cast<const float2*>(p.stats_in + ((long long)lane * R + row) * 2));
                const float mean = warp_sum(st.x) / (float)nt;
                const float d = st.x - mean;
                const float m2 = warp_sum(lane < nt ? st.y + 128.f * d * d : 0.f);
                const float rstd = rsqrtf(m2 / (float)K + p.ln_eps);
                v.x = rstd * (v.x - mean * u4.x); v.y = rstd * (v.y - mean * u4.y);
                v.z = rstd * (v.z - mean * u4.z); v.w = rstd * (v.w - mean * u4.w);
            }
            v.x += bias4.x; v.y += bias4.y; v.z += bias4.z; v.w += bias4.w;
            if (p.act == AVSR_ACT_RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            v.x += res4.x; v.y += res4.y; v.z += res4.z; v.w += res4.w;
            if (fok) {
                if (out != nullptr) {
                    float* op = out + (long long)row * ldo + (fcol - cb);
                    if (vec) *reinterpret_cast<float4*>(op) = v;
                    else {
                        op[0] = v.x;
                        if (fcol + 1 < N) op[1] = v.y;
                        if (fcol + 2 < N) op[2] = v.z;
                        if (fcol + 3 < N) op[3] = v.w;
                    }
                }
                if (split_out != nullptr) avsr_split3c_store4(split_out + (long long)row * 3 * nsw, nsw, fcol, v);
            }
            if (stats_out != nullptr) {                // N % 128 == 0 here: every lane holds valid features
                const float mean = warp_sum((v.x + v.y) + (v.z + v.w)) * (1.f / 128.f);
                const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
                const float m2 = warp_sum((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3));
                if (lane == 0) *reinterpret_cast<float2*>(stats_out + ((long long)blockIdx.x * R + row) * 2) = make_float2(mean, m2);
            }
        }
    }
    // ---- nobody leaves while a peer may still read its shared memory
    __syncwarp();
    cluster_arrive_relaxed();
    cluster_wait_relaxed();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, TMEM_COLS);
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

size_t smem_bytes(int nb) { return (size_t)STAGES * (3 * W_TILE + 3 * nb * BK * 2) + 1024 + 256 + 2 * MAX_NB * sizeof(float); }

const float* g_skd = nullptr;
const float* g_svd = nullptr;
int g_slmax = 0, g_snuh = 0;
const int* g_sstep = nullptr;
const char* g_pf2 = nullptr;
long long g_pf2_bytes = 0;
int g_force_splits = 0;
int g_early_trigger = 0;
int g_sm_budget = 0;                                   // SMs one projection may occupy (0 = all): concurrent decode chains share the GPU
bool g_configured = false;
int configure() {
    if (!g_configured) {
        AVSR_CHECK_CUDA(cudaFuncSetAttribute(dec_proj_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(MAX_NB)));
        AVSR_CHECK_CUDA(cudaFuncSetAttribute(dec_proj_kernel<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        AVSR_CHECK_CUDA(cudaFuncSetAttribute(dec_proj_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(MAX_NB)));
        AVSR_CHECK_CUDA(cudaFuncSetAttribute(dec_proj_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        const char* e = getenv("AVSR_X3C_SPLITS");
        if (e) g_force_splits = atoi(e);
        e = getenv("AVSR_PROJ_EARLY_TRIGGER");
        if (e) g_early_trigger = atoi(e);
        g_configured = true;
    }
    return AVSR_OK;
}

// clusters of `cs` CTAs (smem for `nb` rows) that can be resident at once, cached per (cs, nb / 32)
int max_active_clusters(int cs, int nb) {
    static int cache[MAX_CLUSTER + 1][5];
    static bool init = false;
    if (!init) {
        for (auto& row : cache) for (int& v : row) v = -1;
        init = true;
    }
    int& c = cache[cs][nb / 32];
    if (c >= 0) return c;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1, cs, 1);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = smem_bytes(nb);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = cs;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, dec_proj_kernel<false>, &cfg) != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    c = n;
    return c;
}

// K splits (= cluster size) for a projection: as many CTAs as fit in ONE wave of co-resident clusters.
int plan_splits(int R, int N, int K, int nb) {
    const int force = g_force_splits;                  // dev knob (avsr_dec_proj_force_splits / AVSR_X3C_SPLITS)
    const int tiles = cdiv(N, BM);
    const int nkb = K / BK;
    const int budget = (g_sm_budget > 0 && g_sm_budget < sm_count()) ? g_sm_budget : sm_count();
    const int share = sm_count() / budget;             // chains that run side by side: each needs its clusters resident
    int smax = budget / tiles;
    if (smax > MAX_CLUSTER) smax = MAX_CLUSTER;
    if (smax > nkb) smax = nkb;
    if (smax < 1) smax = 1;
    if (force > 0) return force < smax ? force : smax;
    for (int s = smax; s > 1; --s)
        if (max_active_clusters(s, nb) >= tiles * share) {
            // the critical path is the CTA with the most k blocks: take the SMALLEST cluster that has the same maximum (fewer
            // peers to reduce over and fewer CTAs to place; K = 1024: 8 x 2 k blocks instead of 10 x {1, 2})
            const int worst = cdiv(nkb, s);
            while (s > 1 && cdiv(nkb, s - 1) == worst) --s;
            return s;
        }
    return 1;
}

}  // namespace

// Cluster size (= K splits) avsr_dec_proj uses for this shape on the current device; informational.
extern "C" int avsr_dec_proj_splits(int R, int N, int K) {
    if (R <= 0 || N <= 0 || K <= 0 || (K % BK) != 0) return AVSR_ERR_ARG;
    if (configure() != AVSR_OK) return AVSR_ERR_CUDA;
    const int r32 = ((R + 31) / 32) * 32;
    return plan_splits(R, N, K, r32 < MAX_NB ? r32 : MAX_NB);
}

// Clusters of `cluster_size` CTAs of the projection kernel (operand tiles of `nb` rows, nb in {32, 64, 96, 128}) that the
// current device can keep resident at once (cudaOccupancyMaxActiveClusters); what the split planner consults.
extern "C" int avsr_dec_proj_max_clusters(int cluster_size, int nb) {
    if (cluster_size < 1 || cluster_size > MAX_CLUSTER || nb < 32 || nb > MAX_NB || (nb % 32) != 0) return AVSR_ERR_ARG;
    if (configure() != AVSR_OK) return AVSR_ERR_CUDA;
    return max_active_clusters(cluster_size, nb);
}

// Dev knob: force the cluster size (= K splits) of every following avsr_dec_proj call (0 = plan automatically).  A forced
// size whose clusters do not all fit at once simply runs in more than one wave.
extern "C" int avsr_dec_proj_force_splits(int splits) {
    if (splits < 0 || splits > MAX_CLUSTER) return AVSR_ERR_ARG;
    if (configure() != AVSR_OK) return AVSR_ERR_CUDA;
    g_force_splits = splits;
    return AVSR_OK;
}

// SMs the work of ONE projection launch is planned for (0 = the whole device).  With G decode chains running side by side
// on separate streams, a budget of sm_count / G keeps every launch's clusters co-resident with the other chains' launches
// instead of serialising behind them.
extern "C" int avsr_dec_proj_set_sm_budget(int sms) {
    if (sms < 0) return AVSR_ERR_ARG;
    if (configure() != AVSR_OK) return AVSR_ERR_CUDA;
    g_sm_budget = sms;
    return AVSR_OK;
}

// A second L2 fetch-ahead span for the NEXT avsr_dec_proj / avsr_dec_proj_folded launch only (e.g. the cross-attention K/V
// that an attention kernel two launches later streams); cleared by that launch.
extern "C" int avsr_dec_proj_also_prefetch(const void* span, long long bytes) {
    if ((span != nullptr && (((uintptr_t)span & 15) != 0 || bytes <= 0))) return AVSR_ERR_ARG;
    g_pf2 = (const char*)span;
    g_pf2_bytes = bytes;
    return AVSR_OK;
}

// The NEXT avsr_dec_proj* launch also asks the L2 for positions [0, *step) of the dense self-attention caches kd / vd of one
// layer (layouts of avsr_dec_cache_promote; n_utt_heads = utterances x 16 spans, lmax positions each): issued after that
// launch's griddepcontrol.wait, so the history streams from HBM while the latency-bound projections in between run and the
// self-attention two launches later reads it from L2.  One-shot.
extern "C" int avsr_dec_proj_prefetch_self_kv(const float* kd, const float* vd, int lmax, int n_utt_heads, const int* step) {
    if (kd != nullptr && (vd == nullptr || step == nullptr || lmax <= 0 || n_utt_heads <= 0 || ((uintptr_t)kd & 31) || ((uintptr_t)vd & 15))) return AVSR_ERR_ARG;
    g_skd = kd; g_svd = vd; g_slmax = lmax; g_snuh = n_utt_heads; g_sstep = step;
    return AVSR_OK;
}

// One decoder-step projection with its glue:  y = act(a W^T + bias) + residual  for R rows.
//   operand a:  A3 != NULL: compact bf16x3 rows [R, 3K] (pitch lda elements), loaded with TMA; else a = LayerNorm(x) with
//               x [R, K] fp32 (pitch ldx), stats_in [K/128][R][2] = (mean, M2) of every 128-column tile of x as a previous call
//               wrote them through stats_out, ln_g / ln_b [K], eps (K % 128 == 0).
//   W3:         compact bf16x3 weights [N, 3K] (pitch ldw).  K % 64 == 0, N % 4 == 0.
//   outputs:    out [R, N] fp32 (pitch ldo) and / or split_out [R, 3N] compact bf16x3; stats_out [N/128][R][2] (N % 128 == 0)
//               for a later LayerNorm-mode call.  residual may alias out (each element is read and written by one thread).
//   l2_prefetch: optional span (the weights of the NEXT projection of the chain) that the kernel asks the L2 to fetch.
struct DualRange { int n1; const float* residual2; long long ldr2; float* out2; long long ldo2; };

static int dec_proj_launch(const void* A3, long long lda, const float* x, long long ldx, const float* stats_in, const float* ln_g,
                             const float* ln_b, float ln_eps, const float* fold_u, const void* W3, long long ldw, int R, int N, int K, const float* bias, int act,
                             const float* residual, long long ldr, float* out, long long ldo, void* split_out, float* stats_out,
                             const void* l2_prefetch, long long l2_prefetch_bytes, cudaStream_t stream,
                             const DualRange dual = DualRange{0, nullptr, 0, nullptr, 0}) {
    AVSR_REQUIRE(W3 && R > 0 && N > 0 && K > 0 && (K % BK) == 0, "avsr_dec_proj: bad shape R=%d N=%d K=%d (K must be a multiple of 64)", R, N, K);
    AVSR_REQUIRE(!split_out || (N & 3) == 0, "avsr_dec_proj: split_out needs N %% 4 == 0");
    AVSR_REQUIRE((A3 != nullptr) != (x != nullptr), "avsr_dec_proj: exactly one of A3 (bf16x3 rows) and x (LayerNorm mode) must be given");
    AVSR_REQUIRE(!fold_u || (A3 && stats_in && bias && (K % 128) == 0 && K <= 4096 && ((uintptr_t)fold_u & 15) == 0 && ((uintptr_t)stats_in & 7) == 0),
                 "avsr_dec_proj_folded: needs bf16x3 rows, stats_in, fold_u / fold_c and K %% 128 == 0 (K <= 4096)");
    AVSR_REQUIRE(x == nullptr || (stats_in && ln_g && ln_b && (K % 128) == 0 && K <= 1024 && (ldx & 3) == 0 && ((uintptr_t)x & 15) == 0 &&
                                  ((uintptr_t)ln_g & 15) == 0 && ((uintptr_t)ln_b & 15) == 0 && ((uintptr_t)stats_in & 7) == 0),
                 "avsr_dec_proj: LayerNorm mode needs stats_in / gamma / beta, K %% 128 == 0, K <= 1024 and 16-byte aligned rows");
    AVSR_REQUIRE(out || split_out, "avsr_dec_proj: no output");
    AVSR_REQUIRE(!l2_prefetch || (((uintptr_t)l2_prefetch & 15) == 0 && l2_prefetch_bytes > 0), "avsr_dec_proj: prefetch span must be 16-byte aligned");
    AVSR_REQUIRE(act == AVSR_ACT_NONE || act == AVSR_ACT_RELU, "avsr_dec_proj: activation %d unsupported", act);
    AVSR_REQUIRE(!stats_out || (N % 128) == 0, "avsr_dec_proj: stats_out needs N %% 128 == 0");
    AVSR_REQUIRE(dual.n1 == 0 || (dual.n1 > 0 && dual.n1 < N && (dual.n1 % BM) == 0 && (N & 3) == 0 && dual.out2 && (dual.ldo2 & 3) == 0 &&
                                  ((uintptr_t)dual.out2 & 15) == 0 && (!dual.residual2 || ((dual.ldr2 & 3) == 0 && ((uintptr_t)dual.residual2 & 15) == 0))),
                 "avsr_dec_proj_dual: the second range starts on a 128-column boundary and needs an aligned output");
    AVSR_REQUIRE((!out || ((uintptr_t)out & 15) == 0) && (!residual || ((uintptr_t)residual & 15) == 0) && (!bias || ((uintptr_t)bias & 15) == 0) &&
                     (!split_out || ((uintptr_t)split_out & 7) == 0) && (!stats_out || ((uintptr_t)stats_out & 7) == 0),
                 "avsr_dec_proj: outputs / residual / bias must be 16-byte aligned");
    int rc = configure();
    if (rc != AVSR_OK) return rc;
    const int r32 = ((R + 31) / 32) * 32;
    const int nb = r32 < MAX_NB ? r32 : MAX_NB;
    const int tiles_m = cdiv(N, BM), tiles_n = cdiv(R, nb);
    const int splits = plan_splits(R, N, K, nb);
    CUtensorMap tw, ta;
    rc = tc::make_tmap_2d_bf16(&tw, W3, (uint64_t)N, (uint64_t)3 * K, (uint64_t)ldw, BM, BK);
    if (rc != AVSR_OK) return rc;
    if (A3 != nullptr) {
        rc = tc::make_tmap_2d_bf16(&ta, A3, (uint64_t)R, (uint64_t)3 * K, (uint64_t)lda, (uint32_t)nb, BK);
        if (rc != AVSR_OK) return rc;
    } else {
        ta = tw;
    }
    ProjArgs p = {R, N, K, nb, x != nullptr ? 1 : 0, x, ldx, stats_in, ln_g, ln_b, ln_eps, fold_u, bias, act, residual, ldr, out, ldo,
                  (__nv_bfloat16*)split_out, stats_out, (const char*)l2_prefetch, l2_prefetch ? l2_prefetch_bytes : 0, g_pf2, g_pf2 ? g_pf2_bytes : 0, g_skd, g_svd, g_slmax, g_snuh, g_sstep,
                  dual.n1, dual.residual2, dual.ldr2, dual.out2, dual.ldo2, g_early_trigger};
    g_pf2 = nullptr;                                   // one-shot (avsr_dec_proj_also_prefetch)
    g_skd = g_svd = nullptr;                           // one-shot (avsr_dec_proj_prefetch_self_kv)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(tiles_m, splits, tiles_n);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = smem_bytes(nb);
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = 1;
    attr[1].val.clusterDim.y = splits;
    attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    if (x != nullptr) AVSR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, dec_proj_kernel<true>, tw, ta, p));
    else AVSR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, dec_proj_kernel<false>, tw, ta, p));
    return AVSR_OK;
}

extern "C" int avsr_dec_proj(const void* A3, long long lda, const float* x, long long ldx, const float* stats_in, const float* ln_g,
                             const float* ln_b, float ln_eps, const void* W3, long long ldw, int R, int N, int K, const float* bias, int act,
                             const float* residual, long long ldr, float* out, long long ldo, void* split_out, float* stats_out,
                             const void* l2_prefetch, long long l2_prefetch_bytes, cudaStream_t stream) {
    return dec_proj_launch(A3, lda, x, ldx, stats_in, ln_g, ln_b, ln_eps, nullptr, W3, ldw, R, N, K, bias, act, residual, ldr, out, ldo, split_out,
                           stats_out, l2_prefetch, l2_prefetch_bytes, stream);
}

// LayerNorm FOLDED into the projection: y = act(LayerNorm(x) W^T + bias) + residual computed as
//     y = act( rstd * (x (gamma . W)^T - mean * fold_u) + fold_c ) + residual,   fold_u[n] = sum_k gamma[k] W[n][k],
//     fold_c[n] = sum_k beta[k] W[n][k] + bias[n],
// so the operand is the RAW row x as compact bf16x3 (X3 [R, 3K], written by the projection that produced x through its
// split_out), loaded with TMA like any other operand, and nothing is normalised on the critical path: mean / rstd come from
// stats_in [K/128][R][2] in the epilogue.  W3g = compact bf16x3 of gamma . W (weights.fold_layernorm).  K % 128 == 0.
extern "C" int avsr_dec_proj_folded(const void* X3, long long lda, const float* stats_in, float ln_eps, const float* fold_u, const float* fold_c,
                                    const void* W3g, long long ldw, int R, int N, int K, int act, const float* residual, long long ldr, float* out,
                                    long long ldo, void* split_out, float* stats_out, const void* l2_prefetch, long long l2_prefetch_bytes,
                                    cudaStream_t stream) {
    AVSR_REQUIRE(fold_u && fold_c, "avsr_dec_proj_folded: fold_u / fold_c missing");
    return dec_proj_launch(X3, lda, nullptr, 0, stats_in, nullptr, nullptr, ln_eps, fold_u, W3g, ldw, R, N, K, fold_c, act, residual, ldr, out, ldo,
                           split_out, stats_out, l2_prefetch, l2_prefetch_bytes, stream);
}

// Two projections of the SAME operand in one launch: the weight rows [0, n1) and [n1, N) are two matrices stacked on top of each
// other; output features of the first range behave as in avsr_dec_proj_folded / avsr_dec_proj (folded LayerNorm when fold_u is
// given, bias, activation, residual, out, statistics), those of the second range get bias[n] + residual2 and go to out2
// (both indexed from column n1 on), without LayerNorm.  bias covers all N features.  n1 % 128 == 0.
// Use: the source-attention query of a decoder layer without a launch of its own ("query merge").
extern "C" int avsr_dec_proj_dual(const void* A3, long long lda, const float* stats_in, float ln_eps, const float* fold_u, const float* bias,
                                  const void* W3, long long ldw, int R, int N, int K, int n1, int act, const float* residual, long long ldr,
                                  float* out, long long ldo, const float* residual2, long long ldr2, float* out2, long long ldo2,
                                  float* stats_out, const void* l2_prefetch, long long l2_prefetch_bytes, cudaStream_t stream) {
    AVSR_REQUIRE(A3 && bias && out && out2 && n1 > 0, "avsr_dec_proj_dual: missing argument");
    const DualRange dual = {n1, residual2, ldr2, out2, ldo2};
    return dec_proj_launch(A3, lda, nullptr, 0, stats_in, nullptr, nullptr, ln_eps, fold_u, W3, ldw, R, N, K, bias, act, residual, ldr, out, ldo,
                           nullptr, stats_out, l2_prefetch, l2_prefetch_bytes, stream, dual);
}
