// bf16 GEMM on the 5th-generation tensor cores (tcgen05) for the AV-HuBERT encoder:
//     C[M,N] = epilogue( A[M,K] * B[N,K]^T )        A, B bf16 K-major, fp32 accumulation in TMEM.
// This single kernel carries every dense op of the encoder: the QKV/out/FFN projections of the 24
// transformer layers (reference: HF Wav2Vec2Attention/FeedForward called from
// src/nets/backend/backbones/avhubert.py:747-768), the modality projections and post_extract_proj
// (avhubert.py:187-198,486-502), the grouped positional conv and the ResNet-18 / 3D-conv frontend as im2col GEMMs
// (src/nets/backend/backbones/resnet.py:30-164).
//
// Structure (one persistent CTA per SM, 192 threads):
//   warp 0   : TMA producer  - cp.async.bulk.tensor 2D loads of 128x64 (A) and BNx64 (B) bf16 tiles, 128B swizzle
//   warp 1   : MMA issuer    - one thread issues tcgen05.mma (M=128, N=BN, K=16) x4 per k-block; owns the TMEM allocation
//   warps 2-5: epilogue      - tcgen05.ld the fp32 accumulator (double-buffered in TMEM), fused bias / activation /
//                              residual / row-mask, bf16 and/or fp32 stores
// Pipelines: smem ring full[]/empty[] (TMA <-> MMA) and TMEM tfull[]/tempty[] (MMA <-> epilogue), all mbarriers.
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_TILE_BYTES = BM * BK * 2;

template <int BN>
struct Cfg {
    static constexpr int B_TILE_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
    static constexpr int STAGES = (188 * 1024) / STAGE_BYTES;
    static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
    static constexpr int COLPAR_BYTES = 2 * 2 * BN * 4;        // per accumulator buffer: bias[BN] + prelu[BN] staged for the epilogue
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + COLPAR_BYTES;
};

// Row-wise arithmetic of one 32-column chunk of one accumulator row: v = act(acc + bias) + residual (or act after the residual),
// row mask.  `cb` / `cp` point at this chunk's per-column bias / PReLU slopes staged in shared memory by the epilogue warps
// BEFORE they wait for the accumulator (a global load here would expose a full memory latency per chunk: with two epilogue
// warps per scheduler nothing hides it).  `rbias` is the per-row bias (bias_mode 2), `res` the residual values of this chunk
// already in registers (has_res), both fetched ahead of use as well.
__device__ __forceinline__ void epilogue_math(const AvsrEpilogue& ep, int row, int M, const uint32_t (&r)[32], const float* cb, const float* cp,
                                              float rbias, const float (&res)[32], bool has_res, float (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    if (ep.bias != nullptr) {
        if (ep.bias_mode == 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += rbias;
        } else {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 b = *reinterpret_cast<const float4*>(cb + j);
                v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
            }
        }
    }
    if (ep.act != AVSR_ACT_NONE && !ep.act_after_residual) {
        if (ep.act == AVSR_ACT_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_erf_fast(v[j]);
        } else if (ep.act == AVSR_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = v[j] >= 0.f ? v[j] : v[j] * cp[j];
        }
    }
    if (has_res) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += res[j];
    }
    if (ep.act != AVSR_ACT_NONE && ep.act_after_residual) {
        if (ep.act == AVSR_ACT_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_erf_fast(v[j]);
        } else if (ep.act == AVSR_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = v[j] >= 0.f ? v[j] : v[j] * cp[j];
        }
    }
    if (ep.row_mask != nullptr && row < M && ep.row_mask[row] == 0) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
    }
}

// Stores of one finished chunk: thread = row, 32 consecutive columns.
__device__ __forceinline__ void epilogue_store_direct(const AvsrEpilogue& ep, int row, int col0, int M, int N, const float (&v)[32]) {
    if (row >= M) return;
    const int ncols = min(32, N - col0);
    if (ncols <= 0) return;
    if (ep.out_f32 != nullptr) {
        float* op = ep.out_f32 + (long long)row * ep.ld_f32 + col0;
        if (ncols == 32 && (ep.ld_f32 & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < ncols) op[j] = v[j];
        }
    }
    if (ep.out_bf16 != nullptr) {
        __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(ep.out_bf16) + (long long)row * ep.ld_bf16 + col0;
        if (ncols == 32 && (ep.ld_bf16 & 7) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                uint4 pk;
                __nv_bfloat162 t0 = __floats2bfloat162_rn(v[j], v[j + 1]);
                __nv_bfloat162 t1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
                __nv_bfloat162 t2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]);
                __nv_bfloat162 t3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
                pk.x = *reinterpret_cast<uint32_t*>(&t0);
                pk.y = *reinterpret_cast<uint32_t*>(&t1);
                pk.z = *reinterpret_cast<uint32_t*>(&t2);
                pk.w = *reinterpret_cast<uint32_t*>(&t3);
                *reinterpret_cast<uint4*>(op + j) = pk;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < ncols) op[j] = __float2bfloat16_rn(v[j]);
        }
    }
}

// residual values of one 32-column chunk of one row -> registers (zero where out of range)
__device__ __forceinline__ void load_residual(const AvsrEpilogue& ep, int row, int col0, int M, int N, float (&res)[32]) {
    const int ncols = min(32, N - col0);
    if (row >= M || ncols <= 0) {
#pragma unroll
        for (int j = 0; j < 32; ++j) res[j] = 0.f;
        return;
    }
    if (ep.res_dtype == 0) {
        const float* rp = reinterpret_cast<const float*>(ep.residual) + (long long)row * ep.ldr + col0;
        if (ncols == 32 && (ep.ldr & 3) == 0 && (reinterpret_cast<uintptr_t>(rp) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 t = *reinterpret_cast<const float4*>(rp + j);
                res[j] = t.x; res[j + 1] = t.y; res[j + 2] = t.z; res[j + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) res[j] = (j < ncols) ? rp[j] : 0.f;
        }
    } else {
        const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(ep.residual) + (long long)row * ep.ldr + col0;
        if (ncols == 32 && (ep.ldr & 7) == 0 && (reinterpret_cast<uintptr_t>(rp) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                const uint4 t = *reinterpret_cast<const uint4*>(rp + j);
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 f = __bfloat1622float2(h[q]);
                    res[j + 2 * q] = f.x; res[j + 2 * q + 1] = f.y;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) res[j] = (j < ncols) ? __bfloat162float(rp[j]) : 0.f;
        }
    }
}

// (Round 2 negative result: transposing every chunk through a per-warp shared-memory tile so that global accesses are 4 rows x
// 128 contiguous bytes per instruction instead of 32 rows x 16 bytes was measured SLOWER for the whole encoder, 32.8 ms against
// 30.2 ms per 32 x 375-frame batch: the two shared-memory round trips and warp barriers per chunk cost more than the 32-line
// store instructions they replace.  The direct stores stay.)

// Epilogue of one accumulator tile for one epilogue warp (TMEM lane quadrant `quad`, column half `half`): the first chunk's
// residual is requested BEFORE the wait for the accumulator, the next chunk's while the current one is processed.
template <int BN, bool BOUNDED>
__device__ __forceinline__ void epilogue_tile(const AvsrEpilogue& ep, uint32_t taddr, int m0, int n0, int M, int N, int quad, int half, int lane,
                                              const float* cpar, float rbias, uint64_t* tfull_bar, uint32_t tfull_parity) {
    constexpr int CHUNKS = BN / 32 / 2 > 0 ? BN / 32 / 2 : 1;
    const bool has_res = ep.residual != nullptr;
    const int row = m0 + quad * 32 + lane;
    const int c_first = half * CHUNKS;
    float res[32];
    if (has_res) load_residual(ep, row, n0 + c_first * 32, M, N, res);
    if (BOUNDED) {
        for (uint32_t i = 0; !tc::mbar_try_wait(tfull_bar, tfull_parity); ++i)
            if (i > (1u << 27)) __trap();
    } else {
        tc::mbar_wait(tfull_bar, tfull_parity);
    }
    tc::tc_fence_after();
#pragma unroll 1
    for (int cc = 0; cc < CHUNKS; ++cc) {
        const int c = c_first + cc, col0 = n0 + c * 32;
        uint32_t r[32];
        tc::tmem_ld_32x32(taddr + c * 32, r);
        const bool more = has_res && cc + 1 < CHUNKS;
        float v[32], res_next[32];
        if (more) load_residual(ep, row, col0 + 32, M, N, res_next);       // in flight while this chunk is processed
        tc::tmem_ld_wait();
        epilogue_math(ep, row, M, r, cpar + c * 32, cpar + BN + c * 32, rbias, res, has_res, v);
        epilogue_store_direct(ep, row, col0, M, N, v);
        if (more) {
#pragma unroll
            for (int j = 0; j < 32; ++j) res[j] = res_next[j];
        }
    }
}

constexpr int NUM_EPI_WARPS = 8;                    // two per TMEM lane quadrant, each taking half of the tile's columns
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;

// splits > 1: split-K. Work item = (z, m-tile, n-tile); item z covers k-blocks [z*kb_per_split, (z+1)*kb_per_split) and
// stores its raw fp32 partial sums at ep.out_f32 + z*M*ld_f32 (the caller reduces them in a fixed order).
// Implicit-GEMM convolution (ks x ks, stride s, pad ks/2, NHWC): A is never materialised; row m = output pixel (n, y, x) of the
// Ho x Wo output and k block kb = (filter tap, 64-channel slice), fetched by an im2col-mode TMA load of the activation tensor.
struct ConvA {
    int on, Ho, Wo, C, ks, stride;
    // Positional-conv mode (avsr_posconv_bf16_tc): grouped Conv1d(k = 128, pad = 64, 16 groups of 64 channels) over the packed
    // frames as an implicit banded GEMM.  Tile = (work item w = 128 frames of one utterance, group g); k block j = filter tap j:
    // the A tile is rows [q0 + j - 64, +128) x columns [64 g, +64) of THAT utterance's frames, a plain 2D TMA load through the
    // utterance's own tensor map (rows outside the utterance are zero-filled by the TMA unit); B = W[64 g .., 64 j ..].
    int pc_on, pc_nwork;
    const CUtensorMap* pc_maps;      // [utterances] tensor maps in global memory
    const int* pc_utt;               // [n_work] utterance of the work item
    const int* pc_off;               // [n_work] first packed frame of that utterance
    const int* pc_T;                 // [n_work] its length
    const int* pc_q0;                // [n_work] first frame (within the utterance) of the work item
};

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M_in, int N_in, int K,
               const AvsrEpilogue ep_in, int splits, int kb_per_split, const ConvA conv) {
    const int M = M_in, N = N_in;
    using C = Cfg<BN>;
    constexpr int STAGES = C::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * C::STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* colpar = reinterpret_cast<float*>(smem + STAGES * C::STAGE_BYTES + 256);       // [2 acc][bias BN | prelu BN]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_n = (N + BN - 1) / BN;
    const int tiles_m = (M + BM - 1) / BM;
    const int tiles_mn = conv.pc_on ? conv.pc_nwork * 16 : tiles_m * tiles_n;
    const int num_tiles = tiles_mn * splits;
    const int num_kb_total = (K + BK - 1) / BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            tc::mbar_init(&full[s], 1);
            tc::mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            tc::mbar_init(&tfull[a], 1);
            tc::mbar_init(&tempty[a], NUM_EPI_WARPS);
        }
        tc::fence_barrier_init();
        tc::tma_prefetch_desc(&tmA);
        tc::tma_prefetch_desc(&tmB);
    }
    if (warp == 1) {
        tc::tmem_alloc(tmem_slot, C::TMEM_COLS);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int z = tile / tiles_mn, t2 = tile - z * tiles_mn;
                const int m0 = (t2 / tiles_n) * BM, n0 = (t2 % tiles_n) * BN;
                const int kb0 = z * kb_per_split, kb1 = min(num_kb_total, kb0 + kb_per_split);
                if (conv.pc_on) {
                    const int w = t2 >> 4, g = t2 & 15;
                    const CUtensorMap* um = conv.pc_maps + conv.pc_utt[w];
                    const int q0 = conv.pc_q0[w];
                    for (int kb = kb0; kb < kb1; ++kb) {
                        tc::mbar_wait(&empty[stage], phase ^ 1);
                        uint8_t* sa = smem + stage * C::STAGE_BYTES;
                        tc::mbar_arrive_expect_tx(&full[stage], C::STAGE_BYTES);
                        tc::tma_load_2d(sa, um, &full[stage], g * 64, q0 + kb - 64);
                        tc::tma_load_2d(sa + A_TILE_BYTES, &tmB, &full[stage], kb * BK, g * 64);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    continue;
                }
                int cn = 0, cy = 0, cx = 0, cpb = 1;
                if (conv.on) {
                    const int hw = conv.Ho * conv.Wo, pad = conv.ks / 2;
                    cn = m0 / hw;
                    cy = (m0 - cn * hw) / conv.Wo;
                    cx = m0 - cn * hw - cy * conv.Wo;
                    cy = cy * conv.stride - pad;     // base pixel of the first output pixel, in input coordinates
                    cx = cx * conv.stride - pad;
                    cpb = conv.C / BK;               // k blocks per filter tap
                }
                for (int kb = kb0; kb < kb1; ++kb) {
                    tc::mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * C::STAGE_BYTES;
                    tc::mbar_arrive_expect_tx(&full[stage], C::STAGE_BYTES);
                    if (conv.on) {
                        const int tap = kb / cpb, c0 = (kb - tap * cpb) * BK;
                        tc::tma_load_im2col_4d(sa, &tmA, &full[stage], c0, cx, cy, cn, (uint16_t)(tap % conv.ks), (uint16_t)(tap / conv.ks));
                    } else
                    tc::tma_load_2d(sa, &tmA, &full[stage], kb * BK, m0);
                    tc::tma_load_2d(sa + A_TILE_BYTES, &tmB, &full[stage], kb * BK, n0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = tc::umma_idesc_bf16(BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int z = tile / tiles_mn;
                const int kb0 = z * kb_per_split, kb1 = min(num_kb_total, kb0 + kb_per_split);
                tc::mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    tc::mbar_wait(&full[stage], phase);
                    tc::tc_fence_after();
                    const uint32_t sa = tc::smem_u32(smem + stage * C::STAGE_BYTES);
                    const uint64_t adesc = tc::umma_desc_sw128(sa);
                    const uint64_t bdesc = tc::umma_desc_sw128(sa + A_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        tc::umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > kb0 || k != 0) ? 1u : 0u);
                    tc::umma_commit(&empty[stage]);     // frees the smem slot when these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tc::umma_commit(&tfull[acc]);           // accumulator complete -> epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        const int quad = warp & 3;                      // TMEM lane quadrant this warp may touch
        const int half = (warp - 2) >> 2;               // which half of the tile's columns
        const int etid = threadIdx.x - 64;              // 0 .. 32 * NUM_EPI_WARPS - 1
        const bool col_bias = ep_in.bias != nullptr && ep_in.bias_mode != 2;
        const bool has_prelu = ep_in.act == AVSR_ACT_PRELU && ep_in.prelu != nullptr;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int z = tile / tiles_mn, t2 = tile - z * tiles_mn;
            int m0 = (t2 / tiles_n) * BM, n0 = (t2 % tiles_n) * BN;
            int M = M_in, N = N_in;
            if (conv.pc_on) {                           // rows = packed frames of the work item's utterance, columns = the group's 64 channels
                const int w = t2 >> 4;
                m0 = conv.pc_off[w] + conv.pc_q0[w];
                M = conv.pc_off[w] + conv.pc_T[w];
                n0 = (t2 & 15) * 64;
                N = 1024;
            }
            AvsrEpilogue ep = ep_in;
            if (splits > 1) ep.out_f32 = ep_in.out_f32 + (long long)z * M * ep_in.ld_f32;
            const int row = m0 + quad * 32 + lane;
            // ---- everything the epilogue needs from global memory is requested BEFORE waiting for the accumulator
            float* cpar = colpar + acc * 2 * BN;
            if (col_bias || has_prelu) {
                for (int i = etid; i < BN; i += 32 * NUM_EPI_WARPS) {
                    const int col = n0 + i;
                    if (col_bias) cpar[i] = col < N ? __ldg(ep_in.bias + col) : 0.f;
                    if (has_prelu) cpar[BN + i] = col < N ? __ldg(ep_in.prelu + col) : 0.f;
                }
            }
            const float rbias = (ep_in.bias != nullptr && ep_in.bias_mode == 2 && row < M) ? __ldg(ep_in.bias + row) : 0.f;
            if (col_bias || has_prelu) asm volatile("bar.sync 1, %0;" ::"n"(32 * NUM_EPI_WARPS) : "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN;
            epilogue_tile<BN, false>(ep, taddr, m0, n0, M, N, quad, half, lane, cpar, rbias, &tfull[acc], acc_phase);
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&tempty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// CTA-PAIR variant (tcgen05 cta_group::2) for the big encoder GEMMs.  Why: the single-CTA kernel above streams 48 KB per k
// block (A 128x64 + B 256x64) for 512 tensor-pipe clocks = 96 B/clk per SM; the L2 delivers ~12 TB/s chip-wide (ncu r02:
// lts 1.26 GB per FFN launch in 102 us, the ~6300 B/clk cap of B300_MICROARCH.md), i.e. ~43 B/clk per SM, so the tensor pipe
// waits for operands half of the time (51.7 % active).  A pair of CTAs on the two SMs of a TPC computes a 256 x 256 tile with
// ONE MMA stream: each CTA stages its own 128 A rows and HALF of the B tile (128 of the 256 rows), the leader's
// tcgen05.mma.cta_group::2 reads both halves out of both shared memories and writes each CTA's 128 x 256 accumulator into
// that CTA's own TMEM.  Per CTA and k block: 32 KB for the same 512 clocks = 64 B/clk, and the smaller stage doubles the
// pipeline depth (6 stages instead of 3).
//   * both CTAs' producers signal the LEADER's full barrier (peer bit of the shared::cluster address cleared);
//   * the leader's commits are multicast to both CTAs' empty / accumulator-full barriers;
//   * both CTAs' epilogue warps release the accumulator buffer on the leader's barrier (remote mbarrier arrive).
// Epilogue, persistent tile loop and TMEM double buffering are those of the single-CTA kernel.
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;       // shared::cluster address of the same offset in the pair's even CTA
constexpr int BN2 = 256;
constexpr int B_HALF_BYTES = (BN2 / 2) * BK * 2;
constexpr int STAGE2_BYTES = A_TILE_BYTES + B_HALF_BYTES;     // 32 KB
constexpr int STAGES2 = 6;
constexpr int COLPAR2_BYTES = 2 * 2 * BN2 * 4;
constexpr int SMEM2_BYTES = STAGES2 * STAGE2_BYTES + 1024 + 256 + COLPAR2_BYTES;

__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
    for (uint32_t i = 0; !tc::mbar_try_wait(bar, parity); ++i)
        if (i > (1u << 27)) __trap();                 // a protocol error ends the kernel instead of hanging the GPU
}
__device__ __forceinline__ uint32_t pair_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void pair_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(tc::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(tc::smem_u32(leader_bar) & PEER_BIT_MASK), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {      // arrives on `bar` of BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(tc::smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(tc::smem_u32(bar) & PEER_BIT_MASK) : "memory");
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
                    const AvsrEpilogue ep) {
    constexpr int BN = BN2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES2 * STAGE2_BYTES);
    uint64_t* empty = full + STAGES2;
    uint64_t* tfull = empty + STAGES2;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* colpar = reinterpret_cast<float*>(smem + STAGES2 * STAGE2_BYTES + 256);        // [2 acc][bias BN | prelu BN]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = pair_rank();
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int tiles_n = (N + BN - 1) / BN;
    const int num_tiles = ((M + 2 * BM - 1) / (2 * BM)) * tiles_n;
    const int num_kb = (K + BK - 1) / BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES2; ++s) {
            tc::mbar_init(&full[s], 1);               // the leader's expect_tx arrival; bytes come from both CTAs
            tc::mbar_init(&empty[s], 1);              // the leader's commit, multicast
        }
        for (int a = 0; a < 2; ++a) {
            tc::mbar_init(&tfull[a], 1);
            tc::mbar_init(&tempty[a], 2 * NUM_EPI_WARPS);     // the epilogue warps of both CTAs (leader's copy is the one waited on)
        }
        tc::fence_barrier_init();
        tc::tma_prefetch_desc(&tmA);
        tc::tma_prefetch_desc(&tmB);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc::tc_fence_before();
    pair_sync();                                      // barriers of both CTAs initialised before anybody signals the peer
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = pair; tile < num_tiles; tile += npairs) {
                const int m0 = (tile / tiles_n) * 2 * BM + (int)rank * BM, n0 = (tile % tiles_n) * BN + (int)rank * (BN / 2);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait_bounded(&empty[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * STAGE2_BYTES;
                    if (rank == 0) tc::mbar_arrive_expect_tx(&full[stage], 2 * STAGE2_BYTES);
                    tma_load_2d_pair(sa, &tmA, &full[stage], kb * BK, m0);
                    tma_load_2d_pair(sa + A_TILE_BYTES, &tmB, &full[stage], kb * BK, n0);
                    if (++stage == STAGES2) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = tc::umma_idesc_bf16(2 * BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = pair; tile < num_tiles; tile += npairs) {
                mbar_wait_bounded(&tempty[acc], acc_phase ^ 1);
                tc::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait_bounded(&full[stage], phase);
                    tc::tc_fence_after();
                    const uint32_t sa = tc::smem_u32(smem + stage * STAGE2_BYTES);
                    const uint64_t adesc = tc::umma_desc_sw128(sa);
                    const uint64_t bdesc = tc::umma_desc_sw128(sa + A_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        umma_bf16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > 0 || k != 0) ? 1u : 0u);
                    umma_commit_pair(&empty[stage]);        // frees the slot in both CTAs when these MMAs retire
                    if (++stage == STAGES2) { stage = 0; phase ^= 1; }
                }
                umma_commit_pair(&tfull[acc]);              // accumulators of both CTAs complete -> both epilogues
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;
        const int etid = threadIdx.x - 64;
        const bool col_bias = ep.bias != nullptr && ep.bias_mode != 2;
        const bool has_prelu = ep.act == AVSR_ACT_PRELU && ep.prelu != nullptr;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = pair; tile < num_tiles; tile += npairs) {
            const int m0 = (tile / tiles_n) * 2 * BM + (int)rank * BM, n0 = (tile % tiles_n) * BN;
            const int row = m0 + quad * 32 + lane;
            float* cpar = colpar + acc * 2 * BN;
            if (col_bias || has_prelu) {
                for (int i = etid; i < BN; i += 32 * NUM_EPI_WARPS) {
                    const int col = n0 + i;
                    if (col_bias) cpar[i] = col < N ? __ldg(ep.bias + col) : 0.f;
                    if (has_prelu) cpar[BN + i] = col < N ? __ldg(ep.prelu + col) : 0.f;
                }
            }
            const float rbias = (ep.bias != nullptr && ep.bias_mode == 2 && row < M) ? __ldg(ep.bias + row) : 0.f;
            if (col_bias || has_prelu) asm volatile("bar.sync 1, %0;" ::"n"(32 * NUM_EPI_WARPS) : "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN;
            epilogue_tile<BN, true>(ep, taddr, m0, n0, M, N, quad, half, lane, cpar, rbias, &tfull[acc], acc_phase);
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tempty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc::tc_fence_before();
    pair_sync();                                      // nobody frees TMEM / leaves while the peer may still use this CTA's memory
    if (warp == 1) {
        tc::tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------------------------
// 3x3 / stride 1 / 64 -> 64 convolution with the input HALO staged once (ResNet layer1: four convolutions on 22 x 22 x 64
// maps, resnet.py:30-69).  The generic implicit GEMM above fetches one 128-pixel x 64-channel operand tile per filter tap:
// 9 x 16 KB of activations + 9 x 8 KB of weights per 128 x 64 output tile whose MMAs last 1152 clocks - 190 B/clk per SM
// against the ~43 B/clk the L2 delivers, so the tensor pipe idled 3/4 of the time (ncu r02: 25 % active, 292 us per launch).
// Here the activations live in a PADDED pixel layout (Wp = W + 2 columns, Hp = H + 1 rows per frame, pads are zero and are
// never written): the nine taps of an output pixel p are the pixels p + dy * Wp + dx of the same linear array, so a tile of
// 128 consecutive (padded) pixels needs ONE contiguous halo of 128 + 2 (Wp + 1) pixels = 23 KB, loaded by one TMA request,
// and the A operand of tap (dy, dx) is the SAME shared-memory tile read from row (Wp + 1) + dy * Wp + dx on: a UMMA descriptor
// whose start address is moved by whole 128-byte rows.  (The 128-byte swizzle is a function of the absolute shared-memory
// address bits, and every stage starts on a 1024-byte boundary as the TMA write assumed, so the matrix-base-offset field of
// the descriptor stays 0: setting it to the row phase was measured WRONG, leaving it 0 is bit-exact.)  The weights of all nine taps (72 KB) stay resident in shared memory for the whole persistent CTA.  L2 traffic per
// tile: 23 KB instead of 216 KB; 8 % of the rows computed are pad pixels whose results are dropped.
constexpr int HL_ROWS = 184;                          // halo rows per tile (128 + 2 * 25 = 178 used), a multiple of 8
constexpr int HL_BYTES = HL_ROWS * 128;               // 23 x 1024: every stage starts on a swizzle-pattern boundary
constexpr int HL_STAGES = 4;
constexpr int HL_W_BYTES = 9 * 64 * 128;              // 9 taps x [64 output channels][64 input channels] bf16
constexpr int HL_SMEM = HL_W_BYTES + HL_STAGES * HL_BYTES + 1024 + 256 + 2 * 2 * 64 * 4;

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const AvsrEpilogue ep, int W, int H,
                    long long total_px) {
    constexpr int BN = 64;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sW = smem;
    uint8_t* sH = smem + HL_W_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sH + HL_STAGES * HL_BYTES);
    uint64_t* empty = full + HL_STAGES;
    uint64_t* tfull = empty + HL_STAGES;
    uint64_t* tempty = tfull + 2;
    uint64_t* wfull = tempty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);
    float* colpar = reinterpret_cast<float*>(sH + HL_STAGES * HL_BYTES + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Wp = W + 2, Hp = H + 1, fpx = Wp * Hp;   // padded row / frame pitches in pixels
    const int num_tiles = (int)((total_px + BM - 1) / BM);

    if (threadIdx.x == 0) {
        for (int s = 0; s < HL_STAGES; ++s) {
            tc::mbar_init(&full[s], 1);
            tc::mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            tc::mbar_init(&tfull[a], 1);
            tc::mbar_init(&tempty[a], NUM_EPI_WARPS);
        }
        tc::mbar_init(wfull, 1);
        tc::fence_barrier_init();
        tc::tma_prefetch_desc(&tmX);
        tc::tma_prefetch_desc(&tmW);
    }
    if (warp == 1) {
        tc::tmem_alloc(tmem_slot, 2 * BN);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            tc::mbar_arrive_expect_tx(wfull, HL_W_BYTES);
            for (int tap = 0; tap < 9; ++tap) tc::tma_load_2d(sW + tap * 8192, &tmW, wfull, tap * 64, 0);
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait_bounded(&empty[stage], phase ^ 1);
                tc::mbar_arrive_expect_tx(&full[stage], HL_BYTES);
                // rows before the first / after the last pixel are zero-filled by the TMA unit (the pads of the first frame)
                tc::tma_load_2d(sH + stage * HL_BYTES, &tmX, &full[stage], 0, tile * BM - (Wp + 1));
                if (++stage == HL_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = tc::umma_idesc_bf16(BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            mbar_wait_bounded(wfull, 0);
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait_bounded(&tempty[acc], acc_phase ^ 1);
                mbar_wait_bounded(&full[stage], phase);
                tc::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                const uint32_t sh = tc::smem_u32(sH + stage * HL_BYTES);
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const int off = (Wp + 1) + (tap / 3 - 1) * Wp + (tap % 3 - 1);          // first halo row of this tap's operand
                    const uint64_t adesc = tc::umma_desc_sw128(sh + (uint32_t)off * 128u);
                    const uint64_t bdesc = tc::umma_desc_sw128(tc::smem_u32(sW + tap * 8192));
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) tc::umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (tap | k) != 0 ? 1u : 0u);
                }
                tc::umma_commit(&empty[stage]);
                tc::umma_commit(&tfull[acc]);
                if (++stage == HL_STAGES) { stage = 0; phase ^= 1; }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;
        const int etid = threadIdx.x - 64;
        const bool col_bias = ep.bias != nullptr && ep.bias_mode != 2;
        const bool has_prelu = ep.act == AVSR_ACT_PRELU && ep.prelu != nullptr;
        // bias / PReLU slopes of the 64 output channels are the same for every tile: staged once (both accumulator slots)
        for (int i = etid; i < 2 * BN; i += 32 * NUM_EPI_WARPS) {
            const int a = i / BN, col = i % BN;
            colpar[a * 2 * BN + col] = col_bias ? __ldg(ep.bias + col) : 0.f;
            colpar[a * 2 * BN + BN + col] = has_prelu ? __ldg(ep.prelu + col) : 0.f;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * NUM_EPI_WARPS) : "memory");
        int acc = 0;
        uint32_t acc_phase = 0;
        const bool has_res = ep.residual != nullptr;
        const int col0 = half * 32;                    // BN = 64: one 32-column chunk per warp
        // this thread's padded pixel of a tile, or -1 for a pad pixel: nothing is read or written there (the pads must stay
        // zero: they are the halo of their neighbours)
        auto pixel_of = [&](int tile) -> long long {
            const long long px = (long long)tile * BM + quad * 32 + lane;
            const int q = (int)(px % fpx);
            return (px < total_px && q / Wp < H && q % Wp < W) ? px : -1;
        };
        // the residual of tile i + 1 is requested before tile i is finished: its latency would otherwise be paid once per tile
        // (one warp finishes one tile at a time; measured 149 us per launch with the residual against 110 us without)
        float res[32], res_next[32];
        {
            const long long px = blockIdx.x < num_tiles ? pixel_of(blockIdx.x) : -1;
            if (has_res) load_residual(ep, px >= 0 ? (int)px : 0, col0, px >= 0 ? (int)total_px : 0, BN, res);
        }
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const long long px = pixel_of(tile);
            const int M_eff = px >= 0 ? (int)total_px : 0;
            const int row = px >= 0 ? (int)px : 0;
            const int nt = tile + gridDim.x;
            if (has_res && nt < num_tiles) {
                const long long pn = pixel_of(nt);
                load_residual(ep, pn >= 0 ? (int)pn : 0, col0, pn >= 0 ? (int)total_px : 0, BN, res_next);
            }
            mbar_wait_bounded(&tfull[acc], acc_phase);
            tc::tc_fence_after();
            uint32_t r[32];
            tc::tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN + col0, r);
            tc::tmem_ld_wait();
            float v[32];
            const float* cpar = colpar + acc * 2 * BN;
            epilogue_math(ep, row, M_eff, r, cpar + col0, cpar + BN + col0, 0.f, res, has_res, v);
            epilogue_store_direct(ep, row, col0, M_eff, BN, v);
            if (has_res) {
#pragma unroll
                for (int j = 0; j < 32; ++j) res[j] = res_next[j];
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&tempty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, 2 * BN);
    }
}

int g_sm_count = 0;
int g_pair_on = -1;                                  // dev knob, read once: AVSR_GEMM_PAIR=0 (no CTA-pair kernel)
bool pair_enabled() {
    if (g_pair_on < 0) {
        const char* e = getenv("AVSR_GEMM_PAIR");
        g_pair_on = (e == nullptr || atoi(e) != 0) ? 1 : 0;
    }
    return g_pair_on == 1;
}

template <int BN>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const AvsrEpilogue& ep, int splits, cudaStream_t stream,
           const ConvA& conv = ConvA{0, 0, 0, 0, 0, 0, 0, 0, nullptr, nullptr, nullptr, nullptr, nullptr}) {
    using C = Cfg<BN>;
    static bool configured = false;
    if (!configured) {
        AVSR_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        configured = true;
    }
    const int num_kb = cdiv(K, BK);
    const int kbps = cdiv(num_kb, splits);
    const int tiles = conv.pc_on ? conv.pc_nwork * 16 : cdiv(M, BM) * cdiv(N, BN) * splits;
    const int grid = tiles < g_sm_count ? tiles : g_sm_count;
    gemm_tc_kernel<BN><<<grid, NUM_THREADS, C::SMEM_BYTES, stream>>>(ta, tb, M, N, K, ep, splits, kbps, conv);
    AVSR_LAUNCH_CHECK();
    return AVSR_OK;
}

}  // namespace

namespace tc {
EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_rows,
                      uint32_t box_cols) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) {
        avsr_set_error("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
        return AVSR_ERR_CUDA;
    }
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || ((ld_elems * 2) & 15) != 0) {
        avsr_set_error("TMA operand must be 16-byte aligned with a 16-byte multiple row pitch (base=%p ld=%llu)", base,
                       (unsigned long long)ld_elems);
        return AVSR_ERR_ARG;
    }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstr[1] = {ld_elems * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        avsr_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu ld=%llu box=%ux%u)", (int)r,
                       (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems, box_rows, box_cols);
        return AVSR_ERR_CUDA;
    }
    return AVSR_OK;
}
}  // namespace tc

namespace tc {
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*,
                                   const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap_im2col_bf16(CUtensorMap* out, const void* base, uint64_t n, uint64_t h, uint64_t w, uint64_t c, int ks, int stride,
                          uint32_t pixels, uint32_t channels, uint64_t row_pitch_px, uint64_t frame_pitch_px) {
    static EncodeIm2colFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeIm2colFn>(p);
    }
    if (fn == nullptr) {
        avsr_set_error("cuTensorMapEncodeIm2col entry point not available (no CUDA driver?)");
        return AVSR_ERR_CUDA;
    }
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || ((c * 2) & 15) != 0) {
        avsr_set_error("im2col TMA operand must be 16-byte aligned with a 16-byte multiple pixel pitch (base=%p c=%llu)", base,
                       (unsigned long long)c);
        return AVSR_ERR_ARG;
    }
    cuuint64_t gdim[4] = {c, w, h, n};
    if (row_pitch_px == 0) row_pitch_px = w;          // dense image rows / frames unless the caller says otherwise
    if (frame_pitch_px == 0) frame_pitch_px = h * row_pitch_px;
    cuuint64_t gstr[3] = {c * 2, row_pitch_px * c * 2, frame_pitch_px * c * 2};
    const int pad = ks / 2;
    int lower[2] = {-pad, -pad};             // {W, H}: the filter's base pixel starts `pad` pixels outside the image ...
    int upper[2] = {-pad, -pad};             // ... and stops pad - (ks - 1) = -pad from the far edge
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};       // the base pixel advances by the conv stride
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, lower, upper, channels, pixels, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        avsr_set_error("cuTensorMapEncodeIm2col failed with CUresult %d (n=%llu h=%llu w=%llu c=%llu)", (int)r, (unsigned long long)n,
                       (unsigned long long)h, (unsigned long long)w, (unsigned long long)c);
        return AVSR_ERR_CUDA;
    }
    return AVSR_OK;
}
}  // namespace tc

// ks x ks (1 or 3) / stride 1 or 2 / pad ks/2 convolution of an NHWC bf16 tensor as an implicit GEMM on the tensor cores
// (BasicBlock and downsample convs, src/nets/backend/backbones/resnet.py:30-69): out[(f, y, x), :] = epilogue(sum_{ky,kx,c}
// in[f, y*s+ky-pad, x*s+kx-pad, c] * Wt[:, (ky*ks + kx)*C + c]).  in [F, H, W, C] bf16, Wt [Cout, ks*ks*C] bf16 (the layout the
// explicit im2col path uses), the epilogue's outputs are [F*Ho*Wo, Cout].  C % 64 == 0.  The patch matrix is never written.
// in_row_pitch_px / in_frame_pitch_px: pixel pitches of the input image rows / frames (0 = dense); a padded layout such as the
// one avsr_conv3x3_halo_bf16 works on is read in place.
extern "C" int avsr_conv2d_bf16_tc_pitched(const void* in, const void* Wt, long long F, int H, int W, int C, int Cout, int ks, int stride,
                                           long long in_row_pitch_px, long long in_frame_pitch_px, const AvsrEpilogue* ep,
                                           cudaStream_t stream) {
    AVSR_REQUIRE(in && Wt && ep && F > 0 && H > 0 && W > 0 && C > 0 && Cout > 0, "avsr_conv2d_bf16_tc: bad arguments");
    AVSR_REQUIRE((ks == 1 || ks == 3) && (stride == 1 || stride == 2), "avsr_conv2d_bf16_tc: ks %d / stride %d unsupported", ks, stride);
    AVSR_REQUIRE((C % BK) == 0, "avsr_conv2d_bf16_tc: C = %d must be a multiple of %d", C, BK);
    AVSR_REQUIRE(ep->out_bf16 || ep->out_f32, "avsr_conv2d_bf16_tc: no output buffer");
    const int pad = ks / 2;
    const int Ho = (H + 2 * pad - ks) / stride + 1, Wo = (W + 2 * pad - ks) / stride + 1;
    AVSR_REQUIRE(F * Ho * Wo < (1ll << 31), "avsr_conv2d_bf16_tc: too many output pixels");
    if (g_sm_count == 0) {
        int dev = 0;
        AVSR_CHECK_CUDA(cudaGetDevice(&dev));
        AVSR_CHECK_CUDA(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    const int M = (int)(F * Ho * Wo), K = ks * ks * C;
    const int bn = Cout <= 64 ? 64 : (Cout <= 128 ? 128 : ((cdiv(M, BM) * cdiv(Cout, 256) >= 2 * g_sm_count) ? 256 : 128));
    CUtensorMap ta, tb;
    AVSR_REQUIRE(in_row_pitch_px >= 0 && in_frame_pitch_px >= 0 && (in_row_pitch_px == 0 || in_row_pitch_px >= W), "avsr_conv2d_bf16_tc: bad input pitches");
    int rc = tc::make_tmap_im2col_bf16(&ta, in, (uint64_t)F, (uint64_t)H, (uint64_t)W, (uint64_t)C, ks, stride, BM, BK, (uint64_t)in_row_pitch_px,
                                       (uint64_t)in_frame_pitch_px);
    if (rc != AVSR_OK) return rc;
    rc = tc::make_tmap_2d_bf16(&tb, Wt, (uint64_t)Cout, (uint64_t)K, (uint64_t)K, (uint32_t)bn, BK);
    if (rc != AVSR_OK) return rc;
    const ConvA conv = {1, Ho, Wo, C, ks, stride, 0, 0, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (bn == 64) return launch<64>(ta, tb, M, Cout, K, *ep, 1, stream, conv);
    if (bn == 128) return launch<128>(ta, tb, M, Cout, K, *ep, 1, stream, conv);
    return launch<256>(ta, tb, M, Cout, K, *ep, 1, stream, conv);
}

extern "C" int avsr_conv2d_bf16_tc(const void* in, const void* Wt, long long F, int H, int W, int C, int Cout, int ks, int stride,
                                   const AvsrEpilogue* ep, cudaStream_t stream) {
    return avsr_conv2d_bf16_tc_pitched(in, Wt, F, H, W, C, Cout, ks, stride, 0, 0, ep, stream);
}

// 3x3 / stride 1 / pad 1 convolution, 64 -> 64 channels, on the PADDED layout: activations [F][H + 1][W + 2][64] bf16 whose pad
// cells (columns W, W + 1 of every row; row H of every frame) are zero.  ep.out_bf16 / ep.out_f32 / ep.residual use the same
// layout (row = padded pixel index, 64 columns); pad cells of the output are not written (allocate it zeroed once).
// Wt [64][9 * 64] bf16, k = (ky * 3 + kx) * 64 + cin (the layout avsr_conv2d_bf16_tc takes).
extern "C" int avsr_conv3x3_halo_bf16(const void* in, const void* Wt, long long F, int H, int W, const AvsrEpilogue* ep, cudaStream_t stream) {
    AVSR_REQUIRE(in && Wt && ep && F > 0 && H > 0 && W > 0, "avsr_conv3x3_halo_bf16: bad arguments");
    AVSR_REQUIRE(ep->out_bf16 || ep->out_f32, "avsr_conv3x3_halo_bf16: no output buffer");
    AVSR_REQUIRE(2 * (W + 3) + BM <= HL_ROWS, "avsr_conv3x3_halo_bf16: rows of %d pixels do not fit the halo tile", W);
    const long long total_px = F * (long long)(H + 1) * (W + 2);
    AVSR_REQUIRE(total_px < (1ll << 31) - HL_ROWS, "avsr_conv3x3_halo_bf16: too many pixels");
    if (g_sm_count == 0) {
        int dev = 0;
        AVSR_CHECK_CUDA(cudaGetDevice(&dev));
        AVSR_CHECK_CUDA(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    static bool configured = false;
    if (!configured) {
        AVSR_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HL_SMEM));
        configured = true;
    }
    CUtensorMap tx, tw;
    int rc = tc::make_tmap_2d_bf16(&tx, in, (uint64_t)total_px, 64, 64, HL_ROWS, 64);
    if (rc != AVSR_OK) return rc;
    rc = tc::make_tmap_2d_bf16(&tw, Wt, 64, 576, 576, 64, 64);
    if (rc != AVSR_OK) return rc;
    const long long tiles = cdiv(total_px, BM);
    const int grid = tiles < g_sm_count ? (int)tiles : g_sm_count;
    conv3x3_halo_kernel<<<grid, NUM_THREADS, HL_SMEM, stream>>>(tx, tw, *ep, W, H, total_px);
    AVSR_LAUNCH_CHECK();
    return AVSR_OK;
}

// The CTA-pair kernel takes plain GEMMs that fill the device with 256 x 256 tiles (the transformer's QKV / out / FFN
// projections at batch sizes worth the launch); AVSR_GEMM_PAIR=0 switches it off (dev knob).
static int launch_pair(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K, const AvsrEpilogue& ep, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        AVSR_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
        configured = true;
    }
    CUtensorMap ta, tb;
    int rc = tc::make_tmap_2d_bf16(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM, BK);
    if (rc != AVSR_OK) return rc;
    rc = tc::make_tmap_2d_bf16(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, BN2 / 2, BK);
    if (rc != AVSR_OK) return rc;
    const int tiles = cdiv(M, 2 * BM) * cdiv(N, BN2);
    const int pairs = tiles < g_sm_count / 2 ? tiles : g_sm_count / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = SMEM2_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    AVSR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_pair_kernel, ta, tb, M, N, K, ep));
    return AVSR_OK;
}

static int gemm_entry(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K, const AvsrEpilogue* ep,
                      int bn_hint, int splits, cudaStream_t stream) {
    AVSR_REQUIRE(A && B && ep, "avsr_gemm_bf16_tc: null operand");
    AVSR_REQUIRE(M > 0 && N > 0 && K > 0, "avsr_gemm_bf16_tc: bad shape %dx%dx%d", M, N, K);
    AVSR_REQUIRE(ep->out_bf16 || ep->out_f32, "avsr_gemm_bf16_tc: no output buffer");
    if (g_sm_count == 0) {
        int dev = 0;
        AVSR_CHECK_CUDA(cudaGetDevice(&dev));
        AVSR_CHECK_CUDA(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    int bn = bn_hint;
    if (bn == 0) {
        if (N <= 64) bn = 64;
        else if (N <= 128) bn = 128;
        else bn = (cdiv(M, BM) * cdiv(N, 256) * splits >= 2 * g_sm_count) ? 256 : 128;
    }
    AVSR_REQUIRE(bn == 64 || bn == 128 || bn == 256, "avsr_gemm_bf16_tc: bad bn_hint %d", bn_hint);
    if (bn == 256 && bn_hint == 0 && splits == 1 && pair_enabled() && cdiv(M, 2 * BM) * cdiv(N, BN2) >= g_sm_count)
        return launch_pair(A, lda, B, ldb, M, N, K, *ep, stream);
    CUtensorMap ta, tb;
    int rc = tc::make_tmap_2d_bf16(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM, BK);
    if (rc != AVSR_OK) return rc;
    rc = tc::make_tmap_2d_bf16(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, (uint32_t)bn, BK);
    if (rc != AVSR_OK) return rc;
    if (bn == 64) return launch<64>(ta, tb, M, N, K, *ep, splits, stream);
    if (bn == 128) return launch<128>(ta, tb, M, N, K, *ep, splits, stream);
    return launch<256>(ta, tb, M, N, K, *ep, splits, stream);
}

// C[M,N] = epilogue(A[M,K] * B[N,K]^T); A, B bf16 row-major with leading dimensions lda/ldb (elements, multiples of 8).
// bn_hint: 0 = choose the N tile automatically, else 64 / 128 / 256.
extern "C" int avsr_gemm_bf16_tc(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K,
                                 const AvsrEpilogue* ep, int bn_hint, cudaStream_t stream) {
    return gemm_entry(A, lda, B, ldb, M, N, K, ep, bn_hint, 1, stream);
}

// Split-K form for skinny operands (decode step): part[z][M][N] (fp32) = A[:, Kz] * B[:, Kz]^T for z < splits.
extern "C" int avsr_gemm_bf16_tc_splitk(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K, float* part,
                                        int splits, int bn_hint, cudaStream_t stream) {
    AVSR_REQUIRE(part && splits >= 1 && splits <= cdiv(K, BK), "avsr_gemm_bf16_tc_splitk: bad splits %d", splits);
    AvsrEpilogue ep = {};
    ep.out_f32 = part;
    ep.ld_f32 = N;
    // every split must own at least one k-block, otherwise its partial would stay unwritten
    const int num_kb = cdiv(K, BK), kbps = cdiv(num_kb, splits);
    AVSR_REQUIRE((splits - 1) * kbps < num_kb, "avsr_gemm_bf16_tc_splitk: %d splits leave an empty split for K=%d", splits, K);
    return gemm_entry(A, lda, B, ldb, M, N, K, &ep, bn_hint, splits, stream);
}

// ---- positional convolution as an implicit banded GEMM -----------------------------------------------------------------
// Host helper (no GPU work): one tensor map per utterance over its frames of the packed bf16 activations x [F, 1024]
// (utterance b = rows utt_off[b] .. + utt_T[b]), written to maps_host [B][128 bytes]; the caller uploads them.  Rows a load
// asks for outside [0, utt_T[b]) come back as zeros: the convolution's zero padding at the UTTERANCE boundary.
extern "C" int avsr_posconv_encode_maps(const void* x, const long long* utt_off, const int* utt_T, int B, void* maps_host) {
    AVSR_REQUIRE(x && utt_off && utt_T && maps_host && B > 0, "avsr_posconv_encode_maps: bad arguments");
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap is 128 bytes");
    CUtensorMap* out = reinterpret_cast<CUtensorMap*>(maps_host);
    for (int b = 0; b < B; ++b) {
        AVSR_REQUIRE(utt_T[b] > 0 && utt_off[b] >= 0, "avsr_posconv_encode_maps: utterance %d has no frames", b);
        CUtensorMap m;
        int rc = tc::make_tmap_2d_bf16(&m, reinterpret_cast<const __nv_bfloat16*>(x) + utt_off[b] * 1024, (uint64_t)utt_T[b], 1024, 1024, BM, BK);
        if (rc != AVSR_OK) return rc;
        memcpy(out + b, &m, sizeof(m));
    }
    return AVSR_OK;
}

// x + GELU(Conv1d_{k=128, pad=64, groups=16}(x)[..., :-1] + bias) of HF Wav2Vec2PositionalConvEmbedding (modeling_wav2vec2.py:326-379,
// called from src/nets/backend/backbones/avhubert.py:698-699) as ONE launch over all 16 groups: out tile (128 frames of an
// utterance x the 64 channels of a group) = sum over the 128 taps of A_j W_j^T with A_j = frames [q0 + j - 64, +128) of the
// utterance (TMA load through utt_maps[utterance], zero outside it).  The patch matrix (128 x the input) is never written.
// W = [1024][8192] bf16 with k = tap * 64 + channel-in-group (the explicit path's layout); work arrays as avsr_attention_varlen
// plus work_utt (utterance index); ep = bias [1024] per column, GELU, residual / output rows indexed by PACKED frame.
extern "C" int avsr_posconv_bf16_tc(const void* utt_maps, const void* W, int n_work, const int* work_utt, const int* work_off, const int* work_T,
                                    const int* work_q0, const AvsrEpilogue* ep, cudaStream_t stream) {
    AVSR_REQUIRE(utt_maps && W && work_utt && work_off && work_T && work_q0 && ep && n_work > 0, "avsr_posconv_bf16_tc: bad arguments");
    AVSR_REQUIRE(((uintptr_t)utt_maps & 63) == 0, "avsr_posconv_bf16_tc: tensor maps must be 64-byte aligned");
    AVSR_REQUIRE(ep->out_bf16 || ep->out_f32, "avsr_posconv_bf16_tc: no output buffer");
    if (g_sm_count == 0) {
        int dev = 0;
        AVSR_CHECK_CUDA(cudaGetDevice(&dev));
        AVSR_CHECK_CUDA(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    CUtensorMap tb;
    int rc = tc::make_tmap_2d_bf16(&tb, W, 1024, 8192, 8192, 64, BK);
    if (rc != AVSR_OK) return rc;
    const ConvA conv = {0, 0, 0, 0, 0, 0, 1, n_work, reinterpret_cast<const CUtensorMap*>(utt_maps), work_utt, work_off, work_T, work_q0};
    return launch<64>(tb, tb, n_work * 128, 1024, 8192, *ep, 1, stream, conv);
}
