// Encoder self-attention on tcgen05 tensor cores, variable-length (packed utterances), non-causal, 16 heads x 64.
// Reference: HF Wav2Vec2Attention / eager_attention_forward (transformers/models/wav2vec2/modeling_wav2vec2.py:438-549)
// called from src/nets/backend/backbones/avhubert.py:751-753: softmax(q k^T / 8) v per head, no mask inside an utterance.
//
// One CTA = one (utterance, head, 128-query tile).  Exact two-pass softmax over 128-key blocks:
//   pass A: S = Q K_j^T (tcgen05.mma, fp32 in TMEM) -> row max
//   pass B: S = Q K_j^T again, P = exp(S - max) -> bf16 into swizzled smem, O += P V_j (tcgen05.mma, O in TMEM)
// Q (pre-scaled by 1/8 at weight-pack time) and all K blocks stay resident in shared memory; V^T blocks are TMA loads into
// one or two buffers.  The CTA is one dependent chain (load -> MMA -> TMEM read -> softmax -> MMA ...), so the kernel is
// latency-bound per CTA: the shared-memory layout is sized by the longest utterance of the batch (kcap key blocks) and, up to
// 384 keys, two CTAs share an SM (one V buffer, 112 KB each) so that one's softmax overlaps the other's MMAs and loads.  Inputs: qk [F, 2048] bf16 (q | k), vt [1024, F] bf16 (V transposed: d-major rows),
// output [F, 1024] bf16.  Thread i owns query row i (TMEM lane i).
#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int QT = 128;          // queries per CTA
constexpr int KB = 128;          // keys per block
constexpr int MAX_KB = 6;        // T <= 768 frames (30.7 s)
constexpr int TILE16K = 16384;
constexpr int ATT_TWO_PER_SM = 7 * TILE16K + 256;      // largest footprint that still lets two CTAs share an SM
__host__ __device__ constexpr int att_smem(int kcap, int nvb) { return (1 /*Q*/ + kcap /*K*/ + nvb /*V*/ + 2 /*P*/) * TILE16K + 256; }

__global__ void __launch_bounds__(128, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmV, const int* __restrict__ work_off,
               const int* __restrict__ work_T, const int* __restrict__ work_q0, __nv_bfloat16* __restrict__ out, int kcap, int nvb) {
    extern __shared__ __align__(1024) uint8_t smem[];          // the 128-byte swizzle needs 1024-byte aligned tiles
    if ((reinterpret_cast<uintptr_t>(smem) & 1023) != 0) __trap();
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + TILE16K;
    uint8_t* sV = sK + kcap * TILE16K;
    uint8_t* sP = sV + nvb * TILE16K;
    uint64_t* bar_qk = reinterpret_cast<uint64_t*>(sP + 2 * TILE16K);
    uint64_t* bar_v = bar_qk + 1;      // [2]
    uint64_t* bar_s = bar_v + 2;       // S ready
    uint64_t* bar_o = bar_s + 1;       // P*V retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_o + 1);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int off = work_off[blockIdx.x], T = work_T[blockIdx.x], q0 = work_q0[blockIdx.x];
    const int h = blockIdx.y;
    // TMA needs a 16-byte aligned global address for the innermost coordinate (tokens, in V^T): start the key range at
    // the previous multiple of 8 tokens and mask the `kshift` leading keys, which belong to the previous utterance.
    const int kshift = off & 7;
    const int koff = off - kshift;
    const int nk = T + kshift;
    const int nkb = (nk + KB - 1) / KB;

    if (tid == 0) {
        tc::mbar_init(bar_qk, 1);
        tc::mbar_init(&bar_v[0], 1);
        tc::mbar_init(&bar_v[1], 1);
        tc::mbar_init(bar_s, 1);
        tc::mbar_init(bar_o, 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) {
        tc::tmem_alloc(tmem_slot, 256);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tS = tmem_base, tO = tmem_base + 128;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;

    if (tid == 0) {
        tc::mbar_arrive_expect_tx(bar_qk, (1 + nkb) * TILE16K);
        tc::tma_load_2d(sQ, &tmQK, bar_qk, h * 64, off + q0);
        for (int kb = 0; kb < nkb; ++kb) tc::tma_load_2d(sK + kb * TILE16K, &tmQK, bar_qk, 1024 + h * 64, koff + kb * KB);
        tc::mbar_arrive_expect_tx(&bar_v[0], TILE16K);
        tc::tma_load_2d(sV, &tmV, &bar_v[0], koff, h * 64);
        tc::tma_load_2d(sV + 8192, &tmV, &bar_v[0], koff + 64, h * 64);
    }
    tc::mbar_wait(bar_qk, 0);

    constexpr uint32_t idesc_s = tc::umma_idesc_bf16(128, 128);
    constexpr uint32_t idesc_o = tc::umma_idesc_bf16(128, 64);
    const uint64_t dq = tc::umma_desc_sw128(tc::smem_u32(sQ));
    uint32_t ph_s = 0, ph_o = 0;
    constexpr float L2E = 1.4426950408889634f;

    // ---------------- pass A: row max
    float m = -INFINITY;
    for (int kb = 0; kb < nkb; ++kb) {
        if (tid == 0) {
            tc::tc_fence_after();
            const uint64_t dk = tc::umma_desc_sw128(tc::smem_u32(sK + kb * TILE16K));
#pragma unroll
            for (int k = 0; k < 4; ++k) tc::umma_bf16(tS, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
            tc::umma_commit(bar_s);
        }
        tc::mbar_wait(bar_s, ph_s);
        ph_s ^= 1;
        tc::tc_fence_after();
        const int lo = kshift - kb * KB, hi = nk - kb * KB;       // valid keys of this block: lo <= column < hi
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            uint32_t r[32];
            tc::tmem_ld_32x32(tS + lane_addr + c * 32, r);
            tc::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (c * 32 + j >= lo && c * 32 + j < hi) m = fmaxf(m, __uint_as_float(r[j]));
        }
        tc::tc_fence_before();
        __syncthreads();
    }
    const float mL = m * L2E;

    // ---------------- pass B: P = exp(S - max), O += P V
    float sum = 0.f;
    for (int kb = 0; kb < nkb; ++kb) {
        if (tid == 0) {
            tc::tc_fence_after();
            const uint64_t dk = tc::umma_desc_sw128(tc::smem_u32(sK + kb * TILE16K));
#pragma unroll
            for (int k = 0; k < 4; ++k) tc::umma_bf16(tS, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
            tc::umma_commit(bar_s);
        }
        if (kb > 0) {                       // previous P*V must have retired before sP / the other V buffer are reused
            tc::mbar_wait(bar_o, ph_o);
            ph_o ^= 1;
        }
        // V^T of block kb + nvb - 1: into the other buffer one block ahead (two buffers), or into the single buffer now that the
        // product that read it has retired (it then lands while this block's scores and exponentials are computed)
        const int nxt = kb + nvb - 1;
        if (tid == 0 && nxt > 0 && nxt < nkb) {
            uint8_t* dst = sV + (nxt % nvb) * TILE16K;
            tc::mbar_arrive_expect_tx(&bar_v[nxt % nvb], TILE16K);
            tc::tma_load_2d(dst, &tmV, &bar_v[nxt % nvb], koff + nxt * KB, h * 64);
            tc::tma_load_2d(dst + 8192, &tmV, &bar_v[nxt % nvb], koff + nxt * KB + 64, h * 64);
        }
        tc::mbar_wait(bar_s, ph_s);
        ph_s ^= 1;
        tc::tc_fence_after();
        const int lo = kshift - kb * KB, hi = nk - kb * KB;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            uint32_t r[32];
            tc::tmem_ld_32x32(tS + lane_addr + c * 32, r);
            tc::tmem_ld_wait();
            uint8_t* prow = sP + (c >> 1) * TILE16K + tid * 128;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint32_t pk[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = g * 8 + e * 2;
                    float p0 = (c * 32 + j >= lo && c * 32 + j < hi) ? exp2f(__uint_as_float(r[j]) * L2E - mL) : 0.f;
                    float p1 = (c * 32 + j + 1 >= lo && c * 32 + j + 1 < hi) ? exp2f(__uint_as_float(r[j + 1]) * L2E - mL) : 0.f;
                    __nv_bfloat162 b = __floats2bfloat162_rn(p0, p1);
                    // the row sum uses the bf16-rounded probabilities that the tensor core will actually multiply
                    sum += __bfloat162float(b.x) + __bfloat162float(b.y);
                    pk[e] = *reinterpret_cast<uint32_t*>(&b);
                }
                const int gran = ((c & 1) * 4 + g) ^ (tid & 7);
                *reinterpret_cast<uint4*>(prow + gran * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
        }
        tc::fence_proxy_async();            // make the generic-proxy P writes visible to the tensor core (async proxy)
        tc::tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc::tc_fence_after();
            tc::mbar_wait(&bar_v[kb % nvb], (kb / nvb) & 1);
            const uint32_t pv = tc::smem_u32(sV + (kb % nvb) * TILE16K);
            const uint32_t pp = tc::smem_u32(sP);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint64_t da = tc::umma_desc_sw128(pp + (k >> 2) * TILE16K) + 2 * (k & 3);
                const uint64_t db = tc::umma_desc_sw128(pv + (k >> 2) * 8192) + 2 * (k & 3);
                tc::umma_bf16(tO, da, db, idesc_o, (kb | k) != 0);
            }
            tc::umma_commit(bar_o);
        }
    }
    tc::mbar_wait(bar_o, ph_o);
    tc::tc_fence_after();

    // ---------------- epilogue: O / sum -> bf16
    const float inv = 1.f / sum;
    const int q = q0 + tid;
    __nv_bfloat16* orow = out + (long long)(off + q) * 1024 + h * 64;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tc::tmem_ld_32x32(tO + lane_addr + c * 32, r);
        tc::tmem_ld_wait();
        if (q < T) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint32_t pk[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    __nv_bfloat162 b = __floats2bfloat162_rn(__uint_as_float(r[g * 8 + e * 2]) * inv, __uint_as_float(r[g * 8 + e * 2 + 1]) * inv);
                    pk[e] = *reinterpret_cast<uint32_t*>(&b);
                }
                *reinterpret_cast<uint4*>(orow + c * 32 + g * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, 256);
    }
}

}  // namespace

// qk [F, 2048] bf16 (ld 2048), vt [1024, ld_vt] bf16 with F valid columns, out [F, 1024] bf16.
// work_* are device arrays of n_work (utterance frame offset, utterance length, first query row of the tile).
extern "C" int avsr_attention_varlen(const void* qk, const void* vt, long long ld_vt, void* out, long long F, const int* work_off,
                                     const int* work_T, const int* work_q0, int n_work, int max_T, cudaStream_t stream) {
    AVSR_REQUIRE(qk && vt && out && work_off && work_T && work_q0 && n_work > 0 && F > 0, "avsr_attention_varlen: bad arguments");
    AVSR_REQUIRE(max_T + 7 <= MAX_KB * KB, "avsr_attention_varlen: utterance of %d frames exceeds the supported %d", max_T, MAX_KB * KB - 7);
    static bool configured = false;
    if (!configured) {
        AVSR_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, att_smem(MAX_KB, 2)));
        configured = true;
    }
    // key blocks the longest utterance needs (7 = worst-case alignment shift of the first key); two V buffers when they still
    // leave room for a second CTA on the SM (or when one CTA per SM is all that fits anyway), else one
    const int kcap = (max_T + 7 + KB - 1) / KB;
    const int nvb = (att_smem(kcap, 2) <= ATT_TWO_PER_SM || att_smem(kcap, 1) > ATT_TWO_PER_SM) ? 2 : 1;
    const int smem = att_smem(kcap, nvb);
    CUtensorMap tq, tv;
    int rc = tc::make_tmap_2d_bf16(&tq, qk, (uint64_t)F, 2048, 2048, 128, 64);
    if (rc != AVSR_OK) return rc;
    rc = tc::make_tmap_2d_bf16(&tv, vt, 1024, (uint64_t)F, (uint64_t)ld_vt, 64, 64);
    if (rc != AVSR_OK) return rc;
    attn_tc_kernel<<<dim3(n_work, 16), 128, smem, stream>>>(tq, tv, work_off, work_T, work_q0, (__nv_bfloat16*)out, kcap, nvb);
    AVSR_LAUNCH_CHECK();
    return AVSR_OK;
}
