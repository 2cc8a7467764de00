// Single-query attention of one decode position (MultiHeadedAttention.forward for a query of length 1,
// src/nets/backend/transformer/attention.py:38-106, called from DecoderLayer.forward, decoder_layer.py:82-107), KV-cache
// form: the reference re-projects K and V of every cached token / source frame at every step (decoder.py:153-183).
//
// HBM-streaming design:
//   * CTA (128 threads) = (utterance, head): the whole grid (512 CTAs for 32 utterances) is resident at once, so there are no
//     waves of CTAs that each pay the full dependent chain (wait -> query -> loads -> softmax -> merge), which is what bounds
//     a launch this short.  All live hyps of the utterance are served by the same CTA, so a K / V row shared by several hyps
//     is read from HBM once.  The keys are processed in tiles of 128; the K registers of tile i+1 are requested as soon as
//     the scores of tile i are done and the V tile of i+1 as soon as the P.V product of tile i is done.
//   * K is stored TRANSPOSED in 32-byte groups, K^T[j][row][8] (j = dim / 8): thread = key, its loads are coalesced across
//     the warp (whole 32-byte sectors per row, also when only every beam-th row of the self-attention cache is live) and put
//     the 64-float key in registers; the dot products with the (few) queries are plain FMAs against broadcast
//     shared-memory reads: no shuffles, no redundant work.  V stays row-major [row][64] and goes to shared memory with
//     cp.async (no registers held): a half-warp owns 16 consecutive keys, lane = 4 output dims, and reads back exactly the
//     bytes it copied.  K tile and V tile (64 KB per CTA) are requested before anything is computed.
//   * mode 1 (source attention): rows = the utterance's frames.  The loads are issued BEFORE griddepcontrol.wait (the cross
//     K/V were written before the chain of step kernels started), i.e. while the query projection is still running.
//   * mode 0 (self-attention): a hyp finds its history through the ancestry table anc[row][pos] = slot; the CTA builds the
//     list of DISTINCT (pos, slot) rows its live hyps reference, each with the bit mask of the hyps that use it (beams
//     mostly share their ancestors), and streams exactly those rows, 128 at a time.  k / v of the current position are
//     appended to the cache first.
//   * softmax statistics per tile (exact max, then exponentials), running (max, sum, sum e*v) across the tiles in tile
//     order: the result never depends on scheduling.
//   * The query (and the current k, v) can be taken straight from the split-K partial sums of the projection that produced
//     them (sum over splits in a fixed order + bias), which saves one epilogue launch per attention.
//   * The body is instantiated for the exact number of live hyps: no per-hyp guards in the inner loops.
#include <stdlib.h>
#include "common.cuh"

namespace {

constexpr int D = 1024;
constexpr int HEADS = 16;
constexpr int DH = 64;
constexpr int CK = 128;                  // keys per tile = threads per CTA
constexpr int NHW = 8;                   // half-warps per CTA
constexpr int VR = CK / NHW;             // V rows per half-warp and tile (16 consecutive keys)
constexpr int KG = 8;                    // floats per key group (32 bytes = one sector)

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Dev aid (avsr_dec_attn_debug): phase timestamps of CTA (0, 0) of the self-attention kernel, globaltimer nanoseconds.
__device__ unsigned long long* g_attn_dbg = nullptr;
#ifndef AVSR_ATTN_DEBUG
#define AVSR_ATTN_DEBUG 0
#endif
__device__ __forceinline__ void dbg_stamp(int k) {
    if (AVSR_ATTN_DEBUG && g_attn_dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_attn_dbg[k] = t;
    }
}

struct AttnArgs {
    const float* q_in; long long ldq; int nsplit; const float* q_bias;
    float* kc; float* vc; const unsigned char* anc; int lmax;
    const float* kd; const float* vd; const int* conv_len;      // mode 0: dense copy of the converged history prefix (may be null)
    const int* n_run; const int* utt_off; const int* utt_T; int beam; int R; const int* step_p;
    float* out; long long n_frames; __nv_bfloat16* out_split;
    const char* pf; long long pf_bytes;      // span the next kernel of the chain streams (its weights): fetched into L2 from here
    int kv_ahead;                            // fetch this CTA's remaining K / V into L2 before griddepcontrol.wait (dev knob)
    // source attention whose query projection was merged into the previous projections ("query merge"): q_in holds
    // x (gamma . Wq)^T of the layer-normed row WITHOUT the LayerNorm; the kernel finishes it, q = rstd (q_in - mean u) + c, from
    // the row's tile statistics q_stats [8][R][2] (mean, M2 per 128 columns, as avsr_dec_proj writes them)
    const float* q_stats; const float* q_u; const float* q_c; float q_eps;
};

template <int NH>
struct AttnSmem {
    float qs[NH][DH];                    // finished queries
    float qpart[NH][DH];                 // second half of the split-K sums
    float sc[NH][CK];                    // exponentials of the current tile
    float s_o[NHW][NH][DH];
    float s_redm[4][NH], s_reds[4][NH];
    float s_run[2][NH];                  // running max / sum over the tiles of this CTA
    int s_wcnt[4];
    int s_nrows;
};

// K of one key: 16 x 16-byte loads, two per 32-byte group.  MODE 0 uses plain loads (this CTA may just have written the row).
template <int MODE>
__device__ __forceinline__ void load_k(float4 (&kreg)[DH / 4], const float* kbase, long long nr, long long r, bool ok) {   // nr = rows per plane
#pragma unroll
    for (int j = 0; j < DH / KG; ++j) {
        kreg[2 * j] = make_float4(0.f, 0.f, 0.f, 0.f);
        kreg[2 * j + 1] = kreg[2 * j];
        if (ok) {
            const float4* p = reinterpret_cast<const float4*>(kbase + ((long long)j * nr + r) * KG);
            kreg[2 * j] = (MODE == 1) ? __ldg(p) : p[0];
            kreg[2 * j + 1] = (MODE == 1) ? __ldg(p + 1) : p[1];
        }
    }
}

// Everything after griddepcontrol.wait, for exactly NHT live hyps (NH = hyp slots of the beam).
template <int MODE, int NH, int NHT, bool SPLITQ>
__device__ __forceinline__ void attn_body(const AttnArgs& a, AttnSmem<NH>& sm, unsigned* rlist, float* vtile, float4 (&kreg)[DH / 4],
                                          const float* kbase, const float* vbase, long long nr, int T_utt, int step, int conv,
                                          const float4 (&pre)[3]) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int hw = tid >> 4, l16 = tid & 15;
    const int utt = blockIdx.x, head = blockIdx.y;
    const int beam = a.beam, R = a.R, nsplit = SPLITQ ? a.nsplit : 0;       // SPLITQ = false: the instantiation without the split-K gather code
    const int row0 = utt * beam;
    const int n = (MODE == 1) ? T_utt : step + 1;    // keys (mode 1) / positions (mode 0)
    const float* q_in = a.q_in;
    const long long ldq = a.ldq;
    const uint32_t vt_s = (uint32_t)__cvta_generic_to_shared(vtile) + (uint32_t)((hw * VR) * DH + 4 * l16) * 4u;   // this thread's V slots

    // Positions [0, C) of the history are "converged": every live hyp has the same ancestor there, and that row has been
    // copied to the dense caches kd / vd (avsr_dec_cache_promote), where consecutive positions are consecutive rows.  Only
    // the positions from C on go through the (pos, slot) list.
    int C = 0;
    const float* kdb = nullptr;
    const float* vdb = nullptr;
    if (MODE == 0 && a.conv_len != nullptr) {
        C = conv;
        kdb = a.kd + (long long)(utt * HEADS + head) * a.lmax * DH;
        vdb = a.vd + (long long)(utt * HEADS + head) * a.lmax * DH;
    }
    // A first tile that lies entirely in the dense prefix needs nothing but C: its K / V are requested now, before the query
    // is gathered and the row list is built, so that their latency overlaps both.
    bool early = false;
    if (MODE == 0 && C >= CK) {
        early = true;
        load_k<0>(kreg, kdb, a.lmax, tid, true);
#pragma unroll
        for (int i = 0; i < VR; ++i) cp_async16(vt_s + i * DH * 4, vdb + (long long)(hw * VR + i) * DH + 4 * l16);
    }

    if (MODE == 0) dbg_stamp(2);
    // ---- query: 16-byte groups, all split-K terms of a group requested at once (two threads share a group when there are
    //      few hyps); the current k / v of the chunk that owns this position go straight to the cache
    const long long zstride = (long long)R * ldq;
    auto gather4 = [&](int row, int col, int z0, int z1) -> float4 {       // sum over splits [z0, z1) in split order
        const float* p = q_in + (long long)row * ldq + col;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        int z = z0;
        for (; z + 8 <= z1; z += 8) {
            float4 t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) t[u] = __ldcg(reinterpret_cast<const float4*>(p + (z + u) * zstride));
#pragma unroll
            for (int u = 0; u < 8; ++u) { v.x += t[u].x; v.y += t[u].y; v.z += t[u].z; v.w += t[u].w; }
        }
        for (; z < z1; ++z) {
            const float4 t = __ldcg(reinterpret_cast<const float4*>(p + z * zstride));
            v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
        }
        return v;
    };
    constexpr int G = NHT * (DH / 4);                // 16-byte groups of the query block
    const bool two = nsplit > 1 && 2 * G <= CK;      // two threads per group: splits [0, zh) and [zh, nsplit)
    const int zh = two ? (nsplit + 1) / 2 : nsplit;
    if (nsplit <= 0) {
        // finished projections: q (and k | v) of the NH hypothesis slots were requested together with the beam state, right
        // after griddepcontrol.wait (thread = (slot, 16-byte group)); one round trip instead of two
        if (tid < G) *reinterpret_cast<float4*>(&sm.qs[tid / (DH / 4)][4 * (tid % (DH / 4))]) = pre[0];
    } else {
        for (int g = tid; g < (two ? 2 * G : G); g += CK) {
            const int sub = g / G, gg = g % G;
            const int h = gg / (DH / 4), j = gg % (DH / 4);
            const float4 v = gather4(row0 + h, head * DH + 4 * j, sub == 0 ? 0 : zh, sub == 0 ? zh : nsplit);
            *reinterpret_cast<float4*>(sub == 0 ? &sm.qs[h][4 * j] : &sm.qpart[h][4 * j]) = v;
        }
    }
    if (MODE == 0 && nsplit <= 0) {
        if (tid < G) {
            const int h = tid / (DH / 4), j = tid % (DH / 4);
            const long long rr = (long long)step * beam + h;
            const long long cb = (long long)(utt * HEADS + head) * nr * DH;
            *reinterpret_cast<float4*>(a.kc + cb + ((long long)(j >> 1) * nr + rr) * KG + (j & 1) * 4) = pre[1];
            *reinterpret_cast<float4*>(a.vc + cb + rr * DH + 4 * j) = pre[2];
        }
    } else if (MODE == 0) {
        for (int g = tid; g < 2 * G; g += CK) {
            const int kv = g / G, gg = g % G;
            const int h = gg / (DH / 4), j = gg % (DH / 4);
            const int col = (1 + kv) * D + head * DH + 4 * j;
            float4 v;
            if (nsplit <= 0) v = *reinterpret_cast<const float4*>(q_in + (long long)(row0 + h) * ldq + col);
            else {
                v = gather4(row0 + h, col, 0, nsplit);
                const float4 b4 = *reinterpret_cast<const float4*>(a.q_bias + col);
                v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
            }
            const long long rr = (long long)step * beam + h;
            const long long cb = (long long)(utt * HEADS + head) * nr * DH;
            if (kv == 0) *reinterpret_cast<float4*>(a.kc + cb + ((long long)(j >> 1) * nr + rr) * KG + (j & 1) * 4) = v;
            else *reinterpret_cast<float4*>(a.vc + cb + rr * DH + 4 * j) = v;
        }
    }
    if (MODE == 0) dbg_stamp(3);
    // ---- mode 0: list of the distinct (pos, slot) rows referenced by the live hyps, in (pos, slot) order
    int nrows = (MODE == 1) ? T_utt : 0;
    if (MODE == 0) {
        int total = 0;
        for (int pb = C; pb < n; pb += CK) {
            const int p = pb + tid;
            unsigned msk[NH];
#pragma unroll
            for (int s = 0; s < NH; ++s) msk[s] = 0u;
            if (p < n) {
#pragma unroll
                for (int h = 0; h < NHT; ++h) {
                    const int slot = (p < step) ? (int)a.anc[((long long)(step & 1) * R + row0 + h) * a.lmax + p] : h;
#pragma unroll
                    for (int s = 0; s < NH; ++s) msk[s] |= (slot == s) ? (1u << h) : 0u;
                }
            }
            int cnt = 0;
#pragma unroll
            for (int s = 0; s < NH; ++s) cnt += msk[s] != 0u;
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            __syncthreads();                         // s_wcnt of the previous round has been read; the appended k / v are visible
            if (lane == 31) sm.s_wcnt[warp] = incl;
            __syncthreads();
            int base = total + incl - cnt;
            for (int w = 0; w < warp; ++w) base += sm.s_wcnt[w];
            total += sm.s_wcnt[0] + sm.s_wcnt[1] + sm.s_wcnt[2] + sm.s_wcnt[3];
#pragma unroll
            for (int s = 0; s < NH; ++s)
                if (msk[s] != 0u) rlist[base++] = ((unsigned)(p * beam + s) << 8) | msk[s];
        }
        nrows = C + total;
    }
    if (tid < NHT) { sm.s_run[0][tid] = -INFINITY; sm.s_run[1][tid] = 0.f; }
    __syncthreads();                                 // qs / qpart, rlist, s_run ready
    if (MODE == 0) dbg_stamp(4);
    if (nsplit > 0) {                                // finish the query: (first half + second half) + bias
        for (int g = tid; g < G; g += CK) {
            const int h = g / (DH / 4), j = g % (DH / 4);
            float4 v = *reinterpret_cast<const float4*>(&sm.qs[h][4 * j]);
            if (two) {
                const float4 w = *reinterpret_cast<const float4*>(&sm.qpart[h][4 * j]);
                v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
            }
            const float4 b4 = *reinterpret_cast<const float4*>(a.q_bias + head * DH + 4 * j);
            v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
            *reinterpret_cast<float4*>(&sm.qs[h][4 * j]) = v;
        }
    }
    float acc[NHT][4];
#pragma unroll
    for (int h = 0; h < NHT; ++h) acc[h][0] = acc[h][1] = acc[h][2] = acc[h][3] = 0.f;

    const int ntiles = (nrows + CK - 1) / CK;
    // requests of one tile: K of this thread's key into registers, the half-warp's 16 V rows into shared memory
    auto request_k = [&](int t0) {
        const int g = t0 + tid;
        const bool ok = g < nrows;
        if (MODE == 0) {
            if (g < C) load_k<0>(kreg, kdb, a.lmax, g, true);
            else load_k<0>(kreg, kbase, nr, ok ? (long long)(rlist[g - C] >> 8) : 0, ok);
        } else {
            load_k<1>(kreg, kbase, nr, g, ok);
        }
    };
    auto request_v = [&](int t0) {
        const int g0 = t0 + hw * VR;
        const int nv = nrows - g0;
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < VR; ++i) {
                if (i < nv) {
                    const int g = g0 + i;
                    const float* src = (g < C) ? vdb + (long long)g * DH : vbase + (long long)(rlist[g - C] >> 8) * DH;
                    cp_async16(vt_s + i * DH * 4, src + 4 * l16);
                }
            }
        } else {
            const float* vsrc = vbase + (long long)g0 * DH + 4 * l16;
#pragma unroll
            for (int i = 0; i < VR; ++i)
                if (i < nv) cp_async16(vt_s + i * DH * 4, vsrc + i * DH);
        }
    };
    if (MODE == 0 && !early) { request_k(0); request_v(0); }  // mode 1: tile 0 was requested before griddepcontrol.wait
    __syncthreads();                                 // finished query visible (the loads above are already in flight)
    if (MODE == 0) dbg_stamp(5);
    for (int tile = 0; tile < ntiles; ++tile) {
        const int t0 = tile * CK;
        const bool kvalid = t0 + tid < nrows;
        const unsigned my = (MODE == 0 && t0 + tid >= C) ? (kvalid ? rlist[t0 + tid - C] : 0u) : 0xffu;     // hyp mask of this thread's key
        const int nv = min(VR, nrows - t0 - hw * VR);                // valid V rows of this half-warp in this tile (may be <= 0)
        // ---- scores of this thread's key against every live hyp
        float s[NHT];
        {
            // four independent partial sums per hyp: the 64-term dot product is 4 chains of 16 dependent FMAs, not one of 64
            // (the kernel runs at 3-4 warps per scheduler: dependent-issue latency, not throughput, paces it)
            float p0[NHT], p1[NHT], p2[NHT], p3[NHT];
#pragma unroll
            for (int h = 0; h < NHT; ++h) p0[h] = p1[h] = p2[h] = p3[h] = 0.f;
#pragma unroll
            for (int j = 0; j < DH / 4; ++j) {
#pragma unroll
                for (int h = 0; h < NHT; ++h) {
                    const float4 q = *reinterpret_cast<const float4*>(&sm.qs[h][4 * j]);
                    p0[h] = fmaf(q.x, kreg[j].x, p0[h]); p1[h] = fmaf(q.y, kreg[j].y, p1[h]);
                    p2[h] = fmaf(q.z, kreg[j].z, p2[h]); p3[h] = fmaf(q.w, kreg[j].w, p3[h]);
                }
            }
#pragma unroll
            for (int h = 0; h < NHT; ++h) s[h] = (p0[h] + p1[h]) + (p2[h] + p3[h]);
        }
        if (tile + 1 < ntiles) request_k(t0 + CK);   // the key registers are free: next tile's keys fly during softmax and P.V
#pragma unroll
        for (int h = 0; h < NHT; ++h) {
            s[h] = (kvalid && ((my >> h) & 1u)) ? s[h] * 0.125f : -INFINITY;
            const float mx = warp_max(s[h]);
            if (lane == 0) sm.s_redm[warp][h] = mx;
        }
        __syncthreads();                             // (1) tile maxima visible
        float scale[NHT], mnew[NHT];
#pragma unroll
        for (int h = 0; h < NHT; ++h) {
            const float mt = fmaxf(fmaxf(sm.s_redm[0][h], sm.s_redm[1][h]), fmaxf(sm.s_redm[2][h], sm.s_redm[3][h]));
            const float mo = sm.s_run[0][h];
            const float mn = fmaxf(mo, mt);          // -inf only while no key of hyp h has been seen yet
            float e = 0.f;
            scale[h] = 1.f;
            if (mn > -INFINITY) {
                // ex2.approx (2^-22 relative): far below the fp32 rounding of the 64-term scores themselves
                e = __expf(s[h] - mn);               // masked / invalid keys: exp(-inf) = 0
                scale[h] = __expf(mo - mn);          // mo = -inf -> 0 (nothing accumulated yet)
            }
            mnew[h] = mn;
            sm.sc[h][tid] = e;
            const float sum = warp_sum(e);
            if (lane == 0) sm.s_reds[warp][h] = sum;
        }
        cp_async_wait_all();                         // this thread's V pieces have landed
        __syncthreads();                             // (2) exponentials and warp sums visible; everyone has read s_run
#pragma unroll
        for (int h = 0; h < NHT; ++h) {
            if (tid == h) {
                sm.s_run[1][h] = sm.s_run[1][h] * scale[h] + ((sm.s_reds[0][h] + sm.s_reds[1][h]) + (sm.s_reds[2][h] + sm.s_reds[3][h]));
                sm.s_run[0][h] = mnew[h];
            }
        }
        // ---- acc = acc * scale + sum_keys e * V : half-warp = 16 consecutive keys, lane = 4 output dims
        if (tile > 0) {
#pragma unroll
            for (int h = 0; h < NHT; ++h) { acc[h][0] *= scale[h]; acc[h][1] *= scale[h]; acc[h][2] *= scale[h]; acc[h][3] *= scale[h]; }
        }
        const float* vrow = vtile + (hw * VR) * DH + 4 * l16;
#pragma unroll
        for (int i4 = 0; i4 < VR / 4; ++i4) {
            if (4 * i4 < nv) {                       // uniform per half-warp; rows that were not copied hold stale data
                float w[NHT][4];
#pragma unroll
                for (int h = 0; h < NHT; ++h) {
                    const float4 t = *reinterpret_cast<const float4*>(&sm.sc[h][hw * VR + 4 * i4]);
                    w[h][0] = t.x; w[h][1] = t.y; w[h][2] = t.z; w[h][3] = t.w;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (4 * i4 + u < nv) {
                        const float4 v = *reinterpret_cast<const float4*>(vrow + (4 * i4 + u) * DH);
#pragma unroll
                        for (int h = 0; h < NHT; ++h) {
                            acc[h][0] = fmaf(w[h][u], v.x, acc[h][0]); acc[h][1] = fmaf(w[h][u], v.y, acc[h][1]);
                            acc[h][2] = fmaf(w[h][u], v.z, acc[h][2]); acc[h][3] = fmaf(w[h][u], v.w, acc[h][3]);
                        }
                    }
                }
            }
        }
        if (tile + 1 < ntiles) request_v(t0 + CK);   // this thread's V slots are free again
    }
    if (MODE == 0) dbg_stamp(6);
    // ---- merge the eight half-warp accumulators (fixed order); they all refer to the same running max
#pragma unroll
    for (int h = 0; h < NHT; ++h) *reinterpret_cast<float4*>(&sm.s_o[hw][h][4 * l16]) = make_float4(acc[h][0], acc[h][1], acc[h][2], acc[h][3]);
    __syncthreads();

    auto store_out = [&](int h, int d, float v) {
        const long long row = row0 + h;
        if (a.out) a.out[row * D + head * DH + d] = v;
        if (a.out_split) avsr_split3c_store(a.out_split + row * 3 * D, D, head * DH + d, v);
    };
    for (int i = tid; i < NHT * DH; i += CK) {
        const int h = i / DH, d = i % DH;
        float o = 0.f;
#pragma unroll
        for (int w = 0; w < NHW; ++w) o += sm.s_o[w][h][d];
        store_out(h, d, o / sm.s_run[1][h]);
    }
    if (MODE == 0) dbg_stamp(7);
}

// SPLITQ: q (| k | v) arrive as split-K partial sums (round 1's projections) instead of finished rows.  The decode position
// with the cluster projections always passes finished rows; its instantiation leaves the gather code out (instruction
// fetch is a measurable part of these ~20 us launches: ncu no_instruction 25 % of the stall samples).
template <int MODE, int NH, bool SPLITQ>
__global__ void __launch_bounds__(CK, (NH <= 4) ? 4 : 2)
dec_attn_stream_kernel(const AttnArgs a) {
    extern __shared__ __align__(16) float vtile[];                 // [CK][64] V rows of the current tile (cp.async target), then rlist
    __shared__ __align__(16) AttnSmem<NH> sm;
    unsigned* rlist = reinterpret_cast<unsigned*>(vtile + CK * DH); // mode 0: (row index << 8) | hyp mask, row = pos * beam + slot
    const int tid = threadIdx.x;
    const int hw = tid >> 4, l16 = tid & 15;
    const int utt = blockIdx.x, head = blockIdx.y;
    if (MODE == 0) dbg_stamp(0);
    pdl_trigger();
    if (a.pf != nullptr && tid == 0) {
        constexpr long long PIECE = 16 * 1024;
        const long long ncta = (long long)gridDim.x * gridDim.y, cta = (long long)blockIdx.y * gridDim.x + blockIdx.x;
        const long long share = ((a.pf_bytes + ncta - 1) / ncta + PIECE - 1) / PIECE * PIECE;
        const long long lo = cta * share, hi = min(a.pf_bytes, lo + share);
        for (long long o = lo; o < hi; o += PIECE) {
            const unsigned n = (unsigned)min(PIECE, hi - o) & ~15u;
            if (n) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.pf + o), "r"(n) : "memory");
        }
    }

    // K^T base (groups [j][row][8]) and V base ([row][64]) of this (utterance, head); nr = rows per j plane
    const float* kbase;
    const float* vbase;
    long long nr;
    int T_utt = 0;
    float4 kreg[DH / 4];
    if (MODE == 1) {
        T_utt = a.utt_T[utt];
        nr = a.n_frames;
        const long long uoff = a.utt_off[utt];
        kbase = a.kc + (long long)head * nr * DH + uoff * KG;
        vbase = a.vc + (long long)head * nr * DH + uoff * DH;
        // first tile requested before the wait; rows past the end of the utterance are skipped
        load_k<1>(kreg, kbase, nr, tid, tid < T_utt);
        const int nv = T_utt - hw * VR;
        const float* vsrc = vbase + (long long)(hw * VR) * DH + 4 * l16;
        const uint32_t vt_s = (uint32_t)__cvta_generic_to_shared(vtile) + (uint32_t)((hw * VR) * DH + 4 * l16) * 4u;
#pragma unroll
        for (int i = 0; i < VR; ++i)
            if (i < nv) cp_async16(vt_s + i * DH * 4, vsrc + i * DH);
        // the rest of this (utterance, head)'s K / V does not depend on the query either: ask the L2 for it now, while the
        // query projection this launch waits for is still running and HBM is idle (its weights are a few MB)
        if (a.kv_ahead && tid < 9 && T_utt > CK) {
            const unsigned rows = (unsigned)(T_utt - CK);
            if (tid < 8) {
                const unsigned n = (rows * 32u) & ~15u;
                if (n) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(kbase + ((long long)tid * nr + CK) * KG), "r"(n) : "memory");
            } else {
                for (unsigned o = 0; o < rows * 256u; o += 32768u) {
                    const unsigned n = min(32768u, rows * 256u - o);
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const char*>(vbase + (long long)CK * DH) + o), "r"(n) : "memory");
                }
            }
        }
    } else {
        nr = (long long)a.lmax * a.beam;
        kbase = a.kc + (long long)(utt * HEADS + head) * nr * DH;
        vbase = a.vc + (long long)(utt * HEADS + head) * nr * DH;
        // self-attention: the dense copy of the converged history.  *step_p may still be one position old here (the chain that
        // advances it is upstream of the wait); it only sizes a prefetch hint
        if (a.kv_ahead && a.conv_len != nullptr && tid < 9) {
            const int L = min(*a.step_p, a.lmax);
            if (L > 0) {
                const float* kdb0 = a.kd + (long long)(utt * HEADS + head) * a.lmax * DH;
                const float* vdb0 = a.vd + (long long)(utt * HEADS + head) * a.lmax * DH;
                if (tid < 8) {
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(kdb0 + (long long)tid * a.lmax * KG), "r"((unsigned)L * 32u) : "memory");
                } else {
                    for (unsigned o = 0; o < (unsigned)L * 256u; o += 32768u) {
                        const unsigned n = min(32768u, (unsigned)L * 256u - o);
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const char*>(vdb0) + o), "r"(n) : "memory");
                    }
                }
            }
        }
    }
    pdl_wait();
    if (MODE == 0) dbg_stamp(1);
    // everything the body needs from the beam state is requested in ONE round trip (both parities of conv_len: which one
    // applies depends on *step)
    const int nh = a.n_run[utt];
    const int step = *a.step_p;
    int conv0 = 0, conv1 = 0;
    if (MODE == 0 && a.conv_len != nullptr) {
        conv0 = a.conv_len[utt];
        conv1 = a.conv_len[a.R / a.beam + utt];
    }
    float4 pre[3];
    pre[0] = pre[1] = pre[2] = make_float4(0.f, 0.f, 0.f, 0.f);
    if ((!SPLITQ || a.nsplit <= 0) && tid < NH * (DH / 4)) {
        // same round trip as the beam state: this position's query (self-attention: and key, value) of all NH slots
        const float* qp = a.q_in + (long long)(utt * a.beam + tid / (DH / 4)) * a.ldq + head * DH + 4 * (tid % (DH / 4));
        if (tid / (DH / 4) < a.beam) {
            pre[0] = *reinterpret_cast<const float4*>(qp);
            if (MODE == 0) { pre[1] = *reinterpret_cast<const float4*>(qp + D); pre[2] = *reinterpret_cast<const float4*>(qp + 2 * D); }
            if (MODE == 1 && a.q_stats != nullptr) {
                // same round trip: the row's eight tile statistics and this group's fold vectors
                const int row = utt * a.beam + tid / (DH / 4), col = head * DH + 4 * (tid % (DH / 4));
                float2 st[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) st[u] = __ldcg(reinterpret_cast<const float2*>(a.q_stats + ((long long)u * a.R + row) * 2));
                const float4 u4 = __ldg(reinterpret_cast<const float4*>(a.q_u + col));
                const float4 c4 = __ldg(reinterpret_cast<const float4*>(a.q_c + col));
                float ms = 0.f;
#pragma unroll
                for (int u = 0; u < 8; ++u) ms += st[u].x;
                const float mean = ms * 0.125f;
                float m2 = 0.f;                      // Chan's merge of the equal-sized tiles, in tile order
#pragma unroll
                for (int u = 0; u < 8; ++u) { const float d = st[u].x - mean; m2 += st[u].y + 128.f * d * d; }
                const float rstd = rsqrtf(m2 * (1.f / 1024.f) + a.q_eps);
                pre[0].x = rstd * (pre[0].x - mean * u4.x) + c4.x; pre[0].y = rstd * (pre[0].y - mean * u4.y) + c4.y;
                pre[0].z = rstd * (pre[0].z - mean * u4.z) + c4.z; pre[0].w = rstd * (pre[0].w - mean * u4.w) + c4.w;
            }
        }
    }
    if (nh == 0) { cp_async_wait_all(); return; }
    const int conv = ((step + 1) & 1) ? conv1 : conv0;
#define AVSR_BODY(NHT) attn_body<MODE, NH, NHT, SPLITQ>(a, sm, rlist, vtile, kreg, kbase, vbase, nr, T_utt, step, conv, pre)
    if (NH <= 4) {
        switch (nh) {
            case 1: AVSR_BODY(1); break;
            case 2: AVSR_BODY(2); break;
            case 3: AVSR_BODY(3); break;
            default: AVSR_BODY(4); break;
        }
    } else {
        switch (nh) {
            case 1: AVSR_BODY(1); break;
            case 2: AVSR_BODY(2); break;
            case 3: AVSR_BODY(3); break;
            case 4: AVSR_BODY(4); break;
            case 5: AVSR_BODY(5); break;
            case 6: AVSR_BODY(6); break;
            case 7: AVSR_BODY(7); break;
            default: AVSR_BODY(8); break;
        }
    }
#undef AVSR_BODY
}

// Promotion of converged history positions to the dense caches.  A position p < *step is converged for an utterance when
// all its live hyps have the same ancestor slot there; it stays converged for all descendants (new hyps copy their parent's
// ancestry), so the row (pos, slot) can be copied once to row `pos` of the dense caches and read from there from now on:
// the self-attention then streams consecutive rows instead of every beam-th row of the per-slot cache (whose 32-byte keys
// cost a full 128-byte DRAM line each).  CTA = (utterance, layer); conv_len is double-buffered on step parity because the
// CTAs of one utterance (one per layer) all read the old value.  At most PROMOTE_MAX positions per call.
constexpr int PROMOTE_MAX = 32;

__global__ void __launch_bounds__(256)
dec_cache_promote_kernel(const float* __restrict__ kc, const float* __restrict__ vc, float* __restrict__ kd, float* __restrict__ vd,
                         long long sparse_layer_stride, long long dense_layer_stride, const unsigned char* __restrict__ anc, int lmax,
                         const int* __restrict__ n_run, int beam, int R, const int* __restrict__ step_p, int* __restrict__ conv_len) {
    __shared__ int s_cnt;
    __shared__ unsigned char s_slot[PROMOTE_MAX];
    const int utt = blockIdx.x, layer = blockIdx.y, B = R / beam;
    const int tid = threadIdx.x;
    pdl_trigger();
    pdl_wait();
    const int step = *step_p;
    const int nh = n_run[utt];
    const int C = conv_len[(step & 1) * B + utt];
    if (tid < 32) {
        const int p = C + tid;
        bool ok = nh > 0 && p < step;
        unsigned char s0 = 0;
        if (ok) {
            const unsigned char* ar = anc + ((long long)(step & 1) * R + utt * beam) * lmax + p;
            s0 = ar[0];
            for (int h = 1; h < nh; ++h) ok = ok && (ar[(long long)h * lmax] == s0);
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        const int cnt = (m == 0xffffffffu) ? 32 : __ffs(~m) - 1;      // leading converged positions
        if (tid < cnt) s_slot[tid] = s0;
        if (tid == 0) {
            s_cnt = cnt;
            if (layer == 0) conv_len[((step + 1) & 1) * B + utt] = C + cnt;
        }
    }
    __syncthreads();
    const int cnt = s_cnt;
    if (cnt == 0) return;
    const long long nr = (long long)lmax * beam;
    const float* kcl = kc + layer * sparse_layer_stride + (long long)utt * HEADS * nr * DH;
    const float* vcl = vc + layer * sparse_layer_stride + (long long)utt * HEADS * nr * DH;
    float* kdl = kd + layer * dense_layer_stride + (long long)utt * HEADS * lmax * DH;
    float* vdl = vd + layer * dense_layer_stride + (long long)utt * HEADS * lmax * DH;
    // per (position, head): 16 key pieces + 16 value pieces of 16 bytes
    for (int i = tid; i < cnt * HEADS * 32; i += 256) {
        const int piece = i & 15, kv = (i >> 4) & 1, head = (i >> 5) % HEADS, pi = i / (32 * HEADS);
        const int p = C + pi;
        const long long rs = (long long)p * beam + s_slot[pi];
        if (kv == 0) {
            const int j = piece >> 1, half = piece & 1;
            const float4 v = *reinterpret_cast<const float4*>(kcl + (long long)head * nr * DH + ((long long)j * nr + rs) * KG + half * 4);
            *reinterpret_cast<float4*>(kdl + (long long)head * lmax * DH + ((long long)j * lmax + p) * KG + half * 4) = v;
        } else {
            const float4 v = *reinterpret_cast<const float4*>(vcl + (long long)head * nr * DH + rs * DH + piece * 4);
            *reinterpret_cast<float4*>(vdl + (long long)head * lmax * DH + (long long)p * DH + piece * 4) = v;
        }
    }
}

}  // namespace

// mode 0: self-attention step.  Query / current k / current v = columns [0,1024) / [1024,2048) / [2048,3072) of q_in
//   ([R, ldq] fp32).  kc / vc = this layer's caches; with row = pos*beam + slot and nr = lmax*beam, key element
//   (utt, head, row, d) is at ((utt*16 + head)*8 + d/8)*nr*8 + row*8 + d%8 (transposed in 32-byte groups) and value
//   element at ((utt*16 + head)*nr + row)*64 + d.  anc [2][R][lmax] (uint8 slot per position, double-buffered on step
//   parity).  The current k / v are written to the cache at (pos = *step, slot = the hyp's own slot).
// mode 1: source attention.  Query = columns [0,1024) of q_in; kc / vc = this layer's cross K / V over the n_frames packed
//   frames of all utterances (utt_off / utt_T index them): key element (head, frame, d) at (head*8 + d/8)*n_frames*8 +
//   frame*8 + d%8, value element at (head*n_frames + frame)*64 + d (the layout avsr_kv_head_major writes).
// nsplit > 0: q_in holds the split-K partial sums part[z][R][ldq] of the projection (z < nsplit); they are summed in a fixed
//   order and q_bias[ldq] is added.  nsplit == 0: q_in is the finished projection.
// out (fp32 [R,1024]) and / or out_split (compact bf16x3 [R, 3*1024]).
// kd / vd / conv_len (mode 0, optional): dense caches of the converged history prefix maintained by avsr_dec_cache_promote,
//   key element (utt, head, pos, d) at ((utt*16 + head)*8 + d/8)*lmax*8 + pos*8 + d%8, value element at
//   ((utt*16 + head)*lmax + pos)*64 + d; conv_len [2][R/beam], the kernel reads conv_len[(*step + 1) & 1][utt].
namespace {
const float* g_q_stats = nullptr;
const float* g_q_u = nullptr;
const float* g_q_c = nullptr;
float g_q_eps = 0.f;
}  // namespace

// The NEXT source-attention launch (mode 1, finished-row query) receives its query WITHOUT the LayerNorm of the row it was
// projected from and finishes it itself: q = rstd (q_in - mean u) + c with (mean, rstd) from stats [8][R][2] (tile statistics of
// the 1024-wide row, avsr_dec_proj's stats_out), u / c [1024] = weights.fold_layernorm's vectors.  One-shot.
extern "C" int avsr_dec_attn_fold_query(const float* stats, const float* u, const float* c, float eps) {
    if (stats != nullptr && (!u || !c || (((uintptr_t)stats & 7) | ((uintptr_t)u & 15) | ((uintptr_t)c & 15)) != 0)) return AVSR_ERR_ARG;
    g_q_stats = stats; g_q_u = u; g_q_c = c; g_q_eps = eps;
    return AVSR_OK;
}

extern "C" int avsr_dec_attn_step_pf(int mode, const float* q_in, long long ldq, int nsplit, const float* q_bias, float* kc, float* vc,
                                     const unsigned char* anc, int lmax, const int* n_run, const int* utt_off, const int* utt_T, int beam,
                                     int R, const int* step, float* out, long long n_frames, void* out_split, const float* kd,
                                     const float* vd, const int* conv_len, const void* l2_prefetch, long long l2_prefetch_bytes,
                                     cudaStream_t stream) {
    AVSR_REQUIRE(!l2_prefetch || (((uintptr_t)l2_prefetch & 15) == 0 && l2_prefetch_bytes > 0), "avsr_dec_attn_step: prefetch span must be 16-byte aligned");
    AVSR_REQUIRE((kd != nullptr) == (vd != nullptr) && (kd != nullptr) == (conv_len != nullptr) && (mode == 0 || kd == nullptr),
                 "avsr_dec_attn_step: kd / vd / conv_len go together (self-attention only)");
    AVSR_REQUIRE(((uintptr_t)kd & 31) == 0 && ((uintptr_t)vd & 15) == 0, "avsr_dec_attn_step: kd / vd alignment");
    AVSR_REQUIRE(q_in && kc && vc && n_run && step && (out || out_split) && R > 0 && beam > 0 && lmax > 0, "avsr_dec_attn_step: bad arguments");
    AVSR_REQUIRE(mode == 0 ? (anc != nullptr) : (utt_off && utt_T && n_frames > 0), "avsr_dec_attn_step: missing index arrays for mode %d", mode);
    AVSR_REQUIRE(beam <= 8 && R % beam == 0, "avsr_dec_attn_step: beam %d unsupported (max 8)", beam);
    AVSR_REQUIRE(nsplit >= 0 && (nsplit == 0 || q_bias), "avsr_dec_attn_step: partial-sum input needs the bias");
    AVSR_REQUIRE((ldq & 3) == 0 && ((uintptr_t)q_in & 15) == 0 && ((uintptr_t)q_bias & 15) == 0 && ((uintptr_t)kc & 31) == 0 &&
                     ((uintptr_t)vc & 15) == 0,
                 "avsr_dec_attn_step: q_in / q_bias / vc must be 16-byte aligned, kc 32-byte aligned and ldq a multiple of 4");
    AVSR_REQUIRE(mode == 1 || (long long)lmax * beam < (1 << 24), "avsr_dec_attn_step: cache too long");
    static int kv_ahead = -1;                // dev knob AVSR_ATTN_KV_AHEAD: bit 0 = source attention, bit 1 = self-attention
    if (kv_ahead < 0) { const char* e = getenv("AVSR_ATTN_KV_AHEAD"); kv_ahead = e ? atoi(e) : 0; }
    const int nslots = beam <= 4 ? 4 : 8;
    // V tile + (self-attention) the list of distinct history rows, at most lmax * beam entries
    const size_t smem = (size_t)CK * DH * sizeof(float) + (mode == 0 ? (size_t)lmax * nslots * sizeof(unsigned) : 0);
    AVSR_REQUIRE(smem <= 160 * 1024, "avsr_dec_attn_step: %d positions do not fit the shared-memory row list", lmax);
    static size_t configured[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int ki = ((mode == 0 ? 0 : 2) + (nslots == 4 ? 0 : 1)) * 2 + (nsplit > 0 ? 1 : 0);
    void (*kern)(const AttnArgs) = nullptr;
    switch (ki) {
        case 0: kern = dec_attn_stream_kernel<0, 4, false>; break;
        case 1: kern = dec_attn_stream_kernel<0, 4, true>; break;
        case 2: kern = dec_attn_stream_kernel<0, 8, false>; break;
        case 3: kern = dec_attn_stream_kernel<0, 8, true>; break;
        case 4: kern = dec_attn_stream_kernel<1, 4, false>; break;
        case 5: kern = dec_attn_stream_kernel<1, 4, true>; break;
        case 6: kern = dec_attn_stream_kernel<1, 8, false>; break;
        default: kern = dec_attn_stream_kernel<1, 8, true>; break;
    }
    if (smem > configured[ki]) {
        const int lim = 160 * 1024;
        AVSR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
        configured[ki] = lim;
    }
    const dim3 grid(R / beam, HEADS);
    const AttnArgs a = {q_in, ldq, nsplit, q_bias, kc, vc, anc, lmax, kd, vd, conv_len, n_run, utt_off, utt_T, beam, R, step, out,
                        n_frames, (__nv_bfloat16*)out_split, (const char*)l2_prefetch, l2_prefetch ? l2_prefetch_bytes : 0, (kv_ahead >> (mode == 1 ? 0 : 1)) & 1,
                        g_q_stats, g_q_u, g_q_c, g_q_eps};
    g_q_stats = nullptr;                              // one-shot (avsr_dec_attn_fold_query)
    AVSR_CHECK_CUDA(avsr_launch_pdl(kern, grid, dim3(CK), smem, stream, a));
    return AVSR_OK;
}

extern "C" int avsr_dec_attn_step(int mode, const float* q_in, long long ldq, int nsplit, const float* q_bias, float* kc, float* vc,
                                  const unsigned char* anc, int lmax, const int* n_run, const int* utt_off, const int* utt_T, int beam,
                                  int R, const int* step, float* out, long long n_frames, void* out_split, const float* kd,
                                  const float* vd, const int* conv_len, cudaStream_t stream) {
    return avsr_dec_attn_step_pf(mode, q_in, ldq, nsplit, q_bias, kc, vc, anc, lmax, n_run, utt_off, utt_T, beam, R, step, out, n_frames,
                                 out_split, kd, vd, conv_len, nullptr, 0, stream);
}

// Copies the newly converged history positions of every utterance from the per-slot caches (all n_layers layers, layer
// stride = B*16*lmax*beam*64 floats) to the dense caches (layer stride B*16*lmax*64) and advances
// conv_len[(*step + 1) & 1][utt] = conv_len[*step & 1][utt] + promoted.  Call once per position, before the layers run.
extern "C" int avsr_dec_cache_promote(const float* kc, const float* vc, float* kd, float* vd, int n_layers, const unsigned char* anc,
                                      int lmax, const int* n_run, int beam, int R, const int* step, int* conv_len,
                                      cudaStream_t stream) {
    AVSR_REQUIRE(kc && vc && kd && vd && anc && n_run && step && conv_len && n_layers > 0 && lmax > 0 && beam > 0 && R > 0 && R % beam == 0,
                 "avsr_dec_cache_promote: bad arguments");
    const long long B = R / beam;
    const long long sls = B * HEADS * (long long)lmax * beam * DH, dls = B * HEADS * (long long)lmax * DH;
    AVSR_CHECK_CUDA(avsr_launch_pdl(dec_cache_promote_kernel, dim3((unsigned)B, n_layers), dim3(256), 0, stream, kc, vc, kd, vd, sls, dls, anc,
                                    lmax, n_run, beam, R, step, conv_len));
    return AVSR_OK;
}

// Dev aid: buf = 8 uint64 in device memory (or NULL to switch off); CTA (0, 0) of every self-attention launch then records
// globaltimer stamps at: kernel entry, after griddepcontrol.wait, query gather start, row list start, row list done, tile loop
// start, tile loop done, output stored.
extern "C" int avsr_dec_attn_debug(unsigned long long* buf) {
    AVSR_CHECK_CUDA(cudaMemcpyToSymbol(g_attn_dbg, &buf, sizeof(buf)));
    return AVSR_OK;
}
