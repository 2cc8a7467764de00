// Output-layer softmax + pre-beam top-S of one decoder row, as a device function shared by the stand-alone kernel
// (csrc/decode.cu) and the fused tail of a decode position (csrc/ctc.cu).
// logits = sum_z part[z][row] + bias ; logp = log_softmax(logits) -> dec_logp[row] ; ids = top-S token ids (value descending,
// ties to the lower id).  Reference: decoder.py:176-181 (output layer + log_softmax), batch_beam_search.py:229-235 (pre-beam).
// One CTA of LSM_THREADS threads per row; a thread keeps its <= ITER strided logits in registers for all three passes (max,
// sum of exponentials, top-S), so the row is read once and nothing is staged.
#pragma once
#include "common.cuh"

constexpr int LSM_THREADS = 512;

struct LsmSmem {
    float red[32];
    float s_v[LSM_THREADS / 32];
    int s_i[LSM_THREADS / 32];
    int s_win;
};

// v[]: the caller pre-loads the bias (it does not depend on the previous kernel) before griddepcontrol.wait.
template <int ITER>
__device__ __forceinline__ void lsm_load_bias(float (&v)[ITER], const float* __restrict__ bias, int V) {
#pragma unroll
    for (int k = 0; k < ITER; ++k) {
        const int c = threadIdx.x + k * LSM_THREADS;
        v[k] = (c < V) ? __ldg(bias + c) : 0.f;
    }
}

// ids_g: part_ids row in global memory; ids_s (optional): the same ids in shared memory for a fused consumer.
template <int ITER>
__device__ __forceinline__ void lsm_topk_row(float (&v)[ITER], LsmSmem& sm, const float* __restrict__ part, int nsplit, int R, int V,
                                             int row, float* __restrict__ logp, int* __restrict__ ids_g, int* ids_s, int S) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    // partial sums: split-major so that the ITER loads of one split are all in flight (nsplit round trips, not ITER * nsplit)
    float acc[ITER];
#pragma unroll
    for (int k = 0; k < ITER; ++k) acc[k] = 0.f;
    for (int z = 0; z < nsplit; z += 2) {             // two splits per pass: 2 * ITER independent loads in flight
        const float* pz = part + ((long long)z * R + row) * V;
        const float* pz1 = pz + (long long)R * V;
        const bool two = z + 1 < nsplit;
        float t[ITER], u[ITER];
#pragma unroll
        for (int k = 0; k < ITER; ++k) {
            const int c = tid + k * LSM_THREADS;
            t[k] = (c < V) ? pz[c] : 0.f;
            u[k] = (c < V && two) ? pz1[c] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < ITER; ++k) { acc[k] += t[k]; if (two) acc[k] += u[k]; }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < ITER; ++k) {
        const int c = tid + k * LSM_THREADS;
        if (c < V) {
            v[k] = acc[k] + v[k];
            mx = fmaxf(mx, v[k]);
        } else {
            v[k] = -INFINITY;
        }
    }
    mx = block_max(mx, sm.red);
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < ITER; ++k) sum += expf(v[k] - mx);            // padding: exp(-inf) = 0
    sum = block_sum(sum, sm.red);
    const float lse = logf(sum);
#pragma unroll
    for (int k = 0; k < ITER; ++k) {
        const int c = tid + k * LSM_THREADS;
        if (c < V) {
            v[k] = (v[k] - mx) - lse;
            logp[(long long)row * V + c] = v[k];
        }
    }
    // ---- S rounds of block arg-max over the register values; the winner drops its entry
    for (int r = 0; r < S; ++r) {
        float bv = -INFINITY;
        int bi = 0x7fffffff;
#pragma unroll
        for (int k = 0; k < ITER; ++k) {
            const int c = tid + k * LSM_THREADS;
            if (v[k] > bv) { bv = v[k]; bi = c; }                     // ascending c: the lowest id wins ties
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { sm.s_v[w] = bv; sm.s_i[w] = bi; }
        __syncthreads();
        if (w == 0) {
            bv = lane < LSM_THREADS / 32 ? sm.s_v[lane] : -INFINITY;
            bi = lane < LSM_THREADS / 32 ? sm.s_i[lane] : 0x7fffffff;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) {
                ids_g[r] = bi;
                if (ids_s) ids_s[r] = bi;
                sm.s_win = bi;
            }
        }
        __syncthreads();
        const int win = sm.s_win;
#pragma unroll
        for (int k = 0; k < ITER; ++k)
            if (tid + k * LSM_THREADS == win) v[k] = -INFINITY;
    }
}
