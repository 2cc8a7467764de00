// Row-wise epilogue of the split-K decoder projections: v = sum_z part[z][row][:] + bias ; act ; + residual -> out (fp32);
// optionally LayerNorm of the finished row -> ln_out (fp32) and / or the compact bf16x3 form -> split_out (the operand of the
// next projection).  Reference: the glue between the nn.Linear calls of DecoderLayer.forward
// (src/nets/backend/transformer/decoder_layer.py:58-121; LayerNorm eps 1e-12, layer_norm.py:12-33).
// One CTA works on one row at a time; used by splitk_epilogue_kernel (csrc/sgemm.cu) and, after a grid-wide barrier, by the
// fused tail of gemm_x3_kernel (csrc/gemm_x3.cu).  The partial sums are read with ld.global.cg: in the fused form they were
// written by other SMs during the same launch.
#pragma once
#include "common.cuh"

struct SplitKEpi {
    const float* part;
    int nsplit, M, N;
    const float* bias;
    int act;
    const float* residual;     // may alias out (in place)
    long long ldr;
    float* out;
    long long ldo;
    const float* ln_g;
    const float* ln_b;
    float ln_eps;
    float* ln_out;
    long long ld_ln;
    const int* row_active;
    __nv_bfloat16* split_out;
};

// rowbuf: N floats of shared memory (only touched when ln_g != nullptr); red: 32 floats of shared memory.
// P (1, 2 or 4; blockDim.x % P == 0): threads that share one group of four columns.  Each sums a contiguous range of the
// splits (few loads per thread, all in flight at once: one L2 round trip) and the P sums are combined with xor shuffles in
// a fixed tree, so every thread of the group holds the bit-identical total; thread 0 of the group finishes the columns.
template <int P = 1>
__device__ __forceinline__ void avsr_splitk_epilogue_row(const SplitKEpi& e, int row, float* rowbuf, float* red) {
    const float* part = e.part;
    const int nsplit = e.nsplit, M = e.M, N = e.N, act = e.act;
    const float* bias = e.bias;
    const float* residual = e.residual;
    const long long ldr = e.ldr, ldo = e.ldo, ld_ln = e.ld_ln;
    float* out = e.out;
    const float* ln_g = e.ln_g;
    const float* ln_b = e.ln_b;
    const float ln_eps = e.ln_eps;
    float* ln_out = e.ln_out;
    const int* row_active = e.row_active;
    __nv_bfloat16* split_out = e.split_out;
    if (row_active != nullptr && row_active[row] == 0) return;
    float lsum = 0.f;
    const long long zstride = (long long)M * N;
    if ((N & 3) == 0 && (ldr & 3) == 0 && (ldo & 3) == 0) {
        // vector path: 4 columns per group of P threads
        const int sub = threadIdx.x % P, grp = threadIdx.x / P, ngrp = blockDim.x / P;
        const int zper = (nsplit + P - 1) / P;
        const int z0 = sub * zper, z1 = min(nsplit, z0 + zper);
        const int ncol_iter = (N / 4 + ngrp - 1) / ngrp;              // same trip count for every thread (shuffles inside)
        for (int it = 0; it < ncol_iter; ++it) {
            const int c = (grp + it * ngrp) * 4;
            const bool live = c < N;
            const float* p = part + (long long)row * N + (live ? c : 0);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            int z = z0;
            for (; z + 4 <= z1; z += 4) {
                const float4 a0 = __ldcg(reinterpret_cast<const float4*>(p + (z + 0) * zstride));
                const float4 a1 = __ldcg(reinterpret_cast<const float4*>(p + (z + 1) * zstride));
                const float4 a2 = __ldcg(reinterpret_cast<const float4*>(p + (z + 2) * zstride));
                const float4 a3 = __ldcg(reinterpret_cast<const float4*>(p + (z + 3) * zstride));
                v.x += a0.x; v.y += a0.y; v.z += a0.z; v.w += a0.w;
                v.x += a1.x; v.y += a1.y; v.z += a1.z; v.w += a1.w;
                v.x += a2.x; v.y += a2.y; v.z += a2.z; v.w += a2.w;
                v.x += a3.x; v.y += a3.y; v.z += a3.z; v.w += a3.w;
            }
            for (; z < z1; ++z) {
                const float4 a0 = __ldcg(reinterpret_cast<const float4*>(p + z * zstride));
                v.x += a0.x; v.y += a0.y; v.z += a0.z; v.w += a0.w;
            }
            if (P > 1) {
#pragma unroll
                for (int o = 1; o < P; o <<= 1) {
                    v.x += __shfl_xor_sync(0xffffffffu, v.x, o); v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
                    v.z += __shfl_xor_sync(0xffffffffu, v.z, o); v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
                }
            }
            if (!live || sub != 0) continue;
            if (bias) {
                const float4 b4 = *reinterpret_cast<const float4*>(bias + c);
                v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
            }
            if (act == AVSR_ACT_RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            else if (act == AVSR_ACT_GELU) { v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w); }
            if (residual) {
                const float4 r4 = *reinterpret_cast<const float4*>(residual + (long long)row * ldr + c);
                v.x += r4.x; v.y += r4.y; v.z += r4.z; v.w += r4.w;
            }
            if (out) *reinterpret_cast<float4*>(out + (long long)row * ldo + c) = v;
            if (ln_g) { *reinterpret_cast<float4*>(rowbuf + c) = v; lsum += (v.x + v.y) + (v.z + v.w); }
            else if (split_out) avsr_split3c_store4(split_out + (long long)row * 3 * N, N, c, v);
        }
        if (ln_g == nullptr) return;
        // LayerNorm of the finished row (values of this thread's columns are still in rowbuf; same thread re-reads them)
        const float mean = block_sum(lsum, red) / (float)N;
        float lvar = 0.f;
        for (int c = (sub == 0 ? grp * 4 : N); c < N; c += ngrp * 4) {
            const float4 t = *reinterpret_cast<const float4*>(rowbuf + c);
            const float d0 = t.x - mean, d1 = t.y - mean, d2 = t.z - mean, d3 = t.w - mean;
            lvar += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
        }
        const float rstd = rsqrtf(block_sum(lvar, red) / (float)N + ln_eps);
        for (int c = (sub == 0 ? grp * 4 : N); c < N; c += ngrp * 4) {
            const float4 t = *reinterpret_cast<const float4*>(rowbuf + c);
            const float4 g4 = *reinterpret_cast<const float4*>(ln_g + c);
            const float4 b4 = *reinterpret_cast<const float4*>(ln_b + c);
            const float4 y = make_float4((t.x - mean) * rstd * g4.x + b4.x, (t.y - mean) * rstd * g4.y + b4.y,
                                         (t.z - mean) * rstd * g4.z + b4.z, (t.w - mean) * rstd * g4.w + b4.w);
            if (ln_out) *reinterpret_cast<float4*>(ln_out + (long long)row * ld_ln + c) = y;
            if (split_out) avsr_split3c_store4(split_out + (long long)row * 3 * N, N, c, y);
        }
        return;
    } else {
        for (int c = threadIdx.x; c < N; c += blockDim.x) {
            float v = 0.f;
            for (int z = 0; z < nsplit; ++z) v += __ldcg(part + z * zstride + (long long)row * N + c);
            if (bias) v += bias[c];
            if (act == AVSR_ACT_RELU) v = fmaxf(v, 0.f);
            else if (act == AVSR_ACT_GELU) v = gelu_erf(v);
            if (residual) v += residual[(long long)row * ldr + c];
            if (out) out[(long long)row * ldo + c] = v;
            if (ln_g) { rowbuf[c] = v; lsum += v; }
            else if (split_out) avsr_split3c_store(split_out + (long long)row * 3 * N, N, c, v);
        }
    }
    if (ln_g == nullptr) return;
    const float mean = block_sum(lsum, red) / (float)N;
    float lvar = 0.f;
    for (int c = threadIdx.x; c < N; c += blockDim.x) {
        const float d = rowbuf[c] - mean;
        lvar += d * d;
    }
    const float var = block_sum(lvar, red) / (float)N;
    const float rstd = rsqrtf(var + ln_eps);
    for (int c = threadIdx.x; c < N; c += blockDim.x) {
        const float y = (rowbuf[c] - mean) * rstd * ln_g[c] + ln_b[c];
        if (ln_out) ln_out[(long long)row * ld_ln + c] = y;
        if (split_out) avsr_split3c_store(split_out + (long long)row * 3 * N, N, c, y);
    }
}
