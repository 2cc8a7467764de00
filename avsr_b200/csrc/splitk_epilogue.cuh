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

// ---- the same row epilogue on a GROUP of GT threads of a CTA (named barrier `bar`; tid = index inside the group), for the
// projection kernel that finishes the rows of the PREVIOUS projection before it loads them as its own operand
// (csrc/gemm_x3.cu).  Vector path only: N, ldr, ldo multiples of 4.
template <int GT>
__device__ __forceinline__ float avsr_group_sum(float v, float* red, int tid, int bar) {
    const int lane = tid & 31, w = tid >> 5;
    v = warp_sum(v);
    asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(GT) : "memory");
    if (lane == 0) red[w] = v;
    asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(GT) : "memory");
    float r = (lane < GT / 32) ? red[lane] : 0.f;
    return warp_sum(r);
}

template <int GT>
__device__ __forceinline__ void avsr_splitk_epilogue_row_group(const SplitKEpi& e, int row, float* rowbuf, float* red, int tid, int bar) {
    if (e.row_active != nullptr && e.row_active[row] == 0) return;
    const int nsplit = e.nsplit, N = e.N, act = e.act;
    const long long zstride = (long long)e.M * N;
    float lsum = 0.f;
    // two column groups per pass, eight splits of both requested before the first add: the partial sums of a 1024-wide row
    // with 16 splits cost two L2 round trips per thread (this runs on the critical path of the projection that follows)
    for (int c0 = tid * 4; c0 < N; c0 += GT * 8) {
        const int c1 = c0 + GT * 4;
        const bool two = c1 < N;
        const float* p0 = e.part + (long long)row * N + c0;
        const float* p1 = e.part + (long long)row * N + (two ? c1 : c0);
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
        int z = 0;
        for (; z + 8 <= nsplit; z += 8) {
            float4 t0[8], t1[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                t0[u] = __ldcg(reinterpret_cast<const float4*>(p0 + (z + u) * zstride));
                t1[u] = __ldcg(reinterpret_cast<const float4*>(p1 + (z + u) * zstride));
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                v0.x += t0[u].x; v0.y += t0[u].y; v0.z += t0[u].z; v0.w += t0[u].w;
                v1.x += t1[u].x; v1.y += t1[u].y; v1.z += t1[u].z; v1.w += t1[u].w;
            }
        }
        {
            float4 t0[8], t1[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const bool ok = z + u < nsplit;
                t0[u] = ok ? __ldcg(reinterpret_cast<const float4*>(p0 + (z + u) * zstride)) : make_float4(0.f, 0.f, 0.f, 0.f);
                t1[u] = ok ? __ldcg(reinterpret_cast<const float4*>(p1 + (z + u) * zstride)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (z + u < nsplit) {
                    v0.x += t0[u].x; v0.y += t0[u].y; v0.z += t0[u].z; v0.w += t0[u].w;
                    v1.x += t1[u].x; v1.y += t1[u].y; v1.z += t1[u].z; v1.w += t1[u].w;
                }
            }
        }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            if (g == 1 && !two) break;
            const int c = g == 0 ? c0 : c1;
            float4 v = g == 0 ? v0 : v1;
            if (e.bias) {
                const float4 b4 = *reinterpret_cast<const float4*>(e.bias + c);
                v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
            }
            if (act == AVSR_ACT_RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            else if (act == AVSR_ACT_GELU) { v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w); }
            if (e.residual) {
                const float4 r4 = *reinterpret_cast<const float4*>(e.residual + (long long)row * e.ldr + c);
                v.x += r4.x; v.y += r4.y; v.z += r4.z; v.w += r4.w;
            }
            if (e.out) *reinterpret_cast<float4*>(e.out + (long long)row * e.ldo + c) = v;
            if (e.ln_g) { *reinterpret_cast<float4*>(rowbuf + c) = v; lsum += (v.x + v.y) + (v.z + v.w); }
            else if (e.split_out) avsr_split3c_store4(e.split_out + (long long)row * 3 * N, N, c, v);
        }
    }
    if (e.ln_g == nullptr) return;
    // LayerNorm of the finished row (every thread re-reads only the columns it wrote)
    const float mean = avsr_group_sum<GT>(lsum, red, tid, bar) / (float)N;
    float lvar = 0.f;
    for (int c = tid * 4; c < N; c += GT * 4) {
        const float4 t = *reinterpret_cast<const float4*>(rowbuf + c);
        const float d0 = t.x - mean, d1 = t.y - mean, d2 = t.z - mean, d3 = t.w - mean;
        lvar += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
    const float rstd = rsqrtf(avsr_group_sum<GT>(lvar, red, tid, bar) / (float)N + e.ln_eps);
    for (int c = tid * 4; c < N; c += GT * 4) {
        const float4 t = *reinterpret_cast<const float4*>(rowbuf + c);
        const float4 g4 = *reinterpret_cast<const float4*>(e.ln_g + c);
        const float4 b4 = *reinterpret_cast<const float4*>(e.ln_b + c);
        const float4 y = make_float4((t.x - mean) * rstd * g4.x + b4.x, (t.y - mean) * rstd * g4.y + b4.y,
                                     (t.z - mean) * rstd * g4.z + b4.z, (t.w - mean) * rstd * g4.w + b4.w);
        if (e.ln_out) *reinterpret_cast<float4*>(e.ln_out + (long long)row * e.ld_ln + c) = y;
        if (e.split_out) avsr_split3c_store4(e.split_out + (long long)row * 3 * N, N, c, y);
    }
}
