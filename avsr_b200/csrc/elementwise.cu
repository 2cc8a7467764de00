// Memory-bound helper kernels of the encoder: LayerNorm, im2col gathers that feed the tcgen05 GEMM (3D-conv frontend,
// ResNet 3x3 / 1x1 convs, grouped positional conv), max/avg pooling, casts and transposes.
// Reference ops: src/nets/backend/backbones/resnet.py:30-164, avhubert.py:187-198,486-502,698-734 and HF
// Wav2Vec2PositionalConvEmbedding (transformers/models/wav2vec2/modeling_wav2vec2.py:326-379).
// All activations are channels-last bf16 ([frames, H, W, C]); frames of all utterances are packed back to back and
// frame_t / frame_T give each frame's index inside its utterance and that utterance's length, so that temporal
// zero padding never crosses an utterance boundary.
#include "common.cuh"

namespace {

// ------------------------------------------------------------------ LayerNorm (fp32 in, bf16 and/or fp32 out)
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, long long ldx, int N, const float* __restrict__ g, const float* __restrict__ b,
                 float eps, __nv_bfloat16* __restrict__ out_bf16, long long ld_bf16, float* __restrict__ out_f32, long long ld_f32) {
    extern __shared__ float rowbuf[];
    __shared__ float red[32];
    const long long row = blockIdx.x;
    const float* xr = x + row * ldx;
    float s = 0.f;
    for (int c = threadIdx.x * 4; c < N; c += blockDim.x * 4) {
        const float4 v = *reinterpret_cast<const float4*>(xr + c);
        *reinterpret_cast<float4*>(rowbuf + c) = v;
        s += v.x + v.y + v.z + v.w;
    }
    const float mean = block_sum(s, red) / (float)N;
    float q = 0.f;
    for (int c = threadIdx.x * 4; c < N; c += blockDim.x * 4) {
        const float4 v = *reinterpret_cast<const float4*>(rowbuf + c);
        const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
        q += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
    const float rstd = rsqrtf(block_sum(q, red) / (float)N + eps);
    for (int c = threadIdx.x * 4; c < N; c += blockDim.x * 4) {
        const float4 v = *reinterpret_cast<const float4*>(rowbuf + c);
        const float4 gg = *reinterpret_cast<const float4*>(g + c);
        const float4 bb = *reinterpret_cast<const float4*>(b + c);
        const float o0 = (v.x - mean) * rstd * gg.x + bb.x, o1 = (v.y - mean) * rstd * gg.y + bb.y;
        const float o2 = (v.z - mean) * rstd * gg.z + bb.z, o3 = (v.w - mean) * rstd * gg.w + bb.w;
        if (out_f32) *reinterpret_cast<float4*>(out_f32 + row * ld_f32 + c) = make_float4(o0, o1, o2, o3);
        if (out_bf16) {
            __nv_bfloat162 p0 = __floats2bfloat162_rn(o0, o1), p1 = __floats2bfloat162_rn(o2, o3);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&p0);
            pk.y = *reinterpret_cast<uint32_t*>(&p1);
            *reinterpret_cast<uint2*>(out_bf16 + row * ld_bf16 + c) = pk;
        }
    }
}

// N = 1024 (every LayerNorm of the transformer): one WARP per row, the row lives in registers (8 float4 per lane), both
// reductions are warp shuffles: no shared memory, no block barriers.  The one-CTA-per-row kernel below pays three barrier phases
// per row and ran at 26 us per 12 000 rows (74 MB: 2.8 TB/s); this one is a plain streaming kernel.
__global__ void __launch_bounds__(256)
layernorm1024_kernel(const float* __restrict__ x, long long ldx, long long rows, const float* __restrict__ g, const float* __restrict__ b,
                     float eps, __nv_bfloat16* __restrict__ out_bf16, long long ld_bf16, float* __restrict__ out_f32, long long ld_f32) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* xr = x + row * ldx;
    float4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __ldcs(reinterpret_cast<const float4*>(xr + i * 128 + lane * 4));
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) * (1.f / 1024.f);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float d0 = v[i].x - mean, d1 = v[i].y - mean, d2 = v[i].z - mean, d3 = v[i].w - mean;
        q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / 1024.f) + eps);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = i * 128 + lane * 4;
        const float4 gg = __ldg(reinterpret_cast<const float4*>(g + c));
        const float4 bb = __ldg(reinterpret_cast<const float4*>(b + c));
        const float o0 = (v[i].x - mean) * rstd * gg.x + bb.x, o1 = (v[i].y - mean) * rstd * gg.y + bb.y;
        const float o2 = (v[i].z - mean) * rstd * gg.z + bb.z, o3 = (v[i].w - mean) * rstd * gg.w + bb.w;
        if (out_f32) *reinterpret_cast<float4*>(out_f32 + row * ld_f32 + c) = make_float4(o0, o1, o2, o3);
        if (out_bf16) {
            __nv_bfloat162 p0 = __floats2bfloat162_rn(o0, o1), p1 = __floats2bfloat162_rn(o2, o3);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&p0);
            pk.y = *reinterpret_cast<uint32_t*>(&p1);
            *reinterpret_cast<uint2*>(out_bf16 + row * ld_bf16 + c) = pk;
        }
    }
}

// ------------------------------------------------------------------ frontend 3D conv im2col
// video fp32 [F,88,88] (packed frames) -> A [F*44*44, 256] bf16, k = (dt*7 + dy)*7 + dx for the 5x7x7 patch
// (stride 1x2x2, pad 2x3x3), columns 245..255 zero.  One CTA per (frame, output row).
__global__ void __launch_bounds__(256)
im2col_frontend_kernel(const float* __restrict__ video, const int* __restrict__ frame_t, const int* __restrict__ frame_T,
                       __nv_bfloat16* __restrict__ out, int f0) {
    __shared__ float rows[5][7][97];       // [dt][dy][x + 3], x in [-3, 91); odd pitch: the 35 patch rows start in different banks
    __shared__ int koff[256];              // k = (dt*7 + dy)*7 + dx -> (dt*7 + dy)*97 + dx, -1 for the padding columns
    if (threadIdx.x < 256) {
        const int k = threadIdx.x;
        koff[k] = k < 245 ? (k / 7) * 97 + (k % 7) : -1;
    }
    const int f = f0 + blockIdx.x / 44, oy = blockIdx.x % 44;
    const int t = frame_t[f], T = frame_T[f];
    for (int i = threadIdx.x; i < 5 * 7 * 96; i += blockDim.x) {
        const int dt = i / (7 * 96), dy = (i / 96) % 7, xx = i % 96;
        const int tt = t + dt - 2, y = oy * 2 + dy - 3, x = xx - 3;
        float v = 0.f;
        if (tt >= 0 && tt < T && y >= 0 && y < 88 && x >= 0 && x < 88) v = video[((long long)(f + dt - 2) * 88 + y) * 88 + x];
        rows[dt][dy][xx] = v;
    }
    __syncthreads();
    // 8 consecutive k per thread = one 16-byte store; k -> offset in `rows` through a small table (no divisions per element)
    __nv_bfloat16* o = out + ((long long)(blockIdx.x) * 44) * 256;
    const float* rflat = &rows[0][0][0];
    for (int i = threadIdx.x; i < 44 * 32; i += blockDim.x) {
        const int ox = i >> 5, k0 = (i & 31) * 8;
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int off = koff[k0 + u];
            v[u] = off >= 0 ? rflat[off + ox * 2] : 0.f;
        }
        __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
        uint4 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
        pk.z = *reinterpret_cast<uint32_t*>(&p2); pk.w = *reinterpret_cast<uint32_t*>(&p3);
        *reinterpret_cast<uint4*>(o + (long long)ox * 256 + k0) = pk;
    }
}

// ------------------------------------------------------------------ 2D conv im2col (NHWC bf16)
// in [F,H,W,C] -> out [F*Ho*Wo, ks*ks*C], k = (ky*ks + kx)*C + c; pad = ks/2; 8 channels (16 B) per thread.
__global__ void __launch_bounds__(256)
im2col2d_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, long long total_vec, int H, int W, int C,
                int Ho, int Wo, int ks, int stride) {
    const int cv = C / 8;
    const int pad = ks / 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total_vec; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % cv);
        long long r = i / cv;
        const int kx = (int)(r % ks); r /= ks;
        const int ky = (int)(r % ks); r /= ks;
        const int ox = (int)(r % Wo); r /= Wo;
        const int oy = (int)(r % Ho);
        const long long f = r / Ho;
        const int y = oy * stride + ky - pad, x = ox * stride + kx - pad;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (y >= 0 && y < H && x >= 0 && x < W) v = *reinterpret_cast<const uint4*>(in + (((f * H + y) * W + x) * C + c8 * 8));
        *reinterpret_cast<uint4*>(out + i * 8) = v;
    }
}

// ------------------------------------------------------------------ max pool 3x3 s2 p1 (NHWC bf16), 8 channels/thread
__global__ void __launch_bounds__(256)
maxpool3x3s2_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, long long total_vec, int H, int W, int C,
                    int Ho, int Wo, long long out_row_pitch, long long out_frame_pitch) {
    const int cv = C / 8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total_vec; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % cv);
        long long r = i / cv;
        const int ox = (int)(r % Wo); r /= Wo;
        const int oy = (int)(r % Ho);
        const long long f = r / Ho;
        float m[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
        for (int dy = 0; dy < 3; ++dy) {
            const int y = oy * 2 + dy - 1;
            if (y < 0 || y >= H) continue;
            for (int dx = 0; dx < 3; ++dx) {
                const int x = ox * 2 + dx - 1;
                if (x < 0 || x >= W) continue;
                const uint4 v = *reinterpret_cast<const uint4*>(in + (((f * H + y) * W + x) * C + c8 * 8));
                const __nv_bfloat16* e = reinterpret_cast<const __nv_bfloat16*>(&v);
#pragma unroll
                for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], __bfloat162float(e[j]));
            }
        }
        __align__(16) __nv_bfloat16 o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = __float2bfloat16_rn(m[j]);
        *reinterpret_cast<uint4*>(out + ((f * out_frame_pitch + (long long)oy * out_row_pitch + ox) * C + c8 * 8)) = *reinterpret_cast<const uint4*>(o);
    }
}

// ------------------------------------------------------------------ global average pool: [F, HW, C] bf16 -> [F, C] bf16
__global__ void __launch_bounds__(256)
avgpool_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, long long F, int HW, int C) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= F * C) return;
    const long long f = i / C;
    const int c = (int)(i % C);
    float s = 0.f;
    for (int p = 0; p < HW; ++p) s += __bfloat162float(in[(f * HW + p) * C + c]);
    out[i] = __float2bfloat16_rn(s / (float)HW);
}

// ------------------------------------------------------------------ audio [B,104,T] fp32 -> packed [F,104] bf16
__global__ void __launch_bounds__(256)
audio_pack_kernel(const float* __restrict__ audio, __nv_bfloat16* __restrict__ out, const int* __restrict__ frame_b,
                  const int* __restrict__ frame_t, long long F, int Cin, int Tpad) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= F * Cin) return;
    const long long f = i / Cin;
    const int c = (int)(i % Cin);
    out[i] = __float2bfloat16_rn(audio[((long long)frame_b[f] * Cin + c) * Tpad + frame_t[f]]);
}

// ------------------------------------------------------------------ positional-conv im2col
// x bf16 [F,1024] -> out [16][F][128*64], out[g][f][j*64 + c] = x[f + j - 64][g*64 + c] if the source frame lies in
// the same utterance, else 0 (k=128, pad=64, last output dropped).  8 channels per thread.
__global__ void __launch_bounds__(256)
posconv_im2col_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, const int* __restrict__ frame_t,
                      const int* __restrict__ frame_T, long long F, int g0, int ng) {
    const long long per_g = F * 128 * 8;                    // uint4 vectors per group
    const long long total = per_g * ng;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % 8);
        long long r = i / 8;
        const int j = (int)(r % 128); r /= 128;
        const long long f = r % F;
        const int g = (int)(r / F);
        const int tt = frame_t[f] + j - 64;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (tt >= 0 && tt < frame_T[f]) v = *reinterpret_cast<const uint4*>(x + (f + j - 64) * 1024 + (g0 + g) * 64 + c8 * 8);
        *reinterpret_cast<uint4*>(out + i * 8) = v;
    }
}

// ------------------------------------------------------------------ fp32 -> bf16 cast (2D with leading dims)
__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float* __restrict__ in, long long ldi, __nv_bfloat16* __restrict__ out, long long ldo, long long rows, int cols) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const long long r = i / cols;
    const int c = (int)(i % cols);
    out[r * ldo + c] = __float2bfloat16_rn(in[r * ldi + c]);
}

// fp32 [rows, K] -> bf16 [rows, 6K] (bf16x3 activation layout, see avsr_split3_store)
__global__ void __launch_bounds__(256)
split3_kernel(const float* __restrict__ in, long long ldi, __nv_bfloat16* __restrict__ out, long long rows, int K) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= rows * K) return;
    const long long r = i / K;
    const int c = (int)(i % K);
    avsr_split3_store(out + r * 6 * K, K, c, in[r * ldi + c]);
}

// [F, ncol] -> [ncol/64][Fs][64] (Fs = rows each 64-column block has room for, >= F): float4 per thread.  kt_period > 0: 64-column blocks whose index b has (b / 16) % kt_period == 0
// (the K halves of [k | v] pairs of 16 heads) are written transposed in 32-byte groups, [b][8][F][8], the layout the decode
// step's attention reads keys in (csrc/dec_attn.cu).
__global__ void __launch_bounds__(256)
kv_head_major_kernel(const float* __restrict__ in, float* __restrict__ out, long long F, long long Fs, int ncol, int kt_period) {
    const long long total = F * (ncol / 4);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long f = i / (ncol / 4);
        const int c = (int)(i % (ncol / 4)) * 4;
        const float4 v = *reinterpret_cast<const float4*>(in + f * ncol + c);
        const int b = c / 64;
        if (kt_period > 0 && ((b / 16) % kt_period) == 0)
            *reinterpret_cast<float4*>(out + (long long)b * Fs * 64 + ((long long)((c % 64) >> 3) * Fs + f) * 8 + (c & 4)) = v;
        else
            *reinterpret_cast<float4*>(out + ((long long)b * Fs + f) * 64 + (c % 64)) = v;
    }
}

}  // namespace

#define GRID1D(total) ((int)(((total) + 255) / 256 > 148 * 64 ? 148 * 64 : ((total) + 255) / 256))

extern "C" int avsr_layernorm(const float* x, long long ldx, long long rows, int N, const float* gamma, const float* beta, float eps,
                              void* out_bf16, long long ld_bf16, float* out_f32, long long ld_f32, cudaStream_t stream) {
    AVSR_REQUIRE(x && gamma && beta && rows > 0 && N > 0 && (N & 3) == 0 && (ldx & 3) == 0 && N * 4 <= 48 * 1024,
                 "avsr_layernorm: bad arguments (rows=%lld N=%d)", rows, N);
    AVSR_REQUIRE(out_bf16 || out_f32, "avsr_layernorm: no output");
    const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta) |
                           reinterpret_cast<uintptr_t>(out_f32)) & 15) == 0 && (reinterpret_cast<uintptr_t>(out_bf16) & 7) == 0 &&
                         (ld_f32 & 3) == 0 && (ld_bf16 & 3) == 0;
    if (N == 1024 && aligned)
        layernorm1024_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(x, ldx, rows, gamma, beta, eps, (__nv_bfloat16*)out_bf16, ld_bf16,
                                                                           out_f32, ld_f32);
    else
        layernorm_kernel<<<(unsigned)rows, 256, (size_t)N * 4, stream>>>(x, ldx, N, gamma, beta, eps, (__nv_bfloat16*)out_bf16, ld_bf16,
                                                                        out_f32, ld_f32);
    AVSR_LAUNCH_CHECK();
    return AVSR_OK;
}

// Frames [f0, f0+nf) of the packed video -> out [nf*1936, 256] bf16.
extern "C" int avsr_im2col_frontend(const float* video, const int* frame_t, const int* frame_T, int f0, int nf, void* out,
                                    cudaStream_t stream) {
    AVSR_REQUIRE(video && frame_t && frame_T && out && nf > 0, "avsr_im2col_frontend: bad arguments");
    im2col_frontend_kernel<<<nf * 44, 256, 0, stream>>>(video, frame_t, frame_T, (__nv_bfloat16*)out, f0);
    AVSR_LAUNCH_CHECK();
    return AVSR_OK;
}

extern "C" int avsr_im2col2d(const void* in, void* out, long long F, int H, int W, int C, int ks, int stride, cudaStream_t stream) {
    AVSR_REQUIRE(in && out && F > 0 && (C & 7) == 0 && (ks == 1 || ks == 3) && (stride == 1 || stride == 2),
                 "avsr_im2col2d: bad arguments");
    const int pad = ks / 2;
    const int Ho = (H + 2 * pad - ks) / stride + 1, Wo = (W + 2 * pad - ks) / stride + 1;
    const long long total = F * Ho * Wo * ks * ks * (C / 8);
    im2col2d_kernel<<<GRID1D(total), 256, 0, stream>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, total, H, W, C, Ho, Wo, ks, stride);
    AVSR_LAUNCH_CHECK();
    return AVSR_OK;
}

// out_row_pitch_px / out_frame_pitch_px: pixel pitches of the output rows / frames (0 = dense); only the Ho x Wo valid pixels of
// a frame are written (a padded layout keeps its zero pads).
extern "C" int avsr_maxpool3x3s2_pitched(const void* in, void* out, long long F, int H, int W, int C, long long out_row_pitch_px,
                                         long long out_frame_pitch_px, cudaStream_t stream) {
    AVSR_REQUIRE(in && out && F > 0 && (C & 7) == 0, "avsr_maxpool3x3s2: bad arguments");
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const long long rp = out_row_pitch_px > 0 ? out_row_pitch_px : Wo, fp = out_frame_pitch_px > 0 ? out_frame_pitch_px : rp * Ho;
    AVSR_REQUIRE(rp >= Wo && fp >= rp * (Ho - 1) + Wo, "avsr_maxpool3x3s2: output pitches too small");
    const long long total = F * Ho * Wo * (C / 8);
    maxpool3x3s2_kernel<<<GRID1D(total), 256, 0, stream>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, total, H, W, C, Ho, Wo, rp, fp);
    AVSR_LAUNCH_CHECK();
    return AVSR_OK;
}

extern "C" int avsr_maxpool3x3s2(const void* in, void* out, long long F, int H, int W, int C, cudaStream_t stream) {
    return avsr_maxpool3x3s2_pitched(in, out, F, H, W, C, 0, 0, stream);
}

extern "C" int avsr_avgpool(const void* in, void* out, long long F, int HW, int C, cudaStream_t stream) {
    AVSR_REQUIRE(in && out && F > 0, "avsr_avgpool: bad arguments");
    avgpool_kernel<<<cdiv(F * C, 256), 256, 0, stream>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, F, HW, C);
    AVSR_LAUNCH_CHECK();
    return AVSR_OK;
}

extern "C" int avsr_audio_pack(const float* audio, void* out, const int* frame_b, const int* frame_t, long long F, int Cin, int Tpad,
                               cudaStream_t stream) {
    AVSR_REQUIRE(audio && out && frame_b && frame_t && F > 0, "avsr_audio_pack: bad arguments");
    audio_pack_kernel<<<cdiv(F * Cin, 256), 256, 0, stream>>>(audio, (__nv_bfloat16*)out, frame_b, frame_t, F, Cin, Tpad);
    AVSR_LAUNCH_CHECK();
    return AVSR_OK;
}

extern "C" int avsr_posconv_im2col(const void* x, void* out, const int* frame_t, const int* frame_T, long long F, int g0, int ng,
                                   cudaStream_t stream) {
    AVSR_REQUIRE(x && out && frame_t && frame_T && F > 0 && g0 >= 0 && ng > 0 && g0 + ng <= 16, "avsr_posconv_im2col: bad arguments");
    const long long total = F * 128 * 8 * ng;
    posconv_im2col_kernel<<<GRID1D(total), 256, 0, stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, frame_t, frame_T, F, g0, ng);
    AVSR_LAUNCH_CHECK();
    return AVSR_OK;
}

extern "C" int avsr_cast_bf16(const float* in, long long ldi, void* out, long long ldo, long long rows, int cols, cudaStream_t stream) {
    AVSR_REQUIRE(in && out && rows > 0 && cols > 0, "avsr_cast_bf16: bad arguments");
    cast_bf16_kernel<<<cdiv(rows * cols, 256), 256, 0, stream>>>(in, ldi, (__nv_bfloat16*)out, ldo, rows, cols);
    AVSR_LAUNCH_CHECK();
    return AVSR_OK;
}

extern "C" int avsr_split3(const float* in, long long ldi, void* out, long long rows, int K, cudaStream_t stream) {
    AVSR_REQUIRE(in && out && rows > 0 && K > 0, "avsr_split3: bad arguments");
    split3_kernel<<<cdiv(rows * K, 256), 256, 0, stream>>>(in, ldi, (__nv_bfloat16*)out, rows, K);
    AVSR_LAUNCH_CHECK();
    return AVSR_OK;
}

extern "C" int avsr_kv_head_major(const float* in, float* out, long long F, long long F_capacity, int ncol, int k_transposed,
                                  cudaStream_t stream) {
    AVSR_REQUIRE(in && out && F > 0 && F_capacity >= F && ncol > 0 && (ncol & 63) == 0, "avsr_kv_head_major: bad arguments");
    AVSR_REQUIRE(!k_transposed || (ncol % 2048) == 0, "avsr_kv_head_major: k_transposed needs [k(1024) | v(1024)] column pairs");
    const long long total = F * (ncol / 4);
    kv_head_major_kernel<<<GRID1D(total), 256, 0, stream>>>(in, out, F, F_capacity, ncol, k_transposed ? 2 : 0);
    AVSR_LAUNCH_CHECK();
    return AVSR_OK;
}
