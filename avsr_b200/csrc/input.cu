// Input pipeline right before the hot path: waveform -> stacked, layer-normed log mel filterbank features,
// uint8 mouth crops -> normalised fp32 frames, both written straight into the zero-padded batch layout the reference's
// DataCollator hands to the encoder (/root/reference/src/dataset/avhubert_dataset.py:86-116, 225-246, 277-349).
//
// The reference computes the filterbank on the CPU with python_speech_features 0.6 (numpy, float64 after a float32
// pre-emphasis).  Here one CTA produces one output row = 4 stacked frames: the four 400-sample frames are zero-padded to
// 512, transformed by a shared-memory FFT (radix-2 stages fused in pairs) in fp64 (B200 keeps a full-rate fp64 pipe; the whole batch is ~1 GFLOP),
// reduced through the 26 triangular mel filters, floored at DBL_EPSILON, logged, rounded to fp32 exactly where the
// reference rounds, layer-normed over the 104 values and stored transposed into [B][104][Tmax].  Bytes per row: 3.5 KB of
// samples in, 416 B out: latency-bound, sized by the number of resident CTAs, not by HBM.
#include <float.h>
#include <math.h>

#include "common.cuh"

namespace {

constexpr int FRAME_LEN = 400, FRAME_STEP = 160, NFFT = 512, NFILT = 26, STACK = 4, NFEAT = NFILT * STACK;
constexpr int CROP = 88;

__constant__ int c_bins[NFILT + 2];
__device__ double2 g_twiddle[NFFT / 2];       // exp(-2 pi i k / 512), written once from the host (libm cos / sin)
int h_bins[NFILT + 2];
bool bins_ready = false;

// psf base.get_filterbanks: floor((nfft + 1) * mel2hz(linspace(hz2mel(0), hz2mel(8000), 28)) / 16000)
void compute_bins(int* bins) {
    const double lowmel = 2595.0 * log10(1.0 + 0.0 / 700.0), highmel = 2595.0 * log10(1.0 + 8000.0 / 700.0);
    const double step = (highmel - lowmel) / (NFILT + 1);
    for (int i = 0; i < NFILT + 2; ++i) {
        const double mel = (i == NFILT + 1) ? highmel : lowmel + i * step;
        const double hz = 700.0 * (pow(10.0, mel / 2595.0) - 1.0);
        bins[i] = (int)floor((NFFT + 1) * hz / 16000.0);
    }
}

int ensure_bins() {
    if (!bins_ready) {
        compute_bins(h_bins);
        AVSR_CHECK_CUDA(cudaMemcpyToSymbol(c_bins, h_bins, sizeof(h_bins)));
        static double2 tw[NFFT / 2];
        const double pi = 3.14159265358979323846;
        for (int k = 0; k < NFFT / 2; ++k) tw[k] = make_double2(cos(2.0 * pi * k / NFFT), -sin(2.0 * pi * k / NFFT));
        tw[NFFT / 4] = make_double2(0.0, -1.0);          // exact quarter turn
        AVSR_CHECK_CUDA(cudaMemcpyToSymbol(g_twiddle, tw, sizeof(tw)));
        bins_ready = true;
    }
    return AVSR_OK;
}

__host__ __device__ inline int frames_of(int n) { return n <= FRAME_LEN ? 1 : 1 + (n - FRAME_LEN + FRAME_STEP - 1) / FRAME_STEP; }

__global__ void __launch_bounds__(256) fbank_stack_ln_kernel(const float* __restrict__ wave, const long long* __restrict__ wave_off,
                                                             const int* __restrict__ wave_len, const int* __restrict__ n_samples,
                                                             int Tmax, float* __restrict__ out) {
    __shared__ double2 buf[STACK][NFFT];       // 32 KB: the four frames' spectra, then (in .x) their power spectra
    __shared__ double2 tw[NFFT / 2];           // exp(-2 pi i k / 512)
    __shared__ float feat[NFEAT];
    __shared__ double stats[2];
    const int tid = threadIdx.x, r = blockIdx.x, b = blockIdx.y;
    const int n = n_samples[b];
    const int nf = frames_of(n), rows = (nf + STACK - 1) / STACK;
    float* o = out + (size_t)b * NFEAT * Tmax + r;
    if (r >= rows) {                           // collate_pad: zero rows behind the utterance
        if (tid < NFEAT) o[(size_t)tid * Tmax] = 0.f;
        return;
    }
    const int navail = min(wave_len[b], n);    // cut_or_pad: samples past the waveform read as zero, past n do not exist
    const float* w = wave + wave_off[b];
    tw[tid] = g_twiddle[tid];
    for (int i = tid; i < STACK * NFFT; i += 256) {
        const int fr = i >> 9, k = i & (NFFT - 1), f = STACK * r + fr;
        double v = 0.0;
        const long long p = (long long)f * FRAME_STEP + k;
        if (k < FRAME_LEN && f < nf && p < n) {
            const float x = p < navail ? w[p] : 0.f;
            if (p == 0) {
                v = x;
            } else {                           // float32 pre-emphasis, two roundings like numpy (no fma)
                const float xp = (p - 1) < navail ? w[p - 1] : 0.f;
                v = __fsub_rn(x, __fmul_rn(0.97f, xp));
            }
        }
        buf[fr][__brev((unsigned)k) >> 23] = make_double2(v, 0.0);
    }
    __syncthreads();
    // Radix-2 decimation-in-time stages fused in pairs: a thread owns the four points {a, a+h, a+2h, a+3h} of two frames,
    // applies stage s to (a, a+h), (a+2h, a+3h) and stage s+1 to (a, a+2h), (a+h, a+3h) in registers, so the 32 KB of
    // spectra cross shared memory five times instead of nine (the kernel is bound by that traffic, profiles/ncu_r01_input.txt).
    // Same operations in the same order as the unfused stages: bit-identical results.
    {
        const int pos = tid & 127, f0 = (tid >> 7) * 2;
#pragma unroll 1
        for (int s = 1; s <= 7; s += 2) {
            const int h = 1 << (s - 1), k = pos & (h - 1);
            const int a = ((pos >> (s - 1)) << (s + 1)) + k;
            const double2 w1 = tw[k << (9 - s)], w2 = tw[k << (8 - s)], w3 = tw[(k + h) << (8 - s)];
#pragma unroll
            for (int fr = f0; fr < f0 + 2; ++fr) {
                const double2 x0 = buf[fr][a], x1 = buf[fr][a + h], x2 = buf[fr][a + 2 * h], x3 = buf[fr][a + 3 * h];
                double tr = w1.x * x1.x - w1.y * x1.y, ti = w1.x * x1.y + w1.y * x1.x;
                const double2 y0 = make_double2(x0.x + tr, x0.y + ti), y1 = make_double2(x0.x - tr, x0.y - ti);
                tr = w1.x * x3.x - w1.y * x3.y, ti = w1.x * x3.y + w1.y * x3.x;
                const double2 y2 = make_double2(x2.x + tr, x2.y + ti), y3 = make_double2(x2.x - tr, x2.y - ti);
                tr = w2.x * y2.x - w2.y * y2.y, ti = w2.x * y2.y + w2.y * y2.x;
                buf[fr][a] = make_double2(y0.x + tr, y0.y + ti);
                buf[fr][a + 2 * h] = make_double2(y0.x - tr, y0.y - ti);
                tr = w3.x * y3.x - w3.y * y3.y, ti = w3.x * y3.y + w3.y * y3.x;
                buf[fr][a + h] = make_double2(y1.x + tr, y1.y + ti);
                buf[fr][a + 3 * h] = make_double2(y1.x - tr, y1.y - ti);
            }
            __syncthreads();
        }
        // last stage (half = 256): thread = butterfly position, all four frames
        const double2 t = tw[tid];
#pragma unroll
        for (int fr = 0; fr < STACK; ++fr) {
            const double2 x = buf[fr][tid + NFFT / 2], y = buf[fr][tid];
            const double tr = t.x * x.x - t.y * x.y, ti = t.x * x.y + t.y * x.x;
            buf[fr][tid + NFFT / 2] = make_double2(y.x - tr, y.y - ti);
            buf[fr][tid] = make_double2(y.x + tr, y.y + ti);
        }
        __syncthreads();
    }
    for (int i = tid; i < STACK * (NFFT / 2 + 1); i += 256) {
        const int fr = i / (NFFT / 2 + 1), k = i - fr * (NFFT / 2 + 1);
        const double2 x = buf[fr][k];
        buf[fr][k].x = (1.0 / NFFT) * (x.x * x.x + x.y * x.y);
    }
    __syncthreads();
    if (tid < NFEAT) {
        const int fr = tid / NFILT, j = tid - fr * NFILT;
        float v = 0.f;                         // stacker: frames past the last one are zero rows
        if (STACK * r + fr < nf) {
            const int b0 = c_bins[j], b1 = c_bins[j + 1], b2 = c_bins[j + 2];
            double acc = 0.0;
            for (int i = b0; i < b1; ++i) acc += buf[fr][i].x * ((double)(i - b0) / (double)(b1 - b0));
            for (int i = b1; i < b2; ++i) acc += buf[fr][i].x * ((double)(b2 - i) / (double)(b2 - b1));
            if (acc == 0.0) acc = DBL_EPSILON;
            v = (float)log(acc);
        }
        feat[tid] = v;
    }
    __syncthreads();
    if (tid < 32) {                            // F.layer_norm over the 104 values, no affine, eps 1e-5
        double s = 0.0;
        for (int i = tid; i < NFEAT; i += 32) s += feat[i];
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        const double mean = s / NFEAT;
        double ss = 0.0;
        for (int i = tid; i < NFEAT; i += 32) {
            const double d = feat[i] - mean;
            ss += d * d;
        }
        for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
        if (tid == 0) {
            stats[0] = mean;
            stats[1] = rsqrt(ss / NFEAT + 1e-5);
        }
    }
    __syncthreads();
    if (tid < NFEAT) o[(size_t)tid * Tmax] = (float)((feat[tid] - stats[0]) * stats[1]);
}

// VideoTransform('test'): x / 255 -> CenterCrop(88) -> (x - 0.421) / 0.165, float32 with the reference's three roundings.
// Only 256 inputs exist, so each CTA builds the table once and the kernel is a byte -> float4 stream: 7.7 KB read (the crop
// of a 9.2 KB frame) and 31 KB written per frame, HBM-bound on the writes.
template <bool ALIGNED>
__global__ void __launch_bounds__(256) video_u8_kernel(const unsigned char* __restrict__ frames, const long long* __restrict__ frame_off,
                                                        const int* __restrict__ utt_T, int Tmax, int H, int W, int top, int left,
                                                        float* __restrict__ out) {
    __shared__ float lut[256];
    const int tid = threadIdx.x, b = blockIdx.y;
    lut[tid] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)tid, 255.f), 0.421f), 0.165f);
    __syncthreads();
    const int T = utt_T[b];
    for (int t = blockIdx.x; t < Tmax; t += gridDim.x) {
        float4* o = reinterpret_cast<float4*>(out + ((size_t)b * Tmax + t) * (CROP * CROP));
        if (t >= T) {
            for (int i = tid; i < CROP * CROP / 4; i += 256) o[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            continue;
        }
        const unsigned char* in = frames + (size_t)(frame_off[b] + t) * H * W + (size_t)top * W + left;
        for (int i = tid; i < CROP * CROP / 4; i += 256) {
            const int y = i / (CROP / 4), x = (i - y * (CROP / 4)) * 4;
            const unsigned char* p = in + (size_t)y * W + x;
            uchar4 u;
            if (ALIGNED) {
                u = *reinterpret_cast<const uchar4*>(p);
            } else {
                u = make_uchar4(p[0], p[1], p[2], p[3]);
            }
            o[i] = make_float4(lut[u.x], lut[u.y], lut[u.z], lut[u.w]);
        }
    }
}


// torchaudio.functional.add_noise (the mixing AddMultiSpk / AddNoise do, avhubert_dataset.py:160-222): per utterance
// scale = 10^((10 (log10 Es - log10 En) - snr) / 20), out = wave + scale * noise.  Energies over the first lengths[b] samples.
// One CTA per utterance and signal accumulates in fp64 in a fixed order (deterministic); the mix is one streaming pass.
__global__ void __launch_bounds__(1024) wave_energy_kernel(const float* __restrict__ wave, const float* __restrict__ noise,
                                                           const int* __restrict__ lengths, long long L, double* __restrict__ energy) {
    __shared__ double red[32];
    const int b = blockIdx.x, which = blockIdx.y, tid = threadIdx.x;
    const float* x = (which ? noise : wave) + (size_t)b * L;
    const long long n = lengths ? min((long long)lengths[b], L) : L;
    double acc = 0.0;
    for (long long i = tid; i < n; i += 1024) {
        const double v = x[i];
        acc += v * v;
    }
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((tid & 31) == 0) red[tid >> 5] = acc;
    __syncthreads();
    if (tid < 32) {
        acc = red[tid];
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (tid == 0) energy[2 * b + which] = acc;
    }
}

__global__ void __launch_bounds__(256) wave_mix_kernel(const float* __restrict__ wave, const float* __restrict__ noise,
                                                       const float* __restrict__ snr_db, const double* __restrict__ energy, long long L,
                                                       float* __restrict__ out) {
    const int b = blockIdx.y;
    // float32 from here on, like the torch expression: vector_norm(.)**2, log10, 10 ** (. / 20)
    const float ns = (float)sqrt(energy[2 * b]), nn = (float)sqrt(energy[2 * b + 1]);
    const float snr0 = 10.f * (log10f(ns * ns) - log10f(nn * nn));
    const float scale = powf(10.f, (snr0 - snr_db[b]) / 20.f);
    const float* x = wave + (size_t)b * L;
    const float* z = noise + (size_t)b * L;
    float* o = out + (size_t)b * L;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < L; i += (long long)gridDim.x * 256)
        o[i] = __fadd_rn(x[i], __fmul_rn(scale, z[i]));
}

}  // namespace

extern "C" int avsr_fbank_rows(int n_samples) {
    if (n_samples < 1) return 0;
    return (frames_of(n_samples) + STACK - 1) / STACK;
}

extern "C" int avsr_fbank_bins(int* bins28) {
    AVSR_REQUIRE(bins28, "avsr_fbank_bins: null output");
    compute_bins(bins28);
    return AVSR_OK;
}

extern "C" int avsr_fbank_stack_ln(const float* wave, const long long* wave_off, const int* wave_len, const int* n_samples, int B, int Tmax,
                                   float* out, cudaStream_t stream) {
    AVSR_REQUIRE(wave && wave_off && wave_len && n_samples && out && B > 0 && Tmax > 0 && B <= 65535, "avsr_fbank_stack_ln: bad arguments");
    const int rc = ensure_bins();
    if (rc != AVSR_OK) return rc;
    fbank_stack_ln_kernel<<<dim3(Tmax, B), 256, 0, stream>>>(wave, wave_off, wave_len, n_samples, Tmax, out);
    AVSR_LAUNCH_CHECK();
    return AVSR_OK;
}

extern "C" int avsr_video_u8_transform(const unsigned char* frames, const long long* frame_off, const int* utt_T, int B, int Tmax, int H, int W,
                                       float* out, cudaStream_t stream) {
    AVSR_REQUIRE(frames && frame_off && utt_T && out && B > 0 && Tmax > 0 && B <= 65535, "avsr_video_u8_transform: bad arguments");
    AVSR_REQUIRE(H >= CROP && W >= CROP, "avsr_video_u8_transform: frames of %dx%d are smaller than the %d crop", H, W, CROP);
    // torchvision center_crop: int(round((H - 88) / 2.0)) with Python's round-half-to-even
    auto half_even = [](int d) { return (d & 1) ? ((d / 2) & 1 ? d / 2 + 1 : d / 2) : d / 2; };
    const int top = half_even(H - CROP), left = half_even(W - CROP);
    const bool aligned = (W % 4 == 0) && (left % 4 == 0) && ((uintptr_t)frames % 4 == 0);
    // a few frames per CTA so that the table is built ~once per 4 frames; 148 SMs x 8 resident CTAs cover the grid in waves
    const dim3 grid(cdiv(Tmax, 4), B);
    if (aligned)
        video_u8_kernel<true><<<grid, 256, 0, stream>>>(frames, frame_off, utt_T, Tmax, H, W, top, left, out);
    else
        video_u8_kernel<false><<<grid, 256, 0, stream>>>(frames, frame_off, utt_T, Tmax, H, W, top, left, out);
    AVSR_LAUNCH_CHECK();
    return AVSR_OK;
}

extern "C" int avsr_add_noise(const float* wave, const float* noise, const float* snr_db, const int* lengths, int B, long long L, float* out,
                              double* energy, cudaStream_t stream) {
    AVSR_REQUIRE(wave && noise && snr_db && out && energy && B > 0 && B <= 65535 && L > 0, "avsr_add_noise: bad arguments");
    wave_energy_kernel<<<dim3(B, 2), 1024, 0, stream>>>(wave, noise, lengths, L, energy);
    AVSR_LAUNCH_CHECK();
    const int gx = (int)((L + 256 * 8 - 1) / (256 * 8));
    wave_mix_kernel<<<dim3(gx < 1 ? 1 : gx, B), 256, 0, stream>>>(wave, noise, snr_db, energy, L, out);
    AVSR_LAUNCH_CHECK();
    return AVSR_OK;
}
