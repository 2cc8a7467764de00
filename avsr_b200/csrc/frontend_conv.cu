// Lip-reading frontend: Conv3d(1 -> 64, kernel 5x7x7, stride 1x2x2, padding 2x3x3, no bias) + BatchNorm3d (eval, folded) +
// PReLU(64) as an IMPLICIT GEMM on the tcgen05 tensor cores (src/nets/backend/backbones/resnet.py:132-135, called from
// ResEncoder.forward :151-153).  Round 1 materialised the 5x7x7 patches as a [F*1936, 256] bf16 matrix (11.9 GB written and
// re-read per 32-utterance pass for 372 MB of input: 8.1 ms, 18 % of the encoder); here nothing is materialised.
//
//   out[(f, oy, ox), c] = PReLU_c( sum_{dt,dy,dx} video[f + dt - 2][2 oy + dy - 3][2 ox + dx - 3] * W[c][dt][dy][dx] + b[c] )
//
// with zeros outside the frame and outside the UTTERANCE (frames are packed back to back; frame_t / frame_T give a frame's
// position in its utterance, so the temporal halo never reaches into a neighbour).
//
// GEMM view: M = output pixels (128 per tile, 16 tiles per 44x44 frame), N = 64 channels, K = 5*7*8 = 280 padded to 320:
// k = (dt*7 + dy)*8 + dx with dx = 7 a zero weight, so that the 8 k's of a (dt, dy) pair are ONE 16-byte chunk of the A
// tile = 8 consecutive input pixels, and a chunk is copied with four aligned 32-bit shared-memory loads and one 16-byte
// store.  One persistent CTA per SM:
//   warp 0      : MMA issuer (tcgen05.mma M128 x N64 x K16, 20 per tile), owns TMEM (2 accumulators x 64 columns)
//   warps 1-8   : builders.  Per tile the A operand is built k block by k block in the 128-byte-swizzled UMMA layout from the
//                 bf16 input window in shared memory.  The five k blocks of A are a ring (slot j = k block j): block j of the
//                 next tile is rebuilt as soon as the MMAs that read block j of this tile have retired.
//   warps 9-12  : epilogue.  tcgen05.ld (thread = pixel), bias + PReLU, 64 bf16 channels = one 128-byte store per pixel (NHWC).
//   warps 13-16 : loaders.  The fp32 input window of a tile (5 frames x <= 13 rows x 88, zero borders) is fetched one tile
//                 ahead into registers and stored as bf16 into a three-slot ring of windows (mbarrier hand-off to the builders).
//                 They are separate warps because the builders execute fence.proxy.async (MEMBAR.ALL.CTA) after every k
//                 block, which would wait for global loads in flight in the same thread (measured: 3.3 -> 3.0 us per tile).  What is left
//                 is shared-memory bandwidth: per tile 72 KB of window reads + 72 KB of A stores by the builders and 120 KB of
//                 operand reads by the 20 MMAs (N = 64: 6 KB per 32-cycle MMA), i.e. ~2300 cycles at 128 B / clock.  (One
//                 proxy fence per tile instead of one per k block measured the same.)
// The 40 KB of weights stay in shared memory for the whole kernel.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int IMG = 88, OUT = 44, PIX = OUT * OUT;            // input frame, conv output, pixels per frame
constexpr int TILE = 128;
constexpr int TILES_PER_FRAME = (PIX + TILE - 1) / TILE;      // 16 (the last one holds 16 pixels)
constexpr int KB = 5;                                         // k blocks of 64
constexpr int KTOT = KB * 64;                                 // 320
constexpr int NCH = 64;
constexpr int WROWS = 13, WPITCH = 96;                        // input window rows / bf16 per row (x = -3 .. 92)
constexpr int WIN_ELEMS = 5 * WROWS * WPITCH;                 // 6240
constexpr int A_KB_BYTES = TILE * 128;                        // 16 KB per k block
constexpr int B_KB_BYTES = NCH * 128;                         // 8 KB per k block
constexpr int N_PROD = 256, N_EPI = 128, N_LOAD = 128;
constexpr int NUM_THREADS = 32 + N_PROD + N_EPI + N_LOAD;
constexpr int WIN_SLOTS = 3;
constexpr int WIN_PER_THREAD = (WIN_ELEMS / 2 + N_LOAD - 1) / N_LOAD;      // bf16 pairs per loader thread: 25
constexpr int SMEM_BYTES = KB * A_KB_BYTES + KB * B_KB_BYTES + WIN_SLOTS * WIN_ELEMS * 2 + 2 * NCH * 4 + 256 + 1024;

__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t i = 0; !tc::mbar_try_wait(bar, parity); ++i)
        if (i > (1u << 28)) __trap();
}

struct FrontArgs {
    const float* video;            // [F][88][88] fp32, packed frames
    const int* frame_t;            // [F] position of the frame in its utterance
    const int* frame_T;            // [F] length of that utterance
    const __nv_bfloat16* w;        // [64][320] bf16, k = (dt*7 + dy)*8 + dx (BN folded)
    const float* bias;             // [64] folded BN shift
    const float* prelu;            // [64]
    __nv_bfloat16* out;            // [nf][44][44][64] bf16
    int f0, nf;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
frontend_conv_kernel(const FrontArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                                            // [KB][128 rows][128 B]
    uint8_t* sB = sA + KB * A_KB_BYTES;                            // [KB][64 rows][128 B]
    __nv_bfloat16* win = reinterpret_cast<__nv_bfloat16*>(sB + KB * B_KB_BYTES);       // [WIN_SLOTS][5][13][96]
    float* s_bias = reinterpret_cast<float*>(win + WIN_SLOTS * WIN_ELEMS);
    float* s_prelu = s_bias + NCH;
    uint64_t* afull = reinterpret_cast<uint64_t*>(s_prelu + NCH);
    uint64_t* aempty = afull + KB;
    uint64_t* tfull = aempty + KB;
    uint64_t* tempty = tfull + 2;
    uint64_t* wfull = tempty + 2;
    uint64_t* wempty = wfull + WIN_SLOTS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wempty + WIN_SLOTS);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_items = a.nf * TILES_PER_FRAME;

    if (tid == 0) {
        for (int j = 0; j < KB; ++j) { tc::mbar_init(&afull[j], N_PROD / 32); tc::mbar_init(&aempty[j], 1); }
        for (int j = 0; j < 2; ++j) { tc::mbar_init(&tfull[j], 1); tc::mbar_init(&tempty[j], N_EPI / 32); }
        for (int j = 0; j < WIN_SLOTS; ++j) { tc::mbar_init(&wfull[j], N_LOAD / 32); tc::mbar_init(&wempty[j], N_PROD / 32); }
        tc::fence_barrier_init();
    }
    if (warp == 0) {
        tc::tmem_alloc(tmem_slot, 128);
        tc::tmem_relinquish();
    }
    // weights -> shared memory in the UMMA layout (row = channel, 16-byte chunk c of row n at n*128 + ((c ^ (n & 7)) << 4));
    // the A ring is zeroed once: the chunks of the padding k's (280 .. 319) are never written again
    for (int i = tid; i < KB * NCH * 8; i += NUM_THREADS) {
        const int kb = i / (NCH * 8), n = (i / 8) % NCH, c = i % 8;
        const uint4 v = *reinterpret_cast<const uint4*>(a.w + (long long)n * KTOT + kb * 64 + c * 8);
        *reinterpret_cast<uint4*>(sB + kb * B_KB_BYTES + n * 128 + ((c ^ (n & 7)) << 4)) = v;
    }
    for (int i = tid; i < KB * A_KB_BYTES / 16; i += NUM_THREADS) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
    if (tid < NCH) { s_bias[tid] = a.bias[tid]; s_prelu[tid] = a.prelu[tid]; }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = tc::umma_idesc_bf16(TILE, NCH);
            uint32_t ph_a = 0, ph_t = 0;
            int acc = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                mbar_wait(&tempty[acc], ph_t ^ 1);
                tc::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * NCH;
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(&afull[kb], ph_a);
                    tc::tc_fence_after();
                    const uint64_t adesc = tc::umma_desc_sw128(tc::smem_u32(sA + kb * A_KB_BYTES));
                    const uint64_t bdesc = tc::umma_desc_sw128(tc::smem_u32(sB + kb * B_KB_BYTES));
#pragma unroll
                    for (int k = 0; k < 4; ++k) tc::umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    tc::umma_commit(&aempty[kb]);          // slot kb may be rebuilt for the next tile once these MMAs retire
                }
                tc::umma_commit(&tfull[acc]);
                ph_a ^= 1;
                if (++acc == 2) { acc = 0; ph_t ^= 1; }
            }
        }
    } else if (warp <= N_PROD / 32) {
        // ------------------------------------------------------------------------------------------------ builders
        const int pt = tid - 32;                               // 0 .. 255
        const int p = pt & 127, half = pt >> 7;                // pixel of the tile, which four chunks of a k block
        // shared-memory word offset of the (dt, dy) pair behind each of this thread's 20 chunks (tile independent)
        int coff[KB][4];
#pragma unroll
        for (int kb = 0; kb < KB; ++kb)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const int q = kb * 8 + half * 4 + cc;
                coff[kb][cc] = q < 35 ? ((q / 7) * WROWS + (q % 7)) * (WPITCH / 2) * 4 : -1;       // bytes
            }
        const uint32_t sA_s = tc::smem_u32(sA) + (uint32_t)p * 128u, win_s = tc::smem_u32(win);
        const uint32_t swz = (uint32_t)(p & 7);
        uint32_t ph_a = 0, ph_w = 0;
        int slot = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int ti = item % TILES_PER_FRAME;
            const int P = ti * TILE + p;                       // pixel of the frame
            const bool pok = P < PIX;
            const int oy = P / OUT, ox = P - oy * OUT;
            const int oy0 = (ti * TILE) / OUT;
            // byte address of this pixel's first input column inside a window row: x = 2 ox - 3 -> xx = 2 ox -> word ox
            const uint32_t wpix = win_s + (uint32_t)(slot * WIN_ELEMS * 2) + (uint32_t)(((2 * (oy - oy0)) * (WPITCH / 2) + ox) * 4);
            mbar_wait(&wfull[slot], ph_w);                     // the loaders have finished this tile's window
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) {
                mbar_wait(&aempty[kb], ph_a ^ 1);
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    if (coff[kb][cc] >= 0 && pok) {
                        const uint32_t src = wpix + (uint32_t)coff[kb][cc];
                        const uint32_t v0 = lds32(src), v1 = lds32(src + 4), v2 = lds32(src + 8), v3 = lds32(src + 12);
                        sts128(sA_s + (uint32_t)(kb * A_KB_BYTES) + ((((uint32_t)(half * 4 + cc)) ^ swz) << 4), v0, v1, v2, v3);
                    }
                }
                tc::fence_proxy_async();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&afull[kb]);
            }
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&wempty[slot]);     // this window slot may be refilled
            ph_a ^= 1;
            if (++slot == WIN_SLOTS) { slot = 0; ph_w ^= 1; }
        }
    } else if (warp > (N_PROD + N_EPI) / 32) {
        // ------------------------------------------------------------------------------------------------ loaders
        const int lt = tid - 32 - N_PROD - N_EPI;              // 0 .. 127
        // input window of a tile: frames f-2 .. f+2, input rows iy0 .. iy0 + 12, x = -3 .. 92 (zero outside the image / utterance).
        // A thread always fetches the same WIN_PER_THREAD element pairs of the window: their (dt, row, column) are unpacked once.
        int wdesc[WIN_PER_THREAD];                             // dt | row << 8 | xx << 16, or -1 past the end of the window
#pragma unroll
        for (int u = 0; u < WIN_PER_THREAD; ++u) {
            const int e = (lt + u * N_LOAD) * 2;
            wdesc[u] = e < WIN_ELEMS ? ((e / (WROWS * WPITCH)) | (((e / WPITCH) % WROWS) << 8) | ((e % WPITCH) << 16)) : -1;
        }
        float2 wreg[WIN_PER_THREAD];
        auto window_fetch = [&](int item, int t, int T) {
            const int fl = item / TILES_PER_FRAME, ti = item - fl * TILES_PER_FRAME;
            const int f = a.f0 + fl;
            const int iy0 = 2 * ((ti * TILE) / OUT) - 3;
            const float* fbase = a.video + (long long)(f - 2) * IMG * IMG - 3;
#pragma unroll
            for (int u = 0; u < WIN_PER_THREAD; ++u) {
                const int d = wdesc[u];
                const int dt = d & 0xff, r = (d >> 8) & 0xff, xx = d >> 16;
                const int tt = t + dt - 2, y = iy0 + r;
                float2 v = make_float2(0.f, 0.f);
                if (d >= 0 && tt >= 0 && tt < T && y >= 0 && y < IMG) {
                    const float* src = fbase + (dt * IMG + y) * IMG + xx;       // &video[f + dt - 2][y][xx - 3]
                    if (xx >= 3 && xx < IMG + 3) v.x = __ldg(src);
                    if (xx >= 2 && xx < IMG + 2) v.y = __ldg(src + 1);
                }
                wreg[u] = v;
            }
        };
        // a frame's position / utterance length are fetched TWO items ahead, so that no tile waits for them
        auto frame_info = [&](int item, int& t, int& T) {
            t = 0; T = 1;
            if (item < n_items) {
                const int f = a.f0 + item / TILES_PER_FRAME;
                t = __ldg(a.frame_t + f);
                T = __ldg(a.frame_T + f);
            }
        };
        int item = blockIdx.x;
        int t0, T0, t1, T1;
        frame_info(item, t0, T0);
        frame_info(item + gridDim.x, t1, T1);
        if (item < n_items) window_fetch(item, t0, T0);
        uint32_t ph_w = 0;
        int slot = 0;
        for (; item < n_items; item += gridDim.x) {
            mbar_wait(&wempty[slot], ph_w ^ 1);                // the builders are done with the tile that used this slot
            __nv_bfloat162* w2 = reinterpret_cast<__nv_bfloat162*>(win + slot * WIN_ELEMS);
#pragma unroll
            for (int u = 0; u < WIN_PER_THREAD; ++u)
                if (wdesc[u] >= 0) w2[lt + u * N_LOAD] = __floats2bfloat162_rn(wreg[u].x, wreg[u].y);
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&wfull[slot]);      // release: the stores above are visible to the waiting builders
            if (++slot == WIN_SLOTS) { slot = 0; ph_w ^= 1; }
            const int nxt = item + gridDim.x;
            int t2, T2;
            frame_info(nxt + gridDim.x, t2, T2);
            if (nxt < n_items) window_fetch(nxt, t1, T1);      // in flight until the next slot is free
            t1 = t2; T1 = T2;
        }
    } else {
        // ------------------------------------------------------------------------------------------------ epilogue
        const int quad = warp & 3;                             // TMEM lane quadrant of this warp
        const int p = quad * 32 + lane;                        // pixel of the tile = TMEM lane
        uint32_t ph_t = 0;
        int acc = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int fl = item / TILES_PER_FRAME, ti = item % TILES_PER_FRAME;
            const int P = ti * TILE + p;
            mbar_wait(&tfull[acc], ph_t);
            tc::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * NCH;
            uint32_t r0[32], r1[32];
            tc::tmem_ld_32x32(taddr, r0);
            tc::tmem_ld_32x32(taddr + 32, r1);
            tc::tmem_ld_wait();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&tempty[acc]);      // the accumulator is in registers: the MMAs of the tile after next may start
            if (P < PIX) {
                __nv_bfloat16* o = a.out + ((long long)fl * PIX + P) * NCH;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        const int c0 = h * 32 + j;
                        const float4 b0 = *reinterpret_cast<const float4*>(s_bias + c0), b1 = *reinterpret_cast<const float4*>(s_bias + c0 + 4);
                        const float4 s0 = *reinterpret_cast<const float4*>(s_prelu + c0), s1 = *reinterpret_cast<const float4*>(s_prelu + c0 + 4);
                        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                        const float ss[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                        float v[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const float t = __uint_as_float(h == 0 ? r0[j + u] : r1[j + u]) + bb[u];
                            v[u] = fmaxf(t, 0.f) + ss[u] * fminf(t, 0.f);        // PReLU without a branch
                        }
                        __nv_bfloat162 q0 = __floats2bfloat162_rn(v[0], v[1]), q1 = __floats2bfloat162_rn(v[2], v[3]);
                        __nv_bfloat162 q2 = __floats2bfloat162_rn(v[4], v[5]), q3 = __floats2bfloat162_rn(v[6], v[7]);
                        uint4 pk;
                        pk.x = *reinterpret_cast<uint32_t*>(&q0); pk.y = *reinterpret_cast<uint32_t*>(&q1);
                        pk.z = *reinterpret_cast<uint32_t*>(&q2); pk.w = *reinterpret_cast<uint32_t*>(&q3);
                        *reinterpret_cast<uint4*>(o + c0) = pk;
                    }
                }
            }
            if (++acc == 2) { acc = 0; ph_t ^= 1; }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tmem_base, 128);
    }
}

}  // namespace

// Conv3d(1 -> 64, 5x7x7, stride 1x2x2, pad 2x3x3) + folded BatchNorm + PReLU of frames [f0, f0 + nf) of the packed video
// ([F][88][88] fp32) as an implicit GEMM.  w = [64][320] bf16 with k = (dt*7 + dy)*8 + dx (dx = 7 and k >= 280 zero), bias / prelu
// [64] fp32, frame_t / frame_T [F] = position of a frame in its utterance / the utterance's length (temporal zero padding
// stops at utterance boundaries).  out = [nf][44][44][64] bf16 (NHWC).
extern "C" int avsr_frontend_conv3d(const float* video, const int* frame_t, const int* frame_T, int f0, int nf, const void* w,
                                    const float* bias, const float* prelu, void* out, cudaStream_t stream) {
    AVSR_REQUIRE(video && frame_t && frame_T && w && bias && prelu && out && nf > 0 && f0 >= 0, "avsr_frontend_conv3d: bad arguments");
    AVSR_REQUIRE(((uintptr_t)w & 15) == 0 && ((uintptr_t)out & 15) == 0, "avsr_frontend_conv3d: w / out must be 16-byte aligned");
    static bool configured = false;
    if (!configured) {
        AVSR_CHECK_CUDA(cudaFuncSetAttribute(frontend_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    int dev = 0, sms = 0;
    AVSR_CHECK_CUDA(cudaGetDevice(&dev));
    AVSR_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int items = nf * TILES_PER_FRAME;
    const FrontArgs a = {video, frame_t, frame_T, (const __nv_bfloat16*)w, bias, prelu, (__nv_bfloat16*)out, f0, nf};
    frontend_conv_kernel<<<items < sms ? items : sms, NUM_THREADS, SMEM_BYTES, stream>>>(a);
    AVSR_LAUNCH_CHECK();
    return AVSR_OK;
}
